"""Summarise an .ncu-rep (ncu --set full) into markdown for profiles/.
usage: python tools/ncu_summary.py rep.ncu-rep "title" algo_bytes[,algo_bytes...] [event_us[,event_us...]] > profiles/x.md"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
# algorithmic bytes / CUDA-event times: one value, or a comma list in kernel order ("-" to skip an entry)
algos = [None if a in ("-", "") else float(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else []
evs = [None if a in ("-", "") else float(a) for a in sys.argv[4].split(",")] if len(sys.argv) > 4 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
print(f"# {title}\n")
print(f"Source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`, one GPU, B200). "
      "Durations under ncu are cold-cache and serialised; the bench number is the CUDA-event time in bench.py.\n")
for ki, row in enumerate(rows[2:]):
    name = row[hdr.index("Kernel Name")]
    algo = (algos[ki] if ki < len(algos) else None) if len(algos) != 1 else algos[0]
    ev_us = (evs[ki] if ki < len(evs) else None) if len(evs) != 1 else evs[0]
    print(f"## `{name}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    vals = {}
    for k in keys:
        if k in hdr:
            v = row[hdr.index(k)]
            vals[k] = v
            print(f"| {k} | {v} | {units[hdr.index(k)]} |")
    try:
        rd = float(vals["dram__bytes_read.sum"].replace(",", "")); wr = float(vals["dram__bytes_write.sum"].replace(",", ""))
        ru, wu = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        traffic = rd * mult.get(ru, 1) + wr * mult.get(wu, 1)
        dur = float(vals["gpu__time_duration.sum"].replace(",", "")) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[units[hdr.index("gpu__time_duration.sum")]]
        print(f"\nDRAM traffic (read+write) per launch: **{traffic / 1e6:.2f} MB**", end="")
        if algo:
            print(f"; algorithmic bytes: {algo / 1e6:.2f} MB (traffic/algorithmic = {traffic / algo:.2f})", end="")
        print(f"; under-ncu rate {traffic / dur / 1e9:.0f} GB/s.")
        if ev_us and algo:
            print(f"CUDA-event time in bench.py: {ev_us} us -> {algo / ev_us / 1e3:.0f} GB/s = {algo / ev_us / 1e3 / 6572.2:.3f} of measured HBM peak (6572 GB/s).")
    except Exception:
        pass
    d = {h: row[i] for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")}
    tot = sum(float(v.replace(",", "") or 0) for v in d.values()) or 1
    mix = sorted(((float(v.replace(",", "") or 0) / tot, k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in d.items()), reverse=True)[:6]
    print("\nWarp-stall mix: " + ", ".join(f"{n} {f:.0%}" for f, n in mix) + "\n")

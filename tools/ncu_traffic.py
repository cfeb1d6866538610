"""DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch) of every kernel in an .ncu-rep, in launch order.
usage: python tools/ncu_traffic.py rep.ncu-rep"""
import csv
import subprocess
import sys


def traffic(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    out = []
    for row in rows[2:]:
        t = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            t += float(row[i].replace(",", "")) * mult[units[i]]
        out.append((row[hdr.index("Kernel Name")], int(t)))
    return out


if __name__ == "__main__":
    for name, t in traffic(sys.argv[1]):
        print(t, name)

"""GPU: the NVLink peer-memory exchange (b2_comm_*, b2_allreduce, b2_allgather, b2_gemm_wq_run_allreduce) with every "rank"
on ONE device: the communicators live in one process, are wired with b2_comm_connect_pointers, and each rank's kernels run
on their own stream — exactly the push / flag / wait / sum protocol of the multi-GPU case (there the peer pointers come from
CUDA IPC: tests/test_tp_gpu.py, 2 GPUs).  Results are bit-exact against fp32 sums in rank order."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _comms(n, max_bytes):
    from b200spark import ops
    return ops.Comm.connect_local([ops.Comm(r, n, max_bytes) for r in range(n)])


def _ref_sum(parts, residual=None):
    acc = torch.zeros_like(parts[0], dtype=torch.float32)
    for p in parts:  # rank order, fp32
        acc = acc + p.float()
    if residual is not None:
        acc = acc + residual.float()
    return acc.to(torch.bfloat16)


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("count", [8, 4096, 16 * 8192, 64 * 3584])
def test_allreduce_loopback_bit_exact(nranks, count):
    comms = _comms(nranks, count * 2)
    streams = [torch.cuda.Stream() for _ in range(nranks)]
    g = torch.Generator(device="cuda").manual_seed(count + nranks)
    for it in range(5):  # consecutive exchanges: both buffer parities, epochs advancing
        parts = [torch.randn(count, generator=g, device="cuda").to(torch.bfloat16) for _ in range(nranks)]
        res = torch.randn(count, generator=g, device="cuda").to(torch.bfloat16) if it % 2 else None
        outs = [torch.empty(count, dtype=torch.bfloat16, device="cuda") for _ in range(nranks)]
        torch.cuda.synchronize()
        for r in range(nranks):
            with torch.cuda.stream(streams[r]):
                comms[r].allreduce(parts[r], out=outs[r], residual=res)
        torch.cuda.synchronize()
        for c in comms:
            c.check_error()
        exp = _ref_sum(parts, res)
        for r in range(nranks):
            assert torch.equal(outs[r], exp), (it, r)


def test_allreduce_in_place_and_graph_replay():
    n, count = 2, 16 * 8192
    comms = _comms(n, count * 2)
    streams = [torch.cuda.Stream() for _ in range(n)]
    xs = [torch.zeros(count, dtype=torch.bfloat16, device="cuda") for _ in range(n)]
    src = [(torch.arange(count, device="cuda") % 7 + r).to(torch.bfloat16) for r in range(n)]
    graphs = []
    for r in range(n):  # one graph per rank: copy the source in, all-reduce in place (the epoch lives on the device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(streams[r]):
            xs[r].copy_(src[r])
        torch.cuda.synchronize()
        graphs.append(g)
    # capture on each rank's stream; the peers are not running during capture, nothing is launched
    for r in range(n):
        with torch.cuda.graph(graphs[r], stream=streams[r]):
            xs[r].copy_(src[r])
            comms[r].allreduce(xs[r])
    exp = _ref_sum(src)
    for it in range(4):
        for r in range(n):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        for c in comms:
            c.check_error()
        for r in range(n):
            assert torch.equal(xs[r], exp), (it, r)


def test_allgather_small():
    n = 4
    comms = _comms(n, 4096)
    streams = [torch.cuda.Stream() for _ in range(n)]
    ins = [torch.full((16,), float(r) + 0.5, dtype=torch.float32, device="cuda") for r in range(n)]
    outs = [torch.empty(n, 16, dtype=torch.float32, device="cuda") for _ in range(n)]
    for r in range(n):
        with torch.cuda.stream(streams[r]):
            comms[r].allgather(ins[r], outs[r])
    torch.cuda.synchronize()
    exp = torch.stack(ins)
    for r in range(n):
        assert torch.equal(outs[r], exp)


def test_peer_that_never_arrives_times_out(monkeypatch):
    """A missing rank must not hang the GPU: the waiting kernel gives up after B2_COMM_TIMEOUT_MS and b2_comm_error reports it."""
    from b200spark import ops
    from b200spark._lib import B2Error
    monkeypatch.setenv("B2_COMM_TIMEOUT_MS", "200")
    comms = _comms(2, 4096)
    x = torch.ones(64, dtype=torch.bfloat16, device="cuda")
    comms[0].allreduce(x)  # rank 1 never launches
    torch.cuda.synchronize()
    with pytest.raises(B2Error):
        comms[0].check_error()


@pytest.mark.parametrize("wbits,M,nranks", [(4, 1, 2), (4, 16, 2), (8, 5, 4), (4, 16, 8)])
def test_gemv_fused_allreduce_loopback(wbits, M, nranks):
    """Row-parallel projection (K split over the ranks, per-channel scale/zero replicated: the reference's HSPLIT) with the
    all-reduce fused into the GEMV epilogue == the ranks' plain GEMVs summed in rank order in fp32 (+ residual), bit for bit;
    and within the GEMM tolerance of the unsplit projection."""
    from b200spark import ops, quantize as PQ, tp as TP
    K, N = 2048, 1024 + 128  # 9 n-groups
    g = torch.Generator().manual_seed(wbits + M + nranks)
    w = (torch.randn(K, N, generator=g) * 0.02).to(torch.bfloat16)
    a = (torch.rand(M, K, generator=g) * 2 - 1).to(torch.bfloat16).cuda()
    res = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda()
    q, s, z = (PQ.quantize_a16w4 if wbits == 4 else PQ.quantize_a16w8)(w, -1)
    full = ops.GemmWQ(K, N, wbits, -1, max_m=M).prepare(q.cuda(), s.cuda(), z.cuda())
    ws = ops.Workspace()
    y_full = full(a, ws, residual=res)
    comms = _comms(nranks, M * N * 2)
    streams = [torch.cuda.Stream() for _ in range(nranks)]
    shards, wss = [], []
    for r in range(nranks):
        qr, sr, zr = TP.shard_rows(q, s, z, wbits, -1, r, nranks)
        shards.append(ops.GemmWQ(K // nranks, N, wbits, -1, max_m=M).prepare(qr.cuda(), sr.cuda(), zr.cuda()))
        wss.append(ops.Workspace())
    kr = K // nranks
    parts = [shards[r](a[:, r * kr:(r + 1) * kr], wss[r]) for r in range(nranks)]
    torch.cuda.synchronize()
    exp = _ref_sum(parts, res)
    for it in range(3):
        outs = [torch.empty(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(nranks)]
        torch.cuda.synchronize()
        for r in range(nranks):
            with torch.cuda.stream(streams[r]):
                assert shards[r].run_allreduce(a[:, r * kr:(r + 1) * kr], wss[r], comms[r], out=outs[r], residual=res)
        torch.cuda.synchronize()
        for c in comms:
            c.check_error()
        for r in range(nranks):
            assert torch.equal(outs[r], exp), (it, r)
    assert (exp.float() - y_full.float()).abs().max().item() <= 2e-2 * max(1.0, y_full.float().abs().max().item())

"""CPU, world_size 2, gloo: the tensor-parallel partitioning logic (b200spark/tp.py) — column-split then row-split of a
quantized projection pair with an all-reduce in between must reproduce the unsharded oracle (reference scheme:
qwen_v15.py:125-146, weight_splitter.cpp)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, wbits, group, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "dash-infer_b200", "python"))
    from b200spark import quantize as PQ, tp as TP  # imports the native lib (CPU-safe) — no GPU call is made here
    from oracle import quant_ref as Q
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)  # every rank builds the same full weights, then keeps its shard
    H, I, M = 256, 512, 3
    w1 = (torch.randn(H, I, generator=g) * 0.05).to(torch.bfloat16)
    w2 = (torch.randn(I, H, generator=g) * 0.05).to(torch.bfloat16)
    x = (torch.rand(M, H, generator=g) * 2 - 1).to(torch.bfloat16)
    quant = PQ.quantize_a16w4 if wbits == 4 else PQ.quantize_a16w8
    q1, s1, z1 = quant(w1, group)
    q2, s2, z2 = quant(w2, group)
    unp = (lambda q, n: Q.unpack_u4x2(q.numpy(), n)) if wbits == 4 else (lambda q, n: q.numpy())
    f = lambda t: t.float().numpy()
    # unsharded oracle
    h_full = Q.to_ft(Q.gemm_wq_math(f(x), unp(q1, I), f(s1), f(z1), group, act=Q.ACT_SILU), "bf16")
    y_full = Q.gemm_wq_math(h_full, unp(q2, H), f(s2), f(z2), group)
    # sharded: column split of the up-projection, row split of the down-projection
    q1s, s1s, z1s, _ = TP.shard_cols(q1, s1, z1, None, wbits, TP.col_range_even(I, rank, world))
    q2s, s2s, z2s = TP.shard_rows(q2, s2, z2, wbits, group, rank, world)
    Il = I // world
    h_loc = Q.to_ft(Q.gemm_wq_math(f(x), unp(q1s, Il), f(s1s), f(z1s), group, act=Q.ACT_SILU), "bf16")
    y_part = torch.from_numpy(Q.gemm_wq_math(h_loc, unp(q2s, H), f(s2s), f(z2s), group))
    dist.all_reduce(y_part)
    err = float(np.abs(y_part.numpy() - y_full).max()) / float(np.abs(y_full).max())
    # QKV group split covers every column exactly once across ranks
    cols = torch.zeros((8 + 2 * 2) * 128, dtype=torch.int32)
    for a, b in TP.col_ranges_qkv(8, 2, rank, world):
        cols[a:b] += 1
    dist.all_reduce(cols)
    ret[rank] = (err, bool((cols == 1).all()))
    dist.destroy_process_group()


@pytest.mark.parametrize("wbits,group", [(4, -1), (8, -1), (4, 128)])
def test_tp2_shards_reproduce_unsharded(wbits, group):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, wbits, group, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    for r in range(2):
        err, cover = ret[r]
        assert err < 1e-5, err  # fp32 partial sums: only summation order differs
        assert cover

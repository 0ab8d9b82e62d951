#!/usr/bin/env python
"""A/B timings of single kernels with CUDA events (warm L2 like inside a training step, 20 launches after 5 warm-ups),
at the shapes the bf16 bench step launches: attention forward/backward (FFMA vs tcgen05, FIRA_ATTN_TC) and the GCN layer
(scatter + GEMM + LayerNorm vs the fused kernel, forward and backward).  One JSON line per measurement.

    python tools/bench_kernels.py [--batch 64]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = "cuda:0"
BF = torch.bfloat16


def timeit(fn, reps=10, iters=10, warm=3):
    """GPU time per call: `reps` back-to-back launches captured into ONE CUDA graph (no host launch overhead between
    them, the way the training step replays them), replayed `iters` times between CUDA events."""
    for _ in range(warm):
        fn()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    us = sorted(1e3 * a.elapsed_time(b) / reps for a, b in ev)
    return {"avg_us": sum(us) / len(us), "min_us": us[0], "med_us": us[len(us) // 2], "timing": f"{reps} launches per graph replay"}


def st():
    return torch.cuda.current_stream().cuda_stream


def attention(B, out):
    from fira_icse_b200 import _lib
    H, T, dh, D = 8, 30, 32, 256
    g = torch.Generator().manual_seed(0)
    for name, Lk, causal, valid in (("cross S=304 (127 valid)", 304, 0, 127), ("cross S=370 (all valid)", 370, 0, 370),
                                    ("self T=30 causal", 30, 1, 30)):
        q = torch.randn(B * T, D, generator=g).to(BF).to(DEV)
        kv = torch.randn(B * Lk, 2 * D, generator=g).to(BF).to(DEV)
        mask = torch.zeros(B, Lk, dtype=torch.uint8)
        mask[:, :valid] = 1
        mask = mask.to(DEV)
        ctx = torch.empty(B * T, D, device=DEV, dtype=BF)
        stats = torch.empty(B, H, T, 2, device=DEV)
        go = torch.randn(B * T, D, generator=g).to(BF).to(DEV)
        dq, dkv = torch.empty_like(q), torch.zeros_like(kv)
        ld = 2 * D

        def fwd():
            _lib.call("fira_attn_fwd", q.data_ptr(), D, kv.data_ptr(), ld, kv.data_ptr() + D * 2, ld, mask.data_ptr(), causal,
                      ctx.data_ptr(), D, stats.data_ptr(), B, H, T, Lk, dh, 1, st())

        def bwd():
            _lib.call("fira_attn_bwd", q.data_ptr(), D, kv.data_ptr(), ld, kv.data_ptr() + D * 2, ld, mask.data_ptr(), causal,
                      ctx.data_ptr(), go.data_ptr(), D, stats.data_ptr(), dq.data_ptr(), D, dkv.data_ptr(), ld,
                      dkv.data_ptr() + D * 2, ld, B, H, T, Lk, dh, 1, st())
        for tc in ("0", "1"):
            os.environ["FIRA_ATTN_TC"] = tc
            fwd()
            out({"kernel": "attention fwd", "case": name, "tcgen05": tc == "1", **timeit(fwd)})
            out({"kernel": "attention bwd", "case": name, "tcgen05": tc == "1", **timeit(bwd)})
    os.environ.pop("FIRA_ATTN_TC", None)


def gcn(B, out):
    from fira_icse_b200 import PackedEdges, _lib, ops
    from fira_icse_b200.data import trim_batch_host
    from fira_icse_b200.synth import N_NODES, synth_batch
    ids, coo = synth_batch(0, B, 24650, 71)
    t = {k: torch.from_numpy(v) for k, v in ids.items()}
    rowptr, col, val = PackedEdges.pack_host(coo, N_NODES, pin=False)
    lst = trim_batch_host([t["sou"], t["tar"], None, t["mark"], t["ast_change"], (rowptr, col, val), t["tar_label"],
                           t["sub_token"]], 24650)
    n = (lst[0].shape[1], lst[7].shape[1], lst[4].shape[1])
    N, R, Mc = sum(n), B * sum(n), B * n[0]
    pe = PackedEdges.from_host(*lst[5], B, N, DEV)
    er = pe.rows_csr(*n)
    pr = ops.Prec(True)
    H = torch.randn(R, 256, device=DEV).to(BF)
    Wc = torch.randn(256, 256, device=DEV) / 16
    Wc16, WcT16 = Wc.to(BF), Wc.t().contiguous().to(BF)
    b2, c1 = torch.randn(256, device=DEV) * 0.1, torch.randn(256, device=DEV) * 0.1
    gamma, beta = torch.ones(256, device=DEV), torch.zeros(256, device=DEV)
    rs = pe.rowsum(*n)
    G, Z = torch.empty_like(H), torch.empty_like(H)
    oA, oB = torch.empty(Mc, 256, device=DEV, dtype=BF), torch.empty_like(H)
    stats = torch.empty(2, R, device=DEV)
    p, seed = 0.2, 1234

    def unfused_fwd():
        _lib.call("fira_gcn_aggregate", pe.rowptr.data_ptr(), pe.col.data_ptr(), pe.val.data_ptr(), H.data_ptr(), None,
                  G.data_ptr(), B, n[0], n[1], n[2], 256, 1, st())
        z = pr.linear(G, Wc, b2, rs=rs, rc=c1, out=Z)
        pr.ln_fwd(z, H, gamma, beta, oA, oB, Mc, R, p, seed, 2)

    def fused_fwd():
        _lib.call("fira_gcn_layer_fwd", er[0].data_ptr(), er[1].data_ptr(), er[2].data_ptr(), H.data_ptr(), Wc16.data_ptr(),
                  b2.data_ptr(), c1.data_ptr(), gamma.data_ptr(), beta.data_ptr(), Z.data_ptr(), oA.data_ptr(),
                  oB.data_ptr(), Mc, stats.data_ptr(), stats.data_ptr() + 4 * R, R, 256, p, seed, None, 2, st())
    dZ, dRes = torch.randn(R, 256, device=DEV).to(BF), torch.randn(R, 256, device=DEV).to(BF)
    dG, dH, AdZ = torch.empty_like(H), torch.empty_like(H), torch.empty_like(H)

    def unfused_bwd():
        pr.linear_dx(dZ, 256, Wc, R, out=dG)
        _lib.call("fira_gcn_aggregate", pe.rowptr.data_ptr(), pe.col.data_ptr(), pe.val.data_ptr(), dG.data_ptr(),
                  dRes.data_ptr(), dH.data_ptr(), B, n[0], n[1], n[2], 256, 1, st())

    def fused_bwd():
        _lib.call("fira_gcn_layer_bwd", er[0].data_ptr(), er[1].data_ptr(), er[2].data_ptr(), dZ.data_ptr(),
                  WcT16.data_ptr(), dRes.data_ptr(), AdZ.data_ptr(), dH.data_ptr(), R, 256, st())
    info = {"rows": R, "segments": n, "nnz": pe.nnz}
    out({"kernel": "GCN layer fwd: scatter + tcgen05 GEMM + LayerNorm (3 launches)", **info, **timeit(unfused_fwd)})
    out({"kernel": "GCN layer fwd: fused (1 launch)", **info, **timeit(fused_fwd)})
    out({"kernel": "GCN layer bwd (dX path): GEMM + scatter (2 launches)", **info, **timeit(unfused_bwd)})
    out({"kernel": "GCN layer bwd (dX path): fused (1 launch)", **info, **timeit(fused_bwd)})
    # algorithmic bytes of the fused forward (SURVEY.md 8d with H in / out replacing X1 / X2): read H, write Z and out,
    # CSR metadata, the weight once per launch
    alg = 3 * R * 256 * 2 + (R + 1) * 4 + pe.nnz * 8 + 256 * 256 * 2
    out({"kernel": "GCN layer fwd fused: algorithmic bytes", "bytes": alg})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()

    def out(d):
        print(json.dumps(d), flush=True)
    attention(a.batch, out)
    gcn(a.batch, out)


if __name__ == "__main__":
    main()

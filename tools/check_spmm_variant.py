#!/usr/bin/env python
"""Correctness + timing of one GNN-scatter variant (FIRA_SPMM_VARIANT, read once per process by the library)
against an fp64 index_add reference: DataSet-like 3-segment batches (with and without addend) and the
config-5 stress graphs.  Prints one JSON line per case; exits non-zero on a mismatch.

    FIRA_SPMM_VARIANT=3 python tools/check_spmm_variant.py [--time]      (1, 3, 4; unset = the defaults)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from fira_icse_b200 import _lib  # noqa: E402
from fira_icse_b200.graph import PackedEdges  # noqa: E402
from fira_icse_b200.synth import synth_batch, synth_stress_graphs  # noqa: E402

DEV = torch.device("cuda:0")


def seg_row(b, j, B, n0, n1, n2):
    return torch.where(j < n0, b * n0 + j,
                       torch.where(j < n0 + n1, B * n0 + b * n1 + (j - n0), B * (n0 + n1) + b * n2 + (j - n0 - n1)))


def reference(pe, x, addend, B, segs):
    N = sum(segs)
    deg = (pe.rowptr[1:] - pe.rowptr[:-1]).long()
    g = torch.repeat_interleave(torch.arange(B * N, device=DEV), deg)
    b, i = g // N, g % N
    dst = seg_row(b, i, B, *segs)
    src = seg_row(b, pe.col.long(), B, *segs)
    y = torch.zeros(B * N, 256, dtype=torch.float64, device=DEV)
    y.index_add_(0, dst, pe.val.double().unsqueeze(1) * x.double()[src])
    return y + addend.double() if addend is not None else y


def check(name, coo, N, segs, code, with_addend):
    B = len(coo)
    pe = PackedEdges.from_coo_lists(coo, N, DEV)
    tdt = torch.float32 if code == 0 else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B * N, 256, generator=g).to(DEV).to(tdt)
    add = torch.randn(B * N, 256, generator=g).to(DEV).to(tdt) if with_addend else None
    y = torch.empty_like(x)
    _lib.call("fira_gcn_aggregate", pe.rowptr.data_ptr(), pe.col.data_ptr(), pe.val.data_ptr(), x.data_ptr(),
              add.data_ptr() if add is not None else None, y.data_ptr(), B, *segs, 256, code,
              torch.cuda.current_stream().cuda_stream)
    ref = reference(pe, x, add, B, segs)
    err = (y.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (1e-5 if code == 0 else 2 ** -7) * scale
    ok = err <= tol
    print(json.dumps({"case": name, "variant": os.environ.get("FIRA_SPMM_VARIANT", "default"),
                      "dtype": "f32" if code == 0 else "bf16", "addend": with_addend, "max_abs_err": err,
                      "ref_max": scale, "ok": ok}), flush=True)
    return ok


if __name__ == "__main__":
    ok = True
    _, coo = synth_batch(0, 8)
    stress = synth_stress_graphs(0, 4)
    for code in (0, 1):
        for addend in (False, True):
            ok &= check("dataset-like B=8 (210/160/280 segments)", coo, 650, (210, 160, 280), code, addend)
        ok &= check("stress N=2048 B=4", stress, 2048, (2048, 0, 0), code, True)
    if not ok:
        sys.exit(1)
    if "--time" in sys.argv:
        import bench_spmm
        for B in (32, 256):
            g = synth_stress_graphs(0, B)
            bench_spmm.run(f"stress N=2048 16k edges/relation B={B}", g, 2048, (2048, 0, 0), 0)
            bench_spmm.run(f"stress N=2048 16k edges/relation B={B}", g, 2048, (2048, 0, 0), 1)
        _, coo = synth_batch(0, 64)
        bench_spmm.run("dataset-like B=64", coo, 650, (210, 160, 280), 0)
        bench_spmm.run("dataset-like B=64", coo, 650, (210, 160, 280), 1)

#!/usr/bin/env python
"""A/B of the stand-alone scatter variants (FIRA_SPMM_VARIANT is read once per process: run one process per variant):
the bf16 / fp32 kernel on the padded graphs of 64 and 512 commits, timed like bench.py's roofline (graph-replayed launches
over rotating > L2 buffers).  One JSON line per (variant, dtype, batch).

    for v in 1 4 6 7 8; do FIRA_SPMM_VARIANT=$v python tools/scatter_variants.py; done
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import __graft_entry__
    __graft_entry__.build()
    dev = torch.device("cuda:0")
    v = os.environ.get("FIRA_SPMM_VARIANT", "default")
    for B in (64, 512):
        hb = bench.host_batch(0, B, pin=False, trim=False)
        for bf16 in (True, False):
            if B == 512 and not bf16:
                continue
            r = bench.spmm_roofline(dev, hb, B, bf16=bf16)
            print(json.dumps({"variant": v, "dtype": r["dtype"], "commits": B, "rows": r["rows"], "avg_us": round(r["avg_launch_ms"] * 1e3, 2),
                              "GBps": round(r["achieved"], 1), "frac_of_measured_peak": round(r["frac"], 3)}), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

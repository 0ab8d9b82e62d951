#!/usr/bin/env python
"""Beam-search inference throughput (BASELINE.json config 4: `run_model.py test`, beam 3 / 5).

Synthetic commits with the DataSet node/edge distribution, random-initialised weights.  With random
weights no beam emits <eos>, so every batch runs all tar_len-1 = 29 decoding steps: the worst case the
reference's own loop was timed on in BASELINE.md (72 s for 32 commits on CPU).  One JSON line per
(batch, beam) configuration; timing with CUDA events around whole batches, inputs resident on the device.

    python tools/bench_beam.py [--batches 20,128] [--beams 3,5] [--precision fp32|bf16] [--reps 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="20,128")
    ap.add_argument("--beams", default="3,5")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--trim", action="store_true", help="loader-side padding trimming (data.trim_batch_host)")
    ap.add_argument("--modes", default="full,incremental,graph")
    a = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    import bench
    import fira_icse_b200 as F
    from fira_icse_b200 import _lib
    from fira_icse_b200.beam import beam_search
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = F.TransModel(bench.model_args()).to(dev)
    model.set_precision(a.precision)
    model.eval()
    for B in (int(x) for x in a.batches.split(",")):
        hb = bench.host_batch(10_000, B, pin=False, trim=a.trim)
        b = bench.device_batch(hb, dev, B)
        for K, mode in ((int(x), m) for x in a.beams.split(",") for m in a.modes.split(",")):
            def run():
                return beam_search(model, b[0], b[3], b[4], b[5], b[7], beam_size=K, tar_len=30, start_id=1, eos_id=2,
                                   pad_id=0, mode=mode)
            ref = run()                                              # warm-up (lazy CUDA state, graph capture)
            if mode == "full":
                ref_full = ref
            same = bool(torch.equal(ref[0], ref_full[0])) if "full" in a.modes.split(",") else None
            torch.cuda.synchronize()
            n0 = _lib.LAUNCH_COUNT
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                seq, length, prob = run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            print(json.dumps({
                "metric": "beam-search inference throughput", "unit": "commits/s", "value": B / ms * 1e3,
                "ms_per_batch": ms, "batch": B, "beam": K, "decoded_steps": int(length.max().item()) - 1,
                "precision": a.precision, "mode": mode, "ids_equal_full_mode": same, "trimmed": bool(a.trim), "data": "synthetic (DataSet distribution), random weights",
                "c_abi_calls_per_batch": (_lib.LAUNCH_COUNT - n0) // a.reps,
                "note": "encoder once per batch; full = 30-position decoder re-run per step over all live beams, "
                        "incremental = newest row against K/V caches, graph = the same as CUDA-graph replays"}),
                  flush=True)


if __name__ == "__main__":
    main()

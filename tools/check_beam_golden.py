#!/usr/bin/env python
"""Beam-search id parity against a golden file produced by the unmodified reference's test() loop
(tests/golden/make_golden_beam.py): every decoder mode, any beam size.  The beam-3 case is a regular test
(tests/test_gpu_cli.py); the beam-5 golden (tests/golden/beam5_first16.npz) was generated after round 1's GPU budget
was spent, so it is checked with this tool first and becomes a test once it has run on a GPU.

    python tools/check_beam_golden.py [tests/golden/beam5_first16.npz]
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from fira_testlib import GOLDEN, golden_batch, load_raw_golden, seeded_model
    from fira_icse_b200.beam import beam_search, best_sequences
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(GOLDEN, "beam5_first16.npz")
    gold = np.load(path)
    vocab = load_raw_golden()["word_vocab"]
    dev = "cuda:0"
    model = copy.deepcopy(seeded_model()).to(dev).eval()
    with torch.no_grad():                      # same sharpening as tests/golden/make_golden_beam.py
        k = float(gold["sharpen"])
        model.out_fc.weight *= k; model.out_fc.bias *= k; model.copy_net.LinearRes.weight *= k
    bs, beam = int(gold["batch"]), int(gold["beam"])
    bad = 0
    for mode in ("full", "incremental", "graph"):
        for lo in range(0, gold["beam_ids"].shape[0], bs):
            b = golden_batch(lo, lo + bs)
            seq, length, prob = beam_search(model, b[0], b[3], b[4], b[5].to(dev), b[7], beam_size=beam, tar_len=30,
                                            start_id=vocab["<start>"], eos_id=vocab["<eos>"], pad_id=vocab["<pad>"],
                                            mode=mode)
            best, blen = best_sequences(seq, length, prob)
            for i in range(bs):
                ref = gold["beam_ids"][lo + i]
                ref = ref[ref >= 0]
                mine = best[i, :blen[i]].cpu().numpy()
                if not np.array_equal(mine, ref):
                    bad += 1
                    print(f"[{mode}] commit {lo + i}: mine {mine.tolist()} reference {ref.tolist()}")
        print(f"mode {mode}: done")
    print("beam", beam, "mismatching (mode, commit) pairs:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

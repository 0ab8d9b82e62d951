#!/usr/bin/env python
"""Golden TRAINING CURVE of the UNMODIFIED reference (build container only: imports /root/reference).

`TransModel` under torch.manual_seed(0) (run_model.py:61-69), torch.optim.Adam(lr=1e-4) (run_model.py:396), the loop body
of run_model.py:101-109 (forward -> loss.sum()/mask.sum() -> zero_grad -> backward -> step) on the 128 golden commits
(tests/golden/batch_first128.npz: the reference's own process_data output, dense float64 adjacency as
Dataset.__getitem__ builds it) in batches of 16, four passes = 32 steps, dropout off (model.eval(): Philox streams
differ between torch and the CUDA kernels).  Writes tests/golden/train_curve.npz = the loss of every
step; tests/test_gpu_train_curve.py replays the same steps on the CUDA path.

Usage:  python tests/golden/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BATCH, PASSES = 16, 4


def main():
    torch.set_num_threads(8)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from Model import TransModel                       # the unmodified reference
    from fira_testlib import golden_batch, reference_args
    torch.manual_seed(0)
    model = TransModel(reference_args())
    model.eval()                                       # dropout off; parameters still train
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    batches = [golden_batch(lo, lo + BATCH) for lo in range(0, 128, BATCH)]
    losses = []
    for p in range(PASSES):
        for b in batches:
            loss, mask = model(*b, "train")
            loss = loss.sum() / mask.sum()
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss.item()))
            print(len(losses), losses[-1], flush=True)
    np.savez(os.path.join(HERE, "train_curve.npz"), loss=np.array(losses, np.float64), batch=BATCH, passes=PASSES,
             lr=1e-4, torch_version=torch.__version__)


if __name__ == "__main__":
    main()

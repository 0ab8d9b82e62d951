#!/usr/bin/env python
"""TEST / MEASUREMENT INFRASTRUCTURE -- not product code.  Times the UNMODIFIED reference training step on host
CPU cores: `TransModel` imported from oracle/_ref (Model.py, gnn_transformer.py,
combination_layer.py staged verbatim by oracle/make_ref.sh), `torch.optim.Adam(lr=1e-4)`, fp32, dense float64
[B,650,650] adjacency exactly as Dataset.__getitem__ produces it (Dataset.py:336-343), the loop body of
run_model.py:101-109 (forward -> loss.sum()/mask.sum() -> zero_grad -> backward -> step), model.train() (dropout
on, as run_model.py:87).  Must run with CUDA_VISIBLE_DEVICES="" (the reference branches on
torch.cuda.is_available() globally: run_model.py:20, Model.py:72, gnn_transformer.py:110); bench.py starts it that
way as a subprocess.  Nothing of fira_icse_b200's models, kernels or engine is imported: only the synthetic-commit
generator (fira_icse_b200/synth.py, numpy), loaded by file path, provides the input batch.

    CUDA_VISIBLE_DEVICES="" python oracle/ref_cpu_bench.py --batch 64 --steps 5 --warmup 2 --threads 16
    ... --calibrate 4,8,16,32   # one forward per thread count, prints the timings

Prints one JSON line: {"impl": "reference"|"port", "step_s": [...], "threads": n, ...}.
If oracle/_ref is absent it falls back to the oracle port (oracle/fira_oracle.py) and says so ("port").
"""
import argparse
import importlib.util
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VOCAB, AST_VOCAB = 24650, 71


class DotDict(dict):
    def __getattr__(self, k):
        return self[k]


def reference_args():
    """run_model.py:27-56"""
    return DotDict(sou_len=210, tar_len=30, att_len=25, ast_change_len=280, sub_token_len=160, lr=1e-4,
                   dropout_rate=0.1, num_head=8, embedding_dim=256, vocab_size=VOCAB,
                   ast_change_vocab_size=AST_VOCAB, batch_size=64)


def load_synth():
    spec = importlib.util.spec_from_file_location("fira_synth", os.path.join(ROOT, "fira_icse_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_batch(first_index, batch):
    """The 8-tuple Dataset.__getitem__ + default collate hands the model: int64 ids, attr [B,210,25] (unused),
    dense float64 adjacency."""
    import numpy as np
    import torch
    synth = load_synth()
    ids, coo = synth.synth_batch(first_index, batch, VOCAB, AST_VOCAB)
    t = {k: torch.from_numpy(v) for k, v in ids.items()}
    edge = torch.zeros(batch, 650, 650, dtype=torch.float64)
    for b, (r, c, v) in enumerate(coo):
        edge[b].index_put_((torch.from_numpy(np.asarray(r)).long(), torch.from_numpy(np.asarray(c)).long()),
                           torch.from_numpy(np.asarray(v, dtype=np.float64)), accumulate=True)
    attr = torch.zeros(batch, 210, 25, dtype=torch.int64)
    return [t["sou"], t["tar"], attr, t["mark"], t["ast_change"], edge, t["tar_label"], t["sub_token"]]


def build_reference():
    import torch
    ref_dir = os.path.join(HERE, "_ref")
    if os.path.exists(os.path.join(ref_dir, "Model.py")):
        sys.path.insert(0, ref_dir)
        from Model import TransModel                     # the unmodified reference
        torch.manual_seed(0)
        model = TransModel(reference_args())
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)

        def step(batch):
            loss, mask = model(*batch, "train")
            loss = loss.sum() / mask.sum()
            opt.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.item())

        def fwd(batch):
            with torch.no_grad():
                return model(*batch, "train")
        return "reference", step, fwd
    # fallback: CPU restatement of the same algorithm
    sys.path.insert(0, HERE)
    import fira_oracle as O
    params = {k: v.requires_grad_(True) for k, v in O.random_state_dict(VOCAB, AST_VOCAB).items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)

    def step(batch):
        return O.train_step(params, opt, batch)

    def fwd(batch):
        with torch.no_grad():
            return O.forward(params, *batch, stage="train")
    return "port", step, fwd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--calibrate", default="")
    ap.add_argument("--first-index", type=int, default=10_000)
    a = ap.parse_args()
    import torch
    assert not torch.cuda.is_available(), "run with CUDA_VISIBLE_DEVICES='' (the reference branches on cuda availability)"
    kind, step, fwd = build_reference()
    if a.calibrate:
        b = make_batch(20_000, 8)
        log, best, best_t = {}, None, float("inf")
        for c in [int(x) for x in a.calibrate.split(",")]:
            torch.set_num_threads(c)
            fwd(b)
            t0 = time.perf_counter()
            fwd(b)
            dt = time.perf_counter() - t0
            log[c] = round(dt, 3)
            if dt < best_t:
                best, best_t = c, dt
            if dt > 4 * best_t:
                break
        print(json.dumps({"impl": kind, "calibration_s": log, "best_threads": best}), flush=True)
        return
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    pool = [make_batch(a.first_index + i * a.batch, a.batch) for i in range(2)]
    losses = []
    for i in range(a.warmup):
        losses.append(step(pool[i % 2]))
    times = []
    for i in range(a.steps):
        t0 = time.perf_counter()
        losses.append(step(pool[i % 2]))
        times.append(time.perf_counter() - t0)
    print(json.dumps({"impl": kind, "batch": a.batch, "threads": torch.get_num_threads(), "step_s": times,
                      "total_s": sum(times), "last_loss": losses[-1], "torch": torch.__version__}), flush=True)


if __name__ == "__main__":
    main()

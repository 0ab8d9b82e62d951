#!/usr/bin/env python
"""N-GPU check of the graphed data-parallel step (engine.GraphedTrainStep, split graphs + gradient all-reduce on a
communication stream overlapping the encoder backward) against the plain eager data-parallel step
(parallel.DataParallelStep: one backward, one flat all-reduce) -- same replicas, same shards, NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_dp_engine.py [--layout packed|padded]

Every rank trains `steps` SGD steps (SGD: Adam's g / sqrt(v) would amplify summation-order noise) on ITS shard of the
golden commits with both engines; the global losses must agree step by step and the parameters at the end, on every
rank, and the replicas must stay identical across ranks.  Prints one JSON line on rank 0; exit code 1 on mismatch."""
import argparse
import copy
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="packed", choices=["packed", "padded"])
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--optimizer", default="sgd", choices=["sgd", "flat_adam"],
                    help="flat_adam: optim.FlatAdam on both sides (gradients written into the flat buffers, all-reduced in "
                         "place, Adam of the head/decoder parameters behind the encoder backward); losses compared at 2e-3 "
                         "(Adam amplifies summation-order noise), parameters by the size of one Adam step")
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from fira_testlib import golden_batch, seeded_model
    from fira_icse_b200 import PackedEdges
    from fira_icse_b200.engine import GraphedTrainStep
    from fira_icse_b200.parallel import DataParallelStep
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    from test_packed import GoldenSplit, V
    B = 8
    base = copy.deepcopy(seeded_model()).to(dev)
    base.eval()
    m_ref, m_eng = copy.deepcopy(base), copy.deepcopy(base)
    n_batches = 128 // (B * world)
    if a.layout == "packed":
        tables = PackedTables(GoldenSplit())
        hosts = [pack_from_dataset(tables, np.arange((i * world + rank) * B, (i * world + rank + 1) * B), V)
                 for i in range(n_batches)]
        ref_batches = None
    else:
        hosts = []
        for i in range(n_batches):
            lo = (i * world + rank) * B
            b = golden_batch(lo, lo + B, dense_edge=False)
            hosts.append([b[0], b[1], None, b[3], b[4], PackedEdges.pack_host(b[5], 650), b[6], b[7]])

    def padded_dev(h):
        d = [x.to(dev) if torch.is_tensor(x) else x for x in h]
        d[5] = PackedEdges.from_host(*h[5], B, 650, dev)
        return d
    lr = 2e-3
    if a.optimizer == "flat_adam":
        from fira_icse_b200.optim import FlatAdam
        lr = 1e-4
        dp = DataParallelStep(m_ref, lambda ps: FlatAdam(ps, lr=lr, groups=m_ref.flat_groups()))
        eng = GraphedTrainStep(m_eng, B, lambda ps: FlatAdam(ps, lr=lr, groups=m_eng.flat_groups()), edge_capacity=65536)
    else:
        dp = DataParallelStep(m_ref, lambda ps: torch.optim.SGD(ps, lr=lr))
        eng = GraphedTrainStep(m_eng, B, lambda ps: torch.optim.SGD(ps, lr=lr), edge_capacity=65536)
    loss_tol, param_tol = (2e-3, 6 * lr * a.steps) if a.optimizer == "flat_adam" else (2e-4, 5e-4)
    assert eng.split, "the N > 1 engine should run the split (overlapped) step"
    worst = 0.0
    ok = True
    for s in range(a.steps):
        h = hosts[s % len(hosts)]
        if a.layout == "packed":
            # eager reference of the packed layout: forward_packed + global token count + one flat all-reduce
            pb = h.to(dev)
            dp.bucket.zero()
            dp.optimizer.zero_grad(set_to_none=True)
            ls, nt = m_ref.forward_packed(pb, "train")
            ng = nt.to(torch.float32).reshape(1).clone()
            lg = ls.detach().reshape(1).clone()
            dist.all_reduce(ng)
            (ls / ng.squeeze(0)).backward()
            dp.bucket.all_reduce()
            dist.all_reduce(lg)
            dp.optimizer.step()
            loss_ref = (lg / ng).item()
        else:
            loss, _ = dp.step(padded_dev(h))
            loss_ref = loss.item()
        ls, nt = eng.step(h)
        tot = torch.stack((ls.detach().float(), nt.float()))
        dist.all_reduce(tot)
        loss_eng = (tot[0] / tot[1]).item()
        rel = abs(loss_eng - loss_ref) / abs(loss_ref)
        worst = max(worst, rel)
        ok &= rel <= loss_tol
    pmax = 0.0
    for (k, p), (_, q) in zip(m_ref.named_parameters(), m_eng.named_parameters()):
        pmax = max(pmax, (p - q).abs().max().item())
    # replicas identical across ranks
    flat = torch.cat([p.detach().reshape(-1) for p in m_eng.live_parameters()])
    other = flat.clone()
    dist.broadcast(other, 0)
    drift = (flat - other).abs().max().item()
    res = torch.tensor([worst, pmax, drift], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    ok = ok and res[1].item() <= param_tol and res[2].item() == 0.0
    if rank == 0:
        print(json.dumps({"check": "GraphedTrainStep(split, NCCL) vs eager DataParallelStep", "layout": a.layout,
                          "optimizer": a.optimizer, "world": world, "steps": a.steps, "worst_loss_rel_diff": res[0].item(),
                          "max_param_abs_diff": res[1].item(), "replica_drift": res[2].item(), "ok": bool(ok)}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

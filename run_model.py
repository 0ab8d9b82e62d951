#!/usr/bin/env python
"""`python run_model.py train|test` -- the reference's CLI (run_model.py:417-425) on the B200 path.

Same CWD-relative files (DataSet/*.json, all_index, VOCAB_UPPER_CASE, best_model.pt,
OUTPUT/{output_fira,train_process,dev_output}), same hyper-parameters (run_model.py:27-46), same
training loop semantics (Adam 1e-4, loss = sum(loss)/sum(tokens), dev BLEU every 10 batches from
epoch 15, best checkpoint saved as a plain state_dict with the reference's 338 keys) and the same
beam search ranking.  Differences, all below the module surface:
  * one process per GPU (launch with torchrun for N > 1) with an NCCL gradient all-reduce instead of
    nn.DataParallel; batch 170 PER GPU like upstream (run_model.py:40);
  * training batches come from the native loader (data.PackedBatchLoader: C++ gather + padding trim + CSR
    packing into pinned memory) instead of dense float64 650x650 through DataLoader workers;
  * the training step (zero-grad, forward, backward, Adam) is replayed as a CUDA graph per batch shape
    (engine.GraphedTrainStep); FIRA_ENGINE=eager issues the same kernels launch by launch;
  * optional env overrides, defaults unchanged: FIRA_BATCH, FIRA_TEST_BATCH, FIRA_EPOCHS, FIRA_BEAM,
    FIRA_MAX_BATCHES (smoke runs), FIRA_WORKERS, FIRA_PRECISION (fp32 parity mode | bf16 throughput mode),
    FIRA_MAX_SHAPES (bound on distinct trimmed batch shapes = captured graphs, default 24), FIRA_BUCKET (0 = the
    reference's uniform batching; K > 1 = size-bucketed batches within windows of K batches, see PackedBatchLoader).
"""
import json
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist
from torch.optim import Adam
from torch.utils.data import DataLoader

from fira_icse_b200 import TransModel
from fira_icse_b200.beam import beam_search, best_sequences
from fira_icse_b200.bleu import sentence_bleu_method2
from fira_icse_b200.data import PackedBatchLoader, TransDataset, batch_to_device, collate_packed
from fira_icse_b200.engine import GraphedTrainStep
from fira_icse_b200.parallel import DataParallelStep, shard_range


class DotDict(dict):
    def __getattr__(self, attr):
        return self[attr]


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))

args = DotDict({
    'sou_len': 210, 'tar_len': 30, 'att_len': 25, 'ast_change_len': 280, 'sub_token_len': 160,
    'lr': 1e-4, 'dropout_rate': 0.1, 'num_head': 8, 'embedding_dim': 256,
    'batch_size': int(os.environ.get("FIRA_BATCH", 170)),           # per GPU, as upstream (170 * n_gpu global)
    'test_batch_size': int(os.environ.get("FIRA_TEST_BATCH", 20)),
    'epoches': int(os.environ.get("FIRA_EPOCHS", 150)),
    'beam_size': int(os.environ.get("FIRA_BEAM", 3)),
    'vocab_size': 0, 'ast_change_vocab_size': 0,
})


def load_globals():
    g = {}
    g["vocab"] = json.load(open('DataSet/word_vocab.json'))
    g["r_vocab"] = {v: k for k, v in g["vocab"].items()}
    args.vocab_size = len(g["vocab"])
    args.ast_change_vocab_size = len(json.load(open('DataSet/ast_change_vocab.json')))
    g["var_maps"] = json.load(open("DataSet/variable.json"))
    return g


def seed_everything(seed=0):
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def device():
    if not torch.cuda.is_available():
        raise SystemExit("run_model.py: no CUDA device. This build has no CPU path; the CPU reference is "
                         "the upstream repository (or oracle/ for tests).")
    torch.cuda.set_device(LOCAL_RANK)
    return torch.device("cuda", LOCAL_RANK)


def ids_to_text(ids, r_vocab):
    s = ' '.join(r_vocab[i] for i in ids)
    return s.replace('<start>', "").replace('<eos>', "").replace('<pad>', "").replace('<unkm>', "😅").strip().split()


def deanonymise(tokens, var_map):
    rev = {v: k for k, v in var_map.items()}
    return [rev.get(t, t) for t in tokens]


def loader(ds, batch_size, shuffle, indices=None):
    sampler = None
    if indices is not None:
        ds = torch.utils.data.Subset(ds, indices)
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, sampler=sampler,
                      num_workers=int(os.environ.get("FIRA_WORKERS", 2)),
                      collate_fn=lambda items: collate_packed(items, pin=False), pin_memory=True)


def dev(model, dev_loader, g, valid_index, epoch, dev):
    """Teacher-forced argmax + sentence BLEU (run_model.py:118-184)."""
    vocab, r_vocab, var_maps = g["vocab"], g["r_vocab"], g["var_maps"]
    model.eval()
    out_str, bleus, total = '', 0.0, 0
    with torch.no_grad():
        for idx, batch in enumerate(dev_loader):
            b = batch_to_device(batch, dev)
            out = model(*b, 'dev').cpu().numpy()
            whole, sub, tar = batch[0].numpy(), batch[7].numpy(), batch[1].numpy()
            for i in range(len(out)):
                sen = out[i].tolist()
                if vocab['<eos>'] in sen:
                    sen = sen[:sen.index(vocab['<eos>'])]
                for t in range(len(sen)):
                    if sen[t] >= args.vocab_size + args.sou_len:
                        sen[t] = int(sub[i][sen[t] - args.vocab_size - args.sou_len])
                    elif sen[t] >= args.vocab_size:
                        sen[t] = int(whole[i][sen[t] - args.vocab_size])
                hyp = ' '.join(r_vocab[x] for x in sen).replace('<pad>', "").replace('<unkm>', "😅").strip().split()
                ref = tar[i].tolist()
                ref = [r_vocab[x] for x in ref[1:ref.index(vocab['<eos>'])]]
                bleu = sentence_bleu_method2([ref], hyp)
                bleus += bleu
                out_str += ' '.join(deanonymise(hyp, var_maps[valid_index[total + i]])) + ',' + str(bleu) + '\n'
            total += len(out)
            if idx % 10 == 0:
                print("epoch: %d data: %d/%d bleu: %.4f" % (epoch, total, len(dev_loader.dataset), bleus / total))
    return bleus / max(1, len(dev_loader.dataset)), out_str


def train_epoch(dp, model, train_loader, epoch, best_bleu, dev_loader, g, valid_index, dev_):
    model.train()
    graphed = isinstance(dp, GraphedTrainStep)
    total_data, total_loss = 0, (torch.zeros((), device=dev_) if graphed else 0.0)
    max_batches = int(os.environ.get("FIRA_MAX_BATCHES", 0))
    for idx, batch in enumerate(train_loader):
        if max_batches and idx >= max_batches:
            break
        if epoch >= 15 and idx % 10 == 0:
            if RANK == 0:
                cur_bleu, output_str = dev(model, dev_loader, g, valid_index, epoch, dev_)
                open('OUTPUT/train_process', 'a').write(
                    'epoch: {} batch: {} dev bleu: {} is better: {}\n'.format(epoch, idx, cur_bleu, cur_bleu > best_bleu))
                if cur_bleu > best_bleu:
                    best_bleu = cur_bleu
                    torch.save(model.state_dict(), "best_model.pt")
                    open('OUTPUT/dev_output', 'w').write(output_str)
            if WORLD > 1:
                dist.barrier()
            model.train()
        if graphed:
            loss_sum, n_tok = dp.step(batch)              # pinned host batch -> static buffers -> graph replay
            total_loss += loss_sum / n_tok                # stays on the device: no sync per step
        else:
            loss, _ = dp.step(batch_to_device(batch, dev_))
            total_loss += loss.item()
        total_data += (batch.B if hasattr(batch, "shape_key") else len(batch[0])) * WORLD
        if idx % 10 == 0 and RANK == 0:
            print("epoch: %d batch: %d/%d  data: %d/%d loss: %.4f" % (
                epoch, idx, len(train_loader), total_data, n_train_total(train_loader) * WORLD, float(total_loss) / 10))
            total_loss = torch.zeros((), device=dev_) if graphed else 0.0
    return best_bleu


def n_train_total(train_loader):
    return len(train_loader.indices) if isinstance(train_loader, PackedBatchLoader) else len(train_loader.dataset)


def main_train():
    dev_ = device()
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=dev_)
    g = load_globals()
    train_set = TransDataset(args, 'train')
    dev_set = TransDataset(args, 'valid')
    all_index = json.load(open('all_index'))
    lo, hi = shard_range(len(train_set), RANK, WORLD)               # graphs shard by commit
    if WORLD > 1:                                                   # equal step counts on every rank: floor-sized
        per = len(train_set) // WORLD                               # contiguous shards (never past the end)
        lo, hi = RANK * per, RANK * per + per
    dev_loader = loader(dev_set, args.batch_size, False)
    model = TransModel(args).to(dev_)
    if os.environ.get("FIRA_ENGINE", "graph") == "graph":
        train_loader = PackedBatchLoader(train_set, args.batch_size, args.vocab_size, shuffle=True,
                                         indices=range(lo, hi) if WORLD > 1 else None, multiples=(8, 16, 16),
                                         max_shapes=int(os.environ.get("FIRA_MAX_SHAPES", 24)), drop_last=WORLD > 1,
                                         bucket=int(os.environ.get("FIRA_BUCKET", 0)),
                                         # per-commit packed node rows (packed.py) by default; FIRA_LAYOUT=trimmed keeps
                                         # the padded batches cut to the batch maximum
                                         packed=os.environ.get("FIRA_LAYOUT", "packed") == "packed")
        from fira_icse_b200.optim import FlatAdam
        dp = GraphedTrainStep(model, args.batch_size, lambda ps: FlatAdam(ps, lr=args.lr, groups=model.flat_groups()),
                              edge_capacity=train_loader.edge_cap)
    else:
        train_loader = loader(train_set, args.batch_size, True, list(range(lo, hi)) if WORLD > 1 else None)
        dp = DataParallelStep(model, lambda ps: Adam(ps, args.lr, fused=True))
    best_bleu = -1
    for epoch in range(args.epoches):
        best_bleu = train_epoch(dp, model, train_loader, epoch, best_bleu, dev_loader, g, all_index['valid'], dev_)
    if RANK == 0 and not os.path.exists("best_model.pt"):
        torch.save(model.state_dict(), "best_model.pt")             # short runs never reach the epoch-15 dev gate
    if WORLD > 1:
        dist.destroy_process_group()


def test(model, test_loader, g, test_index, dev_, out_path="OUTPUT/output_fira"):
    """Beam search over the test split, one line per commit (run_model.py:187-380)."""
    vocab, r_vocab, var_maps = g["vocab"], g["r_vocab"], g["var_maps"]
    model.eval()
    total, bleus = 0, 0.0
    with open(out_path, 'w') as f:
        for idx, batch in enumerate(test_loader):
            b = batch_to_device(batch, dev_)
            seq, length, prob = beam_search(model, b[0], b[3], b[4], b[5], b[7], beam_size=args.beam_size,
                                            tar_len=args.tar_len, start_id=vocab['<start>'], eos_id=vocab['<eos>'],
                                            pad_id=vocab['<pad>'])
            best, blen = best_sequences(seq, length, prob)
            best, blen, tar = best.cpu().numpy(), blen.cpu().numpy(), batch[1].numpy()
            bleu_batch = 0.0
            for i in range(len(best)):
                hyp = ids_to_text(best[i][:blen[i]].tolist(), r_vocab)
                ref = tar[i].tolist()
                ref = [r_vocab[x] for x in ref[1:ref.index(vocab['<eos>'])]]
                bl = sentence_bleu_method2([ref], hyp)
                bleus += bl; bleu_batch += bl
                f.write(' '.join(deanonymise(hyp, var_maps[test_index[total + i]])) + '\n')
            f.flush()
            total += len(best)
            print("data: %d/%d bleu: %f" % (total, len(test_loader.dataset), bleu_batch / len(best)))
    return bleus / max(1, total)


def main_test():
    dev_ = device()
    g = load_globals()
    test_set = TransDataset(args, 'test')
    all_index = json.load(open('all_index'))
    model = TransModel(args)
    model.load_state_dict(torch.load("best_model.pt", map_location="cpu"))
    model = model.to(dev_)
    lo, hi = shard_range(len(test_set), RANK, WORLD)                # replicas only: index ranges, files concatenated
    idx = list(range(lo, hi)) if WORLD > 1 else None
    test_loader = loader(test_set, args.test_batch_size, False, idx)
    out = "OUTPUT/output_fira" if WORLD == 1 else f"OUTPUT/output_fira.part{RANK:02d}"
    bleu = test(model, test_loader, g, all_index['test'][lo:hi], dev_, out)
    print("mean sentence bleu: %f" % bleu)


if __name__ == '__main__':
    stage = str(sys.argv[1])
    seed_everything()
    os.makedirs('OUTPUT', exist_ok=True)
    if stage == 'train':
        main_train()
    elif stage == 'test':
        main_test()
    else:
        raise SystemExit("usage: python run_model.py train|test")

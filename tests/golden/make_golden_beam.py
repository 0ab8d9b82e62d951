#!/usr/bin/env python
"""Golden beam-search output of the UNMODIFIED reference `run_model.test()` (run_model.py:187-380).

Runs only in the build container.  The reference imports nltk for BLEU reporting; nltk is not
installed, so `nltk.translate.bleu_score` is stubbed (BLEU is printed, never used for ranking).
Weights: the reference TransModel under torch.manual_seed(0) (no trained checkpoint is shipped) with
the output projections scaled x20 (SHARPEN) so that the distributions are peaked like a trained
model's -- with raw random weights every step has p ~ 1e-4, the fp32 probability PRODUCTS of
run_model.py:271 underflow to 0 after ~12 steps and the ranking degenerates into sort tie-breaking;
inputs: the first N_COMMITS commits of tests/golden/batch_first128.npz, test batch 8, beam 3.
Writes tests/golden/beam_first16.npz (chosen sequence per commit, -1 padded).
`--beam 5` writes tests/golden/beam5_first16.npz (BASELINE.json config 4 also names beam 5).
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
N_COMMITS, BATCH, SHARPEN = 16, 8, 20.0
BEAM = int(sys.argv[sys.argv.index("--beam") + 1]) if "--beam" in sys.argv else 3


def main():
    torch.set_num_threads(8)
    scratch = tempfile.mkdtemp(prefix="fira_beam_")
    os.symlink(os.path.join(REF, "DataSet"), os.path.join(scratch, "DataSet"))
    os.symlink(os.path.join(REF, "VOCAB_UPPER_CASE"), os.path.join(scratch, "VOCAB_UPPER_CASE"))
    shutil.copy(os.path.join(REF, "all_index"), os.path.join(scratch, "all_index"))
    os.makedirs(os.path.join(scratch, "OUTPUT"))
    os.chdir(scratch)
    sys.path.insert(0, REF)
    stub = types.ModuleType("nltk.translate.bleu_score")

    class SmoothingFunction:
        def method2(self, *a, **k):
            return None
    stub.SmoothingFunction = SmoothingFunction
    stub.sentence_bleu = lambda refs, hyp, smoothing_function=None: 0.0
    nltk = types.ModuleType("nltk"); tr = types.ModuleType("nltk.translate")
    nltk.translate = tr; tr.bleu_score = stub
    sys.modules.update({"nltk": nltk, "nltk.translate": tr, "nltk.translate.bleu_score": stub})
    import run_model as R
    R.args['beam_size'] = BEAM

    g = np.load(os.path.join(HERE, "batch_first128.npz"))
    ptr = g["edge_ptr"]

    def dense(i):
        a = np.zeros((650, 650))
        a[g["edge_row"][ptr[i]:ptr[i + 1]], g["edge_col"][ptr[i]:ptr[i + 1]]] = g["edge_val"][ptr[i]:ptr[i + 1]]
        return a

    class Loader(list):
        dataset = list(range(N_COMMITS))
    loader = Loader()
    for lo in range(0, N_COMMITS, BATCH):
        sl = slice(lo, lo + BATCH)
        t = lambda k: torch.from_numpy(g[k][sl].astype(np.int64))
        loader.append([t("sou"), t("tar"), t("attr"), t("mark"), t("ast_change"),
                       torch.from_numpy(np.stack([dense(i) for i in range(lo, lo + BATCH)])), t("tar_label"),
                       t("sub_token")])
    torch.manual_seed(0)
    model = R.TransModel(R.args)
    with torch.no_grad():
        model.out_fc.weight *= SHARPEN; model.out_fc.bias *= SHARPEN; model.copy_net.LinearRes.weight *= SHARPEN
    calls = []
    orig = R.convert_ids_to_tokens

    def spy(ids, r_vocab):
        calls.append(list(int(x) for x in ids))
        return orig(ids, r_vocab)
    R.convert_ids_to_tokens = spy
    R.test(model, loader)
    hyps = calls[0::2]
    assert len(hyps) == N_COMMITS
    out = np.full((N_COMMITS, 30), -1, np.int64)
    for i, h in enumerate(hyps):
        out[i, :len(h)] = h
    name = "beam_first16.npz" if BEAM == 3 else f"beam{BEAM}_first16.npz"
    np.savez_compressed(os.path.join(HERE, name), beam_ids=out, batch=BATCH, beam=BEAM, sharpen=SHARPEN)
    print(out[:4])


if __name__ == "__main__":
    main()

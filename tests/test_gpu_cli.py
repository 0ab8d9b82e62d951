"""`python run_model.py train|test` end to end on a 128-commit DataSet directory, and beam-search id
parity against the reference's own test() loop (tests/golden/beam_first16.npz)."""
import copy
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from fira_testlib import GOLDEN, ROOT, golden_batch, load_raw_golden, seeded_model
from test_data import _write_dataset

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("golden", ["beam_first16.npz", "beam5_first16.npz"])      # beam 3 (run_model.py:43), beam 5
@pytest.mark.parametrize("mode", ["full", "incremental", "graph"])
def test_beam_search_ids_match_reference_test_loop(mode, golden):
    """mode: full decoder re-run per step / KV-cached newest row / the same replayed as CUDA graphs
    (first batch captures, later batches replay); goldens = outputs of the unmodified reference's test() loop
    (tests/golden/make_golden_beam.py) with beam 3 (the reference default) and beam 5 (BASELINE.json configs[3])"""
    from fira_icse_b200.beam import beam_search, best_sequences
    gold = np.load(os.path.join(GOLDEN, golden))
    raw = load_raw_golden()
    vocab = raw["word_vocab"]
    model = copy.deepcopy(seeded_model()).to(DEV).eval()
    with torch.no_grad():                      # same sharpening as tests/golden/make_golden_beam.py
        k = float(gold["sharpen"])
        model.out_fc.weight *= k; model.out_fc.bias *= k; model.copy_net.LinearRes.weight *= k
    bs = int(gold["batch"])
    for lo in range(0, gold["beam_ids"].shape[0], bs):
        b = golden_batch(lo, lo + bs)
        seq, length, prob = beam_search(model, b[0], b[3], b[4], b[5].to(DEV), b[7], beam_size=int(gold["beam"]),
                                        tar_len=30, start_id=vocab["<start>"], eos_id=vocab["<eos>"],
                                        pad_id=vocab["<pad>"], mode=mode)
        best, blen = best_sequences(seq, length, prob)
        for i in range(bs):
            ref = gold["beam_ids"][lo + i]
            ref = ref[ref >= 0]
            mine = best[i, :blen[i]].cpu().numpy()
            assert np.array_equal(mine, ref), (lo + i, mine, ref)


def test_run_model_train_then_test(tmp_path):
    raw = load_raw_golden()
    _write_dataset(str(tmp_path), raw)
    env = dict(os.environ, PYTHONPATH=ROOT, FIRA_EPOCHS="1", FIRA_BATCH="16", FIRA_MAX_BATCHES="3",
               FIRA_WORKERS="0", FIRA_TEST_BATCH="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run_model.py"), "train"], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "loss:" in r.stdout
    sd = torch.load(tmp_path / "best_model.pt", map_location="cpu")
    assert len(sd) == 338 and not any(k.startswith("module.") for k in sd)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run_model.py"), "test"], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    n_test = len(json.load(open(tmp_path / "all_index"))["test"])
    lines = open(tmp_path / "OUTPUT" / "output_fira").read().split("\n")
    assert len(lines) == n_test + 1 and lines[-1] == ""

#!/usr/bin/env python
"""Host-side throughput of the commit builder on the first 128 real DataSet commits
(tests/golden/raw_first128.json.gz), one core, CPU only:
  * product: fira_icse_b200.data.build_commit (id conversion / labels in Python + fira_host_build_adjacency in C++),
  * the adjacency step alone (fira_host_build_adjacency),
  * the pure-Python restatement of the reference's Dataset.process_data (oracle/graph_oracle.py -- test infrastructure,
    timed here only as the baseline: the reference builds a dense 650x650 scipy matrix per commit).
    python tools/bench_builder.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import __graft_entry__
    __graft_entry__.build()
    import graph_oracle as G
    from fira_testlib import load_raw_golden
    from fira_icse_b200 import data
    raw = load_raw_golden()
    upper = set(raw["VOCAB_UPPER_CASE"])
    n = 128

    def timed(fn, reps):
        fn(0)
        t0 = time.perf_counter()
        for _ in range(reps):
            for i in range(n):
                fn(i)
        return (time.perf_counter() - t0) / (reps * n)
    t_prod = timed(lambda i: data.build_commit(raw["raw"], i, raw["word_vocab"], raw["ast_change_vocab"], upper), 10)
    calls = []
    orig = data.build_adjacency

    def spy(*a, **k):
        calls.append((a, k))
        return orig(*a, **k)
    data.build_adjacency = spy
    for i in range(n):
        data.build_commit(raw["raw"], i, raw["word_vocab"], raw["ast_change_vocab"], upper)
    data.build_adjacency = orig
    t0 = time.perf_counter()
    for _ in range(20):
        for a, k in calls:
            orig(*a, **k)
    t_adj = (time.perf_counter() - t0) / (20 * n)
    t_ref = timed(lambda i: G.build_commit(raw["raw"], i, raw["word_vocab"], raw["ast_change_vocab"], raw["VOCAB_UPPER_CASE"]), 2)
    print(json.dumps({"commits": n, "data": "first 128 commits of the real DataSet", "cores": 1,
                      "product_build_commit_ms": round(1e3 * t_prod, 4), "product_commits_per_s": round(1 / t_prod, 1),
                      "native_adjacency_only_ms": round(1e3 * t_adj, 4),
                      "python_restatement_of_reference_ms": round(1e3 * t_ref, 4),
                      "speedup_vs_restatement": round(t_ref / t_prod, 1)}))


if __name__ == "__main__":
    main()

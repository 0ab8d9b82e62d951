#!/usr/bin/env python
"""Summary of a torch.profiler chrome trace of graph-replayed training steps (bench.py --timeline PATH):
per-stream busy time, the span of one step, idle gaps between consecutive kernels of the busiest stream, kernels by
total time.  usage: python tools/timeline_summary.py trace.json [step_index]"""
import collections
import json
import sys


def load(path):
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    return ev


def short(name):
    n = name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("at::native::", "")
    return n.split("(")[0][:64]


def split_steps(ev, n_steps):
    """The trace holds n_steps replays of the same CUDA graph, separated by device synchronisations: the same number of
    kernels each, so the sorted list is cut evenly when the per-chunk kernel-name counts agree; otherwise (an eager run
    with shape-dependent launches) at the largest inter-kernel gaps."""
    if n_steps > 1 and len(ev) % n_steps == 0:
        k = len(ev) // n_steps
        chunks = [ev[i * k:(i + 1) * k] for i in range(n_steps)]
        names = [collections.Counter(short(e["name"]) for e in c) for c in chunks]
        if all(n == names[0] for n in names[1:]):
            return chunks
    gaps = sorted(((ev[i + 1]["ts"] - (ev[i]["ts"] + ev[i]["dur"]), i) for i in range(len(ev) - 1)), reverse=True)
    cuts = sorted(i for _, i in gaps[:n_steps - 1])
    out, lo = [], 0
    for c in cuts:
        out.append(ev[lo:c + 1])
        lo = c + 1
    out.append(ev[lo:])
    return out


def summarize(ev, out=sys.stdout):
    t0 = min(e["ts"] for e in ev)
    t1 = max(e["ts"] + e["dur"] for e in ev)
    span = t1 - t0
    streams = collections.defaultdict(list)
    for e in ev:
        streams[e["args"].get("stream", 0)].append(e)
    print(f"kernels {len(ev)}  span {span:.1f} us  sum of kernel time {sum(e['dur'] for e in ev):.1f} us", file=out)
    # union busy time over all streams
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print(f"GPU busy (union over streams) {busy:.1f} us = {100 * busy / span:.1f}% of the span; idle {span - busy:.1f} us", file=out)
    main = max(streams, key=lambda s: sum(e["dur"] for e in streams[s]))
    for s, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
        print(f"  stream {s}: {len(es)} kernels, busy {sum(e['dur'] for e in es):.1f} us{'  <- main' if s == main else ''}", file=out)
    es = streams[main]
    gaps = [(es[i + 1]["ts"] - (es[i]["ts"] + es[i]["dur"]), short(es[i]["name"]), short(es[i + 1]["name"])) for i in range(len(es) - 1)]
    pos = [g for g in gaps if g[0] > 0]
    print(f"main stream: {len(es)} kernels, sum of gaps {sum(g[0] for g in pos):.1f} us, median gap "
          f"{sorted(g[0] for g in gaps)[len(gaps) // 2]:.2f} us, mean kernel {sum(e['dur'] for e in es) / len(es):.2f} us", file=out)
    hist = collections.Counter(min(int(g[0]), 20) for g in gaps)
    print("  gap histogram (us: count):", " ".join(f"{k}:{hist[k]}" for k in sorted(hist)), file=out)
    print("  largest gaps:", file=out)
    for g in sorted(gaps, reverse=True)[:12]:
        print(f"    {g[0]:7.1f} us  after {g[1]}  before {g[2]}", file=out)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for e in ev:
        a = agg[short(e["name"])]
        a[0] += 1
        a[1] += e["dur"]
        if e["args"].get("stream", 0) == main:
            a[2] += e["dur"]
    tot = sum(a[1] for a in agg.values())
    print("kernels by total time (count, total us, share, of which on the main stream):", file=out)
    for k, (c, t, tm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"  {t:8.1f} us {100 * t / tot:5.1f}% n={c:4d} avg={t / c:6.2f} main={tm:8.1f}  {k}", file=out)


if __name__ == "__main__":
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    steps = split_steps(load(sys.argv[1]), n_steps)
    summarize(steps[len(steps) // 2])

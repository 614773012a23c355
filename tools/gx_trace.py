#!/usr/bin/env python
"""Summarise an LZB_GAP_TRACE log of the gapped scheduler (stderr of lastz_b200 / bench.py): when anchors started, finished
and were committed, and what the late starters waited for.  Reads the log on stdin; a measurement aid, not product code."""
import re
import sys

calls, cur = [], None
for line in sys.stdin:
    m = re.match(r"\[gx ([0-9.]+)\] (\w+) a=(\d+)(.*)", line)
    if not m:
        continue
    t, kind, a, rest = float(m.group(1)), m.group(2), int(m.group(3)), m.group(4)
    if cur is None or t < cur["last"] - 0.05:           # the clock restarts with every call
        cur = {"ev": {}, "last": 0.0, "order": []}
        calls.append(cur)
    cur["last"] = t
    d = cur["ev"].setdefault(a, {"start": [], "done": [], "commit": None, "wait": [], "retire": None, "misc": []})
    if kind == "start":
        d["start"].append(t); d["unsure"] = "unsure=1" in rest
    elif kind == "done":
        d["done"].append(t)
    elif kind == "commit":
        d["commit"] = t; cur["order"].append(a)
    elif kind == "wait":
        d["wait"].append((t, rest.strip()))
    elif kind == "retire":
        d["retire"] = t
    else:
        d["misc"].append((t, kind + rest[:60]))
for ci, c in enumerate(calls):
    ev = c["ev"]
    if len(c["order"]) < 5:
        continue
    commits = sorted(d["commit"] for d in ev.values() if d["commit"] is not None)
    starts = sorted(t for d in ev.values() for t in d["start"])
    print(f"call {ci}: {len(starts)} anchors started, {len(commits)} committed, {sum(1 for d in ev.values() if d['retire'] is not None)} retired while holding a lane, wall {c['last']:.3f}")

    def hist(ts, step=0.05):
        h = {}
        for t in ts:
            h[round(t // step * step, 2)] = h.get(round(t // step * step, 2), 0) + 1
        return " ".join(f"{k:.2f}:{v}" for k, v in sorted(h.items()))
    print("  starts  ", hist(starts))
    print("  commits ", hist(commits))
    first = starts[0] if starts else 0
    late = [(a, d) for a, d in ev.items() if d["commit"] is not None and d["start"] and d["start"][0] > first + 0.1]
    print(f"  committed anchors that started late: {len(late)}")
    for a, d in sorted(late, key=lambda x: x[1]["start"][0])[:25]:
        w = d["wait"][-1][1] if d["wait"] else ""
        print(f"    a={a} start {d['start'][0]:.3f} done {[round(x, 3) for x in d['done']]} commit {d['commit']:.3f}  last wait: {w}")

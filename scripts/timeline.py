#!/usr/bin/env python3
"""Row-block timeline of a single-pair long-mode run (debug hook BA_DEBUG_TS of the engine): start-to-start lag, durations.
usage: scripts/timeline.py <2|5> [col_chunks]     (run through gpurun)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bialign_b200 import workloads
from bialign_b200.batch import BatchAligner

which = sys.argv[1]
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
path = f"/tmp/ba_ts_{which}_{chunks}.txt"
os.environ["BA_DEBUG_TS"] = path
if which == "2":
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "dnapol1.json")))
    al = BatchAligner(**g["params"])
    res, cls, off = al.encode([g["seqA"], g["seqB"]], [g["strA"], g["strB"]])
    pa, pb = np.array([0], np.int32), np.array([1], np.int32)
else:
    al = BatchAligner(max_shift=3, **workloads.PROTEIN_PARAMS)
    res, cls, off, pa, pb = workloads.protein_pairs(1, lo=8192, hi=8192, seed=5)
al.set_option("col_chunks", chunks)
al.configure()
eng = al.engine
eng.load_sequences(res, cls, off); eng.load_pairs(pa, pb)
for _ in range(3):
    eng.run(want_trace=True)
st = eng.stats()
t = np.loadtxt(path, dtype=np.int64).reshape(-1, 9)
start, end = t[:, 1] / 1e3, t[:, 2] / 1e3  # us
marks = t[:, 3:9] / 1e3 - start[:, None]  # us after the tile's start at which iterations 0, 50, 100, 200, 400, 800 were reached
with np.printoptions(precision=1, suppress=True):
    print("  iterations 0 / 50 / 100 / 200 / 400 / 800 reached, us after the start of the row block (median over row blocks 1..):",
          np.median(marks[1:], axis=0))
ntc = max(chunks, 1)
ntiles = len(t)
print(f"cfg{which} chunks={chunks}: fill {st['fill_ms']:.3f} ms, {ntiles} tiles, span {end.max():.1f} us")
dur = end - start
print(f"  tile duration us: median {np.median(dur):.1f}  p10 {np.percentile(dur,10):.1f}  p90 {np.percentile(dur,90):.1f}")
if ntc == 1:
    d = np.diff(start)
    print(f"  start-to-start lag us: median {np.median(d):.2f}  mean {d.mean():.2f}  p10 {np.percentile(d,10):.2f}  p90 {np.percentile(d,90):.2f}")
    for lo in range(0, ntiles, max(1, ntiles // 8)):
        hi = min(ntiles, lo + max(1, ntiles // 8))
        print(f"    row blocks {lo:4d}-{hi-1:4d}: first start {start[lo]:9.1f}  last start {start[hi-1]:9.1f}  mean lag {np.diff(start[lo:hi]).mean() if hi - lo > 1 else 0:7.2f}  mean duration {dur[lo:hi].mean():8.1f}")
else:
    s2 = start.reshape(-1, ntc); e2 = end.reshape(-1, ntc)
    for c in range(ntc):
        d = np.diff(s2[:, c])
        print(f"    chunk {c:2d}: first start {s2[0, c]:9.1f}  last end {e2[-1, c]:9.1f}  lag median {np.median(d):6.2f} mean {d.mean():6.2f}  duration median {np.median(e2[:, c] - s2[:, c]):7.1f}")
    busy = dur.sum()
    print(f"  sum of tile durations {busy/1e3:.1f} ms = {busy / end.max():.1f} CTAs busy on average")

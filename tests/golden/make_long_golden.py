#!/usr/bin/env python3
"""Golden answers for the long single-pair configs, computed by the CPU oracle (oracle/bialign_oracle.c,
itself pinned to the reference by tests/test_oracle.py) -- the reference cannot run these sizes
(config 5 needs 237 GB of int64 tables).  Build container only; ~35 min single-threaded for cfg5.
Writes tests/golden/long_pairs.json: score, trace length, sha256 of the trace (hex string of column codes)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from bialign_b200 import workloads  # noqa: E402

out = {}
for name, length, s, seed in (("cfg5_small", 1024, 3, 5), ("cfg5", 8192, 3, 5)):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    res, cls, off, pa, pb = workloads.protein_pairs(1, lo=length, hi=length, seed=seed)
    a, sa = workloads.decode_protein(res, cls, off, 0)
    b, sb = workloads.decode_protein(res, cls, off, 1)
    params = dict(workloads.PROTEIN_PARAMS, max_shift=s)
    r = oracle.run(a, b, sa, sb, params, mode="codes")
    out[name] = dict(length=length, max_shift=s, seed=seed, score=r["score"], trace_len=len(r["trace"]),
                     trace_sha256=hashlib.sha256(r["trace"].encode()).hexdigest(), complete=r["complete"])
    print(name, out[name], flush=True)
    path = os.path.join(HERE, "long_pairs.json")
    prev = json.load(open(path)) if os.path.exists(path) else {}
    prev.update(out)
    json.dump(prev, open(path, "w"), indent=1)

import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from test_gpu_parity import _random_protein_batch, _decode_codes
from bialign_b200.batch import BatchAligner, trace_hex
warps, pad = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(900 + warps + 10 * pad)
for s in (1, 2, 3):
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50, shift_cost=-150, max_shift=s)
    seqs, structs, pairs = _random_protein_batch(rng, 3, 90, 260)
    al = BatchAligner(**params)
    al.set_option("kernel", 1); al.set_option("pad", pad); al.set_option("warps_per_cta", warps); al.set_option("long", 1)
    for rep in range(3):
        scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
        kind = al.engine.stats()["kernel_kind"]
        W = 2 * s + 1
        for q, (ia, ib) in enumerate(pairs):
            n, m = len(seqs[ia]), len(seqs[ib])
            r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes", want_codes=True)
            try:
                words = al.engine.debug_codes(q, r["codes"].size)
            except Exception as ex:
                print("no codes", ex); continue
            oc = r["codes"]
            valid = oc != np.uint64(0xFFFFFFFFFFFFFFFF)
            want = _decode_codes(oc, 0); got = _decode_codes(words, kind)
            reach = np.stack([((oc >> np.uint64(36 + t)) & np.uint64(1)).astype(bool) for t in range(9)], axis=1)
            reach &= (want != 15) & valid[:, None]
            bad = np.argwhere(reach & (want != got))
            cells = np.unique(bad[:, 0]) if len(bad) else []
            ii = sorted(set(int(c // (W * (m + 1) * W)) for c in cells))
            print(f"s={s} rep={rep} pair {q} n={n} m={m} kind={kind} score {int(scores[q])} vs {r['score']} trace_ok={trace_hex(cols, offsets, q) == r['trace']} bad={len(bad)} rows={ii[:12]}", flush=True)
            for cell, t in bad[:4]:
                bb = cell % W; j = (cell // W) % (m + 1); aa = (cell // (W * (m + 1))) % W; i = cell // (W * (m + 1) * W)
                print(f"    i={i} j={j} a={aa-s} b={bb-s} t={t} want={want[cell,t]} got={got[cell,t]}")

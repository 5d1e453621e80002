"""Debug helper: compare device code tables against the oracle and print the first mismatching cell-states."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from test_gpu_parity import _random_protein_batch, _decode_codes
from bialign_b200.batch import BatchAligner, trace_hex

def main(s, seed, npairs, lo, hi, warps=4, kernel=1, pad=-1, long=-1, extra=None):
    rng = np.random.default_rng(seed)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=s)
    if extra: params.update(extra)
    seqs, structs, pairs = _random_protein_batch(rng, npairs, lo, hi)
    al = BatchAligner(**params)
    al.set_option("kernel", kernel); al.set_option("warps_per_cta", warps)
    al.set_option("pad", pad); al.set_option("long", long)
    scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
    kind = al.engine.stats()["kernel_kind"]
    W = 2 * s + 1
    nbad = 0
    for q, (ia, ib) in enumerate(pairs):
        n, m = len(seqs[ia]), len(seqs[ib])
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes", want_codes=True)
        ok_s = int(scores[q]) == r["score"]; ok_t = trace_hex(cols, offsets, q) == r["trace"]
        words = al.engine.debug_codes(q, r["codes"].size)
        oc = r["codes"]
        valid = oc != np.uint64(0xFFFFFFFFFFFFFFFF)
        want = _decode_codes(oc, 0); got = _decode_codes(words, kind)
        reach = np.stack([((oc >> np.uint64(36 + t)) & np.uint64(1)).astype(bool) for t in range(9)], axis=1)
        reach &= (want != 15) & valid[:, None]
        bad = np.argwhere(reach & (want != got))
        print(f"pair {q} n={n} m={m} score_ok={ok_s} trace_ok={ok_t} bad={len(bad)} of {int(reach.sum())}", flush=True)
        for cell, t in bad[:8]:
            bb = cell % W; j = (cell // W) % (m + 1); aa = (cell // (W * (m + 1))) % W; i = cell // (W * (m + 1) * W)
            print(f"    i={i} j={j} a={aa-s} b={bb-s} t={t} want={want[cell,t]} got={got[cell,t]} word={int(words[cell]):x}")
        nbad += len(bad) + (not ok_s) + (not ok_t)
    print("TOTAL BAD", nbad)

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)

// GPU traceback: a pure pointer chase over the code table written by the fill kernels.
//
// The reference re-enumerates the cases at every visited cell and breaks ties by the "fewest
// shifts" key (pyx:547-571); because that key depends only on (predecessor cell, source state) the
// fill already stored the winner, so each step here is one read of 2-8 bytes, a field extract and a
// table decode.  The walk is a dependent chain of <= 2(n+m) reads, so throughput comes from running
// thousands of pairs side by side (one thread per pair).
#include "common.cuh"
#include "kernels.cuh"

namespace ba {

__constant__ int NA_XBITS_TB[13] = {15, 10, 5, 12, 3, 8, 4, 2, 1, 11, 7, 14, 13};  // pyx:233-248 order

struct Walk {
    long long slot0;  // first slot of the pair (systolic layout)
    int i, j, k, l, state, len, ok;
    bool first;  // the reference's very first termination test never fires (pyx:551, tuple vs list)
    uint8_t* out;
};

// Address of the code word of cell (i, j, k, l) in the table layout of the kernel that filled it.
__device__ __forceinline__ const void* code_addr(const TraceArgs& A, const uint64_t* codes, int m, int s, int nit_all, int nit_na,
                                                 int i, int j, int k, int l) {
    if (A.fmt == 3) return reinterpret_cast<const uint32_t*>(codes) + na_code_index(s, A.sysG, nit_na, i, j, l - j);
    return codes + code_index(m, s, i, j, k - i, l - j);
}

// One step of the walk; returns false when the walk ends (w.ok then tells whether it ended at the origin).
__device__ __forceinline__ bool walk_step(const TraceArgs& A, const uint64_t* codes, int m, int s, int nit_all, int nit_na, Walk& w) {
    int &i = w.i, &j = w.j, &k = w.k, &l = w.l;
    if ((i | j | k | l) == 0) {  // at the origin no case passes the guard (pyx:133-141): the walk ends here
        w.ok = (!w.first && w.state == 8) || A.fmt == 2;
        return false;
    }
    w.first = false;
    const bool planes = A.sysG > 0 && A.fmt != 3;  // systolic kernel: slot index into a 32-bit and a 16-bit plane
    const long long slot = planes ? w.slot0 + sys_code_index(A.R, A.LPR, A.P, s, A.sysG, nit_all, i, j, k - i, l - j) : 0;
    const void* addr = planes ? nullptr : code_addr(A, codes, m, s, nit_all, nit_na, i, j, k, l);
    if (A.fmt == 3) {  // dedicated non-affine kernel: one nibble per cell, the walk ends at the first cell without a case
        const uint32_t w32 = __ldg(reinterpret_cast<const uint32_t*>(addr));
        const int cidx = (int)((w32 >> (4 * (k - i + s))) & 15);
        if (cidx > 12) { w.ok = 1; return false; }
        const int xbn = NA_XBITS_TB[cidx];
        *--w.out = (uint8_t)xbn;
        ++w.len;
        i -= (xbn >> 3) & 1; j -= (xbn >> 2) & 1; k -= (xbn >> 1) & 1; l -= xbn & 1;
        if ((i | j | k | l) < 0 || abs(k - i) > s || abs(l - j) > s) { w.ok = 0; return false; }
        return true;
    }
    const uint64_t wd = planes ? 0 : __ldg(reinterpret_cast<const uint64_t*>(addr));
    int id;
    if (A.fmt == 2) {  // non-affine: the walk ends when no case reproduces the value (origin), pyx:521-528
        const int cidx = planes ? (int)(__ldg(reinterpret_cast<const uint32_t*>(A.codes) + slot) & 15) : (int)(wd & 15);
        if (cidx > 12) { w.ok = 1; return false; }
        const int xbn = NA_XBITS_TB[cidx];
        *--w.out = (uint8_t)xbn;
        ++w.len;
        i -= (xbn >> 3) & 1; j -= (xbn >> 2) & 1; k -= (xbn >> 1) & 1; l -= xbn & 1;
        return true;
    }
    const int state = w.state;
    if (A.fmt == 0) {
        id = (int)((wd >> (4 * state)) & 15);
    } else {
        // systolic format: 5-bit fields, state t < 6 at bit 2 + 5t of the low word, t >= 6 at bit 17 + 5(t-6) of the high
        // word (the fill pushes them in with funnel shifts).  The field is the id part of the winner's tie-break: a source
        // state src carries 27 - src;
        //   19..27        full column, case id = source = 27 - f                     (ids 0-8)
        //    9..17        x=(0,0,t2,t3), source (t01, h): f = 17 - 3*t01 - rank(h)   (ids 9-11, h = 11,10,01)
        //    0..8         x=(t0,t1,0,0), source (h, t23): f = 8 - 3*rank(h) - t23    (ids 12-14)
        // -- of which only the upper half is kept, in the 16-bit plane.  One read per step: 4 or 2 bytes.
        const int f = state < 6 ? (int)((__ldg(reinterpret_cast<const uint32_t*>(A.codes) + slot) >> (2 + 5 * state)) & 31)
                                : (int)((__ldg(A.codes_hi + slot) >> (1 + 5 * (state - 6))) & 31);
        const int t01 = state / 3, t23 = state % 3;
        id = 15;
        if (f >= 19 && f <= 27) id = 27 - f;
        else if (f >= 9 && f <= 17) {
            const int rk = 17 - 3 * t01 - f;
            if (rk >= 0 && rk <= 2) id = 9 + (2 - rk);
        } else if (f <= 8) {
            const int num = 8 - t23 - f;
            if (num >= 0 && num <= 6 && num % 3 == 0) id = 12 + (2 - num / 3);
        }
    }
    if (id == 15) return false;  // no case reproduced the value (pyx:570-571)
    int xb, src;
    decode_case(state, id, xb, src);
    *--w.out = (uint8_t)xb;
    ++w.len;
    i -= (xb >> 3) & 1; j -= (xb >> 2) & 1; k -= (xb >> 1) & 1; l -= xb & 1;
    w.state = src;
    if ((i | j | k | l) < 0 || abs(k - i) > s || abs(l - j) > s) return false;  // never follows a code out of the band
    return true;
}

// One thread per pair: thousands of independent walks hide the read latency of each.  (A warp per pair whose idle lanes
// prefetch the diagonal ahead of the walker was measured on the 928 x 933 and 8192 x 8192 pairs: 0.61 vs 0.55 ms and
// 8.3 vs 7.8 ms -- the path leaves the predicted diagonal too often -- and dropped.  So was a warp per pair whose lanes
// prefetch, while the walker's read is in flight, EVERY predecessor of the current cell in both planes, each followed by
// zero to two match columns: 0.84 vs 0.52 ms and 9.3 vs 7.8 ms -- the index arithmetic of the prefetches costs the lone
// warp more than the hits save.)
__global__ void __launch_bounds__(128) traceback_kernel(TraceArgs A) {
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= A.npairs) return;
    const PairDesc d = A.pairs[pi];
    const uint64_t* codes = A.codes + d.code_off;
    const int s = A.s, n = d.n, m = d.m;
    // iterations per row block in the fast kernels' tables (fill_systolic.cuh: nit + PRE; fill_na.cu likewise)
    const int nit_all = A.sysG ? (m + 1) * A.P + 2 * (A.sysG * A.R - 1) + A.LPR + A.RING + 4 : 0;
    const int Pna = 2 * s + 1 < 2 ? 2 : 2 * s + 1;
    const int nit_na = A.fmt == 3 ? (m + 1) * Pna + A.sysG * 32 + Pna + 1 : 0;
    Walk w;
    w.i = n; w.j = m; w.k = n; w.l = m;
    w.state = A.start_state[d.orig];
    w.out = A.trace + d.trace_off + d.trace_cap;  // one past the end of the slot
    w.len = 0; w.ok = 0; w.first = true; w.slot0 = d.code_off;
    while (w.len < d.trace_cap && walk_step(A, codes, m, s, nit_all, nit_na, w)) {}
    A.trace_len[d.orig] = w.len;
    A.complete[d.orig] = (uint8_t)w.ok;
}

void launch_traceback(const TraceArgs& A, cudaStream_t st) {
    const int threads = 128;
    traceback_kernel<<<(A.npairs + threads - 1) / threads, threads, 0, st>>>(A);
}

}  // namespace ba

// GPU traceback: a pure pointer chase over the 4-bit code table written by the fill kernels.
//
// The reference re-enumerates the cases at every visited cell and breaks ties by the "fewest
// shifts" key (pyx:547-571); because that key depends only on (predecessor cell, source state) the
// fill already stored the winner, so each step here is one 8-byte read, a nibble extract and a
// table decode.  One thread per pair: the walk is a dependent chain of <= 2(n+m) HBM/L2 reads, so
// throughput comes from running thousands of pairs side by side, not from parallelism inside one.
#include "common.cuh"
#include "kernels.cuh"

namespace ba {

__constant__ int NA_XBITS_TB[13] = {15, 10, 5, 12, 3, 8, 4, 2, 1, 11, 7, 14, 13};  // pyx:233-248 order

__global__ void __launch_bounds__(128) traceback_kernel(TraceArgs A) {
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= A.npairs) return;
    const PairDesc d = A.pairs[pi];
    const uint64_t* codes = A.codes + d.code_off;
    const int s = A.s, n = d.n, m = d.m;
    // iterations per row block in the systolic kernel's table (fill_systolic.cuh: nit + PRE)
    const int nit_all = A.sysG ? (m + 1) * A.P + 2 * (A.sysG * A.R - 1) + A.LPR + A.RING + 4 : 0;
    const int nit_na = A.fmt == 3 ? (m + 1) * (2 * s + 1 < 2 ? 2 : 2 * s + 1) + A.sysG * 32 + (2 * s + 1 < 2 ? 2 : 2 * s + 1) + 1 : 0;
    int i = n, j = m, k = n, l = m;
    int state = A.start_state[d.orig];
    uint8_t* out = A.trace + d.trace_off + d.trace_cap;  // one past the end of the slot
    int len = 0, ok = 0;
    bool first = true;  // the reference's very first termination test never fires (pyx:551, tuple vs list)
    while (len < d.trace_cap) {
        if ((i | j | k | l) == 0) {  // at the origin no case passes the guard (pyx:133-141): the walk ends here
            ok = (!first && state == 8) || A.fmt == 2;
            break;
        }
        first = false;
        if (A.fmt == 3) {  // dedicated non-affine kernel: one nibble per cell, walk ends at the first cell without a case
            const uint32_t w32 = __ldg(reinterpret_cast<const uint32_t*>(codes) + na_code_index(s, A.sysG, nit_na, i, j, l - j));
            const int cidx = (int)((w32 >> (4 * (k - i + s))) & 15);
            if (cidx > 12) { ok = 1; break; }
            const int xbn = NA_XBITS_TB[cidx];
            *--out = (uint8_t)xbn;
            ++len;
            i -= (xbn >> 3) & 1; j -= (xbn >> 2) & 1; k -= (xbn >> 1) & 1; l -= xbn & 1;
            if ((i | j | k | l) < 0 || abs(k - i) > s || abs(l - j) > s) { ok = 0; break; }
            continue;
        }
        const uint64_t wd = __ldg(codes + (A.sysG ? sys_code_index(A.R, A.LPR, A.P, s, A.sysG, nit_all, i, j, k - i, l - j)
                                                  : code_index(m, s, i, j, k - i, l - j)));
        int id;
        if (A.fmt == 2) {  // non-affine: the walk ends when no case reproduces the value (origin), pyx:521-528
            const int cidx = (int)(wd & 15);
            if (cidx > 12) { ok = 1; break; }
            const int xbn = NA_XBITS_TB[cidx];
            *--out = (uint8_t)xbn;
            ++len;
            i -= (xbn >> 3) & 1; j -= (xbn >> 2) & 1; k -= (xbn >> 1) & 1; l -= xbn & 1;
            continue;
        }
        if (A.fmt == 0) {
            id = (int)((wd >> (4 * state)) & 15);
        } else {
            // systolic format: 5-bit fields, state t < 6 at bit 2 + 5t of the low word, t >= 6 at bit 17 + 5(t-6) of the high word
            // (the fill pushes them in with funnel shifts).  The field is the
            // id part of the winner's tie-break: a source state src carries 27 - src;
            //   19..27        full column, case id = source = 27 - f                     (ids 0-8)
            //    9..17        x=(0,0,t2,t3), source (t01, h): f = 17 - 3*t01 - rank(h)   (ids 9-11, h = 11,10,01)
            //    0..8         x=(t0,t1,0,0), source (h, t23): f = 8 - 3*rank(h) - t23    (ids 12-14)
            const unsigned half = state < 6 ? (unsigned)wd : (unsigned)(wd >> 32);
            const int f = (int)((half >> (state < 6 ? 2 + 5 * state : 17 + 5 * (state - 6))) & 31);
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
        if (id == 15) break;  // no case reproduced the value (pyx:570-571)
        int xb, src;
        decode_case(state, id, xb, src);
        *--out = (uint8_t)xb;
        ++len;
        i -= (xb >> 3) & 1; j -= (xb >> 2) & 1; k -= (xb >> 1) & 1; l -= xb & 1;
        state = src;
        if ((i | j | k | l) < 0 || abs(k - i) > s || abs(l - j) > s) break;  // never follows a code out of the band
    }
    A.trace_len[d.orig] = len;
    A.complete[d.orig] = (uint8_t)ok;
}

void launch_traceback(const TraceArgs& A, cudaStream_t st) {
    const int threads = 128;
    traceback_kernel<<<(A.npairs + threads - 1) / threads, threads, 0, st>>>(A);
}

}  // namespace ba

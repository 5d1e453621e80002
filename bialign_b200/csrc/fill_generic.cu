// Generic affine fill: one CTA per pair, wavefront over the 4-D diagonal i+j+k+l.
//
// Every recursion case lowers i+j+k+l by 1..4 (a column advances at least one of the four
// indices), so all cells of one level are independent and five rolling level buffers suffice.
// This kernel evaluates the fifteen cases of pyx:255-296 one by one with the fill-time comparator
// (value desc, tie key asc, case id asc) that reproduces pyx:555-564, for any max_shift and any
// int32-safe parameters.  It is the general-purpose device path (odd max_shift, positive gap
// opening, ...) and the in-library cross-check of the systolic kernel; it is not the fast path.
#include "common.cuh"
#include "kernels.cuh"

namespace ba {

__device__ __forceinline__ int half_score(int src_r, int x_r, int mu, int beta, int gamma) {
    // contribution of one of the two coupled alignments to a column (pyx:103-129)
    if (x_r == 2) return mu;
    return gamma + (src_r == x_r ? 0 : beta);
}

// V = int (default) or long long: the reference's tables are int64 (pyx:27-35); the wide instantiation is used when the
// int32 bound (n+m) * (one column's largest magnitude) >= 2^30 does not hold.
template <bool TRACE, class V>
__global__ void __launch_bounds__(256) fill_level_kernel(FillArgs A) {
    __shared__ int s_pair;
    const int s = A.sc.s, W = 2 * s + 1;
    const int beta = A.sc.beta, gamma = A.sc.gamma, Delta = A.sc.delta;
    for (;;) {
        if (threadIdx.x == 0) s_pair = atomicAdd(A.counter, 1);
        __syncthreads();
        const int pi = s_pair;
        __syncthreads();
        if (pi >= A.npairs) return;
        const PairDesc d = A.pairs[pi];
        const uint8_t* ra = A.res + d.offA;
        const uint8_t* rb = A.res + d.offB;
        const uint8_t* ca = A.cls + d.offA;
        const uint8_t* cb = A.cls + d.offB;
        const int n = d.n, m = d.m;
        const int items = (n + 1) * W * W;
        V* lv = reinterpret_cast<V*>(A.scratch) + (size_t)blockIdx.x * A.scratch_stride;
        const size_t lstride = (size_t)items * 9;
        uint64_t* codes = TRACE ? A.codes + d.code_off : nullptr;

        for (int tau = 0; tau <= 2 * (n + m); ++tau) {
            V* cur = lv + (size_t)(tau % 5) * lstride;
            for (int it = threadIdx.x; it < items; it += blockDim.x) {
                const int bb = it % W, aa = (it / W) % W, i = it / (W * W);
                const int a = aa - s, b = bb - s;
                const int twoj = tau - 2 * i - a - b;
                if (twoj < 0 || (twoj & 1)) continue;
                const int j = twoj >> 1;
                const int k = i + a, l = j + b;
                if (j > m || k < 0 || k > n || l < 0 || l > m) continue;
                if (tau == 0) {  // pyx:483-485
                    for (int t = 0; t < 9; ++t) cur[(size_t)it * 9 + t] = (t == 8) ? 0 : NEG;
                    if (TRACE) codes[code_index(m, s, 0, 0, 0, 0)] = 0xFFFFFFFFFULL;
                    continue;
                }
                const int mu1 = (i > 0 && j > 0) ? A.sim[(int)ra[i - 1] * A.sc.nsym + rb[j - 1]] : 0;
                const int mu2 = A.mu2 ? ((k > 0 && l > 0) ? A.mu2[A.mu2_off[d.orig] + (long long)(k - 1) * m + (l - 1)] : 0)
                                      : ((k > 0 && l > 0 && ca[k - 1] == cb[l - 1]) ? A.sc.w : 0);
                uint64_t word = 0;
                for (int t = 0; t < 9; ++t) {
                    const int r01 = t / 3, r23 = t % 3;
                    const int t0 = hb0(r01), t1 = hb1(r01), t2 = hb0(r23), t3 = hb1(r23);
                    V best = NEG;
                    int bid = 15, bk0 = 0, bk1 = 0;
                    // the three groups of pyx:275-296: (x0,x1,x2,x3), first case id
                    for (int g = 0; g < 3; ++g) {
                        const int x0 = (g == 1) ? 0 : t0, x1 = (g == 1) ? 0 : t1;
                        const int x2 = (g == 2) ? 0 : t2, x3 = (g == 2) ? 0 : t3;
                        const int p0 = i - x0, p1 = j - x1, p2 = k - x2, p3 = l - x3;
                        if (p0 < 0 || p1 < 0 || p2 < 0 || p3 < 0) continue;          // pyx:133-141
                        const int pa = p2 - p0, pb = p3 - p1;
                        if (pa > s || pa < -s || pb > s || pb < -s) continue;
                        const int nx = x0 + x1 + x2 + x3;
                        const V* src = lv + (size_t)((tau - nx) % 5) * lstride +
                                         ((size_t)(p0 * W + (pa + s)) * W + (pb + s)) * 9;
                        const int shift = Delta * ((x0 != x2) + (x1 != x3));
                        const int ncase = (g == 0) ? 9 : 3;
                        for (int c = 0; c < ncase; ++c) {
                            int s01, s23, id;
                            if (g == 0) { s01 = c / 3; s23 = c % 3; id = c; }
                            else if (g == 1) { s01 = r01; s23 = 2 - c; id = 9 + c; }
                            else { s01 = 2 - c; s23 = r23; id = 12 + c; }
                            int sc = shift;
                            if (g != 1) sc += half_score(s01, r01, mu1, beta, gamma);
                            if (g != 2) sc += half_score(s23, r23, mu2, beta, gamma);
                            const V v = src[3 * s01 + s23] + sc;
                            // tie key, a function of (predecessor cell, source state): pyx:541-545
                            const int T0 = pa + hb0(s01) - hb0(s23), T1 = pb + hb1(s01) - hb1(s23);
                            const int k1 = abs(T1), k0 = abs(T0) + k1;
                            if (bid == 15 || v > best || (v == best && (k0 < bk0 || (k0 == bk0 && k1 < bk1)))) {
                                best = v; bid = id; bk0 = k0; bk1 = k1;
                            }
                        }
                    }
                    cur[(size_t)it * 9 + t] = best;  // NEG when no case was emitted (pyx:299-303)
                    word |= (uint64_t)bid << (4 * t);
                }
                if (TRACE) codes[code_index(m, s, i, j, a, b)] = word;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const V* fin = lv + (size_t)((2 * (n + m)) % 5) * lstride + ((size_t)(n * W + s) * W + s) * 9;
            V best = fin[0];
            for (int t = 1; t < 9; ++t) best = max(best, fin[t]);
            // start state: first best state with the fewest shifts (pyx:573-582)
            int st = 0, bsh = 99;
            for (int t = 0; t < 9; ++t)
                if (fin[t] == best) {
                    const int r01 = t / 3, r23 = t % 3;
                    const int sh = (hb0(r01) != hb0(r23)) + (hb1(r01) != hb1(r23));
                    if (sh < bsh) { bsh = sh; st = t; }
                }
            A.scores[d.orig] = best;
            A.start_state[d.orig] = (uint8_t)st;
            for (int t = 0; t < 9; ++t) A.end_values[(size_t)d.orig * 9 + t] = (int)fin[t];  // (debug hook: truncated when wide)
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Non-affine model (gap_opening_cost == 0): one value per cell, the 13 cases of pyx:233-248 in that
// order, first case reproducing the maximum wins (pyx:522-528).  Same level wavefront; code word =
// case index 0..12 in the low nibble (15 = none).
// ---------------------------------------------------------------------------------------------
__constant__ int NA_XBITS[13] = {15, 10, 5, 12, 3, 8, 4, 2, 1, 11, 7, 14, 13};

template <bool TRACE, class V>
__global__ void __launch_bounds__(256) fill_level_nonaffine_kernel(FillArgs A) {
    __shared__ int s_pair;
    const int s = A.sc.s, W = 2 * s + 1;
    const int gamma = A.sc.gamma, Delta = A.sc.delta;
    for (;;) {
        if (threadIdx.x == 0) s_pair = atomicAdd(A.counter, 1);
        __syncthreads();
        const int pi = s_pair;
        __syncthreads();
        if (pi >= A.npairs) return;
        const PairDesc d = A.pairs[pi];
        const uint8_t* ra = A.res + d.offA;
        const uint8_t* rb = A.res + d.offB;
        const uint8_t* ca = A.cls + d.offA;
        const uint8_t* cb = A.cls + d.offB;
        const int n = d.n, m = d.m;
        const int items = (n + 1) * W * W;
        V* lv = reinterpret_cast<V*>(A.scratch) + (size_t)blockIdx.x * A.scratch_stride;
        uint64_t* codes = TRACE ? A.codes + d.code_off : nullptr;
        for (int tau = 0; tau <= 2 * (n + m); ++tau) {
            V* cur = lv + (size_t)(tau % 5) * items;
            for (int it = threadIdx.x; it < items; it += blockDim.x) {
                const int bb = it % W, aa = (it / W) % W, i = it / (W * W);
                const int a = aa - s, b = bb - s;
                const int twoj = tau - 2 * i - a - b;
                if (twoj < 0 || (twoj & 1)) continue;
                const int j = twoj >> 1;
                const int k = i + a, l = j + b;
                if (j > m || k < 0 || k > n || l < 0 || l > m) continue;
                if (tau == 0) {  // np.zeros, pyx:27 / pyx:464-465
                    cur[it] = 0;
                    if (TRACE) codes[code_index(m, s, 0, 0, 0, 0)] = 15;
                    continue;
                }
                const int mu1 = (i > 0 && j > 0) ? A.sim[(int)ra[i - 1] * A.sc.nsym + rb[j - 1]] : 0;
                const int mu2 = A.mu2 ? ((k > 0 && l > 0) ? A.mu2[A.mu2_off[d.orig] + (long long)(k - 1) * m + (l - 1)] : 0)
                                      : ((k > 0 && l > 0 && ca[k - 1] == cb[l - 1]) ? A.sc.w : 0);
                V best = NEG;
                int bid = 15;
                for (int c = 0; c < 13; ++c) {
                    const int xb = NA_XBITS[c];
                    const int x0 = (xb >> 3) & 1, x1 = (xb >> 2) & 1, x2 = (xb >> 1) & 1, x3 = xb & 1;
                    const int p0 = i - x0, p1 = j - x1, p2 = k - x2, p3 = l - x3;
                    if (p0 < 0 || p1 < 0 || p2 < 0 || p3 < 0) continue;
                    const int pa = p2 - p0, pb = p3 - p1;
                    if (pa > s || pa < -s || pb > s || pb < -s) continue;
                    int sc;  // pyx:233-248
                    if (c == 0) sc = mu1 + mu2;
                    else if (c <= 2) sc = gamma + gamma;
                    else if (c == 3) sc = mu1 + Delta;
                    else if (c == 4) sc = mu2 + Delta;
                    else if (c <= 8) sc = gamma + Delta;
                    else if (c <= 10) sc = gamma + mu2 + Delta;
                    else sc = gamma + mu1 + Delta;
                    const V v = lv[(size_t)((tau - (x0 + x1 + x2 + x3)) % 5) * items + (p0 * W + (pa + s)) * W + (pb + s)] + sc;
                    if (bid == 15 || v > best) { best = v; bid = c; }
                }
                cur[it] = best;
                if (TRACE) codes[code_index(m, s, i, j, a, b)] = (uint64_t)bid;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const V fin = lv[(size_t)((2 * (n + m)) % 5) * items + (n * W + s) * W + s];
            A.scores[d.orig] = fin;
            A.start_state[d.orig] = 0;
            for (int t = 0; t < 9; ++t) A.end_values[(size_t)d.orig * 9 + t] = (int)fin;
        }
        __syncthreads();
    }
}

void launch_fill_nonaffine(const FillArgs& A, int grid, bool trace, bool wide, cudaStream_t st) {
    if (wide) {
        if (trace) fill_level_nonaffine_kernel<true, long long><<<grid, 256, 0, st>>>(A);
        else fill_level_nonaffine_kernel<false, long long><<<grid, 256, 0, st>>>(A);
    } else {
        if (trace) fill_level_nonaffine_kernel<true, int><<<grid, 256, 0, st>>>(A);
        else fill_level_nonaffine_kernel<false, int><<<grid, 256, 0, st>>>(A);
    }
}

void launch_fill_generic(const FillArgs& A, int grid, bool trace, bool wide, cudaStream_t st) {
    if (wide) {
        if (trace) fill_level_kernel<true, long long><<<grid, 256, 0, st>>>(A);
        else fill_level_kernel<false, long long><<<grid, 256, 0, st>>>(A);
    } else {
        if (trace) fill_level_kernel<true, int><<<grid, 256, 0, st>>>(A);
        else fill_level_kernel<false, int><<<grid, 256, 0, st>>>(A);
    }
}

size_t generic_scratch_ints(int nmax, int s) {
    const int W = 2 * s + 1;
    return (size_t)5 * (nmax + 1) * W * W * 9;
}

}  // namespace ba

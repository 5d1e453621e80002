// Non-affine fill: dedicated kernel for the reference's single-table model (gap_opening_cost == 0; pyx:225-252 case
// list and scores, pyx:443-471 fill, pyx:513-531 traceback rule "first case in case order that reproduces the value").
//
// One value per cell, thirteen cases.  The affine kernel's lane = (row, band offset a) mapping would spend all its time
// moving that single value between lanes; here a LANE OWNS A ROW i AND ALL ITS BAND OFFSETS a = k-i (W = 2s+1 cells per
// iteration, statically unrolled), walks the linearised (j, b = l-j) axis, and lags the row above by ONE iteration:
//      cell (i, j, a, b) is computed at iteration  q = i_in_block + j*P + (b+s)          (P = max(W, 2) cells per column)
// so a column x = (x0,x1,x2,x3) reaches back  D(x) = x0 + (P-1)*x1 + x3  iterations, to band offset a + x0 - x2:
//   x0 = 0 (seven cases)  sources are this lane's own cells: a register delay line of the last P+1 iterations
//                         (x = 0010 is the neighbouring offset of the SAME iteration: offsets are swept in ascending order)
//   x0 = 1 (six cases)    sources are the lane above, 1, 2, P or P+1 iterations ago: 4*W warp shuffles per iteration;
//                         lane 0 takes them from an exchange block written by lane 31 of the warp above (or staged from
//                         the boundary stream the previous row block left in global memory)
// Band edges in a are compile-time (the case is simply not instantiated), band edges in b are a poison on the additive
// constant, the range guard (pyx:133-138) is "cells outside the pair hold minus infinity".  With traceback every value is
// value << 4 | (15 - case index): plain max is then (value desc, case order asc), the low nibble of the winner is the
// code, and it is cleared before the value is handed on.  Codes: one nibble per cell, W nibbles = one uint32 per lane and
// iteration, stored in computation order [row block][warp][iteration][lane] (128 contiguous bytes per warp).
#include <utility>

#include "common.cuh"
#include "kernels.cuh"

#ifndef BA_NA_RECV
#define BA_NA_RECV 1
#endif

namespace ba {
namespace na {

constexpr int PRE_MAX = 12;  // >= P + 2 for every instantiated band

template <int S>
struct Geo {
    static constexpr int W = 2 * S + 1;
    static constexpr int P = W < 2 ? 2 : W;   // S = 0 keeps one pad cell per column so that D(0100) = P-1 >= 1
    static constexpr int NH = P + 1;          // own history depth (iterations)
    static constexpr int RING = 2 * NH;       // exchange-block depth: two history periods (reads reach back P+1 = NH iterations),
                                              // so a steady block that starts on a half boundary has compile-time slots
    static constexpr int XW = 8;              // ints per exchange record (W <= 7 values)
    static constexpr int PRE = P + 1;         // iterations run before position 0 so that the staged row above is primed
    static constexpr int LA = 8;              // cp.async look-ahead of the boundary staging (iterations)
    static constexpr int PB = 16;             // landing-zone depth = 2 LA (a power of two)
};

__device__ __forceinline__ int addmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int lds32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void stg32o_if(const void* p, int v, bool c) {  // [p + OFF] = v iff c (no branch)
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.global.b32 [%0+%3], %1;\n}\n" ::"l"(p), "r"(v), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts32o_if(unsigned addr, int v, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.shared.s32 [%0+%3], %1;\n}\n" ::"r"(addr), "r"(v), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void cp_async4o_if(unsigned smem_dst, const void* gsrc, bool c) {  // 4 bytes, [gsrc + OFF] -> smem_dst iff c
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp cp.async.ca.shared.global [%0], [%1+%3], 4;\n}\n" ::"r"(smem_dst), "l"(gsrc), "r"((int)c), "n"(OFF) : "memory");
}

template <bool V> struct BC_ { static constexpr bool value = V; };
template <int V> struct IC_ { static constexpr int value = V; };
template <bool SEL, class F, int... U>
__device__ __forceinline__ void steady_block(F& f, const int q, std::integer_sequence<int, U...>) {
    ((f(BC_<true>{}, BC_<SEL>{}, IC_<U>{}, q + U), __syncthreads()), ...);
}
// One exchange record (8 ints, W of them used): predicated loads into / stores from exactly W registers
template <int W>
__device__ __forceinline__ void lds_rec(int (&v)[W], unsigned addr, bool c) {
    static_assert(W == 1 || W == 3 || W == 5 || W == 7, "odd band widths only");
    if constexpr (W == 1)
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp ld.shared.s32 %0, [%1];\n}\n" : "+r"(v[0]) : "r"(addr), "r"((int)c) : "memory");
    else if constexpr (W == 3)
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %4, 0;\n @pp ld.shared.v2.s32 {%0, %1}, [%3];\n @pp ld.shared.s32 %2, [%3+8];\n}\n"
                     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]) : "r"(addr), "r"((int)c) : "memory");
    else if constexpr (W == 5)
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %6, 0;\n @pp ld.shared.v4.s32 {%0, %1, %2, %3}, [%5];\n @pp ld.shared.s32 %4, [%5+16];\n}\n"
                     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]) : "r"(addr), "r"((int)c) : "memory");
    else
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %8, 0;\n @pp ld.shared.v4.s32 {%0, %1, %2, %3}, [%7];\n @pp ld.shared.v2.s32 {%4, %5}, [%7+16];\n @pp ld.shared.s32 %6, [%7+24];\n}\n"
                     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]) : "r"(addr), "r"((int)c) : "memory");
}
template <int W>
__device__ __forceinline__ void sts_rec(unsigned addr, const int (&v)[W], bool c) {
    if constexpr (W == 1)
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.shared.s32 [%0], %1;\n}\n" ::"r"(addr), "r"(v[0]), "r"((int)c) : "memory");
    else if constexpr (W == 3)
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %4, 0;\n @pp st.shared.v2.s32 [%0], {%1, %2};\n @pp st.shared.s32 [%0+8], %3;\n}\n" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"((int)c) : "memory");
    else if constexpr (W == 5)
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %6, 0;\n @pp st.shared.v4.s32 [%0], {%1, %2, %3, %4};\n @pp st.shared.s32 [%0+16], %5;\n}\n" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"((int)c) : "memory");
    else
        asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %8, 0;\n @pp st.shared.v4.s32 [%0], {%1, %2, %3, %4};\n @pp st.shared.v2.s32 [%0+16], {%5, %6};\n @pp st.shared.s32 [%0+24], %7;\n}\n" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"((int)c) : "memory");
}

template <int S, bool TRACE>
// (no register cap: a 96-register cap compiles without spills for max_shift <= 2 and gives 20 resident warps per SM instead
// of 12, but measured 236 instead of 356 G cells/s on the config-3 shape: not adopted)
__global__ void __launch_bounds__(256, 1) fill_na_kernel(SysArgs A) {
    using G_ = Geo<S>;
    constexpr int W = G_::W, P = G_::P, NH = G_::NH, RING = G_::RING, XW = G_::XW, PRE = G_::PRE, LA = G_::LA, PB = G_::PB;
    static_assert(W <= 7 && PRE <= PRE_MAX, "band too wide for this kernel");
    constexpr bool RECV = BA_NA_RECV && (S <= 2);  // wider bands: the second delay line would cost 56 registers (occupancy)
    extern __shared__ __align__(16) int smem[];
    const int G = blockDim.x >> 5, RT = G * 32;
    int* xs = smem;                                   // [(G+1)][RING][XW]   xs[0] = staged row above the block
    int* pb = xs + (G + 1) * RING * XW;               // [PB][XW]
    int* ssim = pb + PB * XW;                         // [(nsym+1)][nsym], last row zero
    const int nsym = A.sc.nsym;
    uint8_t* sresB = reinterpret_cast<uint8_t*>(ssim + (nsym + 1) * nsym);
    const int bpad = A.bpad, boff = A.boff;
    uint8_t* sclsB = sresB + bpad;
    __shared__ int s_pair;

    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const int TB = TRACE ? 4 : 0;
    const int NEGP = A.negp;
    const int w_p = A.w_p, kG2 = A.k_2g, kGD = A.k_gd, kD = A.k_d;
    constexpr int T_ = TRACE ? 1 : 0;

    const unsigned xin_b = smem_u32(xs + g * RING * XW), xout_b = smem_u32(xs + (g + 1) * RING * XW);
    for (int q = tid; q < (G + 1) * RING * XW + PB * XW; q += blockDim.x) smem[q] = NEGP;
    for (int q = tid; q < (nsym + 1) * nsym; q += blockDim.x) ssim[q] = (q < nsym * nsym) ? A.sim_p[q] : 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_pair = atomicAdd(A.counter, 1);
        __syncthreads();
        const int pi = s_pair;
        __syncthreads();
        if (pi >= A.npairs) return;
        const PairDesc d = A.pairs[pi];
        const int n = d.n, m = d.m;
        const uint8_t* ra = A.res + d.offA;
        const uint8_t* ca = A.cls + d.offA;
        for (int q = tid; q < bpad; q += blockDim.x) {
            const int l = q - boff;
            sresB[q] = (l >= 1 && l <= m) ? A.res[d.offB + l - 1] : 0;
            sclsB[q] = (l >= 1 && l <= m) ? A.cls[d.offB + l - 1] : 255;
        }
        __syncthreads();

        const int npass = (n + RT) / RT;
        const int nit = (m + 1) * P + RT;                  // iterations 0 .. nit-1 cover every lane's last cell
        const size_t bstride = (size_t)A.bnd_iters * XW;   // ints per boundary buffer
        int* bnd_base = A.bnd + (size_t)blockIdx.x * 2 * bstride;

        for (int pass = 0; pass < npass; ++pass) {
            const int rr = g * 32 + lane, i = pass * RT + rr;
            const bool has_in = pass > 0, has_out = pass + 1 < npass;
            const int* bnd_in = bnd_base + (size_t)((pass + 1) & 1) * bstride;
            int* bnd_out = bnd_base + (size_t)(pass & 1) * bstride;
            const int Ai = (i >= 1 && i <= n) ? ra[i - 1] : nsym;  // zero row for i = 0 and rows outside the pair
            const int* simrow = ssim + Ai * nsym;
            int cA[W];       // structure class of A at k = i + a (254 never matches)
            bool okA[W];     // band offset inside the pair: 0 <= k <= n, row inside the pair
#pragma unroll
            for (int aa = 0; aa < W; ++aa) {
                const int k = i + aa - S;
                okA[aa] = i <= n && k >= 0 && k <= n;
                cA[aa] = (okA[aa] && k >= 1) ? ca[k - 1] : 254;
            }
            const int q_origin = (i == 0) ? S : (int)0x80000000;               // cell (0,0,0,0): j = 0, b = 0
            const int q_end = (i == n) ? m * P + S + rr : (int)0x80000000;     // cell (n,m,n,m)
            uint32_t* cw = nullptr;
            if (TRACE) cw = reinterpret_cast<uint32_t*>(A.codes + d.code_off) + ((long long)pass * G + g) * (long long)(nit + PRE) * 32 + lane;

            // position one iteration before the first one (q = -PRE)
            int pos = -PRE - 1 - rr;
            int j = -((-pos + P - 1) / P);
            int bb = pos - j * P;
            int slot = (((-PRE - 1) % RING) + RING) % RING;
            int mu1 = 0;
            int h[NH][W];  // h[d-1][aa]: this lane's value at band offset aa, d iterations ago
#pragma unroll
            for (int dd = 0; dd < NH; ++dd)
#pragma unroll
                for (int aa = 0; aa < W; ++aa) h[dd][aa] = NEGP;
            // RECV: what the lane above sent, rU[d-1][aa] = its value at band offset aa, d iterations ago (one shuffle per band
            // offset and iteration; the four delays the cases need are taps of this line instead of four shuffles)
            int rU[RECV ? NH : 1][W];
#pragma unroll
            for (int dd = 0; dd < (RECV ? NH : 1); ++dd)
#pragma unroll
                for (int aa = 0; aa < W; ++aa) rU[dd][aa] = NEGP;

            // boundary I/O: thread e < W moves element e of the record of one iteration
            const bool io = tid < W;
            const int io_e = io ? tid : 0;
            const bool do_flush = has_out && io, do_stage = has_in && io;
            // running pointers / shared-memory byte addresses of this thread's record element (threads beyond the record shadow
            // element 0; their stores and copies are predicated off)
            int* fl_g = bnd_out + io_e;                                                   // record q-1 of iteration q = -PRE
            const int* st_g = bnd_in + (size_t)(LA + RT + 1) * XW + io_e;                 // record (q + LA + RT) of iteration q = -PRE
            const unsigned fl_s = smem_u32(xs + G * RING * XW + io_e), st_s = smem_u32(xs + io_e), pb_s = smem_u32(pb + io_e);
            int pq = (-PRE - 1 + 4 * PB) & (PB - 1);                                      // landing-zone slot of the previous iteration
            if (has_in) {  // prime: records of the producer's iterations q + RT for q = -PRE .. -PRE+LA-1
                for (int t0 = -PRE; t0 < LA - PRE; ++t0) {
                    const int rec = t0 + RT;
                    if (io && rec < nit) {
                        const unsigned dst = smem_u32(pb + ((t0 + 4 * PB) & (PB - 1)) * XW + io_e);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(bnd_in + (size_t)(rec + PRE + 1) * XW + io_e) : "memory");
                    }
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                }
            }
            __syncthreads();

            // Steady range of this warp [st_lo, st_hi): every lane has S < j <= m - max(S, 1) throughout (all column range
            // tests true, no origin / end cell) and every staged record exists.
            const int st_lo = (S + 1) * P + g * 32 + 31;
            int st_hi = (m - (S > 0 ? S : 1) + 1) * P + g * 32;
            if (has_in) st_hi = min(st_hi, nit - RT - LA);

            // ---- one iteration: W cells of this lane's row.  ST: steady form (no column tests, no end / origin).
            unsigned xin_cur = 0, xin_oth = 0, xout_cur = 0;  // steady blocks: the two halves of the exchange ring
            unsigned fl_cur = 0, fl_oth = 0, st_cur = 0;      //                ... of the CTA's output block / of the staged row
            bool half = false;
            auto iteration = [&](auto st_, auto sel_, auto u_, const int q) __attribute__((always_inline)) {
                constexpr bool ST = decltype(st_)::value;
                constexpr bool SEL = !ST || decltype(sel_)::value;  // band offsets below row 0 (k < 0) must read "minus infinity"
                constexpr int u = decltype(u_)::value;  // ST: position inside the block = slot inside the current half
                ++bb;
                if (bb == P) { bb = 0; ++j; }
                if (!ST) slot = (slot + 1 == RING) ? 0 : slot + 1;
                const int l = j + bb - S;
                const bool colok = ST || ((bb < W) && ((unsigned)j <= (unsigned)m) && ((unsigned)l <= (unsigned)m));
                if (bb == 0) mu1 = simrow[sresB[j + boff]];
                const int cB = sclsB[l + boff];

                // ---- flush the record of iteration q-1 (exchange block xs[G], written before the last barrier)
                if (has_out) {
                    if constexpr (ST) {  // slot u-1 of the current half, or the last slot of the other one
                        stg32o_if<u * XW * 4>(fl_g, lds32((u >= 1 ? fl_cur : fl_oth) + ((u - 1 + NH) % NH) * (XW * 4)), do_flush);
                    } else {
                        stg32o_if<0>(fl_g, lds32(fl_s + ((slot == 0) ? RING - 1 : slot - 1) * (XW * 4)), do_flush);
                        fl_g += XW;
                    }
                }

                // ---- the row above, 1, 2, P and P+1 iterations ago: shuffles; lane 0 reads the exchange block of the warp
                // above (xs[0]: the staged boundary) with predicated vector loads that overwrite the shuffle results
                int U1[W], U2[W], UP[W], UQ[W];
                if constexpr (RECV) {
#pragma unroll
                    for (int dd = NH - 1; dd >= 1; --dd)
#pragma unroll
                        for (int aa = 0; aa < W; ++aa) rU[dd][aa] = rU[dd - 1][aa];
#pragma unroll
                    for (int aa = 0; aa < W; ++aa) rU[0][aa] = __shfl_up_sync(0xffffffffu, h[0][aa], 1);
                    if constexpr (ST) lds_rec<W>(rU[0], (u >= 1 ? xin_cur : xin_oth) + ((u - 1 + NH) % NH) * (XW * 4), lane == 0);
                    else lds_rec<W>(rU[0], xin_b + (slot == 0 ? RING - 1 : slot - 1) * (XW * 4), lane == 0);
#pragma unroll
                    for (int aa = 0; aa < W; ++aa) { U1[aa] = rU[0][aa]; U2[aa] = rU[1][aa]; UP[aa] = rU[P - 1][aa]; UQ[aa] = rU[P][aa]; }
                } else {
#pragma unroll
                    for (int aa = 0; aa < W; ++aa) {
                        U1[aa] = __shfl_up_sync(0xffffffffu, h[0][aa], 1);
                        U2[aa] = __shfl_up_sync(0xffffffffu, h[1][aa], 1);
                        UP[aa] = __shfl_up_sync(0xffffffffu, h[P - 1][aa], 1);
                        UQ[aa] = __shfl_up_sync(0xffffffffu, h[P][aa], 1);
                    }
                    if constexpr (ST) {
                        // slot u of the current half; d iterations back: the same half when u >= d, else the other one
                        lds_rec<W>(U1, (u >= 1 ? xin_cur : xin_oth) + ((u - 1 + NH) % NH) * (XW * 4), lane == 0);
                        lds_rec<W>(U2, (u >= 2 ? xin_cur : xin_oth) + ((u - 2 + NH) % NH) * (XW * 4), lane == 0);
                        lds_rec<W>(UP, (u >= P ? xin_cur : xin_oth) + ((u - P + NH) % NH) * (XW * 4), lane == 0);
                        lds_rec<W>(UQ, xin_oth + u * (XW * 4), lane == 0);
                    } else {
                        int s1 = slot - 1, s2 = slot - 2, sp = slot - P, sq = slot - (P + 1);
                        if (s1 < 0) s1 += RING;
                        if (s2 < 0) s2 += RING;
                        if (sp < 0) sp += RING;
                        if (sq < 0) sq += RING;
                        lds_rec<W>(U1, xin_b + s1 * (XW * 4), lane == 0);
                        lds_rec<W>(U2, xin_b + s2 * (XW * 4), lane == 0);
                        lds_rec<W>(UP, xin_b + sp * (XW * 4), lane == 0);
                        lds_rec<W>(UQ, xin_b + sq * (XW * 4), lane == 0);
                    }
                }

                // ---- additive constants shared by all band offsets of this iteration (scores of pyx:233-248; the low
                // nibble carries 15 - case index when codes are wanted); pB0 / pB1: the source column b-1 / b+1 is outside the band
                const int pB0 = (bb == 0) ? NEGP : 0, pB1 = (bb == W - 1) ? NEGP : 0;
                const int c1 = kG2 + T_ * 14, c2 = kG2 + T_ * 13;
                const int c3 = mu1 + kD + pB1 + T_ * 12;
                const int c5 = kGD + T_ * 10, c6 = kGD + pB1 + T_ * 9, c7 = kGD + T_ * 8, c8 = kGD + pB0 + T_ * 7;
                const int c11 = mu1 + kGD + pB1 + T_ * 4, c12 = mu1 + kGD + T_ * 3;
                // the four cases whose second alignment is a match column (0, 4, 9, 10) share mu2: it is added once, to their maximum
                const int k0 = mu1 + T_ * 15, k4 = kD + pB0 + T_ * 11, k9 = kGD + pB0 + T_ * 6, k10 = kGD + T_ * 5;
                int M[W];
                uint32_t code = 0;
                // Three independent chains per cell (max is associative: the case order lives in the low nibble); inputs that
                // arrive late (U1: this iteration's shuffle) come last in their chain.  Then the one case that depends on this
                // iteration's neighbouring offset (0010).  (A running maximum over the cleared values, which takes the nibble
                // clearing off that chain, measured 2 % slower: five more instructions on the pipe that limits this kernel.)
                int rest[W];
#pragma unroll
                for (int aa = 0; aa < W; ++aa) {
                    const int mu2 = (cB == cA[aa]) ? w_p : 0;
                    int va = UQ[aa] + k0;                                            // 1111  case 0
                    va = addmax(U2[aa], k9, va);                                     // 1011  case 9
                    if (aa >= 1) va = addmax(h[P - 1][aa - 1], k10, va);             // 0111  case 10
                    if (aa >= 1) va = addmax(h[0][aa - 1], k4, va);                  // 0011  case 4
                    int vb = addmax(h[P - 1][aa], c2, NEGP);                         // 0101  case 2 (and the floor)
                    if (aa + 1 < W) vb = addmax(UP[aa + 1], c3, vb);                 // 1100  case 3
                    vb = addmax(U1[aa], c1, vb);                                     // 1010  case 1
                    if (aa + 1 < W) vb = addmax(U1[aa + 1], c5, vb);                 // 1000  case 5
                    int vc = h[P - 2][aa] + c6;                                      // 0100  case 6
                    vc = addmax(UP[aa], c11, vc);                                    // 1110  case 11
                    if (aa + 1 < W) vc = addmax(UQ[aa + 1], c12, vc);                // 1101  case 12
                    vc = addmax(h[0][aa], c8, vc);                                   // 0001  case 8
                    rest[aa] = addmax(va, mu2, max(vb, vc));
                }
                {
#pragma unroll
                    for (int aa = 0; aa < W; ++aa) {
                        int v = rest[aa];
                        if (aa >= 1) v = addmax(M[aa - 1], c7, v);                   // 0010  case 7 (same iteration)
                        if (SEL) v = (colok && okA[aa]) ? v : NEGP;
                        if (!ST && aa == S && q == q_origin) v = 0;                  // M[0,0,0,0] = 0 (numpy zeros, pyx:27-35)
                        if (TRACE) {
                            code = __funnelshift_r(code, (uint32_t)v, 4);  // low nibble (15 - case index) in at the top
                            v &= ~15;
                        }
                        M[aa] = v;
                    }
                }
                // nibble aa = case index of band offset aa; a cell no case reaches reads 15 ("none": NEGP has a zero low nibble)
                if (TRACE) { *cw = (~code) >> (32 - 4 * W); cw += 32; }

                if (!ST && q == q_end) {
                    const int best = M[S] >> TB;
                    A.scores[d.orig] = (long long)best * A.gscale;
                    A.start_state[d.orig] = 8;
#pragma unroll
                    for (int t = 0; t < 9; ++t) A.end_values[(size_t)d.orig * 9 + t] = best * A.gscale;
                }

                // ---- publish: the last lane of every warp feeds lane 0 of the warp below / the next row block
                if constexpr (ST) sts_rec<W>(xout_cur + u * (XW * 4), M, lane == 31);
                else sts_rec<W>(xout_b + slot * (XW * 4), M, lane == 31);
#pragma unroll
                for (int dd = NH - 1; dd >= 1; --dd)
#pragma unroll
                    for (int aa = 0; aa < W; ++aa) h[dd][aa] = h[dd - 1][aa];
#pragma unroll
                for (int aa = 0; aa < W; ++aa) h[0][aa] = M[aa];

                // ---- stage the row above this block for iteration q (what lane 0 of warp 0 reads from the next one on)
                if (has_in) {
                    asm volatile("cp.async.wait_group %0;\n" ::"n"(LA - 1) : "memory");
                    pq = (pq + 1) & (PB - 1);
                    int val = lds32(pb_s + pq * (XW * 4));
                    const unsigned land = pb_s + (pq ^ LA) * (XW * 4);  // slot of iteration q + LA (PB = 2 LA)
                    if constexpr (ST) {
                        sts32o_if<u * XW * 4>(st_cur, val, do_stage);
                        cp_async4o_if<u * XW * 4>(land, st_g, do_stage);
                    } else {
                        if (q + RT >= nit) val = NEGP;
                        sts32o_if<0>(st_s + slot * (XW * 4), val, do_stage);
                        cp_async4o_if<0>(land, st_g, do_stage && q + LA + RT < nit);
                        st_g += XW;
                    }
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                }
            };

            // whole history periods (NH iterations, statically unrolled: the delay-line registers are renamed, not moved)
            for (int q = -PRE; q < nit;) {
                const bool aligned = (slot == NH - 1) || (slot == RING - 1);  // the next slot starts a half of the exchange ring
                if (aligned && q >= st_lo && q + NH <= st_hi) {  // warp-uniform
                    half = (slot == NH - 1);                     // the block writes the second half
                    xin_cur = xin_b + (half ? NH : 0) * (XW * 4);
                    xin_oth = xin_b + (half ? 0 : NH) * (XW * 4);
                    xout_cur = xout_b + (half ? NH : 0) * (XW * 4);
                    fl_cur = fl_s + (half ? NH : 0) * (XW * 4);
                    fl_oth = fl_s + (half ? 0 : NH) * (XW * 4);
                    st_cur = st_s + (half ? NH : 0) * (XW * 4);
                    // Only band offsets below row 0 (k < 0: the first S rows of a pair) are ever read as "minus infinity" by valid
                    // cells; everything else outside the pair (rows beyond n, k > n) has larger coordinates than any valid cell
                    // and is never a source, so steady blocks elsewhere skip the per-cell validity select.
                    if (S > 0 && pass == 0 && g == 0) steady_block<true>(iteration, q, std::make_integer_sequence<int, NH>{});
                    else steady_block<false>(iteration, q, std::make_integer_sequence<int, NH>{});
                    q += NH;
                    slot = half ? RING - 1 : NH - 1;
                    fl_g += NH * XW;
                    st_g += NH * XW;
                } else {
                    iteration(BC_<false>{}, BC_<true>{}, IC_<0>{}, q);
                    __syncthreads();
                    ++q;
                }
            }
            if (has_out) {  // the record of the last iteration, then make the stream visible to the next row block
                if (io) bnd_out[(size_t)(nit + PRE) * XW + io_e] = xs[(G * RING + slot) * XW + io_e];
                __threadfence();
            }
            if (has_in) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            if (!has_out && has_in)  // leaving a multi-block pair: the staged row must read "minus infinity" again
                for (int qq = tid; qq < RING * XW; qq += blockDim.x) xs[qq] = NEGP;
            __syncthreads();
        }
    }
}

template <int S>
size_t smem_bytes_t(int G, int nsym, int bpad) {
    using G_ = Geo<S>;
    const size_t ints = (size_t)(G + 1) * G_::RING * G_::XW + (size_t)G_::PB * G_::XW + (size_t)(nsym + 1) * nsym;
    return ((ints * 4 + 2 * (size_t)bpad) + 15) & ~(size_t)15;
}

template <int S, bool TRACE>
cudaError_t launch_t(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    auto kern = fill_na_kernel<S, TRACE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, G * 32, smem, st>>>(A);
    return cudaGetLastError();
}
template <int S, bool TRACE>
int occ_t(int G, size_t smem) {
    auto kern = fill_na_kernel<S, TRACE>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G * 32, smem);
    return nb;
}

}  // namespace na

// ---- host-side geometry (mirrors na::Geo) and dispatch over max_shift
static int na_P(int S) { return 2 * S + 1 < 2 ? 2 : 2 * S + 1; }
int na_iters(int S, int G, int m) { return (m + 1) * na_P(S) + G * 32; }
int na_pre(int S) { return na_P(S) + 1; }
int na_boff(int S, int G) { return (G * 32 + 2 * na_P(S) + 16) / na_P(S) + 3 + S; }
int na_bpad(int S, int G, int mmax) { return na_boff(S, G) + mmax + (G * 32 + 2 * na_P(S) + 16) / na_P(S) + S + 8; }
size_t na_boundary_ints(int S, int G, int mmax) { return (size_t)(na_iters(S, G, mmax) + na_pre(S) + 8) * 8; }
long long na_code_words(int S, int G, int n, int m) {  // uint64 units: one uint32 per lane and iteration
    const int RT = G * 32;
    return ((long long)((n + RT) / RT) * G * (na_iters(S, G, m) + na_pre(S)) * 32 + 1) / 2;
}
size_t na_smem_bytes(int S, int G, int nsym, int mmax) {
    const int bpad = na_bpad(S, G, mmax);
    switch (S) {
        case 0: return na::smem_bytes_t<0>(G, nsym, bpad);
        case 1: return na::smem_bytes_t<1>(G, nsym, bpad);
        case 2: return na::smem_bytes_t<2>(G, nsym, bpad);
        default: return na::smem_bytes_t<3>(G, nsym, bpad);
    }
}
int na_occupancy(int S, bool trace, int G, size_t smem) {
    switch (S) {
        case 0: return trace ? na::occ_t<0, true>(G, smem) : na::occ_t<0, false>(G, smem);
        case 1: return trace ? na::occ_t<1, true>(G, smem) : na::occ_t<1, false>(G, smem);
        case 2: return trace ? na::occ_t<2, true>(G, smem) : na::occ_t<2, false>(G, smem);
        case 3: return trace ? na::occ_t<3, true>(G, smem) : na::occ_t<3, false>(G, smem);
    }
    return 0;
}
cudaError_t launch_fill_na(const SysArgs& A, int grid, int G, size_t smem, bool trace, cudaStream_t st) {
    switch (A.sc.s) {
        case 0: return trace ? na::launch_t<0, true>(A, grid, G, smem, st) : na::launch_t<0, false>(A, grid, G, smem, st);
        case 1: return trace ? na::launch_t<1, true>(A, grid, G, smem, st) : na::launch_t<1, false>(A, grid, G, smem, st);
        case 2: return trace ? na::launch_t<2, true>(A, grid, G, smem, st) : na::launch_t<2, false>(A, grid, G, smem, st);
        case 3: return trace ? na::launch_t<3, true>(A, grid, G, smem, st) : na::launch_t<3, false>(A, grid, G, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ba

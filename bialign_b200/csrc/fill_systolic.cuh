// Systolic affine fill: the fast path (template; instantiated per max_shift in fill_systolic_s*.cu).
//
// Mapping.  A lane owns one (row i, row offset a = k-i) pair and walks the linearised (j, b = l-j)
// axis one band cell per iteration; LPR lanes make a row, R = 32/LPR rows make a warp, G warps make
// a CTA that works as ONE systolic array of G*R rows in lock step.  Lane (row r, column c) is
// sigma = 2r + c iterations behind lane (0,0) and the b axis has period P, so every one of the 15
// predecessor columns x = (x0,x1,x2,x3) of pyx:255-296 is a fixed number of iterations
//      D(x) = x0 + x1*(P-1) + x2 + x3   >= 1
// in the past, at lane  lane - (x0*(LPR-1) + x2).  Cells outside the sequences hold "minus
// infinity" (NEGP), which replaces the range half of the reference's guard (pyx:133-138); the band
// half (pyx:139-140) comes in two flavours:
//   PAD = true   one always-invalid pad lane per row and one pad cell per column (LPR = P = 2S+2):
//                band-edge sources are pad cells, nothing to mask;
//   PAD = false  no pads (LPR = P = 2S+1, 20-30% more useful lanes and iterations): band-edge
//                sources are the wrong cells, so the additive constant of exactly those cases carries
//                a "poison" of NEGP; every value is clamped from below at NEGP, so at most three
//                NEGP can pile up -- the host only selects this flavour when that fits the integer
//                range (engine.cu, plan_systolic).
//
// Recurrence.  Push form: after a cell's nine state values M are known it publishes three 3x3
// blocks of partial maxima,
//      R[s01][x23] = max_{s23} M[s01][s23] + open(s23, x23)        (second alignment decided)
//      L[x01][s23] = max_{s01} M[s01][s23] + open(s01, x01)        (first alignment decided)
//      Q[x01][x23] = max_{s01} R[s01][x23] + open(s01, x01)        (both decided)
// with open(s, x) = beta if x is a gap half different from s, else 0.  A target state t then needs
// three values -- Q of the cell at -(t01,t23), R of the cell at -(00,t23), L of the cell at
// -(t01,00) -- plus constants built from mu1, mu2, gamma, Delta: affine_score (pyx:84-131) is
// separable in the two alignments.  BNEG (beta < 0) shortens open() to one fused add-max.
//
// Transport.  x1 = 0 values are 1-3 iterations old: warp shuffles from history registers.  x1 = 1
// values are P-1..P+2 iterations old: a per-warp shared-memory ring indexed by iteration.  Row 0 of
// a warp reads the ring (and the short-delay exchange block xs) of the warp above; row 0 of warp 0
// reads a staging ring fed by cp.async from the boundary stream that the previous row block
// ("pass") left in global memory.  One __syncthreads per iteration orders all of it.
//
// Traceback codes.  With TRACE the integers carry the tie-break of pyx:555-564 in their low bits:
// value << TB | (inverted rank of the tie key (|T0|+|T1|, |T1|) of (cell, source state)) << 5 |
// id field (27 - source state), so plain integer max implements (value desc, key asc, case id asc) exactly.  The id
// field of the winner (5 bits per state, 45 bits per cell) is streamed to HBM in computation order, six bytes
// per lane and iteration: a 32-bit word (states 0-5) and a 16-bit half (states 6-8) in two planes.
#pragma once
#include <utility>

#include "common.cuh"
#include "kernels.cuh"

namespace ba {
namespace sys {

__device__ __forceinline__ int vmax(int a, int b) { return max(a, b); }
__device__ __forceinline__ int vmax3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
__device__ __forceinline__ int addmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }  // max(a+b, c)

// 16-bit pair mode (P16): every int holds two independent signed 16-bit values, one per pair of a two-pair work
// item; the DPX SIMD forms do per-halfword add/max in one instruction.  Plain integer adds are not allowed on
// packed values (carries would cross the halves), hence vadd2 = max(a + b, -32768) per half.
__device__ __forceinline__ int pack2(int x) { return (x & 0xffff) | (x << 16); }
__device__ __forceinline__ int vadd2(int a, int b) { return (int)__viaddmax_s16x2((unsigned)a, (unsigned)b, 0x80008000u); }
template <bool P16>
__device__ __forceinline__ int xadd(int a, int b) { return P16 ? vadd2(a, b) : a + b; }
template <bool P16>
__device__ __forceinline__ int xaddmax(int a, int b, int c) {
    return P16 ? (int)__viaddmax_s16x2((unsigned)a, (unsigned)b, (unsigned)c) : __viaddmax_s32(a, b, c);
}
template <bool P16>
__device__ __forceinline__ int xmax3(int a, int b, int c) {
    return P16 ? (int)__vimax3_s16x2((unsigned)a, (unsigned)b, (unsigned)c) : __vimax3_s32(a, b, c);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ int lds32o(unsigned addr) {  // [addr + OFF]: the offset folds into the LDS immediate
    int v;
    asm volatile("ld.shared.s32 %0, [%1+%2];\n" : "=r"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(unsigned addr, int v) { asm volatile("st.shared.s32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void cp_async4s(unsigned smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    // 4-byte copies go through L1 (.ca); the boundary stream is written by this same SM (write-through
    // stores keep its L1 coherent), so no stale line can be observed in CTA-per-pair mode
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16s(unsigned smem_dst, const void* gsrc) {  // L2 only: coherent across SMs
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
// Predicated forms (no branch, so no divergence bookkeeping in the iteration loop and the compiler is free to move the
// feeding loads away from the store): the operation happens iff c.
__device__ __forceinline__ void stg32_if(void* p, int v, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.global.b32 [%0], %1;\n}\n" ::"l"(p), "r"(v), "r"((int)c) : "memory");
}
__device__ __forceinline__ void stg64_if(void* p, unsigned lo, unsigned hi, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %3, 0;\n @pp st.global.v2.b32 [%0], {%1, %2};\n}\n" ::"l"(p), "r"(lo), "r"(hi), "r"((int)c) : "memory");
}
__device__ __forceinline__ void sts32_if(unsigned addr, int v, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.shared.s32 [%0], %1;\n}\n" ::"r"(addr), "r"(v), "r"((int)c) : "memory");
}
__device__ __forceinline__ void cp_async4s_if(unsigned smem_dst, const void* gsrc, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp cp.async.ca.shared.global [%0], [%1], 4;\n}\n" ::"r"(smem_dst), "l"(gsrc), "r"((int)c) : "memory");
}
// Forms with a compile-time byte offset folded into the instruction (the statically unrolled "steady" iterations)
template <int OFF>
__device__ __forceinline__ void lds32o_if(int& v, unsigned addr, bool c) {  // v = [addr + OFF] iff c, else v keeps its value
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp ld.shared.s32 %0, [%1+%3];\n}\n" : "+r"(v) : "r"(addr), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts32o(unsigned addr, int v) { asm volatile("st.shared.s32 [%0+%2], %1;\n" ::"r"(addr), "r"(v), "n"(OFF) : "memory"); }
template <int OFF>
__device__ __forceinline__ void sts32o_if(unsigned addr, int v, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.shared.s32 [%0+%3], %1;\n}\n" ::"r"(addr), "r"(v), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void stg32o_if(const void* p, int v, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.global.b32 [%0+%3], %1;\n}\n" ::"l"(p), "r"(v), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void stg64o_if(const void* p, unsigned lo, unsigned hi, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %3, 0;\n @pp st.global.v2.b32 [%0+%4], {%1, %2};\n}\n" ::"l"(p), "r"(lo), "r"(hi), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void stg32o(const void* p, unsigned v) {
    asm volatile("st.global.b32 [%0+%2], %1;\n" ::"l"(p), "r"(v), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void stg16o(const void* p, unsigned v) {  // low 16 bits of v
    asm volatile("{\n .reg .b16 hh;\n cvt.u16.u32 hh, %1;\n st.global.b16 [%0+%2], hh;\n}\n" ::"l"(p), "r"(v), "n"(OFF) : "memory");
}
template <int SOFF, int GOFF>
__device__ __forceinline__ void cp_async4so_if(unsigned smem_dst, const void* gsrc, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp cp.async.ca.shared.global [%0+%3], [%1+%4], 4;\n}\n" ::"r"(smem_dst), "l"(gsrc), "r"((int)c), "n"(SOFF), "n"(GOFF) : "memory");
}
// predicated vector forms for the short-delay exchange block (row 0 reads / last row writes)
template <int OFF>
__device__ __forceinline__ void lds64o_if(int& x, int& y, unsigned addr, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %3, 0;\n @pp ld.shared.v2.s32 {%0, %1}, [%2+%4];\n}\n" : "+r"(x), "+r"(y) : "r"(addr), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void lds128o_if(int& x, int& y, int& z, int& w, unsigned addr, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %5, 0;\n @pp ld.shared.v4.s32 {%0, %1, %2, %3}, [%4+%6];\n}\n"
                 : "+r"(x), "+r"(y), "+r"(z), "+r"(w) : "r"(addr), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts64o_if(unsigned addr, int x, int y, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %3, 0;\n @pp st.shared.v2.s32 [%0+%4], {%1, %2};\n}\n" ::"r"(addr), "r"(x), "r"(y), "r"((int)c), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts128o_if(unsigned addr, int x, int y, int z, int w, bool c) {
    asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %5, 0;\n @pp st.shared.v4.s32 [%0+%6], {%1, %2, %3, %4};\n}\n" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w), "r"((int)c), "n"(OFF) : "memory");
}
template <int V> struct IC { static constexpr int value = V; };
template <bool V> struct BC_ { static constexpr bool value = V; };
// one statically unrolled ring period: iteration u uses ring slot u; one CTA barrier per iteration
template <bool IO, int MODE, class F, int... U>
__device__ __forceinline__ void steady_block(F& f, const int q, std::integer_sequence<int, U...>) {
    ((f(IC<MODE>{}, IC<U>{}, q + U, BC_<IO>{}), __syncthreads()), ...);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// One 3-way "gap opening" reduction: out[x] for x = 01, 10, 11 from in[s] for s = 01, 10, 11.
//   x = 11 (match half): best of all three sources (mu is added at the target)
//   x = gap half g     : max(in[g], beta + max(other two))                       (pyx:108-115)
// With beta < 0 the own source may join the inner max (beta + in[g] < in[g] never wins).
template <bool BNEG, bool P16 = false>
__device__ __forceinline__ void open3(int i0, int i1, int i2, int beta, int& o0, int& o1, int& o2) {
    o2 = xmax3<P16>(i0, i1, i2);
    if (BNEG) {
        o0 = xaddmax<P16>(o2, beta, i0);
        o1 = xaddmax<P16>(o2, beta, i1);
    } else {
        o0 = addmax(vmax(i1, i2), beta, i0);
        o1 = addmax(vmax(i0, i2), beta, i1);
    }
}

// Register caps: narrow bands in batch mode fit 128 registers without spills, which (with 54 KB of shared memory per
// 4-warp CTA) allows four CTAs = 16 warps per SM (measured on config 3: +3.6 % over three CTAs at 142 registers);
// wider bands and the long-pair flavour keep 168 (three CTAs of 128 threads: 65536 / 384 = 170).
#ifndef BA_TB_PACKED
#define BA_TB_PACKED 1
#endif
#ifndef BA_LONG_NARROW
#define BA_LONG_NARROW 1
#endif
#ifndef BA_SYS_MAXNREG
#define BA_SYS_MAXNREG 168
#endif
#ifndef BA_SYS_MAXNREG_NARROW
#define BA_SYS_MAXNREG_NARROW 128
#endif
constexpr int LQ = 32;   // long-pair mode: progress flags are published / polled about every LQ iterations (rounded up to whole
                        // ring periods); measured on the 8192 x 8192 pair: 4 -> 158 ms (the extra barrier and the spinning
                        // thread dominate), 16 -> 86 ms, 32 -> 84 ms, 64 -> 84 ms
constexpr int PRE = 4;  // iterations run before position 0: the virtual row above row 0 is 2 iterations ahead,
                        // so its first records must be staged before lane (0,0) reaches its first cell

template <int S, bool PAD>
struct Geo {
    static constexpr int W = 2 * S + 1;                      // band width
    static constexpr int LPR = PAD ? 2 * S + 2 : 2 * S + 1;  // lanes per row
    static constexpr int P = PAD ? 2 * S + 2 : (S == 0 ? 2 : 2 * S + 1);  // cells per column (S = 0 keeps its pad cell: D >= 1)
    static constexpr int R = 32 / LPR;                       // rows per warp
    static constexpr int RING = P + 3;                       // ring depth in iterations (max delay P+2; +1: reads never meet the write)
    // Ring values per lane and iteration, ordered by who reads them:
    //   0..2  Q[11][01,10,11]   3..5  L[11][01,10,11]      row below (x0 = 1): the only ones a boundary record carries
    //   6, 7  Q[01][10], Q[01][11]                          next lane of the same row (x0 = 0, x2 = 1)
    //   8     Q[01][01]         9..11 L[01][01,10,11]      the same lane (x0 = 0, x2 = 0), P resp. P-1 iterations later:
    //                                                       kept in register delay lines instead when P is small (SELFREG)
    static constexpr bool SELFREG = (P <= 5);
    static constexpr int NV = SELFREG ? 8 : 12;
    static constexpr int NVR = 6;                            // ring values in a boundary record
    static constexpr int NX = 6;                             // short-delay values crossing a warp boundary
    static constexpr int LA = RING;                          // cp.async look-ahead (iterations) of the boundary staging
    static constexpr int PB = 2 * RING;                      // prefetch-buffer depth (iterations): two ring periods
    static constexpr int REAL = (NVR + NX) * LPR;            // values per boundary record (one iteration of one row)
    static constexpr int REC = (REAL + 3) & ~3;              // record stride in ints (16-byte multiple)
    static constexpr int RSLOT = NV * 32;
    // short-delay exchange block, per slot: recB [LPR+1][4] = {Q[10][01] of the iteration before, L[10][01,10,11]} then
    // recA [LPR+1][2] = {Q[10][11] of the iteration before, Q[10][10]} (one spare element: the last lane of row 0 reads past its row)
    static constexpr int XSLOT = (6 * (LPR + 1) + 3) & ~3;   // ints per slot (16-byte multiple)
    static constexpr int XA = 4 * (LPR + 1);                 // offset of recA inside a slot
    static constexpr int LQB = ((LQ + RING - 1) / RING) * RING;  // flag period in iterations (whole ring periods)
};

// Shared-memory carve-up (ints unless noted):
//   ring   [(G+1)][RING][NV][32]      ring[0] = staging ring of the virtual warp above warp 0
//   xs     [(G+1)][RING][XSLOT]       short-delay values of the row above each warp (two packed records per lane,
//                                     see Geo); xs[G] = CTA output
//   pb     [PB][REC]                  cp.async landing zone for the incoming boundary stream
//   tb     [9][P][LPR] (+pad)         tie-break constants per (source state, b, lane column)  (TRACE)
//   sim    [(nsym+1)][nsym]           similarity table (<< TB), last row zero
//   resB/clsB  bytes, padded
//
// Iterations come in two forms.  The GENERIC form handles everything (range guards, origin, end cell, pad cells) with
// ring slots computed from the iteration counter.  The STEADY form is used for whole ring periods (RING iterations,
// statically unrolled, slot = position in the block) in which every lane of the warp sits strictly inside its pair
// (S < j <= m - S): there all range tests are true, no end/origin event can occur, every ring / exchange / prefetch
// slot is an immediate offset, the history registers are renamed instead of moved, and lanes outside the pair (rows
// beyond n) simply compute garbage that no valid cell ever reads (sources have smaller coordinates; the band edges are
// poisoned).  Each warp picks the form per ring period on its own; both forms execute one CTA barrier per iteration.
//
// LONG = long pairs spread over several CTAs each (cooperative launch; a gang of NC = cpp CTAs per pair, the whole grid for
// a single pair): CTA b of a gang runs the row blocks ("passes") b, b+NC, b+2NC, ... and the boundary stream of pass p is
// consumed by pass p+1 on another CTA while it is being produced.  Streams live in 2*NC global buffers (pass p -> buffer
// p%NC + NC*((p/NC)&1): by the time it is overwritten, at pass p+2NC, pass p+1 has finished because
// every later pass transitively depends on it).  Progress flags (pass id << 32 | records complete)
// are published every LQB iterations with release semantics and polled with acquire loads; stream
// reads are 16-byte cp.async.cg (L2 only), so no stale L1 line can be seen across SMs.
//
// P16 = two pairs per work item in packed 16-bit halves (score only, pad-free, beta < 0): twice the pairs per
// instruction for batches whose scores provably fit 16 bits (short RNA-like pairs, BASELINE config 4).
//
// NA = the non-affine model (gap_opening_cost == 0, pyx:225-252, 443-471, 513-531) on the same machinery: with beta = 0
// the nine "states" only remember the type of the last column, the best of them is the reference's single value
// M[i,j,k,l].  Three constants differ from the affine score (half-match columns 1100 / 0011 cost Delta, not 2 Delta;
// the double-shift columns 0110 / 1001 do not exist, pyx:250-252), and the tie-break is "first case in the order of
// pyx:233-248": with TRACE the low 4 bits carry 15 - case index, attached at the target through the additive constants;
// the code word of a cell is the case index of its best state (what the traceback of pyx:521-528 would pick).
//
// CHAIN = short pairs (every pair fits one row block): a work item is a CHAIN of up to KCHAIN pairs that run back to back through
// the systolic array along the j axis, separated by one dead column, so the 2-iterations-per-row pipeline skew is paid once per
// chain instead of once per pair.  Every lane switches to the next pair of the chain on its own when it passes the dead column
// (its row of A, its validity, its slice of the staged B molecules, its code stream); steady blocks are agreed by a warp vote.
//
// REBASE = score ranges beyond the packed 32-bit plan (value << TB does not fit although the values themselves do; e.g. scoring
// parameters without a common divisor on pairs of a few hundred residues and more).  Two launches.  The score-only launch
// (TB = 0, exact) also records rowmax[i] = the largest value of any cell-state in row i.  The TRACE launch then runs in
// rebased coordinates V' = V - rowmax[i]: the max-plus recurrence is invariant under a potential, and a potential that depends
// on i only changes nothing but the additive constants of the x0 = 1 cases (by rowmax[i-1] - rowmax[i], a lane constant).
// Every true V' is <= 0, and values further than |NEGP| below their row's maximum are clamped to the floor.  A clamp (or any
// other out-of-range source) can only RAISE a value, and every raised value's argmax chain ends at a floor value -- whose id
// field is 31, which no case has --, at a poisoned case (source outside the band) or outside the matrix: the traceback stops
// there with complete = 0.  A walk that reaches the origin therefore saw exact values only, and (all values being >= the
// true ones) won its ties against a superset of its true competitors: it is the reference's trace.  Pairs whose walk fails,
// whose values rise above 0 or whose final score differs from the first launch's are recomputed by the level kernel (engine.cu).
//
// TILES (I/O-warp flavour of the long-pair mode, one pair per launch, A.ntc > 1): the row blocks are cut into column chunks and
// a free CTA claims the first READY (row block, chunk) tile, lowest chunk first (A.tile_next[c] = next row block of chunk c;
// new row blocks start as early as they can, every other CTA works further to the right), so that a pair with more row
// blocks than the GPU holds CTAs keeps every CTA busy instead of running a second, partly empty round.  A tile is ready when
// the tile to its left is complete and the one above has produced its first records: a claimed tile never waits for an
// unclaimed one, so the scheme cannot deadlock whatever the timing.  Tile (p, c) streams its top boundary
// from tile (p-1, c) exactly like a row block streams from the one above, and starts when tile (p, c-1) is complete.  The
// only state that crosses a chunk boundary are the twelve x1 = 1 values (the ring values) of the chunk's last column: they
// go through a global column buffer indexed by (b, value, row, band offset), written by the cells of column j1-1 and read
// by the cells of column j1 in place of their ring inputs; every other source of a cell lies in its own column (and, where
// it would lie in the previous cell of that column below b = -S, is poisoned anyway).  Every tile has its own boundary
// stream and its own progress flag (zeroed per launch), so nothing is reused and nothing can be overwritten early.
constexpr int KCHAIN = BA_KCHAIN;
template <int S, bool TRACE, bool PAD, bool BNEG, bool LONG, bool P16 = false, bool NA = false, bool CHAIN = false, bool REBASE = false, bool IOW = false,
          bool TILED = false>
__global__ void __maxnreg__((S <= 2 && (!LONG || BA_LONG_NARROW)) ? BA_SYS_MAXNREG_NARROW : BA_SYS_MAXNREG) fill_systolic_kernel(SysArgs A) {
    static_assert(!P16 || (!TRACE && !PAD && BNEG && !LONG), "16-bit pair mode: score only, pad-free, beta < 0, batch mode");
    static_assert(!NA || (BNEG && !LONG && !P16), "non-affine flavour: batch mode, 32-bit");
    static_assert(!CHAIN || (!PAD && BNEG && !LONG && !P16 && !NA), "chained short pairs: plain pad-free affine flavour");
    static_assert(!REBASE || (!PAD && BNEG && !P16 && !NA && !CHAIN), "rebased wide-range flavour: plain pad-free affine flavour");
    static_assert(!IOW || LONG, "the I/O warp exists in the long-pair flavour only");
    static_assert(!TILED || (IOW && !PAD && !REBASE), "column-chunked tiles: pad-free I/O-warp flavour");
    constexpr bool TILES = TILED;                     // column-chunked tiles (a flavour of its own: launched when A.ntc > 1)
    constexpr bool REB = REBASE && TRACE;             // values are relative to the row maxima of the score-only launch
    constexpr bool TBPACK = BA_TB_PACKED && TRACE && !NA && S <= 3;  // tie-break table packed three entries per word (3 TB <= 32 bits)
    using G_ = Geo<S, PAD>;
    constexpr int W = G_::W, P = G_::P, LPR = G_::LPR, R = G_::R, RING = G_::RING, NVR = G_::NVR, PB = G_::PB;
    constexpr bool SELFREG = G_::SELFREG;
    constexpr int RSLOT = G_::RSLOT, REC = G_::REC, REAL = G_::REAL, LA = G_::LA, XSLOT = G_::XSLOT, LQB = G_::LQB, XA = G_::XA;
    constexpr int RSLOTB = RSLOT * 4, XSLOTB = XSLOT * 4, RECB = REC * 4;
    // ring inputs fetched one iteration ahead: pays off only where the long-pair pipeline is latency-bound (wide bands, two CTAs
    // per SM: 8192 x 8192 at max_shift 3 +1.4 %); at max_shift <= 2 with four CTAs per SM it costs 5 % (measured on gangs of 2000-aa pairs)
    constexpr bool PF = LONG && S >= 3;
#ifdef BA_SYS_NO_STEADY
    constexpr bool STEADY_OK = false;
#else
    constexpr bool STEADY_OK = !PAD;                // pad cells need the per-cell validity select
#endif
#ifndef BA_SYS_GUARD
#define BA_SYS_GUARD 1
#endif
// (long-pair flavours: measured with the row-block timeline hook -- the head of a row block gets faster (first 50 iterations of
// the 928 x 933 pair 31 -> 20 us, start-to-start lag 54 -> 35 us), but a consumer that starts closer to its producer only stalls
// at its flag points until the old distance is back; fill times unchanged (2.04 / 53.4 ms), column chunks slower: off)
#ifndef BA_SYS_GUARD_LONG
#define BA_SYS_GUARD_LONG 0
#endif
#ifndef BA_LONG_CREDIT
#define BA_LONG_CREDIT 1
#endif
    // the guarded static form (see the iteration lambda); batch flavours only: in the long-pair flavours it bought nothing
    // (the start-up lag of a row block is not the generic form's cost) and the larger code cost 5-15 %
    // (16-bit pair mode: its per-half masks make the guarded form cost more than it saves, 1688 -> 1624 GCUPS on config 4)
    constexpr bool GUARD_OK = BA_SYS_GUARD && STEADY_OK && !P16 && !CHAIN && (!LONG || BA_SYS_GUARD_LONG);
    extern __shared__ __align__(16) int smem[];
    // IOW (long-pair flavour, launched when A.io_warp is set): one more warp than the G compute warps.  It owns the boundary I/O of the CTA -- the flush
    // of the last row's records, the staging of the incoming stream, the progress flags -- so that no compute warp carries it:
    // in long-pair mode a CTA advances at the pace of its slowest warp (one barrier per iteration), and that was warp 0.
    const int G = LONG ? A.gwarps : (int)(blockDim.x >> 5);  // compute warps (a plain kernel argument: stays in uniform registers)
    const int RT = G * R;  // rows per pass
    int* ring = smem;
    int* xs = ring + (size_t)(G + 1) * RING * RSLOT;
    int* pb = xs + (G + 1) * RING * XSLOT;
    int* tbtab = pb + PB * REC;
    int* ssim = tbtab + P * LPR * 12;
    const int nsym = A.sc.nsym;
    uint8_t* sresB = reinterpret_cast<uint8_t*>(ssim + (nsym + 1) * nsym);
    const int bpad = A.bpad, boff = A.boff;
    uint8_t* sclsB = sresB + bpad;
    uint8_t* sresB_hi = sclsB + bpad;   // P16 only: molecule B of the second pair
    uint8_t* sclsB_hi = sresB_hi + bpad;
    __shared__ int s_pair;
    __shared__ PairDesc s_cd[CHAIN ? KCHAIN : 1];  // CHAIN: the pairs of the current chain
    __shared__ int s_seg[CHAIN ? KCHAIN + 1 : 1];  //        and where position l = 0 of each one's molecule B sits in sresB / sclsB

    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const int r = lane / LPR, c = lane - r * LPR;
    const bool io_warp = IOW && g == G;
    const int ftid = IOW ? G * 32 : 0;  // the thread that publishes / polls progress flags
    const bool lane_real = (r < R) && (c < W) && !io_warp;
    const int a = c - S;
    const int sigma = 2 * (g * R + r) + c;
    const bool row0 = (r == 0), lastrow = (r == R - 1);
    const int TB = TRACE ? A.tb_bits : 0;
    const int NEGP = P16 ? pack2(A.negp) : A.negp;
    const int NEGF = REB ? (NEGP | 31) : NEGP;      // the floor itself: id field 31 = "no case" for the traceback
    const int beta = P16 ? pack2(A.beta_p) : A.beta_p, kGD = P16 ? pack2(A.k_gd) : A.k_gd, k2G = P16 ? pack2(A.k_2g) : A.k_2g;
    const int k2G2D = P16 ? pack2(A.k_2g2d) : A.k_2g2d, k2D = P16 ? pack2(A.k_2d) : A.k_2d;
    // band-edge poisons of the pad-free flavour (lane constants): sources at a+1 (x0=1,x2=0) do not exist for
    // the last column, sources at a-1 (x0=0,x2=1) do not exist for the first one
    const int pU1 = (!PAD && c == W - 1) ? NEGP : 0;
    const int pW = (!PAD && c == 0) ? NEGP : 0;

    // ---- one-time shared-memory initialisation: everything "minus infinity"
    for (int q = tid; q < (int)((G + 1) * RING * RSLOT + (G + 1) * RING * XSLOT + PB * REC); q += blockDim.x) smem[q] = NEGP;
    if (TRACE && !NA) {
        if constexpr (TBPACK) {  // three TB-bit entries per word: word w of (b, column) = states 3w, 3w+1, 3w+2
            for (int q = tid; q < P * LPR * 3; q += blockDim.x) {
                const int w = q / (P * LPR), rem = q - w * (P * LPR);
                const int* src = A.tbtab + (3 * w) * (P * LPR) + rem;
                tbtab[q] = src[0] | (src[P * LPR] << TB) | (src[2 * P * LPR] << (2 * TB));
            }
        } else {
            for (int q = tid; q < P * LPR * 12; q += blockDim.x) tbtab[q] = A.tbtab[q];
        }
    }
    for (int q = tid; q < (nsym + 1) * nsym; q += blockDim.x) ssim[q] = (q < nsym * nsym) ? A.sim_p[q] : 0;
    __syncthreads();

    // per-lane source bases inside the ring array (in ints): U sources sit one row up, W one lane left
    const int own_ring = (g + 1) * RING * RSLOT;
    // The 32 - R*LPR idle lanes behind the last row shadow lane 0's ring loads (same addresses = a broadcast).  With
    // their own lane index they would hit the banks of row 0's loads from the warp above: a 2-way conflict on the six
    // loads that head the iteration's dependency chain (measured: 648 -> 666 GCUPS on config 3).
    const int alane = (r >= R) ? 0 : lane;
    const int up_ring = (row0 || r >= R) ? g * RING * RSLOT + R * LPR : own_ring;  // row 0 reads the last row of the warp above
    const int baseU0 = up_ring + alane - LPR;                          // source lane for x0=1,x2=1 (same column)
    const int baseU1 = up_ring + alane - (LPR - 1);                    // x0=1,x2=0 (column + 1)
    // lane 0 has no left neighbour; lane 31 is a pad or idle lane (PAD) or lane 0's W inputs are poisoned
    // (!PAD), so lane 0 simply reads lane 31 (no special case in the loop)
    const int lsrcW = (lane == 0) ? 31 : lane - 1;
    const int baseW = own_ring + ((alane == 0) ? 31 : alane - 1);      // x0=0,x2=1
    const int baseS = own_ring + alane;                                // self
    // the same as shared-memory byte addresses: slot * RSLOT * 4 is then the only per-iteration address arithmetic
    const unsigned rU0 = smem_u32(ring + baseU0), rU1 = smem_u32(ring + baseU1), rW = smem_u32(ring + baseW), rS = smem_u32(ring + baseS);
    const unsigned wS = smem_u32(ring + own_ring + lane);              // this lane's column of its own ring (writes)
    const int xs_in = g * RING * XSLOT;                                // xs block feeding this warp's row 0
    const int xs_out = (g + 1) * RING * XSLOT;
    // byte addresses of this lane's recB entry (slot 0) in the block it reads as row 0 / writes as the last row
    const unsigned xsi_b = smem_u32(xs + xs_in + 4 * c), xso_b = smem_u32(xs + xs_out + 4 * c);
    const unsigned tb_b = smem_u32(tbtab + c);

    bool long_done = false;
    for (;;) {
        int pi = 0;
        if (LONG) {  // CTAs blockIdx.x / cpp == pi form the gang of pair pi (one pair per gang and launch)
            if (long_done) return;
            long_done = true;
            pi = blockIdx.x / A.cpp;
            if (pi >= A.npairs) return;
        } else {
            if (tid == 0) s_pair = atomicAdd(A.counter, 1);
            __syncthreads();
            pi = s_pair;
            __syncthreads();
            if (pi >= A.npairs) return;
        }
        PairDesc d, dh;  // dh: second pair of a P16 work item (dh.orig < 0: none, the first pair is computed twice)
        int K = 1;       // CHAIN: pairs in this chain
        if (P16) {
            d = A.pairs[2 * pi];
            dh = A.pairs[2 * pi + 1];
        } else if (CHAIN) {
            const int first = A.chains[pi];
            K = A.chains[pi + 1] - first;
            if (tid < K) s_cd[tid] = A.pairs[first + tid];
            __syncthreads();
            if (tid == 0) {  // B segments: front slack, then per pair its columns with a guard of S + 3 on either side
                int at = boff;
                for (int kk = 0; kk < K; ++kk) { s_seg[kk] = at; at += s_cd[kk].m + 2 * S + 6; }
                s_seg[K] = at;  // the "no more pairs" state of a lane reads the tail slack
            }
            d = s_cd[0];
            dh = d;
        } else {
            d = A.pairs[pi];
            dh = d;
        }
        const int n = P16 ? max(d.n, dh.n) : d.n, m = P16 ? max(d.m, dh.m) : d.m;
        const uint8_t* ra = A.res + d.offA;
        const uint8_t* ca = A.cls + d.offA;
        const uint8_t* ra_hi = A.res + dh.offA;
        const uint8_t* ca_hi = A.cls + dh.offA;
        // stage molecule B (bytes); 1-based position l -> sclsB[l + boff]; 255 (B) / 254 (A) never match
        if (CHAIN) {
            for (int q = tid; q < bpad; q += blockDim.x) { sresB[q] = 0; sclsB[q] = 255; }
            __syncthreads();
            for (int kk = 0; kk < K; ++kk)
                for (int l = 1 + tid; l <= s_cd[kk].m; l += blockDim.x) {
                    sresB[s_seg[kk] + l] = A.res[s_cd[kk].offB + l - 1];
                    sclsB[s_seg[kk] + l] = A.cls[s_cd[kk].offB + l - 1];
                }
        } else
        for (int q = tid; q < bpad; q += blockDim.x) {
            const int l = q - boff;
            sresB[q] = (l >= 1 && l <= d.m) ? A.res[d.offB + l - 1] : 0;
            sclsB[q] = (l >= 1 && l <= d.m) ? A.cls[d.offB + l - 1] : 255;
            if (P16) {
                sresB_hi[q] = (l >= 1 && l <= dh.m) ? A.res[dh.offB + l - 1] : 0;
                sclsB_hi[q] = (l >= 1 && l <= dh.m) ? A.cls[dh.offB + l - 1] : 255;
            }
        }
        __syncthreads();

        const int npass = CHAIN ? 1 : (n + RT) / RT;  // ceil((n+1)/RT)
        int nit_full = (m + 1) * P + 2 * (RT - 1) + LPR + RING;
        if (CHAIN) {  // all pairs of the chain back to back, one dead column after each
            nit_full = 2 * (RT - 1) + LPR + RING;
            for (int kk = 0; kk < K; ++kk) nit_full += (s_cd[kk].m + 2) * P;
        }
        const bool tiles = TILES && A.ntc > 1;
        const int ntc = tiles ? A.ntc : 1;            // column chunks per row block
        const int rowsz = tiles ? A.col_rowsz : 0;    // column buffer: ints per (b, value) plane
        const size_t bstride = (size_t)A.bnd_iters * REC;  // ints per boundary buffer
        const int NC = LONG ? A.cpp : (int)gridDim.x;  // LONG: CTAs that share this pair's row blocks
        const int lb = LONG ? (int)blockIdx.x - pi * NC : 0;
        int* bnd_base = A.bnd + (LONG ? (tiles ? (size_t)0 : (size_t)pi * 2 * NC * bstride) : (size_t)blockIdx.x * 2 * bstride);
        unsigned long long* prog_base = LONG ? A.progress + (tiles ? (size_t)0 : (size_t)pi * 2 * NC) : nullptr;

        for (int tix = lb;; tix += LONG ? NC : 1) {
            int tile = tix;
            if constexpr (TILES) {
                if (tiles) {
                    if (tid == ftid) {  // claim the first ready tile, lowest chunk first; -1: every tile has been claimed
                        int got = -2;
                        while (got == -2) {
                            bool any = false;
                            for (int c = 0; c < ntc && got == -2; ++c) {
                                const int p = *reinterpret_cast<volatile int*>(A.tile_next + c);
                                if (p >= npass) continue;
                                any = true;
                                const int cw_ = min(A.chunk_cols, m + 1 - c * A.chunk_cols);
                                const int nit_c = cw_ * P + 2 * (RT - 1) + LPR + RING;
                                if (c > 0 && ld_acquire_u64(prog_base + p * ntc + c - 1) <
                                                 (((unsigned long long)(p + 1) << 32) | (unsigned long long)(A.chunk_cols * P + 2 * (RT - 1) + LPR + RING)))
                                    continue;
                                if (p > 0 && ld_acquire_u64(prog_base + (p - 1) * ntc + c) < (((unsigned long long)p << 32) | (unsigned long long)min(LA + 2 * RT, nit_c)))
                                    continue;
                                if (atomicCAS(A.tile_next + c, p, p + 1) == p) got = p * ntc + c;
                            }
                            if (got == -2) {
                                if (!any) got = -1;
                                else __nanosleep(200);
                            }
                        }
                        s_pair = got;
                    }
                    __syncthreads();
                    tile = s_pair;
                    __syncthreads();
                    if (tile < 0) break;
                } else if (tix >= npass) break;
            } else if (tix >= npass) break;
            const int pass = tiles ? tile / ntc : tile, tc = tiles ? tile - pass * ntc : 0;
            // columns [j0, j1) of this tile (everything without tiles), and its iterations
            const int j0 = tiles ? tc * A.chunk_cols : 0, j1 = tiles ? min(j0 + A.chunk_cols, m + 1) : m + 1;
            const int nit = tiles ? (j1 - j0) * P + 2 * (RT - 1) + LPR + RING : nit_full;
            const bool col_in = tiles && tc > 0, col_out = tiles && tc + 1 < ntc;
            const int i = pass * RT + g * R + r;
            const int k = i + a;
            // per-pair lane state (constant over a pass, except in CHAIN mode where a lane moves from pair to pair)
            int cur_n = d.n, cur_m = d.m, cur_orig = d.orig, kidx = 0, bofs = boff;
            bool lane_ok = lane_real && i <= d.n && k >= 0 && k <= d.n;
            const int Ai = (lane_ok && i >= 1) ? ra[i - 1] : nsym;  // zero row for i = 0
            int Ak = (lane_ok && k >= 1) ? ca[k - 1] : 254;
            const int* simrow = ssim + Ai * nsym;
            const bool lane_ok_hi = P16 && lane_real && i <= dh.n && k >= 0 && k <= dh.n;
            const int Ak_hi = (lane_ok_hi && k >= 1) ? ca_hi[k - 1] : 254;
            const int* simrow_hi = ssim + ((lane_ok_hi && i >= 1) ? ra_hi[i - 1] : nsym) * nsym;
            const int wlo = A.w_p & 0xffff, whi = A.w_p << 16;
            const bool has_in = pass > 0, has_out = pass + 1 < npass;
            // (tiles: one stream and one flag per tile; the producer of tile (p, c) is tile (p-1, c))
            const int buf_in = tiles ? tile - ntc : LONG ? ((pass - 1 + 2 * NC) % NC) + NC * ((((pass - 1 + 2 * NC) / NC) & 1)) : ((pass + 1) & 1);
            const int buf_out = tiles ? tile : LONG ? (pass % NC) + NC * (((pass + 2 * NC) / NC) & 1) : (pass & 1);
            const int* bnd_in = bnd_base + (size_t)buf_in * bstride;
            int* bnd_out = bnd_base + (size_t)buf_out * bstride;
            const unsigned long long* prog_in = LONG ? prog_base + buf_in : nullptr;
            unsigned long long* prog_out = LONG ? prog_base + buf_out : nullptr;
            const unsigned long long tag_in = (unsigned long long)pass << 32;         // producer pass id + 1
            const unsigned long long tag_out = (unsigned long long)(pass + 1) << 32;
            // boundary I/O descriptors: thread e moves record element e (= v*LPR + cs) every iteration
            const bool io_thread = IOW ? io_warp : tid < REAL;
            // Threads beyond the record shadow element 0 for the (harmless) loads; their stores and copies are predicated off.
            // (Letting them duplicate thread 0's work instead, which makes all of this thread-independent, measured 5 % slower:
            // the fourth warp skipping the boundary I/O matters more than the divergence bookkeeping.)
            const int io_e = io_thread ? (IOW ? lane : tid) : 0;
            const int io_v = io_e / LPR, io_cs = io_e - io_v * LPR;
            const bool io_ring = io_v < NVR;
            const int io_stride = io_ring ? RSLOT : XSLOT;
            // record elements beyond the ring values are the exchange block's 4*LPR recB ints, then its 2*LPR recA ints
            const int io_x = io_e - NVR * LPR;
            const int io_col = io_ring ? io_v * 32 + (R - 1) * LPR + io_cs
                                       : (int)((G + 1) * RING * RSLOT) + (io_x < 4 * LPR ? io_x : XA + io_x - 4 * LPR);
            const int fl_src = io_col + (io_ring ? G * RING * RSLOT : G * RING * XSLOT);  // CTA output row
            const int st_dst = io_col;                                                    // ring[0] / xs[0]
            // (the engine guarantees blockDim.x >= REAL, see engine.cu: one record element per thread)
            // Boundary streams hold record `rec` at offset (rec + PRE + 1) * REC, so the unconditional flush of
            // "iteration q-1" in the very first iteration lands in a slack record.  Running pointers, shared-memory
            // addresses in bytes: the fast path is ~7 instructions per direction and iteration.
            const bool do_flush = has_out && io_thread && !IOW, do_stage = !LONG && has_in && io_thread;
            // IOW: the I/O warp moves the whole record, lane l the elements l, l + 32, ...
            // (its address arithmetic is redone every iteration inside the I/O branch: kept in registers across the compute
            // path it would cost every thread of the CTA six registers, and the I/O warp has the time)
            constexpr int NE = (REAL + 31) / 32;
            // LONG staging: thread e4 < REC/4 moves four consecutive record elements (one 16-byte cp.async.cg)
            constexpr int NVEC = REC / 4;
            static_assert(NVEC <= 32, "the staging threads must fit one warp");
            const int stid = IOW ? lane : tid;                                // index among the staging threads
            const bool stg_thread = (IOW ? io_warp : true) && stid < NVEC;
            unsigned lg_dst[4] = {0, 0, 0, 0};
            int lg_ring = 0, lg_real = 0;
            if (LONG && stg_thread) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = 4 * stid + u, v = e / LPR, cs = e - v * LPR;
                    const bool rg = v < NVR;
                    lg_ring |= rg ? (1 << u) : 0;
                    lg_real |= (e < REAL) ? (1 << u) : 0;
                    const int x = e - NVR * LPR;
                    lg_dst[u] = smem_u32(smem + (rg ? v * 32 + (R - 1) * LPR + cs : (int)((G + 1) * RING * RSLOT) + (x < 4 * LPR ? x : XA + x - 4 * LPR)));
                }
            }
            int* fl_g = bnd_out + io_e;                                                  // record q-1 of iteration q = -PRE
            const int* st_g = bnd_in + (size_t)(2 * RT + LA + 1) * REC + io_e;           // record (q + LA + 2RT) of q = -PRE
            const unsigned fl_s = smem_u32(smem + fl_src), st_s = smem_u32(smem + st_dst), pb_s = smem_u32(pb + io_e);
            const int io_stride_b = io_stride * 4;
            const int q_rec_lim = nit - 2 * RT;                                          // records beyond are "minus infinity"
            // iteration at which this lane sits on the origin / on the end cell (INT_MIN: never)
            const int q_origin = (i == 0 && a == 0 && j0 == 0) ? S + sigma : (int)0x80000000;
            const int q_end = (lane_ok && i == d.n && a == 0 && (!tiles || j1 == d.m + 1)) ? (d.m - j0) * P + S + sigma : (int)0x80000000;
            const int q_end_hi = (lane_ok_hi && i == dh.n && a == 0 && dh.orig >= 0) ? dh.m * P + S + sigma : (int)0x80000000;
            // P16 lane constants: additive constants with the lane's band-edge poisons folded in (SIMD adds)
            const int c16_a1 = P16 ? vadd2(k2G2D, pW) : 0, c16_a3 = P16 ? vadd2(k2G2D, pU1) : 0;
            const int c16_a2 = P16 ? vadd2(kGD, pW) : 0, c16_a6 = P16 ? vadd2(kGD, pU1) : 0;
            const int c16_h22 = P16 ? vadd2(k2D, pW) : 0, c16_h12 = P16 ? vadd2(k2D, pU1) : 0;
            // Traceback codes are stored in the order they are computed: row block, warp, iteration, lane -- every warp
            // appends 256 contiguous bytes per iteration to its own stream (two full lines per store instruction; the
            // cell-major layout costs one 32-byte sector per lane).  Every lane stores in every iteration: slots of lanes
            // or iterations outside the pair are simply never read (sys_code_index, kernels.cuh).
            // A slot is 6 bytes in two planes with the same index: a 32-bit word (states 0-5) and a 16-bit half (states 6-8).
            uint32_t* const codes_lo = reinterpret_cast<uint32_t*>(A.codes);
            uint32_t* cw = nullptr;  // this lane's slot of the current iteration (low plane)
            uint16_t* ch = nullptr;  // steady blocks: the same slot in the high plane
            if (TRACE) cw = codes_lo + d.code_off + (((long long)pass * G + g) * (long long)((m + 1) * P + 2 * (RT - 1) + LPR + RING + PRE) + (long long)j0 * P) * 32 + lane;

            // plain affine flavour: a lane at k = 0 poisons its x2 = 1 cases itself (their sources sit at k = -1, lanes that
            // are outside the pair and, in steady blocks, not masked), so the first rows of a pair can run steady blocks too
            constexpr bool K0FIX = !P16 && !NA && !PAD;
            // REBASE: row potential.  rbase = rowmax[i]; the x0 = 1 cases (sources in row i-1) carry rowmax[i-1] - rowmax[i], folded
            // into the two lane constants that every one of them already adds (pU1 for x2 = 0, pK0 for x2 = 1)
            int rbase = 0, dfp = 0, runmax = (int)0x80000000;
            if (REB && lane_ok) {
                const int* rm = A.rowmax + A.row_off[d.orig];
                rbase = rm[i];
                int dfi = i >= 1 ? rbase - rm[i - 1] : 0;
                if (dfi > A.df_max || dfi < -A.df_max) { A.suspect[d.orig] = 1; dfi = 0; }
                dfp = dfi * (1 << TB);
            }
            const int pK0 = ((K0FIX && k == 0) ? NEGP : 0) - dfp;
            const int pU1w = pU1 - dfp;
            const int pWk = K0FIX ? ((c == 0 || k == 0) ? NEGP : 0) : pW;
            const int k2Gk = k2G + pK0, kGDk = kGD + pK0;
            // Steady range of this warp [st_lo, st_hi): every lane of the warp has S < j <= m - max(S,1) throughout
            // (all range tests true, no origin / end cell), the staged records exist, and no lane of the warp
            // is a band offset below row 0 (k < 0, first rows of the first pass: their cells ARE read, as "minus infinity").
            int st_lo = 0, st_hi = 0;
            if (STEADY_OK) {
                const int sig_lo = 2 * (g * R), sig_hi = 2 * (g * R + R - 1) + LPR - 1;
                const int m_eff = P16 ? min(d.m, dh.m) : d.m;
                st_lo = (S + 1) * P + sig_hi;
                st_hi = (m_eff - (S > 0 ? S : 1) + 1) * P + sig_lo;
                // (tiles: iterations count from the tile's first column; the last column of a chunk that hands its state on runs
                // the generic form, and so does, as everywhere, the first S + 1)
                if (tiles) st_hi = (min(m_eff - (S > 0 ? S : 1), col_out ? j1 - 2 : j1 - 1) - j0 + 1) * P + sig_lo;
                if (has_in) st_hi = min(st_hi, q_rec_lim - LA);
                // (REBASE keeps those lanes masked as well: unmasked they would grow without their row's potential)
                if ((!K0FIX || REBASE) && pass == 0 && g * R < S) st_hi = st_lo;
                if (io_warp) {  // immediate-offset I/O wherever the staged records exist
                    st_lo = -PRE;
                    st_hi = has_in ? q_rec_lim - LA : nit;
                }
            }

            // position of this lane one iteration before the first one (q = -PRE)
            int pos = j0 * P - PRE - 1 - sigma;
            int j = pos >= 0 ? pos / P : -((-pos + P - 1) / P);
            int bb = pos - j * P;
            int wslot = (((-PRE - 1) % RING) + RING) % RING;  // slot of the previous iteration: q mod RING
            int pslot = (((-PRE - 1) % PB) + PB) % PB;        // prefetch-buffer slot of the previous iteration: q mod PB
            int mu1 = 0;

            // history registers (outputs of the last 1..3 iterations), all "minus infinity"
            int hQ10[3] = {NEGP, NEGP, NEGP}, hL10[3] = {NEGP, NEGP, NEGP};
            int hR[3][3];
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int y = 0; y < 3; ++y) hR[x][y] = NEGP;
            int h2Q1010 = NEGP, h2Q1001 = NEGP, h2Q1011 = NEGP, h3Q1011 = NEGP;
            int h2R11[3] = {NEGP, NEGP, NEGP};
            int pf[12];  // PF: ring inputs of the next iteration, by ring id
#pragma unroll
            for (int v = 0; v < 12; ++v) pf[v] = NEGP;
            // SELFREG: this lane's own Q[01][01] of the last P iterations and L[01][*] of the last P-1 (oldest first)
            int dQ[P], dL[3][P > 1 ? P - 1 : 1];
#pragma unroll
            for (int k = 0; k < P; ++k) dQ[k] = NEGP;
#pragma unroll
            for (int y = 0; y < 3; ++y)
#pragma unroll
                for (int k = 0; k < P - 1; ++k) dL[y][k] = NEGP;

            if constexpr (TILES) {
                if (tiles) {
                    // the staged row above, its exchange block and the landing zone start from "minus infinity" in every tile
                    for (int qq = tid; qq < RING * RSLOT; qq += blockDim.x) ring[qq] = NEGP;
                    for (int qq = tid; qq < RING * XSLOT; qq += blockDim.x) xs[qq] = NEGP;
                    for (int qq = tid; qq < PB * REC; qq += blockDim.x) pb[qq] = NEGP;
                    if (col_in && tid == ftid) {  // the chunk to the left of this one (same row block) must be complete
                        const unsigned long long want = tag_out | (unsigned long long)(A.chunk_cols * P + 2 * (RT - 1) + LPR + RING);
                        while (ld_acquire_u64(prog_base + tile - 1) < want) __nanosleep(100);
                    }
                    __syncthreads();
                }
            }
            if (LONG && has_in) {  // wait for the first records of the producer pass, then prime with 16-byte copies
                if (tid == ftid) {
                    const unsigned long long want = tag_in | (unsigned long long)min(LA + 2 * RT, nit);
                    while (ld_acquire_u64(prog_in) < want) __nanosleep(100);
                }
                __syncthreads();
                for (int t0 = -PRE; t0 < LA - PRE; ++t0) {
                    const int rec = t0 + 2 * RT;
                    if (stg_thread && rec < nit)
                        cp_async16s(smem_u32(pb + ((t0 + PB) % PB) * REC + 4 * stid), bnd_in + (size_t)(rec + PRE + 1) * REC + 4 * stid);
                    cp_async_commit();
                }
            } else if (has_in) {  // prime the cp.async pipeline: records for iterations -PRE..LA-PRE-1
                for (int t0 = -PRE; t0 < LA - PRE; ++t0) {
                    for (int e = tid; e < REAL; e += blockDim.x) {
                        const int rec = t0 + 2 * RT;
                        if (rec >= 0 && rec < nit) cp_async4(pb + ((t0 + PB) % PB) * REC + e, bnd_in + (size_t)(rec + PRE + 1) * REC + e);
                    }
                    cp_async_commit();
                }
            }
            __syncthreads();

            if (LONG && A.dbg_ts && tid == ftid) {  // debug hook (BA_DEBUG_TS): when a row block (tile) got going and when it ended
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
                A.dbg_ts[8 * (size_t)tile] = t;
            }
            unsigned pb_cur = 0, pb_oth = 0;  // steady blocks: the two halves of the prefetch buffer (this thread's element)
            // ---- long-pair flavour: stage the incoming boundary (virtual row above warp 0) for iteration q and fetch the record of
            // iteration q + LA; run by the staging threads of warp 0, or of the I/O warp (IOW)
            auto stage_long = [&](auto st_, auto u_, const int q) __attribute__((always_inline)) {
                constexpr bool ST = decltype(st_)::value != 0;
                constexpr int u = decltype(u_)::value;
                const int ws = ST ? u : wslot;
                    cp_async_wait<LA - 1>();
                    if (stg_thread) {
                        const int ps_r = ST ? ((pslot + 1 == PB ? 0 : pslot + 1) + u) : pslot;  // prefetch-buffer slot of q
                        const int ps_w = ps_r >= LA ? ps_r - LA : ps_r + LA;                     // ... and of q + LA
                        int4 v4 = make_int4(NEGP, NEGP, NEGP, NEGP);
                        if (ST || q < q_rec_lim) v4 = *reinterpret_cast<const int4*>(pb + ps_r * REC + 4 * stid);
                        const int vals[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                        for (int uu = 0; uu < 4; ++uu)
                            if ((lg_real >> uu) & 1)
                                sts32(lg_dst[uu] + (((lg_ring >> uu) & 1) ? ws * RSLOTB : ws * XSLOTB), vals[uu]);
                        if (ST || q + LA < q_rec_lim)
                            cp_async16s(smem_u32(pb + ps_w * REC + 4 * stid),
                                        bnd_in + (size_t)(q + LA + 2 * RT + PRE + 1) * REC + 4 * stid);
                    }
                    cp_async_commit();
            };
            // ---- one iteration (one band cell per lane).  ST: steady form, u = ring slot (compile time).
            // IO: the instantiation the I/O warp runs (a loop of its own: merging the two roles after every iteration would
            // undo the register renaming of the steady blocks)
            // Forms (st_): 0 generic; 1 steady (no range tests at all); 2 guarded: the static slots and renamed registers of the
            // steady form plus the per-cell validity select -- for whole ring periods that are not strictly inside the pair but
            // hold neither the origin nor the end cell (the head and the tail of every row block: the row block below starts one
            // such head later than the one above, and in generic form a head costs 2.2 x as much).
            auto iteration = [&](auto st_, auto u_, const int q, auto io_) __attribute__((always_inline)) {
                constexpr int MODE = decltype(st_)::value;
                constexpr bool ST = MODE != 0, GUARD = MODE == 2;
                constexpr int u = decltype(u_)::value;
                constexpr bool IO = IOW && decltype(io_)::value;
                // ---- advance position
                ++bb;
                if (bb == P) { bb = 0; ++j; }
                if (!ST) {
                    wslot = (wslot + 1 == RING) ? 0 : wslot + 1;
                    pslot = (pslot + 1 == PB) ? 0 : pslot + 1;
                }
                if constexpr (IO) {
                    {  // the I/O warp: flush the record of iteration q-1, stage the one of iteration q, nothing else
#pragma unroll
                        for (int kk = 0; kk < NE; ++kk) {
                            const int e0 = lane + 32 * kk;
                            const bool on = has_out && e0 < REAL;
                            const int e = e0 < REAL ? e0 : 0, v = e / LPR, cs = e - v * LPR, x = e - NVR * LPR;
                            const bool rg = v < NVR;
                            const int col = rg ? v * 32 + (R - 1) * LPR + cs + G * RING * RSLOT
                                               : (int)((G + 1) * RING * RSLOT) + (x < 4 * LPR ? x : XA + x - 4 * LPR) + G * RING * XSLOT;
                            const int ps = ST ? (u + RING - 1) % RING : ((wslot == 0) ? RING - 1 : wslot - 1);
                            const int val = lds32(smem_u32(smem + col + ps * (rg ? RSLOT : XSLOT)));
                            if constexpr (ST) stg32o_if<u * RECB>(fl_g + 32 * kk, val, on);
                            else stg32_if(fl_g + 32 * kk, val, on);
                        }
                        if (!ST) fl_g += REC;
                        if (has_in) stage_long(st_, u_, q);
                        return;
                    }
                }
                if constexpr (CHAIN && !ST) {
                    if (bb == 0 && j >= cur_m + 2) {  // past the dead column: this lane enters the next pair of the chain
                        ++kidx;
                        j = 0;
                        if (kidx < K) {
                            const PairDesc& dn = s_cd[kidx];
                            cur_n = dn.n; cur_m = dn.m; cur_orig = dn.orig;
                            lane_ok = lane_real && i <= dn.n && k >= 0 && k <= dn.n;
                            simrow = ssim + ((lane_ok && i >= 1) ? A.res[dn.offA + i - 1] : nsym) * nsym;
                            Ak = (lane_ok && k >= 1) ? A.cls[dn.offA + k - 1] : 254;
                            if (TRACE) cw = codes_lo + dn.code_off + ((long long)g * ((dn.m + 1) * P + 2 * (RT - 1) + LPR + RING + PRE) + sigma + PRE) * 32 + lane;
                        } else {  // no more pairs: idle until the array has drained
                            cur_n = -1; cur_m = 0x3fffffff; lane_ok = false; simrow = ssim + nsym * nsym; Ak = 254;
                        }
                        bofs = s_seg[kidx < K ? kidx : K];
                    }
                }
                const int l = j + bb - S;
                // (tiles: the columns of this tile only; j1 - 1 = m without tiles)
                const bool jin = TILES ? ((unsigned)(j - j0) < (unsigned)(j1 - j0)) : ((unsigned)j <= (unsigned)cur_m);
                const bool valid = (ST && !GUARD) || (lane_ok && (bb < W) && jin && ((unsigned)l <= (unsigned)cur_m));
                int vmask = 0, nmask = 0;  // P16: per-half validity
                if (P16 && !ST) {
                    const bool valid_hi = lane_ok_hi && (bb < W) && ((unsigned)j <= (unsigned)dh.m) && ((unsigned)l <= (unsigned)dh.m);
                    vmask = (valid ? 0x0000ffff : 0) | (valid_hi ? (int)0xffff0000 : 0);
                    nmask = NEGP & ~vmask;
                }

                // ---- similarity inputs of the target cell (mu1 changes once per column)
                const int cB = sclsB[l + bofs];
                int mu2;
                if (P16) {
                    const int cBh = sclsB_hi[l + boff];
                    if (bb == 0) mu1 = (simrow[sresB[j + boff]] & 0xffff) | (simrow_hi[sresB_hi[j + boff]] << 16);
                    mu2 = ((cB == Ak) ? wlo : 0) | ((cBh == Ak_hi) ? whi : 0);
                } else {
                    if (bb == 0) mu1 = simrow[sresB[j + bofs]];
                    mu2 = (cB == Ak) ? A.w_p : 0;
                }

                // ---- flush the CTA's last row of iteration q-1 to the outgoing boundary stream
                if constexpr (!IOW) {
                    if constexpr (ST) {
                        constexpr int ps = (u + RING - 1) % RING;
                        stg32o_if<u * RECB>(fl_g, lds32(fl_s + ps * io_stride_b), do_flush);
                    } else {
                        const int ps = (wslot == 0) ? RING - 1 : wslot - 1;
                        stg32_if(fl_g, lds32(fl_s + ps * io_stride_b), do_flush);
                        fl_g += REC;
                    }
                }

                // ---- gather the 27 inputs
                int inF[9], inH2[9], inH1[9];
                // long-delay values from the rings (ids: 0..2 Q[11][01,10,11], 3..5 Q[01][..], 6..8 L[11][..], 9..11 L[01][..]).
                // They are at least P-1 iterations old, so (when P-1 >= 2) they were fetched during the previous
                // iteration: the loads overlap that iteration's tail and the barrier instead of stalling this one.
                // ring id -> (source lane base, delay): 0:(U1,P+1) 1:(U0,P+1) 2:(U0,P+2) 3-5:(U1,P) 6:(W,P) 7:(W,P+1) 8:(S,P) 9-11:(S,P-1)
                int rv[12];
                if (PF) {
#pragma unroll
                    for (int v = 0; v < 12; ++v) rv[v] = pf[v];
                } else if constexpr (ST) {
                    constexpr int oA = ((u + 2 * RING - (P + 2)) % RING) * RSLOTB, oB = ((u + 2 * RING - (P + 1)) % RING) * RSLOTB;
                    constexpr int oC = ((u + 2 * RING - P) % RING) * RSLOTB;
                    [[maybe_unused]] constexpr int oD = ((u + 2 * RING - (P - 1)) % RING) * RSLOTB;
                    rv[2] = lds32o<oA + 2 * 128>(rU0);
                    rv[1] = lds32o<oB + 1 * 128>(rU0);
                    rv[0] = lds32o<oB + 0 * 128>(rU1);
                    rv[7] = lds32o<oB + 7 * 128>(rW);
                    rv[6] = lds32o<oC + 6 * 128>(rW);
                    rv[3] = lds32o<oC + 3 * 128>(rU1); rv[4] = lds32o<oC + 4 * 128>(rU1); rv[5] = lds32o<oC + 5 * 128>(rU1);
                    if constexpr (!SELFREG) {
                        rv[8] = lds32o<oC + 8 * 128>(rS);
                        rv[9] = lds32o<oD + 9 * 128>(rS); rv[10] = lds32o<oD + 10 * 128>(rS); rv[11] = lds32o<oD + 11 * 128>(rS);
                    }
                } else {
                    int rsA, rsB, rsC, rsD;  // ring slots written P+2, P+1, P, P-1 iterations ago
                    rsA = wslot - (P + 2); if (rsA < 0) rsA += RING;
                    rsB = wslot - (P + 1); if (rsB < 0) rsB += RING;
                    rsC = wslot - P;       if (rsC < 0) rsC += RING;
                    rsD = wslot - (P - 1); if (rsD < 0) rsD += RING;
                    rv[2] = lds32o<2 * 128>(rU0 + rsA * RSLOTB);
                    rv[1] = lds32o<1 * 128>(rU0 + rsB * RSLOTB);
                    rv[0] = lds32o<0 * 128>(rU1 + rsB * RSLOTB);
                    rv[7] = lds32o<7 * 128>(rW + rsB * RSLOTB);
                    rv[6] = lds32o<6 * 128>(rW + rsC * RSLOTB);
                    const unsigned aU1C = rU1 + rsC * RSLOTB;
                    rv[3] = lds32o<3 * 128>(aU1C); rv[4] = lds32o<4 * 128>(aU1C); rv[5] = lds32o<5 * 128>(aU1C);
                    if constexpr (!SELFREG) {
                        const unsigned aSD = rS + rsD * RSLOTB;
                        rv[8] = lds32o<8 * 128>(rS + rsC * RSLOTB);
                        rv[9] = lds32o<9 * 128>(aSD); rv[10] = lds32o<10 * 128>(aSD); rv[11] = lds32o<11 * 128>(aSD);
                    }
                }
                if (SELFREG) {  // this lane's own values of P resp. P-1 iterations ago never left its registers
                    rv[8] = dQ[0];
#pragma unroll
                    for (int y = 0; y < 3; ++y) rv[9 + y] = dL[y][0];
                }
                if constexpr (TILES && !ST) {
                    if (col_in && j == j0 && lane_real) {
                        // first column of a chunk: the x1 = 1 sources sit in the last column of the chunk to the left -> column buffer.
                        // Plane (b, value), element (row + 1) * LPR + column; a source at b + 1 beyond the band is poisoned (any sane
                        // value will do: b is clamped), likewise the left neighbour of column 0; row -1 does not exist.
                        const int* cbuf = A.colbuf + (size_t)(tc - 1) * (size_t)(P * 12) * rowsz;
                        const int b1 = min(bb + 1, P - 1);
                        const int eS = (i + 1) * LPR + c, eW = (c == 0) ? eS : eS - 1, eU0 = eS - LPR, eU1 = eU0 + 1;
                        const int* p0 = cbuf + (size_t)(bb * 12) * rowsz;   // same b
                        const int* p1 = cbuf + (size_t)(b1 * 12) * rowsz;   // b + 1
                        const bool up = i >= 1;
                        rv[0] = up ? __ldcg(p0 + 0 * (size_t)rowsz + eU1) : NEGP;
                        rv[1] = up ? __ldcg(p1 + 1 * (size_t)rowsz + eU0) : NEGP;
                        rv[2] = up ? __ldcg(p0 + 2 * (size_t)rowsz + eU0) : NEGP;
#pragma unroll
                        for (int y = 0; y < 3; ++y) rv[3 + y] = up ? __ldcg(p1 + (3 + y) * (size_t)rowsz + eU1) : NEGP;
                        rv[6] = __ldcg(p1 + 6 * (size_t)rowsz + eW);
                        rv[7] = __ldcg(p0 + 7 * (size_t)rowsz + eW);
                        rv[8] = __ldcg(p0 + 8 * (size_t)rowsz + eS);
#pragma unroll
                        for (int y = 0; y < 3; ++y) rv[9 + y] = __ldcg(p1 + (9 + y) * (size_t)rowsz + eS);
                    }
                }
                inF[6] = rv[0]; inF[7] = rv[1]; inF[8] = rv[2];     // x=1101, 1110, 1111   Q[11][*]
                inF[1] = rv[6]; inF[2] = rv[7]; inF[0] = rv[8];     // x=0110, 0111, 0101   Q[01][*]
#pragma unroll
                for (int y = 0; y < 3; ++y) { inH1[6 + y] = rv[3 + y]; inH1[y] = rv[9 + y]; }  // x=1100 L[11][*], x=0100 L[01][*]
                // short-delay values by shuffle
                inF[5] = __shfl_up_sync(0xffffffffu, h3Q1011, LPR);      // x=1011 D=3
                inF[4] = __shfl_up_sync(0xffffffffu, h2Q1010, LPR);      // x=1010 D=2
                inF[3] = __shfl_up_sync(0xffffffffu, h2Q1001, LPR - 1);  // x=1001 D=2
#pragma unroll
                for (int y = 0; y < 3; ++y) {
                    inH1[3 + y] = __shfl_up_sync(0xffffffffu, hL10[y], LPR - 1);  // x=1000 D=1  L[10][y]
                    inH2[3 * y + 2] = __shfl_sync(0xffffffffu, h2R11[y], lsrcW);  // x=0011 D=2  R[y][11]
                    inH2[3 * y + 1] = __shfl_sync(0xffffffffu, hR[y][1], lsrcW);  // x=0010 D=1  R[y][10]
                    inH2[3 * y + 0] = hR[y][0];                                   // x=0001 D=1  R[y][01] (self)
                }
                // row 0: the row above lives in another warp (or in the staged boundary) -> xs, not shuffles.
                // Q[10][11] / Q[10][10] of lane c (x2 = 1), Q[10][01] and L[10][*] of lane c+1 (x2 = 0).  The last
                // lane of row 0 reads one element past its row: in bounds, and either invalid (PAD) or poisoned.
                if constexpr (ST) {
                    constexpr int o1 = ((u + RING - 1) % RING) * XSLOTB, o2 = ((u + RING - 2) % RING) * XSLOTB;
                    lds64o_if<o2 + XA * 4>(inF[5], inF[4], xsi_b - 8 * c, row0);          // recA of lane c, two iterations ago
                    lds128o_if<o1 + 16>(inF[3], inH1[3], inH1[4], inH1[5], xsi_b, row0);  // recB of lane c+1, last iteration
                } else {
                    int s1 = wslot - 1, s2 = wslot - 2;
                    if (s1 < 0) s1 += RING;
                    if (s2 < 0) s2 += RING;
                    lds64o_if<XA * 4>(inF[5], inF[4], xsi_b - 8 * c + s2 * XSLOTB, row0);
                    lds128o_if<16>(inF[3], inH1[3], inH1[4], inH1[5], xsi_b + s1 * XSLOTB, row0);
                    // origin: M[1111][0,0,0,0] = 0 (pyx:485) enters as the F input of state 1111 (mu1 = mu2 = 0 there)
                    // (the origin lane sits at k = 0, whose x2 = 1 cases carry the poison pK0: cancel it for this one input)
                    if (CHAIN ? (i == 0 && a == 0 && j == 0 && bb == S) : (q == q_origin)) inF[8] = -pK0 - rbase * (1 << TB);
                }

                // ---- additive constants per case (affine_score minus its gap-opening part), with the poisons
                // of the pad-free flavour on exactly the cases whose source cell is outside the band:
                //   pU1: x0=1,x2=0 (lane constant)   pW: x0=0,x2=1 (lane constant)
                //   pB0: x1=0,x3=1 at b = -S          pB1: x1=1,x3=0 at b = +S
                const int pB0 = (!PAD && bb == 0) ? NEGP : 0;
                const int pB1 = (!PAD && bb == W - 1) ? NEGP : 0;
                int kF[9], kh2[3], kh1[3];
                constexpr int ADJ2 = TRACE ? -10 : 0, ADJ1 = TRACE ? -19 : 0;
                if (P16) {
                    const int t0 = vadd2(kGD, pB0), t1 = vadd2(kGD, pB1);
                    kF[0] = k2G;                         // x=0101
                    kF[1] = vadd2(c16_a1, pB1);          // x=0110
                    kF[2] = vadd2(mu2, c16_a2);          // x=0111
                    kF[3] = vadd2(c16_a3, pB0);          // x=1001
                    kF[4] = k2G;                         // x=1010
                    kF[5] = vadd2(mu2, t0);              // x=1011
                    kF[6] = vadd2(mu1, c16_a6);          // x=1101
                    kF[7] = vadd2(mu1, t1);              // x=1110
                    kF[8] = vadd2(mu1, mu2);             // x=1111
                    kh2[0] = t0;                         // x=0001
                    kh2[1] = c16_a2;                     // x=0010
                    kh2[2] = vadd2(mu2, vadd2(c16_h22, pB0));  // x=0011
                    kh1[0] = t1;                         // x=0100
                    kh1[1] = c16_a6;                     // x=1000
                    kh1[2] = vadd2(mu1, vadd2(c16_h12, pB1));  // x=1100
                } else if (NA) {
                    // non-affine scores (pyx:233-248) + (15 - case index) in the low four bits when TRACE
                    constexpr int T_ = TRACE ? 1 : 0;
                    kF[0] = k2G + T_ * (15 - 2);                       // x=0101  case 2
                    kF[1] = 0; inF[1] = NEGP;                          // x=0110  does not exist (pyx:250-252)
                    kF[2] = mu2 + kGD + pW + T_ * (15 - 10);           // x=0111  case 10
                    kF[3] = 0; inF[3] = NEGP;                          // x=1001  does not exist
                    kF[4] = k2G + T_ * (15 - 1);                       // x=1010  case 1
                    kF[5] = mu2 + kGD + pB0 + T_ * (15 - 9);           // x=1011  case 9
                    kF[6] = mu1 + kGD + pU1 + T_ * (15 - 12);          // x=1101  case 12
                    kF[7] = mu1 + kGD + pB1 + T_ * (15 - 11);          // x=1110  case 11
                    kF[8] = mu1 + mu2 + T_ * 15;                       // x=1111  case 0
                    kh2[0] = kGD + pB0 + T_ * (15 - 8);                // x=0001  case 8
                    kh2[1] = kGD + pW + T_ * (15 - 7);                 // x=0010  case 7
                    kh2[2] = mu2 + A.k_d + pW + pB0 + T_ * (15 - 4);   // x=0011  case 4: mu2 + Delta
                    kh1[0] = kGD + pB1 + T_ * (15 - 6);                // x=0100  case 6
                    kh1[1] = kGD + pU1 + T_ * (15 - 5);                // x=1000  case 5
                    kh1[2] = mu1 + A.k_d + pU1 + pB1 + T_ * (15 - 3);  // x=1100  case 3: mu1 + Delta
                } else {
                    kF[0] = k2G;                       // x=0101
                    kF[1] = k2G2D + pWk + pB1;         // x=0110
                    kF[2] = mu2 + kGD + pWk;           // x=0111
                    kF[3] = k2G2D + pU1w + pB0;        // x=1001
                    kF[4] = k2Gk;                      // x=1010   (x0 = x2 = 1: source at k - 1)
                    kF[5] = mu2 + kGDk + pB0;          // x=1011
                    kF[6] = mu1 + kGD + pU1w;          // x=1101
                    kF[7] = mu1 + kGDk + pB1;          // x=1110
                    kF[8] = mu1 + mu2 + pK0;           // x=1111
                    // half-column cases also re-base the id field of the winner they carry (TRACE): a full-column
                    // source has field 27 - src (19..27); -10 maps the (t01, h) sources of x=(0,0,t2,t3) to 9..17 and
                    // -19 the (h, t23) sources of x=(t0,t1,0,0) to 0..8, so on equal tie keys the three groups keep the
                    // reference's case order (ids 0-8 < 9-11 < 12-14) and the source stays decodable.
                    kh2[0] = kGD + ADJ2 + pB0;             // x=0001
                    kh2[1] = kGD + ADJ2 + pWk;             // x=0010
                    kh2[2] = mu2 + k2D + ADJ2 + pWk + pB0; // x=0011
                    kh1[0] = kGD + ADJ1 + pB1;             // x=0100
                    kh1[1] = kGD + ADJ1 + pU1w;            // x=1000
                    kh1[2] = mu1 + k2D + ADJ1 + pU1w + pB1;// x=1100
                }
                int M[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int t01 = t / 3, t23 = t % 3;
                    const int v1 = xaddmax<P16>(inH1[t], kh1[t01], NEGF);  // floor: nothing ever drops below "minus infinity"
                    const int v = xaddmax<P16>(inH2[t], kh2[t23], v1);
                    M[t] = xaddmax<P16>(inF[t], kF[t], v);
                    if (!ST || GUARD) M[t] = P16 ? ((M[t] & vmask) | nmask) : (valid ? M[t] : NEGF);
                }
                if (REBASE) runmax = vmax3(vmax3(runmax, M[0], M[1]), vmax3(M[2], M[3], M[4]), vmax3(vmax3(M[5], M[6], M[7]), M[8], runmax));

                // ---- results at the end cell
                if (!ST && (CHAIN ? (lane_ok && i == cur_n && a == 0 && j == cur_m && bb == S) : (q == q_end))) {
                    int Mv[9];  // plain values of the nine states (low half in 16-bit pair mode)
#pragma unroll
                    for (int t = 0; t < 9; ++t) Mv[t] = P16 ? (int)(short)(M[t] & 0xffff) : (M[t] >> TB) + rbase;
                    int best = Mv[0];
#pragma unroll
                    for (int t = 1; t < 9; ++t) best = max(best, Mv[t]);
                    int st = 0, bsh = 99;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int t01 = t / 3, t23 = t % 3;
                        const int sh = (hb0(t01) != hb0(t23)) + (hb1(t01) != hb1(t23));
                        if (Mv[t] == best && sh < bsh) { bsh = sh; st = t; }
                        A.end_values[(size_t)cur_orig * 9 + t] = Mv[t] * A.gscale;
                    }
                    if (REB && A.scores[cur_orig] != (long long)best * A.gscale) A.suspect[cur_orig] = 1;  // the first launch's exact score
                    A.scores[cur_orig] = (long long)best * A.gscale;
                    A.start_state[cur_orig] = (uint8_t)st;
                }

                if (P16 && !ST && q == q_end_hi) {  // the second pair of the work item ends at its own cell
                    int best = M[0] >> 16;
#pragma unroll
                    for (int t = 1; t < 9; ++t) best = max(best, M[t] >> 16);
#pragma unroll
                    for (int t = 0; t < 9; ++t) A.end_values[(size_t)dh.orig * 9 + t] = (M[t] >> 16) * A.gscale;
                    A.scores[dh.orig] = (long long)best * A.gscale;
                    A.start_state[dh.orig] = 0;
                }

                // ---- traceback code word + re-arm the tie-break bits for the role as a source
                if (TRACE && NA) {
                    // the cell's single traceback code: case index of the best state (value desc, case order asc)
                    const int bestp = vmax3(vmax3(M[0], M[1], M[2]), vmax3(M[3], M[4], M[5]), vmax3(M[6], M[7], M[8]));
                    if constexpr (ST) stg32o<u * 128>(cw, 15u - (unsigned)(bestp & 15));
                    else { stg32o<0>(cw, 15u - (unsigned)(bestp & 15)); cw += 32; }
                    const int msk = ~((1 << TB) - 1);
#pragma unroll
                    for (int t = 0; t < 9; ++t) M[t] &= msk;  // as a source a state carries no tie information
                } else if (TRACE) {
                    // nine 5-bit id fields: each funnel shift pushes one field in at the top of its word, so state t < 6 ends
                    // at bit 2 + 5t of the low word and state t >= 6 at bit 17 + 5(t-6) of the high word (traceback.cu)
                    unsigned lo = 0, hi = 0;
#pragma unroll
                    for (int t = 0; t < 6; ++t) lo = __funnelshift_r(lo, (unsigned)M[t], 5);
#pragma unroll
                    for (int t = 6; t < 9; ++t) hi = __funnelshift_r(hi, (unsigned)M[t], 5);
                    // (the three fields of the high word sit in its upper half: only that half is stored)
                    if constexpr (ST) {
                        if (!(GUARD && TILES) || !tiles || jin) {  // (tiles: see the generic form)
                            stg32o<u * 128>(cw, lo);
                            stg16o<u * 64>(ch, hi >> 16);
                        }
                    } else {
                        // (tiles: the slot of a lane and iteration belongs to the tile in whose columns the lane is at that time)
                        if ((!CHAIN || kidx < K) && (!TILES || !tiles || jin)) {
                            stg32o<0>(cw, lo);
                            stg16o<0>(A.codes_hi + (cw - codes_lo), hi >> 16);
                        }
                        cw += 32;
                    }
                    // table layout [source state][b][lane column]: for one source state the lanes of a warp read
                    // (at most P*LPR <= 32) consecutive words -> no bank conflicts
                    const int msk = ~((1 << TB) - 1);
                    {
                        const unsigned tp = tb_b + bb * (LPR * 4);
                        if constexpr (TBPACK) {  // three loads instead of nine (the L1 data pipe is the busiest one), six shifts more
#pragma unroll
                            for (int w = 0; w < 3; ++w) {
                                const unsigned tw = (unsigned)lds32(tp + w * (P * LPR * 4));
                                M[3 * w + 0] = (M[3 * w + 0] & msk) | (int)(tw & ~msk);
                                M[3 * w + 1] = (M[3 * w + 1] & msk) | (int)((tw >> TB) & ~msk);
                                M[3 * w + 2] = (M[3 * w + 2] & msk) | (int)((tw >> (2 * TB)) & ~msk);
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < 9; ++t) M[t] = (M[t] & msk) | lds32(tp + t * (P * LPR * 4));
                        }
                    }
                }

                // ---- publish: R (second alignment decided), L (first decided), Q (both)
                int Rv[3][3], Lv[3][3], Qv[3][3];
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    open3<BNEG, P16>(M[3 * x + 0], M[3 * x + 1], M[3 * x + 2], beta, Rv[x][0], Rv[x][1], Rv[x][2]);  // over s23, s01 = x
                    open3<BNEG, P16>(M[0 + x], M[3 + x], M[6 + x], beta, Lv[0][x], Lv[1][x], Lv[2][x]);              // over s01, s23 = x
                }
#pragma unroll
                for (int y = 0; y < 3; ++y) open3<BNEG, P16>(Rv[0][y], Rv[1][y], Rv[2][y], beta, Qv[0][y], Qv[1][y], Qv[2][y]);

                // prefetch the ring inputs of the next iteration (slots written >= 2 iterations before it)
                if (PF) {
                    if constexpr (ST) {
                        constexpr int oA = ((u + 1 + 2 * RING - (P + 2)) % RING) * RSLOTB, oB = ((u + 1 + 2 * RING - (P + 1)) % RING) * RSLOTB;
                        constexpr int oC = ((u + 1 + 2 * RING - P) % RING) * RSLOTB;
                        [[maybe_unused]] constexpr int oD = ((u + 1 + 2 * RING - (P - 1)) % RING) * RSLOTB;
                        pf[2] = lds32o<oA + 2 * 128>(rU0);
                        pf[1] = lds32o<oB + 1 * 128>(rU0);
                        pf[0] = lds32o<oB + 0 * 128>(rU1);
                        pf[7] = lds32o<oB + 7 * 128>(rW);
                        pf[6] = lds32o<oC + 6 * 128>(rW);
                        pf[3] = lds32o<oC + 3 * 128>(rU1); pf[4] = lds32o<oC + 4 * 128>(rU1); pf[5] = lds32o<oC + 5 * 128>(rU1);
                        if constexpr (!SELFREG) {
                            pf[8] = lds32o<oC + 8 * 128>(rS);
                            pf[9] = lds32o<oD + 9 * 128>(rS); pf[10] = lds32o<oD + 10 * 128>(rS); pf[11] = lds32o<oD + 11 * 128>(rS);
                        }
                    } else {
                        const int ns = (wslot + 1 == RING) ? 0 : wslot + 1;
                        int rsA, rsB, rsC, rsD;
                        rsA = ns - (P + 2); if (rsA < 0) rsA += RING;
                        rsB = ns - (P + 1); if (rsB < 0) rsB += RING;
                        rsC = ns - P;       if (rsC < 0) rsC += RING;
                        rsD = ns - (P - 1); if (rsD < 0) rsD += RING;
                        pf[2] = lds32o<2 * 128>(rU0 + rsA * RSLOTB);
                        pf[1] = lds32o<1 * 128>(rU0 + rsB * RSLOTB);
                        pf[0] = lds32o<0 * 128>(rU1 + rsB * RSLOTB);
                        pf[7] = lds32o<7 * 128>(rW + rsB * RSLOTB);
                        pf[6] = lds32o<6 * 128>(rW + rsC * RSLOTB);
                        const unsigned aU1C = rU1 + rsC * RSLOTB;
                        pf[3] = lds32o<3 * 128>(aU1C); pf[4] = lds32o<4 * 128>(aU1C); pf[5] = lds32o<5 * 128>(aU1C);
                        if constexpr (!SELFREG) {
                            const unsigned aSD = rS + rsD * RSLOTB;
                            pf[8] = lds32o<8 * 128>(rS + rsC * RSLOTB);
                            pf[9] = lds32o<9 * 128>(aSD); pf[10] = lds32o<10 * 128>(aSD); pf[11] = lds32o<11 * 128>(aSD);
                        }
                    }
                }

                // ring: long-delay values; short-delay values for the warp below / the next pass (only the last row's matter)
                if constexpr (ST) {
                    sts32o<u * RSLOTB + 0 * 128>(wS, Qv[2][0]); sts32o<u * RSLOTB + 1 * 128>(wS, Qv[2][1]); sts32o<u * RSLOTB + 2 * 128>(wS, Qv[2][2]);
                    sts32o<u * RSLOTB + 3 * 128>(wS, Lv[2][0]); sts32o<u * RSLOTB + 4 * 128>(wS, Lv[2][1]); sts32o<u * RSLOTB + 5 * 128>(wS, Lv[2][2]);
                    sts32o<u * RSLOTB + 6 * 128>(wS, Qv[0][1]); sts32o<u * RSLOTB + 7 * 128>(wS, Qv[0][2]);
                    if constexpr (!SELFREG) {
                        sts32o<u * RSLOTB + 8 * 128>(wS, Qv[0][0]);
                        sts32o<u * RSLOTB + 9 * 128>(wS, Lv[0][0]); sts32o<u * RSLOTB + 10 * 128>(wS, Lv[0][1]); sts32o<u * RSLOTB + 11 * 128>(wS, Lv[0][2]);
                    }
                    sts128o_if<u * XSLOTB>(xso_b, hQ10[0], Lv[1][0], Lv[1][1], Lv[1][2], lastrow);
                    sts64o_if<u * XSLOTB + XA * 4>(xso_b - 8 * c, hQ10[2], Qv[1][1], lastrow);
                } else {
                    int* wr = ring + own_ring + wslot * RSLOT + lane;
#pragma unroll
                    for (int y = 0; y < 3; ++y) {
                        wr[(0 + y) * 32] = Qv[2][y];
                        wr[(3 + y) * 32] = Lv[2][y];
                        if (!SELFREG) wr[(9 + y) * 32] = Lv[0][y];
                    }
                    wr[6 * 32] = Qv[0][1];
                    wr[7 * 32] = Qv[0][2];
                    if (!SELFREG) wr[8 * 32] = Qv[0][0];
                    if constexpr (TILES) {
                        if (col_out && j == j1 - 1 && lane_real) {  // last column of a chunk: hand the twelve ring values on
                            int* cb = A.colbuf + (size_t)tc * (size_t)(P * 12) * rowsz + (size_t)(bb * 12) * rowsz + (i + 1) * LPR + c;
#pragma unroll
                            for (int y = 0; y < 3; ++y) {
                                cb[(0 + y) * (size_t)rowsz] = Qv[2][y];
                                cb[(3 + y) * (size_t)rowsz] = Lv[2][y];
                                cb[(9 + y) * (size_t)rowsz] = Lv[0][y];
                            }
                            cb[6 * (size_t)rowsz] = Qv[0][1];
                            cb[7 * (size_t)rowsz] = Qv[0][2];
                            cb[8 * (size_t)rowsz] = Qv[0][0];
                        }
                    }
                    sts128o_if<0>(xso_b + wslot * XSLOTB, hQ10[0], Lv[1][0], Lv[1][1], Lv[1][2], lastrow);
                    sts64o_if<XA * 4>(xso_b - 8 * c + wslot * XSLOTB, hQ10[2], Qv[1][1], lastrow);
                }
                // history shift
                if (SELFREG) {
#pragma unroll
                    for (int k = 0; k + 1 < P; ++k) dQ[k] = dQ[k + 1];
                    dQ[P - 1] = Qv[0][0];
#pragma unroll
                    for (int y = 0; y < 3; ++y) {
#pragma unroll
                        for (int k = 0; k + 2 < P; ++k) dL[y][k] = dL[y][k + 1];
                        dL[y][P - 2] = Lv[0][y];
                    }
                }
                h3Q1011 = h2Q1011;
                h2Q1011 = hQ10[2]; h2Q1010 = hQ10[1]; h2Q1001 = hQ10[0];
#pragma unroll
                for (int y = 0; y < 3; ++y) {
                    h2R11[y] = hR[y][2];
                    hQ10[y] = Qv[1][y];
                    hL10[y] = Lv[1][y];
#pragma unroll
                    for (int z = 0; z < 3; ++z) hR[y][z] = Rv[y][z];
                }

                // ---- stage the incoming boundary: virtual row above warp 0, iteration q
                const int ws = ST ? u : wslot;
                if (LONG) {
                    if constexpr (!IOW) {
                        if (has_in) stage_long(st_, u_, q);
                    }
                } else if (has_in) {
                    cp_async_wait<LA - 1>();
                    if constexpr (ST) {
                        // pb_cur / pb_oth: this thread's element in the half of the prefetch buffer that holds iterations
                        // q0..q0+RING-1 / that receives q0+LA.. (set per block below)
                        const int val = lds32o<u * RECB>(pb_cur);
                        sts32_if(st_s + u * io_stride_b, val, do_stage);
                        cp_async4so_if<u * RECB, u * RECB>(pb_oth, st_g, do_stage);
                    } else {
                        int val = lds32(pb_s + pslot * RECB);
                        val = (q < q_rec_lim) ? val : NEGP;
                        sts32_if(st_s + wslot * io_stride_b, val, do_stage);
                        const int pw = pslot >= LA ? pslot - LA : pslot + LA;
                        cp_async4s_if(pb_s + pw * RECB, st_g, do_stage && q + LA < q_rec_lim);
                        st_g += REC;
                    }
                    cp_async_commit();
                }
            };

            int next_flag = 0;  // LONG: next iteration (a multiple of RING) at which progress is published / awaited
            // (IOW: the flags belong to the I/O warp alone -- it wrote the records it publishes and it is the only reader of the
            // incoming stream --, so there is no CTA barrier at a flag point: a late I/O warp simply holds the CTA at the iteration
            // barrier.  Publish period: three ring periods; measured on the 8192 x 8192 pair with the credit-style input side
            // below, fill time: 1 period 50.8 ms, 2: 49.4, 3: 49.5, 4: 50.3, 6: 50.6 (when the input side still waited for a whole
            // period of look-ahead at every flag point: 1 period 76.4 ms, 2: 53.4, 4: 51.7))
            const int lqb = (LONG && A.lq_iters > 0) ? A.lq_iters : (IOW ? 3 * RING : LQB);
            auto pass_loop = [&](auto io_) __attribute__((always_inline)) {
            constexpr bool IO = decltype(io_)::value;
            int avail = min(LA + 2 * RT, nit);  // IOW: records of the producer known to be complete (the first wait saw these)
            int dbg_k = 0;  // debug hook: compute warp 0 stamps when it reaches iteration 0, 50, 100, 200, 400, 800
            for (int q = -PRE; q < nit;) {
                if (LONG && !IO && A.dbg_ts && tid == 0 && dbg_k < 6 && q >= (dbg_k == 0 ? 0 : (25 << dbg_k))) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
                    A.dbg_ts[8 * (size_t)tile + 2 + dbg_k] = t;
                    ++dbg_k;
                }
                const bool aligned = (wslot == RING - 1);  // q is a multiple of RING
                if constexpr (IOW) {
#if BA_LONG_CREDIT
                    // Publish every lqb iterations; on the input side keep the last progress value read and poll only when the
                    // records this ring period fetches are not covered by it: a consumer that starts right behind its producer
                    // waits for exactly what it needs instead of for a whole flag period of look-ahead.
                    if (IO && aligned) {
                        if (q >= next_flag) {
                            __syncwarp();
                            if (lane == 0 && has_out && q > 0)  // records 0..q-2: flushed by this warp in the iterations before this one
                                st_release_u64(prog_out, tag_out | (unsigned long long)(q - 1));  // (release: ordered after them)
                            next_flag = q + lqb;
                        }
                        if (has_in) {
                            const int need = min(q + RING + LA + 2 * RT, nit);
                            if (avail < need) {  // warp-uniform
                                unsigned long long v = 0;
                                if (lane == 0) {
                                    const unsigned long long want = tag_in | (unsigned long long)need;
                                    while ((v = ld_acquire_u64(prog_in)) < want) __nanosleep(40);
                                }
                                v = __shfl_sync(0xffffffffu, v, 0);
                                avail = (v >> 32) > (tag_in >> 32) ? nit : (int)(unsigned)v;  // a later tag: the producer is done
                            }
                        }
                    }
#else
                    if (IO && aligned && q >= next_flag) {
                        // (flags every ring period at the head of a row block, so that the row block below starts sooner, measured
                        // slower: 928 x 933 pair 1.83 -> 2.07 ms -- the release / acquire round trips then pace the I/O warp)
                        const int per = lqb;
                        __syncwarp();
                        if (lane == 0) {
                            if (has_out && q > 0) {  // records 0..q-2: flushed by this warp in the iterations before this one
                                st_release_u64(prog_out, tag_out | (unsigned long long)(q - 1));  // (release: ordered after them)
                            }
                            if (has_in) {
                                const unsigned long long want = tag_in | (unsigned long long)min(q + per + LA + 2 * RT, nit);
                                while (ld_acquire_u64(prog_in) < want) __nanosleep(40);
                            }
                        }
                        __syncwarp();
                        next_flag = q + per;
                    }
#endif
                } else if (LONG && aligned && q >= next_flag) {
                    if (tid == ftid) {
                        if (has_out && q > 0) {  // records 0..q-2 were stored before the last barrier
                            __threadfence();
                            st_release_u64(prog_out, tag_out | (unsigned long long)(q - 1));
                        }
                        if (has_in) {  // the iterations up to the next flag point prefetch records up to q + LQB - 1 + LA + 2RT
                            const unsigned long long want = tag_in | (unsigned long long)min(q + lqb + LA + 2 * RT, nit);
                            while (ld_acquire_u64(prog_in) < want) __nanosleep(100);
                        }
                    }
                    next_flag = q + lqb;
                    __syncthreads();
                }
                bool steady = STEADY_OK && aligned && q >= st_lo && q + RING <= st_hi;  // warp-uniform
                if (CHAIN && STEADY_OK && aligned) {
                    // every lane of the warp stays strictly inside its current pair for the whole block (lanes are in
                    // different pairs around a chain boundary, so the range cannot be precomputed per warp)
                    const int posn = j * P + bb;  // position of the previous iteration
                    const bool mine = kidx < K && posn + 1 >= (S + 1) * P && posn + RING <= (cur_m - (S > 0 ? S : 1) + 1) * P - 1;
                    steady = __all_sync(0xffffffffu, mine);
                }
                bool guard = false;  // warp-uniform
                if constexpr (GUARD_OK && !IO) {
                    if (!steady && aligned && q + RING <= (has_in ? q_rec_lim - LA : nit)) {
                        const bool hit = (q_origin >= q && q_origin < q + RING) || (q_end >= q && q_end < q + RING) ||
                                         (P16 && q_end_hi >= q && q_end_hi < q + RING);
                        guard = !__any_sync(0xffffffffu, hit);
                        if constexpr (TILES) {  // the columns that talk to the column buffer run the generic form
                            const int sg_lo = 2 * (g * R), sg_hi = 2 * (g * R + R - 1) + LPR - 1;
                            if (col_in && q < sg_hi + P + 1) guard = false;
                            if (col_out && q + RING > (j1 - 1 - j0) * P + sg_lo) guard = false;
                        }
                    }
                }
                if (steady || guard) {
                    const int ph = (pslot + 1 == PB) ? 0 : pslot + 1;           // 0 or RING
                    pb_cur = pb_s + ph * RECB;
                    pb_oth = pb_s + (ph ? 0 : LA) * RECB;
                    if (TRACE && !NA) ch = A.codes_hi + (cw - codes_lo);
                    if (steady) steady_block<IO, 1>(iteration, q, std::make_integer_sequence<int, RING>{});
                    else steady_block<IO, GUARD_OK ? 2 : 1>(iteration, q, std::make_integer_sequence<int, RING>{});
                    q += RING;
                    pslot = ph + RING - 1;
                    fl_g += RING * REC;
                    st_g += RING * REC;
                    if (TRACE) cw += RING * 32;
                } else {
                    iteration(IC<0>{}, IC<0>{}, q, io_);
                    __syncthreads();
                    ++q;
                }
            }
            };
            if constexpr (IOW) {
                if (io_warp) pass_loop(BC_<true>{});
                else pass_loop(BC_<false>{});
            } else {
                pass_loop(BC_<false>{});
            }
            if (LONG && A.dbg_ts && tid == ftid) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
                A.dbg_ts[8 * (size_t)tile + 1] = t;
            }
            if (REBASE && lane_ok) {
                if (!TRACE) atomicMax(A.rowmax + A.row_off[d.orig] + i, runmax);       // row maxima for the rebased launch
                else if (runmax >= (1 << TB)) A.suspect[d.orig] = 1;                   // no true value exceeds its row's maximum
            }
            if (has_out) {  // last iteration's record, then make the stream visible to the next pass
                for (int e = tid; e < REAL; e += blockDim.x) {
                    const int v = e / LPR, cs = e - v * LPR;
                    const int val = (v < NVR) ? ring[(G * RING + wslot) * RSLOT + v * 32 + (R - 1) * LPR + cs]
                                              : xs[(G * RING + wslot) * XSLOT + (e - NVR * LPR < 4 * LPR ? e - NVR * LPR : XA + e - NVR * LPR - 4 * LPR)];
                    bnd_out[(size_t)(nit + PRE) * REC + e] = val;
                }
                __threadfence();
                if (LONG) {
                    __syncthreads();
                    if (tid == ftid) {
                        __threadfence();
                        st_release_u64(prog_out, tag_out | (unsigned long long)nit);
                    }
                }
            } else if (TILES && tiles) {  // last row block: the chunk to the right still waits for this tile's column state
                __threadfence();
                __syncthreads();
                if (tid == ftid) {
                    __threadfence();
                    st_release_u64(prog_out, tag_out | (unsigned long long)nit);
                }
            }
            if (has_in) cp_async_wait<0>();
            if (!has_out && has_in) {
                // leaving a multi-pass pair: the staging ring must read "minus infinity" again
                for (int qq = tid; qq < RING * RSLOT; qq += blockDim.x) ring[qq] = NEGP;
                for (int qq = tid; qq < RING * XSLOT; qq += blockDim.x) xs[qq] = NEGP;
            }
            __syncthreads();
        }  // passes
    }
}

// ---- host-side geometry helpers (mirrored by engine.cu through kernels.cuh)
template <int S, bool PAD>
size_t smem_bytes_t(int G, int nsym, int bpad, bool p16 = false) {
    using G_ = Geo<S, PAD>;
    size_t ints = (size_t)(G + 1) * G_::RING * G_::RSLOT + (size_t)(G + 1) * G_::RING * G_::XSLOT + (size_t)G_::PB * G_::REC +
                  (size_t)G_::P * G_::LPR * 12 + (size_t)(nsym + 1) * nsym;
    size_t bytes = ints * 4 + (p16 ? 4 : 2) * (size_t)bpad;  // molecule B residues + classes (of both pairs in 16-bit pair mode)
    return (bytes + 15) & ~(size_t)15;
}

template <int S, bool TRACE, bool PAD, bool BNEG>
cudaError_t launch_t(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    auto kern = fill_systolic_kernel<S, TRACE, PAD, BNEG, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, G * 32, smem, st>>>(A);
    return cudaGetLastError();
}

// 16-bit pair mode: two pairs per work item (score only, pad-free, beta < 0)
template <int S>
cudaError_t launch_p16_t(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    auto kern = fill_systolic_kernel<S, false, false, true, false, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, G * 32, smem, st>>>(A);
    return cudaGetLastError();
}
template <int S>
int occ_p16_t(int G, size_t smem) {
    auto kern = fill_systolic_kernel<S, false, false, true, false, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G * 32, smem);
    return nb;
}

// non-affine flavour (gap_opening_cost == 0)
template <int S, bool TRACE, bool PAD>
cudaError_t launch_na_t(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    auto kern = fill_systolic_kernel<S, TRACE, PAD, true, false, false, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, G * 32, smem, st>>>(A);
    return cudaGetLastError();
}
template <int S, bool TRACE, bool PAD>
int occ_na_t(int G, size_t smem) {
    auto kern = fill_systolic_kernel<S, TRACE, PAD, true, false, false, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G * 32, smem);
    return nb;
}
template <int S>
cudaError_t launch_na_s(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st) {
    if (pad) return trace ? launch_na_t<S, true, true>(A, grid, G, smem, st) : launch_na_t<S, false, true>(A, grid, G, smem, st);
    return trace ? launch_na_t<S, true, false>(A, grid, G, smem, st) : launch_na_t<S, false, false>(A, grid, G, smem, st);
}
template <int S>
int occ_na_s(bool trace, bool pad, int G, size_t smem) {
    if (pad) return trace ? occ_na_t<S, true, true>(G, smem) : occ_na_t<S, false, true>(G, smem);
    return trace ? occ_na_t<S, true, false>(G, smem) : occ_na_t<S, false, false>(G, smem);
}

// CHAIN flavour: short pairs chained along j (pad-free, beta < 0)
template <int S, bool TRACE>
cudaError_t launch_chain_t(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    auto kern = fill_systolic_kernel<S, TRACE, false, true, false, false, false, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, G * 32, smem, st>>>(A);
    return cudaGetLastError();
}
template <int S, bool TRACE>
int occ_chain_t(int G, size_t smem) {
    auto kern = fill_systolic_kernel<S, TRACE, false, true, false, false, false, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G * 32, smem);
    return nb;
}

// REBASE flavour (rebased wide-range runs: a score-only launch that records row maxima, then a TRACE launch), batch and long-pair mode
template <int S, bool TRACE, bool LONG>
cudaError_t launch_rebase_t(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    auto kern = fill_systolic_kernel<S, TRACE, false, true, LONG, false, false, false, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (LONG) {
        SysArgs a = A;
        void* params[] = {&a};
        return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(G * 32), params, smem, st);
    }
    kern<<<grid, G * 32, smem, st>>>(A);
    return cudaGetLastError();
}
template <int S, bool TRACE, bool LONG>
int occ_rebase_t(int G, size_t smem) {
    auto kern = fill_systolic_kernel<S, TRACE, false, true, LONG, false, false, false, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G * 32, smem);
    return nb;
}
template <int S>
cudaError_t launch_rebase_s(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool lng, cudaStream_t st) {
    if (lng) return trace ? launch_rebase_t<S, true, true>(A, grid, G, smem, st) : launch_rebase_t<S, false, true>(A, grid, G, smem, st);
    return trace ? launch_rebase_t<S, true, false>(A, grid, G, smem, st) : launch_rebase_t<S, false, false>(A, grid, G, smem, st);
}
template <int S>
int occ_rebase_s(bool trace, bool lng, int G, size_t smem) {
    if (lng) return trace ? occ_rebase_t<S, true, true>(G, smem) : occ_rebase_t<S, false, true>(G, smem);
    return trace ? occ_rebase_t<S, true, false>(G, smem) : occ_rebase_t<S, false, false>(G, smem);
}

// LONG flavour: cooperative launch (all CTAs must be co-resident: they wait on one another)
// (A.io_warp picks the instantiation with the I/O warp: gwarps + 1 warps per CTA)
template <int S, bool TRACE, bool PAD>
cudaError_t launch_long_t(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    auto kern = A.io_warp ? fill_systolic_kernel<S, TRACE, PAD, true, true, false, false, false, false, true>
                          : fill_systolic_kernel<S, TRACE, PAD, true, true>;
    if constexpr (!PAD) {
        if (A.io_warp && A.ntc > 1) kern = fill_systolic_kernel<S, TRACE, PAD, true, true, false, false, false, false, true, true>;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    SysArgs a = A;
    void* params[] = {&a};
    return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3((G + (A.io_warp ? 1 : 0)) * 32), params, smem, st);
}
template <int S, bool TRACE, bool PAD>
int occ_long_t(int G, size_t smem, bool iow) {  // G: compute warps
    auto kern = iow ? fill_systolic_kernel<S, TRACE, PAD, true, true, false, false, false, false, true>
                    : fill_systolic_kernel<S, TRACE, PAD, true, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, (G + (iow ? 1 : 0)) * 32, smem);
    return nb;
}

template <int S, bool TRACE, bool PAD, bool BNEG>
int occ_t(int G, size_t smem) {
    auto kern = fill_systolic_kernel<S, TRACE, PAD, BNEG, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G * 32, smem);
    return nb;
}

// One translation unit per max_shift instantiates the flavours the engine uses:
// (PAD=false,BNEG=true) fast, (PAD=true,BNEG=true) wide-range, (PAD=true,BNEG=false) positive gap opening.
template <int S>
cudaError_t launch_s(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, bool bneg, cudaStream_t st) {
    if (!pad) return trace ? launch_t<S, true, false, true>(A, grid, G, smem, st) : launch_t<S, false, false, true>(A, grid, G, smem, st);
    if (bneg) return trace ? launch_t<S, true, true, true>(A, grid, G, smem, st) : launch_t<S, false, true, true>(A, grid, G, smem, st);
    return trace ? launch_t<S, true, true, false>(A, grid, G, smem, st) : launch_t<S, false, true, false>(A, grid, G, smem, st);
}
template <int S>
cudaError_t launch_long_s(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st) {
    if (pad) return trace ? launch_long_t<S, true, true>(A, grid, G, smem, st) : launch_long_t<S, false, true>(A, grid, G, smem, st);
    return trace ? launch_long_t<S, true, false>(A, grid, G, smem, st) : launch_long_t<S, false, false>(A, grid, G, smem, st);
}
template <int S>
int occ_long_s(bool trace, bool pad, int G, size_t smem, bool iow) {
    if (pad) return trace ? occ_long_t<S, true, true>(G, smem, iow) : occ_long_t<S, false, true>(G, smem, iow);
    return trace ? occ_long_t<S, true, false>(G, smem, iow) : occ_long_t<S, false, false>(G, smem, iow);
}
template <int S>
int occ_s(bool trace, bool pad, bool bneg, int G, size_t smem) {
    if (!pad) return trace ? occ_t<S, true, false, true>(G, smem) : occ_t<S, false, false, true>(G, smem);
    if (bneg) return trace ? occ_t<S, true, true, true>(G, smem) : occ_t<S, false, true, true>(G, smem);
    return trace ? occ_t<S, true, true, false>(G, smem) : occ_t<S, false, true, false>(G, smem);
}

}  // namespace sys
}  // namespace ba

// Kernel argument blocks and host-side launchers shared by the engine.
#pragma once
#include "common.cuh"

namespace ba {

struct FillArgs {
    const uint8_t* res;       // concatenated residue codes
    const uint8_t* cls;       // concatenated structure classes
    const int* sim;           // nsym x nsym similarity table
    Scoring sc;
    const PairDesc* pairs;    // this wave's pairs
    int npairs;
    int* counter;             // work-queue head (zeroed before launch)
    int* scratch;             // per-CTA scratch (generic kernel: rolling levels; systolic: strip boundaries)
    size_t scratch_stride;    // values (int, or long long in the wide instantiation) per CTA
    uint64_t* codes;          // traceback-code arena (nullptr when score-only)
    long long* scores;        // [n_pairs] in caller order
    uint8_t* start_state;     // [n_pairs] traceback start state
    int* end_values;          // [n_pairs][9]  M[t][n,m,n,m]
    const int* mu2;           // optional per-pair structure-similarity matrices (probabilistic RNA profiles, pyx:414-423):
    const long long* mu2_off; //   pair p (caller order) owns n x m ints at mu2 + mu2_off[p], entry (k-1)*m + (l-1) = mu2(k, l)
};

// Systolic kernel (fill_systolic.cu).  All score constants are divided by the gcd of the scoring
// parameters (gscale) and, with traceback codes, shifted left by tb_bits.
struct SysArgs {
    const uint8_t* res;
    const uint8_t* cls;
    const int* sim_p;         // nsym x nsym, scaled and shifted
    const int* tbtab;         // [P][LPR][12] tie-break constants (TRACE only)
    Scoring sc;               // unscaled, for s / nsym
    int w_p, beta_p;          // structure weight, gap opening (scaled, shifted)
    int k_gd, k_2g, k_2g2d, k_2d;  // gamma+Delta, 2*gamma, 2*gamma+2*Delta, 2*Delta (scaled, shifted)
    int k_d;                  // Delta (scaled, shifted): half-match columns of the non-affine model
    int negp;                 // "minus infinity" in the packed domain
    int tb_bits;              // low bits reserved for the tie-break (0 when score-only)
    int gscale;               // gcd the scores were divided by
    int mmax;                 // longest molecule B of the wave
    int boff, bpad;           // front offset / total size of the staged molecule-B arrays
    const PairDesc* pairs;
    int npairs;               // work items: pairs, or chains of pairs (CHAIN flavour)
    const int* chains;        // CHAIN flavour: chain c = pairs[chains[c] .. chains[c+1])
    int* counter;
    unsigned long long* progress;  // LONG flavour: one progress flag per boundary stream (2 * grid)
    int* bnd;                 // per CTA: two boundary streams of bnd_iters records
    int bnd_iters;
    int lq_iters;             // LONG flavour: iterations between progress-flag exchanges (a multiple of the ring period; 0 = default)
    int cpp;                  // LONG flavour: CTAs per pair (gang size); pair p is run by CTAs [p * cpp, (p + 1) * cpp)
    int gwarps;               // LONG flavour: compute warps per CTA (the launch has gwarps + io_warp warps)
    int ntc, chunk_cols;      // I/O-warp flavour, one pair per launch: column chunks per row block (<= 1: none), columns per chunk
    int* tile_next;           //   [ntc] next unclaimed row block of every chunk (zeroed per launch)
    int* colbuf;              //   column buffer [chunk boundary][b][12 values][col_rowsz]: element (row + 1) * LPR + column
    int col_rowsz;
    unsigned long long* dbg_ts;  // LONG flavour, debug hook: [tile][2] globaltimer at the start / end of every row block (tile), or null
    int io_warp;              // LONG flavour: 1 = the CTA has one warp more than its G compute warps, which owns the boundary I/O
    uint64_t* codes;          // code arena; the systolic kernel uses it as a plane of 32-bit words ...
    uint16_t* codes_hi;       // ... plus this plane of 16-bit halves, same slot index (PairDesc::code_off counts slots)
    long long* scores;
    uint8_t* start_state;
    int* end_values;
    // REBASE flavour (fill_systolic.cuh): row maxima of the score-only launch, in scaled units, pair p (caller order) at row_off[p]
    int* rowmax;
    const long long* row_off;
    uint8_t* suspect;         // [n_pairs] caller order: set when the rebased launch cannot vouch for a pair
    int df_max;               // bound on |rowmax[i] - rowmax[i-1]| the range plan allowed for
};

struct TraceArgs {
    const PairDesc* pairs;
    int npairs;
    int s;
    const uint64_t* codes;
    const uint16_t* codes_hi; // systolic layout (sysG > 0, fmt 1 / 2): the 16-bit plane; codes is then the 32-bit plane
    int fmt;                  // 0: nibble t = case id (generic kernel); 1: 5-bit tie fields (systolic kernel);
                              // 2: non-affine model, low nibble = case index 0..12
                              // 3: non-affine model, dedicated kernel: uint32 per (row, j, b), nibble a+S = case index
    int sysG;                 // > 0: the table has the systolic kernel's layout (sys_code_index) with sysG warps per CTA
    int R, LPR, P, RING;      //      and this geometry
    const uint8_t* start_state;
    uint8_t* trace;           // slots; columns are written backwards from the end of each slot
    int* trace_len;           // [n_pairs] caller order
    uint8_t* complete;        // [n_pairs]
};

void launch_fill_generic(const FillArgs& A, int grid, bool trace, bool wide, cudaStream_t st);    // wide: int64 values
void launch_fill_nonaffine(const FillArgs& A, int grid, bool trace, bool wide, cudaStream_t st);
size_t generic_scratch_ints(int nmax, int s);
void launch_traceback(const TraceArgs& A, cudaStream_t st);

struct SysGeo { int W, LPR, P, R, RING, REC; };
SysGeo sys_geo(int S, bool pad);
size_t sys_smem_bytes(int S, bool pad, int G, int nsym, int mmax, bool p16 = false);
int sys_occupancy_p16(int S, int G, size_t smem);
cudaError_t launch_fill_systolic_p16(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st);
int sys_iters(int S, bool pad, int G, int m);
// Code-table geometry of the systolic kernel: [row block][warp][iteration][lane] slots (6 bytes each, two planes).
long long sys_code_words(int S, bool pad, int G, int n, int m);
__host__ __device__ __forceinline__ long long sys_code_index(int R, int LPR, int P, int S, int G, int nit_all, int i, int j, int a, int b) {
    // nit_all = iterations per row block incl. the PRE warm-up ones; lane (row r of warp g, column c = a+S) computes
    // cell (j, b) at iteration j*P + (b+S) + 2*(g*R + r) + c  (fill_systolic.cuh)
    const int RT = G * R, pass = i / RT, rr = i - pass * RT, g = rr / R, r = rr - g * R, c = a + S;
    const int qq = j * P + (b + S) + 2 * rr + c + 4 /* PRE */;
    return (((long long)pass * G + g) * nit_all + qq) * 32 + r * LPR + c;
}
size_t sys_boundary_ints(int S, bool pad, int G, int mmax);
int sys_occupancy(int S, bool trace, bool pad, bool bneg, int G, size_t smem);
int sys_boff(int S, bool pad, int G);
int sys_bpad(int S, bool pad, int G, int mmax);
cudaError_t launch_fill_systolic(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, bool bneg, cudaStream_t st);
int sys_occupancy_na(int S, bool trace, bool pad, int G, size_t smem);
cudaError_t launch_fill_systolic_na(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st);
// Dedicated non-affine kernel (fill_na.cu): a lane owns a row and all its band offsets; max_shift <= BA_NA_MAX_SHIFT.
constexpr int BA_NA_MAX_SHIFT = 3;
int na_iters(int S, int G, int m);
int na_pre(int S);
int na_boff(int S, int G);
int na_bpad(int S, int G, int mmax);
size_t na_boundary_ints(int S, int G, int mmax);
long long na_code_words(int S, int G, int n, int m);
size_t na_smem_bytes(int S, int G, int nsym, int mmax);
int na_occupancy(int S, bool trace, int G, size_t smem);
cudaError_t launch_fill_na(const SysArgs& A, int grid, int G, size_t smem, bool trace, cudaStream_t st);
// uint32 index of the code word holding cell (i, j, *, b) in that kernel's table (nibble a+S inside it)
__host__ __device__ __forceinline__ long long na_code_index(int S, int G, int nit_all, int i, int j, int b) {
    const int P = 2 * S + 1 < 2 ? 2 : 2 * S + 1, RT = G * 32;
    const int pass = i / RT, rr = i - pass * RT, g = rr >> 5, lane = rr & 31;
    const int qq = j * P + (b + S) + rr + P + 1 /* PRE */;
    return (((long long)pass * G + g) * nit_all + qq) * 32 + lane;
}
int sys_occupancy_long(int S, bool trace, bool pad, int G, size_t smem, bool iow = false);  // G: compute warps
// chained short pairs (pad-free affine flavour); bpad_total = bytes of one staged-B array incl. all slack
constexpr int BA_KCHAIN = 16;
int sys_occupancy_chain(int S, bool trace, int G, size_t smem);
size_t sys_smem_bytes_chain(int S, int G, int nsym, int bpad_total);
cudaError_t launch_fill_systolic_chain(const SysArgs& A, int grid, int G, size_t smem, bool trace, cudaStream_t st);
cudaError_t launch_fill_systolic_long(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st);
// rebased wide-range flavour (pad-free affine): score-only launch records row maxima, TRACE launch consumes them
int sys_occupancy_rebase(int S, bool trace, bool lng, int G, size_t smem);
cudaError_t launch_fill_systolic_rebase(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool lng, cudaStream_t st);

}  // namespace ba

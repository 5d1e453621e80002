// C ABI of the engine (include/bialign_b200.h): device memory, wave scheduling, launches.
//
// Data layout in HBM (all owned here, grow-only):
//   res/cls     concatenated uint8 residue codes / structure classes of the sequence table
//   sim         nsym x nsym int32 similarity table
//   desc        PairDesc per pair, sorted by decreasing cost (LPT order inside the device)
//   codes       traceback-code arena: one uint64 per band cell (level kernel: nibble t = winning case of state t;
//               systolic kernel: 5-bit tie field per state; non-affine: case index in the low nibble);
//               pairs are processed in waves that fit the arena, traceback runs per wave
//   bnd         systolic kernel: boundary streams between row blocks (per CTA, or 2 * grid in long-pair mode)
//   trace       one slot of 2(n+m)+2 bytes per pair, columns written backwards by the traceback
//   scores/start_state/end_values/trace_len/complete   per pair, caller order
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bialign_b200.h"
#include "common.cuh"
#include "kernels.cuh"

using namespace ba;

namespace {
thread_local std::string g_create_error;  // ba_engine_create failures have no engine to hold the text

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t ensure(size_t need) {
        if (need <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(need, 1) * sizeof(T));
        if (e == cudaSuccess) cap = need;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t need) {
        if (need <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc((void**)&p, std::max<size_t>(need, 1) * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) cap = need;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};
}  // namespace

struct ba_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    bool have_scoring = false, have_seqs = false, have_pairs = false, ran = false, ran_trace = false;
    Scoring sc{};
    int64_t max_abs_sim = 0;
    DevBuf<int> d_sim;

    std::vector<int64_t> h_off;
    int64_t n_seq = 0;
    int max_residue = 0;
    DevBuf<uint8_t> d_res, d_cls;

    std::vector<int32_t> h_pa, h_pb;
    int64_t n_pairs = 0;
    bool have_mu2 = false;             // per-pair mu2 matrices loaded (ba_set_pair_mu2): general level kernel only
    int64_t max_abs_mu2 = 0;
    DevBuf<int> d_mu2;
    DevBuf<long long> d_mu2_off;

    std::vector<PairDesc> h_desc;          // sorted order
    std::vector<int64_t> h_slot_off;       // caller order
    std::vector<int32_t> h_slot_cap;       // caller order
    std::vector<int64_t> h_last_code_off;  // caller order; -1 if not in the last wave
    DevBuf<PairDesc> d_desc;
    DevBuf<uint64_t> d_codes;
    DevBuf<int> d_scratch;
    DevBuf<int> d_counter;
    DevBuf<int> d_simp, d_tbtab, d_bnd;
    DevBuf<unsigned long long> d_progress;
    int opt_long = -1;                 // multi-CTA long-pair mode: -1 auto, 0 off, 1 force
    int opt_io_warp = -1;              // long-pair mode: a dedicated I/O warp per CTA: -1 auto, 0 off, 1 on
    int opt_col_chunks = 0;            // long-pair mode with the I/O warp: column chunks per row block: 0 / 1 none, 2..64 that many
    DevBuf<int> d_tile_order, d_colbuf;
    DevBuf<unsigned long long> d_dbg_ts;  // BA_DEBUG_TS=<file>: row-block timeline of a single-pair long-mode run
    int opt_p16 = -1;                  // 16-bit pair mode for score-only batches: -1 auto, 0 off, 1 force
    int opt_na = -1;                   // non-affine model: -1 / 1 dedicated kernel when applicable, 0 systolic NA flavour
    int opt_chain = -1;                // short pairs chained along j: -1 auto, 0 off, 1 force
    DevBuf<int> d_chains;
    int opt_rebase = -1;               // rebased wide-range trace runs: -1 auto (when the packed plan fails), 0 off, 1 force
    int64_t opt_rebase_window = 0;     // test hook: cap on the rebased window (scaled score units below a row's maximum), 0 = none
    DevBuf<int> d_rowmax, d_simp1;
    DevBuf<long long> d_row_off;
    DevBuf<uint8_t> d_suspect;
    DevBuf<PairDesc> d_desc2;          // pairs recomputed by the level kernel after a rebased run
    std::vector<int32_t> h_sim;
    int opt_warps = 0;                 // warps per CTA of the systolic kernel (0 = chosen per batch)
    int opt_pad = -1;                  // systolic flavour: -1 auto, 0 pad-free, 1 padded
    int last_fmt = 0, last_sysG = 0;   // code-table layout of the last run (debug fetch)
    int64_t last_hi_off = 0;
    bool last_pad = false, last_na = false;
    DevBuf<long long> d_scores;
    DevBuf<uint8_t> d_start, d_complete, d_trace;
    DevBuf<int> d_endv, d_tlen;
    PinnedBuf<uint8_t> h_stage;
    std::vector<int32_t> h_tlen;
    bool have_tlen = false;

    // Multi-GPU front (ba_engine_create_multi): this object owns no device state; it shards the pair list over one
    // child engine per device (LPT), drives them from one host thread each and merges results in caller order.
    std::vector<ba_engine*> kids;
    std::vector<std::vector<int32_t>> shard;  // per child: caller-order pair indices it owns
    std::vector<int64_t> m_tlen;              // caller order, after ba_trace_bytes / ba_fetch_traces

    int64_t opt_code_arena_bytes = 0;  // 0 = auto
    bool arena_is_budget = false;      // current arena = the full memory budget (not just "all pairs fit")
    int opt_kernel = -1;               // -1 auto
    ba_stats stats{};
};

namespace {

int fail(ba_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg;
    else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(e, _e == cudaErrorMemoryAllocation ? BA_ERR_OOM : BA_ERR_CUDA,               \
                        std::string(#call) + ": " + cudaGetErrorString(_e));                         \
    } while (0)

constexpr size_t kSysSmemLimit = 220 * 1024;  // dynamic shared memory one systolic CTA may ask for

int64_t band_cells(int64_t n, int64_t m, int64_t s) {
    // C(n,m,s) of SURVEY 8: exact count of band-valid cells (k in [max(0,i-s), min(n,i+s)] etc.)
    auto one = [&](int64_t len) {
        int64_t c = 0;
        for (int64_t d = -s; d <= s; ++d) {
            int64_t ad = d < 0 ? -d : d;
            if (len + 1 - ad > 0) c += len + 1 - ad;
        }
        return c;
    };
    return one(n) * one(m);
}

// Host-side plan of the systolic kernel: gcd scaling, tie-break bit budget, "minus infinity".
struct SysPlan {
    bool ok = false;
    bool pad = true, bneg = false;   // kernel flavour (fill_systolic.cuh)
    int g = 1, tb = 0, negp = 0;
    int64_t colabs = 0;              // bound on what one column (plus tie adjustments) adds, scaled units
    int df_max = 0;                  // rebased plan: bound on the row-potential steps it allows for
    std::vector<int> sim_p, tbtab;
};

int64_t gcd64(int64_t a, int64_t b) {
    a = std::llabs((long long)a);
    b = std::llabs((long long)b);
    while (b) { int64_t t = a % b; a = b; b = t; }
    return a;
}

// Exactness conditions of the packed 32-bit domain (DESIGN.md "tie-break packing"): every finite value
// lies in [-Fn, Fp], "minus infinity" values in [negp - Fn, negp + Fp]; the two ranges must not meet
// and nothing may leave the (32 - tb)-bit field.
SysPlan plan_systolic(const ba_engine* e, int nmax, int mmax, bool trace, bool bits16 = false, bool nonaffine = false) {
    SysPlan pl;
    if (bits16 && (trace || nonaffine)) return pl;
    const Scoring& sc = e->sc;
    const int S = sc.s, nsym = sc.nsym;
    if (nsym > 64 || S > BA_SYSTOLIC_MAX_SHIFT) return pl;
    int64_t g = gcd64(gcd64(sc.w, sc.beta), gcd64(sc.gamma, sc.delta));
    int64_t smax = 0, smin = 0;
    for (int32_t v : e->h_sim) { g = gcd64(g, v); smax = std::max<int64_t>(smax, v); smin = std::min<int64_t>(smin, v); }
    if (g == 0) g = 1;
    const int64_t gb = (int64_t)sc.gamma + sc.beta;
    auto pos = [](std::initializer_list<int64_t> xs) { int64_t m = 0; for (auto x : xs) m = std::max(m, x); return m; };
    const int64_t c1p = pos({smax, sc.gamma, gb}), c1n = pos({-smin, -(int64_t)sc.gamma, -gb});
    const int64_t c2p = pos({sc.w, sc.gamma, gb}), c2n = pos({-(int64_t)sc.w, -(int64_t)sc.gamma, -gb});
    const int64_t dp = pos({sc.delta}), dn = pos({-(int64_t)sc.delta});
    // Every finite value lies in [-Fn, Fp]: a path has at most n+m half-columns per alignment and 2(n+m)
    // shift units; only min(n,m) of the half-columns of an alignment can be matches.
    const int64_t len = (int64_t)nmax + mmax + 2, mlen = std::min(nmax, mmax) + 1;
    const int64_t g1p = pos({(int64_t)sc.gamma, gb}), m1p = pos({smax}), m2p = pos({(int64_t)sc.w});
    const int64_t Fp = (mlen * (std::max(m1p, g1p) + std::max(m2p, g1p)) + (len - mlen) * 2 * g1p + len * 2 * dp) / g + 2;
    const int64_t Fn = len * (c1n + c2n + 2 * dn) / g + 2;
    const int64_t colabs = (c1p + c2p + c1n + c2n + 2 * dp + 2 * dn + 2 * std::llabs((long long)sc.beta)) / g + 64;  // one column + tie adjustments
    (void)c1p; (void)c2p;
    pl.colabs = colabs;
    int kb = 0;
    while ((1 << kb) < (S + 2) * (S + 2)) ++kb;
    pl.tb = trace ? (nonaffine ? 4 : kb + 5) : 0;  // non-affine: 15 - case index of pyx:233-248
    const int vb = (bits16 ? 16 : 32) - pl.tb;
    const int64_t lim = (int64_t)1 << (vb - 1);
    pl.bneg = sc.beta < 0 || nonaffine;  // beta == 0: the beta <= 0 shortcut of open() is exact, no opening term at all
    // pad-free flavour: band-edge cases carry a poison of NEGP and values are floored at NEGP, so the most
    // negative intermediate is (floored source) + two poisons + constants  >=  3*negv - Fn
    const int64_t negv_nopad = -(Fn + Fp + colabs);
    const bool nopad_ok = pl.bneg && (-3 * negv_nopad + 2 * colabs < lim);
    const bool pad_ok = (2 * Fn + Fp + 64 < lim);
    if (bits16) {  // 16-bit pair mode exists for the pad-free flavour only, and the table entries must fit too
        if (!nopad_ok || (smax - smin) / g >= lim / 2) return pl;
    }
    if (e->opt_pad == 0 && !nopad_ok) return pl;
    if (e->opt_pad == 1 && !pad_ok) return pl;
    pl.pad = !bits16 && ((e->opt_pad == 1) || (e->opt_pad < 0 && !nopad_ok));
    if (pl.pad && !pad_ok) return pl;
    int64_t negv = negv_nopad;
    if (pl.pad) {
        negv = trace ? -(lim - Fn - 32) : std::max<int64_t>(-(lim - Fn - 32), -((int64_t)1 << 30));
        if (negv + Fp + 16 >= -Fn) return pl;
    }
    pl.g = (int)g;
    pl.negp = (int)(negv * ((int64_t)1 << pl.tb));
    pl.sim_p.resize((size_t)nsym * nsym);
    for (size_t q = 0; q < pl.sim_p.size(); ++q) pl.sim_p[q] = (int)((e->h_sim[q] / g) * ((int64_t)1 << pl.tb));
    if (trace && !nonaffine) {
        // rank of the tie key (k0,k1) = (|T0|+|T1|, |T1|), T = band offset of the cell + (s0-s2, s1-s3): pyx:541-545
        const SysGeo geo = sys_geo(S, pl.pad);
        const int LPR = geo.LPR, P = geo.P, NK = (S + 2) * (S + 2);
        std::vector<std::pair<int, int>> keys;
        for (int k1 = 0; k1 <= S + 1; ++k1)
            for (int d0 = 0; d0 <= S + 1; ++d0) keys.push_back({d0 + k1, k1});
        std::sort(keys.begin(), keys.end());
        pl.tbtab.assign((size_t)P * LPR * 12, 0);
        for (int bb = 0; bb < 2 * S + 1; ++bb)
            for (int c = 0; c < 2 * S + 1; ++c)
                for (int src = 0; src < 9; ++src) {
                    const int s01 = src / 3, s23 = src % 3;
                    const int T0 = (c - S) + hb0(s01) - hb0(s23), T1 = (bb - S) + hb1(s01) - hb1(s23);
                    const int k1 = std::abs(T1), k0 = std::abs(T0) + k1;
                    const int rank = (int)(std::lower_bound(keys.begin(), keys.end(), std::make_pair(k0, k1)) - keys.begin());
                    pl.tbtab[(size_t)src * P * LPR + (size_t)bb * LPR + c] = ((NK - 1 - rank) << 5) | (27 - src);
                }
    }
    pl.ok = true;
    return pl;
}

// Plans of a rebased wide-range trace run (fill_systolic.cuh, REBASE): p1 = the ordinary score-only plan (exact values, no tie
// bits), p2 = the TRACE launch in coordinates relative to the row maxima of the first: its values only need the window
// [negv, 0] plus the usual slack, whatever the lengths.  False when either does not exist or the window would be too small
// to be useful (fewer than 64 worst-case columns below a row's maximum).
bool plan_rebase(const ba_engine* e, int nmax, int mmax, SysPlan& p1, SysPlan& p2) {
    if (e->sc.beta >= 0 || e->opt_pad == 1) return false;
    p1 = plan_systolic(e, nmax, mmax, false);
    if (!p1.ok || p1.pad || !p1.bneg) return false;
    // tie-break tables, scaled similarity table and bit budget: those of a trace plan for a pair short enough to have one
    p2 = plan_systolic(e, 1, 1, true);
    if (!p2.ok || p2.pad || !p2.bneg || p2.g != p1.g) return false;
    const int64_t lim = (int64_t)1 << (31 - p2.tb);
    const int64_t dfm = (2 * e->sc.s + 4) * p1.colabs;
    int64_t negv = -((lim - 2 * p1.colabs - dfm - 64) / 3);
    if (-negv < 64 * p1.colabs) return false;
    if (e->opt_rebase_window > 0) negv = std::max<int64_t>(negv, -e->opt_rebase_window);
    p2.negp = (int)(negv * ((int64_t)1 << p2.tb));
    p2.colabs = p1.colabs;
    p2.df_max = (int)std::min<int64_t>(dfm, 1 << 28);
    return true;
}

}  // namespace

// Multi-GPU front (defined at the end of this file): every entry point forwards here when the engine has children.
namespace multi {
void destroy(ba_engine* e);
int set_option(ba_engine* e, const char* key, int64_t value);
int set_scoring(ba_engine* e, const int32_t* sim, int nsym, int w, int beta, int gamma, int delta, int s);
int load_sequences(ba_engine* e, const uint8_t* residues, const uint8_t* classes, const int64_t* offsets, int64_t n_seq);
int load_pairs(ba_engine* e, const int32_t* seq_a, const int32_t* seq_b, int64_t n_pairs);
int set_pair_mu2(ba_engine* e, const int32_t* mu2, const int64_t* offsets);
int run(ba_engine* e, int want_trace);
int fetch_scores(ba_engine* e, int64_t* scores);
int trace_bytes(ba_engine* e, int64_t* total);
int fetch_traces(ba_engine* e, uint8_t* cols, int64_t* offsets, uint8_t* complete);
int debug_route(ba_engine* e, int64_t pair, ba_engine** kid, int64_t* local);
}  // namespace multi

extern "C" {

const char* ba_version(void) { return "bialign_b200 0.1 (sm_100a)"; }

int ba_engine_create(int device, ba_engine** out) {
    ba_engine* e = nullptr;
    if (!out) return fail(nullptr, BA_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, BA_ERR_NO_DEVICE,
                    std::string("no CUDA device: ") + (ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce)) +
                        " (bialign_b200 has no CPU path)");
    if (device < 0 || device >= ndev) return fail(nullptr, BA_ERR_INVALID_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, BA_ERR_CUDA, cudaGetErrorString(ce));
    if (prop.major != 10)
        return fail(nullptr, BA_ERR_NO_DEVICE,
                    std::string("device '") + prop.name + "' is not sm_100 (kernels are built for sm_100a only)");
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, BA_ERR_CUDA, cudaGetErrorString(ce));
    e = new ba_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete e;
        return fail(nullptr, BA_ERR_CUDA, cudaGetErrorString(ce));
    }
    e->stats.device = device;
    *out = e;
    return BA_OK;
}

void ba_engine_destroy(ba_engine* e) {
    if (!e) return;
    if (!e->kids.empty()) { multi::destroy(e); return; }
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    e->d_sim.release(); e->d_res.release(); e->d_cls.release(); e->d_desc.release(); e->d_codes.release();
    e->d_scratch.release(); e->d_counter.release(); e->d_scores.release(); e->d_start.release();
    e->d_complete.release(); e->d_trace.release(); e->d_endv.release(); e->d_tlen.release(); e->h_stage.release();
    e->d_simp.release(); e->d_tbtab.release(); e->d_bnd.release(); e->d_progress.release();
    e->d_mu2.release(); e->d_mu2_off.release(); e->d_chains.release(); e->d_tile_order.release(); e->d_colbuf.release();
    e->d_rowmax.release(); e->d_simp1.release(); e->d_row_off.release(); e->d_suspect.release(); e->d_desc2.release();
    cudaStreamDestroy(e->stream);
    delete e;
}

const char* ba_last_error(const ba_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int ba_set_option(ba_engine* e, const char* key, int64_t value) {
    if (!e || !key) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::set_option(e, key, value);
    auto tri = [&](int* dst) {  // -1 auto, 0 off, 1 force
        if (value < -1 || value > 1) return fail(e, BA_ERR_INVALID_ARG, std::string(key) + " must be -1 (auto), 0 or 1");
        *dst = (int)value;
        return (int)BA_OK;
    };
    if (!strcmp(key, "code_arena_bytes")) {
        if (value < 0) return fail(e, BA_ERR_INVALID_ARG, "code_arena_bytes must be >= 0 (0 = auto)");
        e->opt_code_arena_bytes = value;
    }
    else if (!strcmp(key, "kernel")) return tri(&e->opt_kernel);
    else if (!strcmp(key, "pad")) return tri(&e->opt_pad);
    else if (!strcmp(key, "long")) return tri(&e->opt_long);
    else if (!strcmp(key, "io_warp")) return tri(&e->opt_io_warp);
    else if (!strcmp(key, "col_chunks")) {
        if (value < 0 || value > 64) return fail(e, BA_ERR_INVALID_ARG, "col_chunks must be in 0..64 (0 = auto, 1 = none)");
        e->opt_col_chunks = (int)value;
    }
    else if (!strcmp(key, "p16")) return tri(&e->opt_p16);
    else if (!strcmp(key, "na_kernel")) return tri(&e->opt_na);
    else if (!strcmp(key, "chain")) return tri(&e->opt_chain);
    else if (!strcmp(key, "rebase")) return tri(&e->opt_rebase);
    else if (!strcmp(key, "rebase_window")) {
        if (value < 0) return fail(e, BA_ERR_INVALID_ARG, "rebase_window must be >= 0 (0 = the full window)");
        e->opt_rebase_window = value;
    }
    else if (!strcmp(key, "warps_per_cta")) {
        if (value < 0 || value > 8) return fail(e, BA_ERR_INVALID_ARG, "warps_per_cta must be in 0..8 (0 = auto)");
        e->opt_warps = (int)value;
    }
    else return fail(e, BA_ERR_INVALID_ARG, std::string("unknown option ") + key);
    return BA_OK;
}

int ba_set_scoring(ba_engine* e, const int32_t* sim, int nsym, int structure_weight, int gap_opening_cost,
                   int gap_cost, int shift_cost, int max_shift) {
    if (!e) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::set_scoring(e, sim, nsym, structure_weight, gap_opening_cost, gap_cost, shift_cost, max_shift);
    if (!sim || nsym <= 0 || nsym > 256) return fail(e, BA_ERR_INVALID_ARG, "sim is NULL or nsym not in 1..256");
    if (max_shift < 0 || max_shift > BA_MAX_SHIFT)
        return fail(e, BA_ERR_INVALID_ARG, "max_shift must be in 0.." + std::to_string(BA_MAX_SHIFT));
    // (the systolic kernels are instantiated for max_shift <= BA_SYSTOLIC_MAX_SHIFT; larger bands run on the
    // general level kernel)
    CU(cudaSetDevice(e->device));
    e->sc = Scoring{structure_weight, gap_opening_cost, gap_cost, shift_cost, max_shift, nsym};
    e->h_sim.assign(sim, sim + (size_t)nsym * nsym);
    e->max_abs_sim = 0;
    for (int q = 0; q < nsym * nsym; ++q) e->max_abs_sim = std::max<int64_t>(e->max_abs_sim, std::llabs((long long)sim[q]));
    CU(e->d_sim.ensure((size_t)nsym * nsym));
    CU(cudaMemcpyAsync(e->d_sim.p, sim, sizeof(int) * nsym * nsym, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->have_scoring = true;
    e->ran = false;
    return BA_OK;
}

int ba_load_sequences(ba_engine* e, const uint8_t* residues, const uint8_t* classes, const int64_t* offsets,
                      int64_t n_seq) {
    if (!e) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::load_sequences(e, residues, classes, offsets, n_seq);
    if (!offsets || n_seq < 0) return fail(e, BA_ERR_INVALID_ARG, "offsets is NULL or n_seq < 0");
    for (int64_t q = 0; q < n_seq; ++q) {
        if (offsets[q + 1] < offsets[q] || offsets[q] < 0) return fail(e, BA_ERR_INVALID_ARG, "offsets not monotone");
        if (offsets[q + 1] - offsets[q] > BA_MAX_SEQ_LEN)
            return fail(e, BA_ERR_INVALID_ARG, "sequence " + std::to_string(q) + " is longer than BA_MAX_SEQ_LEN");
    }
    const int64_t total = n_seq ? offsets[n_seq] : 0;
    if (total > 0 && (!residues || !classes)) return fail(e, BA_ERR_INVALID_ARG, "residues/classes NULL");
    if (total - (n_seq ? offsets[0] : 0) > (int64_t)1 << 40) return fail(e, BA_ERR_INVALID_ARG, "sequence table too large");
    CU(cudaSetDevice(e->device));
    e->h_off.assign(offsets, offsets + n_seq + 1);
    e->n_seq = n_seq;
    int mx = 0;
    for (int64_t q = 0; q < total; ++q) mx = std::max<int>(mx, residues[q]);
    e->max_residue = mx;
    CU(e->d_res.ensure((size_t)total));
    CU(e->d_cls.ensure((size_t)total));
    if (total) {
        CU(cudaMemcpyAsync(e->d_res.p, residues, (size_t)total, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(e->d_cls.p, classes, (size_t)total, cudaMemcpyHostToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    e->have_seqs = true;
    e->have_pairs = false;
    e->ran = false;
    return BA_OK;
}

int ba_load_pairs(ba_engine* e, const int32_t* seq_a, const int32_t* seq_b, int64_t n_pairs) {
    if (!e) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::load_pairs(e, seq_a, seq_b, n_pairs);
    if (!e->have_seqs) return fail(e, BA_ERR_STATE, "ba_load_sequences first");
    if (n_pairs < 0 || (n_pairs > 0 && (!seq_a || !seq_b))) return fail(e, BA_ERR_INVALID_ARG, "pair arrays NULL");
    if (n_pairs > 0x7fffffff) return fail(e, BA_ERR_INVALID_ARG, "too many pairs for one call");
    for (int64_t p = 0; p < n_pairs; ++p)
        if (seq_a[p] < 0 || seq_a[p] >= e->n_seq || seq_b[p] < 0 || seq_b[p] >= e->n_seq)
            return fail(e, BA_ERR_INVALID_ARG, "pair " + std::to_string(p) + " references a sequence outside the table");
    e->h_pa.assign(seq_a, seq_a + n_pairs);
    e->h_pb.assign(seq_b, seq_b + n_pairs);
    e->n_pairs = n_pairs;
    e->have_pairs = true;
    e->have_mu2 = false;
    e->ran = false;
    return BA_OK;
}

int ba_set_pair_mu2(ba_engine* e, const int32_t* mu2, const int64_t* offsets) {
    if (!e) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::set_pair_mu2(e, mu2, offsets);
    if (!e->have_pairs) return fail(e, BA_ERR_STATE, "ba_load_pairs first");
    e->have_mu2 = false;
    e->ran = false;
    if (!mu2 && !offsets) return BA_OK;  // back to class-equality scoring
    if (!offsets) return fail(e, BA_ERR_INVALID_ARG, "offsets is NULL");
    const int64_t N = e->n_pairs;
    for (int64_t p = 0; p < N; ++p) {
        const int64_t n = e->h_off[e->h_pa[p] + 1] - e->h_off[e->h_pa[p]], m = e->h_off[e->h_pb[p] + 1] - e->h_off[e->h_pb[p]];
        if (offsets[p + 1] - offsets[p] != n * m)
            return fail(e, BA_ERR_INVALID_ARG, "mu2 matrix of pair " + std::to_string(p) + " must hold len(A) x len(B) entries");
    }
    const int64_t total = N ? offsets[N] - offsets[0] : 0;
    if (total > 0 && !mu2) return fail(e, BA_ERR_INVALID_ARG, "mu2 is NULL");
    CU(cudaSetDevice(e->device));
    e->max_abs_mu2 = 0;
    for (int64_t q = 0; q < total; ++q) e->max_abs_mu2 = std::max<int64_t>(e->max_abs_mu2, std::llabs((long long)mu2[offsets[0] + q]));
    std::vector<long long> off((size_t)N + 1);
    for (int64_t p = 0; p <= N; ++p) off[p] = offsets[p] - offsets[0];
    CU(e->d_mu2.ensure((size_t)std::max<int64_t>(total, 1)));
    CU(e->d_mu2_off.ensure((size_t)N + 1));
    if (total) CU(cudaMemcpyAsync(e->d_mu2.p, mu2 + offsets[0], sizeof(int) * total, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_mu2_off.p, off.data(), sizeof(long long) * (N + 1), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->have_mu2 = true;
    return BA_OK;
}

int ba_run(ba_engine* e, int want_trace) {
    if (!e) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::run(e, want_trace);
    if (!e->have_scoring) return fail(e, BA_ERR_STATE, "ba_set_scoring first");
    if (!e->have_pairs) return fail(e, BA_ERR_STATE, "ba_load_sequences and ba_load_pairs first");
    const bool affine = e->sc.beta != 0;  // pyx:203-205
    CU(cudaSetDevice(e->device));
    const int64_t N = e->n_pairs;
    const int s = e->sc.s;
    const bool timing = getenv("BA_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto t_start = tnow();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t = tnow();
        fprintf(stderr, "[ba_run] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_start).count());
        t_start = t;
    };
    e->ran = false;
    e->have_tlen = false;
    e->stats = ba_stats{};
    e->stats.device = e->device;
    e->stats.pairs = N;
    if (e->max_residue >= e->sc.nsym)
        return fail(e, BA_ERR_ALPHABET, "residue code " + std::to_string(e->max_residue) + " >= nsym " + std::to_string(e->sc.nsym));

    // ---- schedule: descending cost (LPT inside the device), waves that fit the code arena
    std::vector<int32_t> order((size_t)N);
    std::iota(order.begin(), order.end(), 0);
    std::vector<int32_t> ln((size_t)N), lm((size_t)N);
    int nmax = 0, mmax = 0;
    int64_t cs = 0;
    for (int64_t p = 0; p < N; ++p) {
        ln[p] = (int32_t)(e->h_off[e->h_pa[p] + 1] - e->h_off[e->h_pa[p]]);
        lm[p] = (int32_t)(e->h_off[e->h_pb[p] + 1] - e->h_off[e->h_pb[p]]);
        nmax = std::max(nmax, ln[p]);
        mmax = std::max(mmax, lm[p]);
        cs += (affine ? 9 : 1) * band_cells(ln[p], lm[p], s);
    }
    e->stats.cell_states = cs;
    bool wide = false;  // values need the reference's int64 tables (pyx:27-35): general level kernel, 64-bit instantiation
    {   // int32 exactness bound of SURVEY 8a-6
        const int64_t col = e->max_abs_sim + std::max<int64_t>(std::llabs((long long)e->sc.w), e->have_mu2 ? e->max_abs_mu2 : 0) +
                            2 * std::llabs((long long)e->sc.beta) +
                            2 * std::llabs((long long)e->sc.gamma) + 2 * std::llabs((long long)e->sc.delta);
        if ((int64_t)(nmax + mmax + 2) * col >= ((int64_t)1 << 30)) {
            if (e->opt_kernel == 1)
                return fail(e, BA_ERR_SCORE_RANGE, "score bound exceeds int32 range: only the general level kernel has a 64-bit instantiation (kernel = 1 requested)");
            if ((double)(nmax + mmax + 2) * (double)col >= 4.0e18)
                return fail(e, BA_ERR_SCORE_RANGE, "score bound exceeds int64 range");
            wide = true;
        }
    }
    lap("lengths + range check");
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return (int64_t)(ln[x] + 1) * (lm[x] + 1) > (int64_t)(ln[y] + 1) * (lm[y] + 1);
    });

    e->h_desc.resize((size_t)N);
    e->h_slot_off.assign((size_t)N, 0);
    e->h_slot_cap.assign((size_t)N, 0);
    e->h_last_code_off.assign((size_t)N, -1);
    int64_t trace_total = 0;
    if (want_trace)
        for (int64_t p = 0; p < N; ++p) {
            e->h_slot_off[p] = trace_total;
            e->h_slot_cap[p] = 2 * (ln[p] + lm[p]) + 2;
            trace_total += (e->h_slot_cap[p] + 15) & ~15;
        }

    lap("sort + slots");
    // ---- kernel choice: systolic when its exactness conditions hold, else the generic level kernel
    SysPlan plan;
    bool p16 = false;
    if (e->opt_kernel != 0 && affine && !want_trace && e->opt_p16 != 0 && e->opt_pad != 1 && N >= 2 && !e->have_mu2 && !wide &&
        !(e->opt_chain == 1 && e->opt_p16 != 1)) {  // a forced chained flavour keeps the 32-bit kernel
        plan = plan_systolic(e, nmax, mmax, false, true);  // do the scores provably fit 16 bits?
        p16 = plan.ok;
    }
    if (e->opt_p16 == 1 && !p16 && !want_trace && affine && N >= 2)
        return fail(e, BA_ERR_SCORE_RANGE, "16-bit pair mode requested but the score range does not fit");
    if (wide) {
        p16 = false;
        plan = SysPlan{};
    } else
    if (e->have_mu2) {  // per-pair mu2 matrices: only the general level kernel reads them
        if (e->opt_kernel == 1) return fail(e, BA_ERR_INVALID_ARG, "per-pair mu2 matrices run on the general level kernel (kernel = 1 requested)");
        p16 = false;
        plan = SysPlan{};
    } else
    if (!p16 && e->opt_kernel != 0) plan = plan_systolic(e, nmax, mmax, want_trace != 0, false, !affine);
    // Traces whose packed plan (value << tie bits in 32 bits) fails although the plain values fit: two launches, the second in
    // coordinates relative to the row maxima recorded by the first (REBASE flavour); pairs it cannot vouch for are recomputed
    // by the level kernel at the end of this function.
    bool rebase = false;
    SysPlan plan1;  // the score-only launch of a rebased run
    // (also preferred over the padded flavour when only that one fits: measured 526 vs 437 GCUPS on 200-500 aa pairs)
    if (want_trace && affine && !wide && !e->have_mu2 && e->opt_kernel != 0 && e->opt_rebase != 0 &&
        (!plan.ok || e->opt_rebase == 1 || (plan.pad && e->opt_pad < 0))) {
        SysPlan p2;
        if (plan_rebase(e, nmax, mmax, plan1, p2)) {
            plan = p2;
            rebase = true;
        }
    }
    if (e->opt_rebase == 1 && !rebase && want_trace)
        return fail(e, BA_ERR_SCORE_RANGE, "rebased trace run requested but not applicable (model, range or window)");
    if (e->opt_kernel == 1 && !plan.ok)
        return fail(e, BA_ERR_SCORE_RANGE, "systolic kernel requested but its packed-integer range conditions do not hold");
    int kernel = plan.ok ? 1 : 0;
    // non-affine model: the dedicated row-per-lane kernel (fill_na.cu) when the pad-free range conditions hold
    bool na_ded = kernel == 1 && !affine && !plan.pad && s <= BA_NA_MAX_SHIFT && e->opt_na != 0;
    if (e->opt_na == 1 && !affine && !na_ded && e->opt_kernel != 0)
        return fail(e, BA_ERR_SCORE_RANGE, "dedicated non-affine kernel requested but not applicable (range or max_shift)");
    int max_grid = e->sm_count * 2;
    size_t scratch_stride = 0, sys_smem = 0;
    int sysG = e->opt_warps;
    SysArgs SA{}, SA1{};  // SA1: the score-only launch of a rebased run
    int64_t hi_plane_off = 0;  // systolic code arena: slot index at which the 16-bit plane starts (in 32-bit words of the arena)
    bool long_mode = false;
    int io_warp = 0;
    size_t dbg_nts = 0;
    int ntc = 1, chunk_cols = 0, tile_iters = 0;  // column-chunked tiles (long-pair mode with the I/O warp, one pair per launch)
    int long_grid_max = 0, sys_occ = 0;
    if (na_ded) {
        // CTA width: a row block is 32 rows per warp and lags 32 iterations per warp; estimate warp-iterations per pair
        if (sysG == 0) {
            double best = 0;
            for (int G = 1; G <= 8; ++G) {
                const size_t sm = na_smem_bytes(s, G, e->sc.nsym, mmax);
                if (sm > kSysSmemLimit) continue;
                const int occ = na_occupancy(s, want_trace != 0, G, sm);
                if (occ < 1) continue;
                double cost = 0;
                const int64_t stride = std::max<int64_t>(1, N / 4096);
                for (int64_t p = 0; p < N; p += stride) cost += (double)((ln[p] + 32 * G) / (32 * G)) * (na_iters(s, G, lm[p]) + na_pre(s)) * G;
                cost /= std::min(occ * G, 16);
                if (sysG == 0 || cost < best * 0.97) { best = cost; sysG = G; }
            }
            if (sysG == 0) sysG = 1;
        }
        sys_smem = na_smem_bytes(s, sysG, e->sc.nsym, mmax);
        sys_occ = sys_smem > kSysSmemLimit ? 0 : na_occupancy(s, want_trace != 0, sysG, sys_smem);
        if (sys_occ < 1) na_ded = false, sysG = e->opt_warps;  // molecule B too long for shared memory: systolic flavour / level kernel below
    }
    if (kernel == 1 && !na_ded) {
        if (sysG == 0) {
            // Pick the CTA width that minimises the estimated warp-iterations per resident warp:
            // a pair costs passes(G) * iterations(G) on G warps; an SM runs occ(G) CTAs, and more than ~16
            // resident warps do not add throughput.  Short pairs want a CTA that
            // covers all rows in one pass; long ones want few warps per CTA and several CTAs per SM.
            const SysGeo geo = sys_geo(s, plan.pad);
            double best = 0;
            for (int G = std::max(2, (12 * geo.LPR + 31) / 32); G <= 8; ++G) {
                const size_t sm = sys_smem_bytes(s, plan.pad, G, e->sc.nsym, mmax, p16);
                if (sm > kSysSmemLimit) continue;
                const int occ = p16 ? sys_occupancy_p16(s, G, sm)
                                    : !affine ? sys_occupancy_na(s, want_trace != 0, plan.pad, G, sm)
                                              : sys_occupancy(s, want_trace != 0, plan.pad, plan.bneg, G, sm);
                if (occ < 1) continue;
                const double eff = std::min(occ * G, 16);
                double cost = 0;
                const int64_t stride = std::max<int64_t>(1, N / 4096);  // sample large batches
                for (int64_t p = 0; p < N; p += stride) {
                    const int rows = G * geo.R;
                    const double passes = (ln[p] + rows) / rows;
                    cost += passes * (double)sys_iters(s, plan.pad, G, lm[p]) * G;
                }
                cost /= eff;
                if (sysG == 0 || cost < best * 0.97) { best = cost; sysG = G; }  // prefer the narrower CTA on near-ties
            }
            if (sysG == 0) sysG = 2;
            // a handful of long pairs will run in long-pair mode, where the pipeline fill (CTAs x lag) matters as much
            // as the per-CTA rate: 4 warps measured best on both the 928 x 933 and the 8192 x 8192 pair
            if (N <= 4 && affine && !p16 && e->opt_long != 0 && (nmax + 1) > 8 * 4 * geo.R) sysG = std::max(4, (12 * geo.LPR + 31) / 32);
        }
        // one boundary-record element per thread: a CTA needs at least 12 * LPR threads, and never fewer -- a narrower
        // CTA would silently drop record elements between the row blocks of a multi-pass pair
        const int minG = (12 * sys_geo(s, plan.pad).LPR + 31) / 32;
        sysG = std::max(sysG, minG);
        while (sysG > minG && sys_smem_bytes(s, plan.pad, sysG, e->sc.nsym, mmax, p16) > kSysSmemLimit) --sysG;
        sys_smem = sys_smem_bytes(s, plan.pad, sysG, e->sc.nsym, mmax, p16);
        sys_occ = sys_smem > kSysSmemLimit ? 0
                        : rebase ? std::min(sys_occupancy_rebase(s, false, false, sysG, sys_smem), sys_occupancy_rebase(s, true, false, sysG, sys_smem))
                        : p16 ? sys_occupancy_p16(s, sysG, sys_smem)
                              : !affine ? sys_occupancy_na(s, want_trace != 0, plan.pad, sysG, sys_smem)
                                        : sys_occupancy(s, want_trace != 0, plan.pad, plan.bneg, sysG, sys_smem);
        if (sys_occ < 1) {
            // molecule B is staged in shared memory: a very long B does not fit next to the rings even at the
            // minimum CTA width.  The general level kernel has no such limit.
            if (e->opt_kernel == 1)
                return fail(e, BA_ERR_CUDA, "systolic kernel does not fit on an SM (shared memory " + std::to_string(sys_smem) + ")");
            kernel = 0;
            p16 = false;
            rebase = false;
        }
    }
    // Short pairs: when every pair fits ONE row block of a CTA of at most 8 warps, chains of pairs run back to back through
    // the systolic array (CHAIN flavour): the pipeline skew is paid per chain, not per pair.
    bool chain_mode = false;
    constexpr int kChainBudget = 3072;  // staged-B bytes of one chain (sum of m + 2s + 6 over its pairs)
    if (kernel == 1 && !na_ded && affine && !p16 && !rebase && !plan.pad && plan.bneg && e->opt_chain != 0 && e->opt_long != 1 && N >= 2 &&
        mmax + 2 * s + 6 <= kChainBudget) {
        const SysGeo geo = sys_geo(s, false);
        const int minG = (12 * geo.LPR + 31) / 32;
        const int Gc = std::max({(nmax + geo.R) / geo.R, minG, e->opt_warps});  // ceil((nmax+1)/R) rows in one block
        if (Gc <= 8) {
            const size_t sm = sys_smem_bytes_chain(s, Gc, e->sc.nsym, sys_bpad(s, false, Gc, kChainBudget));
            const int occ = sm > kSysSmemLimit ? 0 : sys_occupancy_chain(s, want_trace != 0, Gc, sm);
            if (occ >= 1) {
                chain_mode = true;
                sysG = Gc;
                sys_smem = sm;
                sys_occ = occ;
            }
        }
    }
    if (e->opt_chain == 1 && !chain_mode && N >= 2)
        return fail(e, BA_ERR_INVALID_ARG, "chained short-pair mode requested but not applicable (a pair needs several row blocks, or another flavour is forced)");
    // Code words of one pair: the level kernel uses the cell-major table of common.cuh; the systolic kernel writes in the
    // order it computes -- [row block][warp][iteration][lane], 256 contiguous bytes per warp and iteration -- so its table
    // also holds the skewed pipeline's idle slots (about 10-15 % more memory, 15x fewer store transactions).
    // The arena is counted in units: 8-byte words for the level kernels and the dedicated non-affine kernel, 6-byte slots
    // (a 32-bit plane followed by a 16-bit plane) for the systolic kernel.
    const bool planes = kernel == 1 && !na_ded;
    const int64_t unit = planes ? 6 : 8;
    auto pair_code_words = [&](int n, int m) -> int64_t {
        if (kernel != 1) return code_words(n, m, s);
        if (na_ded) return na_code_words(s, sysG, n, m);
        return sys_code_words(s, plan.pad, sysG, n, m);
    };
    auto arena_cap_units = [&]() -> int64_t { return planes ? std::max<int64_t>(0, (int64_t)e->d_codes.cap * 8 / 6 - 64) : (int64_t)e->d_codes.cap; };
    auto arena_alloc_words = [&](int64_t units) -> size_t { return planes ? (size_t)(((units + 64) * 6 + 7) / 8) : (size_t)units; };
    // code arena
    size_t arena_words = 0;
    std::vector<int64_t> wave_begin;  // indices into sorted order
    if (want_trace && N) {
        int64_t total_words = 0, max_words = 0;
        for (int64_t p = 0; p < N; ++p) {
            const int64_t wds = pair_code_words(ln[p], lm[p]);
            total_words += wds;
            max_words = std::max(max_words, wds);
        }
        // The arena is sticky: once allocated it is reused as long as the largest pair fits (a 90 GB
        // cudaFree + cudaMalloc costs tens of milliseconds and free memory drifts from run to run).
        // An explicit "code_arena_bytes" is an upper bound on what a wave may use (at least one pair always fits).
        const int64_t limit_words = e->opt_code_arena_bytes > 0 ? std::max<int64_t>(e->opt_code_arena_bytes / unit, max_words)
                                                                 : std::numeric_limits<int64_t>::max();
        const int64_t want_words = std::min(total_words, limit_words);
        if (arena_cap_units() >= max_words &&
            (arena_cap_units() >= want_words || (e->opt_code_arena_bytes <= 0 && e->arena_is_budget))) {
            arena_words = (size_t)std::min<int64_t>(arena_cap_units(), want_words);
        } else {
            int64_t budget_words = limit_words;
            if (e->opt_code_arena_bytes <= 0) {
                size_t fr = 0, tot = 0;
                CU(cudaMemGetInfo(&fr, &tot));
                budget_words = std::max<int64_t>((int64_t)(fr + e->d_codes.cap * 8) / 2 / unit - 64, max_words);
            }
            arena_words = (size_t)std::min(budget_words, total_words);
            e->arena_is_budget = e->opt_code_arena_bytes <= 0 && budget_words <= total_words;
        }
        cudaError_t ce = e->d_codes.ensure(arena_alloc_words((int64_t)arena_words));
        // automatic sizing only: a failed allocation (fragmentation, another tenant on the GPU) is retried with half
        // the arena, i.e. more waves, down to the largest single pair
        while (ce == cudaErrorMemoryAllocation && e->opt_code_arena_bytes <= 0 && (int64_t)arena_words > max_words) {
            cudaGetLastError();  // clear the sticky allocation error
            arena_words = (size_t)std::max<int64_t>((int64_t)arena_words / 2, max_words);
            e->arena_is_budget = true;
            ce = e->d_codes.ensure(arena_alloc_words((int64_t)arena_words));
        }
        if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "traceback-code arena: " + std::string(cudaGetErrorString(ce)));
    }
    {
        int64_t used = 0;
        wave_begin.push_back(0);
        for (int64_t q = 0; q < N; ++q) {
            const int32_t p = order[q];
            PairDesc& d = e->h_desc[q];
            d.offA = e->h_off[e->h_pa[p]];
            d.offB = e->h_off[e->h_pb[p]];
            d.n = ln[p];
            d.m = lm[p];
            d.orig = p;
            d.trace_off = e->h_slot_off[p];
            d.trace_cap = e->h_slot_cap[p];
            d.code_off = 0;
            if (want_trace) {
                const int64_t wds = pair_code_words(d.n, d.m);
                if (used + wds > (int64_t)arena_words) {
                    wave_begin.push_back(q);
                    used = 0;
                }
                d.code_off = used;
                used += wds;
            }
        }
        wave_begin.push_back(N);
    }
    const int n_waves = (int)wave_begin.size() - 1;
    std::vector<int> h_chains;          // per wave: chain start indices (relative to the wave's first pair), one extra end entry
    std::vector<int64_t> chain_begin;   // per wave: offset of its entries in h_chains
    if (chain_mode) {
        for (int w = 0; w < n_waves; ++w) {
            chain_begin.push_back((int64_t)h_chains.size());
            const int64_t b = wave_begin[w], cnt = wave_begin[w + 1] - b;
            int64_t q = 0;
            while (q < cnt) {
                h_chains.push_back((int)q);
                int used = 0, len = 0;
                while (q < cnt && len < BA_KCHAIN && used + e->h_desc[b + q].m + 2 * s + 6 <= kChainBudget) {
                    used += e->h_desc[b + q].m + 2 * s + 6;
                    ++len;
                    ++q;
                }
            }
            h_chains.push_back((int)cnt);
        }
        chain_begin.push_back((int64_t)h_chains.size());
    }
    lap("arena + waves");

    // ---- buffers
    CU(e->d_desc.ensure((size_t)N));
    CU(e->d_scores.ensure((size_t)N));
    CU(e->d_start.ensure((size_t)N));
    CU(e->d_endv.ensure((size_t)N * 9));
    CU(e->d_counter.ensure((size_t)std::max(2 * n_waves, 1)));
    if (want_trace) {
        CU(e->d_tlen.ensure((size_t)N));
        CU(e->d_complete.ensure((size_t)N));
        CU(e->d_trace.ensure((size_t)trace_total));
    }
    if (N == 0) {
        e->ran = true;
        e->ran_trace = want_trace != 0;
        return BA_OK;
    }
    lap("result buffers");
    int64_t biggest_wave = 0;
    for (int w = 0; w < n_waves; ++w) biggest_wave = std::max(biggest_wave, wave_begin[w + 1] - wave_begin[w]);
    if (na_ded) {
        max_grid = e->sm_count * sys_occ;
        const int grid = (int)std::min<int64_t>(biggest_wave, max_grid);
        CU(e->d_simp.ensure(plan.sim_p.size()));
        CU(cudaMemcpyAsync(e->d_simp.p, plan.sim_p.data(), plan.sim_p.size() * 4, cudaMemcpyHostToDevice, e->stream));
        if (nmax + 1 > sysG * 32) {  // pairs with several row blocks hand a boundary stream from block to block
            cudaError_t ce = e->d_bnd.ensure((size_t)grid * 2 * na_boundary_ints(s, sysG, mmax));
            if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "boundary streams: " + std::string(cudaGetErrorString(ce)));
        }
        const int64_t sh = (int64_t)1 << plan.tb;
        SA.res = e->d_res.p; SA.cls = e->d_cls.p; SA.sim_p = e->d_simp.p; SA.sc = e->sc;
        SA.w_p = (int)(e->sc.w / plan.g * sh);
        SA.k_gd = (int)((e->sc.gamma + e->sc.delta) / plan.g * sh); SA.k_2g = (int)(2 * e->sc.gamma / plan.g * sh);
        SA.k_d = (int)(e->sc.delta / plan.g * sh);
        SA.negp = plan.negp; SA.tb_bits = plan.tb; SA.gscale = plan.g; SA.mmax = mmax;
        SA.boff = na_boff(s, sysG); SA.bpad = na_bpad(s, sysG, mmax);
        SA.bnd = e->d_bnd.p; SA.bnd_iters = (int)(na_boundary_ints(s, sysG, mmax) / 8);
        SA.codes = want_trace ? e->d_codes.p : nullptr;
        SA.scores = e->d_scores.p; SA.start_state = e->d_start.p; SA.end_values = e->d_endv.p;
    } else if (kernel == 1) {
        max_grid = e->sm_count * sys_occ;
        const int grid = (int)std::min<int64_t>(biggest_wave, max_grid);
        const int rows_pass = sysG * sys_geo(s, plan.pad).R;
        const bool multi_pass = nmax + 1 > rows_pass;
        const int biters = sys_iters(s, plan.pad, sysG, mmax);
        CU(e->d_simp.ensure(plan.sim_p.size()));
        CU(cudaMemcpyAsync(e->d_simp.p, plan.sim_p.data(), plan.sim_p.size() * 4, cudaMemcpyHostToDevice, e->stream));
        if (want_trace) {
            CU(e->d_tbtab.ensure(plan.tbtab.size()));
            CU(cudaMemcpyAsync(e->d_tbtab.p, plan.tbtab.data(), plan.tbtab.size() * 4, cudaMemcpyHostToDevice, e->stream));
        }
        // Long pairs, fewer per wave than half the CTAs the GPU holds (a wave is what fits the code arena): spread the row
        // blocks of each pair over a gang of CTAs (LONG flavour, cooperative launch), the whole grid for a single pair
        const int npass_max = (nmax + rows_pass) / rows_pass;
        if (!chain_mode && !p16 && affine && plan.bneg && e->opt_long != 0 && npass_max >= 2 && sysG >= 2 &&
            (e->opt_long == 1 || (npass_max >= 8 && biggest_wave * 2 <= max_grid))) {
            // The I/O warp (one more warp per CTA that owns the boundary I/O and the progress flags) pays where the pipeline is
            // latency-bound: a handful of pairs whose row blocks are (nearly) all resident.  Measured: 928 x 933 pair 2.69 ->
            // 2.16 ms, 8192 x 8192 +7 %; gangs of 2000-aa pairs at max_shift 2 lose 13 % (three CTAs per SM instead of four).
            io_warp = (!rebase && (e->opt_io_warp == 1 || (e->opt_io_warp < 0 && N <= 4))) ? 1 : 0;
            const int occl = rebase ? std::min(sys_occupancy_rebase(s, false, true, sysG, sys_smem), sys_occupancy_rebase(s, true, true, sysG, sys_smem))
                                    : sys_occupancy_long(s, want_trace != 0, plan.pad, sysG, sys_smem, io_warp != 0);
            int coop = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
            if (occl >= 1 && coop) {
                long_mode = true;
                long_grid_max = e->sm_count * occl;
            } else if (e->opt_long == 1) {
                return fail(e, BA_ERR_CUDA, "long-pair mode requested but cooperative launch is unavailable");
            }
        }
        // Column chunks: when one pair has more row blocks than the GPU holds CTAs, a second (partly empty) round of row blocks
        // would cost a whole pass; instead the row blocks are cut into column chunks and the tiles are dealt in start order.
        // Runs of a single pair only.
        if (long_mode && io_warp && !plan.pad && !rebase && N == 1 && e->opt_col_chunks != 1) {
            const SysGeo geo = sys_geo(s, false);
            const int64_t its = (int64_t)(mmax + 1) * geo.P;
            // (not automatic: measured on the 8192 x 8192 pair, 14 chunks 57.5 ms against 52 ms without -- every row block of
            // every chunk starts one pipeline lag (~50 us) after the one above, and with chunks that lag is paid 512 times on
            // the critical path instead of 216 times)
            (void)its;
            int want = e->opt_col_chunks >= 2 ? e->opt_col_chunks : 1;
            const char* ov = getenv("BA_COL_CHUNKS");
            if (ov) want = std::max(1, atoi(ov));
            want = (int)std::min<int64_t>(want, (mmax + 1) / 16);  // at least 16 columns per chunk
            if (want >= 2) {
                chunk_cols = (mmax + 1 + want - 1) / want;
                ntc = (mmax + 1 + chunk_cols - 1) / chunk_cols;
                tile_iters = chunk_cols * geo.P + 2 * (rows_pass - 1) + geo.LPR + geo.RING;
            }
        }
        if (long_mode && ntc > 1) {
            const SysGeo geo = sys_geo(s, false);
            const size_t ntiles = (size_t)npass_max * ntc;
            cudaError_t ce = e->d_bnd.ensure(ntiles * (size_t)(tile_iters + 8) * geo.REC);
            if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "boundary streams (tiles): " + std::string(cudaGetErrorString(ce)));
            CU(e->d_progress.ensure(std::max<size_t>(ntiles, (size_t)long_grid_max * 2)));
            const size_t rowsz = (((size_t)npass_max * rows_pass + 2) * geo.LPR + 31) & ~(size_t)31;
            ce = e->d_colbuf.ensure((size_t)(ntc - 1) * geo.P * 12 * rowsz);
            if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "column buffer: " + std::string(cudaGetErrorString(ce)));
            CU(e->d_tile_order.ensure((size_t)ntc));
            SA.ntc = ntc; SA.chunk_cols = chunk_cols; SA.tile_next = e->d_tile_order.p; SA.colbuf = e->d_colbuf.p; SA.col_rowsz = (int)rowsz;
        } else if (long_mode) {
            const int lg = (int)std::min<int64_t>(long_grid_max, (int64_t)npass_max * biggest_wave);
            cudaError_t ce = e->d_bnd.ensure((size_t)lg * 2 * sys_boundary_ints(s, plan.pad, sysG, mmax));
            if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "boundary streams: " + std::string(cudaGetErrorString(ce)));
            CU(e->d_progress.ensure((size_t)long_grid_max * 2));
        } else if (multi_pass) {
            cudaError_t ce = e->d_bnd.ensure((size_t)grid * 2 * sys_boundary_ints(s, plan.pad, sysG, mmax));
            if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "boundary streams: " + std::string(cudaGetErrorString(ce)));
        }
        const int64_t sh = (int64_t)1 << plan.tb;
        SA.res = e->d_res.p; SA.cls = e->d_cls.p; SA.sim_p = e->d_simp.p; SA.tbtab = e->d_tbtab.p; SA.sc = e->sc;
        SA.w_p = (int)(e->sc.w / plan.g * sh); SA.beta_p = (int)(e->sc.beta / plan.g * sh);
        SA.k_gd = (int)((e->sc.gamma + e->sc.delta) / plan.g * sh); SA.k_2g = (int)(2 * e->sc.gamma / plan.g * sh);
        SA.k_2g2d = (int)((2 * e->sc.gamma + 2 * e->sc.delta) / plan.g * sh); SA.k_2d = (int)(2 * e->sc.delta / plan.g * sh);
        SA.k_d = (int)(e->sc.delta / plan.g * sh);
        SA.negp = plan.negp; SA.tb_bits = plan.tb; SA.gscale = plan.g; SA.mmax = mmax;
        SA.boff = sys_boff(s, plan.pad, sysG); SA.bpad = sys_bpad(s, plan.pad, sysG, chain_mode ? kChainBudget : mmax);
        if (chain_mode) {
            CU(e->d_chains.ensure(h_chains.size()));
            CU(cudaMemcpyAsync(e->d_chains.p, h_chains.data(), sizeof(int) * h_chains.size(), cudaMemcpyHostToDevice, e->stream));
        }
        SA.progress = e->d_progress.p;
        SA.io_warp = long_mode ? io_warp : 0;
        if (long_mode && N == 1 && getenv("BA_DEBUG_TS")) {
            const size_t nts = (size_t)npass_max * std::max(ntc, 1) * 8;
            CU(e->d_dbg_ts.ensure(nts));
            CU(cudaMemsetAsync(e->d_dbg_ts.p, 0, nts * 8, e->stream));
            SA.dbg_ts = e->d_dbg_ts.p;
            dbg_nts = nts;
        }
        SA.gwarps = sysG;
        {   // Flag period of the long-pair pipeline.  Many row blocks (a CTA per block, several per SM): a flag exchange costs an
            // extra barrier and a spinning thread, so it is rare (default, ~32 iterations).  Few row blocks (one CTA per SM, the pair is
            // latency-bound): every row block starts one flag period later than it could, so the period is three ring periods
            // (measured on the 928 x 933 pair, fill time: 1 period 2.65 ms, 2: 2.42, 3: 2.30, 6: 2.48).
            const SysGeo geo = sys_geo(s, plan.pad);
            const char* ov = getenv("BA_LONG_LQ");
            SA.lq_iters = ov ? std::max(1, atoi(ov)) * geo.RING : (long_mode && !io_warp && npass_max <= e->sm_count ? 3 * geo.RING : 0);
        }
        SA.bnd = e->d_bnd.p; SA.bnd_iters = (ntc > 1 ? tile_iters : biters) + 8;  // matches sys_boundary_ints: slack records in front
        SA.codes = want_trace ? e->d_codes.p : nullptr;
        hi_plane_off = ((int64_t)arena_words + 63) & ~(int64_t)63;  // the 16-bit plane starts behind the 32-bit one (128-byte aligned)
        SA.codes_hi = want_trace ? reinterpret_cast<uint16_t*>(reinterpret_cast<uint32_t*>(e->d_codes.p) + hi_plane_off) : nullptr;
        SA.scores = e->d_scores.p; SA.start_state = e->d_start.p; SA.end_values = e->d_endv.p;
        if (rebase) {
            // row maxima of every pair (caller order), "minus infinity" until the score-only launch has raised them
            std::vector<long long> roff((size_t)N);
            long long rows = 0;
            for (int64_t p = 0; p < N; ++p) { roff[p] = rows; rows += ln[p] + 1; }
            CU(e->d_rowmax.ensure((size_t)rows));
            CU(e->d_row_off.ensure((size_t)N));
            CU(e->d_suspect.ensure((size_t)N));
            CU(e->d_simp1.ensure(plan1.sim_p.size()));
            CU(cudaMemcpyAsync(e->d_row_off.p, roff.data(), sizeof(long long) * N, cudaMemcpyHostToDevice, e->stream));
            CU(cudaMemcpyAsync(e->d_simp1.p, plan1.sim_p.data(), plan1.sim_p.size() * 4, cudaMemcpyHostToDevice, e->stream));
            CU(cudaMemsetAsync(e->d_rowmax.p, 0x80, sizeof(int) * (size_t)rows, e->stream));
            CU(cudaMemsetAsync(e->d_suspect.p, 0, (size_t)N, e->stream));
            CU(cudaStreamSynchronize(e->stream));  // roff is a local
            SA.rowmax = e->d_rowmax.p; SA.row_off = e->d_row_off.p; SA.suspect = e->d_suspect.p; SA.df_max = plan.df_max;
            SA1 = SA;
            SA1.sim_p = e->d_simp1.p; SA1.tbtab = nullptr;
            SA1.w_p = (int)(e->sc.w / plan1.g); SA1.beta_p = (int)(e->sc.beta / plan1.g);
            SA1.k_gd = (int)((e->sc.gamma + e->sc.delta) / plan1.g); SA1.k_2g = (int)(2 * e->sc.gamma / plan1.g);
            SA1.k_2g2d = (int)((2 * e->sc.gamma + 2 * e->sc.delta) / plan1.g); SA1.k_2d = (int)(2 * e->sc.delta / plan1.g);
            SA1.k_d = (int)(e->sc.delta / plan1.g);
            SA1.negp = plan1.negp; SA1.tb_bits = 0; SA1.codes = nullptr;
        }
    } else {
        scratch_stride = generic_scratch_ints(nmax, s);  // sized for nine states; the non-affine kernel uses a ninth
        const int grid = (int)std::min<int64_t>(biggest_wave, max_grid);
        cudaError_t ce = e->d_scratch.ensure(scratch_stride * grid * (wide ? 2 : 1));
        if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "fill scratch: " + std::string(cudaGetErrorString(ce)));
    }
    lap("kernel plan + scratch");
    if (p16 && (N & 1)) {  // odd batch: the last work item computes its only pair in both halves
        PairDesc extra = e->h_desc[N - 1];
        extra.orig = -1;
        e->h_desc.push_back(extra);
        CU(e->d_desc.ensure((size_t)N + 1));
    }
    CU(cudaMemcpyAsync(e->d_desc.p, e->h_desc.data(), sizeof(PairDesc) * e->h_desc.size(), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(e->d_counter.p, 0, sizeof(int) * 2 * n_waves, e->stream));

    struct EventSet {  // destroyed on every exit path (the CU macro returns early on errors)
        std::vector<cudaEvent_t> v;
        ~EventSet() { for (auto x : v) if (x) cudaEventDestroy(x); }
        cudaEvent_t& operator[](size_t i) { return v[i]; }
    } ev;
    ev.v.assign((size_t)n_waves * 3 + 1, nullptr);
    for (auto& x : ev.v) CU(cudaEventCreate(&x));
    CU(cudaEventRecord(ev[0], e->stream));
    for (int w = 0; w < n_waves; ++w) {
        const int64_t b = wave_begin[w], cnt = wave_begin[w + 1] - b;
        FillArgs A{};
        A.res = e->d_res.p; A.cls = e->d_cls.p; A.sim = e->d_sim.p; A.sc = e->sc;
        A.pairs = e->d_desc.p + b; A.npairs = (int)cnt; A.counter = e->d_counter.p + w;
        A.scratch = e->d_scratch.p; A.scratch_stride = scratch_stride;
        A.codes = want_trace ? e->d_codes.p : nullptr;
        A.scores = e->d_scores.p; A.start_state = e->d_start.p; A.end_values = e->d_endv.p;
        A.mu2 = e->have_mu2 ? e->d_mu2.p : nullptr; A.mu2_off = e->d_mu2_off.p;
        const int grid = (int)std::min<int64_t>(cnt, max_grid);
        if (kernel == 1 && long_mode) {
            const int rows_pass = sysG * sys_geo(s, plan.pad).R;
            // a gang of cpp CTAs per pair; as many pairs per cooperative launch as the GPU holds gangs (pairs are sorted by
            // cost, so the pairs of one launch are of similar length)
            int64_t q = 0;
            while (q < cnt) {
                const int np0 = (e->h_desc[b + q].n + rows_pass) / rows_pass;  // row blocks of the longest pair of this launch
                // (tiles: cnt == 1; the gang is as large as the GPU holds, every tile has a flag of its own)
                const int64_t units = ntc > 1 ? (int64_t)np0 * ntc : np0;
                const int cpp = (int)std::max<int64_t>(1, std::min<int64_t>(units, long_grid_max / std::min<int64_t>(cnt - q, long_grid_max)));
                const int np = (int)std::min<int64_t>(cnt - q, long_grid_max / cpp);
                const int lg = np * cpp;
                CU(cudaMemsetAsync(e->d_progress.p, 0, sizeof(unsigned long long) * (ntc > 1 ? (size_t)units : (size_t)2 * lg), e->stream));
                if (ntc > 1) CU(cudaMemsetAsync(e->d_tile_order.p, 0, sizeof(int) * ntc, e->stream));
                SA.pairs = e->d_desc.p + b + q; SA.npairs = np; SA.cpp = cpp; SA.counter = e->d_counter.p + w;
                if (rebase) {
                    SA1.pairs = SA.pairs; SA1.npairs = np; SA1.cpp = cpp; SA1.counter = SA.counter;
                    CU(launch_fill_systolic_rebase(SA1, lg, sysG, sys_smem, false, true, e->stream));
                    CU(cudaMemsetAsync(e->d_progress.p, 0, sizeof(unsigned long long) * 2 * lg, e->stream));
                    CU(launch_fill_systolic_rebase(SA, lg, sysG, sys_smem, true, true, e->stream));
                    e->stats.kernel_launches++;
                } else
                CU(launch_fill_systolic_long(SA, lg, sysG, sys_smem, want_trace != 0, plan.pad, e->stream));
                q += np;
                if (q < cnt) e->stats.kernel_launches++;
            }
        } else if (kernel == 1 && p16) {  // two pairs per work item (score only: a single wave)
            SA.pairs = e->d_desc.p; SA.npairs = (int)((N + 1) / 2); SA.counter = e->d_counter.p + w;
            CU(launch_fill_systolic_p16(SA, (int)std::min<int64_t>((N + 1) / 2, max_grid), sysG, sys_smem, e->stream));
        } else if (na_ded) {
            SA.pairs = e->d_desc.p + b; SA.npairs = (int)cnt; SA.counter = e->d_counter.p + w;
            CU(launch_fill_na(SA, grid, sysG, sys_smem, want_trace != 0, e->stream));
        } else if (kernel == 1 && !affine) {
            SA.pairs = e->d_desc.p + b; SA.npairs = (int)cnt; SA.counter = e->d_counter.p + w;
            CU(launch_fill_systolic_na(SA, grid, sysG, sys_smem, want_trace != 0, plan.pad, e->stream));
        } else if (kernel == 1 && chain_mode) {
            const int nch = (int)(chain_begin[w + 1] - chain_begin[w]) - 1;
            SA.pairs = e->d_desc.p + b; SA.npairs = nch; SA.chains = e->d_chains.p + chain_begin[w]; SA.counter = e->d_counter.p + w;
            CU(launch_fill_systolic_chain(SA, std::min(nch, max_grid), sysG, sys_smem, want_trace != 0, e->stream));
        } else if (kernel == 1 && rebase) {
            SA1.pairs = e->d_desc.p + b; SA1.npairs = (int)cnt; SA1.counter = e->d_counter.p + n_waves + w;
            CU(launch_fill_systolic_rebase(SA1, grid, sysG, sys_smem, false, false, e->stream));
            SA.pairs = e->d_desc.p + b; SA.npairs = (int)cnt; SA.counter = e->d_counter.p + w;
            CU(launch_fill_systolic_rebase(SA, grid, sysG, sys_smem, true, false, e->stream));
            e->stats.kernel_launches++;
        } else if (kernel == 1) {
            SA.pairs = e->d_desc.p + b; SA.npairs = (int)cnt; SA.counter = e->d_counter.p + w;
            CU(launch_fill_systolic(SA, grid, sysG, sys_smem, want_trace != 0, plan.pad, plan.bneg, e->stream));
        } else if (affine) {
            launch_fill_generic(A, grid, want_trace != 0, wide, e->stream);
        } else {
            launch_fill_nonaffine(A, grid, want_trace != 0, wide, e->stream);
        }
        e->stats.kernel_launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(ev[1 + 3 * w], e->stream));
        if (want_trace) {
            TraceArgs T{};
            T.pairs = e->d_desc.p + b; T.npairs = (int)cnt; T.s = s; T.codes = e->d_codes.p; T.fmt = affine ? kernel : 2;
            T.codes_hi = SA.codes_hi;
            T.start_state = e->d_start.p; T.trace = e->d_trace.p; T.trace_len = e->d_tlen.p; T.complete = e->d_complete.p;
            if (na_ded) {
                T.fmt = 3; T.sysG = sysG;
            } else if (kernel == 1) {
                const SysGeo geo = sys_geo(s, plan.pad);
                T.sysG = sysG; T.R = geo.R; T.LPR = geo.LPR; T.P = geo.P; T.RING = geo.RING;
            }
            launch_traceback(T, e->stream);
            e->stats.kernel_launches++;
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(ev[2 + 3 * w], e->stream));
        CU(cudaEventRecord(ev[3 + 3 * w], e->stream));
    }
    lap("launches");
    CU(cudaStreamSynchronize(e->stream));
    lap("device");
    if (dbg_nts) {  // debug hook: "<tile> <start ns> <end ns>" per line, relative to the earliest start
        std::vector<unsigned long long> ts(dbg_nts);
        CU(cudaMemcpy(ts.data(), e->d_dbg_ts.p, dbg_nts * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (size_t q = 0; q < dbg_nts; q += 8) if (ts[q]) t0 = std::min(t0, ts[q]);
        if (FILE* f = fopen(getenv("BA_DEBUG_TS"), "w")) {  // tile, start, end, then when iterations 0, 50, 100, 200, 400, 800 were reached
            for (size_t q = 0; q < dbg_nts; q += 8) {
                fprintf(f, "%zu", q / 8);
                for (int k = 0; k < 8; ++k) fprintf(f, " %lld", ts[q + k] ? (long long)(ts[q + k] - t0) : -1ll);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    {
        float ms = 0;
        for (int w = 0; w < n_waves; ++w) {
            const cudaEvent_t prev = (w == 0) ? ev[0] : ev[3 * w];
            CU(cudaEventElapsedTime(&ms, prev, ev[1 + 3 * w]));
            e->stats.fill_ms += ms;
            CU(cudaEventElapsedTime(&ms, ev[1 + 3 * w], ev[2 + 3 * w]));
            e->stats.traceback_ms += ms;
        }
        CU(cudaEventElapsedTime(&ms, ev[0], ev[3 * n_waves]));
        e->stats.total_ms = ms;
    }
    if (want_trace) {
        for (int64_t q = wave_begin[n_waves - 1]; q < N; ++q) e->h_last_code_off[e->h_desc[q].orig] = e->h_desc[q].code_off;
        int64_t cb = 0;
        for (int64_t p = 0; p < N; ++p) cb += pair_code_words(ln[p], lm[p]) * unit;
        e->stats.code_bytes = cb;
    }
    if (rebase) {
        // pairs the rebased launch could not vouch for (walk stopped at a floor / out-of-band source, values above a row
        // maximum, score differing from the exact first launch): recompute them with the level kernel
        std::vector<uint8_t> sus((size_t)N), comp((size_t)N);
        CU(cudaMemcpyAsync(sus.data(), e->d_suspect.p, (size_t)N, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(comp.data(), e->d_complete.p, (size_t)N, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        std::vector<PairDesc> redo;
        for (int64_t q = 0; q < N; ++q)
            if (sus[e->h_desc[q].orig] || !comp[e->h_desc[q].orig]) redo.push_back(e->h_desc[q]);
        e->stats.fallback_pairs = (int32_t)std::min<size_t>(redo.size(), 0x7fffffff);
        if (!redo.empty()) {
            int64_t max_words = 0;
            int rn = 0;
            for (auto& d : redo) { max_words = std::max<int64_t>(max_words, code_words(d.n, d.m, s)); rn = std::max(rn, d.n); }
            if ((int64_t)e->d_codes.cap < max_words) {
                cudaError_t ce = e->d_codes.ensure((size_t)max_words);
                if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "traceback-code arena (level-kernel recomputation): " + std::string(cudaGetErrorString(ce)));
            }
            const size_t stride = generic_scratch_ints(rn, s);
            const int lgrid = (int)std::min<size_t>(redo.size(), (size_t)e->sm_count * 2);
            {
                cudaError_t ce = e->d_scratch.ensure(stride * lgrid);
                if (ce != cudaSuccess) return fail(e, BA_ERR_OOM, "fill scratch: " + std::string(cudaGetErrorString(ce)));
            }
            CU(e->d_desc2.ensure(redo.size()));
            cudaEvent_t r0, r1;
            CU(cudaEventCreate(&r0));
            CU(cudaEventCreate(&r1));
            CU(cudaEventRecord(r0, e->stream));
            size_t at = 0;
            while (at < redo.size()) {  // waves that fit the arena, cell-major code tables
                int64_t used = 0;
                size_t end = at;
                while (end < redo.size() && (end == at || used + code_words(redo[end].n, redo[end].m, s) <= (int64_t)e->d_codes.cap)) {
                    redo[end].code_off = used;
                    used += code_words(redo[end].n, redo[end].m, s);
                    ++end;
                }
                CU(cudaMemcpyAsync(e->d_desc2.p + at, redo.data() + at, sizeof(PairDesc) * (end - at), cudaMemcpyHostToDevice, e->stream));
                CU(cudaMemsetAsync(e->d_counter.p, 0, sizeof(int), e->stream));
                FillArgs A{};
                A.res = e->d_res.p; A.cls = e->d_cls.p; A.sim = e->d_sim.p; A.sc = e->sc;
                A.pairs = e->d_desc2.p + at; A.npairs = (int)(end - at); A.counter = e->d_counter.p;
                A.scratch = e->d_scratch.p; A.scratch_stride = stride;
                A.codes = e->d_codes.p;
                A.scores = e->d_scores.p; A.start_state = e->d_start.p; A.end_values = e->d_endv.p;
                launch_fill_generic(A, (int)std::min<size_t>(end - at, (size_t)lgrid), true, false, e->stream);
                CU(cudaGetLastError());
                TraceArgs T{};
                T.pairs = e->d_desc2.p + at; T.npairs = (int)(end - at); T.s = s; T.codes = e->d_codes.p; T.fmt = 0;
                T.start_state = e->d_start.p; T.trace = e->d_trace.p; T.trace_len = e->d_tlen.p; T.complete = e->d_complete.p;
                launch_traceback(T, e->stream);
                CU(cudaGetLastError());
                e->stats.kernel_launches += 2;
                at = end;
            }
            CU(cudaEventRecord(r1, e->stream));
            CU(cudaStreamSynchronize(e->stream));
            float ms = 0;
            cudaEventElapsedTime(&ms, r0, r1);
            cudaEventDestroy(r0);
            cudaEventDestroy(r1);
            e->stats.fill_ms += ms;
            e->stats.total_ms += ms;
            std::fill(e->h_last_code_off.begin(), e->h_last_code_off.end(), -1);  // the arena now holds the recomputed pairs' tables
        }
    }
    e->stats.waves = n_waves;
    e->stats.kernel_kind = kernel == 0 ? (wide ? 9 : 0) : na_ded ? 8 : chain_mode ? 10 : rebase ? (long_mode ? 12 : 11) : (p16 ? 5 : !affine ? (plan.pad ? 7 : 6) : (plan.pad ? 2 : 1) + (long_mode ? 2 : 0));
    e->stats.warps_per_cta = kernel == 0 ? 0 : sysG;
    e->last_fmt = kernel;
    e->last_sysG = kernel == 1 ? sysG : 0;
    e->last_na = na_ded;
    e->last_pad = kernel == 1 && plan.pad;
    e->last_hi_off = hi_plane_off;
    e->ran = true;
    e->ran_trace = want_trace != 0;
    return BA_OK;
}

int ba_fetch_scores(ba_engine* e, int64_t* scores) {
    if (!e) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::fetch_scores(e, scores);
    if (!e->ran) return fail(e, BA_ERR_STATE, "ba_run first");
    if (e->n_pairs == 0) return BA_OK;
    if (!scores) return fail(e, BA_ERR_INVALID_ARG, "scores is NULL");
    CU(cudaSetDevice(e->device));
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    CU(cudaMemcpyAsync(scores, e->d_scores.p, sizeof(int64_t) * e->n_pairs, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BA_OK;
}

static int fetch_lens(ba_engine* e) {
    if (e->have_tlen) return BA_OK;
    e->h_tlen.resize((size_t)e->n_pairs);
    if (e->n_pairs) {
        CU(cudaMemcpyAsync(e->h_tlen.data(), e->d_tlen.p, sizeof(int32_t) * e->n_pairs, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    e->have_tlen = true;
    return BA_OK;
}

// D2H of all trace slots into the engine's pinned staging buffer (pair p's columns end at
// h_stage + h_slot_off[p] + h_slot_cap[p]) and of the completeness flags into `complete` (engine order).
static int stage_traces(ba_engine* e, uint8_t* complete) {
    const int64_t N = e->n_pairs;
    if (N == 0) return BA_OK;
    const size_t slot_bytes = (size_t)e->h_slot_off[N - 1] + ((e->h_slot_cap[N - 1] + 15) & ~15);
    CU(e->h_stage.ensure(slot_bytes));
    CU(cudaMemcpyAsync(e->h_stage.p, e->d_trace.p, slot_bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(complete, e->d_complete.p, (size_t)N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BA_OK;
}

int ba_trace_bytes(ba_engine* e, int64_t* total) {
    if (!e || !total) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::trace_bytes(e, total);
    if (!e->ran || !e->ran_trace) return fail(e, BA_ERR_STATE, "ba_run(want_trace=1) first");
    CU(cudaSetDevice(e->device));
    int rc = fetch_lens(e);
    if (rc) return rc;
    int64_t t = 0;
    for (int32_t x : e->h_tlen) t += x;
    *total = t;
    return BA_OK;
}

int ba_fetch_traces(ba_engine* e, uint8_t* cols, int64_t* offsets, uint8_t* complete) {
    if (!e) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) return multi::fetch_traces(e, cols, offsets, complete);
    if (!e->ran || !e->ran_trace) return fail(e, BA_ERR_STATE, "ba_run(want_trace=1) first");
    if (!offsets) return fail(e, BA_ERR_INVALID_ARG, "offsets is NULL");
    CU(cudaSetDevice(e->device));
    int rc = fetch_lens(e);
    if (rc) return rc;
    const int64_t N = e->n_pairs;
    offsets[0] = 0;
    if (N == 0) return BA_OK;
    if (!complete) return fail(e, BA_ERR_INVALID_ARG, "complete is NULL");
    rc = stage_traces(e, complete);
    if (rc) return rc;
    int64_t pos = 0;
    for (int64_t p = 0; p < N; ++p) {
        const int32_t len = e->h_tlen[p];
        if (len && !cols) return fail(e, BA_ERR_INVALID_ARG, "cols is NULL");
        if (len) memcpy(cols + pos, e->h_stage.p + e->h_slot_off[p] + e->h_slot_cap[p] - len, (size_t)len);
        pos += len;
        offsets[p + 1] = pos;
    }
    return BA_OK;
}

int ba_align_batch(ba_engine* e, const uint8_t* residues, const uint8_t* classes, const int64_t* offsets,
                   int64_t n_seq, const int32_t* seq_a, const int32_t* seq_b, int64_t n_pairs, int want_trace,
                   int64_t* scores) {
    int rc;
    if ((rc = ba_load_sequences(e, residues, classes, offsets, n_seq))) return rc;
    if ((rc = ba_load_pairs(e, seq_a, seq_b, n_pairs))) return rc;
    if ((rc = ba_run(e, want_trace))) return rc;
    return ba_fetch_scores(e, scores);
}

int ba_get_stats(const ba_engine* e, ba_stats* out) {
    if (!e || !out) return BA_ERR_INVALID_ARG;
    *out = e->stats;
    return BA_OK;
}

int ba_debug_fetch_codes(ba_engine* e, int64_t pair, uint64_t* out, int64_t words) {
    if (!e || !out) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) {
        ba_engine* kid = nullptr;
        int64_t local = 0;
        int rc = multi::debug_route(e, pair, &kid, &local);
        if (rc) return rc;
        rc = ba_debug_fetch_codes(kid, local, out, words);
        if (rc) e->err = kid->err;
        return rc;
    }
    if (!e->ran || !e->ran_trace) return fail(e, BA_ERR_STATE, "ba_run(want_trace=1) first");
    if (pair < 0 || pair >= e->n_pairs || e->h_last_code_off[pair] < 0)
        return fail(e, BA_ERR_INVALID_ARG, "pair not in the last wave");
    CU(cudaSetDevice(e->device));
    const int n = (int)(e->h_off[e->h_pa[pair] + 1] - e->h_off[e->h_pa[pair]]);
    const int m = (int)(e->h_off[e->h_pb[pair] + 1] - e->h_off[e->h_pb[pair]]);
    const int S = e->sc.s;
    const int64_t need = code_words(n, m, S);
    if (words < need) return fail(e, BA_ERR_INVALID_ARG, "buffer too small");
    if (e->last_sysG == 0) {
        CU(cudaMemcpyAsync(out, e->d_codes.p + e->h_last_code_off[pair], (size_t)need * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        return BA_OK;
    }
    if (e->last_na) {  // dedicated non-affine kernel: a nibble per cell -> the cell-major table, case index in the low nibble
        const int G = e->last_sysG, nit_all = na_iters(S, G, m) + na_pre(S);
        const int64_t raw = na_code_words(S, G, n, m);
        std::vector<uint64_t> tmp((size_t)raw);
        CU(cudaMemcpyAsync(tmp.data(), e->d_codes.p + e->h_last_code_off[pair], (size_t)raw * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(tmp.data());
        for (int i = 0; i <= n; ++i)
            for (int a = -S; a <= S; ++a)
                for (int j = 0; j <= m; ++j)
                    for (int b = -S; b <= S; ++b)
                        out[code_index(m, S, i, j, a, b)] = (w32[na_code_index(S, G, nit_all, i, j, b)] >> (4 * (a + S))) & 15;
        return BA_OK;
    }
    // systolic layout -> the cell-major table this hook promises (cells the kernel never visits read as all-ones)
    const SysGeo geo = sys_geo(S, e->last_pad);
    const int G = e->last_sysG, nit_all = sys_iters(S, e->last_pad, G, m) + 4;
    const int64_t raw = sys_code_words(S, e->last_pad, G, n, m);
    // slots live in two planes: the low word and the upper half of the high word (the fill's 64-bit code word is rebuilt here)
    std::vector<uint32_t> lo((size_t)raw);
    std::vector<uint16_t> hi((size_t)raw);
    const uint32_t* plane_lo = reinterpret_cast<const uint32_t*>(e->d_codes.p);
    const uint16_t* plane_hi = reinterpret_cast<const uint16_t*>(plane_lo + e->last_hi_off);
    CU(cudaMemcpyAsync(lo.data(), plane_lo + e->h_last_code_off[pair], (size_t)raw * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(hi.data(), plane_hi + e->h_last_code_off[pair], (size_t)raw * 2, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (int i = 0; i <= n; ++i)
        for (int a = -S; a <= S; ++a)
            for (int j = 0; j <= m; ++j)
                for (int b = -S; b <= S; ++b) {
                    const size_t q = (size_t)sys_code_index(geo.R, geo.LPR, geo.P, S, G, nit_all, i, j, a, b);
                    out[code_index(m, S, i, j, a, b)] = (uint64_t)lo[q] | ((uint64_t)hi[q] << 48);
                }
    return BA_OK;
}

int ba_debug_fetch_end_values(ba_engine* e, int64_t pair, int32_t* out9) {
    if (!e || !out9) return BA_ERR_INVALID_ARG;
    if (!e->kids.empty()) {
        ba_engine* kid = nullptr;
        int64_t local = 0;
        int rc = multi::debug_route(e, pair, &kid, &local);
        if (rc) return rc;
        rc = ba_debug_fetch_end_values(kid, local, out9);
        if (rc) e->err = kid->err;
        return rc;
    }
    if (!e->ran) return fail(e, BA_ERR_STATE, "ba_run first");
    if (pair < 0 || pair >= e->n_pairs) return fail(e, BA_ERR_INVALID_ARG, "pair out of range");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(out9, e->d_endv.p + pair * 9, sizeof(int) * 9, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BA_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Multi-GPU front.  Pairs are independent units, so the whole box is driven from one process without any
// collective: the pair list is dealt over the devices by cost (longest first; a snake deal for large batches,
// exact LPT for small ones -- the same policy as bialign_b200/batch.py lpt_shards), every device gets the
// sequence table, one host thread per device runs the ordinary single-device engine, and the results are
// written into the caller's arrays in caller order.
// ------------------------------------------------------------------------------------------------
namespace multi {

template <class F>
int for_each_kid(ba_engine* e, F f) {  // one host thread per device; the first error wins
    const size_t K = e->kids.size();
    std::vector<int> rc(K, BA_OK);
    std::vector<std::thread> th;
    th.reserve(K);
    for (size_t k = 0; k < K; ++k) th.emplace_back([&, k] { rc[k] = f(k, e->kids[k]); });
    for (auto& t : th) t.join();
    for (size_t k = 0; k < K; ++k)
        if (rc[k]) {
            e->err = "device " + std::to_string(e->kids[k]->device) + ": " + e->kids[k]->err;
            return rc[k];
        }
    return BA_OK;
}

void destroy(ba_engine* e) {
    for (ba_engine* k : e->kids) ba_engine_destroy(k);
    e->kids.clear();
    delete e;
}

int set_option(ba_engine* e, const char* key, int64_t value) {
    for (ba_engine* k : e->kids) {
        const int rc = ba_set_option(k, key, value);
        if (rc) { e->err = k->err; return rc; }
    }
    return BA_OK;
}

int set_scoring(ba_engine* e, const int32_t* sim, int nsym, int w, int beta, int gamma, int delta, int s) {
    const int rc = for_each_kid(e, [&](size_t, ba_engine* k) { return ba_set_scoring(k, sim, nsym, w, beta, gamma, delta, s); });
    if (rc) return rc;
    e->sc = e->kids[0]->sc;
    e->have_scoring = true;
    e->ran = false;
    return BA_OK;
}

int load_sequences(ba_engine* e, const uint8_t* residues, const uint8_t* classes, const int64_t* offsets, int64_t n_seq) {
    const int rc = for_each_kid(e, [&](size_t, ba_engine* k) { return ba_load_sequences(k, residues, classes, offsets, n_seq); });
    if (rc) return rc;
    e->h_off = e->kids[0]->h_off;
    e->n_seq = n_seq;
    e->have_seqs = true;
    e->have_pairs = false;
    e->ran = false;
    return BA_OK;
}

int load_pairs(ba_engine* e, const int32_t* seq_a, const int32_t* seq_b, int64_t n_pairs) {
    if (!e->have_seqs) return fail(e, BA_ERR_STATE, "ba_load_sequences first");
    if (n_pairs < 0 || (n_pairs > 0 && (!seq_a || !seq_b))) return fail(e, BA_ERR_INVALID_ARG, "pair arrays NULL");
    if (n_pairs > 0x7fffffff) return fail(e, BA_ERR_INVALID_ARG, "too many pairs for one call");
    for (int64_t p = 0; p < n_pairs; ++p)
        if (seq_a[p] < 0 || seq_a[p] >= e->n_seq || seq_b[p] < 0 || seq_b[p] >= e->n_seq)
            return fail(e, BA_ERR_INVALID_ARG, "pair " + std::to_string(p) + " references a sequence outside the table");
    const size_t K = e->kids.size();
    std::vector<int64_t> cost((size_t)n_pairs);
    for (int64_t p = 0; p < n_pairs; ++p)
        cost[p] = (e->h_off[seq_a[p] + 1] - e->h_off[seq_a[p]] + 1) * (e->h_off[seq_b[p] + 1] - e->h_off[seq_b[p]] + 1);
    std::vector<int32_t> order((size_t)n_pairs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cost[x] > cost[y]; });
    e->shard.assign(K, {});
    if (n_pairs > 4096) {  // snake deal of the sorted list
        for (int64_t q = 0; q < n_pairs; ++q) {
            const int64_t rnd = q / (int64_t)K, pos = q % (int64_t)K;
            e->shard[(rnd & 1) ? K - 1 - pos : pos].push_back(order[q]);
        }
    } else {  // exact LPT: next pair to the least loaded device
        std::vector<int64_t> load(K, 0);
        for (int64_t q = 0; q < n_pairs; ++q) {
            const size_t k = std::min_element(load.begin(), load.end()) - load.begin();
            e->shard[k].push_back(order[q]);
            load[k] += cost[order[q]];
        }
    }
    for (auto& sh : e->shard) std::sort(sh.begin(), sh.end());
    const int rc = for_each_kid(e, [&](size_t k, ba_engine* kid) {
        const auto& sh = e->shard[k];
        std::vector<int32_t> a(sh.size()), b(sh.size());
        for (size_t q = 0; q < sh.size(); ++q) { a[q] = seq_a[sh[q]]; b[q] = seq_b[sh[q]]; }
        return ba_load_pairs(kid, a.data(), b.data(), (int64_t)sh.size());
    });
    if (rc) return rc;
    e->n_pairs = n_pairs;
    e->have_pairs = true;
    e->ran = false;
    return BA_OK;
}

int set_pair_mu2(ba_engine* e, const int32_t* mu2, const int64_t* offsets) {
    if (!e->have_pairs) return fail(e, BA_ERR_STATE, "ba_load_pairs first");
    if (!mu2 && !offsets) return for_each_kid(e, [&](size_t, ba_engine* kid) { return ba_set_pair_mu2(kid, nullptr, nullptr); });
    if (!mu2 || !offsets) return fail(e, BA_ERR_INVALID_ARG, "mu2 / offsets NULL");
    return for_each_kid(e, [&](size_t k, ba_engine* kid) {  // every device gets the matrices of its own pairs
        const auto& sh = e->shard[k];
        std::vector<int64_t> off(sh.size() + 1, 0);
        for (size_t q = 0; q < sh.size(); ++q) off[q + 1] = off[q] + (offsets[sh[q] + 1] - offsets[sh[q]]);
        std::vector<int32_t> buf((size_t)std::max<int64_t>(off.back(), 1));
        for (size_t q = 0; q < sh.size(); ++q)
            std::copy(mu2 + offsets[sh[q]], mu2 + offsets[sh[q] + 1], buf.begin() + off[q]);
        return ba_set_pair_mu2(kid, buf.data(), off.data());
    });
}

int run(ba_engine* e, int want_trace) {
    if (!e->have_scoring) return fail(e, BA_ERR_STATE, "ba_set_scoring first");
    if (!e->have_pairs) return fail(e, BA_ERR_STATE, "ba_load_sequences and ba_load_pairs first");
    e->ran = false;
    e->m_tlen.clear();
    const int rc = for_each_kid(e, [&](size_t, ba_engine* k) { return ba_run(k, want_trace); });
    if (rc) return rc;
    ba_stats st{};
    st.device = e->kids[0]->device;
    st.kernel_kind = e->kids[0]->stats.kernel_kind;
    st.warps_per_cta = e->kids[0]->stats.warps_per_cta;
    for (ba_engine* k : e->kids) {  // sums of work, max of device times (the devices run side by side)
        st.pairs += k->stats.pairs;
        st.cell_states += k->stats.cell_states;
        st.kernel_launches += k->stats.kernel_launches;
        st.code_bytes += k->stats.code_bytes;
        st.fallback_pairs += k->stats.fallback_pairs;
        st.waves = std::max(st.waves, k->stats.waves);
        st.fill_ms = std::max(st.fill_ms, k->stats.fill_ms);
        st.traceback_ms = std::max(st.traceback_ms, k->stats.traceback_ms);
        st.total_ms = std::max(st.total_ms, k->stats.total_ms);
    }
    e->stats = st;
    e->ran = true;
    e->ran_trace = want_trace != 0;
    return BA_OK;
}

int fetch_scores(ba_engine* e, int64_t* scores) {
    if (!e->ran) return fail(e, BA_ERR_STATE, "ba_run first");
    if (e->n_pairs == 0) return BA_OK;
    if (!scores) return fail(e, BA_ERR_INVALID_ARG, "scores is NULL");
    return for_each_kid(e, [&](size_t k, ba_engine* kid) {
        const auto& sh = e->shard[k];
        std::vector<int64_t> tmp(sh.size());
        const int rc = ba_fetch_scores(kid, tmp.data());
        if (rc) return rc;
        for (size_t q = 0; q < sh.size(); ++q) scores[sh[q]] = tmp[q];
        return (int)BA_OK;
    });
}

static int gather_lens(ba_engine* e) {
    if (!e->m_tlen.empty() || e->n_pairs == 0) return BA_OK;
    std::vector<int64_t> tl((size_t)e->n_pairs, 0);
    const int rc = for_each_kid(e, [&](size_t k, ba_engine* kid) {
        cudaSetDevice(kid->device);
        const int r = fetch_lens(kid);
        if (r) return r;
        const auto& sh = e->shard[k];
        for (size_t q = 0; q < sh.size(); ++q) tl[sh[q]] = kid->h_tlen[q];
        return (int)BA_OK;
    });
    if (rc) return rc;
    e->m_tlen.swap(tl);
    return BA_OK;
}

int trace_bytes(ba_engine* e, int64_t* total) {
    if (!e->ran || !e->ran_trace) return fail(e, BA_ERR_STATE, "ba_run(want_trace=1) first");
    const int rc = gather_lens(e);
    if (rc) return rc;
    int64_t t = 0;
    for (int64_t x : e->m_tlen) t += x;
    *total = t;
    return BA_OK;
}

int fetch_traces(ba_engine* e, uint8_t* cols, int64_t* offsets, uint8_t* complete) {
    if (!e->ran || !e->ran_trace) return fail(e, BA_ERR_STATE, "ba_run(want_trace=1) first");
    if (!offsets) return fail(e, BA_ERR_INVALID_ARG, "offsets is NULL");
    int rc = gather_lens(e);
    if (rc) return rc;
    const int64_t N = e->n_pairs;
    offsets[0] = 0;
    if (N == 0) return BA_OK;
    if (!complete) return fail(e, BA_ERR_INVALID_ARG, "complete is NULL");
    for (int64_t p = 0; p < N; ++p) offsets[p + 1] = offsets[p] + e->m_tlen[p];
    if (offsets[N] && !cols) return fail(e, BA_ERR_INVALID_ARG, "cols is NULL");
    // every device copies its trace slots to its own pinned staging buffer, then its host thread places each
    // pair's columns directly at the pair's position in the caller's array
    return for_each_kid(e, [&](size_t k, ba_engine* kid) {
        const auto& sh = e->shard[k];
        std::vector<uint8_t> comp(sh.size() + 1);
        cudaSetDevice(kid->device);
        const int r = stage_traces(kid, comp.data());
        if (r) return r;
        for (size_t q = 0; q < sh.size(); ++q) {
            const int32_t len = kid->h_tlen[q];
            if (len) memcpy(cols + offsets[sh[q]], kid->h_stage.p + kid->h_slot_off[q] + kid->h_slot_cap[q] - len, (size_t)len);
            complete[sh[q]] = comp[q];
        }
        return (int)BA_OK;
    });
}

int debug_route(ba_engine* e, int64_t pair, ba_engine** kid, int64_t* local) {
    if (!e->ran) return fail(e, BA_ERR_STATE, "ba_run first");
    for (size_t k = 0; k < e->kids.size(); ++k) {
        const auto& sh = e->shard[k];
        const auto it = std::lower_bound(sh.begin(), sh.end(), (int32_t)pair);
        if (it != sh.end() && *it == pair) {
            *kid = e->kids[k];
            *local = it - sh.begin();
            return BA_OK;
        }
    }
    return fail(e, BA_ERR_INVALID_ARG, "pair out of range");
}

}  // namespace multi

extern "C" int ba_engine_create_multi(const int* devices, int n_devices, ba_engine** out) {
    if (!out) return fail(nullptr, BA_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (n_devices == 0 || !devices) {  // all visible devices
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            return fail(nullptr, BA_ERR_NO_DEVICE, "no CUDA device (bialign_b200 has no CPU path)");
    } else if (n_devices < 0) {
        return fail(nullptr, BA_ERR_INVALID_ARG, "n_devices < 0");
    }
    const int K = ndev ? ndev : n_devices;
    for (int k = 0; k < K; ++k)
        for (int q = 0; q < k; ++q)
            if (!ndev && devices[k] == devices[q]) return fail(nullptr, BA_ERR_INVALID_ARG, "device listed twice");
    ba_engine* e = new ba_engine();
    for (int k = 0; k < K; ++k) {
        ba_engine* kid = nullptr;
        const int rc = ba_engine_create(ndev ? k : devices[k], &kid);
        if (rc) {  // g_create_error holds the text
            for (ba_engine* x : e->kids) ba_engine_destroy(x);
            delete e;
            return rc;
        }
        e->kids.push_back(kid);
    }
    e->device = e->kids[0]->device;
    e->sm_count = e->kids[0]->sm_count;
    e->stats.device = e->device;
    *out = e;
    return BA_OK;
}

extern "C" int ba_engine_device_count(const ba_engine* e) { return e ? (e->kids.empty() ? 1 : (int)e->kids.size()) : 0; }


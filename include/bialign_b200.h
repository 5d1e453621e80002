/*
 * bialign_b200.h -- C ABI of the B200-native bi-alignment engine.
 *
 * The reference (s-will/BiAlign) has no FFI: its hot path lives inside the Cython extension class
 * `bialignment.BiAligner` (src/bialignment.pyx:155-586).  This header is the boundary a maintainer
 * would bind instead of that class' two DP methods; every entry point names the reference code it
 * replaces.  Plain pointers and sizes only; no torch / numpy / CUDA types.  All buffers passed in
 * are owned by the caller and never freed or retained past the call (except where stated); the
 * library owns all device memory.  Calls on one engine must be serialised by the caller; the
 * Python binding (ctypes) releases the GIL for the duration of every call.
 *
 * There is NO CPU path: ba_engine_create fails with BA_ERR_NO_DEVICE when no sm_100 device is
 * usable, and nothing else works without an engine.
 *
 * Encodings (mirrored by oracle/bialign_oracle.c):
 *   residues  uint8 codes < nsym; mu1(i,j) = sim[res_a[i-1]*nsym + res_b[j-1]]   (pyx:405-412)
 *   classes   uint8 structure symbols; mu2(k,l) = (cls_a[k-1]==cls_b[l-1]) ? w : 0 (pyx:414-429;
 *             for RNA with supplied dot-bracket strings the class is up/down/unpaired, pyx:366-392)
 *   column    8*x0 + 4*x1 + 2*x2 + x3 in 1..15, x = which of A1,B1,A2,B2 advance (pyx:526, 568)
 *   states    index 0..8 = 0101 0110 0111 1001 1010 1011 1101 1110 1111          (pyx:61-65)
 */
#ifndef BIALIGN_B200_H
#define BIALIGN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BA_OK 0
#define BA_ERR_INVALID_ARG 1 /* NULL pointer, negative size, max_shift out of range, bad pair index */
#define BA_ERR_NO_DEVICE 2   /* no CUDA device / not sm_100: the product has no CPU fallback       */
#define BA_ERR_CUDA 3        /* a CUDA call failed; ba_last_error has the text                    */
#define BA_ERR_OOM 4         /* device (or pinned host) allocation failed                         */
#define BA_ERR_SCORE_RANGE 5 /* a fast kernel was forced ("kernel" = 1) although its integer range conditions do not hold,
                                or the score bound exceeds even int64.  In automatic mode batches whose bound
                                (n+m)*(max|mu1|+|w|+2|beta|+2|gamma|+2|Delta|) reaches 2^30 run on the 64-bit
                                instantiation of the general level kernel (the reference's int64 tables, pyx:27-35) */
#define BA_ERR_ALPHABET 6    /* a residue code >= nsym (the reference raises KeyError, pyx:407)    */
#define BA_ERR_STATE 7       /* call order: scoring / sequences / pairs not set, nothing run yet   */

#if defined(__GNUC__)
#define BA_API __attribute__((visibility("default")))
#else
#define BA_API
#endif

#define BA_MAX_SHIFT 16         /* largest band half-width accepted (general level kernel)          */
#define BA_SYSTOLIC_MAX_SHIFT 4 /* the fast systolic kernels are instantiated for max_shift 0..4    */
#define BA_MAX_SEQ_LEN (1 << 24) /* longest single sequence accepted by ba_load_sequences             */

typedef struct ba_engine ba_engine;

typedef struct ba_stats {
    int64_t pairs;            /* pairs processed by the last ba_run                              */
    int64_t cell_states;      /* sum over pairs of 9*C(n,m,s) (affine) or C(n,m,s) (non-affine)  */
    int64_t kernel_launches;  /* CUDA kernels launched by the last ba_run                        */
    int64_t waves;            /* traceback-memory waves                                          */
    double fill_ms;           /* device time of the fill kernels (CUDA events)                   */
    double traceback_ms;      /* device time of the traceback kernels                            */
    double total_ms;          /* device time of the whole ba_run                                 */
    int64_t code_bytes;       /* traceback-code bytes written to HBM                             */
    int32_t kernel_kind;      /* 0 = generic level kernel, 1 = systolic pad-free, 2 = systolic padded,
                                 3 / 4 = the same two in multi-CTA long-pair mode,
                                 5 = systolic pad-free, two pairs per lane in packed 16-bit halves,
                                 6 / 7 = systolic pad-free / padded running the non-affine model,
                                 8 = dedicated non-affine kernel (a lane owns a row and all its band offsets),
                                 9 = general level kernel, 64-bit values (score bound >= 2^30),
                                 10 = systolic pad-free, short pairs chained along j (every pair fits one row block),
                                 11 / 12 = systolic pad-free, rebased wide-range trace run (batch / long-pair mode): scores
                                 whose packed form (value << tie bits) exceeds 32 bits; a score-only launch records
                                 row maxima, the trace launch works relative to them */
    int32_t device;
    int32_t warps_per_cta;    /* CTA width the systolic kernel ran with (0 for the general kernel)      */
    int32_t fallback_pairs;   /* rebased runs: pairs the fast kernel could not vouch for, recomputed by the level kernel */
} ba_stats;

/* One engine per process per GPU (device = CUDA ordinal, normally LOCAL_RANK). */
BA_API int ba_engine_create(int device, ba_engine** out);
/* One engine for several GPUs of the box, driven from this one process (replaces nothing in the
 * reference, which is single-threaded: pairs are independent, so the pair list is sharded).  devices = n_devices
 * CUDA ordinals, or NULL / 0 for all visible devices.  The returned handle takes every call below:
 * ba_load_pairs deals the pairs over the devices by cost (longest first), ba_load_sequences uploads the sequence
 * table to each, ba_run drives one host thread per device, and the fetch calls write scores / traces straight into the
 * caller's arrays in the caller's pair order.  No collective and no peer traffic: results travel device -> host only. */
BA_API int ba_engine_create_multi(const int* devices, int n_devices, ba_engine** out);
/* Number of devices behind a handle (1 for ba_engine_create). */
BA_API int ba_engine_device_count(const ba_engine* e);
BA_API void ba_engine_destroy(ba_engine* e);
/* Text of the last error on this engine (or of the last failed ba_engine_create when e == NULL). */
BA_API const char* ba_last_error(const ba_engine* e);

/* Scoring model.  Replaces BiAligner.__init__'s parameter capture (pyx:179-193) and the per-cell
 * mu1/mu2/affine_score evaluation (pyx:84-131, 405-440).  sim is nsym*nsym int32, copied.
 * gap_opening_cost != 0 selects the affine model (pyx:203-205, 474-509, 535-586), == 0 the
 * non-affine one (pyx:225-252, 443-471, 513-531). */
BA_API int ba_set_scoring(ba_engine* e, const int32_t* sim, int nsym, int structure_weight, int gap_opening_cost,
                   int gap_cost, int shift_cost, int max_shift);

/* Sequence table: n_seq sequences concatenated; sequence q is [offsets[q], offsets[q+1]).
 * Copied to the device (one H2D each).  Replaces _preprocess_seq's per-object storage (pyx:340-376). */
BA_API int ba_load_sequences(ba_engine* e, const uint8_t* residues, const uint8_t* classes, const int64_t* offsets,
                      int64_t n_seq);
/* Pair list: pair p aligns sequence seq_a[p] (molecule A) with seq_b[p] (molecule B). */
BA_API int ba_load_pairs(ba_engine* e, const int32_t* seq_a, const int32_t* seq_b, int64_t n_pairs);

/* Optional per-pair structure-similarity matrices: pair p of the loaded pair list (caller order) owns len(A) x len(B)
 * int32 entries at mu2 + offsets[p], entry (k-1)*len(B) + (l-1) = mu2(k, l).  Replaces the probabilistic branch of
 * _structure_similarity (pyx:414-423: int(w * (sqrt(upA*upB) + sqrt(downA*downB) + sqrt(unpA*unpB))) for base-pair
 * probability profiles, e.g. ViennaRNA's when no structure is supplied, pyx:345-353): the host evaluates that formula, the
 * device reads the integers.  Pairs with matrices run on the general level kernel (small batches; not the fast path).
 * ba_load_pairs drops the matrices; ba_set_pair_mu2(e, NULL, NULL) returns to class-equality scoring. */
BA_API int ba_set_pair_mu2(ba_engine* e, const int32_t* mu2, const int64_t* offsets);

/* Forward fill (+ traceback when want_trace) of every loaded pair on device-resident inputs.
 * Replaces optimize() (pyx:443-509) and traceback() (pyx:513-586).  Blocks until the device is done. */
BA_API int ba_run(ba_engine* e, int want_trace);

/* scores[p] = optimize() of pair p (the reference returns numpy.int64, pyx:509). */
BA_API int ba_fetch_scores(ba_engine* e, int64_t* scores);
/* Bytes needed for ba_fetch_traces' cols buffer after a ba_run(want_trace=1). */
BA_API int ba_trace_bytes(ba_engine* e, int64_t* total);
/* cols: one byte per alignment column, forward order (what traceback() returns, pyx:531/586);
 * pair p owns [offsets[p], offsets[p+1]); complete[p] == 0 reproduces the reference's
 * "WARNING: incomplete traceback" condition (pyx:584-585). */
BA_API int ba_fetch_traces(ba_engine* e, uint8_t* cols, int64_t* offsets, uint8_t* complete);

/* Host-buffer convenience = load_sequences + load_pairs + run + fetch (the end-to-end call). */
BA_API int ba_align_batch(ba_engine* e, const uint8_t* residues, const uint8_t* classes, const int64_t* offsets,
                   int64_t n_seq, const int32_t* seq_a, const int32_t* seq_b, int64_t n_pairs, int want_trace,
                   int64_t* scores);

BA_API int ba_get_stats(const ba_engine* e, ba_stats* out);
/* Tuning knobs: "code_arena_bytes" (traceback-code arena, default: 1/2 of free HBM),
 * "kernel" (0 generic, 1 systolic, -1 auto), "pad" (systolic flavour: 0 pad-free, 1 padded, -1 auto),
 * "warps_per_cta" (1..8, 0 = chosen per batch), "long" (multi-CTA long-pair mode: 0 off, 1 force, -1 auto),
 * "io_warp" (long-pair mode: one more warp per CTA that owns the boundary I/O and the progress flags: 0 off, 1 on,
 * -1 = for a handful of long pairs), "col_chunks" (a single long pair with the I/O warp: column chunks per row block, tiles dealt
 * claimed by readiness: 0 / 1 = none (the default: measured slower than plain row blocks), 2..64 = that many),
 * "p16" (16-bit pair mode for score-only batches: 0 off, 1 force, -1 auto),
 * "na_kernel" (non-affine model: 0 = systolic kernel's non-affine flavour, 1 / -1 = dedicated kernel when applicable),
 * "chain" (batches of short pairs run as chains through the systolic array: 0 off, 1 force, -1 auto),
 * "rebase" (trace runs relative to row maxima: -1 = when the packed 32-bit plan fails, 0 off, 1 force),
 * "rebase_window" (test hook: cap on the rebased value window in scaled score units; 0 = the full window). */
BA_API int ba_set_option(ba_engine* e, const char* key, int64_t value);

/* Test hook: copy the raw 4-bit code table of pair p of the last wave (uint64 per cell, index
 * ((i*W + a+s)*(m+1) + j)*W + b+s, nibble t = case id of state t, 15 = none).  words = capacity. */
BA_API int ba_debug_fetch_codes(ba_engine* e, int64_t pair, uint64_t* out, int64_t words);
/* Test hook: the nine end values M[t][n,m,n,m] of pair p. */
BA_API int ba_debug_fetch_end_values(ba_engine* e, int64_t pair, int32_t* out9);

/* Roofline denominator: measured thread-instructions per second of the integer pipes on `device`
 * (kind 0: add, 1: xor+max pairs, 2: fused add+max VIADDMNMX).  Used by bench.py only. */
BA_API int ba_microbench_int(int device, int kind, double* instr_per_s, int* sm_count);

BA_API const char* ba_version(void);

#ifdef __cplusplus
}
#endif
#endif

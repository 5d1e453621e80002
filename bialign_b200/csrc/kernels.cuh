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
    size_t scratch_stride;    // ints per CTA
    uint64_t* codes;          // traceback-code arena (nullptr when score-only)
    long long* scores;        // [n_pairs] in caller order
    uint8_t* start_state;     // [n_pairs] traceback start state
    int* end_values;          // [n_pairs][9]  M[t][n,m,n,m]
};

struct TraceArgs {
    const PairDesc* pairs;
    int npairs;
    int s;
    const uint64_t* codes;
    const uint8_t* start_state;
    uint8_t* trace;           // slots; columns are written backwards from the end of each slot
    int* trace_len;           // [n_pairs] caller order
    uint8_t* complete;        // [n_pairs]
};

void launch_fill_generic(const FillArgs& A, int grid, bool trace, cudaStream_t st);
size_t generic_scratch_ints(int nmax, int s);
void launch_traceback(const TraceArgs& A, cudaStream_t st);

}  // namespace ba

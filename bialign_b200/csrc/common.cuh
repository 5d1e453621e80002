// Shared definitions of the bi-alignment device code.
//
// Recurrence and tie-breaking follow the reference's behaviour (src/bialignment.pyx, "pyx"):
// states pyx:61-65, column score pyx:84-131, band guard pyx:133-141, case order pyx:255-296,
// fill pyx:474-509, traceback tie rule pyx:541-564.  Nothing here is translated from it; the
// formulation (half-state ranks, push/pull factorisation, 4-bit codes) is this project's own.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ba {

constexpr int NEG = -(1 << 30);  // pyx:303, pyx:484

// A half state (what the last column of one of the two coupled alignments was) is stored as a
// rank r in {0,1,2} = {01, 10, 11} = {gap over B-residue, A-residue over gap, match}; its two
// bits are (r+1)>>1 and (r+1)&1.  A state index t in 0..8 is 3*r01 + r23, which is exactly the
// reference's itertools.product order 0101 0110 0111 1001 1010 1011 1101 1110 1111.
__host__ __device__ __forceinline__ constexpr int hb0(int r) { return (r + 1) >> 1; }
__host__ __device__ __forceinline__ constexpr int hb1(int r) { return (r + 1) & 1; }

struct Scoring {
    int w, beta, gamma, delta;  // structure_weight, gap_opening_cost, gap_cost, shift_cost
    int s;                      // max_shift
    int nsym;                   // alphabet size of the similarity table
};

// One pair of a wave, as the kernels see it.
struct PairDesc {
    long long offA, offB;   // start of molecule A / B in the residue and class arrays
    int n, m;               // lengths
    long long code_off;     // first code word of this pair in the arena (uint64 units)
    long long trace_off;    // first byte of this pair's trace slot
    int trace_cap;          // slot size in bytes (2(n+m)+2)
    int orig;               // index of the pair in the caller's order
};

// Code-table index of cell (i, j, a = k-i, b = l-j); one uint64 per cell, nibble t = case id.
__host__ __device__ __forceinline__ long long code_index(int m, int s, int i, int j, int a, int b) {
    const int W = 2 * s + 1;
    return (((long long)i * W + (a + s)) * (m + 1) + j) * W + (b + s);
}
__host__ __device__ __forceinline__ long long code_words(int n, int m, int s) {
    const int W = 2 * s + 1;
    return (long long)(n + 1) * W * (m + 1) * W;
}

// Decode case id (0..14) at target state t into the column x (bit 3 = x0 .. bit 0 = x3) and the
// source state index.  ids 0-8: full column, source = id; 9-11: x = (0,0,t2,t3), source half
// (t01, h) with h in order 11,10,01; 12-14: x = (t0,t1,0,0), source (h, t23)  (pyx:275-296).
__host__ __device__ __forceinline__ void decode_case(int t, int id, int& xbits, int& src) {
    const int r01 = t / 3, r23 = t % 3;
    const int x01 = r01 + 1, x23 = r23 + 1;  // two-bit column halves
    if (id < 9) {
        xbits = (x01 << 2) | x23;
        src = id;
    } else if (id < 12) {
        xbits = x23;
        src = 3 * r01 + (2 - (id - 9));
    } else {
        xbits = x01 << 2;
        src = 3 * (2 - (id - 12)) + r23;
    }
}

}  // namespace ba

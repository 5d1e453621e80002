// Run-time dispatch over max_shift for the systolic fill, plus its host-side geometry
// (must mirror sys::Geo in fill_systolic.cuh).
#include "kernels.cuh"

namespace ba {

#define DECL(s)                                                                                                         \
    cudaError_t launch_fill_systolic_s##s(const SysArgs&, int, int, size_t, bool, bool, bool, cudaStream_t);            \
    int sys_occupancy_s##s(bool, bool, bool, int, size_t);                                                              \
    cudaError_t launch_fill_systolic_long_s##s(const SysArgs&, int, int, size_t, bool, bool, cudaStream_t);             \
    int sys_occupancy_long_s##s(bool, bool, int, size_t, bool);                                                              \
    size_t sys_smem_bytes_s##s(bool, int, int, int, bool);                                                              \
    cudaError_t launch_fill_systolic_p16_s##s(const SysArgs&, int, int, size_t, cudaStream_t);                          \
    int sys_occupancy_p16_s##s(int, size_t);                                                                            \
    cudaError_t launch_fill_systolic_na_s##s(const SysArgs&, int, int, size_t, bool, bool, cudaStream_t);               \
    int sys_occupancy_na_s##s(bool, bool, int, size_t);                                                                 \
    cudaError_t launch_fill_systolic_chain_s##s(const SysArgs&, int, int, size_t, bool, cudaStream_t);                  \
    int sys_occupancy_chain_s##s(bool, int, size_t);                                                                    \
    cudaError_t launch_fill_systolic_rebase_s##s(const SysArgs&, int, int, size_t, bool, bool, cudaStream_t);             \
    int sys_occupancy_rebase_s##s(bool, bool, int, size_t);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4)
#undef DECL

SysGeo sys_geo(int S, bool pad) {
    SysGeo g;
    g.W = 2 * S + 1;
    g.LPR = pad ? 2 * S + 2 : 2 * S + 1;
    g.P = pad ? 2 * S + 2 : (S == 0 ? 2 : 2 * S + 1);
    g.R = 32 / g.LPR;
    g.RING = g.P + 3;
    g.REC = (12 * g.LPR + 3) & ~3;  // six ring values + six exchange values per lane of the row
    return g;
}

int sys_iters(int S, bool pad, int G, int m) {
    const SysGeo g = sys_geo(S, pad);
    return (m + 1) * g.P + 2 * (G * g.R - 1) + g.LPR + g.RING;
}
long long sys_code_words(int S, bool pad, int G, int n, int m) {
    const SysGeo g = sys_geo(S, pad);
    const int RT = G * g.R;
    return (long long)((n + RT) / RT) * G * (sys_iters(S, pad, G, m) + 4 /* PRE */) * 32;
}
// records -(PRE+1) .. nit-1 (PRE = 4 warm-up iterations of the kernel, one slack record in front)
size_t sys_boundary_ints(int S, bool pad, int G, int mmax) { return (size_t)(sys_iters(S, pad, G, mmax) + 8) * sys_geo(S, pad).REC; }

// Molecule B is staged with slack on both sides for every lane's position during the pipeline
// fill (negative columns) and drain (columns beyond m).
int sys_boff(int S, bool pad, int G) {
    const SysGeo g = sys_geo(S, pad);
    return (2 * G * g.R + 4 * g.P + 16) / g.P + 3 + S;
}
int sys_bpad(int S, bool pad, int G, int mmax) {
    const SysGeo g = sys_geo(S, pad);
    return sys_boff(S, pad, G) + mmax + (2 * G * g.R + 4 * g.P + 16) / g.P + S + 8;
}

size_t sys_smem_bytes(int S, bool pad, int G, int nsym, int mmax, bool p16) {
    const int bpad = sys_bpad(S, pad, G, mmax);
    switch (S) {
        case 0: return sys_smem_bytes_s0(pad, G, nsym, bpad, p16);
        case 1: return sys_smem_bytes_s1(pad, G, nsym, bpad, p16);
        case 2: return sys_smem_bytes_s2(pad, G, nsym, bpad, p16);
        case 3: return sys_smem_bytes_s3(pad, G, nsym, bpad, p16);
        default: return sys_smem_bytes_s4(pad, G, nsym, bpad, p16);
    }
}

int sys_occupancy_p16(int S, int G, size_t smem) {
    switch (S) {
        case 0: return sys_occupancy_p16_s0(G, smem);
        case 1: return sys_occupancy_p16_s1(G, smem);
        case 2: return sys_occupancy_p16_s2(G, smem);
        case 3: return sys_occupancy_p16_s3(G, smem);
        default: return sys_occupancy_p16_s4(G, smem);
    }
}

cudaError_t launch_fill_systolic_p16(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    switch (A.sc.s) {
        case 0: return launch_fill_systolic_p16_s0(A, grid, G, smem, st);
        case 1: return launch_fill_systolic_p16_s1(A, grid, G, smem, st);
        case 2: return launch_fill_systolic_p16_s2(A, grid, G, smem, st);
        case 3: return launch_fill_systolic_p16_s3(A, grid, G, smem, st);
        case 4: return launch_fill_systolic_p16_s4(A, grid, G, smem, st);
    }
    return cudaErrorInvalidValue;
}

int sys_occupancy(int S, bool trace, bool pad, bool bneg, int G, size_t smem) {
    switch (S) {
        case 0: return sys_occupancy_s0(trace, pad, bneg, G, smem);
        case 1: return sys_occupancy_s1(trace, pad, bneg, G, smem);
        case 2: return sys_occupancy_s2(trace, pad, bneg, G, smem);
        case 3: return sys_occupancy_s3(trace, pad, bneg, G, smem);
        default: return sys_occupancy_s4(trace, pad, bneg, G, smem);
    }
}

cudaError_t launch_fill_systolic(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, bool bneg, cudaStream_t st) {
    switch (A.sc.s) {
        case 0: return launch_fill_systolic_s0(A, grid, G, smem, trace, pad, bneg, st);
        case 1: return launch_fill_systolic_s1(A, grid, G, smem, trace, pad, bneg, st);
        case 2: return launch_fill_systolic_s2(A, grid, G, smem, trace, pad, bneg, st);
        case 3: return launch_fill_systolic_s3(A, grid, G, smem, trace, pad, bneg, st);
        case 4: return launch_fill_systolic_s4(A, grid, G, smem, trace, pad, bneg, st);
    }
    return cudaErrorInvalidValue;
}

int sys_occupancy_na(int S, bool trace, bool pad, int G, size_t smem) {
    switch (S) {
        case 0: return sys_occupancy_na_s0(trace, pad, G, smem);
        case 1: return sys_occupancy_na_s1(trace, pad, G, smem);
        case 2: return sys_occupancy_na_s2(trace, pad, G, smem);
        case 3: return sys_occupancy_na_s3(trace, pad, G, smem);
        default: return sys_occupancy_na_s4(trace, pad, G, smem);
    }
}

cudaError_t launch_fill_systolic_na(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st) {
    switch (A.sc.s) {
        case 0: return launch_fill_systolic_na_s0(A, grid, G, smem, trace, pad, st);
        case 1: return launch_fill_systolic_na_s1(A, grid, G, smem, trace, pad, st);
        case 2: return launch_fill_systolic_na_s2(A, grid, G, smem, trace, pad, st);
        case 3: return launch_fill_systolic_na_s3(A, grid, G, smem, trace, pad, st);
        case 4: return launch_fill_systolic_na_s4(A, grid, G, smem, trace, pad, st);
    }
    return cudaErrorInvalidValue;
}

int sys_occupancy_long(int S, bool trace, bool pad, int G, size_t smem, bool iow) {
    switch (S) {
        case 0: return sys_occupancy_long_s0(trace, pad, G, smem, iow);
        case 1: return sys_occupancy_long_s1(trace, pad, G, smem, iow);
        case 2: return sys_occupancy_long_s2(trace, pad, G, smem, iow);
        case 3: return sys_occupancy_long_s3(trace, pad, G, smem, iow);
        default: return sys_occupancy_long_s4(trace, pad, G, smem, iow);
    }
}

cudaError_t launch_fill_systolic_long(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st) {
    switch (A.sc.s) {
        case 0: return launch_fill_systolic_long_s0(A, grid, G, smem, trace, pad, st);
        case 1: return launch_fill_systolic_long_s1(A, grid, G, smem, trace, pad, st);
        case 2: return launch_fill_systolic_long_s2(A, grid, G, smem, trace, pad, st);
        case 3: return launch_fill_systolic_long_s3(A, grid, G, smem, trace, pad, st);
        case 4: return launch_fill_systolic_long_s4(A, grid, G, smem, trace, pad, st);
    }
    return cudaErrorInvalidValue;
}

int sys_occupancy_chain(int S, bool trace, int G, size_t smem) {
    switch (S) {
        case 0: return sys_occupancy_chain_s0(trace, G, smem);
        case 1: return sys_occupancy_chain_s1(trace, G, smem);
        case 2: return sys_occupancy_chain_s2(trace, G, smem);
        case 3: return sys_occupancy_chain_s3(trace, G, smem);
        default: return sys_occupancy_chain_s4(trace, G, smem);
    }
}

// shared memory of the chained flavour: the ordinary carve-up with one staged-B array of bpad_total bytes per kind
size_t sys_smem_bytes_chain(int S, int G, int nsym, int bpad_total) {
    switch (S) {
        case 0: return sys_smem_bytes_s0(false, G, nsym, bpad_total, false);
        case 1: return sys_smem_bytes_s1(false, G, nsym, bpad_total, false);
        case 2: return sys_smem_bytes_s2(false, G, nsym, bpad_total, false);
        case 3: return sys_smem_bytes_s3(false, G, nsym, bpad_total, false);
        default: return sys_smem_bytes_s4(false, G, nsym, bpad_total, false);
    }
}

cudaError_t launch_fill_systolic_chain(const SysArgs& A, int grid, int G, size_t smem, bool trace, cudaStream_t st) {
    switch (A.sc.s) {
        case 0: return launch_fill_systolic_chain_s0(A, grid, G, smem, trace, st);
        case 1: return launch_fill_systolic_chain_s1(A, grid, G, smem, trace, st);
        case 2: return launch_fill_systolic_chain_s2(A, grid, G, smem, trace, st);
        case 3: return launch_fill_systolic_chain_s3(A, grid, G, smem, trace, st);
        case 4: return launch_fill_systolic_chain_s4(A, grid, G, smem, trace, st);
    }
    return cudaErrorInvalidValue;
}

int sys_occupancy_rebase(int S, bool trace, bool lng, int G, size_t smem) {
    switch (S) {
        case 0: return sys_occupancy_rebase_s0(trace, lng, G, smem);
        case 1: return sys_occupancy_rebase_s1(trace, lng, G, smem);
        case 2: return sys_occupancy_rebase_s2(trace, lng, G, smem);
        case 3: return sys_occupancy_rebase_s3(trace, lng, G, smem);
        default: return sys_occupancy_rebase_s4(trace, lng, G, smem);
    }
}

cudaError_t launch_fill_systolic_rebase(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool lng, cudaStream_t st) {
    switch (A.sc.s) {
        case 0: return launch_fill_systolic_rebase_s0(A, grid, G, smem, trace, lng, st);
        case 1: return launch_fill_systolic_rebase_s1(A, grid, G, smem, trace, lng, st);
        case 2: return launch_fill_systolic_rebase_s2(A, grid, G, smem, trace, lng, st);
        case 3: return launch_fill_systolic_rebase_s3(A, grid, G, smem, trace, lng, st);
        case 4: return launch_fill_systolic_rebase_s4(A, grid, G, smem, trace, lng, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ba

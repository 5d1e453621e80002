// Instantiations of the systolic fill for max_shift = 1 (see fill_systolic.cuh).
#include "fill_systolic.cuh"
namespace ba {
cudaError_t launch_fill_systolic_s1(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, bool bneg, cudaStream_t st) {
    return sys::launch_s<1>(A, grid, G, smem, trace, pad, bneg, st);
}
cudaError_t launch_fill_systolic_long_s1(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st) {
    return sys::launch_long_s<1>(A, grid, G, smem, trace, pad, st);
}
int sys_occupancy_long_s1(bool trace, bool pad, int G, size_t smem, bool iow) { return sys::occ_long_s<1>(trace, pad, G, smem, iow); }
int sys_occupancy_s1(bool trace, bool pad, bool bneg, int G, size_t smem) { return sys::occ_s<1>(trace, pad, bneg, G, smem); }
size_t sys_smem_bytes_s1(bool pad, int G, int nsym, int bpad, bool p16) {
    return pad ? sys::smem_bytes_t<1, true>(G, nsym, bpad, p16) : sys::smem_bytes_t<1, false>(G, nsym, bpad, p16);
}
cudaError_t launch_fill_systolic_p16_s1(const SysArgs& A, int grid, int G, size_t smem, cudaStream_t st) {
    return sys::launch_p16_t<1>(A, grid, G, smem, st);
}
int sys_occupancy_p16_s1(int G, size_t smem) { return sys::occ_p16_t<1>(G, smem); }
cudaError_t launch_fill_systolic_na_s1(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool pad, cudaStream_t st) {
    return sys::launch_na_s<1>(A, grid, G, smem, trace, pad, st);
}
int sys_occupancy_na_s1(bool trace, bool pad, int G, size_t smem) { return sys::occ_na_s<1>(trace, pad, G, smem); }
cudaError_t launch_fill_systolic_chain_s1(const SysArgs& A, int grid, int G, size_t smem, bool trace, cudaStream_t st) {
    return trace ? sys::launch_chain_t<1, true>(A, grid, G, smem, st) : sys::launch_chain_t<1, false>(A, grid, G, smem, st);
}
int sys_occupancy_chain_s1(bool trace, int G, size_t smem) { return trace ? sys::occ_chain_t<1, true>(G, smem) : sys::occ_chain_t<1, false>(G, smem); }
cudaError_t launch_fill_systolic_rebase_s1(const SysArgs& A, int grid, int G, size_t smem, bool trace, bool lng, cudaStream_t st) {
    return sys::launch_rebase_s<1>(A, grid, G, smem, trace, lng, st);
}
int sys_occupancy_rebase_s1(bool trace, bool lng, int G, size_t smem) { return sys::occ_rebase_s<1>(trace, lng, G, smem); }
}  // namespace ba

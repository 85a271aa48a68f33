// Instances of the persistent SQP kernel (solve_kernel.cuh): the block size of the block-cyclic-reduction factor
// of its QP step is a compile-time constant (2*D).
#include <cuda_runtime.h>

#include "solve_kernel.cuh"
#include "kernels.h"

namespace tb200 {
SolveKernelFn solve_pair_kernel_for(int D);  // solve_kernels_pair.cu: rows over two consecutive waypoints
SolveKernelFn solve_kernel_for(int D, bool pair_rows) {
  if (pair_rows) return solve_pair_kernel_for(D);
  switch (D) {
    case 2: return solve_kernel<2, 0>;
    case 3: return solve_kernel<3, 0>;
    case 6: return solve_kernel<6, 0>;
    case 7: return solve_kernel<7, 0>;
    default: return nullptr;
  }
}
int qp_debug_prof(unsigned long long* out, int reset) {
#ifdef TB200_PROFILE
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_prof, z, sizeof(z));
    return 0;
  }
  cudaMemcpyFromSymbol(out, g_prof, 16 * sizeof(unsigned long long));
  return 0;
#else
  (void)out; (void)reset;
  return -1;
#endif
}
}  // namespace tb200

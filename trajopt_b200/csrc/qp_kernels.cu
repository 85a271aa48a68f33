// Instances of the QP subproblem kernel (qp_cta_kernel.cuh): the block size of the block-cyclic-reduction
// factor is a compile-time constant (2*D).
#include <cuda_runtime.h>

#include "qp_cta_kernel.cuh"
#include "kernels.h"

namespace tb200 {
QpKernelFn qp_pair_kernel_for(int D);  // qp_kernels_pair.cu: rows over two consecutive waypoints
QpKernelFn qp_kernel_for(int D, bool pair_rows) {
  if (pair_rows) return qp_pair_kernel_for(D);
  switch (D) {
    case 2: return qp_kernel<2, 0>;
    case 3: return qp_kernel<3, 0>;
    case 6: return qp_kernel<6, 0>;
    case 7: return qp_kernel<7, 0>;
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

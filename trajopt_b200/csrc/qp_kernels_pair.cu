// Instances of the QP subproblem kernel whose rows may span two consecutive waypoints (CartVel and continuous
// collision rows: 2*D coefficients per padded row).  Own translation unit: it builds beside qp_kernels.cu.
#include <cuda_runtime.h>

#include "qp_cta_kernel.cuh"
#include "kernels.h"

namespace tb200 {
QpKernelFn qp_pair_kernel_for(int D) {
  switch (D) {
    case 7: return qp_kernel<7, 1>;
    default: return nullptr;
  }
}
}  // namespace tb200

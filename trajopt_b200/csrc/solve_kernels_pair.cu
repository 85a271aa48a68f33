// Instances of the persistent SQP kernel whose QP rows may span two consecutive waypoints (CartVel and continuous
// collision rows: 2*D coefficients per padded row).  Own translation unit: it builds beside solve_kernels.cu.
#include <cuda_runtime.h>

#include "solve_kernel.cuh"
#include "kernels.h"

namespace tb200 {
SolveKernelFn solve_pair_kernel_for(int D) {
  switch (D) {
    case 7: return solve_kernel<7, 1>;
    default: return nullptr;
  }
}
}  // namespace tb200

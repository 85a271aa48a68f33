// Instances of the convexify + exact evaluation + SQP decision kernel (eval_kernel.cuh), one per joint count.
#include "eval_kernel.cuh"
#include "kernels.h"

namespace tb200 {
EvalKernelFn eval_kernel_for(int D) {
  switch (D) {
    case 2: return eval_convexify_decide_kernel<2>;
    case 3: return eval_convexify_decide_kernel<3>;
    case 6: return eval_convexify_decide_kernel<6>;
    case 7: return eval_convexify_decide_kernel<7>;
    default: return nullptr;
  }
}
}  // namespace tb200

// Exact values of the joint-space terms (JointPos/Vel/Acc Eq/Ineq cost and constraint objects,
// trajopt/src/trajectory_costs.cpp:139-623), shared by the evaluation kernel and the QP kernel's model values.
#pragma once
#include "device_types.cuh"

namespace tb200 {

__device__ inline double joint_err(const double* x, int D, int order, int t, int d, double target) {
  double e;
  if (order == 0)
    e = x[t * D + d];
  else if (order == 1)
    e = x[(t + 1) * D + d] - x[t * D + d];
  else
    e = x[t * D + d] - 2.0 * x[(t + 1) * D + d] + x[(t + 2) * D + d];
  return e - target;
}

// exact Cost::value / Constraint::violation of a joint-space object at x
__device__ inline double joint_obj_value(const DevProblem& p, const DevObj& o, const double* x) {
  const DevJointTerm& jt = p.joint_terms[o.term];
  double s = 0.0;
  for (int t = o.first; t < o.first + o.n_steps; ++t)
    for (int d = 0; d < p.D; ++d) {
      const double e = joint_err(x, p.D, o.order, t, d, jt.targets[d]);
      if (o.kind == OBJ_JOINT_EQ_COST)
        s += e * e * jt.coeffs[d];
      else if (o.kind == OBJ_JOINT_EQ_CNT)
        s += fabs(e * e * jt.coeffs[d]);  // value() is c*e^2 while the row is c*e (trajectory_costs.cpp:160 vs 173)
      else {
        s += fmax((e - jt.upper[d]) * jt.coeffs[d], 0.0);
        s += fmax((jt.lower[d] - e) * jt.coeffs[d], 0.0);
      }
    }
  return s;
}

}  // namespace tb200

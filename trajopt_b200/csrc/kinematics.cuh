// Forward kinematics, geometric Jacobian pieces and the Cartesian pose error on the device.
// Replaces (for the batched path) tesseract::kinematics::JointGroup::calcFwdKin / calcJacobian and
// tesseract::common::calcTransformError / calcJacobianTransformErrorDiff as they are used by
// trajopt/src/kinematic_terms.cpp:250-263, 348-366 and trajopt/src/collision_terms.cpp:203-250.
#pragma once
#include "device_types.cuh"

namespace tb200 {

struct Frame {
  double R[9];
  double p[3];
};

__device__ __forceinline__ void frame_mul(const Frame& a, const Frame& b, Frame& o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o.R[i * 3 + j] = a.R[i * 3] * b.R[j] + a.R[i * 3 + 1] * b.R[3 + j] + a.R[i * 3 + 2] * b.R[6 + j];
    o.p[i] = a.R[i * 3] * b.p[0] + a.R[i * 3 + 1] * b.p[1] + a.R[i * 3 + 2] * b.p[2] + a.p[i];
  }
}

// Local transform of one segment: origin * motion(q).
__device__ __forceinline__ void segment_local_q(const DevSegment& g, double qv, Frame& t);
__device__ __forceinline__ void segment_local(const DevSegment& g, const double* q, Frame& t) {
  segment_local_q(g, g.q_index >= 0 ? q[g.q_index] : 0.0, t);
}
__device__ __forceinline__ void segment_local_q(const DevSegment& g, double qv, Frame& t) {
#pragma unroll
  for (int i = 0; i < 9; ++i) t.R[i] = g.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t.p[i] = g.p[i];
  if (g.joint_type == 1) {  // revolute: Rodrigues about the joint axis
    double sn, c;
    sincos(qv, &sn, &c);
    const double v = 1.0 - c, x = g.axis[0], y = g.axis[1], z = g.axis[2];
    double m[9];
    m[0] = c + x * x * v;      m[1] = x * y * v - z * sn;  m[2] = x * z * v + y * sn;
    m[3] = y * x * v + z * sn; m[4] = c + y * y * v;       m[5] = y * z * v - x * sn;
    m[6] = z * x * v - y * sn; m[7] = z * y * v + x * sn;  m[8] = c + z * z * v;
    double r[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) r[i * 3 + j] = g.R[i * 3] * m[j] + g.R[i * 3 + 1] * m[3 + j] + g.R[i * 3 + 2] * m[6 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) t.R[i] = r[i];
  } else if (g.joint_type == 2) {  // prismatic
    const double d = qv;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      t.p[i] = g.p[i] + (g.R[i * 3] * g.axis[0] + g.R[i * 3 + 1] * g.axis[1] + g.R[i * 3 + 2] * g.axis[2]) * d;
  }
}

// rotation matrix -> (unit axis, signed angle in [-pi, pi]); identity -> axis (1,0,0), angle 0.
// Same convention as tesseract::common::calcRotationalError (axis = +v/|v|, sign carried by the angle).
__device__ __forceinline__ void rot_err_decomposed(const double* m, double* axis, double& angle) {
  double q0, q1, q2, q3;
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q0 = 0.5 * t;
    t = 0.5 / t;
    q1 = (m[7] - m[5]) * t;
    q2 = (m[2] - m[6]) * t;
    q3 = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    double qv[3];
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    q0 = (m[k * 3 + j] - m[j * 3 + k]) * t;
    qv[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    qv[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    q1 = qv[0];
    q2 = qv[1];
    q3 = qv[2];
  }
  const double n = sqrt(q1 * q1 + q2 * q2 + q3 * q3);
  if (n == 0.0) {
    axis[0] = 1.0;
    axis[1] = 0.0;
    axis[2] = 0.0;
    angle = 0.0;
    return;
  }
  double ang = 2.0 * atan2(n, fabs(q0));
  if (q0 < 0) ang = -ang;
  axis[0] = q1 / n;
  axis[1] = q2 / n;
  axis[2] = q3 / n;
  const double two_pi = 6.283185307179586476925286766559;
  const double pi = 3.14159265358979323846;
  // (tesseract reduces with fmod(|angle|, 2 pi) first; |angle| = 2 atan2(n, |q0|) <= pi never needs it, and fp64
  // fmod is a long loop on the GPU)
  if (ang < -pi)
    ang += two_pi;
  else if (ang > pi)
    ang -= two_pi;
  angle = ang;
}

// e = target^-1 * source : translation + rotation matrix
__device__ __forceinline__ void rel_pose(const Frame& tgt, const Frame& src, Frame& e) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) e.R[i * 3 + j] = tgt.R[i] * src.R[j] + tgt.R[3 + i] * src.R[3 + j] + tgt.R[6 + i] * src.R[6 + j];
    const double d0 = src.p[0] - tgt.p[0], d1 = src.p[1] - tgt.p[1], d2 = src.p[2] - tgt.p[2];
    e.p[i] = tgt.R[i] * d0 + tgt.R[3 + i] * d1 + tgt.R[6 + i] * d2;
  }
}

__device__ __forceinline__ void quat_to_frame(const double* pose7, Frame& f) {
  double w = pose7[3], x = pose7[4], y = pose7[5], z = pose7[6];
  const double n = sqrt(w * w + x * x + y * y + z * z);
  w /= n; x /= n; y /= n; z /= n;
  f.R[0] = 1 - 2 * (y * y + z * z); f.R[1] = 2 * (x * y - z * w);     f.R[2] = 2 * (x * z + y * w);
  f.R[3] = 2 * (x * y + z * w);     f.R[4] = 1 - 2 * (x * x + z * z); f.R[5] = 2 * (y * z - x * w);
  f.R[6] = 2 * (x * z - y * w);     f.R[7] = 2 * (y * z + x * w);     f.R[8] = 1 - 2 * (x * x + y * y);
  f.p[0] = pose7[0];
  f.p[1] = pose7[1];
  f.p[2] = pose7[2];
}

}  // namespace tb200

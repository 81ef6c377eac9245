// rot6D_to_angular for one joint, shared by the ROMP parse kernel and the BEV unpack kernel.
#pragma once
#include "common.cuh"

namespace b200romp {

// rotation_matrix_to_quaternion (utils.py:606-682) + quaternion_to_angle_axis (:554-604) on m = R^T (row-major), NaN -> 0 (:551)
__device__ __forceinline__ void rt_to_aa(const float* m, float* aa) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
  const bool d2 = m22 < 1e-6f, d01 = m00 > m11, d0n1 = m00 < -m11;   // :640-643
  float q0, q1, q2, q3, t;
  if (d2 && d01) {          // :645-649
    t = 1.f + m00 - m11 - m22;
    q0 = m12 - m21; q1 = t; q2 = m01 + m10; q3 = m20 + m02;
  } else if (d2 && !d01) {  // :651-655
    t = 1.f - m00 + m11 - m22;
    q0 = m20 - m02; q1 = m01 + m10; q2 = t; q3 = m12 + m21;
  } else if (d0n1) {        // :657-661
    t = 1.f - m00 - m11 + m22;
    q0 = m01 - m10; q1 = m20 + m02; q2 = m12 + m21; q3 = t;
  } else {                  // :663-667
    t = 1.f + m00 + m11 + m22;
    q0 = t; q1 = m12 - m21; q2 = m20 - m02; q3 = m01 - m10;
  }
  const float r = sqrtf(t);                                      // :679
  q0 = q0 / r * 0.5f; q1 = q1 / r * 0.5f; q2 = q2 / r * 0.5f; q3 = q3 / r * 0.5f;   // :681
  const float s2 = q1 * q1 + q2 * q2 + q3 * q3;                  // :587
  const float s = sqrtf(s2);
  const float two_theta = 2.f * (q0 < 0.f ? atan2f(-s, -q0) : atan2f(s, q0));   // :591-594
  const float k = s2 > 0.f ? two_theta / s : 2.f;                // :596-598
  float o0 = q1 * k, o1 = q2 * k, o2 = q3 * k;
  aa[0] = isnan(o0) ? 0.f : o0;                                  // :551
  aa[1] = isnan(o1) ? 0.f : o1;
  aa[2] = isnan(o2) ? 0.f : o2;
}

// ---- rot6D_to_angular for one joint (fp32, same op order as the reference) ------------------------
__device__ __forceinline__ void rot6d_to_aa(const float* x, float* aa) {
  // x.view(3,2): column 0 = (x0,x2,x4), column 1 = (x1,x3,x5)   utils.py:478
  float a0 = x[0], a1 = x[2], a2 = x[4];
  float c0 = x[1], c1 = x[3], c2 = x[5];
  float n1 = fmaxf(sqrtf(a0 * a0 + a1 * a1 + a2 * a2), 1e-6f);   // F.normalize eps, :481
  const float b10 = a0 / n1, b11 = a1 / n1, b12 = a2 / n1;
  const float dot = b10 * c0 + b11 * c1 + b12 * c2;              // :483
  const float u0 = c0 - dot * b10, u1 = c1 - dot * b11, u2 = c2 - dot * b12;
  const float n2 = fmaxf(sqrtf(u0 * u0 + u1 * u1 + u2 * u2), 1e-6f);
  const float b20 = u0 / n2, b21 = u1 / n2, b22 = u2 / n2;       // :485
  const float b30 = b11 * b22 - b12 * b21, b31 = b12 * b20 - b10 * b22, b32 = b10 * b21 - b11 * b20;  // :488
  // R = [b1 b2 b3] (columns); the quaternion code works on Rt = R^T: m(i,j) = R[j][i]   :489,638
  const float Rt[9] = {b10, b11, b12, b20, b21, b22, b30, b31, b32};   // rows of Rt = b1, b2, b3
  rt_to_aa(Rt, aa);
}

// rotation_matrix_to_angle_axis (utils.py:535-552) of a row-major 3x3 matrix R (not necessarily orthonormal: the
// temporal path feeds it a low-pass filtered matrix, utils.py:188-192)
__device__ __forceinline__ void rotmat_to_aa(const float* R, float* aa) {
  const float Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
  rt_to_aa(Rt, aa);
}

}  // namespace b200romp

// ob_math.h — scalar / small-vector arithmetic shared by host model and CUDA kernels.
//
// Numeric contract: every expression is evaluated in
// dReal, left to right, WITHOUT fused multiply-add, with IEEE division and
// square root, exactly as the reference's scalar SSE build does.  The CUDA
// side is therefore compiled with  -fmad=false -prec-div=true -prec-sqrt=true
// -ftz=false  and the host side with  -ffp-contract=off.  Never replace the
// compare-branch clamps with fmin/fmax (NaN semantics differ).
//
// Matrices are ODE's 3x4 row-major layout (element (i,j) at [4*i+j]); vectors
// may be 3 or 4 wide.  Reference: include/ode/odemath.h:150-410,
// ode/src/odemath.cpp:42-177, ode/src/rotation.cpp:191-317.
#pragma once
#include <math.h>
#include <stdint.h>

// Lock-step helpers for per-lane loops whose lanes would otherwise drift apart (a BVH walk followed by a triangle test: lanes
// that `continue` early re-enter the walk while their neighbours are still in the test, and every region of the loop runs
// with a fraction of the warp).  The lanes that enter such a loop together vote at the top of every iteration; the vote
// reconverges them, and a lane that is done idles until all are.  On the host a lane is alone.
#if defined(__CUDA_ARCH__)
#define OB_LANES_TOGETHER() __activemask()
#define OB_ALL_LANES(mask, pred) (__all_sync((mask), (pred)) != 0)
#else
#define OB_LANES_TOGETHER() 0u
#define OB_ALL_LANES(mask, pred) (pred)
#endif
#if defined(__CUDACC__)
#define OB_HD __host__ __device__ __forceinline__
#define OB_HDN static __host__ __device__ __noinline__
#else
#define OB_HD inline
#define OB_HDN inline
#endif

#if !defined(dSINGLE) && !defined(dDOUBLE)
#define dSINGLE 1
#endif
#if defined(dSINGLE)
typedef float real;
#define OB_REAL(x) (x##f)
#define OB_INF (__builtin_inff())
#else
typedef double real;
#define OB_REAL(x) (x)
#define OB_INF (__builtin_inf())
#endif

#define OB_PI 3.14159265358979323846
#define OB_SQRT1_2 0.70710678118654752440

OB_HD real ob_sqrt(real x) {
#if defined(dSINGLE)
  return sqrtf(x);
#else
  return sqrt(x);
#endif
}
OB_HD real ob_fabs(real x) {
#if defined(dSINGLE)
  return fabsf(x);
#else
  return fabs(x);
#endif
}
OB_HD real ob_recip(real x) { return OB_REAL(1.0) / x; }
#if defined(dSINGLE)
#define OB_EPSILON 1.1920928955078125e-7f      // FLT_EPSILON (dEpsilon, ode/src/config.h:93)
#else
#define OB_EPSILON 2.2204460492503131e-16      // DBL_EPSILON (config.h:95)
#endif
OB_HD real ob_recipsqrt(real x) { return OB_REAL(1.0) / ob_sqrt(x); }

// a.b with strides (odemath.h:175-178): a0*b0 + a1*b1 + a2*b2, left to right
OB_HD real ob_dot(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
OB_HD real ob_dot14(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[4] + a[2] * b[8]; }
OB_HD real ob_dot41(const real *a, const real *b) { return a[0] * b[0] + a[4] * b[1] + a[8] * b[2]; }
OB_HD real ob_dot44(const real *a, const real *b) { return a[0] * b[0] + a[4] * b[4] + a[8] * b[8]; }

// res = a x b (odemath.h:196-206)
OB_HD void ob_cross(real *res, const real *a, const real *b) {
  real r0 = a[1] * b[2] - a[2] * b[1];
  real r1 = a[2] * b[0] - a[0] * b[2];
  real r2 = a[0] * b[1] - a[1] * b[0];
  res[0] = r0; res[1] = r1; res[2] = r2;
}

// res = A*b, A 3x4 (dMultiply0_331)
OB_HD void ob_mul0_331(real *res, const real *A, const real *b) {
  real r0 = ob_dot(A, b), r1 = ob_dot(A + 4, b), r2 = ob_dot(A + 8, b);
  res[0] = r0; res[1] = r1; res[2] = r2;
}
// res = A^T*b (dMultiply1_331)
OB_HD void ob_mul1_331(real *res, const real *A, const real *b) {
  real r0 = ob_dot41(A, b), r1 = ob_dot41(A + 1, b), r2 = ob_dot41(A + 2, b);
  res[0] = r0; res[1] = r1; res[2] = r2;
}
// res = A*B (dMultiply0_333): row i of res = (B col j . A row i), expression b[j]*a0 + b[4+j]*a1 + b[8+j]*a2
OB_HD void ob_mul0_333(real *res, const real *A, const real *B) {
  for (int i = 0; i < 3; i++) {
    const real *a = A + 4 * i;
    real r0 = ob_dot41(B, a), r1 = ob_dot41(B + 1, a), r2 = ob_dot41(B + 2, a);
    res[4 * i] = r0; res[4 * i + 1] = r1; res[4 * i + 2] = r2;
  }
}
// res = A*B^T (dMultiply2_333): res(i,j) = B row j . A row i
OB_HD void ob_mul2_333(real *res, const real *A, const real *B) {
  for (int i = 0; i < 3; i++) {
    const real *a = A + 4 * i;
    real r0 = ob_dot(B, a), r1 = ob_dot(B + 4, a), r2 = ob_dot(B + 8, a);
    res[4 * i] = r0; res[4 * i + 1] = r1; res[4 * i + 2] = r2;
  }
}
// res = A^T*B (dMultiply1_333): res(i,j) = A col i . B col j
OB_HD void ob_mul1_333(real *res, const real *A, const real *B) {
  for (int i = 0; i < 3; i++) {
    real r0 = ob_dot44(B, A + i), r1 = ob_dot44(B + 1, A + i), r2 = ob_dot44(B + 2, A + i);
    res[4 * i] = r0; res[4 * i + 1] = r1; res[4 * i + 2] = r2;
  }
}

// _dSafeNormalize3 (odemath.cpp:42-85)
OB_HD int ob_safe_normalize3(real *a) {
  int idx;
  real aa0 = ob_fabs(a[0]), aa1 = ob_fabs(a[1]), aa2 = ob_fabs(a[2]), l, s;
  if (aa1 > aa0) {
    if (aa2 > aa1) idx = 2; else idx = 1;
  } else {
    if (aa2 > aa0) idx = 2;
    else {
      if (aa0 <= 0) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
      idx = 0;
    }
  }
  s = (idx == 0) ? aa0 : ((idx == 1) ? aa1 : aa2);
  a[0] /= s; a[1] /= s; a[2] /= s;
  l = ob_recipsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  a[0] *= l; a[1] *= l; a[2] *= l;
  return 1;
}
// dRFrom2Axes (rotation.cpp:94-133); returns 0 (R untouched) for a zero-length vector
OB_HD int ob_Rfrom2axes(real *R, real ax, real ay, real az, real bx, real by, real bz) {
  real l = ob_sqrt(ax * ax + ay * ay + az * az);
  if (l <= OB_REAL(0.0)) return 0;
  l = ob_recip(l);
  ax *= l; ay *= l; az *= l;
  const real k = ax * bx + ay * by + az * bz;
  bx -= k * ax; by -= k * ay; bz -= k * az;
  l = ob_sqrt(bx * bx + by * by + bz * bz);
  if (l <= OB_REAL(0.0)) return 0;
  l = ob_recip(l);
  bx *= l; by *= l; bz *= l;
  R[0] = ax; R[4] = ay; R[8] = az;
  R[1] = bx; R[5] = by; R[9] = bz;
  R[2] = -by * az + ay * bz;
  R[6] = -bz * ax + az * bx;
  R[10] = -bx * ay + ax * by;
  R[3] = R[7] = R[11] = OB_REAL(0.0);
  return 1;
}
// dQfromR (rotation.cpp:258-307)
OB_HD void ob_QfromR(real *q, const real *R) {
  real tr = R[0] + R[5] + R[10], s;
  if (tr >= 0) {
    s = ob_sqrt(tr + 1);
    q[0] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[1] = (R[9] - R[6]) * s; q[2] = (R[2] - R[8]) * s; q[3] = (R[4] - R[1]) * s;
    return;
  }
  int c;
  if (R[5] > R[0]) c = (R[10] > R[5]) ? 2 : 1;
  else c = (R[10] > R[0]) ? 2 : 0;
  if (c == 0) {
    s = ob_sqrt((R[0] - (R[5] + R[10])) + 1);
    q[1] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[2] = (R[1] + R[4]) * s; q[3] = (R[8] + R[2]) * s; q[0] = (R[9] - R[6]) * s;
  } else if (c == 1) {
    s = ob_sqrt((R[5] - (R[10] + R[0])) + 1);
    q[2] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[3] = (R[6] + R[9]) * s; q[1] = (R[1] + R[4]) * s; q[0] = (R[2] - R[8]) * s;
  } else {
    s = ob_sqrt((R[10] - (R[0] + R[5])) + 1);
    q[3] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[1] = (R[8] + R[2]) * s; q[2] = (R[6] + R[9]) * s; q[0] = (R[4] - R[1]) * s;
  }
}
// _dSafeNormalize4 (odemath.cpp:119-139)
OB_HD int ob_safe_normalize4(real *a) {
  real l = ob_dot(a, a) + a[3] * a[3];
  if (l > 0) {
    l = ob_recipsqrt(l);
    a[0] *= l; a[1] *= l; a[2] *= l; a[3] *= l;
    return 1;
  }
  a[0] = 1; a[1] = 0; a[2] = 0; a[3] = 0;
  return 0;
}
// dPlaneSpace (odemath.cpp:151-177); the fabs compare is against the double M_SQRT1_2
OB_HD void ob_plane_space(const real *n, real *p, real *q) {
  if ((double)ob_fabs(n[2]) > OB_SQRT1_2) {
    real a = n[1] * n[1] + n[2] * n[2];
    real k = ob_recipsqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    real a = n[0] * n[0] + n[1] * n[1];
    real k = ob_recipsqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}
// dRfromQ (rotation.cpp:236-256)
OB_HD void ob_RfromQ(real *R, const real *q) {
  real qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
  R[0] = 1 - qq2 - qq3;
  R[1] = 2 * (q[1] * q[2] - q[0] * q[3]);
  R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
  R[3] = 0;
  R[4] = 2 * (q[1] * q[2] + q[0] * q[3]);
  R[5] = 1 - qq1 - qq3;
  R[6] = 2 * (q[2] * q[3] - q[0] * q[1]);
  R[7] = 0;
  R[8] = 2 * (q[1] * q[3] - q[0] * q[2]);
  R[9] = 2 * (q[2] * q[3] + q[0] * q[1]);
  R[10] = 1 - qq1 - qq2;
  R[11] = 0;
}
// dDQfromW (rotation.cpp:310-317)
OB_HD void ob_DQfromW(real *dq, const real *w, const real *q) {
  dq[0] = OB_REAL(0.5) * (-w[0] * q[1] - w[1] * q[2] - w[2] * q[3]);
  dq[1] = OB_REAL(0.5) * (w[0] * q[0] + w[1] * q[3] - w[2] * q[2]);
  dq[2] = OB_REAL(0.5) * (-w[0] * q[3] + w[1] * q[0] + w[2] * q[1]);
  dq[3] = OB_REAL(0.5) * (w[0] * q[2] - w[1] * q[1] + w[2] * q[0]);
}
// dQMultiply0 (rotation.cpp:191-198)
OB_HD void ob_qmul0(real *qa, const real *qb, const real *qc) {
  real a0 = qb[0] * qc[0] - qb[1] * qc[1] - qb[2] * qc[2] - qb[3] * qc[3];
  real a1 = qb[0] * qc[1] + qb[1] * qc[0] + qb[2] * qc[3] - qb[3] * qc[2];
  real a2 = qb[0] * qc[2] + qb[2] * qc[0] + qb[3] * qc[1] - qb[1] * qc[3];
  real a3 = qb[0] * qc[3] + qb[3] * qc[0] + qb[1] * qc[2] - qb[2] * qc[1];
  qa[0] = a0; qa[1] = a1; qa[2] = a2; qa[3] = a3;
}

// dQMultiply1/2/3 (rotation.cpp:201-228)
OB_HD void ob_qmul1(real *qa, const real *qb, const real *qc) {
  real a0 = qb[0] * qc[0] + qb[1] * qc[1] + qb[2] * qc[2] + qb[3] * qc[3];
  real a1 = qb[0] * qc[1] - qb[1] * qc[0] - qb[2] * qc[3] + qb[3] * qc[2];
  real a2 = qb[0] * qc[2] - qb[2] * qc[0] - qb[3] * qc[1] + qb[1] * qc[3];
  real a3 = qb[0] * qc[3] - qb[3] * qc[0] - qb[1] * qc[2] + qb[2] * qc[1];
  qa[0] = a0; qa[1] = a1; qa[2] = a2; qa[3] = a3;
}
OB_HD void ob_qmul2(real *qa, const real *qb, const real *qc) {
  real a0 = qb[0] * qc[0] + qb[1] * qc[1] + qb[2] * qc[2] + qb[3] * qc[3];
  real a1 = -qb[0] * qc[1] + qb[1] * qc[0] - qb[2] * qc[3] + qb[3] * qc[2];
  real a2 = -qb[0] * qc[2] + qb[2] * qc[0] - qb[3] * qc[1] + qb[1] * qc[3];
  real a3 = -qb[0] * qc[3] + qb[3] * qc[0] - qb[1] * qc[2] + qb[2] * qc[1];
  qa[0] = a0; qa[1] = a1; qa[2] = a2; qa[3] = a3;
}
OB_HD void ob_qmul3(real *qa, const real *qb, const real *qc) {
  real a0 = qb[0] * qc[0] - qb[1] * qc[1] - qb[2] * qc[2] - qb[3] * qc[3];
  real a1 = -qb[0] * qc[1] - qb[1] * qc[0] + qb[2] * qc[3] - qb[3] * qc[2];
  real a2 = -qb[0] * qc[2] - qb[2] * qc[0] + qb[3] * qc[1] - qb[1] * qc[3];
  real a3 = -qb[0] * qc[3] - qb[3] * qc[0] + qb[1] * qc[2] - qb[2] * qc[1];
  qa[0] = a0; qa[1] = a1; qa[2] = a2; qa[3] = a3;
}
// atan2 with the host libm's results.  The reference calls glibc's atan2f (dAtan2,
// include/ode/common.h:158) in cullPoints (box.cpp:280), getHingeAngle (joint.cpp:391-393) and
// hinge2 measureAngle (hinge2.cpp:42); the value itself enters limit rows (limit_err), so a
// merely faithful atan2 is not enough for bit parity.  glibc 2.39 (the libm pinned in this
// image; third-party, not part of /root/reference) implements atan2f / atanf with the
// fdlibm single-precision algorithm (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c):
// argument reduction to one of four breakpoints + an odd/even split degree-11 polynomial,
// all in float arithmetic.  Restated here; tests/test_abi.py checks it bit-for-bit against
// the host atan2f on 10^6 inputs.  dDOUBLE uses the double atan2 of the platform
// (glibc's is correctly rounded, CUDA's is <= 2 ulp): tolerance-only there.
OB_HD int32_t ob_f2i(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  union { float f; int32_t i; } u; u.f = x; return u.i;
#endif
}
OB_HD float ob_i2f(int32_t i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  union { float f; int32_t i; } u; u.i = i; return u.f;
#endif
}
OB_HD float ob_atanf_glibc(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  float w, s1, s2, z;
  int32_t ix, hx, id;
  hx = ob_f2i(x);
  ix = hx & 0x7fffffff;
  if (ix >= 0x4c000000) {
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return atanhi[3] + atanlo[3];
    return -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
      else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
      else { id = 3; x = -1.0f / x; }
    }
  }
  z = x * x;
  w = z * z;
  s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}
OB_HD float ob_atan2f_glibc(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  float z;
  int32_t k, m, hx, hy, ix, iy;
  hx = ob_f2i(x); ix = hx & 0x7fffffff;
  hy = ob_f2i(y); iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return ob_atanf_glibc(y);
  m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    if (m == 0 || m == 1) return y;
    if (m == 2) return pi + tiny;
    return -pi - tiny;
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      if (m == 0) return pi_o_4 + tiny;
      if (m == 1) return -pi_o_4 - tiny;
      if (m == 2) return 3.0f * pi_o_4 + tiny;
      return -3.0f * pi_o_4 - tiny;
    }
    if (m == 0) return 0.0f;
    if (m == 1) return -0.0f;
    if (m == 2) return pi + tiny;
    return -pi - tiny;
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  k = (iy - ix) >> 23;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = ob_atanf_glibc(fabsf(y / x));
  if (m == 0) return z;
  if (m == 1) return ob_i2f(ob_f2i(z) ^ (int32_t)0x80000000);
  if (m == 2) return pi - (z - pi_lo);
  return (z - pi_lo) - pi;
}
OB_HD real ob_atan2(real y, real x) {
#if defined(dSINGLE)
  return ob_atan2f_glibc(y, x);
#else
  return atan2(y, x);
#endif
}

// sinf / cosf with the host libm's results: dxStepBody's finite-rotation integrator (util.cpp:288-330) calls dSin / dCos
// (sinf / cosf in dSINGLE) on the half rotation angle, and the value lands in the body's quaternion.  glibc 2.39
// (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf.h — the ARM optimized-routines algorithm; third-party, not part
// of /root/reference) evaluates both in DOUBLE precision: |x| < pi/4 a degree-7 / degree-8 polynomial, |x| < 120 a
// multiply-and-round quadrant reduction, beyond that a 2/pi bit-table reduction.  x86-64 libm dispatches (ifunc) to its
// FMA build on every CPU with FMA3, i.e. the polynomial and the reduction contract a*b+c; restated here with explicit
// fma() in exactly those places.  Checked against the host's sinf / cosf on ALL 2^32 float inputs (0 mismatches with
// fma, tests/test_abi.py runs a sampled version; the coefficient table was read back from this image's libm.so.6).
// dDOUBLE uses the platform's sin / cos (glibc's and CUDA's differ in the last bits): tolerance class, like atan2.
OB_HD double ob_fma64(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return fma(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
OB_HD uint32_t ob_abstop12(float x) { return ((uint32_t)ob_f2i(x) >> 20) & 0x7ffu; }
// sign folded into the polynomial's coefficients (table entry 1 of __sincosf_table negates the cosine part)
OB_HD float ob_sincosf_poly(double x, double x2, int neg, int n) {
  const double sg = neg ? -1.0 : 1.0;
  if ((n & 1) == 0) {
    const double S1 = -0.16666654943701084 /* 0x1.555545995a603p-3 */, S2 = 0.008332178146138854 /* 0x1.1107605230bc4p-7 */, S3 = -0.00019517298981385725 /* 0x1.994eb3774cf24p-13 */;
    const double x3 = x * x2, s1 = ob_fma64(x2, S3, S2), x7 = x3 * x2, s = ob_fma64(x3, S1, x);
    return (float)ob_fma64(x7, s1, s);
  } else {
    const double C0 = sg * 1.0 /* 0x1p0 */, C1 = sg * -0.49999999725108224 /* 0x1.ffffffd0c621cp-2 */, C2 = sg * 0.041666623324344516 /* 0x1.55553e1068f19p-5 */, C3 = sg * -0.001388676379437604 /* 0x1.6c087e89a359dp-10 */,
                 C4 = sg * 2.4390450703564542e-05 /* 0x1.99343027bf8c3p-16 */;
    const double x4 = x2 * x2, c2 = ob_fma64(x2, C4, C3), c1 = ob_fma64(x2, C1, C0), x6 = x4 * x2, c = ob_fma64(x4, C2, c1);
    return (float)ob_fma64(x6, c2, c);
  }
}
OB_HD double ob_sincosf_reduce(float y, int *np, int *signp) {
  const double hpi_inv = 10680707.430881744 /* 0x1.45F306DC9C883p+23 */, hpi = 1.5707963267948966 /* 0x1.921FB54442D18p0 */;
  *signp = 0;
  if (ob_abstop12(y) < ob_abstop12(120.0f)) {
    const double x = (double)y, r = x * hpi_inv;
    const int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return ob_fma64(-(double)n, hpi, x);
  }
  const uint32_t inv_pio4[24] = {0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27,
                                 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295,
                                 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
  uint32_t xi = (uint32_t)ob_f2i(y);
  *signp = (int)(xi >> 31);
  const uint32_t *arr = &inv_pio4[(xi >> 26) & 15];
  const int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint64_t)(uint32_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  const double x = (double)(int64_t)res0;
  *np = (int)n;
  return x * 3.4061215800865545e-19 /* 0x1.921FB54442D18p-62 */;
}
OB_HD float ob_sinf_glibc(float y) {
  if (ob_abstop12(y) < ob_abstop12(0.785398185f)) {
    if (ob_abstop12(y) < ob_abstop12(0.000244140625f)) return y;
    const double x = (double)y;
    return ob_sincosf_poly(x, x * x, 0, 0);
  }
  if (!(ob_abstop12(y) < ob_abstop12(ob_i2f(0x7f800000)))) return y - y;   // inf / nan -> nan
  int n, sign;
  const double x = ob_sincosf_reduce(y, &n, &sign);
  const int q = n + sign;
  const double s = ((q & 3) == 1 || (q & 3) == 2) ? -1.0 : 1.0;
  return ob_sincosf_poly(x * s, x * x, (q & 2) != 0, n);
}
OB_HD float ob_cosf_glibc(float y) {
  if (ob_abstop12(y) < ob_abstop12(0.785398185f)) {
    if (ob_abstop12(y) < ob_abstop12(0.000244140625f)) return 1.0f;
    const double x = (double)y;
    return ob_sincosf_poly(x, x * x, 0, 1);
  }
  if (!(ob_abstop12(y) < ob_abstop12(ob_i2f(0x7f800000)))) return y - y;
  int n, sign;
  const double x = ob_sincosf_reduce(y, &n, &sign);
  const int q = n + sign;
  const double s = ((q & 3) == 1 || (q & 3) == 2) ? -1.0 : 1.0;
  return ob_sincosf_poly(x * s, x * x, (q & 2) != 0, n ^ 1);
}

// ---- LCG behind the SOR row shuffle (misc.cpp:33-38, 66-117) -----------------
OB_HD uint32_t ob_lcg_next(uint32_t s) { return 1664525u * s + 1013904223u; }
// fold + modulus part of dRandInt applied to an already-advanced state r
OB_HD int ob_randint_fold(uint32_t r, uint32_t un) {
  if (un <= 0x10u) {
    r ^= (r >> 16); r ^= (r >> 8); r ^= (r >> 4);
    if (un <= 0x2u) { r ^= (r >> 2); r ^= (r >> 1); }
    else if (un <= 0x4u) { r ^= (r >> 2); }
  } else if (un <= 0x100u) {
    r ^= (r >> 16); r ^= (r >> 8);
  } else if (un <= 0x10000u) {
    r ^= (r >> 16);
  }
  return (int)(r % un);
}
// advance the LCG by k steps in O(log k): returns (A^k, C_k) with s_k = A^k s + C_k
OB_HD void ob_lcg_skip(uint32_t k, uint32_t *A, uint32_t *C) {
  uint32_t a = 1664525u, c = 1013904223u, ra = 1u, rc = 0u;
  while (k) {
    if (k & 1u) { ra = ra * a; rc = rc * a + c; }
    c = c * a + c;  // (a,c) o (a,c): s -> a(a s + c) + c
    a = a * a;
    k >>= 1;
  }
  *A = ra; *C = rc;
}

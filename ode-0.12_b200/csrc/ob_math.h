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

#if defined(__CUDACC__)
#define OB_HD __host__ __device__ __forceinline__
#define OB_HDN __host__ __device__ __noinline__
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
// glibc's atan2f/atan2 are not available on the device.  For dSINGLE we evaluate atan2 in
// double and round once (a faithful float result) — it only feeds
// discrete decisions (cullPoints ranking, joint-limit activation), checked against the
// reference in tests.  For dDOUBLE the CUDA/glibc double routines are both <1 ulp.
OB_HD real ob_atan2(real y, real x) { return (real)atan2((double)y, (double)x); }

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

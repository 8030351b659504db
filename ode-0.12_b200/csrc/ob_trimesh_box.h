// ob_trimesh_box.h — box vs trimesh, per-thread device function.
//   dCollideBTL                     ode/src/collision_trimesh_box.cpp:1217-1274
//   _cldTestSeparatingAxes          :389-636 (13 axes; edge axes penalised x1.5, :324)
//   _cldClipping                    :676-1053 (edge-edge point / box face clipped by the triangle / triangle clipped by a box face)
//   GenerateContact                 :1357-1446 (dedupe by position and normal within dEpsilon, keeps the deepest; the
//                                   triangle loop keeps scanning after max_contacts is reached, :1128-1136)
//   OBB query                       OPCODE/OPC_OBBCollider.cpp:177-345 (InitQuery), OPC_BoxBoxOverlap.h:68-116,
//                                   OBBContainsBox :348-397, TriBoxOverlap OPC_TriBoxOverlap.h:198-240 (primitive tests ON)
// The query runs in float in the mesh's model space, contact generation in dReal, as in the reference.
#pragma once
#include "ob_trimesh.h"

struct ObMat44f { float m[4][4]; };
OB_HD void ob_m44_from_pr(const real *pos, const real *R, ObMat44f *o) {   // MakeMatrix, collision_trimesh_internal.h:419-442
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o->m[i][j] = (float)R[4 * j + i];
  for (int j = 0; j < 3; j++) o->m[3][j] = (float)pos[j];
  o->m[0][3] = 0.0f; o->m[1][3] = 0.0f; o->m[2][3] = 0.0f; o->m[3][3] = 1.0f;
}
OB_HD void ob_m44_invert_pr(const ObMat44f &s, ObMat44f *d) {   // InvertPRMatrix, Ice/IceMatrix4x4.cpp:54-75
  d->m[0][0] = s.m[0][0]; d->m[1][0] = s.m[0][1]; d->m[2][0] = s.m[0][2];
  d->m[3][0] = -(s.m[3][0] * s.m[0][0] + s.m[3][1] * s.m[0][1] + s.m[3][2] * s.m[0][2]);
  d->m[0][1] = s.m[1][0]; d->m[1][1] = s.m[1][1]; d->m[2][1] = s.m[1][2];
  d->m[3][1] = -(s.m[3][0] * s.m[1][0] + s.m[3][1] * s.m[1][1] + s.m[3][2] * s.m[1][2]);
  d->m[0][2] = s.m[2][0]; d->m[1][2] = s.m[2][1]; d->m[2][2] = s.m[2][2];
  d->m[3][2] = -(s.m[3][0] * s.m[2][0] + s.m[3][1] * s.m[2][1] + s.m[3][2] * s.m[2][2]);
  d->m[0][3] = 0.0f; d->m[1][3] = 0.0f; d->m[2][3] = 0.0f; d->m[3][3] = 1.0f;
}
OB_HD void ob_m44_mul(const ObMat44f &a, const ObMat44f &b, ObMat44f *o) {   // Matrix4x4::operator*, Ice/IceMatrix4x4.h:278-300
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      o->m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
}

struct ObObbQuery {
  float ext[3];          // mBoxExtents
  float rM2B[3][3], tM2B[3];   // mRModelToBox, mTModelToBox
  float rB2M[3][3], tB2M[3];   // mRBoxToModel, mTBoxToModel
  float AR[3][3];
  float B0[3], B1[3];
  float BBx1, BBy1, BBz1, BB[9];
  OB_HD static bool greater(float x, float y) { return fabsf(x) > y; }   // GREATER, OPC_Common.h:27
  OB_HD bool overlap(const ObBvNode &n) const {   // OBBCollider::BoxBoxOverlap, full test
    const float *e = n.e;
    float t, t2;
    const float Tx = tB2M[0] - n.c[0]; t = e[0] + BBx1; if (greater(Tx, t)) return false;
    const float Ty = tB2M[1] - n.c[1]; t = e[1] + BBy1; if (greater(Ty, t)) return false;
    const float Tz = tB2M[2] - n.c[2]; t = e[2] + BBz1; if (greater(Tz, t)) return false;
    t = Tx * rB2M[0][0] + Ty * rB2M[0][1] + Tz * rB2M[0][2];
    t2 = e[0] * AR[0][0] + e[1] * AR[0][1] + e[2] * AR[0][2] + ext[0];
    if (greater(t, t2)) return false;
    t = Tx * rB2M[1][0] + Ty * rB2M[1][1] + Tz * rB2M[1][2];
    t2 = e[0] * AR[1][0] + e[1] * AR[1][1] + e[2] * AR[1][2] + ext[1];
    if (greater(t, t2)) return false;
    t = Tx * rB2M[2][0] + Ty * rB2M[2][1] + Tz * rB2M[2][2];
    t2 = e[0] * AR[2][0] + e[1] * AR[2][1] + e[2] * AR[2][2] + ext[2];
    if (greater(t, t2)) return false;
    t = Tz * rB2M[0][1] - Ty * rB2M[0][2]; t2 = e[1] * AR[0][2] + e[2] * AR[0][1] + BB[0]; if (greater(t, t2)) return false;
    t = Tz * rB2M[1][1] - Ty * rB2M[1][2]; t2 = e[1] * AR[1][2] + e[2] * AR[1][1] + BB[1]; if (greater(t, t2)) return false;
    t = Tz * rB2M[2][1] - Ty * rB2M[2][2]; t2 = e[1] * AR[2][2] + e[2] * AR[2][1] + BB[2]; if (greater(t, t2)) return false;
    t = Tx * rB2M[0][2] - Tz * rB2M[0][0]; t2 = e[0] * AR[0][2] + e[2] * AR[0][0] + BB[3]; if (greater(t, t2)) return false;
    t = Tx * rB2M[1][2] - Tz * rB2M[1][0]; t2 = e[0] * AR[1][2] + e[2] * AR[1][0] + BB[4]; if (greater(t, t2)) return false;
    t = Tx * rB2M[2][2] - Tz * rB2M[2][0]; t2 = e[0] * AR[2][2] + e[2] * AR[2][0] + BB[5]; if (greater(t, t2)) return false;
    t = Ty * rB2M[0][0] - Tx * rB2M[0][1]; t2 = e[0] * AR[0][1] + e[1] * AR[0][0] + BB[6]; if (greater(t, t2)) return false;
    t = Ty * rB2M[1][0] - Tx * rB2M[1][1]; t2 = e[0] * AR[1][1] + e[1] * AR[1][0] + BB[7]; if (greater(t, t2)) return false;
    t = Ty * rB2M[2][0] - Tx * rB2M[2][1]; t2 = e[0] * AR[2][1] + e[1] * AR[2][0] + BB[8]; if (greater(t, t2)) return false;
    return true;
  }
  OB_HD bool contains(const ObBvNode &n) const {   // OBBContainsBox
    const float *bc = n.c, *be = n.e;
    const float NCx = bc[0] * rM2B[0][0] + bc[1] * rM2B[1][0] + bc[2] * rM2B[2][0];
    const float NEx = fabsf(rM2B[0][0] * be[0]) + fabsf(rM2B[1][0] * be[1]) + fabsf(rM2B[2][0] * be[2]);
    if (B0[0] < NCx + NEx) return false;
    if (B1[0] > NCx - NEx) return false;
    const float NCy = bc[0] * rM2B[0][1] + bc[1] * rM2B[1][1] + bc[2] * rM2B[2][1];
    const float NEy = fabsf(rM2B[0][1] * be[0]) + fabsf(rM2B[1][1] * be[1]) + fabsf(rM2B[2][1] * be[2]);
    if (B0[1] < NCy + NEy) return false;
    if (B1[1] > NCy - NEy) return false;
    const float NCz = bc[0] * rM2B[0][2] + bc[1] * rM2B[1][2] + bc[2] * rM2B[2][2];
    const float NEz = fabsf(rM2B[0][2] * be[0]) + fabsf(rM2B[1][2] * be[1]) + fabsf(rM2B[2][2] * be[2]);
    if (B0[2] < NCz + NEz) return false;
    if (B1[2] > NCz - NEz) return false;
    return true;
  }
  OB_HD static float min3(float a, float b, float c) { return (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c); }
  OB_HD static float max3(float a, float b, float c) { return (a > b) ? ((a > c) ? a : c) : ((b > c) ? b : c); }
  OB_HD bool plane_box(const float *normal, float d) const {   // planeBoxOverlap
    float vmin[3], vmax[3];
    for (int q = 0; q <= 2; q++) {
      if (normal[q] > 0.0f) { vmin[q] = -ext[q]; vmax[q] = ext[q]; }
      else { vmin[q] = ext[q]; vmax[q] = -ext[q]; }
    }
    if ((normal[0] * vmin[0] + normal[1] * vmin[1] + normal[2] * vmin[2]) + d > 0.0f) return false;
    if ((normal[0] * vmax[0] + normal[1] * vmax[1] + normal[2] * vmax[2]) + d >= 0.0f) return true;
    return false;
  }
  // OBB_PRIM: triangle to box space, then OBBCollider::TriBoxOverlap
  OB_HD bool prim(const ObMeshDev &m, int tri) const {
    float v[3][3];
    for (int i = 0; i < 3; i++) {
      const float *s = m.verts + 3 * (size_t)m.tris[3 * (size_t)tri + i];
      v[i][0] = tM2B[0] + s[0] * rM2B[0][0] + s[1] * rM2B[1][0] + s[2] * rM2B[2][0];
      v[i][1] = tM2B[1] + s[0] * rM2B[0][1] + s[1] * rM2B[1][1] + s[2] * rM2B[2][1];
      v[i][2] = tM2B[2] + s[0] * rM2B[0][2] + s[1] * rM2B[1][2] + s[2] * rM2B[2][2];
    }
    const float *v0 = v[0], *v1 = v[1], *v2 = v[2];
    if (min3(v0[0], v1[0], v2[0]) > ext[0]) return false;
    if (max3(v0[0], v1[0], v2[0]) < -ext[0]) return false;
    if (min3(v0[1], v1[1], v2[1]) > ext[1]) return false;
    if (max3(v0[1], v1[1], v2[1]) < -ext[1]) return false;
    if (min3(v0[2], v1[2], v2[2]) > ext[2]) return false;
    if (max3(v0[2], v1[2], v2[2]) < -ext[2]) return false;
    const float e0[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
    const float e1[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
    const float normal[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
    const float d = (-normal[0]) * v0[0] + (-normal[1]) * v0[1] + (-normal[2]) * v0[2];
    if (!plane_box(normal, d)) return false;
    float rad, mn, mx;
#define OB_AX(EXPR_MIN, EXPR_MAX, RAD) mn = EXPR_MIN; mx = EXPR_MAX; if (mn > mx) { const float tmp = mx; mx = mn; mn = tmp; } rad = RAD; if (mn > rad || mx < -rad) return false;
#define OB_X01(a, b, fa, fb) OB_AX(a * v0[1] - b * v0[2], a * v2[1] - b * v2[2], fa * ext[1] + fb * ext[2])
#define OB_X2(a, b, fa, fb) OB_AX(a * v0[1] - b * v0[2], a * v1[1] - b * v1[2], fa * ext[1] + fb * ext[2])
#define OB_Y02(a, b, fa, fb) OB_AX(b * v0[2] - a * v0[0], b * v2[2] - a * v2[0], fa * ext[0] + fb * ext[2])
#define OB_Y1(a, b, fa, fb) OB_AX(b * v0[2] - a * v0[0], b * v1[2] - a * v1[0], fa * ext[0] + fb * ext[2])
#define OB_Z12(a, b, fa, fb) OB_AX(a * v1[0] - b * v1[1], a * v2[0] - b * v2[1], fa * ext[0] + fb * ext[1])
#define OB_Z0(a, b, fa, fb) OB_AX(a * v0[0] - b * v0[1], a * v1[0] - b * v1[1], fa * ext[0] + fb * ext[1])
    const float fey0 = fabsf(e0[1]), fez0 = fabsf(e0[2]);
    OB_X01(e0[2], e0[1], fez0, fey0)
    const float fex0 = fabsf(e0[0]);
    OB_Y02(e0[2], e0[0], fez0, fex0)
    OB_Z12(e0[1], e0[0], fey0, fex0)
    const float fey1 = fabsf(e1[1]), fez1 = fabsf(e1[2]);
    OB_X01(e1[2], e1[1], fez1, fey1)
    const float fex1 = fabsf(e1[0]);
    OB_Y02(e1[2], e1[0], fez1, fex1)
    OB_Z0(e1[1], e1[0], fey1, fex1)
    const float e2[3] = {v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2]};
    const float fey2 = fabsf(e2[1]), fez2 = fabsf(e2[2]);
    OB_X2(e2[2], e2[1], fez2, fey2)
    const float fex2 = fabsf(e2[0]);
    OB_Y1(e2[2], e2[0], fez2, fex2)
    OB_Z12(e2[1], e2[0], fey2, fex2)
#undef OB_AX
#undef OB_X01
#undef OB_X2
#undef OB_Y02
#undef OB_Y1
#undef OB_Z12
#undef OB_Z0
    return true;
  }
};

OB_HD void ob_obb_query_init(ObObbQuery &q, const real *boxpos, const real *boxR, const real *halfsize, const real *meshpos, const real *meshR) {
  for (int k = 0; k < 3; k++) q.ext[k] = (float)halfsize[k];
  ObMat44f WorldB, worldm, InvWorldB, InvWorldM, BtoM, MtoB;
  ob_m44_from_pr(boxpos, boxR, &WorldB);   // Box.mRot (+ centre as translation) has the same layout as MakeMatrix
  ob_m44_from_pr(meshpos, meshR, &worldm);
  ob_m44_invert_pr(WorldB, &InvWorldB);
  ob_m44_invert_pr(worldm, &InvWorldM);
  ob_m44_mul(WorldB, InvWorldM, &BtoM);
  ob_m44_mul(worldm, InvWorldB, &MtoB);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { q.rM2B[i][j] = MtoB.m[i][j]; q.rB2M[i][j] = BtoM.m[i][j]; }
  for (int j = 0; j < 3; j++) { q.tM2B[j] = MtoB.m[3][j]; q.tB2M[j] = BtoM.m[3][j]; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) q.AR[i][j] = 1e-6f + fabsf(q.rB2M[i][j]);
  for (int k = 0; k < 3; k++) { q.B0[k] = q.ext[k] - q.tM2B[k]; q.B1[k] = -q.ext[k] - q.tM2B[k]; }
  const float *E = q.ext;
  q.BBx1 = E[0] * q.AR[0][0] + E[1] * q.AR[1][0] + E[2] * q.AR[2][0];
  q.BBy1 = E[0] * q.AR[0][1] + E[1] * q.AR[1][1] + E[2] * q.AR[2][1];
  q.BBz1 = E[0] * q.AR[0][2] + E[1] * q.AR[1][2] + E[2] * q.AR[2][2];
  q.BB[0] = E[1] * q.AR[2][0] + E[2] * q.AR[1][0];
  q.BB[1] = E[0] * q.AR[2][0] + E[2] * q.AR[0][0];
  q.BB[2] = E[0] * q.AR[1][0] + E[1] * q.AR[0][0];
  q.BB[3] = E[1] * q.AR[2][1] + E[2] * q.AR[1][1];
  q.BB[4] = E[0] * q.AR[2][1] + E[2] * q.AR[0][1];
  q.BB[5] = E[0] * q.AR[1][1] + E[1] * q.AR[0][1];
  q.BB[6] = E[1] * q.AR[2][2] + E[2] * q.AR[1][2];
  q.BB[7] = E[0] * q.AR[2][2] + E[2] * q.AR[0][2];
  q.BB[8] = E[0] * q.AR[1][2] + E[1] * q.AR[0][2];
}

// OB_EPSILON (dEpsilon) comes from ob_math.h
#if defined(dSINGLE)
#define OB_MAXVALUE 3.402823466e+38f
#else
#define OB_MAXVALUE 1.7976931348623157e+308
#endif
#define OB_CONTACTS_UNIMPORTANT 0x80000000u

struct ObBtlData {   // sTrimeshBoxColliderData
  real boxR[12], boxPos[3], half[3];
  real bestNormal[3], bestDepth;
  int bestAxis;
  real E0[3], E1[3], E2[3], N[3];
  unsigned flags;
  ObCg *contacts;
  int ct;
};

OB_HD real ob_len3(const real *a) { return ob_sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
OB_HD void ob_getcol(const real *R, int a, real *v) { v[0] = R[a]; v[1] = R[4 + a]; v[2] = R[8 + a]; }

// GenerateContact, collision_trimesh_box.cpp:1357-1446
OB_HD void ob_btl_generate_contact(ObBtlData &D, int tri, const real *pos, const real *normal, real depth) {
  const int maxc = (int)(D.flags & 0xffffu);
  if (!(D.flags & OB_CONTACTS_UNIMPORTANT)) {
    bool duplicate = false;
    for (int i = 0; i < D.ct; i++) {
      ObCg *c = D.contacts + i;
      real diff[3];
      for (int j = 0; j < 3; j++) diff[j] = pos[j] - c->pos[j];
      if (ob_dot(diff, diff) < (real)OB_EPSILON) {
        if (OB_REAL(1.0) - ob_fabs(ob_dot(normal, c->normal)) < (real)OB_EPSILON) {
          if (depth > c->depth) c->depth = depth;
          duplicate = true;
        }
      }
    }
    if (duplicate || D.ct == maxc) return;
  }
  ObCg *c = D.contacts + D.ct;
  for (int j = 0; j < 3; j++) { c->pos[j] = pos[j]; c->normal[j] = normal[j]; }
  c->depth = depth;
  c->side1 = tri; c->side2 = -1;
  D.ct++;
}

OB_HD bool ob_btl_test_normal(ObBtlData &D, real fp0, real fR, const real *vNormal, int iAxis) {
  real fDepth = fR + fp0;
  if (fDepth < 0) return false;
  const real fLength = ob_len3(vNormal);
  if (fLength > 0.0f) {
    const real fOneOverLength = 1.0f / fLength;
    fDepth = fDepth * fOneOverLength;
    if (fDepth < D.bestDepth) {
      D.bestNormal[0] = -vNormal[0] * fOneOverLength;
      D.bestNormal[1] = -vNormal[1] * fOneOverLength;
      D.bestNormal[2] = -vNormal[2] * fOneOverLength;
      D.bestAxis = iAxis;
      D.bestDepth = fDepth;
    }
  }
  return true;
}
OB_HD bool ob_btl_test_face(ObBtlData &D, real fp0, real fp1, real fp2, real fR, real *vNormal, int iAxis) {
  real fMin, fMax;
  if (fp0 < fp1) { if (fp0 < fp2) fMin = fp0; else fMin = fp2; }
  else { if (fp1 < fp2) fMin = fp1; else fMin = fp2; }
  if (fp0 > fp1) { if (fp0 > fp2) fMax = fp0; else fMax = fp2; }
  else { if (fp1 > fp2) fMax = fp1; else fMax = fp2; }
  const real fDepthMin = fR - fMin, fDepthMax = fMax + fR;
  if (fDepthMin < 0 || fDepthMax < 0) return false;
  real fDepth;
  if (fDepthMin > fDepthMax) { fDepth = fDepthMax; vNormal[0] = -vNormal[0]; vNormal[1] = -vNormal[1]; vNormal[2] = -vNormal[2]; }
  else fDepth = fDepthMin;
  if (fDepth < D.bestDepth) {
    D.bestNormal[0] = vNormal[0]; D.bestNormal[1] = vNormal[1]; D.bestNormal[2] = vNormal[2];
    D.bestAxis = iAxis;
    D.bestDepth = fDepth;
  }
  return true;
}
OB_HD bool ob_btl_test_edge(ObBtlData &D, real fp0, real fp1, real fR, real *vNormal, int iAxis) {
  real fMin, fMax;
  fMin = vNormal[0] * vNormal[0] + vNormal[1] * vNormal[1] + vNormal[2] * vNormal[2];
  if (fMin <= (real)OB_EPSILON) return true;
  if (fp0 < fp1) { fMin = fp0; fMax = fp1; } else { fMin = fp1; fMax = fp0; }
  const real fDepthMin = fR - fMin, fDepthMax = fMax + fR;
  if (fDepthMin < 0 || fDepthMax < 0) return false;
  real fDepth;
  if (fDepthMin > fDepthMax) { fDepth = fDepthMax; vNormal[0] = -vNormal[0]; vNormal[1] = -vNormal[1]; vNormal[2] = -vNormal[2]; }
  else fDepth = fDepthMin;
  const real fLength = ob_len3(vNormal);
  if (fLength > 0.0f) {
    const real fOneOverLength = 1.0f / fLength;
    fDepth = fDepth * fOneOverLength;
    if (fDepth * 1.5f < D.bestDepth) {
      D.bestNormal[0] = vNormal[0] * fOneOverLength;
      D.bestNormal[1] = vNormal[1] * fOneOverLength;
      D.bestNormal[2] = vNormal[2] * fOneOverLength;
      D.bestAxis = iAxis;
      D.bestDepth = fDepth;
    }
  }
  return true;
}

OB_HD bool ob_btl_separating_axes(ObBtlData &D, const real *v0, const real *v1, const real *v2) {
  D.bestAxis = 0;
  D.bestDepth = (real)OB_MAXVALUE;
  for (int k = 0; k < 3; k++) { D.E0[k] = v1[k] - v0[k]; D.E1[k] = v2[k] - v0[k]; }
  for (int k = 0; k < 3; k++) D.E2[k] = D.E1[k] - D.E0[k];
  ob_cross(D.N, D.E0, D.E1);
  const real fNLen = ob_len3(D.N);
  if (!fNLen) return false;
  real vA[3][3];
  ob_getcol(D.boxR, 0, vA[0]); ob_getcol(D.boxR, 1, vA[1]); ob_getcol(D.boxR, 2, vA[2]);
  const real fa[3] = {D.half[0], D.half[1], D.half[2]};
  real vD[3] = {v0[0] - D.boxPos[0], v0[1] - D.boxPos[1], v0[2] - D.boxPos[2]};
  real vL[3], fp0, fp1, fp2, fR;
  // axis 1: triangle normal
  vL[0] = D.N[0]; vL[1] = D.N[1]; vL[2] = D.N[2];
  fp0 = ob_dot(vL, vD);
  fR = fa[0] * ob_fabs(ob_dot(D.N, vA[0])) + fa[1] * ob_fabs(ob_dot(D.N, vA[1])) + fa[2] * ob_fabs(ob_dot(D.N, vA[2]));
  if (!ob_btl_test_normal(D, fp0, fR, vL, 1)) return false;
  // axes 2-4: box faces
  for (int a = 0; a < 3; a++) {
    vL[0] = vA[a][0]; vL[1] = vA[a][1]; vL[2] = vA[a][2];
    fp0 = ob_dot(vL, vD);
    fp1 = fp0 + ob_dot(vA[a], D.E0);
    fp2 = fp0 + ob_dot(vA[a], D.E1);
    fR = fa[a];
    if (!ob_btl_test_face(D, fp0, fp1, fp2, fR, vL, 2 + a)) return false;
  }
  // axes 5-13: box axis x triangle edge
  for (int a = 0; a < 3; a++) {
    const int b = a == 0 ? 1 : 0, c = a == 2 ? 1 : 2;   // the two other box axes (b < c)
    const real *Es[3] = {D.E0, D.E1, D.E2};
    for (int e = 0; e < 3; e++) {
      ob_cross(vL, vA[a], Es[e]);
      fp0 = ob_dot(vL, vD);
      const real an = ob_dot(vA[a], D.N);
      real q0, q1;
      if (e == 0) { q0 = fp0; q1 = fp0 + an; }          // (fp1|fp0, fp2) = (fp0, fp0 + A.N)
      else { q0 = fp0; q1 = fp0 - an; }                 // edge 1: (fp0, fp0 - A.N); edge 2: (fp0 - A.N twice -> fp0, fp1)
      fR = fa[b] * ob_fabs(ob_dot(vA[c], Es[e])) + fa[c] * ob_fabs(ob_dot(vA[b], Es[e]));
      if (!ob_btl_test_edge(D, q0, q1, fR, vL, 5 + 3 * a + e)) return false;
    }
  }
  return true;
}

// _cldClipPolyToPlane, :331-381
OB_HD void ob_btl_clip_poly(const real (*in)[3], int ctIn, real (*out)[3], int *ctOut, const real *pl) {
  int n = 0;
  int i0 = ctIn - 1;
  for (int i1 = 0; i1 < ctIn; i0 = i1, i1++) {
    const real d0 = (pl[0] * in[i0][0] + pl[1] * in[i0][1] + pl[2] * in[i0][2] + pl[3]);
    const real d1 = (pl[0] * in[i1][0] + pl[1] * in[i1][1] + pl[2] * in[i1][2] + pl[3]);
    if (d0 >= 0) { out[n][0] = in[i0][0]; out[n][1] = in[i0][1]; out[n][2] = in[i0][2]; n++; }
    if ((d0 > 0 && d1 < 0) || (d0 < 0 && d1 > 0)) {
      out[n][0] = in[i0][0] - (in[i0][0] - in[i1][0]) * d0 / (d0 - d1);
      out[n][1] = in[i0][1] - (in[i0][1] - in[i1][1]) * d0 / (d0 - d1);
      out[n][2] = in[i0][2] - (in[i0][2] - in[i1][2]) * d0 / (d0 - d1);
      n++;
    }
  }
  *ctOut = n;
}

OB_HD bool ob_btl_done(const ObBtlData &D) {
  return (((unsigned)D.ct | OB_CONTACTS_UNIMPORTANT) == (D.flags & (0xffffu | OB_CONTACTS_UNIMPORTANT)));
}

// _cldClipping, :676-1053
OB_HD void ob_btl_clipping(ObBtlData &D, const real *v0, const real *v1, const real *v2, int tri) {
  if (D.bestAxis > 4) {
    real vub[3], vPb[3], vPa[3] = {D.boxPos[0], D.boxPos[1], D.boxPos[2]};
    for (int i = 0; i < 3; i++) {
      real col[3];
      ob_getcol(D.boxR, i, col);
      const real fSign = ob_dot(D.bestNormal, col) > 0 ? 1.0f : -1.0f;
      vPa[0] += fSign * D.half[i] * col[0];
      vPa[1] += fSign * D.half[i] * col[1];
      vPa[2] += fSign * D.half[i] * col[2];
    }
    const int iEdge = (D.bestAxis - 5) % 3;
    if (iEdge == 0) { for (int k = 0; k < 3; k++) { vPb[k] = v0[k]; vub[k] = D.E0[k]; } }
    else if (iEdge == 1) { for (int k = 0; k < 3; k++) { vPb[k] = v2[k]; vub[k] = D.E1[k]; } }
    else { for (int k = 0; k < 3; k++) { vPb[k] = v1[k]; vub[k] = D.E2[k]; } }
    ob_safe_normalize3(vub);
    real vua[3];
    ob_getcol(D.boxR, (D.bestAxis - 5) / 3, vua);
    // _cldClosestPointOnTwoLines, :644-670
    real fParam1, fParam2;
    {
      real vp[3] = {vPb[0] - vPa[0], vPb[1] - vPa[1], vPb[2] - vPa[2]};
      const real fuaub = ob_dot(vua, vub), fq1 = ob_dot(vua, vp), fq2 = -ob_dot(vub, vp);
      real fd = 1.0f - fuaub * fuaub;
      if (fd > 0.0f) { fd = 1.0f / fd; fParam1 = (fq1 + fuaub * fq2) * fd; fParam2 = (fuaub * fq1 + fq2) * fd; }
      else { fParam1 = 0.0f; fParam2 = 0.0f; }
    }
    for (int k = 0; k < 3; k++) { vPa[k] += vua[k] * fParam1; vPb[k] += vub[k] * fParam2; }
    real vPnt[3] = {vPa[0] + vPb[0], vPa[1] + vPb[1], vPa[2] + vPb[2]};
    vPnt[0] *= 0.5f; vPnt[1] *= 0.5f; vPnt[2] *= 0.5f;
    ob_btl_generate_contact(D, tri, vPnt, D.bestNormal, D.bestDepth);
  } else if (D.bestAxis == 1) {
    real vNormal2[3] = {-D.bestNormal[0], -D.bestNormal[1], -D.bestNormal[2]};
    const real *R = D.boxR;
    real vNr[3];
    vNr[0] = R[0] * vNormal2[0] + R[4] * vNormal2[1] + R[8] * vNormal2[2];
    vNr[1] = R[1] * vNormal2[0] + R[5] * vNormal2[1] + R[9] * vNormal2[2];
    vNr[2] = R[2] * vNormal2[0] + R[6] * vNormal2[1] + R[10] * vNormal2[2];
    const real an[3] = {ob_fabs(vNr[0]), ob_fabs(vNr[1]), ob_fabs(vNr[2])};
    int iB0, iB1, iB2;
    if (an[1] > an[0]) {
      if (an[1] > an[2]) { iB1 = 0; iB0 = 1; iB2 = 2; } else { iB1 = 0; iB2 = 1; iB0 = 2; }
    } else {
      if (an[0] > an[2]) { iB0 = 0; iB1 = 1; iB2 = 2; } else { iB1 = 0; iB2 = 1; iB0 = 2; }
    }
    real vCenter[3], col[3], col2[3];
    ob_getcol(R, iB0, col);
    if (vNr[iB0] > 0) { for (int k = 0; k < 3; k++) vCenter[k] = D.boxPos[k] - v0[k] - D.half[iB0] * col[k]; }
    else { for (int k = 0; k < 3; k++) vCenter[k] = D.boxPos[k] - v0[k] + D.half[iB0] * col[k]; }
    real avPoints[4][3];
    ob_getcol(R, iB1, col);
    ob_getcol(R, iB2, col2);
    for (int x = 0; x < 3; x++) {
      avPoints[0][x] = vCenter[x] + (D.half[iB1] * col[x]) - (D.half[iB2] * col2[x]);
      avPoints[1][x] = vCenter[x] - (D.half[iB1] * col[x]) - (D.half[iB2] * col2[x]);
      avPoints[2][x] = vCenter[x] - (D.half[iB1] * col[x]) + (D.half[iB2] * col2[x]);
      avPoints[3][x] = vCenter[x] + (D.half[iB1] * col[x]) + (D.half[iB2] * col2[x]);
    }
    real t1[9][3], t2[9][3], pl[4], vTemp[3], vTemp2[3];
    int c1 = 0, c2 = 0;
    vTemp[0] = -D.N[0]; vTemp[1] = -D.N[1]; vTemp[2] = -D.N[2];
    ob_safe_normalize3(vTemp);
    pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = 0;
    ob_btl_clip_poly(avPoints, 4, t1, &c1, pl);
    for (int k = 0; k < 3; k++) vTemp2[k] = v1[k] - v0[k];
    ob_cross(vTemp, D.N, vTemp2);
    ob_safe_normalize3(vTemp);
    pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = 0;
    ob_btl_clip_poly(t1, c1, t2, &c2, pl);
    for (int k = 0; k < 3; k++) vTemp2[k] = v2[k] - v1[k];
    ob_cross(vTemp, D.N, vTemp2);
    ob_safe_normalize3(vTemp);
    for (int k = 0; k < 3; k++) vTemp2[k] = v0[k] - v2[k];
    pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = ob_dot(vTemp2, vTemp);
    ob_btl_clip_poly(t2, c2, t1, &c1, pl);
    for (int k = 0; k < 3; k++) vTemp2[k] = v0[k] - v2[k];
    ob_cross(vTemp, D.N, vTemp2);
    ob_safe_normalize3(vTemp);
    pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = 0;
    ob_btl_clip_poly(t1, c1, t2, &c2, pl);
    for (int i = 0; i < c2; i++) {
      real fTempDepth = ob_dot(vNormal2, t2[i]);
      if (fTempDepth > 0) fTempDepth = 0;
      real vPnt[3] = {t2[i][0] + v0[0], t2[i][1] + v0[1], t2[i][2] + v0[2]};
      ob_btl_generate_contact(D, tri, vPnt, D.bestNormal, -fTempDepth);
      if (ob_btl_done(D)) break;
    }
  } else {
    real vNormal2[3] = {D.bestNormal[0], D.bestNormal[1], D.bestNormal[2]};
    int iA0 = D.bestAxis - 2, iA1, iA2;
    if (iA0 == 0) { iA1 = 1; iA2 = 2; } else if (iA0 == 1) { iA1 = 0; iA2 = 2; } else { iA1 = 0; iA2 = 1; }
    real avPoints[3][3];
    for (int k = 0; k < 3; k++) { avPoints[0][k] = v0[k] - D.boxPos[k]; avPoints[1][k] = v1[k] - D.boxPos[k]; avPoints[2][k] = v2[k] - D.boxPos[k]; }
    real t1[9][3], t2[9][3], pl[4], vTemp[3];
    int c1 = 0, c2 = 0;
    pl[0] = -vNormal2[0]; pl[1] = -vNormal2[1]; pl[2] = -vNormal2[2]; pl[3] = D.half[iA0];
    ob_btl_clip_poly(avPoints, 3, t1, &c1, pl);
    ob_getcol(D.boxR, iA1, vTemp);
    pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = D.half[iA1];
    ob_btl_clip_poly(t1, c1, t2, &c2, pl);
    pl[0] = -vTemp[0]; pl[1] = -vTemp[1]; pl[2] = -vTemp[2]; pl[3] = D.half[iA1];
    ob_btl_clip_poly(t2, c2, t1, &c1, pl);
    ob_getcol(D.boxR, iA2, vTemp);
    pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = D.half[iA2];
    ob_btl_clip_poly(t1, c1, t2, &c2, pl);
    pl[0] = -vTemp[0]; pl[1] = -vTemp[1]; pl[2] = -vTemp[2]; pl[3] = D.half[iA2];
    ob_btl_clip_poly(t2, c2, t1, &c1, pl);
    for (int i = 0; i < c1; i++) {
      real fTempDepth = ob_dot(vNormal2, t1[i]) - D.half[iA0];
      if (fTempDepth > 0) fTempDepth = 0;
      real vPnt[3] = {t1[i][0] + D.boxPos[0], t1[i][1] + D.boxPos[1], t1[i][2] + D.boxPos[2]};
      ob_btl_generate_contact(D, tri, vPnt, D.bestNormal, -fTempDepth);
      if (ob_btl_done(D)) break;
    }
  }
}

// dCollideBTL: o1 = trimesh, o2 = box
OB_HD int ob_collide_trimesh_box(const ObPose &o1, const ObPose &o2, const ObMeshDev &m, int flags, ObCg *contact, int *bverr) {
  ObBtlData D;
  for (int k = 0; k < 12; k++) D.boxR[k] = o2.R[k];
  for (int k = 0; k < 3; k++) { D.boxPos[k] = o2.pos[k]; D.half[k] = o2.p[k]; }
  D.half[0] *= 0.5f; D.half[1] *= 0.5f; D.half[2] *= 0.5f;
  D.flags = (unsigned)flags; D.contacts = contact; D.ct = 0;
  D.bestDepth = (real)OB_MAXVALUE; D.bestAxis = 0;
  D.bestNormal[0] = D.bestNormal[1] = D.bestNormal[2] = 0;
  ObObbQuery q;
  ob_obb_query_init(q, o2.pos, o2.R, D.half, o1.pos, o1.R);
  ObBvIter it;
  ob_bv_begin(it);
  // walk and triangle test in lock-step over the lanes that came here together (ob_math.h: OB_ALL_LANES), as in the sphere collider
  const unsigned together = OB_LANES_TOGETHER();
  bool fin = false;
  for (;;) {
    int tri = -1;
    if (!fin) { tri = ob_bv_next(m, it, q); fin = tri < 0; }
    if (OB_ALL_LANES(together, fin)) break;
    if (tri >= 0) {
      real dv[3][3];
      ob_fetch_triangle(m, tri, o1.pos, o1.R, dv);
      if (ob_btl_separating_axes(D, dv[0], dv[1], dv[2]) && D.bestAxis != 0) ob_btl_clipping(D, dv[0], dv[1], dv[2], tri);
      if (ob_btl_done(D)) fin = true;
    }
  }
  if (it.overflow) *bverr = 1;
  return D.ct;
}

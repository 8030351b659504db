// ob_trimesh_capsule.h — dCollideCCTL (ode/src/collision_trimesh_ccylinder.cpp:1038-1110): capsule vs trimesh.
//
// OBB query around the capsule (extents radius, radius, length/2 + radius; :930-1000) through the same
// OPCODE-equivalent OBB collider the box uses (ob_trimesh_box.h), then per touched triangle, in visit
// order: one-sided test, the capsule's separating-axis search over up to 19 axes (:466-735, best axis
// = LARGEST negative depth), the capsule's axis segment shifted by the radius along the best normal
// clipped against the triangle plane and its three edge planes (:738-910) -> two contacts per triangle.
// The local contact list is finally de-duplicated (_OptimizeLocalContacts :238-264) and copied out.
// Edge / vertex use flags (dGeomTriMeshDataPreprocess, ob_trimesh_build.cpp) select the axes per triangle;
// without preprocessing every axis is tested (UseFlags == NULL -> kUseAll, :1089).
#pragma once
#include "ob_trimesh_box.h"

struct ObCctlData {   // sTrimeshCapsuleColliderData
  real capPos[3], capAxis[3], radius, size;
  real E0[3], E1[3], E2[3], N[3];
  real V0[3], V1[3], V2[3];
  real normal[3], bestDepth, bestCenter, bestrt;
  int bestAxis;
  int maxc;
  real lpos[OB_MAXC_LOCAL][3], lnormal[OB_MAXC_LOCAL][3], ldepth[OB_MAXC_LOCAL];
  int ltri[OB_MAXC_LOCAL], lflag[OB_MAXC_LOCAL];
  int ct;
};

// _cldClipEdgeToPlane :301-351
OB_HD bool ob_cctl_clip_edge(real *p0, real *p1, const real *pl) {
  const real d0 = pl[0] * p0[0] + pl[1] * p0[1] + pl[2] * p0[2] + pl[3];
  const real d1 = pl[0] * p1[0] + pl[1] * p1[1] + pl[2] * p1[2] + pl[3];
  if (d0 < 0 && d1 < 0) return false;
  else if (d0 > 0 && d1 > 0) return true;
  else if ((d0 > 0 && d1 < 0) || (d0 < 0 && d1 > 0)) {
    real ip[3];
    ip[0] = p0[0] - (p0[0] - p1[0]) * d0 / (d0 - d1);
    ip[1] = p0[1] - (p0[1] - p1[1]) * d0 / (d0 - d1);
    ip[2] = p0[2] - (p0[2] - p1[2]) * d0 / (d0 - d1);
    if (d0 < 0) { p0[0] = ip[0]; p0[1] = ip[1]; p0[2] = ip[2]; }
    else { p1[0] = ip[0]; p1[1] = ip[1]; p1[2] = ip[2]; }
    return true;
  }
  return true;
}

// _cldTestAxis :353-448
OB_HD bool ob_cctl_test_axis(ObCctlData &D, real *vAxis, int iAxis, bool bNoFlip) {
  const real fL = ob_len3(vAxis);
  if (fL < OB_REAL(1e-5)) return true;
  ob_safe_normalize3(vAxis);
  const real frc = ob_fabs(ob_dot(D.capAxis, vAxis)) * (D.size * OB_REAL(0.5) - D.radius) + D.radius;
  real afv[3];
  afv[0] = ob_dot(D.V0, vAxis); afv[1] = ob_dot(D.V1, vAxis); afv[2] = ob_dot(D.V2, vAxis);
  real fMin = (real)OB_MAXVALUE, fMax = -(real)OB_MAXVALUE;
  for (int i = 0; i < 3; i++) {
    if (afv[i] < fMin) fMin = afv[i];
    if (afv[i] > fMax) fMax = afv[i];
  }
  const real fCenter = (fMin + fMax) * OB_REAL(0.5);
  const real fTriangleRadius = (fMax - fMin) * OB_REAL(0.5);
  if (ob_fabs(fCenter) > (frc + fTriangleRadius)) return false;
  const real fDepth = ob_fabs(fCenter) - (frc + fTriangleRadius);
  if (fDepth > D.bestDepth) {
    D.bestDepth = fDepth; D.bestCenter = fCenter; D.bestrt = fTriangleRadius;
    D.normal[0] = vAxis[0]; D.normal[1] = vAxis[1]; D.normal[2] = vAxis[2];
    D.bestAxis = iAxis;
    if (fCenter < 0 && !bNoFlip) {
      D.normal[0] = -D.normal[0]; D.normal[1] = -D.normal[1]; D.normal[2] = -D.normal[2];
      D.bestCenter = -fCenter;
    }
  }
  return true;
}
// _CalculateAxis :451-463: r = ((v1 - v2) x v3) x v4
OB_HD void ob_cctl_axis(const real *v1, const real *v2, const real *v3, const real *v4, real *r) {
  real t1[3] = {v1[0] - v2[0], v1[1] - v2[1], v1[2] - v2[2]}, t2[3];
  ob_cross(t2, t1, v3);
  ob_cross(r, t2, v4);
}
OB_HD real ob_len2_3(const real *a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }

// _cldTestSeparatingAxesOfCapsule :466-735; flags = the triangle's use flags (edge e: 1<<e, vertex e: 8<<e)
OB_HDN bool ob_cctl_separating_axes(ObCctlData &D, const real *v0, const real *v1, const real *v2, unsigned flags) {
  const real hl = D.size * OB_REAL(0.5) - D.radius;
  real vCp0[3], vCp1[3];
  for (int k = 0; k < 3; k++) { vCp0[k] = D.capPos[k] + D.capAxis[k] * hl; vCp1[k] = D.capPos[k] - D.capAxis[k] * hl; }
  D.bestAxis = 0;
  D.bestDepth = -(real)OB_MAXVALUE;
  real vAxis[3] = {0, 0, 0};
  const real fEpsilon = (real)1e-6f;
  for (int k = 0; k < 3; k++) { D.V0[k] = v0[k] - D.capPos[k]; D.V1[k] = v1[k] - D.capPos[k]; D.V2[k] = v2[k] - D.capPos[k]; }
  vAxis[0] = -D.N[0]; vAxis[1] = -D.N[1]; vAxis[2] = -D.N[2];
  if (!ob_cctl_test_axis(D, vAxis, 1, true)) return false;
  const real *E[3] = {D.E0, D.E1, D.E2};
  const real *V[3] = {v0, v1, v2};
  for (int e = 0; e < 3; e++) {   // axes C x E_e
    if (!(flags & (1u << e))) continue;
    ob_cross(vAxis, D.capAxis, E[e]);
    if (ob_len2_3(vAxis) > fEpsilon) if (!ob_cctl_test_axis(D, vAxis, 2 + e, false)) return false;
  }
  for (int e = 0; e < 3; e++) {   // ((Cp0 - V_e) x E_e) x E_e
    if (!(flags & (1u << e))) continue;
    ob_cctl_axis(vCp0, V[e], E[e], E[e], vAxis);
    if (ob_len2_3(vAxis) > fEpsilon) if (!ob_cctl_test_axis(D, vAxis, 5 + e, false)) return false;
  }
  for (int e = 0; e < 3; e++) {   // ((Cp1 - V_e) x E_e) x E_e
    if (!(flags & (1u << e))) continue;
    ob_cctl_axis(vCp1, V[e], E[e], E[e], vAxis);
    if (ob_len2_3(vAxis) > fEpsilon) if (!ob_cctl_test_axis(D, vAxis, 8 + e, false)) return false;
  }
  for (int e = 0; e < 3; e++) {   // ((V_e - Cp0) x C) x C
    if (!(flags & (8u << e))) continue;
    ob_cctl_axis(V[e], vCp0, D.capAxis, D.capAxis, vAxis);
    if (ob_len2_3(vAxis) > fEpsilon) if (!ob_cctl_test_axis(D, vAxis, 11 + e, false)) return false;
  }
  for (int e = 0; e < 3; e++) {   // V_e - Cp0
    if (!(flags & (8u << e))) continue;
    vAxis[0] = V[e][0] - vCp0[0]; vAxis[1] = V[e][1] - vCp0[1]; vAxis[2] = V[e][2] - vCp0[2];
    if (ob_len2_3(vAxis) > fEpsilon) if (!ob_cctl_test_axis(D, vAxis, 14 + e, false)) return false;
  }
  for (int e = 0; e < 3; e++) {   // V_e - Cp1
    if (!(flags & (8u << e))) continue;
    vAxis[0] = V[e][0] - vCp1[0]; vAxis[1] = V[e][1] - vCp1[1]; vAxis[2] = V[e][2] - vCp1[2];
    if (ob_len2_3(vAxis) > fEpsilon) if (!ob_cctl_test_axis(D, vAxis, 17 + e, false)) return false;
  }
  return true;
}

// _cldTestOneTriangleVSCapsule :738-910
OB_HDN void ob_cctl_one_triangle(ObCctlData &D, const real *v0, const real *v1, const real *v2, unsigned flags) {
  for (int k = 0; k < 3; k++) { D.E0[k] = v1[k] - v0[k]; D.E1[k] = v2[k] - v1[k]; D.E2[k] = v0[k] - v2[k]; }
  real mE0[3] = {v0[0] - v1[0], v0[1] - v1[1], v0[2] - v1[2]};
  ob_cross(D.N, D.E1, mE0);
  if (!ob_safe_normalize3(D.N)) return;
  const real plDistance = -ob_dot(v0, D.N);
  const real dist = D.N[0] * D.capPos[0] + D.N[1] * D.capPos[1] + D.N[2] * D.capPos[2] + plDistance;
  if (dist < 0) return;   // capsule must be over the positive side of the triangle
  if (!ob_cctl_separating_axes(D, v0, v1, v2, flags)) return;
  if (D.bestAxis == 0) return;
  const real hl = D.size * OB_REAL(0.5) - D.radius;
  real ct[3], p0[3], p1[3];
  for (int k = 0; k < 3; k++) ct[k] = D.capPos[k] + D.normal[k] * D.radius;
  for (int k = 0; k < 3; k++) { p0[k] = ct[k] + D.capAxis[k] * hl; p1[k] = ct[k] - D.capAxis[k] * hl; }
  for (int k = 0; k < 3; k++) { p0[k] -= v0[k]; p1[k] -= v0[k]; }
  real pl[4], vTemp[3];
  pl[0] = -D.N[0]; pl[1] = -D.N[1]; pl[2] = -D.N[2]; pl[3] = 0;
  if (!ob_cctl_clip_edge(p0, p1, pl)) return;
  ob_cross(vTemp, D.N, D.E0);
  pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = OB_REAL(1e-5);
  if (!ob_cctl_clip_edge(p0, p1, pl)) return;
  ob_cross(vTemp, D.N, D.E1);
  pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = -(ob_dot(D.E0, vTemp) - OB_REAL(1e-5));
  if (!ob_cctl_clip_edge(p0, p1, pl)) return;
  ob_cross(vTemp, D.N, D.E2);
  pl[0] = vTemp[0]; pl[1] = vTemp[1]; pl[2] = vTemp[2]; pl[3] = OB_REAL(1e-5);
  if (!ob_cctl_clip_edge(p0, p1, pl)) return;
  for (int k = 0; k < 3; k++) { p0[k] += v0[k]; p1[k] += v0[k]; }
  for (int k = 0; k < 3; k++) vTemp[k] = p0[k] - D.capPos[k];
  real fDepth0 = ob_dot(vTemp, D.normal) - (D.bestCenter - D.bestrt);
  for (int k = 0; k < 3; k++) vTemp[k] = p1[k] - D.capPos[k];
  real fDepth1 = ob_dot(vTemp, D.normal) - (D.bestCenter - D.bestrt);
  if (fDepth0 < 0) fDepth0 = 0;
  if (fDepth1 < 0) fDepth1 = 0;
  D.ldepth[D.ct] = fDepth0;
  for (int k = 0; k < 3; k++) { D.lnormal[D.ct][k] = D.normal[k]; D.lpos[D.ct][k] = p0[k]; }
  D.lflag[D.ct] = 1;
  D.ct++;
  if (D.ct < D.maxc) {
    D.ldepth[D.ct] = fDepth1;
    for (int k = 0; k < 3; k++) { D.lnormal[D.ct][k] = D.normal[k]; D.lpos[D.ct][k] = p1[k]; }
    D.lflag[D.ct] = 1;
    D.ct++;
  }
}

// o1 = trimesh, o2 = capsule
OB_HD int ob_collide_trimesh_capsule(const ObPose &o1, const ObPose &o2, const ObMeshDev &m, int flags, ObCg *contact, int *bverr) {
  ObCctlData D;
  int maxc = flags & 0xffff;
  if (maxc > OB_MAXC_LOCAL) maxc = OB_MAXC_LOCAL;
  D.maxc = maxc;
  for (int k = 0; k < 3; k++) { D.capPos[k] = o2.pos[k]; D.capAxis[k] = o2.R[4 * k + 2]; D.normal[k] = 0; D.N[k] = 0; }
  D.radius = o2.p[0];
  D.size = o2.p[1];
  D.size += 2 * D.radius;
  D.ct = 0;
  D.bestDepth = -(real)OB_MAXVALUE; D.bestCenter = 0; D.bestrt = 0; D.bestAxis = 0;
  if (maxc < 1) return 0;
  ObObbQuery q;
  const real ext[3] = {D.radius, D.radius, D.size / 2};
  ob_obb_query_init(q, o2.pos, o2.R, ext, o1.pos, o1.R);
  ObBvIter it;
  ob_bv_begin(it);
  int ct0 = 0;
  for (;;) {
    const int tri = ob_bv_next(m, it, q);
    if (tri < 0) break;
    real dv[3][3];
    ob_fetch_triangle(m, tri, o1.pos, o1.R, dv);
    ob_cctl_one_triangle(D, dv[0], dv[1], dv[2], m.useflags ? (unsigned)m.useflags[tri] : 0xFFu);
    for (; ct0 < D.ct; ct0++) D.ltri[ct0] = tri;
    if (D.ct >= maxc) break;
  }
  if (it.overflow) *bverr = 1;
  if (D.ct == 0) return 0;
  // _ProcessLocalContacts :266-299 with _OptimizeLocalContacts :238-264
  if (D.ct > 1 && !((unsigned)flags & OB_CONTACTS_UNIMPORTANT)) {
    const real ePos = OB_REAL(0.0001), eN = OB_REAL(0.0001);
    for (int i = 0; i < D.ct - 1; i++)
      for (int j = i + 1; j < D.ct; j++) {
        const bool posNear = ob_fabs(D.lpos[i][0] - D.lpos[j][0]) < ePos && ob_fabs(D.lpos[i][1] - D.lpos[j][1]) < ePos && ob_fabs(D.lpos[i][2] - D.lpos[j][2]) < ePos;
        const bool sameDir = ob_fabs(D.lnormal[i][0] - D.lnormal[j][0]) < eN && ob_fabs(D.lnormal[i][1] - D.lnormal[j][1]) < eN && ob_fabs(D.lnormal[i][2] - D.lnormal[j][2]) < eN;
        if (posNear && sameDir) {
          if (D.ldepth[j] > D.ldepth[i]) D.lflag[i] = 0;
          else D.lflag[j] = 0;
        }
      }
  }
  int n = 0;
  for (int i = 0; i < D.ct; i++) {
    if (n >= maxc) break;
    if (D.lflag[i] == 1) {
      contact[n].depth = D.ldepth[i];
      for (int k = 0; k < 3; k++) { contact[n].normal[k] = D.lnormal[i][k]; contact[n].pos[k] = D.lpos[i][k]; }
      contact[n].side1 = D.ltri[i]; contact[n].side2 = -1;
      n++;
    }
  }
  return n;
}

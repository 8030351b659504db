// ob_collide.h — AABBs and primitive narrowphase, one geom pair per thread.
//
// Each function follows the reference collider it replaces expression by
// expression (same association order, same compare-branch structure) so that
// contact counts are exact and contact fields are bit-identical:
//   computeAABB      ode/src/sphere.cpp:59-67, box.cpp:60-77, plane.cpp:68-96, capsule.cpp:60-74
//   sphere-sphere    ode/src/sphere.cpp:110-128 + collision_util.cpp:38-67
//   sphere-box       ode/src/sphere.cpp:131-219
//   sphere-plane     ode/src/sphere.cpp:222-251
//   box-box          ode/src/box.cpp:331-742 (intersectRectQuad :187-238, cullPoints :249-311)
//   box-plane        ode/src/box.cpp:745-878
//   capsule-*        ode/src/capsule.cpp:130-411 + collision_util.cpp:107-391
//   dCollide         ode/src/collision_kernel.cpp:292-339 (class table + reverse fix-up :167-268)
// No virtual dispatch: the class pair selects a switch arm.
#pragma once
#include "ob_types.h"
#include "ob_collide_types.h"
#include "ob_trimesh.h"
#include "ob_trimesh_box.h"
#include "ob_trimesh_capsule.h"

OB_HD void ob_aabb(const ObPose &g, real *aabb, const ObMeshDev *meshes = 0) {
  switch (g.type) {
    case OB_GEOM_SPHERE: {
      real r = g.p[0];
      aabb[0] = g.pos[0] - r; aabb[1] = g.pos[0] + r;
      aabb[2] = g.pos[1] - r; aabb[3] = g.pos[1] + r;
      aabb[4] = g.pos[2] - r; aabb[5] = g.pos[2] + r;
    } break;
    case OB_GEOM_BOX: {
      const real *R = g.R; const real *s = g.p;
      real xr = OB_REAL(0.5) * (ob_fabs(R[0] * s[0]) + ob_fabs(R[1] * s[1]) + ob_fabs(R[2] * s[2]));
      real yr = OB_REAL(0.5) * (ob_fabs(R[4] * s[0]) + ob_fabs(R[5] * s[1]) + ob_fabs(R[6] * s[2]));
      real zr = OB_REAL(0.5) * (ob_fabs(R[8] * s[0]) + ob_fabs(R[9] * s[1]) + ob_fabs(R[10] * s[2]));
      aabb[0] = g.pos[0] - xr; aabb[1] = g.pos[0] + xr;
      aabb[2] = g.pos[1] - yr; aabb[3] = g.pos[1] + yr;
      aabb[4] = g.pos[2] - zr; aabb[5] = g.pos[2] + zr;
    } break;
    case OB_GEOM_CAPSULE: {
      // capsule.cpp:60-74: radius + |R(:,2)| * lz/2
      const real *R = g.R; real radius = g.p[0], lz = g.p[1];
      real xr = ob_fabs(R[2] * lz) * OB_REAL(0.5) + radius;
      real yr = ob_fabs(R[6] * lz) * OB_REAL(0.5) + radius;
      real zr = ob_fabs(R[10] * lz) * OB_REAL(0.5) + radius;
      aabb[0] = g.pos[0] - xr; aabb[1] = g.pos[0] + xr;
      aabb[2] = g.pos[1] - yr; aabb[3] = g.pos[1] + yr;
      aabb[4] = g.pos[2] - zr; aabb[5] = g.pos[2] + zr;
    } break;
    case OB_GEOM_CYLINDER: {
      // dxCylinder::computeAABB, cylinder.cpp:60-78
      const real *R = g.R; const real radius = g.p[0], lz = g.p[1];
      real xr = ob_fabs(R[0] * radius) + ob_fabs(R[1] * radius) + OB_REAL(0.5) * ob_fabs(R[2] * lz);
      real yr = ob_fabs(R[4] * radius) + ob_fabs(R[5] * radius) + OB_REAL(0.5) * ob_fabs(R[6] * lz);
      real zr = ob_fabs(R[8] * radius) + ob_fabs(R[9] * radius) + OB_REAL(0.5) * ob_fabs(R[10] * lz);
      aabb[0] = g.pos[0] - xr; aabb[1] = g.pos[0] + xr;
      aabb[2] = g.pos[1] - yr; aabb[3] = g.pos[1] + yr;
      aabb[4] = g.pos[2] - zr; aabb[5] = g.pos[2] + zr;
    } break;
    case OB_GEOM_PLANE: {
      const real *p = g.p;
      aabb[0] = -OB_INF; aabb[1] = OB_INF; aabb[2] = -OB_INF; aabb[3] = OB_INF; aabb[4] = -OB_INF; aabb[5] = OB_INF;
      if (p[1] == 0.0f && p[2] == 0.0f) {
        aabb[0] = (p[0] > 0) ? -OB_INF : -p[3];
        aabb[1] = (p[0] > 0) ? p[3] : OB_INF;
      } else if (p[0] == 0.0f && p[2] == 0.0f) {
        aabb[2] = (p[1] > 0) ? -OB_INF : -p[3];
        aabb[3] = (p[1] > 0) ? p[3] : OB_INF;
      } else if (p[0] == 0.0f && p[1] == 0.0f) {
        aabb[4] = (p[2] > 0) ? -OB_INF : -p[3];
        aabb[5] = (p[2] > 0) ? p[3] : OB_INF;
      }
    } break;
    case OB_GEOM_TRIMESH: {
      // dxTriMesh::computeAABB, collision_trimesh_opcode.cpp:668-692
      const ObMeshDev &d = meshes[g.mesh];
      const real *R = g.R;
      real c[3];
      ob_mul0_331(c, R, d.aabbc);
      real xrange = ob_fabs(R[0] * d.aabbe[0]) + ob_fabs(R[1] * d.aabbe[1]) + ob_fabs(R[2] * d.aabbe[2]);
      real yrange = ob_fabs(R[4] * d.aabbe[0]) + ob_fabs(R[5] * d.aabbe[1]) + ob_fabs(R[6] * d.aabbe[2]);
      real zrange = ob_fabs(R[8] * d.aabbe[0]) + ob_fabs(R[9] * d.aabbe[1]) + ob_fabs(R[10] * d.aabbe[2]);
      aabb[0] = c[0] + g.pos[0] - xrange; aabb[1] = c[0] + g.pos[0] + xrange;
      aabb[2] = c[1] + g.pos[1] - yrange; aabb[3] = c[1] + g.pos[1] + yrange;
      aabb[4] = c[2] + g.pos[2] - zrange; aabb[5] = c[2] + g.pos[2] + zrange;
    } break;
    case OB_GEOM_RAY: {
      // dxRay::computeAABB, ray.cpp:56-88
      const real length = g.p[0];
      for (int k = 0; k < 3; k++) {
        const real e = g.pos[k] + g.R[4 * k + 2] * length;
        if (g.pos[k] < e) { aabb[2 * k] = g.pos[k]; aabb[2 * k + 1] = e; }
        else { aabb[2 * k] = e; aabb[2 * k + 1] = g.pos[k]; }
      }
    } break;
    case OB_GEOM_SPACE:   // dxSpace::computeAABB (collision_space.cpp:116-137) ran on the host: the union of the members' boxes
      for (int k = 0; k < 6; k++) aabb[k] = g.R[k];
      break;
    default:
      aabb[0] = aabb[2] = aabb[4] = -OB_INF; aabb[1] = aabb[3] = aabb[5] = OB_INF;
  }
}

// ---- sphere colliders -----------------------------------------------------------
OB_HD int ob_collide_spheres(const real *p1, real r1, const real *p2, real r2, ObCg *c) {
  real t0 = p1[0] - p2[0], t1 = p1[1] - p2[1], t2 = p1[2] - p2[2];
  real d = ob_sqrt(t0 * t0 + t1 * t1 + t2 * t2);
  if (d > (r1 + r2)) return 0;
  if (d <= 0) {
    c->pos[0] = p1[0]; c->pos[1] = p1[1]; c->pos[2] = p1[2];
    c->normal[0] = 1; c->normal[1] = 0; c->normal[2] = 0;
    c->depth = r1 + r2;
  } else {
    real d1 = ob_recip(d);
    c->normal[0] = (p1[0] - p2[0]) * d1;
    c->normal[1] = (p1[1] - p2[1]) * d1;
    c->normal[2] = (p1[2] - p2[2]) * d1;
    real k = OB_REAL(0.5) * (r2 - r1 - d);
    c->pos[0] = p1[0] + c->normal[0] * k;
    c->pos[1] = p1[1] + c->normal[1] * k;
    c->pos[2] = p1[2] + c->normal[2] * k;
    c->depth = r1 + r2 - d;
  }
  return 1;
}

OB_HD int ob_collide_sphere_box(const ObPose &o1, const ObPose &o2, ObCg *contact) {
  real l[3], t[3], p[3], q[3], r[3];
  real depth;
  int onborder = 0;
  const real *R2 = o2.R;
  p[0] = o1.pos[0] - o2.pos[0];
  p[1] = o1.pos[1] - o2.pos[1];
  p[2] = o1.pos[2] - o2.pos[2];
  l[0] = o2.p[0] * OB_REAL(0.5);
  t[0] = ob_dot14(p, R2);
  if (t[0] < -l[0]) { t[0] = -l[0]; onborder = 1; }
  if (t[0] > l[0]) { t[0] = l[0]; onborder = 1; }
  l[1] = o2.p[1] * OB_REAL(0.5);
  t[1] = ob_dot14(p, R2 + 1);
  if (t[1] < -l[1]) { t[1] = -l[1]; onborder = 1; }
  if (t[1] > l[1]) { t[1] = l[1]; onborder = 1; }
  t[2] = ob_dot14(p, R2 + 2);
  l[2] = o2.p[2] * OB_REAL(0.5);
  if (t[2] < -l[2]) { t[2] = -l[2]; onborder = 1; }
  if (t[2] > l[2]) { t[2] = l[2]; onborder = 1; }
  if (!onborder) {
    real min_distance = l[0] - ob_fabs(t[0]);
    int mini = 0;
    for (int i = 1; i < 3; i++) {
      real face_distance = l[i] - ob_fabs(t[i]);
      if (face_distance < min_distance) { min_distance = face_distance; mini = i; }
    }
    contact->pos[0] = o1.pos[0]; contact->pos[1] = o1.pos[1]; contact->pos[2] = o1.pos[2];
    real tmp[3] = {0, 0, 0};
    real sgn = (t[mini] > 0) ? OB_REAL(1.0) : OB_REAL(-1.0);
    if (mini == 0) tmp[0] = sgn; else if (mini == 1) tmp[1] = sgn; else tmp[2] = sgn;
    ob_mul0_331(contact->normal, R2, tmp);
    contact->depth = min_distance + o1.p[0];
    return 1;
  }
  ob_mul0_331(q, R2, t);
  r[0] = p[0] - q[0]; r[1] = p[1] - q[1]; r[2] = p[2] - q[2];
  depth = o1.p[0] - ob_sqrt(ob_dot(r, r));
  if (depth < 0) return 0;
  contact->pos[0] = q[0] + o2.pos[0];
  contact->pos[1] = q[1] + o2.pos[1];
  contact->pos[2] = q[2] + o2.pos[2];
  contact->normal[0] = r[0]; contact->normal[1] = r[1]; contact->normal[2] = r[2];
  ob_safe_normalize3(contact->normal);
  contact->depth = depth;
  return 1;
}

OB_HD int ob_collide_sphere_plane(const ObPose &o1, const ObPose &o2, ObCg *contact) {
  const real *pl = o2.p;
  real k = ob_dot(o1.pos, pl);
  real depth = pl[3] - k + o1.p[0];
  if (depth >= 0) {
    contact->normal[0] = pl[0]; contact->normal[1] = pl[1]; contact->normal[2] = pl[2];
    contact->pos[0] = o1.pos[0] - pl[0] * o1.p[0];
    contact->pos[1] = o1.pos[1] - pl[1] * o1.p[0];
    contact->pos[2] = o1.pos[2] - pl[2] * o1.p[0];
    contact->depth = depth;
    return 1;
  }
  return 0;
}

// ---- box colliders --------------------------------------------------------------
// dLineClosestApproach, collision_util.cpp:70-92
OB_HD void ob_line_closest_approach(const real *pa, const real *ua, const real *pb, const real *ub,
                                    real *alpha, real *beta) {
  real p[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
  real uaub = ob_dot(ua, ub);
  real q1 = ob_dot(ua, p);
  real q2 = -ob_dot(ub, p);
  real d = 1 - uaub * uaub;
  if (d <= OB_REAL(0.0001)) { *alpha = 0; *beta = 0; }
  else {
    d = ob_recip(d);
    *alpha = (q1 + uaub * q2) * d;
    *beta = (uaub * q1 + q2) * d;
  }
}

// intersectRectQuad, box.cpp:187-238: the quad p[4] clipped against the rectangle |x| < h[0], |y| < h[1]; ret must
// hold 16 reals.  Same clip arithmetic and point order as the reference; the control flow is structured for a warp:
// the reference leaves all three loops with a goto when the 8th point is written, here a flag closes the loops so
// that the lanes of a warp (each clipping its own pair) meet again after every pass.
OB_HDN int ob_intersect_rect_quad(const real h[2], real p[8], real ret[16]) {
  real buffer[16];
  real *q = p, *r = ret;
  int nq = 4, nr = 0;
  bool full = false;
  for (int pass = 0; pass < 4 && !full; pass++) {
    const int dir = pass >> 1, sign = (pass & 1) ? 1 : -1;
    real *pr = r;
    nr = 0;
    for (int i = nq; i > 0 && !full; i--) {
      const real *pq = q + 2 * (nq - i);
      const bool inside = sign * pq[dir] < h[dir];
      if (inside) {
        pr[0] = pq[0]; pr[1] = pq[1];
        pr += 2; nr++;
        full = (nr & 8) != 0;
      }
      const real *nextq = (i > 1) ? pq + 2 : q;
      if (!full && (inside ^ (sign * nextq[dir] < h[dir]))) {
        pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
        pr[dir] = sign * h[dir];
        pr += 2; nr++;
        full = (nr & 8) != 0;
      }
    }
    q = r;
    if (!full) { r = (q == ret) ? buffer : ret; nq = nr; }
  }
  if (q != ret) for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
  return nr;
}

// cullPoints, box.cpp:249-311
OB_HDN void ob_cull_points(int n, const real p[], int m, int i0, int iret[]) {
  int i, j;
  real a, cx, cy, q;
  if (n == 1) { cx = p[0]; cy = p[1]; }
  else if (n == 2) { cx = OB_REAL(0.5) * (p[0] + p[2]); cy = OB_REAL(0.5) * (p[1] + p[3]); }
  else {
    a = 0; cx = 0; cy = 0;
    for (i = 0; i < (n - 1); i++) {
      q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
      a += q;
      cx += q * (p[i * 2] + p[i * 2 + 2]);
      cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
    }
    q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
    a = ob_recip(OB_REAL(3.0) * (a + q));
    cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
    cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
  }
  real A[8];
  for (i = 0; i < n; i++) A[i] = ob_atan2(p[i * 2 + 1] - cy, p[i * 2] - cx);
  int avail[8];
  for (i = 0; i < n; i++) avail[i] = 1;
  avail[i0] = 0;
  iret[0] = i0;
  int w = 1;
  for (j = 1; j < m; j++) {
    a = (real)((double)(real)j * (2 * OB_PI / m) + (double)A[i0]);
    if ((double)a > OB_PI) a -= (real)(2 * OB_PI);
    real maxdiff = OB_REAL(1e9), diff;
    int pick = i0;
    for (i = 0; i < n; i++) {
      if (avail[i]) {
        diff = ob_fabs(A[i] - a);
        if ((double)diff > OB_PI) diff = (real)(2 * OB_PI - (double)diff);
        if (diff < maxdiff) { maxdiff = diff; pick = i; }
      }
    }
    avail[pick] = 0;
    iret[w++] = pick;
  }
}

OB_HD real ob_sel3(real v0, real v1, real v2, int i) { return i == 0 ? v0 : (i == 1 ? v1 : v2); }

// dBoxBox, box.cpp:331-712.  Returns number of contacts (pos/depth filled), normal, depth, code.
//
// The reference scans its 15 candidate separating axes with a macro that returns from the middle of the function on
// a separating axis and breaks out on the first hit of an "unimportant" query.  One pair per lane, that control flow
// leaves the lanes of a warp on their own for the rest of the scan (ncu r02d: the scan ran with 1.2 active threads
// per instruction and was half of k_collide's issued instructions).  Here the scan is straight-line: every axis is
// evaluated with the reference's expressions in the reference's order and three flags carry what the returns and
// breaks did (sep: a separating axis was met while the scan was live; stop: the unimportant query took its first
// hit; neither: keep the deepest axis so far).  The edge axes keep the winning unnormalised axis and its length and
// normalise once after the scan -- the same three divisions the reference performs when that axis takes the lead.
// The values compared and stored are the reference's, so the result is bit-identical (tests: every box scene).
OB_HDN int ob_box_box(const real *p1, const real *R1, const real *side1, const real *p2, const real *R2,
                      const real *side2, real *normal, real *depth, int *return_code, int flags, ObCg *contact) {
  const real fudge_factor = OB_REAL(1.05);
  real p[3], pp[3];
  real R11, R12, R13, R21, R22, R23, R31, R32, R33, Q11, Q12, Q13, Q21, Q22, Q23, Q31, Q32, Q33;
  int i, j;
  const bool unimportant = (flags & 0x80000000u) != 0;

  p[0] = p2[0] - p1[0]; p[1] = p2[1] - p1[1]; p[2] = p2[2] - p1[2];
  ob_mul1_331(pp, R1, p);
  const real A0 = side1[0] * OB_REAL(0.5), A1 = side1[1] * OB_REAL(0.5), A2 = side1[2] * OB_REAL(0.5);
  const real B0 = side2[0] * OB_REAL(0.5), B1 = side2[1] * OB_REAL(0.5), B2 = side2[2] * OB_REAL(0.5);

  R11 = ob_dot44(R1 + 0, R2 + 0); R12 = ob_dot44(R1 + 0, R2 + 1); R13 = ob_dot44(R1 + 0, R2 + 2);
  R21 = ob_dot44(R1 + 1, R2 + 0); R22 = ob_dot44(R1 + 1, R2 + 1); R23 = ob_dot44(R1 + 1, R2 + 2);
  R31 = ob_dot44(R1 + 2, R2 + 0); R32 = ob_dot44(R1 + 2, R2 + 1); R33 = ob_dot44(R1 + 2, R2 + 2);
  Q11 = ob_fabs(R11); Q12 = ob_fabs(R12); Q13 = ob_fabs(R13);
  Q21 = ob_fabs(R21); Q22 = ob_fabs(R22); Q23 = ob_fabs(R23);
  Q31 = ob_fabs(R31); Q32 = ob_fabs(R32); Q33 = ob_fabs(R33);

  real s = -OB_INF;
  int code = 0;
  bool invert_normal = false, sep = false, stop = false;
  real en1 = 0, en2 = 0, en3 = 0, el = 1;   // leading edge axis (unnormalised) and its length
  // face axes (box.cpp:389-411)
#define OB_TST(expr1, expr2, cc)                                   \
  {                                                                \
    const real e_ = (expr1);                                       \
    const real s2_ = ob_fabs(e_) - (expr2);                        \
    const bool live_ = !(sep | stop), out_ = s2_ > 0;              \
    sep |= live_ & out_;                                           \
    const bool upd_ = live_ & !out_ & (s2_ > s);                   \
    if (upd_) { s = s2_; invert_normal = e_ < 0; code = (cc); }    \
    stop |= upd_ & unimportant;                                    \
  }
  OB_TST(pp[0], (A0 + B0 * Q11 + B1 * Q12 + B2 * Q13), 1);
  OB_TST(pp[1], (A1 + B0 * Q21 + B1 * Q22 + B2 * Q23), 2);
  OB_TST(pp[2], (A2 + B0 * Q31 + B1 * Q32 + B2 * Q33), 3);
  OB_TST(ob_dot41(R2 + 0, p), (A0 * Q11 + A1 * Q21 + A2 * Q31 + B0), 4);
  OB_TST(ob_dot41(R2 + 1, p), (A0 * Q12 + A1 * Q22 + A2 * Q32 + B1), 5);
  OB_TST(ob_dot41(R2 + 2, p), (A0 * Q13 + A1 * Q23 + A2 * Q33 + B2), 6);
#undef OB_TST
  // edge x edge axes (box.cpp:413-452): the penetration is measured along the normalised axis
#define OB_TST(expr1, expr2, n1, n2, n3, cc)                                                  \
  {                                                                                           \
    const real e_ = (expr1);                                                                  \
    real s2_ = ob_fabs(e_) - (expr2);                                                         \
    const bool live_ = !(sep | stop), out_ = s2_ > 0;                                         \
    sep |= live_ & out_;                                                                      \
    const real l_ = ob_sqrt((n1) * (n1) + (n2) * (n2) + (n3) * (n3));                         \
    s2_ /= l_;                                                                                \
    const bool upd_ = live_ & !out_ & (l_ > 0) & (s2_ * fudge_factor > s);                    \
    if (upd_) { s = s2_; en1 = (n1); en2 = (n2); en3 = (n3); el = l_; invert_normal = e_ < 0; code = (cc); } \
    stop |= upd_ & unimportant;                                                               \
  }
  OB_TST(pp[2] * R21 - pp[1] * R31, (A1 * Q31 + A2 * Q21 + B1 * Q13 + B2 * Q12), 0, -R31, R21, 7);
  OB_TST(pp[2] * R22 - pp[1] * R32, (A1 * Q32 + A2 * Q22 + B0 * Q13 + B2 * Q11), 0, -R32, R22, 8);
  OB_TST(pp[2] * R23 - pp[1] * R33, (A1 * Q33 + A2 * Q23 + B0 * Q12 + B1 * Q11), 0, -R33, R23, 9);
  OB_TST(pp[0] * R31 - pp[2] * R11, (A0 * Q31 + A2 * Q11 + B1 * Q23 + B2 * Q22), R31, 0, -R11, 10);
  OB_TST(pp[0] * R32 - pp[2] * R12, (A0 * Q32 + A2 * Q12 + B0 * Q23 + B2 * Q21), R32, 0, -R12, 11);
  OB_TST(pp[0] * R33 - pp[2] * R13, (A0 * Q33 + A2 * Q13 + B0 * Q22 + B1 * Q21), R33, 0, -R13, 12);
  OB_TST(pp[1] * R11 - pp[0] * R21, (A0 * Q21 + A1 * Q11 + B1 * Q33 + B2 * Q32), -R21, R11, 0, 13);
  OB_TST(pp[1] * R12 - pp[0] * R22, (A0 * Q22 + A1 * Q12 + B0 * Q33 + B2 * Q31), -R22, R12, 0, 14);
  OB_TST(pp[1] * R13 - pp[0] * R23, (A0 * Q23 + A1 * Q13 + B0 * Q32 + B1 * Q31), -R23, R13, 0, 15);
#undef OB_TST

  if (sep || !code) return 0;

  if (code <= 6) {
    const real *normalR = code <= 3 ? R1 + (code - 1) : R2 + (code - 4);
    normal[0] = normalR[0]; normal[1] = normalR[4]; normal[2] = normalR[8];
  } else {
    real normalC[3];
    normalC[0] = en1 / el; normalC[1] = en2 / el; normalC[2] = en3 / el;
    ob_mul0_331(normal, R1, normalC);
  }
  if (invert_normal) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
  *depth = -s;

  if (code > 6) {
    // edge-edge contact: the point midway between the closest points of the two edges (box.cpp:474-512)
    real pa[3], pb[3], sign;
    const real A[3] = {A0, A1, A2}, B[3] = {B0, B1, B2};
    for (i = 0; i < 3; i++) pa[i] = p1[i];
    for (j = 0; j < 3; j++) {
      sign = (ob_dot14(normal, R1 + j) > 0) ? OB_REAL(1.0) : OB_REAL(-1.0);
      for (i = 0; i < 3; i++) pa[i] += sign * A[j] * R1[i * 4 + j];
    }
    for (i = 0; i < 3; i++) pb[i] = p2[i];
    for (j = 0; j < 3; j++) {
      sign = (ob_dot14(normal, R2 + j) > 0) ? OB_REAL(-1.0) : OB_REAL(1.0);
      for (i = 0; i < 3; i++) pb[i] += sign * B[j] * R2[i * 4 + j];
    }
    real alpha, beta, ua[3], ub[3];
    for (i = 0; i < 3; i++) ua[i] = R1[((code) - 7) / 3 + i * 4];
    for (i = 0; i < 3; i++) ub[i] = R2[((code) - 7) % 3 + i * 4];
    ob_line_closest_approach(pa, ua, pb, ub, &alpha, &beta);
    for (i = 0; i < 3; i++) pa[i] += ua[i] * alpha;
    for (i = 0; i < 3; i++) pb[i] += ub[i] * beta;
    for (i = 0; i < 3; i++) contact[0].pos[i] = OB_REAL(0.5) * (pa[i] + pb[i]);
    contact[0].depth = *depth;
    *return_code = code;
    return 1;
  }

  // face-something contact (box.cpp:514-712): a = the box of the reference face, b = the other one.  The poses stay where
  // they are (pointer selects into memory); the half sizes are selected by value so that they stay in registers
  const bool face1 = code <= 3;
  const real *Ra = face1 ? R1 : R2, *Rb = face1 ? R2 : R1, *pa = face1 ? p1 : p2, *pb = face1 ? p2 : p1;
  const real Sa0 = face1 ? A0 : B0, Sa1 = face1 ? A1 : B1, Sa2 = face1 ? A2 : B2;
  const real Sb0 = face1 ? B0 : A0, Sb1 = face1 ? B1 : A1, Sb2 = face1 ? B2 : A2;

  real normal2[3], nr[3], anr[3];
  if (face1) { normal2[0] = normal[0]; normal2[1] = normal[1]; normal2[2] = normal[2]; }
  else { normal2[0] = -normal[0]; normal2[1] = -normal[1]; normal2[2] = -normal[2]; }
  ob_mul1_331(nr, Rb, normal2);
  anr[0] = ob_fabs(nr[0]); anr[1] = ob_fabs(nr[1]); anr[2] = ob_fabs(nr[2]);

  // largest component of the normal in b's frame -> incident face lanr, its two in-plane axes a1, a2
  int lanr, a1, a2;
  if (anr[1] > anr[0]) {
    if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; }
    else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; }
    else { a1 = 0; a2 = 1; lanr = 2; }
  }

  real center[3];
  const real Sbl = ob_sel3(Sb0, Sb1, Sb2, lanr), nrl = ob_sel3(nr[0], nr[1], nr[2], lanr);
  if (nrl < 0) { for (i = 0; i < 3; i++) center[i] = pb[i] - pa[i] + Sbl * Rb[i * 4 + lanr]; }
  else { for (i = 0; i < 3; i++) center[i] = pb[i] - pa[i] - Sbl * Rb[i * 4 + lanr]; }

  int codeN, code1, code2;
  if (face1) codeN = code - 1; else codeN = code - 4;
  if (codeN == 0) { code1 = 1; code2 = 2; }
  else if (codeN == 1) { code1 = 0; code2 = 2; }
  else { code1 = 0; code2 = 1; }

  real quad[8];
  real c1, c2, m11, m12, m21, m22;
  c1 = ob_dot14(center, Ra + code1);
  c2 = ob_dot14(center, Ra + code2);
  m11 = ob_dot44(Ra + code1, Rb + a1);
  m12 = ob_dot44(Ra + code1, Rb + a2);
  m21 = ob_dot44(Ra + code2, Rb + a1);
  m22 = ob_dot44(Ra + code2, Rb + a2);
  {
    const real Sba1 = ob_sel3(Sb0, Sb1, Sb2, a1), Sba2 = ob_sel3(Sb0, Sb1, Sb2, a2);
    real k1 = m11 * Sba1, k2 = m21 * Sba1, k3 = m12 * Sba2, k4 = m22 * Sba2;
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  real rect[2] = {ob_sel3(Sa0, Sa1, Sa2, code1), ob_sel3(Sa0, Sa1, Sa2, code2)};
  real ret[16];
  int n = ob_intersect_rect_quad(rect, quad, ret);
  if (n < 1) return 0;

  real point[3 * 8];
  real dep[8];
  real det1 = ob_recip(m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  const real SaN = ob_sel3(Sa0, Sa1, Sa2, codeN);
  int cnum = 0;
  for (j = 0; j < n; j++) {
    real k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    real k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    for (i = 0; i < 3; i++) point[cnum * 3 + i] = center[i] + k1 * Rb[i * 4 + a1] + k2 * Rb[i * 4 + a2];
    dep[cnum] = SaN - ob_dot(normal2, point + cnum * 3);
    if (dep[cnum] >= 0) {
      ret[cnum * 2] = ret[j * 2];
      ret[cnum * 2 + 1] = ret[j * 2 + 1];
      cnum++;
      if ((unsigned)(cnum | 0x80000000u) == ((unsigned)flags & (0xffffu | 0x80000000u))) break;
    }
  }
  if (cnum < 1) return 0;

  int maxc = flags & 0xffff;
  if (maxc > cnum) maxc = cnum;
  if (maxc < 1) maxc = 1;

  if (cnum <= maxc) {
    for (j = 0; j < cnum; j++) {
      for (i = 0; i < 3; i++) contact[j].pos[i] = point[j * 3 + i] + pa[i];
      contact[j].depth = dep[j];
    }
  } else {
    int i1 = 0;
    real maxdepth = dep[0];
    for (i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
    int iret[8];
    ob_cull_points(cnum, ret, maxc, i1, iret);
    for (j = 0; j < maxc; j++) {
      for (i = 0; i < 3; i++) contact[j].pos[i] = point[iret[j] * 3 + i] + pa[i];
      contact[j].depth = dep[iret[j]];
    }
    cnum = maxc;
  }
  *return_code = code;
  return cnum;
}

OB_HD int ob_collide_box_box(const ObPose &o1, const ObPose &o2, int flags, ObCg *contact) {
  real normal[3], depth;
  int code;
  int num = ob_box_box(o1.pos, o1.R, o1.p, o2.pos, o2.R, o2.p, normal, &depth, &code, flags, contact);
  for (int i = 0; i < num; i++) {
    contact[i].normal[0] = -normal[0];
    contact[i].normal[1] = -normal[1];
    contact[i].normal[2] = -normal[2];
  }
  return num;
}

// dCollideBoxPlane, box.cpp:745-878
OB_HDN int ob_collide_box_plane(const ObPose &o1, const ObPose &o2, int flags, ObCg *contact) {
  int ret = 0;
  const real *R = o1.R;
  const real *n = o2.p;
  const real *side = o1.p;
  real Q1 = ob_dot14(n, R + 0), Q2 = ob_dot14(n, R + 1), Q3 = ob_dot14(n, R + 2);
  real A1 = side[0] * Q1, A2 = side[1] * Q2, A3 = side[2] * Q3;
  real B1 = ob_fabs(A1), B2 = ob_fabs(A2), B3 = ob_fabs(A3);
  real depth = n[3] + OB_REAL(0.5) * (B1 + B2 + B3) - ob_dot(n, o1.pos);
  if (depth < 0) return 0;
  int maxc = flags & 0xffff;
  if (maxc > 4) maxc = 4;
  real p[3] = {o1.pos[0], o1.pos[1], o1.pos[2]};
#define OB_FOO(i, op)                          \
  p[0] op OB_REAL(0.5) * side[i] * R[0 + i];   \
  p[1] op OB_REAL(0.5) * side[i] * R[4 + i];   \
  p[2] op OB_REAL(0.5) * side[i] * R[8 + i];
#define OB_BAR(i, AA) if (AA > 0) { OB_FOO(i, -=) } else { OB_FOO(i, +=) }
  OB_BAR(0, A1);
  OB_BAR(1, A2);
  OB_BAR(2, A3);
#undef OB_FOO
#undef OB_BAR
  contact[0].pos[0] = p[0]; contact[0].pos[1] = p[1]; contact[0].pos[2] = p[2];
  contact[0].depth = depth;
  ret = 1;
  if (maxc == 1) goto done;
#define OB_FOO(i, j, op)                           \
  contact[i].pos[0] = p[0] op side[j] * R[0 + j];  \
  contact[i].pos[1] = p[1] op side[j] * R[4 + j];  \
  contact[i].pos[2] = p[2] op side[j] * R[8 + j];
#define OB_BAR(ctact, sd, AA, BB)                                    \
  if (depth - BB < 0) goto done;                                     \
  if (AA > 0) { OB_FOO(ctact, sd, +); } else { OB_FOO(ctact, sd, -); } \
  contact[ctact].depth = depth - BB;                                 \
  ret++;
  if (B1 < B2) {
    if (B3 < B1) goto use_side_3;
    else {
      OB_BAR(1, 0, A1, B1);
      if (maxc == 2) goto done;
      if (B2 < B3) goto contact2_2; else goto contact2_3;
    }
  } else {
    if (B3 < B2) {
    use_side_3:
      OB_BAR(1, 2, A3, B3);
      if (maxc == 2) goto done;
      if (B1 < B2) goto contact2_1; else goto contact2_2;
    } else {
      OB_BAR(1, 1, A2, B2);
      if (maxc == 2) goto done;
      if (B1 < B3) goto contact2_1; else goto contact2_3;
    }
  }
contact2_1: OB_BAR(2, 0, A1, B1); goto done;
contact2_2: OB_BAR(2, 1, A2, B2); goto done;
contact2_3: OB_BAR(2, 2, A3, B3); goto done;
#undef OB_FOO
#undef OB_BAR
done:
  if (maxc == 4 && ret == 3) {
    real d4 = contact[1].depth + contact[2].depth - depth;
    if (d4 > 0) {
      contact[3].pos[0] = contact[1].pos[0] + contact[2].pos[0] - p[0];
      contact[3].pos[1] = contact[1].pos[1] + contact[2].pos[1] - p[1];
      contact[3].pos[2] = contact[1].pos[2] + contact[2].pos[2] - p[2];
      contact[3].depth = d4;
      ret++;
    }
  }
  for (int i = 0; i < ret; i++) { contact[i].normal[0] = n[0]; contact[i].normal[1] = n[1]; contact[i].normal[2] = n[2]; }
  return ret;
}

// ---- capsule colliders (capsule.cpp:130-411) ------------------------------------
// dClosestLineSegmentPoints, collision_util.cpp:107-219
OB_HD void ob_closest_segment_points(const real *a1, const real *a2, const real *b1, const real *b2, real *cp1, real *cp2) {
  real a1a2[3], b1b2[3], a1b1[3], a1b2[3], a2b1[3], a2b2[3], n[3];
  real la, lb, k, da1, da2, da3, da4, db1, db2, db3, db4, det;
  for (int i = 0; i < 3; i++) { a1a2[i] = a2[i] - a1[i]; b1b2[i] = b2[i] - b1[i]; a1b1[i] = b1[i] - a1[i]; }
  da1 = ob_dot(a1a2, a1b1);
  db1 = ob_dot(b1b2, a1b1);
  if (da1 <= 0 && db1 >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b1[i]; } return; }
  for (int i = 0; i < 3; i++) a1b2[i] = b2[i] - a1[i];
  da2 = ob_dot(a1a2, a1b2);
  db2 = ob_dot(b1b2, a1b2);
  if (da2 <= 0 && db2 <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b2[i]; } return; }
  for (int i = 0; i < 3; i++) a2b1[i] = b1[i] - a2[i];
  da3 = ob_dot(a1a2, a2b1);
  db3 = ob_dot(b1b2, a2b1);
  if (da3 >= 0 && db3 >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a2[i]; cp2[i] = b1[i]; } return; }
  for (int i = 0; i < 3; i++) a2b2[i] = b2[i] - a2[i];
  da4 = ob_dot(a1a2, a2b2);
  db4 = ob_dot(b1b2, a2b2);
  if (da4 >= 0 && db4 <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a2[i]; cp2[i] = b2[i]; } return; }
  la = ob_dot(a1a2, a1a2);
  if (da1 >= 0 && da3 <= 0) {
    k = da1 / la;
    for (int i = 0; i < 3; i++) n[i] = a1b1[i] - k * a1a2[i];
    if (ob_dot(b1b2, n) >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i] + k * a1a2[i]; cp2[i] = b1[i]; } return; }
  }
  if (da2 >= 0 && da4 <= 0) {
    k = da2 / la;
    for (int i = 0; i < 3; i++) n[i] = a1b2[i] - k * a1a2[i];
    if (ob_dot(b1b2, n) <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i] + k * a1a2[i]; cp2[i] = b2[i]; } return; }
  }
  lb = ob_dot(b1b2, b1b2);
  if (db1 <= 0 && db2 >= 0) {
    k = -db1 / lb;
    for (int i = 0; i < 3; i++) n[i] = -a1b1[i] - k * b1b2[i];
    if (ob_dot(a1a2, n) >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b1[i] + k * b1b2[i]; } return; }
  }
  if (db3 <= 0 && db4 >= 0) {
    k = -db3 / lb;
    for (int i = 0; i < 3; i++) n[i] = -a2b1[i] - k * b1b2[i];
    if (ob_dot(a1a2, n) <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a2[i]; cp2[i] = b1[i] + k * b1b2[i]; } return; }
  }
  k = ob_dot(a1a2, b1b2);
  det = la * lb - k * k;
  if (det <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b1[i]; } return; }
  det = ob_recip(det);
  real alpha = (lb * da1 - k * db1) * det;
  real beta = (k * da1 - la * db1) * det;
  for (int i = 0; i < 3; i++) { cp1[i] = a1[i] + alpha * a1a2[i]; cp2[i] = b1[i] + beta * b1b2[i]; }
}

// dClosestLineBoxPoints, collision_util.cpp:244-391
OB_HD void ob_closest_line_box_points(const real *p1, const real *p2, const real *c, const real *R, const real *side,
                                      real *lret, real *bret) {
  real tmp[3], s[3], v[3], sign[3], v2[3], h[3], tanchor[3];
  int region[3];
  tmp[0] = p1[0] - c[0]; tmp[1] = p1[1] - c[1]; tmp[2] = p1[2] - c[2];
  ob_mul1_331(s, R, tmp);
  tmp[0] = p2[0] - p1[0]; tmp[1] = p2[1] - p1[1]; tmp[2] = p2[2] - p1[2];
  ob_mul1_331(v, R, tmp);
  for (int i = 0; i < 3; i++) {
    if (v[i] < 0) { s[i] = -s[i]; v[i] = -v[i]; sign[i] = -1; }
    else sign[i] = 1;
  }
  for (int i = 0; i < 3; i++) { v2[i] = v[i] * v[i]; h[i] = OB_REAL(0.5) * side[i]; }
#if defined(dSINGLE)
  const real tanchor_eps = OB_REAL(1e-19);
#else
  const real tanchor_eps = OB_REAL(1e-307);
#endif
  for (int i = 0; i < 3; i++) {
    if (v[i] > tanchor_eps) {
      if (s[i] < -h[i]) { region[i] = -1; tanchor[i] = (-h[i] - s[i]) / v[i]; }
      else { region[i] = (s[i] > h[i]); tanchor[i] = (h[i] - s[i]) / v[i]; }
    } else { region[i] = 0; tanchor[i] = 2; }
  }
  real t = 0;
  real dd2dt = 0;
  bool done = false;
  for (int i = 0; i < 3; i++) dd2dt -= (region[i] ? v2[i] : (real)0) * tanchor[i];
  if (dd2dt >= 0) done = true;
  if (!done) {
    do {
      real next_t = 1;
      for (int i = 0; i < 3; i++)
        if (tanchor[i] > t && tanchor[i] < 1 && tanchor[i] < next_t) next_t = tanchor[i];
      real next_dd2dt = 0;
      for (int i = 0; i < 3; i++) next_dd2dt += (region[i] ? v2[i] : (real)0) * (next_t - tanchor[i]);
      if (next_dd2dt >= 0) {
        real m = (next_dd2dt - dd2dt) / (next_t - t);
        t -= dd2dt / m;
        done = true;
        break;
      }
      for (int i = 0; i < 3; i++) {
        if (tanchor[i] == next_t) { tanchor[i] = (h[i] - s[i]) / v[i]; region[i]++; }
      }
      t = next_t;
      dd2dt = next_dd2dt;
    } while (t < 1);
    if (!done) t = 1;
  }
  for (int i = 0; i < 3; i++) lret[i] = p1[i] + t * tmp[i];
  for (int i = 0; i < 3; i++) {
    tmp[i] = sign[i] * (s[i] + t * v[i]);
    if (tmp[i] < -h[i]) tmp[i] = -h[i];
    else if (tmp[i] > h[i]) tmp[i] = h[i];
  }
  ob_mul0_331(s, R, tmp);
  for (int i = 0; i < 3; i++) bret[i] = s[i] + c[i];
}

// dCollideCapsuleSphere, capsule.cpp:130-161
OB_HD int ob_collide_capsule_sphere(const ObPose &o1, const ObPose &o2, ObCg *c) {
  real alpha = o1.R[2] * (o2.pos[0] - o1.pos[0]) + o1.R[6] * (o2.pos[1] - o1.pos[1]) + o1.R[10] * (o2.pos[2] - o1.pos[2]);
  real lz2 = o1.p[1] * OB_REAL(0.5);
  if (alpha > lz2) alpha = lz2;
  if (alpha < -lz2) alpha = -lz2;
  real p[3];
  p[0] = o1.pos[0] + alpha * o1.R[2];
  p[1] = o1.pos[1] + alpha * o1.R[6];
  p[2] = o1.pos[2] + alpha * o1.R[10];
  return ob_collide_spheres(p, o1.p[0], o2.pos, o2.p[0], c);
}

// dCollideCapsuleBox, capsule.cpp:178-231 (dCollideSpheresZeroDist :165-176)
OB_HD int ob_collide_capsule_box(const ObPose &o1, const ObPose &o2, ObCg *contact) {
  real p1[3], p2[3];
  real clen = o1.p[1] * OB_REAL(0.5);
  p1[0] = o1.pos[0] + clen * o1.R[2]; p1[1] = o1.pos[1] + clen * o1.R[6]; p1[2] = o1.pos[2] + clen * o1.R[10];
  p2[0] = o1.pos[0] - clen * o1.R[2]; p2[1] = o1.pos[1] - clen * o1.R[6]; p2[2] = o1.pos[2] - clen * o1.R[10];
  real radius = o1.p[0];
  const real *c = o2.pos;
  real pl[3], pb[3];
  ob_closest_line_box_points(p1, p2, c, o2.R, o2.p, pl, pb);
#if defined(dSINGLE)
  const real mindist = OB_REAL(1e-9);
#else
  const real mindist = OB_REAL(1e-18);
#endif
  real dist = ob_sqrt((pl[0] - pb[0]) * (pl[0] - pb[0]) + (pl[1] - pb[1]) * (pl[1] - pb[1]) + (pl[2] - pb[2]) * (pl[2] - pb[2]));
  if (dist < mindist) {
    real normal[3];
    for (int i = 0; i < 3; i++) normal[i] = pb[i] - c[i];
    ob_safe_normalize3(normal);
    contact->normal[0] = normal[0]; contact->normal[1] = normal[1]; contact->normal[2] = normal[2];
    contact->depth = radius + 0;
    real k = OB_REAL(0.5) * (0 - radius);
    contact->pos[0] = pl[0] + contact->normal[0] * k;
    contact->pos[1] = pl[1] + contact->normal[1] * k;
    contact->pos[2] = pl[2] + contact->normal[2] * k;
    return 1;
  }
  return ob_collide_spheres(pl, radius, pb, 0, contact);
}

// dCollideCapsuleCapsule, capsule.cpp:234-349
OB_HD int ob_collide_capsule_capsule(const ObPose &o1, const ObPose &o2, int flags, ObCg *contact) {
  const real tolerance = OB_REAL(1e-5);
  real lz1 = o1.p[1] * OB_REAL(0.5), lz2 = o2.p[1] * OB_REAL(0.5);
  const real *pos1 = o1.pos, *pos2 = o2.pos;
  real axis1[3] = {o1.R[2], o1.R[6], o1.R[10]}, axis2[3] = {o2.R[2], o2.R[6], o2.R[10]};
  real sphere1[3], sphere2[3];
  real a1a2 = ob_dot(axis1, axis2);
  real det = OB_REAL(1.0) - a1a2 * a1a2;
  if (det < tolerance) {
    if (a1a2 < 0) { axis2[0] = -axis2[0]; axis2[1] = -axis2[1]; axis2[2] = -axis2[2]; }
    real q[3];
    for (int i = 0; i < 3; i++) q[i] = pos1[i] - pos2[i];
    real k = ob_dot(axis1, q);
    real a1lo = -lz1, a1hi = lz1, a2lo = -lz2 - k, a2hi = lz2 - k;
    real lo = (a1lo > a2lo) ? a1lo : a2lo;
    real hi = (a1hi < a2hi) ? a1hi : a2hi;
    if (lo <= hi) {
      int num_contacts = flags & 0xffff;
      if (num_contacts >= 2 && lo < hi) {
        for (int i = 0; i < 3; i++) sphere1[i] = pos1[i] + lo * axis1[i];
        for (int i = 0; i < 3; i++) sphere2[i] = pos2[i] + (lo + k) * axis2[i];
        int n1 = ob_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
        if (n1) {
          for (int i = 0; i < 3; i++) sphere1[i] = pos1[i] + hi * axis1[i];
          for (int i = 0; i < 3; i++) sphere2[i] = pos2[i] + (hi + k) * axis2[i];
          int n2 = ob_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact + 1);
          if (n2) return 2;
        }
      }
      real alpha1 = (lo + hi) * OB_REAL(0.5);
      real alpha2 = alpha1 + k;
      for (int i = 0; i < 3; i++) sphere1[i] = pos1[i] + alpha1 * axis1[i];
      for (int i = 0; i < 3; i++) sphere2[i] = pos2[i] + alpha2 * axis2[i];
      return ob_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
    }
  }
  real a1[3], a2[3], b1[3], b2[3];
  for (int i = 0; i < 3; i++) {
    a1[i] = o1.pos[i] + axis1[i] * lz1; a2[i] = o1.pos[i] - axis1[i] * lz1;
    b1[i] = o2.pos[i] + axis2[i] * lz2; b2[i] = o2.pos[i] - axis2[i] * lz2;
  }
  ob_closest_segment_points(a1, a2, b1, b2, sphere1, sphere2);
  return ob_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
}

// dCollideCapsulePlane, capsule.cpp:352-411
OB_HD int ob_collide_capsule_plane(const ObPose &o1, const ObPose &o2, int flags, ObCg *contact) {
  const real *pp = o2.p;
  const real radius = o1.p[0], lz = o1.p[1];
  real sign = (pp[0] * o1.R[2] + pp[1] * o1.R[6] + pp[2] * o1.R[10] > 0) ? OB_REAL(-1.0) : OB_REAL(1.0);
  real p[3];
  p[0] = o1.pos[0] + o1.R[2] * lz * OB_REAL(0.5) * sign;
  p[1] = o1.pos[1] + o1.R[6] * lz * OB_REAL(0.5) * sign;
  p[2] = o1.pos[2] + o1.R[10] * lz * OB_REAL(0.5) * sign;
  real k = ob_dot(p, pp);
  real depth = pp[3] - k + radius;
  if (depth < 0) return 0;
  for (int i = 0; i < 3; i++) { contact->normal[i] = pp[i]; contact->pos[i] = p[i] - pp[i] * radius; }
  contact->depth = depth;
  int ncontacts = 1;
  if ((flags & 0xffff) >= 2) {
    p[0] = o1.pos[0] - o1.R[2] * lz * OB_REAL(0.5) * sign;
    p[1] = o1.pos[1] - o1.R[6] * lz * OB_REAL(0.5) * sign;
    p[2] = o1.pos[2] - o1.R[10] * lz * OB_REAL(0.5) * sign;
    k = ob_dot(p, pp);
    depth = pp[3] - k + radius;
    if (depth >= 0) {
      ObCg *c2 = contact + 1;
      for (int i = 0; i < 3; i++) { c2->normal[i] = pp[i]; c2->pos[i] = p[i] - pp[i] * radius; }
      c2->depth = depth;
      ncontacts = 2;
    }
  }
  return ncontacts;
}

// ---- flat-ended cylinder colliders --------------------------------------------------------------
// cylinder pose: p[0] = radius, p[1] = length, axis = column 2 of the rotation.
#if defined(dSINGLE)
#define OB_CYL_TOL OB_REAL(0.0001)
#else
#define OB_CYL_TOL OB_REAL(0.0000001)
#endif
// dCollideCylinderPlane, collision_cylinder_plane.cpp:38-266 (o1 = cylinder, o2 = plane)
OB_HDN int ob_collide_cylinder_plane(const ObPose &o1, const ObPose &o2, int flags, ObCg *contact) {
  const int maxc = flags & 0xffff;
  int n = 0;
  const real radius = o1.p[0], length = o1.p[1];
  const real *cylpos = o1.pos, *pv = o2.p;
  const real vDir1[3] = {o1.R[2], o1.R[6], o1.R[10]};
  real s = length * OB_REAL(0.5);
  real G1Pos1[3], G1Pos2[3];
  for (int i = 0; i < 3; i++) { G1Pos2[i] = vDir1[i] * s + cylpos[i]; G1Pos1[i] = vDir1[i] * -s + cylpos[i]; }
  s = vDir1[0] * pv[0] + vDir1[1] * pv[1] + vDir1[2] * pv[2];
  if (s < 0) s += OB_REAL(1.0); else s -= OB_REAL(1.0);
#define OB_CYLPL_EMIT(COND)                                                                       \
  {                                                                                               \
    ObCg *c = contact + n;                                                                        \
    c->depth = pv[3] - ob_dot(pv, c->pos);                                                        \
    if (c->depth COND 0) {                                                                        \
      c->normal[0] = pv[0]; c->normal[1] = pv[1]; c->normal[2] = pv[2];                           \
      c->side1 = -1; c->side2 = -1;                                                               \
      n++;                                                                                        \
      if (n >= maxc) return n;                                                                    \
    }                                                                                             \
  }
  if (s < OB_CYL_TOL && s > (-OB_CYL_TOL)) {
    // the axis is parallel to the normal: the deeper disc touches with up to four rim points
    real P[3];
    s = pv[3] - ob_dot(pv, G1Pos1);
    real t = pv[3] - ob_dot(pv, G1Pos2);
    if (s >= t) { if (s >= 0) { P[0] = G1Pos1[0]; P[1] = G1Pos1[1]; P[2] = G1Pos1[2]; } else return n; }
    else { if (t >= 0) { P[0] = G1Pos2[0]; P[1] = G1Pos2[1]; P[2] = G1Pos2[2]; } else return n; }
    real V1[3], V2[3];
    if (vDir1[0] < OB_CYL_TOL && vDir1[0] > (-OB_CYL_TOL)) { V1[0] = vDir1[0] + OB_REAL(1.0); V1[1] = vDir1[1]; V1[2] = vDir1[2]; }
    else { V1[0] = vDir1[0]; V1[1] = vDir1[1] + OB_REAL(1.0); V1[2] = vDir1[2]; }
    ob_cross(V2, V1, vDir1);
    t = ob_sqrt(V2[0] * V2[0] + V2[1] * V2[1] + V2[2] * V2[2]);
    t = radius / t;
    V2[0] *= t; V2[1] *= t; V2[2] *= t;
    ob_cross(V1, V2, vDir1);
    for (int i = 0; i < 3; i++) contact[n].pos[i] = P[i] + V1[i];
    OB_CYLPL_EMIT(>)
    for (int i = 0; i < 3; i++) contact[n].pos[i] = P[i] - V1[i];
    OB_CYLPL_EMIT(>)
    for (int i = 0; i < 3; i++) contact[n].pos[i] = P[i] + V2[i];
    OB_CYLPL_EMIT(>)
    for (int i = 0; i < 3; i++) contact[n].pos[i] = P[i] - V2[i];
    OB_CYLPL_EMIT(>)
  } else {
    real C[3];
    const real t = ob_dot(pv, vDir1);
    for (int i = 0; i < 3; i++) C[i] = vDir1[i] * t - pv[i];
    s = ob_sqrt(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]);
    s = radius / s;
    C[0] *= s; C[1] *= s; C[2] *= s;
    for (int i = 0; i < 3; i++) contact[n].pos[i] = C[i] + G1Pos1[i];
    OB_CYLPL_EMIT(>=)
    for (int i = 0; i < 3; i++) contact[n].pos[i] = C[i] + G1Pos2[i];
    {   // the second depth is written out term by term in the reference (:250)
      ObCg *c = contact + n;
      c->depth = pv[3] - pv[0] * c->pos[0] - pv[1] * c->pos[1] - pv[2] * c->pos[2];
      if (c->depth >= 0) {
        c->normal[0] = pv[0]; c->normal[1] = pv[1]; c->normal[2] = pv[2];
        c->side1 = -1; c->side2 = -1;
        n++;
        if (n >= maxc) return n;
      }
    }
  }
#undef OB_CYLPL_EMIT
  return n;
}

// dCollideCylinderSphere, collision_cylinder_sphere.cpp:51-277 (o1 = cylinder, o2 = sphere); one contact or none
OB_HDN int ob_collide_cylinder_sphere(const ObPose &o1, const ObPose &o2, ObCg *contact) {
  const real radius = o1.p[0], length = o1.p[1], radius2 = o2.p[0];
  const real *cylpos = o1.pos, *sp = o2.pos;
  const real vDir1[3] = {o1.R[2], o1.R[6], o1.R[10]};
  real s = length * OB_REAL(0.5);
  real G1Pos1[3], G1Pos2[3], C[3];
  for (int i = 0; i < 3; i++) { G1Pos2[i] = vDir1[i] * s + cylpos[i]; G1Pos1[i] = vDir1[i] * -s + cylpos[i]; }
  s = (sp[0] - G1Pos1[0]) * vDir1[0] - (G1Pos1[1] - sp[1]) * vDir1[1] - (G1Pos1[2] - sp[2]) * vDir1[2];
  if (s < (-radius2) || s > (length + radius2)) return 0;
  for (int i = 0; i < 3; i++) C[i] = s * vDir1[i] + G1Pos1[i] - sp[i];
  const real t = ob_sqrt(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]);
  if (t > (radius + radius2)) return 0;
  contact->side1 = -1; contact->side2 = -1;
  if (t > radius && (s < 0 || s > length)) {
    // the sphere touches a rim
    const real *G = s <= 0 ? G1Pos1 : G1Pos2;
    const real ds = s <= 0 ? s : s - length;
    contact->depth = radius2 - ob_sqrt(ds * ds + (t - radius) * (t - radius));
    if (contact->depth < 0) return 0;
    for (int i = 0; i < 3; i++) contact->pos[i] = C[i] / t * -radius + G[i];
    for (int i = 0; i < 3; i++) contact->normal[i] = (contact->pos[i] - sp[i]) / (radius2 - contact->depth);
    return 1;
  } else if ((radius - t) <= s && (radius - t) <= (length - s)) {
    // the sphere touches the mantle
    contact->depth = (radius2 + radius) - t;
    if (contact->depth < 0) return 0;
    if (t > (radius2 + OB_CYL_TOL)) {
      C[0] /= t; C[1] /= t; C[2] /= t;
      for (int i = 0; i < 3; i++) { contact->pos[i] = C[i] * radius2 + sp[i]; contact->normal[i] = C[i]; }
    } else {
      for (int i = 0; i < 3; i++) { contact->pos[i] = C[i] + sp[i]; contact->normal[i] = C[i] / t; }
    }
    return 1;
  } else {
    // the sphere touches a disc
    if (s <= (length * OB_REAL(0.5))) {
      contact->depth = s + radius2;
      if (contact->depth < 0) return 0;
      for (int i = 0; i < 3; i++) { contact->pos[i] = radius2 * vDir1[i] + sp[i]; contact->normal[i] = vDir1[i]; }
    } else {
      contact->depth = (radius2 + length - s);
      if (contact->depth < 0) return 0;
      for (int i = 0; i < 3; i++) { contact->pos[i] = radius2 * -vDir1[i] + sp[i]; contact->normal[i] = -vDir1[i]; }
    }
    return 1;
  }
}

// dCollideCylinderBox, collision_cylinder_box.cpp (o1 = cylinder, o2 = box): separating axes (3 box axes, the cylinder
// axis, 3 cross products, 8 vertex axes, 2 x 12 edge-circle axes), then either the cylinder's nearest mantle line
// clipped to the box (two contacts) or the box's nearest face clipped to the cylinder's cap octagon.
struct ObCylBox {
  real cylR[12], cylPos[3], cylAxis[3], radius, size;
  real boxR[12], boxPos[3], boxHalf[3], vert[8][3];
  real diff[3], normal[3], bestDepth, bestrb, bestrc;
  int bestAxis;
};
// the octagon's outward normals in the cap frame: -cos / -sin of pi/8 + i*pi/4 accumulated in dReal as :186-196 does
// (glibc values of this image; gcc folds the same constants at -O2)
OB_HD void ob_cyl_segment_normal(int i, real *n) {
#if defined(dSINGLE)
  const float t[8][2] = {{-0.923879504f, -0.382683456f}, {-0.382683426f, -0.923879504f}, {0.382683516f, -0.923879504f}, {0.923879623f, -0.382683277f}, {0.923879445f, 0.382683665f}, {0.382683128f, 0.923879683f}, {-0.382683605f, 0.923879445f}, {-0.923879564f, 0.382683426f}};
#else
  const double t[8][2] = {{-0.92387953251128674, -0.38268343236508978}, {-0.38268343236508984, -0.92387953251128674}, {0.38268343236508973, -0.92387953251128674}, {0.92387953251128674, -0.38268343236508989}, {0.92387953251128685, 0.38268343236508967}, {0.38268343236509034, 0.92387953251128652}, {-0.38268343236508917, 0.92387953251128696}, {-0.92387953251128652, 0.38268343236509039}};
#endif
  n[0] = t[i][0]; n[1] = t[i][1]; n[2] = 0;
}
// _cldTestAxis :206-286
OB_HDN int ob_cylbox_test_axis(ObCylBox &D, real *vIn, int iAxis) {
  const real fL = ob_sqrt(vIn[0] * vIn[0] + vIn[1] * vIn[1] + vIn[2] * vIn[2]);
  if (fL < OB_REAL(1e-5)) return 1;
  ob_safe_normalize3(vIn);
  const real fdot1 = ob_dot(D.cylAxis, vIn);
  real frc;
  if (fdot1 > OB_REAL(1.0)) frc = D.size * OB_REAL(0.5);
  else if (fdot1 < OB_REAL(-1.0)) frc = D.size * OB_REAL(0.5);
  else frc = ob_fabs(fdot1 * (D.size * OB_REAL(0.5))) + D.radius * ob_sqrt(OB_REAL(1.0) - (fdot1 * fdot1));
  real frb = ob_fabs(ob_dot41(D.boxR + 0, vIn)) * D.boxHalf[0];
  frb += ob_fabs(ob_dot41(D.boxR + 1, vIn)) * D.boxHalf[1];
  frb += ob_fabs(ob_dot41(D.boxR + 2, vIn)) * D.boxHalf[2];
  const real fd = ob_dot(D.diff, vIn);
  real fDepth = frc + frb;
  if (ob_fabs(fd) > fDepth) return 0;
  fDepth -= ob_fabs(fd);
  if (fDepth < D.bestDepth) {
    D.bestDepth = fDepth;
    D.normal[0] = vIn[0]; D.normal[1] = vIn[1]; D.normal[2] = vIn[2];
    D.bestAxis = iAxis; D.bestrb = frb; D.bestrc = frc;
    if (fd > 0) { D.normal[0] = -D.normal[0]; D.normal[1] = -D.normal[1]; D.normal[2] = -D.normal[2]; }
  }
  return 1;
}
// _cldTestEdgeCircleAxis :289-333
OB_HDN int ob_cylbox_test_edge_circle(ObCylBox &D, const real *vcc, const real *v0, const real *v1, int iAxis) {
  real dirE[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
  ob_safe_normalize3(dirE);
  const real fdot2 = ob_dot(dirE, D.cylAxis);
  if (ob_fabs(fdot2) < OB_REAL(1e-5)) return 1;
  real t1[3] = {vcc[0] - v0[0], vcc[1] - v0[1], vcc[2] - v0[2]};
  const real fdot1 = ob_dot(t1, D.cylAxis);
  real vpnt[3];
  for (int i = 0; i < 3; i++) vpnt[i] = v0[i] + dirE[i] * (fdot1 / fdot2);
  real tangent[3], axis[3];
  for (int i = 0; i < 3; i++) t1[i] = vcc[i] - vpnt[i];
  ob_cross(tangent, t1, D.cylAxis);
  ob_cross(axis, tangent, dirE);
  return ob_cylbox_test_axis(D, axis, iAxis);
}
// dClipEdgeToPlane, collision_util.cpp:471-509
OB_HD int ob_clip_edge_to_plane(real *e0, real *e1, const real *pl) {
  const real d0 = pl[0] * e0[0] + pl[1] * e0[1] + pl[2] * e0[2] + pl[3];
  const real d1 = pl[0] * e1[0] + pl[1] * e1[1] + pl[2] * e1[2] + pl[3];
  if (d0 < 0 && d1 < 0) return 0;
  else if (d0 > 0 && d1 > 0) return 1;
  else if ((d0 > 0 && d1 < 0) || (d0 < 0 && d1 > 0)) {
    real ip[3];
    for (int i = 0; i < 3; i++) ip[i] = e0[i] - (e0[i] - e1[i]) * d0 / (d0 - d1);
    if (d0 < 0) { e0[0] = ip[0]; e0[1] = ip[1]; e0[2] = ip[2]; }
    else { e1[0] = ip[0]; e1[1] = ip[1]; e1[2] = ip[2]; }
    return 1;
  }
  return 1;
}
// dClipPolyToPlane, collision_util.cpp:512-557
OB_HDN int ob_clip_poly_to_plane(const real (*in)[3], int ctIn, real (*out)[3], const real *pl) {
  int ctOut = 0;
  int i0 = ctIn - 1;
  for (int i1 = 0; i1 < ctIn; i0 = i1, i1++) {
    const real d0 = pl[0] * in[i0][0] + pl[1] * in[i0][1] + pl[2] * in[i0][2] + pl[3];
    const real d1 = pl[0] * in[i1][0] + pl[1] * in[i1][1] + pl[2] * in[i1][2] + pl[3];
    if (d0 >= 0) { out[ctOut][0] = in[i0][0]; out[ctOut][1] = in[i0][1]; out[ctOut][2] = in[i0][2]; ctOut++; }
    if ((d0 > 0 && d1 < 0) || (d0 < 0 && d1 > 0)) {
      for (int i = 0; i < 3; i++) out[ctOut][i] = in[i0][i] - (in[i0][i] - in[i1][i]) * d0 / (d0 - d1);
      ctOut++;
    }
  }
  return ctOut;
}
// dMatrix3Inv, collision_util.h:192-213 -- including its operator precedence: only the second product of the
// three "cofactors" without parentheses is divided by the determinant
OB_HD void ob_matrix3_inv_ref(const real *ma, real *dst) {
  const real det = ma[0] * (ma[5] * ma[10] - ma[9] * ma[6]) - ma[1] * (ma[4] * ma[10] - ma[8] * ma[6]) + ma[2] * (ma[4] * ma[9] - ma[8] * ma[5]);
  for (int i = 0; i < 12; i++) dst[i] = 0;
  if (ob_fabs(det) < OB_REAL(0.0005)) { dst[0] = 1; dst[5] = 1; dst[10] = 1; return; }
  dst[0] = ma[5] * ma[10] - ma[6] * ma[9] / det;
  dst[1] = -(ma[1] * ma[10] - ma[9] * ma[2]) / det;
  dst[2] = ma[1] * ma[6] - ma[5] * ma[2] / det;
  dst[4] = -(ma[4] * ma[10] - ma[6] * ma[8]) / det;
  dst[5] = ma[0] * ma[10] - ma[8] * ma[2] / det;
  dst[6] = -(ma[0] * ma[6] - ma[4] * ma[2]) / det;
  dst[8] = ma[4] * ma[9] - ma[8] * ma[5] / det;
  dst[9] = -(ma[0] * ma[9] - ma[8] * ma[1]) / det;
  dst[10] = ma[0] * ma[5] - ma[1] * ma[4] / det;
}
OB_HDN int ob_collide_cylinder_box(const ObPose &o1, const ObPose &o2, int flags, ObCg *contact) {
  const int maxc = flags & 0xffff;
  ObCylBox D;
  // _cldInitCylinderBox :100-203
  for (int i = 0; i < 12; i++) { D.cylR[i] = o1.R[i]; D.boxR[i] = o2.R[i]; }
  for (int i = 0; i < 3; i++) { D.cylPos[i] = o1.pos[i]; D.boxPos[i] = o2.pos[i]; D.cylAxis[i] = o1.R[4 * i + 2]; D.boxHalf[i] = o2.p[i] * OB_REAL(0.5); }
  D.radius = o1.p[0]; D.size = o1.p[1];
  {
    const real sx[8] = {-1, 1, -1, 1, 1, 1, -1, -1}, sy[8] = {1, 1, -1, -1, 1, -1, -1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    for (int i = 0; i < 8; i++) {
      const real v[3] = {sx[i] < 0 ? -D.boxHalf[0] : D.boxHalf[0], sy[i] < 0 ? -D.boxHalf[1] : D.boxHalf[1], sz[i] < 0 ? -D.boxHalf[2] : D.boxHalf[2]};
      real t[3];
      ob_mul0_331(t, D.boxR, v);
      for (int k = 0; k < 3; k++) D.vert[i][k] = t[k] + D.boxPos[k];
    }
  }
  for (int i = 0; i < 3; i++) { D.diff[i] = D.cylPos[i] - D.boxPos[i]; D.normal[i] = 0; }
  D.bestDepth = OB_INF; D.bestrb = 0; D.bestrc = 0; D.bestAxis = 0;
  int n = 0;
  // _cldTestSeparatingAxes :336-573
  {
    real ax[3];
    const real eps = OB_REAL(1e-6);
    for (int a = 0; a < 3; a++) {
      ax[0] = D.boxR[a]; ax[1] = D.boxR[4 + a]; ax[2] = D.boxR[8 + a];
      if (!ob_cylbox_test_axis(D, ax, 1 + a)) return 0;
    }
    ax[0] = D.cylAxis[0]; ax[1] = D.cylAxis[1]; ax[2] = D.cylAxis[2];
    if (!ob_cylbox_test_axis(D, ax, 4)) return 0;
    for (int a = 0; a < 3; a++) {
      const real col[3] = {D.boxR[a], D.boxR[4 + a], D.boxR[8 + a]};
      ob_cross(ax, D.cylAxis, col);     // dVector3CrossMat3Col(m, col, v, r): r = v x m(:,col)
      if (ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2] > eps) { if (!ob_cylbox_test_axis(D, ax, 5 + a)) return 0; }
    }
    for (int i = 0; i < 8; i++) {
      real t1[3] = {D.vert[i][0] - D.cylPos[0], D.vert[i][1] - D.cylPos[1], D.vert[i][2] - D.cylPos[2]}, t2[3];
      ob_cross(t2, D.cylAxis, t1);
      ob_cross(ax, D.cylAxis, t2);
      if (ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2] > eps) { if (!ob_cylbox_test_axis(D, ax, 8 + i)) return 0; }
    }
    const unsigned char ea[12] = {1, 1, 2, 2, 4, 4, 0, 5, 5, 2, 4, 6}, eb[12] = {0, 3, 3, 0, 1, 7, 7, 3, 6, 6, 5, 7};
    real vcc[3];
    for (int i = 0; i < 3; i++) vcc[i] = D.cylPos[i] + D.cylAxis[i] * (D.size * OB_REAL(0.5));
    for (int e = 0; e < 12; e++) if (!ob_cylbox_test_edge_circle(D, vcc, D.vert[ea[e]], D.vert[eb[e]], 16 + e)) return 0;
    for (int i = 0; i < 3; i++) vcc[i] = D.cylPos[i] - D.cylAxis[i] * (D.size * OB_REAL(0.5));
    for (int e = 0; e < 12; e++) if (!ob_cylbox_test_edge_circle(D, vcc, D.vert[ea[e]], D.vert[eb[e]], 28 + e)) return 0;
  }
  if (D.bestAxis == 0) return 0;
  const real fdot = ob_dot(D.normal, D.cylAxis);
  if (ob_fabs(fdot) < OB_REAL(0.9)) {
    // _cldClipCylinderToBox :576-717
    real vN[3];
    const real ft = ob_dot(D.cylAxis, D.normal);
    for (int i = 0; i < 3; i++) vN[i] = D.normal[i] - D.cylAxis[i] * ft;
    ob_safe_normalize3(vN);
    real cpt[3], ep0[3], ep1[3];
    for (int i = 0; i < 3; i++) cpt[i] = D.cylPos[i] + vN[i] * D.radius;
    for (int i = 0; i < 3; i++) { ep0[i] = cpt[i] + D.cylAxis[i] * (D.size * OB_REAL(0.5)); ep1[i] = cpt[i] - D.cylAxis[i] * (D.size * OB_REAL(0.5)); }
    for (int i = 0; i < 3; i++) { ep0[i] -= D.boxPos[i]; ep1[i] -= D.boxPos[i]; }
    for (int k = 0; k < 6; k++) {
      const int a = k % 3;
      real pl[4] = {D.boxR[a], D.boxR[4 + a], D.boxR[8 + a], D.boxHalf[a]};
      if (k >= 3) { pl[0] = -pl[0]; pl[1] = -pl[1]; pl[2] = -pl[2]; }
      if (!ob_clip_edge_to_plane(ep0, ep1, pl)) return 0;
    }
    real d0 = D.bestrb + ob_dot(ep0, D.normal), d1 = D.bestrb + ob_dot(ep1, D.normal);
    if (d0 < 0) d0 = 0;
    if (d1 < 0) d1 = 0;
    for (int i = 0; i < 3; i++) { ep0[i] += D.boxPos[i]; ep1[i] += D.boxPos[i]; }
    ObCg *c = contact + n;
    c->depth = d0; c->side1 = -1; c->side2 = -1;
    for (int i = 0; i < 3; i++) { c->normal[i] = -D.normal[i]; c->pos[i] = ep0[i]; }
    n++;
    if (n != maxc) {
      c = contact + n;
      c->depth = d1; c->side1 = -1; c->side2 = -1;
      for (int i = 0; i < 3; i++) { c->normal[i] = -D.normal[i]; c->pos[i] = ep1[i]; }
      n++;
    }
    return n;
  }
  // _cldClipBoxToCylinder :720-982
  real circlePos[3], circleN[3] = {0, 0, 0};
  if (ob_dot(D.cylAxis, D.normal) > OB_REAL(0.0)) {
    for (int i = 0; i < 3; i++) circlePos[i] = D.cylPos[i] + D.cylAxis[i] * (D.size * OB_REAL(0.5));
    circleN[2] = OB_REAL(-1.0);
  } else {
    for (int i = 0; i < 3; i++) circlePos[i] = D.cylPos[i] - D.cylAxis[i] * (D.size * OB_REAL(0.5));
    circleN[2] = OB_REAL(1.0);
  }
  real vNr[3], inv[12];
  ob_matrix3_inv_ref(D.boxR, inv);
  ob_mul0_331(vNr, inv, D.normal);
  const real an[3] = {ob_fabs(vNr[0]), ob_fabs(vNr[1]), ob_fabs(vNr[2])};
  int iB0, iB1, iB2;
  if (an[1] > an[0]) {
    if (an[0] > an[2]) { iB0 = 1; iB1 = 0; iB2 = 2; }
    else if (an[1] > an[2]) { iB0 = 1; iB1 = 2; iB2 = 0; }
    else { iB0 = 2; iB1 = 1; iB2 = 0; }
  } else {
    if (an[1] > an[2]) { iB0 = 0; iB1 = 1; iB2 = 2; }
    else if (an[0] > an[2]) { iB0 = 0; iB1 = 2; iB2 = 1; }
    else { iB0 = 2; iB1 = 0; iB2 = 1; }
  }
  real center[3];
  {
    const real col[3] = {D.boxR[iB0], D.boxR[4 + iB0], D.boxR[8 + iB0]};
    if (vNr[iB0] > 0) { for (int i = 0; i < 3; i++) center[i] = D.boxPos[i] - D.boxHalf[iB0] * col[i]; }
    else { for (int i = 0; i < 3; i++) center[i] = D.boxPos[i] + D.boxHalf[iB0] * col[i]; }
  }
  real pts[4][3], A1[16][3], A2[16][3];
  for (int i = 0; i < 16; i++) for (int k = 0; k < 3; k++) { A1[i][k] = 0; A2[i][k] = 0; }
  {
    const real a1[3] = {D.boxR[iB1], D.boxR[4 + iB1], D.boxR[8 + iB1]}, a2[3] = {D.boxR[iB2], D.boxR[4 + iB2], D.boxR[8 + iB2]};
    for (int k = 0; k < 3; k++) {
      pts[0][k] = center[k] + D.boxHalf[iB1] * a1[k] - D.boxHalf[iB2] * a2[k];
      pts[1][k] = center[k] - D.boxHalf[iB1] * a1[k] - D.boxHalf[iB2] * a2[k];
      pts[2][k] = center[k] - D.boxHalf[iB1] * a1[k] + D.boxHalf[iB2] * a2[k];
      pts[3][k] = center[k] + D.boxHalf[iB1] * a1[k] + D.boxHalf[iB2] * a2[k];
    }
  }
  real cinv[12];
  ob_matrix3_inv_ref(D.cylR, cinv);
  for (int i = 0; i < 4; i++) {
    const real t[3] = {pts[i][0] - circlePos[0], pts[i][1] - circlePos[1], pts[i][2] - circlePos[2]};
    ob_mul0_331(pts[i], cinv, t);
  }
  int c1 = 0, c2 = 0;
  {
    const real pl[4] = {circleN[0], circleN[1], circleN[2], OB_REAL(0.0)};
    c1 = ob_clip_poly_to_plane(pts, 4, A1, pl);
  }
  for (int sgm = 0; sgm < 8; sgm++) {
    real pl[4];
    ob_cyl_segment_normal(sgm, pl);
    pl[3] = D.radius;
    if (0 == (sgm % 2)) c2 = ob_clip_poly_to_plane(A1, c1, A2, pl);
    else c1 = ob_clip_poly_to_plane(A2, c2, A1, pl);
  }
  // eight segments: the result is in A1 (nCircleSegment % 2 == 0 after the loop)
  for (int i = 0; i < c1; i++) {
    real vp[3];
    ob_mul0_331(vp, D.cylR, A1[i]);
    vp[0] += circlePos[0]; vp[1] += circlePos[1]; vp[2] += circlePos[2];
    const real t[3] = {vp[0] - D.cylPos[0], vp[1] - D.cylPos[1], vp[2] - D.cylPos[2]};
    const real ftmpdot = ob_dot(t, D.normal);
    const real fd = D.bestrc - ftmpdot;
    if (fd > OB_REAL(0.0)) {
      ObCg *c = contact + n;
      c->depth = fd; c->side1 = -1; c->side2 = -1;
      for (int k = 0; k < 3; k++) { c->normal[k] = -D.normal[k]; c->pos[k] = vp[k]; }
      n++;
      if (n == maxc) break;
    }
  }
  return n;
}

// ---- ray colliders (ode/src/ray.cpp) ----------------------------------------------------------
// ray pose: pos = origin, R(:,2) = direction, p[0] = length.  o1 = ray.
// ray_sphere_helper, ray.cpp:192-232; mode 1 = use the exit point
OB_HD int ob_ray_sphere_helper(const ObPose &ray, const real *sphere_pos, real radius, ObCg *contact, int mode) {
  real q[3] = {ray.pos[0] - sphere_pos[0], ray.pos[1] - sphere_pos[1], ray.pos[2] - sphere_pos[2]};
  const real B = ob_dot14(q, ray.R + 2);
  const real C = ob_dot(q, q) - radius * radius;
  real k = B * B - C;
  if (k < 0) return 0;
  k = ob_sqrt(k);
  real alpha;
  if (mode && C >= 0) {
    alpha = -B + k;
    if (alpha < 0) return 0;
  } else {
    alpha = -B - k;
    if (alpha < 0) {
      alpha = -B + k;
      if (alpha < 0) return 0;
    }
  }
  if (alpha > ray.p[0]) return 0;
  contact->pos[0] = ray.pos[0] + alpha * ray.R[2];
  contact->pos[1] = ray.pos[1] + alpha * ray.R[6];
  contact->pos[2] = ray.pos[2] + alpha * ray.R[10];
  const real nsign = (C < 0 || mode) ? OB_REAL(-1.0) : OB_REAL(1.0);
  contact->normal[0] = nsign * (contact->pos[0] - sphere_pos[0]);
  contact->normal[1] = nsign * (contact->pos[1] - sphere_pos[1]);
  contact->normal[2] = nsign * (contact->pos[2] - sphere_pos[2]);
  ob_safe_normalize3(contact->normal);
  contact->depth = alpha;
  return 1;
}
// dCollideRaySphere, ray.cpp:235-251
OB_HD int ob_collide_ray_sphere(const ObPose &o1, const ObPose &o2, ObCg *contact) {
  return ob_ray_sphere_helper(o1, o2.pos, o2.p[0], contact, 0);
}
// dCollideRayBox, ray.cpp:254-350
OB_HD int ob_collide_ray_box(const ObPose &ray, const ObPose &box, ObCg *contact) {
  real tmp[3], s[3], v[3], sign[3];
  tmp[0] = ray.pos[0] - box.pos[0]; tmp[1] = ray.pos[1] - box.pos[1]; tmp[2] = ray.pos[2] - box.pos[2];
  ob_mul1_331(s, box.R, tmp);
  tmp[0] = ray.R[2]; tmp[1] = ray.R[6]; tmp[2] = ray.R[10];
  ob_mul1_331(v, box.R, tmp);
  for (int i = 0; i < 3; i++) {
    if (v[i] < 0) { s[i] = -s[i]; v[i] = -v[i]; sign[i] = 1; }
    else sign[i] = -1;
  }
  real h[3] = {OB_REAL(0.5) * box.p[0], OB_REAL(0.5) * box.p[1], OB_REAL(0.5) * box.p[2]};
  if ((s[0] < -h[0] && v[0] <= 0) || s[0] > h[0] || (s[1] < -h[1] && v[1] <= 0) || s[1] > h[1] ||
      (s[2] < -h[2] && v[2] <= 0) || s[2] > h[2] || (v[0] == 0 && v[1] == 0 && v[2] == 0))
    return 0;
  real lo = -OB_INF, hi = OB_INF;
  int nlo = 0, nhi = 0;
  for (int i = 0; i < 3; i++) {
    if (v[i] != 0) {
      real k = (-h[i] - s[i]) / v[i];
      if (k > lo) { lo = k; nlo = i; }
      k = (h[i] - s[i]) / v[i];
      if (k < hi) { hi = k; nhi = i; }
    }
  }
  if (lo > hi) return 0;
  real alpha;
  int n;
  if (lo >= 0) { alpha = lo; n = nlo; }
  else { alpha = hi; n = nhi; }
  if (alpha < 0 || alpha > ray.p[0]) return 0;
  contact->pos[0] = ray.pos[0] + alpha * ray.R[2];
  contact->pos[1] = ray.pos[1] + alpha * ray.R[6];
  contact->pos[2] = ray.pos[2] + alpha * ray.R[10];
  contact->normal[0] = box.R[0 * 4 + n] * sign[n];
  contact->normal[1] = box.R[1 * 4 + n] * sign[n];
  contact->normal[2] = box.R[2 * 4 + n] * sign[n];
  contact->depth = alpha;
  return 1;
}
// dCollideRayCapsule, ray.cpp:353-470
OB_HD int ob_collide_ray_capsule(const ObPose &ray, const ObPose &ccyl, ObCg *contact) {
  const real radius = ccyl.p[0], lz2 = ccyl.p[1] * OB_REAL(0.5);
  real cs[3], q[3], r[3], C, k;
  cs[0] = ray.pos[0] - ccyl.pos[0]; cs[1] = ray.pos[1] - ccyl.pos[1]; cs[2] = ray.pos[2] - ccyl.pos[2];
  k = ob_dot41(ccyl.R + 2, cs);
  q[0] = k * ccyl.R[2] - cs[0]; q[1] = k * ccyl.R[6] - cs[1]; q[2] = k * ccyl.R[10] - cs[2];
  C = ob_dot(q, q) - radius * radius;
  int inside_ccyl = 0;
  if (C < 0) {
    if (k < -lz2) k = -lz2;
    else if (k > lz2) k = lz2;
    r[0] = ccyl.pos[0] + k * ccyl.R[2]; r[1] = ccyl.pos[1] + k * ccyl.R[6]; r[2] = ccyl.pos[2] + k * ccyl.R[10];
    if ((ray.pos[0] - r[0]) * (ray.pos[0] - r[0]) + (ray.pos[1] - r[1]) * (ray.pos[1] - r[1]) + (ray.pos[2] - r[2]) * (ray.pos[2] - r[2]) <
        radius * radius)
      inside_ccyl = 1;
  }
  if (!inside_ccyl && C < 0) {
    if (k < 0) k = -lz2; else k = lz2;
  } else {
    const real uv = ob_dot44(ccyl.R + 2, ray.R + 2);
    r[0] = uv * ccyl.R[2] - ray.R[2]; r[1] = uv * ccyl.R[6] - ray.R[6]; r[2] = uv * ccyl.R[10] - ray.R[10];
    real A = ob_dot(r, r);
    const real B = 2 * ob_dot(q, r);
    k = B * B - 4 * A * C;
    if (k < 0) {
      if (!inside_ccyl) return 0;
      if (uv < 0) k = -lz2; else k = lz2;
    } else {
      k = ob_sqrt(k);
      A = ob_recip(2 * A);
      real alpha = (-B - k) * A;
      if (alpha < 0) {
        alpha = (-B + k) * A;
        if (alpha < 0) return 0;
      }
      if (alpha > ray.p[0]) return 0;
      contact->pos[0] = ray.pos[0] + alpha * ray.R[2];
      contact->pos[1] = ray.pos[1] + alpha * ray.R[6];
      contact->pos[2] = ray.pos[2] + alpha * ray.R[10];
      q[0] = contact->pos[0] - ccyl.pos[0]; q[1] = contact->pos[1] - ccyl.pos[1]; q[2] = contact->pos[2] - ccyl.pos[2];
      k = ob_dot14(q, ccyl.R + 2);
      const real nsign = inside_ccyl ? OB_REAL(-1.0) : OB_REAL(1.0);
      if (k >= -lz2 && k <= lz2) {
        contact->normal[0] = nsign * (contact->pos[0] - (ccyl.pos[0] + k * ccyl.R[2]));
        contact->normal[1] = nsign * (contact->pos[1] - (ccyl.pos[1] + k * ccyl.R[6]));
        contact->normal[2] = nsign * (contact->pos[2] - (ccyl.pos[2] + k * ccyl.R[10]));
        ob_safe_normalize3(contact->normal);
        contact->depth = alpha;
        return 1;
      }
      if (k < 0) k = -lz2; else k = lz2;
    }
  }
  q[0] = ccyl.pos[0] + k * ccyl.R[2]; q[1] = ccyl.pos[1] + k * ccyl.R[6]; q[2] = ccyl.pos[2] + k * ccyl.R[10];
  return ob_ray_sphere_helper(ray, q, radius, contact, inside_ccyl);
}
// dCollideRayCylinder, ray.cpp:500-620 (ray vs flat cylinder: caps when the ray is parallel to the axis, else the
// mantle between the caps)
OB_HD int ob_collide_ray_cylinder(const ObPose &ray, const ObPose &cyl, ObCg *contact) {
  contact->side1 = -1; contact->side2 = -1;
  const real half_length = cyl.p[1] * OB_REAL(0.5), radius = cyl.p[0], length = ray.p[0];
  real q[3], r[3];
  for (int i = 0; i < 3; i++) r[i] = ray.pos[i] - cyl.pos[i];
  real d = ob_dot41(cyl.R + 2, r);
  for (int i = 0; i < 3; i++) q[i] = (d * cyl.R[i * 4 + 2]) - r[i];
  const real C = ob_dot(q, q) - (radius * radius);
  const real uv = ob_dot44(cyl.R + 2, ray.R + 2);
  for (int i = 0; i < 3; i++) r[i] = (uv * cyl.R[i * 4 + 2]) - ray.R[i * 4 + 2];
  real A = ob_dot(r, r);
  const real B = 2 * ob_dot(q, r);
  real k = B * B - 4 * A * C;
  if (k < OB_EPSILON && C <= 0) {
    // the ray is parallel to the axis and inside the infinite cylinder: it can only meet a cap
    const real uvsign = (uv < 0) ? OB_REAL(-1.0) : OB_REAL(1.0);
    const real internal = (d >= -half_length && d <= +half_length) ? OB_REAL(-1.0) : OB_REAL(1.0);
    if (((uv > 0) && (d + (uvsign * length) < half_length * internal)) || ((uv < 0) && (d + (uvsign * length) > half_length * internal))) return 0;
    contact->depth = ((-uvsign * d) - (internal * half_length));
    for (int i = 0; i < 3; i++) { contact->pos[i] = ray.pos[i] + (contact->depth * ray.R[i * 4 + 2]); contact->normal[i] = uvsign * (cyl.R[i * 4 + 2]); }
    return 1;
  }
  if (k > 0) {
    k = ob_sqrt(k);
    A = ob_recip(2 * A);
    real alpha = (-B - k) * A;
    if (alpha < 0) alpha = (-B + k) * A;
    if (alpha >= 0 && alpha <= length) {
      for (int i = 0; i < 3; i++) contact->pos[i] = ray.pos[i] + (alpha * ray.R[i * 4 + 2]);
      for (int i = 0; i < 3; i++) q[i] = contact->pos[i] - cyl.pos[i];
      d = ob_dot14(q, cyl.R + 2);
      if (d >= -half_length && d <= +half_length) {
        const real nsign = (C < 0) ? OB_REAL(-1.0) : OB_REAL(1.0);
        for (int i = 0; i < 3; i++) contact->normal[i] = nsign * (contact->pos[i] - (cyl.pos[i] + d * cyl.R[i * 4 + 2]));
        ob_safe_normalize3(contact->normal);
        contact->depth = alpha;
        return 1;
      }
    }
  }
  return 0;
}

// dCollideRayPlane, ray.cpp:473-502
OB_HD int ob_collide_ray_plane(const ObPose &ray, const ObPose &plane, ObCg *contact) {
  real alpha = plane.p[3] - ob_dot(plane.p, ray.pos);
  const real nsign = (alpha > 0) ? OB_REAL(-1.0) : OB_REAL(1.0);
  const real k = ob_dot14(plane.p, ray.R + 2);
  if (k == 0) return 0;
  alpha /= k;
  if (alpha < 0 || alpha > ray.p[0]) return 0;
  contact->pos[0] = ray.pos[0] + alpha * ray.R[2];
  contact->pos[1] = ray.pos[1] + alpha * ray.R[6];
  contact->pos[2] = ray.pos[2] + alpha * ray.R[10];
  contact->normal[0] = nsign * plane.p[0];
  contact->normal[1] = nsign * plane.p[1];
  contact->normal[2] = nsign * plane.p[2];
  contact->depth = alpha;
  return 1;
}

// upper bound on contacts a class pair can emit (used to lay out contact slots)
OB_HD int ob_pair_max_contacts(int t1, int t2, int maxc) {
  int lo = t1 < t2 ? t1 : t2, hi = t1 < t2 ? t2 : t1;
  int cap;
  if (hi == OB_GEOM_RAY) cap = (lo == OB_GEOM_SPHERE || lo == OB_GEOM_BOX || lo == OB_GEOM_CAPSULE || lo == OB_GEOM_CYLINDER || lo == OB_GEOM_PLANE) ? 1 : 0;
  else if (hi == OB_GEOM_TRIMESH) cap = (lo == OB_GEOM_SPHERE || lo == OB_GEOM_BOX || lo == OB_GEOM_CAPSULE || lo == OB_GEOM_PLANE || lo == OB_GEOM_RAY) ? (1 << 15) : 0;   // bounded by the caller's max_contacts only
  else if (lo == OB_GEOM_SPHERE) cap = (hi == OB_GEOM_SPHERE || hi == OB_GEOM_BOX || hi == OB_GEOM_PLANE || hi == OB_GEOM_CAPSULE || hi == OB_GEOM_CYLINDER) ? 1 : 0;
  else if (lo == OB_GEOM_CYLINDER && hi == OB_GEOM_PLANE) cap = 4;
  else if (lo == OB_GEOM_BOX && hi == OB_GEOM_CYLINDER) cap = 16;   // a clipped face polygon: bounded by the caller's max_contacts / the contact buffer
  else if (lo == OB_GEOM_BOX && hi == OB_GEOM_BOX) cap = 8;
  else if (lo == OB_GEOM_BOX && hi == OB_GEOM_PLANE) cap = 4;
  else if (lo == OB_GEOM_BOX && hi == OB_GEOM_CAPSULE) cap = 1;
  else if (lo == OB_GEOM_CAPSULE && hi == OB_GEOM_CAPSULE) cap = 2;
  else if (lo == OB_GEOM_CAPSULE && hi == OB_GEOM_PLANE) cap = 2;
  else cap = 0;
  return cap < maxc ? cap : maxc;
}

// dCollide for primitive class pairs: table lookup + reverse fix-up.
// Returns the contact count; `swapped` tells the caller g1/g2 were exchanged.
// MESH = false compiles the trimesh arms out (kernels for worlds without trimesh geoms); CGCAP = size of c[]
template <bool MESH, int CGCAP>
OB_HD int ob_collide_pair_t(const ObPose &o1, const ObPose &o2, int flags, ObCg *c, int *swapped, const ObMeshDev *meshes, int *bverr) {
  int t1 = o1.type, t2 = o2.type, n = 0, rev = 0;
  int bve = 0;
  for (int i = 0; i < CGCAP; i++) { c[i].side1 = -1; c[i].side2 = -1; }
  if (t1 == OB_GEOM_SPHERE && t2 == OB_GEOM_SPHERE) n = ob_collide_spheres(o1.pos, o1.p[0], o2.pos, o2.p[0], c);
  else if (t1 == OB_GEOM_SPHERE && t2 == OB_GEOM_BOX) n = ob_collide_sphere_box(o1, o2, c);
  else if (t1 == OB_GEOM_BOX && t2 == OB_GEOM_SPHERE) { n = ob_collide_sphere_box(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_SPHERE && t2 == OB_GEOM_PLANE) n = ob_collide_sphere_plane(o1, o2, c);
  else if (t1 == OB_GEOM_PLANE && t2 == OB_GEOM_SPHERE) { n = ob_collide_sphere_plane(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_BOX && t2 == OB_GEOM_BOX) n = ob_collide_box_box(o1, o2, flags, c);
  else if (t1 == OB_GEOM_BOX && t2 == OB_GEOM_PLANE) n = ob_collide_box_plane(o1, o2, flags, c);
  else if (t1 == OB_GEOM_PLANE && t2 == OB_GEOM_BOX) { n = ob_collide_box_plane(o2, o1, flags, c); rev = 1; }
  else if (t1 == OB_GEOM_CAPSULE && t2 == OB_GEOM_SPHERE) n = ob_collide_capsule_sphere(o1, o2, c);
  else if (t1 == OB_GEOM_SPHERE && t2 == OB_GEOM_CAPSULE) { n = ob_collide_capsule_sphere(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_CAPSULE && t2 == OB_GEOM_BOX) n = ob_collide_capsule_box(o1, o2, c);
  else if (t1 == OB_GEOM_BOX && t2 == OB_GEOM_CAPSULE) { n = ob_collide_capsule_box(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_CAPSULE && t2 == OB_GEOM_CAPSULE) n = ob_collide_capsule_capsule(o1, o2, flags, c);
  else if (t1 == OB_GEOM_CAPSULE && t2 == OB_GEOM_PLANE) n = ob_collide_capsule_plane(o1, o2, flags, c);
  else if (t1 == OB_GEOM_PLANE && t2 == OB_GEOM_CAPSULE) { n = ob_collide_capsule_plane(o2, o1, flags, c); rev = 1; }
  else if (t1 == OB_GEOM_CYLINDER && t2 == OB_GEOM_PLANE) n = ob_collide_cylinder_plane(o1, o2, flags, c);
  else if (t1 == OB_GEOM_PLANE && t2 == OB_GEOM_CYLINDER) { n = ob_collide_cylinder_plane(o2, o1, flags, c); rev = 1; }
  else if (t1 == OB_GEOM_CYLINDER && t2 == OB_GEOM_BOX) n = ob_collide_cylinder_box(o1, o2, (flags & ~0xffff) | ((flags & 0xffff) < CGCAP ? (flags & 0xffff) : CGCAP), c);
  else if (t1 == OB_GEOM_BOX && t2 == OB_GEOM_CYLINDER) { n = ob_collide_cylinder_box(o2, o1, (flags & ~0xffff) | ((flags & 0xffff) < CGCAP ? (flags & 0xffff) : CGCAP), c); rev = 1; }
  else if (t1 == OB_GEOM_CYLINDER && t2 == OB_GEOM_SPHERE) n = ob_collide_cylinder_sphere(o1, o2, c);
  else if (t1 == OB_GEOM_SPHERE && t2 == OB_GEOM_CYLINDER) { n = ob_collide_cylinder_sphere(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_RAY && t2 == OB_GEOM_SPHERE) n = ob_collide_ray_sphere(o1, o2, c);
  else if (t1 == OB_GEOM_SPHERE && t2 == OB_GEOM_RAY) { n = ob_collide_ray_sphere(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_RAY && t2 == OB_GEOM_BOX) n = ob_collide_ray_box(o1, o2, c);
  else if (t1 == OB_GEOM_BOX && t2 == OB_GEOM_RAY) { n = ob_collide_ray_box(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_RAY && t2 == OB_GEOM_CAPSULE) n = ob_collide_ray_capsule(o1, o2, c);
  else if (t1 == OB_GEOM_CAPSULE && t2 == OB_GEOM_RAY) { n = ob_collide_ray_capsule(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_RAY && t2 == OB_GEOM_CYLINDER) n = ob_collide_ray_cylinder(o1, o2, c);
  else if (t1 == OB_GEOM_CYLINDER && t2 == OB_GEOM_RAY) { n = ob_collide_ray_cylinder(o2, o1, c); rev = 1; }
  else if (t1 == OB_GEOM_RAY && t2 == OB_GEOM_PLANE) n = ob_collide_ray_plane(o1, o2, c);
  else if (t1 == OB_GEOM_PLANE && t2 == OB_GEOM_RAY) { n = ob_collide_ray_plane(o2, o1, c); rev = 1; }
  else if (MESH && t1 == OB_GEOM_TRIMESH && t2 == OB_GEOM_PLANE) n = ob_collide_trimesh_plane(o1, o2, meshes[o1.mesh], (flags & ~0xffff) | ((flags & 0xffff) < CGCAP ? (flags & 0xffff) : CGCAP), c);
  else if (MESH && t1 == OB_GEOM_PLANE && t2 == OB_GEOM_TRIMESH) { n = ob_collide_trimesh_plane(o2, o1, meshes[o2.mesh], (flags & ~0xffff) | ((flags & 0xffff) < CGCAP ? (flags & 0xffff) : CGCAP), c); rev = 1; }
  else if (MESH && t1 == OB_GEOM_TRIMESH && t2 == OB_GEOM_CAPSULE) n = ob_collide_trimesh_capsule(o1, o2, meshes[o1.mesh], flags, c, &bve);
  else if (MESH && t1 == OB_GEOM_CAPSULE && t2 == OB_GEOM_TRIMESH) { n = ob_collide_trimesh_capsule(o2, o1, meshes[o2.mesh], flags, c, &bve); rev = 1; }
  else if (MESH && t1 == OB_GEOM_TRIMESH && t2 == OB_GEOM_RAY) n = ob_collide_trimesh_ray(o1, o2, meshes[o1.mesh], flags, c, &bve);
  else if (MESH && t1 == OB_GEOM_RAY && t2 == OB_GEOM_TRIMESH) { n = ob_collide_trimesh_ray(o2, o1, meshes[o2.mesh], flags, c, &bve); rev = 1; }
  else if (MESH && t1 == OB_GEOM_TRIMESH && t2 == OB_GEOM_SPHERE) n = ob_collide_trimesh_sphere(o1, o2, meshes[o1.mesh], flags, c, &bve);
  else if (MESH && t1 == OB_GEOM_SPHERE && t2 == OB_GEOM_TRIMESH) { n = ob_collide_trimesh_sphere(o2, o1, meshes[o2.mesh], flags, c, &bve); rev = 1; }
  else if (MESH && t1 == OB_GEOM_TRIMESH && t2 == OB_GEOM_BOX) n = ob_collide_trimesh_box(o1, o2, meshes[o1.mesh], flags, c, &bve);
  else if (MESH && t1 == OB_GEOM_BOX && t2 == OB_GEOM_TRIMESH) { n = ob_collide_trimesh_box(o2, o1, meshes[o2.mesh], flags, c, &bve); rev = 1; }
  if (bve && bverr) *bverr = 1;
  if (rev) {
    for (int i = 0; i < n; i++) {
      c[i].normal[0] = -c[i].normal[0]; c[i].normal[1] = -c[i].normal[1]; c[i].normal[2] = -c[i].normal[2];
      int t = c[i].side1; c[i].side1 = c[i].side2; c[i].side2 = t;
    }
  }
  *swapped = rev;
  return n;
}
// Geom transforms (collision_transform.cpp:115-160; setAllColliders(dGeomTransformClass, ..), collision_kernel.cpp:267):
// a transform arrives as its encapsulated geom posed at T o local, tagged OB_POSE_XFORM.  The transform class's table
// entries give dCollide(X, T) = reverse(dCollideTransform(T, X)) = reverse(dCollide(inner, X)) and dCollide(T1, T2) =
// dCollide(inner1, T2) = reverse(dCollide(inner2, inner1)): whenever the SECOND geom is a transform the pair is
// collided the other way round and reversed once more (box-box is not symmetric bit for bit).  any_xf is the
// batch-wide "some geom is a transform" flag (ObBatchDev::any_xf), uniform per launch, so batches without transforms
// pay one uniform branch and no dependent load.
OB_HD bool ob_pose_is_xform(const ObPose &p) { return p.type != OB_GEOM_TRIMESH && p.type != OB_GEOM_RAY && (p.mesh & OB_POSE_XFORM); }
// The reversed path is a real (not inlined) call: batches without transforms keep exactly the code they had, behind
// one uniform branch on the kernel parameter, and the rare path does not double the size of the collide kernels.
template <bool MESH, int CGCAP>
OB_HDN int ob_collide_pair_flipped_t(const ObPose *a1, const ObPose *a2, int flags, ObCg *c, int *swapped, const ObMeshDev *meshes, int *bverr) {
  const int n = ob_collide_pair_t<MESH, CGCAP>(*a2, *a1, flags, c, swapped, meshes, bverr);
  for (int i = 0; i < n; i++) {
    c[i].normal[0] = -c[i].normal[0]; c[i].normal[1] = -c[i].normal[1]; c[i].normal[2] = -c[i].normal[2];
    int t = c[i].side1; c[i].side1 = c[i].side2; c[i].side2 = t;
  }
  *swapped ^= 1;
  return n;
}
template <bool MESH, int CGCAP>
OB_HD int ob_collide_pair_xf_t(const ObPose *a1, const ObPose *a2, int any_xf, int flags, ObCg *c, int *swapped, const ObMeshDev *meshes, int *bverr) {
  if (any_xf && ob_pose_is_xform(*a2)) return ob_collide_pair_flipped_t<MESH, CGCAP>(a1, a2, flags, c, swapped, meshes, bverr);
  return ob_collide_pair_t<MESH, CGCAP>(*a1, *a2, flags, c, swapped, meshes, bverr);
}
// XF is a compile-time property of the launch (kernels are instantiated <MESH, XF>; a batch with transforms runs the
// <true, true> instantiation): measured on B200, even a never-taken call to the flipped path cost k_collide 2 % (config 2)
// to 7 % (config 4) through register allocation (80 -> 96 registers), so batches without transforms run the code they
// had before (profiles/job_ab.sh).
// Row of the contact-policy table that serves a pair (dBatchContactPolicy, ode.h): the first row whose category masks accept
// the two geoms in either order; -1 = no row, the pair gets no contacts (a near callback that returns early).  A table of one
// row serves every pair, as a callback without a class test does (the masks of a single row are not consulted).
OB_HD int ob_policy_row(const ObPolicy *tab, uint32_t cat1, uint32_t cat2) {
  const int n = tab[0].nrows;
  if (n <= 1) return 0;
  for (int r = 0; r < n && r < OB_MAXPOLICY; r++) {
    const uint32_t m1 = tab[r].cat_mask1, m2 = tab[r].cat_mask2;
    if (((cat1 & m1) && (cat2 & m2)) || ((cat2 & m1) && (cat1 & m2))) return r;
  }
  return -1;
}
template <bool MESH, int CGCAP, bool XF>
OB_HD int ob_collide_pair_sel_t(const ObPose *a1, const ObPose *a2, int flags, ObCg *c, int *swapped, const ObMeshDev *meshes, int *bverr) {
  if (XF) return ob_collide_pair_xf_t<MESH, CGCAP>(a1, a2, 1, flags, c, swapped, meshes, bverr);
  return ob_collide_pair_t<MESH, CGCAP>(*a1, *a2, flags, c, swapped, meshes, bverr);
}
OB_HD int ob_collide_pair(const ObPose &o1, const ObPose &o2, int flags, ObCg *c, int *swapped, const ObMeshDev *meshes = 0,
                          int *bverr = 0) {
  return ob_collide_pair_xf_t<true, OB_MAXC_LOCAL>(&o1, &o2, 1, flags, c, swapped, meshes, bverr);
}


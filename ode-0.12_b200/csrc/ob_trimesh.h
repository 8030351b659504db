// ob_trimesh.h — OPCODE-equivalent trimesh colliders, per-thread device functions.
//
// The mesh is a flattened "no-leaf" AABB tree (the layout OPCODE's AABBNoLeafTree has,
// OPCODE/OPC_OptimizedTree.cpp:150-202,325-345) built on the host at dGeomTriMeshDataBuild* time
// (ob_trimesh_build.cpp) with the reference's own float arithmetic, so node boxes and — what the
// contact set depends on — the depth-first triangle visit order are identical.  Volume queries run
// in the mesh's model space in float, exactly like OPCODE; contact generation runs in dReal like
// the ODE colliders that consume the touched-triangle list:
//   sphere  : dCollideSTL  ode/src/collision_trimesh_sphere.cpp:244-538 over
//             SphereCollider::_CollideNoPrimitiveTest OPCODE/OPC_SphereCollider.cpp:474-488
//             (primitive tests are off, collision_trimesh_opcode.cpp:46)
//   box     : ob_trimesh_box.h
// Traversal is pos-then-neg depth first with an explicit stack; a volume that fully contains a node
// box dumps the whole subtree (OPC_VolumeCollider.cpp:70-80).  Triangles are consumed in visit order
// until the caller's max-contacts is reached, as the reference does.
#pragma once
#include "ob_collide_types.h"

struct ObBvNode {      // 32 bytes
  float c[3], e[3];    // box centre / extents (model space)
  uint32_t pos, neg;   // (triangle << 1) | 1  or  (node index << 1)
};
struct ObMeshDev {     // one dTriMeshData on the execution side
  const float *verts;  // [nverts*3]
  const int *tris;     // [ntris*3]
  const ObBvNode *nodes;   // [ntris-1]
  int nverts, ntris;
  real aabbc[3], aabbe[3];   // model-space AABB centre / extents (collision_trimesh_opcode.cpp:123-157)
};

#define OB_BV_STACK 96
#define OB_BV_DUMP 0x80000000u
struct ObBvIter {
  uint32_t stack[OB_BV_STACK];
  int sp;
  int overflow;
};
OB_HD void ob_bv_begin(ObBvIter &it) { it.sp = 0; it.overflow = 0; it.stack[it.sp++] = 0u; }
// next touched triangle in the reference's visit order, or -1.  Q: overlap(node), contains(node), prim(mesh, tri)
template <class Q>
OB_HD int ob_bv_next(const ObMeshDev &m, ObBvIter &it, const Q &q) {
  while (it.sp > 0) {
    const uint32_t item = it.stack[--it.sp];
    const uint32_t dump = item & OB_BV_DUMP, ref = item & ~OB_BV_DUMP;
    if (ref & 1u) {   // leaf: primitive test unless the subtree is being dumped
      if (dump || q.prim(m, (int)(ref >> 1))) return (int)(ref >> 1);
      continue;
    }
    const ObBvNode nd = m.nodes[ref >> 1];
    uint32_t d = dump;
    if (!d) {
      if (!q.overlap(nd)) continue;
      if (q.contains(nd)) d = OB_BV_DUMP;
    }
    if (it.sp + 2 > OB_BV_STACK) { it.overflow = 1; continue; }
    it.stack[it.sp++] = nd.neg | d;
    it.stack[it.sp++] = nd.pos | d;
  }
  return -1;
}

// FetchTriangle, collision_trimesh_internal.h:394-411: vertices (float) to world space in dReal
OB_HD void ob_fetch_triangle(const ObMeshDev &m, int tri, const real *pos, const real *R, real dv[3][3]) {
  for (int i = 0; i < 3; i++) {
    const float *p = m.verts + 3 * (size_t)m.tris[3 * (size_t)tri + i];
    real v[3] = {(real)p[0], (real)p[1], (real)p[2]};
    ob_mul0_331(dv[i], R, v);
    dv[i][0] += pos[0]; dv[i][1] += pos[1]; dv[i][2] += pos[2];
  }
}

// world -> model transform of a point the way OPCODE does it: MakeMatrix (float casts,
// collision_trimesh_internal.h:419-442), InvertPRMatrix (Ice/IceMatrix4x4.cpp:54-75), Point *= Matrix4x4
struct ObInvPR { float r[9]; float t[3]; };   // r[3*i+j] = dest.m[i][j], t[j] = dest.m[3][j]
OB_HD void ob_inv_pr(const real *pos, const real *R, ObInvPR *o) {
  const float p0 = (float)pos[0], p1 = (float)pos[1], p2 = (float)pos[2];
  // src.m[i][j] = (float)R[4*j+i];  dest.m[i][j] = src.m[j][i] = (float)R[4*i+j]
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o->r[3 * i + j] = (float)R[4 * i + j];
  // dest.m[3][j] = -(src.m[3][0]*src.m[j][0] + src.m[3][1]*src.m[j][1] + src.m[3][2]*src.m[j][2]), src.m[j][k] = (float)R[4*k+j]
  for (int j = 0; j < 3; j++) o->t[j] = -(p0 * (float)R[j] + p1 * (float)R[4 + j] + p2 * (float)R[8 + j]);
}
OB_HD void ob_point_mul(const ObInvPR &M, float *p) {
  const float x = p[0], y = p[1], z = p[2];
  const float xp = x * M.r[0] + y * M.r[3] + z * M.r[6] + M.t[0];
  const float yp = x * M.r[1] + y * M.r[4] + z * M.r[7] + M.t[1];
  const float zp = x * M.r[2] + y * M.r[5] + z * M.r[8] + M.t[2];
  p[0] = xp; p[1] = yp; p[2] = zp;
}

// ---- sphere query (OPC_SphereCollider.cpp:179-216, OPC_SphereAABBOverlap.h, SphereContainsBox :322-338)
struct ObSphereQuery {
  float c[3], r2;
  OB_HD bool overlap(const ObBvNode &n) const {
    float d = 0.0f, tmp, s;
    tmp = c[0] - n.c[0]; s = tmp + n.e[0];
    if (s < 0.0f) { d += s * s; if (d > r2) return false; }
    else { s = tmp - n.e[0]; if (s > 0.0f) { d += s * s; if (d > r2) return false; } }
    tmp = c[1] - n.c[1]; s = tmp + n.e[1];
    if (s < 0.0f) { d += s * s; if (d > r2) return false; }
    else { s = tmp - n.e[1]; if (s > 0.0f) { d += s * s; if (d > r2) return false; } }
    tmp = c[2] - n.c[2]; s = tmp + n.e[2];
    if (s < 0.0f) { d += s * s; if (d > r2) return false; }
    else { s = tmp - n.e[2]; if (s > 0.0f) { d += s * s; if (d > r2) return false; } }
    return d <= r2;
  }
  OB_HD bool prim(const ObMeshDev &, int) const { return true; }   // SetPrimitiveTests(false)
  OB_HD float sqd(float px, float py, float pz) const {
    return ((c[0] - px) * (c[0] - px) + (c[1] - py) * (c[1] - py) + (c[2] - pz) * (c[2] - pz));
  }
  OB_HD bool contains(const ObBvNode &n) const {
    const float xp = n.c[0] + n.e[0], xm = n.c[0] - n.e[0], yp = n.c[1] + n.e[1], ym = n.c[1] - n.e[1];
    const float zp = n.c[2] + n.e[2], zm = n.c[2] - n.e[2];
    if (sqd(xp, yp, zp) >= r2) return false;
    if (sqd(xm, yp, zp) >= r2) return false;
    if (sqd(xp, ym, zp) >= r2) return false;
    if (sqd(xm, ym, zp) >= r2) return false;
    if (sqd(xp, yp, zm) >= r2) return false;
    if (sqd(xm, yp, zm) >= r2) return false;
    if (sqd(xp, ym, zm) >= r2) return false;
    if (sqd(xm, ym, zm) >= r2) return false;
    return true;
  }
};

// GetContactData, collision_trimesh_sphere.cpp:42-242 (closest point on triangle, 7 regions)
OB_HD bool ob_stl_contact_data(const real *Center, real Radius, const real *Origin, const real *Edge0, const real *Edge1,
                               real *Dist, real *pu, real *pv) {
  real Diff[3] = {Origin[0] - Center[0], Origin[1] - Center[1], Origin[2] - Center[2]};
  const real A00 = ob_dot(Edge0, Edge0), A01 = ob_dot(Edge0, Edge1), A11 = ob_dot(Edge1, Edge1);
  const real B0 = ob_dot(Diff, Edge0), B1 = ob_dot(Diff, Edge1);
  const real C = ob_dot(Diff, Diff);
  const real Det = ob_fabs(A00 * A11 - A01 * A01);
  real u = A01 * B1 - A11 * B0;
  real v = A01 * B0 - A00 * B1;
  real DistSq;
  if (u + v <= Det) {
    if (u < OB_REAL(0.0)) {
      if (v < OB_REAL(0.0)) {   // region 4
        if (B0 < OB_REAL(0.0)) {
          v = OB_REAL(0.0);
          if (-B0 >= A00) { u = OB_REAL(1.0); DistSq = A00 + OB_REAL(2.0) * B0 + C; }
          else { u = -B0 / A00; DistSq = B0 * u + C; }
        } else {
          u = OB_REAL(0.0);
          if (B1 >= OB_REAL(0.0)) { v = OB_REAL(0.0); DistSq = C; }
          else if (-B1 >= A11) { v = OB_REAL(1.0); DistSq = A11 + OB_REAL(2.0) * B1 + C; }
          else { v = -B1 / A11; DistSq = B1 * v + C; }
        }
      } else {   // region 3
        u = OB_REAL(0.0);
        if (B1 >= OB_REAL(0.0)) { v = OB_REAL(0.0); DistSq = C; }
        else if (-B1 >= A11) { v = OB_REAL(1.0); DistSq = A11 + OB_REAL(2.0) * B1 + C; }
        else { v = -B1 / A11; DistSq = B1 * v + C; }
      }
    } else if (v < OB_REAL(0.0)) {   // region 5
      v = OB_REAL(0.0);
      if (B0 >= OB_REAL(0.0)) { u = OB_REAL(0.0); DistSq = C; }
      else if (-B0 >= A00) { u = OB_REAL(1.0); DistSq = A00 + OB_REAL(2.0) * B0 + C; }
      else { u = -B0 / A00; DistSq = B0 * u + C; }
    } else {   // region 0
      if (Det == OB_REAL(0.0)) { u = OB_REAL(0.0); v = OB_REAL(0.0); DistSq = (real)3.402823466e+38f; }
      else {
        const real InvDet = OB_REAL(1.0) / Det;
        u *= InvDet; v *= InvDet;
        DistSq = u * (A00 * u + A01 * v + OB_REAL(2.0) * B0) + v * (A01 * u + A11 * v + OB_REAL(2.0) * B1) + C;
      }
    }
  } else {
    real Tmp0, Tmp1, Numer, Denom;
    if (u < OB_REAL(0.0)) {   // region 2
      Tmp0 = A01 + B0; Tmp1 = A11 + B1;
      if (Tmp1 > Tmp0) {
        Numer = Tmp1 - Tmp0; Denom = A00 - OB_REAL(2.0) * A01 + A11;
        if (Numer >= Denom) { u = OB_REAL(1.0); v = OB_REAL(0.0); DistSq = A00 + OB_REAL(2.0) * B0 + C; }
        else {
          u = Numer / Denom; v = OB_REAL(1.0) - u;
          DistSq = u * (A00 * u + A01 * v + OB_REAL(2.0) * B0) + v * (A01 * u + A11 * v + OB_REAL(2.0) * B1) + C;
        }
      } else {
        u = OB_REAL(0.0);
        if (Tmp1 <= OB_REAL(0.0)) { v = OB_REAL(1.0); DistSq = A11 + OB_REAL(2.0) * B1 + C; }
        else if (B1 >= OB_REAL(0.0)) { v = OB_REAL(0.0); DistSq = C; }
        else { v = -B1 / A11; DistSq = B1 * v + C; }
      }
    } else if (v < OB_REAL(0.0)) {   // region 6
      Tmp0 = A01 + B1; Tmp1 = A00 + B0;
      if (Tmp1 > Tmp0) {
        Numer = Tmp1 - Tmp0; Denom = A00 - OB_REAL(2.0) * A01 + A11;
        if (Numer >= Denom) { v = OB_REAL(1.0); u = OB_REAL(0.0); DistSq = A11 + OB_REAL(2.0) * B1 + C; }
        else {
          v = Numer / Denom; u = OB_REAL(1.0) - v;
          DistSq = u * (A00 * u + A01 * v + OB_REAL(2.0) * B0) + v * (A01 * u + A11 * v + OB_REAL(2.0) * B1) + C;
        }
      } else {
        v = OB_REAL(0.0);
        if (Tmp1 <= OB_REAL(0.0)) { u = OB_REAL(1.0); DistSq = A00 + OB_REAL(2.0) * B0 + C; }
        else if (B0 >= OB_REAL(0.0)) { u = OB_REAL(0.0); DistSq = C; }
        else { u = -B0 / A00; DistSq = B0 * u + C; }
      }
    } else {   // region 1
      Numer = A11 + B1 - A01 - B0;
      if (Numer <= OB_REAL(0.0)) { u = OB_REAL(0.0); v = OB_REAL(1.0); DistSq = A11 + OB_REAL(2.0) * B1 + C; }
      else {
        Denom = A00 - OB_REAL(2.0) * A01 + A11;
        if (Numer >= Denom) { u = OB_REAL(1.0); v = OB_REAL(0.0); DistSq = A00 + OB_REAL(2.0) * B0 + C; }
        else {
          u = Numer / Denom; v = OB_REAL(1.0) - u;
          DistSq = u * (A00 * u + A01 * v + OB_REAL(2.0) * B0) + v * (A01 * u + A11 * v + OB_REAL(2.0) * B1) + C;
        }
      }
    }
  }
  real d = ob_sqrt(ob_fabs(DistSq));
  *pu = u; *pv = v;
  if (d <= Radius) { *Dist = Radius - d; return true; }
  *Dist = d;
  return false;
}

// dCollideSTL, collision_trimesh_sphere.cpp:244-538 (default: no contact merging,
// collision_trimesh_internal.h:343-345).  o1 = trimesh, o2 = sphere.  *bverr: traversal stack overflow
OB_HD int ob_collide_trimesh_sphere(const ObPose &o1, const ObPose &o2, const ObMeshDev &m, int flags, ObCg *contact, int *bverr) {
  const int maxc = flags & 0xffff;
  const real *TLPosition = o1.pos, *TLRotation = o1.R, *Position = o2.pos;
  const real Radius = o2.p[0];
  ObSphereQuery q;
  {
    const float sr = (float)Radius;
    q.r2 = sr * sr;
    q.c[0] = (float)Position[0]; q.c[1] = (float)Position[1]; q.c[2] = (float)Position[2];
    ObInvPR inv;
    ob_inv_pr(TLPosition, TLRotation, &inv);
    ob_point_mul(inv, q.c);
  }
  ObBvIter it;
  ob_bv_begin(it);
  int out = 0;
  for (;;) {
    if (out == maxc) break;
    const int tri = ob_bv_next(m, it, q);
    if (tri < 0) break;
    real dv[3][3];
    ob_fetch_triangle(m, tri, TLPosition, TLRotation, dv);
    const real *v0 = dv[0], *v1 = dv[1], *v2 = dv[2];
    real vu[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
    real vv[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    real Plane[3];
    ob_cross(Plane, vu, vv);
    if (!ob_safe_normalize3(Plane)) continue;
    const real side = ob_dot(Plane, Position) - ob_dot(Plane, v0);
    if (side < OB_REAL(0.0)) continue;
    real Depth, u, v;
    if (!ob_stl_contact_data(Position, Radius, v0, vu, vv, &Depth, &u, &v)) continue;
    if (Depth < OB_REAL(0.0)) continue;
    real ContactPos[3];
    const real w = OB_REAL(1.0) - u - v;
    ContactPos[0] = (v0[0] * w) + (v1[0] * u) + (v2[0] * v);
    ContactPos[1] = (v0[1] * w) + (v1[1] * u) + (v2[1] * v);
    ContactPos[2] = (v0[2] * w) + (v1[2] * u) + (v2[2] * v);
    real dir[3] = {Position[0] - ContactPos[0], Position[1] - ContactPos[1], Position[2] - ContactPos[2]};
    const real dirProj = ob_dot(dir, Plane) / ob_sqrt(ob_dot(dir, dir));
    if (dirProj < OB_REAL(0.0)) continue;
    ObCg *c = contact + out;
    c->pos[0] = ContactPos[0]; c->pos[1] = ContactPos[1]; c->pos[2] = ContactPos[2];
    c->normal[0] = -Plane[0]; c->normal[1] = -Plane[1]; c->normal[2] = -Plane[2];
    c->depth = Depth * dirProj;
    c->side1 = tri; c->side2 = -1;
    out++;
  }
  if (it.overflow) *bverr = 1;
  return out;
}

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
  const int *vfirst;   // [nverts] flattened corner index (3*tri + corner) of the vertex's first use, -1 if unused
  const unsigned char *useflags;   // [ntris] dxTriMeshData::UseFlags (kEdge0..2 = 1,2,4; kVert0..2 = 8,16,32), null = kUseAll
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

// one candidate triangle of dCollideSTL's loop (collision_trimesh_sphere.cpp:318-420): 1 and the contact in *c, or 0
OB_HD int ob_stl_triangle(const ObMeshDev &m, int tri, const real *TLPosition, const real *TLRotation, const real *Position, real Radius, ObCg *c) {
  real dv[3][3];
  ob_fetch_triangle(m, tri, TLPosition, TLRotation, dv);
  const real *v0 = dv[0], *v1 = dv[1], *v2 = dv[2];
  real vu[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
  real vv[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
  real Plane[3];
  ob_cross(Plane, vu, vv);
  if (!ob_safe_normalize3(Plane)) return 0;
  const real side = ob_dot(Plane, Position) - ob_dot(Plane, v0);
  if (side < OB_REAL(0.0)) return 0;
  real Depth, u, v;
  if (!ob_stl_contact_data(Position, Radius, v0, vu, vv, &Depth, &u, &v)) return 0;
  if (Depth < OB_REAL(0.0)) return 0;
  real ContactPos[3];
  const real w = OB_REAL(1.0) - u - v;
  ContactPos[0] = (v0[0] * w) + (v1[0] * u) + (v2[0] * v);
  ContactPos[1] = (v0[1] * w) + (v1[1] * u) + (v2[1] * v);
  ContactPos[2] = (v0[2] * w) + (v1[2] * u) + (v2[2] * v);
  real dir[3] = {Position[0] - ContactPos[0], Position[1] - ContactPos[1], Position[2] - ContactPos[2]};
  const real dirProj = ob_dot(dir, Plane) / ob_sqrt(ob_dot(dir, dir));
  if (dirProj < OB_REAL(0.0)) return 0;
  c->pos[0] = ContactPos[0]; c->pos[1] = ContactPos[1]; c->pos[2] = ContactPos[2];
  c->normal[0] = -Plane[0]; c->normal[1] = -Plane[1]; c->normal[2] = -Plane[2];
  c->depth = Depth * dirProj;
  c->side1 = tri; c->side2 = -1;
  return 1;
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
  // the walk and the triangle test alternate in lock-step over the lanes that came here together (OB_ALL_LANES): every lane
  // visits its triangles in the reference's order, the warp runs each of the two regions with all the lanes that are in it
  const unsigned together = OB_LANES_TOGETHER();
  bool fin = false;
  for (;;) {
    int tri = -1;
    if (!fin) {
      if (out == maxc) fin = true;
      else { tri = ob_bv_next(m, it, q); fin = tri < 0; }
    }
    if (OB_ALL_LANES(together, fin)) break;
    if (tri >= 0) out += ob_stl_triangle(m, tri, TLPosition, TLRotation, Position, Radius, contact + out);
  }
  if (it.overflow) *bverr = 1;
  return out;
}


// ---- ray query: dCollideRTL, collision_trimesh_ray.cpp:37-150 over RayCollider::_SegmentStab
// (OPCODE/OPC_RayCollider.cpp:263-460, :550-568, OPC_RayAABBOverlap.h, OPC_RayTriOverlap.h).
// o1 = trimesh, o2 = ray (o2.p[0] = length, o2.mesh = OB_RAY_* flag bits).
enum { OB_RAY_FIRSTCONTACT = 1, OB_RAY_BACKFACECULL = 2, OB_RAY_CLOSEST_HIT = 4 };
struct ObRayQuery {
  float dir[3], org[3];        // model space
  float data[3], data2[3], fdir[3];
  float maxdist;
  int culling;
  // result of the last successful primitive test (mStabbedFace)
  mutable float dist, u, v;
  OB_HD bool overlap(const ObBvNode &n) const {   // SegmentAABBOverlap
    const float Dx = data2[0] - n.c[0]; if (fabsf(Dx) > n.e[0] + fdir[0]) return false;
    const float Dy = data2[1] - n.c[1]; if (fabsf(Dy) > n.e[1] + fdir[1]) return false;
    const float Dz = data2[2] - n.c[2]; if (fabsf(Dz) > n.e[2] + fdir[2]) return false;
    float f;
    f = data[1] * Dz - data[2] * Dy; if (fabsf(f) > n.e[1] * fdir[2] + n.e[2] * fdir[1]) return false;
    f = data[2] * Dx - data[0] * Dz; if (fabsf(f) > n.e[0] * fdir[2] + n.e[2] * fdir[0]) return false;
    f = data[0] * Dy - data[1] * Dx; if (fabsf(f) > n.e[0] * fdir[1] + n.e[1] * fdir[0]) return false;
    return true;
  }
  OB_HD bool contains(const ObBvNode &) const { return false; }
  // SEGMENT_PRIM: RayTriOverlap (integer-representation compares, IceFPU.h:40) + distance < segment length
  OB_HD bool prim(const ObMeshDev &m, int tri) const {
    const float *v0 = m.verts + 3 * (size_t)m.tris[3 * (size_t)tri], *v1 = m.verts + 3 * (size_t)m.tris[3 * (size_t)tri + 1],
                *v2 = m.verts + 3 * (size_t)m.tris[3 * (size_t)tri + 2];
    const float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    const float pv[3] = {dir[1] * e2[2] - dir[2] * e2[1], dir[2] * e2[0] - dir[0] * e2[2], dir[0] * e2[1] - dir[1] * e2[0]};
    const float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    const float LOCAL_EPSILON = 0.000001f;
    if (culling) {
      if (det < LOCAL_EPSILON) return false;
      const float tv[3] = {org[0] - v0[0], org[1] - v0[1], org[2] - v0[2]};
      u = tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2];
      if (((uint32_t)ob_f2i(u) & 0x80000000u) || (uint32_t)ob_f2i(u) > (uint32_t)ob_f2i(det)) return false;
      const float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
      v = dir[0] * qv[0] + dir[1] * qv[1] + dir[2] * qv[2];
      if (((uint32_t)ob_f2i(v) & 0x80000000u) || u + v > det) return false;
      dist = e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2];
      if ((uint32_t)ob_f2i(dist) & 0x80000000u) return false;
      const float ood = 1.0f / det;
      dist *= ood; u *= ood; v *= ood;
    } else {
      if (det > -LOCAL_EPSILON && det < LOCAL_EPSILON) return false;
      const float ood = 1.0f / det;
      const float tv[3] = {org[0] - v0[0], org[1] - v0[1], org[2] - v0[2]};
      u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * ood;
      if (((uint32_t)ob_f2i(u) & 0x80000000u) || (uint32_t)ob_f2i(u) > 0x3f800000u) return false;
      const float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
      v = (dir[0] * qv[0] + dir[1] * qv[1] + dir[2] * qv[2]) * ood;
      if (((uint32_t)ob_f2i(v) & 0x80000000u) || u + v > 1.0f) return false;
      dist = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * ood;
      if ((uint32_t)ob_f2i(dist) & 0x80000000u) return false;
    }
    return (uint32_t)ob_f2i(dist) < (uint32_t)ob_f2i(maxdist);
  }
};

OB_HD int ob_collide_trimesh_ray(const ObPose &o1, const ObPose &o2, const ObMeshDev &m, int flags, ObCg *contact, int *bverr) {
  const int maxc = flags & 0xffff;
  const real *TLPosition = o1.pos, *TLRotation = o1.R;
  const real Length = o2.p[0];
  const int first = (o2.mesh & OB_RAY_FIRSTCONTACT) != 0, closest = (o2.mesh & OB_RAY_CLOSEST_HIT) != 0;
  const real Origin[3] = {o2.pos[0], o2.pos[1], o2.pos[2]}, Direction[3] = {o2.R[2], o2.R[6], o2.R[10]};
  ObRayQuery q;
  q.culling = (o2.mesh & OB_RAY_BACKFACECULL) != 0;
  q.maxdist = (float)Length;
  q.dist = q.u = q.v = 0;
  {
    // InitQuery (:339-361): mDir = Matrix3x3(world) * dir, mOrigin = orig * InvertPRMatrix(world)
    const float wd[3] = {(float)Direction[0], (float)Direction[1], (float)Direction[2]};
    for (int i = 0; i < 3; i++) q.dir[i] = (float)TLRotation[i] * wd[0] + (float)TLRotation[4 + i] * wd[1] + (float)TLRotation[8 + i] * wd[2];
    q.org[0] = (float)Origin[0]; q.org[1] = (float)Origin[1]; q.org[2] = (float)Origin[2];
    ObInvPR inv;
    ob_inv_pr(TLPosition, TLRotation, &inv);
    ob_point_mul(inv, q.org);
    for (int i = 0; i < 3; i++) {   // :426-434
      q.data[i] = (0.5f * q.dir[i]) * q.maxdist;
      q.data2[i] = q.org[i] + q.data[i];
      q.fdir[i] = fabsf(q.data[i]);
    }
  }
  // the stab: every hit in visit order, or only the closest one (strict <, ties keep the first), or the first
  ObBvIter it;
  ob_bv_begin(it);
  int out = 0;
  int ctri = -1; float cdist = 0;
  for (;;) {
    if (!closest && out == maxc) break;
    const int tri = ob_bv_next(m, it, q);
    if (tri < 0) break;
    int use = tri; float T = q.dist;
    if (closest) {
      if (ctri < 0 || q.dist < cdist) { ctri = tri; cdist = q.dist; }
      if (first) break;
      continue;
    }
    real dv[3][3];
    ob_fetch_triangle(m, use, TLPosition, TLRotation, dv);
    real vu[3] = {dv[1][0] - dv[0][0], dv[1][1] - dv[0][1], dv[1][2] - dv[0][2]};
    real vv[3] = {dv[2][0] - dv[0][0], dv[2][1] - dv[0][1], dv[2][2] - dv[0][2]};
    ob_cross(contact[out].normal, vv, vu);   // reversed
    if (ob_safe_normalize3(contact[out].normal)) {
      const real Tr = (real)T;
      for (int e = 0; e < 3; e++) contact[out].pos[e] = Origin[e] + (Direction[e] * Tr);
      contact[out].depth = Tr; contact[out].side1 = use; contact[out].side2 = -1;
      out++;
    }
    if (first) break;
  }
  if (closest && ctri >= 0 && maxc > 0) {
    real dv[3][3];
    ob_fetch_triangle(m, ctri, TLPosition, TLRotation, dv);
    real vu[3] = {dv[1][0] - dv[0][0], dv[1][1] - dv[0][1], dv[1][2] - dv[0][2]};
    real vv[3] = {dv[2][0] - dv[0][0], dv[2][1] - dv[0][1], dv[2][2] - dv[0][2]};
    ob_cross(contact[0].normal, vv, vu);
    if (ob_safe_normalize3(contact[0].normal)) {
      const real Tr = (real)cdist;
      for (int e = 0; e < 3; e++) contact[0].pos[e] = Origin[e] + (Direction[e] * Tr);
      contact[0].depth = Tr; contact[0].side1 = ctri; contact[0].side2 = -1;
      out = 1;
    }
  }
  if (it.overflow && bverr) *bverr = 1;
  return out;
}


// ---- dCollideTrimeshPlane, collision_trimesh_plane.cpp:38-168: every mesh vertex, in triangle order, each
// vertex once (the reference's VertexUseCache bit set == "this corner is the vertex's first use", which is a
// property of the index list and is precomputed at upload: ObMeshDev::vfirst).  o1 = trimesh, o2 = plane.
OB_HD int ob_collide_trimesh_plane(const ObPose &o1, const ObPose &o2, const ObMeshDev &m, int flags, ObCg *contact) {
  const int contact_max = flags & 0xffff;
  int count = 0;
  for (int t = 0; t < m.ntris; ++t) {
    for (int v = 0; v < 3; ++v) {
      const int vi = m.tris[3 * (size_t)t + v];
      if (m.vfirst[vi] != 3 * t + v) continue;
      const float *p = m.verts + 3 * (size_t)vi;
      const real iv[3] = {(real)p[0], (real)p[1], (real)p[2]};
      real vertex[3];
      ob_mul0_331(vertex, o1.R, iv);
      vertex[0] += o1.pos[0]; vertex[1] += o1.pos[1]; vertex[2] += o1.pos[2];
      const real alpha = o2.p[3] - ob_dot(o2.p, vertex);
      if (alpha > 0) {
        ObCg &c = contact[count];
        c.pos[0] = vertex[0]; c.pos[1] = vertex[1]; c.pos[2] = vertex[2];
        c.normal[0] = o2.p[0]; c.normal[1] = o2.p[1]; c.normal[2] = o2.p[2];
        c.depth = alpha; c.side1 = t; c.side2 = -1;
        ++count;
        if (count >= contact_max) return count;
      }
    }
  }
  return count;
}

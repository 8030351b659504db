// ob_trimesh_build.cpp — dGeomTriMeshData* / dCreateTriMesh (include/ode/collision_trimesh.h) and the
// setup-time BVH build that stands in for OPCODE's Model::Build.
//
// What is reproduced, and why it has to be exact: colliders consume the touched-triangle list in
// the tree's depth-first order and stop at max_contacts, so the contact SET depends on the tree.
//   dxTriMeshData::Build           ode/src/collision_trimesh_opcode.cpp:85-166 (rules SPLIT_BEST_AXIS |
//                                  SPLIT_SPLATTER_POINTS | SPLIT_GEOM_CENTER, mNoLeaf, !mQuantized, mLimit 1)
//   AABBTreeNode::Subdivide/Split  OPCODE/OPC_AABBTree.cpp:99-330: splatter-points branch (:176-210, tested before
//                                  best-axis): axis of largest variance of the triangle centroids, float sums in
//                                  primitive-list order; split value = mean of all 3n vertex coordinates
//                                  (OPC_TreeBuilders.cpp:172-190); partition by centroid > split, swapping to the front;
//                                  an invalid split falls back to 50/50 (:272-284)
//   ComputeGlobalBox               OPC_TreeBuilders.cpp:109-131 (float min/max -> centre/extents, Ice/IceAABB.h:285)
//   _BuildNoLeafTree               OPC_OptimizedTree.cpp:150-202: pre-order, positive subtree laid out completely
//                                  before the negative child's id is allocated
// All of it is float arithmetic compiled like the reference (x86-64 SSE, no contraction), at setup time
// on the host exactly where the reference does it; the per-step queries run on the GPU (ob_trimesh.h).
#include <float.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "ob_backend.h"
#include "ob_host.h"
#include "ob_trimesh_host.h"

namespace {
struct Builder {
  const float *verts;
  const int *tris;
  std::vector<int> idx;
  std::vector<ObBvNode> nodes;
  unsigned cur;
  float coord(int tri, int k, int axis) const { return verts[3 * (size_t)tris[3 * (size_t)tri + k] + axis]; }
  float centroid(int tri, int axis) const {   // GetSplittingValue(index, axis), OPC_TreeBuilders.cpp:140-160
    return (coord(tri, 0, axis) + coord(tri, 1, axis) + coord(tri, 2, axis)) * 0.33333333333333333333f;
  }
  void box(const int *prims, int n, ObBvNode &nd) const {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++)
        for (int a = 0; a < 3; a++) {
          const float v = coord(prims[i], k, a);
          if (v < mn[a]) mn[a] = v;     // MIN(x, p.x) = (x < p.x) ? x : p.x with x the accumulator
          if (v > mx[a]) mx[a] = v;
        }
    for (int a = 0; a < 3; a++) { nd.c[a] = (mx[a] + mn[a]) * 0.5f; nd.e[a] = (mx[a] - mn[a]) * 0.5f; }
  }
  int split(int *prims, int n) const {
    // means / variances of the centroids (OPC_AABBTree.cpp:176-204)
    float means[3] = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) means[a] += centroid(prims[i], a);
    { const float s = 1.0f / float(n); for (int a = 0; a < 3; a++) means[a] *= s; }
    float vars[3] = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < n; i++)
      for (int a = 0; a < 3; a++) { const float c = centroid(prims[i], a); vars[a] += (c - means[a]) * (c - means[a]); }
    { const float s = 1.0f / float(n - 1); for (int a = 0; a < 3; a++) vars[a] *= s; }
    int axis = 0;                                   // Point::LargestAxis, Ice/IcePoint.h:341-348
    if (vars[1] > vars[axis]) axis = 1;
    if (vars[2] > vars[axis]) axis = 2;
    // split value (OPC_TreeBuilders.cpp:172-190)
    float sv = 0.0f;
    for (int i = 0; i < n; i++) { sv += coord(prims[i], 0, axis); sv += coord(prims[i], 1, axis); sv += coord(prims[i], 2, axis); }
    sv = sv / float(n * 3);
    int nbpos = 0;                                  // Split, OPC_AABBTree.cpp:99-129
    for (int i = 0; i < n; i++) {
      if (centroid(prims[i], axis) > sv) { const int t = prims[i]; prims[i] = prims[nbpos]; prims[nbpos] = t; nbpos++; }
    }
    if (!nbpos || nbpos == n) nbpos = n >> 1;       // invalid split on a complete tree (:272-284)
    return nbpos;
  }
  // node `id` covers prims[0..n), n >= 2
  void build(unsigned id, int *prims, int n) {
    box(prims, n, nodes[id]);
    const int nbpos = split(prims, n);
    if (nbpos == 1) nodes[id].pos = ((unsigned)prims[0] << 1) | 1u;
    else { const unsigned c = cur++; nodes[id].pos = c << 1; build(c, prims, nbpos); }
    if (n - nbpos == 1) nodes[id].neg = ((unsigned)prims[nbpos] << 1) | 1u;
    else { const unsigned c = cur++; nodes[id].neg = c << 1; build(c, prims + nbpos, n - nbpos); }
  }
};
}  // namespace

static int build_common(dxTriMeshData *d, int nverts, int ntris) {
  if (ntris < 2) { ob_set_last_error("dGeomTriMeshDataBuild: meshes with fewer than 2 triangles are not supported"); return -1; }
  d->nverts = nverts; d->ntris = ntris;
  Builder b;
  b.verts = d->verts.data(); b.tris = d->tris.data();
  b.idx.resize(ntris);
  for (int i = 0; i < ntris; i++) b.idx[i] = i;
  b.nodes.resize(ntris - 1);
  b.cur = 1;
  b.build(0, b.idx.data(), ntris);
  d->nodes.swap(b.nodes);
  d->useflags.clear();
  for (size_t i = 0; i < d->dev.size(); i++) obk_mesh_free(&d->dev[i].m);
  d->dev.clear();
  return 0;
}

// lazily upload to the execution side; one copy per device
const ObMeshDev *ob_trimesh_device(dxTriMeshData *d, int device) {
  for (size_t i = 0; i < d->dev.size(); i++) if (d->dev[i].device == device) return &d->dev[i].m;
  dxTriMeshData::DevCopy c;
  c.device = device;
  memset(&c.m, 0, sizeof c.m);
  for (int k = 0; k < 3; k++) { c.m.aabbc[k] = d->aabbc[k]; c.m.aabbe[k] = d->aabbe[k]; }
  if (obk_mesh_upload(d->verts.data(), d->nverts, d->tris.data(), d->ntris, d->nodes.data(), d->useflags.empty() ? 0 : d->useflags.data(), device, &c.m)) return 0;
  d->dev.push_back(c);
  return &d->dev.back().m;
}

// model-space AABB exactly as dxTriMeshData::Build accumulates it (collision_trimesh_opcode.cpp:123-157)
template <class T> static void mesh_aabb(dxTriMeshData *d, const void *Vertices, int stride, int count) {
  dReal mx[3] = {-dInfinity, -dInfinity, -dInfinity}, mn[3] = {dInfinity, dInfinity, dInfinity};
  const char *p = (const char *)Vertices;
  for (int i = 0; i < count; i++, p += stride) {
    const T *v = (const T *)p;
    for (int k = 0; k < 3; k++) { if (v[k] > mx[k]) mx[k] = (dReal)v[k]; if (v[k] < mn[k]) mn[k] = (dReal)v[k]; }
  }
  for (int k = 0; k < 3; k++) { d->aabbc[k] = (mn[k] + mx[k]) * (dReal)0.5; d->aabbe[k] = mx[k] - d->aabbc[k]; }
}

static void copy_indices(dxTriMeshData *d, const void *Indices, int IndexCount, int TriStride) {
  const int nt = IndexCount / 3;
  d->tris.resize((size_t)nt * 3);
  const char *p = (const char *)Indices;
  for (int t = 0; t < nt; t++, p += TriStride) { const dTriIndex *ix = (const dTriIndex *)p; for (int k = 0; k < 3; k++) d->tris[3 * (size_t)t + k] = (int)ix[k]; }
}

extern "C" {

dTriMeshDataID dGeomTriMeshDataCreate(void) { dxTriMeshData *d = new dxTriMeshData; d->nverts = d->ntris = 0; for (int k = 0; k < 3; k++) d->aabbc[k] = d->aabbe[k] = 0; return d; }
void dGeomTriMeshDataDestroy(dTriMeshDataID d) {
  if (!d) return;
  for (size_t i = 0; i < d->dev.size(); i++) obk_mesh_free(&d->dev[i].m);
  delete d;
}

void dGeomTriMeshDataBuildSingle1(dTriMeshDataID d, const void *Vertices, int VertexStride, int VertexCount, const void *Indices,
                                  int IndexCount, int TriStride, const void *Normals) {
  (void)Normals;
  OB_AASSERT(d && Vertices && Indices);
  d->verts.resize((size_t)VertexCount * 3);
  const char *p = (const char *)Vertices;
  for (int i = 0; i < VertexCount; i++, p += VertexStride) { const float *v = (const float *)p; for (int k = 0; k < 3; k++) d->verts[3 * (size_t)i + k] = v[k]; }
  copy_indices(d, Indices, IndexCount, TriStride);
  mesh_aabb<float>(d, Vertices, VertexStride, VertexCount);
  build_common(d, VertexCount, IndexCount / 3);
}
void dGeomTriMeshDataBuildSingle(dTriMeshDataID d, const void *Vertices, int VertexStride, int VertexCount, const void *Indices,
                                 int IndexCount, int TriStride) {
  dGeomTriMeshDataBuildSingle1(d, Vertices, VertexStride, VertexCount, Indices, IndexCount, TriStride, 0);
}
void dGeomTriMeshDataBuildDouble1(dTriMeshDataID d, const void *Vertices, int VertexStride, int VertexCount, const void *Indices,
                                  int IndexCount, int TriStride, const void *Normals) {
  (void)Normals;
  OB_AASSERT(d && Vertices && Indices);
  // OPCODE reads double vertices through a float conversion area (OPC_MeshInterface.cpp FetchTriangleFromDoubles)
  d->verts.resize((size_t)VertexCount * 3);
  const char *p = (const char *)Vertices;
  for (int i = 0; i < VertexCount; i++, p += VertexStride) { const double *v = (const double *)p; for (int k = 0; k < 3; k++) d->verts[3 * (size_t)i + k] = (float)v[k]; }
  copy_indices(d, Indices, IndexCount, TriStride);
  mesh_aabb<double>(d, Vertices, VertexStride, VertexCount);
  build_common(d, VertexCount, IndexCount / 3);
}
void dGeomTriMeshDataBuildDouble(dTriMeshDataID d, const void *Vertices, int VertexStride, int VertexCount, const void *Indices,
                                 int IndexCount, int TriStride) {
  dGeomTriMeshDataBuildDouble1(d, Vertices, VertexStride, VertexCount, Indices, IndexCount, TriStride, 0);
}
void dGeomTriMeshDataBuildSimple(dTriMeshDataID d, const dReal *Vertices, int VertexCount, const dTriIndex *Indices, int IndexCount) {
  // collision_trimesh_opcode.cpp: dReal[4] vertices, 3 indices per triangle
#if defined(dSINGLE)
  dGeomTriMeshDataBuildSingle(d, Vertices, 4 * sizeof(dReal), VertexCount, Indices, IndexCount, 3 * sizeof(dTriIndex));
#else
  dGeomTriMeshDataBuildDouble(d, Vertices, 4 * sizeof(dReal), VertexCount, Indices, IndexCount, 3 * sizeof(dTriIndex));
#endif
}
void dGeomTriMeshDataBuildSimple1(dTriMeshDataID d, const dReal *Vertices, int VertexCount, const dTriIndex *Indices, int IndexCount, const int *Normals) {
  // collision_trimesh_opcode.cpp:484-498: as BuildSimple, with per-face normals handed to the *1 builder
#if defined(dSINGLE)
  dGeomTriMeshDataBuildSingle1(d, Vertices, 4 * sizeof(dReal), VertexCount, Indices, IndexCount, 3 * sizeof(dTriIndex), Normals);
#else
  dGeomTriMeshDataBuildDouble1(d, Vertices, 4 * sizeof(dReal), VertexCount, Indices, IndexCount, 3 * sizeof(dTriIndex), Normals);
#endif
}
// dxTriMeshData::Preprocess (collision_trimesh_opcode.cpp:256-363): per triangle, which edges and vertices the
// capsule collider should test.  Edges shared by two triangles are paired after sorting (the reference's qsort
// on (VertIdx1, VertIdx2); glibc's is a stable merge sort, so equal records keep their order); convex and
// boundary edges mark their edge + vertices, and finally every vertex of a concave edge is cleared wherever it
// is used (done here with a vertex set instead of the reference's all-pairs loop: same result).
void dGeomTriMeshDataPreprocess(dTriMeshDataID d) {
  if (!d || !d->useflags.empty() || d->ntris <= 0) return;
  struct Edge { int v1, v2, tri; unsigned char ef, f1, f2; bool concave; };
  const int nt = d->ntris, ne = 3 * nt;
  std::vector<Edge> rec(ne);
  static const unsigned char EF[3] = {1, 2, 4}, VF[3] = {8, 16, 32};
  for (int t = 0; t < nt; t++)
    for (int e = 0; e < 3; e++) {
      Edge &r = rec[3 * t + e];
      r.ef = EF[e]; r.f1 = VF[e]; r.f2 = VF[(e + 1) % 3];
      r.v1 = d->tris[3 * (size_t)t + e]; r.v2 = d->tris[3 * (size_t)t + (e + 1) % 3];
      if (r.v1 > r.v2) { int tv = r.v1; r.v1 = r.v2; r.v2 = tv; unsigned char tf = r.f1; r.f1 = r.f2; r.f2 = tf; }
      r.tri = t; r.concave = false;
    }
  std::stable_sort(rec.begin(), rec.end(), [](const Edge &a, const Edge &b) { return a.v1 == b.v1 ? a.v2 < b.v2 : a.v1 < b.v1; });
  d->useflags.assign(nt, 0);
  auto vert = [&](int tri, int k) { return d->verts.data() + 3 * (size_t)d->tris[3 * (size_t)tri + k]; };
  auto opposite = [&](const Edge &r) {   // GetOppositeVert
    if ((r.f1 == 8 && r.f2 == 16) || (r.f1 == 16 && r.f2 == 8)) return vert(r.tri, 2);
    if ((r.f1 == 16 && r.f2 == 32) || (r.f1 == 32 && r.f2 == 16)) return vert(r.tri, 0);
    return vert(r.tri, 1);
  };
  for (int i = 0; i < ne; i++) {
    Edge &r1 = rec[i];
    if (i < ne - 1 && r1.v1 == rec[i + 1].v1 && r1.v2 == rec[i + 1].v2) {
      const Edge &r2 = rec[i + 1];
      const float *p0 = vert(r1.tri, 0), *p1 = vert(r1.tri, 1), *p2 = vert(r1.tri, 2);
      const float a[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, b[3] = {p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]};
      float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
      float M = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
      if (M) { M = 1.0f / sqrtf(M); n[0] *= M; n[1] *= M; n[2] *= M; }
      const float *o1 = opposite(r1), *o2 = opposite(r2);
      float dv[3] = {o2[0] - o1[0], o2[1] - o1[1], o2[2] - o1[2]};
      float M2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
      if (M2) { M2 = 1.0f / sqrtf(M2); dv[0] *= M2; dv[1] *= M2; dv[2] *= M2; }
      const float dot = n[0] * dv[0] + n[1] * dv[1] + n[2] * dv[2];
      if (dot >= -0.000001f) r1.concave = true;
      else d->useflags[r1.tri] |= r1.f1 | r1.f2 | r1.ef;
      i++;
    } else {
      d->useflags[r1.tri] |= r1.f1 | r1.f2 | r1.ef;
    }
  }
  std::vector<char> cv(d->nverts > 0 ? d->nverts : 1, 0);
  bool any = false;
  for (int i = 0; i < ne; i++) if (rec[i].concave) { cv[rec[i].v1] = 1; cv[rec[i].v2] = 1; any = true; }
  if (any)
    for (int j = 0; j < ne; j++) {
      if (cv[rec[j].v1]) d->useflags[rec[j].tri] &= (unsigned char)~rec[j].f1;
      if (cv[rec[j].v2]) d->useflags[rec[j].tri] &= (unsigned char)~rec[j].f2;
    }
  for (size_t i = 0; i < d->dev.size(); i++) obk_mesh_free(&d->dev[i].m);   // device copies are rebuilt with the flags
  d->dev.clear();
}
void dGeomTriMeshDataUpdate(dTriMeshDataID) {}

dGeomID dCreateTriMesh(dSpaceID space, dTriMeshDataID Data, dTriCallback *Callback, dTriArrayCallback *ArrayCallback, dTriRayCallback *RayCallback) {
  if (Callback || ArrayCallback || RayCallback) ob_message(0, "dCreateTriMesh: per-triangle callbacks are not supported by the GPU colliders and are ignored");
  dxGeom *g = ob_geom_create(space, 1, dTriMeshClass);
  g->tmdata = Data;
  return g;
}
void dGeomTriMeshSetData(dGeomID g, dTriMeshDataID Data) { OB_UASSERT(g && g->type == dTriMeshClass, "argument not a trimesh"); g->tmdata = Data; ob_geom_moved(g); }
dTriMeshDataID dGeomTriMeshGetData(dGeomID g) { OB_UASSERT(g && g->type == dTriMeshClass, "argument not a trimesh"); return g->tmdata; }
int dGeomTriMeshGetTriangleCount(dGeomID g) { OB_UASSERT(g && g->type == dTriMeshClass, "argument not a trimesh"); return g->tmdata ? g->tmdata->ntris : 0; }

}  // extern "C"

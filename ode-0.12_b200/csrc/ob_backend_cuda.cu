// ob_backend_cuda.cu — sm_100a execution backend: device memory + the kernels.
//
// One CTA owns one world for a whole phase and keeps that world's working set
// in shared memory; independent worlds are spread over the grid (grid-stride
// over world ids, so a grid of k*148 CTAs covers any batch size).  Two kernels
// per step:
//   k_collide : pose+AABB per geom -> order-exact dxHashSpace pair list
//               (ob_broad.h) -> thread-per-pair primitive narrowphase
//               (ob_collide.h) -> contact joints in creation order.
//   k_step    : joint graph -> island DFS (order-exact) -> row assembly ->
//               SOR with the reference's seeded row shuffle -> integration ->
//               space-list reordering (what dGeomMoved does).
// Numerics: compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
// so every per-element function in ob_*.h evaluates exactly like the
// reference's scalar SSE build.
//
// SOR parallelisation that keeps results bit-identical: rows are processed in
// the reference's order[] sequence, 32 consecutive positions ("window") per
// warp pass; inside a window a row may run as soon as every earlier row of the
// window that shares a body with it has run (rows that touch disjoint bodies
// commute exactly: they read and write disjoint fc[] entries, and a friction
// row's lambda[findex] lives on the same two bodies as the row itself).  The
// per-window round numbers are computed once per shuffle epoch with shuffles.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ob_backend.h"
#include "ob_broad.h"
#include "ob_rows.h"
#include "ob_solver.h"
#include <unistd.h>
#include "ob_large.h"
#include "ob_step_kernel.cuh"

#define OB_THREADS 128
static long long g_launches = 0;

// ------------------------------------------------------------------------------------
// shared-memory layouts (one function for host sizing and device carving)
struct CollideSmem {
  size_t pose, aabb, cb, gid, body, cat, col, en, hr, br, walk_of, sapkey, sapinit, sappos, sapwalk, key, o12, sorted, misc, total;
};
__host__ __device__ inline size_t ob_al16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline CollideSmem collide_smem(int NG, int NP) {
  CollideSmem s; size_t o = 0;
  s.pose = o; o = ob_al16(o + sizeof(ObPose) * NG);
  s.aabb = o; o = ob_al16(o + sizeof(real) * 6 * NG);
  s.cb = o; o = ob_al16(o + sizeof(ObCellBox) * NG);
  s.gid = o; o = ob_al16(o + sizeof(int) * NG);
  s.body = o; o = ob_al16(o + sizeof(int) * NG);
  s.cat = o; o = ob_al16(o + sizeof(uint32_t) * NG);
  s.col = o; o = ob_al16(o + sizeof(uint32_t) * NG);
  s.en = o; o = ob_al16(o + sizeof(int) * NG);
  s.hr = o; o = ob_al16(o + sizeof(int) * NG);
  s.br = o; o = ob_al16(o + sizeof(int) * NG);
  s.walk_of = o; o = ob_al16(o + sizeof(int) * NG);
  s.sapkey = o; o = ob_al16(o + sizeof(float) * (NG + 1));
  s.sapinit = o; o = ob_al16(o + sizeof(int) * (NG + 1));
  s.sappos = o; o = ob_al16(o + sizeof(int) * (NG + 1));
  s.sapwalk = o; o = ob_al16(o + sizeof(int) * (NG + 1));
  s.key = o; o = ob_al16(o + sizeof(ObPairKey) * NP);
  s.o12 = o; o = ob_al16(o + sizeof(int2) * NP);
  s.sorted = o; o = ob_al16(o + sizeof(int2) * NP);
  s.misc = o; o = ob_al16(o + sizeof(int) * 64);
  s.total = o;
  return s;
}

// exclusive scan of one int per thread across the CTA; *total = sum.  s_w: >= 33 ints of smem
__device__ inline int block_excl_scan(int v, int *s_w, int *total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int x = v;
  for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) s_w[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int t = lane < nw ? s_w[lane] : 0;
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
    s_w[lane] = t;   // inclusive warp totals
  }
  __syncthreads();
  int base = wid ? s_w[wid - 1] : 0;
  *total = s_w[nw - 1];
  __syncthreads();
  return base + x - v;
}

__device__ inline void geom_pose_dev(const ObGeom &g, const ObBodyDyn *bd, ObPose *o) {
  o->type = g.type;
  o->mesh = g.mesh;
  for (int k = 0; k < 4; k++) o->p[k] = g.p[k];
  if (g.body >= 0) {
    const ObBodyDyn &b = bd[g.body];
    if (g.flags & OB_GEOM_HAS_OFFSET) {
      ob_mul0_331(o->pos, b.R, g.pos);
      o->pos[0] += b.pos[0]; o->pos[1] += b.pos[1]; o->pos[2] += b.pos[2];
      ob_mul0_333(o->R, b.R, g.R);
      o->R[3] = o->R[7] = o->R[11] = 0;
    } else {
      for (int k = 0; k < 3; k++) o->pos[k] = b.pos[k];
      for (int k = 0; k < 12; k++) o->R[k] = b.R[k];
    }
  } else {
    for (int k = 0; k < 3; k++) o->pos[k] = g.pos[k];
    for (int k = 0; k < 12; k++) o->R[k] = g.R[k];
  }
}

// ------------------------------------------------------------------------------------
// MESH: the batch has trimesh geoms (narrowphase with the BVH colliders, up to OB_MAXC_LOCAL contacts per
// pair); otherwise the primitive-only narrowphase with 8 contact slots per pair (box-box emits at most 8)
template <bool MESH, bool XF>
__global__ void __launch_bounds__(OB_THREADS) k_collide(ObBatchDev d) {
  constexpr int CGCAP = MESH ? OB_MAXC_LOCAL : 8;
  extern __shared__ __align__(16) unsigned char smem[];
  const CollideSmem L = collide_smem(d.NG, d.NP);
  ObPose *s_pose = (ObPose *)(smem + L.pose);
  real *s_aabb = (real *)(smem + L.aabb);
  ObCellBox *s_cb = (ObCellBox *)(smem + L.cb);
  int *s_gid = (int *)(smem + L.gid);
  int *s_body = (int *)(smem + L.body);
  uint32_t *s_cat = (uint32_t *)(smem + L.cat);
  uint32_t *s_col = (uint32_t *)(smem + L.col);
  int *s_en = (int *)(smem + L.en);
  int *s_hr = (int *)(smem + L.hr);
  int *s_br = (int *)(smem + L.br);
  int *s_walk_of = (int *)(smem + L.walk_of);
  float *s_sapkey = (float *)(smem + L.sapkey);
  int *s_sapinit = (int *)(smem + L.sapinit);
  int *s_sappos = (int *)(smem + L.sappos);
  int *s_sapwalk = (int *)(smem + L.sapwalk);
  ObPairKey *s_key = (ObPairKey *)(smem + L.key);
  int2 *s_o12 = (int2 *)(smem + L.o12);
  int2 *s_sorted = (int2 *)(smem + L.sorted);
  int *s_misc = (int *)(smem + L.misc);   // [0]=npairs raw, [1]=nh, [2]=nbig, [3]=contact base, [8..40]=scan scratch
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int w = d.wbeg + blockIdx.x; w < d.wend; w += gridDim.x) {
    ObWorld &W = d.world[w];
    const int ng = W.ng;
    const ObGeom *geoms = d.geom + (size_t)w * d.NG;
    const ObBodyDyn *bd = d.bdyn + (size_t)w * d.NB;
    const int *glist = d.glist + (size_t)w * d.NG;
    if (tid < 8) s_misc[tid] = 0;
    const int stype = W.space_type;
    // SAP: cleanGeoms appends the DirtyList to the GeomList (collision_sapspace.cpp:394-423), so the walk
    // order is glist rotated by sap_ndirty; the cleaned order is written back below
    const int rot = stype == OB_SPACE_SAP ? W.sap_ndirty : 0;
    int ax0 = 0, ax1 = 2, ax2 = 4;
    if (stype == OB_SPACE_SAP) ob_sap_axes(W.sap_axes, &ax0, &ax1, &ax2);
    // (1) pose, AABB, cell box per geom in walk order
    for (int i = tid; i < ng; i += nt) {
      int gi = glist[i + rot < ng ? i + rot : i + rot - ng];
      const ObGeom g = geoms[gi];
      s_gid[i] = gi; s_body[i] = g.body; s_cat[i] = g.cat; s_col[i] = g.col;
      s_walk_of[gi] = i;
      s_en[i] = (g.flags & OB_GEOM_ENABLED) && !(g.flags & OB_GEOM_ZERO_SIZED);
      ObPose p;
      geom_pose_dev(g, bd, &p);
      s_pose[i] = p;
      real ab[6];
      ob_aabb(p, ab, d.meshes);
      for (int k = 0; k < 6; k++) s_aabb[6 * i + k] = ab[k];
      ObCellBox cb;
      cb.level = 0;
      for (int k = 0; k < 6; k++) cb.db[k] = 0;
      if (stype == OB_SPACE_HASH) ob_hash_cellbox(ab, W.hash_minlevel, W.hash_maxlevel, &cb);
      else if (stype == OB_SPACE_SAP && ab[ax0 + 1] == OB_INF) cb.level = OB_LEVEL_BIG;   // TmpInfGeomList (:446-449)
      s_cb[i] = cb;
    }
    __syncthreads();
    if (stype == OB_SPACE_SAP) {
      int *gl = d.glist + (size_t)w * d.NG;
      for (int i = tid; i < ng; i += nt) gl[i] = s_gid[i];
      if (tid == 0) W.sap_ndirty = 0;
    }
    // (2) ranks among hashed / big geoms in walk order (SAP: finite / infinite on axis 0)
    for (int i = tid; i < ng; i += nt) {
      int h = 0, b = 0;
      for (int j = 0; j < i; j++)
        if (s_en[j]) { if (s_cb[j].level == OB_LEVEL_BIG) b++; else h++; }
      s_hr[i] = h; s_br[i] = b;
      if (s_en[i] && s_cb[i].level != OB_LEVEL_BIG) { s_sapwalk[h] = i; s_sapkey[h] = (float)s_aabb[6 * i + ax0]; }
      if (i == ng - 1) {
        if (s_en[i]) { if (s_cb[i].level == OB_LEVEL_BIG) b++; else h++; }
        s_misc[1] = h; s_misc[2] = b;
      }
    }
    __syncthreads();
    const int nh = s_misc[1], nbig = s_misc[2];
    // (2b) SAP: sorted position of every finite geom = RadixSort's output order (ob_broad.h)
    if (stype == OB_SPACE_SAP && nh > 0) {
      int *st = d.sapstate + (size_t)w * (d.NG + 3);
      const int nbk = nh + 1;                       // + FLT_MAX sentinel, element index nh
      const bool valid = st[0] != 0 && st[1] == nbk;
      if (tid == 0) s_sapkey[nh] = 3.402823466e+38f;
      for (int p = tid; p < nbk; p += nt) { if (valid) s_sapinit[st[2 + p]] = p; else s_sapinit[p] = p; }
      __syncthreads();
      for (int p = 1 + tid; p < nbk; p += nt) {
        const int e = valid ? st[2 + p] : p, e0 = valid ? st[1 + p] : p - 1;
        if (s_sapkey[e] < s_sapkey[e0]) s_misc[4] = 1;   // not already sorted
      }
      __syncthreads();
      const bool unsorted = s_misc[4] != 0;
      for (int t = tid; t < nbk; t += nt) {
        int pos = s_sapinit[t];
        if (unsorted) {
          const uint32_t ot = ob_sap_keyorder(s_sapkey[t]);
          pos = 0;
          for (int u = 0; u < nbk; u++)
            if (u != t && ob_sap_precedes(ob_sap_keyorder(s_sapkey[u]), ot, s_sapinit[u], s_sapinit[t])) pos++;
        }
        s_sappos[t] = pos;
      }
      __syncthreads();
      if (unsorted) for (int t = tid; t < nbk; t += nt) st[2 + s_sappos[t]] = t;
      if (tid == 0) { st[1] = nbk; if (unsorted) st[0] = 1; else if (!valid) st[0] = 0; }
    }
    // (3) candidate pairs: the space's filter + the sequence key of the pair's callback
    for (int idx = tid; idx < ng * ng; idx += nt) {
      int a = idx / ng, b = idx - a * ng;
      if (a >= b || !s_en[a] || !s_en[b]) continue;
      ObPairKey key;
      int first_is_a;
      if (stype == OB_SPACE_HASH) {
        if (!ob_aabb_pair_filter(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b], s_aabb + 6 * a, s_aabb + 6 * b))
          continue;
        if (!ob_hash_pair_key(a, b, s_cb[a], s_cb[b], s_hr[a], s_hr[b], s_br[a], s_br[b], nh, nbig, &key, &first_is_a)) continue;
      } else if (stype == OB_SPACE_SAP) {
        if (!ob_pair_filter_noaabb(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b])) continue;
        const bool ia = s_cb[a].level == OB_LEVEL_BIG, ib = s_cb[b].level == OB_LEVEL_BIG;
        for (int k = 0; k < 7; k++) key.k[k] = 0;
        if (!ia && !ib) {
          const int pa = s_sappos[s_hr[a]], pb = s_sappos[s_hr[b]];
          first_is_a = pa < pb;
          const int K = first_is_a ? a : b, J = first_is_a ? b : a;
          if (!ob_sap_sweep_test(s_sapkey[s_hr[J]], s_aabb + 6 * K, s_aabb + 6 * J, ax0, ax1, ax2)) continue;
          key.k[1] = first_is_a ? pa : pb; key.k[2] = first_is_a ? pb : pa;
        } else if (ia && ib) { key.k[0] = 1; key.k[1] = s_br[a]; key.k[3] = s_br[b]; first_is_a = 1; }
        else { key.k[0] = 1; key.k[1] = ia ? s_br[a] : s_br[b]; key.k[2] = 1; key.k[3] = ia ? s_hr[b] : s_hr[a]; first_is_a = ia; }
      } else {   // dxSimpleSpace::collide (collision_space.cpp:247-268): nested walk of the list
        if (!ob_aabb_pair_filter(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b], s_aabb + 6 * a, s_aabb + 6 * b))
          continue;
        for (int k = 0; k < 7; k++) key.k[k] = 0;
        key.k[1] = a; key.k[2] = b; first_is_a = 1;
      }
      int slot = atomicAdd(&s_misc[0], 1);
      if (slot < d.NP) {
        s_key[slot] = key;
        s_o12[slot] = first_is_a ? make_int2(s_gid[a], s_gid[b]) : make_int2(s_gid[b], s_gid[a]);
      }
    }
    __syncthreads();
    int np = s_misc[0];
    if (np > d.NP) { np = d.NP; if (tid == 0) atomicOr(&W.status, OB_ERR_PAIR_OVERFLOW); }
    // (4) order: rank of every pair = number of pairs with a smaller key (keys are unique)
    int *gpairs = d.pairs + (size_t)w * d.NP * 2;
    for (int p = tid; p < np; p += nt) {
      const ObPairKey kp = s_key[p];
      int rank = 0;
      for (int q = 0; q < np; q++) rank += ob_key_less(s_key[q], kp) ? 1 : 0;
      s_sorted[rank] = s_o12[p];
      gpairs[2 * rank] = s_o12[p].x; gpairs[2 * rank + 1] = s_o12[p].y;
    }
    __syncthreads();
    // (5) narrowphase per pair in callback order, ordered compaction into contact joints
    const ObPolicy pol = d.policy[0];
    ObContact *cout = d.contacts + (size_t)w * d.NC;
    const int maxc = pol.max_contacts > CGCAP ? CGCAP : pol.max_contacts;
    for (int base = 0; base < np; base += nt) {
      int p = base + tid;
      ObCg cg[CGCAP];
      int n = 0, o1 = 0, o2 = 0;
      if (p < np) {
        o1 = s_sorted[p].x; o2 = s_sorted[p].y;
        bool connected = false;
        if (pol.skip_if_connected && d.NJ) {   // dAreConnectedExcluding(b1, b2, dJointTypeContact), ode.cpp:1529-1537
          const int b1 = geoms[o1].body, b2 = geoms[o2].body;
          if (b1 >= 0 && b2 >= 0) {
            const unsigned short *ps = d.padjstart + (size_t)w * (d.NB + 1), *pa = d.padj + (size_t)w * 2 * d.NJ;
            const ObJoint *pj = d.joint + (size_t)w * d.NJ;
            for (int k = ps[b1]; k < ps[b1 + 1]; k++) {
              const ObJoint &jj = pj[pa[k]];
              const int other = jj.b1 == b1 ? jj.b2 : jj.b1;
              if (other == b2) connected = true;
            }
          }
        }
        int swapped;
        int bverr = 0;
        if (!connected) n = ob_collide_pair_sel_t<MESH, CGCAP, XF>(&s_pose[s_walk_of[o1]], &s_pose[s_walk_of[o2]], maxc, cg, &swapped, d.meshes, &bverr);
        if (bverr) atomicOr(&W.status, OB_ERR_BVH_STACK);
      }
      int total;
      int off = block_excl_scan(n, s_misc + 8, &total);
      int cbase = s_misc[3];
      for (int k = 0; k < n; k++) {
        int j = cbase + off + k;
        if (j < d.NC) {
          ObContact c;
          for (int e = 0; e < 3; e++) { c.pos[e] = cg[k].pos[e]; c.normal[e] = cg[k].normal[e]; }
          c.depth = cg[k].depth; c.g1 = o1; c.g2 = o2; c.side1 = cg[k].side1; c.side2 = cg[k].side2; c.policy = 0;
          cout[j] = c;
        }
      }
      __syncthreads();
      if (tid == 0) s_misc[3] = cbase + total;
      __syncthreads();
    }
    if (tid == 0) {
      int nc = s_misc[3];
      if (nc > d.NC) { nc = d.NC; atomicOr(&W.status, OB_ERR_CONTACT_OVERFLOW); }
      d.ncontacts[w] = nc;
      d.npairs[w] = np;
      atomicAdd(&d.counters->pairs, (unsigned long long)np);
    }
    __syncthreads();
  }
}

#define OB_TILE_WPC 8   // worlds per CTA of k_collide_tile
struct CollideTileSmem { size_t np, cb, first, scan, cls, perm, stoff, cnt, stage, total; };
__host__ __device__ inline CollideTileSmem collide_tile_smem(int NG, int NP, int WPC, int stage_cap) {
  CollideTileSmem s; size_t o = 0;
  s.np = o; o = ob_al16(o + sizeof(int) * (WPC + 1));
  s.cb = o; o = ob_al16(o + sizeof(int) * WPC);
  s.first = o; o = ob_al16(o + sizeof(int) * WPC);
  s.scan = o; o = ob_al16(o + sizeof(int) * 40);
  s.cls = o; o = ob_al16(o + sizeof(int) * 8);
  s.perm = o; o = ob_al16(o + sizeof(unsigned short) * WPC * NP);
  s.stoff = o; o = ob_al16(o + sizeof(unsigned short) * WPC * NP);
  s.cnt = o; o = ob_al16(o + (size_t)WPC * NP);
  s.stage = o; o = ob_al16(o + sizeof(ObCg) * stage_cap);
  s.total = o;
  return s;
}
// k_collide_tile: the same products as k_collide for batches of SMALL worlds (a handful of geoms, e.g. the buggies of
// BASELINE.json configs[2]).  With one warp per world the narrowphase runs on 5-9 lanes of 32 (ncu, r01z: 89 % of the
// kernel's warp instructions execute with <= 4 active threads).  Here a CTA takes WPC worlds: every warp stages its
// own world (poses, AABBs, ordered pair list: phases 1-4, warp-synchronous), then the pairs of all WPC worlds are
// pooled, grouped by collider class so that a warp runs ONE collider on full lanes, and their contacts go through a
// shared-memory staging area into the per-world contact arrays in callback order (phase 5).
template <bool MESH, bool XF, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_collide_tile(ObBatchDev d, int stage_cap) {
  constexpr int CGCAP = MESH ? OB_MAXC_LOCAL : 8;
  extern __shared__ __align__(16) unsigned char smem[];
  const CollideSmem L = collide_smem(d.NG, d.NP);
  const int warp = threadIdx.x >> 5;
  unsigned char *sm = smem + (size_t)warp * L.total;
  ObPose *s_pose = (ObPose *)(sm + L.pose);
  real *s_aabb = (real *)(sm + L.aabb);
  ObCellBox *s_cb = (ObCellBox *)(sm + L.cb);
  int *s_gid = (int *)(sm + L.gid);
  int *s_body = (int *)(sm + L.body);
  uint32_t *s_cat = (uint32_t *)(sm + L.cat);
  uint32_t *s_col = (uint32_t *)(sm + L.col);
  int *s_en = (int *)(sm + L.en);
  int *s_hr = (int *)(sm + L.hr);
  int *s_br = (int *)(sm + L.br);
  int *s_walk_of = (int *)(sm + L.walk_of);
  float *s_sapkey = (float *)(sm + L.sapkey);
  int *s_sapinit = (int *)(sm + L.sapinit);
  int *s_sappos = (int *)(sm + L.sappos);
  int *s_sapwalk = (int *)(sm + L.sapwalk);
  ObPairKey *s_key = (ObPairKey *)(sm + L.key);
  int2 *s_o12 = (int2 *)(sm + L.o12);
  int2 *s_sorted = (int2 *)(sm + L.sorted);
  int *s_misc = (int *)(sm + L.misc);   // [0]=npairs raw, [1]=nh, [2]=nbig, [3]=contact base, [8..40]=scan scratch
  const int tid = threadIdx.x & 31, nt = 32;
  // CTA-wide area behind the WPC per-world slices
  const CollideTileSmem T = collide_tile_smem(d.NG, d.NP, WPC, stage_cap);
  unsigned char *cm = smem + (size_t)WPC * L.total;
  int *c_np = (int *)(cm + T.np);                    // [WPC+1] prefix of the worlds' pair counts
  int *c_cb = (int *)(cm + T.cb);                    // [WPC] contacts written so far per world
  int *c_first = (int *)(cm + T.first);              // [WPC] scan value at a world's first pair of the chunk
  int *c_scan = (int *)(cm + T.scan);                // [40] block_excl_scan scratch
  int *c_cls = (int *)(cm + T.cls);                  // [8] pairs per collider class -> class starts -> fill cursors; [7] = staged contacts
  unsigned short *c_perm = (unsigned short *)(cm + T.perm);     // [WPC*NP] class-grouped order -> pooled pair
  unsigned short *c_stoff = (unsigned short *)(cm + T.stoff);   // [WPC*NP] staging offset of a pooled pair
  unsigned char *c_n = cm + T.cnt;                              // [WPC*NP] contacts of a pooled pair
  ObCg *c_stage = (ObCg *)(cm + T.stage);                       // [stage_cap]

  for (int wb = d.wbeg + blockIdx.x * WPC; wb < d.wend; wb += gridDim.x * WPC) {
    const int w = wb + warp;
    const bool valid = w < d.wend;
    int np = 0;
    if (valid) {
    ObWorld &W = d.world[w];
    const int ng = W.ng;
    const ObGeom *geoms = d.geom + (size_t)w * d.NG;
    const ObBodyDyn *bd = d.bdyn + (size_t)w * d.NB;
    const int *glist = d.glist + (size_t)w * d.NG;
    if (tid < 8) s_misc[tid] = 0;
    const int stype = W.space_type;
    // SAP: cleanGeoms appends the DirtyList to the GeomList (collision_sapspace.cpp:394-423), so the walk
    // order is glist rotated by sap_ndirty; the cleaned order is written back below
    const int rot = stype == OB_SPACE_SAP ? W.sap_ndirty : 0;
    int ax0 = 0, ax1 = 2, ax2 = 4;
    if (stype == OB_SPACE_SAP) ob_sap_axes(W.sap_axes, &ax0, &ax1, &ax2);
    // (1) pose, AABB, cell box per geom in walk order
    for (int i = tid; i < ng; i += nt) {
      int gi = glist[i + rot < ng ? i + rot : i + rot - ng];
      const ObGeom g = geoms[gi];
      s_gid[i] = gi; s_body[i] = g.body; s_cat[i] = g.cat; s_col[i] = g.col;
      s_walk_of[gi] = i;
      s_en[i] = (g.flags & OB_GEOM_ENABLED) && !(g.flags & OB_GEOM_ZERO_SIZED);
      ObPose p;
      geom_pose_dev(g, bd, &p);
      s_pose[i] = p;
      real ab[6];
      ob_aabb(p, ab, d.meshes);
      for (int k = 0; k < 6; k++) s_aabb[6 * i + k] = ab[k];
      ObCellBox cb;
      cb.level = 0;
      for (int k = 0; k < 6; k++) cb.db[k] = 0;
      if (stype == OB_SPACE_HASH) ob_hash_cellbox(ab, W.hash_minlevel, W.hash_maxlevel, &cb);
      else if (stype == OB_SPACE_SAP && ab[ax0 + 1] == OB_INF) cb.level = OB_LEVEL_BIG;   // TmpInfGeomList (:446-449)
      s_cb[i] = cb;
    }
    __syncwarp();
    if (stype == OB_SPACE_SAP) {
      int *gl = d.glist + (size_t)w * d.NG;
      for (int i = tid; i < ng; i += nt) gl[i] = s_gid[i];
      if (tid == 0) W.sap_ndirty = 0;
    }
    // (2) ranks among hashed / big geoms in walk order (SAP: finite / infinite on axis 0)
    for (int i = tid; i < ng; i += nt) {
      int h = 0, b = 0;
      for (int j = 0; j < i; j++)
        if (s_en[j]) { if (s_cb[j].level == OB_LEVEL_BIG) b++; else h++; }
      s_hr[i] = h; s_br[i] = b;
      if (s_en[i] && s_cb[i].level != OB_LEVEL_BIG) { s_sapwalk[h] = i; s_sapkey[h] = (float)s_aabb[6 * i + ax0]; }
      if (i == ng - 1) {
        if (s_en[i]) { if (s_cb[i].level == OB_LEVEL_BIG) b++; else h++; }
        s_misc[1] = h; s_misc[2] = b;
      }
    }
    __syncwarp();
    const int nh = s_misc[1], nbig = s_misc[2];
    // (2b) SAP: sorted position of every finite geom = RadixSort's output order (ob_broad.h)
    if (stype == OB_SPACE_SAP && nh > 0) {
      int *st = d.sapstate + (size_t)w * (d.NG + 3);
      const int nbk = nh + 1;                       // + FLT_MAX sentinel, element index nh
      const bool valid = st[0] != 0 && st[1] == nbk;
      if (tid == 0) s_sapkey[nh] = 3.402823466e+38f;
      for (int p = tid; p < nbk; p += nt) { if (valid) s_sapinit[st[2 + p]] = p; else s_sapinit[p] = p; }
      __syncwarp();
      for (int p = 1 + tid; p < nbk; p += nt) {
        const int e = valid ? st[2 + p] : p, e0 = valid ? st[1 + p] : p - 1;
        if (s_sapkey[e] < s_sapkey[e0]) s_misc[4] = 1;   // not already sorted
      }
      __syncwarp();
      const bool unsorted = s_misc[4] != 0;
      for (int t = tid; t < nbk; t += nt) {
        int pos = s_sapinit[t];
        if (unsorted) {
          const uint32_t ot = ob_sap_keyorder(s_sapkey[t]);
          pos = 0;
          for (int u = 0; u < nbk; u++)
            if (u != t && ob_sap_precedes(ob_sap_keyorder(s_sapkey[u]), ot, s_sapinit[u], s_sapinit[t])) pos++;
        }
        s_sappos[t] = pos;
      }
      __syncwarp();
      if (unsorted) for (int t = tid; t < nbk; t += nt) st[2 + s_sappos[t]] = t;
      if (tid == 0) { st[1] = nbk; if (unsorted) st[0] = 1; else if (!valid) st[0] = 0; }
    }
    // (3) candidate pairs: the space's filter + the sequence key of the pair's callback
    for (int idx = tid; idx < ng * ng; idx += nt) {
      int a = idx / ng, b = idx - a * ng;
      if (a >= b || !s_en[a] || !s_en[b]) continue;
      ObPairKey key;
      int first_is_a;
      if (stype == OB_SPACE_HASH) {
        if (!ob_aabb_pair_filter(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b], s_aabb + 6 * a, s_aabb + 6 * b))
          continue;
        if (!ob_hash_pair_key(a, b, s_cb[a], s_cb[b], s_hr[a], s_hr[b], s_br[a], s_br[b], nh, nbig, &key, &first_is_a)) continue;
      } else if (stype == OB_SPACE_SAP) {
        if (!ob_pair_filter_noaabb(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b])) continue;
        const bool ia = s_cb[a].level == OB_LEVEL_BIG, ib = s_cb[b].level == OB_LEVEL_BIG;
        for (int k = 0; k < 7; k++) key.k[k] = 0;
        if (!ia && !ib) {
          const int pa = s_sappos[s_hr[a]], pb = s_sappos[s_hr[b]];
          first_is_a = pa < pb;
          const int K = first_is_a ? a : b, J = first_is_a ? b : a;
          if (!ob_sap_sweep_test(s_sapkey[s_hr[J]], s_aabb + 6 * K, s_aabb + 6 * J, ax0, ax1, ax2)) continue;
          key.k[1] = first_is_a ? pa : pb; key.k[2] = first_is_a ? pb : pa;
        } else if (ia && ib) { key.k[0] = 1; key.k[1] = s_br[a]; key.k[3] = s_br[b]; first_is_a = 1; }
        else { key.k[0] = 1; key.k[1] = ia ? s_br[a] : s_br[b]; key.k[2] = 1; key.k[3] = ia ? s_hr[b] : s_hr[a]; first_is_a = ia; }
      } else {   // dxSimpleSpace::collide (collision_space.cpp:247-268): nested walk of the list
        if (!ob_aabb_pair_filter(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b], s_aabb + 6 * a, s_aabb + 6 * b))
          continue;
        for (int k = 0; k < 7; k++) key.k[k] = 0;
        key.k[1] = a; key.k[2] = b; first_is_a = 1;
      }
      int slot = atomicAdd(&s_misc[0], 1);
      if (slot < d.NP) {
        s_key[slot] = key;
        s_o12[slot] = first_is_a ? make_int2(s_gid[a], s_gid[b]) : make_int2(s_gid[b], s_gid[a]);
      }
    }
    __syncwarp();
    np = s_misc[0];
    if (np > d.NP) { np = d.NP; if (tid == 0) atomicOr(&W.status, OB_ERR_PAIR_OVERFLOW); }
    // (4) order: rank of every pair = number of pairs with a smaller key (keys are unique)
    int *gpairs = d.pairs + (size_t)w * d.NP * 2;
    for (int p = tid; p < np; p += nt) {
      const ObPairKey kp = s_key[p];
      int rank = 0;
      for (int q = 0; q < np; q++) rank += ob_key_less(s_key[q], kp) ? 1 : 0;
      s_sorted[rank] = s_o12[p];
      gpairs[2 * rank] = s_o12[p].x; gpairs[2 * rank + 1] = s_o12[p].y;
    }
    __syncwarp();
    }   // valid
    // (5) narrowphase over the pooled pairs of the CTA's worlds
    const ObPolicy pol = d.policy[0];
    const int maxc = pol.max_contacts > CGCAP ? CGCAP : pol.max_contacts;
    const int nthr = 32 * WPC;
    if (tid == 0) c_np[warp + 1] = np;
    if (threadIdx.x < 8) c_cls[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) { c_np[0] = 0; for (int v = 0; v < WPC; v++) { c_np[v + 1] += c_np[v]; c_cb[v] = 0; } }
    __syncthreads();
    const int total = c_np[WPC];
    // (5a) group the pooled pairs by collider class (order inside a class is irrelevant: results are staged)
    for (int f = threadIdx.x; f < total; f += nthr) {
      int v = 0;
      while (f >= c_np[v + 1]) v++;
      const unsigned char *smv = smem + (size_t)v * L.total;
      const int2 o12 = ((const int2 *)(smv + L.sorted))[f - c_np[v]];
      const ObGeom *gv = d.geom + (size_t)(wb + v) * d.NG;
      const int t1 = gv[o12.x].type, t2 = gv[o12.y].type;
      const int lo = t1 < t2 ? t1 : t2, hi = t1 < t2 ? t2 : t1;
      int cls = hi == OB_GEOM_TRIMESH ? (lo == OB_GEOM_SPHERE ? 1 : (lo == OB_GEOM_BOX ? 2 : 3)) : ((lo == OB_GEOM_BOX && hi == OB_GEOM_BOX) ? 4 : 5);
      if (pol.skip_if_connected && d.NJ && gv[o12.x].body >= 0 && gv[o12.y].body >= 0) cls = 0;   // mostly jointed pairs: the cheap test
      c_n[f] = (unsigned char)cls;
      atomicAdd(&c_cls[cls], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int k = 0; k < 7; k++) { const int c = c_cls[k]; c_cls[k] = run; run += c; } c_cls[7] = 0; }
    __syncthreads();
    for (int f = threadIdx.x; f < total; f += nthr) c_perm[atomicAdd(&c_cls[c_n[f]], 1)] = (unsigned short)f;
    __syncthreads();
    // (5b) one pair per thread in class order; contacts into the staging area
    for (int s0 = 0; s0 < total; s0 += nthr) {
      const int sidx = s0 + threadIdx.x;
      if (sidx < total) {
        const int f = c_perm[sidx];
        int v = 0;
        while (f >= c_np[v + 1]) v++;
        const int wv = wb + v;
        const unsigned char *smv = smem + (size_t)v * L.total;
        const int2 o12 = ((const int2 *)(smv + L.sorted))[f - c_np[v]];
        const ObPose *pose_v = (const ObPose *)(smv + L.pose);
        const int *walk_v = (const int *)(smv + L.walk_of);
        const ObGeom *gv = d.geom + (size_t)wv * d.NG;
        ObCg cg[CGCAP];
        int n = 0;
        bool connected = false;
        if (pol.skip_if_connected && d.NJ) {   // dAreConnectedExcluding(b1, b2, dJointTypeContact), ode.cpp:1529-1537
          const int b1 = gv[o12.x].body, b2 = gv[o12.y].body;
          if (b1 >= 0 && b2 >= 0) {
            const unsigned short *ps = d.padjstart + (size_t)wv * (d.NB + 1), *pa = d.padj + (size_t)wv * 2 * d.NJ;
            const ObJoint *pj = d.joint + (size_t)wv * d.NJ;
            for (int k = ps[b1]; k < ps[b1 + 1]; k++) {
              const ObJoint &jj = pj[pa[k]];
              const int other = jj.b1 == b1 ? jj.b2 : jj.b1;
              if (other == b2) connected = true;
            }
          }
        }
        int swapped, bverr = 0;
        if (!connected) n = ob_collide_pair_sel_t<MESH, CGCAP, XF>(&pose_v[walk_v[o12.x]], &pose_v[walk_v[o12.y]], maxc, cg, &swapped, d.meshes, &bverr);
        if (bverr) atomicOr(&d.world[wv].status, OB_ERR_BVH_STACK);
        int off = 0;
        if (n > 0) {
          off = atomicAdd(&c_cls[7], n);
          if (off + n > stage_cap) { atomicOr(&d.world[wv].status, OB_ERR_CONTACT_OVERFLOW); n = 0; }
        }
        for (int k = 0; k < n; k++) c_stage[off + k] = cg[k];
        c_stoff[f] = (unsigned short)off;
        c_n[f] = (unsigned char)n;
      }
    }
    __syncthreads();
    // (5c) ordered compaction: pooled pairs in (world, callback) order, contacts in pair order
    for (int base = 0; base < total; base += nthr) {
      const int f = base + threadIdx.x;
      const int n = f < total ? c_n[f] : 0;
      int tot;
      const int off = block_excl_scan(n, c_scan, &tot);
      int v = 0;
      if (f < total) {
        while (f >= c_np[v + 1]) v++;
        const int firstf = c_np[v] > base ? c_np[v] : base;
        if (f == firstf) c_first[v] = off;
      }
      __syncthreads();
      int j0 = 0;
      if (f < total) {
        j0 = c_cb[v] + off - c_first[v];
        const unsigned char *smv = smem + (size_t)v * L.total;
        const int2 o12 = ((const int2 *)(smv + L.sorted))[f - c_np[v]];
        ObContact *cout = d.contacts + (size_t)(wb + v) * d.NC;
        const ObCg *src = c_stage + c_stoff[f];
        for (int k = 0; k < n; k++) {
          const int j = j0 + k;
          if (j < d.NC) {
            ObContact c;
            for (int e = 0; e < 3; e++) { c.pos[e] = src[k].pos[e]; c.normal[e] = src[k].normal[e]; }
            c.depth = src[k].depth; c.g1 = o12.x; c.g2 = o12.y; c.side1 = src[k].side1; c.side2 = src[k].side2; c.policy = 0;
            cout[j] = c;
          }
        }
      }
      __syncthreads();
      if (f < total) {
        const int lastf = (c_np[v + 1] < base + nthr ? c_np[v + 1] : base + nthr) - 1;
        if (f == lastf) c_cb[v] = j0 + n;
      }
      __syncthreads();
    }
    if (tid == 0 && valid) {
      int nc = c_cb[warp];
      if (nc > d.NC) { nc = d.NC; atomicOr(&d.world[w].status, OB_ERR_CONTACT_OVERFLOW); }
      d.ncontacts[w] = nc;
      d.npairs[w] = np;
      atomicAdd(&d.counters->pairs, (unsigned long long)np);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// bulk state I/O kernels: API order is [world][creation index], device order is newest-first
__global__ void k_pack_state(ObBatchDev d, real *pos3, real *quat4, real *lvel3, real *avel3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.W * d.NB) return;
  const int w = i / d.NB, c = i - w * d.NB, nb = d.world[w].nb;
  if (c >= nb) return;
  const ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
  for (int k = 0; k < 3; k++) { pos3[(size_t)i * 3 + k] = s.pos[k]; lvel3[(size_t)i * 3 + k] = s.lvel[k]; avel3[(size_t)i * 3 + k] = s.avel[k]; }
  for (int k = 0; k < 4; k++) quat4[(size_t)i * 4 + k] = s.q[k];
}
__global__ void k_unpack_state(ObBatchDev d, const real *pos3, const real *quat4, const real *lvel3, const real *avel3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.W * d.NB) return;
  const int w = i / d.NB, c = i - w * d.NB, nb = d.world[w].nb;
  if (c >= nb) return;
  ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
  for (int k = 0; k < 3; k++) {
    if (pos3) s.pos[k] = pos3[(size_t)i * 3 + k];
    if (lvel3) s.lvel[k] = lvel3[(size_t)i * 3 + k];
    if (avel3) s.avel[k] = avel3[(size_t)i * 3 + k];
  }
  if (quat4) {
    // q is stored as given (must be unit): a state read with dBatchGetBodyState restores bit for bit
    for (int k = 0; k < 4; k++) s.q[k] = quat4[(size_t)i * 4 + k];
    ob_RfromQ(s.R, s.q);
  }
}
__global__ void k_add_forces(ObBatchDev d, const real *f3, const real *t3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.W * d.NB) return;
  const int w = i / d.NB, c = i - w * d.NB, nb = d.world[w].nb;
  if (c >= nb) return;
  ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
  for (int k = 0; k < 3; k++) {
    if (f3) s.facc[k] += f3[(size_t)i * 3 + k];
    if (t3) s.tacc[k] += t3[(size_t)i * 3 + k];
  }
}

// ------------------------------------------------------------------------------------
struct ObBackend {
  ObBatchDev d;
  int device;
  cudaStream_t stream;
  std::vector<void *> allocs;
  real *st_dev;      // packed state staging on the device: pos3|quat4|lvel3|avel3
  real *st_host;     // pinned
  size_t st_elems;   // W*NB
  size_t smem_collide, smem_prep, smem_sched, smem_sched_lane, smem_sor, smem_post, smem_collide_tile;
  int prep_tile;                      // tile width of k_prep (defaults to `tile`)
  int collide_tile, tile_stage_cap;   // k_collide_tile serves the batch (worlds of <= 8 geoms)
  int sor_deep;     // 1: k_sor with the deep index prefetch (worlds with many rows)
  int sched_lane;   // 1: k_sched_lane (one lane per world) fits shared memory
  // independent worlds are stepped in nchunks chunks, each on its own stream: the chunks drift apart, so the
  // ALU-bound collide of one chunk overlaps the latency-bound SOR of another instead of running back to back
  int nchunks;
  cudaStream_t cstream[8];
  cudaEvent_t cev[9];
  int grid, grid_step, grid_sor, tile;
  cudaEvent_t ev[8];   // 0,1: user timer; 2..7 per-kernel timing
  int ktiming;
  double kms[OBK_NKERNELS];
  long long klaunch[OBK_NKERNELS];
  // one large world (ob_large.h)
  int large;
  ObLargeDev L;
  int *lw_host;        // pinned: scalars + segment table read back for launch sizing
  int lw_rounds, lw_ncol, lw_stat[8], lw_sor_grid[3];
  double lw_ms[8];     // geoms+sort, pairs, narrow, colour, assemble, sor, post (CUDA events, when kernel timing is on)
  cudaEvent_t lw_ev[9];
  // SOR phase split over the GPUs of one box (ObLwSplit): flag words behind fc in ONE allocation (one IPC handle)
  unsigned *lw_flags;
  size_t lw_flags_off;          // bytes from L.fc to lw_flags
  int lw_split_on, lw_split_grid[3], lw_split_threads;
  ObLwSplit lw_split;
  void *lw_peer_base[OB_LW_MAXRANKS];   // cudaIpcOpenMemHandle mappings to close
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_)); goto fail; } } while (0)

template <class T> static cudaError_t dalloc(ObBackend *b, T **p, size_t n) {
  void *q = 0;
  cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
  if (e == cudaSuccess) { b->allocs.push_back(q); cudaMemset(q, 0, (n ? n : 1) * sizeof(T)); }
  *p = (T *)q;
  return e;
}
#include "ob_large_kernels.cuh"

ObBackend *obk_create(const ObBatchDev &caps, int device, char *err, size_t errlen) {
  ObBackend *b = new ObBackend;
  b->d = caps; b->device = device; b->stream = 0; b->st_dev = 0; b->st_host = 0;
  b->ktiming = 0;
  for (int k = 0; k < OBK_NKERNELS; k++) { b->kms[k] = 0; b->klaunch[k] = 0; }
  for (int k = 0; k < 8; k++) b->ev[k] = 0;
  ObBatchDev &d = b->d;
  const size_t W = d.W;
  int ndev = 0;
  cudaDeviceProp prop;
  b->nchunks = 1;
  for (int k = 0; k < 8; k++) b->cstream[k] = 0;
  for (int k = 0; k < 9; k++) b->cev[k] = 0;
  b->lw_flags = 0; b->lw_flags_off = 0; b->lw_split_on = 0; memset(&b->lw_split, 0, sizeof b->lw_split);
  for (int k = 0; k < OB_LW_MAXRANKS; k++) b->lw_peer_base[k] = 0;
  b->large = d.large; b->lw_host = 0; b->lw_rounds = 0; b->lw_ncol = 0;
  for (int k = 0; k < 8; k++) b->lw_stat[k] = 0;
  memset(&b->L, 0, sizeof b->L);
  for (int k = 0; k < 9; k++) b->lw_ev[k] = 0;
  for (int k = 0; k < 8; k++) b->lw_ms[k] = 0;
  if (!d.large && (d.NB > 32000 || d.NC > 32000 || d.NR > 65000)) { snprintf(err, errlen, "world too large for the CTA-per-world path (NB=%d NC=%d NR=%d)", d.NB, d.NC, d.NR); goto fail2; }
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { snprintf(err, errlen, "no CUDA device available (this library has no CPU fallback)"); goto fail2; }
  CK(cudaSetDevice(device));
  CK(cudaGetDeviceProperties(&prop, device));
  CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  for (int k = 0; k < 8; k++) CK(cudaEventCreate(&b->ev[k]));
  for (int k = 0; k < 8; k++) CK(cudaStreamCreateWithFlags(&b->cstream[k], cudaStreamNonBlocking));
  for (int k = 0; k < 9; k++) CK(cudaEventCreateWithFlags(&b->cev[k], cudaEventDisableTiming));
  d.wbeg = 0; d.wend = d.W;
  if (d.large) {
    if (lw_create(b, err, errlen)) goto fail;
    return b;
  }
  CK(dalloc(b, &d.world, W));
  CK(dalloc(b, &d.bdyn, W * d.NB));
  CK(dalloc(b, &d.bconst, W * d.NB));
  CK(dalloc(b, &d.geom, W * d.NG));
  CK(dalloc(b, &d.glist, W * d.NG));
  CK(dalloc(b, &d.sapstate, W * (d.NG + 3)));
  CK(dalloc(b, &d.policy, (size_t)d.npolicy));
  CK(dalloc(b, &d.meshes, (size_t)(d.nmesh ? d.nmesh : 1)));
  CK(dalloc(b, &d.joint, W * (d.NJ ? d.NJ : 1)));
  CK(dalloc(b, &d.njoints, W));
  CK(dalloc(b, &d.padjstart, W * (d.NB + 1)));
  CK(dalloc(b, &d.padj, W * 2 * (d.NJ ? d.NJ : 1)));
  CK(dalloc(b, &d.npairs, W));
  CK(dalloc(b, &d.pairs, W * d.NP * 2));
  CK(dalloc(b, &d.ncontacts, W));
  CK(dalloc(b, &d.contacts, W * d.NC));
  if (d.NEP > OB_MAXEPOCH) { snprintf(err, errlen, "more than %d SOR iterations are not supported", 8 * OB_MAXEPOCH); goto fail; }
  CK(dalloc(b, &d.rows, W * d.NR * OB_ROWW));
  CK(dalloc(b, &d.stepinfo, W * SI_WORDS));
  CK(dalloc(b, &d.ibody, W * d.NB));
  CK(dalloc(b, &d.isz, W * d.NB * 4));
  CK(dalloc(b, &d.jrow, W * (d.NC + d.NJ + 1)));
  CK(dalloc(b, &d.ijoint, W * (d.NC + d.NJ)));
  CK(dalloc(b, &d.jside, W * (d.NJ ? d.NJ : 1) * 4 * OB_NSIDE));
  CK(dalloc(b, &d.sched, W * d.NEP * d.NR));
  CK(dalloc(b, &d.pstart, W * d.NEP * (d.NR + 1)));
  CK(dalloc(b, &d.invIw, W * d.NB * 12));
  CK(dalloc(b, &d.tmp1, W * d.NB * 8));
  d.rowJ = d.rowiMJ = d.rowJc = d.rowS = 0; d.rowI = 0;
  CK(dalloc(b, &d.lambda, W * d.NR));
  CK(dalloc(b, &d.nrows, W));
  CK(dalloc(b, &d.fback, W * (d.NC + d.NJ) * 12));
  d.csurf = 0; d.cfdir1 = 0;
  if (d.dropin) { CK(dalloc(b, &d.csurf, W * d.NC)); CK(dalloc(b, &d.cfdir1, W * d.NC * 4)); }
  CK(dalloc(b, &d.counters, (size_t)1));
  b->st_elems = W * d.NB;
  CK(dalloc(b, &b->st_dev, b->st_elems * 13));
  b->smem_collide = collide_smem(d.NG, d.NP).total;
  {
    // tile width: G lanes per world, 32/G worlds per warp.  Narrow tiles waste fewer lanes in the
    // dependency rounds of the SOR sweep; wide tiles finish one world sooner.  Heuristic on batch size.
    int G = W >= 1024 ? 8 : (W >= 256 ? 16 : 32);
    if (W >= 4096 && d.NB <= 8) G = 4;   // tiny worlds (config 3: 5 bodies, <= 3 rows per level): 8 worlds per warp, 10.5 -> 8.3 ms/step
    const char *e = getenv("OB_TILE");
    if (e && (atoi(e) == 4 || atoi(e) == 8 || atoi(e) == 16 || atoi(e) == 32)) G = atoi(e);
    b->tile = G;
    // row assembly (k_prep) has its own tile width: it is a throughput kernel (lane per joint / row) that waits on
    // scattered loads, so more, narrower-batched warps can pay even where the sweep prefers few lanes per world
    b->prep_tile = G;
    // measured on B200: contact-only worlds with hundreds of rows (config 2) 0.44 -> 0.38 ms with one world per warp;
    // jointed / tiny worlds (configs 3, 4) are fastest at the sweep's own width (config 3: 1.26 / 1.35 / 1.50 / 2.17 ms at 4 / 8 / 16 / 32)
    if (d.NJ == 0 && d.NC >= 96) b->prep_tile = 32;
    { const char *pe = getenv("OB_PREP_TILE"); if (pe && (atoi(pe) == 4 || atoi(pe) == 8 || atoi(pe) == 16 || atoi(pe) == 32)) b->prep_tile = atoi(pe); }
    b->smem_prep = prep_tile_smem(d.NB, d.NC, d.NJ, d.NR).total * (32 / b->prep_tile);
    b->smem_sor = sor_tile_smem(d.NB, d.NR).total * (32 / G);
    b->smem_sched = sched_smem(d.NB, d.NR).total;
    b->smem_post = post_tile_smem(d.NG).total * (32 / G);
    b->grid_step = (int)((W + (32 / G) - 1) / (32 / G));
    b->grid_sor = b->grid_step;
    { const char *g = getenv("OB_GRID_SOR"); if (g && atoi(g) > 0 && atoi(g) < b->grid_sor) b->grid_sor = atoi(g); }
  }
  if (d.NB > 254 || d.NG > 255 || d.NC + d.NJ > 65000) { snprintf(err, errlen, "world too large for the tile-per-world step kernel (NB=%d NG=%d NR=%d)", d.NB, d.NG, d.NR); goto fail; }
  if (b->smem_collide > (size_t)prop.sharedMemPerBlockOptin || b->smem_prep > (size_t)prop.sharedMemPerBlockOptin ||
      b->smem_sor > (size_t)prop.sharedMemPerBlockOptin) {
    snprintf(err, errlen, "world does not fit one CTA's shared memory (collide %zu B, prep %zu B, sor %zu B, limit %zu B)",
             b->smem_collide, b->smem_prep, b->smem_sor, (size_t)prop.sharedMemPerBlockOptin);
    goto fail;
  }
  CK(cudaFuncSetAttribute(k_collide<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_collide));
  CK(cudaFuncSetAttribute(k_collide<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_collide));
  CK(cudaFuncSetAttribute(k_collide<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_collide));
  // small worlds: OB_TILE_WPC worlds per CTA with a pooled, class-grouped narrowphase (k_collide_tile)
  b->collide_tile = 0;
  if (d.NG <= 8 && !getenv("OB_COLLIDE_NOTILE")) {
    long long cap = (long long)OB_TILE_WPC * d.NC;
    b->tile_stage_cap = (int)(cap < 60000 ? cap : 60000);
    b->smem_collide_tile = (size_t)OB_TILE_WPC * collide_smem(d.NG, d.NP).total + collide_tile_smem(d.NG, d.NP, OB_TILE_WPC, b->tile_stage_cap).total;
    if (b->smem_collide_tile <= (size_t)prop.sharedMemPerBlockOptin && (long long)OB_TILE_WPC * d.NP < 65000) {
      b->collide_tile = 1;
      CK(cudaFuncSetAttribute(k_collide_tile<false, false, OB_TILE_WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_collide_tile));
      CK(cudaFuncSetAttribute(k_collide_tile<true, false, OB_TILE_WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_collide_tile));
      CK(cudaFuncSetAttribute(k_collide_tile<true, true, OB_TILE_WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_collide_tile));
    }
  }
#define OB_SETSMEM(GG) \
  CK(cudaFuncSetAttribute(k_prep<GG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_prep)); \
  CK(cudaFuncSetAttribute(k_prep<GG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_prep)); \
  CK(cudaFuncSetAttribute(k_sor<GG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_sor)); \
  CK(cudaFuncSetAttribute(k_sor<GG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_sor)); \
  CK(cudaFuncSetAttribute(k_post<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_post));
  OB_SETSMEM(4) OB_SETSMEM(8) OB_SETSMEM(16) OB_SETSMEM(32)
#undef OB_SETSMEM
  CK(cudaFuncSetAttribute(k_sched<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_sched));
  CK(cudaFuncSetAttribute(k_sched<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_sched));
  CK(cudaFuncSetAttribute(k_sched<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_sched));
  b->sor_deep = d.NR > 256 ? 1 : 0;
  { const char *e = getenv("OB_SOR_DEEP"); if (e) b->sor_deep = atoi(e) != 0; }
  b->smem_sched_lane = sched_lane_smem(d.NB, d.NR).total;
  // measured on B200: one lane per world wins for many small worlds (config 3: 65536 worlds x 56 rows, 0.70 -> 0.43 ms),
  // the warp per world for fewer, larger ones (config 2: 4096 x 377 rows, 0.40 vs 2.7 ms: too few warps to hide the chain latency)
  b->sched_lane = b->smem_sched_lane <= (size_t)prop.sharedMemPerBlockOptin && ((W >= 8192 && d.NR <= 256) || getenv("OB_SCHED_LANE")) && !getenv("OB_SCHED_WARP");
  if (b->sched_lane) CK(cudaFuncSetAttribute(k_sched_lane, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_sched_lane));
  {
    int nch = 1;   // measured on B200 (configs 2-4): 2-8 chunks change the step time by -4 % .. +10 %, so off unless OB_CHUNKS asks
    const char *e = getenv("OB_CHUNKS");
    if (e && atoi(e) >= 1 && atoi(e) <= 8) nch = atoi(e);
    b->nchunks = nch;
  }
  {
    // grid: every world gets its own CTA up to 16 resident CTAs per SM worth of blocks, beyond that grid-stride
    int cap = prop.multiProcessorCount * 32;
    b->grid = (int)W < cap ? (int)W : cap;
  }
  return b;
fail:
  for (int k = 0; k < OB_LW_MAXRANKS; k++) if (b->lw_peer_base[k]) cudaIpcCloseMemHandle(b->lw_peer_base[k]);
  for (size_t i = 0; i < b->allocs.size(); i++) cudaFree(b->allocs[i]);
  if (b->st_host) cudaFreeHost(b->st_host);
  if (b->lw_host) cudaFreeHost(b->lw_host);
  if (b->stream) cudaStreamDestroy(b->stream);
fail2:
  delete b;
  return 0;
}

void obk_destroy(ObBackend *b) {
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  for (int k = 0; k < OB_LW_MAXRANKS; k++) if (b->lw_peer_base[k]) cudaIpcCloseMemHandle(b->lw_peer_base[k]);
  for (size_t i = 0; i < b->allocs.size(); i++) cudaFree(b->allocs[i]);
  if (b->st_host) cudaFreeHost(b->st_host);
  if (b->lw_host) cudaFreeHost(b->lw_host);
  for (int k = 0; k < 9; k++) if (b->lw_ev[k]) cudaEventDestroy(b->lw_ev[k]);
  for (int k = 0; k < 8; k++) if (b->ev[k]) cudaEventDestroy(b->ev[k]);
  for (int k = 0; k < 8; k++) if (b->cstream[k]) cudaStreamDestroy(b->cstream[k]);
  for (int k = 0; k < 9; k++) if (b->cev[k]) cudaEventDestroy(b->cev[k]);
  cudaStreamDestroy(b->stream);
  delete b;
}
ObBatchDev *obk_arrays(ObBackend *b) { return &b->d; }
int obk_h2d(ObBackend *b, void *dst, const void *src, size_t n) {
  cudaSetDevice(b->device);
  if (cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, b->stream) != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_d2h(ObBackend *b, void *dst, const void *src, size_t n) {
  cudaSetDevice(b->device);
  if (cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, b->stream) != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_memset(ObBackend *b, void *dst, int v, size_t n) {
  cudaSetDevice(b->device);
  if (cudaMemsetAsync(dst, v, n, b->stream) != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_sync(ObBackend *b) { cudaSetDevice(b->device); return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1; }
void *obk_stream(ObBackend *b) { return (void *)b->stream; }
long long obk_launch_count(void) { return g_launches; }
int obk_timer_start(ObBackend *b) { cudaSetDevice(b->device); return cudaEventRecord(b->ev[0], b->stream) == cudaSuccess ? 0 : -1; }
int obk_timer_stop(ObBackend *b, float *ms) {
  cudaSetDevice(b->device);
  if (cudaEventRecord(b->ev[1], b->stream) != cudaSuccess) return -1;
  if (cudaEventSynchronize(b->ev[1]) != cudaSuccess) return -1;
  return cudaEventElapsedTime(ms, b->ev[0], b->ev[1]) == cudaSuccess ? 0 : -1;
}
void obk_set_kernel_timing(ObBackend *b, int enable) {
  b->ktiming = enable;
  for (int k = 0; k < 8; k++) b->lw_ms[k] = 0;
  b->lw_stat[7] = 0;
  for (int k = 0; k < OBK_NKERNELS; k++) { b->kms[k] = 0; b->klaunch[k] = 0; }
}
void obk_get_kernel_times(ObBackend *b, double *ms, long long *l) {
  for (int k = 0; k < OBK_NKERNELS; k++) { ms[k] = b->kms[k]; l[k] = b->klaunch[k]; }
}
int obk_large_stats(ObBackend *b, int *ints8, double *ms8) {
  if (!b->large) return -1;
  for (int k = 0; k < 8; k++) { ints8[k] = b->lw_stat[k]; ms8[k] = b->lw_ms[k]; }
  return 0;
}
const char *obk_kernel_name(int k) { static const char *n[] = {"k_collide", "k_prep", "k_sched", "k_sor", "k_post"}; return k >= 0 && k < 5 ? n[k] : ""; }

// one step of the world range [w0, w1) on stream st
template <int G> static void launch_step(ObBackend *b, real h, int taps, int phases, int w0, int w1, cudaStream_t st, bool timing) {
  cudaEvent_t *ev = b->ev + 2;
  constexpr int T = 32 / G;
  ObBatchDev d = b->d;
  d.wbeg = w0; d.wend = w1;
  const int W = w1 - w0;
  const int cap = b->grid;   // resident-CTA cap computed for the whole batch
  if (timing) cudaEventRecord(ev[0], st);
  if (phases & OBK_PHASE_COLLIDE) {
    // CTA width follows the world size: the widest loop is the ng*ng candidate-pair scan
    int ct = d.NG <= 8 ? 32 : (d.NG <= 20 ? 64 : OB_THREADS);
    { static const char *e = getenv("OB_COLLIDE_THREADS"); if (e && (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 96 || atoi(e) == 128)) ct = atoi(e); }
    const int grid = W < cap ? W : cap;
    if (b->collide_tile) {
      const int tiles = (W + OB_TILE_WPC - 1) / OB_TILE_WPC;
      const int tgrid = tiles < cap ? tiles : cap;
      // batches with geom transforms run the <MESH = true, XF = true> instantiation (a superset: the mesh arms only fire for trimesh geoms)
      if (d.any_xf) k_collide_tile<true, true, OB_TILE_WPC><<<tgrid, 32 * OB_TILE_WPC, b->smem_collide_tile, st>>>(d, b->tile_stage_cap);
      else if (d.nmesh) k_collide_tile<true, false, OB_TILE_WPC><<<tgrid, 32 * OB_TILE_WPC, b->smem_collide_tile, st>>>(d, b->tile_stage_cap);
      else k_collide_tile<false, false, OB_TILE_WPC><<<tgrid, 32 * OB_TILE_WPC, b->smem_collide_tile, st>>>(d, b->tile_stage_cap);
    } else if (d.any_xf) k_collide<true, true><<<grid, ct, b->smem_collide, st>>>(d);
    else if (d.nmesh) k_collide<true, false><<<grid, ct, b->smem_collide, st>>>(d);
    else k_collide<false, false><<<grid, ct, b->smem_collide, st>>>(d);
    g_launches++;
  }
  if (timing) cudaEventRecord(ev[1], st);
  if (phases & OBK_PHASE_STEP) {
    const int gstep = (W + T - 1) / T;
    int gsor = gstep;
    if (b->grid_sor < b->grid_step) gsor = gsor < b->grid_sor ? gsor : b->grid_sor;
#define OB_LAUNCH_PREP(GP) { const int gp = (W + (32 / GP) - 1) / (32 / GP); \
      if (d.NJ > 0) k_prep<GP, true><<<gp, 32, b->smem_prep, st>>>(d, h, taps); else k_prep<GP, false><<<gp, 32, b->smem_prep, st>>>(d, h, taps); }
    if (b->prep_tile == 4) OB_LAUNCH_PREP(4) else if (b->prep_tile == 8) OB_LAUNCH_PREP(8) else if (b->prep_tile == 16) OB_LAUNCH_PREP(16) else OB_LAUNCH_PREP(32)
#undef OB_LAUNCH_PREP
    if (timing) cudaEventRecord(ev[2], st);
    if (b->sched_lane) k_sched_lane<<<(W + 31) / 32, 32, b->smem_sched_lane, st>>>(d, G);
    else if (d.NB <= 64) k_sched<2><<<W, 32, b->smem_sched, st>>>(d, G, taps);
    else if (d.NB <= 128) k_sched<4><<<W, 32, b->smem_sched, st>>>(d, G, taps);
    else k_sched<8><<<W, 32, b->smem_sched, st>>>(d, G, taps);
    if (timing) cudaEventRecord(ev[3], st);
    if (b->sor_deep) k_sor<G, true><<<gsor, 32, b->smem_sor, st>>>(d, taps);
    else k_sor<G, false><<<gsor, 32, b->smem_sor, st>>>(d, taps);
    if (timing) cudaEventRecord(ev[4], st);
    k_post<G><<<gstep, 32, b->smem_post, st>>>(d, h);
    g_launches += 4;
  }
  if (timing && phases == (OBK_PHASE_COLLIDE | OBK_PHASE_STEP)) {
    cudaEventRecord(ev[5], st);
    if (cudaEventSynchronize(ev[5]) == cudaSuccess)
      for (int k = 0; k < 5; k++) { float m = 0; cudaEventElapsedTime(&m, ev[k], ev[k + 1]); b->kms[k] += m; b->klaunch[k]++; }
  }
}
template <int G> static void launch_steps(ObBackend *b, real h, int nsteps, int taps, int phases) {
  const int W = b->d.W;
  // per-kernel timing and the parity taps run unchunked on the main stream
  const int nch = (b->ktiming || (taps & 1) || b->d.dropin || nsteps < 2) ? 1 : b->nchunks;
  if (nch <= 1) {
    for (int s = 0; s < nsteps; s++) {
      // parity tap: joints that enter no island (attached to no body / to disabled bodies) report zero feedback
      if ((taps & 1) && !b->d.dropin && b->d.fback) cudaMemsetAsync(b->d.fback, 0, sizeof(real) * 12 * (size_t)W * (b->d.NC + b->d.NJ), b->stream);
      launch_step<G>(b, h, taps, phases, 0, W, b->stream, b->ktiming != 0);
    }
    return;
  }
  // fork: every chunk stream starts after what is already queued on the main stream; worlds are independent,
  // so the chunks never wait for each other between steps; join: the main stream waits for all of them
  cudaEventRecord(b->cev[8], b->stream);
  const int per = ((W + nch - 1) / nch + 31) / 32 * 32;   // multiple of 32 keeps warp tiles whole
  for (int c = 0; c < nch; c++) {
    const int w0 = c * per, w1 = (c + 1) * per < W ? (c + 1) * per : W;
    if (w0 >= w1) continue;
    cudaStreamWaitEvent(b->cstream[c], b->cev[8], 0);
    for (int s = 0; s < nsteps; s++) launch_step<G>(b, h, taps, phases, w0, w1, b->cstream[c], false);
    cudaEventRecord(b->cev[c], b->cstream[c]);
    cudaStreamWaitEvent(b->stream, b->cev[c], 0);
  }
}

static int run_steps(ObBackend *b, real h, int nsteps, int taps, int phases, char *err, size_t errlen) {
  cudaSetDevice(b->device);
  if (getenv("OB_SEQ")) taps |= 2;
  if (getenv("OB_CHECK")) taps |= 4;
  if (getenv("OB_SYNC2")) taps |= 8;
  if (b->large) {
    if (phases != (OBK_PHASE_COLLIDE | OBK_PHASE_STEP)) { snprintf(err, errlen, "the large-world path runs whole steps only"); return -1; }
    for (int s = 0; s < nsteps; s++) if (lw_step(b, h, taps, err, errlen)) return -1;
    return 0;
  }
  if (b->tile == 4) launch_steps<4>(b, h, nsteps, taps, phases);
  else if (b->tile == 8) launch_steps<8>(b, h, nsteps, taps, phases);
  else if (b->tile == 16) launch_steps<16>(b, h, nsteps, taps, phases);
  else launch_steps<32>(b, h, nsteps, taps, phases);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  if (e != cudaSuccess) { snprintf(err, errlen, "kernel launch/exec failed: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
int obk_step(ObBackend *b, real h, int nsteps, int taps, char *err, size_t errlen) {
  return run_steps(b, h, nsteps, taps, OBK_PHASE_COLLIDE | OBK_PHASE_STEP, err, errlen);
}
int obk_run_phases(ObBackend *b, real h, int phases, int taps, char *err, size_t errlen) {
  return run_steps(b, h, 1, taps, phases, err, errlen);
}

// dCollide outside a batch: one pair, one thread (the per-element collider functions are the same
// ones k_collide runs; there is no host implementation to fall back to)
struct PairCtx { ObPose *pose; ObCg *cg; int *n; cudaStream_t stream; bool ok; };
__global__ void k_collide_pair(const ObPose *pose, int flags, ObCg *out, int *n, ObMeshDev m0, ObMeshDev m1) {
  int swapped, bverr = 0;
  ObCg cg[OB_MAXC_LOCAL];
  ObMeshDev meshes[2] = {m0, m1};
  const int c = ob_collide_pair(pose[0], pose[1], flags, cg, &swapped, meshes, &bverr);
  for (int i = 0; i < c; i++) out[i] = cg[i];
  *n = bverr ? -2 : c;
}
int obk_collide_pair(const ObPose *a, const ObPose *b, int flags, ObCg *out, const ObMeshDev *meshes2, char *err, size_t errlen) {
  static PairCtx C = {0, 0, 0, 0, false};
  if (!C.ok) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { snprintf(err, errlen, "no CUDA device available (this library has no CPU fallback)"); return -1; }
    if (cudaMallocHost((void **)&C.pose, 2 * sizeof(ObPose)) != cudaSuccess || cudaMallocHost((void **)&C.cg, OB_MAXC_LOCAL * sizeof(ObCg)) != cudaSuccess ||
        cudaMallocHost((void **)&C.n, sizeof(int)) != cudaSuccess || cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking) != cudaSuccess) {
      snprintf(err, errlen, "obk_collide_pair: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
      return -1;
    }
    C.ok = true;
  }
  C.pose[0] = *a; C.pose[1] = *b;
  int maxc = flags & 0xffff;
  if (maxc > OB_MAXC_LOCAL) maxc = OB_MAXC_LOCAL;
  ObMeshDev m0, m1;
  memset(&m0, 0, sizeof m0); memset(&m1, 0, sizeof m1);
  if (meshes2) { m0 = meshes2[0]; m1 = meshes2[1]; }
  // page-locked buffers are mapped into the device address space (unified addressing): the kernel reads and writes them directly
  k_collide_pair<<<1, 1, 0, C.stream>>>(C.pose, (flags & ~0xffff) | maxc, C.cg, C.n, m0, m1);
  g_launches++;
  cudaError_t e = cudaStreamSynchronize(C.stream);
  if (e != cudaSuccess) { snprintf(err, errlen, "k_collide_pair failed: %s", cudaGetErrorString(e)); return -1; }
  const int n = *C.n;
  if (n == -2) { snprintf(err, errlen, "trimesh tree deeper than the traversal stack"); return -1; }
  for (int i = 0; i < n; i++) out[i] = C.cg[i];
  return n;
}

// dSpaceCollide2: thread per (space geom, query geom)
struct ObQueryGeom { ObPose pose; ObMeshDev mesh; int body; uint32_t cat, col; int pad; };
__global__ void k_collide2(ObBatchDev d, const ObQueryGeom *q, int nq, unsigned char *hit) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ng = d.world[0].ng;
  if (idx >= ng * nq) return;
  const int qi = idx / ng, g = idx - qi * ng;
  const ObGeom G = d.geom[g];
  unsigned char h = 0;
  if ((G.flags & OB_GEOM_ENABLED) && !(G.flags & OB_GEOM_ZERO_SIZED)) {   // GEOM_ENABLED(g), collision_kernel.h:75
    ObPose p;
    geom_pose_dev(G, d.bdyn, &p);
    real a[6], b[6];
    ob_aabb(p, a, d.meshes);
    ob_aabb(q[qi].pose, b, &q[qi].mesh);
    h = ob_aabb_pair_filter(G.body, q[qi].body, G.cat, G.col, q[qi].cat, q[qi].col, a, b) ? 1 : 0;
  }
  hit[(size_t)qi * d.NG + g] = h;
}
int obk_collide2(ObBackend *b, const ObPose *q, const int *qbody, const uint32_t *qcat, const uint32_t *qcol, const ObMeshDev *qmesh,
                 int nq, unsigned char *hit, char *err, size_t errlen) {
  cudaSetDevice(b->device);
  std::vector<ObQueryGeom> hq(nq);
  for (int i = 0; i < nq; i++) { hq[i].pose = q[i]; hq[i].mesh = qmesh[i]; hq[i].body = qbody[i]; hq[i].cat = qcat[i]; hq[i].col = qcol[i]; hq[i].pad = 0; }
  ObQueryGeom *dq = 0; unsigned char *dh = 0;
  const size_t nh = (size_t)nq * b->d.NG;
  cudaError_t e = cudaMalloc((void **)&dq, sizeof(ObQueryGeom) * nq);
  if (e == cudaSuccess) e = cudaMalloc((void **)&dh, nh);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dq, hq.data(), sizeof(ObQueryGeom) * nq, cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) {
    k_collide2<<<(unsigned)((nh + 127) / 128), 128, 0, b->stream>>>(b->d, dq, nq, dh);
    g_launches++;
    e = cudaMemcpyAsync(hit, dh, nh, cudaMemcpyDeviceToHost, b->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  if (dq) cudaFree(dq);
  if (dh) cudaFree(dh);
  if (e != cudaSuccess) { snprintf(err, errlen, "k_collide2 failed: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

int obk_mesh_upload(const float *verts, int nverts, const int *tris, int ntris, const ObBvNode *nodes, const unsigned char *useflags, int device, ObMeshDev *io) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  float *dv = 0; int *dt = 0; ObBvNode *dn = 0; int *df = 0;
  std::vector<int> vfirst((size_t)(nverts > 0 ? nverts : 1), -1);
  for (int c = 0; c < 3 * ntris; c++) { const int vi = tris[c]; if (vi >= 0 && vi < nverts && vfirst[vi] < 0) vfirst[vi] = c; }
  if (cudaMalloc((void **)&df, sizeof(int) * vfirst.size()) != cudaSuccess) return -1;
  if (cudaMemcpy(df, vfirst.data(), sizeof(int) * vfirst.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(df); return -1; }
  io->vfirst = df;
  io->useflags = 0;
  if (useflags) {
    unsigned char *du = 0;
    if (cudaMalloc((void **)&du, (size_t)ntris) != cudaSuccess) return -1;
    if (cudaMemcpy(du, useflags, (size_t)ntris, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(du); return -1; }
    io->useflags = du;
  }
  if (cudaMalloc((void **)&dv, sizeof(float) * 3 * (size_t)nverts) != cudaSuccess) return -1;
  if (cudaMalloc((void **)&dt, sizeof(int) * 3 * (size_t)ntris) != cudaSuccess) { cudaFree(dv); return -1; }
  if (cudaMalloc((void **)&dn, sizeof(ObBvNode) * (size_t)(ntris - 1)) != cudaSuccess) { cudaFree(dv); cudaFree(dt); return -1; }
  cudaError_t e = cudaMemcpy(dv, verts, sizeof(float) * 3 * (size_t)nverts, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dt, tris, sizeof(int) * 3 * (size_t)ntris, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dn, nodes, sizeof(ObBvNode) * (size_t)(ntris - 1), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(dv); cudaFree(dt); cudaFree(dn); return -1; }
  io->verts = dv; io->tris = dt; io->nodes = dn; io->nverts = nverts; io->ntris = ntris;
  return 0;
}
void obk_mesh_free(ObMeshDev *m) {
  if (m->verts) cudaFree((void *)m->verts);
  if (m->tris) cudaFree((void *)m->tris);
  if (m->nodes) cudaFree((void *)m->nodes);
  if (m->vfirst) cudaFree((void *)m->vfirst);
  if (m->useflags) cudaFree((void *)m->useflags);
  m->verts = 0; m->tris = 0; m->nodes = 0; m->vfirst = 0; m->useflags = 0;
}

// Bulk state I/O copies straight between the caller's buffers and the packed device staging
// array: with buffers from dBatchHostAlloc (pinned) these are plain DMA transfers, with pageable
// memory the driver stages them.
int obk_get_state(ObBackend *b, real *pos3, real *quat4, real *lvel3, real *avel3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *dp = b->st_dev, *dq = dp + n * 3, *dl = dq + n * 4, *da = dl + n * 3;
  k_pack_state<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, dp, dq, dl, da);
  g_launches++;
  cudaError_t e = cudaSuccess;
  if (pos3 && e == cudaSuccess) e = cudaMemcpyAsync(pos3, dp, n * 3 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (quat4 && e == cudaSuccess) e = cudaMemcpyAsync(quat4, dq, n * 4 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (lvel3 && e == cudaSuccess) e = cudaMemcpyAsync(lvel3, dl, n * 3 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (avel3 && e == cudaSuccess) e = cudaMemcpyAsync(avel3, da, n * 3 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (e != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_set_state(ObBackend *b, const real *pos3, const real *quat4, const real *lvel3, const real *avel3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *dp = b->st_dev, *dq = dp + n * 3, *dl = dq + n * 4, *da = dl + n * 3;
  cudaError_t e = cudaSuccess;
  if (pos3 && e == cudaSuccess) e = cudaMemcpyAsync(dp, pos3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (quat4 && e == cudaSuccess) e = cudaMemcpyAsync(dq, quat4, n * 4 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (lvel3 && e == cudaSuccess) e = cudaMemcpyAsync(dl, lvel3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (avel3 && e == cudaSuccess) e = cudaMemcpyAsync(da, avel3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (e != cudaSuccess) return -1;
  k_unpack_state<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, pos3 ? dp : 0, quat4 ? dq : 0, lvel3 ? dl : 0, avel3 ? da : 0);
  g_launches++;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_add_forces(ObBackend *b, const real *f3, const real *t3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *df = b->st_dev, *dt = df + n * 3;
  cudaError_t e = cudaSuccess;
  if (f3) e = cudaMemcpyAsync(df, f3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (t3 && e == cudaSuccess) e = cudaMemcpyAsync(dt, t3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (e != cudaSuccess) return -1;
  k_add_forces<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, f3 ? df : 0, t3 ? dt : 0);
  g_launches++;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
void *obk_host_alloc(size_t bytes) { void *p = 0; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : 0; }
void obk_host_free(void *p) { if (p) cudaFreeHost(p); }

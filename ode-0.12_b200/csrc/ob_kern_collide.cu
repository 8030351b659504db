// ob_kern_collide.cu — broadphase + narrowphase kernels of the batched path (k_collide, k_collide_tile), the
// one-pair kernel behind dCollide and the dSpaceCollide2 filter kernel.
#include "ob_backend_cuda.h"
#include "ob_ray.h"

// ------------------------------------------------------------------------------------
// Phases (1)-(4) of the collide kernels, shared by k_collide (one CTA per world, CTA = true) and k_collide_tile (one warp
// per world, CTA = false): poses + AABBs + cell boxes in walk order, ranks, the space's candidate filter, the sequence key
// of every surviving pair and the ordered pair list (V.sorted, d.pairs).  Returns the number of pairs.
//   (3) runs in two steps so that lanes stay busy: (3a) the cheap overlap filter over the strict triangle of geom pairs,
//       survivors compacted by warp ballots; (3b) the key of every survivor (ncu r02k: computing the key inside the scan
//       left 3 of 32 lanes live in every iteration).
//   (4) pairs are ranked inside their GROUP (stage, query rank) -- the two leading words of the key -- after a counting
//       sort over the groups, so a pair is compared with the few pairs of its own query instead of all np (np^2 / 2
//       seven-word compares before; ncu r02k: a quarter of the kernel's instructions).
struct ObKeyTail { int k[5]; };   // level, cx, cy, cz, node walk index (words 2..6 of ObPairKey)
__device__ __forceinline__ bool ob_tail_less(const ObKeyTail &a, const ObKeyTail &b) {
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (a.k[i] < b.k[i]) return true;
    if (a.k[i] > b.k[i]) return false;
  }
  return false;
}
struct CollideView {
  ObPose *pose; real *aabb; ObCellBox *cb; int *gid, *body; uint32_t *cat, *col; int *en, *hr, *br, *walk_of;
  float *sapkey; int *sapinit, *sappos, *sapwalk; ObKeyTail *key; int2 *o12, *sorted; int *misc;
  unsigned short *grp, *member; int *gstart, *gcur;
};
__device__ __forceinline__ CollideView collide_view(unsigned char *smem, const CollideSmem &L) {
  CollideView V;
  V.pose = (ObPose *)(smem + L.pose); V.aabb = (real *)(smem + L.aabb); V.cb = (ObCellBox *)(smem + L.cb);
  V.gid = (int *)(smem + L.gid); V.body = (int *)(smem + L.body); V.cat = (uint32_t *)(smem + L.cat); V.col = (uint32_t *)(smem + L.col);
  V.en = (int *)(smem + L.en); V.hr = (int *)(smem + L.hr); V.br = (int *)(smem + L.br); V.walk_of = (int *)(smem + L.walk_of);
  V.sapkey = (float *)(smem + L.sapkey); V.sapinit = (int *)(smem + L.sapinit); V.sappos = (int *)(smem + L.sappos); V.sapwalk = (int *)(smem + L.sapwalk);
  V.key = (ObKeyTail *)(smem + L.key); V.o12 = (int2 *)(smem + L.o12); V.sorted = (int2 *)(smem + L.sorted);
  V.misc = (int *)(smem + L.misc);   // [0]=npairs raw, [1]=nh, [2]=nbig, [3]=contact base, [4]=SAP unsorted, [5]=candidates, [8..40]=scan scratch
  V.grp = (unsigned short *)(smem + L.grp); V.member = (unsigned short *)(smem + L.member);
  V.gstart = (int *)(smem + L.gstart); V.gcur = (int *)(smem + L.gcur);
  return V;
}
template <bool CTA> __device__ __forceinline__ void collide_sync() { if (CTA) __syncthreads(); else __syncwarp(); }
// idx = b (b - 1) / 2 + a  ->  (a, b) with a < b: the strict lower triangle, so a candidate scan touches every unordered
// pair once (the float square root is only a first guess, the two loops make it exact)
__device__ __forceinline__ void tri_pair(int idx, int *a, int *b) {
  int r = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)idx)) * 0.5f);
  while (r * (r - 1) / 2 > idx) r--;
  while ((r + 1) * r / 2 <= idx) r++;
  *b = r; *a = idx - r * (r - 1) / 2;
}
template <bool CTA>
__device__ __forceinline__ int collide_broad(const ObBatchDev &d, int w, const CollideView &V, int tid, int nt) {
  ObPose *s_pose = V.pose; real *s_aabb = V.aabb; ObCellBox *s_cb = V.cb; int *s_gid = V.gid, *s_body = V.body;
  uint32_t *s_cat = V.cat, *s_col = V.col; int *s_en = V.en, *s_hr = V.hr, *s_br = V.br, *s_walk_of = V.walk_of;
  float *s_sapkey = V.sapkey; int *s_sapinit = V.sapinit, *s_sappos = V.sappos, *s_sapwalk = V.sapwalk;
  int2 *s_o12 = V.o12, *s_sorted = V.sorted; int *s_misc = V.misc;
  const unsigned FULL = 0xffffffffu;
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
    collide_sync<CTA>();
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
    collide_sync<CTA>();
    const int nh = s_misc[1], nbig = s_misc[2];
    // (2b) SAP: sorted position of every finite geom = RadixSort's output order (ob_broad.h)
    if (stype == OB_SPACE_SAP && nh > 0) {
      int *st = d.sapstate + (size_t)w * (d.NG + 3);
      const int nbk = nh + 1;                       // + FLT_MAX sentinel, element index nh
      const bool valid = st[0] != 0 && st[1] == nbk;
      if (tid == 0) s_sapkey[nh] = 3.402823466e+38f;
      for (int p = tid; p < nbk; p += nt) { if (valid) s_sapinit[st[2 + p]] = p; else s_sapinit[p] = p; }
      collide_sync<CTA>();
      for (int p = 1 + tid; p < nbk; p += nt) {
        const int e = valid ? st[2 + p] : p, e0 = valid ? st[1 + p] : p - 1;
        if (s_sapkey[e] < s_sapkey[e0]) s_misc[4] = 1;   // not already sorted
      }
      collide_sync<CTA>();
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
      collide_sync<CTA>();
      if (unsorted) for (int t = tid; t < nbk; t += nt) st[2 + s_sappos[t]] = t;
      if (tid == 0) { st[1] = nbk; if (unsorted) st[0] = 1; else if (!valid) st[0] = 0; }
    }
    // (3a) the space's overlap filter over every unordered geom pair; survivors compacted into a candidate list (in s_sorted,
    // which the ordering step fills only after the candidates are consumed)
    int *cand = (int *)s_sorted;
    const int candcap = 2 * d.NP, T = ng * (ng - 1) / 2;
    for (int base = 0; base < T; base += nt) {
      const int idx = base + tid;
      bool ok = false;
      int a = 0, b = 0;
      if (idx < T) {
        tri_pair(idx, &a, &b);
        if (s_en[a] && s_en[b]) {
          if (stype == OB_SPACE_SAP) {
            ok = ob_pair_filter_noaabb(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b]);
            if (ok && s_cb[a].level != OB_LEVEL_BIG && s_cb[b].level != OB_LEVEL_BIG) {
              const bool fa = s_sappos[s_hr[a]] < s_sappos[s_hr[b]];
              const int K = fa ? a : b, J = fa ? b : a;
              ok = ob_sap_sweep_test(s_sapkey[s_hr[J]], s_aabb + 6 * K, s_aabb + 6 * J, ax0, ax1, ax2);
            }
          } else ok = ob_aabb_pair_filter(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b], s_aabb + 6 * a, s_aabb + 6 * b);
        }
      }
      const unsigned m = __ballot_sync(FULL, ok);
      if (m) {
        const int lane = tid & 31, leader = __ffs(m) - 1;
        int pos = 0;
        if (lane == leader) pos = atomicAdd(&s_misc[5], __popc(m));
        pos = __shfl_sync(FULL, pos, leader) + __popc(m & ((1u << lane) - 1u));
        if (ok && pos < candcap) cand[pos] = a | (b << 16);
      }
    }
    collide_sync<CTA>();
    int ncand = s_misc[5];
    const bool cand_over = ncand > candcap;
    if (cand_over) ncand = candcap;
    // (3b) the sequence key of every candidate's callback: group = (stage, query rank), tail = the rest
    for (int c = tid; c < ncand; c += nt) {
      const int a = cand[c] & 0xffff, b = cand[c] >> 16;
      ObPairKey key;
      int first_is_a;
      if (stype == OB_SPACE_HASH) {
        if (!ob_hash_pair_key(a, b, s_cb[a], s_cb[b], s_hr[a], s_hr[b], s_br[a], s_br[b], nh, nbig, &key, &first_is_a)) continue;
      } else if (stype == OB_SPACE_SAP) {
        const bool ia = s_cb[a].level == OB_LEVEL_BIG, ib = s_cb[b].level == OB_LEVEL_BIG;
        for (int k = 0; k < 7; k++) key.k[k] = 0;
        if (!ia && !ib) {
          const int pa = s_sappos[s_hr[a]], pb = s_sappos[s_hr[b]];
          first_is_a = pa < pb;
          key.k[1] = first_is_a ? pa : pb; key.k[2] = first_is_a ? pb : pa;
        } else if (ia && ib) { key.k[0] = 1; key.k[1] = s_br[a]; key.k[3] = s_br[b]; first_is_a = 1; }
        else { key.k[0] = 1; key.k[1] = ia ? s_br[a] : s_br[b]; key.k[2] = 1; key.k[3] = ia ? s_hr[b] : s_hr[a]; first_is_a = ia; }
      } else {   // dxSimpleSpace::collide (collision_space.cpp:247-268): nested walk of the list
        for (int k = 0; k < 7; k++) key.k[k] = 0;
        key.k[1] = a; key.k[2] = b; first_is_a = 1;
      }
      const int slot = atomicAdd(&s_misc[0], 1);
      if (slot < d.NP) {
        ObKeyTail t;
        for (int k = 0; k < 5; k++) t.k[k] = key.k[2 + k];
        V.key[slot] = t;
        V.grp[slot] = (unsigned short)(key.k[0] * (ng + 1) + key.k[1]);   // stage <= 2, query rank <= ng
        s_o12[slot] = first_is_a ? make_int2(s_gid[a], s_gid[b]) : make_int2(s_gid[b], s_gid[a]);
      }
    }
    const int ngrp = 3 * (ng + 1);
    for (int g = tid; g <= ngrp; g += nt) V.gcur[g] = 0;
    collide_sync<CTA>();
    int np = s_misc[0];
    if (np > d.NP || cand_over) { if (np > d.NP) np = d.NP; if (tid == 0) atomicOr(&W.status, OB_ERR_PAIR_OVERFLOW); }
    // (4) order: counting sort over the groups, then the rank of a pair inside its group = the number of members with a
    // smaller tail (keys are unique)
    int *gpairs = d.pairs + (size_t)w * d.NP * 2;
    for (int p = tid; p < np; p += nt) atomicAdd(&V.gcur[V.grp[p]], 1);
    collide_sync<CTA>();
    if (tid < 32) {
      int carry = 0;
      for (int base = 0; base < ngrp; base += 32) {
        const int g = base + tid;
        const int v = g < ngrp ? V.gcur[g] : 0;
        int x = v;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd); if (tid >= dd) x += y; }
        if (g < ngrp) { V.gstart[g] = carry + x - v; V.gcur[g] = carry + x - v; }
        carry += __shfl_sync(FULL, x, 31);
      }
    }
    collide_sync<CTA>();
    for (int p = tid; p < np; p += nt) V.member[atomicAdd(&V.gcur[V.grp[p]], 1)] = (unsigned short)p;
    collide_sync<CTA>();
    for (int p = tid; p < np; p += nt) {
      const int g = V.grp[p];
      const ObKeyTail kp = V.key[p];
      const int g0 = V.gstart[g], g1 = V.gcur[g];
      int rank = g0;
      for (int mm = g0; mm < g1; mm++) rank += ob_tail_less(V.key[V.member[mm]], kp) ? 1 : 0;
      s_sorted[rank] = s_o12[p];
      gpairs[2 * rank] = s_o12[p].x; gpairs[2 * rank + 1] = s_o12[p].y;
    }
    collide_sync<CTA>();
    return np;
}

// ------------------------------------------------------------------------------------
// MESH: the batch has trimesh geoms (narrowphase with the BVH colliders, up to OB_MAXC_LOCAL contacts per
// pair); otherwise the primitive-only narrowphase with 8 contact slots per pair (box-box emits at most 8)
template <bool MESH, bool XF>
__global__ void __launch_bounds__(OB_THREADS) k_collide(ObBatchDev d) {
  constexpr int CGCAP = MESH ? OB_MAXC_LOCAL : 8;
  extern __shared__ __align__(16) unsigned char smem[];
  const CollideSmem L = collide_smem(d.NG, d.NP);
  const CollideView V = collide_view(smem, L);
  ObPose *s_pose = V.pose;
  int *s_walk_of = V.walk_of;
  int2 *s_sorted = V.sorted;
  int *s_misc = V.misc;
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int w = d.wbeg + blockIdx.x; w < d.wend; w += gridDim.x) {
    ObWorld &W = d.world[w];
    const ObGeom *geoms = d.geom + (size_t)w * d.NG;
    const int np = collide_broad<true>(d, w, V, tid, nt);
    // (5) narrowphase per pair in callback order, ordered compaction into contact joints
    const int nrows = d.policy[0].nrows;   // > 1: the row is chosen per pair by the geoms' category bits
    ObContact *cout = d.contacts + (size_t)w * d.NC;
    for (int base = 0; base < np; base += nt) {
      int p = base + tid;
      ObCg cg[CGCAP];
      int n = 0, o1 = 0, o2 = 0, row = 0;
      if (p < np) {
        o1 = s_sorted[p].x; o2 = s_sorted[p].y;
        if (nrows > 1) row = ob_policy_row(d.policy, geoms[o1].cat, geoms[o2].cat);
        const ObPolicy &pol = d.policy[row < 0 ? 0 : row];
        const int maxc = pol.max_contacts > CGCAP ? CGCAP : pol.max_contacts;
        bool connected = row < 0 || (pol.skip_static_pairs && geoms[o1].body < 0 && geoms[o2].body < 0);
        if (!connected && pol.skip_if_connected && d.NJ) {   // dAreConnectedExcluding(b1, b2, dJointTypeContact), ode.cpp:1529-1537
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
          c.depth = cg[k].depth; c.g1 = o1; c.g2 = o2; c.side1 = cg[k].side1; c.side2 = cg[k].side2; c.policy = row;
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
  const CollideView V = collide_view(sm, L);
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
    if (valid) np = collide_broad<false>(d, w, V, tid, nt);
    // (5) narrowphase over the pooled pairs of the CTA's worlds
    const int nrows = d.policy[0].nrows;   // > 1: the row is chosen per pair by the geoms' category bits
    const bool any_skip_connected = d.policy[0].skip_if_connected != 0;
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
      if (any_skip_connected && d.NJ && gv[o12.x].body >= 0 && gv[o12.y].body >= 0) cls = 0;   // mostly jointed pairs: the cheap test
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
        const int row = nrows > 1 ? ob_policy_row(d.policy, gv[o12.x].cat, gv[o12.y].cat) : 0;
        const ObPolicy &pol = d.policy[row < 0 ? 0 : row];
        const int maxc = pol.max_contacts > CGCAP ? CGCAP : pol.max_contacts;
        bool connected = row < 0 || (pol.skip_static_pairs && gv[o12.x].body < 0 && gv[o12.y].body < 0);
        if (!connected && pol.skip_if_connected && d.NJ) {   // dAreConnectedExcluding(b1, b2, dJointTypeContact), ode.cpp:1529-1537
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
        const ObGeom *gv = d.geom + (size_t)(wb + v) * d.NG;
        const int row = (n > 0 && nrows > 1) ? ob_policy_row(d.policy, gv[o12.x].cat, gv[o12.y].cat) : 0;
        for (int k = 0; k < n; k++) {
          const int j = j0 + k;
          if (j < d.NC) {
            ObContact c;
            for (int e = 0; e < 3; e++) { c.pos[e] = src[k].pos[e]; c.normal[e] = src[k].normal[e]; }
            c.depth = src[k].depth; c.g1 = o12.x; c.g2 = o12.y; c.side1 = src[k].side1; c.side2 = src[k].side2; c.policy = row;
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


// =====================================================================================
// Decoupled narrowphase: k_broad -> k_narrow -> k_contacts.
//
// ncu of the fused kernels (r02m): the warps of a CTA wait at the barrier behind the narrowphase for the slowest pair of
// their world(s) -- 28 % (configs[1]), 39 % ([3]) and 66 % ([2]) of all warp samples sit there, at 25 % occupancy.  Here the
// narrowphase is a kernel of its own that knows no worlds: k_broad (phases 1-4 per world as before) leaves the poses in
// global memory and files every pair that needs a collider call into the work list of its collider class; k_narrow is a
// persistent grid whose warps pull 32 items of ONE class at a time from a queue (heavy classes first), so a warp runs one
// collider on full lanes and no warp ever waits for another; contacts go into the world's pool in order of completion
// and k_contacts (a warp per world) moves them into the contact-joint array in callback order, which is what fixes
// dJointCreateContact's creation order.
__device__ __forceinline__ int collide_class(int t1, int t2) {
  const int lo = t1 < t2 ? t1 : t2, hi = t1 < t2 ? t2 : t1;
  return hi == OB_GEOM_TRIMESH ? (lo == OB_GEOM_SPHERE ? 1 : (lo == OB_GEOM_BOX ? 2 : 3)) : ((lo == OB_GEOM_BOX && hi == OB_GEOM_BOX) ? 4 : 5);
}
// class of pair p of world w (0: the policy makes no dCollide call for it) and the policy row that serves it
__device__ __forceinline__ int collide_classify(const ObBatchDev &d, int w, int o1, int o2, int nrows, int *row_out) {
  const ObGeom *geoms = d.geom + (size_t)w * d.NG;
  const ObGeom &G1 = geoms[o1], &G2 = geoms[o2];
  const int row = nrows > 1 ? ob_policy_row(d.policy, G1.cat, G2.cat) : 0;
  *row_out = row < 0 ? 0 : row;
  if (row < 0) return 0;
  const ObPolicy &pol = d.policy[row];
  if (pol.skip_static_pairs && G1.body < 0 && G2.body < 0) return 0;
  if (pol.skip_if_connected && d.NJ) {   // dAreConnectedExcluding(b1, b2, dJointTypeContact), ode.cpp:1529-1537
    const int b1 = G1.body, b2 = G2.body;
    if (b1 >= 0 && b2 >= 0) {
      const unsigned short *ps = d.padjstart + (size_t)w * (d.NB + 1), *pa = d.padj + (size_t)w * 2 * d.NJ;
      const ObJoint *pj = d.joint + (size_t)w * d.NJ;
      for (int k = ps[b1]; k < ps[b1 + 1]; k++) {
        const ObJoint &jj = pj[pa[k]];
        const int other = jj.b1 == b1 ? jj.b2 : jj.b1;
        if (other == b2) return 0;
      }
    }
  }
  return collide_class(G1.type, G2.type);
}
// Filing the pairs of one world (ordered list in V.sorted, np of them) into the class lists, in three steps so that the
// space in the lists can be reserved once per CTA: (A) classify, (B) reserve, (C) fill.  misc[41..46] = items per class,
// [47..52] = first slot in the class list, [53..58] = fill cursor (misc[8..40] is scan scratch in the fused kernels).
__device__ __forceinline__ void collide_emit_classify(const ObBatchDev &d, int w, const CollideView &V, int np, int tid, int nt, int nrows) {
  int *s_cnt = V.misc + 41;
  unsigned short *s_code = V.member;   // class | row << 4 per pair (the ranking step is done with its member list)
  for (int p = tid; p < np; p += nt) {
    int row;
    const int cls = collide_classify(d, w, V.sorted[p].x, V.sorted[p].y, nrows, &row);
    s_code[p] = (unsigned short)(cls | (row << 4));
    if (cls) atomicAdd(&s_cnt[cls], 1);
    else d.pn[(size_t)w * d.NP + p] = 0;
  }
  const int ng = d.world[w].ng;
  for (int i = tid; i < ng; i += nt) d.gpose[(size_t)w * d.NG + V.gid[i]] = V.pose[i];
}
__device__ __forceinline__ void collide_emit_fill(const ObBatchDev &d, int w, const CollideView &V, int np, int tid, int nt) {
  const int *s_base = V.misc + 47;
  int *s_cur = V.misc + 53;
  const unsigned short *s_code = V.member;
  const size_t wlcap = (size_t)d.W * d.NP;
  for (int p = tid; p < np; p += nt) {
    const int code = s_code[p], cls = code & 15;
    if (!cls) continue;
    const size_t slot = (size_t)cls * wlcap + (size_t)s_base[cls] + (size_t)atomicAdd(&s_cur[cls], 1);
    d.wl[2 * slot] = (unsigned)w; d.wl[2 * slot + 1] = (unsigned)p | ((unsigned)(code >> 4) << 24);
  }
  if (tid == 0) { d.pcount[w] = 0; d.npairs[w] = np; atomicAdd(&d.counters->pairs, (unsigned long long)np); }
}
// k_broad: one CTA per world
__global__ void __launch_bounds__(OB_THREADS) k_broad(ObBatchDev d) {
  extern __shared__ __align__(16) unsigned char smem[];
  const CollideSmem L = collide_smem(d.NG, d.NP);
  const CollideView V = collide_view(smem, L);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nrows = d.policy[0].nrows;
  for (int w = d.wbeg + blockIdx.x; w < d.wend; w += gridDim.x) {
    const int np = collide_broad<true>(d, w, V, tid, nt);
    if (tid < 18) V.misc[41 + tid] = 0;
    __syncthreads();
    collide_emit_classify(d, w, V, np, tid, nt, nrows);
    __syncthreads();
    if (tid >= 1 && tid < OB_NCLS && V.misc[41 + tid]) V.misc[47 + tid] = (int)atomicAdd(&d.wlcnt[tid], (unsigned)V.misc[41 + tid]);
    __syncthreads();
    collide_emit_fill(d, w, V, np, tid, nt);
    __syncthreads();
  }
}
// k_broad_tile: a warp per world, WPC worlds per CTA (worlds of a handful of geoms); one reservation per class and CTA
template <int WPC>
__global__ void __launch_bounds__(32 * WPC) k_broad_tile(ObBatchDev d) {
  extern __shared__ __align__(16) unsigned char smem[];
  const CollideSmem L = collide_smem(d.NG, d.NP);
  const int warp = threadIdx.x >> 5, tid = threadIdx.x & 31;
  const CollideView V = collide_view(smem + (size_t)warp * L.total, L);
  const int nrows = d.policy[0].nrows;
  for (int wb = d.wbeg + blockIdx.x * WPC; wb < d.wend; wb += gridDim.x * WPC) {
    const int w = wb + warp;
    const bool valid = w < d.wend;
    int np = 0;
    if (valid) np = collide_broad<false>(d, w, V, tid, 32);
    if (tid < 18) V.misc[41 + tid] = 0;
    __syncwarp();
    if (valid) collide_emit_classify(d, w, V, np, tid, 32, nrows);
    __syncthreads();
    if (threadIdx.x >= 1 && threadIdx.x < OB_NCLS) {   // thread c reserves class c for all WPC worlds of the CTA
      const int c = threadIdx.x;
      int tot = 0;
      for (int v = 0; v < WPC; v++) tot += ((const int *)(smem + (size_t)v * L.total + L.misc))[41 + c];
      if (tot) {
        int base = (int)atomicAdd(&d.wlcnt[c], (unsigned)tot);
        for (int v = 0; v < WPC; v++) { int *mv = (int *)(smem + (size_t)v * L.total + L.misc); mv[47 + c] = base; base += mv[41 + c]; }
      }
    }
    __syncthreads();
    if (valid) collide_emit_fill(d, w, V, np, tid, 32);
    __syncthreads();
  }
}
// k_narrow: persistent; a warp takes 32 items of one class at a time.  Queue order: the mesh classes, box-box, the rest.
template <bool MESH, bool XF>
__global__ void __launch_bounds__(128) k_narrow(ObBatchDev d) {
  constexpr int CGCAP = MESH ? OB_MAXC_LOCAL : 8;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const size_t wlcap = (size_t)d.W * d.NP;
  const int order[OB_NCLS - 1] = {3, 2, 1, 4, 5};
  unsigned cnt[OB_NCLS];
#pragma unroll
  for (int c = 1; c < OB_NCLS; c++) cnt[c] = d.wlcnt[c];
  for (;;) {
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(&d.wlcnt[8], 1u);
    v = __shfl_sync(FULL, v, 0);
    int cls = 0;
#pragma unroll
    for (int k = 0; k < OB_NCLS - 1; k++) {
      const int c = order[k];
      const unsigned nch = (cnt[c] + 31u) >> 5;
      if (!cls) { if (v < nch) cls = c; else v -= nch; }
    }
    if (!cls) break;
    const unsigned idx = v * 32u + (unsigned)lane;
    if (idx < cnt[cls]) {
      const size_t slot = (size_t)cls * wlcap + idx;
      const int w = (int)d.wl[2 * slot];
      const unsigned pw = d.wl[2 * slot + 1];
      const int p = (int)(pw & 0xffffffu), row = (int)(pw >> 24);
      const int *pr = d.pairs + ((size_t)w * d.NP + p) * 2;
      const int o1 = pr[0], o2 = pr[1];
      const ObPolicy &pol = d.policy[row];
      const int maxc = pol.max_contacts > CGCAP ? CGCAP : pol.max_contacts;
      const ObPose *gp = d.gpose + (size_t)w * d.NG;
      ObCg cg[CGCAP];
      int swapped, bverr = 0;
      const int n = ob_collide_pair_sel_t<MESH, CGCAP, XF>(&gp[o1], &gp[o2], maxc, cg, &swapped, d.meshes, &bverr);
      if (bverr) atomicOr(&d.world[w].status, OB_ERR_BVH_STACK);
      int off = 0;
      if (n > 0) {
        off = atomicAdd(&d.pcount[w], n);
        ObContact *out = d.pool + (size_t)w * d.NC;
        for (int k = 0; k < n; k++) {
          if (off + k >= d.NC) break;   // k_contacts reports the overflow
          ObContact c;
          for (int e = 0; e < 3; e++) { c.pos[e] = cg[k].pos[e]; c.normal[e] = cg[k].normal[e]; }
          c.depth = cg[k].depth; c.g1 = o1; c.g2 = o2; c.side1 = cg[k].side1; c.side2 = cg[k].side2; c.policy = row;
          out[off + k] = c;
        }
      }
      d.pn[(size_t)w * d.NP + p] = (unsigned char)n;
      d.poff[(size_t)w * d.NP + p] = off;
    }
  }
}
// k_contacts<G>: G lanes per world (32 / G worlds per warp) move the pool into the contact-joint array, pairs in callback order.
// G follows the pair capacity: a warp per world would leave most lanes idle on worlds of a dozen pairs (configs[2]: 65536 worlds).
template <int G>
__global__ void __launch_bounds__(128) k_contacts(ObBatchDev d) {
  constexpr int T = 32 / G;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
  const int w = d.wbeg + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * T + grp;
  const bool valid = w < d.wend;
  const int wc = valid ? w : d.wbeg;
  const int np = valid ? d.npairs[wc] : 0;
  const unsigned char *pn = d.pn + (size_t)wc * d.NP;
  const int *poff = d.poff + (size_t)wc * d.NP;
  const ObContact *pool = d.pool + (size_t)wc * d.NC;
  ObContact *out = d.contacts + (size_t)wc * d.NC;
  int np_max = np;   // warp-uniform trip count
#pragma unroll
  for (int dd = 16; dd >= 1; dd >>= 1) { const int o = __shfl_xor_sync(FULL, np_max, dd); np_max = o > np_max ? o : np_max; }
  int carry = 0;
  for (int base = 0; base < np_max; base += G) {
    const int p = base + gl;
    const int n = p < np ? (int)pn[p] : 0;
    int x = n;
#pragma unroll
    for (int dd = 1; dd < G; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd, G); if (gl >= dd) x += y; }
    const int dst = carry + x - n;
    const int src = n ? poff[p] : 0;
    for (int k = 0; k < n; k++) if (dst + k < d.NC && src + k < d.NC) out[dst + k] = pool[src + k];
    carry += __shfl_sync(FULL, x, G - 1, G);
  }
  if (gl == 0 && valid) {
    int nc = carry;
    if (nc > d.NC || d.pcount[w] > d.NC) { if (nc > d.NC) nc = d.NC; atomicOr(&d.world[w].status, OB_ERR_CONTACT_OVERFLOW); }
    d.ncontacts[w] = nc;
  }
}

int obk_collide_setup(ObBackend *b, const cudaDeviceProp &prop, char *err, size_t errlen) {
  ObBatchDev &d = b->d;
  b->smem_collide = collide_smem(d.NG, d.NP).total;
  if (b->smem_collide > (size_t)prop.sharedMemPerBlockOptin) {
    snprintf(err, errlen, "world does not fit one CTA's shared memory (collide %zu B, limit %zu B)", b->smem_collide, (size_t)prop.sharedMemPerBlockOptin);
    goto fail;
  }
  CK(ob_func_smem((const void *)k_collide<false, false>, (int)b->smem_collide));
  CK(ob_func_smem((const void *)k_collide<true, false>, (int)b->smem_collide));
  CK(ob_func_smem((const void *)k_collide<true, true>, (int)b->smem_collide));
  // small worlds: OB_TILE_WPC worlds per CTA with a pooled, class-grouped narrowphase (k_collide_tile)
  b->collide_tile = 0;
  if (d.NG <= 8 && !getenv("OB_COLLIDE_NOTILE")) {
    long long cap = (long long)OB_TILE_WPC * d.NC;
    b->tile_stage_cap = (int)(cap < 60000 ? cap : 60000);
    b->smem_collide_tile = (size_t)OB_TILE_WPC * collide_smem(d.NG, d.NP).total + collide_tile_smem(d.NG, d.NP, OB_TILE_WPC, b->tile_stage_cap).total;
    if (b->smem_collide_tile <= (size_t)prop.sharedMemPerBlockOptin && (long long)OB_TILE_WPC * d.NP < 65000) {
      b->collide_tile = 1;
      CK(ob_func_smem((const void *)k_collide_tile<false, false, OB_TILE_WPC>, (int)b->smem_collide_tile));
      CK(ob_func_smem((const void *)k_collide_tile<true, false, OB_TILE_WPC>, (int)b->smem_collide_tile));
      CK(ob_func_smem((const void *)k_collide_tile<true, true, OB_TILE_WPC>, (int)b->smem_collide_tile));
    }
  }
  // decoupled narrowphase (k_broad -> k_narrow -> k_contacts): the default; OB_COLLIDE_FUSED=1 keeps the one-kernel path
  b->collide_split = 0;
  { const char *e = getenv("OB_COLLIDE_FUSED"); if (!(e && atoi(e) != 0) && d.W < (1 << 24) && d.NP < (1 << 24) && b->nchunks <= 1) b->collide_split = 1; }
  if (b->collide_split) {
    CK(dalloc(b, &d.gpose, (size_t)d.W * d.NG));
    CK(dalloc(b, &d.wl, (size_t)OB_NCLS * d.W * d.NP * 2));
    CK(dalloc(b, &d.wlcnt, (size_t)16));
    CK(dalloc(b, &d.pn, (size_t)d.W * d.NP));
    CK(dalloc(b, &d.poff, (size_t)d.W * d.NP));
    CK(dalloc(b, &d.pool, (size_t)d.W * d.NC));
    CK(dalloc(b, &d.pcount, (size_t)d.W));
    CK(ob_func_smem((const void *)k_broad, (int)b->smem_collide));
    b->smem_broad_tile = (size_t)OB_TILE_WPC * collide_smem(d.NG, d.NP).total;
    if (b->collide_tile && b->smem_broad_tile <= (size_t)prop.sharedMemPerBlockOptin) CK(ob_func_smem((const void *)k_broad_tile<OB_TILE_WPC>, (int)b->smem_broad_tile));
    int per = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_narrow<true, true>, 128, 0));
    b->narrow_grid[2] = per * prop.multiProcessorCount;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_narrow<true, false>, 128, 0));
    b->narrow_grid[1] = per * prop.multiProcessorCount;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_narrow<false, false>, 128, 0));
    b->narrow_grid[0] = per * prop.multiProcessorCount;
  }
  return 0;
fail:
  return -1;
}

void obk_collide_launch(ObBackend *b, const ObBatchDev &d, int W, int cap, cudaStream_t st) {
  if (b->collide_split && b->nchunks <= 1) {   // chunks of the batch on streams of their own would share the work lists
    cudaMemsetAsync(d.wlcnt, 0, 16 * sizeof(unsigned), st);
    int ct = d.NG <= 8 ? 32 : (d.NG <= 20 ? 64 : OB_THREADS);
    if (b->collide_tile) {
      const int tiles = (W + OB_TILE_WPC - 1) / OB_TILE_WPC;
      k_broad_tile<OB_TILE_WPC><<<tiles < cap ? tiles : cap, 32 * OB_TILE_WPC, b->smem_broad_tile, st>>>(d);
    } else k_broad<<<W < cap ? W : cap, ct, b->smem_collide, st>>>(d);
    // the queue is as long as the pairs that need a collider call; a grid that fills the machine drains it
    const long long maxitems = (long long)W * d.NP;
    int ngrid = d.any_xf ? b->narrow_grid[2] : (d.nmesh ? b->narrow_grid[1] : b->narrow_grid[0]);
    if ((long long)ngrid * 128 > maxitems + 127) ngrid = (int)((maxitems + 127) / 128);
    if (ngrid < 1) ngrid = 1;
    if (d.any_xf) k_narrow<true, true><<<ngrid, 128, 0, st>>>(d);
    else if (d.nmesh) k_narrow<true, false><<<ngrid, 128, 0, st>>>(d);
    else k_narrow<false, false><<<ngrid, 128, 0, st>>>(d);
    if (d.NP <= 16) k_contacts<8><<<(W + 15) / 16, 128, 0, st>>>(d);          // 4 worlds per warp
    else if (d.NP <= 64) k_contacts<16><<<(W + 7) / 8, 128, 0, st>>>(d);    // 2 worlds per warp
    else k_contacts<32><<<(W + 3) / 4, 128, 0, st>>>(d);
    g_launches += 3;
    return;
  }
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


// dBatchRayCast: thread per (world, ray); the world's geoms are walked in the space's list order (SAP: GeomList then DirtyList,
// what cleanGeoms leaves), so that of two equally near hits the first in list order wins, like a near callback keeping `depth <`
__global__ void __launch_bounds__(128) k_raycast(ObBatchDev d, int nrays, const real *origin, const real *dir, const real *length, int ray_flags,
                                                 uint32_t rcat, uint32_t rcol, ObRayHit *hits) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)d.W * nrays) return;
  const int w = (int)(t / nrays);
  ObWorld &W = d.world[w];
  const int ng = W.ng;
  const ObGeom *geoms = d.geom + (size_t)w * d.NG;
  const ObBodyDyn *bd = d.bdyn + (size_t)w * d.NB;
  const int *glist = d.glist + (size_t)w * d.NG;
  const int rot = W.space_type == OB_SPACE_SAP ? W.sap_ndirty : 0;
  ObPose ray;
  ob_ray_pose(origin + 3 * t, dir + 3 * t, length[t], ray_flags, &ray);
  real rab[6];
  ob_aabb(ray, rab, d.meshes);
  ObRayHit h;
  for (int k = 0; k < 3; k++) { h.pos[k] = 0; h.normal[k] = 0; }
  h.depth = length[t]; h.geom = -1;
  bool have = false;
  int bverr = 0;
  for (int i = 0; i < ng; i++) {
    const int gi = glist[i + rot < ng ? i + rot : i + rot - ng];
    const ObGeom g = geoms[gi];
    if (!(g.flags & OB_GEOM_ENABLED) || (g.flags & OB_GEOM_ZERO_SIZED) || g.type == OB_GEOM_RAY || g.type == OB_GEOM_SPACE) continue;
    ObPose p;
    geom_pose_dev(g, bd, &p);
    ObCg c;
    if (ob_ray_vs_geom(ray, rab, rcat, rcol, p, g.body, g.cat, g.col, d.meshes, &c, &bverr) && (!have || c.depth < h.depth)) {
      have = true;
      for (int k = 0; k < 3; k++) { h.pos[k] = c.pos[k]; h.normal[k] = c.normal[k]; }
      h.depth = c.depth; h.geom = gi;
    }
  }
  if (bverr) atomicOr(&W.status, OB_ERR_BVH_STACK);
  ObRayHit &o = hits[t];   // field by field: the buffer was zeroed, so the padding of the dDOUBLE layout stays zero
  for (int k = 0; k < 3; k++) { o.pos[k] = h.pos[k]; o.normal[k] = h.normal[k]; }
  o.depth = h.depth; o.geom = h.geom;
}
int obk_raycast(ObBackend *b, int nrays, const real *origin3, const real *dir3, const real *length, int ray_flags, uint32_t cat, uint32_t col,
                ObRayHit *hits, char *err, size_t errlen) {
  cudaSetDevice(b->device);
  if (b->large) { snprintf(err, errlen, "dBatchRayCast is not served on the large-world path"); return -1; }
  const size_t n = (size_t)b->d.W * nrays;
  if (n == 0) return 0;
  real *dbuf = 0; ObRayHit *dh = 0;
  cudaError_t e = cudaMalloc((void **)&dbuf, sizeof(real) * 7 * n);
  if (e == cudaSuccess) e = cudaMalloc((void **)&dh, sizeof(ObRayHit) * n);
  real *dor = dbuf, *ddir = dbuf + 3 * n, *dlen = dbuf + 6 * n;
  if (e == cudaSuccess) e = cudaMemcpyAsync(dor, origin3, sizeof(real) * 3 * n, cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ddir, dir3, sizeof(real) * 3 * n, cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dlen, length, sizeof(real) * n, cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dh, 0, sizeof(ObRayHit) * n, b->stream);
  if (e == cudaSuccess) {
    k_raycast<<<(unsigned)((n + 127) / 128), 128, 0, b->stream>>>(b->d, nrays, dor, ddir, dlen, ray_flags, cat, col, dh);
    g_launches++;
    e = cudaMemcpyAsync(hits, dh, sizeof(ObRayHit) * n, cudaMemcpyDeviceToHost, b->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  if (dbuf) cudaFree(dbuf);
  if (dh) cudaFree(dh);
  if (e != cudaSuccess) { snprintf(err, errlen, "k_raycast failed: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

// ob_large_kernels.cuh — grid-wide kernels of the large-world path (ob_large.h explains the design).
// Included by ob_backend_cuda.cu after the definition of ObBackend.
//
// Per step (all launches on the backend's stream, three small device->host reads for launch sizing):
//   k_lw_geom      thread/geom   pose, AABB, sort key                       (collision_kernel.cpp:454-465, sapspace.cpp:441-452)
//   lw_radix_sort  4 passes      stable LSD radix sort of (key, geom)       (RadixSort, sapspace.cpp:600-830)
//   k_lw_gather    thread/geom   sorted boxes
//   k_lw_sweep<0>  thread/geom   count pairs   } BoxPruning sweep + infinite-geom list (sapspace.cpp:478-493, :537-566)
//   lw_scan                      offsets       }
//   k_lw_sweep<1>  thread/geom   fill pairs    }
//   k_lw_narrow    thread/pair   dCollide -> contacts in fixed per-pair slots
//   lw_scan x2                   contact creation index, contact-pair compaction
//   k_lw_cpairs    thread/pair   contact pairs (body1, body2, contacts)
//   k_lw_col_*     rounds        deterministic greedy colouring (hashed priorities)
//   lw_radix_sort  2 passes      pairs by (colour, contacts descending); k_lw_segtab: segment table
//   k_lw_body_pre  thread/body   quickstep.cpp:610-665, :840-846
//   k_lw_assemble  thread/pair   contact.cpp:74-256 + quickstep.cpp:849-857, :117-136, :370-402 -> SoA rows
//   k_lw_sor       iters x colours launches, thread/pair: quickstep.cpp:490-581 over the pair's rows
//   k_lw_body_post thread/body   quickstep.cpp:905-975, util.cpp:255-360
#pragma once
#include "ob_large.h"

#define LW_T 128
static inline unsigned lw_blocks(size_t n, int t = LW_T) { return (unsigned)((n + t - 1) / t); }

// ---- exclusive scan of uint32 (out[n] = total) -------------------------------------------------
#define LW_SCAN_T 256
#define LW_SCAN_E 4
__global__ void __launch_bounds__(LW_SCAN_T) k_lw_scan1(const uint32_t *in, uint32_t *out, uint32_t *bsum, int n, int single) {
  __shared__ uint32_t s_w[LW_SCAN_T / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const size_t base = ((size_t)blockIdx.x * LW_SCAN_T + tid) * LW_SCAN_E;
  uint32_t v[LW_SCAN_E], sum = 0;
#pragma unroll
  for (int e = 0; e < LW_SCAN_E; e++) { v[e] = base + e < (size_t)n ? in[base + e] : 0u; sum += v[e]; }
  uint32_t x = sum;
  for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) s_w[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = lane < LW_SCAN_T / 32 ? s_w[lane] : 0u;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
    if (lane < LW_SCAN_T / 32) s_w[lane] = t;
  }
  __syncthreads();
  uint32_t run = (wid ? s_w[wid - 1] : 0u) + x - sum;
#pragma unroll
  for (int e = 0; e < LW_SCAN_E; e++) { if (base + e < (size_t)n) out[base + e] = run; run += v[e]; }
  if (tid == LW_SCAN_T - 1) {
    bsum[blockIdx.x] = s_w[LW_SCAN_T / 32 - 1];
    if (single) out[n] = s_w[LW_SCAN_T / 32 - 1];
  }
}
__global__ void __launch_bounds__(LW_SCAN_T) k_lw_scan_add(uint32_t *out, const uint32_t *bscan, int n, int nblk) {
  const size_t base = ((size_t)blockIdx.x * LW_SCAN_T + threadIdx.x) * LW_SCAN_E;
  const uint32_t add = bscan[blockIdx.x];
#pragma unroll
  for (int e = 0; e < LW_SCAN_E; e++) if (base + e < (size_t)n) out[base + e] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bscan[nblk];
}
// tmp: scratch of at least 2*(n/1024+2)+... words (levels shrink by 1024x)
static void lw_scan(cudaStream_t st, const uint32_t *in, uint32_t *out, int n, uint32_t *tmp) {
  const int per = LW_SCAN_T * LW_SCAN_E;
  const int nblk = n > 0 ? (n + per - 1) / per : 1;
  k_lw_scan1<<<nblk, LW_SCAN_T, 0, st>>>(in, out, tmp, n, nblk == 1);
  g_launches++;
  if (nblk > 1) {
    uint32_t *bscan = tmp + nblk + 1;
    lw_scan(st, tmp, bscan, nblk, bscan + nblk + 2);
    k_lw_scan_add<<<nblk, LW_SCAN_T, 0, st>>>(out, bscan, n, nblk);
    g_launches++;
  }
}

// ---- stable LSD radix sort, 8-bit digits, (uint32 key, int value) ----------------------------------
#define LW_RS_T 256
#define LW_RS_E 8
__global__ void __launch_bounds__(LW_RS_T) k_lw_rs_hist(const uint32_t *key, int n, int shift, uint32_t *bh, int nblk) {
  __shared__ uint32_t s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * LW_RS_T * LW_RS_E;
  for (int r = 0; r < LW_RS_E; r++) {
    const size_t e = base + (size_t)r * LW_RS_T + threadIdx.x;
    if (e < (size_t)n) atomicAdd(&s_h[(key[e] >> shift) & 255u], 1u);
  }
  __syncthreads();
  bh[(size_t)threadIdx.x * nblk + blockIdx.x] = s_h[threadIdx.x];
}
__global__ void __launch_bounds__(LW_RS_T) k_lw_rs_scatter(const uint32_t *key, const int *val, uint32_t *key2, int *val2, int n,
                                                          int shift, const uint32_t *bhscan, int nblk) {
  __shared__ uint32_t s_base[256], s_run[256];
  __shared__ unsigned short s_wc[LW_RS_T / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  s_base[tid] = bhscan[(size_t)tid * nblk + blockIdx.x];
  s_run[tid] = 0;
  for (int w = 0; w < LW_RS_T / 32; w++) s_wc[w][tid] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * LW_RS_T * LW_RS_E;
  for (int r = 0; r < LW_RS_E; r++) {
    const size_t e = base + (size_t)r * LW_RS_T + tid;
    const bool act = e < (size_t)n;
    uint32_t k = 0; int v = 0;
    if (act) { k = key[e]; v = val[e]; }
    const unsigned d = act ? ((k >> shift) & 255u) : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank_w = __popc(peers & ((1u << lane) - 1u));
    if (act && rank_w == 0) s_wc[wid][d] = (unsigned short)__popc(peers);
    __syncthreads();
    if (act) {
      uint32_t pre = s_run[d];
      for (int w = 0; w < wid; w++) pre += s_wc[w][d];
      const size_t dst = (size_t)s_base[d] + pre + rank_w;
      key2[dst] = k; val2[dst] = v;
    }
    __syncthreads();
    {
      uint32_t t = 0;
      for (int w = 0; w < LW_RS_T / 32; w++) { t += s_wc[w][tid]; s_wc[w][tid] = 0; }
      s_run[tid] += t;
    }
    __syncthreads();
  }
}
// sorts (key[0], val[0]) using the [1] buffers; returns the index (0/1) of the buffers holding the result
static int lw_radix_sort(cudaStream_t st, uint32_t *key[2], int *val[2], int n, int npasses, uint32_t *tmp) {
  const int per = LW_RS_T * LW_RS_E;
  const int nblk = n > 0 ? (n + per - 1) / per : 1;
  int cur = 0;
  for (int p = 0; p < npasses; p++) {
    uint32_t *bh = tmp, *bhs = tmp + 256 * (size_t)nblk + 1;
    k_lw_rs_hist<<<nblk, LW_RS_T, 0, st>>>(key[cur], n, 8 * p, bh, nblk);
    lw_scan(st, bh, bhs, 256 * nblk, bhs + 256 * (size_t)nblk + 2);
    k_lw_rs_scatter<<<nblk, LW_RS_T, 0, st>>>(key[cur], val[cur], key[cur ^ 1], val[cur ^ 1], n, 8 * p, bhs, nblk);
    g_launches += 2;
    cur ^= 1;
  }
  return cur;
}

// ---- geoms ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LW_T) k_lw_geom(ObBatchDev d, ObLargeDev L) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const ObWorld &W = d.world[0];
  if (g >= W.ng) return;
  const ObGeom G = d.geom[g];
  ObPose p;
  geom_pose_dev(G, d.bdyn, &p);
  L.pose[g] = p;
  real ab[6];
  ob_aabb(p, ab, d.meshes);
  for (int k = 0; k < 6; k++) L.aabb[(size_t)g * 6 + k] = ab[k];
  int ax0, ax1, ax2;
  ob_sap_axes(W.sap_axes, &ax0, &ax1, &ax2);
  float minf;
  const int en = (G.flags & OB_GEOM_ENABLED) && !(G.flags & OB_GEOM_ZERO_SIZED);
  L.gkey[0][g] = ob_lw_geomkey(ab, en, ax0, &minf);
  L.gidx[0][g] = g;
}
__global__ void __launch_bounds__(LW_T) k_lw_gather(ObBatchDev d, ObLargeDev L, const uint32_t *skey, const int *sidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const ObWorld &W = d.world[0];
  const int ng = W.ng;
  if (i >= ng) return;
  const uint32_t key = skey[i];
  const int g = sidx[i];
  int ax0, ax1, ax2;
  ob_sap_axes(W.sap_axes, &ax0, &ax1, &ax2);
  const real *ab = L.aabb + (size_t)g * 6;
  const ObGeom &G = d.geom[g];
  ObLwBox b;
  b.maxx = ab[ax0 + 1]; b.miny = ab[ax1]; b.maxy = ab[ax1 + 1]; b.minz = ab[ax2]; b.maxz = ab[ax2 + 1];
  b.minx = (float)ab[ax0]; b.body = G.body; b.cat = G.cat; b.col = G.col; b.geom = g; b.pad[0] = b.pad[1] = 0;
  L.sminx[i] = b.minx; L.smaxx[i] = b.maxx;
  L.syz[(size_t)4 * i] = b.miny; L.syz[(size_t)4 * i + 1] = b.maxy; L.syz[(size_t)4 * i + 2] = b.minz; L.syz[(size_t)4 * i + 3] = b.maxz;
  L.smeta[i] = make_int4(G.body, (int)G.cat, (int)G.col, g);
  const uint32_t nxt = i + 1 < ng ? skey[i + 1] : OB_LW_KEY_OFF;
  if (key < OB_LW_KEY_BIG && (i + 1 == ng || nxt >= OB_LW_KEY_BIG)) L.scal[LW_NFIN] = i + 1;
  if (key == OB_LW_KEY_BIG && (i + 1 == ng || nxt != OB_LW_KEY_BIG)) L.scal[LW_NBIG] = i + 1;   // end of the infinite list (start = NFIN)
}

// ---- pairs ---------------------------------------------------------------------------------------
// cnt / off layout: [0, ng) sweep hits of sorted position i; [ng, 2ng) hits of i against the infinite
// list; [2ng] infinite x infinite.  Pair orientation: the geom met first is o1 (sapspace.cpp:478-493, :553).
// Front-end split (several GPUs, ObLwSplit): a launch handles the sorted positions [i0, i1) only; k_lw_push_cnt / k_lw_push_pairs
// then copy what it produced into the peers' arrays.  The infinite x infinite block is evaluated by every rank for itself.
template <int FILL>
__global__ void __launch_bounds__(LW_T) k_lw_sweep(ObBatchDev d, ObLargeDev L, int i0, int i1) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = i0 + t;
  const int ng = d.world[0].ng;
#define LW_PUT_PAIR(IDX, A, B) { const size_t x_ = (IDX); if (x_ < (size_t)L.NP) *(int2 *)(L.pairs + 2 * x_) = make_int2((A), (B)); }
#define LW_PUT_CNT(IDX, V) { L.cnt[(IDX)] = (V); }
  const int nfin = L.scal[LW_NFIN];
  const int bigend = L.scal[LW_NBIG] > nfin ? L.scal[LW_NBIG] : nfin;
  const int *sidx = L.gidx[0];   // the 4-pass sort leaves the result in buffer 0
  if (t == 0) {   // infinite x infinite (collideGeomsNoAABBs on the TmpInfGeomList, :479-485)
    uint32_t h = 0;
    const size_t o = FILL ? L.off[2 * ng] : 0;
    for (int a = nfin; a < bigend; a++)
      for (int b = a + 1; b < bigend; b++) {
        const int ga = sidx[a], gb = sidx[b];
        const ObGeom &A = d.geom[ga], &B = d.geom[gb];
        if (ob_pair_filter_noaabb(A.body, B.body, A.cat, A.col, B.cat, B.col)) {
          if (FILL && o + h < (size_t)L.NP) { L.pairs[2 * (o + h)] = ga; L.pairs[2 * (o + h) + 1] = gb; }
          h++;
        }
      }
    if (!FILL) L.cnt[2 * ng] = h;
    if (FILL) {
      uint32_t np = L.off[2 * ng + 1];
      if (np > (uint32_t)L.NP) { np = (uint32_t)L.NP; atomicOr(&d.world[0].status, OB_ERR_PAIR_OVERFLOW); }
      L.scal[LW_NP] = (int)np;
      // the pairs this launch fills: sweep hits of [i0, i1), then their hits against the infinite list
      L.scal[LW_SEG0] = (int)L.off[i0]; L.scal[LW_SEG0 + 1] = (int)(L.off[i1] - L.off[i0]);
      L.scal[LW_SEG0 + 2] = (int)L.off[ng + i0]; L.scal[LW_SEG0 + 3] = (int)(L.off[ng + i1] - L.off[ng + i0]);
    }
  }
  if (i >= i1) return;
  if (i >= nfin) { if (!FILL) { LW_PUT_CNT(i, 0u) LW_PUT_CNT(ng + i, 0u) } return; }
  const real Kmaxx = L.smaxx[i];
  const real Kminy = L.syz[(size_t)4 * i], Kmaxy = L.syz[(size_t)4 * i + 1], Kminz = L.syz[(size_t)4 * i + 2], Kmaxz = L.syz[(size_t)4 * i + 3];
  const int4 Km = L.smeta[i];
  int *hb = L.hits + (size_t)i * OB_LW_HITBUF;
  if (!FILL) {
    // counting pass: BoxPruning's inner loop (:545-560); the first OB_LW_HITBUF hits are remembered
    uint32_t h = 0;
    for (int j = i + 1; j < nfin; j++) {
      if (!((real)L.sminx[j] <= Kmaxx)) break;
#if defined(dSINGLE)
      const float4 yz = *(const float4 *)(L.syz + (size_t)4 * j);
      const real Jminy = yz.x, Jmaxy = yz.y, Jminz = yz.z, Jmaxz = yz.w;
#else
      const double2 y2 = *(const double2 *)(L.syz + (size_t)4 * j), z2 = *(const double2 *)(L.syz + (size_t)4 * j + 2);
      const real Jminy = y2.x, Jmaxy = y2.y, Jminz = z2.x, Jmaxz = z2.y;
#endif
      if (!(Kmaxy >= Jminy && Jmaxy >= Kminy)) continue;
      if (!(Kmaxz >= Jminz && Jmaxz >= Kminz)) continue;
      const int4 Jm = L.smeta[j];
      if (!ob_pair_filter_noaabb(Km.x, Jm.x, (uint32_t)Km.y, (uint32_t)Km.z, (uint32_t)Jm.y, (uint32_t)Jm.z)) continue;
      if (h < OB_LW_HITBUF) hb[h] = Jm.w;
      h++;
    }
    LW_PUT_CNT(i, h)
  } else {
    const uint32_t n = L.cnt[i];
    const size_t o = L.off[i];
    if (n <= OB_LW_HITBUF) {
      for (uint32_t h = 0; h < n; h++) LW_PUT_PAIR(o + h, Km.w, hb[h])
    } else {
      uint32_t h = 0;
      for (int j = i + 1; j < nfin; j++) {
        if (!((real)L.sminx[j] <= Kmaxx)) break;
        const real Jminy = L.syz[(size_t)4 * j], Jmaxy = L.syz[(size_t)4 * j + 1], Jminz = L.syz[(size_t)4 * j + 2], Jmaxz = L.syz[(size_t)4 * j + 3];
        if (!(Kmaxy >= Jminy && Jmaxy >= Kminy)) continue;
        if (!(Kmaxz >= Jminz && Jmaxz >= Kminz)) continue;
        const int4 Jm = L.smeta[j];
        if (!ob_pair_filter_noaabb(Km.x, Jm.x, (uint32_t)Km.y, (uint32_t)Km.z, (uint32_t)Jm.y, (uint32_t)Jm.z)) continue;
        LW_PUT_PAIR(o + h, Km.w, Jm.w)
        h++;
      }
    }
  }
  uint32_t h = 0;
  const size_t o = FILL ? L.off[ng + i] : 0;
  for (int a = nfin; a < bigend; a++) {   // collideGeomsNoAABBs: no AABB test against the infinite list (:486-491)
    const int ga = sidx[a];
    const ObGeom &A = d.geom[ga];
    if (ob_pair_filter_noaabb(A.body, Km.x, A.cat, A.col, (uint32_t)Km.y, (uint32_t)Km.z)) {
      if (FILL) LW_PUT_PAIR(o + h, ga, Km.w)
      h++;
    }
  }
  if (!FILL) LW_PUT_CNT(ng + i, h)
#undef LW_PUT_PAIR
#undef LW_PUT_CNT
}

// the pairs one launch collides: up to three ranges of the pair list (one GPU: [0, np); front-end split: the two ranges this
// rank filled, written to every rank, and the infinite x infinite range, which every rank keeps to itself)
struct ObLwSeg3 { int start[3], len[3]; };
__device__ __forceinline__ int lw_seg_pair(const ObLwSeg3 &G, int t) {   // t-th pair of the ranges, -1 behind the end
  int p = -1;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (p < 0 && t >= 0 && t < G.len[k]) p = G.start[k] + t;
    t -= G.len[k];
  }
  return p;
}
template <bool MESH, bool XF>
__global__ void __launch_bounds__(LW_T) k_lw_narrow(ObBatchDev d, ObLargeDev L, ObLwSeg3 G, int np, int maxc) {
  const int p = lw_seg_pair(G, blockIdx.x * blockDim.x + threadIdx.x);
  if (p < 0 || p >= np) return;
  const int o1 = L.pairs[2 * p], o2 = L.pairs[2 * p + 1];
  ObCg cg[OB_LW_MAXC];
  int swapped, bverr = 0;
  const int n = ob_collide_pair_sel_t<MESH, OB_LW_MAXC, XF>(&L.pose[o1], &L.pose[o2], maxc, cg, &swapped, d.meshes, &bverr);
  if (bverr) atomicOr(&d.world[0].status, OB_ERR_BVH_STACK);
  ObContact *out = L.pc + (size_t)p * maxc;
  for (int k = 0; k < n; k++) {
    ObContact c;
    for (int e = 0; e < 3; e++) { c.pos[e] = cg[k].pos[e]; c.normal[e] = cg[k].normal[e]; }
    c.depth = cg[k].depth; c.g1 = o1; c.g2 = o2; c.side1 = cg[k].side1; c.side2 = cg[k].side2; c.policy = 0;
    out[k] = c;
  }
  L.ncp[p] = (uint32_t)n;
  const int b1 = d.geom[o1].body, b2 = d.geom[o2].body;
  L.cpflag[p] = (n > 0 && (b1 >= 0 || b2 >= 0)) ? 1u : 0u;
}
// ---- front-end split: what a rank computed goes to its peers as bulk copies -- consecutive threads store consecutive 4 / 8 / 16-byte
// words through the NVLink peer mappings, so the links carry full lines (r02l: storing every count / pair / contact from inside the
// sweep and narrowphase kernels cost more than the split saved).  The flag barrier behind a push makes the copies visible.
__global__ void __launch_bounds__(LW_T) k_lw_push_cnt(ObLargeDev L, ObLwSplit S, int ng, int i0, int i1) {
  const int n = i1 - i0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * n; t += gridDim.x * blockDim.x) {
    const int idx = t < n ? i0 + t : ng + i0 + (t - n);
    const uint32_t v = L.cnt[idx];
    for (int r = 0; r < S.nranks; r++) if (r != S.rank) S.cnt[r][idx] = v;
  }
}
// thread (pair, j): j == 0 copies the pair's geom ids, contact count and solver flag; j >= 1 the (j - 1)-th 16-byte chunk of its contact slots
__global__ void __launch_bounds__(LW_T) k_lw_push_pairs(ObLargeDev L, ObLwSplit S, ObLwSeg3 G, int np, int maxc) {
  static_assert(sizeof(ObContact) % 16 == 0, "contacts are copied in 16-byte chunks");
  constexpr int CC = (int)sizeof(ObContact) / 16;
  const int per = 1 + CC * maxc;
  const long long total = (long long)(G.len[0] + G.len[1]) * per;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t / per), j = (int)(t - (long long)q * per);
    const int p = lw_seg_pair(G, q);
    if (p < 0 || p >= np) continue;
    if (j == 0) {
      const int2 o = *(const int2 *)(L.pairs + 2 * (size_t)p);
      const uint32_t n = L.ncp[p], f = L.cpflag[p];
      for (int r = 0; r < S.nranks; r++) if (r != S.rank) { *(int2 *)(S.pairs[r] + 2 * (size_t)p) = o; S.ncp[r][p] = n; S.cpflag[r][p] = f; }
    } else {
      const int k = (j - 1) / CC;
      if ((uint32_t)k >= L.ncp[p]) continue;
      const size_t off = ((size_t)p * maxc) * CC + (size_t)(j - 1);   // in 16-byte chunks from the start of pc
      const float4 v = ((const float4 *)L.pc)[off];
      for (int r = 0; r < S.nranks; r++) if (r != S.rank) ((float4 *)S.pc[r])[off] = v;
    }
  }
}
// Barrier of the ranks between two kernels of the step (front-end split): one thread.  The kernels before it have
// completed, so their peer stores are performed; this rank raises its phase word in every peer's flag array and waits
// until every peer has raised its own here.  The kernels behind it on the stream then read what the peers wrote into
// this GPU's memory.  A peer that never arrives trips the timeout flag (the host reports it) instead of hanging the GPU.
__device__ __forceinline__ unsigned long long lw_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void k_lw_xbarrier(ObLwSplit S, unsigned phase) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  volatile unsigned *mine = (volatile unsigned *)S.flags[S.rank];
  __threadfence_system();
  for (int r = 0; r < S.nranks; r++) if (r != S.rank) *((volatile unsigned *)S.flags[r] + S.rank) = phase;
  const unsigned long long t0 = lw_globaltimer();
  for (int r = 0; r < S.nranks; r++) {
    if (r == S.rank) continue;
    while ((int)(mine[r] - phase) < 0 && !mine[OB_LW_FLAG_TIMEOUT]) {
      if (lw_globaltimer() - t0 > (unsigned long long)S.timeout_ms * 1000000ull) mine[OB_LW_FLAG_TIMEOUT] = 1;
    }
  }
  __threadfence_system();
}
__global__ void __launch_bounds__(LW_T) k_lw_cpairs(ObBatchDev d, ObLargeDev L, int np, int maxc, int taps) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) {
    L.scal[LW_NCONTACTS] = (int)L.coff[np];
    L.scal[LW_NCP] = (int)L.cpoff[np];
    d.npairs[0] = np;
    d.ncontacts[0] = (int)L.coff[np] < d.NC ? (int)L.coff[np] : d.NC;
  }
  if (p >= np) return;
  const int n = (int)L.ncp[p];
  if (L.cpflag[p]) {
    const int o1 = L.pairs[2 * p], o2 = L.pairs[2 * p + 1];
    int b1 = d.geom[o1].body, b2 = d.geom[o2].body, rev = 0;
    if (b1 < 0) { b1 = b2; b2 = -1; rev = 1; }   // dJointAttach swap rule (ode.cpp:1368-1377)
    ObLwPair P;
    P.b1 = b1; P.b2 = b2; P.info = n | (rev << 8) | (255 << 16); P.src = p;
    L.cp[0][L.cpoff[p]] = P;
  }
  if (taps) {
    const size_t o = L.coff[p];
    for (int k = 0; k < n; k++) if (o + k < (size_t)d.NC) d.contacts[o + k] = L.pc[(size_t)p * maxc + k];
  }
}

// ---- colouring ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(LW_T) k_lw_col_claim(ObLargeDev L, int ncp, uint32_t round) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ncp) return;
  const ObLwPair P = L.cp[0][p];
  if ((P.info >> 16) != 255) return;
  const unsigned long long pr = ob_lw_prio((uint32_t)p, round);
  atomicMin(&L.claim[P.b1], pr);
  if (P.b2 >= 0) atomicMin(&L.claim[P.b2], pr);
}
__global__ void __launch_bounds__(LW_T) k_lw_col_take(ObLargeDev L, int ncp, uint32_t round) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  bool left = false;
  if (p < ncp) {
    ObLwPair P = L.cp[0][p];
    if ((P.info >> 16) == 255) {
      const unsigned long long pr = ob_lw_prio((uint32_t)p, round);
      if (L.claim[P.b1] == pr && (P.b2 < 0 || L.claim[P.b2] == pr)) {
        unsigned long long u = L.used[P.b1];
        if (P.b2 >= 0) u |= L.used[P.b2];
        int c = ob_lw_first_free(u);
        if (c >= OB_LW_MAXCOL) { c = OB_LW_MAXCOL - 1; atomicOr(&L.scal[LW_ERR], 1); }
        L.used[P.b1] |= 1ull << c;
        if (P.b2 >= 0) L.used[P.b2] |= 1ull << c;
        L.cp[0][p].info = (P.info & 0xffff) | (c << 16);
      } else left = true;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, left);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&L.scal[LW_LEFT0 + (round & 15u)], __popc(m));
}
__global__ void __launch_bounds__(LW_T) k_lw_pairkey(ObLargeDev L, int ncp) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ncp) return;
  const int info = L.cp[0][p].info;
  L.pkey[0][p] = (uint32_t)(((info >> 16) & 255) * 8 + (8 - (info & 255)));
  L.pidx[0][p] = p;
}
__global__ void __launch_bounds__(LW_T) k_lw_pairgather(ObLargeDev L, int ncp, const int *sidx) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ncp) return;
  L.cp[1][p] = L.cp[0][sidx[p]];
}
// segment table from the sorted keys: one block of 512 threads
__global__ void __launch_bounds__(512) k_lw_segtab(ObLargeDev L, int ncp, const uint32_t *skey) {
  __shared__ int s_lb[OB_LW_MAXCOL * 8 + 1];
  const int t = threadIdx.x;
  {   // lower_bound(t) in skey[0..ncp)
    int lo = 0, hi = ncp;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (skey[mid] < (uint32_t)t) lo = mid + 1; else hi = mid; }
    s_lb[t] = lo;
    if (t == 0) s_lb[OB_LW_MAXCOL * 8] = ncp;
  }
  __syncthreads();
  if (t == 0) {
    int cbase = 0, ncol = 0;
    for (int c = 0; c < OB_LW_MAXCOL; c++) {
      int *row = L.segtab + c * (2 + OB_LW_MAXC);
      const int start = s_lb[c * 8], count = s_lb[c * 8 + 8] - start;
      row[0] = start; row[1] = count;
      for (int k = 0; k < OB_LW_MAXC; k++) {
        row[2 + k] = cbase;
        cbase += s_lb[c * 8 + (8 - k)] - start;   // pairs of this colour with more than k contacts
      }
      if (count > 0) ncol = c + 1;
    }
    L.scal[LW_NCOL] = ncol;
    L.scal[LW_NSOLVED] = cbase;
  }
}

// ---- solver ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LW_T) k_lw_body_pre(ObBatchDev d, ObLargeDev L, real h) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const ObWorld &W = d.world[0];
  if (b >= W.nb) return;
  ObBodyDyn &B = d.bdyn[b];
  const ObBodyConst &C = d.bconst[b];
  const real stepsize1 = ob_recip(h);
  real R[12], I[12], invI[12], iw[12], avel[3], lvel[3], facc[3], tacc[3], t1[6];
  for (int k = 0; k < 12; k++) { R[k] = B.R[k]; I[k] = C.I[k]; invI[k] = C.invI[k]; }
  for (int k = 0; k < 3; k++) { avel[k] = B.avel[k]; lvel[k] = B.lvel[k]; facc[k] = B.facc[k]; tacc[k] = B.tacc[k]; }
  ob_body_preamble(R, I, invI, avel, B.flags, C.mass, W.gravity, iw, facc, tacc);
  for (int k = 0; k < 3; k++) { B.facc[k] = facc[k]; B.tacc[k] = tacc[k]; }
  for (int k = 0; k < 12; k++) d.invIw[(size_t)12 * b + k] = iw[k];
  ob_body_tmp1(facc, tacc, lvel, avel, C.invMass, iw, stepsize1, t1);
  for (int k = 0; k < 6; k++) d.tmp1[(size_t)8 * b + k] = t1[k];
  for (int k = 0; k < 8; k++) L.fc[(size_t)8 * b + k] = 0;
  L.invM[b] = C.invMass;
  L.hasrow[b] = 0;
}

__device__ __forceinline__ void lw_store_row(const ObLargeDev &L, int q, size_t cs, const real *rw, unsigned meta) {
#if defined(dSINGLE)
#pragma unroll
  for (int s = 0; s < 4; s++) *(float4 *)(L.rows + ob_lw_row_index(q, s, L.NC, cs)) = make_float4(rw[4 * s], rw[4 * s + 1], rw[4 * s + 2], rw[4 * s + 3]);
  *(float4 *)(L.rows + ob_lw_row_index(q, 4, L.NC, cs)) = make_float4(rw[16], rw[17], rw[18], __uint_as_float(meta));
#else
#pragma unroll
  for (int s = 0; s < 9; s++) *(double2 *)(L.rows + ob_lw_row_index(q, s, L.NC, cs)) = make_double2(rw[2 * s], rw[2 * s + 1]);
  *(double2 *)(L.rows + ob_lw_row_index(q, 9, L.NC, cs)) = make_double2(rw[18], __hiloint2double(0, (int)meta));
#endif
}
__device__ __forceinline__ void lw_load_row(const ObLargeDev &L, int q, size_t cs, real *rw, unsigned *meta) {
#if defined(dSINGLE)
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const float4 t = __ldg((const float4 *)(L.rows + ob_lw_row_index(q, s, L.NC, cs)));
    rw[4 * s] = t.x; rw[4 * s + 1] = t.y; rw[4 * s + 2] = t.z; rw[4 * s + 3] = t.w;
  }
  const float4 t = __ldg((const float4 *)(L.rows + ob_lw_row_index(q, 4, L.NC, cs)));
  rw[16] = t.x; rw[17] = t.y; rw[18] = t.z; *meta = __float_as_uint(t.w);
#else
#pragma unroll
  for (int s = 0; s < 9; s++) {
    const double2 t = __ldg((const double2 *)(L.rows + ob_lw_row_index(q, s, L.NC, cs)));
    rw[2 * s] = t.x; rw[2 * s + 1] = t.y;
  }
  const double2 t = __ldg((const double2 *)(L.rows + ob_lw_row_index(q, 9, L.NC, cs)));
  rw[18] = t.x; *meta = (unsigned)__double2loint(t.y);
#endif
}

__global__ void __launch_bounds__(LW_T) k_lw_assemble(ObBatchDev d, ObLargeDev L, int ncp, int maxc, int m, real h) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ncp) return;
  const ObWorld &W = d.world[0];
  const ObLwPair P = L.cp[1][p];
  const int nc = P.info & 255, rev = (P.info >> 8) & 1, col = (P.info >> 16) & 255;
  const int *seg = L.segtab + col * (2 + OB_LW_MAXC);
  const int i = p - seg[0];
  const real stepsize1 = ob_recip(h);
  ObSurface surf = d.policy[0].surface;
  ob_contact_info1(surf);
  const int b1 = P.b1, b2 = P.b2;
  real p1[3], l1[3], a1[3], t1a[6], iw1[12], p2[3] = {0, 0, 0}, l2[3] = {0, 0, 0}, a2[3] = {0, 0, 0}, t1b[6] = {0, 0, 0, 0, 0, 0}, iw2[12];
  for (int e = 0; e < 3; e++) { p1[e] = d.bdyn[b1].pos[e]; l1[e] = d.bdyn[b1].lvel[e]; a1[e] = d.bdyn[b1].avel[e]; }
  for (int e = 0; e < 6; e++) t1a[e] = d.tmp1[(size_t)8 * b1 + e];
  for (int e = 0; e < 12; e++) iw1[e] = d.invIw[(size_t)12 * b1 + e];
  real k2 = 0;
  for (int e = 0; e < 12; e++) iw2[e] = 0;
  if (b2 >= 0) {
    for (int e = 0; e < 3; e++) { p2[e] = d.bdyn[b2].pos[e]; l2[e] = d.bdyn[b2].lvel[e]; a2[e] = d.bdyn[b2].avel[e]; }
    for (int e = 0; e < 6; e++) t1b[e] = d.tmp1[(size_t)8 * b2 + e];
    for (int e = 0; e < 12; e++) iw2[e] = d.invIw[(size_t)12 * b2 + e];
    k2 = L.invM[b2];
    L.hasrow[b2] = 1;
  }
  L.hasrow[b1] = 1;
  const real k1 = L.invM[b1];
  for (int k = 0; k < nc; k++) {
    const size_t cs = (size_t)seg[2 + k] + i;
    if (cs >= (size_t)L.NC) { atomicOr(&d.world[0].status, OB_ERR_CONTACT_OVERFLOW); break; }
    const ObContact c = L.pc[(size_t)P.src * maxc + k];
    real rw[3][OB_LW_ROWW];
    unsigned meta[3];
    if (!ob_lw_contact_rows(c, rev, surf, m, W, p1, l1, a1, t1a, iw1, k1, b2 >= 0, p2, l2, a2, t1b, iw2, k2, stepsize1, rw, meta))
      atomicOr(&d.world[0].status, OB_ERR_ROW_OVERFLOW);
    for (int q = 0; q < m; q++) {
      lw_store_row(L, q, cs, rw[q], meta[q]);
      L.lambda[(size_t)q * L.NC + cs] = 0;
    }
  }
}

// one contact pair: its contacts in order, the rows of a contact in order.  fc and lambda go through L2
// (ld.cg / st.cg): inside the persistent kernel other SMs wrote them in an earlier colour.
template <int M, bool SPLIT = false>
__device__ __forceinline__ void lw_sor_pair(const ObLargeDev &L, const int *seg, int i, const ObLwSplit *S = 0) {
  const ObLwPair P = L.cp[1][seg[0] + i];
  const int nc = P.info & 255, b1 = P.b1, b2 = P.b2;
  real f1[6], f2[6] = {0, 0, 0, 0, 0, 0};
  real *fp1 = L.fc + (size_t)8 * b1, *fp2 = L.fc + (size_t)8 * (b2 >= 0 ? b2 : b1);
  // contacts in order; the rows of contact k+1 are loaded (registers) while contact k is applied
  real rwA[M][OB_LW_ROWW], rwB[M][OB_LW_ROWW], lamA[M], lamB[M];
  unsigned metaA[M], metaB[M];
#define LW_LOAD(RW, META, LAM, KK)                                                                                   \
  {                                                                                                                  \
    const size_t cs_ = (size_t)seg[2 + (KK)] + i;                                                                    \
    _Pragma("unroll") for (int q = 0; q < M; q++) { lw_load_row(L, q, cs_, RW[q], &META[q]); LAM[q] = __ldcg(L.lambda + (size_t)q * L.NC + cs_); } \
  }
  LW_LOAD(rwA, metaA, lamA, 0)   // every contact pair has at least one contact: no need to wait for P
#if defined(dSINGLE)
  { const float4 a = __ldcg((const float4 *)fp1); const float2 c = __ldcg((const float2 *)(fp1 + 4)); f1[0] = a.x; f1[1] = a.y; f1[2] = a.z; f1[3] = a.w; f1[4] = c.x; f1[5] = c.y; }
#else
  for (int e = 0; e < 6; e++) f1[e] = __ldcg(fp1 + e);
#endif
  const real k1 = L.invM[b1];
  real k2 = 0;
  if (b2 >= 0) {
#if defined(dSINGLE)
    const float4 a = __ldcg((const float4 *)fp2); const float2 c = __ldcg((const float2 *)(fp2 + 4));
    f2[0] = a.x; f2[1] = a.y; f2[2] = a.z; f2[3] = a.w; f2[4] = c.x; f2[5] = c.y;
#else
    for (int e = 0; e < 6; e++) f2[e] = __ldcg(fp2 + e);
#endif
    k2 = L.invM[b2];
  }
#define LW_APPLY(RW, META, LAM, KK)                                                                                  \
  {                                                                                                                  \
    const size_t cs_ = (size_t)seg[2 + (KK)] + i;                                                                    \
    _Pragma("unroll") for (int q = 0; q < M; q++) {                                                                  \
      const int fio = (META[q] >> 16) & 255;                                                                         \
      real lam_f = 0;                                                                                                \
      _Pragma("unroll") for (int r = 0; r < M; r++) if (fio && r == q - fio) lam_f = LAM[r];                         \
      LAM[q] = ob_lw_row_update(RW[q], META[q], k1, k2, b2 >= 0, lam_f, LAM[q], f1, f2);                             \
      __stcg(L.lambda + (size_t)q * L.NC + cs_, LAM[q]);                                                             \
    }                                                                                                                \
  }
  for (int k = 0; k < nc; k += 2) {
    if (k + 1 < nc) LW_LOAD(rwB, metaB, lamB, k + 1)
    LW_APPLY(rwA, metaA, lamA, k)
    if (k + 1 < nc) {
      if (k + 2 < nc) LW_LOAD(rwA, metaA, lamA, k + 2)
      LW_APPLY(rwB, metaB, lamB, k + 1)
    }
  }
#undef LW_LOAD
#undef LW_APPLY
  if (SPLIT) {
    // the split sweep: this rank's result goes into every rank's fc array (own copy included); the
    // peer copies are plain stores through the NVLink mapping, made visible by the flag barrier
    for (int r = 0; r < S->nranks; r++) {
      real *q1 = S->fc[r] + (size_t)8 * b1, *q2 = S->fc[r] + (size_t)8 * (b2 >= 0 ? b2 : b1);
#if defined(dSINGLE)
      __stcg((float4 *)q1, make_float4(f1[0], f1[1], f1[2], f1[3])); __stcg((float2 *)(q1 + 4), make_float2(f1[4], f1[5]));
      if (b2 >= 0) { __stcg((float4 *)q2, make_float4(f2[0], f2[1], f2[2], f2[3])); __stcg((float2 *)(q2 + 4), make_float2(f2[4], f2[5])); }
#else
      for (int e = 0; e < 6; e++) __stcg(q1 + e, f1[e]);
      if (b2 >= 0) for (int e = 0; e < 6; e++) __stcg(q2 + e, f2[e]);
#endif
    }
    return;
  }
#if defined(dSINGLE)
  __stcg((float4 *)fp1, make_float4(f1[0], f1[1], f1[2], f1[3])); __stcg((float2 *)(fp1 + 4), make_float2(f1[4], f1[5]));
  if (b2 >= 0) { __stcg((float4 *)fp2, make_float4(f2[0], f2[1], f2[2], f2[3])); __stcg((float2 *)(fp2 + 4), make_float2(f2[4], f2[5])); }
#else
  for (int e = 0; e < 6; e++) __stcg(fp1 + e, f1[e]);
  if (b2 >= 0) for (int e = 0; e < 6; e++) __stcg(fp2 + e, f2[e]);
#endif
}
// one launch per (iteration, colour): debugging path (OB_LW_SOR_LAUNCHES=1)
template <int M>
__global__ void __launch_bounds__(LW_T) k_lw_sor(ObLargeDev L, int col) {
  const int *seg = L.segtab + col * (2 + OB_LW_MAXC);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= seg[1]) return;
  lw_sor_pair<M>(L, seg, i);
}
// the whole SOR phase in one cooperative launch: every CTA walks (iteration, colour) and the grid meets
// at a barrier after every colour (arrive counter in global memory, monotonically increasing)
#define LW_SOR_T 256
__device__ __forceinline__ void lw_grid_barrier(unsigned *bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (*(volatile unsigned *)bar < target) { }
    __threadfence();
  }
  __syncthreads();
}
// pull the rows a pair will read into L2 (no registers held): issued for the NEXT colour before the
// grid barrier, so that after the barrier the sweep's row loads are L2 hits
template <int M>
__device__ __forceinline__ void lw_prefetch_pair(const ObLargeDev &L, const int *seg, int i) {
  const int nc = L.cp[1][seg[0] + i].info & 255;
  for (int k = 0; k < nc; k++) {
    const size_t cs = (size_t)seg[2 + k] + i;
#pragma unroll
    for (int q = 0; q < M; q++) {
#pragma unroll
      for (int s2 = 0; s2 < OB_LW_SLOTS; s2++) asm volatile("prefetch.global.L2 [%0];" ::"l"(L.rows + ob_lw_row_index(q, s2, L.NC, cs)));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(L.lambda + (size_t)q * L.NC + cs));
    }
  }
}
template <int M>
__global__ void __launch_bounds__(LW_SOR_T) k_lw_sor_all(ObLargeDev L, int iters, int ncol, unsigned *bar) {
  unsigned epoch = 0;
  const int stride = gridDim.x * blockDim.x;
  // warp tiles of 32 consecutive pairs are dealt round-robin over the CTAs: pairs are sorted by contact
  // count, so consecutive tiles are equally heavy and every SM gets the same share of the bytes
  const int t = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
  for (int it = 0; it < iters; it++)
    for (int c = 0; c < ncol; c++) {
      const int *seg = L.segtab + c * (2 + OB_LW_MAXC);
      const int cnt = seg[1];
      for (int i = t; i < cnt; i += stride) lw_sor_pair<M>(L, seg, i);
      if (c + 1 < ncol || it + 1 < iters) {
        const int *segn = L.segtab + (c + 1 < ncol ? c + 1 : 0) * (2 + OB_LW_MAXC);
        const int cntn = segn[1];
        for (int i = t; i < cntn; i += stride) lw_prefetch_pair<M>(L, segn, i);
      }
      epoch++;
      lw_grid_barrier(bar, epoch * gridDim.x);
    }
}

// ---- the SOR phase split over the GPUs of one box (ObLwSplit, ob_large.h) -------------------------
// Barrier of all CTAs of all ranks.  Every CTA arrives at the local counter behind a system-scope fence
// (its peer stores of fc are ordered before the arrival); CTA 0 waits for the local arrivals, raises this
// rank's phase word in every peer's flag array, waits until every peer has raised its own here, and then
// releases the local CTAs.  A wait that outlasts timeout_ms (a rank that never launched) raises the
// timeout flag, after which no barrier waits any more: the step finishes and the host reports the error.
__device__ __forceinline__ void lw_split_barrier(unsigned *bar, unsigned local_target, const ObLwSplit &S, unsigned phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned *mine = (volatile unsigned *)S.flags[S.rank];
    // the ONE system-scope fence of the barrier: it returns when this CTA's peer stores have been performed at the
    // peers, so everything ordered after it (the arrival, CTA 0's flag stores) reaches a peer after the data
    __threadfence_system();
    atomicAdd(bar, 1u);
    if (blockIdx.x == 0) {
      while (*(volatile unsigned *)bar < local_target) { }
      __threadfence();
      for (int r = 0; r < S.nranks; r++) if (r != S.rank) *((volatile unsigned *)S.flags[r] + S.rank) = phase;
      const unsigned long long t0 = lw_globaltimer();
      for (int r = 0; r < S.nranks; r++) {
        if (r == S.rank) continue;
        while ((int)(mine[r] - phase) < 0 && !mine[OB_LW_FLAG_TIMEOUT]) {
          if (lw_globaltimer() - t0 > (unsigned long long)S.timeout_ms * 1000000ull) mine[OB_LW_FLAG_TIMEOUT] = 1;
        }
      }
      // the peers' data arrived in this GPU's L2 before their flags did; the sweep reads fc with ld.cg (L2)
      __threadfence();
      mine[OB_LW_FLAG_RELEASE] = phase;
    } else {
      while ((int)(mine[OB_LW_FLAG_RELEASE] - phase) < 0) { }
    }
    __threadfence();
  }
  __syncthreads();
}
template <int M>
__global__ void __launch_bounds__(LW_SOR_T) k_lw_sor_split(ObLargeDev L, ObLwSplit S, int iters, int ncol, unsigned *bar) {
  unsigned epoch = 0, phase = S.base;
  // the ranks' grids form one virtual grid with this rank's CTAs at positions rank, rank + nranks, ...:
  // warp tile w of a colour is swept by rank w % nranks (ob_lw_split_owner)
  const int vgrid = gridDim.x * S.nranks, vblock = blockIdx.x * S.nranks + S.rank;
  const int stride = vgrid * blockDim.x;
  const int t = ((threadIdx.x >> 5) * vgrid + vblock) * 32 + (threadIdx.x & 31);
  // entry: no rank may store into a peer's fc before that peer's k_lw_body_pre has initialised it
  epoch++; phase++;
  lw_split_barrier(bar, epoch * gridDim.x, S, phase);
  for (int it = 0; it < iters; it++)
    for (int c = 0; c < ncol; c++) {
      const int *seg = L.segtab + c * (2 + OB_LW_MAXC);
      const int cnt = seg[1];
      for (int i = t; i < cnt; i += stride) lw_sor_pair<M, true>(L, seg, i, &S);
      if (c + 1 < ncol || it + 1 < iters) {
        const int *segn = L.segtab + (c + 1 < ncol ? c + 1 : 0) * (2 + OB_LW_MAXC);
        const int cntn = segn[1];
        for (int i = t; i < cntn; i += stride) lw_prefetch_pair<M>(L, segn, i);
      }
      epoch++; phase++;
      lw_split_barrier(bar, epoch * gridDim.x, S, phase);
    }
}

// parity tap (taps & 1): dJointFeedback of every contact joint, creation order -- f1 / t1 / f2 / t2 = J^T lambda of the contact's rows
// (quickstep.cpp:918-957, Multiply1_12q1 order: rows of the contact in sequence)
__global__ void __launch_bounds__(LW_T) k_lw_feedback(ObBatchDev d, ObLargeDev L, int ncp, int maxc, int m) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ncp) return;
  const ObLwPair P = L.cp[1][p];
  const int nc = P.info & 255, col = (P.info >> 16) & 255;
  const int *seg = L.segtab + col * (2 + OB_LW_MAXC);
  const int i = p - seg[0];
  for (int k = 0; k < nc; k++) {
    const size_t cs = (size_t)seg[2 + k] + i;
    if (cs >= (size_t)L.NC) break;
    const size_t ci = (size_t)L.coff[P.src] + k;
    if (ci >= (size_t)d.NC) continue;
    real acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = 0; q < m; q++) {
      real rw[OB_LW_ROWW];
      unsigned meta;
      lw_load_row(L, q, cs, rw, &meta);
      const real lam = __ldcg(L.lambda + (size_t)q * L.NC + cs);
      for (int e = 0; e < 6; e++) acc[e] += rw[e] * lam;
      for (int e = 0; e < 3; e++) { acc[6 + e] += (-rw[e]) * lam; acc[9 + e] += rw[6 + e] * lam; }
    }
    if (P.b2 < 0) for (int e = 6; e < 12; e++) acc[e] = 0;
    for (int e = 0; e < 12; e++) d.fback[ci * 12 + e] = acc[e];
  }
}

__global__ void __launch_bounds__(LW_T) k_lw_body_post(ObBatchDev d, ObLargeDev L, real h) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const ObWorld &W = d.world[0];
  if (b == 0) {
    d.nrows[0] = 0;
    atomicAdd(&d.counters->steps, 1ull);
    atomicAdd(&d.counters->body_steps, (unsigned long long)W.nb);
    atomicAdd(&d.counters->pairs, (unsigned long long)L.scal[LW_NP]);
    atomicAdd(&d.counters->contacts, (unsigned long long)L.scal[LW_NSOLVED]);
    if (W.status || L.scal[LW_ERR]) atomicAdd(&d.counters->overflow_worlds, 1ull);
  }
  if (b >= W.nb) return;
  ObBodyDyn &B = d.bdyn[b];
  const ObBodyConst &C = d.bconst[b];
  real pos[3], q[4], R[12], lvel[3], avel[3], facc[3], tacc[3], iw[12], fcb[6];
  for (int k = 0; k < 3; k++) { pos[k] = B.pos[k]; lvel[k] = B.lvel[k]; avel[k] = B.avel[k]; facc[k] = B.facc[k]; tacc[k] = B.tacc[k]; }
  for (int k = 0; k < 4; k++) q[k] = B.q[k];
  for (int k = 0; k < 12; k++) iw[k] = d.invIw[(size_t)12 * b + k];
  for (int k = 0; k < 6; k++) fcb[k] = L.fc[(size_t)8 * b + k];
  ob_body_velocity_update(lvel, avel, L.hasrow[b] ? fcb : (real *)0, facc, tacc, C.invMass, iw, h);
  real fra[3] = {C.finite_rot_axis[0], C.finite_rot_axis[1], C.finite_rot_axis[2]};
  ob_step_body(pos, q, R, lvel, avel, B.flags, h, C.max_angular_speed, fra, C.damp_lin_scale, C.damp_ang_scale, C.damp_lin_thr,
               C.damp_ang_thr);
  for (int k = 0; k < 3; k++) { B.pos[k] = pos[k]; B.lvel[k] = lvel[k]; B.avel[k] = avel[k]; }
  for (int k = 0; k < 4; k++) { B.q[k] = q[k]; B.facc[k] = 0; B.tacc[k] = 0; }
  for (int k = 0; k < 12; k++) B.R[k] = R[k];
}

// ---- host side -----------------------------------------------------------------------------------
#define LWCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_)); return -1; } } while (0)
#define LW_HOST_WORDS (LW_WORDS + OB_LW_MAXCOL * (2 + OB_LW_MAXC))

// layout of the peer-visible allocation: byte offsets from its start (= ObLargeDev::fc); a function of the capacities
// only, so every rank derives a peer's pointers from that peer's base address
struct ObLwArena { size_t flags, cnt, pairs, ncp, cpflag, pc, total; };
static ObLwArena lw_arena(size_t NB, size_t NG, size_t NP) {
  ObLwArena A; size_t o = 0;
  auto take = [&o](size_t bytes) { const size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
  take(NB * 8 * sizeof(real));
  A.flags = take(OB_LW_FLAG_WORDS * sizeof(unsigned));
  A.cnt = take((2 * NG + 2) * sizeof(uint32_t));
  A.pairs = take(NP * 2 * sizeof(int));
  A.ncp = take((NP + 1) * sizeof(uint32_t));
  A.cpflag = take((NP + 1) * sizeof(uint32_t));
  A.pc = take(NP * OB_LW_MAXC * sizeof(ObContact));
  A.total = o;
  return A;
}

int lw_create(ObBackend *b, char *err, size_t errlen) {
  ObBatchDev &d = b->d;
  ObLargeDev &L = b->L;
  if (d.W != 1) { snprintf(err, errlen, "the large-world path takes exactly one world"); return -1; }
  L.NG = d.NG; L.NB = d.NB; L.NP = d.NP; L.NC = d.NC;
  const size_t NG = d.NG, NB = d.NB, NP = d.NP, NC = d.NC;
  LWCK(dalloc(b, &d.world, (size_t)1));
  LWCK(dalloc(b, &d.bdyn, NB));
  LWCK(dalloc(b, &d.bconst, NB));
  LWCK(dalloc(b, &d.geom, NG));
  LWCK(dalloc(b, &d.glist, NG));
  LWCK(dalloc(b, &d.policy, (size_t)d.npolicy));
  LWCK(dalloc(b, &d.meshes, (size_t)(d.nmesh ? d.nmesh : 1)));
  LWCK(dalloc(b, &d.njoints, (size_t)1));
  LWCK(dalloc(b, &d.npairs, (size_t)1));
  LWCK(dalloc(b, &d.ncontacts, (size_t)1));
  LWCK(dalloc(b, &d.contacts, NC));
  LWCK(dalloc(b, &d.invIw, NB * 12));
  LWCK(dalloc(b, &d.tmp1, NB * 8));
  LWCK(dalloc(b, &d.nrows, (size_t)1));
  LWCK(dalloc(b, &d.counters, (size_t)1));
  d.joint = 0; d.padjstart = 0; d.padj = 0; d.sapstate = 0; d.rows = 0; d.stepinfo = 0; d.ibody = 0; d.isz = 0; d.jrow = 0;
  d.ijoint = 0; d.jside = 0; d.sched = 0; d.pstart = 0; d.rowJ = d.rowiMJ = d.rowJc = d.rowS = 0; d.rowI = 0; d.lambda = 0;
  d.csurf = 0; d.cfdir1 = 0;
  LWCK(dalloc(b, &d.fback, NC * 12));   // parity tap: joint feedback per contact (creation order), k_lw_feedback
  b->st_elems = NB;
  LWCK(dalloc(b, &b->st_dev, NB * 13));
  LWCK(dalloc(b, &L.pose, NG));
  LWCK(dalloc(b, &L.aabb, NG * 6));
  for (int k = 0; k < 2; k++) { LWCK(dalloc(b, &L.gkey[k], NG)); LWCK(dalloc(b, &L.gidx[k], NG)); }
  L.sbox = 0;
  LWCK(dalloc(b, &L.sminx, NG));
  LWCK(dalloc(b, &L.smaxx, NG));
  LWCK(dalloc(b, &L.syz, NG * 4));
  LWCK(dalloc(b, &L.smeta, NG));
  LWCK(dalloc(b, &L.hits, NG * OB_LW_HITBUF));
  LWCK(dalloc(b, &L.scal, (size_t)LW_WORDS));
  LWCK(dalloc(b, &L.off, 2 * NG + 2));
  LWCK(dalloc(b, &L.coff, NP + 1));
  LWCK(dalloc(b, &L.cpoff, NP + 1));
  for (int k = 0; k < 2; k++) { LWCK(dalloc(b, &L.cp[k], NP)); LWCK(dalloc(b, &L.pkey[k], NP)); LWCK(dalloc(b, &L.pidx[k], NP)); }
  LWCK(dalloc(b, &L.claim, NB));
  LWCK(dalloc(b, &L.used, NB));
  LWCK(dalloc(b, &L.segtab, (size_t)OB_LW_MAXCOL * (2 + OB_LW_MAXC)));
  LWCK(dalloc(b, &L.rows, (size_t)3 * OB_LW_SLOTS * OB_LW_SLOTW * NC));
  LWCK(dalloc(b, &L.lambda, 3 * NC));
  {   // everything a peer rank writes into (ObLwSplit) lives in ONE allocation, so that one IPC handle exports it all
    ObLwArena A = lw_arena(NB, NG, NP);
    unsigned char *raw = 0;
    LWCK(dalloc(b, &raw, A.total));
    L.fc = (real *)raw; b->lw_flags = (unsigned *)(raw + A.flags); b->lw_flags_off = A.flags;
    L.cnt = (uint32_t *)(raw + A.cnt); L.pairs = (int *)(raw + A.pairs); d.pairs = L.pairs;
    L.ncp = (uint32_t *)(raw + A.ncp); L.cpflag = (uint32_t *)(raw + A.cpflag); L.pc = (ObContact *)(raw + A.pc);
  }
  LWCK(dalloc(b, &L.invM, NB));
  LWCK(dalloc(b, &L.hasrow, NB));
  L.tmp_words = (NP > 2 * NG ? NP : 2 * NG) / 2 + 65536;
  LWCK(dalloc(b, &L.tmp, L.tmp_words));
  LWCK(cudaMallocHost((void **)&b->lw_host, sizeof(int) * (LW_HOST_WORDS + 4)));
  for (int k = 0; k < 9; k++) LWCK(cudaEventCreate(&b->lw_ev[k]));
  {   // persistent SOR kernel: as many CTAs as are co-resident
    cudaDeviceProp prop;
    LWCK(cudaGetDeviceProperties(&prop, b->device));
    if (!prop.cooperativeLaunch) { snprintf(err, errlen, "device does not support cooperative launches"); return -1; }
    int per = 0;
    LWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lw_sor_all<1>, LW_SOR_T, 0)); b->lw_sor_grid[0] = per * prop.multiProcessorCount;
    LWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lw_sor_all<2>, LW_SOR_T, 0)); b->lw_sor_grid[1] = per * prop.multiProcessorCount;
    LWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lw_sor_all<3>, LW_SOR_T, 0)); b->lw_sor_grid[2] = per * prop.multiProcessorCount;
    const char *e = getenv("OB_LW_SOR_CTAS_PER_SM");
    if (e && atoi(e) > 0) for (int k = 0; k < 3; k++) if (atoi(e) * prop.multiProcessorCount < b->lw_sor_grid[k]) b->lw_sor_grid[k] = atoi(e) * prop.multiProcessorCount;
    // the split kernel (several GPUs): same sizing; OB_LW_SOR_THREADS narrows its CTAs (two ranks on ONE GPU
    // must be co-resident: the loop-back test uses 128 threads and one CTA per SM for each)
    const char *t = getenv("OB_LW_SOR_THREADS");
    b->lw_split_threads = t && atoi(t) >= 32 && atoi(t) <= LW_SOR_T ? atoi(t) & ~31 : LW_SOR_T;
    LWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lw_sor_split<1>, b->lw_split_threads, 0)); b->lw_split_grid[0] = per * prop.multiProcessorCount;
    LWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lw_sor_split<2>, b->lw_split_threads, 0)); b->lw_split_grid[1] = per * prop.multiProcessorCount;
    LWCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lw_sor_split<3>, b->lw_split_threads, 0)); b->lw_split_grid[2] = per * prop.multiProcessorCount;
    if (e && atoi(e) > 0) for (int k = 0; k < 3; k++) if (atoi(e) * prop.multiProcessorCount < b->lw_split_grid[k]) b->lw_split_grid[k] = atoi(e) * prop.multiProcessorCount;
  }
  return 0;
}

// ---- split over GPUs: export / attach ------------------------------------------------------------
// What a rank publishes about its fc + flag allocation.  Ranks in other processes map it with the IPC
// handle; a rank in the same process (loop-back tests, or one process driving several GPUs) uses the
// pointer itself.  128 bytes, opaque to the caller (dBatchSplitExport / dBatchSplitAttach).
struct ObLwSplitHandle {
  int pid, device;
  unsigned long long ptr;        // L.fc in the exporting process
  unsigned long long alloc_off;  // L.fc - base of the cudaMalloc block the IPC handle names
  unsigned long long flags_off;  // flag words - L.fc
  unsigned long long nb;         // body capacity (must match)
  cudaIpcMemHandle_t ipc;        // 64 bytes
  unsigned long long ng, np;     // geom and pair capacities (must match: they fix the layout behind fc, lw_arena)
  unsigned char pad[128 - 56 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(ObLwSplitHandle) == OBK_SPLIT_HANDLE_BYTES, "split handle is 128 bytes");

int obk_split_export(ObBackend *b, void *handle128, char *err, size_t errlen) {
  if (!b->large) { snprintf(err, errlen, "only the large-world path splits over GPUs"); return -1; }
  cudaSetDevice(b->device);
  ObLwSplitHandle H;
  memset(&H, 0, sizeof H);
  H.pid = (int)getpid(); H.device = b->device; H.ptr = (unsigned long long)(uintptr_t)b->L.fc; H.flags_off = b->lw_flags_off; H.nb = (unsigned long long)b->L.NB; H.ng = (unsigned long long)b->L.NG; H.np = (unsigned long long)b->L.NP;
  // the IPC handle names the whole block cudaMalloc carved the buffer from: find our offset in it
  typedef int (*getrange_t)(unsigned long long *, size_t *, unsigned long long);
  void *fn = 0;
  cudaDriverEntryPointQueryResult qr;
  LWCK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
  unsigned long long base = 0; size_t size = 0;
  if (!fn || ((getrange_t)fn)(&base, &size, H.ptr) != 0) { snprintf(err, errlen, "cuMemGetAddressRange failed"); return -1; }
  H.alloc_off = H.ptr - base;
  LWCK(cudaIpcGetMemHandle(&H.ipc, (void *)(uintptr_t)base));
  memcpy(handle128, &H, sizeof H);
  return 0;
}

int obk_split_attach(ObBackend *b, int rank, int nranks, const void *handles, char *err, size_t errlen) {
  if (!b->large) { snprintf(err, errlen, "only the large-world path splits over GPUs"); return -1; }
  if (nranks < 1 || nranks > OB_LW_MAXRANKS || rank < 0 || rank >= nranks) { snprintf(err, errlen, "bad rank %d of %d (at most %d ranks)", rank, nranks, OB_LW_MAXRANKS); return -1; }
  if (b->lw_split_on) { snprintf(err, errlen, "the batch is already attached to a split"); return -1; }
  cudaSetDevice(b->device);
  const ObLwSplitHandle *H = (const ObLwSplitHandle *)handles;
  ObLwSplit S;
  memset(&S, 0, sizeof S);
  S.rank = rank; S.nranks = nranks; S.base = 0;
  const char *to = getenv("OB_LW_SPLIT_TIMEOUT_MS");
  S.timeout_ms = to && atoi(to) > 0 ? (unsigned)atoi(to) : 5000u;
  for (int r = 0; r < nranks; r++) {
    if (H[r].nb != (unsigned long long)b->L.NB || H[r].flags_off != b->lw_flags_off || H[r].ng != (unsigned long long)b->L.NG || H[r].np != (unsigned long long)b->L.NP) { snprintf(err, errlen, "rank %d holds a world of another size", r); return -1; }
    real *fc = 0;
    if (r == rank) {
      if (H[r].ptr != (unsigned long long)(uintptr_t)b->L.fc || H[r].pid != (int)getpid()) { snprintf(err, errlen, "handle %d is not this batch's own export", r); return -1; }
      fc = b->L.fc;
    } else if (H[r].pid == (int)getpid()) {
      if (H[r].device != b->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(H[r].device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) { snprintf(err, errlen, "no peer access from device %d to %d: %s", b->device, H[r].device, cudaGetErrorString(e)); return -1; }
      }
      fc = (real *)(uintptr_t)H[r].ptr;
    } else {
      void *base = 0;
      LWCK(cudaIpcOpenMemHandle(&base, H[r].ipc, cudaIpcMemLazyEnablePeerAccess));
      b->lw_peer_base[r] = base;
      fc = (real *)((unsigned char *)base + H[r].alloc_off);
    }
    S.fc[r] = fc;
    S.flags[r] = (unsigned *)((unsigned char *)fc + H[r].flags_off);
    const ObLwArena A = lw_arena((size_t)b->L.NB, (size_t)b->L.NG, (size_t)b->L.NP);
    unsigned char *raw = (unsigned char *)fc;
    S.cnt[r] = (uint32_t *)(raw + A.cnt); S.pairs[r] = (int *)(raw + A.pairs); S.ncp[r] = (uint32_t *)(raw + A.ncp);
    S.cpflag[r] = (uint32_t *)(raw + A.cpflag); S.pc[r] = (ObContact *)(raw + A.pc);
  }
  { const char *e = getenv("OB_LW_SPLIT_SOR"); b->lw_split_sor = e && atoi(e) != 0; }
  { const char *e = getenv("OB_LW_SPLIT_FRONT"); b->lw_split_front = e ? atoi(e) != 0 : 1; }
  LWCK(cudaMemsetAsync(b->lw_flags, 0, OB_LW_FLAG_WORDS * sizeof(unsigned), b->stream));
  LWCK(cudaStreamSynchronize(b->stream));
  b->lw_split = S;
  b->lw_split_on = nranks > 1;
  return 0;
}

int lw_step(ObBackend *b, real h, int taps, char *err, size_t errlen) {
  ObBatchDev &d = b->d;
  ObLargeDev &L = b->L;
  cudaStream_t st = b->stream;
  int *hs = b->lw_host;
  // world + policy as the device holds them (small reads; also orders this step after any upload)
  ObWorld hw; ObPolicy hp;
  LWCK(cudaMemcpyAsync(&hw, d.world, sizeof hw, cudaMemcpyDeviceToHost, st));
  LWCK(cudaMemcpyAsync(&hp, d.policy, sizeof hp, cudaMemcpyDeviceToHost, st));
  LWCK(cudaStreamSynchronize(st));
  const int ng = hw.ng, nb = hw.nb;
  if (hw.space_type != OB_SPACE_SAP) { snprintf(err, errlen, "the large-world path implements dSweepAndPruneSpace only"); return -1; }
  const int maxc = hp.max_contacts > OB_LW_MAXC ? OB_LW_MAXC : (hp.max_contacts < 1 ? 1 : hp.max_contacts);
  ObSurface sf = hp.surface;
  const int m = ob_contact_info1(sf);
  const bool tm = b->ktiming != 0;
  int evi = 0;
#define LW_MARK() do { if (tm) cudaEventRecord(b->lw_ev[evi], st); evi++; } while (0)
  LW_MARK();
  // (1) geoms, sort by axis-0 minimum
  LWCK(cudaMemsetAsync(L.scal, 0, sizeof(int) * LW_WORDS, st));
  k_lw_geom<<<lw_blocks(ng), LW_T, 0, st>>>(d, L);
  const int gcur = lw_radix_sort(st, L.gkey, L.gidx, ng, 4, L.tmp);
  if (gcur != 0) { snprintf(err, errlen, "internal: sort parity"); return -1; }
  k_lw_gather<<<lw_blocks(ng), LW_T, 0, st>>>(d, L, L.gkey[0], L.gidx[0]);
  g_launches += 2;
  LW_MARK();
  // (2) pairs: count, scan, fill.  Front-end split: this rank's share of the sorted positions, counts exchanged before the scan
  const bool front = b->lw_split_on && b->lw_split_front;
  ObLwSplit S = b->lw_split;
  int i0 = 0, i1 = ng;
  if (front) ob_lw_split_range(ng, S.rank, S.nranks, &i0, &i1);
  const int nsw = i1 - i0 > 1 ? i1 - i0 : 1;   // thread 0 always runs (infinite x infinite block)
  k_lw_sweep<0><<<lw_blocks(nsw), LW_T, 0, st>>>(d, L, i0, i1);
  if (front) {
    if (i1 > i0) k_lw_push_cnt<<<lw_blocks(2 * (size_t)(i1 - i0)), LW_T, 0, st>>>(L, S, ng, i0, i1);
    k_lw_xbarrier<<<1, 32, 0, st>>>(S, ++b->lw_split.base);
    g_launches += 2;
  }
  lw_scan(st, L.cnt, L.off, 2 * ng + 1, L.tmp);
  k_lw_sweep<1><<<lw_blocks(nsw), LW_T, 0, st>>>(d, L, i0, i1);
  g_launches += 2;
  LWCK(cudaMemcpyAsync(hs, L.scal, sizeof(int) * LW_WORDS, cudaMemcpyDeviceToHost, st));
  if (front) LWCK(cudaMemcpyAsync(hs + LW_HOST_WORDS, L.off + 2 * ng, sizeof(int), cudaMemcpyDeviceToHost, st));
  LWCK(cudaStreamSynchronize(st));
  const int np = hs[LW_NP];
  LW_MARK();
  // (3) narrowphase, contact pairs.  Front-end split: the pairs this rank has just filled (nobody else's are needed yet), then
  // pairs + contacts go to the peers in one bulk copy and the second barrier of the step follows
  {
    ObLwSeg3 G;
    memset(&G, 0, sizeof G);
    if (front) {
      G.start[0] = hs[LW_SEG0]; G.len[0] = hs[LW_SEG0 + 1];
      G.start[1] = hs[LW_SEG0 + 2]; G.len[1] = hs[LW_SEG0 + 3];
      // infinite x infinite: behind the two lists (off[2 ng] = end of the second one); every rank for itself
      G.start[2] = hs[LW_HOST_WORDS];
      G.len[2] = np > G.start[2] ? np - G.start[2] : 0;
    } else { G.start[0] = 0; G.len[0] = np; }
    const int nt = G.len[0] + G.len[1] + G.len[2];
    if (nt > 0) {
      if (d.any_xf) k_lw_narrow<true, true><<<lw_blocks(nt), LW_T, 0, st>>>(d, L, G, np, maxc);
      else if (d.nmesh) k_lw_narrow<true, false><<<lw_blocks(nt), LW_T, 0, st>>>(d, L, G, np, maxc);
      else k_lw_narrow<false, false><<<lw_blocks(nt), LW_T, 0, st>>>(d, L, G, np, maxc);
      g_launches++;
    }
    if (front) {
      const size_t work = (size_t)(G.len[0] + G.len[1]) * (1 + (sizeof(ObContact) / 16) * maxc);
      if (work > 0) { k_lw_push_pairs<<<lw_blocks(work < ((size_t)1 << 24) ? work : ((size_t)1 << 24)), LW_T, 0, st>>>(L, S, G, np, maxc); g_launches++; }
      k_lw_xbarrier<<<1, 32, 0, st>>>(S, ++b->lw_split.base);
      g_launches++;
    }
  }
  lw_scan(st, L.ncp, L.coff, np, L.tmp);
  lw_scan(st, L.cpflag, L.cpoff, np, L.tmp);
  k_lw_cpairs<<<lw_blocks(np > 0 ? np : 1), LW_T, 0, st>>>(d, L, np, maxc, taps);
  g_launches++;
  LWCK(cudaMemcpyAsync(hs, L.scal, sizeof(int) * LW_WORDS, cudaMemcpyDeviceToHost, st));
  LWCK(cudaStreamSynchronize(st));
  const int ncp = hs[LW_NCP];
  LW_MARK();
  // (4) colouring: rounds of claim / take until every pair has a colour
  int ncol = 0, rounds = 0;
  if (ncp > 0) {
    LWCK(cudaMemsetAsync(L.used, 0, sizeof(unsigned long long) * nb, st));
    LWCK(cudaMemsetAsync(L.claim, 0xff, sizeof(unsigned long long) * nb, st));
    int left = ncp;
    while (left > 0) {
      const int chunk = rounds == 0 ? 12 : 4;   // rounds between two looks at the remaining count
      LWCK(cudaMemsetAsync(L.scal + LW_LEFT0, 0, sizeof(int) * 16, st));
      for (int r = 0; r < chunk; r++, rounds++) {
        k_lw_col_claim<<<lw_blocks(ncp), LW_T, 0, st>>>(L, ncp, (uint32_t)rounds);
        k_lw_col_take<<<lw_blocks(ncp), LW_T, 0, st>>>(L, ncp, (uint32_t)rounds);
        g_launches += 2;
      }
      LWCK(cudaMemcpyAsync(hs, L.scal, sizeof(int) * LW_WORDS, cudaMemcpyDeviceToHost, st));
      LWCK(cudaStreamSynchronize(st));
      left = hs[LW_LEFT0 + ((rounds - 1) & 15)];
      if (rounds > 240) { snprintf(err, errlen, "colouring did not converge"); return -1; }
    }
    if (hs[LW_ERR]) { snprintf(err, errlen, "more than %d colours needed (a body with more than %d contact pairs)", OB_LW_MAXCOL, OB_LW_MAXCOL / 2); return -1; }
    k_lw_pairkey<<<lw_blocks(ncp), LW_T, 0, st>>>(L, ncp);
    const int pcur = lw_radix_sort(st, L.pkey, L.pidx, ncp, 2, L.tmp);
    if (pcur != 0) { snprintf(err, errlen, "internal: sort parity"); return -1; }
    k_lw_pairgather<<<lw_blocks(ncp), LW_T, 0, st>>>(L, ncp, L.pidx[0]);
    k_lw_segtab<<<1, 512, 0, st>>>(L, ncp, L.pkey[0]);
    g_launches += 3;
    LWCK(cudaMemcpyAsync(hs, L.scal, sizeof(int) * LW_WORDS, cudaMemcpyDeviceToHost, st));
    LWCK(cudaMemcpyAsync(hs + LW_WORDS, L.segtab, sizeof(int) * OB_LW_MAXCOL * (2 + OB_LW_MAXC), cudaMemcpyDeviceToHost, st));
    LWCK(cudaStreamSynchronize(st));
    ncol = hs[LW_NCOL];
    if (hs[LW_NSOLVED] > L.NC) { snprintf(err, errlen, "contact capacity exceeded (%d > %d): raise dBatchDesc.max_contacts_per_world", hs[LW_NSOLVED], L.NC); return -1; }
  }
  b->lw_rounds = rounds; b->lw_ncol = ncol;
  LW_MARK();
  // (5) bodies, rows
  k_lw_body_pre<<<lw_blocks(nb), LW_T, 0, st>>>(d, L, h);
  g_launches++;
  if (ncp > 0) { k_lw_assemble<<<lw_blocks(ncp), LW_T, 0, st>>>(d, L, ncp, maxc, m, h); g_launches++; }
  LW_MARK();
  // (6) SOR: colours in ascending order, every iteration
  int sor_launches = 0;
  if (ncp > 0 && hw.iters > 0 && !getenv("OB_LW_SOR_LAUNCHES")) {
    LWCK(cudaMemsetAsync(L.scal + LW_BARRIER, 0, sizeof(int), st));
    unsigned *bar = (unsigned *)(L.scal + LW_BARRIER);
    int iters = hw.iters, nc_ = ncol;
    void *kargs[] = {(void *)&L, (void *)&iters, (void *)&nc_, (void *)&bar};
    const void *fn = m == 3 ? (const void *)k_lw_sor_all<3> : (m == 2 ? (const void *)k_lw_sor_all<2> : (const void *)k_lw_sor_all<1>);
    if (b->lw_split_on && b->lw_split_sor) {
      ObLwSplit S = b->lw_split;
      void *sargs[] = {(void *)&L, (void *)&S, (void *)&iters, (void *)&nc_, (void *)&bar};
      const void *sfn = m == 3 ? (const void *)k_lw_sor_split<3> : (m == 2 ? (const void *)k_lw_sor_split<2> : (const void *)k_lw_sor_split<1>);
      LWCK(cudaLaunchCooperativeKernel(sfn, dim3(b->lw_split_grid[m - 1]), dim3(b->lw_split_threads), sargs, 0, st));
      b->lw_split.base += 1u + (unsigned)iters * (unsigned)nc_;   // the same on every rank: the phases before the solver are replicated
      LWCK(cudaMemcpyAsync(hs + LW_ERR, b->lw_flags + OB_LW_FLAG_TIMEOUT, sizeof(int), cudaMemcpyDeviceToHost, st));
    } else
    LWCK(cudaLaunchCooperativeKernel(fn, dim3(b->lw_sor_grid[m - 1]), dim3(LW_SOR_T), kargs, 0, st));
    g_launches++; sor_launches = 1;
  } else
  for (int it = 0; it < hw.iters && ncp > 0; it++)
    for (int c = 0; c < ncol; c++) {
      const int cnt = hs[LW_WORDS + c * (2 + OB_LW_MAXC) + 1];
      if (cnt <= 0) continue;
      if (m == 3) k_lw_sor<3><<<lw_blocks(cnt), LW_T, 0, st>>>(L, c);
      else if (m == 2) k_lw_sor<2><<<lw_blocks(cnt), LW_T, 0, st>>>(L, c);
      else k_lw_sor<1><<<lw_blocks(cnt), LW_T, 0, st>>>(L, c);
      g_launches++; sor_launches++;
    }
  LW_MARK();
  if ((taps & 1) && d.fback) {
    LWCK(cudaMemsetAsync(d.fback, 0, sizeof(real) * 12 * (size_t)d.NC, st));
    if (ncp > 0) { k_lw_feedback<<<lw_blocks(ncp), LW_T, 0, st>>>(d, L, ncp, maxc, m); g_launches++; }
  }
  // (7) integrate
  k_lw_body_post<<<lw_blocks(nb), LW_T, 0, st>>>(d, L, h);
  g_launches++;
  LW_MARK();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { snprintf(err, errlen, "large-world step failed: %s", cudaGetErrorString(e)); return -1; }
  if (b->lw_split_on) { LWCK(cudaMemcpy(hs + LW_ERR, b->lw_flags + OB_LW_FLAG_TIMEOUT, sizeof(int), cudaMemcpyDeviceToHost)); }
  if (b->lw_split_on && hs[LW_ERR]) { snprintf(err, errlen, "split step: a peer rank did not reach a barrier within the timeout (rank %d of %d)", b->lw_split.rank, b->lw_split.nranks); return -1; }
  if (tm) for (int k = 0; k + 1 < evi && k < 8; k++) { float ms = 0; cudaEventElapsedTime(&ms, b->lw_ev[k], b->lw_ev[k + 1]); b->lw_ms[k] += ms; }
  b->lw_stat[0] = np; b->lw_stat[1] = hs[LW_NCONTACTS]; b->lw_stat[2] = ncp; b->lw_stat[3] = ncp > 0 ? hs[LW_NSOLVED] : 0;
  b->lw_stat[4] = ncol; b->lw_stat[5] = rounds; b->lw_stat[6] = sor_launches; if (tm) b->lw_stat[7]++;
  if (getenv("OB_LW_VERBOSE")) fprintf(stderr, "lw: pairs %d contacts %d cpairs %d solved %d colours %d rounds %d sor launches %d\n", np, hs[LW_NCONTACTS], ncp, b->lw_stat[3], ncol, rounds, sor_launches);
#undef LW_MARK
  return 0;
}

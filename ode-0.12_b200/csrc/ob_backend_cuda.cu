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
#include <string.h>
#include <vector>
#include "ob_backend.h"
#include "ob_broad.h"
#include "ob_collide.h"
#include "ob_rows.h"
#include "ob_solver.h"

#define OB_THREADS 128
static long long g_launches = 0;

// ------------------------------------------------------------------------------------
// shared-memory layouts (one function for host sizing and device carving)
struct CollideSmem {
  size_t pose, aabb, cb, gid, body, cat, col, en, hr, br, walk_of, key, o12, sorted, misc, total;
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
  s.key = o; o = ob_al16(o + sizeof(ObPairKey) * NP);
  s.o12 = o; o = ob_al16(o + sizeof(int2) * NP);
  s.sorted = o; o = ob_al16(o + sizeof(int2) * NP);
  s.misc = o; o = ob_al16(o + sizeof(int) * 64);
  s.total = o;
  return s;
}

struct StepSmem {
  size_t bd, invIw, tmp1, fc, jb1, jb2, deg, adjstart, adj, btag, jtag, ibody, ijoint, isz, stack, jrow, lam, order,
      rnd, rowb, find, moved, gflag, misc, total;
};
__host__ __device__ inline StepSmem step_smem(int NB, int NG, int NC, int NR) {
  StepSmem s; size_t o = 0;
  s.bd = o; o = ob_al16(o + sizeof(ObBodyDyn) * NB);
  s.invIw = o; o = ob_al16(o + sizeof(real) * 12 * NB);
  s.tmp1 = o; o = ob_al16(o + sizeof(real) * 6 * NB);
  s.fc = o; o = ob_al16(o + sizeof(real) * 8 * NB);
  s.jb1 = o; o = ob_al16(o + sizeof(short) * NC);
  s.jb2 = o; o = ob_al16(o + sizeof(short) * NC);
  s.deg = o; o = ob_al16(o + sizeof(int) * NB);
  s.adjstart = o; o = ob_al16(o + sizeof(int) * (NB + 1));
  s.adj = o; o = ob_al16(o + sizeof(short) * 2 * NC);
  s.btag = o; o = ob_al16(o + sizeof(signed char) * NB);
  s.jtag = o; o = ob_al16(o + sizeof(signed char) * NC);
  s.ibody = o; o = ob_al16(o + sizeof(short) * NB);
  s.ijoint = o; o = ob_al16(o + sizeof(short) * NC);
  s.isz = o; o = ob_al16(o + sizeof(int) * 4 * NB);   // per island: body start, body count, joint start, joint count
  s.stack = o; o = ob_al16(o + sizeof(short) * NB);
  s.jrow = o; o = ob_al16(o + sizeof(int) * (NC + 1));   // row offset of joint k (ijoint order)
  s.lam = o; o = ob_al16(o + sizeof(real) * NR);
  s.order = o; o = ob_al16(o + sizeof(unsigned short) * NR);
  s.rnd = o; o = ob_al16(o + sizeof(unsigned char) * NR);
  s.rowb = o; o = ob_al16(o + sizeof(short) * 2 * NR);
  s.find = o; o = ob_al16(o + sizeof(int) * NR);
  s.moved = o; o = ob_al16(o + sizeof(short) * NG);
  s.gflag = o; o = ob_al16(o + sizeof(int) * 2 * NG);
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
__global__ void __launch_bounds__(OB_THREADS) k_collide(ObBatchDev d) {
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
  ObPairKey *s_key = (ObPairKey *)(smem + L.key);
  int2 *s_o12 = (int2 *)(smem + L.o12);
  int2 *s_sorted = (int2 *)(smem + L.sorted);
  int *s_misc = (int *)(smem + L.misc);   // [0]=npairs raw, [1]=nh, [2]=nbig, [3]=contact base, [8..40]=scan scratch
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int w = blockIdx.x; w < d.W; w += gridDim.x) {
    ObWorld &W = d.world[w];
    const int ng = W.ng;
    const ObGeom *geoms = d.geom + (size_t)w * d.NG;
    const ObBodyDyn *bd = d.bdyn + (size_t)w * d.NB;
    const int *glist = d.glist + (size_t)w * d.NG;
    if (tid < 8) s_misc[tid] = 0;
    // (1) pose, AABB, cell box per geom in walk order
    for (int i = tid; i < ng; i += nt) {
      int gi = glist[i];
      const ObGeom g = geoms[gi];
      s_gid[i] = gi; s_body[i] = g.body; s_cat[i] = g.cat; s_col[i] = g.col;
      s_walk_of[gi] = i;
      s_en[i] = (g.flags & OB_GEOM_ENABLED) && !(g.flags & OB_GEOM_ZERO_SIZED);
      ObPose p;
      geom_pose_dev(g, bd, &p);
      s_pose[i] = p;
      real ab[6];
      ob_aabb(p, ab);
      for (int k = 0; k < 6; k++) s_aabb[6 * i + k] = ab[k];
      ObCellBox cb;
      cb.level = 0;
      for (int k = 0; k < 6; k++) cb.db[k] = 0;
      ob_hash_cellbox(ab, W.hash_minlevel, W.hash_maxlevel, &cb);
      s_cb[i] = cb;
    }
    __syncthreads();
    // (2) ranks among hashed / big geoms in walk order
    for (int i = tid; i < ng; i += nt) {
      int h = 0, b = 0;
      for (int j = 0; j < i; j++)
        if (s_en[j]) { if (s_cb[j].level == OB_LEVEL_BIG) b++; else h++; }
      s_hr[i] = h; s_br[i] = b;
      if (i == ng - 1) {
        if (s_en[i]) { if (s_cb[i].level == OB_LEVEL_BIG) b++; else h++; }
        s_misc[1] = h; s_misc[2] = b;
      }
    }
    __syncthreads();
    const int nh = s_misc[1], nbig = s_misc[2];
    // (3) candidate pairs: collideAABBs filter + first-encounter key
    for (int idx = tid; idx < ng * ng; idx += nt) {
      int a = idx / ng, b = idx - a * ng;
      if (a >= b || !s_en[a] || !s_en[b]) continue;
      if (!ob_aabb_pair_filter(s_body[a], s_body[b], s_cat[a], s_col[a], s_cat[b], s_col[b], s_aabb + 6 * a, s_aabb + 6 * b))
        continue;
      ObPairKey key;
      int first_is_a;
      if (!ob_hash_pair_key(a, b, s_cb[a], s_cb[b], s_hr[a], s_hr[b], s_br[a], s_br[b], nh, nbig, &key, &first_is_a)) continue;
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
    const int maxc = pol.max_contacts > OB_MAXC_LOCAL ? OB_MAXC_LOCAL : pol.max_contacts;
    for (int base = 0; base < np; base += nt) {
      int p = base + tid;
      ObCg cg[OB_MAXC_LOCAL];
      int n = 0, o1 = 0, o2 = 0;
      if (p < np) {
        o1 = s_sorted[p].x; o2 = s_sorted[p].y;
        int swapped;
        n = ob_collide_pair(s_pose[s_walk_of[o1]], s_pose[s_walk_of[o2]], maxc, cg, &swapped);
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

// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OB_THREADS) k_step(ObBatchDev d, real h, int taps) {
  extern __shared__ __align__(16) unsigned char smem[];
  const StepSmem L = step_smem(d.NB, d.NG, d.NC, d.NR);
  ObBodyDyn *s_bd = (ObBodyDyn *)(smem + L.bd);
  real *s_invIw = (real *)(smem + L.invIw);
  real *s_tmp1 = (real *)(smem + L.tmp1);
  real *s_fc = (real *)(smem + L.fc);          // [body][8]
  short *s_jb1 = (short *)(smem + L.jb1);
  short *s_jb2 = (short *)(smem + L.jb2);
  int *s_deg = (int *)(smem + L.deg);
  int *s_adjstart = (int *)(smem + L.adjstart);
  short *s_adj = (short *)(smem + L.adj);
  signed char *s_btag = (signed char *)(smem + L.btag);
  signed char *s_jtag = (signed char *)(smem + L.jtag);
  short *s_ibody = (short *)(smem + L.ibody);
  short *s_ijoint = (short *)(smem + L.ijoint);
  int *s_isz = (int *)(smem + L.isz);
  short *s_stack = (short *)(smem + L.stack);
  int *s_jrow = (int *)(smem + L.jrow);
  real *s_lam = (real *)(smem + L.lam);
  unsigned short *s_order = (unsigned short *)(smem + L.order);
  unsigned char *s_rnd = (unsigned char *)(smem + L.rnd);
  short *s_rowb = (short *)(smem + L.rowb);
  short *s_moved = (short *)(smem + L.moved);
  int *s_find = (int *)(smem + L.find);
  int *s_gflag = (int *)(smem + L.gflag);
  int *s_misc = (int *)(smem + L.misc);   // [0]=nislands [1]=nib [2]=nij [3]=m total [4]=nmoved [8..]=scan scratch
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarps = nt >> 5;
  const real stepsize1 = ob_recip(h);

  for (int w = blockIdx.x; w < d.W; w += gridDim.x) {
    ObWorld &W = d.world[w];
    const int nb = W.nb, ng = W.ng;
    int nc = d.ncontacts[w];
    ObBodyDyn *gbd = d.bdyn + (size_t)w * d.NB;
    const ObBodyConst *bc = d.bconst + (size_t)w * d.NB;
    const ObGeom *geoms = d.geom + (size_t)w * d.NG;
    const ObContact *con = d.contacts + (size_t)w * d.NC;
    real *rowJ = d.rowJ + (size_t)w * d.NR * 12, *rowiMJ = d.rowiMJ + (size_t)w * d.NR * 12;
    real *rowJc = d.rowJc + (size_t)w * d.NR * 12;
    real *rowS = d.rowS + (size_t)w * d.NR * 4;
    int *rowI = d.rowI + (size_t)w * d.NR * 4;

    // (0) stage body state in shared memory (coalesced 128-bit copy)
    {
      const uint4 *src = (const uint4 *)gbd;
      uint4 *dst = (uint4 *)s_bd;
      const int n16 = (int)(sizeof(ObBodyDyn) / 16) * nb;
      for (int i = tid; i < n16; i += nt) dst[i] = src[i];
    }
    if (tid < 8) s_misc[tid] = 0;
    for (int b = tid; b < nb; b += nt) { s_deg[b] = 0; s_btag[b] = 0; }
    // (1) contact joint -> bodies (dJointAttach swap rule, ode.cpp:1368-1377)
    for (int j = tid; j < nc; j += nt) {
      int b1 = geoms[con[j].g1].body, b2 = geoms[con[j].g2].body;
      if (b1 < 0) { b1 = b2; b2 = -1; }
      s_jb1[j] = (short)b1; s_jb2[j] = (short)b2; s_jtag[j] = 0;
    }
    __syncthreads();
    // (2) per-body joint lists, newest joint first
    for (int b = tid; b < nb; b += nt) {
      int c = 0;
      for (int j = 0; j < nc; j++) c += (s_jb1[j] == b || s_jb2[j] == b) ? 1 : 0;
      s_deg[b] = c;
    }
    __syncthreads();
    if (tid == 0) { int a = 0; for (int b = 0; b < nb; b++) { s_adjstart[b] = a; a += s_deg[b]; } s_adjstart[nb] = a; }
    __syncthreads();
    for (int b = tid; b < nb; b += nt) {
      int o = s_adjstart[b];
      for (int j = nc - 1; j >= 0; j--) if (s_jb1[j] == b || s_jb2[j] == b) s_adj[o++] = (short)j;
    }
    // (3) auto-disable, instantaneous-sample mode (util.cpp:99-233)
    for (int b = tid; b < nb; b += nt) {
      ObBodyDyn &B = s_bd[b];
      if (s_deg[b] == 0) continue;
      if ((B.flags & (OB_BODY_AUTO_DISABLE | OB_BODY_DISABLED)) != OB_BODY_AUTO_DISABLE) continue;
      const ObBodyConst &C = bc[b];
      if (C.adis_samples == 0) continue;
      int idle = 1;
      real ls = ob_dot(B.lvel, B.lvel);
      if (ls > C.adis_lin_thr) idle = 0;
      else { real as = ob_dot(B.avel, B.avel); if (as > C.adis_ang_thr) idle = 0; }
      if (idle) { B.adis_stepsleft--; B.adis_timeleft -= h; }
      else { B.adis_stepsleft = C.adis_idle_steps; B.adis_timeleft = C.adis_idle_time; }
      if (B.adis_stepsleft <= 0 && B.adis_timeleft <= 0) {
        B.flags |= OB_BODY_DISABLED;
        for (int k = 0; k < 3; k++) { B.lvel[k] = 0; B.avel[k] = 0; }
      }
    }
    __syncthreads();
    // (4) islands: the reference's DFS, one thread (util.cpp:411-487)
    if (tid == 0) {
      int nib = 0, nij = 0, nis = 0;
      for (int bb = 0; bb < nb; bb++) {
        if (s_btag[bb]) continue;
        if (s_bd[bb].flags & OB_BODY_DISABLED) { s_btag[bb] = -1; continue; }
        s_btag[bb] = 1;
        const int b0 = nib, j0 = nij;
        s_ibody[nib++] = (short)bb;
        int sp = 0, b = bb;
        while (true) {
          const int e = s_adjstart[b + 1];
          for (int k = s_adjstart[b]; k < e; k++) {
            const int j = s_adj[k];
            if (!s_jtag[j]) {
              const int j1 = s_jb1[j], j2 = s_jb2[j];
              const bool enabled = (bc[j1].invMass > 0) || (j2 >= 0 && bc[j2].invMass > 0);
              if (enabled) {
                s_jtag[j] = 1;
                s_ijoint[nij++] = (short)j;
                const int other = (j1 == b) ? j2 : j1;
                if (other >= 0 && s_btag[other] <= 0) {
                  s_btag[other] = 1;
                  s_bd[other].flags &= ~OB_BODY_DISABLED;
                  s_stack[sp++] = (short)other;
                }
              } else s_jtag[j] = -1;
            }
          }
          if (sp == 0) break;
          b = s_stack[--sp];
          s_ibody[nib++] = (short)b;
        }
        s_isz[4 * nis + 0] = b0; s_isz[4 * nis + 1] = nib - b0; s_isz[4 * nis + 2] = j0; s_isz[4 * nis + 3] = nij - j0;
        nis++;
      }
      s_misc[0] = nis; s_misc[1] = nib; s_misc[2] = nij;
    }
    __syncthreads();
    const int nis = s_misc[0], nib = s_misc[1], nij = s_misc[2];

    // (5) per-body preamble for every island body (quickstep.cpp:610-665)
    for (int i = tid; i < nib; i += nt) {
      const int b = s_ibody[i];
      ObBodyDyn &B = s_bd[b];
      const ObBodyConst &C = bc[b];
      real I[12], invI[12], iw[12];
      for (int k = 0; k < 12; k++) { I[k] = C.I[k]; invI[k] = C.invI[k]; }
      ob_body_preamble(B.R, I, invI, B.avel, B.flags, C.mass, W.gravity, iw, B.facc, B.tacc);
      for (int k = 0; k < 12; k++) s_invIw[12 * b + k] = iw[k];
      ob_body_tmp1(B.facc, B.tacc, B.lvel, B.avel, C.invMass, iw, stepsize1, s_tmp1 + 6 * b);
      for (int k = 0; k < 8; k++) s_fc[8 * b + k] = 0;
    }
    // (6) rows per joint (getInfo1) and row offsets in island joint order
    const ObSurface surf0 = d.policy[0].surface;
    int mtot = 0;
    for (int base = 0; base < nij || base == 0; base += nt) {
      const int k = base + tid;
      int m = 0;
      if (k < nij) { ObSurface sf = surf0; m = ob_contact_info1(sf); }
      int total;
      const int off = block_excl_scan(m, s_misc + 8, &total);
      if (k < nij) s_jrow[k] = mtot + off;
      mtot += total;
      if (base + nt >= nij) break;
    }
    if (tid == 0) s_jrow[nij] = mtot;
    __syncthreads();
    if (mtot > d.NR) {   // capacity: solve nothing rather than corrupt memory; flagged per world
      if (tid == 0) atomicOr(&W.status, OB_ERR_ROW_OVERFLOW);
      mtot = 0;
    }
    const bool have_rows = mtot > 0;
    // (7) row assembly (getInfo2) + finalisation, one thread per joint
    if (have_rows) {
      for (int k = tid; k < nij; k += nt) {
        const int j = s_ijoint[k];
        ObSurface sf = surf0;
        const int jm = ob_contact_info1(sf);
        ObRowOut3 r;
        ob_rows_defaults(r, jm, W.cfm);
        const ObContact c = con[j];
        const int b1 = s_jb1[j], b2 = s_jb2[j];
        const int rev = geoms[c.g1].body < 0;
        const real zero3[3] = {0, 0, 0};
        const real fdir1[3] = {0, 0, 0};
        ob_contact_info2(r, jm, sf, c.pos, c.normal, c.depth, fdir1, rev, s_bd[b1].pos, s_bd[b1].lvel, s_bd[b1].avel,
                         b2 >= 0, b2 >= 0 ? s_bd[b2].pos : zero3, b2 >= 0 ? s_bd[b2].lvel : zero3,
                         b2 >= 0 ? s_bd[b2].avel : zero3, stepsize1, W.erp, W.min_depth, W.max_vel);
        const int r0 = s_jrow[k];
        const real invM1 = bc[b1].invMass, invM2 = b2 >= 0 ? bc[b2].invMass : (real)0;
        for (int q = 0; q < jm; q++) {
          const int ri = r0 + q;
          if (taps) for (int e = 0; e < 12; e++) rowJc[(size_t)ri * 12 + e] = r.J[q][e];
          real iMJ[12], b_out, adcfm;
          ob_row_finalize(r.J[q], r.c[q], r.cfm[q], b2, s_tmp1 + 6 * b1, b2 >= 0 ? s_tmp1 + 6 * b2 : s_tmp1, invM1,
                          s_invIw + 12 * b1, invM2, b2 >= 0 ? s_invIw + 12 * b2 : s_invIw, stepsize1, W.sor_w, iMJ,
                          &b_out, &adcfm);
          for (int e = 0; e < 12; e++) { rowJ[(size_t)ri * 12 + e] = r.J[q][e]; rowiMJ[(size_t)ri * 12 + e] = iMJ[e]; }
          rowS[ri * 4 + 0] = b_out; rowS[ri * 4 + 1] = adcfm; rowS[ri * 4 + 2] = r.lo[q]; rowS[ri * 4 + 3] = r.hi[q];
          rowI[ri * 4 + 0] = r.findex[q] >= 0 ? r.findex[q] + r0 : -1;
          rowI[ri * 4 + 1] = b1; rowI[ri * 4 + 2] = b2; rowI[ri * 4 + 3] = j;
          s_rowb[2 * ri] = (short)b1; s_rowb[2 * ri + 1] = (short)b2;
          s_find[ri] = r.findex[q] >= 0 ? r.findex[q] + r0 : -1;
          s_lam[ri] = 0;
        }
      }
    }
    __syncthreads();   // rows visible to the whole CTA (global writes by this CTA + smem)

    // (8) SOR, one warp per island (islands round-robin over the CTA's warps)
    // per-island LCG offset: islands consume the world's stream in island order
    for (int isl = wid; isl < nis; isl += nwarps) {
      const int j0 = s_isz[4 * isl + 2], jn = s_isz[4 * isl + 3];
      if (!have_rows || jn == 0) continue;
      const int r0 = s_jrow[j0], m = s_jrow[j0 + jn] - r0;
      if (m == 0) continue;
      const int nshuf = (W.iters + 7) >> 3;
      // draws consumed by earlier islands
      uint32_t seed = W.seed;
      {
        unsigned skip = 0;
        for (int p = 0; p < isl; p++) {
          const int pj0 = s_isz[4 * p + 2], pjn = s_isz[4 * p + 3];
          const int pm = pjn ? (s_jrow[pj0 + pjn] - s_jrow[pj0]) : 0;
          if (pm >= 2) skip += (unsigned)nshuf * (unsigned)(pm - 1);
        }
        uint32_t A, C;
        ob_lcg_skip(skip, &A, &C);
        seed = A * seed + C;
      }
      unsigned short *ord = s_order + r0;
      unsigned char *rnd = s_rnd + r0;
      // initial order: findex==-1 rows ascending at the head, the rest descending at the tail
      if (lane == 0) {
        int head = 0, tail = m - 1;
        for (int i = 0; i < m; i++) { if (s_find[r0 + i] == -1) ord[head++] = (unsigned short)i; else ord[tail--] = (unsigned short)i; }
      }
      __syncwarp();
      const int nwin = (m + 31) >> 5;
      for (int it = 0; it < W.iters; it++) {
        if ((it & 7) == 0) {
          if (lane == 0) {
            for (int i = 1; i < m; i++) {
              seed = ob_lcg_next(seed);
              const int swapi = ob_randint_fold(seed, (uint32_t)(i + 1));
              const unsigned short t = ord[i]; ord[i] = ord[swapi]; ord[swapi] = t;
            }
          }
          __syncwarp();
          // dependency rounds inside each 32-row window
          for (int win = 0; win < nwin; win++) {
            const int k = win * 32 + lane;
            int b1 = -2, b2 = -2;
            if (k < m) { const int idx = r0 + ord[k]; b1 = s_rowb[2 * idx]; b2 = s_rowb[2 * idx + 1]; }
            int myr = 1;
            for (int l2 = 0; l2 < 31; l2++) {
              const int pb1 = __shfl_sync(0xffffffffu, b1, l2), pb2 = __shfl_sync(0xffffffffu, b2, l2);
              const int pr = __shfl_sync(0xffffffffu, myr, l2);
              if (l2 < lane && pb1 >= 0 && b1 >= 0) {
                const bool share = (b1 == pb1) || (b1 == pb2) || (b2 >= 0 && (b2 == pb1 || b2 == pb2));
                if (share && pr + 1 > myr) myr = pr + 1;
              }
            }
            if (k < m) rnd[k] = (unsigned char)myr;
          }
          __syncwarp();
        }
        for (int win = 0; win < nwin; win++) {
          const int k = win * 32 + lane;
          const bool act = k < m;
          int idx = 0, myr = 0, fi = -1, b1 = 0, b2 = -1;
          real J[12], iMJ[12], sb = 0, sad = 0, slo = 0, shi = 0;
          if (act) {
            idx = r0 + ord[k];
            myr = rnd[k];
            const float4 *pj = (const float4 *)(rowJ + (size_t)idx * 12);
            const float4 *pm = (const float4 *)(rowiMJ + (size_t)idx * 12);
#if defined(dSINGLE)
            float4 a0 = pj[0], a1 = pj[1], a2 = pj[2], m0 = pm[0], m1 = pm[1], m2 = pm[2];
            J[0] = a0.x; J[1] = a0.y; J[2] = a0.z; J[3] = a0.w; J[4] = a1.x; J[5] = a1.y; J[6] = a1.z; J[7] = a1.w;
            J[8] = a2.x; J[9] = a2.y; J[10] = a2.z; J[11] = a2.w;
            iMJ[0] = m0.x; iMJ[1] = m0.y; iMJ[2] = m0.z; iMJ[3] = m0.w; iMJ[4] = m1.x; iMJ[5] = m1.y; iMJ[6] = m1.z;
            iMJ[7] = m1.w; iMJ[8] = m2.x; iMJ[9] = m2.y; iMJ[10] = m2.z; iMJ[11] = m2.w;
            const float4 s4 = *(const float4 *)(rowS + (size_t)idx * 4);
            sb = s4.x; sad = s4.y; slo = s4.z; shi = s4.w;
#else
            (void)pj; (void)pm;
            for (int e = 0; e < 12; e++) { J[e] = rowJ[(size_t)idx * 12 + e]; iMJ[e] = rowiMJ[(size_t)idx * 12 + e]; }
            sb = rowS[idx * 4]; sad = rowS[idx * 4 + 1]; slo = rowS[idx * 4 + 2]; shi = rowS[idx * 4 + 3];
#endif
            fi = s_find[idx];
            b1 = s_rowb[2 * idx]; b2 = s_rowb[2 * idx + 1];
          }
          int maxr = myr;
          for (int dd = 16; dd; dd >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, maxr, dd); maxr = o > maxr ? o : maxr; }
          for (int r = 1; r <= maxr; r++) {
            if (act && myr == r) {
              real *fc1 = s_fc + 8 * b1;
              real *fc2 = b2 >= 0 ? s_fc + 8 * b2 : (real *)0;
              s_lam[idx] = ob_sor_row(J, iMJ, sb, sad, slo, shi, fi, fi >= 0 ? s_lam[fi] : (real)0, s_lam[idx], fc1, fc2);
            }
            __syncwarp();
          }
        }
      }
      // joint feedback tap (quickstep.cpp:918-957) for parity tests: f1,t1 = Jcopy^T lambda
      if (taps) {
        real *fb = d.fback + (size_t)w * d.NC * 6;
        for (int k = j0 + lane; k < j0 + jn; k += 32) {
          const int jr0 = s_jrow[k], jm = s_jrow[k + 1] - jr0;
          real acc[6] = {0, 0, 0, 0, 0, 0};
          for (int q = 0; q < jm; q++) {
            const real s = s_lam[jr0 + q];
            for (int e = 0; e < 6; e++) acc[e] += rowJc[(size_t)(jr0 + q) * 12 + e] * s;
          }
          for (int e = 0; e < 6; e++) fb[s_ijoint[k] * 6 + e] = acc[e];
        }
      }
    }
    __syncthreads();
    // the world's LCG advances by every island's draws
    if (tid == 0) {
      unsigned skip = 0;
      const int nshuf = (W.iters + 7) >> 3;
      if (have_rows)
        for (int p = 0; p < nis; p++) {
          const int pj0 = s_isz[4 * p + 2], pjn = s_isz[4 * p + 3];
          const int pm = pjn ? (s_jrow[pj0 + pjn] - s_jrow[pj0]) : 0;
          if (pm >= 2) skip += (unsigned)nshuf * (unsigned)(pm - 1);
        }
      uint32_t A, C;
      ob_lcg_skip(skip, &A, &C);
      W.seed = A * W.seed + C;
    }
    // (9) velocity update + integration per island body; (10) accumulators cleared
    for (int i = tid; i < nib; i += nt) {
      const int b = s_ibody[i];
      ObBodyDyn &B = s_bd[b];
      const ObBodyConst &C = bc[b];
      // which island is this body in -> did it have rows?
      bool island_rows = false;
      if (have_rows) {
        for (int p = 0; p < nis; p++) {
          const int pb0 = s_isz[4 * p], pbn = s_isz[4 * p + 1];
          if (i >= pb0 && i < pb0 + pbn) {
            const int pj0 = s_isz[4 * p + 2], pjn = s_isz[4 * p + 3];
            island_rows = pjn && (s_jrow[pj0 + pjn] - s_jrow[pj0]) > 0;
            break;
          }
        }
      }
      ob_body_velocity_update(B.lvel, B.avel, island_rows ? s_fc + 8 * b : (real *)0, B.facc, B.tacc, C.invMass,
                              s_invIw + 12 * b, h);
      real fra[3] = {C.finite_rot_axis[0], C.finite_rot_axis[1], C.finite_rot_axis[2]};
      ob_step_body(B.pos, B.q, B.R, B.lvel, B.avel, B.flags, h, C.max_angular_speed, fra, C.damp_lin_scale,
                   C.damp_ang_scale, C.damp_lin_thr, C.damp_ang_thr);
      for (int k = 0; k < 4; k++) { B.facc[k] = 0; B.tacc[k] = 0; }
    }
    __syncthreads();
    // (11) space list: every geom of a stepped body moves to the head, in stepping order
    int *glist = d.glist + (size_t)w * d.NG;
    if (tid == 0) {
      int nm = 0;
      for (int i = 0; i < nib; i++)
        for (int g = bc[s_ibody[i]].geom_first; g >= 0; g = geoms[g].body_next) s_moved[nm++] = (short)g;
      s_misc[4] = nm;
    }
    __syncthreads();
    {
      const int nm = s_misc[4];
      int *s_flag = s_gflag;
      for (int g = tid; g < d.NG; g += nt) s_flag[g] = 0;
      __syncthreads();
      for (int i = tid; i < nm; i += nt) s_flag[s_moved[i]] = 1;
      __syncthreads();
      int *s_old = s_flag + d.NG;   // old list snapshot
      for (int i = tid; i < ng; i += nt) s_old[i] = glist[i];
      __syncthreads();
      for (int i = tid; i < ng; i += nt) {
        const int g = s_old[i];
        if (!s_flag[g]) {
          int before = 0;
          for (int j2 = 0; j2 < i; j2++) before += s_flag[s_old[j2]] ? 0 : 1;
          glist[nm + before] = g;
        }
      }
      for (int i = tid; i < nm; i += nt) glist[nm - 1 - i] = s_moved[i];
    }
    __syncthreads();
    // (12) write back body state, lambda, counters
    {
      uint4 *dst = (uint4 *)gbd;
      const uint4 *src = (const uint4 *)s_bd;
      const int n16 = (int)(sizeof(ObBodyDyn) / 16) * nb;
      for (int i = tid; i < n16; i += nt) dst[i] = src[i];
    }
    if (taps) { real *gl = d.lambda + (size_t)w * d.NR; for (int i = tid; i < mtot; i += nt) gl[i] = s_lam[i]; }
    if (tid == 0) {
      d.nrows[w] = mtot;
      atomicAdd(&d.counters->steps, 1ull);
      atomicAdd(&d.counters->body_steps, (unsigned long long)nib);
      atomicAdd(&d.counters->contacts, (unsigned long long)(have_rows ? nij : 0));
      atomicAdd(&d.counters->rows, (unsigned long long)mtot);
      atomicAdd(&d.counters->islands, (unsigned long long)nis);
      if (W.status) atomicAdd(&d.counters->overflow_worlds, 1ull);
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
    for (int k = 0; k < 4; k++) s.q[k] = quat4[(size_t)i * 4 + k];
    ob_safe_normalize4(s.q);
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
  size_t smem_collide, smem_step;
  int grid;
  cudaEvent_t ev[4];   // 0,1: user timer; 1..3 reused for per-kernel timing
  int ktiming;
  double kms[OBK_NKERNELS];
  long long klaunch[OBK_NKERNELS];
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_)); goto fail; } } while (0)

template <class T> static cudaError_t dalloc(ObBackend *b, T **p, size_t n) {
  void *q = 0;
  cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
  if (e == cudaSuccess) { b->allocs.push_back(q); cudaMemset(q, 0, (n ? n : 1) * sizeof(T)); }
  *p = (T *)q;
  return e;
}

ObBackend *obk_create(const ObBatchDev &caps, int device, char *err, size_t errlen) {
  ObBackend *b = new ObBackend;
  b->d = caps; b->device = device; b->stream = 0; b->st_dev = 0; b->st_host = 0;
  b->ktiming = 0;
  for (int k = 0; k < OBK_NKERNELS; k++) { b->kms[k] = 0; b->klaunch[k] = 0; }
  for (int k = 0; k < 4; k++) b->ev[k] = 0;
  ObBatchDev &d = b->d;
  const size_t W = d.W;
  int ndev = 0;
  cudaDeviceProp prop;
  if (d.NB > 32000 || d.NC > 32000 || d.NR > 65000) { snprintf(err, errlen, "world too large for the CTA-per-world path (NB=%d NC=%d NR=%d)", d.NB, d.NC, d.NR); goto fail2; }
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { snprintf(err, errlen, "no CUDA device available (this library has no CPU fallback)"); goto fail2; }
  CK(cudaSetDevice(device));
  CK(cudaGetDeviceProperties(&prop, device));
  CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  for (int k = 0; k < 4; k++) CK(cudaEventCreate(&b->ev[k]));
  CK(dalloc(b, &d.world, W));
  CK(dalloc(b, &d.bdyn, W * d.NB));
  CK(dalloc(b, &d.bconst, W * d.NB));
  CK(dalloc(b, &d.geom, W * d.NG));
  CK(dalloc(b, &d.glist, W * d.NG));
  CK(dalloc(b, &d.policy, (size_t)d.npolicy));
  CK(dalloc(b, &d.npairs, W));
  CK(dalloc(b, &d.pairs, W * d.NP * 2));
  CK(dalloc(b, &d.ncontacts, W));
  CK(dalloc(b, &d.contacts, W * d.NC));
  CK(dalloc(b, &d.rowJ, W * d.NR * 12));
  CK(dalloc(b, &d.rowiMJ, W * d.NR * 12));
  CK(dalloc(b, &d.rowJc, W * d.NR * 12));
  CK(dalloc(b, &d.rowS, W * d.NR * 4));
  CK(dalloc(b, &d.rowI, W * d.NR * 4));
  CK(dalloc(b, &d.lambda, W * d.NR));
  CK(dalloc(b, &d.nrows, W));
  CK(dalloc(b, &d.fback, W * d.NC * 6));
  CK(dalloc(b, &d.counters, (size_t)1));
  b->st_elems = W * d.NB;
  CK(dalloc(b, &b->st_dev, b->st_elems * 13));
  CK(cudaMallocHost((void **)&b->st_host, b->st_elems * 13 * sizeof(real)));
  b->smem_collide = collide_smem(d.NG, d.NP).total;
  b->smem_step = step_smem(d.NB, d.NG, d.NC, d.NR).total;
  if (b->smem_collide > (size_t)prop.sharedMemPerBlockOptin || b->smem_step > (size_t)prop.sharedMemPerBlockOptin) {
    snprintf(err, errlen, "world does not fit one CTA's shared memory (collide %zu B, step %zu B, limit %zu B)",
             b->smem_collide, b->smem_step, (size_t)prop.sharedMemPerBlockOptin);
    goto fail;
  }
  CK(cudaFuncSetAttribute(k_collide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_collide));
  CK(cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_step));
  {
    // grid: every world gets its own CTA up to 16 resident CTAs per SM worth of blocks, beyond that grid-stride
    int cap = prop.multiProcessorCount * 16;
    b->grid = (int)W < cap ? (int)W : cap;
  }
  return b;
fail:
  for (size_t i = 0; i < b->allocs.size(); i++) cudaFree(b->allocs[i]);
  if (b->st_host) cudaFreeHost(b->st_host);
  if (b->stream) cudaStreamDestroy(b->stream);
fail2:
  delete b;
  return 0;
}

void obk_destroy(ObBackend *b) {
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  for (size_t i = 0; i < b->allocs.size(); i++) cudaFree(b->allocs[i]);
  if (b->st_host) cudaFreeHost(b->st_host);
  for (int k = 0; k < 4; k++) if (b->ev[k]) cudaEventDestroy(b->ev[k]);
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
  for (int k = 0; k < OBK_NKERNELS; k++) { b->kms[k] = 0; b->klaunch[k] = 0; }
}
void obk_get_kernel_times(ObBackend *b, double *ms, long long *l) {
  for (int k = 0; k < OBK_NKERNELS; k++) { ms[k] = b->kms[k]; l[k] = b->klaunch[k]; }
}
const char *obk_kernel_name(int k) { return k == 0 ? "k_collide" : (k == 1 ? "k_step" : ""); }

int obk_step(ObBackend *b, real h, int nsteps, int taps, char *err, size_t errlen) {
  cudaSetDevice(b->device);
  for (int s = 0; s < nsteps; s++) {
    if (b->ktiming) cudaEventRecord(b->ev[1], b->stream);
    k_collide<<<b->grid, OB_THREADS, b->smem_collide, b->stream>>>(b->d);
    if (b->ktiming) cudaEventRecord(b->ev[2], b->stream);
    k_step<<<b->grid, OB_THREADS, b->smem_step, b->stream>>>(b->d, h, taps);
    g_launches += 2;
    if (b->ktiming) {
      cudaEventRecord(b->ev[3], b->stream);
      if (cudaEventSynchronize(b->ev[3]) == cudaSuccess) {
        float m0 = 0, m1 = 0;
        cudaEventElapsedTime(&m0, b->ev[1], b->ev[2]);
        cudaEventElapsedTime(&m1, b->ev[2], b->ev[3]);
        b->kms[0] += m0; b->kms[1] += m1; b->klaunch[0]++; b->klaunch[1]++;
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  if (e != cudaSuccess) { snprintf(err, errlen, "kernel launch/exec failed: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

int obk_get_state(ObBackend *b, real *pos3, real *quat4, real *lvel3, real *avel3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *dp = b->st_dev, *dq = dp + n * 3, *dl = dq + n * 4, *da = dl + n * 3;
  k_pack_state<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, dp, dq, dl, da);
  g_launches++;
  if (cudaMemcpyAsync(b->st_host, b->st_dev, n * 13 * sizeof(real), cudaMemcpyDeviceToHost, b->stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) return -1;
  const real *hp = b->st_host, *hq = hp + n * 3, *hl = hq + n * 4, *ha = hl + n * 3;
  if (pos3) memcpy(pos3, hp, n * 3 * sizeof(real));
  if (quat4) memcpy(quat4, hq, n * 4 * sizeof(real));
  if (lvel3) memcpy(lvel3, hl, n * 3 * sizeof(real));
  if (avel3) memcpy(avel3, ha, n * 3 * sizeof(real));
  return 0;
}
int obk_set_state(ObBackend *b, const real *pos3, const real *quat4, const real *lvel3, const real *avel3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *hp = b->st_host, *hq = hp + n * 3, *hl = hq + n * 4, *ha = hl + n * 3;
  if (pos3) memcpy(hp, pos3, n * 3 * sizeof(real));
  if (quat4) memcpy(hq, quat4, n * 4 * sizeof(real));
  if (lvel3) memcpy(hl, lvel3, n * 3 * sizeof(real));
  if (avel3) memcpy(ha, avel3, n * 3 * sizeof(real));
  if (cudaMemcpyAsync(b->st_dev, b->st_host, n * 13 * sizeof(real), cudaMemcpyHostToDevice, b->stream) != cudaSuccess) return -1;
  real *dp = b->st_dev, *dq = dp + n * 3, *dl = dq + n * 4, *da = dl + n * 3;
  k_unpack_state<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, pos3 ? dp : 0, quat4 ? dq : 0, lvel3 ? dl : 0, avel3 ? da : 0);
  g_launches++;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_add_forces(ObBackend *b, const real *f3, const real *t3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *hf = b->st_host, *ht = hf + n * 3;
  if (f3) memcpy(hf, f3, n * 3 * sizeof(real));
  if (t3) memcpy(ht, t3, n * 3 * sizeof(real));
  if (cudaMemcpyAsync(b->st_dev, b->st_host, n * 6 * sizeof(real), cudaMemcpyHostToDevice, b->stream) != cudaSuccess) return -1;
  k_add_forces<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, f3 ? b->st_dev : 0, t3 ? b->st_dev + n * 3 : 0);
  g_launches++;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}

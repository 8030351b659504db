// ob_step_kernel.cuh — dWorldQuickStep for a batch of small worlds, three kernels:
//
//   k_prep<G> : contact joints -> body/joint graph -> island DFS (order-exact,
//               ode/src/util.cpp:411-487) -> per-body preamble -> row assembly and
//               finalisation (quickstep.cpp:670-857, SOR_LCP :355-402) -> for every shuffle
//               epoch the reference's seeded row order (:409-482) and from it a LEVEL
//               SCHEDULE (below), written to global memory.
//   k_sor<G>  : the SOR sweeps (quickstep.cpp:484-582).  Shared memory holds only the state
//               the sweep's dependency chain runs through: fc (6 reals per body) and lambda.
//               Rows are compact 24-word records streamed from L2 with a one-pass-ahead
//               register prefetch.
//   k_post<G> : velocity update, dxStepBody, accumulator clear, space-list reorder.
//
// Mapping: one world per G-lane tile, 32/G worlds per warp, one warp per CTA.  Control flow is
// warp-uniform (bounds are maxima over the warp's tiles; idle tiles are predicated off).
//
// Level schedule = the result-preserving parallelisation of the sweep.  Walk the rows in the
// reference's order[]; level(row) = 1 + max(level of the previous row on body1, on body2).
// Rows of one level touch pairwise disjoint bodies, hence disjoint fc[] entries, and a
// friction row's lambda[findex] belongs to a row on the same two bodies (an earlier or later
// level, never the same), so executing level after level — rows of a level in any order or
// in parallel — reads and writes exactly the values the sequential sweep does, bit for bit.
// Islands of one world never share bodies, so their sweeps are scheduled together.
//
// Row record (24 words): J1l[3] J1a[3] J2a[3] (unscaled) | iMJ1a[3] iMJ2a[3] | Ad b Ad*cfm lo hi
// | meta (b1, b2, findex offset) | pad.  Every joint type on the path has J2l == -J1l
// (contact.cpp:86-93, joint.cpp:91-102,146-163, hinge.cpp:101-116) and iMJ*l == invMass*J*l,
// so those values are rebuilt per use with the reference's own multiplications.
#pragma once
#include <cuda_runtime.h>
#include "ob_rows.h"
#include "ob_solver.h"

#define OB_ROWF 20                                   // reals per row record
#define OB_ROWW (OB_ROWF + 16 / (int)sizeof(real))   // reals per row incl. the 16-byte meta chunk
#define OB_MAXEPOCH 8                                // shuffle epochs per step ((iters+7)/8)

__host__ __device__ inline size_t ob_al(size_t x, size_t a) { return (x + a - 1) & ~(a - 1); }

struct PrepTileSmem {   // byte offsets inside one world's shared-memory slice (k_prep)
  size_t invM, rowb, ord, lvl, X, jb1, jb2, adjstart, cursor, adj, btag, jtag, stack, ibody, ijoint, jrow, isz, last, misc, total;
};
__host__ __device__ inline PrepTileSmem prep_tile_smem(int NB, int NC, int NR) {
  PrepTileSmem s; size_t o = 0;
  s.invM = o; o = ob_al(o + sizeof(real) * NB, 16);
  s.rowb = o; o = ob_al(o + 2 * (size_t)NR, 16);                       // b1,b2 per row
  s.ord = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);
  s.lvl = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);           // swap indices, then level per position
  s.X = o; o = ob_al(o + sizeof(unsigned short) * (NR + 2), 16);       // rows per level -> level ends
  s.jb1 = o; o = ob_al(o + (size_t)NC, 4);
  s.jb2 = o; o = ob_al(o + (size_t)NC, 4);
  s.adjstart = o; o = ob_al(o + sizeof(unsigned short) * (NB + 1), 4);
  s.cursor = o; o = ob_al(o + sizeof(unsigned short) * NB, 4);
  s.adj = o; o = ob_al(o + sizeof(unsigned short) * 2 * NC, 4);
  s.btag = o; o = ob_al(o + (size_t)NB, 4);
  s.jtag = o; o = ob_al(o + (size_t)NC, 4);
  s.stack = o; o = ob_al(o + (size_t)NB, 4);
  s.ibody = o; o = ob_al(o + (size_t)NB, 4);
  s.ijoint = o; o = ob_al(o + sizeof(unsigned short) * NC, 4);
  s.jrow = o; o = ob_al(o + sizeof(unsigned short) * (NC + 1), 4);
  s.isz = o; o = ob_al(o + sizeof(unsigned short) * 4 * NB, 4);
  s.last = o; o = ob_al(o + sizeof(unsigned short) * NB, 4);
  s.misc = o; o = ob_al(o + sizeof(int) * 8, 16);
  s.total = ob_al(o, 16);
  return s;
}
struct SorTileSmem { size_t fc, lam, invM, total; };
__host__ __device__ inline SorTileSmem sor_tile_smem(int NB, int NR) {
  SorTileSmem s; size_t o = 0;
  s.fc = o; o = ob_al(o + sizeof(real) * 8 * NB, 16);
  s.lam = o; o = ob_al(o + sizeof(real) * NR, 16);
  s.invM = o; o = ob_al(o + sizeof(real) * NB, 16);
  s.total = ob_al(o, 16);
  return s;
}
struct PostTileSmem { size_t moved, flag, old, misc, total; };
__host__ __device__ inline PostTileSmem post_tile_smem(int NG) {
  PostTileSmem s; size_t o = 0;
  s.moved = o; o = ob_al(o + (size_t)NG, 4);
  s.flag = o; o = ob_al(o + (size_t)NG, 4);
  s.old = o; o = ob_al(o + sizeof(unsigned short) * NG, 4);
  s.misc = o; o = ob_al(o + sizeof(int) * 4, 16);
  s.total = ob_al(o, 16);
  return s;
}

__device__ __forceinline__ int warp_max_i(int v) {
  for (int d = 16; d; d >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
  return v;
}

struct ObRowReg {   // one row in registers
  real v[OB_ROWF];
  unsigned meta;    // b1 | b2<<8 | findex offset<<16
};
__device__ __forceinline__ void load_row(const real *__restrict__ p, ObRowReg &r) {
#if defined(dSINGLE)
  const float4 *q = (const float4 *)p;
#pragma unroll
  for (int i = 0; i < 5; i++) { const float4 t = __ldg(q + i); r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w; }
  r.meta = __ldg((const unsigned *)(q + 5));
#else
  const double2 *q = (const double2 *)p;
#pragma unroll
  for (int i = 0; i < 10; i++) { const double2 t = __ldg(q + i); r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y; }
  r.meta = __ldg((const unsigned *)(q + 10));
#endif
}
__device__ __forceinline__ void store_row(real *p, const real *rw, unsigned meta) {
#if defined(dSINGLE)
  float4 *q = (float4 *)p;
#pragma unroll
  for (int i = 0; i < 5; i++) q[i] = make_float4(rw[4 * i], rw[4 * i + 1], rw[4 * i + 2], rw[4 * i + 3]);
  ((uint4 *)q)[5] = make_uint4(meta, 0u, 0u, 0u);
#else
  double2 *q = (double2 *)p;
#pragma unroll
  for (int i = 0; i < 10; i++) q[i] = make_double2(rw[2 * i], rw[2 * i + 1]);
  ((uint4 *)q)[10] = make_uint4(meta, 0u, 0u, 0u);
#endif
}

// per-world hand-off between the three kernels (ObBatchDev::stepinfo), ints
enum { SI_NIS = 0, SI_NIB, SI_NIJ, SI_MTOT, SI_HAVEROWS, SI_NPASS0, SI_WORDS = SI_NPASS0 + OB_MAXEPOCH };

// =====================================================================================
template <int G>
__global__ void __launch_bounds__(32) k_prep(ObBatchDev d, real h, int taps) {
  constexpr int T = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const PrepTileSmem L = prep_tile_smem(d.NB, d.NC, d.NR);
  const int lane = threadIdx.x, grp = lane / G, gl = lane % G;
  const unsigned FULL = 0xffffffffu;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  real *s_invM = (real *)(smem + L.invM);
  unsigned char *s_rowb = smem + L.rowb;
  unsigned short *s_ord = (unsigned short *)(smem + L.ord);
  unsigned short *s_lvl = (unsigned short *)(smem + L.lvl);
  unsigned short *s_X = (unsigned short *)(smem + L.X);
  unsigned char *s_jb1 = smem + L.jb1, *s_jb2 = smem + L.jb2;
  unsigned short *s_adjstart = (unsigned short *)(smem + L.adjstart);
  unsigned short *s_cursor = (unsigned short *)(smem + L.cursor);
  unsigned short *s_adj = (unsigned short *)(smem + L.adj);
  signed char *s_btag = (signed char *)(smem + L.btag);
  signed char *s_jtag = (signed char *)(smem + L.jtag);
  unsigned char *s_stack = smem + L.stack;
  unsigned char *s_ibody = smem + L.ibody;
  unsigned short *s_ijoint = (unsigned short *)(smem + L.ijoint);
  unsigned short *s_jrow = (unsigned short *)(smem + L.jrow);
  unsigned short *s_isz = (unsigned short *)(smem + L.isz);
  unsigned short *s_last = (unsigned short *)(smem + L.last);
  int *s_misc = (int *)(smem + L.misc);
  const real stepsize1 = ob_recip(h);

  for (int wbase = blockIdx.x * T; wbase < d.W; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.W;
    const int wc = valid ? w : 0;
    ObWorld &W = d.world[wc];
    const int nb = valid ? W.nb : 0;
    const int nc = valid ? d.ncontacts[wc] : 0;
    const int iters = W.iters;
    ObBodyDyn *bd = d.bdyn + (size_t)wc * d.NB;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const ObGeom *geoms = d.geom + (size_t)wc * d.NG;
    const ObContact *con = d.contacts + (size_t)wc * d.NC;
    real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    real *g_invIw = d.invIw + (size_t)wc * d.NB * 12;
    real *g_tmp1 = d.tmp1 + (size_t)wc * d.NB * 8;
    int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nb_max = warp_max_i(nb), nc_max = warp_max_i(nc);

    // ---- (1) graph: contact joint -> bodies (dJointAttach swap rule, ode.cpp:1368-1377)
    for (int b = gl; b < nb; b += G) { s_cursor[b] = 0; s_btag[b] = 0; }
    __syncwarp();
    for (int base = 0; base < nc_max; base += G) {
      const int j = base + gl;
      if (j < nc) {
        int b1 = geoms[con[j].g1].body, b2 = geoms[con[j].g2].body;
        if (b1 < 0) { b1 = b2; b2 = -1; }
        s_jb1[j] = (unsigned char)b1; s_jb2[j] = (unsigned char)(b2 < 0 ? 255 : b2); s_jtag[j] = 0;
      }
    }
    __syncwarp();
    // per-body joint lists, newest joint first (one lane per world)
    if (gl == 0 && valid) {
      for (int j = 0; j < nc; j++) { s_cursor[s_jb1[j]]++; if (s_jb2[j] != 255) s_cursor[s_jb2[j]]++; }
      int a = 0;
      for (int b = 0; b < nb; b++) { const int c = s_cursor[b]; s_adjstart[b] = (unsigned short)a; s_cursor[b] = (unsigned short)a; a += c; }
      s_adjstart[nb] = (unsigned short)a;
      for (int j = nc - 1; j >= 0; j--) {
        const int b1 = s_jb1[j], b2 = s_jb2[j];
        s_adj[s_cursor[b1]++] = (unsigned short)j;
        if (b2 != 255) s_adj[s_cursor[b2]++] = (unsigned short)j;
      }
    }
    __syncwarp();
    // ---- (2) auto-disable, instantaneous-sample mode (util.cpp:99-233); invMass to smem
    for (int base = 0; base < nb_max; base += G) {
      const int b = base + gl;
      if (b < nb) {
        const ObBodyConst &C = bc[b];
        s_invM[b] = C.invMass;
        ObBodyDyn &B = bd[b];
        const unsigned fl = B.flags;
        if (s_adjstart[b + 1] != s_adjstart[b] && (fl & (OB_BODY_AUTO_DISABLE | OB_BODY_DISABLED)) == OB_BODY_AUTO_DISABLE &&
            C.adis_samples != 0) {
          int idle = 1;
          const real ls = ob_dot(B.lvel, B.lvel);
          if (ls > C.adis_lin_thr) idle = 0;
          else { const real as = ob_dot(B.avel, B.avel); if (as > C.adis_ang_thr) idle = 0; }
          if (idle) { B.adis_stepsleft--; B.adis_timeleft -= h; }
          else { B.adis_stepsleft = C.adis_idle_steps; B.adis_timeleft = C.adis_idle_time; }
          if (B.adis_stepsleft <= 0 && B.adis_timeleft <= 0) {
            B.flags = fl | OB_BODY_DISABLED;
            for (int k = 0; k < 3; k++) { B.lvel[k] = 0; B.avel[k] = 0; }
          }
        }
        if (B.flags & OB_BODY_DISABLED) s_btag[b] = -2;   // remembered so the DFS needs no global reads
      }
    }
    __syncwarp();
    // ---- (3) islands: the reference's DFS (util.cpp:411-487), one lane per world
    if (gl == 0 && valid) {
      int nib = 0, nij = 0, nis = 0;
      for (int bb = 0; bb < nb; bb++) {
        const int t0 = s_btag[bb];
        if (t0 == 1 || t0 == -1) continue;
        if (t0 == -2) { s_btag[bb] = -1; continue; }
        s_btag[bb] = 1;
        const int b0 = nib, j0 = nij;
        s_ibody[nib++] = (unsigned char)bb;
        int sp = 0, b = bb;
        while (true) {
          const int e = s_adjstart[b + 1];
          for (int k = s_adjstart[b]; k < e; k++) {
            const int j = s_adj[k];
            if (!s_jtag[j]) {
              const int j1 = s_jb1[j], j2 = s_jb2[j];
              const bool enabled = (s_invM[j1] > 0) || (j2 != 255 && s_invM[j2] > 0);
              if (enabled) {
                s_jtag[j] = 1;
                s_ijoint[nij++] = (unsigned short)j;
                const int other = (j1 == b) ? j2 : j1;
                if (other != 255 && s_btag[other] != 1) {
                  if (s_btag[other] < 0) bd[other].flags &= ~OB_BODY_DISABLED;   // re-enable (util.cpp:447-451)
                  s_btag[other] = 1;
                  s_stack[sp++] = (unsigned char)other;
                }
              } else s_jtag[j] = -1;
            }
          }
          if (sp == 0) break;
          b = s_stack[--sp];
          s_ibody[nib++] = (unsigned char)b;
        }
        s_isz[4 * nis + 0] = (unsigned short)b0; s_isz[4 * nis + 1] = (unsigned short)(nib - b0);
        s_isz[4 * nis + 2] = (unsigned short)j0; s_isz[4 * nis + 3] = (unsigned short)(nij - j0);
        nis++;
      }
      s_misc[0] = nis; s_misc[1] = nib; s_misc[2] = nij;
    }
    __syncwarp();
    const int nis = valid ? s_misc[0] : 0, nib = valid ? s_misc[1] : 0, nij = valid ? s_misc[2] : 0;
    const int nib_max = warp_max_i(nib), nij_max = warp_max_i(nij), nis_max = warp_max_i(nis);

    // ---- (4) per-body preamble (quickstep.cpp:610-665) + tmp1 (:840-846), island bodies only
    unsigned char *g_ibody = d.ibody + (size_t)wc * d.NB;
    for (int base = 0; base < nib_max; base += G) {
      const int i = base + gl;
      if (i < nib) {
        const int b = s_ibody[i];
        g_ibody[i] = (unsigned char)b;
        ObBodyDyn &B = bd[b];
        const ObBodyConst &C = bc[b];
        real R[12], I[12], invI[12], iw[12], avel[3], lvel[3], facc[3], tacc[3], t1[6];
        for (int k = 0; k < 12; k++) { R[k] = B.R[k]; I[k] = C.I[k]; invI[k] = C.invI[k]; }
        for (int k = 0; k < 3; k++) { avel[k] = B.avel[k]; lvel[k] = B.lvel[k]; facc[k] = B.facc[k]; tacc[k] = B.tacc[k]; }
        ob_body_preamble(R, I, invI, avel, B.flags, C.mass, W.gravity, iw, facc, tacc);
        ob_body_tmp1(facc, tacc, lvel, avel, C.invMass, iw, stepsize1, t1);
        for (int k = 0; k < 3; k++) { B.facc[k] = facc[k]; B.tacc[k] = tacc[k]; }
        for (int k = 0; k < 12; k++) g_invIw[12 * b + k] = iw[k];
        for (int k = 0; k < 6; k++) g_tmp1[8 * b + k] = t1[k];
      }
    }
    // ---- (5) rows per joint (getInfo1) -> row offsets in island joint order (tile-wide scan)
    const ObSurface surf0 = d.policy[0].surface;
    int mtot = 0;
    for (int base = 0; base < nij_max; base += G) {
      const int k = base + gl;
      int m = 0;
      if (k < nij) { ObSurface sf = surf0; m = ob_contact_info1(sf); }
      int x = m;
      for (int dd = 1; dd < G; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd, G); if (gl >= dd) x += y; }
      if (k < nij) s_jrow[k] = (unsigned short)(mtot + x - m);
      mtot += __shfl_sync(FULL, x, G - 1, G);
    }
    if (gl == 0 && valid) s_jrow[nij] = (unsigned short)mtot;
    __syncwarp();
    if (mtot > d.NR) {   // capacity: solve nothing rather than corrupt memory; flagged per world
      if (gl == 0) atomicOr(&W.status, OB_ERR_ROW_OVERFLOW);
      mtot = 0;
    }
    const bool have_rows = mtot > 0;
    // ---- (6) row assembly (getInfo2) + finalisation, one lane per joint
    unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + 1);
    unsigned short *g_ijoint = d.ijoint + (size_t)wc * d.NC;
    for (int base = 0; base < nij_max; base += G) {
      const int k = base + gl;
      if (k < nij && have_rows) {
        const int j = s_ijoint[k];
        ObSurface sf = surf0;
        const int jm = ob_contact_info1(sf);
        ObRowOut3 r;
        ob_rows_defaults(r, jm, W.cfm);
        const ObContact c = con[j];
        const int b1 = s_jb1[j], b2 = s_jb2[j] == 255 ? -1 : (int)s_jb2[j];
        const int rev = geoms[c.g1].body < 0;
        real p1[3], l1[3], a1[3], p2[3] = {0, 0, 0}, l2[3] = {0, 0, 0}, a2[3] = {0, 0, 0};
        for (int e = 0; e < 3; e++) { p1[e] = bd[b1].pos[e]; l1[e] = bd[b1].lvel[e]; a1[e] = bd[b1].avel[e]; }
        if (b2 >= 0) for (int e = 0; e < 3; e++) { p2[e] = bd[b2].pos[e]; l2[e] = bd[b2].lvel[e]; a2[e] = bd[b2].avel[e]; }
        const real fdir1[3] = {0, 0, 0};
        ob_contact_info2(r, jm, sf, c.pos, c.normal, c.depth, fdir1, rev, p1, l1, a1, b2 >= 0, p2, l2, a2, stepsize1,
                         W.erp, W.min_depth, W.max_vel);
        real t1a[6], t1b[6], iw1[12], iw2[12];
        for (int e = 0; e < 6; e++) t1a[e] = __ldcg(g_tmp1 + 8 * b1 + e);
        for (int e = 0; e < 12; e++) iw1[e] = __ldcg(g_invIw + 12 * b1 + e);
        if (b2 >= 0) {
          for (int e = 0; e < 6; e++) t1b[e] = __ldcg(g_tmp1 + 8 * b2 + e);
          for (int e = 0; e < 12; e++) iw2[e] = __ldcg(g_invIw + 12 * b2 + e);
        }
        const int r0 = s_jrow[k];
        const real invM1 = s_invM[b1], invM2 = b2 >= 0 ? s_invM[b2] : (real)0;
        for (int q = 0; q < jm; q++) {
          const int ri = r0 + q;
          real rw[OB_ROWF];
          for (int e = 0; e < 6; e++) rw[e] = r.J[q][e];
          for (int e = 0; e < 3; e++) rw[6 + e] = r.J[q][9 + e];
          real iMJ[12], b_out, adcfm, Ad;
          ob_row_finalize2(r.J[q], r.c[q], r.cfm[q], b2, t1a, t1b, invM1, iw1, invM2, iw2, stepsize1, W.sor_w, iMJ, &b_out,
                           &adcfm, &Ad);
          for (int e = 0; e < 3; e++) { rw[9 + e] = iMJ[3 + e]; rw[12 + e] = iMJ[9 + e]; }
          rw[15] = Ad; rw[16] = b_out; rw[17] = adcfm; rw[18] = r.lo[q]; rw[19] = r.hi[q];
          const unsigned fio = (unsigned)(r.findex[q] >= 0 ? q - r.findex[q] : 0);
          const unsigned ub2 = (unsigned)(b2 < 0 ? 255 : b2);
          store_row(rows + (size_t)ri * OB_ROWW, rw, (unsigned)b1 | (ub2 << 8) | (fio << 16));
          s_rowb[2 * ri] = (unsigned char)b1; s_rowb[2 * ri + 1] = (unsigned char)ub2;
          s_ord[ri] = (unsigned short)fio;   // parked: "has findex" flag for the initial order below
        }
      }
      if (k < nij) { g_ijoint[k] = s_ijoint[k]; g_jrow[k] = s_jrow[k]; }
    }
    if (gl == 0 && valid) g_jrow[nij] = (unsigned short)(have_rows ? mtot : 0);
    __syncwarp();

    // ---- (7) per shuffle epoch: the reference's row order, then the level schedule
    const int nep = (iters + 7) >> 3;
    const uint32_t seed = W.seed;
    if (gl == 0) s_misc[5] = 0;
    // initial order per island (quickstep.cpp:409-424): findex==-1 rows ascending at the head,
    // the others descending at the tail.  s_ord currently holds the per-row findex flag; the
    // result goes to s_lvl first, then is copied back (one lane per world).
    if (gl == 0 && have_rows) {
      for (int isl = 0; isl < nis; isl++) {
        const int j0 = s_isz[4 * isl + 2], jn = s_isz[4 * isl + 3];
        if (!jn) continue;
        const int r0 = s_jrow[j0], m = s_jrow[j0 + jn] - r0;
        int head = 0, tail = m - 1;
        for (int i = 0; i < m; i++) { if (s_ord[r0 + i] == 0) s_lvl[r0 + head++] = (unsigned short)i; else s_lvl[r0 + tail--] = (unsigned short)i; }
      }
      for (int i = 0; i < mtot; i++) s_ord[i] = s_lvl[i];
    }
    __syncwarp();
    const int mtot_max = warp_max_i(mtot);
    const int nep_max = warp_max_i(valid ? nep : 0);
    for (int ep = 0; ep < nep_max && ep < d.NEP; ep++) {
      const bool epv = have_rows && ep < nep;
      // (a) shuffle every island's segment (quickstep.cpp:474-481).  The reference runs ALL
      // iterations of island 0 before island 1, so the draws of (island i, epoch e) start at
      // offset  sum_{j<i} nep*(m_j-1) + e*(m_i-1)  of the world's LCG stream.
      unsigned draws_before = 0;
      for (int isl = 0; isl < nis_max; isl++) {
        int r0 = 0, m = 0;
        if (isl < nis && have_rows) {
          const int j0 = s_isz[4 * isl + 2], jn = s_isz[4 * isl + 3];
          if (jn) { r0 = s_jrow[j0]; m = s_jrow[j0 + jn] - r0; }
        }
        const unsigned my_off = draws_before + (unsigned)ep * (unsigned)(m >= 2 ? m - 1 : 0);
        if (m >= 2) draws_before += (unsigned)nep * (unsigned)(m - 1);
        if (!epv) m = 0;
        const int m_max = warp_max_i(m);
        if (m_max < 2) continue;
        {   // swap indices for i = 1..m-1 by all lanes (misc.cpp:66-117), LCG skip-ahead by G
          uint32_t A, C, A0, C0;
          ob_lcg_skip(my_off, &A0, &C0);
          uint32_t s = A0 * seed + C0;
          ob_lcg_skip((uint32_t)G, &A, &C);
          for (int q = 0; q <= gl; q++) s = ob_lcg_next(s);
          for (int base = 1; base < m_max; base += G) {
            const int i = base + gl;
            if (i < m) s_lvl[r0 + i] = (unsigned short)ob_randint_fold(s, (uint32_t)(i + 1));
            s = A * s + C;
          }
        }
        __syncwarp();
        if (gl == 0) {
          unsigned short *ord = s_ord + r0;
          for (int i = 1; i < m; i++) {
            const int sj = s_lvl[r0 + i];
            const unsigned short t = ord[i]; ord[i] = ord[sj]; ord[sj] = t;
          }
        }
        __syncwarp();
      }
      if (ep == 0) s_misc[5] = (int)draws_before;   // total draws of this step (same every epoch)
      // (b) level of every position, walking islands and positions in sweep order (one lane)
      int nlev = 0;
      if (gl == 0 && epv) {
        for (int b = 0; b < nb; b++) s_last[b] = 0;
        for (int i = 0; i <= mtot + 1; i++) s_X[i] = 0;
        for (int isl = 0; isl < nis; isl++) {
          const int j0 = s_isz[4 * isl + 2], jn = s_isz[4 * isl + 3];
          if (!jn) continue;
          const int r0 = s_jrow[j0], m = s_jrow[j0 + jn] - r0;
          for (int k = 0; k < m; k++) {
            const int idx = r0 + s_ord[r0 + k];
            const int b1 = s_rowb[2 * idx], b2 = s_rowb[2 * idx + 1];
            int lv = s_last[b1];
            if (b2 != 255) { const int l2 = s_last[b2]; lv = l2 > lv ? l2 : lv; }
            lv++;
            if (taps & 2) lv = nlev + 1;   // debug: strictly sequential schedule (one row per level)
            s_last[b1] = (unsigned short)lv;
            if (b2 != 255) s_last[b2] = (unsigned short)lv;
            s_lvl[r0 + k] = (unsigned short)lv;      // levels are 1-based
            s_X[lv]++;
            nlev = lv > nlev ? lv : nlev;
          }
        }
        // exclusive prefix: s_X[l] = first slot of level l (l = 1..nlev)
        int a = 0;
        for (int l = 1; l <= nlev; l++) { const int c = s_X[l]; s_X[l] = (unsigned short)a; a += c; }
        s_misc[4] = nlev;
      }
      __syncwarp();
      nlev = epv ? s_misc[4] : 0;
      // (c) scatter rows into level order (any order inside a level), then the pass table
      unsigned short *g_sched = d.sched + ((size_t)wc * d.NEP + ep) * d.NR;
      unsigned short *g_pstart = d.pstart + ((size_t)wc * d.NEP + ep) * (d.NR + 1);
      if (gl == 0 && epv) {
        // sequential scatter (keeps s_X as "end of level" afterwards) and pass table:
        // a pass is a chunk of <= G consecutive slots that does not cross a level boundary
        for (int isl = 0; isl < nis; isl++) {
          const int j0 = s_isz[4 * isl + 2], jn = s_isz[4 * isl + 3];
          if (!jn) continue;
          const int r0 = s_jrow[j0], m = s_jrow[j0 + jn] - r0;
          for (int k = 0; k < m; k++) {
            const int lv = s_lvl[r0 + k];
            const int pos = s_X[lv]++;
            g_sched[pos] = (unsigned short)(r0 + s_ord[r0 + k]);
          }
        }
        int np = 0, start = 0;
        for (int l = 1; l <= nlev; l++) {
          const int end = s_X[l];
          for (int p = start; p < end; p += G) g_pstart[np++] = (unsigned short)p;
          start = end;
        }
        g_pstart[np] = (unsigned short)start;
        si[SI_NPASS0 + ep] = np;
      }
      if (gl == 0 && valid && !epv) si[SI_NPASS0 + ep] = 0;
      __syncwarp();
    }
    (void)mtot_max;
    if (gl == 0 && valid) {
      { uint32_t A, C; ob_lcg_skip(have_rows && nep > 0 ? (unsigned)s_misc[5] : 0u, &A, &C); W.seed = A * seed + C; }
      si[SI_NIS] = nis; si[SI_NIB] = nib; si[SI_NIJ] = nij; si[SI_MTOT] = mtot; si[SI_HAVEROWS] = have_rows ? 1 : 0;
      unsigned short *g_isz = d.isz + (size_t)wc * 4 * d.NB;
      for (int i = 0; i < 4 * nis; i++) g_isz[i] = s_isz[i];
    }
    __syncwarp();
  }
}

// =====================================================================================
template <int G>
__global__ void __launch_bounds__(32) k_sor(ObBatchDev d, int taps) {
  constexpr int T = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const SorTileSmem L = sor_tile_smem(d.NB, d.NR);
  const int lane = threadIdx.x, grp = lane / G, gl = lane % G;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  real *s_fc = (real *)(smem + L.fc);
  real *s_lam = (real *)(smem + L.lam);
  real *s_invM = (real *)(smem + L.invM);

  for (int wbase = blockIdx.x * T; wbase < d.W; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.W;
    const int wc = valid ? w : 0;
    const int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nb = valid ? d.world[wc].nb : 0;
    const int iters = valid ? d.world[wc].iters : 0;
    const int mtot = valid && si[SI_HAVEROWS] ? si[SI_MTOT] : 0;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    for (int b = gl; b < nb; b += G) {
      s_invM[b] = bc[b].invMass;
      for (int k = 0; k < 8; k++) s_fc[8 * b + k] = 0;
    }
    for (int i = gl; i < mtot; i += G) s_lam[i] = 0;
    __syncwarp();
    const int iters_max = warp_max_i(mtot > 0 ? iters : 0);
    for (int it = 0; it < iters_max; it++) {
      const int ep = it >> 3;
      const bool itv = mtot > 0 && it < iters;
      const unsigned short *sched = d.sched + ((size_t)wc * d.NEP + ep) * d.NR;
      const unsigned short *pstart = d.pstart + ((size_t)wc * d.NEP + ep) * (d.NR + 1);
      const int np = itv ? si[SI_NPASS0 + ep] : 0;
      const int np_max = warp_max_i(np);
      // software pipeline: row of pass p+1 is fetched while pass p computes
      ObRowReg cur, nxt;
      int cur_idx = -1, nxt_idx = -1;
      if (np > 0) {
        const int p0 = pstart[0], p1 = pstart[1];
        if (p0 + gl < p1) { nxt_idx = sched[p0 + gl]; load_row(rows + (size_t)nxt_idx * OB_ROWW, nxt); }
      }
      for (int p = 0; p < np_max; p++) {
        cur = nxt; cur_idx = nxt_idx;
        nxt_idx = -1;
        if (p + 1 < np) {
          const int q0 = pstart[p + 1], q1 = pstart[p + 2];
          if (q0 + gl < q1) { nxt_idx = sched[q0 + gl]; load_row(rows + (size_t)nxt_idx * OB_ROWW, nxt); }
        }
        const bool act = p < np && cur_idx >= 0;
        if (taps & 8) __syncwarp();
        if (taps & 4) {   // debug: verify that the rows of this pass touch pairwise disjoint bodies
          const int mb1 = act ? (int)(cur.meta & 255) : -1, mb2 = act ? (int)((cur.meta >> 8) & 255) : -1;
          for (int l2 = 0; l2 < G; l2++) {
            const int o1 = __shfl_sync(0xffffffffu, mb1, l2, G), o2 = __shfl_sync(0xffffffffu, mb2, l2, G);
            if (act && l2 != gl && o1 >= 0) {
              if (mb1 == o1 || mb1 == o2 || (mb2 != 255 && (mb2 == o1 || mb2 == o2))) atomicOr(&d.world[wc].status, 256);
            }
          }
        }
        if (act) {
          const int b1 = cur.meta & 255, b2r = (cur.meta >> 8) & 255, fio = (cur.meta >> 16) & 255;
          const int b2 = b2r == 255 ? -1 : b2r;
          const int fi = fio ? cur_idx - fio : -1;
          const real Ad = cur.v[15], k1 = s_invM[b1];
          real J[12], iMJ[12];
          // rebuild what SOR_LCP keeps per row: J scaled by Ad (quickstep.cpp:393-401), iMJ (:117-136)
#pragma unroll
          for (int e = 0; e < 3; e++) {
            iMJ[e] = k1 * cur.v[e];
            iMJ[3 + e] = cur.v[9 + e];
            J[e] = cur.v[e] * Ad;
            J[3 + e] = cur.v[3 + e] * Ad;
          }
          real f1[6], f2[6];
          real *fp1 = s_fc + 8 * b1;
#if defined(dSINGLE)
          { const float4 a = *(const float4 *)fp1; const float2 c = *(const float2 *)(fp1 + 4); f1[0] = a.x; f1[1] = a.y; f1[2] = a.z; f1[3] = a.w; f1[4] = c.x; f1[5] = c.y; }
#else
          for (int e = 0; e < 6; e++) f1[e] = fp1[e];
#endif
          real *fp2 = s_fc;
          if (b2 >= 0) {
            const real k2 = s_invM[b2];
#pragma unroll
            for (int e = 0; e < 3; e++) {
              const real j2l = -cur.v[e];
              iMJ[6 + e] = k2 * j2l;
              iMJ[9 + e] = cur.v[12 + e];
              J[6 + e] = j2l * Ad;
              J[9 + e] = cur.v[6 + e] * Ad;
            }
            fp2 = s_fc + 8 * b2;
#if defined(dSINGLE)
            { const float4 a = *(const float4 *)fp2; const float2 c = *(const float2 *)(fp2 + 4); f2[0] = a.x; f2[1] = a.y; f2[2] = a.z; f2[3] = a.w; f2[4] = c.x; f2[5] = c.y; }
#else
            for (int e = 0; e < 6; e++) f2[e] = fp2[e];
#endif
          }
          const real lam_new = ob_sor_row(J, iMJ, cur.v[16], cur.v[17], cur.v[18], cur.v[19], fi, fi >= 0 ? s_lam[fi] : (real)0,
                                          s_lam[cur_idx], f1, b2 >= 0 ? f2 : (real *)0);
          s_lam[cur_idx] = lam_new;
#if defined(dSINGLE)
          *(float4 *)fp1 = make_float4(f1[0], f1[1], f1[2], f1[3]); *(float2 *)(fp1 + 4) = make_float2(f1[4], f1[5]);
          if (b2 >= 0) { *(float4 *)fp2 = make_float4(f2[0], f2[1], f2[2], f2[3]); *(float2 *)(fp2 + 4) = make_float2(f2[4], f2[5]); }
#else
          for (int e = 0; e < 6; e++) fp1[e] = f1[e];
          if (b2 >= 0) for (int e = 0; e < 6; e++) fp2[e] = f2[e];
#endif
        }
        __syncwarp();
      }
    }
    // cforce per body for k_post; lambda + joint feedback taps (quickstep.cpp:918-957)
    real *g_fc = d.tmp1 + (size_t)wc * d.NB * 8;   // tmp1 is dead after row assembly: reuse as cforce
    for (int b = gl; b < nb; b += G)
      for (int k = 0; k < 6; k++) g_fc[8 * b + k] = s_fc[8 * b + k];
    if (taps && valid) {
      const int nij = si[SI_NIJ];
      const unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + 1);
      const unsigned short *g_ijoint = d.ijoint + (size_t)wc * d.NC;
      real *fb = d.fback + (size_t)wc * d.NC * 6;
      real *gl_lam = d.lambda + (size_t)wc * d.NR;
      if (mtot > 0)
        for (int k = gl; k < nij; k += G) {
          const int jr0 = g_jrow[k], jm = g_jrow[k + 1] - jr0;
          real acc[6] = {0, 0, 0, 0, 0, 0};
          for (int q = 0; q < jm; q++) {
            const real s = s_lam[jr0 + q];
            for (int e = 0; e < 6; e++) acc[e] += rows[(size_t)(jr0 + q) * OB_ROWW + e] * s;
          }
          for (int e = 0; e < 6; e++) fb[g_ijoint[k] * 6 + e] = acc[e];
        }
      for (int i = gl; i < mtot; i += G) gl_lam[i] = s_lam[i];
    }
    __syncwarp();
  }
}

// =====================================================================================
template <int G>
__global__ void __launch_bounds__(32) k_post(ObBatchDev d, real h) {
  constexpr int T = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const PostTileSmem L = post_tile_smem(d.NG);
  const int lane = threadIdx.x, grp = lane / G, gl = lane % G;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  unsigned char *s_moved = smem + L.moved;
  unsigned char *s_flag = smem + L.flag;
  unsigned short *s_old = (unsigned short *)(smem + L.old);
  int *s_misc = (int *)(smem + L.misc);

  for (int wbase = blockIdx.x * T; wbase < d.W; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.W;
    const int wc = valid ? w : 0;
    ObWorld &W = d.world[wc];
    const int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int ng = valid ? W.ng : 0;
    const int nis = valid ? si[SI_NIS] : 0, nib = valid ? si[SI_NIB] : 0, nij = valid ? si[SI_NIJ] : 0;
    const int mtot = valid ? si[SI_MTOT] : 0;
    const bool have_rows = valid && si[SI_HAVEROWS];
    ObBodyDyn *bd = d.bdyn + (size_t)wc * d.NB;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const ObGeom *geoms = d.geom + (size_t)wc * d.NG;
    const unsigned char *g_ibody = d.ibody + (size_t)wc * d.NB;
    const unsigned short *g_isz = d.isz + (size_t)wc * 4 * d.NB;
    const unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + 1);
    const real *g_invIw = d.invIw + (size_t)wc * d.NB * 12;
    const real *g_fc = d.tmp1 + (size_t)wc * d.NB * 8;
    // velocity update + integration per island body (quickstep.cpp:905-1021, util.cpp:255-360)
    for (int i = gl; i < nib; i += G) {
      const int b = g_ibody[i];
      bool island_rows = false;
      if (have_rows) {
        for (int p = 0; p < nis; p++) {
          const int pb0 = g_isz[4 * p], pbn = g_isz[4 * p + 1];
          if (i >= pb0 && i < pb0 + pbn) {
            const int pj0 = g_isz[4 * p + 2], pjn = g_isz[4 * p + 3];
            island_rows = pjn && (g_jrow[pj0 + pjn] - g_jrow[pj0]) > 0;
            break;
          }
        }
      }
      ObBodyDyn &B = bd[b];
      const ObBodyConst &C = bc[b];
      real pos[3], q[4], R[12], lvel[3], avel[3], facc[3], tacc[3], iw[12], fcb[6];
      for (int k = 0; k < 3; k++) { pos[k] = B.pos[k]; lvel[k] = B.lvel[k]; avel[k] = B.avel[k]; facc[k] = B.facc[k]; tacc[k] = B.tacc[k]; }
      for (int k = 0; k < 4; k++) q[k] = B.q[k];
      for (int k = 0; k < 12; k++) iw[k] = g_invIw[12 * b + k];
      for (int k = 0; k < 6; k++) fcb[k] = g_fc[8 * b + k];
      ob_body_velocity_update(lvel, avel, island_rows ? fcb : (real *)0, facc, tacc, C.invMass, iw, h);
      real fra[3] = {C.finite_rot_axis[0], C.finite_rot_axis[1], C.finite_rot_axis[2]};
      ob_step_body(pos, q, R, lvel, avel, B.flags, h, C.max_angular_speed, fra, C.damp_lin_scale, C.damp_ang_scale,
                   C.damp_lin_thr, C.damp_ang_thr);
      for (int k = 0; k < 3; k++) { B.pos[k] = pos[k]; B.lvel[k] = lvel[k]; B.avel[k] = avel[k]; }
      for (int k = 0; k < 4; k++) { B.q[k] = q[k]; B.facc[k] = 0; B.tacc[k] = 0; }
      for (int k = 0; k < 12; k++) B.R[k] = R[k];
    }
    // space list: every geom of a stepped body moves to the head, in stepping order
    int *glist = d.glist + (size_t)wc * d.NG;
    if (gl == 0 && valid) {
      int nm = 0;
      for (int i = 0; i < nib; i++)
        for (int g = bc[g_ibody[i]].geom_first; g >= 0; g = geoms[g].body_next) s_moved[nm++] = (unsigned char)g;
      s_misc[0] = nm;
    }
    for (int g = gl; g < d.NG; g += G) s_flag[g] = 0;
    __syncwarp();
    const int nm = valid ? s_misc[0] : 0;
    for (int i = gl; i < nm; i += G) s_flag[s_moved[i]] = 1;
    for (int i = gl; i < ng; i += G) s_old[i] = (unsigned short)glist[i];
    __syncwarp();
    for (int i = gl; i < ng; i += G) {
      const int g = s_old[i];
      if (!s_flag[g]) {
        int before = 0;
        for (int j2 = 0; j2 < i; j2++) before += s_flag[s_old[j2]] ? 0 : 1;
        glist[nm + before] = g;
      }
    }
    for (int i = gl; i < nm; i += G) glist[nm - 1 - i] = s_moved[i];
    if (gl == 0 && valid) {
      d.nrows[w] = mtot;
      atomicAdd(&d.counters->steps, 1ull);
      atomicAdd(&d.counters->body_steps, (unsigned long long)nib);
      atomicAdd(&d.counters->contacts, (unsigned long long)(have_rows ? nij : 0));
      atomicAdd(&d.counters->rows, (unsigned long long)mtot);
      atomicAdd(&d.counters->islands, (unsigned long long)nis);
      if (W.status) atomicAdd(&d.counters->overflow_worlds, 1ull);
    }
    __syncwarp();
  }
}

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
// Row record (20 words): J1l[3] J1a[3] J2a[3] (unscaled) | iMJ1a[3] iMJ2a[3] | Ad b Ad*cfm bound
// | meta (b1, b2, findex offset, bound mode: lo=-hi / lo=0 / hi=0).  Every joint type on the path has J2l == -J1l
// (contact.cpp:86-93, joint.cpp:91-102,146-163, hinge.cpp:101-116) and iMJ*l == invMass*J*l,
// so those values are rebuilt per use with the reference's own multiplications.
#pragma once
#include <cuda_runtime.h>
#include "ob_rows.h"
#include "ob_solver.h"


__host__ __device__ inline size_t ob_al(size_t x, size_t a) { return (x + a - 1) & ~(a - 1); }

struct PrepTileSmem {   // byte offsets inside one world's shared-memory slice (k_prep)
  size_t invM, erpsrc, jb1, jb2, adjstart, cursor, adj, btag, jtag, stack, ibody, ijoint, jrow, isz, misc, total;
};
__host__ __device__ inline PrepTileSmem prep_tile_smem(int NB, int NCin, int NJ, int NR) {
  PrepTileSmem s; size_t o = 0;
  const int NC = NCin + NJ;   // joint id space: contacts, then permanent joints
  s.invM = o; o = ob_al(o + sizeof(real) * NB, 16);
  s.erpsrc = o; o = ob_al(o + sizeof(unsigned short) * NC, 4);
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
  unsigned meta;    // b1 | b2<<8 | findex offset<<16 | bound mode<<24
};
__device__ __forceinline__ void load_row(const real *__restrict__ p, ObRowReg &r) {
#if defined(dSINGLE)
  const float4 *q = (const float4 *)p;
#pragma unroll
  for (int i = 0; i < 4; i++) { const float4 t = __ldg(q + i); r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w; }
  const float4 t = __ldg(q + 4);
  r.v[16] = t.x; r.v[17] = t.y; r.v[18] = t.z; r.meta = __float_as_uint(t.w);
#else
  const double2 *q = (const double2 *)p;
#pragma unroll
  for (int i = 0; i < 9; i++) { const double2 t = __ldg(q + i); r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y; }
  const double2 t = __ldg(q + 9);
  r.v[18] = t.x; r.meta = (unsigned)__double2loint(t.y);
#endif
}
__device__ __forceinline__ void store_row(real *p, const real *rw, unsigned meta) {
#if defined(dSINGLE)
  float4 *q = (float4 *)p;
#pragma unroll
  for (int i = 0; i < 4; i++) q[i] = make_float4(rw[4 * i], rw[4 * i + 1], rw[4 * i + 2], rw[4 * i + 3]);
  q[4] = make_float4(rw[16], rw[17], rw[18], __uint_as_float(meta));
#else
  double2 *q = (double2 *)p;
#pragma unroll
  for (int i = 0; i < 9; i++) q[i] = make_double2(rw[2 * i], rw[2 * i + 1]);
  q[9] = make_double2(rw[18], __hiloint2double(0, (int)meta));
#endif
}
// bounds of a row are stored as one value + a mode (all joint types on the path produce one of these)
__device__ __forceinline__ bool encode_bounds(real lo, real hi, real *v, unsigned *mode) {
  if (lo == -hi) { *v = hi; *mode = 0; return true; }
  if (lo == 0) { *v = hi; *mode = 1; return true; }
  if (hi == 0) { *v = lo; *mode = 2; return true; }
  *v = hi; *mode = 0;
  return false;
}


// =====================================================================================
// PJ = false: the batch has no permanent joints (contact joints only), the getInfo1/2 code of every joint type is compiled out
// (tried on B200: capping the contact-only variant at 128 registers for 16 warps per SM instead of 11 -- spills in the row
// assembly, config 2 k_prep 0.43 -> 0.58 ms; kept at 168 registers)
// PH = 0: the whole kernel.  PH = 1 / 2: its two halves as separate launches -- (1) joint graph, islands, body preamble, rows
// per joint and everything k_sched* needs (island / row tables, the rows' body + findex bytes in d.rowmeta); (2) row assembly
// and finalisation -- so that the schedule (k_sched*, which needs only the topology) runs on a second stream WHILE the rows
// are assembled (ob_kern_step.cu).  Half 2 re-stages the few index tables it needs from what half 1 left in global memory.
template <int G, bool PJ, int PH>
__global__ void __launch_bounds__(32) k_prep(ObBatchDev d, real h, int taps) {
  constexpr int T = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const PrepTileSmem L = prep_tile_smem(d.NB, d.NC, d.NJ, d.NR);
  const int lane = threadIdx.x, grp = lane / G, gl = lane % G;
  const unsigned FULL = 0xffffffffu;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  real *s_invM = (real *)(smem + L.invM);
  unsigned short *s_erpsrc = (unsigned short *)(smem + L.erpsrc);
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
  int *s_misc = (int *)(smem + L.misc);
  const real stepsize1 = ob_recip(h);

  for (int wbase = d.wbeg + blockIdx.x * T; wbase < d.wend; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.wend;
    const int wc = valid ? w : 0;
    ObWorld &W = d.world[wc];
    const int nb = valid ? W.nb : 0;
    const int nc = valid ? d.ncontacts[wc] : 0;
    ObBodyDyn *bd = d.bdyn + (size_t)wc * d.NB;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const ObGeom *geoms = d.geom + (size_t)wc * d.NG;
    const ObContact *con = d.contacts + (size_t)wc * d.NC;
    real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    real *g_invIw = d.invIw + (size_t)wc * d.NB * 12;
    real *g_tmp1 = d.tmp1 + (size_t)wc * d.NB * 8;
    int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nb_max = warp_max_i(nb), nc_max = warp_max_i(nc);

    // ---- (1) graph.  Joint id space: this step's contact joints 0..nc-1 (creation order), then the
    // world's permanent joints nc..nc+nj-1.  Contact joint -> bodies by the dJointAttach swap rule
    // (ode.cpp:1368-1377); permanent joints carry their bodies from the host.
    const int nj = valid ? d.njoints[wc] : 0;
    const int njall = nc + nj;
    const ObJoint *pjoint = d.joint + (size_t)wc * (d.NJ ? d.NJ : 1);
    const unsigned short *padjstart = d.padjstart + (size_t)wc * (d.NB + 1);
    const unsigned short *padj = d.padj + (size_t)wc * 2 * (d.NJ ? d.NJ : 1);
    const int njall_max = warp_max_i(njall);
    int nis = 0, nib = 0, nij = 0, nib_max = 0, nij_max = 0, mtot = 0, anyball = 0;
    bool have_rows = false;
    const ObPolicy *ptab = d.policy;   // batched path: a contact carries the row of the policy table that made it (drop-in: per-contact surfaces)
    const ObSurface *csurf = d.csurf ? d.csurf + (size_t)wc * d.NC : (const ObSurface *)0;   // drop-in: per-contact surfaces
    unsigned char *g_ibody = d.ibody + (size_t)wc * d.NB;
    unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + d.NJ + 1);
    unsigned short *g_ijoint = d.ijoint + (size_t)wc * (d.NC + d.NJ);
    unsigned short *g_isz = d.isz + (size_t)wc * 4 * d.NB;
    if constexpr (PH != 2) {
    for (int b = gl; b < nb; b += G) { s_cursor[b] = 0; s_btag[b] = 0; }
    __syncwarp();
    for (int base = 0; base < njall_max; base += G) {
      const int j = base + gl;
      if (j < nc) {
        int b1, b2;
        if (d.dropin) { b1 = con[j].side1; b2 = con[j].side2; }   // drop-in: the caller attached the joint (bodies after the swap rule)
        else { b1 = geoms[con[j].g1].body; b2 = geoms[con[j].g2].body; if (b1 < 0) { b1 = b2; b2 = -1; } }
        s_jb1[j] = (unsigned char)b1; s_jb2[j] = (unsigned char)(b2 < 0 ? 255 : b2); s_jtag[j] = 0;
      } else if (j < njall) {
        const ObJoint &pj = pjoint[j - nc];
        s_jb1[j] = (unsigned char)pj.b1; s_jb2[j] = (unsigned char)(pj.b2 < 0 ? 255 : pj.b2);
        s_jtag[j] = (pj.flags & OB_JF_DISABLED) ? -1 : 0;   // disabled joints are never traversed (joint.cpp:66-71)
      }
    }
    __syncwarp();
    // per-body joint lists: contacts newest first, then the permanent joints in the body's list order
    if (gl == 0 && valid) {
      // a contact between two body-less geoms is a joint attached to nothing (ode.cpp:1348-1399): it is on no body's list
      for (int j = 0; j < nc; j++) { if (s_jb1[j] == 255) continue; s_cursor[s_jb1[j]]++; if (s_jb2[j] != 255) s_cursor[s_jb2[j]]++; }
      int a = 0;
      for (int b = 0; b < nb; b++) {
        const int c = s_cursor[b] + (padjstart[b + 1] - padjstart[b]);
        s_adjstart[b] = (unsigned short)a; s_cursor[b] = (unsigned short)a; a += c;
      }
      s_adjstart[nb] = (unsigned short)a;
      for (int j = nc - 1; j >= 0; j--) {
        const int b1 = s_jb1[j], b2 = s_jb2[j];
        if (b1 == 255) continue;
        s_adj[s_cursor[b1]++] = (unsigned short)j;
        if (b2 != 255) s_adj[s_cursor[b2]++] = (unsigned short)j;
      }
      if (nj)
        for (int b = 0; b < nb; b++)
          for (int k = padjstart[b]; k < padjstart[b + 1]; k++) s_adj[s_cursor[b]++] = (unsigned short)(nc + padj[k]);
    }
    __syncwarp();
    // ---- (2) auto-disable (util.cpp:99-233: instantaneous or averaged samples, ob_auto_disable); invMass to smem
    for (int base = 0; base < nb_max; base += G) {
      const int b = base + gl;
      if (b < nb) {
        const ObBodyConst &C = bc[b];
        s_invM[b] = C.invMass;
        ObBodyDyn &B = bd[b];
        if (s_adjstart[b + 1] != s_adjstart[b]) {   // bodies without joints never fall asleep (util.cpp:104)
          const size_t bi = (size_t)wc * d.NB + b;
          ob_auto_disable(B, C, h, d.NADIS ? d.adisbuf + bi * d.NADIS * 6 : (real *)0, d.NADIS ? d.adisctl + bi * 2 : (int *)0);
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
      nis = valid ? s_misc[0] : 0; nib = valid ? s_misc[1] : 0; nij = valid ? s_misc[2] : 0;
      nib_max = warp_max_i(nib); nij_max = warp_max_i(nij);

    // ---- (4) per-body preamble (quickstep.cpp:610-665), island bodies only
    for (int base = 0; base < nib_max; base += G) {
      const int i = base + gl;
      if (i < nib) {
        const int b = s_ibody[i];
        g_ibody[i] = (unsigned char)b;
        ObBodyDyn &B = bd[b];
        const ObBodyConst &C = bc[b];
        real R[12], I[12], invI[12], iw[12], avel[3], facc[3], tacc[3];
        for (int k = 0; k < 12; k++) { R[k] = B.R[k]; I[k] = C.I[k]; invI[k] = C.invI[k]; }
        for (int k = 0; k < 3; k++) { avel[k] = B.avel[k]; facc[k] = B.facc[k]; tacc[k] = B.tacc[k]; }
        ob_body_preamble(R, I, invI, avel, B.flags, C.mass, W.gravity, iw, facc, tacc);
        for (int k = 0; k < 3; k++) { B.facc[k] = facc[k]; B.tacc[k] = tacc[k]; }
        for (int k = 0; k < 12; k++) g_invIw[12 * b + k] = iw[k];
      }
    }
    // ---- (5) rows per joint (getInfo1) -> row offsets in island joint order (tile-wide scan)
    for (int base = 0; base < nij_max; base += G) {
      const int k = base + gl;
      int m = 0;
      if (k < nij) {
        const int j = s_ijoint[k];
        if (!PJ || j < nc) { ObSurface sf = csurf ? csurf[j] : ptab[con[j].policy].surface; m = ob_contact_info1(sf); }
        else {
          ObJoint pj = pjoint[j - nc];
          const int b1 = pj.b1, b2 = pj.b2;
          ObBodyView B1 = {bd[b1].pos, bd[b1].R, bd[b1].q, bd[b1].lvel, bd[b1].avel}, B2 = B1;
          if (b2 >= 0) { B2.pos = bd[b2].pos; B2.R = bd[b2].R; B2.q = bd[b2].q; B2.lvel = bd[b2].lvel; B2.avel = bd[b2].avel; }
          m = ob_joint_info1(pj, B1, b2 >= 0 ? &B2 : (const ObBodyView *)0);
          if (pj.type == OB_JOINT_BALL || pj.type == OB_JOINT_FIXED) anyball = 1;   // joints that overwrite Info2.erp (ball.cpp:59, fixed.cpp:72)
        }
      }
      int x = m;
      for (int dd = 1; dd < G; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd, G); if (gl >= dd) x += y; }
      if (k < nij) s_jrow[k] = (unsigned short)(mtot + x - m);
      mtot += __shfl_sync(FULL, x, G - 1, G);
    }
    for (int dd = 1; dd < G; dd <<= 1) anyball |= __shfl_xor_sync(FULL, anyball, dd, G);
    if (gl == 0 && valid) s_jrow[nij] = (unsigned short)mtot;
    __syncwarp();
    if (mtot > d.NR) {   // capacity: solve nothing rather than corrupt memory; flagged per world
      if (gl == 0) atomicOr(&W.status, OB_ERR_ROW_OVERFLOW);
      mtot = 0;
    }
    have_rows = mtot > 0;
    if constexpr (PH == 1) {
      // hand-off: island / row tables for half 2 and k_sched*, the rows' body + findex bytes for k_sched*
      unsigned *g_meta = d.rowmeta + (size_t)wc * d.NR;
      for (int base = 0; base < nij_max; base += G) {
        const int k = base + gl;
        if (k < nij) {
          const int j = s_ijoint[k];
          g_ijoint[k] = s_ijoint[k]; g_jrow[k] = s_jrow[k];
          if (have_rows) {
            const int r0 = s_jrow[k], jm = s_jrow[k + 1] - r0;
            const unsigned bb = (unsigned)s_jb1[j] | ((unsigned)s_jb2[j] << 8);
            unsigned mode = 0;
            if (!PJ || j < nc) { const ObSurface sf = csurf ? csurf[j] : ptab[con[j].policy].surface; mode = (unsigned)sf.mode; }
            for (int q = 0; q < jm; q++) {
              // findex offset of a contact's friction rows (contact.cpp:211, :235); every other row has findex -1
              const unsigned fio = (q == 1 && (mode & 0x1000u)) ? 1u : ((q == 2 && (mode & 0x2000u)) ? 2u : 0u);
              g_meta[r0 + q] = bb | (fio << 16);
            }
          }
        }
      }
      if (gl == 0 && valid) {
        g_jrow[nij] = (unsigned short)(have_rows ? mtot : 0);
        si[SI_NIS] = nis; si[SI_NIB] = nib; si[SI_NIJ] = nij; si[SI_MTOT] = mtot; si[SI_HAVEROWS] = have_rows ? 1 : 0;
        si[SI_ANYBALL] = anyball;
        for (int i = 0; i < 4 * nis; i++) g_isz[i] = s_isz[i];
      }
      __syncwarp();
      continue;
    }
    } else {
      // half 2: re-stage what half 1 computed
      nis = valid ? si[SI_NIS] : 0; nib = valid ? si[SI_NIB] : 0; nij = valid ? si[SI_NIJ] : 0;
      mtot = valid ? si[SI_MTOT] : 0; have_rows = valid && si[SI_HAVEROWS] != 0; anyball = valid ? si[SI_ANYBALL] : 0;
      anyball = __shfl_sync(FULL, anyball, 0, G);
      nib_max = warp_max_i(nib); nij_max = warp_max_i(nij);
      for (int b = gl; b < nb; b += G) s_invM[b] = bc[b].invMass;
      for (int i = gl; i < nib; i += G) s_ibody[i] = g_ibody[i];
      for (int k = gl; k <= nij; k += G) { s_jrow[k] = g_jrow[k]; if (k < nij) s_ijoint[k] = g_ijoint[k]; }
      for (int i = gl; i < 4 * nis; i += G) s_isz[i] = g_isz[i];
      for (int base = 0; base < njall_max; base += G) {
        const int j = base + gl;
        if (j < nc) {
          int b1, b2;
          if (d.dropin) { b1 = con[j].side1; b2 = con[j].side2; }
          else { b1 = geoms[con[j].g1].body; b2 = geoms[con[j].g2].body; if (b1 < 0) { b1 = b2; b2 = -1; } }
          s_jb1[j] = (unsigned char)b1; s_jb2[j] = (unsigned char)(b2 < 0 ? 255 : b2);
        } else if (j < njall) {
          const ObJoint &pj = pjoint[j - nc];
          s_jb1[j] = (unsigned char)pj.b1; s_jb2[j] = (unsigned char)(pj.b2 < 0 ? 255 : pj.b2);
        }
      }
      __syncwarp();
    }
    // Info2.erp is one variable shared by all joints of an island and a ball joint overwrites it
    // (ball.cpp:60, quickstep.cpp:764-786): joint k sees the erp of the last ball joint before it
    if (anyball && gl == 0 && have_rows) {
      for (int isl = 0; isl < nis; isl++) {
        const int j0 = s_isz[4 * isl + 2], jn = s_isz[4 * isl + 3];
        int src = 0xffff;
        for (int k = j0; k < j0 + jn; k++) {
          s_erpsrc[k] = (unsigned short)src;
          const int j = s_ijoint[k];
          if (j >= nc && (pjoint[j - nc].type == OB_JOINT_BALL || pjoint[j - nc].type == OB_JOINT_FIXED)) src = k;
        }
      }
    }
    __syncwarp();
    // ---- (6) row assembly + finalisation.  Order of the sub-steps (r02r): (6a) getInfo2 of the PERMANENT joints, one lane per joint,
    // raw rows {J[12], c, cfm, lo, hi} staged in the row records, motor-at-limit side effects collected (joint.cpp:638-657); (6b) the side
    // effects applied in joint order; (4b) tmp1 per body; (6c) CONTACT joints, one lane per joint: getInfo2 and the finalisation of its
    // rows in one go (contacts have no side effects, so nothing they produce can change tmp1) -- their rows never pass through memory as
    // raw rows and the bodies' tmp1 / invI are loaded once per contact instead of once per row; (6d) finalisation of the staged rows of
    // the permanent joints.  Before r02r every row was staged raw and finalised by a lane per row: ncu (configs[3]) 66 % of the warp
    // samples of this half waited on those L2 round trips at 11 % occupancy.
    real *g_side = d.jside + (size_t)wc * (d.NJ ? d.NJ : 1) * 4 * OB_NSIDE;
    int anyside = 0;
    if (PJ) {
    for (int base = 0; base < nij_max; base += G) {
      const int k = base + gl;
      if (k < nij && have_rows && (int)s_ijoint[k] >= nc) {
        const int j = s_ijoint[k];
        const int b1 = s_jb1[j], b2 = s_jb2[j] == 255 ? -1 : (int)s_jb2[j];
        const int r0 = s_jrow[k];
        const unsigned ub2 = (unsigned)(b2 < 0 ? 255 : b2);
        real erp_in = W.erp;
        if (anyball) { const int src = s_erpsrc[k]; if (src != 0xffff) erp_in = pjoint[s_ijoint[src] - nc].erp; }
        ObJoint pj = pjoint[j - nc];
        ObBodyView B1 = {bd[b1].pos, bd[b1].R, bd[b1].q, bd[b1].lvel, bd[b1].avel}, B2 = B1;
        if (b2 >= 0) { B2.pos = bd[b2].pos; B2.R = bd[b2].R; B2.q = bd[b2].q; B2.lvel = bd[b2].lvel; B2.avel = bd[b2].avel; }
        const int jm = ob_joint_info1(pj, B1, b2 >= 0 ? &B2 : (const ObBodyView *)0);
        ObRowOut r;
        ob_rows_defaults(r, jm, W.cfm);
        real side[OB_NSIDE][4];
        real erp_io = erp_in;
        ob_joint_info2(r, pj, B1, b2 >= 0 ? &B2 : (const ObBodyView *)0, stepsize1, &erp_io, side);
        for (int sx = 0; sx < OB_NSIDE; sx++) {
          for (int e = 0; e < 4; e++) g_side[(size_t)(j - nc) * 4 * OB_NSIDE + 4 * sx + e] = side[sx][e];
          if (side[sx][0] != 0) anyside = 1;
        }
        for (int q = 0; q < jm; q++) {
          real rw[OB_ROWW];
          for (int e = 0; e < 12; e++) rw[e] = r.J[q][e];
          rw[12] = r.c[q]; rw[13] = r.cfm[q]; rw[14] = r.lo[q]; rw[15] = r.hi[q]; rw[16] = rw[17] = rw[18] = 0;
          const unsigned fio = (unsigned)(r.findex[q] >= 0 ? q - r.findex[q] : 0);
          store_row(rows + (size_t)(r0 + q) * OB_ROWW, rw, (unsigned)b1 | (ub2 << 8) | (fio << 16));
        }
      }
    }
    }
    for (int k = gl; k < nij; k += G) { g_ijoint[k] = s_ijoint[k]; g_jrow[k] = s_jrow[k]; }
    if (gl == 0 && valid) g_jrow[nij] = (unsigned short)(have_rows ? mtot : 0);
    for (int dd = 1; dd < G; dd <<= 1) anyside |= __shfl_xor_sync(FULL, anyside, dd, G);
    __syncwarp();
    // (6b) dBodyAddTorque / dBodyAddForce side effects in joint order (they change the accumulators before the rhs is formed)
    if (PJ && anyside && gl == 0) {
      for (int k = 0; k < nij; k++) {
        const int j = s_ijoint[k];
        if (j < nc) continue;
        const int b1 = s_jb1[j], b2 = s_jb2[j];
        real side[OB_NSIDE][4];
        bool any = false;
        for (int sx = 0; sx < OB_NSIDE; sx++) { for (int e = 0; e < 4; e++) side[sx][e] = __ldcg(g_side + (size_t)(j - nc) * 4 * OB_NSIDE + 4 * sx + e); any |= side[sx][0] != 0; }
        if (!any) continue;
        ob_apply_joint_side(pjoint[j - nc].type, side, bd[b1].facc, bd[b1].tacc, b2 != 255 ? bd[b2].facc : (real *)0, b2 != 255 ? bd[b2].tacc : (real *)0);
      }
    }
    __syncwarp();
    // (4b) tmp1 = v/h + invM*f_ext per island body (quickstep.cpp:840-846), after the side effects
    for (int base = 0; base < nib_max; base += G) {
      const int i = base + gl;
      if (i < nib && have_rows) {
        const int b = s_ibody[i];
        const ObBodyDyn &B = bd[b];
        real iw[12], avel[3], lvel[3], facc[3], tacc[3], t1[6];
        for (int k = 0; k < 12; k++) iw[k] = __ldcg(g_invIw + 12 * b + k);
        for (int k = 0; k < 3; k++) { avel[k] = B.avel[k]; lvel[k] = B.lvel[k]; facc[k] = __ldcg(&B.facc[k]); tacc[k] = __ldcg(&B.tacc[k]); }
        ob_body_tmp1(facc, tacc, lvel, avel, s_invM[b], iw, stepsize1, t1);
        for (int k = 0; k < 6; k++) g_tmp1[8 * b + k] = t1[k];
      }
    }
    __syncwarp();
    // one raw row {J[12], c, cfm, lo, hi} -> the compact record (rhs, cfm/h, iMJ, Ad: quickstep.cpp:849-857, :355-402)
#define OB_FINALIZE_ROW(RAW, RP, B1_, B2_, META_)                                                                               \
    {                                                                                                                            \
      real rw_[OB_ROWW];                                                                                                         \
      for (int e = 0; e < 6; e++) rw_[e] = (RAW)[e];                                                                             \
      for (int e = 0; e < 3; e++) rw_[6 + e] = (RAW)[9 + e];                                                                     \
      real iMJ_[12], b_out_, adcfm_, Ad_;                                                                                        \
      ob_row_finalize2((RAW), (RAW)[12], (RAW)[13], (B2_), t1a, t1b, s_invM[(B1_)], iw1, (B2_) >= 0 ? s_invM[(B2_)] : (real)0, iw2, stepsize1, \
                       W.sor_w, iMJ_, &b_out_, &adcfm_, &Ad_);                                                                   \
      for (int e = 0; e < 3; e++) { rw_[9 + e] = iMJ_[3 + e]; rw_[12 + e] = iMJ_[9 + e]; }                                       \
      rw_[15] = Ad_; rw_[16] = b_out_; rw_[17] = adcfm_;                                                                         \
      unsigned bmode_;                                                                                                           \
      if (!encode_bounds((RAW)[14], (RAW)[15], &rw_[18], &bmode_)) atomicOr(&W.status, OB_ERR_ROW_OVERFLOW);                     \
      store_row((RP), rw_, ((META_) & 0x00ffffffu) | (bmode_ << 24));                                                           \
    }
#define OB_LOAD_BODY_TERMS(B1_, B2_)                                                                                             \
      real t1a[6], t1b[6], iw1[12], iw2[12];                                                                                     \
      for (int e = 0; e < 6; e++) t1a[e] = __ldcg(g_tmp1 + 8 * (B1_) + e);                                                       \
      for (int e = 0; e < 12; e++) iw1[e] = __ldcg(g_invIw + 12 * (B1_) + e);                                                    \
      if ((B2_) >= 0) {                                                                                                          \
        for (int e = 0; e < 6; e++) t1b[e] = __ldcg(g_tmp1 + 8 * (B2_) + e);                                                     \
        for (int e = 0; e < 12; e++) iw2[e] = __ldcg(g_invIw + 12 * (B2_) + e);                                                  \
      }
    // ---- (6c) contact joints: getInfo2 + finalisation, one lane per joint
    for (int base = 0; base < nij_max; base += G) {
      const int k = base + gl;
      if (k < nij && have_rows && (!PJ || (int)s_ijoint[k] < nc)) {
        const int j = s_ijoint[k];
        const int b1 = s_jb1[j], b2 = s_jb2[j] == 255 ? -1 : (int)s_jb2[j];
        const int r0 = s_jrow[k];
        const unsigned ub2 = (unsigned)(b2 < 0 ? 255 : b2);
        real erp_in = W.erp;
        if (anyball) { const int src = s_erpsrc[k]; if (src != 0xffff) erp_in = pjoint[s_ijoint[src] - nc].erp; }
        ObSurface sf = csurf ? csurf[j] : ptab[con[j].policy].surface;
        const int jm = ob_contact_info1(sf);
        ObRowOut3 r;
        ob_rows_defaults(r, jm, W.cfm);
        const ObContact c = con[j];
        const int rev = d.dropin ? c.policy : (geoms[c.g1].body < 0);   // dJOINT_REVERSE
        real p1[3], l1[3], a1[3], p2[3] = {0, 0, 0}, l2[3] = {0, 0, 0}, a2[3] = {0, 0, 0};
        for (int e = 0; e < 3; e++) { p1[e] = bd[b1].pos[e]; l1[e] = bd[b1].lvel[e]; a1[e] = bd[b1].avel[e]; }
        if (b2 >= 0) for (int e = 0; e < 3; e++) { p2[e] = bd[b2].pos[e]; l2[e] = bd[b2].lvel[e]; a2[e] = bd[b2].avel[e]; }
        real fdir1[3] = {0, 0, 0};
        if (csurf) for (int e = 0; e < 3; e++) fdir1[e] = d.cfdir1[((size_t)wc * d.NC + j) * 4 + e];
        ob_contact_info2(r, jm, sf, c.pos, c.normal, c.depth, fdir1, rev, p1, l1, a1, b2 >= 0, p2, l2, a2, stepsize1,
                         erp_in, W.min_depth, W.max_vel);
        OB_LOAD_BODY_TERMS(b1, b2)
        for (int q = 0; q < jm; q++) {
          real raw[16];
          for (int e = 0; e < 12; e++) raw[e] = r.J[q][e];
          raw[12] = r.c[q]; raw[13] = r.cfm[q]; raw[14] = r.lo[q]; raw[15] = r.hi[q];
          const unsigned fio = (unsigned)(r.findex[q] >= 0 ? q - r.findex[q] : 0);
          OB_FINALIZE_ROW(raw, rows + (size_t)(r0 + q) * OB_ROWW, b1, b2, (unsigned)b1 | (ub2 << 8) | (fio << 16))
        }
      }
    }
    // ---- (6d) the staged rows of the permanent joints, one lane per joint
    if (PJ) {
    __syncwarp();
    for (int base = 0; base < nij_max; base += G) {
      const int k = base + gl;
      if (k < nij && have_rows && (int)s_ijoint[k] >= nc) {
        const int j = s_ijoint[k];
        const int b1 = s_jb1[j], b2 = s_jb2[j] == 255 ? -1 : (int)s_jb2[j];
        const int r0 = s_jrow[k], jm = (int)s_jrow[k + 1] - r0;
        OB_LOAD_BODY_TERMS(b1, b2)
        for (int q = 0; q < jm; q++) {
          real *rp = rows + (size_t)(r0 + q) * OB_ROWW;
          real raw[16];
          for (int e = 0; e < 16; e++) raw[e] = __ldcg(rp + e);
          const unsigned meta = *(const volatile unsigned *)(rp + OB_ROWF);
          OB_FINALIZE_ROW(raw, rp, b1, b2, meta)
        }
      }
    }
    }
#undef OB_FINALIZE_ROW
#undef OB_LOAD_BODY_TERMS
    __syncwarp();

    // (7) the row order and the level schedule of every shuffle epoch are built by k_sched
    if (gl == 0 && valid) {
      si[SI_NIS] = nis; si[SI_NIB] = nib; si[SI_NIJ] = nij; si[SI_MTOT] = mtot; si[SI_HAVEROWS] = have_rows ? 1 : 0;
      for (int i = 0; i < 4 * nis; i++) g_isz[i] = s_isz[i];
    }
    __syncwarp();
  }
}


// =====================================================================================
// k_sched: one warp per world.  For every shuffle epoch: the reference's row order
// (quickstep.cpp:409-482, LCG offsets per island as the reference consumes them) and the level
// schedule derived from it, written as sched[] (rows in level order) + pstart[] (first slot of
// every pass; a pass = at most G rows of one level).  The level recurrence is a serial chain
// over the order; it runs with the per-body "last level" table spread over the lanes'
// registers (body b -> lane b&31, register b>>5) so one step costs two shuffles, not a
// shared-memory round trip.
// the rows' body + findex bytes: from d.rowmeta when the split k_prep wrote them, else from the row records
__device__ __forceinline__ unsigned ob_row_meta(const ObBatchDev &d, int w, const real *rows, int i) {
  return d.rowmeta ? d.rowmeta[(size_t)w * d.NR + i] : *(const unsigned *)(rows + (size_t)i * OB_ROWW + OB_ROWF);
}
struct SchedSmem { size_t ord, rowb, fio, lvl, X, isl, last, total; };
__host__ __device__ inline SchedSmem sched_smem(int NB, int NR) {
  SchedSmem s; size_t o = 0;
  s.ord = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);
  s.rowb = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);
  s.fio = o; o = ob_al(o + (size_t)NR, 16);
  s.lvl = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);      // swap indices, then level per position
  s.X = o; o = ob_al(o + sizeof(int) * (NR + 2), 16);              // rows per level -> level starts -> level ends
  s.isl = o; o = ob_al(o + sizeof(unsigned short) * 2 * NB, 16);   // (r0, m) per island that has rows
  s.last = o; o = ob_al(o + sizeof(int) * 257, 16);                // level of the last row on every body (+ slot 255/256: "no body 2")
  s.total = ob_al(o, 16);
  return s;
}

template <int NBR>
__global__ void __launch_bounds__(32) k_sched(ObBatchDev d, int G, int taps) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SchedSmem L = sched_smem(d.NB, d.NR);
  unsigned short *s_ord = (unsigned short *)(smem + L.ord);
  unsigned short *s_rowb = (unsigned short *)(smem + L.rowb);
  unsigned char *s_fio = smem + L.fio;
  unsigned short *s_lvl = (unsigned short *)(smem + L.lvl);
  int *s_X = (int *)(smem + L.X);
  unsigned short *s_isl = (unsigned short *)(smem + L.isl);
  int *s_last = (int *)(smem + L.last);
  const int lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int w = d.wbeg + blockIdx.x; w < d.wend; w += gridDim.x) {
    ObWorld &W = d.world[w];
    int *si = d.stepinfo + (size_t)w * SI_WORDS;
    const int nis = si[SI_NIS];
    const bool have_rows = si[SI_HAVEROWS] != 0;
    const int mtot = have_rows ? si[SI_MTOT] : 0;
    const int nep = (W.iters + 7) >> 3;
    const real *rows = d.rows + (size_t)w * d.NR * OB_ROWW;
    const unsigned short *g_isz = d.isz + (size_t)w * 4 * d.NB;
    const unsigned short *g_jrow = d.jrow + (size_t)w * (d.NC + d.NJ + 1);
    if (mtot == 0) {
      if (lane == 0) for (int ep = 0; ep < d.NEP; ep++) si[SI_NPASS0 + ep] = 0;
      continue;
    }
    // islands that own rows, in island order
    int nri = 0;
    if (lane == 0) {
      for (int isl = 0; isl < nis; isl++) {
        const int j0 = g_isz[4 * isl + 2], jn = g_isz[4 * isl + 3];
        if (!jn) continue;
        const int r0 = g_jrow[j0], m = g_jrow[j0 + jn] - r0;
        if (m > 0) { s_isl[2 * nri] = (unsigned short)r0; s_isl[2 * nri + 1] = (unsigned short)m; nri++; }
      }
    }
    nri = __shfl_sync(FULL, nri, 0);
    for (int i = lane; i < mtot; i += 32) {
      const unsigned meta = ob_row_meta(d, w, rows, i);
      s_rowb[i] = (unsigned short)(meta & 0xffffu);
      s_fio[i] = (unsigned char)((meta >> 16) & 255u);
    }
    __syncwarp();
    // initial order per island (quickstep.cpp:409-424): findex==-1 rows ascending at the head,
    // the others descending at the tail
    for (int q = 0; q < nri; q++) {
      const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
      int head = 0, tail = 0;
      for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        const bool v = i < m;
        const bool hf = v && s_fio[r0 + i] == 0;
        const unsigned bh = __ballot_sync(FULL, hf), bt = __ballot_sync(FULL, v && !hf);
        if (hf) s_ord[r0 + head + __popc(bh & lt_mask)] = (unsigned short)i;
        else if (v) s_ord[r0 + m - 1 - (tail + __popc(bt & lt_mask))] = (unsigned short)i;
        head += __popc(bh); tail += __popc(bt);
      }
    }
    __syncwarp();
    const uint32_t seed = W.seed;
    unsigned total_draws = 0;
    for (int ep = 0; ep < nep && ep < d.NEP; ep++) {
      // (a) shuffle every island's segment (quickstep.cpp:474-481).  The reference runs ALL
      // iterations of island 0 before island 1, so the draws of (island i, epoch e) start at
      // offset  sum_{j<i} nep*(m_j-1) + e*(m_i-1)  of the world's LCG stream.
      unsigned draws_before = 0;
      for (int q = 0; q < nri; q++) {
        const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
        if (m < 2) continue;
        const unsigned my_off = draws_before + (unsigned)ep * (unsigned)(m - 1);
        draws_before += (unsigned)nep * (unsigned)(m - 1);
        uint32_t A, C, A0, C0;
        ob_lcg_skip(my_off + (unsigned)lane + 1u, &A0, &C0);   // lane handles i = lane+1, lane+33, ...
        uint32_t s = A0 * seed + C0;
        ob_lcg_skip(32u, &A, &C);
        for (int i = 1 + lane; i < m; i += 32) {
          s_lvl[r0 + i] = (unsigned short)ob_randint_fold(s, (uint32_t)(i + 1));
          s = A * s + C;
        }
        __syncwarp();
        if (lane == 0) {
          unsigned short *ord = s_ord + r0;
          int sj = s_lvl[r0 + 1];
          for (int i = 1; i < m; i++) {
            const int sjn = s_lvl[r0 + (i + 1 < m ? i + 1 : i)];   // the next swap index is fetched ahead of the dependent chain
            const unsigned short t = ord[i], u = ord[sj];
            ord[i] = u; ord[sj] = t;
            sj = sjn;
          }
        }
        __syncwarp();
      }
      total_draws = draws_before;
      // (b) level of every position, islands and positions in sweep order.  The recurrence
      // level(k) = 1 + max(last[b1], last[b2]) is one dependent chain per world: every lane runs it in lock-step on
      // the shared `last` table (uniform addresses: broadcast loads, same-value stores), ~10 instructions per row;
      // the rows' body bytes are gathered into s_lvl in parallel first and fetched one row ahead of the chain
      for (int i = lane; i <= mtot + 1; i += 32) s_X[i] = 0;
      for (int i = lane; i < 257; i += 32) s_last[i] = 0;
      for (int q = 0; q < nri; q++) {
        const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
        for (int k = lane; k < m; k += 32) s_lvl[r0 + k] = s_rowb[r0 + s_ord[r0 + k]];
      }
      __syncwarp();
      int nlev = 0;
      for (int q = 0; q < nri; q++) {
        const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
        unsigned rb = s_lvl[r0];
        for (int k = 0; k < m; k++) {
          const unsigned rbn = s_lvl[r0 + (k + 1 < m ? k + 1 : k)];
          const int b1 = rb & 255, b2 = (rb >> 8) & 255;      // b2 == 255: slot 255 is read (always 0) and slot 256 written
          const int l1 = s_last[b1], l2 = s_last[b2];
          const int lv = (l2 > l1 ? l2 : l1) + 1;
          s_last[b1] = lv;
          s_last[b2 + (b2 == 255)] = lv;
          s_lvl[r0 + k] = (unsigned short)lv;
          nlev = lv > nlev ? lv : nlev;
          rb = rbn;
        }
      }
      __syncwarp();
      for (int q = 0; q < nri; q++) {
        const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
        for (int k = lane; k < m; k += 32) atomicAdd(&s_X[s_lvl[r0 + k]], 1);
      }
      __syncwarp();
      // exclusive prefix over levels 1..nlev: s_X[l] = first slot of level l
      {
        int carry = 0;
        for (int base = 1; base <= nlev; base += 32) {
          const int l = base + lane;
          const int c = l <= nlev ? s_X[l] : 0;
          int x = c;
          for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd); if (lane >= dd) x += y; }
          if (l <= nlev) s_X[l] = carry + x - c;
          carry += __shfl_sync(FULL, x, 31);
        }
      }
      __syncwarp();
      // (c) scatter rows into level order (any order inside a level); afterwards s_X[l] = end of level l
      unsigned short *g_sched = d.sched + ((size_t)w * d.NEP + ep) * d.NR;
      unsigned short *g_pstart = d.pstart + ((size_t)w * d.NEP + ep) * (d.NR + 1);
      for (int q = 0; q < nri; q++) {
        const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
        for (int k = lane; k < m; k += 32) {
          const int pos = atomicAdd(&s_X[s_lvl[r0 + k]], 1);
          g_sched[pos] = (unsigned short)(r0 + s_ord[r0 + k]);
        }
      }
      __syncwarp();
      // pass table: a pass is a chunk of <= G consecutive slots that does not cross a level boundary
      {
        int carry = 0;
        for (int base = 1; base <= nlev; base += 32) {
          const int l = base + lane;
          int start = 0, n = 0;
          if (l <= nlev) { start = l > 1 ? s_X[l - 1] : 0; n = s_X[l] - start; }
          const int c = (n + G - 1) / G;
          int x = c;
          for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd); if (lane >= dd) x += y; }
          const int off = carry + x - c;
          for (int j = 0; j < c; j++) g_pstart[off + j] = (unsigned short)(start + j * G);
          carry += __shfl_sync(FULL, x, 31);
        }
        if (lane == 0) { g_pstart[carry] = (unsigned short)mtot; si[SI_NPASS0 + ep] = carry; }
      }
      __syncwarp();
    }
    if (lane == 0) {
      uint32_t A, C;
      ob_lcg_skip(total_draws, &A, &C);
      W.seed = A * seed + C;
    }
    __syncwarp();
  }
}


// k_sched_lane: the same products as k_sched, but ONE LANE PER WORLD.  Everything k_sched does is a serial
// chain per world (the m-1 dependent swaps of the shuffle, the level recurrence over the order), which a
// warp per world executes with 31 idle lanes; here 32 worlds advance in lock-step, each lane in its own
// column of shared memory (element i of lane l at [i*32 + l]: bank == lane, conflict-free for any index).
// Used when the per-warp working set fits shared memory (ob_backend_cuda.cu), else k_sched.
struct SchedLaneSmem { size_t ord, rowb, lvl, X, fio, last, isl, total; };
__host__ __device__ inline SchedLaneSmem sched_lane_smem(int NB, int NR) {
  SchedLaneSmem s; size_t o = 0;
  s.ord = o; o = ob_al(o + sizeof(unsigned short) * 32 * NR, 16);
  s.rowb = o; o = ob_al(o + sizeof(unsigned short) * 32 * NR, 16);
  s.lvl = o; o = ob_al(o + sizeof(unsigned short) * 32 * NR, 16);
  s.X = o; o = ob_al(o + sizeof(unsigned short) * 32 * (NR + 2), 16);
  s.fio = o; o = ob_al(o + (size_t)32 * NR, 16);
  s.last = o; o = ob_al(o + sizeof(unsigned short) * 32 * (NB + 1), 16);
  s.isl = o; o = ob_al(o + sizeof(unsigned short) * 32 * 2 * NB, 16);
  s.total = ob_al(o, 16);
  return s;
}
__global__ void __launch_bounds__(32) k_sched_lane(ObBatchDev d, int G) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SchedLaneSmem L = sched_lane_smem(d.NB, d.NR);
  const int lane = threadIdx.x;
  unsigned short *s_ord = (unsigned short *)(smem + L.ord) + lane;
  unsigned short *s_rowb = (unsigned short *)(smem + L.rowb) + lane;
  unsigned short *s_lvl = (unsigned short *)(smem + L.lvl) + lane;
  unsigned short *s_X = (unsigned short *)(smem + L.X) + lane;
  unsigned char *s_fio = smem + L.fio + lane;
  unsigned short *s_last = (unsigned short *)(smem + L.last) + lane;
  unsigned short *s_isl = (unsigned short *)(smem + L.isl) + lane;
#define LN(a, i) a[(size_t)(i) * 32]
  for (int wbase = d.wbeg + blockIdx.x * 32; wbase < d.wend; wbase += gridDim.x * 32) {
    const int w = wbase + lane;
    if (w >= d.wend) continue;
    ObWorld &W = d.world[w];
    int *si = d.stepinfo + (size_t)w * SI_WORDS;
    const int nis = si[SI_NIS];
    const int mtot = si[SI_HAVEROWS] ? si[SI_MTOT] : 0;
    const int nep = (W.iters + 7) >> 3;
    if (mtot == 0) { for (int ep = 0; ep < d.NEP; ep++) si[SI_NPASS0 + ep] = 0; continue; }
    const real *rows = d.rows + (size_t)w * d.NR * OB_ROWW;
    const unsigned short *g_isz = d.isz + (size_t)w * 4 * d.NB;
    const unsigned short *g_jrow = d.jrow + (size_t)w * (d.NC + d.NJ + 1);
    const int nb = W.nb;
    int nri = 0;
    for (int isl = 0; isl < nis; isl++) {
      const int j0 = g_isz[4 * isl + 2], jn = g_isz[4 * isl + 3];
      if (!jn) continue;
      const int r0 = g_jrow[j0], m = g_jrow[j0 + jn] - r0;
      if (m > 0) { LN(s_isl, 2 * nri) = (unsigned short)r0; LN(s_isl, 2 * nri + 1) = (unsigned short)m; nri++; }
    }
    for (int i = 0; i < mtot; i++) {
      const unsigned meta = ob_row_meta(d, w, rows, i);
      LN(s_rowb, i) = (unsigned short)(meta & 0xffffu);
      LN(s_fio, i) = (unsigned char)((meta >> 16) & 255u);
    }
    // initial order per island (quickstep.cpp:409-424)
    for (int q = 0; q < nri; q++) {
      const int r0 = LN(s_isl, 2 * q), m = LN(s_isl, 2 * q + 1);
      int head = 0, tail = 0;
      for (int i = 0; i < m; i++) {
        if (LN(s_fio, r0 + i) == 0) LN(s_ord, r0 + head++) = (unsigned short)i;
        else LN(s_ord, r0 + m - 1 - tail++) = (unsigned short)i;
      }
    }
    const uint32_t seed = W.seed;
    unsigned total_draws = 0;
    for (int ep = 0; ep < nep && ep < d.NEP; ep++) {
      // (a) shuffle (quickstep.cpp:474-481); draw offsets as in k_sched
      unsigned draws_before = 0;
      for (int q = 0; q < nri; q++) {
        const int r0 = LN(s_isl, 2 * q), m = LN(s_isl, 2 * q + 1);
        if (m < 2) continue;
        const unsigned my_off = draws_before + (unsigned)ep * (unsigned)(m - 1);
        draws_before += (unsigned)nep * (unsigned)(m - 1);
        uint32_t A, C;
        ob_lcg_skip(my_off, &A, &C);
        uint32_t s = A * seed + C;
        for (int i = 1; i < m; i++) {
          s = 1664525u * s + 1013904223u;
          const int sj = ob_randint_fold(s, (uint32_t)(i + 1));
          const unsigned short t = LN(s_ord, r0 + i); LN(s_ord, r0 + i) = LN(s_ord, r0 + sj); LN(s_ord, r0 + sj) = t;
        }
      }
      total_draws = draws_before;
      // (b) levels
      for (int b = 0; b < nb; b++) LN(s_last, b) = 0;
      for (int i = 0; i <= mtot + 1; i++) LN(s_X, i) = 0;
      int nlev = 0;
      for (int q = 0; q < nri; q++) {
        const int r0 = LN(s_isl, 2 * q), m = LN(s_isl, 2 * q + 1);
        for (int k = 0; k < m; k++) {
          const unsigned rb = LN(s_rowb, r0 + LN(s_ord, r0 + k));
          const int b1 = rb & 255, b2 = (rb >> 8) & 255;
          int lv = LN(s_last, b1);
          if (b2 != 255) { const int l2 = LN(s_last, b2); lv = l2 > lv ? l2 : lv; }
          lv++;
          LN(s_last, b1) = (unsigned short)lv;
          if (b2 != 255) LN(s_last, b2) = (unsigned short)lv;
          LN(s_lvl, r0 + k) = (unsigned short)lv;
          LN(s_X, lv) = LN(s_X, lv) + 1;
          nlev = lv > nlev ? lv : nlev;
        }
      }
      {
        int run = 0;
        for (int l = 1; l <= nlev; l++) { const int c = LN(s_X, l); LN(s_X, l) = (unsigned short)run; run += c; }
      }
      // (c) rows into level order; afterwards X[l] = end of level l
      unsigned short *g_sched = d.sched + ((size_t)w * d.NEP + ep) * d.NR;
      unsigned short *g_pstart = d.pstart + ((size_t)w * d.NEP + ep) * (d.NR + 1);
      for (int q = 0; q < nri; q++) {
        const int r0 = LN(s_isl, 2 * q), m = LN(s_isl, 2 * q + 1);
        for (int k = 0; k < m; k++) {
          const int lv = LN(s_lvl, r0 + k);
          const int pos = LN(s_X, lv);
          LN(s_X, lv) = (unsigned short)(pos + 1);
          g_sched[pos] = (unsigned short)(r0 + LN(s_ord, r0 + k));
        }
      }
      int total = 0, start = 0;
      for (int l = 1; l <= nlev; l++) {
        const int end = LN(s_X, l);
        for (int p0 = start; p0 < end; p0 += G) g_pstart[total++] = (unsigned short)p0;
        start = end;
      }
      g_pstart[total] = (unsigned short)mtot;
      si[SI_NPASS0 + ep] = total;
    }
    uint32_t A, C;
    ob_lcg_skip(total_draws, &A, &C);
    W.seed = A * seed + C;
  }
#undef LN
}


// k_sched_tile<GS>: the same products as k_sched, GS lanes per world and 32/GS worlds per warp.  What k_sched spends
// its instructions on are the two serial chains per epoch (the m-1 dependent swaps of the shuffle and the level
// recurrence): a warp per world issues them as warp instructions with one live lane (ncu r01z: 237 M warp
// instructions, 70 % issue-active, as many as the whole sweep).  Here lane 0 of every tile runs its world's chain, so
// one warp instruction advances 32/GS worlds; the data-parallel parts (draws, gather, scatter, pass table) use the
// tile's GS lanes.  k_sched_lane is the GS = 1 end of the same idea (too few warps for batches of big worlds).
struct SchedTileSmem { size_t ord, rowb, fio, lvl, X, isl, last, total; };
__host__ __device__ inline SchedTileSmem sched_tile_smem(int NB, int NR) {
  SchedTileSmem s; size_t o = 0;
  s.ord = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);
  s.rowb = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);
  s.fio = o; o = ob_al(o + (size_t)NR, 16);
  s.lvl = o; o = ob_al(o + sizeof(unsigned short) * NR, 16);      // swap indices, then body bytes, then level per position
  s.X = o; o = ob_al(o + sizeof(int) * (NR + 2), 16);              // rows per level -> level starts -> level ends
  s.isl = o; o = ob_al(o + sizeof(unsigned short) * 2 * NB, 16);   // (r0, m) per island that has rows
  s.last = o; o = ob_al(o + sizeof(unsigned short) * 258, 16);     // level of the last row on every body (+ slots 255/256: "no body 2")
  s.total = ob_al(o, 16);
  return s;
}
// Tried and rejected (r02p, B200): running the shuffle chain of epoch e + 1 and the level chain of epoch e interleaved in one serial
// loop (two independent dependency chains, nep + 1 chain lengths instead of 2 nep, order[] double-buffered): 0.31 -> 0.54 ms on
// configs[1], 0.82 -> 1.20 ms on configs[3] -- the longer loop body and 4 NR more bytes of shared memory per world cost more than the
// overlap gains.
template <int GS>
__global__ void __launch_bounds__(32) k_sched_tile(ObBatchDev d, int G) {
  constexpr int T = 32 / GS;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const SchedTileSmem L = sched_tile_smem(d.NB, d.NR);
  const int lane = threadIdx.x, grp = lane / GS, gl = lane % GS;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  unsigned short *s_ord = (unsigned short *)(smem + L.ord);
  unsigned short *s_rowb = (unsigned short *)(smem + L.rowb);
  unsigned char *s_fio = smem + L.fio;
  unsigned short *s_lvl = (unsigned short *)(smem + L.lvl);
  int *s_X = (int *)(smem + L.X);
  unsigned short *s_isl = (unsigned short *)(smem + L.isl);
  unsigned short *s_last = (unsigned short *)(smem + L.last);
  const unsigned FULL = 0xffffffffu;
  const unsigned gmask = GS == 32 ? FULL : ((1u << GS) - 1u);
  const unsigned lt_mask = (1u << gl) - 1u;
  const int gsh = grp * GS;

  for (int wbase = d.wbeg + blockIdx.x * T; wbase < d.wend; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.wend;
    const int wc = valid ? w : d.wbeg;
    ObWorld &W = d.world[wc];
    int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nis = valid ? si[SI_NIS] : 0;
    const int mtot = (valid && si[SI_HAVEROWS]) ? si[SI_MTOT] : 0;
    int nep = valid ? (W.iters + 7) >> 3 : 0;
    if (nep > d.NEP) nep = d.NEP;
    const real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    const unsigned short *g_isz = d.isz + (size_t)wc * 4 * d.NB;
    const unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + d.NJ + 1);
    if (valid && mtot == 0 && gl == 0) for (int ep = 0; ep < d.NEP; ep++) si[SI_NPASS0 + ep] = 0;
    // islands that own rows, in island order
    int nri = 0;
    if (gl == 0 && mtot > 0) {
      for (int isl = 0; isl < nis; isl++) {
        const int j0 = g_isz[4 * isl + 2], jn = g_isz[4 * isl + 3];
        if (!jn) continue;
        const int r0 = g_jrow[j0], m = g_jrow[j0 + jn] - r0;
        if (m > 0) { s_isl[2 * nri] = (unsigned short)r0; s_isl[2 * nri + 1] = (unsigned short)m; nri++; }
      }
    }
    nri = __shfl_sync(FULL, nri, 0, GS);
    for (int i = gl; i < mtot; i += GS) {
      const unsigned meta = ob_row_meta(d, wc, rows, i);
      s_rowb[i] = (unsigned short)(meta & 0xffffu);
      s_fio[i] = (unsigned char)((meta >> 16) & 255u);
    }
    __syncwarp();
    // initial order per island (quickstep.cpp:409-424): findex==-1 rows ascending at the head, the others descending at the tail
    const int nri_max = warp_max_i(nri);
    for (int q = 0; q < nri_max; q++) {
      const int r0 = q < nri ? s_isl[2 * q] : 0, m = q < nri ? s_isl[2 * q + 1] : 0;
      const int m_max = warp_max_i(m);
      int head = 0, tail = 0;
      for (int base = 0; base < m_max; base += GS) {
        const int i = base + gl;
        const bool v = i < m;
        const bool hf = v && s_fio[r0 + i] == 0;
        const unsigned bh = (__ballot_sync(FULL, hf) >> gsh) & gmask, bt = (__ballot_sync(FULL, v && !hf) >> gsh) & gmask;
        if (hf) s_ord[r0 + head + __popc(bh & lt_mask)] = (unsigned short)i;
        else if (v) s_ord[r0 + m - 1 - (tail + __popc(bt & lt_mask))] = (unsigned short)i;
        head += __popc(bh); tail += __popc(bt);
      }
    }
    __syncwarp();
    const uint32_t seed = W.seed;
    unsigned total_draws = 0;
    for (int q = 0; q < nri; q++) { const int m = s_isl[2 * q + 1]; if (m >= 2) total_draws += (unsigned)nep * (unsigned)(m - 1); }
    const int nep_max = warp_max_i(mtot > 0 ? nep : 0);
    for (int ep = 0; ep < nep_max; ep++) {
      const bool epv = mtot > 0 && ep < nep;
      const int nri_e = epv ? nri : 0, mtot_e = epv ? mtot : 0;
      // (a) shuffle every island's segment (quickstep.cpp:474-481).  The reference runs ALL iterations of island 0
      // before island 1, so the draws of (island i, epoch e) start at offset sum_{j<i} nep*(m_j-1) + e*(m_i-1)
      {
        unsigned draws_before = 0;
        for (int q = 0; q < nri_e; q++) {
          const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
          if (m < 2) continue;
          const unsigned my_off = draws_before + (unsigned)ep * (unsigned)(m - 1);
          draws_before += (unsigned)nep * (unsigned)(m - 1);
          uint32_t A, C, A0, C0;
          ob_lcg_skip(my_off + (unsigned)gl + 1u, &A0, &C0);   // lane handles i = gl+1, gl+1+GS, ...
          uint32_t s = A0 * seed + C0;
          ob_lcg_skip((unsigned)GS, &A, &C);
          for (int i = 1 + gl; i < m; i += GS) {
            s_lvl[r0 + i] = (unsigned short)(r0 + ob_randint_fold(s, (uint32_t)(i + 1)));   // absolute position of the swap partner
            s = A * s + C;
          }
        }
        // the first position of every island swaps with itself: the chain below needs no island boundaries
        for (int q = gl; q < nri_e; q += GS) s_lvl[s_isl[2 * q]] = s_isl[2 * q];
      }
      __syncwarp();
      // the m-1 dependent swaps of every island, as ONE flat loop over the world's positions with a warp-uniform trip count:
      // the leader lanes of the warp's tiles stay converged, one warp instruction advances every tile's chain
      // (r02b: with per-island loops of different lengths the leaders drifted apart, 1.4 live lanes per instruction)
      {
        const int kmax = warp_max_i(mtot_e);
        if (gl == 0) {
          int tg = mtot_e > 0 ? (int)s_lvl[0] : 0;
          for (int k = 0; k < kmax; k++) {
            const bool on = k < mtot_e;
            const int kk = on ? k : 0, tj = on ? tg : 0;
            const int tgn = s_lvl[(on && k + 1 < mtot_e) ? k + 1 : kk];   // the next partner is fetched ahead of the dependent chain
            const unsigned short t = s_ord[kk], u = s_ord[tj];
            if (on) { s_ord[kk] = u; s_ord[tj] = t; }
            tg = tgn;
          }
        }
      }
      __syncwarp();
      // (b) level of every position, islands and positions in sweep order: level(k) = 1 + max(last[b1], last[b2])
      for (int i = gl; i <= mtot_e + 1; i += GS) s_X[i] = 0;
      if (epv) for (int i = gl; i < 258; i += GS) s_last[i] = 0;
      for (int q = 0; q < nri_e; q++) {
        const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
        for (int k = gl; k < m; k += GS) s_lvl[r0 + k] = s_rowb[r0 + s_ord[r0 + k]];
      }
      __syncwarp();
      int nlev = 0;
      {
        // islands own disjoint bodies and rows [r0, r0 + m) are contiguous in island order, so the recurrence is one flat
        // loop over the world's positions; warp-uniform trip count as above
        const int kmax = warp_max_i(mtot_e);
        if (gl == 0) {
          unsigned rb = mtot_e > 0 ? (unsigned)s_lvl[0] : 0u;
          for (int k = 0; k < kmax; k++) {
            const bool on = k < mtot_e;
            const int kk = on ? k : 0;
            const unsigned rbn = s_lvl[(on && k + 1 < mtot_e) ? k + 1 : kk];
            const int b1 = rb & 255, b2 = (rb >> 8) & 255;      // b2 == 255: slot 255 is read (always 0) and slot 256 written
            const int l1 = s_last[b1], l2 = s_last[b2];
            const int lv = (l2 > l1 ? l2 : l1) + 1;
            if (on) {
              s_last[b1] = (unsigned short)lv;
              s_last[b2 + (b2 == 255)] = (unsigned short)lv;
              s_lvl[kk] = (unsigned short)lv;
              nlev = lv > nlev ? lv : nlev;
            }
            rb = rbn;
          }
        }
      }
      nlev = __shfl_sync(FULL, nlev, 0, GS);
      __syncwarp();
      // rows per level, counted by the tile's lanes afterwards: inside the serial loop the read-modify-write of s_X put a second
      // shared-memory round trip into every step of an in-order instruction stream (ncu r02j: as many stall samples as the chain itself)
      for (int k = gl; k < mtot_e; k += GS) atomicAdd(&s_X[s_lvl[k]], 1);
      __syncwarp();
      const int nlev_max = warp_max_i(nlev);
      // exclusive prefix over levels 1..nlev: s_X[l] = first slot of level l
      {
        int carry = 0;
        for (int base = 1; base <= nlev_max; base += GS) {
          const int l = base + gl;
          const int c = l <= nlev ? s_X[l] : 0;
          int x = c;
          for (int dd = 1; dd < GS; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd, GS); if (gl >= dd) x += y; }
          if (l <= nlev) s_X[l] = carry + x - c;
          carry += __shfl_sync(FULL, x, GS - 1, GS);
        }
      }
      __syncwarp();
      // (c) scatter rows into level order (any order inside a level); afterwards s_X[l] = end of level l
      unsigned short *g_sched = d.sched + ((size_t)wc * d.NEP + ep) * d.NR;
      unsigned short *g_pstart = d.pstart + ((size_t)wc * d.NEP + ep) * (d.NR + 1);
      for (int q = 0; q < nri_e; q++) {
        const int r0 = s_isl[2 * q], m = s_isl[2 * q + 1];
        for (int k = gl; k < m; k += GS) {
          const int pos = atomicAdd(&s_X[s_lvl[r0 + k]], 1);
          g_sched[pos] = (unsigned short)(r0 + s_ord[r0 + k]);
        }
      }
      __syncwarp();
      // pass table: a pass is a chunk of <= G consecutive slots that does not cross a level boundary
      {
        int carry = 0;
        for (int base = 1; base <= nlev_max; base += GS) {
          const int l = base + gl;
          int start = 0, n = 0;
          if (l <= nlev) { start = l > 1 ? s_X[l - 1] : 0; n = s_X[l] - start; }
          const int c = (n + G - 1) / G;
          int x = c;
          for (int dd = 1; dd < GS; dd <<= 1) { const int y = __shfl_up_sync(FULL, x, dd, GS); if (gl >= dd) x += y; }
          const int off = carry + x - c;
          for (int j = 0; j < c; j++) g_pstart[off + j] = (unsigned short)(start + j * G);
          carry += __shfl_sync(FULL, x, GS - 1, GS);
        }
        if (gl == 0 && epv) { g_pstart[carry] = (unsigned short)mtot; si[SI_NPASS0 + ep] = carry; }
      }
      __syncwarp();
    }
    if (gl == 0 && mtot > 0) {
      uint32_t A, C;
      ob_lcg_skip(total_draws, &A, &C);
      W.seed = A * seed + C;
    }
    __syncwarp();
  }
}


// one row update of the sweep (quickstep.cpp:490-581) on a row held in registers
__device__ __forceinline__ void sor_pass(const ObRowReg &cur, int cur_idx, real *s_fc, real *s_lam, const real *s_invM) {
  const int b1 = cur.meta & 255, b2r = (cur.meta >> 8) & 255, fio = (cur.meta >> 16) & 255, bmode = cur.meta >> 24;
  const int b2 = b2r == 255 ? -1 : b2r;
  const int fi = fio ? cur_idx - fio : -1;
  const real Ad = cur.v[15], k1 = s_invM[b1];
  const real bv = cur.v[18];
  const real lo = bmode == 0 ? -bv : (bmode == 1 ? (real)0 : bv);
  const real hi = bmode == 2 ? (real)0 : bv;
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
  const real lam_new = ob_sor_row(J, iMJ, cur.v[16], cur.v[17], lo, hi, fi, fi >= 0 ? s_lam[fi] : (real)0, s_lam[cur_idx], f1,
                                  b2 >= 0 ? f2 : (real *)0);
  s_lam[cur_idx] = lam_new;
#if defined(dSINGLE)
  *(float4 *)fp1 = make_float4(f1[0], f1[1], f1[2], f1[3]); *(float2 *)(fp1 + 4) = make_float2(f1[4], f1[5]);
  if (b2 >= 0) { *(float4 *)fp2 = make_float4(f2[0], f2[1], f2[2], f2[3]); *(float2 *)(fp2 + 4) = make_float2(f2[4], f2[5]); }
#else
  for (int e = 0; e < 6; e++) fp1[e] = f1[e];
  if (b2 >= 0) for (int e = 0; e < 6; e++) fp2[e] = f2[e];
#endif
}
// debug (taps & 4): verify that the rows of one pass touch pairwise disjoint bodies
template <int G>
__device__ __noinline__ void sor_check_pass(bool act, unsigned meta, int gl, int *status) {
  const int mb1 = act ? (int)(meta & 255) : -1, mb2 = act ? (int)((meta >> 8) & 255) : -1;
  for (int l2 = 0; l2 < G; l2++) {
    const int o1 = __shfl_sync(0xffffffffu, mb1, l2, G), o2 = __shfl_sync(0xffffffffu, mb2, l2, G);
    if (act && l2 != gl && o1 >= 0) {
      if (mb1 == o1 || mb1 == o2 || (mb2 != 255 && (mb2 == o1 || mb2 == o2))) atomicOr(status, 256);
    }
  }
}

// fc slot swizzle of the ring / pair sweep kernels (see SorRingSmem)
__host__ __device__ __forceinline__ int ob_fc4(int b) { return 8 * b + ((b & 4) ? 4 : 0); }   // word offset of fc[0..3]
__host__ __device__ __forceinline__ int ob_fc2(int b) { return 8 * b + ((b & 4) ? 0 : 4); }   // word offset of fc[4..5]
// after the sweeps: cforce per body for k_post; lambda + joint feedback taps (quickstep.cpp:918-957)
template <int G>
__device__ __forceinline__ void sor_epilogue(const ObBatchDev &d, int taps, int wc, bool valid, int gl, int nb, int mtot, const int *si,
                                             const real *rows, const real *s_fc, const real *s_lam, int fcs = 8) {
  real *g_fc = d.tmp1 + (size_t)wc * d.NB * 8;   // tmp1 is dead after row assembly: reuse as cforce
  for (int b = gl; b < nb; b += G)
    for (int k = 0; k < 6; k++) g_fc[8 * b + k] = fcs == 1 ? s_fc[(k < 4 ? ob_fc4(b) : ob_fc2(b) - 4) + k] : s_fc[8 * b + k];   // fcs == 1: swizzled slots (ring / pair kernels)
  if (taps && valid) {
    const int nij = si[SI_NIJ];
    const unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + d.NJ + 1);
    const unsigned short *g_ijoint = d.ijoint + (size_t)wc * (d.NC + d.NJ);
    const int ncw = d.ncontacts[wc];
    real *fb = d.fback + (size_t)wc * (d.NC + d.NJ) * 12;
    real *gl_lam = d.lambda + (size_t)wc * d.NR;
    if (mtot > 0)
      for (int k = gl; k < nij; k += G) {
        const int jr0 = g_jrow[k], jm = g_jrow[k + 1] - jr0;
        // Multiply1_12q1 (quickstep.cpp:70-101): data = J^T lambda, body 1 then body 2 (J2l == -J1l)
        real acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool two = false;
        for (int q = 0; q < jm; q++) {
          const real *rp = rows + (size_t)(jr0 + q) * OB_ROWW;
          const real s = s_lam[jr0 + q];
          const unsigned meta = *(const unsigned *)(rp + OB_ROWF);
          two = ((meta >> 8) & 255u) != 255u;
          for (int e = 0; e < 6; e++) acc[e] += rp[e] * s;
          for (int e = 0; e < 3; e++) { acc[6 + e] += (-rp[e]) * s; acc[9 + e] += rp[6 + e] * s; }
        }
        if (!two) for (int e = 6; e < 12; e++) acc[e] = 0;
        const int jid = g_ijoint[k];
        real *o = fb + (size_t)(jid < ncw ? jid : d.NC + (jid - ncw)) * 12;
        for (int e = 0; e < 12; e++) o[e] = acc[e];
      }
    for (int i = gl; i < mtot; i += G) gl_lam[i] = s_lam[i];
  }
}

// =====================================================================================
// DEEP: index prefetch five passes ahead (worlds with many rows: the sweep is long and the pipeline's
// prologue is amortised); otherwise the shallow pipeline with its short prologue (many tiny worlds).
// Measured on B200: config 2 (377 rows/world) 1.07 -> 0.93 ms with DEEP (index 6 passes, rows 3 passes ahead, four
// row buffers), config 3 (56 rows/world) 5.19 -> 5.59 ms.
// (tried on B200: capping k_sor<4, false> at 128 registers for 16 warps per SM instead of 10 -- 176 B of spills in the
// sweep loop, config 3 k_sor 3.07 -> 4.04 ms; kept at 186 registers)
template <int G, bool DEEP>
__global__ void __launch_bounds__(32) k_sor(ObBatchDev d, int taps) {
  constexpr int T = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const SorTileSmem L = sor_tile_smem(d.NB, d.NR);
  const int lane = threadIdx.x, grp = lane / G, gl = lane % G;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  real *s_fc = (real *)(smem + L.fc);
  real *s_lam = (real *)(smem + L.lam);
  real *s_invM = (real *)(smem + L.invM);

  for (int wbase = d.wbeg + blockIdx.x * T; wbase < d.wend; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.wend;
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
      if constexpr (DEEP) {
      // software pipeline, no load waits on another load issued less than three passes earlier:
      // pstart[p+10] -> row index of pass p+6 -> row record of pass p+3 (four register buffers used in
      // rotation) -> compute pass p.  (With one pass of slack the pass time settles at the L2 latency of
      // the index loads instead of the update's own dependent chain.)
#define OB_PS(K) (((K) <= np) ? (int)pstart[K] : 0)
#define OB_IDX(K, PA, PB) (((K) < np && (PA) + gl < (PB)) ? (int)sched[(PA) + gl] : -1)
      ObRowReg A, B, C, D;
      int ci = -1, i1 = -1, i2 = -1, i3 = -1, i4 = -1, i5 = -1, ps6 = 0, ps7 = 0, ps8 = 0, ps9 = 0;
      if (np > 0) {
        const int ps0 = OB_PS(0), ps1 = OB_PS(1), ps2 = OB_PS(2), ps3 = OB_PS(3), ps4 = OB_PS(4), ps5 = OB_PS(5);
        ps6 = OB_PS(6); ps7 = OB_PS(7); ps8 = OB_PS(8); ps9 = OB_PS(9);
        ci = OB_IDX(0, ps0, ps1); i1 = OB_IDX(1, ps1, ps2); i2 = OB_IDX(2, ps2, ps3); i3 = OB_IDX(3, ps3, ps4);
        i4 = OB_IDX(4, ps4, ps5); i5 = OB_IDX(5, ps5, ps6);
        if (ci >= 0) load_row(rows + (size_t)ci * OB_ROWW, A);
        if (i1 >= 0) load_row(rows + (size_t)i1 * OB_ROWW, B);
        if (i2 >= 0) load_row(rows + (size_t)i2 * OB_ROWW, C);
      }
      int p = 0;
#define OB_SOR_PASS(CB, LB)                                                                        \
      {                                                                                            \
        const int ps10 = OB_PS(p + 10);                                                            \
        const int i6 = OB_IDX(p + 6, ps6, ps7);                                                    \
        if (i3 >= 0) load_row(rows + (size_t)i3 * OB_ROWW, LB);                                    \
        if (taps & 4) sor_check_pass<G>(p < np && ci >= 0, CB.meta, gl, &d.world[wc].status);      \
        if (p < np && ci >= 0) sor_pass(CB, ci, s_fc, s_lam, s_invM);                              \
        __syncwarp();                                                                              \
        ci = i1; i1 = i2; i2 = i3; i3 = i4; i4 = i5; i5 = i6; ps6 = ps7; ps7 = ps8; ps8 = ps9; ps9 = ps10; p++; \
      }
      while (p < np_max) {
        OB_SOR_PASS(A, D)
        if (p >= np_max) break;
        OB_SOR_PASS(B, A)
        if (p >= np_max) break;
        OB_SOR_PASS(C, B)
        if (p >= np_max) break;
        OB_SOR_PASS(D, C)
      }
#undef OB_SOR_PASS
#undef OB_PS
#undef OB_IDX
      } else {
      // software pipeline: pstart[p+5] -> row index of pass p+3 -> row record of pass p+2 (three
      // register buffers used in rotation) -> compute pass p.  No load waits on another load.
      ObRowReg A, B, C;
      int ci = -1, i1 = -1, i2 = -1, ps3 = 0, ps4 = 0;
      if (np > 0) {
        const int ps0 = pstart[0], ps1 = pstart[1], ps2 = np >= 2 ? (int)pstart[2] : 0;
        ps3 = np >= 3 ? (int)pstart[3] : 0; ps4 = np >= 4 ? (int)pstart[4] : 0;
        if (ps0 + gl < ps1) { ci = sched[ps0 + gl]; load_row(rows + (size_t)ci * OB_ROWW, A); }
        if (np >= 2 && ps1 + gl < ps2) { i1 = sched[ps1 + gl]; load_row(rows + (size_t)i1 * OB_ROWW, B); }
        if (np >= 3 && ps2 + gl < ps3) i2 = sched[ps2 + gl];
      }
      int p = 0;
#define OB_SOR_PASS(CB, LB)                                                                        \
      {                                                                                            \
        const int ps5 = (p + 5 <= np) ? (int)pstart[p + 5] : 0;                                    \
        int i3 = -1;                                                                               \
        if (p + 3 < np && ps3 + gl < ps4) i3 = sched[ps3 + gl];                                    \
        if (i2 >= 0) load_row(rows + (size_t)i2 * OB_ROWW, LB);                                    \
        if (taps & 4) sor_check_pass<G>(p < np && ci >= 0, CB.meta, gl, &d.world[wc].status);      \
        if (p < np && ci >= 0) sor_pass(CB, ci, s_fc, s_lam, s_invM);                              \
        __syncwarp();                                                                              \
        ci = i1; i1 = i2; i2 = i3; ps3 = ps4; ps4 = ps5; p++;                                      \
      }
      while (p < np_max) {
        OB_SOR_PASS(A, C)
        if (p >= np_max) break;
        OB_SOR_PASS(B, A)
        if (p >= np_max) break;
        OB_SOR_PASS(C, B)
      }
#undef OB_SOR_PASS
      }
    }
    sor_epilogue<G>(d, taps, wc, valid, gl, nb, mtot, si, rows, s_fc, s_lam);
    __syncwarp();
  }
}

// =====================================================================================
// k_sor_ring<G, D>: the same sweep as k_sor, restructured around what actually bounds it.  ncu of k_sor<8, DEEP> (r01z)
// and of the first ring version (r02a/b) agree: one pass costs a warp ~1050 cycles for ~190 instructions although the
// math of a row update is ~60 flops -- the warp executes ONE long dependent chain per pass (schedule lookup -> row
// record -> meta decode -> invM / fc addresses -> fc loads -> J*Ad, invM*J products -> dot products -> clamp -> update
// -> stores), and with 4096 worlds = 1024 warps there are < 2 warps per scheduler to hide it behind.  So:
//   * rows come through a shared-memory ring filled by cp.async (LDGSTS): every lane copies ITS row of pass v + D - 1
//     straight from global/L2 into its own ring slot and later waits only for its own groups -- no register staging,
//     no cross-lane hand-off; the prefetcher also leaves the row's index in a header slot, so the consumer never looks
//     at the schedule;
//   * the epoch's schedule is staged in shared memory once per epoch (reused by its 8 iterations) as ONE u16 array:
//     row index | 0x8000 on the first slot of a pass -- the pass length is a ballot away, the pass table is not kept;
//   * SOFTWARE PIPELINE over passes: while pass v waits for its fc / lambda loads, the warp decodes the row of pass
//     v + 1 (ring -> registers, J*Ad, invM*J, bounds, addresses).  Only fc and lambda carry a dependency from pass to
//     pass, so the dependent chain of a pass shrinks to  fc loads -> dots -> clamp -> update -> stores;
//   * no branches in the pass: one-body rows and idle lanes run the same instructions on harmless operands, their
//     results are dropped by selects and predicated stores (a select never lets a garbage NaN through).
// The arithmetic of a row update is ob_sor_row()'s, operation for operation (same products, same association order):
// bit-identical results, checked by the parity suite and the OB_CHECK pass-disjointness tap.
struct SorRingSmem { size_t fc, invM, lam, idx, hdr, ring, total; };
// fc slot of body b: 8 words.  Bank-conflict swizzle (r02c/r02e: the shared-memory pipe is the busiest unit of the sweep, 58-65 % of
// its peak, half of the wavefronts are conflict replays): with the natural layout the 16-byte part of every body starts at bank
// 8b mod 32 -- four start positions for the lanes of a wavefront.  Bodies with bit 2 of b set keep their halves swapped
// ([4..7] = fc[0..3], [0..1] = fc[4..5]), which gives eight start positions for both the 16-byte and the 8-byte access.
__host__ __device__ inline SorRingSmem sor_ring_smem(int NB, int NR, int G, int D) {
  SorRingSmem s; size_t o = 0;
  s.fc = o; o = ob_al(o + sizeof(real) * 8 * NB, 16);                      // per body: fc[6] + 2 spare words, swizzled (ob_fc4 / ob_fc2)
  s.invM = o; o = ob_al(o + sizeof(real) * NB, 16);                        // dense: bank = b mod 32
  s.lam = o; o = ob_al(o + sizeof(real) * NR, 16);
  s.idx = o; o = ob_al(o + sizeof(unsigned short) * (NR + G + 2), 16);     // schedule of the epoch + sentinels
  s.hdr = o; o = ob_al(o + sizeof(unsigned short) * D * G, 16);            // row index of every ring slot (0xffff: idle)
  s.ring = o; o = ob_al(o + sizeof(real) * OB_ROWW * D * G, 16);
  s.total = ob_al(o, 16);
  return s;
}
__device__ __forceinline__ void ob_cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ob_cp_async16_sa(unsigned sa, const void *gsrc) {   // destination given as a shared-window address
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ob_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void ob_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// operands of one row update, prepared one pass ahead of their use
struct ObRowPrep {
  real Js[9];       // J1l, J1a, J2a scaled by Ad (quickstep.cpp:393-401); J2l*Ad == -Js[0..2]
  real iM1[3];      // invM1 * J1l      (iMJ of body 2's linear part is invM2 * (-J1l) == -iM2[])
  real iM2[3];      // invM2 * J1l
  real iMa[6];      // iMJ1a, iMJ2a as stored
  real b, adcfm, lo, hi;
  int o1, o2;       // the bodies (o2 == o1 for one-body rows)
  int li, lf;       // lambda index of the row, of its friction normal (== li when none)
  bool act, has2, fric;
};
__device__ __forceinline__ void sor_prep(const real *slot, int ci, const real *s_invM, ObRowPrep &P) {
  ObRowReg r;
#if defined(dSINGLE)
  const float4 *q = (const float4 *)slot;
#pragma unroll
  for (int i = 0; i < 4; i++) { const float4 t = q[i]; r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w; }
  { const float4 t = q[4]; r.v[16] = t.x; r.v[17] = t.y; r.v[18] = t.z; r.meta = __float_as_uint(t.w); }
#else
  const double2 *q = (const double2 *)slot;
#pragma unroll
  for (int i = 0; i < 9; i++) { const double2 t = q[i]; r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y; }
  { const double2 t = q[9]; r.v[18] = t.x; r.meta = (unsigned)__double2loint(t.y); }
#endif
  P.act = ci != 0xffff;
  const unsigned meta = P.act ? r.meta : 0u;
  const int b1 = meta & 255, b2r = (meta >> 8) & 255, fio = (meta >> 16) & 255, bmode = meta >> 24;
  P.has2 = b2r != 255;
  P.fric = fio != 0;
  const int b2 = P.has2 ? b2r : b1;
  P.o1 = b1; P.o2 = b2;
  P.li = P.act ? ci : 0;
  P.lf = P.fric ? P.li - fio : P.li;
  const real Ad = r.v[15], k1 = s_invM[b1], k2 = s_invM[b2];
  const real bv = r.v[18];
  P.lo = bmode == 0 ? -bv : (bmode == 1 ? (real)0 : bv);
  P.hi = bmode == 2 ? (real)0 : bv;
  P.b = r.v[16]; P.adcfm = r.v[17];
#pragma unroll
  for (int e = 0; e < 9; e++) P.Js[e] = r.v[e] * Ad;
#pragma unroll
  for (int e = 0; e < 3; e++) { P.iM1[e] = k1 * r.v[e]; P.iM2[e] = k2 * r.v[e]; }
#pragma unroll
  for (int e = 0; e < 6; e++) P.iMa[e] = r.v[9 + e];
}

template <int G, int D>
__global__ void __launch_bounds__(32) k_sor_ring(ObBatchDev d, int taps) {
  constexpr int T = 32 / G;
  constexpr int ROWB = OB_ROWW * (int)sizeof(real);
  extern __shared__ __align__(16) unsigned char smem_all[];
  const SorRingSmem L = sor_ring_smem(d.NB, d.NR, G, D);
  int lane = threadIdx.x;
  asm volatile("" : "+r"(lane));   // opaque: keeps lane-derived values in registers instead of re-reading SR_TID.X inside the pass (r02c: S2R + short-scoreboard stalls)
  const int grp = lane / G, gl = lane % G;
  const unsigned FULL = 0xffffffffu;
  const unsigned gmask = G == 32 ? FULL : ((1u << G) - 1u);
  const int gsh = grp * G;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  real *s_fc = (real *)(smem + L.fc);
  real *s_invM = (real *)(smem + L.invM);
  real *s_lam = (real *)(smem + L.lam);
  unsigned short *s_idx = (unsigned short *)(smem + L.idx);
  unsigned short *s_hdr = (unsigned short *)(smem + L.hdr) + gl;   // this lane's column: slot k at k * G
  unsigned char *s_ring = smem + L.ring + (size_t)gl * ROWB;       // this lane's column of the ring: slot k at k * G * ROWB
  unsigned ring_sa = (unsigned)__cvta_generic_to_shared(s_ring);   // the same as a shared-window address, converted once (cp.async destination)
  asm volatile("" : "+r"(ring_sa));

  for (int wbase = d.wbeg + blockIdx.x * T; wbase < d.wend; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.wend;
    const int wc = valid ? w : d.wbeg;
    const int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nb = valid ? d.world[wc].nb : 0;
    const int iters = valid ? d.world[wc].iters : 0;
    const int mtot = valid && si[SI_HAVEROWS] ? si[SI_MTOT] : 0;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    for (int b = gl; b < d.NB; b += G) {
#pragma unroll
      for (int k = 0; k < 8; k++) s_fc[8 * b + k] = 0;
      s_invM[b] = b < nb ? bc[b].invMass : (real)0;
    }
    for (int i = gl; i < mtot; i += G) s_lam[i] = 0;
    int nep = mtot > 0 ? (iters + 7) >> 3 : 0;
    if (nep > d.NEP) nep = d.NEP;
    const int nep_max = warp_max_i(nep);
    for (int ep = 0; ep < nep_max; ep++) {
      const bool epv = ep < nep;
      const int np = epv ? si[SI_NPASS0 + ep] : 0;
      int nit = epv ? iters - 8 * ep : 0;
      if (nit > 8) nit = 8;
      // stage the epoch's schedule: row index per slot, bit 15 on the first slot of every pass, sentinels behind the end
      __syncwarp();
      if (epv) {
        const unsigned short *sched = d.sched + ((size_t)wc * d.NEP + ep) * d.NR;
        for (int i = gl; i < mtot; i += G) s_idx[i] = sched[i];
        for (int i = gl; i <= G + 1; i += G) s_idx[mtot + i] = 0x8000;
      }
      __syncwarp();
      if (epv) {
        const unsigned short *pstart = d.pstart + ((size_t)wc * d.NEP + ep) * (d.NR + 1);
        for (int i = gl; i < np; i += G) s_idx[pstart[i]] |= 0x8000;
      }
      __syncwarp();
      const int vtot = np * nit;               // passes of this tile in this epoch (virtual pass v = it * np + p)
      const int vmax = warp_max_i(vtot);
      int pf_v = 0, pf_slot = 0, pf_s = 0;      // prefetcher: virtual pass, ring slot, first schedule slot of the pass
      // one prefetch step: find the row of this lane in the next pass, start its copy, leave its index in the header
#define OB_RING_ISSUE()                                                                            \
      {                                                                                            \
        const bool on_ = pf_v < vtot;                                                              \
        const unsigned me_ = on_ ? (unsigned)s_idx[pf_s + gl] : 0u;                                \
        const unsigned nx_ = on_ ? (unsigned)s_idx[pf_s + gl + 1] : 0x8000u;                       \
        const unsigned bits_ = (__ballot_sync(FULL, (nx_ & 0x8000u) != 0) >> gsh) & gmask;         \
        const int len_ = __ffs(bits_);           /* slots of this pass: next start within G slots */ \
        const bool mine_ = on_ && gl < len_;                                                       \
        const unsigned ri_ = me_ & 0x7fffu;                                                        \
        if (mine_) {                                                                               \
          const unsigned char *src_ = (const unsigned char *)(rows + (size_t)ri_ * OB_ROWW);       \
          const unsigned dst_ = ring_sa + (unsigned)pf_slot * (unsigned)(G * ROWB);                 \
          _Pragma("unroll") for (int c_ = 0; c_ < ROWB / 16; c_++) ob_cp_async16_sa(dst_ + 16u * c_, src_ + 16 * c_); \
        }                                                                                          \
        s_hdr[pf_slot * G] = (unsigned short)(mine_ ? ri_ : 0xffffu);                              \
        ob_cp_async_commit();                                                                      \
        if (on_) { pf_s += len_; if (pf_s >= mtot) pf_s = 0; }                                     \
        pf_v++;                                                                                    \
        if (++pf_slot == D) pf_slot = 0;                                                           \
      }
#pragma unroll 1
      for (int k = 0; k < D - 1; k++) OB_RING_ISSUE()
      ObRowPrep pa, pb;
      ob_cp_async_wait<D - 2>();                // group 0 has landed
      sor_prep((const real *)s_ring, (int)s_hdr[0], s_invM, pa);
      int n_slot = 1;                           // ring slot of pass v + 1
      // one pass: CUR is applied, NXT is decoded meanwhile; the loop alternates (pa, pb) / (pb, pa) so no operand set is copied
#define OB_RING_PASS(CUR, NXT)                                                                     \
      {                                                                                            \
        OB_RING_ISSUE()                                                                            \
        /* (B) this pass's dependent loads: fc of both bodies, lambda, lambda of the friction normal */ \
        real f1[6], f2[6];                                                                         \
        OB_LOAD_FC(f1, CUR.o1) OB_LOAD_FC(f2, CUR.o2)                                              \
        const real old_lambda = s_lam[CUR.li], lam_f = s_lam[CUR.lf];                              \
        /* (C) meanwhile: decode the row of the next pass */                                       \
        ob_cp_async_wait<D - 2>();              /* this lane's copy of pass v + 1 has landed (groups complete in order) */ \
        sor_prep((const real *)(s_ring + (size_t)n_slot * (G * ROWB)), (int)s_hdr[n_slot * G], s_invM, NXT); \
        if (++n_slot == D) n_slot = 0;                                                             \
        if (taps & 4) sor_check_pass<G>(CUR.act, (unsigned)CUR.o1 | ((CUR.has2 ? (unsigned)CUR.o2 : 255u) << 8), gl, &d.world[wc].status); \
        /* (D) the update, ob_sor_row() operation for operation (quickstep.cpp:490-581) */         \
        real delta = CUR.b - old_lambda * CUR.adcfm;                                               \
        delta -= f1[0] * CUR.Js[0] + f1[1] * CUR.Js[1] + f1[2] * CUR.Js[2] + f1[3] * CUR.Js[3] + f1[4] * CUR.Js[4] + f1[5] * CUR.Js[5]; \
        const real delta2 = delta - (f2[0] * -CUR.Js[0] + f2[1] * -CUR.Js[1] + f2[2] * -CUR.Js[2] + f2[3] * CUR.Js[6] + f2[4] * CUR.Js[7] + f2[5] * CUR.Js[8]); \
        delta = CUR.has2 ? delta2 : delta;                                                         \
        const real hf = ob_fabs(CUR.hi * lam_f);                                                   \
        const real hi_act = CUR.fric ? hf : CUR.hi, lo_act = CUR.fric ? -hf : CUR.lo;              \
        const real new_lambda = old_lambda + delta;                                                \
        real out = new_lambda;                                                                     \
        if (new_lambda < lo_act) { delta = lo_act - old_lambda; out = lo_act; }                    \
        else if (new_lambda > hi_act) { delta = hi_act - old_lambda; out = hi_act; }               \
        f1[0] += delta * CUR.iM1[0]; f1[1] += delta * CUR.iM1[1]; f1[2] += delta * CUR.iM1[2];     \
        f1[3] += delta * CUR.iMa[0]; f1[4] += delta * CUR.iMa[1]; f1[5] += delta * CUR.iMa[2];     \
        f2[0] += delta * -CUR.iM2[0]; f2[1] += delta * -CUR.iM2[1]; f2[2] += delta * -CUR.iM2[2];  \
        f2[3] += delta * CUR.iMa[3]; f2[4] += delta * CUR.iMa[4]; f2[5] += delta * CUR.iMa[5];     \
        if (CUR.act) {                                                                             \
          s_lam[CUR.li] = out;                                                                     \
          OB_STORE_FC(CUR.o1, f1)                                                                  \
          if (CUR.has2) OB_STORE_FC(CUR.o2, f2)                                                    \
        }                                                                                          \
        __syncwarp();                                                                              \
      }
#if defined(dSINGLE)
#define OB_LOAD_FC(F, B_) { const float4 a_ = *(const float4 *)(s_fc + ob_fc4(B_)); const float2 c_ = *(const float2 *)(s_fc + ob_fc2(B_)); F[0] = a_.x; F[1] = a_.y; F[2] = a_.z; F[3] = a_.w; F[4] = c_.x; F[5] = c_.y; }
#define OB_STORE_FC(B_, F) { *(float4 *)(s_fc + ob_fc4(B_)) = make_float4(F[0], F[1], F[2], F[3]); *(float2 *)(s_fc + ob_fc2(B_)) = make_float2(F[4], F[5]); }
#else
#define OB_LOAD_FC(F, B_) { const real *p4_ = s_fc + ob_fc4(B_), *p2_ = s_fc + ob_fc2(B_); F[0] = p4_[0]; F[1] = p4_[1]; F[2] = p4_[2]; F[3] = p4_[3]; F[4] = p2_[0]; F[5] = p2_[1]; }
#define OB_STORE_FC(B_, F) { real *p4_ = s_fc + ob_fc4(B_), *p2_ = s_fc + ob_fc2(B_); p4_[0] = F[0]; p4_[1] = F[1]; p4_[2] = F[2]; p4_[3] = F[3]; p2_[0] = F[4]; p2_[1] = F[5]; }
#endif
#pragma unroll 1
      for (int v = 0; v < vmax; v += 2) {
        OB_RING_PASS(pa, pb)
        if (v + 1 >= vmax) break;
        OB_RING_PASS(pb, pa)
      }
#undef OB_RING_PASS
#undef OB_LOAD_FC
#undef OB_STORE_FC
      ob_cp_async_wait<0>();
#undef OB_RING_ISSUE
    }
    __syncwarp();
    sor_epilogue<G>(d, taps, wc, valid, gl, nb, mtot, si, rows, s_fc, s_lam, 1);
    __syncwarp();
  }
}

// =====================================================================================
// k_sor_lane: ONE LANE PER WORLD, for batches of many tiny worlds (configs[2]: 65536 buggies of 5 bodies and ~56 rows).
// A tile of lanes per world finds 1-3 rows per level there, so most of a pass's ~170 instructions serve idle lanes, and at 228
// registers only 8 such warps fit an SM (r02o: 3.07 ms, 16 % of the issue slots).  Here a lane walks its world's rows one after the
// other in the schedule's order -- levels in sequence, rows of a level in any order: exactly the values of the reference's sweep,
// as in the tiled kernels -- so 32 worlds advance per warp instruction, no shuffles, no pass table.  fc and invMass sit in shared
// memory interleaved by lane (word (8 b + k) * 32 + lane: every lane owns its bank, whichever body it touches); lambda stays in
// global memory (each lane re-reads what it wrote itself; r02u: lambda in lane-interleaved shared memory costs the occupancy the
// loads need -- 27 KB per warp -- and was slower, 2.84 -> 3.87 ms); the row records are read straight from global memory, the index
// two rows and the record one row ahead of their use (r02v: an extra prefetch.global.L2 six rows ahead, 2.84 -> 3.75 ms; r02t: fewer
// worlds in flight so that their records stay in the L2, 2.99 / 3.55 / 4.00 ms at 1024 / 683 / 512 warps -- the kernel wants every warp it can get).
struct SorLaneSmem { size_t fc, invM, total; };
__host__ __device__ inline SorLaneSmem sor_lane_smem(int NB) {
  SorLaneSmem s; size_t o = 0;
  s.fc = o; o = ob_al(o + sizeof(real) * 8 * NB * 32, 16);
  s.invM = o; o = ob_al(o + sizeof(real) * NB * 32, 16);
  s.total = ob_al(o, 16);
  return s;
}
__device__ __forceinline__ void sor_pass_lane(const ObRowReg &cur, int cur_idx, real *s_fc, real *g_lam, const real *s_invM) {
  const int b1 = cur.meta & 255, b2r = (cur.meta >> 8) & 255, fio = (cur.meta >> 16) & 255, bmode = cur.meta >> 24;
  const int b2 = b2r == 255 ? -1 : b2r;
  const int fi = fio ? cur_idx - fio : -1;
  const real Ad = cur.v[15], k1 = s_invM[b1 * 32];
  const real bv = cur.v[18];
  const real lo = bmode == 0 ? -bv : (bmode == 1 ? (real)0 : bv);
  const real hi = bmode == 2 ? (real)0 : bv;
  real J[12], iMJ[12];
  // rebuild what SOR_LCP keeps per row: J scaled by Ad (quickstep.cpp:393-401), iMJ (:117-136) -- as sor_pass does
#pragma unroll
  for (int e = 0; e < 3; e++) {
    iMJ[e] = k1 * cur.v[e];
    iMJ[3 + e] = cur.v[9 + e];
    J[e] = cur.v[e] * Ad;
    J[3 + e] = cur.v[3 + e] * Ad;
  }
  real f1[6], f2[6];
  real *fp1 = s_fc + (size_t)(8 * b1) * 32;
#pragma unroll
  for (int e = 0; e < 6; e++) f1[e] = fp1[e * 32];
  real *fp2 = s_fc;
  if (b2 >= 0) {
    const real k2 = s_invM[b2 * 32];
#pragma unroll
    for (int e = 0; e < 3; e++) {
      const real j2l = -cur.v[e];
      iMJ[6 + e] = k2 * j2l;
      iMJ[9 + e] = cur.v[12 + e];
      J[6 + e] = j2l * Ad;
      J[9 + e] = cur.v[6 + e] * Ad;
    }
    fp2 = s_fc + (size_t)(8 * b2) * 32;
#pragma unroll
    for (int e = 0; e < 6; e++) f2[e] = fp2[e * 32];
  }
  const real lam_new = ob_sor_row(J, iMJ, cur.v[16], cur.v[17], lo, hi, fi, fi >= 0 ? g_lam[fi] : (real)0, g_lam[cur_idx], f1,
                                  b2 >= 0 ? f2 : (real *)0);
  g_lam[cur_idx] = lam_new;
#pragma unroll
  for (int e = 0; e < 6; e++) fp1[e * 32] = f1[e];
  if (b2 >= 0) {
#pragma unroll
    for (int e = 0; e < 6; e++) fp2[e * 32] = f2[e];
  }
}
__global__ void __launch_bounds__(32) k_sor_lane(ObBatchDev d, int taps) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SorLaneSmem L = sor_lane_smem(d.NB);
  const int lane = threadIdx.x;
  real *s_fc = (real *)(smem + L.fc) + lane;
  real *s_invM = (real *)(smem + L.invM) + lane;
  for (int wbase = d.wbeg + blockIdx.x * 32; wbase < d.wend; wbase += gridDim.x * 32) {
    const int w = wbase + lane;
    const bool valid = w < d.wend;
    const int wc = valid ? w : d.wbeg;
    const int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nb = valid ? d.world[wc].nb : 0;
    const int iters = valid ? d.world[wc].iters : 0;
    const int mtot = valid && si[SI_HAVEROWS] ? si[SI_MTOT] : 0;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    real *g_lam = d.lambda + (size_t)wc * d.NR;
    for (int b = 0; b < d.NB; b++) {
#pragma unroll
      for (int k = 0; k < 8; k++) s_fc[(size_t)(8 * b + k) * 32] = 0;
      s_invM[b * 32] = b < nb ? bc[b].invMass : (real)0;
    }
    for (int i = 0; i < mtot; i++) g_lam[i] = 0;
    const int iters_max = warp_max_i(mtot > 0 ? iters : 0);
    for (int it = 0; it < iters_max; it++) {
      const int ep = it >> 3;
      const int m = (mtot > 0 && it < iters) ? mtot : 0;
      const unsigned short *sched = d.sched + ((size_t)wc * d.NEP + (ep < d.NEP ? ep : d.NEP - 1)) * d.NR;
      const int m_max = warp_max_i(m);
      // the index two rows ahead, the record one row ahead
      int i1 = m > 0 ? (int)sched[0] : 0, i2 = m > 1 ? (int)sched[1] : 0;
      ObRowReg nxt;
      load_row(rows + (size_t)i1 * OB_ROWW, nxt);
      int ni = i1;
      for (int sx = 0; sx < m_max; sx++) {
        const ObRowReg cur = nxt;
        const int ci = ni;
        const int i3 = sx + 2 < m ? (int)sched[sx + 2] : 0;
        ni = i2;
        if (sx + 1 < m) load_row(rows + (size_t)i2 * OB_ROWW, nxt);
        i2 = i3;
        if (sx < m) sor_pass_lane(cur, ci, s_fc, g_lam, s_invM);
      }
    }
    // cforce per body for k_post (tmp1 is dead after row assembly); lambda is where the taps expect it already
    real *g_fc = d.tmp1 + (size_t)wc * d.NB * 8;
    if (valid)
      for (int b = 0; b < nb; b++)
        for (int k = 0; k < 6; k++) g_fc[8 * b + k] = s_fc[(size_t)(8 * b + k) * 32];
    if (taps && valid && mtot > 0) {   // joint feedback (quickstep.cpp:918-957, Multiply1_12q1), as sor_epilogue
      const int nij = si[SI_NIJ];
      const unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + d.NJ + 1);
      const unsigned short *g_ijoint = d.ijoint + (size_t)wc * (d.NC + d.NJ);
      const int ncw = d.ncontacts[wc];
      real *fb = d.fback + (size_t)wc * (d.NC + d.NJ) * 12;
      for (int k = 0; k < nij; k++) {
        const int jr0 = g_jrow[k], jm = g_jrow[k + 1] - jr0;
        real acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool two = false;
        for (int q = 0; q < jm; q++) {
          const real *rp = rows + (size_t)(jr0 + q) * OB_ROWW;
          const real sl = g_lam[jr0 + q];
          const unsigned meta = *(const unsigned *)(rp + OB_ROWF);
          two = ((meta >> 8) & 255u) != 255u;
          for (int e = 0; e < 6; e++) acc[e] += rp[e] * sl;
          for (int e = 0; e < 3; e++) { acc[6 + e] += (-rp[e]) * sl; acc[9 + e] += rp[6 + e] * sl; }
        }
        if (!two) for (int e = 6; e < 12; e++) acc[e] = 0;
        const int jid = g_ijoint[k];
        real *o = fb + (size_t)(jid < ncw ? jid : d.NC + (jid - ncw)) * 12;
        for (int e = 0; e < 12; e++) o[e] = acc[e];
      }
    }
    __syncwarp();
  }
}

// =====================================================================================
// k_sor_reg<G>: the pipelined, branch-free pass of k_sor_ring with the row records prefetched into REGISTERS again.
// ncu of the ring kernel (r02c-r02f): the shared-memory pipe is its busiest unit (55-65 % of peak wavefronts) and more than
// half of that is the ring itself -- every 80-byte record crosses the pipe twice (LDGSTS in, LDS out) -- while the loads
// that depend on shared memory (short-scoreboard) top the stall list.  Here every lane loads ITS row of pass v + 3 with
// five 16-byte LDG.cg straight into one of three rotating register buffers; the decode of pass v + 1 reads the buffer that was
// filled two passes earlier (the hardware scoreboard is the only synchronisation).  What stays in shared memory is what
// the dependency chain runs through (fc, lambda) plus the staged schedule.  The loop is unrolled 6x = lcm(3 buffers, 2
// operand sets) so that every buffer and operand set is addressed statically.
struct SorRegSmem { size_t fc, invM, lam, idx, total; };
__host__ __device__ inline SorRegSmem sor_reg_smem(int NB, int NR, int G) {
  SorRegSmem s; size_t o = 0;
  s.fc = o; o = ob_al(o + sizeof(real) * 8 * NB, 16);
  s.invM = o; o = ob_al(o + sizeof(real) * NB, 16);
  s.lam = o; o = ob_al(o + sizeof(real) * NR, 16);
  s.idx = o; o = ob_al(o + sizeof(unsigned short) * (NR + G + 2), 16);
  s.total = ob_al(o, 16);
  return s;
}
struct ObRowBuf { ObRowReg r; int ci; };   // a prefetched row + its index (0xffff: this lane idles in that pass)
__device__ __forceinline__ void load_row_cg(const real *__restrict__ p, ObRowReg &r) {
#if defined(dSINGLE)
  const float4 *q = (const float4 *)p;
#pragma unroll
  for (int i = 0; i < 4; i++) { const float4 t = __ldcg(q + i); r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w; }
  const float4 t = __ldcg(q + 4);
  r.v[16] = t.x; r.v[17] = t.y; r.v[18] = t.z; r.meta = __float_as_uint(t.w);
#else
  const double2 *q = (const double2 *)p;
#pragma unroll
  for (int i = 0; i < 9; i++) { const double2 t = __ldcg(q + i); r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y; }
  const double2 t = __ldcg(q + 9);
  r.v[18] = t.x; r.meta = (unsigned)__double2loint(t.y);
#endif
}
__device__ __forceinline__ void sor_prep_reg(const ObRowBuf &B, const real *s_invM, ObRowPrep &P) {
  const ObRowReg &r = B.r;
  const int ci = B.ci;
  P.act = ci != 0xffff;
  const unsigned meta = P.act ? r.meta : 0u;
  const int b1 = meta & 255, b2r = (meta >> 8) & 255, fio = (meta >> 16) & 255, bmode = meta >> 24;
  P.has2 = b2r != 255;
  P.fric = fio != 0;
  const int b2 = P.has2 ? b2r : b1;
  P.o1 = b1; P.o2 = b2;
  P.li = P.act ? ci : 0;
  P.lf = P.fric ? P.li - fio : P.li;
  const real Ad = r.v[15], k1 = s_invM[b1], k2 = s_invM[b2];
  const real bv = r.v[18];
  P.lo = bmode == 0 ? -bv : (bmode == 1 ? (real)0 : bv);
  P.hi = bmode == 2 ? (real)0 : bv;
  P.b = r.v[16]; P.adcfm = r.v[17];
#pragma unroll
  for (int e = 0; e < 9; e++) P.Js[e] = r.v[e] * Ad;
#pragma unroll
  for (int e = 0; e < 3; e++) { P.iM1[e] = k1 * r.v[e]; P.iM2[e] = k2 * r.v[e]; }
#pragma unroll
  for (int e = 0; e < 6; e++) P.iMa[e] = r.v[9 + e];
}

template <int G>
__global__ void __launch_bounds__(32) k_sor_reg(ObBatchDev d, int taps) {
  constexpr int T = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const SorRegSmem L = sor_reg_smem(d.NB, d.NR, G);
  int lane = threadIdx.x;
  asm volatile("" : "+r"(lane));
  const int grp = lane / G, gl = lane % G;
  const unsigned FULL = 0xffffffffu;
  const unsigned gmask = G == 32 ? FULL : ((1u << G) - 1u);
  const int gsh = grp * G;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  real *s_fc = (real *)(smem + L.fc);
  real *s_invM = (real *)(smem + L.invM);
  real *s_lam = (real *)(smem + L.lam);
  unsigned short *s_idx = (unsigned short *)(smem + L.idx);

  for (int wbase = d.wbeg + blockIdx.x * T; wbase < d.wend; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.wend;
    const int wc = valid ? w : d.wbeg;
    const int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nb = valid ? d.world[wc].nb : 0;
    const int iters = valid ? d.world[wc].iters : 0;
    const int mtot = valid && si[SI_HAVEROWS] ? si[SI_MTOT] : 0;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    for (int b = gl; b < d.NB; b += G) {
#pragma unroll
      for (int k = 0; k < 8; k++) s_fc[8 * b + k] = 0;
      s_invM[b] = b < nb ? bc[b].invMass : (real)0;
    }
    for (int i = gl; i < mtot; i += G) s_lam[i] = 0;
    int nep = mtot > 0 ? (iters + 7) >> 3 : 0;
    if (nep > d.NEP) nep = d.NEP;
    const int nep_max = warp_max_i(nep);
    for (int ep = 0; ep < nep_max; ep++) {
      const bool epv = ep < nep;
      const int np = epv ? si[SI_NPASS0 + ep] : 0;
      int nit = epv ? iters - 8 * ep : 0;
      if (nit > 8) nit = 8;
      __syncwarp();
      if (epv) {
        const unsigned short *sched = d.sched + ((size_t)wc * d.NEP + ep) * d.NR;
        for (int i = gl; i < mtot; i += G) s_idx[i] = sched[i];
        for (int i = gl; i <= G + 1; i += G) s_idx[mtot + i] = 0x8000;
      }
      __syncwarp();
      if (epv) {
        const unsigned short *pstart = d.pstart + ((size_t)wc * d.NEP + ep) * (d.NR + 1);
        for (int i = gl; i < np; i += G) s_idx[pstart[i]] |= 0x8000;
      }
      __syncwarp();
      const int vtot = np * nit;
      const int vmax = warp_max_i(vtot);
      int pf_v = 0, pf_s = 0;
      // one prefetch step into buffer BUF: find this lane's row of the next pass and start its five 16-byte loads
#define OB_REG_FETCH(BUF)                                                                          \
      {                                                                                            \
        const bool on_ = pf_v < vtot;                                                              \
        const unsigned me_ = on_ ? (unsigned)s_idx[pf_s + gl] : 0u;                                \
        const unsigned nx_ = on_ ? (unsigned)s_idx[pf_s + gl + 1] : 0x8000u;                       \
        const unsigned bits_ = (__ballot_sync(FULL, (nx_ & 0x8000u) != 0) >> gsh) & gmask;         \
        const int len_ = __ffs(bits_);                                                             \
        const bool mine_ = on_ && gl < len_;                                                       \
        const unsigned ri_ = mine_ ? (me_ & 0x7fffu) : 0u;   /* idle lanes load row 0 of their world (valid memory, result unused) */ \
        load_row_cg(rows + (size_t)ri_ * OB_ROWW, BUF.r);                                          \
        BUF.ci = mine_ ? (int)ri_ : 0xffff;                                                        \
        if (on_) { pf_s += len_; if (pf_s >= mtot) pf_s = 0; }                                     \
        pf_v++;                                                                                    \
      }
      ObRowBuf r0, r1, r2;
      ObRowPrep pa, pb;
      if (vmax > 0) {
        OB_REG_FETCH(r0) OB_REG_FETCH(r1) OB_REG_FETCH(r2)
        sor_prep_reg(r0, s_invM, pa);
      }
      // pass v: CUR is applied; FB (the buffer of pass v, already decoded) is refilled with the row of pass v + 3; NXT is decoded from NB
#define OB_REG_PASS(CUR, NXT, FB, NBUF)                                                            \
      {                                                                                            \
        real f1[6], f2[6];                                                                         \
        OB_LOAD_FC(f1, CUR.o1) OB_LOAD_FC(f2, CUR.o2)                                              \
        const real old_lambda = s_lam[CUR.li], lam_f = s_lam[CUR.lf];                              \
        OB_REG_FETCH(FB)                                                                           \
        sor_prep_reg(NBUF, s_invM, NXT);                                                           \
        if (taps & 4) sor_check_pass<G>(CUR.act, (unsigned)CUR.o1 | ((CUR.has2 ? (unsigned)CUR.o2 : 255u) << 8), gl, &d.world[wc].status); \
        real delta = CUR.b - old_lambda * CUR.adcfm;                                               \
        delta -= f1[0] * CUR.Js[0] + f1[1] * CUR.Js[1] + f1[2] * CUR.Js[2] + f1[3] * CUR.Js[3] + f1[4] * CUR.Js[4] + f1[5] * CUR.Js[5]; \
        const real delta2 = delta - (f2[0] * -CUR.Js[0] + f2[1] * -CUR.Js[1] + f2[2] * -CUR.Js[2] + f2[3] * CUR.Js[6] + f2[4] * CUR.Js[7] + f2[5] * CUR.Js[8]); \
        delta = CUR.has2 ? delta2 : delta;                                                         \
        const real hf = ob_fabs(CUR.hi * lam_f);                                                   \
        const real hi_act = CUR.fric ? hf : CUR.hi, lo_act = CUR.fric ? -hf : CUR.lo;              \
        const real new_lambda = old_lambda + delta;                                                \
        real out = new_lambda;                                                                     \
        if (new_lambda < lo_act) { delta = lo_act - old_lambda; out = lo_act; }                    \
        else if (new_lambda > hi_act) { delta = hi_act - old_lambda; out = hi_act; }               \
        f1[0] += delta * CUR.iM1[0]; f1[1] += delta * CUR.iM1[1]; f1[2] += delta * CUR.iM1[2];     \
        f1[3] += delta * CUR.iMa[0]; f1[4] += delta * CUR.iMa[1]; f1[5] += delta * CUR.iMa[2];     \
        f2[0] += delta * -CUR.iM2[0]; f2[1] += delta * -CUR.iM2[1]; f2[2] += delta * -CUR.iM2[2];  \
        f2[3] += delta * CUR.iMa[3]; f2[4] += delta * CUR.iMa[4]; f2[5] += delta * CUR.iMa[5];     \
        if (CUR.act) {                                                                             \
          s_lam[CUR.li] = out;                                                                     \
          OB_STORE_FC(CUR.o1, f1)                                                                  \
          if (CUR.has2) OB_STORE_FC(CUR.o2, f2)                                                    \
        }                                                                                          \
        __syncwarp();                                                                              \
      }
#if defined(dSINGLE)
#define OB_LOAD_FC(F, B_) { const float4 a_ = *(const float4 *)(s_fc + ob_fc4(B_)); const float2 c_ = *(const float2 *)(s_fc + ob_fc2(B_)); F[0] = a_.x; F[1] = a_.y; F[2] = a_.z; F[3] = a_.w; F[4] = c_.x; F[5] = c_.y; }
#define OB_STORE_FC(B_, F) { *(float4 *)(s_fc + ob_fc4(B_)) = make_float4(F[0], F[1], F[2], F[3]); *(float2 *)(s_fc + ob_fc2(B_)) = make_float2(F[4], F[5]); }
#else
#define OB_LOAD_FC(F, B_) { const real *p4_ = s_fc + ob_fc4(B_), *p2_ = s_fc + ob_fc2(B_); F[0] = p4_[0]; F[1] = p4_[1]; F[2] = p4_[2]; F[3] = p4_[3]; F[4] = p2_[0]; F[5] = p2_[1]; }
#define OB_STORE_FC(B_, F) { real *p4_ = s_fc + ob_fc4(B_), *p2_ = s_fc + ob_fc2(B_); p4_[0] = F[0]; p4_[1] = F[1]; p4_[2] = F[2]; p4_[3] = F[3]; p2_[0] = F[4]; p2_[1] = F[5]; }
#endif
      // pass v: operands (pa, pb) alternate with v mod 2, buffers with v mod 3: pass v refills buffer v mod 3, decodes buffer (v + 1) mod 3
#pragma unroll 1
      for (int v = 0; v < vmax; v += 6) {
        OB_REG_PASS(pa, pb, r0, r1)
        if (v + 1 >= vmax) break;
        OB_REG_PASS(pb, pa, r1, r2)
        if (v + 2 >= vmax) break;
        OB_REG_PASS(pa, pb, r2, r0)
        if (v + 3 >= vmax) break;
        OB_REG_PASS(pb, pa, r0, r1)
        if (v + 4 >= vmax) break;
        OB_REG_PASS(pa, pb, r1, r2)
        if (v + 5 >= vmax) break;
        OB_REG_PASS(pb, pa, r2, r0)
      }
#undef OB_REG_PASS
#undef OB_LOAD_FC
#undef OB_STORE_FC
#undef OB_REG_FETCH
    }
    __syncwarp();
    sor_epilogue<G>(d, taps, wc, valid, gl, nb, mtot, si, rows, s_fc, s_lam, 1);
    __syncwarp();
  }
}

// =====================================================================================
// k_sor_pair<G, D>: k_sor_ring with TWO LANES PER ROW (G lanes per world = G/2 rows per pass; lane 2r works on body 1 of the
// pass's r-th row, lane 2r+1 on body 2).  The ring kernel showed what bounds the sweep on configs[1]: 4096 worlds x 4 per
// warp = 1024 warps, 1.7 per scheduler, every one of them an in-order chain of ~215 instructions per pass at ~5 cycles
// each (ncu r02c/r02d: wait + short-scoreboard + branch stalls, 0.29 IPC per scheduler, memory idle).  Splitting a row
// over a lane pair halves the chain a lane runs (half the decode, one 6-term dot product, one 6-vector update) and
// doubles the warps that hide it.  The arithmetic is unchanged: lane A forms fc1.J1, lane B fc2.J2 (same association
// order), one shuffle exchanges them and BOTH lanes evaluate  delta -= dot1; delta -= dot2; clamp  exactly as
// ob_sor_row() does, so they agree bit for bit and each applies delta to its own body.
// Row copies: lane A starts the copy of the row's first 48 bytes, lane B of the last 32; a lane may read its partner's
// part only after the partner's wait_group AND a __syncwarp, so the wait for pass v + 2 sits in iteration v (the pass's
// closing __syncwarp publishes it) and the decode of pass v + 1 runs one iteration behind the wait.  D >= 4.
struct ObHalfPrep {
  real Jo[6];       // this lane's half of the row's J, scaled by Ad (body 2: linear part negated)
  real iMo[6];      // this lane's half of iMJ (body 2: linear part negated)
  real b, adcfm, lo, hi;
  int off;          // this lane's body
  int li, lf;       // lambda index of the row, of its friction normal (== li when none)
  bool act, has2, fric;
};
__device__ __forceinline__ void sor_prep_half(const real *slot, int ci, int h, const real *s_invM, ObHalfPrep &P) {
  // words: [0-2 J1l][3-5 J1a][6-8 J2a][9-11 iMJ1a][12-14 iMJ2a][15 Ad][16 b][17 Ad*cfm][18 bound][19 meta]
  real lin[3], ang[3], ima[3];
#pragma unroll
  for (int e = 0; e < 3; e++) { lin[e] = slot[e]; ang[e] = h ? slot[6 + e] : slot[3 + e]; ima[e] = h ? slot[12 + e] : slot[9 + e]; }
  const real Ad = slot[15], bv = slot[18];
  P.b = slot[16]; P.adcfm = slot[17];
#if defined(dSINGLE)
  const unsigned rmeta = __float_as_uint(slot[19]);
#else
  const unsigned rmeta = (unsigned)__double2loint(slot[19]);
#endif
  P.act = ci != 0xffff;
  const unsigned meta = P.act ? rmeta : 0u;
  const int b1 = meta & 255, b2r = (meta >> 8) & 255, fio = (meta >> 16) & 255, bmode = meta >> 24;
  P.has2 = b2r != 255;
  P.fric = fio != 0;
  const int bo = (h && P.has2) ? b2r : b1;
  P.off = bo;
  P.li = P.act ? ci : 0;
  P.lf = P.fric ? P.li - fio : P.li;
  const real k = s_invM[bo];
  P.lo = bmode == 0 ? -bv : (bmode == 1 ? (real)0 : bv);
  P.hi = bmode == 2 ? (real)0 : bv;
#pragma unroll
  for (int e = 0; e < 3; e++) {
    const real l = h ? -lin[e] : lin[e];      // J2l == -J1l
    P.Jo[e] = l * Ad;                          // (-x) * Ad == -(x * Ad): same bits as the reference's J2l * Ad
    P.Jo[3 + e] = ang[e] * Ad;
    P.iMo[e] = k * l;                          // invM2 * (-J1l)
    P.iMo[3 + e] = ima[e];
  }
}

template <int G, int D>
__global__ void __launch_bounds__(32) k_sor_pair(ObBatchDev d, int taps) {
  constexpr int T = 32 / G, R = G / 2;
  constexpr int ROWB = OB_ROWW * (int)sizeof(real);
  extern __shared__ __align__(16) unsigned char smem_all[];
  const SorRingSmem L = sor_ring_smem(d.NB, d.NR, R, D);
  int lane = threadIdx.x;
  asm volatile("" : "+r"(lane));
  const int grp = lane / G, gl = lane % G, r = gl >> 1, h = gl & 1;
  const unsigned FULL = 0xffffffffu;
  const unsigned gmask = G == 32 ? FULL : ((1u << G) - 1u);
  const int gsh = grp * G;
  unsigned char *smem = smem_all + (size_t)grp * L.total;
  real *s_fc = (real *)(smem + L.fc);
  real *s_invM = (real *)(smem + L.invM);
  real *s_lam = (real *)(smem + L.lam);
  unsigned short *s_idx = (unsigned short *)(smem + L.idx);
  unsigned short *s_hdr = (unsigned short *)(smem + L.hdr) + r;    // this row slot's column: ring slot k at k * R
  unsigned char *s_ring = smem + L.ring + (size_t)r * ROWB;        // this row slot's column of the ring: ring slot k at k * R * ROWB
  unsigned ring_sa = (unsigned)__cvta_generic_to_shared(s_ring) + (h ? 48u : 0u);   // lane A copies bytes [0, 48), lane B [48, ROWB)
  asm volatile("" : "+r"(ring_sa));

  for (int wbase = d.wbeg + blockIdx.x * T; wbase < d.wend; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.wend;
    const int wc = valid ? w : d.wbeg;
    const int *si = d.stepinfo + (size_t)wc * SI_WORDS;
    const int nb = valid ? d.world[wc].nb : 0;
    const int iters = valid ? d.world[wc].iters : 0;
    const int mtot = valid && si[SI_HAVEROWS] ? si[SI_MTOT] : 0;
    const ObBodyConst *bc = d.bconst + (size_t)wc * d.NB;
    const real *rows = d.rows + (size_t)wc * d.NR * OB_ROWW;
    for (int b = gl; b < d.NB; b += G) {
#pragma unroll
      for (int k = 0; k < 8; k++) s_fc[8 * b + k] = 0;
      s_invM[b] = b < nb ? bc[b].invMass : (real)0;
    }
    for (int i = gl; i < mtot; i += G) s_lam[i] = 0;
    int nep = mtot > 0 ? (iters + 7) >> 3 : 0;
    if (nep > d.NEP) nep = d.NEP;
    const int nep_max = warp_max_i(nep);
    for (int ep = 0; ep < nep_max; ep++) {
      const bool epv = ep < nep;
      const int np = epv ? si[SI_NPASS0 + ep] : 0;
      int nit = epv ? iters - 8 * ep : 0;
      if (nit > 8) nit = 8;
      __syncwarp();
      if (epv) {
        const unsigned short *sched = d.sched + ((size_t)wc * d.NEP + ep) * d.NR;
        for (int i = gl; i < mtot; i += G) s_idx[i] = sched[i];
        for (int i = gl; i <= R + 1; i += G) s_idx[mtot + i] = 0x8000;
      }
      __syncwarp();
      if (epv) {
        const unsigned short *pstart = d.pstart + ((size_t)wc * d.NEP + ep) * (d.NR + 1);
        for (int i = gl; i < np; i += G) s_idx[pstart[i]] |= 0x8000;
      }
      __syncwarp();
      const int vtot = np * nit;
      const int vmax = warp_max_i(vtot);
      int pf_v = 0, pf_slot = 0, pf_s = 0;
#define OB_PAIR_ISSUE()                                                                            \
      {                                                                                            \
        const bool on_ = pf_v < vtot;                                                              \
        const unsigned me_ = on_ ? (unsigned)s_idx[pf_s + r] : 0u;                                 \
        const unsigned nx_ = on_ ? (unsigned)s_idx[pf_s + r + 1] : 0x8000u;                        \
        const unsigned bits_ = (__ballot_sync(FULL, h == 0 && (nx_ & 0x8000u) != 0) >> gsh) & gmask; \
        const int len_ = (__ffs(bits_) + 1) >> 1;   /* rows of this pass: first even lane whose next slot starts a pass */ \
        const bool mine_ = on_ && r < len_;                                                        \
        const unsigned ri_ = me_ & 0x7fffu;                                                        \
        if (mine_) {                                                                               \
          const unsigned char *src_ = (const unsigned char *)(rows + (size_t)ri_ * OB_ROWW) + (h ? 48 : 0); \
          const unsigned dst_ = ring_sa + (unsigned)pf_slot * (unsigned)(R * ROWB);                \
          if (h == 0) { _Pragma("unroll") for (int c_ = 0; c_ < 3; c_++) ob_cp_async16_sa(dst_ + 16u * c_, src_ + 16 * c_); } \
          else { _Pragma("unroll") for (int c_ = 0; c_ < (ROWB - 48) / 16; c_++) ob_cp_async16_sa(dst_ + 16u * c_, src_ + 16 * c_); } \
        }                                                                                          \
        if (h == 0) s_hdr[pf_slot * R] = (unsigned short)(mine_ ? ri_ : 0xffffu);                  \
        ob_cp_async_commit();                                                                      \
        if (on_) { pf_s += len_; if (pf_s >= mtot) pf_s = 0; }                                     \
        pf_v++;                                                                                    \
        if (++pf_slot == D) pf_slot = 0;                                                           \
      }
#pragma unroll 1
      for (int k = 0; k < D - 1; k++) OB_PAIR_ISSUE()
      ObHalfPrep pa, pb;
      ob_cp_async_wait<D - 3>();                // groups 0 and 1 have landed (this lane's parts)
      __syncwarp();                             // ... and the partner's
      sor_prep_half((const real *)s_ring, (int)s_hdr[0], h, s_invM, pa);
      int n_slot = 1;
#define OB_PAIR_PASS(CUR, NXT)                                                                     \
      {                                                                                            \
        OB_PAIR_ISSUE()                                                                            \
        real f[6];                                                                                 \
        OB_LOAD_FC(f, CUR.off)                                                                     \
        const real old_lambda = s_lam[CUR.li], lam_f = s_lam[CUR.lf];                              \
        ob_cp_async_wait<D - 3>();              /* pass v + 2 landed; published by this pass's closing __syncwarp */ \
        sor_prep_half((const real *)(s_ring + (size_t)n_slot * (R * ROWB)), (int)s_hdr[n_slot * R], h, s_invM, NXT); \
        if (++n_slot == D) n_slot = 0;                                                             \
        const real dot = f[0] * CUR.Jo[0] + f[1] * CUR.Jo[1] + f[2] * CUR.Jo[2] + f[3] * CUR.Jo[3] + f[4] * CUR.Jo[4] + f[5] * CUR.Jo[5]; \
        const real oth = __shfl_xor_sync(FULL, dot, 1);                                            \
        const real d1 = h ? oth : dot, d2 = h ? dot : oth;                                         \
        real delta = CUR.b - old_lambda * CUR.adcfm;                                               \
        delta -= d1;                                                                               \
        const real delta2 = delta - d2;                                                            \
        delta = CUR.has2 ? delta2 : delta;                                                         \
        const real hf = ob_fabs(CUR.hi * lam_f);                                                   \
        const real hi_act = CUR.fric ? hf : CUR.hi, lo_act = CUR.fric ? -hf : CUR.lo;              \
        const real new_lambda = old_lambda + delta;                                                \
        real out = new_lambda;                                                                     \
        if (new_lambda < lo_act) { delta = lo_act - old_lambda; out = lo_act; }                    \
        else if (new_lambda > hi_act) { delta = hi_act - old_lambda; out = hi_act; }               \
        f[0] += delta * CUR.iMo[0]; f[1] += delta * CUR.iMo[1]; f[2] += delta * CUR.iMo[2];        \
        f[3] += delta * CUR.iMo[3]; f[4] += delta * CUR.iMo[4]; f[5] += delta * CUR.iMo[5];        \
        if (CUR.act && (h == 0 || CUR.has2)) {                                                     \
          if (h == 0) s_lam[CUR.li] = out;                                                         \
          OB_STORE_FC(CUR.off, f)                                                                  \
        }                                                                                          \
        __syncwarp();                                                                              \
      }
#if defined(dSINGLE)
#define OB_LOAD_FC(F, B_) { const float4 a_ = *(const float4 *)(s_fc + ob_fc4(B_)); const float2 c_ = *(const float2 *)(s_fc + ob_fc2(B_)); F[0] = a_.x; F[1] = a_.y; F[2] = a_.z; F[3] = a_.w; F[4] = c_.x; F[5] = c_.y; }
#define OB_STORE_FC(B_, F) { *(float4 *)(s_fc + ob_fc4(B_)) = make_float4(F[0], F[1], F[2], F[3]); *(float2 *)(s_fc + ob_fc2(B_)) = make_float2(F[4], F[5]); }
#else
#define OB_LOAD_FC(F, B_) { const real *p4_ = s_fc + ob_fc4(B_), *p2_ = s_fc + ob_fc2(B_); F[0] = p4_[0]; F[1] = p4_[1]; F[2] = p4_[2]; F[3] = p4_[3]; F[4] = p2_[0]; F[5] = p2_[1]; }
#define OB_STORE_FC(B_, F) { real *p4_ = s_fc + ob_fc4(B_), *p2_ = s_fc + ob_fc2(B_); p4_[0] = F[0]; p4_[1] = F[1]; p4_[2] = F[2]; p4_[3] = F[3]; p2_[0] = F[4]; p2_[1] = F[5]; }
#endif
#pragma unroll 1
      for (int v = 0; v < vmax; v += 2) {
        OB_PAIR_PASS(pa, pb)
        if (v + 1 >= vmax) break;
        OB_PAIR_PASS(pb, pa)
      }
#undef OB_PAIR_PASS
#undef OB_LOAD_FC
#undef OB_STORE_FC
      ob_cp_async_wait<0>();
#undef OB_PAIR_ISSUE
    }
    __syncwarp();
    sor_epilogue<G>(d, taps, wc, valid, gl, nb, mtot, si, rows, s_fc, s_lam, 1);
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

  for (int wbase = d.wbeg + blockIdx.x * T; wbase < d.wend; wbase += gridDim.x * T) {
    const int w = wbase + grp;
    const bool valid = w < d.wend;
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
    const unsigned short *g_jrow = d.jrow + (size_t)wc * (d.NC + d.NJ + 1);
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
    const bool sap = valid && W.space_type == OB_SPACE_SAP;
    for (int i = gl; i < ng; i += G) s_old[i] = (unsigned short)glist[i];
    __syncwarp();
    if (sap) {
      // dxSAPSpace::dirty per moved geom (collision_sapspace.cpp:363-387): swap-remove from the GeomList,
      // append to the DirtyList; stored as DirtyList followed by GeomList.  (All geoms are clean here:
      // k_collide ran cleanGeoms.)  The removals form a serial chain, one lane walks it.
      if (gl == 0) {
        for (int i = 0; i < ng; i++) s_flag[s_old[i]] = (unsigned char)i;   // position in the GeomList
        int n = ng;
        for (int i = 0; i < nm; i++) {
          const int g = s_moved[i], idx = s_flag[g], last = s_old[n - 1];
          s_old[idx] = (unsigned short)last; s_flag[last] = (unsigned char)idx;
          n--;
        }
        W.sap_ndirty = nm;
      }
      __syncwarp();
      for (int i = gl; i < nm; i += G) glist[i] = s_moved[i];
      for (int i = gl; i < ng - nm; i += G) glist[nm + i] = s_old[i];
    } else {
      for (int i = gl; i < nm; i += G) s_flag[s_moved[i]] = 1;
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
    }
    // contacts solved = contact joints that entered an island with rows (permanent joints are not contacts)
    int ncj = 0;
    if (have_rows) {
      const unsigned short *g_ijoint = d.ijoint + (size_t)wc * (d.NC + d.NJ);
      const int ncw = d.ncontacts[wc];
      for (int k = gl; k < nij; k += G) ncj += g_ijoint[k] < ncw ? 1 : 0;
    }
    for (int dd = 1; dd < G; dd <<= 1) ncj += __shfl_xor_sync(0xffffffffu, ncj, dd, G);
    if (gl == 0 && valid) {
      d.nrows[w] = mtot;
      atomicAdd(&d.counters->steps, 1ull);
      atomicAdd(&d.counters->body_steps, (unsigned long long)nib);
      atomicAdd(&d.counters->contacts, (unsigned long long)ncj);
      atomicAdd(&d.counters->rows, (unsigned long long)mtot);
      atomicAdd(&d.counters->islands, (unsigned long long)nis);
      if (W.status) atomicAdd(&d.counters->overflow_worlds, 1ull);
    }
    __syncwarp();
  }
}

// ob_large.h — one LARGE world (200 k-body pile, dSweepAndPruneSpace):
// data layout and the per-element functions of the grid-wide pipeline.  The CUDA kernels
// (ob_large_kernels.cuh) and the test-only sequential mirror (tests/hostsim) call the SAME
// functions, so the mirror reproduces the GPU result bit for bit and can be checked against the
// unmodified reference on the CPU.
//
// What differs from the batched small-world path, and why:
//   * the world does not fit a CTA, so every phase is a grid-wide kernel over geoms / pairs /
//     bodies and the rows live in HBM;
//   * broadphase = dxSAPSpace::collide (ode/src/collision_sapspace.cpp:425-496): device radix sort
//     of the float-cast axis-0 minima, one thread per sorted position sweeps forward while
//     key[j] <= max0[i] (BoxPruning :521-567), pairs come out in the reference's (i, j) order by
//     count -> scan -> fill, so the pair SET and per-pair contacts equal the reference's;
//   * SOR_LCP (ode/src/quickstep.cpp:342-584) sweeps the rows in a seeded random order.  A single
//     200 k-body island has millions of rows in one dependency chain, so here the order is a GRAPH
//     COLOURING instead: contact pairs (all contacts between two bodies) are coloured so that no two
//     pairs of one colour share a body; colours are swept in ascending order, the pairs of one colour
//     in parallel (they touch disjoint bodies, hence disjoint fc[] entries), the rows of one pair in
//     sequence.  This is exactly a sequential Gauss-Seidel sweep in the order
//     (colour, pair, contact, row) -- which is what the host mirror executes -- but it is NOT the
//     reference's random order, so body state is compared with the reference within a stated
//     tolerance (tests/test_large_world.py), as BASELINE.json's north_star prescribes.
//
// Row storage: the 20-word row record of ob_step_kernel.cuh, but structure-of-arrays by 16-byte
// slot so that consecutive threads (consecutive pairs of one colour) load consecutive addresses:
//   rows[(q * OB_LW_SLOTS + s) * NC + cs], cs = contact slot.  Pairs are sorted by
//   (colour, contacts descending); the k-th contacts of one colour's pairs are contiguous:
//   cs = segbase[colour][k] + (pair - colstart[colour]).
#pragma once
#include "ob_types.h"
#include "ob_collide.h"
#include "ob_broad.h"
#include "ob_rows.h"
#include "ob_solver.h"

#define OB_LW_MAXC 8          // contacts per geom pair kept by the narrowphase (policy.max_contacts is clamped)
#define OB_LW_MAXCOL 64       // colours (one bit each in the per-body mask)
#define OB_LW_KEY_BIG 0xFFFFFFFEu     // sort key of geoms with an infinite axis-0 maximum (TmpInfGeomList)
#define OB_LW_KEY_OFF 0xFFFFFFFFu     // disabled geoms
#if defined(dSINGLE)
#define OB_LW_SLOTS 5         // float4 slots per row record
#define OB_LW_SLOTW 4
#else
#define OB_LW_SLOTS 10        // double2 slots
#define OB_LW_SLOTW 2
#endif
#define OB_LW_ROWW 20
#define OB_LW_HITBUF 16
#if !defined(__CUDACC__)
struct int4 { int x, y, z, w; };
#endif

struct __attribute__((aligned(16))) ObLwBox { real maxx, miny, maxy, minz, maxz; float minx; int body; uint32_t cat, col; int geom; int pad[2]; };
struct __attribute__((aligned(16))) ObLwPair { int b1, b2; int info; int src; };   // info = nc | rev<<8 | colour<<16 ; src = geom-pair index

struct ObLargeDev {
  int NG, NB, NP, NC;          // capacities: geoms, bodies, geom pairs, contact slots
  // geoms
  ObPose *pose;                // [NG]
  real *aabb;                  // [NG*6]
  uint32_t *gkey[2];           // [NG] radix sort ping-pong
  int *gidx[2];
  ObLwBox *sbox;               // [NG] sorted order (host mirror layout; the kernels use the split arrays below)
  float *sminx;                // [NG] sorted: float-cast axis-0 minimum (the sweep's stop test reads only this)
  real *smaxx;                 // [NG] axis-0 maximum
  real *syz;                   // [NG*4] miny maxy minz maxz
  int4 *smeta;                 // [NG] body, category bits, collide bits, geom index
  int *hits;                   // [NG*OB_LW_HITBUF] first hits of every sorted position, recorded by the counting pass
  int *scal;                   // device scalars, see LW_* below
  // pairs
  uint32_t *cnt;               // [2*NG+2] per sorted geom: sweep hits, then hits against the infinite list
  uint32_t *off;               // [2*NG+2]
  int *pairs;                  // == ObBatchDev::pairs [NP*2]
  ObContact *pc;               // [NP*maxc] contacts of pair p at p*maxc
  uint32_t *ncp;               // [NP+1] contacts per pair
  uint32_t *coff;              // [NP+1] exclusive scan: contact creation index
  uint32_t *cpflag;            // [NP+1] 1 if the pair enters the solver
  uint32_t *cpoff;             // [NP+1]
  ObLwPair *cp[2];             // [NP] contact pairs (creation order), then sorted by (colour, -nc)
  uint32_t *pkey[2];           // [NP]
  int *pidx[2];
  unsigned long long *claim;   // [NB] colouring: lowest-priority claimant of the round
  unsigned long long *used;    // [NB] colours taken at this body
  int *segtab;                 // [OB_LW_MAXCOL*(2+OB_LW_MAXC)] colstart, colcount, segbase[k]
  // solver
  real *rows;                  // [3*OB_LW_SLOTS*OB_LW_SLOTW*NC]
  real *lambda;                // [3*NC]
  real *fc;                    // [NB*8]
  real *invM;                  // [NB]
  int *hasrow;                 // [NB]
  uint32_t *tmp;               // scan / sort scratch
  size_t tmp_words;
};
// ---- one large world over the GPUs of one box ---------------------------
// Every rank (one process per GPU) holds the whole world.  What is divided, and what is not:
//   * FRONT END (default after dBatchSplitAttach): the pair sweep and the narrowphase are throughput-bound and
//     independent per sorted position, so rank r takes the sorted positions [ng r / N, ng (r + 1) / N): it counts
//     their hits, and -- after one exchange of the counts -- fills and collides exactly the pairs those positions
//     emit.  Counts, pairs and contacts are written into EVERY rank's arrays (plain stores through the NVLink peer
//     mappings) and a flag barrier (k_lw_xbarrier) follows, so afterwards all ranks hold the single-GPU arrays bit for
//     bit.  Two barriers per step.
//   * sort, colouring, row assembly, SOR and integration run on every rank redundantly: they are deterministic, and the
//     SOR phase is bound by the dependent row chain of one pair per colour (one wave of threads per colour at 200 k
//     bodies), not by bytes, so dividing its pairs does not shorten it.
//   * SOR split (OB_LW_SPLIT_SOR=1, the first design, kept as a measured negative result): the pairs of a colour are
//     dealt over the ranks; a rank writes the fc[] it produced into every rank's fc array and the ranks meet at a flag
//     barrier after each colour.
#define OB_LW_MAXRANKS 8
#define OB_LW_FLAG_WORDS 64      // per rank: [0..7] phase reached by rank r, [16] local release word, [17] timeout flag
#define OB_LW_FLAG_RELEASE 16
#define OB_LW_FLAG_TIMEOUT 17
struct ObLwSplit {
  int rank, nranks;
  unsigned base;                       // phase number before this launch (monotonic over the steps, equal on all ranks)
  unsigned timeout_ms;
  real *fc[OB_LW_MAXRANKS];            // every rank's fc array as mapped into THIS process (fc[rank] == ObLargeDev::fc)
  unsigned *flags[OB_LW_MAXRANKS];     // every rank's flag words (flags[rank] is local)
  // front-end split: every rank's pair-count, pair and contact arrays (the local ones are ObLargeDev's)
  uint32_t *cnt[OB_LW_MAXRANKS];
  int *pairs[OB_LW_MAXRANKS];
  uint32_t *ncp[OB_LW_MAXRANKS];
  uint32_t *cpflag[OB_LW_MAXRANKS];
  ObContact *pc[OB_LW_MAXRANKS];
};
// sorted positions [lo, hi) of rank r (multiples of 32 so that warps stay whole)
OB_HD void ob_lw_split_range(int ng, int rank, int nranks, int *lo, int *hi) {
  const int per = ((ng + nranks - 1) / nranks + 31) & ~31;
  *lo = rank * per < ng ? rank * per : ng;
  *hi = (rank + 1) * per < ng ? (rank + 1) * per : ng;
}
// which rank sweeps warp tile `tile` (32 consecutive pairs of a colour): tiles are dealt round-robin, like the
// single-GPU kernel deals them over its CTAs, so every rank gets the same mix of heavy and light pairs
OB_HD int ob_lw_split_owner(int pair_in_colour, int nranks) { return (pair_in_colour >> 5) % nranks; }

enum { LW_NFIN = 0, LW_NBIG, LW_NP, LW_NCONTACTS, LW_NCP, LW_UNCOLOURED, LW_NCOL, LW_ERR, LW_NSOLVED, LW_BARRIER, LW_LEFT0 /* 16 per-round counters */,
       LW_SEG0 = 26 /* front-end split: start / length of this rank's two pair ranges */, LW_WORDS = 32 };

// ---- broadphase --------------------------------------------------------------------------------
// sort key of one geom (collision_sapspace.cpp:441-452, :531-535): float-cast axis-0 minimum
OB_HD uint32_t ob_lw_geomkey(const real *aabb, int enabled, int ax0, float *minf) {
  *minf = (float)aabb[ax0];
  if (!enabled) return OB_LW_KEY_OFF;
  if (aabb[ax0 + 1] == OB_INF) return OB_LW_KEY_BIG;
  const uint32_t k = ob_sap_keyorder(*minf);
  return k < OB_LW_KEY_BIG ? k : OB_LW_KEY_BIG - 1;
}
// BoxPruning inner test for sorted positions i < j (:545-560) + collideGeomsNoAABBs filter (:234-258)
OB_HD bool ob_lw_sweep_hit(const ObLwBox &K, const ObLwBox &J) {
  if (!(K.maxy >= J.miny && J.maxy >= K.miny)) return false;
  if (!(K.maxz >= J.minz && J.maxz >= K.minz)) return false;
  return ob_pair_filter_noaabb(K.body, J.body, K.cat, K.col, J.cat, J.col);
}

// ---- colouring ---------------------------------------------------------------------------------
OB_HD uint32_t ob_lw_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// priority of a pair in a round (lower wins).  The top byte counts DOWN with the round, so a claim left
// in the table by an earlier round always loses against any claim of the current round: the claim table
// needs no clearing between rounds (it is cleared once per step; at most 255 rounds).
OB_HD unsigned long long ob_lw_prio(uint32_t pair, uint32_t round) {
  return ((unsigned long long)(254u - (round & 255u)) << 56) | ((unsigned long long)(ob_lw_hash(pair * 0x9E3779B9u + round) >> 8) << 32) | pair;
}
OB_HD int ob_lw_first_free(unsigned long long used) {   // lowest clear bit, OB_LW_MAXCOL if none
  for (int c = 0; c < OB_LW_MAXCOL; c++) if (!((used >> c) & 1ull)) return c;
  return OB_LW_MAXCOL;
}

// ---- rows --------------------------------------------------------------------------------------
// contact -> m finalised row records (20 words each, the format of ob_step_kernel.cuh):
// J1l[3] J1a[3] J2a[3] | iMJ1a[3] iMJ2a[3] | Ad b Ad*cfm bound | meta (findex offset<<16 | bound mode<<24)
// Returns false when a row's bounds are not one of the three encodable shapes.
OB_HD bool ob_lw_contact_rows(const ObContact &c, int rev, const ObSurface &surf, int m, const ObWorld &W,
                              const real *p1, const real *l1, const real *a1, const real *t1a, const real *iw1, real k1,
                              int has_b2, const real *p2, const real *l2, const real *a2, const real *t1b, const real *iw2, real k2,
                              real stepsize1, real (*out)[OB_LW_ROWW], unsigned *meta) {
  ObRowOut3 r;
  ob_rows_defaults(r, m, W.cfm);
  const real fdir1[3] = {0, 0, 0};
  ob_contact_info2(r, m, surf, c.pos, c.normal, c.depth, fdir1, rev, p1, l1, a1, has_b2, p2, l2, a2, stepsize1, W.erp,
                   W.min_depth, W.max_vel);
  bool ok = true;
  for (int q = 0; q < m; q++) {
    real iMJ[12], b_out, adcfm, Ad;
    ob_row_finalize2(r.J[q], r.c[q], r.cfm[q], has_b2 ? 0 : -1, t1a, t1b, k1, iw1, k2, iw2, stepsize1, W.sor_w, iMJ, &b_out,
                     &adcfm, &Ad);
    real *rw = out[q];
    for (int e = 0; e < 6; e++) rw[e] = r.J[q][e];
    for (int e = 0; e < 3; e++) { rw[6 + e] = r.J[q][9 + e]; rw[9 + e] = iMJ[3 + e]; rw[12 + e] = iMJ[9 + e]; }
    rw[15] = Ad; rw[16] = b_out; rw[17] = adcfm;
    unsigned bmode = 0;
    const real lo = r.lo[q], hi = r.hi[q];
    if (lo == -hi) { rw[18] = hi; bmode = 0; }
    else if (lo == 0) { rw[18] = hi; bmode = 1; }
    else if (hi == 0) { rw[18] = lo; bmode = 2; }
    else { rw[18] = hi; ok = false; }
    rw[19] = 0;
    const unsigned fio = (unsigned)(r.findex[q] >= 0 ? q - r.findex[q] : 0);
    meta[q] = (fio << 16) | (bmode << 24);
  }
  return ok;
}

// one row update (quickstep.cpp:490-581) from a row record; lam_f = lambda of the row's findex row
OB_HD real ob_lw_row_update(const real *v, unsigned meta, real k1, real k2, int has_b2, real lam_f, real lam_old, real *f1, real *f2) {
  const int fio = (meta >> 16) & 255, bmode = meta >> 24;
  const real Ad = v[15], bv = v[18];
  const real lo = bmode == 0 ? -bv : (bmode == 1 ? (real)0 : bv);
  const real hi = bmode == 2 ? (real)0 : bv;
  real J[12], iMJ[12];
  for (int e = 0; e < 3; e++) {
    iMJ[e] = k1 * v[e];
    iMJ[3 + e] = v[9 + e];
    J[e] = v[e] * Ad;
    J[3 + e] = v[3 + e] * Ad;
  }
  if (has_b2) {
    for (int e = 0; e < 3; e++) {
      const real j2l = -v[e];
      iMJ[6 + e] = k2 * j2l;
      iMJ[9 + e] = v[12 + e];
      J[6 + e] = j2l * Ad;
      J[9 + e] = v[6 + e] * Ad;
    }
  }
  return ob_sor_row(J, iMJ, v[16], v[17], lo, hi, fio ? 0 : -1, lam_f, lam_old, f1, has_b2 ? f2 : (real *)0);
}

OB_HD size_t ob_lw_row_index(int q, int s, size_t NC, size_t cs) { return ((size_t)(q * OB_LW_SLOTS + s) * NC + cs) * OB_LW_SLOTW; }

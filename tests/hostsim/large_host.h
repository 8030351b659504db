// large_host.h — TEST-ONLY sequential mirror of the large-world pipeline (ode-0.12_b200/csrc/ob_large.h,
// ob_large_kernels.cuh).  Same per-element functions, same orders: pairs in (sorted position i, j) order,
// colouring by synchronous rounds with the same hashed priorities, pairs sorted by (colour, contacts
// descending), and the SOR sweep executed as a plain sequential Gauss-Seidel loop in the order
// (iteration, colour, pair, contact, row).  The CUDA path must reproduce this bit for bit (tests, -m gpu);
// this mirror itself is compared with the unmodified reference within the stated tolerance on the CPU.
#pragma once
#include <algorithm>
#include <vector>
#include "../../ode-0.12_b200/csrc/ob_large.h"

struct LargeHostStats { int np, ncontacts, ncp, ncol, rounds, nsolved; };

// Mirror of the split SOR phase (ObLwSplit, ob_large.h): each rank is a batch of its own, stepped by its own
// host thread; the ranks write their fc results into every rank's array and meet at a flag barrier after each
// colour -- the protocol of k_lw_sor_split / lw_split_barrier with std::atomic flag words.
#include <atomic>
#include <chrono>
struct LargeHostSplit {
  int rank, nranks;
  unsigned base, timeout_ms;
  real *fc[OB_LW_MAXRANKS];                       // [nb*6] per rank
  std::atomic<unsigned> *flags[OB_LW_MAXRANKS];   // [OB_LW_FLAG_WORDS] per rank
};
static bool large_host_barrier(LargeHostSplit &S, unsigned phase) {
  std::atomic<unsigned> *mine = S.flags[S.rank];
  for (int r = 0; r < S.nranks; r++) if (r != S.rank) S.flags[r][S.rank].store(phase, std::memory_order_release);
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < S.nranks; r++) {
    if (r == S.rank) continue;
    while ((int)(mine[r].load(std::memory_order_acquire) - phase) < 0 && !mine[OB_LW_FLAG_TIMEOUT].load()) {
      if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > (long long)S.timeout_ms) mine[OB_LW_FLAG_TIMEOUT].store(1);
    }
  }
  return !mine[OB_LW_FLAG_TIMEOUT].load();
}

static int large_step_host(ObBatchDev &d, real h, int taps, LargeHostStats *stats, LargeHostSplit *split = 0) {
  ObWorld &W = d.world[0];
  const int ng = W.ng, nb = W.nb;
  if (W.space_type != OB_SPACE_SAP) return -1;
  const ObPolicy pol = d.policy[0];
  const int maxc = pol.max_contacts > OB_LW_MAXC ? OB_LW_MAXC : (pol.max_contacts < 1 ? 1 : pol.max_contacts);
  ObSurface surf = pol.surface;
  const int m = ob_contact_info1(surf);
  int ax0, ax1, ax2;
  ob_sap_axes(W.sap_axes, &ax0, &ax1, &ax2);
  // (1) geoms
  std::vector<ObPose> pose(ng);
  std::vector<real> aabb((size_t)ng * 6);
  std::vector<std::pair<uint32_t, int> > ks(ng);
  for (int g = 0; g < ng; g++) {
    geom_pose(d, 0, g, &pose[g]);
    ob_aabb(pose[g], &aabb[(size_t)g * 6], d.meshes);
    const ObGeom &G = d.geom[g];
    const int en = (G.flags & OB_GEOM_ENABLED) && !(G.flags & OB_GEOM_ZERO_SIZED);
    float minf;
    ks[g] = std::make_pair(ob_lw_geomkey(&aabb[(size_t)g * 6], en, ax0, &minf), g);
  }
  std::stable_sort(ks.begin(), ks.end(), [](const std::pair<uint32_t, int> &a, const std::pair<uint32_t, int> &b) { return a.first < b.first; });
  int nfin = 0, bigend = 0;
  for (int i = 0; i < ng; i++) { if (ks[i].first < OB_LW_KEY_BIG) nfin = i + 1; if (ks[i].first == OB_LW_KEY_BIG) bigend = i + 1; }
  if (bigend < nfin) bigend = nfin;
  std::vector<ObLwBox> sbox(ng);
  for (int i = 0; i < ng; i++) {
    const int g = ks[i].second;
    const real *ab = &aabb[(size_t)g * 6];
    const ObGeom &G = d.geom[g];
    ObLwBox b;
    b.maxx = ab[ax0 + 1]; b.miny = ab[ax1]; b.maxy = ab[ax1 + 1]; b.minz = ab[ax2]; b.maxz = ab[ax2 + 1];
    b.minx = (float)ab[ax0]; b.body = G.body; b.cat = G.cat; b.col = G.col; b.geom = g; b.pad[0] = b.pad[1] = 0;
    sbox[i] = b;
  }
  // (2) pairs
  std::vector<int> pairs;
  for (int i = 0; i < nfin; i++) {
    const ObLwBox &K = sbox[i];
    for (int j = i + 1; j < nfin; j++) {
      const ObLwBox &J = sbox[j];
      if (!((real)J.minx <= K.maxx)) break;
      if (ob_lw_sweep_hit(K, J)) { pairs.push_back(K.geom); pairs.push_back(J.geom); }
    }
  }
  for (int i = 0; i < nfin; i++) {
    const ObLwBox &K = sbox[i];
    for (int a = nfin; a < bigend; a++) {
      const int ga = ks[a].second;
      const ObGeom &A = d.geom[ga];
      if (ob_pair_filter_noaabb(A.body, K.body, A.cat, A.col, K.cat, K.col)) { pairs.push_back(ga); pairs.push_back(K.geom); }
    }
  }
  for (int a = nfin; a < bigend; a++)
    for (int b = a + 1; b < bigend; b++) {
      const int ga = ks[a].second, gb = ks[b].second;
      const ObGeom &A = d.geom[ga], &B = d.geom[gb];
      if (ob_pair_filter_noaabb(A.body, B.body, A.cat, A.col, B.cat, B.col)) { pairs.push_back(ga); pairs.push_back(gb); }
    }
  int np = (int)pairs.size() / 2;
  if (np > d.NP) { np = d.NP; W.status |= OB_ERR_PAIR_OVERFLOW; }
  for (int i = 0; i < 2 * np; i++) d.pairs[i] = pairs[i];
  d.npairs[0] = np;
  // (3) narrowphase + contact pairs
  std::vector<ObContact> pc((size_t)np * maxc);
  std::vector<int> ncp(np);
  std::vector<ObLwPair> cp;
  int ncontacts = 0;
  for (int p = 0; p < np; p++) {
    const int o1 = pairs[2 * p], o2 = pairs[2 * p + 1];
    ObCg cg[OB_LW_MAXC];
    int swapped, bverr = 0;
    const int n = d.nmesh ? ob_collide_pair_xf_t<true, OB_LW_MAXC>(&pose[o1], &pose[o2], d.any_xf, maxc, cg, &swapped, d.meshes, &bverr)
                          : ob_collide_pair_xf_t<false, OB_LW_MAXC>(&pose[o1], &pose[o2], d.any_xf, maxc, cg, &swapped, d.meshes, &bverr);
    if (bverr) W.status |= OB_ERR_BVH_STACK;
    for (int k = 0; k < n; k++) {
      ObContact c;
      for (int e = 0; e < 3; e++) { c.pos[e] = cg[k].pos[e]; c.normal[e] = cg[k].normal[e]; }
      c.depth = cg[k].depth; c.g1 = o1; c.g2 = o2; c.side1 = cg[k].side1; c.side2 = cg[k].side2; c.policy = 0;
      pc[(size_t)p * maxc + k] = c;
      if (taps && ncontacts + k < d.NC) d.contacts[ncontacts + k] = c;
    }
    ncp[p] = n;
    ncontacts += n;
    int b1 = d.geom[o1].body, b2 = d.geom[o2].body, rev = 0;
    if (n > 0 && (b1 >= 0 || b2 >= 0)) {
      if (b1 < 0) { b1 = b2; b2 = -1; rev = 1; }
      ObLwPair P;
      P.b1 = b1; P.b2 = b2; P.info = n | (rev << 8) | (255 << 16); P.src = p;
      cp.push_back(P);
    }
  }
  d.ncontacts[0] = ncontacts < d.NC ? ncontacts : d.NC;
  const int ncpairs = (int)cp.size();
  // (4) colouring, synchronous rounds
  std::vector<unsigned long long> used(nb, 0ull), claim(nb);
  int left = ncpairs, rounds = 0;
  bool colerr = false;
  while (left > 0) {
    std::fill(claim.begin(), claim.end(), ~0ull);
    for (int p = 0; p < ncpairs; p++) {
      if ((cp[p].info >> 16) != 255) continue;
      const unsigned long long pr = ob_lw_prio((uint32_t)p, (uint32_t)rounds);
      if (pr < claim[cp[p].b1]) claim[cp[p].b1] = pr;
      if (cp[p].b2 >= 0 && pr < claim[cp[p].b2]) claim[cp[p].b2] = pr;
    }
    left = 0;
    for (int p = 0; p < ncpairs; p++) {
      ObLwPair &P = cp[p];
      if ((P.info >> 16) != 255) continue;
      const unsigned long long pr = ob_lw_prio((uint32_t)p, (uint32_t)rounds);
      if (claim[P.b1] == pr && (P.b2 < 0 || claim[P.b2] == pr)) {
        unsigned long long u = used[P.b1];
        if (P.b2 >= 0) u |= used[P.b2];
        int c = ob_lw_first_free(u);
        if (c >= OB_LW_MAXCOL) { c = OB_LW_MAXCOL - 1; colerr = true; }
        used[P.b1] |= 1ull << c;
        if (P.b2 >= 0) used[P.b2] |= 1ull << c;
        P.info = (P.info & 0xffff) | (c << 16);
      } else left++;
    }
    rounds++;
    if (rounds > 240) return -1;
  }
  if (colerr) return -2;
  // (5) order pairs by (colour, contacts descending), stable
  std::vector<std::pair<uint32_t, int> > pk(ncpairs);
  for (int p = 0; p < ncpairs; p++) pk[p] = std::make_pair((uint32_t)(((cp[p].info >> 16) & 255) * 8 + (8 - (cp[p].info & 255))), p);
  std::stable_sort(pk.begin(), pk.end(), [](const std::pair<uint32_t, int> &a, const std::pair<uint32_t, int> &b) { return a.first < b.first; });
  std::vector<ObLwPair> scp(ncpairs);
  for (int p = 0; p < ncpairs; p++) scp[p] = cp[pk[p].second];
  int ncol = 0;
  for (int p = 0; p < ncpairs; p++) ncol = std::max(ncol, ((scp[p].info >> 16) & 255) + 1);
  // (6) bodies
  const real stepsize1 = ob_recip(h);
  std::vector<real> fc_own(split ? (size_t)0 : (size_t)nb * 6, (real)0);
  real *fc = split ? split->fc[split->rank] : fc_own.data();
  if (split) for (size_t k = 0; k < (size_t)nb * 6; k++) fc[k] = 0;
  std::vector<int> hasrow(nb, 0);
  for (int b = 0; b < nb; b++) {
    ObBodyDyn &B = d.bdyn[b];
    const ObBodyConst &C = d.bconst[b];
    real iw[12], facc[3], tacc[3], t1[6];
    for (int k = 0; k < 3; k++) { facc[k] = B.facc[k]; tacc[k] = B.tacc[k]; }
    ob_body_preamble(B.R, C.I, C.invI, B.avel, B.flags, C.mass, W.gravity, iw, facc, tacc);
    for (int k = 0; k < 3; k++) { B.facc[k] = facc[k]; B.tacc[k] = tacc[k]; }
    for (int k = 0; k < 12; k++) d.invIw[(size_t)12 * b + k] = iw[k];
    ob_body_tmp1(facc, tacc, B.lvel, B.avel, C.invMass, iw, stepsize1, t1);
    for (int k = 0; k < 6; k++) d.tmp1[(size_t)8 * b + k] = t1[k];
  }
  // (7) rows, in sweep order: [pair][contact][row]
  struct Row { real v[OB_LW_ROWW]; unsigned meta; };
  std::vector<Row> rows;
  std::vector<size_t> rstart(ncpairs + 1, 0);
  int nsolved = 0;
  for (int p = 0; p < ncpairs; p++) {
    const ObLwPair &P = scp[p];
    const int nc = P.info & 255, rev = (P.info >> 8) & 1, b1 = P.b1, b2 = P.b2;
    rstart[p] = rows.size();
    real z3[3] = {0, 0, 0}, z6[6] = {0, 0, 0, 0, 0, 0}, z12[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    hasrow[b1] = 1;
    if (b2 >= 0) hasrow[b2] = 1;
    for (int k = 0; k < nc; k++) {
      real rw[3][OB_LW_ROWW];
      unsigned meta[3];
      const ObContact &c = pc[(size_t)P.src * maxc + k];
      const bool ok = ob_lw_contact_rows(c, rev, surf, m, W, d.bdyn[b1].pos, d.bdyn[b1].lvel, d.bdyn[b1].avel, &d.tmp1[(size_t)8 * b1],
                                         &d.invIw[(size_t)12 * b1], d.bconst[b1].invMass, b2 >= 0, b2 >= 0 ? d.bdyn[b2].pos : z3,
                                         b2 >= 0 ? d.bdyn[b2].lvel : z3, b2 >= 0 ? d.bdyn[b2].avel : z3, b2 >= 0 ? &d.tmp1[(size_t)8 * b2] : z6,
                                         b2 >= 0 ? &d.invIw[(size_t)12 * b2] : z12, b2 >= 0 ? d.bconst[b2].invMass : (real)0, stepsize1, rw, meta);
      if (!ok) W.status |= OB_ERR_ROW_OVERFLOW;
      for (int q = 0; q < m; q++) { Row r; for (int e = 0; e < OB_LW_ROWW; e++) r.v[e] = rw[q][e]; r.meta = meta[q]; rows.push_back(r); }
      nsolved++;
    }
  }
  rstart[ncpairs] = rows.size();
  std::vector<real> lam(rows.size(), (real)0);
  // (8') split over ranks: the pairs of a colour are dealt over the ranks by warp tile (ob_lw_split_owner), results
  // go into every rank's fc, barrier after every colour (and one before the first store: the peers' fc must be cleared)
  if (split && ncpairs > 0 && W.iters > 0) {
    LargeHostSplit &S = *split;
    unsigned phase = S.base;
    bool ok = large_host_barrier(S, ++phase);
    std::vector<int> colstart(ncol + 1, ncpairs);
    for (int p = ncpairs - 1; p >= 0; p--) colstart[(scp[p].info >> 16) & 255] = p;
    for (int c = ncol - 1; c >= 0; c--) if (colstart[c] > colstart[c + 1]) colstart[c] = colstart[c + 1];
    for (int it = 0; it < W.iters; it++)
      for (int c = 0; c < ncol; c++) {
        for (int p = colstart[c]; p < colstart[c + 1]; p++) {
          if (ob_lw_split_owner(p - colstart[c], S.nranks) != S.rank) continue;
          const ObLwPair &P = scp[p];
          const int nc = P.info & 255, b1 = P.b1, b2 = P.b2;
          real f1[6], f2[6] = {0, 0, 0, 0, 0, 0};
          for (int e = 0; e < 6; e++) { f1[e] = fc[(size_t)6 * b1 + e]; if (b2 >= 0) f2[e] = fc[(size_t)6 * b2 + e]; }
          const real k1 = d.bconst[b1].invMass, k2 = b2 >= 0 ? d.bconst[b2].invMass : (real)0;
          for (int k = 0; k < nc; k++)
            for (int q = 0; q < m; q++) {
              const size_t ri = rstart[p] + (size_t)k * m + q;
              const int fio = (rows[ri].meta >> 16) & 255;
              const real lam_f = fio ? lam[ri - fio] : (real)0;
              lam[ri] = ob_lw_row_update(rows[ri].v, rows[ri].meta, k1, k2, b2 >= 0, lam_f, lam[ri], f1, b2 >= 0 ? f2 : (real *)0);
            }
          for (int r = 0; r < S.nranks; r++)
            for (int e = 0; e < 6; e++) { S.fc[r][(size_t)6 * b1 + e] = f1[e]; if (b2 >= 0) S.fc[r][(size_t)6 * b2 + e] = f2[e]; }
        }
        ok = large_host_barrier(S, ++phase) && ok;
      }
    S.base = phase;
    if (!ok) return -7;
  } else
  // (8) SOR: sequential Gauss-Seidel in (colour, pair, contact, row) order
  for (int it = 0; it < W.iters; it++)
    for (int p = 0; p < ncpairs; p++) {
      const ObLwPair &P = scp[p];
      const int nc = P.info & 255, b1 = P.b1, b2 = P.b2;
      real *f1 = &fc[(size_t)6 * b1], *f2 = b2 >= 0 ? &fc[(size_t)6 * b2] : (real *)0;
      const real k1 = d.bconst[b1].invMass, k2 = b2 >= 0 ? d.bconst[b2].invMass : (real)0;
      for (int k = 0; k < nc; k++)
        for (int q = 0; q < m; q++) {
          const size_t ri = rstart[p] + (size_t)k * m + q;
          const int fio = (rows[ri].meta >> 16) & 255;
          const real lam_f = fio ? lam[ri - fio] : (real)0;
          lam[ri] = ob_lw_row_update(rows[ri].v, rows[ri].meta, k1, k2, b2 >= 0, lam_f, lam[ri], f1, f2);
        }
    }
  // parity tap: joint feedback per contact joint in creation order (as k_lw_feedback)
  if ((taps & 1) && d.fback) {
    memset(d.fback, 0, sizeof(real) * 12 * (size_t)d.NC);
    std::vector<size_t> coff(np + 1, 0);
    for (int p = 0; p < np; p++) coff[p + 1] = coff[p] + ncp[p];
    for (int p = 0; p < ncpairs; p++) {
      const ObLwPair &P = scp[p];
      const int nc = P.info & 255;
      for (int k = 0; k < nc; k++) {
        const size_t ci = coff[P.src] + k;
        if (ci >= (size_t)d.NC) continue;
        real acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int q = 0; q < m; q++) {
          const size_t ri = rstart[p] + (size_t)k * m + q;
          const real *rw = rows[ri].v;
          for (int e = 0; e < 6; e++) acc[e] += rw[e] * lam[ri];
          for (int e = 0; e < 3; e++) { acc[6 + e] += (-rw[e]) * lam[ri]; acc[9 + e] += rw[6 + e] * lam[ri]; }
        }
        if (P.b2 < 0) for (int e = 6; e < 12; e++) acc[e] = 0;
        for (int e = 0; e < 12; e++) d.fback[ci * 12 + e] = acc[e];
      }
    }
  }
  // (9) integrate
  for (int b = 0; b < nb; b++) {
    ObBodyDyn &B = d.bdyn[b];
    const ObBodyConst &C = d.bconst[b];
    real facc[3], tacc[3], fra[3] = {C.finite_rot_axis[0], C.finite_rot_axis[1], C.finite_rot_axis[2]};
    for (int k = 0; k < 3; k++) { facc[k] = B.facc[k]; tacc[k] = B.tacc[k]; }
    ob_body_velocity_update(B.lvel, B.avel, hasrow[b] ? &fc[(size_t)6 * b] : (real *)0, facc, tacc, C.invMass, &d.invIw[(size_t)12 * b], h);
    ob_step_body(B.pos, B.q, B.R, B.lvel, B.avel, B.flags, h, C.max_angular_speed, fra, C.damp_lin_scale, C.damp_ang_scale,
                 C.damp_lin_thr, C.damp_ang_thr);
    for (int k = 0; k < 4; k++) { B.facc[k] = 0; B.tacc[k] = 0; }
  }
  d.nrows[0] = 0;
  d.counters->steps += 1;
  d.counters->body_steps += nb;
  d.counters->pairs += np;
  d.counters->contacts += nsolved;
  if (stats) { stats->np = np; stats->ncontacts = ncontacts; stats->ncp = ncpairs; stats->ncol = ncol; stats->rounds = rounds; stats->nsolved = nsolved; }
  return 0;
}

// backend_host.cpp — TEST-ONLY execution backend (see ob_backend.h).
//
// Runs the per-element device functions of ode-0.12_b200/csrc/ob_*.h in plain
// sequential loops on the CPU so that their arithmetic and the ordering rules
// can be compared with the unmodified reference in a container without a GPU.
// It is linked only into tests/hostsim/libode_b200_hostsim_*.so and is never part
// of, loaded by, or a fallback for libode_b200_*.so.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <vector>
#include "../../ode-0.12_b200/csrc/ob_backend.h"
#include "../../ode-0.12_b200/csrc/ob_large.h"
#include "../../ode-0.12_b200/csrc/ob_broad.h"
#include "../../ode-0.12_b200/csrc/ob_collide.h"
#include "../../ode-0.12_b200/csrc/ob_rows.h"
#include "../../ode-0.12_b200/csrc/ob_solver.h"

struct LargeHostSplit;
struct ObBackend {
  ObBatchDev d;
  std::vector<void *> allocs;
  LargeHostSplit *split;      // large world split over ranks (threads here), see large_host.h
  void *split_fc, *split_flags;
};

template <class T> static T *halloc(ObBackend *b, size_t n) {
  void *p = calloc(n ? n : 1, sizeof(T));
  b->allocs.push_back(p);
  return (T *)p;
}

ObBackend *obk_create(const ObBatchDev &caps, int, char *, size_t) {
  ObBackend *b = new ObBackend;
  b->d = caps; b->split = 0; b->split_fc = 0; b->split_flags = 0;
  ObBatchDev &d = b->d;
  size_t W = d.W;
  d.world = halloc<ObWorld>(b, W);
  d.bdyn = halloc<ObBodyDyn>(b, W * d.NB);
  d.bconst = halloc<ObBodyConst>(b, W * d.NB);
  d.geom = halloc<ObGeom>(b, W * d.NG);
  d.glist = halloc<int>(b, W * d.NG);
  d.sapstate = halloc<int>(b, W * (d.NG + 3));
  d.policy = halloc<ObPolicy>(b, d.npolicy);
  d.meshes = halloc<ObMeshDev>(b, d.nmesh ? d.nmesh : 1);
  d.joint = halloc<ObJoint>(b, W * (d.NJ ? d.NJ : 1));
  d.njoints = halloc<int>(b, W);
  d.padjstart = halloc<unsigned short>(b, W * (d.NB + 1));
  d.padj = halloc<unsigned short>(b, W * 2 * (d.NJ ? d.NJ : 1));
  d.npairs = halloc<int>(b, W);
  d.pairs = halloc<int>(b, W * d.NP * 2);
  d.ncontacts = halloc<int>(b, W);
  d.contacts = halloc<ObContact>(b, W * d.NC);
  d.rowJ = halloc<real>(b, W * d.NR * 12);
  d.rowiMJ = halloc<real>(b, W * d.NR * 12);
  d.rowJc = halloc<real>(b, W * d.NR * 12);
  d.rowS = halloc<real>(b, W * d.NR * 4);
  d.rowI = halloc<int>(b, W * d.NR * 4);
  d.lambda = halloc<real>(b, W * d.NR);
  d.nrows = halloc<int>(b, W);
  d.fback = halloc<real>(b, W * (d.NC + d.NJ) * 12);
  d.ibody = halloc<unsigned char>(b, W * d.NB);
  d.stepinfo = halloc<int>(b, W * SI_WORDS);
  d.csurf = d.dropin ? halloc<ObSurface>(b, W * d.NC) : 0;
  d.cfdir1 = d.dropin ? halloc<real>(b, W * d.NC * 4) : 0;
  d.counters = halloc<ObCounters>(b, 1);
  d.adisbuf = 0; d.adisctl = 0; d.rowmeta = 0;
  if (d.NADIS > 0) { d.adisbuf = halloc<real>(b, (size_t)d.W * d.NB * d.NADIS * 6); d.adisctl = halloc<int>(b, (size_t)d.W * d.NB * 2); }
  if (d.large) { d.invIw = halloc<real>(b, (size_t)d.NB * 12); d.tmp1 = halloc<real>(b, (size_t)d.NB * 8); }
  return b;
}
void obk_destroy(ObBackend *b) { for (size_t i = 0; i < b->allocs.size(); i++) free(b->allocs[i]); delete b; }
ObBatchDev *obk_arrays(ObBackend *b) { return &b->d; }
int obk_h2d(ObBackend *, void *dst, const void *src, size_t n) { memcpy(dst, src, n); return 0; }
int obk_d2h(ObBackend *, void *dst, const void *src, size_t n) { memcpy(dst, src, n); return 0; }
int obk_memset(ObBackend *, void *dst, int v, size_t n) { memset(dst, v, n); return 0; }
int obk_sync(ObBackend *) { return 0; }
void *obk_stream(ObBackend *) { return 0; }
long long obk_launch_count(void) { return 0; }
int obk_timer_start(ObBackend *) { return 0; }
int obk_timer_stop(ObBackend *, float *ms) { *ms = 0; return 0; }
void obk_set_kernel_timing(ObBackend *, int) {}
void obk_get_kernel_times(ObBackend *, double *ms, long long *l) { for (int k = 0; k < OBK_NKERNELS; k++) { ms[k] = 0; l[k] = 0; } }
const char *obk_kernel_name(int) { return "host"; }
static int g_lw_stat[8];
int obk_large_stats(ObBackend *b, int *ints8, double *ms8) {
  if (!b->d.large) return -1;
  for (int k = 0; k < 8; k++) { ints8[k] = g_lw_stat[k]; ms8[k] = 0; }
  return 0;
}

int obk_get_state(ObBackend *b, real *pos3, real *quat4, real *lvel3, real *avel3) {
  ObBatchDev &d = b->d;
  for (int w = 0; w < d.W; w++) {
    int nb = d.world[w].nb;
    for (int c = 0; c < d.NB; c++) {
      size_t o = (size_t)w * d.NB + c;
      if (c >= nb) continue;
      const ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
      for (int k = 0; k < 3; k++) { if (pos3) pos3[o * 3 + k] = s.pos[k]; if (lvel3) lvel3[o * 3 + k] = s.lvel[k]; if (avel3) avel3[o * 3 + k] = s.avel[k]; }
      for (int k = 0; k < 4; k++) if (quat4) quat4[o * 4 + k] = s.q[k];
    }
  }
  return 0;
}
int obk_set_state(ObBackend *b, const real *pos3, const real *quat4, const real *lvel3, const real *avel3) {
  ObBatchDev &d = b->d;
  for (int w = 0; w < d.W; w++) {
    int nb = d.world[w].nb;
    for (int c = 0; c < nb; c++) {
      size_t o = (size_t)w * d.NB + c;
      ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
      for (int k = 0; k < 3; k++) { if (pos3) s.pos[k] = pos3[o * 3 + k]; if (lvel3) s.lvel[k] = lvel3[o * 3 + k]; if (avel3) s.avel[k] = avel3[o * 3 + k]; }
      if (quat4) { for (int k = 0; k < 4; k++) s.q[k] = quat4[o * 4 + k]; ob_RfromQ(s.R, s.q); }
    }
  }
  return 0;
}
void *obk_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void obk_host_free(void *p) { free(p); }
int obk_add_forces(ObBackend *b, const real *f3, const real *t3) {
  ObBatchDev &d = b->d;
  for (int w = 0; w < d.W; w++) {
    int nb = d.world[w].nb;
    for (int c = 0; c < nb; c++) {
      size_t o = (size_t)w * d.NB + c;
      ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
      for (int k = 0; k < 3; k++) { if (f3) s.facc[k] += f3[o * 3 + k]; if (t3) s.tacc[k] += t3[o * 3 + k]; }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
static void geom_pose(const ObBatchDev &d, int w, int gi, ObPose *o) {
  const ObGeom &g = d.geom[(size_t)w * d.NG + gi];
  o->type = g.type;
  o->mesh = g.mesh;
  for (int k = 0; k < 4; k++) o->p[k] = g.p[k];
  if (g.body >= 0) {
    const ObBodyDyn &b = d.bdyn[(size_t)w * d.NB + g.body];
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

#include "large_host.h"

struct PairRec { ObPairKey key; int o1, o2; };
static bool pair_less(const PairRec &a, const PairRec &b) { return ob_key_less(a.key, b.key); }

static void collide_world(ObBatchDev &d, int w) {
  ObWorld &W = d.world[w];
  int ng = W.ng;
  int *glist = d.glist + (size_t)w * d.NG;
  const int stype = W.space_type;
  int ax0 = 0, ax1 = 2, ax2 = 4;
  if (stype == OB_SPACE_SAP) {
    ob_sap_axes(W.sap_axes, &ax0, &ax1, &ax2);
    // cleanGeoms: GeomList += DirtyList
    std::vector<int> cleaned;
    for (int i = W.sap_ndirty; i < ng; i++) cleaned.push_back(glist[i]);
    for (int i = 0; i < W.sap_ndirty; i++) cleaned.push_back(glist[i]);
    for (int i = 0; i < ng; i++) glist[i] = cleaned[i];
    W.sap_ndirty = 0;
  }
  std::vector<ObPose> pose(ng);
  std::vector<real> aabb(6 * ng);
  std::vector<ObCellBox> cb(ng);
  std::vector<int> hr(ng), br(ng), en(ng);
  int nh = 0, nbig = 0;
  for (int i = 0; i < ng; i++) {   // i = walk index
    int gi = glist[i];
    const ObGeom &g = d.geom[(size_t)w * d.NG + gi];
    en[i] = (g.flags & OB_GEOM_ENABLED) && !(g.flags & OB_GEOM_ZERO_SIZED);
    geom_pose(d, w, gi, &pose[i]);
    ob_aabb(pose[i], &aabb[6 * i], d.meshes);
    cb[i].level = 0;
    if (stype == OB_SPACE_HASH) ob_hash_cellbox(&aabb[6 * i], W.hash_minlevel, W.hash_maxlevel, &cb[i]);
    else if (stype == OB_SPACE_SAP && aabb[6 * i + ax0 + 1] == OB_INF) cb[i].level = OB_LEVEL_BIG;
    hr[i] = nh; br[i] = nbig;
    if (en[i]) { if (cb[i].level == OB_LEVEL_BIG) nbig++; else nh++; }
  }
  // SAP: sorted position of every finite geom (RadixSort's output order, ob_broad.h)
  std::vector<float> sapkey(nh + 1);
  std::vector<int> sappos(nh + 1), sapinit(nh + 1);
  if (stype == OB_SPACE_SAP && nh > 0) {
    for (int i = 0; i < ng; i++) if (en[i] && cb[i].level != OB_LEVEL_BIG) sapkey[hr[i]] = (float)aabb[6 * i + ax0];
    sapkey[nh] = 3.402823466e+38f;
    int *st = d.sapstate + (size_t)w * (d.NG + 3);
    const int nbk = nh + 1;
    const bool valid = st[0] != 0 && st[1] == nbk;
    for (int p = 0; p < nbk; p++) { if (valid) sapinit[st[2 + p]] = p; else sapinit[p] = p; }
    bool unsorted = false;
    for (int p = 1; p < nbk; p++) {
      const int e = valid ? st[2 + p] : p, e0 = valid ? st[1 + p] : p - 1;
      if (sapkey[e] < sapkey[e0]) unsorted = true;
    }
    for (int t = 0; t < nbk; t++) {
      int pos = sapinit[t];
      if (unsorted) {
        pos = 0;
        for (int u = 0; u < nbk; u++)
          if (u != t && ob_sap_precedes(ob_sap_keyorder(sapkey[u]), ob_sap_keyorder(sapkey[t]), sapinit[u], sapinit[t])) pos++;
      }
      sappos[t] = pos;
    }
    if (unsorted) { for (int t = 0; t < nbk; t++) st[2 + sappos[t]] = t; st[0] = 1; }
    else if (!valid) st[0] = 0;
    st[1] = nbk;
  }
  std::vector<PairRec> recs;
  for (int a = 0; a < ng; a++) {
    if (!en[a]) continue;
    const ObGeom &ga = d.geom[(size_t)w * d.NG + glist[a]];
    for (int b = a + 1; b < ng; b++) {
      if (!en[b]) continue;
      const ObGeom &gb = d.geom[(size_t)w * d.NG + glist[b]];
      PairRec r;
      int first_is_a;
      if (stype == OB_SPACE_HASH) {
        if (!ob_aabb_pair_filter(ga.body, gb.body, ga.cat, ga.col, gb.cat, gb.col, &aabb[6 * a], &aabb[6 * b])) continue;
        if (!ob_hash_pair_key(a, b, cb[a], cb[b], hr[a], hr[b], br[a], br[b], nh, nbig, &r.key, &first_is_a)) continue;
      } else if (stype == OB_SPACE_SAP) {
        if (!ob_pair_filter_noaabb(ga.body, gb.body, ga.cat, ga.col, gb.cat, gb.col)) continue;
        const bool ia = cb[a].level == OB_LEVEL_BIG, ib = cb[b].level == OB_LEVEL_BIG;
        for (int k = 0; k < 7; k++) r.key.k[k] = 0;
        if (!ia && !ib) {
          const int pa = sappos[hr[a]], pb = sappos[hr[b]];
          first_is_a = pa < pb;
          const int K = first_is_a ? a : b, J = first_is_a ? b : a;
          if (!ob_sap_sweep_test(sapkey[hr[J]], &aabb[6 * K], &aabb[6 * J], ax0, ax1, ax2)) continue;
          r.key.k[1] = first_is_a ? pa : pb; r.key.k[2] = first_is_a ? pb : pa;
        } else if (ia && ib) { r.key.k[0] = 1; r.key.k[1] = br[a]; r.key.k[3] = br[b]; first_is_a = 1; }
        else { r.key.k[0] = 1; r.key.k[1] = ia ? br[a] : br[b]; r.key.k[2] = 1; r.key.k[3] = ia ? hr[b] : hr[a]; first_is_a = ia; }
      } else {
        if (!ob_aabb_pair_filter(ga.body, gb.body, ga.cat, ga.col, gb.cat, gb.col, &aabb[6 * a], &aabb[6 * b])) continue;
        for (int k = 0; k < 7; k++) r.key.k[k] = 0;
        r.key.k[1] = a; r.key.k[2] = b; first_is_a = 1;
      }
      r.o1 = first_is_a ? glist[a] : glist[b];
      r.o2 = first_is_a ? glist[b] : glist[a];
      recs.push_back(r);
    }
  }
  std::sort(recs.begin(), recs.end(), pair_less);
  int np = (int)recs.size();
  if (np > d.NP) { W.status |= OB_ERR_PAIR_OVERFLOW; np = d.NP; }
  d.npairs[w] = np;
  int *pairs = d.pairs + (size_t)w * d.NP * 2;
  for (int i = 0; i < np; i++) { pairs[2 * i] = recs[i].o1; pairs[2 * i + 1] = recs[i].o2; }

  // narrowphase + policy, contact joints in creation order
  ObContact *cout = d.contacts + (size_t)w * d.NC;
  int nc = 0;
  std::vector<int> walk_of(d.NG, -1);
  for (int i = 0; i < ng; i++) walk_of[glist[i]] = i;
  for (int i = 0; i < np; i++) {
    int o1 = pairs[2 * i], o2 = pairs[2 * i + 1];
    const ObGeom &G1 = d.geom[(size_t)w * d.NG + o1], &G2 = d.geom[(size_t)w * d.NG + o2];
    const int row = ob_policy_row(d.policy, G1.cat, G2.cat);   // the row of the policy table that serves this pair
    if (row < 0) continue;
    const ObPolicy &pol = d.policy[row];
    if (pol.skip_static_pairs && G1.body < 0 && G2.body < 0) continue;
    if (pol.skip_if_connected && d.NJ) {   // dAreConnectedExcluding(b1, b2, dJointTypeContact), ode.cpp:1529-1537
      const int b1 = G1.body, b2 = G2.body;
      bool connected = false;
      if (b1 >= 0 && b2 >= 0) {
        const unsigned short *ps = d.padjstart + (size_t)w * (d.NB + 1), *pa = d.padj + (size_t)w * 2 * d.NJ;
        const ObJoint *pj = d.joint + (size_t)w * d.NJ;
        for (int k = ps[b1]; k < ps[b1 + 1]; k++) {
          const ObJoint &jj = pj[pa[k]];
          const int other = jj.b1 == b1 ? jj.b2 : jj.b1;
          if (other == b2) connected = true;
        }
      }
      if (connected) continue;
    }
    ObCg cg[OB_MAXC_LOCAL];
    int swapped;
    int flags = pol.max_contacts > OB_MAXC_LOCAL ? OB_MAXC_LOCAL : pol.max_contacts;
    int bverr = 0;
    int n = ob_collide_pair_xf_t<true, OB_MAXC_LOCAL>(&pose[walk_of[o1]], &pose[walk_of[o2]], d.any_xf, flags, cg, &swapped, d.meshes, &bverr);
    if (bverr) W.status |= OB_ERR_BVH_STACK;
    for (int k = 0; k < n; k++) {
      if (nc >= d.NC) { W.status |= OB_ERR_CONTACT_OVERFLOW; break; }
      ObContact &c = cout[nc++];
      for (int j = 0; j < 3; j++) { c.pos[j] = cg[k].pos[j]; c.normal[j] = cg[k].normal[j]; }
      c.depth = cg[k].depth; c.g1 = o1; c.g2 = o2; c.side1 = cg[k].side1; c.side2 = cg[k].side2; c.policy = row;
    }
  }
  d.ncontacts[w] = nc;
  d.counters->pairs += np;
}

static void step_world(ObBatchDev &d, int w, real h, int taps) {
  ObWorld &W = d.world[w];
  int nb = W.nb, nc = d.ncontacts[w];
  ObBodyDyn *bd = d.bdyn + (size_t)w * d.NB;
  const ObBodyConst *bc = d.bconst + (size_t)w * d.NB;
  const ObGeom *geoms = d.geom + (size_t)w * d.NG;
  const ObContact *con = d.contacts + (size_t)w * d.NC;
  const real stepsize1 = ob_recip(h);

  // joint id space: contact joints 0..nc-1 (creation order), permanent joints nc..nc+nj-1
  const int nj = d.njoints[w];
  const ObJoint *pjoint = d.joint + (size_t)w * (d.NJ ? d.NJ : 1);
  const int njall = nc + nj;
  // contact joint -> bodies (dJointAttach: body1==0 -> swap + REVERSE), ode.cpp:1368-1377
  std::vector<int> jb1(njall), jb2(njall), jrev(njall);
  for (int j = 0; j < nc; j++) {
    int b1, b2;
    jrev[j] = 0;
    if (d.dropin) { b1 = con[j].side1; b2 = con[j].side2; jrev[j] = con[j].policy; }
    else { b1 = geoms[con[j].g1].body; b2 = geoms[con[j].g2].body; if (b1 < 0) { b1 = b2; b2 = -1; jrev[j] = 1; } }
    jb1[j] = b1; jb2[j] = b2;
  }
  for (int j = 0; j < nj; j++) { jb1[nc + j] = pjoint[j].b1; jb2[nc + j] = pjoint[j].b2; jrev[nc + j] = 0; }
  // body joint lists: this step's contacts newest first, then the permanent joints in list order
  std::vector<std::vector<int> > adj(nb);   // joint ids
  for (int j = nc - 1; j >= 0; j--) {
    if (jb1[j] >= 0) adj[jb1[j]].push_back(j);
    if (jb2[j] >= 0) adj[jb2[j]].push_back(j);
  }
  {
    const unsigned short *ps = d.padjstart + (size_t)w * (d.NB + 1), *pa = d.padj + (size_t)w * 2 * (d.NJ ? d.NJ : 1);
    for (int b = 0; b < nb; b++) for (int k = ps[b]; k < ps[b + 1]; k++) adj[b].push_back(nc + pa[k]);
  }
  // auto-disable (util.cpp:99-233)
  for (int b = 0; b < nb; b++) {
    if (adj[b].empty()) continue;
    const size_t bi = (size_t)w * d.NB + b;
    ob_auto_disable(bd[b], bc[b], h, d.NADIS ? d.adisbuf + bi * d.NADIS * 6 : (real *)0, d.NADIS ? d.adisctl + bi * 2 : (int *)0);
  }
  // islands (util.cpp:411-487)
  std::vector<int> btag(nb, 0), jtag(njall, 0), ibody, ijoint, isz;
  std::vector<int> stack;
  for (int bb = 0; bb < nb; bb++) {
    if (btag[bb]) continue;
    if (bd[bb].flags & OB_BODY_DISABLED) { btag[bb] = -1; continue; }
    btag[bb] = 1;
    size_t b0 = ibody.size(), j0 = ijoint.size();
    ibody.push_back(bb);
    int b = bb;
    while (true) {
      for (size_t k = 0; k < adj[b].size(); k++) {
        int j = adj[b][k];
        if (!jtag[j]) {
          // isEnabled: at least one attached body with invMass > 0 (joint.cpp:66-71)
          bool enabled = (bc[jb1[j]].invMass > 0) || (jb2[j] >= 0 && bc[jb2[j]].invMass > 0);
          if (j >= nc && (pjoint[j - nc].flags & OB_JF_DISABLED)) enabled = false;
          if (enabled) {
            jtag[j] = 1;
            ijoint.push_back(j);
            int other = (jb1[j] == b) ? jb2[j] : jb1[j];
            if (other >= 0 && btag[other] <= 0) {
              btag[other] = 1;
              bd[other].flags &= ~OB_BODY_DISABLED;
              stack.push_back(other);
            }
          } else jtag[j] = -1;
        }
      }
      if (stack.empty()) break;
      b = stack.back(); stack.pop_back();
      ibody.push_back(b);
    }
    isz.push_back((int)(ibody.size() - b0));
    isz.push_back((int)(ijoint.size() - j0));
  }

  // per island: dxQuickStepper
  real *rowJ = d.rowJ + (size_t)w * d.NR * 12, *rowiMJ = d.rowiMJ + (size_t)w * d.NR * 12;
  real *rowS = d.rowS + (size_t)w * d.NR * 4, *lambda = d.lambda + (size_t)w * d.NR;
  int *rowI = d.rowI + (size_t)w * d.NR * 4;
  real *fb = d.fback + (size_t)w * (d.NC + d.NJ) * 12;
  int rowbase = 0;
  size_t bpos = 0, jpos = 0;
  std::vector<int> moved;   // geoms in dGeomMoved order
  uint32_t seed = W.seed;
  for (size_t isl = 0; isl < isz.size() / 2; isl++) {
    int inb = isz[2 * isl], inj = isz[2 * isl + 1];
    const int *ib = &ibody[bpos];
    const int *ij = inj ? &ijoint[jpos] : 0;
    std::vector<int> tag(nb, -1);
    for (int i = 0; i < inb; i++) tag[ib[i]] = i;
    std::vector<real> invIw(12 * inb), tmp1(6 * inb), fc(6 * inb, 0);
    for (int i = 0; i < inb; i++) {
      int b = ib[i];
      ob_body_preamble(bd[b].R, bc[b].I, bc[b].invI, bd[b].avel, bd[b].flags, bc[b].mass, W.gravity, &invIw[12 * i],
                       bd[b].facc, bd[b].tacc);
    }
    // rows: getInfo1 per joint (quickstep.cpp:670-688; joints with m == 0 are dropped)
    auto view = [&](int b) { ObBodyView v; v.pos = bd[b].pos; v.R = bd[b].R; v.q = bd[b].q; v.lvel = bd[b].lvel; v.avel = bd[b].avel; return v; };
    std::vector<int> jm(inj), jofs(inj);
    int m = 0;
    std::vector<ObSurface> surf(inj);
    std::vector<ObJoint> pj(inj);
    for (int k = 0; k < inj; k++) {
      const int j = ij[k];
      if (j < nc) {
        surf[k] = d.csurf ? d.csurf[(size_t)w * d.NC + j] : d.policy[con[j].policy].surface;
        jm[k] = ob_contact_info1(surf[k]);
      } else {
        pj[k] = pjoint[j - nc];
        ObBodyView B1 = view(jb1[j]), B2;
        if (jb2[j] >= 0) B2 = view(jb2[j]);
        jm[k] = ob_joint_info1(pj[k], B1, jb2[j] >= 0 ? &B2 : 0);
      }
      jofs[k] = m;
      m += jm[k];
    }
    if (rowbase + m > d.NR) { W.status |= OB_ERR_ROW_OVERFLOW; m = 0; inj = 0; }
    real *J = rowJ + (size_t)rowbase * 12, *iMJ = rowiMJ + (size_t)rowbase * 12, *S = rowS + (size_t)rowbase * 4;
    int *RI = rowI + (size_t)rowbase * 4;
    real *lam = lambda + rowbase;
    if (m > 0) {
      // getInfo2 for every joint first (it may add motor torques to tacc, joint.cpp:638-657), then tmp1
      std::vector<ObRowOut> rr(inj);
      real erp_io = W.erp;   // Info2.erp is shared by the island's joints; a ball joint overwrites it (ball.cpp:60)
      for (int k = 0; k < inj; k++) {
        int j = ij[k];
        ObRowOut &r = rr[k];
        ob_rows_defaults(r, jm[k], W.cfm);
        int b1 = jb1[j], b2 = jb2[j];
        if (j < nc) {
          real zero3[3] = {0, 0, 0};
          real fdir1[3] = {0, 0, 0};
          if (d.csurf) for (int e = 0; e < 3; e++) fdir1[e] = d.cfdir1[((size_t)w * d.NC + j) * 4 + e];
          ob_contact_info2(r, jm[k], surf[k], con[j].pos, con[j].normal, con[j].depth, fdir1, jrev[j], bd[b1].pos,
                           bd[b1].lvel, bd[b1].avel, b2 >= 0, b2 >= 0 ? bd[b2].pos : zero3, b2 >= 0 ? bd[b2].lvel : zero3,
                           b2 >= 0 ? bd[b2].avel : zero3, stepsize1, erp_io, W.min_depth, W.max_vel);
        } else {
          ObBodyView B1 = view(b1), B2;
          if (b2 >= 0) B2 = view(b2);
          real side[OB_NSIDE][4];
          ob_joint_info2(r, pj[k], B1, b2 >= 0 ? &B2 : 0, stepsize1, &erp_io, side);
          if ((side[0][0] != 0 || side[1][0] != 0 || side[2][0] != 0) && getenv("OB_HOST_VERBOSE")) fprintf(stderr, "side effect: joint type %d fm %g\n", pj[k].type, (double)side[0][0]);
          if (side[0][0] != 0 || side[1][0] != 0 || side[2][0] != 0)
            ob_apply_joint_side(pj[k].type, side, bd[b1].facc, bd[b1].tacc, b2 >= 0 ? bd[b2].facc : (real *)0, b2 >= 0 ? bd[b2].tacc : (real *)0);
        }
      }
      for (int i = 0; i < inb; i++) {
        int b = ib[i];
        ob_body_tmp1(bd[b].facc, bd[b].tacc, bd[b].lvel, bd[b].avel, bc[b].invMass, &invIw[12 * i], stepsize1, &tmp1[6 * i]);
      }
      std::vector<real> Jcopy;
      if (taps) Jcopy.resize((size_t)m * 12);
      for (int k = 0; k < inj; k++) {
        int j = ij[k];
        ObRowOut &r = rr[k];
        int b1 = jb1[j], b2 = jb2[j];
        for (int q = 0; q < jm[k]; q++) {
          int ri = jofs[k] + q;
          if (taps) memcpy(&Jcopy[(size_t)ri * 12], r.J[q], 12 * sizeof(real));
          int t1 = tag[b1], t2 = b2 >= 0 ? tag[b2] : -1;
          real b_out, adcfm;
          ob_row_finalize(r.J[q], r.c[q], r.cfm[q], t2, &tmp1[6 * t1], t2 >= 0 ? &tmp1[6 * t2] : 0, bc[b1].invMass,
                          &invIw[12 * t1], b2 >= 0 ? bc[b2].invMass : 0, t2 >= 0 ? &invIw[12 * t2] : 0, stepsize1,
                          W.sor_w, &iMJ[(size_t)ri * 12], &b_out, &adcfm);
          memcpy(&J[(size_t)ri * 12], r.J[q], 12 * sizeof(real));
          S[ri * 4 + 0] = b_out; S[ri * 4 + 1] = adcfm; S[ri * 4 + 2] = r.lo[q]; S[ri * 4 + 3] = r.hi[q];
          RI[ri * 4 + 0] = r.findex[q] >= 0 ? r.findex[q] + jofs[k] : -1;
          RI[ri * 4 + 1] = t1; RI[ri * 4 + 2] = t2; RI[ri * 4 + 3] = j;
          lam[ri] = 0;
        }
      }
      // order (quickstep.cpp:409-424)
      std::vector<int> order(m);
      int head = 0, tail = m - 1;
      for (int i = 0; i < m; i++) { if (RI[i * 4] == -1) order[head++] = i; else order[tail--] = i; }
      for (int it = 0; it < W.iters; it++) {
        if ((it & 7) == 0) {
          for (int i = 1; i < m; i++) {
            seed = ob_lcg_next(seed);
            int swapi = ob_randint_fold(seed, (uint32_t)(i + 1));
            int t = order[i]; order[i] = order[swapi]; order[swapi] = t;
          }
        }
        // mirror of the CUDA level schedule (ob_step_kernel.cuh): execute rows level by level;
        // inside a level in an arbitrary (here: reversed) order — results must not change
        std::vector<int> lvl(m), lastl(inb, 0), sched;
        int nlev = 0;
        for (int i = 0; i < m; i++) {
          int idx = order[i];
          int t1 = RI[idx * 4 + 1], t2 = RI[idx * 4 + 2];
          int lv = lastl[t1];
          if (t2 >= 0 && lastl[t2] > lv) lv = lastl[t2];
          lv++;
          lastl[t1] = lv; if (t2 >= 0) lastl[t2] = lv;
          lvl[i] = lv; if (lv > nlev) nlev = lv;
        }
        for (int l = 1; l <= nlev; l++) for (int i = m - 1; i >= 0; i--) if (lvl[i] == l) sched.push_back(order[i]);
        for (int i = 0; i < m; i++) {
          int idx = getenv("OB_HOST_LEVELS") ? sched[i] : order[i];
          int t1 = RI[idx * 4 + 1], t2 = RI[idx * 4 + 2], fi = RI[idx * 4];
          lam[idx] = ob_sor_row(&J[(size_t)idx * 12], &iMJ[(size_t)idx * 12], S[idx * 4], S[idx * 4 + 1], S[idx * 4 + 2],
                                S[idx * 4 + 3], fi, fi >= 0 ? lam[fi] : 0, lam[idx], &fc[6 * t1], t2 >= 0 ? &fc[6 * t2] : 0);
        }
      }
      if (taps) {
        for (int k = 0; k < inj; k++) {   // Multiply1_12q1 (quickstep.cpp:70-101), body 1 then body 2
          real acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
          for (int q = 0; q < jm[k]; q++) {
            real s = lam[jofs[k] + q];
            for (int e = 0; e < 12; e++) acc[e] += Jcopy[(size_t)(jofs[k] + q) * 12 + e] * s;
          }
          if (jb2[ij[k]] < 0) for (int e = 6; e < 12; e++) acc[e] = 0;
          real *o = fb + (size_t)(ij[k] < nc ? ij[k] : d.NC + (ij[k] - nc)) * 12;
          for (int e = 0; e < 12; e++) o[e] = acc[e];
        }
      }
    }
    for (int i = 0; i < inb; i++) {
      int b = ib[i];
      ob_body_velocity_update(bd[b].lvel, bd[b].avel, m > 0 ? &fc[6 * i] : 0, bd[b].facc, bd[b].tacc, bc[b].invMass,
                              &invIw[12 * i], h);
    }
    for (int i = 0; i < inb; i++) {
      int b = ib[i];
      ob_step_body(bd[b].pos, bd[b].q, bd[b].R, bd[b].lvel, bd[b].avel, bd[b].flags, h, bc[b].max_angular_speed,
                   bc[b].finite_rot_axis, bc[b].damp_lin_scale, bc[b].damp_ang_scale, bc[b].damp_lin_thr,
                   bc[b].damp_ang_thr);
      for (int g = bc[b].geom_first; g >= 0; g = geoms[g].body_next) moved.push_back(g);
    }
    for (int i = 0; i < inb; i++) {
      int b = ib[i];
      for (int k = 0; k < 4; k++) { bd[b].facc[k] = 0; bd[b].tacc[k] = 0; }
    }
    rowbase += m;
    bpos += isz[2 * isl]; jpos += isz[2 * isl + 1];
    d.counters->body_steps += inb;
    d.counters->rows += m;
    for (int k = 0; k < inj; k++) if (ij[k] < nc) d.counters->contacts += 1;
    d.counters->islands += 1;
  }
  W.seed = seed;
  for (size_t i = 0; i < ibody.size(); i++) d.ibody[(size_t)w * d.NB + i] = (unsigned char)ibody[i];
  d.stepinfo[(size_t)w * SI_WORDS + SI_NIB] = (int)ibody.size();
  d.nrows[w] = rowbase;
  // space list update: every moved (clean) geom goes to the head, in order (collision_space.cpp:47-75)
  int *glist = d.glist + (size_t)w * d.NG;
  int ng = W.ng;
  std::vector<char> ismoved(d.NG, 0);
  std::vector<int> nl;
  if (W.space_type == OB_SPACE_SAP) {
    // dxSAPSpace::dirty per moved geom (collision_sapspace.cpp:363-387); stored as DirtyList + GeomList
    std::vector<int> G(glist, glist + ng), posof(d.NG, -1);
    for (int i = 0; i < ng; i++) posof[G[i]] = i;
    for (size_t i = 0; i < moved.size(); i++) {
      const int g = moved[i], idx = posof[g], last = G.back();
      G[idx] = last; posof[last] = idx; G.pop_back();
      nl.push_back(g);
    }
    for (size_t i = 0; i < G.size(); i++) nl.push_back(G[i]);
    W.sap_ndirty = (int)moved.size();
  } else {
    for (int i = (int)moved.size() - 1; i >= 0; i--) { nl.push_back(moved[i]); ismoved[moved[i]] = 1; }
    for (int i = 0; i < ng; i++) if (!ismoved[glist[i]]) nl.push_back(glist[i]);
  }
  for (int i = 0; i < ng; i++) glist[i] = nl[i];
  d.counters->steps += 1;
}

int obk_run_phases(ObBackend *b, real h, int phases, int taps, char *, size_t) {
  ObBatchDev &d = b->d;
  if (d.large) return -1;
  for (int w = 0; w < d.W; w++) {
    if (phases & OBK_PHASE_COLLIDE) collide_world(d, w);
    if (phases & OBK_PHASE_STEP) step_world(d, w, h, taps);
  }
  return 0;
}
int obk_collide_pair(const ObPose *a, const ObPose *b, int flags, ObCg *out, const ObMeshDev *meshes2, char *, size_t) {
  int swapped, bverr = 0;
  int maxc = flags & 0xffff;
  if (maxc > OB_MAXC_LOCAL) maxc = OB_MAXC_LOCAL;
  const int n = ob_collide_pair(*a, *b, (flags & ~0xffff) | maxc, out, &swapped, meshes2, &bverr);
  return bverr ? -1 : n;
}
int obk_collide2(ObBackend *b, const ObPose *q, const int *qbody, const uint32_t *qcat, const uint32_t *qcol, const ObMeshDev *qmesh,
                 int nq, unsigned char *hit, char *, size_t) {
  ObBatchDev &d = b->d;
  const int ng = d.world[0].ng;
  for (int qi = 0; qi < nq; qi++)
    for (int g = 0; g < ng; g++) {
      const ObGeom &G = d.geom[g];
      unsigned char h = 0;
      if ((G.flags & OB_GEOM_ENABLED) && !(G.flags & OB_GEOM_ZERO_SIZED)) {
        ObPose p;
        geom_pose(d, 0, g, &p);
        real a[6], bb[6];
        ob_aabb(p, a, d.meshes);
        ob_aabb(q[qi], bb, &qmesh[qi]);
        h = ob_aabb_pair_filter(G.body, qbody[qi], G.cat, G.col, qcat[qi], qcol[qi], a, bb) ? 1 : 0;
      }
      hit[(size_t)qi * d.NG + g] = h;
    }
  return 0;
}
int obk_mesh_upload(const float *verts, int nverts, const int *tris, int ntris, const ObBvNode *nodes, const unsigned char *useflags, int, ObMeshDev *io) {
  float *v = (float *)malloc(sizeof(float) * 3 * (size_t)nverts);
  int *t = (int *)malloc(sizeof(int) * 3 * (size_t)ntris);
  ObBvNode *n = (ObBvNode *)malloc(sizeof(ObBvNode) * (size_t)(ntris - 1));
  memcpy(v, verts, sizeof(float) * 3 * (size_t)nverts);
  memcpy(t, tris, sizeof(int) * 3 * (size_t)ntris);
  memcpy(n, nodes, sizeof(ObBvNode) * (size_t)(ntris - 1));
  int *vf = (int *)malloc(sizeof(int) * (size_t)(nverts > 0 ? nverts : 1));
  for (int i = 0; i < nverts; i++) vf[i] = -1;
  for (int c = 0; c < 3 * ntris; c++) { const int vi = tris[c]; if (vi >= 0 && vi < nverts && vf[vi] < 0) vf[vi] = c; }
  io->verts = v; io->tris = t; io->nodes = n; io->nverts = nverts; io->ntris = ntris; io->vfirst = vf;
  io->useflags = 0;
  if (useflags) { unsigned char *u = (unsigned char *)malloc((size_t)ntris); memcpy(u, useflags, (size_t)ntris); io->useflags = u; }
  return 0;
}
void obk_mesh_free(ObMeshDev *m) { free((void *)m->verts); free((void *)m->tris); free((void *)m->nodes); free((void *)m->vfirst); free((void *)m->useflags); m->useflags = 0; m->verts = 0; m->tris = 0; m->nodes = 0; m->vfirst = 0; }
int obk_step(ObBackend *b, real h, int nsteps, int taps, char *err, size_t errlen) {
  ObBatchDev &d = b->d;
  if (d.large) {
    for (int s = 0; s < nsteps; s++) {
      LargeHostStats st;
      const int rc = large_step_host(d, h, taps, &st, b->split);
      if (rc == -7) { snprintf(err, errlen, "split SOR: a peer rank did not reach the barrier within the timeout"); return -1; }
      if (rc) { snprintf(err, errlen, "large-world step failed (%d)", rc); return -1; }
      g_lw_stat[0] = st.np; g_lw_stat[1] = st.ncontacts; g_lw_stat[2] = st.ncp; g_lw_stat[3] = st.nsolved; g_lw_stat[4] = st.ncol; g_lw_stat[5] = st.rounds;
      if (getenv("OB_LW_VERBOSE")) fprintf(stderr, "lw: pairs %d contacts %d cpairs %d colours %d rounds %d\n", st.np, st.ncontacts, st.ncp, st.ncol, st.rounds);
    }
    return 0;
  }
  for (int s = 0; s < nsteps; s++) {
    if ((taps & 1) && !d.dropin) memset(d.fback, 0, sizeof(real) * 12 * (size_t)d.W * (d.NC + d.NJ));
    for (int w = 0; w < d.W; w++) {
      collide_world(d, w);
      step_world(d, w, h, taps);
    }
  }
  return 0;
}

// split mirror: the "handle" is the exporting batch's fc / flag pointers (ranks are threads of one process)
struct HostSplitHandle { void *fc; void *flags; unsigned long long nb; };
int obk_split_export(ObBackend *b, void *handle128, char *err, size_t errlen) {
  if (!b->d.large) { snprintf(err, errlen, "only the large-world path splits over GPUs"); return -1; }
  if (!b->split_fc) {
    b->split_fc = halloc<real>(b, (size_t)b->d.NB * 6);
    b->split_flags = new std::atomic<unsigned>[OB_LW_FLAG_WORDS];
    for (int k = 0; k < OB_LW_FLAG_WORDS; k++) ((std::atomic<unsigned> *)b->split_flags)[k].store(0);
  }
  memset(handle128, 0, OBK_SPLIT_HANDLE_BYTES);
  HostSplitHandle H = {b->split_fc, b->split_flags, (unsigned long long)b->d.NB};
  memcpy(handle128, &H, sizeof H);
  return 0;
}
int obk_split_attach(ObBackend *b, int rank, int nranks, const void *handles, char *err, size_t errlen) {
  if (!b->d.large || !b->split_fc) { snprintf(err, errlen, "export before attach (large-world batches only)"); return -1; }
  if (nranks < 1 || nranks > OB_LW_MAXRANKS || rank < 0 || rank >= nranks) { snprintf(err, errlen, "bad rank %d of %d", rank, nranks); return -1; }
  LargeHostSplit *S = new LargeHostSplit;
  S->rank = rank; S->nranks = nranks; S->base = 0;
  S->timeout_ms = getenv("OB_LW_SPLIT_TIMEOUT_MS") ? (unsigned)atoi(getenv("OB_LW_SPLIT_TIMEOUT_MS")) : 20000u;
  for (int r = 0; r < nranks; r++) {
    HostSplitHandle H;
    memcpy(&H, (const unsigned char *)handles + (size_t)r * OBK_SPLIT_HANDLE_BYTES, sizeof H);
    if (H.nb != (unsigned long long)b->d.NB) { snprintf(err, errlen, "rank %d holds a world of another size", r); delete S; return -1; }
    S->fc[r] = (real *)H.fc; S->flags[r] = (std::atomic<unsigned> *)H.flags;
  }
  if (S->fc[rank] != b->split_fc) { snprintf(err, errlen, "handle %d is not this batch's own export", rank); delete S; return -1; }
  b->split = nranks > 1 ? S : 0;
  return 0;
}


int obk_libm(int, int, const float *, const float *, float *) { return -1; }   // device-only diagnostic

#include "../../ode-0.12_b200/csrc/ob_ray.h"
int obk_raycast(ObBackend *b, int nrays, const real *origin3, const real *dir3, const real *length, int ray_flags, uint32_t cat, uint32_t col,
                ObRayHit *hits, char *, size_t) {
  ObBatchDev &d = b->d;
  for (int w = 0; w < d.W; w++) {
    ObWorld &W = d.world[w];
    const int ng = W.ng;
    const int *glist = d.glist + (size_t)w * d.NG;
    const int rot = W.space_type == OB_SPACE_SAP ? W.sap_ndirty : 0;
    for (int r = 0; r < nrays; r++) {
      const size_t t = (size_t)w * nrays + r;
      ObPose ray;
      ob_ray_pose(origin3 + 3 * t, dir3 + 3 * t, length[t], ray_flags, &ray);
      real rab[6];
      ob_aabb(ray, rab, d.meshes);
      ObRayHit h;
      for (int k = 0; k < 3; k++) { h.pos[k] = 0; h.normal[k] = 0; }
      h.depth = length[t]; h.geom = -1;
      bool have = false;
      int bverr = 0;
      for (int i = 0; i < ng; i++) {
        const int gi = glist[i + rot < ng ? i + rot : i + rot - ng];
        const ObGeom &g = d.geom[(size_t)w * d.NG + gi];
        if (!(g.flags & OB_GEOM_ENABLED) || (g.flags & OB_GEOM_ZERO_SIZED) || g.type == OB_GEOM_RAY || g.type == OB_GEOM_SPACE) continue;
        ObPose p;
        geom_pose(d, w, gi, &p);
        ObCg c;
        if (ob_ray_vs_geom(ray, rab, cat, col, p, g.body, g.cat, g.col, d.meshes, &c, &bverr) && (!have || c.depth < h.depth)) {
          have = true;
          for (int k = 0; k < 3; k++) { h.pos[k] = c.pos[k]; h.normal[k] = c.normal[k]; }
          h.depth = c.depth; h.geom = gi;
        }
      }
      memset(&hits[t], 0, sizeof(ObRayHit));
      for (int k = 0; k < 3; k++) { hits[t].pos[k] = h.pos[k]; hits[t].normal[k] = h.normal[k]; }
      hits[t].depth = h.depth; hits[t].geom = h.geom;
    }
  }
  return 0;
}

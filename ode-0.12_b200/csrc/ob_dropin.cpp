// ob_dropin.cpp — the compute entry points of the classic ODE API (dSpaceCollide, dCollide,
// dWorldQuickStep; ode/src/collision_space.cpp:748, collision_kernel.cpp:292, ode.cpp:1807)
// served by the CUDA kernels through a hidden batch of ONE world.
//
//   dSpaceCollide   : upload the world/space state, run k_collide (broadphase in the reference's
//                     callback order + narrowphase for every pair with the max-contacts value the
//                     caller used last time), copy pairs + contacts back, then call the user's
//                     near callback per pair, in order, on the calling thread.
//   dCollide        : inside that callback, served from those results when the request matches
//                     (same pair orientation, same effective max-contacts); otherwise one pair
//                     is collided on the GPU on demand (k_collide_pair).
//   dWorldQuickStep : upload state + the contact joints the callback created (with their own
//                     dSurfaceParameters / fdir1) and run k_prep -> k_sched -> k_sor -> k_post;
//                     copy body state, the process-global dRand seed and dJointFeedback back, and
//                     apply dGeomMoved in stepping order on the host lists.
// Nothing is computed on the CPU here; without a usable GPU these calls raise dError.
// Semantics are the point of this path, not speed: one small world cannot fill a GPU.
#include <string.h>
#include <algorithm>
#include <map>
#include <vector>
#include "ob_batch.h"
#include "ob_trimesh_host.h"

void ob_joints_prestep_bookkeeping(dxWorld *w);   // ob_joints.cpp

struct ObDropin {
  dxBatch *B;
  dxWorld *world;      // world stepped through this context (own_world when the space holds no bodies)
  dxSpace *space;      // space collided through this context (own_space when only a world is stepped)
  dxWorld *own_world;
  dxSpace *own_space;
  int maxc_hint;       // max-contacts value the near callback passed to dCollide last time (NEXT frame's narrowphase runs with it)
  int maxc_used;       // max-contacts value this frame's batched narrowphase actually ran with (what the cached contacts obey)
  int kcap;            // per-pair contact capacity of the collide kernel that produced them (8 without trimeshes, else OB_MAXC_LOCAL)
  bool in_collide;     // results below are valid (only while the callbacks run)
  int pairs_cap;       // 0: the batch layer's default pair capacity; else the capacity to bind with (grown after an overflow)
  std::vector<int> pairs;              // (o1,o2) geom indices in callback order
  std::vector<ObContact> contacts;     // grouped by pair, pair order
  std::map<std::pair<int, int>, std::pair<int, int> > pair_contacts;   // (o1,o2) -> (first contact, count)
};
static std::vector<ObDropin *> g_ctx;

static void ctx_free_batch(ObDropin *c) {
  if (c->B) { dBatchDestroy(c->B); c->B = 0; }
}
static void ctx_drop(size_t i) {
  ObDropin *c = g_ctx[i];
  ctx_free_batch(c);
  g_ctx.erase(g_ctx.begin() + i);
  if (c->own_world) { dxWorld *w = c->own_world; c->own_world = 0; dWorldDestroy(w); }
  if (c->own_space) { dxSpace *s = c->own_space; c->own_space = 0; dSpaceDestroy(s); }
  delete c;
}
// called by dWorldDestroy / dGeomDestroy(space) so no context keeps a dangling pointer
void ob_dropin_forget_world(dxWorld *w) {
  for (size_t i = 0; i < g_ctx.size();) { if (g_ctx[i]->world == w && g_ctx[i]->own_world != w) ctx_drop(i); else i++; }
}
void ob_dropin_forget_space(dxSpace *s) {
  for (size_t i = 0; i < g_ctx.size();) { if (g_ctx[i]->space == s && g_ctx[i]->own_space != s) ctx_drop(i); else i++; }
}

static void world_of_space_rec(dxSpace *s, dxWorld **w, bool *mixed) {
  for (dxGeom *g = s->first; g; g = g->next) {
    if (g->is_space) world_of_space_rec((dxSpace *)g, w, mixed);
    else if (g->body) { if (!*w) *w = g->body->world; else if (*w != g->body->world) *mixed = true; }
  }
}
static dxWorld *world_of_space(dxSpace *s, bool *mixed) {
  dxWorld *w = 0;
  *mixed = false;
  world_of_space_rec(s, &w, mixed);   // sub-spaces included (demo_buggy's car_space inside the main space)
  return w;
}

static ObDropin *ctx_new(dxWorld *w, dxSpace *s) {
  ObDropin *c = new ObDropin;
  c->B = 0; c->world = w; c->space = s; c->own_world = 0; c->own_space = 0; c->maxc_hint = 8; c->maxc_used = 8; c->kcap = 8; c->in_collide = false; c->pairs_cap = 0;
  g_ctx.push_back(c);
  return c;
}

// does the bound batch still describe the world/space (same objects, enough capacity)?
static bool batch_matches(ObDropin *c, int need_contacts) {
  dxBatch *B = c->B;
  if (!B) return false;
  if (need_contacts > B->caps.NC) return false;
  if ((c->world->qs_iterations + 7) / 8 > B->caps.NEP) return false;
  int i = 0;
  for (dxBody *b = c->world->firstbody; b; b = b->next, i++) {
    if (i >= B->nb[0] || B->bodies[0][i] != b || b->batch_index != i) return false;
    if (b->adis.average_samples > 1 && (int)b->adis.average_samples > B->caps.NADIS) return false;   // deeper sample buffer than the batch has
  }
  if (i != B->nb[0]) return false;
  if (c->space->count != B->ng[0]) return false;
  for (dxGeom *g = c->space->first; g; g = g->next)
    if (g->batch_index < 0 || g->batch_index >= B->ng[0] || B->geoms[0][g->batch_index] != g) return false;
  for (dxGeom *g = c->space->first; g; g = g->next)
    if (g->type == dTriMeshClass) {
      bool have = false;
      for (size_t mi = 0; mi < B->meshes.size(); mi++) have |= B->meshes[mi] == g->tmdata && !g->tmdata->dev.empty();
      if (!have) return false;
    }
  std::vector<dxJoint *> js;
  for (dxJoint *j = c->world->firstjoint; j; j = j->next) if (j->type != dJointTypeContact) js.push_back(j);
  std::reverse(js.begin(), js.end());
  if (js != B->joints[0]) return false;
  return true;
}

static bool ctx_ensure(ObDropin *c, int need_contacts, const char *who) {
  if (batch_matches(c, need_contacts)) return true;
  ctx_free_batch(c);
  dBatchDesc desc;
  memset(&desc, 0, sizeof desc);
  int cap = std::max(64, 16 * std::max(1, c->space->count));
  while (cap < need_contacts) cap *= 2;
  desc.max_contacts_per_world = cap;
  desc.max_pairs_per_world = c->pairs_cap;
  c->B = ob_batch_create(1, &c->world, &c->space, &desc, 1);
  if (!c->B) { ob_error(0, "%s: %s", who, dB200LastError()); return false; }
  return true;
}

// the context (bound batch of one world) that serves `space`, created or re-bound as needed
static ObDropin *ctx_for_space(dxSpace *space, const char *who) {
  bool mixed;
  dxWorld *w = world_of_space(space, &mixed);
  if (mixed) { ob_error(0, "%s: geoms of one space attached to bodies of different worlds are not supported", who); return 0; }
  ObDropin *c = 0;
  for (size_t i = 0; i < g_ctx.size(); i++) if (g_ctx[i]->space == space) c = g_ctx[i];
  if (c && w && c->world != w && c->own_world != c->world) { ob_dropin_forget_space(space); c = 0; }
  if (!c) {
    // a context created by dWorldQuickStep for this world (with a placeholder space) is superseded
    for (size_t i = 0; i < g_ctx.size();) { if (w && g_ctx[i]->world == w && g_ctx[i]->own_space) ctx_drop(i); else i++; }
    c = ctx_new(w, space);
    if (!w) { c->own_world = dWorldCreate(); c->world = c->own_world; }
  } else if (w && c->own_world) {   // bodies appeared in a space that had none
    dxWorld *ow = c->own_world; ctx_free_batch(c); c->own_world = 0; c->world = w; dWorldDestroy(ow);
  }
  if (!ctx_ensure(c, 0, who)) return 0;
  return c;
}

void ob_dropin_space_collide(dxSpace *space, void *data, dNearCallback *cb) {
  ObDropin *c = ctx_for_space(space, "dSpaceCollide");
  if (!c) return;
  dxBatch *B = c->B;
  char err[512] = "";
  ObPolicy pol;
  memset(&pol, 0, sizeof pol);
  pol.cat_mask1 = pol.cat_mask2 = ~0u;
  pol.max_contacts = c->maxc_hint; pol.skip_if_connected = 0;   // the callback decides, not a policy
  c->maxc_used = c->maxc_hint;
  c->kcap = (B->caps.nmesh || B->caps.any_xf) ? OB_MAXC_LOCAL : 8;   // CGCAP of the k_collide instantiation that serves this batch
  int rc = ob_batch_upload(B);
  rc |= obk_h2d(B->bk, B->caps.policy, &pol, sizeof pol);
  if (rc) { ob_error(0, "dSpaceCollide: upload failed"); return; }
  if (obk_run_phases(B->bk, 0, OBK_PHASE_COLLIDE, 0, err, sizeof err)) { ob_error(0, "dSpaceCollide: %s", err); return; }
  int np = 0, nc = 0;
  ObWorld hw;
  rc = obk_d2h(B->bk, &np, B->caps.npairs, sizeof(int));
  rc |= obk_d2h(B->bk, &nc, B->caps.ncontacts, sizeof(int));
  rc |= obk_d2h(B->bk, &hw, B->caps.world, sizeof hw);
  if (rc) { ob_error(0, "dSpaceCollide: download failed"); return; }
  if (hw.status & OB_ERR_PAIR_OVERFLOW) {
    // more overlapping pairs than the bound capacity (default max(256, 12 per geom)): the reference has no such limit, so rebind
    // with room for every pair of the space and collide again
    const long long all = (long long)c->space->count * (c->space->count - 1) / 2 + 1;
    if (c->pairs_cap >= all || all > 60000000) { hw.status = 0; obk_h2d(B->bk, B->caps.world, &hw, sizeof hw); ob_error(0, "dSpaceCollide: more than %d overlapping pairs", B->caps.NP); return; }
    c->pairs_cap = (int)all;
    ctx_free_batch(c);
    ob_dropin_space_collide(space, data, cb);
    return;
  }
  c->pairs.resize((size_t)2 * np);
  c->contacts.resize(nc);
  if (np) rc |= obk_d2h(B->bk, c->pairs.data(), B->caps.pairs, sizeof(int) * 2 * np);
  if (nc) rc |= obk_d2h(B->bk, c->contacts.data(), B->caps.contacts, sizeof(ObContact) * nc);
  if (rc) { ob_error(0, "dSpaceCollide: download failed"); return; }
  c->pair_contacts.clear();
  const bool contacts_complete = !(hw.status & OB_ERR_CONTACT_OVERFLOW);
  if (hw.status) { hw.status = 0; obk_h2d(B->bk, B->caps.world, &hw, sizeof hw); }
  if (contacts_complete) {
    int k = 0;
    for (int p = 0; p < np; p++) {
      const int o1 = c->pairs[2 * p], o2 = c->pairs[2 * p + 1];
      const int k0 = k;
      while (k < nc && c->contacts[k].g1 == o1 && c->contacts[k].g2 == o2) k++;
      c->pair_contacts[std::make_pair(o1, o2)] = std::make_pair(k0, k - k0);
    }
  }
  // cleanGeoms (collision_space.cpp:405-417): dirty flags cleared, the space is locked while the callbacks run
  dSpaceClean(space);
  space->lock_count++;
  c->in_collide = true;
  std::vector<int> pairs = c->pairs;   // the callback may re-enter (dSpaceCollide2 on sub-spaces)
  for (int p = 0; p < np; p++) cb(data, B->geoms[0][pairs[2 * p]], B->geoms[0][pairs[2 * p + 1]]);
  c->in_collide = false;
  space->lock_count--;
}

static void geom_pose_host(dxGeom *g, ObPose *o) {
  dxGeom *sh = ob_geom_shape(g);   // geom transform: the encapsulated geom at T o local (collision_transform.cpp:101-108)
  o->type = sh->type;
  o->mesh = g->type == dRayClass ? ob_ray_flags(g) : (g->type == dGeomTransformClass ? OB_POSE_XFORM : 0);
  for (int i = 0; i < 4; i++) o->p[i] = sh->p[i];
  dxPosR f;
  ob_geom_final_pose(g, &f);
  for (int i = 0; i < 3; i++) o->pos[i] = f.pos[i];
  for (int i = 0; i < 12; i++) o->R[i] = f.R[i];
}
// g1 / g2 of a generated contact: the encapsulated geom of a transform unless its info mode is on (collision_transform.cpp:143-151)
static inline dxGeom *contact_geom(dxGeom *g) { return (g->type == dGeomTransformClass && !g->xf_info) ? g->xf_obj : g; }
static inline int shape_type(dxGeom *g) { dxGeom *sh = ob_geom_shape(g); return sh ? sh->type : -1; }

#define OB_CONTACT_AT(p, skip, i) ((dContactGeom *)(((char *)(p)) + (size_t)(i) * (skip)))

// dCollide with a space as an argument: dCollideSpaceGeom (collision_kernel.cpp:104-131) -- dSpaceCollide2 over the pair with a
// collector callback that calls dCollide for every reported pair while contact slots remain
struct SpaceGeomColliderData { int flags; dContactGeom *contact; int skip; };
static void space_geom_collider(void *data, dGeomID o1, dGeomID o2) {
  SpaceGeomColliderData *d = (SpaceGeomColliderData *)data;
  if (d->flags & 0xffff) {
    const int n = dCollide(o1, o2, d->flags, d->contact, d->skip);
    d->contact = (dContactGeom *)(((char *)d->contact) + (size_t)d->skip * n);
    d->flags -= n;
  }
}
static int collide_space_geom(dxGeom *o1, dxGeom *o2, int flags, dContactGeom *contact, int skip) {
  SpaceGeomColliderData data = {flags, contact, skip};
  ob_dropin_space_collide2(o1, o2, &data, &space_geom_collider);
  return (flags & 0xffff) - (data.flags & 0xffff);
}

int ob_dropin_collide(dxGeom *o1, dxGeom *o2, int flags, dContactGeom *contact, int skip) {
  const int want = flags & 0xffff;
  if (o1->is_space || o2->is_space) {
    // collider table of the space classes (dInitColliders, collision_kernel.cpp:177-182): (space, anything) straight, (geom, space)
    // reversed, of two spaces the one of the higher class number is the reversed one
    if (o1 == o2) return 0;
    const bool rev = o1->is_space ? (o2->is_space && o1->type > o2->type) : true;
    if (!rev) return collide_space_geom(o1, o2, flags, contact, skip);
    const int n = collide_space_geom(o2, o1, flags, contact, skip);
    for (int i = 0; i < n; i++) {
      dContactGeom *c = (dContactGeom *)(((char *)contact) + (size_t)skip * i);
      c->normal[0] = -c->normal[0]; c->normal[1] = -c->normal[1]; c->normal[2] = -c->normal[2];
      dGeomID tg = c->g1; c->g1 = c->g2; c->g2 = tg;
      const int ts = c->side1; c->side1 = c->side2; c->side2 = ts;
    }
    return n;
  }
  if (!(o1->gflags & GEOM_ENABLED) || !(o2->gflags & GEOM_ENABLED)) { /* dCollide itself does not test enable flags */ }
  // (1) inside dSpaceCollide's callback: serve from the batch results when they are the same computation
  for (size_t i = 0; i < g_ctx.size(); i++) {
    ObDropin *c = g_ctx[i];
    if (!c->in_collide || o1->parent_space != c->space || o2->parent_space != c->space) continue;
    const int cap = ob_pair_max_contacts(shape_type(o1), shape_type(o2), 1 << 15);
    // the cached contacts were computed with maxc_used (clamped by the kernel's per-pair capacity), not with this call's
    // value: they are served only when both give the same computation; anything else goes to the on-demand path
    const int eff_want = std::min(want, std::min(cap, OB_MAXC_LOCAL)), eff_have = std::min(c->maxc_used, std::min(cap, c->kcap));
    if (want != c->maxc_hint) c->maxc_hint = std::min(want, OB_MAXC_LOCAL);   // NEXT frame's batch narrowphase uses the caller's value
    std::map<std::pair<int, int>, std::pair<int, int> >::iterator it = c->pair_contacts.find(std::make_pair(o1->batch_index, o2->batch_index));
    if (it == c->pair_contacts.end() || eff_want != eff_have) break;
    const int k0 = it->second.first, n = it->second.second;
    if (n > want) break;   // never write past the caller's array
    for (int k = 0; k < n; k++) {
      const ObContact &s = c->contacts[k0 + k];
      dContactGeom *d = OB_CONTACT_AT(contact, skip, k);
      for (int e = 0; e < 3; e++) { d->pos[e] = s.pos[e]; d->normal[e] = s.normal[e]; }
      d->depth = s.depth; d->g1 = contact_geom(o1); d->g2 = contact_geom(o2); d->side1 = s.side1; d->side2 = s.side2;
    }
    return n;
  }
  // (2) on demand: one pair on the GPU
  if ((o1->type == dGeomTransformClass && !o1->xf_obj) || (o2->type == dGeomTransformClass && !o2->xf_obj)) return 0;   // collision_transform.cpp:122
  if (ob_pair_max_contacts(shape_type(o1), shape_type(o2), 1 << 15) == 0) return 0;   // no collider for this class pair (collision_kernel.cpp:329)
  ObPose a, b;
  geom_pose_host(o1, &a);
  geom_pose_host(o2, &b);
  ObCg cg[OB_MAXC_LOCAL];
  char err[512] = "";
  ObMeshDev m2[2];
  memset(m2, 0, sizeof m2);
  dxGeom *og[2] = {o1, o2};
  ObPose *op[2] = {&a, &b};
  for (int k = 0; k < 2; k++)
    if (og[k]->type == dTriMeshClass) {
      const ObMeshDev *md = og[k]->tmdata ? ob_trimesh_device(og[k]->tmdata, 0) : 0;
      if (!md) { ob_error(0, "dCollide: trimesh data missing or upload failed"); return 0; }
      m2[k] = *md; op[k]->mesh = k;
    }
  if (want > OB_MAXC_LOCAL && ob_pair_max_contacts(shape_type(o1), shape_type(o2), 1 << 15) > OB_MAXC_LOCAL) {
    static bool warned = false;   // the reference returns up to `want` contacts for these pairs; this build stops at OB_MAXC_LOCAL (ode.h)
    if (!warned) { warned = true; ob_message(0, "dCollide: %d contacts requested, at most %d per pair are generated (documented limit)", want, OB_MAXC_LOCAL); }
  }
  const int n = obk_collide_pair(&a, &b, flags, cg, m2, err, sizeof err);
  if (n < 0) { ob_error(0, "dCollide: %s", err); return 0; }
  for (int k = 0; k < n; k++) {
    dContactGeom *d = OB_CONTACT_AT(contact, skip, k);
    for (int e = 0; e < 3; e++) { d->pos[e] = cg[k].pos[e]; d->normal[e] = cg[k].normal[e]; }
    d->depth = cg[k].depth; d->g1 = contact_geom(o1); d->g2 = contact_geom(o2); d->side1 = cg[k].side1; d->side2 = cg[k].side2;
  }
  return n;
}

// dxSpace::collide2 (collision_space.cpp:269-286, :586-604, collision_sapspace.cpp:498-518): every enabled geom of
// the space, in list order, against the query geoms; AABBs and the collideAABBs filter run on the GPU (k_collide2).
// swap: call cb(query, member) instead of cb(member, query) (swap_callback, collision_space.cpp:764-769).
static void space_collide2(dxSpace *space, const std::vector<dxGeom *> &queries, void *data, dNearCallback *cb, bool swap) {
  if (queries.empty()) return;
  ObDropin *c = ctx_for_space(space, "dSpaceCollide2");
  if (!c) return;
  dxBatch *B = c->B;
  if (ob_batch_upload(B)) { ob_error(0, "dSpaceCollide2: upload failed"); return; }
  const int nq = (int)queries.size(), NG = B->caps.NG;
  std::vector<ObPose> qp(nq);
  std::vector<ObMeshDev> qm(nq);
  std::vector<int> qb(nq);
  std::vector<uint32_t> qcat(nq), qcol(nq);
  for (int i = 0; i < nq; i++) {
    dxGeom *g = queries[i];
    if (g->is_space) {   // a space as the query: its union box (dxSpace::computeAABB); no pose, no body
      memset(&qp[i], 0, sizeof(ObPose));
      qp[i].type = OB_GEOM_SPACE;
      dReal a[6];
      dGeomGetAABB(g, a);
      for (int k = 0; k < 6; k++) qp[i].R[k] = a[k];
      memset(&qm[i], 0, sizeof(ObMeshDev));
      qb[i] = -1; qcat[i] = (uint32_t)g->category_bits; qcol[i] = (uint32_t)g->collide_bits;
      continue;
    }
    geom_pose_host(g, &qp[i]);
    memset(&qm[i], 0, sizeof(ObMeshDev));
    if (g->type == dTriMeshClass) {
      if (!g->tmdata) { ob_error(0, "dSpaceCollide2: trimesh geom without data"); return; }
      for (int k = 0; k < 3; k++) { qm[i].aabbc[k] = g->tmdata->aabbc[k]; qm[i].aabbe[k] = g->tmdata->aabbe[k]; }
      qp[i].mesh = 0;
    }
    qb[i] = !g->body ? -1 : (g->body->world == c->world && g->body->batch_index >= 0 && g->body->batch_index < B->nb[0] &&
                             B->bodies[0][g->body->batch_index] == g->body ? g->body->batch_index : -2);
    qcat[i] = (uint32_t)g->category_bits; qcol[i] = (uint32_t)g->collide_bits;
  }
  std::vector<unsigned char> hit((size_t)nq * NG);
  char err[512] = "";
  if (obk_collide2(B->bk, qp.data(), qb.data(), qcat.data(), qcol.data(), qm.data(), nq, hit.data(), err, sizeof err)) { ob_error(0, "dSpaceCollide2: %s", err); return; }
  dSpaceClean(space);
  // member order: the space's list (SAP: GeomList, which cleanGeoms just completed)
  std::vector<dxGeom *> members;
  if (space->type == dSweepAndPruneSpaceClass) members = space->sap_geoms;
  else for (dxGeom *g = space->first; g; g = g->next) members.push_back(g);
  space->lock_count++;
  for (int i = 0; i < nq; i++)
    for (size_t k = 0; k < members.size(); k++) {
      dxGeom *g = members[k];
      if (g->batch_index < 0 || !hit[(size_t)i * NG + g->batch_index] || g == queries[i]) continue;
      if (swap) cb(data, queries[i], g); else cb(data, g, queries[i]);
    }
  space->lock_count--;
}

// dSpaceCollide2, collision_space.cpp:772-833
void ob_dropin_space_collide2(dxGeom *g1, dxGeom *g2, void *data, dNearCallback *cb) {
  dxSpace *s1 = g1->is_space ? (dxSpace *)g1 : 0, *s2 = g2->is_space ? (dxSpace *)g2 : 0;
  if (s1 && s2) {
    // sublevel rule (:778-788): of two spaces on different sublevels the deeper one is traversed, the other is taken as one geom
    const int l1 = s1->sublevel, l2 = s2->sublevel;
    if (l1 != l2) { if (l1 > l2) s2 = 0; else s1 = 0; }
  }
  if (s1 && s2) {
    if (s1 == s2) { ob_dropin_space_collide(s1, data, cb); return; }
    // iterate through the space that has the fewest geoms, calling collide2 in the other space for each one
    std::vector<dxGeom *> q;
    if (s1->count < s2->count) {
      for (dxGeom *g = s1->first; g; g = g->next) q.push_back(g);
      space_collide2(s2, q, data, cb, true);
    } else {
      for (dxGeom *g = s2->first; g; g = g->next) q.push_back(g);
      space_collide2(s1, q, data, cb, false);
    }
  } else if (s1) {
    space_collide2(s1, std::vector<dxGeom *>(1, g2), data, cb, false);
  } else if (s2) {
    space_collide2(s2, std::vector<dxGeom *>(1, g1), data, cb, true);
  } else {
    // two geoms: collideAABBs; served by a one-geom query against a throw-away space would cost more than it
    // is worth: the callback's dCollide is the computation, the AABB pre-test only prunes
    if (g1->body == g2->body && g1->body) return;
    if ((((unsigned long)g1->category_bits & (unsigned long)g2->collide_bits) || ((unsigned long)g2->category_bits & (unsigned long)g1->collide_bits)) == 0) return;
    dReal a[6], b[6];
    dGeomGetAABB(g1, a); dGeomGetAABB(g2, b);
    if (a[0] > b[1] || a[1] < b[0] || a[2] > b[3] || a[3] < b[2] || a[4] > b[5] || a[5] < b[4]) return;
    cb(data, g1, g2);
  }
}

int ob_dropin_quickstep(dxWorld *w, dReal h) {
  ObDropin *c = 0;
  for (size_t i = 0; i < g_ctx.size(); i++) if (g_ctx[i]->world == w) c = g_ctx[i];
  if (!c) {
    // world stepped without a collided space (free bodies / joints only): placeholder space.  If the
    // world's geoms live in a space that was never collided through this API, bind that one.
    dxSpace *s = 0;
    for (dxBody *b = w->firstbody; b && !s; b = b->next) if (b->geom && b->geom->parent_space) s = b->geom->parent_space;
    c = ctx_new(w, s);
    if (!s) { c->own_space = dSimpleSpaceCreate(0); c->space = c->own_space; }
  }
  // contact joints on the world list, creation order (the list is newest-first, ode.cpp:1162-1177)
  std::vector<dxJoint *> cj;
  bool want_fb = false;
  for (dxJoint *j = w->firstjoint; j; j = j->next) {
    if (j->feedback) want_fb = true;
    if (j->type == dJointTypeContact && j->node[0].body) cj.push_back(j);
  }
  std::reverse(cj.begin(), cj.end());
  const int nc = (int)cj.size();
  if (!ctx_ensure(c, nc, "dWorldQuickStep")) return 0;
  dxBatch *B = c->B;
  const ObBatchDev &D = B->caps;
  B->seeds[0] = ob_global_seed;
  std::vector<ObContact> hc(std::max(nc, 1));
  std::vector<ObSurface> hs(std::max(nc, 1));
  std::vector<dReal> hf((size_t)std::max(nc, 1) * 4, 0);
  for (int i = 0; i < nc; i++) {
    const dxJoint *j = cj[i];
    const dContact &ct = j->contact;
    ObContact &o = hc[i];
    memset(&o, 0, sizeof o);
    for (int e = 0; e < 3; e++) { o.pos[e] = ct.geom.pos[e]; o.normal[e] = ct.geom.normal[e]; }
    o.depth = ct.geom.depth;
    o.g1 = (ct.geom.g1 && ct.geom.g1->parent_space == c->space) ? ct.geom.g1->batch_index : -1;
    o.g2 = (ct.geom.g2 && ct.geom.g2->parent_space == c->space) ? ct.geom.g2->batch_index : -1;
    o.side1 = j->node[0].body->batch_index;                          // bodies as attached (after the swap rule)
    o.side2 = j->node[1].body ? j->node[1].body->batch_index : -1;
    o.policy = (j->flags & dJOINT_REVERSE) ? 1 : 0;
    ob_fill_surface(hs[i], ct.surface);
    for (int e = 0; e < 3; e++) hf[(size_t)4 * i + e] = ct.fdir1[e];
    if (j->node[0].body->world != w) { ob_error(0, "dWorldQuickStep: contact joint attached to a body of another world"); return 0; }
  }
  ob_joints_prestep_bookkeeping(w);
  int rc = ob_batch_upload(B);
  rc |= obk_h2d(B->bk, D.ncontacts, &nc, sizeof(int));
  if (nc) {
    rc |= obk_h2d(B->bk, D.contacts, hc.data(), sizeof(ObContact) * nc);
    rc |= obk_h2d(B->bk, D.csurf, hs.data(), sizeof(ObSurface) * nc);
    rc |= obk_h2d(B->bk, D.cfdir1, hf.data(), sizeof(dReal) * 4 * nc);
  }
  // feedback slots start as all-ones (NaN) so joints that did not enter a solved island keep the caller's values
  if (want_fb) rc |= obk_memset(B->bk, D.fback, 0xff, sizeof(dReal) * 12 * (size_t)(D.NC + D.NJ));
  if (rc) { ob_error(0, "dWorldQuickStep: upload failed"); return 0; }
  char err[512] = "";
  if (obk_run_phases(B->bk, h, OBK_PHASE_STEP, want_fb ? 1 : 0, err, sizeof err)) { ob_error(0, "dWorldQuickStep: %s", err); return 0; }
  // results: body state, stepping order (for dGeomMoved), seed, feedback
  const int nb = B->nb[0];
  std::vector<ObBodyDyn> hd(std::max(nb, 1));
  std::vector<unsigned char> ib(std::max(D.NB, 1));
  std::vector<int> si(SI_WORDS);
  ObWorld hw;
  rc = obk_d2h(B->bk, hd.data(), D.bdyn, sizeof(ObBodyDyn) * std::max(nb, 1));
  rc |= obk_d2h(B->bk, ib.data(), D.ibody, ib.size());
  rc |= obk_d2h(B->bk, si.data(), D.stepinfo, sizeof(int) * SI_WORDS);
  rc |= obk_d2h(B->bk, &hw, D.world, sizeof hw);
  if (rc) { ob_error(0, "dWorldQuickStep: download failed"); return 0; }
  if (hw.status) {
    const int st = hw.status;
    hw.status = 0; obk_h2d(B->bk, D.world, &hw, sizeof hw);
    ob_error(0, "dWorldQuickStep: capacity exceeded on the device (status %d)", st);
    return 0;
  }
  const int NA = D.NADIS;   // averaged auto-disable: the sample buffers live in the host bodies between calls
  std::vector<dReal> hb((size_t)(NA > 0 ? (size_t)nb * NA * 6 : 0));
  std::vector<int> hcw((size_t)(NA > 0 ? (size_t)nb * 2 : 0));
  if (NA > 0 && nb > 0) {
    rc = obk_d2h(B->bk, hb.data(), D.adisbuf, hb.size() * sizeof(dReal));
    rc |= obk_d2h(B->bk, hcw.data(), D.adisctl, hcw.size() * sizeof(int));
    if (rc) { ob_error(0, "dWorldQuickStep: download failed"); return 0; }
  }
  for (int i = 0; i < nb; i++) {
    dxBody *b = B->bodies[0][i];
    const ObBodyDyn &d = hd[i];
    for (int k = 0; k < 3; k++) { b->pos[k] = d.pos[k]; b->lvel[k] = d.lvel[k]; b->avel[k] = d.avel[k]; b->facc[k] = d.facc[k]; b->tacc[k] = d.tacc[k]; }
    for (int k = 0; k < 4; k++) b->q[k] = d.q[k];
    for (int k = 0; k < 12; k++) b->R[k] = d.R[k];
    b->flags = d.flags; b->adis_stepsleft = d.adis_stepsleft; b->adis_timeleft = d.adis_timeleft;
    if (NA > 0) {
      for (size_t k = 0; k < b->average_buf.size() && k < (size_t)NA * 6; k++) b->average_buf[k] = hb[(size_t)i * NA * 6 + k];
      b->average_counter = (unsigned)hcw[2 * i]; b->average_ready = hcw[2 * i + 1];
    }
  }
  // dxStepBody: every geom of a stepped body is reported moved, in stepping order (util.cpp:331-337)
  for (int i = 0; i < si[SI_NIB] && i < nb; i++)
    for (dxGeom *g = B->bodies[0][ib[i]]->geom; g; g = g->body_next) ob_geom_moved(g);
  for (int i = 0; i < si[SI_NIB] && i < nb; i++) {   // moved callbacks, stepping order (util.cpp:338-340)
    dxBody *b = B->bodies[0][ib[i]];
    if (b->moved_callback) b->moved_callback(b);
  }
  ob_global_seed = hw.seed;
  if (want_fb) {
    std::vector<dReal> fb((size_t)(D.NC + D.NJ) * 12);
    if (obk_d2h(B->bk, fb.data(), D.fback, fb.size() * sizeof(dReal))) { ob_error(0, "dWorldQuickStep: download failed"); return 0; }
    for (int i = 0; i < nc + (int)B->joints[0].size(); i++) {
      dxJoint *j = i < nc ? cj[i] : B->joints[0][i - nc];
      if (!j->feedback) continue;
      const dReal *s = &fb[(size_t)(i < nc ? i : D.NC + (i - nc)) * 12];
      if (s[0] != s[0] && s[11] != s[11]) continue;   // untouched slot
      for (int e = 0; e < 3; e++) { j->feedback->f1[e] = s[e]; j->feedback->t1[e] = s[3 + e]; j->feedback->f2[e] = s[6 + e]; j->feedback->t2[e] = s[9 + e]; }
    }
  }
  return 1;
}

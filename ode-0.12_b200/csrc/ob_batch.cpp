// ob_batch.cpp — the added batched-world entry points (dBatch*, include/ode_b200/ode.h):
// marshals worlds built through the ODE C API into the device layout of ob_types.h,
// and exposes bulk I/O, counters and parity taps.  No physics here.
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "ob_backend.h"
#include "ob_host.h"
#include "ob_batch.h"
#include "ob_trimesh_host.h"


void ob_fill_surface(ObSurface &d, const dSurfaceParameters &s) {
  d.mode = s.mode; d.mu = s.mu; d.mu2 = s.mu2; d.bounce = s.bounce; d.bounce_vel = s.bounce_vel;
  d.soft_erp = s.soft_erp; d.soft_cfm = s.soft_cfm; d.motion1 = s.motion1; d.motion2 = s.motion2;
  d.motionN = s.motionN; d.slip1 = s.slip1; d.slip2 = s.slip2;
}

void ob_marshal_body(const dxBody *b, ObBodyDyn &d, ObBodyConst &c) {
  memset(&d, 0, sizeof d);
  memset(&c, 0, sizeof c);
  for (int k = 0; k < 3; k++) { d.pos[k] = b->pos[k]; d.lvel[k] = b->lvel[k]; d.avel[k] = b->avel[k]; d.facc[k] = b->facc[k]; d.tacc[k] = b->tacc[k]; }
  for (int k = 0; k < 4; k++) d.q[k] = b->q[k];
  for (int k = 0; k < 12; k++) d.R[k] = b->R[k];
  d.flags = b->flags;
  d.adis_stepsleft = b->adis_stepsleft;
  d.adis_timeleft = b->adis_timeleft;
  c.mass = b->mass.mass; c.invMass = b->invMass; c.max_angular_speed = b->max_angular_speed;
  for (int k = 0; k < 12; k++) { c.I[k] = b->mass.I[k]; c.invI[k] = b->invI[k]; }
  for (int k = 0; k < 3; k++) c.finite_rot_axis[k] = b->finite_rot_axis[k];
  c.damp_lin_scale = b->dampingp.linear_scale; c.damp_ang_scale = b->dampingp.angular_scale;
  c.damp_lin_thr = b->dampingp.linear_threshold; c.damp_ang_thr = b->dampingp.angular_threshold;
  c.adis_lin_thr = b->adis.linear_average_threshold; c.adis_ang_thr = b->adis.angular_average_threshold;
  c.adis_idle_time = b->adis.idle_time; c.adis_idle_steps = b->adis.idle_steps; c.adis_samples = (int)b->adis.average_samples;
  c.geom_first = -1;
}

static void fill_limot(ObLimot &d, const dxLimot &l) {
  d.vel = l.vel; d.fmax = l.fmax; d.lostop = l.lostop; d.histop = l.histop; d.fudge_factor = l.fudge_factor;
  d.normal_cfm = l.normal_cfm; d.stop_erp = l.stop_erp; d.stop_cfm = l.stop_cfm; d.bounce = l.bounce;
  d.limit = l.limit; d.limit_err = l.limit_err; d.pad = 0;
}
void ob_marshal_joint(const dxJoint *j, ObJoint &d) {
  memset(&d, 0, sizeof d);
  d.type = j->type;
  d.b1 = j->node[0].body ? j->node[0].body->batch_index : -1;
  d.b2 = j->node[1].body ? j->node[1].body->batch_index : -1;
  d.flags = ((j->flags & dJOINT_DISABLED) ? OB_JF_DISABLED : 0) | ((j->flags & dJOINT_REVERSE) ? OB_JF_REVERSE : 0);
  for (int k = 0; k < 4; k++) {
    d.anchor1[k] = j->anchor1[k]; d.anchor2[k] = j->anchor2[k]; d.axis1[k] = j->axis1[k]; d.axis2[k] = j->axis2[k];
    d.qrel[k] = j->qrel[k]; d.v1[k] = j->v1[k]; d.v2[k] = j->v2[k];
  }
  if (j->type == dJointTypePU) for (int k = 0; k < 4; k++) { d.v1[k] = j->qrel2[k]; d.v2[k] = j->axis3[k]; }   // universal part as below; v2 = axisP1
  if (j->type == dJointTypeUniversal) for (int k = 0; k < 4; k++) d.v1[k] = j->qrel2[k];   // ObJoint::v1 doubles as qrel2
  if (j->type == dJointTypePR) for (int k = 0; k < 4; k++) { d.anchor1[k] = k < 3 ? j->offset[k] : 0; d.v1[k] = j->axis3[k]; }   // offset -> anchor1, axisP1 -> v1
  if (j->type == dJointTypeSlider || j->type == dJointTypeFixed) for (int k = 0; k < 3; k++) d.anchor1[k] = j->offset[k];   // ObJoint::anchor1 doubles as the offset
  d.erp = j->erp; d.cfm = j->cfm; d.susp_erp = j->susp_erp; d.susp_cfm = j->susp_cfm; d.c0 = j->c0; d.s0 = j->s0;
  fill_limot(d.limot1, j->limot);
  fill_limot(d.limot2, j->limot2);
  fill_limot(d.limot3, j->limot3);
  if (j->type == dJointTypeAMotor || j->type == dJointTypeLMotor) {
    d.flags |= (j->num << 8) | ((j->mode == dAMotorEuler ? 1 : 0) << 12) | (j->rel[0] << 16) | (j->rel[1] << 20) | (j->rel[2] << 24);
    for (int k = 0; k < 3; k++) { d.anchor1[k] = j->axis3[k]; d.anchor2[k] = j->reference1[k]; d.v1[k] = j->reference2[k]; d.qrel[k] = j->angle[k]; }
    d.qrel[3] = 0;
  }
}

void ob_marshal_geom(dxGeom *g, ObGeom &d) {
  memset(&d, 0, sizeof d);
  if (g->is_space) {
    // a sub-space is a member like any geom: enabled flag, category / collide bits and the union box of its members; the near
    // callback decides what to do with the pair (dSpaceCollide2 / dCollide on the space), there is no collider for it
    d.type = OB_GEOM_SPACE; d.body = -1; d.body_next = -1;
    d.cat = (uint32_t)g->category_bits; d.col = (uint32_t)g->collide_bits;
    d.flags = (g->gflags & GEOM_ENABLED) ? OB_GEOM_ENABLED : 0;
    dReal a[6];
    dGeomGetAABB(g, a);
    for (int k = 0; k < 6; k++) d.R[k] = a[k];
    return;
  }
  // a geom transform is uploaded as its encapsulated geom: class, parameters and zero-size flag of the inner geom,
  // everything else (body, bits, enable flag) of the transform; the inner geom's own pose is the "offset"
  // (computeFinalTx, collision_transform.cpp:101-108, is the same arithmetic as computePosr)
  dxGeom *sh = ob_geom_shape(g);
  const bool xf = g->type == dGeomTransformClass;
  if (xf) {
    if (!sh) { ob_error(0, "geom transform without an encapsulated geom is not served on this path"); sh = g; }
    else if (sh->type != dSphereClass && sh->type != dBoxClass && sh->type != dCapsuleClass && sh->type != dCylinderClass)
      ob_error(0, "geom transform: encapsulated geom class %d is not served on this path (sphere, box, capsule, cylinder are)", sh->type);
    else if (g->offset_posr) ob_error(0, "geom transform with its own offset is not served on this path");
    if (sh->parent_space || sh->body) ob_debug(2, "GeomTransform encapsulated object must not be in a space or attached to a body");
  }
  d.type = sh->type;
  d.body = g->body ? g->body->batch_index : -1;
  d.cat = (uint32_t)g->category_bits; d.col = (uint32_t)g->collide_bits;
  d.flags = ((g->gflags & GEOM_ENABLED) ? OB_GEOM_ENABLED : 0) | ((g->offset_posr || (xf && g->body)) ? OB_GEOM_HAS_OFFSET : 0) |
            ((sh->gflags & GEOM_ZERO_SIZED) ? OB_GEOM_ZERO_SIZED : 0);
  d.body_next = -1;
  if (g->type == dRayClass) d.mesh = ob_ray_flags(g);   // rays carry their mode bits where trimeshes carry the data index
  if (xf) d.mesh = OB_POSE_XFORM;
  for (int k = 0; k < 4; k++) d.p[k] = sh->p[k];
  const dxPosR *src = 0;
  dxPosR composed;
  if (xf && g->body) src = sh->final_posr;
  else if (xf) { ob_geom_final_pose(g, &composed); src = &composed; }
  else if (g->offset_posr) src = g->offset_posr;
  else if (!g->body && (g->gflags & GEOM_PLACEABLE)) src = g->final_posr;
  if (src) { for (int k = 0; k < 3; k++) d.pos[k] = src->pos[k]; for (int k = 0; k < 12; k++) d.R[k] = src->R[k]; }
  else { d.R[0] = d.R[5] = d.R[10] = 1; }
}

// Re-marshal the bound worlds into the device layout and upload (everything except the policy
// table, seeds and per-step scratch).  Contact joints on the world's joint list are skipped: on
// the batched path they do not exist, on the drop-in path they are uploaded per step.
static int ob_batch_upload_check_large(dxBatch *) { return 0; }

int ob_batch_upload(dxBatch *B) {
  const int nworlds = B->caps.W, NB = B->caps.NB, NG = B->caps.NG, NJ = B->caps.NJ;
  std::vector<ObWorld> hw(nworlds);
  std::vector<ObBodyDyn> hd((size_t)nworlds * NB);
  std::vector<ObBodyConst> hc((size_t)nworlds * NB);
  std::vector<ObGeom> hg((size_t)nworlds * NG);
  std::vector<int> hl((size_t)nworlds * NG, -1);
  memset(hd.data(), 0, hd.size() * sizeof(ObBodyDyn));
  memset(hc.data(), 0, hc.size() * sizeof(ObBodyConst));
  memset(hg.data(), 0, hg.size() * sizeof(ObGeom));
  int any_xf = 0;
  for (int w = 0; w < nworlds; w++) {
    dxWorld *W = B->worlds[w]; dxSpace *S = B->spaces[w];
    ObWorld &o = hw[w];
    memset(&o, 0, sizeof o);
    for (int k = 0; k < 3; k++) o.gravity[k] = W->gravity[k];
    o.erp = W->global_erp; o.cfm = W->global_cfm; o.sor_w = W->qs_w; o.max_vel = W->contact_max_vel;
    o.min_depth = W->contact_min_depth; o.iters = W->qs_iterations; o.nb = B->nb[w]; o.ng = B->ng[w];
    o.seed = B->seeds[w]; o.hash_minlevel = S->minlevel; o.hash_maxlevel = S->maxlevel;
    o.space_type = S->type == dHashSpaceClass ? OB_SPACE_HASH : (S->type == dSweepAndPruneSpaceClass ? OB_SPACE_SAP : OB_SPACE_SIMPLE);
    for (int i = 0; i < B->nb[w]; i++) {
      dxBody *b = B->bodies[w][i];
      ob_marshal_body(b, hd[(size_t)w * NB + i], hc[(size_t)w * NB + i]);
      hc[(size_t)w * NB + i].geom_first = (b->geom && b->geom->parent_space == S) ? b->geom->batch_index : -1;
    }
    int pos = 0;
    for (dxGeom *g = S->first; g; g = g->next, pos++) {
      ObGeom &d = hg[(size_t)w * NG + g->batch_index];
      ob_marshal_geom(g, d);
      if (g->type == dGeomTransformClass) any_xf = 1;
      if (g->type == dTriMeshClass) {
        d.mesh = -1;
        for (size_t mi = 0; mi < B->meshes.size(); mi++) if (B->meshes[mi] == g->tmdata) d.mesh = (int)mi;
        if (d.mesh < 0) d.flags &= ~OB_GEOM_ENABLED;   // data changed after binding: the drop-in layer rebinds, see batch_matches
      }
      d.body_next = (g->body_next && g->body_next->parent_space == S) ? g->body_next->batch_index : -1;
      hl[(size_t)w * NG + pos] = g->batch_index;
    }
    if (S->type == dSweepAndPruneSpaceClass) {   // order state of a SAP space: DirtyList then GeomList
      pos = 0;
      for (size_t i = 0; i < S->sap_dirty.size(); i++) hl[(size_t)w * NG + pos++] = S->sap_dirty[i]->batch_index;
      for (size_t i = 0; i < S->sap_geoms.size(); i++) hl[(size_t)w * NG + pos++] = S->sap_geoms[i]->batch_index;
      o.sap_ndirty = (int)S->sap_dirty.size();
      o.sap_axes = S->axisorder;
    }
  }
  std::vector<ObJoint> hj((size_t)nworlds * std::max(NJ, 1));
  std::vector<int> hnj(nworlds, 0);
  std::vector<unsigned short> hps((size_t)nworlds * (NB + 1), 0), hpa((size_t)nworlds * 2 * std::max(NJ, 1), 0);
  memset(hj.data(), 0, hj.size() * sizeof(ObJoint));
  for (int w = 0; w < nworlds; w++) {
    const std::vector<dxJoint *> &J = B->joints[w];
    hnj[w] = (int)J.size();
    for (size_t i = 0; i < J.size(); i++) { ob_marshal_joint(J[i], hj[(size_t)w * NJ + i]); J[i]->tag = (int)i; }
    unsigned short *ps = &hps[(size_t)w * (NB + 1)], *pa = &hpa[(size_t)w * 2 * std::max(NJ, 1)];
    int a = 0;
    for (int b = 0; b < B->nb[w]; b++) {
      ps[b] = (unsigned short)a;
      for (dxJointNode *n = B->bodies[w][b]->firstjoint; n; n = n->next)
        if (n->joint->type != dJointTypeContact) pa[a++] = (unsigned short)n->joint->tag;
    }
    for (int b = B->nb[w]; b <= NB; b++) ps[b] = (unsigned short)a;
  }
  int rc = 0;
  B->caps.any_xf = any_xf; obk_arrays(B->bk)->any_xf = any_xf;   // kernel parameter, not device memory
  rc |= obk_h2d(B->bk, B->caps.world, hw.data(), hw.size() * sizeof(ObWorld));
  rc |= obk_h2d(B->bk, B->caps.bdyn, hd.data(), hd.size() * sizeof(ObBodyDyn));
  rc |= obk_h2d(B->bk, B->caps.bconst, hc.data(), hc.size() * sizeof(ObBodyConst));
  rc |= obk_h2d(B->bk, B->caps.geom, hg.data(), hg.size() * sizeof(ObGeom));
  rc |= obk_h2d(B->bk, B->caps.glist, hl.data(), hl.size() * sizeof(int));
  if (B->caps.NADIS > 0) {   // averaged auto-disable: sample ring buffers + (write index, full flag) per body
    const int NA = B->caps.NADIS;
    std::vector<dReal> hb((size_t)nworlds * NB * NA * 6, 0);
    std::vector<int> hcw((size_t)nworlds * NB * 2, 0);
    for (int w = 0; w < nworlds; w++)
      for (int i = 0; i < B->nb[w]; i++) {
        const dxBody *b = B->bodies[w][i];
        const size_t bi = (size_t)w * NB + i;
        if ((int)b->adis.average_samples > NA) { ob_set_last_error("a body's auto-disable sample count (%u) was raised above the bound batch's buffer depth (%d): re-create the batch", b->adis.average_samples, NA); return -1; }
        for (size_t k = 0; k < b->average_buf.size(); k++) hb[bi * NA * 6 + k] = b->average_buf[k];
        hcw[bi * 2] = (int)b->average_counter; hcw[bi * 2 + 1] = b->average_ready;
      }
    rc |= obk_h2d(B->bk, B->caps.adisbuf, hb.data(), hb.size() * sizeof(dReal));
    rc |= obk_h2d(B->bk, B->caps.adisctl, hcw.data(), hcw.size() * sizeof(int));
  }
  if (B->caps.large) return rc;   // no permanent joints on the large-world path
  if (NJ) rc |= obk_h2d(B->bk, B->caps.joint, hj.data(), hj.size() * sizeof(ObJoint));
  rc |= obk_h2d(B->bk, B->caps.njoints, hnj.data(), hnj.size() * sizeof(int));
  rc |= obk_h2d(B->bk, B->caps.padjstart, hps.data(), hps.size() * sizeof(unsigned short));
  if (NJ) rc |= obk_h2d(B->bk, B->caps.padj, hpa.data(), hpa.size() * sizeof(unsigned short));
  return rc;
}

// dropin: the batch serves the classic per-call API (ob_dropin.cpp): per-contact surface arrays are
// allocated, the worlds are not marked as owned by a user batch, contact joints present at bind time are ignored
dxBatch *ob_batch_create(int nworlds, const dWorldID *worlds, const dSpaceID *spaces, const dBatchDesc *desc, int dropin) {
  if (nworlds <= 0 || !worlds || !spaces) { ob_set_last_error("dBatchCreate: bad arguments"); return 0; }
  dxBatch *B = new dxBatch;
  B->bk = 0; B->debug_taps = 1; B->dropin = dropin; B->invalid = 0;
  B->worlds.assign(worlds, worlds + nworlds);
  B->spaces.assign(spaces, spaces + nworlds);
  B->bodies.resize(nworlds); B->geoms.resize(nworlds); B->nb.resize(nworlds); B->ng.resize(nworlds);
  B->seeds.assign(nworlds, 0);
  int NB = 1, NG = 1, NJ = 0;
  B->joints.resize(nworlds);
  for (int w = 0; w < nworlds; w++) {
    dxWorld *W = worlds[w]; dxSpace *S = spaces[w];
    if (!W || !S || !S->is_space) { ob_set_last_error("dBatchCreate: world %d: bad world/space", w); delete B; return 0; }
    if (S->type != dHashSpaceClass && S->type != dSimpleSpaceClass && S->type != dSweepAndPruneSpaceClass) {
      ob_set_last_error("dBatchCreate: world %d: unsupported space class %d", w, S->type); delete B; return 0;
    }
    int i = 0;
    for (dxBody *b = W->firstbody; b; b = b->next) { b->batch_index = i++; B->bodies[w].push_back(b); }
    B->nb[w] = i;
    int ng = S->count;
    B->geoms[w].resize(ng);
    int pos = 0;
    for (dxGeom *g = S->first; g; g = g->next, pos++) {
      if (g->is_space && !dropin) { ob_set_last_error("dBatchCreate: world %d: nested spaces are served by the classic API (dSpaceCollide / dSpaceCollide2 / dCollide), not by a bound batch", w); delete B; return 0; }
      if (g->body && g->body->world != W) { ob_set_last_error("dBatchCreate: world %d: geom attached to a body of another world", w); delete B; return 0; }
      if (g->type == dRayClass && !dropin) { ob_set_last_error("dBatchCreate: world %d: ray geoms are served by dSpaceCollide / dCollide (a ray contact is a query result, not a contact joint)", w); delete B; return 0; }
      g->batch_index = ng - 1 - pos;
      B->geoms[w][g->batch_index] = g;
    }
    B->ng[w] = ng;
    // permanent joints: everything on the world's joint list at bind time, creation order
    for (dxJoint *j = W->firstjoint; j; j = j->next) {
      if (j->type == dJointTypeContact) {
        if (dropin) continue;
        ob_set_last_error("dBatchCreate: world %d: contact joints present at bind time (call dJointGroupEmpty first)", w); delete B; return 0;
      }
      if (j->type != dJointTypeBall && j->type != dJointTypeHinge && j->type != dJointTypeHinge2 && j->type != dJointTypeSlider && j->type != dJointTypeFixed && j->type != dJointTypeUniversal && j->type != dJointTypeAMotor && j->type != dJointTypeLMotor && j->type != dJointTypePlane2D && j->type != dJointTypePiston && j->type != dJointTypePR && j->type != dJointTypePU && j->type != dJointTypeNull) { ob_set_last_error("dBatchCreate: world %d: unsupported joint type %d", w, j->type); delete B; return 0; }
      // plane2d constrains body 1 against the static environment (plane2d.cpp:95-118 never fills J2); a second body is refused
      if (j->type == dJointTypePlane2D && j->node[1].body) { ob_set_last_error("dBatchCreate: world %d: a plane2d joint takes one body", w); delete B; return 0; }
      B->joints[w].push_back(j);
    }
    std::reverse(B->joints[w].begin(), B->joints[w].end());
    NB = std::max(NB, B->nb[w]); NG = std::max(NG, ng); NJ = std::max(NJ, (int)B->joints[w].size());
  }
  for (int w = 0; w < nworlds; w++)
    for (dxGeom *g = spaces[w]->first; g; g = g->next)
      if (g->type == dTriMeshClass) {
        if (!g->tmdata || g->tmdata->ntris < 2) { ob_set_last_error("dBatchCreate: world %d: trimesh geom without built data", w); delete B; return 0; }
        bool have = false;
        for (size_t mi = 0; mi < B->meshes.size(); mi++) have |= B->meshes[mi] == g->tmdata;
        if (!have) B->meshes.push_back(g->tmdata);
      }
  ObBatchDev caps;
  memset(&caps, 0, sizeof caps);
  caps.nmesh = (int)B->meshes.size();
  caps.W = nworlds; caps.NB = NB; caps.NG = NG;
  caps.NC = (desc && desc->max_contacts_per_world > 0) ? desc->max_contacts_per_world : std::max(64, 16 * NG);
  caps.NP = (desc && desc->max_pairs_per_world > 0) ? desc->max_pairs_per_world : (int)std::min((long long)NG * (NG - 1) / 2 + 1, (long long)std::max(256, 12 * NG));
  caps.large = (!dropin && nworlds == 1 && ((desc && desc->large_world) || NB > 254 || NG > 255)) ? 1 : 0;
  if (caps.large) {
    // the grid-wide path (ob_large.h): contact joints only, SAP space, every body enabled
    if (NJ) { ob_set_last_error("dBatchCreate: the large-world path supports contact joints only (%d permanent joints bound)", NJ); delete B; return 0; }
    if (spaces[0]->type != dSweepAndPruneSpaceClass) { ob_set_last_error("dBatchCreate: the large-world path implements dSweepAndPruneSpace only"); delete B; return 0; }
    for (int i = 0; i < B->nb[0]; i++)
      if (B->bodies[0][i]->flags & (OB_BODY_DISABLED | OB_BODY_AUTO_DISABLE)) { ob_set_last_error("dBatchCreate: the large-world path does not support disabled / auto-disabling bodies"); delete B; return 0; }
  }
  caps.NJ = NJ;
  caps.NADIS = 0;   // deepest auto-disable sample buffer among the bound bodies (only averaging bodies, average_samples > 1, need one)
  for (int w = 0; w < nworlds; w++)
    for (int i = 0; i < B->nb[w]; i++) { const int n = (int)B->bodies[w][i]->adis.average_samples; if (n > 1 && n > caps.NADIS) caps.NADIS = n; }
  caps.NR = 3 * caps.NC + 6 * NJ;
  caps.npolicy = OB_MAXPOLICY;
  caps.dropin = dropin;
  {
    int it = 1;
    for (int w = 0; w < nworlds; w++) it = std::max(it, worlds[w]->qs_iterations);
    caps.NEP = (it + 7) / 8;
  }
  char err[512] = "";
  B->bk = obk_create(caps, desc ? desc->device : 0, err, sizeof err);
  if (!B->bk) { ob_set_last_error("dBatchCreate: %s", err); delete B; return 0; }
  B->caps = *obk_arrays(B->bk);
  if (!dropin)
    for (int w = 0; w < nworlds; w++) { worlds[w]->bound_batch = B; spaces[w]->bound_batch = B; }
  ObPolicy pol;
  memset(&pol, 0, sizeof pol);
  pol.cat_mask1 = pol.cat_mask2 = ~0u; pol.max_contacts = 8; pol.skip_if_connected = 1;
  pol.surface.mode = 0; pol.surface.mu = OB_INF;
  int rc = 0;
  if (!B->meshes.empty()) {
    std::vector<ObMeshDev> mt(B->meshes.size());
    for (size_t mi = 0; mi < mt.size(); mi++) {
      const ObMeshDev *md = ob_trimesh_device(B->meshes[mi], desc ? desc->device : 0);
      if (!md) { ob_set_last_error("dBatchCreate: trimesh upload failed"); dBatchDestroy(B); return 0; }
      mt[mi] = *md;
    }
    rc |= obk_h2d(B->bk, B->caps.meshes, mt.data(), mt.size() * sizeof(ObMeshDev));
  }
  rc |= ob_batch_upload(B);
  rc |= obk_h2d(B->bk, B->caps.policy, &pol, sizeof pol);
  rc |= obk_memset(B->bk, B->caps.counters, 0, sizeof(ObCounters));
  rc |= obk_memset(B->bk, B->caps.npairs, 0, sizeof(int) * nworlds);
  rc |= obk_memset(B->bk, B->caps.ncontacts, 0, sizeof(int) * nworlds);
  rc |= obk_memset(B->bk, B->caps.nrows, 0, sizeof(int) * nworlds);
  if (rc) { ob_set_last_error("dBatchCreate: upload failed"); dBatchDestroy(B); return 0; }
  if (caps.large && ob_batch_upload_check_large(B)) { dBatchDestroy(B); return 0; }
  return B;
}

void ob_batch_invalidate(dxBatch *B, dxWorld *gone_world, dxSpace *gone_space) {
  if (!B || B->dropin) return;
  B->invalid = 1;
  for (size_t w = 0; w < B->worlds.size(); w++) {
    if (gone_world && B->worlds[w] == gone_world) B->worlds[w] = 0;
    if (gone_space && B->spaces[w] == gone_space) B->spaces[w] = 0;
  }
}
static bool batch_usable(dxBatch *B, const char *who) {
  if (!B) { ob_set_last_error("%s: null batch", who); return false; }
  if (B->invalid) { ob_set_last_error("%s: a world, space, body or geom bound into this batch was destroyed or added / removed after dBatchCreate; destroy the batch and create it again", who); return false; }
  return true;
}

extern "C" {

dBatchID dBatchCreate(int nworlds, const dWorldID *worlds, const dSpaceID *spaces, const dBatchDesc *desc) {
  return ob_batch_create(nworlds, worlds, spaces, desc, 0);
}

void dBatchDestroy(dBatchID B) {
  if (!B) return;
  for (size_t w = 0; w < B->worlds.size() && !B->dropin; w++) {
    if (B->worlds[w] && B->worlds[w]->bound_batch == B) B->worlds[w]->bound_batch = 0;
    if (B->spaces[w] && B->spaces[w]->bound_batch == B) B->spaces[w]->bound_batch = 0;
  }
  if (B->bk) obk_destroy(B->bk);
  delete B;
}

int dBatchSetContactPolicy(dBatchID B, const dBatchContactPolicy *table, int n) {
  if (!B || !table || n < 1 || n > OB_MAXPOLICY) { ob_set_last_error("dBatchSetContactPolicy: 1 to %d policy rows", OB_MAXPOLICY); return -1; }
  if (n > 1 && B->caps.large) { ob_set_last_error("dBatchSetContactPolicy: the large-world path takes one policy row"); return -1; }
  ObPolicy pol[OB_MAXPOLICY];
  memset(pol, 0, sizeof pol);
  for (int r = 0; r < n; r++) {
    pol[r].cat_mask1 = (uint32_t)table[r].cat_mask1; pol[r].cat_mask2 = (uint32_t)table[r].cat_mask2;
    pol[r].max_contacts = table[r].max_contacts; pol[r].skip_if_connected = table[r].skip_if_connected;
    pol[r].skip_static_pairs = table[r].skip_static_pairs;
    ob_fill_surface(pol[r].surface, table[r].surface);
  }
  pol[0].nrows = n;
  return obk_h2d(B->bk, B->caps.policy, pol, sizeof(ObPolicy) * (B->caps.npolicy < OB_MAXPOLICY ? 1 : OB_MAXPOLICY));
}

int dBatchSetSeeds(dBatchID B, const uint32_t *seeds) {
  int W = B->caps.W;
  std::vector<ObWorld> hw(W);
  if (obk_d2h(B->bk, hw.data(), B->caps.world, W * sizeof(ObWorld))) return -1;
  for (int w = 0; w < W; w++) { hw[w].seed = seeds[w]; B->seeds[w] = seeds[w]; }
  return obk_h2d(B->bk, B->caps.world, hw.data(), W * sizeof(ObWorld));
}
int dBatchGetSeeds(dBatchID B, uint32_t *seeds) {
  int W = B->caps.W;
  std::vector<ObWorld> hw(W);
  if (obk_d2h(B->bk, hw.data(), B->caps.world, W * sizeof(ObWorld))) return -1;
  for (int w = 0; w < W; w++) seeds[w] = hw[w].seed;
  return 0;
}

int dBatchCollideAndQuickStep(dBatchID B, dReal h, int nsteps, int *status_per_world) {
  if (!B || !(h > 0) || nsteps < 0) { ob_set_last_error("dBatchCollideAndQuickStep: bad arguments"); return -1; }
  if (!batch_usable(B, "dBatchCollideAndQuickStep")) return -1;
  char err[512] = "";
  int rc = obk_step(B->bk, h, nsteps, B->debug_taps, err, sizeof err);
  if (rc) { ob_set_last_error("dBatchCollideAndQuickStep: %s", err); return rc; }
  if (status_per_world) {
    int W = B->caps.W;
    std::vector<ObWorld> hw(W);
    if (obk_d2h(B->bk, hw.data(), B->caps.world, W * sizeof(ObWorld))) return -1;
    bool any = false;
    for (int w = 0; w < W; w++) { status_per_world[w] = hw[w].status; any |= hw[w].status != 0; hw[w].status = 0; }
    if (any) obk_h2d(B->bk, B->caps.world, hw.data(), W * sizeof(ObWorld));
  }
  return 0;
}

int dBatchNumBodies(dBatchID B) { return B->caps.NB; }
int dBatchGetBodyState(dBatchID B, dReal *pos3, dReal *quat4, dReal *lvel3, dReal *avel3) {
  return obk_get_state(B->bk, pos3, quat4, lvel3, avel3);
}
int dBatchSetBodyState(dBatchID B, const dReal *pos3, const dReal *quat4, const dReal *lvel3, const dReal *avel3) {
  return obk_set_state(B->bk, pos3, quat4, lvel3, avel3);
}
int dBatchAddForces(dBatchID B, const dReal *force3, const dReal *torque3) { return obk_add_forces(B->bk, force3, torque3); }
void *dBatchHostAlloc(size_t bytes) { return obk_host_alloc(bytes); }
void dBatchHostFree(void *p) { obk_host_free(p); }

int dBatchDownload(dBatchID B) {
  if (!batch_usable(B, "dBatchDownload")) return -1;
  int W = B->caps.W, NB = B->caps.NB, NG = B->caps.NG;
  std::vector<ObBodyDyn> hd((size_t)W * NB);
  std::vector<int> hl((size_t)W * NG);
  std::vector<ObWorld> hw(W);
  if (obk_d2h(B->bk, hd.data(), B->caps.bdyn, hd.size() * sizeof(ObBodyDyn))) return -1;
  if (obk_d2h(B->bk, hl.data(), B->caps.glist, hl.size() * sizeof(int))) return -1;
  if (obk_d2h(B->bk, hw.data(), B->caps.world, W * sizeof(ObWorld))) return -1;
  const int NA = B->caps.NADIS;
  std::vector<dReal> hb((size_t)(NA > 0 ? (size_t)W * NB * NA * 6 : 0));
  std::vector<int> hcw((size_t)(NA > 0 ? (size_t)W * NB * 2 : 0));
  if (NA > 0) {
    if (obk_d2h(B->bk, hb.data(), B->caps.adisbuf, hb.size() * sizeof(dReal))) return -1;
    if (obk_d2h(B->bk, hcw.data(), B->caps.adisctl, hcw.size() * sizeof(int))) return -1;
  }
  for (int w = 0; w < W; w++) {
    for (int i = 0; i < B->nb[w]; i++) {
      dxBody *b = B->bodies[w][i];
      const ObBodyDyn &d = hd[(size_t)w * NB + i];
      if (NA > 0) {
        const size_t bi = (size_t)w * NB + i;
        for (size_t k = 0; k < b->average_buf.size() && k < (size_t)NA * 6; k++) b->average_buf[k] = hb[bi * NA * 6 + k];
        b->average_counter = (unsigned)hcw[bi * 2]; b->average_ready = hcw[bi * 2 + 1];
      }
      for (int k = 0; k < 3; k++) { b->pos[k] = d.pos[k]; b->lvel[k] = d.lvel[k]; b->avel[k] = d.avel[k]; b->facc[k] = d.facc[k]; b->tacc[k] = d.tacc[k]; }
      for (int k = 0; k < 4; k++) b->q[k] = d.q[k];
      for (int k = 0; k < 12; k++) b->R[k] = d.R[k];
      b->flags = d.flags; b->adis_stepsleft = d.adis_stepsleft; b->adis_timeleft = d.adis_timeleft;
    }
    // rebuild the space's linked list in device order (all geoms clean after a step's collide,
    // dirty again after its integration: mark the body geoms dirty like dGeomMoved does)
    dxSpace *S = B->spaces[w];
    int ng = B->ng[w];
    if (S->type == dSweepAndPruneSpaceClass) {
      const int nd = hw[w].sap_ndirty;
      S->sap_dirty.clear(); S->sap_geoms.clear();
      for (int pos = 0; pos < ng; pos++) {
        dxGeom *g = B->geoms[w][hl[(size_t)w * NG + pos]];
        if (pos < nd) { g->sap_didx = pos; g->sap_gidx = -1; g->gflags |= GEOM_DIRTY | GEOM_AABB_BAD; S->sap_dirty.push_back(g); }
        else { g->sap_didx = -1; g->sap_gidx = pos - nd; g->gflags &= ~(GEOM_DIRTY | GEOM_AABB_BAD); S->sap_geoms.push_back(g); }
      }
      continue;
    }
    S->first = 0;
    dxGeom **link = &S->first;
    for (int pos = 0; pos < ng; pos++) {
      dxGeom *g = B->geoms[w][hl[(size_t)w * NG + pos]];
      *link = g; g->tome = link; g->next = 0; link = &g->next;
    }
  }
  return 0;
}

int dBatchGetCounters(dBatchID B, dBatchCounters *out) {
  ObCounters c;
  if (obk_d2h(B->bk, &c, B->caps.counters, sizeof c)) return -1;
  out->steps = (long long)c.steps; out->body_steps = (long long)c.body_steps; out->pairs = (long long)c.pairs;
  out->contacts = (long long)c.contacts; out->rows = (long long)c.rows; out->islands = (long long)c.islands;
  out->overflow_worlds = (long long)c.overflow_worlds;
  return 0;
}
int dBatchResetCounters(dBatchID B) { return obk_memset(B->bk, B->caps.counters, 0, sizeof(ObCounters)); }

int dBatchDebugPairs(dBatchID B, int w, int *g1g2, int cap) {
  int n = 0;
  if (obk_d2h(B->bk, &n, B->caps.npairs + w, sizeof(int))) return -1;
  int m = std::min(n, cap);
  if (m > 0 && obk_d2h(B->bk, g1g2, B->caps.pairs + (size_t)w * B->caps.NP * 2, sizeof(int) * 2 * m)) return -1;
  return n;
}
int dBatchDebugContacts(dBatchID B, int w, dReal *pnd7, int *g1g2, int cap) {
  int n = 0;
  if (obk_d2h(B->bk, &n, B->caps.ncontacts + w, sizeof(int))) return -1;
  int m = std::min(n, cap);
  if (m <= 0) return n;
  std::vector<ObContact> c(m);
  if (obk_d2h(B->bk, c.data(), B->caps.contacts + (size_t)w * B->caps.NC, sizeof(ObContact) * m)) return -1;
  for (int i = 0; i < m; i++) {
    for (int k = 0; k < 3; k++) { pnd7[7 * i + k] = c[i].pos[k]; pnd7[7 * i + 3 + k] = c[i].normal[k]; }
    pnd7[7 * i + 6] = c[i].depth;
    g1g2[2 * i] = c[i].g1; g1g2[2 * i + 1] = c[i].g2;
  }
  return n;
}
int dBatchDebugLambda(dBatchID B, int w, dReal *lambda, int cap) {
  int n = 0;
  if (B->caps.large) return 0;
  if (obk_d2h(B->bk, &n, B->caps.nrows + w, sizeof(int))) return -1;
  int m = std::min(n, cap);
  if (m > 0 && obk_d2h(B->bk, lambda, B->caps.lambda + (size_t)w * B->caps.NR, sizeof(dReal) * m)) return -1;
  return n;
}
int dBatchDebugFeedback(dBatchID B, int w, dReal *f1t1, int cap) {
  int n = 0;
  if (obk_d2h(B->bk, &n, B->caps.ncontacts + w, sizeof(int))) return -1;
  int m = std::min(n, cap);
  if (m <= 0) return n;
  std::vector<dReal> fb((size_t)m * 12);
  if (obk_d2h(B->bk, fb.data(), B->caps.fback + (size_t)w * (B->caps.NC + B->caps.NJ) * 12, sizeof(dReal) * 12 * m)) return -1;
  for (int i = 0; i < m; i++) for (int k = 0; k < 6; k++) f1t1[6 * i + k] = fb[(size_t)12 * i + k];
  return n;
}
int dBatchDebugGeomOrder(dBatchID B, int w, int *order, int cap) {
  int n = B->ng[w];
  int m = std::min(n, cap);
  if (m > 0 && obk_d2h(B->bk, order, B->caps.glist + (size_t)w * B->caps.NG, sizeof(int) * m)) return -1;
  return n;
}
// Order state: what the next step's callback / row order depends on besides body state and seeds — the space
// list order (rewritten every step like dGeomMoved does), the SAP space's RadixSort ranks and dirty count.
// An opaque int blob for snapshots: [W*NG glist][W*(NG+3) sapstate][W sap_ndirty].
int dBatchOrderStateSize(dBatchID B) {
  const ObBatchDev &D = B->caps;
  return D.W * D.NG + (D.sapstate ? D.W * (D.NG + 3) : 0) + D.W;
}
int dBatchGetOrderState(dBatchID B, int *buf) {
  const ObBatchDev &D = B->caps;
  size_t o = 0;
  if (obk_d2h(B->bk, buf + o, D.glist, sizeof(int) * (size_t)D.W * D.NG)) return -1;
  o += (size_t)D.W * D.NG;
  if (D.sapstate) { if (obk_d2h(B->bk, buf + o, D.sapstate, sizeof(int) * (size_t)D.W * (D.NG + 3))) return -1; o += (size_t)D.W * (D.NG + 3); }
  std::vector<ObWorld> hw(D.W);
  if (obk_d2h(B->bk, hw.data(), D.world, D.W * sizeof(ObWorld))) return -1;
  for (int w = 0; w < D.W; w++) buf[o + w] = hw[w].sap_ndirty;
  return 0;
}
int dBatchSetOrderState(dBatchID B, const int *buf) {
  const ObBatchDev &D = B->caps;
  size_t o = 0;
  if (obk_h2d(B->bk, D.glist, buf + o, sizeof(int) * (size_t)D.W * D.NG)) return -1;
  o += (size_t)D.W * D.NG;
  if (D.sapstate) { if (obk_h2d(B->bk, D.sapstate, buf + o, sizeof(int) * (size_t)D.W * (D.NG + 3))) return -1; o += (size_t)D.W * (D.NG + 3); }
  std::vector<ObWorld> hw(D.W);
  if (obk_d2h(B->bk, hw.data(), D.world, D.W * sizeof(ObWorld))) return -1;
  for (int w = 0; w < D.W; w++) hw[w].sap_ndirty = buf[o + w];
  return obk_h2d(B->bk, D.world, hw.data(), D.W * sizeof(ObWorld));
}
int dBatchSetDebugTaps(dBatchID B, int enable) { B->debug_taps = enable != 0; return 0; }
int dBatchTimerStart(dBatchID B) { return obk_timer_start(B->bk); }
int dBatchTimerStop(dBatchID B, float *ms) { return obk_timer_stop(B->bk, ms); }
int dBatchSetKernelTiming(dBatchID B, int enable) { obk_set_kernel_timing(B->bk, enable); return 0; }
int dBatchGetKernelTimes(dBatchID B, double *ms, long long *launches, int nk) {
  double m[OBK_NKERNELS]; long long l[OBK_NKERNELS];
  obk_get_kernel_times(B->bk, m, l);
  for (int k = 0; k < nk && k < OBK_NKERNELS; k++) { ms[k] = m[k]; launches[k] = l[k]; }
  return OBK_NKERNELS;
}
const char *dBatchKernelName(int k) { return obk_kernel_name(k); }
int dBatchGetLargeWorldStats(dBatchID B, dBatchLargeWorldStats *out) {
  int iv[8]; double ms[8];
  if (!B || !out || obk_large_stats(B->bk, iv, ms)) return -1;
  out->pairs = iv[0]; out->contacts = iv[1]; out->contact_pairs = iv[2]; out->solved_contacts = iv[3];
  out->colours = iv[4]; out->colouring_rounds = iv[5]; out->sor_launches = iv[6]; out->steps_timed = iv[7];
  for (int k = 0; k < 7; k++) out->phase_ms[k] = ms[k];
  return 0;
}
int dBatchSplitExport(dBatchID B, void *handle) {
  char err[256] = "";
  if (!B || !handle) return -1;
  if (obk_split_export(B->bk, handle, err, sizeof err)) { ob_set_last_error("dBatchSplitExport: %s", err); return -1; }
  return 0;
}
int dBatchSplitAttach(dBatchID B, int rank, int nranks, const void *handles) {
  char err[256] = "";
  if (!B || !handles) return -1;
  if (obk_split_attach(B->bk, rank, nranks, handles, err, sizeof err)) { ob_set_last_error("dBatchSplitAttach: %s", err); return -1; }
  return 0;
}
void *dBatchGetStream(dBatchID B) { return obk_stream(B->bk); }
int dBatchRayCast(dBatchID B, int rays_per_world, const dReal *origin3, const dReal *dir3, const dReal *length, int ray_flags,
                  unsigned long category_bits, unsigned long collide_bits, dBatchRayHit *hits) {
  if (!B || rays_per_world < 0 || !origin3 || !dir3 || !length || !hits) { ob_set_last_error("dBatchRayCast: bad arguments"); return -1; }
  static_assert(sizeof(dBatchRayHit) == sizeof(ObRayHit), "dBatchRayHit layout");
  char err[512] = "";
  if (obk_raycast(B->bk, rays_per_world, origin3, dir3, length, ray_flags, (uint32_t)category_bits, (uint32_t)collide_bits, (ObRayHit *)hits, err, sizeof err)) {
    ob_set_last_error("dBatchRayCast: %s", err);
    return -1;
  }
  return 0;
}
long long dB200KernelLaunchCount(void) { return obk_launch_count(); }
int dB200LibmHost(int fn, int n, const float *a, const float *b, float *out) {
  for (int i = 0; i < n; i++) out[i] = fn == 0 ? ob_atan2f_glibc(a[i], b[i]) : (fn == 1 ? ob_sinf_glibc(a[i]) : ob_cosf_glibc(a[i]));
  return 0;
}
int dB200LibmDevice(int fn, int n, const float *a, const float *b, float *out) { return obk_libm(fn, n, a, b, out); }
}  // extern "C"

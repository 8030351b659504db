// scene_driver — runs a scene through the ODE C API and writes a binary trace.
//
// The same source is linked twice: against the unmodified reference
// (oracle/_ref/libode_ref_<prec>.so  -> oracle/_ref/driver_ref_<prec>) and against
// the product (libode_b200_<prec>.so -> build/driver_b200_<prec>).  With
// --mode callback both run the classic per-frame loop of the reference demos
// (ode/demo/demo_boxstack.cpp:553-597): dSpaceCollide + near callback +
// dWorldQuickStep + dJointGroupEmpty.  With --mode batch (product only, -DHAVE_BATCH)
// the dBatch* entry points are used and the same trace is produced from the
// library's parity taps.  --time prints throughput instead of writing a trace
// (the CPU-baseline leg of bench.py uses this with the reference build).
//
// Trace layout (little endian): header {magic 'ODTR', realsize, nworlds, nsteps},
// then for every step, for every world:
//   i32 ng, i32 glist[ng]           geom creation indices in space-list order before collide
//   i32 nb, real state0[nb*13]      pos3 quat4 lvel3 avel3 before the step
//   i32 np, i32 pairs[2*np]         near-callback (o1,o2) in call order
//   i32 nc, {i32 g1,g2; real pos3,normal3,depth} x nc   contact joints in creation order
//   real fb[nc*6]                   joint feedback f1,t1 per contact joint (lambda tap)
//   real state1[nb*13]              after the step
//   u32 seed_after
#include <stdio.h>
#include <time.h>
#include <algorithm>
#include <string>
#include "scenes.h"

struct Trace {
  FILE *f;
  void i32(int v) { fwrite(&v, 4, 1, f); }
  void u32(uint32_t v) { fwrite(&v, 4, 1, f); }
  void reals(const dReal *p, size_t n) { fwrite(p, sizeof(dReal), n, f); }
};

struct CbCtx {
  SceneWorld *sw;
  ScenePolicy *pol;
  std::vector<int> pairs;
  std::vector<int> cg;
  std::vector<dReal> cdata;
  std::vector<dJointFeedback *> fbs;
  long long ncontacts;
  bool record;
  unsigned calls, frame;   // near-callback invocations so far, frames so far (ScenePolicy::varmaxc)
};

static void near_cb(void *data, dGeomID o1, dGeomID o2) {
  CbCtx *c = (CbCtx *)data;
  if (c->record) {
    c->pairs.push_back((int)(intptr_t)dGeomGetData(o1));
    c->pairs.push_back((int)(intptr_t)dGeomGetData(o2));
  }
  const bool sp1 = dGeomIsSpace(o1) != 0, sp2 = dGeomIsSpace(o2) != 0;
  if ((sp1 || sp2) && c->pol->nested != 2) {
    // the manual's idiom for sub-spaces: collide the pair, then the interior of each space
    dSpaceCollide2(o1, o2, data, &near_cb);
    if (sp1) dSpaceCollide((dSpaceID)o1, data, &near_cb);
    if (sp2) dSpaceCollide((dSpaceID)o2, data, &near_cb);
    return;
  }
  dBodyID b1 = dGeomGetBody(o1), b2 = dGeomGetBody(o2);
  if (c->pol->skip_if_connected && b1 && b2 && dAreConnectedExcluding(b1, b2, dJointTypeContact)) return;
  enum { MAXC = 64 };
  dContact contact[MAXC];
  int maxc = c->pol->max_contacts;
  // varmaxc: a different max-contacts value from call to call and from frame to frame (an application that budgets contacts per
  // pair type); exercises dCollide against results the batched narrowphase computed with another value
  if (c->pol->varmaxc) maxc = 2 + (int)((c->calls * 5 + c->frame * 3) % 7);
  c->calls++;
  for (int i = 0; i < maxc; i++) {
    memset(&contact[i], 0, sizeof(dContact));
    contact[i].surface = c->pol->surface;
  }
  if (c->pol->sphere_mu > 0 && (dGeomGetClass(o1) == dSphereClass || dGeomGetClass(o2) == dSphereClass))
    for (int i = 0; i < maxc; i++) contact[i].surface.mu = c->pol->sphere_mu;   // demo_crash.cpp:128-131
  int n = dCollide(o1, o2, maxc, &contact[0].geom, sizeof(dContact));
  if (c->pol->fdir1)
    for (int i = 0; i < n; i++) {
      // first friction direction: the world x axis projected into the contact plane (falls back to y for normals along x), unit length
      const dReal *nn = contact[i].geom.normal;
      dVector3 a = {1, 0, 0, 0};
      if (nn[0] > (dReal)0.9 || nn[0] < (dReal)-0.9) { a[0] = 0; a[1] = 1; }
      const dReal k = a[0] * nn[0] + a[1] * nn[1] + a[2] * nn[2];
      dVector3 t = {a[0] - k * nn[0], a[1] - k * nn[1], a[2] - k * nn[2], 0};
      dNormalize3(t);
      contact[i].fdir1[0] = t[0]; contact[i].fdir1[1] = t[1]; contact[i].fdir1[2] = t[2];
    }
  const bool ray = dGeomGetClass(o1) == dRayClass || dGeomGetClass(o2) == dRayClass;   // query result, not a contact joint
  // contact geom identity (collision_kernel.cpp:331-343, collision_transform.cpp:143-151): g1 / g2 name the geoms of the
  // call, a geom transform being replaced by its encapsulated geom unless its info mode is on
  for (int i = 0; i < n && !sp1 && !sp2; i++) {
    dGeomID e[2] = {o1, o2};
    for (int k = 0; k < 2; k++)
      if (dGeomGetClass(e[k]) == dGeomTransformClass && !dGeomTransformGetInfo(e[k])) e[k] = dGeomTransformGetGeom(e[k]);
    if (contact[i].geom.g1 != e[0] || contact[i].geom.g2 != e[1]) { fprintf(stderr, "contact %d: g1 / g2 do not name the expected geoms\n", i); exit(3); }
  }
  for (int i = 0; i < n; i++) {
    dJointID j = 0;
    if (!ray) {
      if (sp1 || sp2) { b1 = dGeomGetBody(contact[i].geom.g1); b2 = dGeomGetBody(contact[i].geom.g2); }   // demo_buggy.cpp:104-107: dCollide on a space names the member geoms
      j = dJointCreateContact(c->sw->world, c->sw->cgroup, &contact[i]);
      dJointAttach(j, b1, b2);
    }
    if (c->record) {
      c->cg.push_back((int)(intptr_t)dGeomGetData(contact[i].geom.g1));
      c->cg.push_back((int)(intptr_t)dGeomGetData(contact[i].geom.g2));
      for (int k = 0; k < 3; k++) c->cdata.push_back(contact[i].geom.pos[k]);
      for (int k = 0; k < 3; k++) c->cdata.push_back(contact[i].geom.normal[k]);
      c->cdata.push_back(contact[i].geom.depth);
      dJointFeedback *fb = new dJointFeedback;
      memset(fb, 0, sizeof(*fb));
      if (j) dJointSetFeedback(j, fb);
      c->fbs.push_back(fb);
    }
  }
  if (!ray) c->ncontacts += n;
}

// scenes with a second space: dSpaceCollide2 in its three forms (space x space, geom x space, space x geom)
static void collide_second_space(SceneWorld &sw, CbCtx &ctx) {
  if (!sw.space2) return;
  dSpaceCollide2((dGeomID)sw.space, (dGeomID)sw.space2, &ctx, &near_cb);
  const int n2 = dSpaceGetNumGeoms(sw.space2);
  for (int i = 0; i < 3 && i < n2; i++) dSpaceCollide2(dSpaceGetGeom(sw.space2, i), (dGeomID)sw.space, &ctx, &near_cb);
  for (int i = 3; i < 6 && i < n2; i++) dSpaceCollide2((dGeomID)sw.space, dSpaceGetGeom(sw.space2, i), &ctx, &near_cb);
  if (n2 > 7) dSpaceCollide2(dSpaceGetGeom(sw.space2, 6), dSpaceGetGeom(sw.space, 2), &ctx, &near_cb);
}

static void dump_state(Trace &t, SceneWorld &sw) {
  t.i32((int)sw.bodies.size());
  for (size_t i = 0; i < sw.bodies.size(); i++) {
    dBodyID b = sw.bodies[i];
    t.reals(dBodyGetPosition(b), 3);
    t.reals(dBodyGetQuaternion(b), 4);
    t.reals(dBodyGetLinearVel(b), 3);
    t.reals(dBodyGetAngularVel(b), 3);
  }
}

// --resync: read the pre-step body state of every world for the next step from a reference trace
struct Resync {
  FILE *f;
  Resync() : f(NULL) {}
  bool open(const std::string &path, int nworlds) {
    f = fopen(path.c_str(), "rb");
    int hdr[4];
    return f && fread(hdr, 4, 4, f) == 4 && hdr[1] == (int)sizeof(dReal) && hdr[2] == nworlds;
  }
  // st: [nb*13] of world w; call for w = 0..nworlds-1 in order, once per step
  bool next(int nb, std::vector<dReal> &st) {
    int n;
    if (fread(&n, 4, 1, f) != 1) return false;
    fseek(f, 4L * n, SEEK_CUR);
    if (fread(&n, 4, 1, f) != 1 || n != nb) return false;
    st.resize((size_t)nb * 13);
    if (fread(st.data(), sizeof(dReal), st.size(), f) != st.size()) return false;
    if (fread(&n, 4, 1, f) != 1) return false;
    fseek(f, 8L * n, SEEK_CUR);                                       // pairs
    if (fread(&n, 4, 1, f) != 1) return false;
    fseek(f, (long)n * (8 + 7 * (long)sizeof(dReal)), SEEK_CUR);       // contacts
    fseek(f, (long)n * 6 * (long)sizeof(dReal), SEEK_CUR);             // feedback
    fseek(f, (long)nb * 13 * (long)sizeof(dReal) + 4, SEEK_CUR);       // state1 + seed
    return true;
  }
};


// --rays: the wheel probe of the ray-cast vehicle (demos/raycar/car.cpp:353-371) for a deterministic set of rays per world: a ray geom
// outside any space, dGeomRaySet / SetParams / SetClosestHit, dSpaceCollide2(space, ray) with a callback keeping the nearest
// dCollide(ray, geom, 1) result.  The batched build answers the same rays with ONE dBatchRayCast call; the two outputs must be equal.
struct RayProbe { dReal o[3], d[3], len; };
struct RayHitOut { dReal pos[3]; dReal depth; dReal normal[3]; int geom; };
static void make_probes(int w, int n, std::vector<RayProbe> &out) {
  xs32 rng(0xC0FFEEu ^ (uint32_t)(w * 7919 + 13));
  out.resize(n);
  for (int i = 0; i < n; i++) {
    RayProbe &r = out[i];
    r.o[0] = rng.uni(-2.5, 2.5); r.o[1] = rng.uni(-2.5, 2.5); r.o[2] = rng.uni(0.3, 4.0);
    dVector3 d = {rng.uni(-0.6, 0.6), rng.uni(-0.6, 0.6), (dReal)(i % 5 == 4 ? 0.4 : -1.0), 0};
    dNormalize3(d);
    r.d[0] = d[0]; r.d[1] = d[1]; r.d[2] = d[2];
    r.len = rng.uni(0.5, 6.0);
  }
}
struct RayCb { dGeomID ray; RayHitOut best; bool have; };
static void ray_cb(void *data, dGeomID o1, dGeomID o2) {
  RayCb *c = (RayCb *)data;
  dGeomID g = o1 == c->ray ? o2 : o1;
  dContactGeom cg;
  if (dCollide(c->ray, g, 1, &cg, sizeof cg) == 1 && (!c->have || cg.depth < c->best.depth)) {
    c->have = true;
    for (int k = 0; k < 3; k++) { c->best.pos[k] = cg.pos[k]; c->best.normal[k] = cg.normal[k]; }
    c->best.depth = cg.depth; c->best.geom = (int)(intptr_t)dGeomGetData(g);
  }
}

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// --moved-log: every body reports through dBodySetMovedCallback; the log holds, per step, the order of the calls
static FILE *g_moved_log = NULL;
static void moved_cb(dBodyID b) { if (g_moved_log) fprintf(g_moved_log, " %d", (int)(intptr_t)dBodyGetData(b)); }

int main(int argc, char **argv) {
  std::string scene = "stack32", out = "", mode = "callback", resync = "", export_dif = "", moved_log = "";
  int nworlds = 1, nsteps = 10, world0 = 0, timing = 0, settle = 0, maxc_world = 0, large = 0;
  uint32_t seed_xor = 0;
  double h = 0.01;
  int rays = 0;
  std::string rays_out = "";
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--scene") scene = argv[++i];
    else if (a == "--out") out = argv[++i];
    else if (a == "--mode") mode = argv[++i];
    else if (a == "--worlds") nworlds = atoi(argv[++i]);
    else if (a == "--world0") world0 = atoi(argv[++i]);
    else if (a == "--steps") nsteps = atoi(argv[++i]);
    else if (a == "--h") h = atof(argv[++i]);
    else if (a == "--time") timing = 1;
    else if (a == "--settle") settle = atoi(argv[++i]);
    else if (a == "--contacts-cap") maxc_world = atoi(argv[++i]);
    else if (a == "--resync") resync = argv[++i];
    else if (a == "--seed-xor") seed_xor = (uint32_t)strtoul(argv[++i], 0, 0);   // other SOR shuffle stream, same scene
    else if (a == "--moved-log") moved_log = argv[++i];                          // callback mode: order of the body moved-callbacks per step
    else if (a == "--export-dif") export_dif = argv[++i];                        // callback mode: dWorldExportDIF of every world inside the last step (contact joints alive)
    else if (a == "--large") large = 1;                                          // force the large-world path (batch mode)
    else if (a == "--rays") { rays = atoi(argv[++i]); rays_out = argv[++i]; }    // after the steps: N probe rays per world, nearest hits written to a file
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
  }
  dInitODE2(0);
  std::vector<SceneWorld> worlds(nworlds);
  ScenePolicy pol;
  for (int w = 0; w < nworlds; w++)
    if (scene_build(scene.c_str(), worlds[w], world0 + w, pol)) { fprintf(stderr, "bad scene\n"); return 2; }
  for (int w = 0; w < nworlds; w++) worlds[w].seed ^= seed_xor;

  if (!moved_log.empty()) {
    g_moved_log = fopen(moved_log.c_str(), "w");
    for (int w = 0; w < nworlds; w++)
      for (size_t i = 0; i < worlds[w].bodies.size(); i++) { dBodySetData(worlds[w].bodies[i], (void *)(intptr_t)i); dBodySetMovedCallback(worlds[w].bodies[i], &moved_cb); }
  }
  Trace t;
  t.f = NULL;
  if (!out.empty()) {
    t.f = fopen(out.c_str(), "wb");
    if (!t.f) { perror("open"); return 2; }
    t.u32(0x5254444f);
    t.i32((int)sizeof(dReal));
    t.i32(nworlds);
    t.i32(nsteps);
  }
  long long body_steps = 0, contacts = 0, pairs = 0;
  double t0 = now_s();

  if (mode == "callback") {
    CbCtx ctx;
    ctx.pol = &pol;
    ctx.ncontacts = 0;
    ctx.record = false;
    ctx.calls = 0; ctx.frame = 0;
    for (int s = 0; s < settle; s++, ctx.frame++)
      for (int w = 0; w < nworlds; w++) {
        SceneWorld &sw = worlds[w];
        ctx.sw = &sw;
        dRandSetSeed(sw.seed);
        dSpaceCollide(sw.space, &ctx, &near_cb);
        collide_second_space(sw, ctx);
        dWorldQuickStep(sw.world, (dReal)h);
        sw.seed = (uint32_t)dRandGetSeed();
        dJointGroupEmpty(sw.cgroup);
      }
    ctx.ncontacts = 0;
    ctx.record = t.f != NULL;
    Resync rsc;
    if (!resync.empty() && !rsc.open(resync, nworlds)) { fprintf(stderr, "bad --resync trace\n"); return 2; }
    t0 = now_s();
    for (int s = 0; s < nsteps; s++, ctx.frame++) {
      for (int w = 0; w < nworlds; w++) {
        SceneWorld &sw = worlds[w];
        ctx.sw = &sw;
        if (rsc.f) {
          std::vector<dReal> st;
          if (!rsc.next((int)sw.bodies.size(), st)) { fprintf(stderr, "resync trace too short\n"); return 2; }
          for (size_t i = 0; i < sw.bodies.size(); i++) {
            const dReal *p = &st[i * 13];
            dBodySetPosition(sw.bodies[i], p[0], p[1], p[2]);
            dBodySetQuaternion(sw.bodies[i], p + 3);
            dBodySetLinearVel(sw.bodies[i], p[7], p[8], p[9]);
            dBodySetAngularVel(sw.bodies[i], p[10], p[11], p[12]);
          }
        }
        if (t.f) {
          int ng = dSpaceGetNumGeoms(sw.space);
          t.i32(ng);
          for (int i = 0; i < ng; i++) t.i32((int)(intptr_t)dGeomGetData(dSpaceGetGeom(sw.space, i)));
          dump_state(t, sw);
        }
        dRandSetSeed(sw.seed);
        dSpaceCollide(sw.space, &ctx, &near_cb);
        collide_second_space(sw, ctx);
        if (!export_dif.empty() && s == nsteps - 1) {   // inside the last step: the contact joints of this step are alive
          FILE *ef = fopen(export_dif.c_str(), w == 0 ? "w" : "a");
          if (!ef) { perror("export"); return 2; }
          char prefix[32];
          snprintf(prefix, sizeof prefix, "w%d_", w);
          dWorldExportDIF(sw.world, ef, prefix);
          fclose(ef);
        }
        dWorldQuickStep(sw.world, (dReal)h);
        if (g_moved_log) fprintf(g_moved_log, "\n");
        sw.seed = (uint32_t)dRandGetSeed();
        if (t.f) {
          t.i32((int)ctx.pairs.size() / 2);
          fwrite(ctx.pairs.data(), 4, ctx.pairs.size(), t.f);
          int nc = (int)ctx.cg.size() / 2;
          t.i32(nc);
          for (int i = 0; i < nc; i++) {
            t.i32(ctx.cg[2 * i]);
            t.i32(ctx.cg[2 * i + 1]);
            t.reals(&ctx.cdata[7 * i], 7);
          }
          for (int i = 0; i < nc; i++) {
            t.reals(ctx.fbs[i]->f1, 3);
            t.reals(ctx.fbs[i]->t1, 3);
            delete ctx.fbs[i];
          }
          pairs += ctx.pairs.size() / 2;
          ctx.pairs.clear(); ctx.cg.clear(); ctx.cdata.clear(); ctx.fbs.clear();
        }
        dJointGroupEmpty(sw.cgroup);
        if (t.f) {
          for (size_t i = 0; i < sw.bodies.size(); i++) {
            dBodyID b = sw.bodies[i];
            t.reals(dBodyGetPosition(b), 3);
            t.reals(dBodyGetQuaternion(b), 4);
            t.reals(dBodyGetLinearVel(b), 3);
            t.reals(dBodyGetAngularVel(b), 3);
          }
          t.u32(sw.seed);
        }
        body_steps += (long long)sw.bodies.size();
      }
    }
    contacts = ctx.ncontacts;
    if (rays > 0) {
      FILE *rf = fopen(rays_out.c_str(), "wb");
      if (!rf) { perror("rays"); return 2; }
      for (int w = 0; w < nworlds; w++) {
        SceneWorld &sw = worlds[w];
        std::vector<RayProbe> pr;
        make_probes(world0 + w, rays, pr);
        dGeomID ray = dCreateRay(0, 1);
        for (int i = 0; i < rays; i++) {
          dGeomRaySetLength(ray, pr[i].len);
          dGeomRaySet(ray, pr[i].o[0], pr[i].o[1], pr[i].o[2], pr[i].d[0], pr[i].d[1], pr[i].d[2]);
          dGeomRaySetParams(ray, 0, 0);
          dGeomRaySetClosestHit(ray, 1);
          RayCb cb;
          cb.ray = ray; cb.have = false;
          memset(&cb.best, 0, sizeof cb.best);
          cb.best.depth = pr[i].len; cb.best.geom = -1;
          dSpaceCollide2((dGeomID)sw.space, ray, &cb, &ray_cb);
          fwrite(&cb.best, sizeof cb.best, 1, rf);
        }
        dGeomDestroy(ray);
      }
      fclose(rf);
    }
  }
#ifdef HAVE_BATCH
  else if (mode == "batch") {
    std::vector<dWorldID> wv(nworlds);
    std::vector<dSpaceID> sv(nworlds);
    std::vector<uint32_t> seeds(nworlds);
    for (int w = 0; w < nworlds; w++) { wv[w] = worlds[w].world; sv[w] = worlds[w].space; seeds[w] = worlds[w].seed; }
    dBatchDesc desc;
    memset(&desc, 0, sizeof(desc));
    desc.max_contacts_per_world = maxc_world;
    desc.large_world = large;
    dBatchID B = dBatchCreate(nworlds, wv.data(), sv.data(), &desc);
    if (!B) { fprintf(stderr, "dBatchCreate failed: %s\n", dB200LastError()); return 3; }
    dBatchContactPolicy bp[2];
    memset(bp, 0, sizeof(bp));
    int nrows = 0;
    if (pol.sphere_mu > 0) {   // pairs with a sphere first (category bit SCENE_CAT_SPHERE), then the catch-all row
      bp[0].cat_mask1 = SCENE_CAT_SPHERE; bp[0].cat_mask2 = ~0ul;
      bp[0].max_contacts = pol.max_contacts; bp[0].skip_if_connected = pol.skip_if_connected;
      bp[0].surface = pol.surface; bp[0].surface.mu = pol.sphere_mu;
      nrows = 1;
    }
    bp[nrows].cat_mask1 = bp[nrows].cat_mask2 = ~0ul;
    bp[nrows].max_contacts = pol.max_contacts;
    bp[nrows].skip_if_connected = pol.skip_if_connected;
    bp[nrows].surface = pol.surface;
    nrows++;
    dBatchSetContactPolicy(B, bp, nrows);
    dBatchSetSeeds(B, seeds.data());
    int nb = dBatchNumBodies(B);
    std::vector<dReal> pos(nworlds * nb * 3), quat(nworlds * nb * 4), lv(nworlds * nb * 3), av(nworlds * nb * 3);
    std::vector<int> status(nworlds);
    if (settle > 0) {
      if (dBatchCollideAndQuickStep(B, (dReal)h, settle, status.data())) { fprintf(stderr, "settle failed: %s\n", dB200LastError()); return 3; }
      dBatchResetCounters(B);
    }
    t0 = now_s();
    if (!t.f) {
      // timing mode: one call, everything device-resident
      int rc = dBatchCollideAndQuickStep(B, (dReal)h, nsteps, status.data());
      if (rc) { fprintf(stderr, "step failed: %s\n", dB200LastError()); return 3; }
      dBatchCounters c;
      dBatchGetCounters(B, &c);
      body_steps = c.body_steps; contacts = c.contacts; pairs = c.pairs;
    } else {
      const int tapcap = std::max(65536, 16 * nb);
      std::vector<int> pbuf(2 * (size_t)tapcap), cg(2 * (size_t)tapcap);
      std::vector<dReal> cd(7 * (size_t)tapcap), lam(12 * (size_t)tapcap);
      // lock-step parity (SURVEY 8d "parity protocol", K = 1): before every step load the body state the
      // reference had before ITS step from a reference trace, so chaos cannot amplify rounding differences
      Resync rsc;
      if (!resync.empty() && !rsc.open(resync, nworlds)) { fprintf(stderr, "bad --resync trace\n"); return 2; }
      for (int s = 0; s < nsteps; s++) {
        if (rsc.f) {
          for (int w = 0; w < nworlds; w++) {
            std::vector<dReal> st;
            if (!rsc.next(nb, st)) { fprintf(stderr, "resync trace too short / body count\n"); return 2; }
            for (int b = 0; b < nb; b++) {
              memcpy(&pos[(w * nb + b) * 3], &st[b * 13], 3 * sizeof(dReal));
              memcpy(&quat[(w * nb + b) * 4], &st[b * 13 + 3], 4 * sizeof(dReal));
              memcpy(&lv[(w * nb + b) * 3], &st[b * 13 + 7], 3 * sizeof(dReal));
              memcpy(&av[(w * nb + b) * 3], &st[b * 13 + 10], 3 * sizeof(dReal));
            }
          }
          dBatchSetBodyState(B, pos.data(), quat.data(), lv.data(), av.data());
        }
        // the trace is world-major inside a step, so gather per world
        dBatchGetBodyState(B, pos.data(), quat.data(), lv.data(), av.data());
        std::vector<dReal> st0(pos.size() + quat.size() + lv.size() + av.size());
        for (int w = 0; w < nworlds; w++)
          for (int b = 0; b < nb; b++) {
            dReal *d = &st0[(size_t)(w * nb + b) * 13];
            memcpy(d, &pos[(w * nb + b) * 3], 3 * sizeof(dReal));
            memcpy(d + 3, &quat[(w * nb + b) * 4], 4 * sizeof(dReal));
            memcpy(d + 7, &lv[(w * nb + b) * 3], 3 * sizeof(dReal));
            memcpy(d + 10, &av[(w * nb + b) * 3], 3 * sizeof(dReal));
          }
        std::vector<std::vector<int> > glists(nworlds);
        for (int w = 0; w < nworlds; w++) {
          glists[w].resize(std::max(4096, nb + 64));
          int ng = dBatchDebugGeomOrder(B, w, glists[w].data(), (int)glists[w].size());
          glists[w].resize(ng);
        }
        int rc = dBatchCollideAndQuickStep(B, (dReal)h, 1, status.data());
        if (rc) { fprintf(stderr, "step failed: %s\n", dB200LastError()); return 3; }
        dBatchGetBodyState(B, pos.data(), quat.data(), lv.data(), av.data());
        dBatchGetSeeds(B, seeds.data());
        for (int w = 0; w < nworlds; w++) {
          if (status[w]) fprintf(stderr, "world %d step %d status %d\n", w, s, status[w]);
          t.i32((int)glists[w].size());
          fwrite(glists[w].data(), 4, glists[w].size(), t.f);
          t.i32(nb);
          t.reals(&st0[(size_t)w * nb * 13], (size_t)nb * 13);
          int np = dBatchDebugPairs(B, w, pbuf.data(), tapcap);
          t.i32(np);
          fwrite(pbuf.data(), 4, 2 * np, t.f);
          int nc = dBatchDebugContacts(B, w, cd.data(), cg.data(), tapcap);
          t.i32(nc);
          for (int i = 0; i < nc; i++) { t.i32(cg[2 * i]); t.i32(cg[2 * i + 1]); t.reals(&cd[7 * i], 7); }
          int nl = dBatchDebugFeedback(B, w, lam.data(), tapcap);
          (void)nl;
          t.reals(lam.data(), (size_t)nc * 6);
          for (int b = 0; b < nb; b++) {
            t.reals(&pos[(w * nb + b) * 3], 3);
            t.reals(&quat[(w * nb + b) * 4], 4);
            t.reals(&lv[(w * nb + b) * 3], 3);
            t.reals(&av[(w * nb + b) * 3], 3);
          }
          t.u32(seeds[w]);
          body_steps += nb; pairs += np; contacts += nc;
        }
      }
    }
    if (rays > 0) {
      std::vector<dReal> ro((size_t)nworlds * rays * 3), rd((size_t)nworlds * rays * 3), rl((size_t)nworlds * rays);
      for (int w = 0; w < nworlds; w++) {
        std::vector<RayProbe> pr;
        make_probes(world0 + w, rays, pr);
        for (int i = 0; i < rays; i++) {
          const size_t k = (size_t)w * rays + i;
          for (int e = 0; e < 3; e++) { ro[3 * k + e] = pr[i].o[e]; rd[3 * k + e] = pr[i].d[e]; }
          rl[k] = pr[i].len;
        }
      }
      std::vector<dBatchRayHit> hits((size_t)nworlds * rays);
      if (dBatchRayCast(B, rays, ro.data(), rd.data(), rl.data(), 4 /* closest hit */, ~0ul, ~0ul, hits.data())) { fprintf(stderr, "dBatchRayCast failed: %s\n", dB200LastError()); return 3; }
      FILE *rf = fopen(rays_out.c_str(), "wb");
      if (!rf) { perror("rays"); return 2; }
      fwrite(hits.data(), sizeof(dBatchRayHit), hits.size(), rf);
      fclose(rf);
    }
    dBatchDestroy(B);
  }
#endif
  else { fprintf(stderr, "mode %s not available in this build\n", mode.c_str()); return 2; }

  double dt = now_s() - t0;
  if (t.f) fclose(t.f);
  if (g_moved_log) fclose(g_moved_log);
  if (timing)
    printf("{\"scene\":\"%s\",\"mode\":\"%s\",\"worlds\":%d,\"steps\":%d,\"seconds\":%.6f,\"body_steps\":%lld,"
           "\"contacts\":%lld,\"body_steps_per_sec\":%.1f,\"contacts_per_sec\":%.1f,\"realsize\":%d}\n",
           scene.c_str(), mode.c_str(), nworlds, nsteps, dt, body_steps, contacts, body_steps / dt, contacts / dt,
           (int)sizeof(dReal));
  dCloseODE();
  return 0;
}

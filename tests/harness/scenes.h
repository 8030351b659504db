// Deterministic scene generators for BASELINE.json's configs, written only
// against the public ODE C API (include/ode_b200/ode.h) so that the SAME call
// sequence drives both the unmodified reference (oracle/_ref/libode_ref_*.so)
// and libode_b200_*.so.  Creation ORDER matters, so do
// not reorder calls.  Test/bench infrastructure, not product code.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ode_b200/ode.h"

struct SceneWorld {
  dWorldID world;
  dSpaceID space;
  dJointGroupID cgroup;
  std::vector<dBodyID> bodies;   // creation order
  std::vector<dGeomID> geoms;    // creation order (index stored in geom data)
  std::vector<dJointID> joints;  // permanent joints, creation order
  uint32_t seed;                 // per-world LCG stream for the SOR shuffle
};

// policy = what the near callback does (mirrors the reference demos)
struct ScenePolicy {
  int max_contacts;
  int skip_if_connected;
  dSurfaceParameters surface;
};

struct xs32 {  // scene jitter RNG (not ODE's)
  uint32_t s;
  explicit xs32(uint32_t seed) : s(seed ? seed : 0x1234567u) {}
  uint32_t next() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
  // uniform in [a,b), computed in double then rounded to dReal once
  dReal uni(double a, double b) { return (dReal)(a + (b - a) * ((next() >> 8) * (1.0 / 16777216.0))); }
};

static inline uint32_t scene_world_seed(int w) { return 0x9E3779B9u * (uint32_t)(w + 1); }

static inline dGeomID scene_add_geom(SceneWorld &sw, dGeomID g) {
  dGeomSetData(g, (void *)(intptr_t)sw.geoms.size());
  sw.geoms.push_back(g);
  return g;
}

static inline dBodyID scene_add_box(SceneWorld &sw, dReal density, dReal lx, dReal ly, dReal lz,
                                    dReal x, dReal y, dReal z) {
  dBodyID b = dBodyCreate(sw.world);
  dBodySetPosition(b, x, y, z);
  dMass m;
  dMassSetBox(&m, density, lx, ly, lz);
  dBodySetMass(b, &m);
  dGeomID g = scene_add_geom(sw, dCreateBox(sw.space, lx, ly, lz));
  dGeomSetBody(g, b);
  sw.bodies.push_back(b);
  return b;
}

static inline dBodyID scene_add_sphere(SceneWorld &sw, dReal density, dReal r, dReal x, dReal y, dReal z) {
  dBodyID b = dBodyCreate(sw.world);
  dBodySetPosition(b, x, y, z);
  dMass m;
  dMassSetSphere(&m, density, r);
  dBodySetMass(b, &m);
  dGeomID g = scene_add_geom(sw, dCreateSphere(sw.space, r));
  dGeomSetBody(g, b);
  sw.bodies.push_back(b);
  return b;
}

static inline void scene_world_base(SceneWorld &sw, int w, bool sap = false) {
  sw.world = dWorldCreate();
  sw.space = sap ? dSweepAndPruneSpaceCreate(0, dSAP_AXES_XYZ) : dHashSpaceCreate(0);
  sw.cgroup = dJointGroupCreate(0);
  sw.seed = scene_world_seed(w);
  dWorldSetGravity(sw.world, 0, 0, (dReal)-9.81);
  dWorldSetCFM(sw.world, (dReal)1e-5);
  dWorldSetERP(sw.world, (dReal)0.2);
  dWorldSetQuickStepNumIterations(sw.world, 20);
  dWorldSetQuickStepW(sw.world, (dReal)1.3);
}

static inline ScenePolicy policy_boxstack() {
  // ode/demo/demo_boxstack.cpp:142-152
  ScenePolicy p;
  memset(&p, 0, sizeof(p));
  p.max_contacts = 8;
  p.skip_if_connected = 1;
  p.surface.mode = dContactBounce | dContactSoftCFM;
  p.surface.mu = dInfinity;
  p.surface.mu2 = 0;
  p.surface.bounce = (dReal)0.1;
  p.surface.bounce_vel = (dReal)0.1;
  p.surface.soft_cfm = (dReal)0.01;
  return p;
}

static inline ScenePolicy policy_crash() {
  // ode/demo/demo_crash.cpp:128-137
  ScenePolicy p;
  memset(&p, 0, sizeof(p));
  p.max_contacts = 4;
  p.skip_if_connected = 1;
  p.surface.mode = dContactSlip1 | dContactSlip2 | dContactSoftERP | dContactSoftCFM | dContactApprox1;
  p.surface.mu = (dReal)0.5;
  p.surface.slip1 = (dReal)0.0;
  p.surface.slip2 = (dReal)0.0;
  p.surface.soft_erp = (dReal)0.8;
  p.surface.soft_cfm = (dReal)0.01;
  return p;
}

// config 1a: 64 unit boxes as a resting 4x4x4 block on a plane
static inline void scene_block64(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  sw.seed = 0;
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int z = 0; z < 4; z++)
    for (int y = 0; y < 4; y++)
      for (int x = 0; x < 4; x++)
        scene_add_box(sw, 1, 1, 1, 1, (dReal)x, (dReal)y, (dReal)(0.5 + z));
}

// config 1b: 1x1x64 tower with 0.01 gaps (topples)
static inline void scene_tower64(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  sw.seed = 0;
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int i = 0; i < 64; i++) scene_add_box(sw, 1, 1, 1, 1, 0, 0, (dReal)(0.5 + i * 1.01));
}

// config 2: plane + 32 boxes (sides U[0.2,0.6]) in a loose jittered 4x4x2 stack
// + 8 spheres r in U[0.1,0.3] dropped from z in [3,4]
static inline void scene_stack32(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int k = 0; k < 2; k++)
    for (int j = 0; j < 4; j++)
      for (int i = 0; i < 4; i++) {
        dReal lx = rng.uni(0.2, 0.6), ly = rng.uni(0.2, 0.6), lz = rng.uni(0.2, 0.6);
        dReal x = (dReal)((i - 1.5) * 0.45) + rng.uni(-0.02, 0.02);
        dReal y = (dReal)((j - 1.5) * 0.45) + rng.uni(-0.02, 0.02);
        dReal z = (dReal)(0.32 + k * 0.62) + rng.uni(-0.02, 0.02);
        scene_add_box(sw, 5, lx, ly, lz, x, y, z);
      }
  for (int i = 0; i < 8; i++) {
    dReal r = rng.uni(0.1, 0.3);
    scene_add_sphere(sw, 5, r, rng.uni(-0.8, 0.8), rng.uni(-0.8, 0.8), rng.uni(3.0, 4.0));
  }
}

// small mixed scene used by unit parity tests: random boxes + spheres dropped
// with random orientation and spin (exercises edge-edge box contacts, sphere-box)
static inline void scene_mixed(SceneWorld &sw, int w, int nbox, int nsph) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0xA5A5A5A5u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int i = 0; i < nbox; i++) {
    dBodyID b = scene_add_box(sw, 2, rng.uni(0.2, 0.9), rng.uni(0.2, 0.9), rng.uni(0.2, 0.9),
                              rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(0.5, 4));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dBodySetAngularVel(b, rng.uni(-2, 2), rng.uni(-2, 2), rng.uni(-2, 2));
    dBodySetLinearVel(b, rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1));
  }
  for (int i = 0; i < nsph; i++) {
    dBodyID b = scene_add_sphere(sw, 2, rng.uni(0.15, 0.5), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(0.5, 4));
    dBodySetAngularVel(b, rng.uni(-2, 2), rng.uni(-2, 2), rng.uni(-2, 2));
  }
}

static inline int scene_build(const char *name, SceneWorld &sw, int w, ScenePolicy &pol) {
  pol = policy_boxstack();
  if (!strcmp(name, "block64")) { scene_block64(sw, w); return 0; }
  if (!strcmp(name, "tower64")) { scene_tower64(sw, w); return 0; }
  if (!strcmp(name, "stack32")) { scene_stack32(sw, w); return 0; }
  if (!strcmp(name, "mixed")) { scene_mixed(sw, w, 12, 6); return 0; }
  if (!strcmp(name, "mixed_maxc4")) { scene_mixed(sw, w, 12, 6); pol = policy_crash(); return 0; }
  if (!strcmp(name, "free6")) {  // no contacts: integrator + gyroscopic term only
    scene_world_base(sw, w);
    xs32 rng(sw.seed);
    for (int i = 0; i < 6; i++) {
      dBodyID b = scene_add_box(sw, 1, rng.uni(0.2, 1), rng.uni(0.2, 1), rng.uni(0.2, 1),
                                (dReal)(3 * i), 0, 100);
      dBodySetAngularVel(b, rng.uni(-5, 5), rng.uni(-5, 5), rng.uni(-5, 5));
      dBodySetLinearVel(b, rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1));
    }
    return 0;
  }
  return -1;
}

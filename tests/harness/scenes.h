// Deterministic scene generators for BASELINE.json's configs, written only
// against the public ODE C API (include/ode_b200/ode.h) so that the SAME call
// sequence drives both the unmodified reference (oracle/_ref/libode_ref_*.so)
// and libode_b200_*.so.  Creation ORDER matters, so do
// not reorder calls.  Test/bench infrastructure, not product code.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ode_b200/ode.h"

struct SceneWorld {
  dWorldID world;
  dSpaceID space;
  dSpaceID space2;               // optional second space (query geoms), collided with dSpaceCollide2
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
  // knobs only a near callback can express (classic loop; the batched path's policy table has no equivalent):
  int fdir1;     // 1: dContactFDir1 with a first friction direction computed from the contact normal
  int varmaxc;   // 1: the max-contacts value passed to dCollide varies from call to call
  dReal sphere_mu;   // > 0: contacts of a pair with a sphere geom take this mu instead of surface.mu (demo_crash.cpp:128-131); the scene gives
                     // its spheres category bit SCENE_CAT_SPHERE so that the batched path's policy table can tell them apart
  int nested;    // sub-spaces among the pairs: 1 = the manual's idiom (dSpaceCollide2 on the pair, then dSpaceCollide on each space
                 // for its interior pairs), 2 = demo_buggy's way (dCollide straight on the (space, geom) pair, bodies from the contacts)
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

// space class used by scene_world_base: 0 hash (default), 1 SAP (XYZ), 2 simple, 3 SAP (ZXY)
static int g_scene_space_kind = 0;
static inline void scene_world_base(SceneWorld &sw, int w) {
  sw.world = dWorldCreate();
  sw.space2 = 0;
  sw.space = g_scene_space_kind == 1 ? dSweepAndPruneSpaceCreate(0, dSAP_AXES_XYZ)
           : g_scene_space_kind == 3 ? dSweepAndPruneSpaceCreate(0, dSAP_AXES_ZXY)
           : g_scene_space_kind == 2 ? dSimpleSpaceCreate(0) : dHashSpaceCreate(0);
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

#define SCENE_CAT_SPHERE 2ul
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

// chain of boxes linked by ball joints, first link pinned to the world, swinging onto a plane
static inline void scene_chain(SceneWorld &sw, int w, int n) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x00C0FFEEu);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID prev = 0;
  for (int i = 0; i < n; i++) {
    dBodyID b = scene_add_box(sw, 3, (dReal)0.35, (dReal)0.12, (dReal)0.12, (dReal)(0.4 * i), 0, (dReal)1.6 + rng.uni(-0.01, 0.01));
    dJointID j = dJointCreateBall(sw.world, 0);
    dJointAttach(j, b, prev);
    dJointSetBallAnchor(j, (dReal)(0.4 * i - 0.2), 0, (dReal)1.6);
    if (i == 2) { dJointSetBallParam(j, dParamERP, (dReal)0.5); dJointSetBallParam(j, dParamCFM, (dReal)1e-3); }
    sw.joints.push_back(j);
    prev = b;
  }
}

// boxes linked by hinges: stops, a free motor and a motor that runs into its stop (joint.cpp:638-657)
static inline void scene_hinges(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0B1E55EDu);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID prev = 0;
  for (int i = 0; i < 6; i++) {
    dBodyID b = scene_add_box(sw, 2, (dReal)0.5, (dReal)0.2, (dReal)0.1, (dReal)(0.55 * i), rng.uni(-0.01, 0.01), (dReal)1.2);
    dJointID j = dJointCreateHinge(sw.world, 0);
    dJointAttach(j, prev, b);   // first one: (0, b) -> reversed attach
    dJointSetHingeAnchor(j, (dReal)(0.55 * i - 0.275), 0, (dReal)1.2);
    dJointSetHingeAxis(j, 0, 1, (dReal)(i == 3 ? 0.2 : 0));
    if (i == 1 || i == 4) { dJointSetHingeParam(j, dParamLoStop, (dReal)-0.4); dJointSetHingeParam(j, dParamHiStop, (dReal)0.3); }
    if (i == 2) { dJointSetHingeParam(j, dParamVel, (dReal)1.5); dJointSetHingeParam(j, dParamFMax, (dReal)4); }
    if (i == 5) {   // powered and limited: reaches the stop, exercises the torque side effect + bounce
      dJointSetHingeParam(j, dParamLoStop, (dReal)-0.2); dJointSetHingeParam(j, dParamHiStop, (dReal)0.2);
      dJointSetHingeParam(j, dParamVel, (dReal)3); dJointSetHingeParam(j, dParamFMax, (dReal)6);
      dJointSetHingeParam(j, dParamBounce, (dReal)0.3); dJointSetHingeParam(j, dParamFudgeFactor, (dReal)0.5);
    }
    sw.joints.push_back(j);
    prev = b;
  }
}

// demo_buggy-style vehicle: box chassis + 4 sphere wheels on hinge2 joints (demo_buggy.cpp:226-294),
// rear wheels locked by stops, front wheels steered by a limited motor, all wheels driven
static inline void scene_buggy(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0000B066u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  const dReal L = (dReal)0.7, Wd = (dReal)0.5, H = (dReal)0.2, R = (dReal)0.18, Z = (dReal)0.5;
  dBodyID chassis = scene_add_box(sw, 1 / (L * Wd * H), L, Wd, H, 0, 0, Z);
  const dReal wx[4] = {(dReal)(0.5 * L), (dReal)(0.5 * L), (dReal)(-0.5 * L), (dReal)(-0.5 * L)};
  const dReal wy[4] = {(dReal)(0.5 * Wd), (dReal)(-0.5 * Wd), (dReal)(0.5 * Wd), (dReal)(-0.5 * Wd)};
  for (int i = 0; i < 4; i++) {
    dBodyID wheel = scene_add_sphere(sw, (dReal)(0.2 / (4.0 / 3.0 * 3.14159265358979 * R * R * R)), R, wx[i], wy[i], Z - H * (dReal)0.5);
    dQuaternion q;
    dQFromAxisAndAngle(q, 1, 0, 0, (dReal)(3.14159265358979 * 0.5));
    dBodySetQuaternion(wheel, q);
    dJointID j = dJointCreateHinge2(sw.world, 0);
    dJointAttach(j, chassis, wheel);
    const dReal *a = dBodyGetPosition(wheel);
    dJointSetHinge2Anchor(j, a[0], a[1], a[2]);
    dJointSetHinge2Axis1(j, 0, 0, 1);
    dJointSetHinge2Axis2(j, 0, 1, 0);
    dJointSetHinge2Param(j, dParamSuspensionERP, (dReal)0.4);
    dJointSetHinge2Param(j, dParamSuspensionCFM, (dReal)0.8);
    if (i >= 2) { dJointSetHinge2Param(j, dParamLoStop, 0); dJointSetHinge2Param(j, dParamHiStop, 0); }
    else {
      dJointSetHinge2Param(j, dParamVel, rng.uni(-0.5, 0.5)); dJointSetHinge2Param(j, dParamFMax, (dReal)0.2);
      dJointSetHinge2Param(j, dParamLoStop, (dReal)-0.75); dJointSetHinge2Param(j, dParamHiStop, (dReal)0.75);
      dJointSetHinge2Param(j, dParamFudgeFactor, (dReal)0.1);
    }
    dJointSetHinge2Param(j, dParamVel2, (dReal)-2.0); dJointSetHinge2Param(j, dParamFMax2, (dReal)0.1);
    sw.joints.push_back(j);
  }
}

// slider and fixed joints (slider.cpp, fixed.cpp) mixed with ball / hinge joints in one island: a fixed joint
// overwrites the island's shared Info2.erp like a ball does (fixed.cpp:72), sliders with stops, a free motor, a
// motor driven into its stop (force + torque-decoupling side effects, joint.cpp:646-657), a slider and a fixed
// joint attached to the world, one of them with the bodies given in reversed order
static inline void scene_sliders(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0051DE5u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID prev = 0;
  for (int i = 0; i < 7; i++) {
    dBodyID b = scene_add_box(sw, 2, (dReal)0.4, (dReal)0.25, (dReal)0.2, (dReal)(0.6 * i), rng.uni(-0.02, 0.02), (dReal)(1.0 + 0.05 * i));
    dQuaternion q = {1, rng.uni(-0.1, 0.1), rng.uni(-0.1, 0.1), rng.uni(-0.1, 0.1)};
    dBodySetQuaternion(b, q);
    dJointID j;
    if (i == 0) {           // slider to the world, bodies reversed: dJOINT_REVERSE
      j = dJointCreateSlider(sw.world, 0);
      dJointAttach(j, 0, b);
      dJointSetSliderAxis(j, (dReal)0.1, 0, 1);
      dJointSetSliderParam(j, dParamLoStop, (dReal)-0.3); dJointSetSliderParam(j, dParamHiStop, (dReal)0.15);
      dJointSetSliderParam(j, dParamBounce, (dReal)0.4);
    } else if (i == 1 || i == 4) {   // two-body sliders: stops / motor into the stop
      j = dJointCreateSlider(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetSliderAxis(j, 1, (dReal)(i == 4 ? 0.3 : 0), (dReal)0.2);
      dJointSetSliderParam(j, dParamLoStop, (dReal)-0.1); dJointSetSliderParam(j, dParamHiStop, (dReal)0.12);
      if (i == 4) {
        dJointSetSliderParam(j, dParamVel, (dReal)1.0); dJointSetSliderParam(j, dParamFMax, (dReal)8);
        dJointSetSliderParam(j, dParamFudgeFactor, (dReal)0.4); dJointSetSliderParam(j, dParamStopERP, (dReal)0.5);
        dJointSetSliderParam(j, dParamStopCFM, (dReal)1e-3);
      }
    } else if (i == 2) {    // fixed, own erp/cfm
      j = dJointCreateFixed(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetFixed(j);
      dJointSetFixedParam(j, dParamERP, (dReal)0.6); dJointSetFixedParam(j, dParamCFM, (dReal)1e-4);
    } else if (i == 3) {    // hinge after the fixed joint: sees the fixed joint's erp
      j = dJointCreateHinge(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetHingeAnchor(j, (dReal)(0.6 * i - 0.3), 0, (dReal)1.1);
      dJointSetHingeAxis(j, 0, 1, 0);
    } else if (i == 5) {    // free slider motor
      j = dJointCreateSlider(sw.world, 0);
      dJointAttach(j, b, prev);
      dJointSetSliderAxis(j, 0, 1, 0);
      dJointSetSliderParam(j, dParamVel, (dReal)-0.5); dJointSetSliderParam(j, dParamFMax, (dReal)3);
    } else {                // ball, then a second chain end welded to the world below
      j = dJointCreateBall(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetBallAnchor(j, (dReal)(0.6 * i - 0.3), 0, (dReal)1.25);
    }
    sw.joints.push_back(j);
    prev = b;
  }
  // a separate box welded to the world (one-body fixed joint), hit by a falling sphere
  dBodyID wb = scene_add_box(sw, 2, (dReal)0.5, (dReal)0.5, (dReal)0.3, (dReal)-1.5, 0, (dReal)0.8);
  dJointID jf = dJointCreateFixed(sw.world, 0);
  dJointAttach(jf, wb, 0);
  dJointSetFixed(jf);
  sw.joints.push_back(jf);
  scene_add_sphere(sw, 3, (dReal)0.25, (dReal)-1.45, (dReal)0.05, (dReal)2.5);
  scene_add_sphere(sw, 3, (dReal)0.2, (dReal)1.3, (dReal)0.02, (dReal)2.8);
}

// piston, PR and plane2d joints (piston.cpp, pr.cpp, plane2d.cpp): a chain hanging from the world by a reversed
// one-body piston with stops on both limit-motors; two-body pistons / PR joints with prismatic stops, rotoide stops,
// a rotoide motor, a prismatic motor driven into its stop (force + torque-decoupling + nothing for the rotoide);
// one body kept in the plane z = const by a plane2d joint with all three motors, one with none
static inline void scene_pistons(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0915704u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID prev = 0;
  for (int i = 0; i < 7; i++) {
    dBodyID b = scene_add_box(sw, 2, (dReal)0.4, (dReal)0.25, (dReal)0.2, (dReal)(0.6 * i), rng.uni(-0.02, 0.02), (dReal)(1.2 + 0.05 * i));
    dQuaternion q = {1, rng.uni(-0.1, 0.1), rng.uni(-0.1, 0.1), rng.uni(-0.1, 0.1)};
    dBodySetQuaternion(b, q);
    dJointID j;
    if (i == 0) {           // piston to the world, bodies reversed
      j = dJointCreatePiston(sw.world, 0);
      dJointAttach(j, 0, b);
      dJointSetPistonAnchor(j, 0, 0, (dReal)1.5);
      dJointSetPistonAxis(j, (dReal)0.1, 0, 1);
      dJointSetPistonParam(j, dParamLoStop, (dReal)-0.2); dJointSetPistonParam(j, dParamHiStop, (dReal)0.1);
      dJointSetPistonParam(j, dParamBounce, (dReal)0.3);
      dJointSetPistonParam(j, dParamLoStop2, (dReal)-0.4); dJointSetPistonParam(j, dParamHiStop2, (dReal)0.3);
    } else if (i == 1 || i == 4) {   // two-body pistons: stops / prismatic motor into its stop + rotoide motor
      j = dJointCreatePiston(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetPistonAnchor(j, (dReal)(0.6 * i - 0.3), (dReal)0.05, (dReal)1.3);
      dJointSetPistonAxis(j, 1, (dReal)(i == 4 ? 0.3 : 0), (dReal)0.2);
      dJointSetPistonParam(j, dParamLoStop, (dReal)-0.1); dJointSetPistonParam(j, dParamHiStop, (dReal)0.12);
      if (i == 4) {
        dJointSetPistonParam(j, dParamVel, (dReal)1.0); dJointSetPistonParam(j, dParamFMax, (dReal)8);
        dJointSetPistonParam(j, dParamFudgeFactor, (dReal)0.4); dJointSetPistonParam(j, dParamStopERP, (dReal)0.5);
        dJointSetPistonParam(j, dParamStopCFM, (dReal)1e-3);
        dJointSetPistonParam(j, dParamVel2, (dReal)-0.8); dJointSetPistonParam(j, dParamFMax2, (dReal)2);
        dJointSetPistonParam(j, dParamLoStop2, (dReal)-0.15); dJointSetPistonParam(j, dParamHiStop2, (dReal)0.2);   // rotoide motor into its stop
        dJointSetPistonParam(j, dParamFudgeFactor2, (dReal)0.7);
      } else {
        dJointSetPistonParam(j, dParamLoStop2, (dReal)-0.05); dJointSetPistonParam(j, dParamHiStop2, (dReal)0.05);
        dJointSetPistonParam(j, dParamBounce2, (dReal)0.5);
      }
    } else if (i == 2 || i == 5) {   // PR: prismatic axis and rotoide axis differ
      j = dJointCreatePR(sw.world, 0);
      if (i == 5) dJointAttach(j, b, prev); else dJointAttach(j, prev, b);
      dJointSetPRAnchor(j, (dReal)(0.6 * i - 0.3), 0, (dReal)1.35);
      dJointSetPRAxis1(j, (dReal)0.2, 0, 1);
      dJointSetPRAxis2(j, 0, 1, (dReal)(i == 5 ? 0.1 : 0));
      dJointSetPRParam(j, dParamLoStop, (dReal)-0.08); dJointSetPRParam(j, dParamHiStop, (dReal)0.06);
      if (i == 5) {
        dJointSetPRParam(j, dParamVel, (dReal)-0.6); dJointSetPRParam(j, dParamFMax, (dReal)5);
        dJointSetPRParam(j, dParamLoStop2, (dReal)-0.2); dJointSetPRParam(j, dParamHiStop2, (dReal)0.25);
      } else {
        dJointSetPRParam(j, dParamVel2, (dReal)0.7); dJointSetPRParam(j, dParamFMax2, (dReal)1.5);
      }
    } else if (i == 3) {    // hinge in between
      j = dJointCreateHinge(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetHingeAnchor(j, (dReal)(0.6 * i - 0.3), 0, (dReal)1.3);
      dJointSetHingeAxis(j, 0, 1, 0);
    } else {                // ball
      j = dJointCreateBall(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetBallAnchor(j, (dReal)(0.6 * i - 0.3), 0, (dReal)1.45);
    }
    sw.joints.push_back(j);
    prev = b;
  }
  // PR to the world, reversed, with a prismatic motor
  {
    dBodyID b = scene_add_box(sw, 2, (dReal)0.3, (dReal)0.3, (dReal)0.3, (dReal)-1.2, (dReal)0.8, (dReal)1.0);
    dJointID j = dJointCreatePR(sw.world, 0);
    dJointAttach(j, 0, b);
    dJointSetPRAnchor(j, (dReal)-1.2, (dReal)0.8, (dReal)1.4);
    dJointSetPRAxis1(j, 0, 0, 1);
    dJointSetPRAxis2(j, 1, 0, 0);
    dJointSetPRParam(j, dParamLoStop, (dReal)-0.3); dJointSetPRParam(j, dParamHiStop, (dReal)0.05);
    dJointSetPRParam(j, dParamVel, (dReal)0.4); dJointSetPRParam(j, dParamFMax, (dReal)30);
    sw.joints.push_back(j);
  }
  // plane2d: two boxes sliding on the plane z = 0 ... they rest on the ground plane geom; spheres drop on them
  for (int k = 0; k < 2; k++) {
    dBodyID b = scene_add_box(sw, 2, (dReal)0.5, (dReal)0.4, (dReal)0.3, (dReal)(-1.5 + 0.9 * k), (dReal)-0.9, (dReal)0.16);
    dBodySetLinearVel(b, (dReal)(0.5 - k), (dReal)0.3, 0);
    dBodySetAngularVel(b, 0, 0, (dReal)(1.0 + k));
    dJointID j = dJointCreatePlane2D(sw.world, 0);
    dJointAttach(j, b, 0);
    if (k == 0) {
      dJointSetPlane2DXParam(j, dParamVel, (dReal)0.3); dJointSetPlane2DXParam(j, dParamFMax, (dReal)4);
      dJointSetPlane2DYParam(j, dParamVel, (dReal)-0.2); dJointSetPlane2DYParam(j, dParamFMax, (dReal)2);
      dJointSetPlane2DAngleParam(j, dParamVel, (dReal)0.9); dJointSetPlane2DAngleParam(j, dParamFMax, (dReal)1);
    } else {
      dJointSetPlane2DYParam(j, dParamVel, (dReal)0.5); dJointSetPlane2DYParam(j, dParamFMax, (dReal)3);
    }
    sw.joints.push_back(j);
    scene_add_sphere(sw, 3, (dReal)0.15, (dReal)(-1.45 + 0.9 * k), (dReal)-0.85, (dReal)(1.5 + 0.4 * k));
  }
  scene_add_sphere(sw, 3, (dReal)0.2, (dReal)1.3, (dReal)0.02, (dReal)2.8);
}

// PU joints (pu.cpp: prismatic + universal): a chain hanging from the world by a reversed one-body PU; two-body PUs
// with prismatic stops, universal-axis stops, motors on each of the three limit-motors (one driven into its stop),
// bodies given in both orders; a hinge and a ball in between so the shared Info2.erp varies
static inline void scene_pus(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0009055u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID prev = 0;
  for (int i = 0; i < 7; i++) {
    dBodyID b = scene_add_box(sw, 2, (dReal)0.4, (dReal)0.25, (dReal)0.2, (dReal)(0.6 * i), rng.uni(-0.02, 0.02), (dReal)(1.2 + 0.05 * i));
    dQuaternion q = {1, rng.uni(-0.1, 0.1), rng.uni(-0.1, 0.1), rng.uni(-0.1, 0.1)};
    dBodySetQuaternion(b, q);
    dJointID j;
    if (i == 3) {
      j = dJointCreateHinge(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetHingeAnchor(j, (dReal)(0.6 * i - 0.3), 0, (dReal)1.3);
      dJointSetHingeAxis(j, 0, 1, 0);
    } else if (i == 6) {
      j = dJointCreateBall(sw.world, 0);
      dJointAttach(j, prev, b);
      dJointSetBallAnchor(j, (dReal)(0.6 * i - 0.3), 0, (dReal)1.45);
    } else {
      j = dJointCreatePU(sw.world, 0);
      if (i == 0) dJointAttach(j, 0, b);            // world, reversed
      else if (i == 5) dJointAttach(j, b, prev);
      else dJointAttach(j, prev, b);
      dJointSetPUAnchor(j, (dReal)(0.6 * i - 0.3), (dReal)0.03, (dReal)(i == 0 ? 1.6 : 1.3));
      dJointSetPUAxis1(j, 0, 1, (dReal)0.1);
      dJointSetPUAxis2(j, (dReal)0.1, 0, 1);
      dJointSetPUAxisP(j, 1, (dReal)(i == 4 ? 0.2 : 0), (dReal)(i == 0 ? 0.5 : 0.1));
      dJointSetPUParam(j, dParamLoStop3, (dReal)-0.1); dJointSetPUParam(j, dParamHiStop3, (dReal)0.08);
      if (i == 0) {
        dJointSetPUParam(j, dParamBounce3, (dReal)0.3);
        dJointSetPUParam(j, dParamLoStop1, (dReal)-0.3); dJointSetPUParam(j, dParamHiStop1, (dReal)0.2);
      } else if (i == 1) {
        dJointSetPUParam(j, dParamLoStop1, (dReal)-0.1); dJointSetPUParam(j, dParamHiStop1, (dReal)0.1);
        dJointSetPUParam(j, dParamLoStop2, (dReal)-0.15); dJointSetPUParam(j, dParamHiStop2, (dReal)0.05);
        dJointSetPUParam(j, dParamBounce2, (dReal)0.4);
      } else if (i == 2) {     // prismatic motor driven into its stop
        dJointSetPUParam(j, dParamVel3, (dReal)0.9); dJointSetPUParam(j, dParamFMax3, (dReal)7);
        dJointSetPUParam(j, dParamFudgeFactor3, (dReal)0.5); dJointSetPUParam(j, dParamStopERP3, (dReal)0.4);
      } else if (i == 4) {     // universal motors, axis 1 into its stop
        dJointSetPUParam(j, dParamVel1, (dReal)-0.7); dJointSetPUParam(j, dParamFMax1, (dReal)1.5);
        dJointSetPUParam(j, dParamLoStop1, (dReal)-0.2); dJointSetPUParam(j, dParamHiStop1, (dReal)0.2);
        dJointSetPUParam(j, dParamVel2, (dReal)0.5); dJointSetPUParam(j, dParamFMax2, (dReal)0.8);
      } else {                 // i == 5: reversed body order, stops on axis 2 + free prismatic motor
        dJointSetPUParam(j, dParamLoStop2, (dReal)-0.25); dJointSetPUParam(j, dParamHiStop2, (dReal)0.15);
        dJointSetPUParam(j, dParamVel3, (dReal)-0.3); dJointSetPUParam(j, dParamFMax3, (dReal)2);
        dJointSetPUParam(j, dParamLoStop3, -dInfinity); dJointSetPUParam(j, dParamHiStop3, dInfinity);
      }
    }
    sw.joints.push_back(j);
    prev = b;
  }
  scene_add_sphere(sw, 3, (dReal)0.25, (dReal)0.3, (dReal)0.05, (dReal)2.5);
  scene_add_sphere(sw, 3, (dReal)0.2, (dReal)2.3, (dReal)0.02, (dReal)2.8);
}

// kinematic bodies (dBodySetKinematic: zero inverse mass / inertia, ode.cpp:841-846): a kinematic slab sweeping through
// a loose pile at constant velocity while spinning slowly, a kinematic sphere dropping at constant speed onto a box, a
// hinge between a kinematic and a dynamic body; one body is switched back with dBodySetDynamic
static inline void scene_kinematic(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x00C1AEu);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID slab = scene_add_box(sw, 2, (dReal)0.3, (dReal)1.6, (dReal)0.5, (dReal)-1.6, 0, (dReal)0.3);
  dBodySetKinematic(slab);
  dBodySetLinearVel(slab, (dReal)0.8, 0, 0);
  dBodySetAngularVel(slab, 0, 0, (dReal)0.3);
  for (int i = 0; i < 10; i++) scene_add_box(sw, 2, rng.uni(0.2, 0.5), rng.uni(0.2, 0.5), rng.uni(0.2, 0.5), rng.uni(-0.8, 0.8), rng.uni(-0.6, 0.6), (dReal)(0.3 + 0.45 * (i % 3)));
  for (int i = 0; i < 4; i++) scene_add_sphere(sw, 2, rng.uni(0.15, 0.3), rng.uni(-0.8, 0.8), rng.uni(-0.6, 0.6), rng.uni(1.6, 2.2));
  dBodyID ball = scene_add_sphere(sw, 3, (dReal)0.2, (dReal)0.1, (dReal)0.05, (dReal)2.8);
  dBodySetKinematic(ball);
  dBodySetLinearVel(ball, 0, 0, (dReal)-0.9);
  dBodyID arm = scene_add_box(sw, 2, (dReal)0.6, (dReal)0.1, (dReal)0.1, (dReal)-1.3, 0, (dReal)0.9);
  dJointID j = dJointCreateHinge(sw.world, 0);
  dJointAttach(j, slab, arm);
  dJointSetHingeAnchor(j, (dReal)-1.6, 0, (dReal)0.9);
  dJointSetHingeAxis(j, 0, 1, 0);
  sw.joints.push_back(j);
  dBodyID back = scene_add_box(sw, 2, (dReal)0.3, (dReal)0.3, (dReal)0.3, (dReal)1.5, (dReal)1.2, (dReal)1.5);
  dBodySetKinematic(back);
  dBodySetDynamic(back);
}

// null joints (joints/null.cpp): no constraint rows, but a null joint puts its two bodies into ONE island, which changes
// the island order, hence the order of the dRandInt draws and every later bit: two separate piles tied by a null joint
static inline void scene_nulljoint(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0000A11u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID first[2] = {0, 0};
  for (int pile = 0; pile < 2; pile++)
    for (int i = 0; i < 6; i++) {
      dBodyID b = scene_add_box(sw, 2, rng.uni(0.2, 0.5), rng.uni(0.2, 0.5), rng.uni(0.2, 0.5), (dReal)(-2 + 4 * pile) + rng.uni(-0.2, 0.2), rng.uni(-0.2, 0.2), (dReal)(0.3 + 0.5 * i));
      if (i == 2) first[pile] = b;
    }
  dJointID j = dJointCreateNull(sw.world, 0);
  dJointAttach(j, first[0], first[1]);
  sw.joints.push_back(j);
  dJointID j1 = dJointCreateNull(sw.world, 0);   // and one to the world
  dJointAttach(j1, sw.bodies[0], 0);
  sw.joints.push_back(j1);
  for (int i = 0; i < 3; i++) scene_add_sphere(sw, 2, rng.uni(0.15, 0.3), rng.uni(-0.5, 0.5), (dReal)1.5, rng.uni(0.5, 2));
}

// universal joints (universal.cpp): free, with stops on both axes (getAngles: dRFrom2Axes + dQfromR + atan2),
// with a motor on axis 2, attached to the world, and one with the bodies given in reversed order
static inline void scene_universals(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x00041E5u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID prev = 0;
  for (int i = 0; i < 6; i++) {
    dBodyID b = scene_add_box(sw, 2, (dReal)0.45, (dReal)0.2, (dReal)0.15, (dReal)(0.55 * i), rng.uni(-0.02, 0.02), (dReal)1.4);
    dBodySetAngularVel(b, rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1));
    dJointID j = dJointCreateUniversal(sw.world, 0);
    if (i == 3) dJointAttach(j, b, prev); else if (i == 0) dJointAttach(j, 0, b); else dJointAttach(j, prev, b);
    dJointSetUniversalAnchor(j, (dReal)(0.55 * i - 0.275), 0, (dReal)1.4);
    dJointSetUniversalAxis1(j, 0, 1, (dReal)(i == 2 ? 0.3 : 0));
    dJointSetUniversalAxis2(j, 0, (dReal)(i == 4 ? 0.2 : 0), 1);
    if (i == 1 || i == 3 || i == 5) {
      dJointSetUniversalParam(j, dParamLoStop, (dReal)-0.3); dJointSetUniversalParam(j, dParamHiStop, (dReal)0.25);
      dJointSetUniversalParam(j, dParamLoStop2, (dReal)-0.2); dJointSetUniversalParam(j, dParamHiStop2, (dReal)0.35);
      if (i == 5) dJointSetUniversalParam(j, dParamBounce, (dReal)0.3);
    }
    if (i == 2) { dJointSetUniversalParam(j, dParamVel2, (dReal)1.2); dJointSetUniversalParam(j, dParamFMax2, (dReal)2); }
    if (i == 3) {   // powered into its stop
      dJointSetUniversalParam(j, dParamVel, (dReal)2.0); dJointSetUniversalParam(j, dParamFMax, (dReal)3);
      dJointSetUniversalParam(j, dParamFudgeFactor, (dReal)0.5);
    }
    sw.joints.push_back(j);
    prev = b;
  }
  scene_add_sphere(sw, 3, (dReal)0.2, (dReal)1.0, (dReal)0.03, (dReal)2.6);
}

// angular / linear motors (amotor.cpp, lmotor.cpp): an Euler-mode amotor with stops next to a ball joint (the
// ragdoll idiom), a user-mode amotor to the world whose user-set angle sits beyond its stop while powered
// (motor-at-limit torque side effect on up to three axes), lmotors with axes relative to the world, body 1 and body 2
static inline void scene_motors(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0A4070u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  dBodyID b[5];
  for (int i = 0; i < 5; i++) {
    b[i] = scene_add_box(sw, 2, (dReal)0.4, (dReal)0.2, (dReal)0.2, (dReal)(0.5 * i), rng.uni(-0.02, 0.02), (dReal)1.3);
    dBodySetAngularVel(b[i], rng.uni(-1.5, 1.5), rng.uni(-1.5, 1.5), rng.uni(-1.5, 1.5));
  }
  // b0-b1: ball + Euler amotor with stops
  dJointID j = dJointCreateBall(sw.world, 0);
  dJointAttach(j, b[0], b[1]);
  dJointSetBallAnchor(j, (dReal)0.25, 0, (dReal)1.3);
  sw.joints.push_back(j);
  j = dJointCreateAMotor(sw.world, 0);
  dJointAttach(j, b[0], b[1]);
  dJointSetAMotorMode(j, dAMotorEuler);
  dJointSetAMotorAxis(j, 0, 1, 1, 0, 0);
  dJointSetAMotorAxis(j, 2, 2, 0, 0, 1);
  dJointSetAMotorParam(j, dParamLoStop, (dReal)-0.3); dJointSetAMotorParam(j, dParamHiStop, (dReal)0.3);
  dJointSetAMotorParam(j, dParamLoStop2, (dReal)-0.2); dJointSetAMotorParam(j, dParamHiStop2, (dReal)0.25);
  dJointSetAMotorParam(j, dParamLoStop3, (dReal)-0.4); dJointSetAMotorParam(j, dParamHiStop3, (dReal)0.15);
  dJointSetAMotorParam(j, dParamBounce2, (dReal)0.3);
  sw.joints.push_back(j);
  // b1-b2: hinge + 3-axis lmotor (axes relative to world / body 1 / body 2)
  j = dJointCreateHinge(sw.world, 0);
  dJointAttach(j, b[1], b[2]);
  dJointSetHingeAnchor(j, (dReal)0.75, 0, (dReal)1.3);
  dJointSetHingeAxis(j, 0, 1, 0);
  sw.joints.push_back(j);
  j = dJointCreateLMotor(sw.world, 0);
  dJointAttach(j, b[1], b[2]);
  dJointSetLMotorNumAxes(j, 3);
  dJointSetLMotorAxis(j, 0, 0, 1, 0, 0);
  dJointSetLMotorAxis(j, 1, 1, 0, 1, (dReal)0.2);
  dJointSetLMotorAxis(j, 2, 2, 0, 0, 1);
  dJointSetLMotorParam(j, dParamVel, (dReal)0.2); dJointSetLMotorParam(j, dParamFMax, (dReal)1.5);
  dJointSetLMotorParam(j, dParamVel3, (dReal)-0.1); dJointSetLMotorParam(j, dParamFMax3, (dReal)0.8);
  sw.joints.push_back(j);
  // b3 - world: user-mode amotor, powered, angle beyond the stop on axis 0, plain motor on axis 1
  j = dJointCreateAMotor(sw.world, 0);
  dJointAttach(j, b[3], 0);
  dJointSetAMotorNumAxes(j, 2);
  dJointSetAMotorAxis(j, 0, 1, 0, 0, 1);
  dJointSetAMotorAxis(j, 1, 0, 1, 0, 0);
  dJointSetAMotorAngle(j, 0, (dReal)0.5);
  dJointSetAMotorParam(j, dParamLoStop, (dReal)-0.3); dJointSetAMotorParam(j, dParamHiStop, (dReal)0.3);
  dJointSetAMotorParam(j, dParamVel, (dReal)1.0); dJointSetAMotorParam(j, dParamFMax, (dReal)2.0);
  dJointSetAMotorParam(j, dParamFudgeFactor, (dReal)0.5);
  dJointSetAMotorParam(j, dParamVel2, (dReal)-2.0); dJointSetAMotorParam(j, dParamFMax2, (dReal)1.0);
  sw.joints.push_back(j);
  // b4 - world: one-body lmotor holding the box up against gravity along z, free otherwise
  j = dJointCreateLMotor(sw.world, 0);
  dJointAttach(j, b[4], 0);
  dJointSetLMotorNumAxes(j, 1);
  dJointSetLMotorAxis(j, 0, 0, 0, 0, 1);
  dJointSetLMotorParam(j, dParamVel, (dReal)0.3); dJointSetLMotorParam(j, dParamFMax, (dReal)1.0);
  sw.joints.push_back(j);
  scene_add_sphere(sw, 3, (dReal)0.2, (dReal)0.6, (dReal)0.03, (dReal)2.6);
}

static inline dBodyID scene_add_capsule(SceneWorld &sw, dReal density, dReal r, dReal l, dReal x, dReal y, dReal z) {
  dBodyID b = dBodyCreate(sw.world);
  dBodySetPosition(b, x, y, z);
  dMass m;
  dMassSetCapsule(&m, density, 3, r, l);
  dBodySetMass(b, &m);
  dGeomID g = scene_add_geom(sw, dCreateCapsule(sw.space, r, l));
  dGeomSetBody(g, b);
  sw.bodies.push_back(b);
  return b;
}

static inline dBodyID scene_add_cylinder(SceneWorld &sw, dReal density, dReal r, dReal l, dReal x, dReal y, dReal z) {
  dBodyID b = dBodyCreate(sw.world);
  dBodySetPosition(b, x, y, z);
  dMass m;
  dMassSetCylinder(&m, density, 3, r, l);
  dBodySetMass(b, &m);
  dGeomID g = scene_add_geom(sw, dCreateCylinder(sw.space, r, l));
  dGeomSetBody(g, b);
  sw.bodies.push_back(b);
  return b;
}

// flat cylinder colliders (collision_cylinder_plane.cpp, collision_cylinder_sphere.cpp): random cylinders and spheres
// tumbling onto a plane, two cylinders standing exactly upright (the axis-parallel four-point branch), spheres dropped
// on discs, rims and mantles.  Cylinder-cylinder pairs have no collider in the reference build (no libccd): they
// reach the near callback and produce no contacts, here too.  `boxes`: also boxes (collision_cylinder_box.cpp)
static inline void scene_cylmix_impl(SceneWorld &sw, int w, bool boxes) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x00C71u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int i = 0; i < 8; i++) {
    dBodyID b = scene_add_cylinder(sw, 2, rng.uni(0.1, 0.3), rng.uni(0.1, 0.8), rng.uni(-0.8, 0.8), rng.uni(-0.8, 0.8), rng.uni(0.4, 3.5));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dBodySetAngularVel(b, rng.uni(-2, 2), rng.uni(-2, 2), rng.uni(-2, 2));
  }
  for (int i = 0; i < 2; i++) scene_add_cylinder(sw, 2, (dReal)0.3, (dReal)0.4, (dReal)(1.6 + 0.9 * i), (dReal)1.5, (dReal)(0.19 + 0.4 * i));   // upright
  for (int i = 0; i < 6; i++) scene_add_sphere(sw, 2, rng.uni(0.12, 0.35), rng.uni(-0.8, 0.8), rng.uni(-0.8, 0.8), rng.uni(0.4, 3.5));
  scene_add_sphere(sw, 3, (dReal)0.15, (dReal)1.62, (dReal)1.53, (dReal)1.2);     // onto the upright cylinder's disc
  scene_add_sphere(sw, 3, (dReal)0.2, (dReal)(2.5 + 0.31), (dReal)1.5, (dReal)1.6);   // onto a rim
  if (boxes) {
    for (int i = 0; i < 5; i++) {
      dBodyID b = scene_add_box(sw, 2, rng.uni(0.2, 0.7), rng.uni(0.2, 0.7), rng.uni(0.2, 0.7), rng.uni(-0.8, 0.8), rng.uni(-0.8, 0.8), rng.uni(0.4, 3.5));
      dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
      dBodySetQuaternion(b, q);
    }
    scene_add_box(sw, 2, (dReal)1.2, (dReal)1.2, (dReal)0.3, (dReal)-2.0, (dReal)-1.5, (dReal)0.16);        // a slab ...
    scene_add_cylinder(sw, 2, (dReal)0.25, (dReal)0.5, (dReal)-2.0, (dReal)-1.5, (dReal)0.58);              // ... with a cylinder standing on it
    dBodyID lying = scene_add_cylinder(sw, 2, (dReal)0.2, (dReal)0.9, (dReal)-1.9, (dReal)-1.4, (dReal)1.2);  // and one lying across
    dMatrix3 R;
    dRFromAxisAndAngle(R, 0, 1, 0, (dReal)(3.14159265358979 * 0.5));
    dBodySetRotation(lying, R);
  }
}
// geom transforms (collision_transform.cpp): composite bodies in the manner of demo_boxstack's 'x' objects - three
// transform geoms per body encapsulating a sphere, a box and a capsule or cylinder at their own local poses, mass
// assembled with dMassRotate / dMassTranslate / dMassAdd and recentred - tumbling onto a plane among plain boxes and
// spheres, plus a static transform (no body) around a tilted box.  Covers T x X, X x T and T x T collider order.
// The encapsulated geoms carry their transform's data id, so traces name the same geom in every mode.
static inline dGeomID scene_add_xf(SceneWorld &sw, dGeomID inner, int info) {
  dGeomID t = scene_add_geom(sw, dCreateGeomTransform(sw.space));
  dGeomTransformSetCleanup(t, 1);
  dGeomTransformSetInfo(t, info);
  dGeomTransformSetGeom(t, inner);
  dGeomSetData(inner, dGeomGetData(t));
  return t;
}
static inline void scene_transforms(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x7F0A3u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int i = 0; i < 7; i++) {
    dBodyID b = dBodyCreate(sw.world);
    dBodySetPosition(b, rng.uni(-1.2, 1.2), rng.uni(-1.2, 1.2), (dReal)(0.8 + 0.75 * i));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dMass m, m2;
    dMassSetZero(&m);
    dGeomID inner[3], t[3];
    dReal dpos[3][3];
    for (int k = 0; k < 3; k++) for (int e = 0; e < 3; e++) dpos[k][e] = rng.uni(-0.3, 0.3);
    for (int k = 0; k < 3; k++) {
      if (k == 0) {
        dReal r = rng.uni(0.08, 0.3);
        inner[k] = dCreateSphere(0, r);
        dMassSetSphere(&m2, 2, r);
      } else if (k == 1) {
        dReal lx = rng.uni(0.15, 0.6), ly = rng.uni(0.15, 0.6), lz = rng.uni(0.15, 0.6);
        inner[k] = dCreateBox(0, lx, ly, lz);
        dMassSetBox(&m2, 2, lx, ly, lz);
      } else if (i & 1) {
        dReal r = rng.uni(0.06, 0.15), l = rng.uni(0.2, 0.9);
        inner[k] = dCreateCapsule(0, r, l);
        dMassSetCapsule(&m2, 2, 3, r, l);
      } else {
        dReal r = rng.uni(0.1, 0.25), l = rng.uni(0.15, 0.6);
        inner[k] = dCreateCylinder(0, r, l);
        dMassSetCylinder(&m2, 2, 3, r, l);
      }
      t[k] = scene_add_xf(sw, inner[k], (i + k) % 3 == 0);
      dGeomSetPosition(inner[k], dpos[k][0], dpos[k][1], dpos[k][2]);
      dMatrix3 Rtx;
      dRFromAxisAndAngle(Rtx, rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-5, 5));
      dGeomSetRotation(inner[k], Rtx);
      dMassRotate(&m2, Rtx);
      dMassTranslate(&m2, dpos[k][0], dpos[k][1], dpos[k][2]);
      dMassAdd(&m, &m2);
    }
    // move all encapsulated objects so that the centre of mass is (0,0,0), as the demo does
    for (int k = 0; k < 3; k++) dGeomSetPosition(inner[k], dpos[k][0] - m.c[0], dpos[k][1] - m.c[1], dpos[k][2] - m.c[2]);
    dMassTranslate(&m, -m.c[0], -m.c[1], -m.c[2]);
    for (int k = 0; k < 3; k++) dGeomSetBody(t[k], b);
    dBodySetMass(b, &m);
    sw.bodies.push_back(b);
  }
  for (int i = 0; i < 4; i++) {
    dBodyID b = scene_add_box(sw, 2, rng.uni(0.2, 0.6), rng.uni(0.2, 0.6), rng.uni(0.2, 0.6), rng.uni(-1.0, 1.0), rng.uni(-1.0, 1.0), rng.uni(0.4, 4.0));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
  }
  for (int i = 0; i < 4; i++) scene_add_sphere(sw, 2, rng.uni(0.12, 0.3), rng.uni(-1.0, 1.0), rng.uni(-1.0, 1.0), rng.uni(0.4, 4.0));
  // static ramp: a transform without a body, itself placed and rotated, around a box with its own local pose
  dGeomID rb = dCreateBox(0, (dReal)1.6, (dReal)1.2, (dReal)0.2);
  dGeomID rt = scene_add_xf(sw, rb, 1);
  dMatrix3 R1, R2;
  dRFromAxisAndAngle(R1, 0, 1, 0, (dReal)0.35);
  dGeomSetRotation(rb, R1);
  dGeomSetPosition(rb, (dReal)0.1, (dReal)-0.05, (dReal)0.2);
  dRFromAxisAndAngle(R2, 0, 0, 1, (dReal)0.6);
  dGeomSetRotation(rt, R2);
  dGeomSetPosition(rt, (dReal)0.3, (dReal)0.2, (dReal)0.15);
}

// the same with ray geoms in the space (drop-in path only): ray x transform and transform x ray in both callback orders
static inline void scene_transforms_rays(SceneWorld &sw, int w) {
  scene_transforms(sw, w);
  xs32 rng(sw.seed ^ 0x5A17Eu);
  for (int i = 0; i < 8; i++) {
    dGeomID r = scene_add_geom(sw, dCreateRay(sw.space, rng.uni(2, 7)));
    if (i < 5) dGeomRaySet(r, rng.uni(-1.2, 1.2), rng.uni(-1.2, 1.2), (dReal)6, rng.uni(-0.2, 0.2), rng.uni(-0.2, 0.2), -1);
    else dGeomRaySet(r, (dReal)-3, rng.uni(-1, 1), rng.uni(0.1, 0.6), 1, rng.uni(-0.2, 0.2), rng.uni(-0.05, 0.1));
    if (i & 1) dGeomRaySetClosestHit(r, 1);
  }
}

static inline void scene_cylspheres(SceneWorld &sw, int w) { scene_cylmix_impl(sw, w, false); }
static inline void scene_cylmix(SceneWorld &sw, int w) { scene_cylmix_impl(sw, w, true); }

// capsule collider coverage: random capsules / boxes / spheres tumbling onto a plane, plus two
// pairs of exactly parallel capsules (the two-contact branch of capsule.cpp:262-316)
static inline void scene_capsmix(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0CA95017u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int i = 0; i < 10; i++) {
    dBodyID b = scene_add_capsule(sw, 2, rng.uni(0.08, 0.25), rng.uni(0.1, 0.8), rng.uni(-0.8, 0.8), rng.uni(-0.8, 0.8), rng.uni(0.4, 3.5));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dBodySetAngularVel(b, rng.uni(-2, 2), rng.uni(-2, 2), rng.uni(-2, 2));
  }
  for (int i = 0; i < 4; i++) {
    dBodyID b = scene_add_box(sw, 2, rng.uni(0.2, 0.7), rng.uni(0.2, 0.7), rng.uni(0.2, 0.7), rng.uni(-0.8, 0.8), rng.uni(-0.8, 0.8), rng.uni(0.4, 3.5));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
  }
  for (int i = 0; i < 4; i++) scene_add_sphere(sw, 2, rng.uni(0.15, 0.4), rng.uni(-0.8, 0.8), rng.uni(-0.8, 0.8), rng.uni(0.4, 3.5));
  for (int i = 0; i < 2; i++) {   // parallel pairs, axis along x, stacked with a small overlap
    dMatrix3 R;
    dRFromAxisAndAngle(R, 0, 1, 0, (dReal)(3.14159265358979 * 0.5));
    dBodyID a = scene_add_capsule(sw, 2, (dReal)0.15, (dReal)0.6, (dReal)(2.5 + i), (dReal)2.5, (dReal)0.15);
    dBodyID b = scene_add_capsule(sw, 2, (dReal)0.15, (dReal)0.4, (dReal)(2.55 + i), (dReal)2.5, (dReal)0.44);
    dBodySetRotation(a, R);
    dBodySetRotation(b, R);
  }
}

// config 4: 20-link ragdoll-like chain (capsule / box links alternating, 19 joints alternating
// ball / hinge with +-1 rad stops) laid out as a serpentine and dropped onto a 4x4 pile of boxes
// resting on a plane; crash contact policy (demo_crash.cpp:128-137)
static inline void scene_ragdoll(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x00D011u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++)
      scene_add_box(sw, 2, (dReal)0.5, (dReal)0.5, rng.uni(0.3, 0.5), (dReal)((i - 1.5) * 0.56) + rng.uni(-0.02, 0.02),
                    (dReal)((j - 1.5) * 0.56) + rng.uni(-0.02, 0.02), (dReal)0.26);
  const dReal L = (dReal)0.3;
  dBodyID prev = 0;
  dReal px = 0, py = 0, pz = 0;
  for (int k = 0; k < 20; k++) {
    const int r = k / 5, c = k % 5;
    const dReal x = (dReal)(((r & 1) ? 4 - c : c) * 0.3 - 0.6) + rng.uni(-0.005, 0.005);
    const dReal y = (dReal)(r * 0.3 - 0.45) + rng.uni(-0.005, 0.005);
    const dReal z = (dReal)(1.2 + 0.03 * k);
    dBodyID b;
    if (k & 1) b = scene_add_box(sw, 2, (dReal)0.26, (dReal)0.12, (dReal)0.1, x, y, z);
    else {
      b = scene_add_capsule(sw, 2, (dReal)0.06, (dReal)0.14, x, y, z);
      dMatrix3 R;
      dRFromAxisAndAngle(R, 0, 1, 0, (dReal)(3.14159265358979 * 0.5));   // capsule axis (local z) along x
      dBodySetRotation(b, R);
    }
    dBodySetAngularVel(b, rng.uni(-0.5, 0.5), rng.uni(-0.5, 0.5), rng.uni(-0.5, 0.5));
    if (prev) {
      const dReal ax = (px + x) * (dReal)0.5, ay = (py + y) * (dReal)0.5, az = (pz + z) * (dReal)0.5;
      dJointID j;
      if (k & 1) {
        j = dJointCreateBall(sw.world, 0);
        dJointAttach(j, prev, b);
        dJointSetBallAnchor(j, ax, ay, az);
      } else {
        j = dJointCreateHinge(sw.world, 0);
        dJointAttach(j, prev, b);
        dJointSetHingeAnchor(j, ax, ay, az);
        if (c == 0) dJointSetHingeAxis(j, 1, 0, 0); else dJointSetHingeAxis(j, 0, 1, 0);
        dJointSetHingeParam(j, dParamLoStop, (dReal)-1.0);
        dJointSetHingeParam(j, dParamHiStop, (dReal)1.0);
      }
      sw.joints.push_back(j);
    }
    prev = b; px = x; py = y; pz = z;
    (void)L;
  }
}

// ---- trimesh scenes ---------------------------------------------------------------------------
// One shared dTriMeshData per (n, spacing) in the process, like config 3's shared terrain: an n x n
// vertex grid, two triangles per cell, height = amp*sin(fx*x)*cos(fy*y) + noise (hashed per vertex).
struct SceneTerrain { int n; double spacing; int pre; dTriMeshDataID data; std::vector<float> verts; std::vector<dTriIndex> idx; };
static inline dTriMeshDataID scene_terrain_data(int n, double spacing, double amp, double fx, double fy, double noise, int preprocess = 0) {
  static std::vector<SceneTerrain *> cache;
  for (size_t i = 0; i < cache.size(); i++) if (cache[i]->n == n && cache[i]->spacing == spacing && cache[i]->pre == preprocess) return cache[i]->data;
  SceneTerrain *t = new SceneTerrain;
  t->n = n; t->spacing = spacing; t->pre = preprocess;
  xs32 rng(7);
  const double half = 0.5 * (n - 1) * spacing;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) {
      const double x = i * spacing - half, y = j * spacing - half;
      const double z = amp * sin(fx * x) * cos(fy * y) + noise * ((rng.next() >> 8) * (1.0 / 16777216.0) - 0.5);
      t->verts.push_back((float)x); t->verts.push_back((float)y); t->verts.push_back((float)z);
    }
  for (int j = 0; j + 1 < n; j++)
    for (int i = 0; i + 1 < n; i++) {
      const dTriIndex a = (dTriIndex)(j * n + i), b = a + 1, c = a + n, d = c + 1;
      t->idx.push_back(a); t->idx.push_back(b); t->idx.push_back(c);
      t->idx.push_back(b); t->idx.push_back(d); t->idx.push_back(c);
    }
  t->data = dGeomTriMeshDataCreate();
  dGeomTriMeshDataBuildSingle(t->data, t->verts.data(), 3 * sizeof(float), n * n, t->idx.data(), (int)t->idx.size(), 3 * sizeof(dTriIndex));
  if (preprocess) dGeomTriMeshDataPreprocess(t->data);   // edge / vertex use flags for the capsule collider
  cache.push_back(t);
  return t->data;
}

// spheres of assorted sizes dropped on a rotated, translated terrain mesh (dCollideSTL + the BVH order)
static inline void scene_terrain_spheres(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x7E44A1u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, (dReal)-3));
  dGeomID mesh = scene_add_geom(sw, dCreateTriMesh(sw.space, scene_terrain_data(25, 0.5, 0.45, 0.9, 0.7, 0.12), 0, 0, 0));
  dMatrix3 R;
  dRFromAxisAndAngle(R, (dReal)0.1, (dReal)-0.05, 1, (dReal)0.6);
  dGeomSetRotation(mesh, R);
  dGeomSetPosition(mesh, (dReal)0.3, (dReal)-0.2, (dReal)0.1);
  for (int i = 0; i < 14; i++) {
    dBodyID b = scene_add_sphere(sw, 2, rng.uni(0.12, 0.7), rng.uni(-4, 4), rng.uni(-4, 4), rng.uni(1.5, 5));
    dBodySetLinearVel(b, rng.uni(-1, 1), rng.uni(-1, 1), 0);
  }
}

// boxes (and a few spheres) tumbling on the terrain mesh: dCollideBTL (all three clipping cases)
static inline void scene_terrain_boxes(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0xB0C5E5u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, (dReal)-3));
  dGeomID mesh = scene_add_geom(sw, dCreateTriMesh(sw.space, scene_terrain_data(25, 0.5, 0.45, 0.9, 0.7, 0.12), 0, 0, 0));
  dGeomSetPosition(mesh, (dReal)-0.1, (dReal)0.25, 0);
  for (int i = 0; i < 10; i++) {
    dBodyID b = scene_add_box(sw, 2, rng.uni(0.2, 1.2), rng.uni(0.2, 1.2), rng.uni(0.2, 0.9), rng.uni(-3.5, 3.5), rng.uni(-3.5, 3.5), rng.uni(1.2, 4));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dBodySetAngularVel(b, rng.uni(-2, 2), rng.uni(-2, 2), rng.uni(-2, 2));
  }
  for (int i = 0; i < 4; i++) scene_add_sphere(sw, 2, rng.uni(0.15, 0.5), rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(1.5, 4));
}

// capsules (and a few boxes) tumbling on the terrain mesh: dCollideCCTL (collision_trimesh_ccylinder.cpp)
static inline void scene_terrain_capsules(SceneWorld &sw, int w, int preprocess) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0xCA95E5u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, (dReal)-3));
  dGeomID mesh = scene_add_geom(sw, dCreateTriMesh(sw.space, scene_terrain_data(25, 0.5, 0.45, 0.9, 0.7, 0.12, preprocess), 0, 0, 0));
  dMatrix3 R;
  dRFromAxisAndAngle(R, (dReal)-0.1, (dReal)0.15, 1, (dReal)-0.5);
  dGeomSetRotation(mesh, R);
  dGeomSetPosition(mesh, (dReal)0.15, (dReal)0.1, (dReal)-0.05);
  for (int i = 0; i < 12; i++) {
    dBodyID b = scene_add_capsule(sw, 2, rng.uni(0.1, 0.35), rng.uni(0.2, 1.2), rng.uni(-3.5, 3.5), rng.uni(-3.5, 3.5), rng.uni(1.2, 4));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dBodySetAngularVel(b, rng.uni(-2, 2), rng.uni(-2, 2), rng.uni(-2, 2));
  }
  for (int i = 0; i < 3; i++) scene_add_box(sw, 2, rng.uni(0.3, 0.8), rng.uni(0.3, 0.8), rng.uni(0.3, 0.8), rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(1.5, 4));
}

// a terrain mesh cut by a tilted plane (dCollideTrimeshPlane: static mesh vertices below the plane -> contact
// joints attached to no body, which the stepper must ignore) with spheres and boxes rolling on both
static inline void scene_terrain_plane(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x91A4E5u);
  scene_add_geom(sw, dCreatePlane(sw.space, (dReal)0.05, (dReal)-0.03, 1, (dReal)0.1));
  dGeomID mesh = scene_add_geom(sw, dCreateTriMesh(sw.space, scene_terrain_data(25, 0.5, 0.45, 0.9, 0.7, 0.12), 0, 0, 0));
  dGeomSetPosition(mesh, (dReal)0.1, (dReal)-0.2, (dReal)0.0);
  for (int i = 0; i < 6; i++) scene_add_sphere(sw, 2, rng.uni(0.2, 0.5), rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(1.5, 3));
  for (int i = 0; i < 4; i++) scene_add_box(sw, 2, rng.uni(0.3, 0.8), rng.uni(0.3, 0.8), rng.uni(0.3, 0.8), rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(1.5, 3));
}

// config 3: demo_buggy-style vehicle (box chassis + 4 sphere wheels on hinge2, demo_buggy.cpp:226-294)
// dropped on a shared trimesh terrain (n x n vertex grid, 1 m spacing,
// 0.5*sin(0.07x)*cos(0.05y) + 0.15*noise); world w spawns on a lattice over the terrain
static inline void scene_buggy_terrain(SceneWorld &sw, int w, int n) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x00C0B066u);
  scene_add_geom(sw, dCreateTriMesh(sw.space, scene_terrain_data(n, 1.0, 0.5, 0.07, 0.05, 0.15), 0, 0, 0));
  const int lat = (n - 8) / 4 > 1 ? (n - 8) / 4 : 1;
  const dReal ox = (dReal)(((w % lat) - 0.5 * (lat - 1)) * 4.0) + rng.uni(-0.5, 0.5);
  const dReal oy = (dReal)((((w / lat) % lat) - 0.5 * (lat - 1)) * 4.0) + rng.uni(-0.5, 0.5);
  const dReal L = (dReal)0.7, Wd = (dReal)0.5, H = (dReal)0.2, R = (dReal)0.18, Z = (dReal)1.1;
  dBodyID chassis = scene_add_box(sw, 1 / (L * Wd * H), L, Wd, H, ox, oy, Z);
  const dReal wx[4] = {(dReal)(0.5 * L), (dReal)(0.5 * L), (dReal)(-0.5 * L), (dReal)(-0.5 * L)};
  const dReal wy[4] = {(dReal)(0.5 * Wd), (dReal)(-0.5 * Wd), (dReal)(0.5 * Wd), (dReal)(-0.5 * Wd)};
  for (int i = 0; i < 4; i++) {
    dBodyID wheel = scene_add_sphere(sw, (dReal)(0.2 / (4.0 / 3.0 * 3.14159265358979 * R * R * R)), R, ox + wx[i], oy + wy[i], Z - H * (dReal)0.5);
    dQuaternion q;
    dQFromAxisAndAngle(q, 1, 0, 0, (dReal)(3.14159265358979 * 0.5));
    dBodySetQuaternion(wheel, q);
    dJointID j = dJointCreateHinge2(sw.world, 0);
    dJointAttach(j, chassis, wheel);
    const dReal *a = dBodyGetPosition(wheel);
    dJointSetHinge2Anchor(j, a[0], a[1], a[2]);
    dJointSetHinge2Axis1(j, 0, 0, 1);
    dJointSetHinge2Axis2(j, 0, 1, 0);
    dJointSetHinge2Param(j, dParamSuspensionERP, (dReal)0.4);
    dJointSetHinge2Param(j, dParamSuspensionCFM, (dReal)0.8);
    if (i >= 2) { dJointSetHinge2Param(j, dParamLoStop, 0); dJointSetHinge2Param(j, dParamHiStop, 0); }
    else {
      dJointSetHinge2Param(j, dParamVel, rng.uni(-0.5, 0.5)); dJointSetHinge2Param(j, dParamFMax, (dReal)0.2);
      dJointSetHinge2Param(j, dParamLoStop, (dReal)-0.75); dJointSetHinge2Param(j, dParamHiStop, (dReal)0.75);
      dJointSetHinge2Param(j, dParamFudgeFactor, (dReal)0.1);
    }
    dJointSetHinge2Param(j, dParamVel2, (dReal)-3.0); dJointSetHinge2Param(j, dParamFMax2, (dReal)0.1);
    sw.joints.push_back(j);
  }
}

// config 5: nx*ny*nz bodies, alternating spheres (r 0.25) and boxes (0.5^3), jittered lattice (1 m in x/y,
// 0.62 m in z) dropped into a walled box (floor + 4 wall planes); dSweepAndPruneSpace, crash policy.
// One world; scene name pile_NXxNYxNZ.
static inline void scene_pile(SceneWorld &sw, int w, int nx, int ny, int nz) {
  g_scene_space_kind = 1;
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0091E5u);
  const dReal hx = (dReal)(0.5 * nx + 0.5), hy = (dReal)(0.5 * ny + 0.5);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  scene_add_geom(sw, dCreatePlane(sw.space, 1, 0, 0, -hx));
  scene_add_geom(sw, dCreatePlane(sw.space, -1, 0, 0, -hx));
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 1, 0, -hy));
  scene_add_geom(sw, dCreatePlane(sw.space, 0, -1, 0, -hy));
  for (int k = 0; k < nz; k++)
    for (int j = 0; j < ny; j++)
      for (int i = 0; i < nx; i++) {
        const dReal x = (dReal)(i - 0.5 * (nx - 1)) + rng.uni(-0.1, 0.1);
        const dReal y = (dReal)(j - 0.5 * (ny - 1)) + rng.uni(-0.1, 0.1);
        const dReal z = (dReal)(0.4 + 0.62 * k) + rng.uni(-0.02, 0.02);
        if ((i + j + k) & 1) scene_add_sphere(sw, 1, (dReal)0.25, x, y, z);
        else scene_add_box(sw, 1, (dReal)0.5, (dReal)0.5, (dReal)0.5, x, y, z);
      }
}

// ray colliders (ray.cpp, collision_trimesh_ray.cpp): bodies of every primitive class tumbling over a plane
// and a rotated terrain mesh, watched by rays — free-standing ones in every mode (all hits / first contact /
// closest hit, with and without backface culling) and "sensor" rays riding on bodies (raycar style,
// RC/car.cpp:353-371).  The near callback records ray contacts and creates no joints for them.
static inline void scene_raycast(SceneWorld &sw, int w, int two_spaces, bool cylinders = false) {
  scene_world_base(sw, w);
  // two_spaces: the rays live in their own space and are collided with dSpaceCollide2 (space x space, geom x space)
  if (two_spaces) sw.space2 = two_spaces == 2 ? dHashSpaceCreate(0) : dSimpleSpaceCreate(0);
  dSpaceID rs = two_spaces ? sw.space2 : sw.space;
  xs32 rng(sw.seed ^ 0x00BA7CA5u);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, (dReal)-1.5));
  dGeomID mesh = scene_add_geom(sw, dCreateTriMesh(sw.space, scene_terrain_data(25, 0.5, 0.45, 0.9, 0.7, 0.12), 0, 0, 0));
  dMatrix3 R;
  dRFromAxisAndAngle(R, (dReal)0.2, (dReal)0.1, 1, (dReal)0.4);
  dGeomSetRotation(mesh, R);
  dGeomSetPosition(mesh, (dReal)0.2, (dReal)-0.1, (dReal)0.05);
  std::vector<dBodyID> carriers;
  for (int i = 0; i < 5; i++) {
    dBodyID b = scene_add_box(sw, 2, rng.uni(0.3, 0.9), rng.uni(0.3, 0.9), rng.uni(0.2, 0.6), rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(1.2, 3));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dBodySetAngularVel(b, rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1));
    carriers.push_back(b);
  }
  for (int i = 0; i < 4; i++) scene_add_sphere(sw, 2, rng.uni(0.2, 0.5), rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(1.5, 3));
  for (int i = 0; i < 4; i++) {
    dBodyID b = scene_add_capsule(sw, 2, rng.uni(0.12, 0.3), rng.uni(0.3, 0.9), rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(1.5, 3));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
  }
  if (cylinders) {   // ray.cpp dCollideRayCylinder: tumbling ones, one standing upright under vertical rays, one lying
    // (static geoms hung in the air: a cylinder BODY would land on the terrain mesh, and cylinder-trimesh is not built)
    for (int i = 0; i < 8; i++) {
      dGeomID c = scene_add_geom(sw, dCreateCylinder(sw.space, rng.uni(0.3, 0.8), rng.uni(0.4, 1.6)));
      dGeomSetPosition(c, rng.uni(-4, 4), rng.uni(-4, 4), rng.uni(0.5, 4));
      dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
      dReal l = (dReal)sqrt((double)(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]));
      for (int k = 0; k < 4; k++) q[k] /= l;
      dGeomSetQuaternion(c, q);
    }
    dGeomID c0 = scene_add_geom(sw, dCreateCylinder(sw.space, (dReal)0.6, (dReal)1.0));
    dGeomSetPosition(c0, (dReal)3.5, (dReal)3.5, (dReal)-1.0);
    for (int i = 0; i < 4; i++) {   // exactly parallel to the axis: the cap branch, from above, below, inside and beside
      dGeomID r = scene_add_geom(sw, dCreateRay(rs, (dReal)(i == 2 ? 0.3 : 4)));
      dGeomRaySet(r, (dReal)(3.5 + 0.1 * i), (dReal)(i == 3 ? 4.3 : 3.4), (dReal)(i == 1 ? -3 : (i == 2 ? -1.1 : 2)), 0, 0, (dReal)(i == 1 ? 1 : -1));
    }
  }
  // free-standing rays, mode = i % 6
  for (int i = 0; i < 42; i++) {
    dGeomID r = scene_add_geom(sw, dCreateRay(rs, rng.uni(2, 9)));
    const dReal px = rng.uni(-4, 4), py = rng.uni(-4, 4), pz = rng.uni(-0.5, 5);
    if (i < 18) dGeomRaySet(r, px, py, pz, rng.uni(-0.6, 0.6), rng.uni(-0.6, 0.6), (dReal)((i / 6) & 1 ? 1 : -1));
    else if (i < 30) dGeomRaySet(r, px, py, rng.uni(0.1, 1.0), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-0.15, 0.15));   // skim the settled bodies
    else dGeomRaySet(r, px, py, pz, rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1));
    const int mode = i % 6;
    dGeomRaySetParams(r, mode == 1 || mode == 4, mode >= 3);
    dGeomRaySetClosestHit(r, mode == 2 || mode == 5);
  }
  // sensor rays on bodies: straight down the body's -z through an offset rotation, and one without offset
  for (size_t i = 0; i < carriers.size(); i++) {
    dGeomID r = scene_add_geom(sw, dCreateRay(rs, (dReal)2.5));
    dGeomSetBody(r, carriers[i]);
    if (i != 1) {
      dMatrix3 Ro;
      dRFromAxisAndAngle(Ro, 1, 0, 0, (dReal)3.14159265358979);
      dGeomSetOffsetRotation(r, Ro);
      dGeomSetOffsetPosition(r, rng.uni(-0.1, 0.1), rng.uni(-0.1, 0.1), 0);
    }
    dGeomRaySetClosestHit(r, i & 1);
  }
}

// ode/demo/demo_crash.cpp as shipped (CARS + WALL + CANNON, :47-78, :165-231, :252-290): dSweepAndPruneSpace XYZ, gravity -1.5,
// a brick wall of unit boxes, a car (box chassis, four sphere wheels on hinge2 joints with the demo's suspension and motor
// parameters, a counterweight body on a fixed joint below the chassis) driving into it and a cannon ball on its way to the wall.
// The demo's callback gives pairs with a sphere mu = 20 and all others mu = 0.5 (:128-131): ScenePolicy::sphere_mu.  A smaller
// wall than the demo's 12 x 10 keeps the trace short; the body count (wall + 6 car bodies + ball) stays on the CTA-per-world path.
static inline void scene_crashwall(SceneWorld &sw, int w, int wallw, int wallh) {
  sw.world = dWorldCreate();
  sw.space2 = 0;
  sw.space = dSweepAndPruneSpaceCreate(0, dSAP_AXES_XYZ);
  sw.cgroup = dJointGroupCreate(0);
  sw.seed = scene_world_seed(w);
  dWorldSetGravity(sw.world, 0, 0, (dReal)-1.5);
  dWorldSetCFM(sw.world, (dReal)1e-5);
  dWorldSetERP(sw.world, (dReal)0.8);
  dWorldSetQuickStepNumIterations(sw.world, 20);
  xs32 rng(sw.seed ^ 0x00C8A54u);
  dGeomSetCategoryBits(scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0)), 1);   // every geom that is not a sphere: category bit 1 only
  const dReal LENGTH = (dReal)3.5, WIDTH = (dReal)2.5, HEIGHT = 1, RADIUS = (dReal)0.5, STARTZ = 1;
  const dReal x = -12, y = rng.uni(-0.2, 0.2);
  dMass m;
  dBodyID chassis = dBodyCreate(sw.world);
  dBodySetPosition(chassis, x, y, STARTZ);
  dMassSetBox(&m, 1, LENGTH, WIDTH, HEIGHT);
  dMassAdjust(&m, (dReal)0.5);
  dBodySetMass(chassis, &m);
  { dGeomID g = scene_add_geom(sw, dCreateBox(sw.space, LENGTH, WIDTH, HEIGHT)); dGeomSetCategoryBits(g, 1); dGeomSetBody(g, chassis); }
  sw.bodies.push_back(chassis);
  const dReal wx[4] = {(dReal)(x + 0.4 * LENGTH - 0.5 * RADIUS), (dReal)(x + 0.4 * LENGTH - 0.5 * RADIUS), (dReal)(x - 0.4 * LENGTH + 0.5 * RADIUS), (dReal)(x - 0.4 * LENGTH + 0.5 * RADIUS)};
  const dReal wy[4] = {(dReal)(y + WIDTH * 0.5), (dReal)(y - WIDTH * 0.5), (dReal)(y + WIDTH * 0.5), (dReal)(y - WIDTH * 0.5)};
  for (int i = 0; i < 4; i++) {
    dBodyID wheel = dBodyCreate(sw.world);
    dQuaternion q;
    dQFromAxisAndAngle(q, 1, 0, 0, (dReal)(3.14159265358979 * 0.5));
    dBodySetQuaternion(wheel, q);
    dMassSetSphere(&m, 1, RADIUS);
    dMassAdjust(&m, 1);
    dBodySetMass(wheel, &m);
    dGeomID g = scene_add_geom(sw, dCreateSphere(sw.space, RADIUS));
    dGeomSetCategoryBits(g, SCENE_CAT_SPHERE);
    dGeomSetBody(g, wheel);
    dBodySetPosition(wheel, wx[i], wy[i], STARTZ - HEIGHT * (dReal)0.5);
    sw.bodies.push_back(wheel);
    dJointID j = dJointCreateHinge2(sw.world, 0);
    dJointAttach(j, chassis, wheel);
    const dReal *a = dBodyGetPosition(wheel);
    dJointSetHinge2Anchor(j, a[0], a[1], a[2]);
    dJointSetHinge2Axis1(j, 0, 0, (dReal)(i < 2 ? 1 : -1));
    dJointSetHinge2Axis2(j, 0, 1, 0);
    dJointSetHinge2Param(j, dParamSuspensionERP, (dReal)0.8);
    dJointSetHinge2Param(j, dParamSuspensionCFM, (dReal)1e-5);
    dJointSetHinge2Param(j, dParamVel2, (dReal)-6);       // the demo's 'a' key a few times: drive towards the wall (-x)
    dJointSetHinge2Param(j, dParamFMax2, 25);
    dJointSetHinge2Param(j, dParamLoStop, 0); dJointSetHinge2Param(j, dParamHiStop, 0);   // steering straight (simLoop's lock of the rear wheels, applied to all four)
    dJointSetHinge2Param(j, dParamFudgeFactor, (dReal)0.1);
    sw.joints.push_back(j);
  }
  {   // centre-of-mass offset body on a fixed joint (:214-222)
    dBodyID b = dBodyCreate(sw.world);
    dBodySetPosition(b, x, y, STARTZ - 5);
    dMassSetBox(&m, 1, LENGTH, WIDTH, HEIGHT);
    dMassAdjust(&m, (dReal)0.5);
    dBodySetMass(b, &m);
    sw.bodies.push_back(b);
    dJointID j = dJointCreateFixed(sw.world, 0);
    dJointAttach(j, chassis, b);
    dJointSetFixed(j);
    sw.joints.push_back(j);
  }
  bool offset = false;   // the wall (:274-291)
  for (dReal z = (dReal)0.5; z <= (dReal)wallh; z += 1) {
    offset = !offset;
    for (dReal yy = (-(dReal)wallw + z) / 2; yy <= ((dReal)wallw - z) / 2; yy += 1) {
      dBodyID b = dBodyCreate(sw.world);
      dBodySetPosition(b, -20, yy, z);
      dMassSetBox(&m, 1, 1, 1, 1);
      dMassAdjust(&m, 1);
      dBodySetMass(b, &m);
      { dGeomID g = scene_add_geom(sw, dCreateBox(sw.space, 1, 1, 1)); dGeomSetCategoryBits(g, 1); dGeomSetBody(g, b); }
      sw.bodies.push_back(b);
    }
  }
  {   // the cannon ball in flight (:432-452: mass 10, radius 0.5, fired at the wall)
    dBodyID b = dBodyCreate(sw.world);
    dMassSetSphereTotal(&m, 10, (dReal)0.5);
    dBodySetMass(b, &m);
    dBodySetPosition(b, -14, rng.uni(-1.0, 1.0), (dReal)2.5);
    dBodySetLinearVel(b, -10, 0, (dReal)0.5);
    dGeomID g = scene_add_geom(sw, dCreateSphere(sw.space, (dReal)0.5));
    dGeomSetCategoryBits(g, SCENE_CAT_SPHERE);
    dGeomSetBody(g, b);
    sw.bodies.push_back(b);
  }
}

static inline ScenePolicy policy_buggy() {
  // ode/demo/demo_buggy.cpp:96-103
  ScenePolicy p;
  memset(&p, 0, sizeof(p));
  p.max_contacts = 8;
  p.skip_if_connected = 0;
  p.surface.mode = dContactSlip1 | dContactSlip2 | dContactSoftERP | dContactSoftCFM | dContactApprox1;
  p.surface.mu = dInfinity;
  p.surface.slip1 = (dReal)0.1;
  p.surface.slip2 = (dReal)0.1;
  p.surface.soft_erp = (dReal)0.5;
  p.surface.soft_cfm = (dReal)0.3;
  return p;
}


// Body-level switches of dxStepBody / the quickstep preamble that no other scene turns on: finite rotation (both modes,
// util.cpp:288-330), linear / angular damping with thresholds (:337-359), max angular speed (:258-268), gravity mode off,
// gyroscopic term off (quickstep.cpp:633-652), and the world's contact correction limits (contact.cpp:130-141:
// dWorldSetContactMaxCorrectingVel / dWorldSetContactSurfaceLayer).  Tumbling boxes and spheres dropped on a plane.
static inline void scene_bodyflags(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0BADF00Du);
  dWorldSetContactMaxCorrectingVel(sw.world, (dReal)1.5);
  dWorldSetContactSurfaceLayer(sw.world, (dReal)0.002);
  dWorldSetLinearDamping(sw.world, (dReal)0.002);           // defaults copied into every body created afterwards
  dWorldSetAngularDampingThreshold(sw.world, (dReal)0.05);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int i = 0; i < 14; i++) {
    dBodyID b = (i % 3 == 2) ? scene_add_sphere(sw, 2, rng.uni(0.2, 0.45), rng.uni(-1.2, 1.2), rng.uni(-1.2, 1.2), rng.uni(0.6, 4))
                             : scene_add_box(sw, 2, rng.uni(0.25, 0.8), rng.uni(0.25, 0.8), rng.uni(0.25, 0.8), rng.uni(-1.2, 1.2), rng.uni(-1.2, 1.2), rng.uni(0.6, 4));
    dQuaternion q = {rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1)};
    dBodySetQuaternion(b, q);
    dBodySetAngularVel(b, rng.uni(-6, 6), rng.uni(-6, 6), rng.uni(-6, 6));
    dBodySetLinearVel(b, rng.uni(-1, 1), rng.uni(-1, 1), rng.uni(-1, 1));
    switch (i % 7) {
      case 0: dBodySetFiniteRotationMode(b, 1); break;                                              // finite rotation about the whole angular velocity
      case 1: dBodySetFiniteRotationMode(b, 1); dBodySetFiniteRotationAxis(b, 0, 0, 1); break;     // finite about z, infinitesimal for the rest
      case 2: dBodySetLinearDamping(b, (dReal)0.02); dBodySetAngularDamping(b, (dReal)0.05); dBodySetLinearDampingThreshold(b, (dReal)0.3); break;
      case 3: dBodySetMaxAngularSpeed(b, (dReal)2.5); break;
      case 4: dBodySetGravityMode(b, 0); dBodySetLinearVel(b, 0, 0, (dReal)-0.8); break;
      case 5: dBodySetGyroscopicMode(b, 0); break;
      default: dBodySetFiniteRotationMode(b, 1); dBodySetFiniteRotationAxis(b, (dReal)0.6, 0, (dReal)0.8); dBodySetAngularDamping(b, (dReal)0.01); dBodySetMaxAngularSpeed(b, 4); break;
    }
  }
}

// Auto-disable (util.cpp:99-233, instantaneous-sample mode) and re-enabling through the island walk (util.cpp:447-451):
// a small pile settles and falls asleep body by body; late spheres land on it and wake it; one body has auto-disable off,
// one was put to sleep by hand, one has its own thresholds.
static inline void scene_autodisable(SceneWorld &sw, int w, int avg = 0) {
  scene_world_base(sw, w);
  if (avg) dWorldSetAutoDisableAverageSamplesCount(sw.world, 5);   // averaged mode: bodies created below inherit 5 samples
  xs32 rng(sw.seed ^ 0x51EE9u);
  dWorldSetAutoDisableFlag(sw.world, 1);
  dWorldSetAutoDisableLinearThreshold(sw.world, (dReal)0.05);
  dWorldSetAutoDisableAngularThreshold(sw.world, (dReal)0.05);
  dWorldSetAutoDisableSteps(sw.world, 8);
  dWorldSetAutoDisableTime(sw.world, (dReal)0.05);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  for (int i = 0; i < 9; i++) {
    dBodyID b = scene_add_box(sw, 3, rng.uni(0.4, 0.7), rng.uni(0.4, 0.7), rng.uni(0.3, 0.5),
                              (dReal)((i % 3 - 1) * 0.75) + rng.uni(-0.03, 0.03), (dReal)((i / 3 - 1) * 0.75) + rng.uni(-0.03, 0.03), (dReal)0.3);
    if (i == 4) dBodySetAutoDisableFlag(b, 0);
    if (avg && i == 2) dBodySetAutoDisableAverageSamplesCount(b, 3);
    if (avg && i == 5) dBodySetAutoDisableAverageSamplesCount(b, 1);
    if (i == 7) { dBodySetAutoDisableLinearThreshold(b, (dReal)0.5); dBodySetAutoDisableSteps(b, 2); dBodySetAutoDisableTime(b, 0); }
    if (i == 8) dBodyDisable(b);
  }
  for (int i = 0; i < 4; i++) scene_add_box(sw, 3, (dReal)0.5, (dReal)0.5, (dReal)0.4, (dReal)((i % 2) * 0.8 - 0.4), (dReal)((i / 2) * 0.8 - 0.4), (dReal)0.85);
  // late arrivals: they fall for 0.7 .. 1.6 s and hit the sleeping pile
  for (int i = 0; i < 4; i++) scene_add_sphere(sw, 4, (dReal)0.22, rng.uni(-0.7, 0.7), rng.uni(-0.7, 0.7), (dReal)(3.5 + 3.0 * i));
}

// Contact surface modes no other policy uses (contact.cpp:74-256): Mu2, Motion1 / Motion2 / MotionN (conveyor-belt
// terms in the right-hand side), Slip1 / Slip2 with non-zero slip, Bounce, SoftERP + SoftCFM, pyramid approximation on
// direction 2 only.  The callback variant adds FDir1 (ScenePolicy::fdir1).
static inline ScenePolicy policy_contactmodes(int fdir1) {
  ScenePolicy p;
  memset(&p, 0, sizeof(p));
  p.max_contacts = 6;
  p.skip_if_connected = 1;
  p.surface.mode = dContactMu2 | dContactMotion1 | dContactMotion2 | dContactMotionN | dContactSlip1 | dContactSlip2 | dContactBounce |
                   dContactSoftERP | dContactSoftCFM | dContactApprox1_2 | (fdir1 ? dContactFDir1 : 0);
  p.surface.mu = (dReal)0.9;
  p.surface.mu2 = (dReal)0.3;
  p.surface.bounce = (dReal)0.3;
  p.surface.bounce_vel = (dReal)0.2;
  p.surface.soft_erp = (dReal)0.6;
  p.surface.soft_cfm = (dReal)0.003;
  p.surface.motion1 = (dReal)0.4;
  p.surface.motion2 = (dReal)-0.25;
  p.surface.motionN = (dReal)0.05;
  p.surface.slip1 = (dReal)0.02;
  p.surface.slip2 = (dReal)0.05;
  p.fdir1 = fdir1;
  return p;
}


// Nested spaces (collision_space.cpp:772-833, collision_kernel.cpp:104-131), after demo_buggy.cpp:226-310: two cars, each in its own
// simple space (sublevel 1) inside the main space, one of them carrying a further sub-space (sublevel 2) with two "antenna" spheres;
// a ground plane and a ramp box in the main space; a few loose boxes.  The cars start close enough to run into each other:
// space x space pairs, space x geom pairs and the sublevel rule all occur.
static inline void scene_nested(SceneWorld &sw, int w) {
  scene_world_base(sw, w);
  xs32 rng(sw.seed ^ 0x0E57EDu);
  scene_add_geom(sw, dCreatePlane(sw.space, 0, 0, 1, 0));
  {
    dGeomID ramp = scene_add_geom(sw, dCreateBox(sw.space, 2, (dReal)1.5, 1));
    dMatrix3 R;
    dRFromAxisAndAngle(R, 0, 1, 0, (dReal)-0.15);
    dGeomSetPosition(ramp, 2, 0, (dReal)-0.34);
    dGeomSetRotation(ramp, R);
  }
  const dReal L = (dReal)0.7, Wd = (dReal)0.5, H = (dReal)0.2, Rw = (dReal)0.18, Z = (dReal)0.5;
  for (int car = 0; car < 2; car++) {
    dSpaceID main_space = sw.space;
    dSpaceID car_space = dSimpleSpaceCreate(main_space);
    dSpaceSetCleanup(car_space, 0);
    dSpaceSetSublevel(car_space, 1);
    scene_add_geom(sw, (dGeomID)car_space);
    const dReal x0 = (dReal)(car == 0 ? 0.0 : 1.35), y0 = rng.uni(-0.05, 0.05);
    sw.space = car_space;   // scene_add_box / scene_add_sphere create their geoms in sw.space
    dBodyID chassis = scene_add_box(sw, 1 / (L * Wd * H), L, Wd, H, x0, y0, Z);
    for (int i = 0; i < 3; i++) {
      const dReal wx = (dReal)(i == 0 ? 0.5 * L : -0.5 * L), wy = (dReal)(i == 0 ? 0 : (i == 1 ? 0.5 * Wd : -0.5 * Wd));
      dBodyID wheel = scene_add_sphere(sw, (dReal)(0.2 / (4.0 / 3.0 * 3.14159265358979 * Rw * Rw * Rw)), Rw, x0 + wx, y0 + wy, Z - H * (dReal)0.5);
      dQuaternion q;
      dQFromAxisAndAngle(q, 1, 0, 0, (dReal)(3.14159265358979 * 0.5));
      dBodySetQuaternion(wheel, q);
      dJointID j = dJointCreateHinge2(sw.world, 0);
      dJointAttach(j, chassis, wheel);
      const dReal *a = dBodyGetPosition(wheel);
      dJointSetHinge2Anchor(j, a[0], a[1], a[2]);
      dJointSetHinge2Axis1(j, 0, 0, 1);
      dJointSetHinge2Axis2(j, 0, 1, 0);
      dJointSetHinge2Param(j, dParamSuspensionERP, (dReal)0.4);
      dJointSetHinge2Param(j, dParamSuspensionCFM, (dReal)0.8);
      if (i > 0) { dJointSetHinge2Param(j, dParamLoStop, 0); dJointSetHinge2Param(j, dParamHiStop, 0); }
      dJointSetHinge2Param(j, dParamVel2, (dReal)(car == 0 ? -3.0 : 1.0)); dJointSetHinge2Param(j, dParamFMax2, (dReal)0.1);
      sw.joints.push_back(j);
    }
    if (car == 0) {
      dSpaceID ant = dHashSpaceCreate(car_space);
      dSpaceSetCleanup(ant, 0);
      dSpaceSetSublevel(ant, 2);
      scene_add_geom(sw, (dGeomID)ant);
      sw.space = ant;
      for (int k = 0; k < 2; k++) {
        dBodyID b = scene_add_sphere(sw, 1, (dReal)0.06, x0 + (dReal)(0.2 * k - 0.1), y0, Z + (dReal)0.4);
        dJointID j = dJointCreateFixed(sw.world, 0);
        dJointAttach(j, chassis, b);
        dJointSetFixed(j);
        sw.joints.push_back(j);
      }
    }
    sw.space = main_space;
  }
  for (int i = 0; i < 4; i++) scene_add_box(sw, 1, (dReal)0.3, (dReal)0.3, (dReal)0.3, rng.uni(0.3, 1.2), rng.uni(-0.6, 0.6), (dReal)(1.2 + 0.5 * i));
}

static inline int scene_build(const char *name_in, SceneWorld &sw, int w, ScenePolicy &pol) {
  // suffixes select the space class: NAME@sap, NAME@sapz (axis order ZXY), NAME@simple
  char name[64];
  strncpy(name, name_in, sizeof name - 1); name[sizeof name - 1] = 0;
  g_scene_space_kind = 0;
  if (char *at = strchr(name, '@')) {
    if (!strcmp(at, "@sap")) g_scene_space_kind = 1;
    else if (!strcmp(at, "@sapz")) g_scene_space_kind = 3;
    else if (!strcmp(at, "@simple")) g_scene_space_kind = 2;
    else return -1;
    *at = 0;
  }
  pol = policy_boxstack();
  if (!strcmp(name, "block64")) { scene_block64(sw, w); return 0; }
  if (!strcmp(name, "tower64")) { scene_tower64(sw, w); return 0; }
  if (!strcmp(name, "stack32")) { scene_stack32(sw, w); return 0; }
  if (!strcmp(name, "mixed")) { scene_mixed(sw, w, 12, 6); return 0; }
  if (!strcmp(name, "mixed_maxc4")) { scene_mixed(sw, w, 12, 6); pol = policy_crash(); return 0; }
  if (!strcmp(name, "mixed_varmaxc")) { scene_mixed(sw, w, 12, 6); pol.varmaxc = 1; return 0; }   // callback loop only
  if (!strcmp(name, "nested")) { scene_nested(sw, w); pol = policy_buggy(); pol.nested = 1; return 0; }          // callback loop only
  if (!strcmp(name, "nested_dcollide")) { scene_nested(sw, w); pol = policy_buggy(); pol.nested = 2; return 0; } // callback loop only
  if (!strcmp(name, "crashwall")) { scene_crashwall(sw, w, 8, 6); pol = policy_crash(); pol.sphere_mu = 20; return 0; }
  if (!strcmp(name, "bodyflags")) { scene_bodyflags(sw, w); return 0; }
  if (!strcmp(name, "autodisable")) { scene_autodisable(sw, w); return 0; }
  if (!strcmp(name, "autodisable_avg")) { scene_autodisable(sw, w, 1); return 0; }   // averaged samples (util.cpp:139-205)
  if (!strcmp(name, "contactmodes")) { scene_mixed(sw, w, 10, 5); pol = policy_contactmodes(0); return 0; }
  if (!strcmp(name, "contactmodes_fdir1")) { scene_mixed(sw, w, 10, 5); pol = policy_contactmodes(1); return 0; }   // callback loop only
  if (!strcmp(name, "chain")) { scene_chain(sw, w, 8); return 0; }
  if (!strcmp(name, "hinges")) { scene_hinges(sw, w); return 0; }
  if (!strcmp(name, "sliders")) { scene_sliders(sw, w); return 0; }
  if (!strcmp(name, "pistons")) { scene_pistons(sw, w); return 0; }
  if (!strcmp(name, "pus")) { scene_pus(sw, w); return 0; }
  if (!strcmp(name, "kinematic")) { scene_kinematic(sw, w); return 0; }
  if (!strcmp(name, "nulljoint")) { scene_nulljoint(sw, w); return 0; }
  if (!strcmp(name, "transforms_rays")) { scene_transforms_rays(sw, w); return 0; }
  if (!strcmp(name, "transforms")) { scene_transforms(sw, w); return 0; }
  if (!strcmp(name, "cylspheres")) { scene_cylspheres(sw, w); return 0; }
  if (!strcmp(name, "cylmix")) { scene_cylmix(sw, w); return 0; }
  if (!strcmp(name, "universals")) { scene_universals(sw, w); return 0; }
  if (!strcmp(name, "motors")) { scene_motors(sw, w); return 0; }
  if (!strcmp(name, "buggy_terrain")) { scene_buggy_terrain(sw, w, 48); pol = policy_buggy(); pol.max_contacts = 10; return 0; }
  if (!strcmp(name, "buggy_terrain256")) { scene_buggy_terrain(sw, w, 256); pol = policy_buggy(); pol.max_contacts = 10; return 0; }
  if (!strcmp(name, "terrain_boxes")) { scene_terrain_boxes(sw, w); return 0; }
  if (!strcmp(name, "terrain_plane")) { scene_terrain_plane(sw, w); return 0; }
  if (!strcmp(name, "terrain_capsules")) { scene_terrain_capsules(sw, w, 0); return 0; }
  if (!strcmp(name, "terrain_capsules_pre")) { scene_terrain_capsules(sw, w, 1); return 0; }
  if (!strcmp(name, "terrain_spheres")) { scene_terrain_spheres(sw, w); return 0; }
  if (!strcmp(name, "capsmix")) { scene_capsmix(sw, w); return 0; }
  if (!strcmp(name, "raycast")) { scene_raycast(sw, w, 0); return 0; }
  if (!strcmp(name, "raycast2")) { scene_raycast(sw, w, 1); return 0; }
  if (!strcmp(name, "raycast2h")) { scene_raycast(sw, w, 2); return 0; }
  if (!strcmp(name, "raycyl")) { scene_raycast(sw, w, 0, true); return 0; }
  if (!strcmp(name, "ragdoll")) { scene_ragdoll(sw, w); pol = policy_crash(); return 0; }
  if (!strcmp(name, "buggy")) { scene_buggy(sw, w); pol = policy_buggy(); return 0; }
  {
    int nx, ny, nz;
    if (sscanf(name, "pile_%dx%dx%d", &nx, &ny, &nz) == 3 && nx > 0 && ny > 0 && nz > 0) { scene_pile(sw, w, nx, ny, nz); pol = policy_crash(); return 0; }
  }
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

// ob_api_extra.cpp — the remaining small entry points of the reference's public C API for the object families this
// library serves (include/ode/objects.h, collision.h, rotation.h, misc.h, mass.h): accessors, force helpers, joint
// conveniences, rotation utilities.  Host bookkeeping only; each function follows the reference's arithmetic so that
// what it stores or returns is bit-identical (tests/test_api_probe.py runs the same probe program against both).
#include <stdio.h>
#include <string.h>
#include "ob_host.h"
#include "ob_rows.h"
#include "ob_collide.h"
#include "ob_trimesh_host.h"

void ob_marshal_joint(const dxJoint *j, ObJoint &d);

extern "C" {
// ---- bodies (ode/src/ode.cpp:640-1110) ------------------------------------------------------------------
void dBodyAddForceAtRelPos(dBodyID b, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz) {
  dReal prel[4] = {px, py, pz, 0}, f[4] = {fx, fy, fz, 0}, p[4], c[4];
  ob_mul0_331(p, b->R, prel);
  b->facc[0] += f[0]; b->facc[1] += f[1]; b->facc[2] += f[2];
  ob_cross(c, p, f);
  b->tacc[0] += c[0]; b->tacc[1] += c[1]; b->tacc[2] += c[2];
}
void dBodyAddRelForceAtPos(dBodyID b, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz) {
  dReal frel[4] = {fx, fy, fz, 0}, f[4], q[4], c[4];
  ob_mul0_331(f, b->R, frel);
  b->facc[0] += f[0]; b->facc[1] += f[1]; b->facc[2] += f[2];
  q[0] = px - b->pos[0]; q[1] = py - b->pos[1]; q[2] = pz - b->pos[2];
  ob_cross(c, q, f);
  b->tacc[0] += c[0]; b->tacc[1] += c[1]; b->tacc[2] += c[2];
}
void dBodyAddRelForceAtRelPos(dBodyID b, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz) {
  dReal frel[4] = {fx, fy, fz, 0}, prel[4] = {px, py, pz, 0}, f[4], p[4], c[4];
  ob_mul0_331(f, b->R, frel);
  ob_mul0_331(p, b->R, prel);
  b->facc[0] += f[0]; b->facc[1] += f[1]; b->facc[2] += f[2];
  ob_cross(c, p, f);
  b->tacc[0] += c[0]; b->tacc[1] += c[1]; b->tacc[2] += c[2];
}
void dBodyCopyPosition(dBodyID b, dVector3 pos) { pos[0] = b->pos[0]; pos[1] = b->pos[1]; pos[2] = b->pos[2]; }
void dBodyCopyQuaternion(dBodyID b, dQuaternion q) { q[0] = b->q[0]; q[1] = b->q[1]; q[2] = b->q[2]; q[3] = b->q[3]; }
void dBodyCopyRotation(dBodyID b, dMatrix3 R) { for (int i = 0; i < 12; i++) R[i] = b->R[i]; }
dReal dBodyGetLinearDamping(dBodyID b) { return b->dampingp.linear_scale; }
dReal dBodyGetAngularDamping(dBodyID b) { return b->dampingp.angular_scale; }
dReal dBodyGetLinearDampingThreshold(dBodyID b) { return ob_sqrt(b->dampingp.linear_threshold); }
dReal dBodyGetAngularDampingThreshold(dBodyID b) { return ob_sqrt(b->dampingp.angular_threshold); }
void dBodySetLinearDampingThreshold(dBodyID b, dReal t) { b->dampingp.linear_threshold = t * t; }
void dBodySetAngularDampingThreshold(dBodyID b, dReal t) { b->dampingp.angular_threshold = t * t; }
void dBodySetDamping(dBodyID b, dReal linear_scale, dReal angular_scale) { dBodySetLinearDamping(b, linear_scale); dBodySetAngularDamping(b, angular_scale); }
dReal dBodyGetAutoDisableLinearThreshold(dBodyID b) { return ob_sqrt(b->adis.linear_average_threshold); }
dReal dBodyGetAutoDisableAngularThreshold(dBodyID b) { return ob_sqrt(b->adis.angular_average_threshold); }
void dBodySetAutoDisableLinearThreshold(dBodyID b, dReal t) { b->adis.linear_average_threshold = t * t; }
void dBodySetAutoDisableAngularThreshold(dBodyID b, dReal t) { b->adis.angular_average_threshold = t * t; }
int dBodyGetAutoDisableAverageSamplesCount(dBodyID b) { return (int)b->adis.average_samples; }
int dBodyGetAutoDisableSteps(dBodyID b) { return b->adis.idle_steps; }
void dBodySetAutoDisableSteps(dBodyID b, int steps) { b->adis.idle_steps = steps; }
dReal dBodyGetAutoDisableTime(dBodyID b) { return b->adis.idle_time; }
void dBodySetAutoDisableTime(dBodyID b, dReal time) { b->adis.idle_time = time; }
int dBodyGetFiniteRotationMode(dBodyID b) { return (b->flags & OB_BODY_FINITE_ROT) != 0; }
void dBodyGetFiniteRotationAxis(dBodyID b, dVector3 result) { result[0] = b->finite_rot_axis[0]; result[1] = b->finite_rot_axis[1]; result[2] = b->finite_rot_axis[2]; }
dReal dBodyGetMaxAngularSpeed(dBodyID b) { return b->max_angular_speed; }
// kinematic bodies (ode.cpp:834-852): infinite mass -- the inverse mass and inertia are zero, nothing else is special;
// the row assembly, the solver and the integrator read them like any other body's
void dBodySetKinematic(dBodyID b) { for (int i = 0; i < 12; i++) b->invI[i] = 0; b->invMass = 0; }
void dBodySetDynamic(dBodyID b) { dBodySetMass(b, &b->mass); }
int dBodyIsKinematic(dBodyID b) { return b->invMass == 0; }
// the reference calls it from dxStepBody (util.cpp:338-340) for every body it stepped; here the drop-in dWorldQuickStep
// calls it for the same bodies in the same order AFTER the whole step came back from the GPU (the callback sees the
// final, damped state of all bodies).  The batched path never calls into user code.
void dBodySetMovedCallback(dBodyID b, void (*callback)(dBodyID)) { b->moved_callback = callback; }
dJointID dBodyGetJoint(dBodyID b, int index) {
  int i = 0;
  for (dxJointNode *n = b->firstjoint; n; n = n->next, i++) if (i == index) return n->joint;
  return 0;
}

// ---- worlds (ode/src/ode.cpp:1840-2010) -----------------------------------------------------------------
dReal dWorldGetLinearDamping(dWorldID w) { return w->dampingp.linear_scale; }
dReal dWorldGetAngularDamping(dWorldID w) { return w->dampingp.angular_scale; }
dReal dWorldGetLinearDampingThreshold(dWorldID w) { return ob_sqrt(w->dampingp.linear_threshold); }
dReal dWorldGetAngularDampingThreshold(dWorldID w) { return ob_sqrt(w->dampingp.angular_threshold); }
dReal dWorldGetMaxAngularSpeed(dWorldID w) { return w->max_angular_speed; }
dReal dWorldGetAutoDisableLinearThreshold(dWorldID w) { return ob_sqrt(w->adis.linear_average_threshold); }
dReal dWorldGetAutoDisableAngularThreshold(dWorldID w) { return ob_sqrt(w->adis.angular_average_threshold); }
int dWorldGetAutoDisableAverageSamplesCount(dWorldID w) { return (int)w->adis.average_samples; }
int dWorldGetAutoDisableSteps(dWorldID w) { return w->adis.idle_steps; }
dReal dWorldGetAutoDisableTime(dWorldID w) { return w->adis.idle_time; }
// step working memory is device memory owned by the batch: the reference's arena controls are accepted and ignored
int dWorldUseSharedWorkingMemory(dWorldID, dWorldID) { return 1; }
void dWorldCleanupWorkingMemory(dWorldID) {}
int dWorldSetStepMemoryReservationPolicy(dWorldID, const void *) { return 1; }
int dWorldSetStepMemoryManager(dWorldID, const void *) { return 1; }

// ---- joints (ode/src/ode.cpp:1380-1560, joints/*.cpp) -----------------------------------------------------
void dJointSetData(dJointID j, void *data) { j->userdata = data; }
void *dJointGetData(dJointID j) { return j->userdata; }
int dJointGetNumBodies(dJointID j) { return !j->node[0].body ? 0 : (!j->node[1].body ? 1 : 2); }
dJointID dConnectingJoint(dBodyID in_b1, dBodyID in_b2) {
  dBodyID b1 = in_b1 ? in_b1 : in_b2, b2 = in_b1 ? in_b2 : in_b1;
  for (dxJointNode *n = b1->firstjoint; n; n = n->next) if (n->body == b2) return n->joint;
  return 0;
}
int dConnectingJointList(dBodyID in_b1, dBodyID in_b2, dJointID *out_list) {
  dBodyID b1 = in_b1 ? in_b1 : in_b2, b2 = in_b1 ? in_b2 : in_b1;
  int n_out = 0;
  for (dxJointNode *n = b1->firstjoint; n; n = n->next) if (n->body == b2) out_list[n_out++] = n->joint;
  return n_out;
}
dReal dJointGetBallParam(dJointID j, int parameter) {   // ball.cpp:132-150
  if (parameter == dParamCFM) return j->cfm;
  if (parameter == dParamERP) return j->erp;
  return 0;
}
void dJointAddHingeTorque(dJointID j, dReal torque) {   // hinge.cpp:326-345
  dReal axis[4] = {0, 0, 0, 0};
  if (j->flags & dJOINT_REVERSE) torque = -torque;
  if (j->node[0].body) ob_mul0_331(axis, j->node[0].body->R, j->axis1);
  axis[0] *= torque; axis[1] *= torque; axis[2] *= torque;
  if (j->node[0].body) dBodyAddTorque(j->node[0].body, axis[0], axis[1], axis[2]);
  if (j->node[1].body) dBodyAddTorque(j->node[1].body, -axis[0], -axis[1], -axis[2]);
}
static void universal_axis(dxJoint *j, int second, dReal *axis) {   // getAxis / getAxis2 (joint.cpp)
  if (!second) { if (j->node[0].body) ob_mul0_331(axis, j->node[0].body->R, j->axis1); }
  else if (j->node[1].body) ob_mul0_331(axis, j->node[1].body->R, j->axis2);
  else { axis[0] = j->axis2[0]; axis[1] = j->axis2[1]; axis[2] = j->axis2[2]; }
}
static dReal universal_rate(dxJoint *j, int second) {   // universal.cpp:682-727
  if (!j->node[0].body) return 0;
  dReal axis[4] = {0, 0, 0, 0};
  universal_axis(j, (j->flags & dJOINT_REVERSE) ? !second : second, axis);
  dReal rate = ob_dot(axis, j->node[0].body->avel);
  if (j->node[1].body) rate -= ob_dot(axis, j->node[1].body->avel);
  return rate;
}
dReal dJointGetUniversalAngle1Rate(dJointID j) { return universal_rate(j, 0); }
dReal dJointGetUniversalAngle2Rate(dJointID j) { return universal_rate(j, 1); }
void dJointAddUniversalTorques(dJointID j, dReal torque1, dReal torque2) {   // universal.cpp:729-753
  dReal a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  if (j->flags & dJOINT_REVERSE) { const dReal t = torque1; torque1 = -torque2; torque2 = -t; }
  universal_axis(j, 0, a1);
  universal_axis(j, 1, a2);
  a1[0] = a1[0] * torque1 + a2[0] * torque2;
  a1[1] = a1[1] * torque1 + a2[1] * torque2;
  a1[2] = a1[2] * torque1 + a2[2] * torque2;
  if (j->node[0].body) dBodyAddTorque(j->node[0].body, a1[0], a1[1], a1[2]);
  if (j->node[1].body) dBodyAddTorque(j->node[1].body, -a1[0], -a1[1], -a1[2]);
}
void dJointAddAMotorTorques(dJointID j, dReal torque1, dReal torque2, dReal torque3) {   // amotor.cpp:476-511
  if (j->num == 0 || !j->node[0].body) return;
  dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
  ObJoint o;
  ob_marshal_joint(j, o);
  ObBodyView B1 = {b0->pos, b0->R, b0->q, b0->lvel, b0->avel}, B2 = B1;
  if (b1) { B2.pos = b1->pos; B2.R = b1->R; B2.q = b1->q; B2.lvel = b1->lvel; B2.avel = b1->avel; }
  real ax[3][3];
  ob_amotor_axes(o, B1, b1 ? &B2 : (const ObBodyView *)0, ax);
  ax[0][0] *= torque1; ax[0][1] *= torque1; ax[0][2] *= torque1;
  if (j->num >= 2) {
    ax[0][0] += ax[1][0] * torque2; ax[0][1] += ax[1][1] * torque2; ax[0][2] += ax[1][2] * torque2;
    if (j->num >= 3) { ax[0][0] += ax[2][0] * torque3; ax[0][1] += ax[2][1] * torque3; ax[0][2] += ax[2][2] * torque3; }
  }
  dBodyAddTorque(b0, ax[0][0], ax[0][1], ax[0][2]);
  if (b1) dBodyAddTorque(b1, -ax[0][0], -ax[0][1], -ax[0][2]);
}

// ---- geoms (ode/src/collision_kernel.cpp:640-700, 1130-1200) -----------------------------------------------
void dGeomCopyPosition(dGeomID g, dVector3 pos) { const dReal *p = dGeomGetPosition(g); pos[0] = p[0]; pos[1] = p[1]; pos[2] = p[2]; }
void dGeomCopyRotation(dGeomID g, dMatrix3 R) { const dReal *r = dGeomGetRotation(g); for (int i = 0; i < 12; i++) R[i] = r[i]; }
static const dReal g_zero3[4] = {0, 0, 0, 0};
static const dReal g_ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
const dReal *dGeomGetOffsetPosition(dGeomID g) { return g->offset_posr ? g->offset_posr->pos : g_zero3; }
const dReal *dGeomGetOffsetRotation(dGeomID g) { return g->offset_posr ? g->offset_posr->R : g_ident; }
void dGeomCopyOffsetPosition(dGeomID g, dVector3 pos) { const dReal *p = dGeomGetOffsetPosition(g); pos[0] = p[0]; pos[1] = p[1]; pos[2] = p[2]; }
void dGeomCopyOffsetRotation(dGeomID g, dMatrix3 R) { const dReal *r = dGeomGetOffsetRotation(g); for (int i = 0; i < 12; i++) R[i] = r[i]; }
void dGeomGetOffsetQuaternion(dGeomID g, dQuaternion result) {
  if (g->offset_posr) ob_QfromR(result, g->offset_posr->R);
  else { result[0] = 1; result[1] = 0; result[2] = 0; result[3] = 0; }
}
// trimesh accessors (collision_trimesh_opcode.cpp:820-882, collision_trimesh_internal.h:394-411, :585-592)
dTriMeshDataID dGeomTriMeshGetTriMeshDataID(dGeomID g) { return g->tmdata; }
static void fetch_triangle(dxGeom *g, int index, dReal out[3][4]) {
  const dReal *pos = dGeomGetPosition(g), *R = dGeomGetRotation(g);
  const dxTriMeshData *d = g->tmdata;
  for (int i = 0; i < 3; i++) {
    const float *vf = &d->verts[(size_t)3 * d->tris[(size_t)3 * index + i]];
    const dReal v[4] = {vf[0], vf[1], vf[2], 0};
    ob_mul0_331(out[i], R, v);
    out[i][0] += pos[0]; out[i][1] += pos[1]; out[i][2] += pos[2]; out[i][3] = 0;
  }
}
void dGeomTriMeshGetTriangle(dGeomID g, int index, dVector3 *v0, dVector3 *v1, dVector3 *v2) {
  dReal v[3][4];
  fetch_triangle(g, index, v);
  if (v0) for (int k = 0; k < 4; k++) (*v0)[k] = v[0][k];
  if (v1) for (int k = 0; k < 4; k++) (*v1)[k] = v[1][k];
  if (v2) for (int k = 0; k < 4; k++) (*v2)[k] = v[2][k];
}
void dGeomTriMeshGetPoint(dGeomID g, int index, dReal u, dReal v, dVector3 out) {
  dReal dv[3][4];
  fetch_triangle(g, index, dv);
  const dReal w = OB_REAL(1.0) - u - v;
  for (int k = 0; k < 4; k++) out[k] = (dv[0][k] * w) + (dv[1][k] * u) + (dv[2][k] * v);
}
int dSpaceGetClass(dSpaceID space) { return space->type; }                       // collision_space.cpp:740
void dSpaceSetManualCleanup(dSpaceID space, int mode) { space->manual_cleanup = mode != 0; }   // :679; spaces here are flat: the flag is only stored
int dSpaceGetManualCleanup(dSpaceID space) { return space->manual_cleanup; }
void dInfiniteAABB(dGeomID, dReal aabb[6]) { aabb[0] = -dInfinity; aabb[1] = dInfinity; aabb[2] = -dInfinity; aabb[3] = dInfinity; aabb[4] = -dInfinity; aabb[5] = dInfinity; }

// ---- rotation / random / mass utilities (rotation.cpp, misc.cpp, mass.cpp) ------------------------------------
void dQMultiply1(dQuaternion qa, const dQuaternion qb, const dQuaternion qc) { ob_qmul1(qa, qb, qc); }
void dQMultiply2(dQuaternion qa, const dQuaternion qb, const dQuaternion qc) { ob_qmul2(qa, qb, qc); }
void dQMultiply3(dQuaternion qa, const dQuaternion qb, const dQuaternion qc) { ob_qmul3(qa, qb, qc); }
void dRFrom2Axes(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal bx, dReal by, dReal bz) { ob_Rfrom2axes(R, ax, ay, az, bx, by, bz); }
void dRFromZAxis(dMatrix3 R, dReal ax, dReal ay, dReal az) {   // rotation.cpp:136-156
  dReal n[4] = {ax, ay, az, 0}, p[4], q[4];
  ob_safe_normalize3(n);
  ob_plane_space(n, p, q);
  R[0] = p[0]; R[4] = p[1]; R[8] = p[2];
  R[1] = q[0]; R[5] = q[1]; R[9] = q[2];
  R[2] = n[0]; R[6] = n[1]; R[10] = n[2];
  R[3] = 0; R[7] = 0; R[11] = 0;
}
dReal dRandReal(void) { return ((dReal)dRand()) / ((dReal)0xffffffff); }
#define OB_PAD(a) (((a) > 1) ? ((((a) - 1) | 3) + 1) : (a))   // dPAD, common.h
void dMakeRandomVector(dReal *A, int n, dReal range) { for (int i = 0; i < n; i++) A[i] = (dRandReal() * OB_REAL(2.0) - OB_REAL(1.0)) * range; }
void dMakeRandomMatrix(dReal *A, int n, int m, dReal range) {
  const int skip = OB_PAD(m);
  dReal *row = A;
  for (int i = 0; i < n; row += skip, ++i) for (int j = 0; j < m; ++j) row[j] = (dRandReal() * OB_REAL(2.0) - OB_REAL(1.0)) * range;
}
void dClearUpperTriangle(dReal *A, int n) {
  const int skip = OB_PAD(n);
  dReal *row = A;
  for (int i = 0; i < n; row += skip, ++i) for (int j = i + 1; j < n; ++j) row[j] = 0;
}
dReal dMaxDifference(const dReal *A, const dReal *B, int n, int m) {
  const int skip = OB_PAD(m);
  dReal mx = 0;
  for (int i = 0; i < n; A += skip, B += skip, ++i) for (int j = 0; j < m; ++j) { const dReal d = ob_fabs(A[j] - B[j]); if (d > mx) mx = d; }
  return mx;
}
dReal dMaxDifferenceLowerTriangle(const dReal *A, const dReal *B, int n) {
  const int skip = OB_PAD(n);
  dReal mx = 0;
  for (int i = 0; i < n; A += skip, B += skip, ++i) for (int j = 0; j <= i; ++j) { const dReal d = ob_fabs(A[j] - B[j]); if (d > mx) mx = d; }
  return mx;
}
int dAllocateODEDataForThread(unsigned int) { return 1; }   // odeinit.h: no per-thread data in this library
void dCleanupODEAllDataForThread(void) {}
// point depth queries (sphere.cpp:95, plane.cpp:141, capsule.cpp:104, box.cpp:109): positive inside
dReal dGeomSpherePointDepth(dGeomID g, dReal x, dReal y, dReal z) {
  const dReal *pos = dGeomGetPosition(g);
  return g->p[0] - ob_sqrt((x - pos[0]) * (x - pos[0]) + (y - pos[1]) * (y - pos[1]) + (z - pos[2]) * (z - pos[2]));
}
dReal dGeomPlanePointDepth(dGeomID g, dReal x, dReal y, dReal z) { return g->p[3] - g->p[0] * x - g->p[1] * y - g->p[2] * z; }
dReal dGeomCapsulePointDepth(dGeomID g, dReal x, dReal y, dReal z) {
  const dReal *pos = dGeomGetPosition(g), *R = dGeomGetRotation(g);
  dReal a[3] = {x - pos[0], y - pos[1], z - pos[2]};
  dReal beta = ob_dot14(a, R + 2);
  const dReal lz2 = g->p[1] * OB_REAL(0.5);
  if (beta < -lz2) beta = -lz2; else if (beta > lz2) beta = lz2;
  a[0] = pos[0] + beta * R[0 * 4 + 2]; a[1] = pos[1] + beta * R[1 * 4 + 2]; a[2] = pos[2] + beta * R[2 * 4 + 2];
  return g->p[0] - ob_sqrt((x - a[0]) * (x - a[0]) + (y - a[1]) * (y - a[1]) + (z - a[2]) * (z - a[2]));
}
dReal dGeomBoxPointDepth(dGeomID g, dReal x, dReal y, dReal z) {
  const dReal *pos = dGeomGetPosition(g), *R = dGeomGetRotation(g);
  dReal p[4] = {x - pos[0], y - pos[1], z - pos[2], 0}, q[4], dist[6];
  ob_mul1_331(q, R, p);
  bool inside = true;
  for (int i = 0; i < 3; i++) {
    const dReal side = g->p[i] * OB_REAL(0.5);
    dist[i] = side - q[i]; dist[i + 3] = side + q[i];
    if (dist[i] < 0 || dist[i + 3] < 0) inside = false;
  }
  if (inside) {
    dReal smallest = (dReal)(unsigned)-1;
    for (int i = 0; i < 6; i++) if (dist[i] < smallest) smallest = dist[i];
    return smallest;
  }
  dReal largest = 0;
  for (int i = 0; i < 6; i++) if (dist[i] > largest) largest = dist[i];
  return -largest;
}
void dJointAddHinge2Torques(dJointID j, dReal torque1, dReal torque2) {   // hinge2.cpp:393-410
  if (j->node[0].body && j->node[1].body) {
    dReal a1[4], a2[4];
    ob_mul0_331(a1, j->node[0].body->R, j->axis1);
    ob_mul0_331(a2, j->node[1].body->R, j->axis2);
    a1[0] = a1[0] * torque1 + a2[0] * torque2;
    a1[1] = a1[1] * torque1 + a2[1] * torque2;
    a1[2] = a1[2] * torque1 + a2[2] * torque2;
    dBodyAddTorque(j->node[0].body, a1[0], a1[1], a1[2]);
    dBodyAddTorque(j->node[1].body, -a1[0], -a1[1], -a1[2]);
  }
}
// body / geom frame conversions (ode.cpp:714-765, collision_kernel.cpp:784-866)
void dBodyGetPointVel(dBodyID b, dReal px, dReal py, dReal pz, dVector3 result) {
  OB_AASSERT(b);
  dReal p[4] = {px - b->pos[0], py - b->pos[1], pz - b->pos[2], 0}, c[4];
  result[0] = b->lvel[0]; result[1] = b->lvel[1]; result[2] = b->lvel[2];
  ob_cross(c, b->avel, p);
  result[0] += c[0]; result[1] += c[1]; result[2] += c[2];
}
void dBodyGetPosRelPoint(dBodyID b, dReal px, dReal py, dReal pz, dVector3 result) {
  OB_AASSERT(b);
  dReal prel[4] = {px - b->pos[0], py - b->pos[1], pz - b->pos[2], 0};
  ob_mul1_331(result, b->R, prel);
}
void dBodyVectorFromWorld(dBodyID b, dReal px, dReal py, dReal pz, dVector3 result) {
  OB_AASSERT(b);
  dReal p[4] = {px, py, pz, 0};
  ob_mul1_331(result, b->R, p);
}
void dWorldImpulseToForce(dWorldID w, dReal stepsize, dReal ix, dReal iy, dReal iz, dVector3 force) {
  OB_AASSERT(w);
  stepsize = OB_REAL(1.0) / stepsize;
  force[0] = stepsize * ix; force[1] = stepsize * iy; force[2] = stepsize * iz;
}
static bool geom_frame(dGeomID g, dReal px, dReal py, dReal pz, dVector3 result) {   // non-placeable: identity
  OB_AASSERT(g);
  if (!(g->gflags & GEOM_PLACEABLE)) { result[0] = px; result[1] = py; result[2] = pz; return false; }
  ob_geom_recompute_posr(g);
  return true;
}
void dGeomGetRelPointPos(dGeomID g, dReal px, dReal py, dReal pz, dVector3 result) {
  if (!geom_frame(g, px, py, pz, result)) return;
  dReal prel[4] = {px, py, pz, 0}, p[4];
  ob_mul0_331(p, g->final_posr->R, prel);
  result[0] = p[0] + g->final_posr->pos[0]; result[1] = p[1] + g->final_posr->pos[1]; result[2] = p[2] + g->final_posr->pos[2];
}
void dGeomGetPosRelPoint(dGeomID g, dReal px, dReal py, dReal pz, dVector3 result) {
  if (!geom_frame(g, px, py, pz, result)) return;
  dReal prel[4] = {px - g->final_posr->pos[0], py - g->final_posr->pos[1], pz - g->final_posr->pos[2], 0};
  ob_mul1_331(result, g->final_posr->R, prel);
}
void dGeomVectorToWorld(dGeomID g, dReal px, dReal py, dReal pz, dVector3 result) {
  if (!geom_frame(g, px, py, pz, result)) return;
  dReal p[4] = {px, py, pz, 0};
  ob_mul0_331(result, g->final_posr->R, p);
}
void dGeomVectorFromWorld(dGeomID g, dReal px, dReal py, dReal pz, dVector3 result) {
  if (!geom_frame(g, px, py, pz, result)) return;
  dReal p[4] = {px, py, pz, 0};
  ob_mul1_331(result, g->final_posr->R, p);
}
void dClosestLineSegmentPoints(const dVector3 a1, const dVector3 a2, const dVector3 b1, const dVector3 b2, dVector3 cp1, dVector3 cp2) {
  ob_closest_segment_points(a1, a2, b1, b2, cp1, cp2);   // the same function the capsule colliders run on the device
}
void dPrintMatrix(const dReal *A, int n, int m, char *fmt, FILE *f) {   // misc.cpp:128-136, rows padded to a multiple of 4
  const int skip = m > 1 ? ((m - 1) | 3) + 1 : m;
  for (int i = 0; i < n; i++, A += skip) {
    for (int j = 0; j < m; j++) fprintf(f, fmt, A[j]);
    fprintf(f, "\n");
  }
}
// dMassSetTrimesh (mass.cpp:234-427): Mirtich's polyhedral mass properties ("Fast and Accurate Computation of Polyhedral
// Mass Properties", jgt 1(2) 1996) over the triangles in world coordinates.  Per face: projection integrals over the
// plane of the two minor axes (A, B) of the normal, face integrals from them, volume integrals accumulated; every
// expression keeps the reference's operand order so that the result is the same bit for bit.
namespace {
struct ObProj { dReal P1, Pa, Pb, Paa, Pab, Pbb, Paaa, Paab, Pabb, Pbbb; };
static void mirtich_projection(const dVector3 *v, int A, int B, ObProj &P) {
  memset(&P, 0, sizeof P);
  for (int j = 0; j < 3; j++) {   // edges v[j] -> v[j+1]
    const int jn = j == 2 ? 0 : j + 1;
    const dReal a0 = v[j][A], b0 = v[j][B], a1 = v[jn][A], b1 = v[jn][B];
    const dReal da = a1 - a0, db = b1 - b0;
    const dReal a0_2 = a0 * a0, a0_3 = a0_2 * a0, a0_4 = a0_3 * a0;
    const dReal b0_2 = b0 * b0, b0_3 = b0_2 * b0, b0_4 = b0_3 * b0;
    const dReal a1_2 = a1 * a1, a1_3 = a1_2 * a1, b1_2 = b1 * b1, b1_3 = b1_2 * b1;
    const dReal C1 = a1 + a0;
    const dReal Ca = a1 * C1 + a0_2, Caa = a1 * Ca + a0_3, Caaa = a1 * Caa + a0_4;
    const dReal Cb = b1 * (b1 + b0) + b0_2, Cbb = b1 * Cb + b0_3, Cbbb = b1 * Cbb + b0_4;
    const dReal Cab = 3 * a1_2 + 2 * a1 * a0 + a0_2, Kab = a1_2 + 2 * a1 * a0 + 3 * a0_2;
    const dReal Caab = a0 * Cab + 4 * a1_3, Kaab = a1 * Kab + 4 * a0_3;
    const dReal Cabb = 4 * b1_3 + 3 * b1_2 * b0 + 2 * b1 * b0_2 + b0_3;
    const dReal Kabb = b1_3 + 2 * b1_2 * b0 + 3 * b1 * b0_2 + 4 * b0_3;
    P.P1 += db * C1; P.Pa += db * Ca; P.Paa += db * Caa; P.Paaa += db * Caaa;
    P.Pb += da * Cb; P.Pbb += da * Cbb; P.Pbbb += da * Cbbb;
    P.Pab += db * (b1 * Cab + b0 * Kab);
    P.Paab += db * (b1 * Caab + b0 * Kaab);
    P.Pabb += da * (a1 * Cabb + a0 * Kabb);
  }
  P.P1 /= 2.0; P.Pa /= 6.0; P.Paa /= 12.0; P.Paaa /= 20.0;
  P.Pb /= -6.0; P.Pbb /= -12.0; P.Pbbb /= -20.0;
  P.Pab /= 24.0; P.Paab /= 60.0; P.Pabb /= -60.0;
}
}  // namespace
void dMassSetTrimesh(dMass *m, dReal density, dGeomID g) {
  OB_AASSERT(m);
  OB_UASSERT(g && g->type == dTriMeshClass, "argument not a trimesh");
  dMassSetZero(m);
  const int ntri = dGeomTriMeshGetTriangleCount(g);
  dReal T0 = 0, T1[3] = {0, 0, 0}, T2[3] = {0, 0, 0}, TP[3] = {0, 0, 0};
  for (int i = 0; i < ntri; i++) {
    dVector3 v[3], n, ea, eb;
    dGeomTriMeshGetTriangle(g, i, &v[0], &v[1], &v[2]);
    for (int k = 0; k < 3; k++) { ea[k] = v[1][k] - v[0][k]; eb[k] = v[2][k] - v[0][k]; }
    ob_cross(n, eb, ea);
    const dReal nx = ob_fabs(n[0]), ny = ob_fabs(n[1]), nz = ob_fabs(n[2]);
    const int C = (nx > ny && nx > nz) ? 0 : ((ny > nz) ? 1 : 2);
    if (n[C] == 0) continue;   // a triangle the pose transform degenerated
    const int A = (C + 1) % 3, B = (A + 1) % 3;
    ObProj P;
    mirtich_projection(v, A, B, P);
    const dReal nA = n[A], nB = n[B];
    const dReal w = -ob_dot(n, v[0]);
    const dReal k1 = 1 / n[C], k2 = k1 * k1, k3 = k2 * k1, k4 = k3 * k1;
    const dReal Fa = k1 * P.Pa, Fb = k1 * P.Pb;
    const dReal Fc = -k2 * (nA * P.Pa + nB * P.Pb + w * P.P1);
    const dReal Faa = k1 * P.Paa, Fbb = k1 * P.Pbb;
    const dReal Fcc = k3 * ((nA * nA) * P.Paa + 2 * nA * nB * P.Pab + (nB * nB) * P.Pbb + w * (2 * (nA * P.Pa + nB * P.Pb) + w * P.P1));
    const dReal Faaa = k1 * P.Paaa, Fbbb = k1 * P.Pbbb;
    const dReal Fccc = -k4 * ((nA * nA * nA) * P.Paaa + 3 * (nA * nA) * nB * P.Paab + 3 * nA * (nB * nB) * P.Pabb + (nB * nB * nB) * P.Pbbb +
                              3 * w * ((nA * nA) * P.Paa + 2 * nA * nB * P.Pab + (nB * nB) * P.Pbb) + w * w * (3 * (nA * P.Pa + nB * P.Pb) + w * P.P1));
    const dReal Faab = k1 * P.Paab;
    const dReal Fbbc = -k2 * (nA * P.Pabb + nB * P.Pbbb + w * P.Pbb);
    const dReal Fcca = k3 * ((nA * nA) * P.Paaa + 2 * nA * nB * P.Paab + (nB * nB) * P.Pabb + w * (2 * (nA * P.Paa + nB * P.Pab) + w * P.Pa));
    T0 += n[0] * ((A == 0) ? Fa : ((B == 0) ? Fb : Fc));
    T1[A] += n[A] * Faa; T1[B] += n[B] * Fbb; T1[C] += n[C] * Fcc;
    T2[A] += n[A] * Faaa; T2[B] += n[B] * Fbbb; T2[C] += n[C] * Fccc;
    TP[A] += n[A] * Faab; TP[B] += n[B] * Fbbc; TP[C] += n[C] * Fcca;
  }
  for (int k = 0; k < 3; k++) { T1[k] /= 2; T2[k] /= 3; TP[k] /= 2; }
  m->mass = density * T0;
  m->I[0] = density * (T2[1] + T2[2]);
  m->I[5] = density * (T2[2] + T2[0]);
  m->I[10] = density * (T2[0] + T2[1]);
  m->I[1] = m->I[4] = -density * TP[0];
  m->I[9] = m->I[6] = -density * TP[1];
  m->I[8] = m->I[2] = -density * TP[2];
  dMassTranslate(m, T1[0] / T0, T1[1] / T0, T1[2] / T0);   // as the reference does (its SF bug 1729095 fix)
}
void dMassSetTrimeshTotal(dMass *m, dReal total_mass, dGeomID g) {
  OB_AASSERT(m);
  OB_UASSERT(g && g->type == dTriMeshClass, "argument not a trimesh");
  dMassSetTrimesh(m, 1.0, g);
  dMassAdjust(m, total_mass);
}
void dMassSetCappedCylinder(dMass *m, dReal density, int direction, dReal radius, dReal length) { dMassSetCapsule(m, density, direction, radius, length); }
void dMassSetCappedCylinderTotal(dMass *m, dReal total_mass, int direction, dReal radius, dReal length) { dMassSetCapsuleTotal(m, total_mass, direction, radius, length); }
}  // extern "C"

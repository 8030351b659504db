// api_probe — calls the small accessor / convenience entry points of the ODE C API on a fixed little scene and prints
// every result as raw bits.  Built three times (against the unmodified reference, the host mirror and the CUDA
// library); tests/test_api_probe.py requires identical output.  Public C API only, no stepping: runs without a GPU.
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#ifdef PROBE_REFERENCE_HEADERS
#include <ode/ode.h>
#else
#include <ode_b200/ode.h>
#endif

static void pr(const char *name, const dReal *v, int n) {
  printf("%s", name);
  for (int i = 0; i < n; i++) {
    if (sizeof(dReal) == 4) { uint32_t u; memcpy(&u, &v[i], 4); printf(" %08x", u); }
    else { uint64_t u; memcpy(&u, &v[i], 8); printf(" %016llx", (unsigned long long)u); }
  }
  printf("\n");
}
static void pr1(const char *name, dReal v) { pr(name, &v, 1); }
static void body_acc(const char *name, dBodyID b) {
  const dReal *f = dBodyGetForce(b), *t = dBodyGetTorque(b);
  dReal v[6] = {f[0], f[1], f[2], t[0], t[1], t[2]};
  pr(name, v, 6);
}

int main() {
  dInitODE2(0);
  dWorldID w = dWorldCreate();
  dSpaceID s = dHashSpaceCreate(0);
  dWorldSetLinearDamping(w, (dReal)0.01); dWorldSetAngularDamping(w, (dReal)0.02);
  dWorldSetLinearDampingThreshold(w, (dReal)0.3); dWorldSetAngularDampingThreshold(w, (dReal)0.4);
  dWorldSetMaxAngularSpeed(w, (dReal)50);
  dWorldSetAutoDisableLinearThreshold(w, (dReal)0.05); dWorldSetAutoDisableAngularThreshold(w, (dReal)0.06);
  dWorldSetAutoDisableSteps(w, 7); dWorldSetAutoDisableTime(w, (dReal)0.25); dWorldSetAutoDisableAverageSamplesCount(w, 3);
  pr1("wlin", dWorldGetLinearDamping(w)); pr1("wang", dWorldGetAngularDamping(w));
  pr1("wlint", dWorldGetLinearDampingThreshold(w)); pr1("wangt", dWorldGetAngularDampingThreshold(w));
  pr1("wmax", dWorldGetMaxAngularSpeed(w));
  pr1("wadl", dWorldGetAutoDisableLinearThreshold(w)); pr1("wada", dWorldGetAutoDisableAngularThreshold(w));
  printf("wad %d %d\n", dWorldGetAutoDisableSteps(w), dWorldGetAutoDisableAverageSamplesCount(w));
  pr1("wadt", dWorldGetAutoDisableTime(w));

  dBodyID b[3];
  for (int i = 0; i < 3; i++) {
    b[i] = dBodyCreate(w);
    dBodySetPosition(b[i], (dReal)(0.7 * i), (dReal)(0.1 * i), (dReal)(1 + 0.3 * i));
    dQuaternion q = {(dReal)0.9, (dReal)(0.1 + 0.1 * i), (dReal)-0.2, (dReal)(0.3 - 0.1 * i)};
    dBodySetQuaternion(b[i], q);
    dBodySetLinearVel(b[i], (dReal)0.1, (dReal)(0.2 * i), (dReal)-0.3);
    dBodySetAngularVel(b[i], (dReal)(0.5 - i), (dReal)0.7, (dReal)(0.2 * i));
  }
  dReal v3[4], m12[12], q4[4];
  dBodyCopyPosition(b[1], v3); pr("bpos", v3, 3);
  dBodyCopyQuaternion(b[1], q4); pr("bq", q4, 4);
  dBodyCopyRotation(b[1], m12); pr("bR", m12, 12);
  pr1("blin", dBodyGetLinearDamping(b[0])); pr1("bang", dBodyGetAngularDamping(b[0]));
  pr1("blint", dBodyGetLinearDampingThreshold(b[0])); pr1("bangt", dBodyGetAngularDampingThreshold(b[0]));
  dBodySetLinearDampingThreshold(b[0], (dReal)0.11); dBodySetAngularDampingThreshold(b[0], (dReal)0.13);
  dBodySetDamping(b[0], (dReal)0.03, (dReal)0.04);
  pr1("blin2", dBodyGetLinearDamping(b[0])); pr1("bang2", dBodyGetAngularDamping(b[0]));
  pr1("blint2", dBodyGetLinearDampingThreshold(b[0])); pr1("bangt2", dBodyGetAngularDampingThreshold(b[0]));
  pr1("bmax", dBodyGetMaxAngularSpeed(b[0]));
  pr1("badl", dBodyGetAutoDisableLinearThreshold(b[0])); pr1("bada", dBodyGetAutoDisableAngularThreshold(b[0]));
  dBodySetAutoDisableLinearThreshold(b[0], (dReal)0.07); dBodySetAutoDisableAngularThreshold(b[0], (dReal)0.09);
  dBodySetAutoDisableSteps(b[0], 11); dBodySetAutoDisableTime(b[0], (dReal)0.5);
  pr1("badl2", dBodyGetAutoDisableLinearThreshold(b[0])); pr1("bada2", dBodyGetAutoDisableAngularThreshold(b[0]));
  printf("bad %d %d\n", dBodyGetAutoDisableSteps(b[0]), dBodyGetAutoDisableAverageSamplesCount(b[0]));
  pr1("badt", dBodyGetAutoDisableTime(b[0]));
  dBodySetFiniteRotationMode(b[2], 1); dBodySetFiniteRotationAxis(b[2], (dReal)0.1, (dReal)0.2, (dReal)0.9);
  dBodyGetFiniteRotationAxis(b[2], v3); pr("bfra", v3, 3);
  printf("bfrm %d %d\n", dBodyGetFiniteRotationMode(b[2]), dBodyGetFiniteRotationMode(b[0]));
  dBodyAddForceAtRelPos(b[0], (dReal)0.3, (dReal)-0.2, (dReal)0.5, (dReal)0.1, (dReal)0.2, (dReal)-0.1); body_acc("f_atrel", b[0]);
  dBodyAddRelForceAtPos(b[0], (dReal)0.3, (dReal)-0.2, (dReal)0.5, (dReal)0.1, (dReal)0.2, (dReal)1.1); body_acc("frel_at", b[0]);
  dBodyAddRelForceAtRelPos(b[0], (dReal)0.3, (dReal)-0.2, (dReal)0.5, (dReal)0.1, (dReal)0.2, (dReal)-0.1); body_acc("frel_atrel", b[0]);

  // joints
  dJointID jh = dJointCreateHinge(w, 0), ju = dJointCreateUniversal(w, 0), ja = dJointCreateAMotor(w, 0), jb = dJointCreateBall(w, 0);
  dJointAttach(jh, b[0], b[1]); dJointSetHingeAnchor(jh, (dReal)0.3, 0, (dReal)1.1); dJointSetHingeAxis(jh, (dReal)0.1, 1, (dReal)0.2);
  dJointAttach(ju, b[1], b[2]); dJointSetUniversalAnchor(ju, 1, (dReal)0.1, (dReal)1.4); dJointSetUniversalAxis1(ju, 1, 0, (dReal)0.1); dJointSetUniversalAxis2(ju, 0, 1, 0);
  dJointAttach(ja, b[0], b[2]); dJointSetAMotorMode(ja, dAMotorUser); dJointSetAMotorNumAxes(ja, 3);
  dJointSetAMotorAxis(ja, 0, 1, 1, 0, (dReal)0.2); dJointSetAMotorAxis(ja, 1, 2, 0, 1, 0); dJointSetAMotorAxis(ja, 2, 0, (dReal)0.1, 0, 1);
  dJointAttach(jb, 0, b[2]); dJointSetBallAnchor(jb, (dReal)1.4, (dReal)0.2, (dReal)1.9);
  dJointSetBallParam(jb, dParamCFM, (dReal)0.001); dJointSetBallParam(jb, dParamERP, (dReal)0.7);
  pr1("ballcfm", dJointGetBallParam(jb, dParamCFM)); pr1("ballerp", dJointGetBallParam(jb, dParamERP));
  static int tagdata = 5;
  dJointSetData(ju, &tagdata);
  printf("jdata %d nb %d %d %d\n", *(int *)dJointGetData(ju), dJointGetNumBodies(jh), dJointGetNumBodies(jb), dJointGetNumBodies(ja));
  printf("conn %d %d %d\n", dConnectingJoint(b[0], b[1]) == jh, dConnectingJoint(b[1], b[2]) == ju, dConnectingJoint(b[0], b[2]) == ja);
  dJointID list[8];
  printf("connlist %d %d\n", dConnectingJointList(b[0], b[2], list), dConnectingJointList(0, b[2], list));
  printf("bodyjoint %d %d %d\n", dBodyGetJoint(b[2], 0) == jb, dBodyGetJoint(b[2], 1) == ja, dBodyGetJoint(b[2], 5) == 0);
  pr1("urate1", dJointGetUniversalAngle1Rate(ju)); pr1("urate2", dJointGetUniversalAngle2Rate(ju));
  for (int i = 0; i < 3; i++) { dBodySetForce(b[i], 0, 0, 0); dBodySetTorque(b[i], 0, 0, 0); }
  dJointAddHingeTorque(jh, (dReal)0.7); body_acc("hingeT0", b[0]); body_acc("hingeT1", b[1]);
  dJointAddUniversalTorques(ju, (dReal)0.4, (dReal)-0.9); body_acc("univT1", b[1]); body_acc("univT2", b[2]);
  dJointAddAMotorTorques(ja, (dReal)0.2, (dReal)0.3, (dReal)-0.5); body_acc("amotT0", b[0]); body_acc("amotT2", b[2]);

  // geoms
  dGeomID g = dCreateBox(s, (dReal)0.4, (dReal)0.5, (dReal)0.6);
  dGeomSetBody(g, b[1]);
  pr("goff0", dGeomGetOffsetPosition(g), 3); pr("goffR0", dGeomGetOffsetRotation(g), 12);
  dGeomGetOffsetQuaternion(g, q4); pr("goffq0", q4, 4);
  dGeomSetOffsetPosition(g, (dReal)0.1, (dReal)-0.2, (dReal)0.05);
  dMatrix3 Ro; dRFromAxisAndAngle(Ro, (dReal)0.3, 1, (dReal)0.2, (dReal)0.8); dGeomSetOffsetRotation(g, Ro);
  dGeomCopyOffsetPosition(g, v3); pr("goff", v3, 3);
  dGeomCopyOffsetRotation(g, m12); pr("goffR", m12, 12);
  dGeomGetOffsetQuaternion(g, q4); pr("goffq", q4, 4);
  dGeomCopyPosition(g, v3); pr("gpos", v3, 3);
  dGeomCopyRotation(g, m12); pr("gR", m12, 12);
  dReal ab[6]; dInfiniteAABB(g, ab); pr("infaabb", ab, 6);

  // rotation / random / mass helpers
  dQuaternion qa, qb = {(dReal)0.8, (dReal)0.1, (dReal)-0.5, (dReal)0.3}, qc = {(dReal)0.2, (dReal)0.9, (dReal)0.1, (dReal)-0.4};
  dQMultiply1(qa, qb, qc); pr("qm1", qa, 4);
  dQMultiply2(qa, qb, qc); pr("qm2", qa, 4);
  dQMultiply3(qa, qb, qc); pr("qm3", qa, 4);
  dMatrix3 R; memset(R, 0, sizeof R);
  dRFrom2Axes(R, (dReal)0.3, (dReal)0.9, (dReal)-0.2, (dReal)0.5, (dReal)-0.1, (dReal)0.8); pr("r2ax", R, 12);
  dRFromZAxis(R, (dReal)0.3, (dReal)-0.6, (dReal)0.7); pr("rzax", R, 12);
  dRFromZAxis(R, 0, 0, 1); pr("rzax2", R, 12);
  dRandSetSeed(12345);
  pr1("rand1", dRandReal()); pr1("rand2", dRandReal());
  dMass m;
  dMassSetCappedCylinder(&m, (dReal)2.5, 2, (dReal)0.3, (dReal)1.2); pr1("ccmass", m.mass); pr("ccI", m.I, 12);
  dMassSetCappedCylinderTotal(&m, (dReal)4, 3, (dReal)0.2, (dReal)0.9); pr("ccIt", m.I, 12);
  dMassSetCylinder(&m, (dReal)1.5, 1, (dReal)0.25, (dReal)0.8); pr1("cylmass", m.mass); pr("cylI", m.I, 12);
  // second batch: matrix helpers, point depths, hinge2 torques
  dReal A[12], Bm[12];
  dMakeRandomVector(v3, 3, (dReal)2.5); pr("rvec", v3, 3);
  dMakeRandomMatrix(A, 3, 3, (dReal)1.5); dMakeRandomMatrix(Bm, 3, 3, (dReal)1.5);
  pr1("maxdiff", dMaxDifference(A, Bm, 3, 3)); pr1("maxdiffL", dMaxDifferenceLowerTriangle(A, Bm, 3));
  dClearUpperTriangle(A, 3); pr("cleared", A, 12 - 1);
  dGeomID gs = dCreateSphere(s, (dReal)0.4), gc = dCreateCapsule(s, (dReal)0.2, (dReal)0.8), gp = dCreatePlane(s, 0, (dReal)0.6, (dReal)0.8, (dReal)0.3);
  dGeomSetPosition(gs, (dReal)0.2, (dReal)0.1, (dReal)0.5); dGeomSetBody(gc, b[2]);
  pr1("sdepth", dGeomSpherePointDepth(gs, (dReal)0.3, (dReal)0.2, (dReal)0.6)); pr1("pdepth", dGeomPlanePointDepth(gp, 1, 2, (dReal)-0.5));
  pr1("cdepth", dGeomCapsulePointDepth(gc, (dReal)1.5, (dReal)0.3, (dReal)1.7)); pr1("cdepth2", dGeomCapsulePointDepth(gc, 3, 0, 0));
  pr1("bdepth_in", dGeomBoxPointDepth(g, v3[0] * 0 + dGeomGetPosition(g)[0] + (dReal)0.01, dGeomGetPosition(g)[1], dGeomGetPosition(g)[2]));
  pr1("bdepth_out", dGeomBoxPointDepth(g, 2, 2, 2));
  dJointID j2 = dJointCreateHinge2(w, 0);
  dJointAttach(j2, b[0], b[1]); dJointSetHinge2Anchor(j2, (dReal)0.2, (dReal)0.1, 1); dJointSetHinge2Axis1(j2, 0, 0, 1); dJointSetHinge2Axis2(j2, 0, 1, (dReal)0.1);
  for (int i = 0; i < 3; i++) { dBodySetForce(b[i], 0, 0, 0); dBodySetTorque(b[i], 0, 0, 0); }
  dJointAddHinge2Torques(j2, (dReal)0.6, (dReal)-0.3); body_acc("h2T0", b[0]); body_acc("h2T1", b[1]);
  dJointID jh2 = dJointCreateHinge(w, 0), jh3 = dJointCreateHinge(w, 0);
  dJointAttach(jh2, b[1], 0); dJointSetHingeAnchorDelta(jh2, (dReal)0.5, (dReal)0.2, (dReal)1.3, (dReal)0.1, (dReal)-0.05, (dReal)0.2);
  dJointGetHingeAnchor(jh2, v3); pr("hdelta_a1", v3, 3); dJointGetHingeAnchor2(jh2, v3); pr("hdelta_a2", v3, 3);
  dJointAttach(jh3, b[1], b[2]); dJointSetHingeAnchor(jh3, 1, 0, (dReal)1.2); dJointSetHingeAxisOffset(jh3, (dReal)0.2, 1, (dReal)0.1, (dReal)0.35);
  pr1("hoff_angle", dJointGetHingeAngle(jh3)); dJointGetHingeAxis(jh3, v3); pr("hoff_axis", v3, 3);
  dGeomID gw = dCreateBox(s, (dReal)0.2, (dReal)0.3, (dReal)0.4);
  dGeomSetBody(gw, b[2]);
  dGeomSetOffsetWorldPosition(gw, (dReal)1.7, (dReal)0.4, (dReal)1.9); pr("gwo_pos", dGeomGetOffsetPosition(gw), 3); pr("gwo_wpos", dGeomGetPosition(gw), 3);
  dMatrix3 Rw; dRFromAxisAndAngle(Rw, 1, (dReal)0.5, (dReal)-0.3, (dReal)1.1);
  dGeomSetOffsetWorldRotation(gw, Rw); pr("gwo_R", dGeomGetOffsetRotation(gw), 12); pr("gwo_pos2", dGeomGetOffsetPosition(gw), 3); pr("gwo_wR", dGeomGetRotation(gw), 12);
  dQuaternion qw = {(dReal)0.7, (dReal)-0.1, (dReal)0.6, (dReal)0.2}; { dReal l = (dReal)sqrt((double)(qw[0]*qw[0]+qw[1]*qw[1]+qw[2]*qw[2]+qw[3]*qw[3])); for (int k = 0; k < 4; k++) qw[k] /= l; }
  dGeomSetOffsetWorldQuaternion(gw, qw); pr("gwo_R2", dGeomGetOffsetRotation(gw), 12); pr("gwo_pos3", dGeomGetOffsetPosition(gw), 3);
  // trimesh accessors on a small rotated mesh
  static float tverts[5 * 3] = {0, 0, 0, 1, 0, (float)0.2, 0, 1, (float)0.1, 1, 1, (float)-0.3, (float)0.5, (float)0.5, 1};
  static dTriIndex tidx[4 * 3] = {0, 1, 2, 1, 3, 2, 0, 4, 1, 2, 3, 4};
  dTriMeshDataID td = dGeomTriMeshDataCreate();
  dGeomTriMeshDataBuildSingle(td, tverts, 3 * sizeof(float), 5, tidx, 12, 3 * sizeof(dTriIndex));
  dGeomID tm = dCreateTriMesh(s, td, 0, 0, 0);
  dGeomSetPosition(tm, (dReal)0.3, (dReal)-0.2, (dReal)0.7); dGeomSetRotation(tm, Rw);
  printf("tmdata %d\n", dGeomTriMeshGetTriMeshDataID(tm) == td);
  dVector3 t0, t1, t2;
  dGeomTriMeshGetTriangle(tm, 2, &t0, &t1, &t2); pr("tri2_v0", t0, 4); pr("tri2_v1", t1, 4); pr("tri2_v2", t2, 4);
  dGeomTriMeshGetPoint(tm, 3, (dReal)0.25, (dReal)0.6, v3); pr("tri3_pt", v3, 3);
  dJointID ju2 = dJointCreateUniversal(w, 0), ju3 = dJointCreateUniversal(w, 0);
  dJointAttach(ju2, b[0], b[2]); dJointSetUniversalAnchor(ju2, (dReal)0.6, 0, (dReal)1.2);
  dJointSetUniversalAxis2(ju2, 0, 1, (dReal)0.1); dJointSetUniversalAxis1Offset(ju2, 1, (dReal)0.1, 0, (dReal)0.3, (dReal)-0.2);
  dReal ua1, ua2; dJointGetUniversalAngles(ju2, &ua1, &ua2); pr1("uoff1_a1", ua1); pr1("uoff1_a2", ua2);
  dJointAttach(ju3, 0, b[1]); dJointSetUniversalAnchor(ju3, (dReal)0.6, 0, (dReal)1.2);
  dJointSetUniversalAxis1(ju3, 1, 0, (dReal)0.2); dJointSetUniversalAxis2Offset(ju3, 0, 1, 0, (dReal)0.25, (dReal)0.15);
  dJointGetUniversalAngles(ju3, &ua1, &ua2); pr1("uoff2_a1", ua1); pr1("uoff2_a2", ua2);
  // geom transform: accessors, class, AABB of the encapsulated geom at T o local (with and without a body)
  dGeomID xin = dCreateCapsule(0, (dReal)0.15, (dReal)0.7), xt = dCreateGeomTransform(s);
  printf("xf0 %d %d %d %d\n", dGeomGetClass(xt) == dGeomTransformClass, dGeomTransformGetGeom(xt) == 0, dGeomTransformGetCleanup(xt), dGeomTransformGetInfo(xt));
  dReal xab[6]; dGeomGetAABB(xt, xab); pr("xf_aabb_empty", xab, 6);
  dGeomTransformSetGeom(xt, xin); dGeomTransformSetCleanup(xt, 1); dGeomTransformSetInfo(xt, 1);
  printf("xf1 %d %d %d\n", dGeomTransformGetGeom(xt) == xin, dGeomTransformGetCleanup(xt), dGeomTransformGetInfo(xt));
  dGeomSetPosition(xin, (dReal)0.2, (dReal)-0.1, (dReal)0.35); dGeomSetRotation(xin, Rw);
  dGeomSetPosition(xt, (dReal)-1.3, (dReal)0.8, (dReal)2.1); dGeomSetQuaternion(xt, qw);
  dGeomGetAABB(xt, xab); pr("xf_aabb_static", xab, 6);
  dGeomSetBody(xt, b[1]); dGeomGetAABB(xt, xab); pr("xf_aabb_body", xab, 6); pr("xf_pos", dGeomGetPosition(xt), 3);
  dGeomDestroy(xt);
  // frame conversions, impulse to force, closest segment points, dPrintMatrix
  dBodyGetPointVel(b[1], (dReal)0.4, (dReal)-0.7, (dReal)1.9, v3); pr("bpointvel", v3, 3);
  dBodyGetPosRelPoint(b[1], (dReal)0.4, (dReal)-0.7, (dReal)1.9, v3); pr("bposrel", v3, 3);
  dBodyVectorFromWorld(b[1], (dReal)0.3, (dReal)0.9, (dReal)-0.2, v3); pr("bvecfrom", v3, 3);
  dWorldImpulseToForce(w, (dReal)0.013, (dReal)0.3, (dReal)-1.1, (dReal)2.7, v3); pr("imp2f", v3, 3);
  dGeomGetRelPointPos(gw, (dReal)0.1, (dReal)0.2, (dReal)-0.3, v3); pr("grelpt", v3, 3);
  dGeomGetPosRelPoint(gw, (dReal)0.1, (dReal)0.2, (dReal)-0.3, v3); pr("gposrel", v3, 3);
  dGeomVectorToWorld(gw, (dReal)0.1, (dReal)0.2, (dReal)-0.3, v3); pr("gvecto", v3, 3);
  dGeomVectorFromWorld(gw, (dReal)0.1, (dReal)0.2, (dReal)-0.3, v3); pr("gvecfrom", v3, 3);
  { dGeomID pl = dCreatePlane(s, 0, 0, 1, 0); dGeomGetRelPointPos(pl, 1, 2, 3, v3); pr("plrelpt", v3, 3); dGeomVectorFromWorld(pl, 1, 2, 3, v3); pr("plvec", v3, 3); dGeomDestroy(pl); }
  { dVector3 sa1 = {0, 0, 0}, sa2 = {1, (dReal)0.2, 0}, sb1 = {(dReal)0.3, 1, (dReal)0.4}, sb2 = {(dReal)0.6, (dReal)-0.5, (dReal)0.1}, c1, c2;
    dClosestLineSegmentPoints(sa1, sa2, sb1, sb2, c1, c2); pr("clsp1", c1, 3); pr("clsp2", c2, 3);
    dVector3 sb3 = {2, 1, 1}, sb4 = {3, 2, (dReal)1.5};
    dClosestLineSegmentPoints(sa1, sa2, sb3, sb4, c1, c2); pr("clsp3", c1, 3); pr("clsp4", c2, 3); }
  { char fmt[] = "%8.3f "; dPrintMatrix(Rw, 3, 3, fmt, stdout); }
  // dMassSetTrimesh / Total on the rotated open mesh above and on a closed tetrahedron
  { dMass mt; dMassSetTrimesh(&mt, (dReal)2.5, tm); pr1("mtm_mass", mt.mass); pr("mtm_c", mt.c, 3); pr("mtm_I", mt.I, 12);
    static float qv[4 * 3] = {0, 0, 0, (float)1.1, 0, 0, 0, (float)0.9, 0, (float)0.2, (float)0.3, (float)1.3};
    static dTriIndex qi[4 * 3] = {0, 2, 1, 0, 1, 3, 1, 2, 3, 2, 0, 3};
    dTriMeshDataID qd = dGeomTriMeshDataCreate();
    dGeomTriMeshDataBuildSingle(qd, qv, 3 * sizeof(float), 4, qi, 12, 3 * sizeof(dTriIndex));
    dGeomID qm = dCreateTriMesh(0, qd, 0, 0, 0);
    dGeomSetPosition(qm, (dReal)-0.4, (dReal)0.25, (dReal)0.6); dGeomSetQuaternion(qm, qw);
    dMassSetTrimesh(&mt, (dReal)1.7, qm); pr1("mtet_mass", mt.mass); pr("mtet_c", mt.c, 3); pr("mtet_I", mt.I, 12);
    dMassSetTrimeshTotal(&mt, (dReal)3.2, qm); pr1("mtett_mass", mt.mass); pr("mtett_I", mt.I, 12);
    dGeomDestroy(qm); dGeomTriMeshDataDestroy(qd); }
  printf("thr %d\n", dAllocateODEDataForThread(0xffffffffu));
  printf("wsm %d %d %d\n", dWorldUseSharedWorkingMemory(w, 0), dWorldSetStepMemoryReservationPolicy(w, 0), dWorldSetStepMemoryManager(w, 0));
  dWorldCleanupWorkingMemory(w);
  dCloseODE();
  return 0;
}

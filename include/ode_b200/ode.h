/*
 * ode_b200 — C ABI of the B200-native ODE 0.12 hot path.
 *
 * Two groups of entry points:
 *
 *  (1) the drop-in subset of ODE 0.12's public C API (same names, argument
 *      meaning, struct layouts and error behaviour) for the per-step path
 *      dSpaceCollide -> dCollide -> dJointCreateContact -> dWorldQuickStep.
 *      Each declaration cites the reference declaration it replaces
 *      (paths relative to /root/reference/ode-0.12/include/ode/).
 *  (2) the added batched-world entry points (dBatch*), SURVEY.md §8(b):
 *      thousands of independent worlds stepped device-resident, the near
 *      callback replaced by a data table (dBatchContactPolicy).
 *
 * Plain C, no torch types, pointers + sizes only.  The library is built twice,
 * like the reference: -DdSINGLE (libode_b200_single.so) or -DdDOUBLE
 * (libode_b200_double.so); dReal is fixed at compile time (common.h:103-112).
 * There is no CPU fallback: every compute entry point returns an error /
 * calls the error handler when no CUDA device is usable.
 */
#ifndef ODE_B200_ODE_H
#define ODE_B200_ODE_H

#include <stddef.h>
#include <stdio.h>
#include <stdint.h>

#if !defined(dSINGLE) && !defined(dDOUBLE)
#define dSINGLE 1
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar / vector types (common.h:103-144) --------------------------- */
#if defined(dSINGLE)
typedef float dReal;
#else
typedef double dReal;
#endif
typedef dReal dVector3[4];    /* 4-wide padded */
typedef dReal dVector4[4];
typedef dReal dMatrix3[12];   /* 3 rows x 4, row-major, 4th column padding */
typedef dReal dQuaternion[4]; /* (w,x,y,z) */
typedef uint32_t dTriIndex;

#ifndef dInfinity
#if defined(dSINGLE)
#define dInfinity ((float)__builtin_inff())
#else
#define dInfinity (__builtin_inf())
#endif
#endif

/* ---- opaque handles (common.h:222-237) ---------------------------------- */
typedef struct dxWorld *dWorldID;
typedef struct dxSpace *dSpaceID;
typedef struct dxBody *dBodyID;
typedef struct dxGeom *dGeomID;
typedef struct dxJoint *dJointID;
typedef struct dxJointGroup *dJointGroupID;
typedef struct dxTriMeshData *dTriMeshDataID;

/* ---- enums --------------------------------------------------------------- */
/* joint type numbers, common.h:251-267 */
typedef enum {
  dJointTypeNone = 0, dJointTypeBall, dJointTypeHinge, dJointTypeSlider,
  dJointTypeContact, dJointTypeUniversal, dJointTypeHinge2, dJointTypeFixed,
  dJointTypeNull, dJointTypeAMotor, dJointTypeLMotor, dJointTypePlane2D,
  dJointTypePR, dJointTypePU, dJointTypePiston
} dJointType;

/* joint parameter names, common.h:303-355: group g, name k -> 0x100*g + k */
enum {
  dParamLoStop = 0, dParamHiStop, dParamVel, dParamFMax, dParamFudgeFactor,
  dParamBounce, dParamCFM, dParamStopERP, dParamStopCFM,
  dParamSuspensionERP, dParamSuspensionCFM, dParamERP,
  dParamsInGroup,
  dParamGroup = 0x100,
  dParamGroup1 = 0x000, dParamGroup2 = 0x100, dParamGroup3 = 0x200,
  dParamLoStop1 = 0x000, dParamHiStop1, dParamVel1, dParamFMax1, dParamFudgeFactor1,
  dParamBounce1, dParamCFM1, dParamStopERP1, dParamStopCFM1,
  dParamSuspensionERP1, dParamSuspensionCFM1, dParamERP1,
  dParamLoStop2 = 0x100, dParamHiStop2, dParamVel2, dParamFMax2, dParamFudgeFactor2,
  dParamBounce2, dParamCFM2, dParamStopERP2, dParamStopCFM2,
  dParamSuspensionERP2, dParamSuspensionCFM2, dParamERP2,
  dParamLoStop3 = 0x200, dParamHiStop3, dParamVel3, dParamFMax3, dParamFudgeFactor3,
  dParamBounce3, dParamCFM3, dParamStopERP3, dParamStopCFM3,
  dParamSuspensionERP3, dParamSuspensionCFM3, dParamERP3
};

/* geom class numbers, collision.h:879-902 */
enum {
  dSphereClass = 0, dBoxClass, dCapsuleClass, dCylinderClass, dPlaneClass,
  dRayClass, dConvexClass, dGeomTransformClass, dTriMeshClass, dHeightfieldClass,
  dFirstSpaceClass,
  dSimpleSpaceClass = dFirstSpaceClass, dHashSpaceClass, dSweepAndPruneSpaceClass,
  dQuadTreeSpaceClass,
  dLastSpaceClass = dQuadTreeSpaceClass,
  dFirstUserClass, dLastUserClass = dFirstUserClass + 3,
  dGeomNumClasses
};

/* contact surface mode bits, contact.h:33-49 */
enum {
  dContactMu2 = 0x001, dContactFDir1 = 0x002, dContactBounce = 0x004,
  dContactSoftERP = 0x008, dContactSoftCFM = 0x010, dContactMotion1 = 0x020,
  dContactMotion2 = 0x040, dContactMotionN = 0x080, dContactSlip1 = 0x100,
  dContactSlip2 = 0x200,
  dContactApprox0 = 0x0000, dContactApprox1_1 = 0x1000, dContactApprox1_2 = 0x2000,
  dContactApprox1 = 0x3000
};

/* dCollide flags, collision.h:746 */
#define CONTACTS_UNIMPORTANT 0x80000000

/* SAP axis orders, collision_space.h:95-101 */
#define dSAP_AXES_XYZ ((0) | (1 << 2) | (2 << 4))
#define dSAP_AXES_XZY ((0) | (2 << 2) | (1 << 4))
#define dSAP_AXES_YXZ ((1) | (0 << 2) | (2 << 4))
#define dSAP_AXES_YZX ((1) | (2 << 2) | (0 << 4))
#define dSAP_AXES_ZXY ((2) | (0 << 2) | (1 << 4))
#define dSAP_AXES_ZYX ((2) | (1 << 2) | (0 << 4))

/* ---- plain structs shared with callers (layout == reference) ------------- */
/* mass.h:88-92 */
typedef struct dMass {
  dReal mass;
  dVector3 c;
  dMatrix3 I;
} dMass;

/* contact.h:52-65 */
typedef struct dSurfaceParameters {
  int mode;
  dReal mu;
  dReal mu2;
  dReal bounce;
  dReal bounce_vel;
  dReal soft_erp;
  dReal soft_cfm;
  dReal motion1, motion2, motionN;
  dReal slip1, slip2;
} dSurfaceParameters;

/* contact.h:81-87 */
typedef struct dContactGeom {
  dVector3 pos;
  dVector3 normal;
  dReal depth;
  dGeomID g1, g2;
  int side1, side2;
} dContactGeom;

/* contact.h:92-96 */
typedef struct dContact {
  dSurfaceParameters surface;
  dContactGeom geom;
  dVector3 fdir1;
} dContact;

/* common.h:368-373 */
enum { dAMotorUser = 0, dAMotorEuler = 1 };   /* include/ode/common.h:261-264 */

typedef struct dJointFeedback {
  dVector3 f1, t1, f2, t2;
} dJointFeedback;

/* collision_space.h:49 */
typedef void dNearCallback(void *data, dGeomID o1, dGeomID o2);

/* ---- init / misc (odeinit.h:119,236; misc.h:45-60; error.h:52-60) --------- */
int dInitODE2(unsigned int uiInitFlags);
void dInitODE(void);
void dCloseODE(void);
const char *dGetConfiguration(void);
unsigned long dRand(void);
unsigned long dRandGetSeed(void);
void dRandSetSeed(unsigned long s);
int dRandInt(int n);
int dTestRand(void);
typedef void dErrorHandlerFn(int errnum, const char *msg);
/* simplified handler hook: the reference takes (int, const char*, va_list)
 * (error.h:43-50); ours receives the formatted message. */
void dB200SetErrorHandler(dErrorHandlerFn *fn);

/* ---- small math that scene construction needs (odemath.h, rotation.h) ---- */
int dSafeNormalize3(dVector3 a);
int dSafeNormalize4(dVector4 a);
void dNormalize3(dVector3 a);
void dNormalize4(dVector4 a);
void dPlaneSpace(const dVector3 n, dVector3 p, dVector3 q);
int dOrthogonalizeR(dMatrix3 m);
void dRSetIdentity(dMatrix3 R);
void dRFromAxisAndAngle(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal angle);
void dRFromEulerAngles(dMatrix3 R, dReal phi, dReal theta, dReal psi);
void dQSetIdentity(dQuaternion q);
void dQFromAxisAndAngle(dQuaternion q, dReal ax, dReal ay, dReal az, dReal angle);
void dQMultiply0(dQuaternion qa, const dQuaternion qb, const dQuaternion qc);
void dRfromQ(dMatrix3 R, const dQuaternion q);
void dQfromR(dQuaternion q, const dMatrix3 R);
void dDQfromW(dReal dq[4], const dVector3 w, const dQuaternion q);
int dInvertPDMatrix(const dReal *A, dReal *Ainv, int n);

/* ---- further accessors and conveniences of the reference API (ob_api_extra.cpp); the line numbers are those of the
 * reference's include/ode/objects.h unless another header is named ------------------------------------------------ */
void dBodyAddForceAtRelPos(dBodyID, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz);      /* :1040 */
void dBodyAddRelForceAtPos(dBodyID, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz);      /* :1042 */
void dBodyAddRelForceAtRelPos(dBodyID, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz);   /* :1044 */
void dBodyCopyPosition(dBodyID body, dVector3 pos);            /* :940 */
void dBodyCopyRotation(dBodyID, dMatrix3 R);                   /* :957 */
void dBodyCopyQuaternion(dBodyID body, dQuaternion quat);      /* :974 */
dReal dBodyGetLinearDamping(dBodyID b);                        /* :1351 */
dReal dBodyGetAngularDamping(dBodyID b);                       /* :1368 */
dReal dBodyGetLinearDampingThreshold(dBodyID b);
dReal dBodyGetAngularDampingThreshold(dBodyID b);
void dBodySetLinearDampingThreshold(dBodyID b, dReal threshold);
void dBodySetAngularDampingThreshold(dBodyID b, dReal threshold);
void dBodySetDamping(dBodyID b, dReal linear_scale, dReal angular_scale);
dReal dBodyGetAutoDisableLinearThreshold(dBodyID);             /* :753 */
void dBodySetAutoDisableLinearThreshold(dBodyID, dReal linear_average_threshold);
dReal dBodyGetAutoDisableAngularThreshold(dBodyID);
void dBodySetAutoDisableAngularThreshold(dBodyID, dReal angular_average_threshold);
int dBodyGetAutoDisableAverageSamplesCount(dBodyID);
int dBodyGetAutoDisableSteps(dBodyID);
void dBodySetAutoDisableSteps(dBodyID, int steps);
dReal dBodyGetAutoDisableTime(dBodyID);
void dBodySetAutoDisableTime(dBodyID, dReal time);
int dBodyGetFiniteRotationMode(dBodyID);                       /* :1209 */
void dBodyGetFiniteRotationAxis(dBodyID, dVector3 result);
dReal dBodyGetMaxAngularSpeed(dBodyID b);
dJointID dBodyGetJoint(dBodyID, int index);                    /* :1231 */
void dBodySetMovedCallback(dBodyID b, void (*callback)(dBodyID));   /* :1308; drop-in path: called after the step, stepping order */
void dBodySetDynamic(dBodyID);                                 /* :1245 */
void dBodySetKinematic(dBodyID);                               /* :1254 */
int dBodyIsKinematic(dBodyID);                                 /* :1261 */
dReal dWorldGetLinearDamping(dWorldID w);                      /* :663 */
dReal dWorldGetAngularDamping(dWorldID w);
dReal dWorldGetLinearDampingThreshold(dWorldID w);
dReal dWorldGetAngularDampingThreshold(dWorldID w);
dReal dWorldGetMaxAngularSpeed(dWorldID w);
dReal dWorldGetAutoDisableLinearThreshold(dWorldID);           /* :487 */
dReal dWorldGetAutoDisableAngularThreshold(dWorldID);
int dWorldGetAutoDisableAverageSamplesCount(dWorldID);
int dWorldGetAutoDisableSteps(dWorldID);
dReal dWorldGetAutoDisableTime(dWorldID);
/* step working memory lives on the device and belongs to the batch: the arena controls (:157-290) are accepted, return
 * success and change nothing; the info structs are taken as opaque pointers */
int dWorldUseSharedWorkingMemory(dWorldID w, dWorldID from_world);
void dWorldCleanupWorkingMemory(dWorldID w);
int dWorldSetStepMemoryReservationPolicy(dWorldID w, const void *policyinfo);
int dWorldSetStepMemoryManager(dWorldID w, const void *memfuncs);
void dJointSetData(dJointID, void *data);                      /* :1722 */
void *dJointGetData(dJointID);
int dJointGetNumBodies(dJointID);                              /* :1678 */
dJointID dConnectingJoint(dBodyID, dBodyID);                   /* :2940 */
int dConnectingJointList(dBodyID, dBodyID, dJointID *);
dReal dJointGetBallParam(dJointID, int parameter);             /* :2369 */
void dJointAddHingeTorque(dJointID joint, dReal torque);       /* :1857 */
void dJointSetHingeAnchorDelta(dJointID, dReal x, dReal y, dReal z, dReal ax, dReal ay, dReal az);   /* :1812 */
void dJointSetHingeAxisOffset(dJointID j, dReal x, dReal y, dReal z, dReal angle);                   /* :1840 */
dReal dJointGetUniversalAngle1Rate(dJointID);                  /* :2578 */
dReal dJointGetUniversalAngle2Rate(dJointID);
void dJointAddUniversalTorques(dJointID joint, dReal torque1, dReal torque2);               /* :2028 */
void dJointSetUniversalAxis1Offset(dJointID, dReal x, dReal y, dReal z, dReal offset1, dReal offset2);   /* :1948 */
void dJointSetUniversalAxis2Offset(dJointID, dReal x, dReal y, dReal z, dReal offset1, dReal offset2);   /* :1968 */
void dJointAddAMotorTorques(dJointID, dReal torque1, dReal torque2, dReal torque3);         /* :2301 */
void dGeomCopyPosition(dGeomID geom, dVector3 pos);            /* collision.h:194 */
void dGeomCopyRotation(dGeomID geom, dMatrix3 R);
const dReal *dGeomGetOffsetPosition(dGeomID geom);             /* collision.h:680 */
void dGeomCopyOffsetPosition(dGeomID geom, dVector3 pos);
const dReal *dGeomGetOffsetRotation(dGeomID geom);
void dGeomCopyOffsetRotation(dGeomID geom, dMatrix3 R);
void dGeomGetOffsetQuaternion(dGeomID geom, dQuaternion result);
void dGeomSetOffsetWorldPosition(dGeomID geom, dReal x, dReal y, dReal z);    /* collision.h:616 */
void dGeomSetOffsetWorldRotation(dGeomID geom, const dMatrix3 R);             /* collision.h:632 */
void dGeomSetOffsetWorldQuaternion(dGeomID geom, const dQuaternion Q);        /* collision.h:648 */
void dInfiniteAABB(dGeomID geom, dReal aabb[6]);               /* collision.h:1481 */
int dSpaceGetClass(dSpaceID space);                            /* collision_space.h:175 */
void dSpaceSetManualCleanup(dSpaceID space, int mode);         /* collision_space.h:112 */
int dSpaceGetManualCleanup(dSpaceID space);
dTriMeshDataID dGeomTriMeshGetTriMeshDataID(dGeomID g);        /* collision_trimesh.h:180 */
void dGeomTriMeshGetTriangle(dGeomID g, int index, dVector3 *v0, dVector3 *v1, dVector3 *v2);   /* collision_trimesh.h:198, world coordinates */
void dGeomTriMeshGetPoint(dGeomID g, int index, dReal u, dReal v, dVector3 out);                 /* collision_trimesh.h:204 */
void dQMultiply1(dQuaternion qa, const dQuaternion qb, const dQuaternion qc);   /* rotation.h:55-57 */
void dQMultiply2(dQuaternion qa, const dQuaternion qb, const dQuaternion qc);
void dQMultiply3(dQuaternion qa, const dQuaternion qb, const dQuaternion qc);
void dRFrom2Axes(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal bx, dReal by, dReal bz);   /* rotation.h:41 */
void dRFromZAxis(dMatrix3 R, dReal ax, dReal ay, dReal az);                                 /* rotation.h:44 */
dReal dRandReal(void);                                         /* misc.h:54 */
void dMakeRandomVector(dReal *A, int n, dReal range);          /* misc.h:60-75 */
void dMakeRandomMatrix(dReal *A, int n, int m, dReal range);
void dClearUpperTriangle(dReal *A, int n);
dReal dMaxDifference(const dReal *A, const dReal *B, int n, int m);
dReal dMaxDifferenceLowerTriangle(const dReal *A, const dReal *B, int n);
int dAllocateODEDataForThread(unsigned int uiAllocateFlags);   /* odeinit.h:190 */
void dCleanupODEAllDataForThread(void);
dReal dGeomSpherePointDepth(dGeomID sphere, dReal x, dReal y, dReal z);    /* collision.h:851 */
dReal dGeomBoxPointDepth(dGeomID box, dReal x, dReal y, dReal z);
dReal dGeomPlanePointDepth(dGeomID plane, dReal x, dReal y, dReal z);
dReal dGeomCapsulePointDepth(dGeomID ccylinder, dReal x, dReal y, dReal z);
void dJointAddHinge2Torques(dJointID joint, dReal torque1, dReal torque2);   /* objects.h:1997 */
void dMassSetCappedCylinder(dMass *m, dReal density, int direction, dReal radius, dReal length);        /* mass.h:84, deprecated alias of the capsule */
void dMassSetCappedCylinderTotal(dMass *m, dReal total_mass, int direction, dReal radius, dReal length);

/* ---- export (include/ode/export-dif.h:31, ode/src/export-dif.cpp): text dump of a world in the reference's
 * "Dynamics Interchange Format v0.1"; every name is prefixed with `world_name` */
void dWorldExportDIF(dWorldID w, FILE *file, const char *world_name);

/* ---- mass (mass.h:43-140) ------------------------------------------------- */
int dMassCheck(const dMass *m);
void dMassSetZero(dMass *m);
void dMassSetParameters(dMass *m, dReal themass, dReal cgx, dReal cgy, dReal cgz,
                        dReal I11, dReal I22, dReal I33, dReal I12, dReal I13, dReal I23);
void dMassSetSphere(dMass *m, dReal density, dReal radius);
void dMassSetSphereTotal(dMass *m, dReal total_mass, dReal radius);
void dMassSetCapsule(dMass *m, dReal density, int direction, dReal radius, dReal length);
void dMassSetCapsuleTotal(dMass *m, dReal total_mass, int direction, dReal radius, dReal length);
void dMassSetCylinder(dMass *m, dReal density, int direction, dReal radius, dReal length);          /* include/ode/mass.h:60 */
void dMassSetCylinderTotal(dMass *m, dReal total_mass, int direction, dReal radius, dReal length);   /* include/ode/mass.h:62 */
void dMassSetBox(dMass *m, dReal density, dReal lx, dReal ly, dReal lz);
void dMassSetBoxTotal(dMass *m, dReal total_mass, dReal lx, dReal ly, dReal lz);
void dMassAdjust(dMass *m, dReal newmass);
void dMassSetTrimesh(dMass *m, dReal density, dGeomID g);             /* include/ode/mass.h:83, mass.cpp:234-427 */
void dMassSetTrimeshTotal(dMass *m, dReal total_mass, dGeomID g);     /* include/ode/mass.h:84 */
void dMassTranslate(dMass *m, dReal x, dReal y, dReal z);
void dMassRotate(dMass *m, const dMatrix3 R);
void dMassAdd(dMass *a, const dMass *b);

/* ---- world (objects.h:54-560) --------------------------------------------- */
dWorldID dWorldCreate(void);
void dWorldDestroy(dWorldID world);
void dWorldSetGravity(dWorldID, dReal x, dReal y, dReal z);
void dWorldGetGravity(dWorldID, dVector3 gravity);
void dWorldSetERP(dWorldID, dReal erp);
dReal dWorldGetERP(dWorldID);
void dWorldSetCFM(dWorldID, dReal cfm);
dReal dWorldGetCFM(dWorldID);
int dWorldQuickStep(dWorldID w, dReal stepsize); /* objects.h:352 */
void dWorldSetQuickStepNumIterations(dWorldID, int num);
int dWorldGetQuickStepNumIterations(dWorldID);
void dWorldSetQuickStepW(dWorldID, dReal over_relaxation);
dReal dWorldGetQuickStepW(dWorldID);
void dWorldSetContactMaxCorrectingVel(dWorldID, dReal vel);
dReal dWorldGetContactMaxCorrectingVel(dWorldID);
void dWorldSetContactSurfaceLayer(dWorldID, dReal depth);
dReal dWorldGetContactSurfaceLayer(dWorldID);
void dWorldSetAutoDisableFlag(dWorldID, int do_auto_disable);
int dWorldGetAutoDisableFlag(dWorldID);
void dWorldSetAutoDisableLinearThreshold(dWorldID, dReal v);
void dWorldSetAutoDisableAngularThreshold(dWorldID, dReal v);
void dWorldSetAutoDisableAverageSamplesCount(dWorldID, unsigned int n);
void dWorldSetAutoDisableSteps(dWorldID, int steps);
void dWorldSetAutoDisableTime(dWorldID, dReal time);
void dWorldSetLinearDamping(dWorldID, dReal scale);
void dWorldSetAngularDamping(dWorldID, dReal scale);
void dWorldSetLinearDampingThreshold(dWorldID, dReal threshold);
void dWorldSetAngularDampingThreshold(dWorldID, dReal threshold);
void dWorldSetDamping(dWorldID, dReal linear_scale, dReal angular_scale);
void dWorldSetMaxAngularSpeed(dWorldID, dReal max_speed);

/* ---- body (objects.h:560-1500) -------------------------------------------- */
dBodyID dBodyCreate(dWorldID);
void dBodyDestroy(dBodyID);
dWorldID dBodyGetWorld(dBodyID);
void dBodySetData(dBodyID, void *data);
void *dBodyGetData(dBodyID);
void dBodySetPosition(dBodyID, dReal x, dReal y, dReal z);
void dBodySetRotation(dBodyID, const dMatrix3 R);
void dBodySetQuaternion(dBodyID, const dQuaternion q);
void dBodySetLinearVel(dBodyID, dReal x, dReal y, dReal z);
void dBodySetAngularVel(dBodyID, dReal x, dReal y, dReal z);
const dReal *dBodyGetPosition(dBodyID);
const dReal *dBodyGetRotation(dBodyID);
const dReal *dBodyGetQuaternion(dBodyID);
const dReal *dBodyGetLinearVel(dBodyID);
const dReal *dBodyGetAngularVel(dBodyID);
void dBodySetMass(dBodyID, const dMass *mass);
void dBodyGetMass(dBodyID, dMass *mass);
void dBodyAddForce(dBodyID, dReal fx, dReal fy, dReal fz);
void dBodyGetRelPointVel(dBodyID, dReal px, dReal py, dReal pz, dVector3 result);   /* include/ode/objects.h:1113 */
void dBodyAddTorque(dBodyID, dReal fx, dReal fy, dReal fz);
void dBodyAddRelForce(dBodyID, dReal fx, dReal fy, dReal fz);
void dBodyAddRelTorque(dBodyID, dReal fx, dReal fy, dReal fz);
void dBodyAddForceAtPos(dBodyID, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz);
const dReal *dBodyGetForce(dBodyID);
const dReal *dBodyGetTorque(dBodyID);
void dBodySetForce(dBodyID, dReal x, dReal y, dReal z);
void dBodySetTorque(dBodyID, dReal x, dReal y, dReal z);
void dBodyEnable(dBodyID);
void dBodyDisable(dBodyID);
int dBodyIsEnabled(dBodyID);
void dBodySetGravityMode(dBodyID, int mode);
int dBodyGetGravityMode(dBodyID);
void dBodySetFiniteRotationMode(dBodyID, int mode);
void dBodySetFiniteRotationAxis(dBodyID, dReal x, dReal y, dReal z);
void dBodySetGyroscopicMode(dBodyID, int enabled);
int dBodyGetGyroscopicMode(dBodyID);
void dBodySetAutoDisableFlag(dBodyID, int do_auto_disable);
int dBodyGetAutoDisableFlag(dBodyID);
void dBodySetAutoDisableDefaults(dBodyID);
void dBodySetAutoDisableAverageSamplesCount(dBodyID, unsigned int n);
void dBodySetLinearDamping(dBodyID, dReal scale);
void dBodySetAngularDamping(dBodyID, dReal scale);
void dBodySetDampingDefaults(dBodyID);
void dBodySetMaxAngularSpeed(dBodyID, dReal max_speed);
int dBodyGetNumJoints(dBodyID);
dGeomID dBodyGetFirstGeom(dBodyID);
dGeomID dBodyGetNextGeom(dGeomID);
void dBodyGetRelPointPos(dBodyID, dReal px, dReal py, dReal pz, dVector3 result);
void dBodyVectorToWorld(dBodyID, dReal px, dReal py, dReal pz, dVector3 result);
void dBodyVectorFromWorld(dBodyID, dReal px, dReal py, dReal pz, dVector3 result);   /* include/ode/objects.h:1154 */
void dBodyGetPointVel(dBodyID, dReal px, dReal py, dReal pz, dVector3 result);       /* include/ode/objects.h:1121 */
void dBodyGetPosRelPoint(dBodyID, dReal px, dReal py, dReal pz, dVector3 result);    /* include/ode/objects.h:1132 */
void dWorldImpulseToForce(dWorldID, dReal stepsize, dReal ix, dReal iy, dReal iz, dVector3 force);   /* ode.cpp:1828-1838 */
/* geom-frame conversions, collision_kernel.cpp:784-866 (non-placeable geoms return the argument) */
void dGeomGetRelPointPos(dGeomID geom, dReal px, dReal py, dReal pz, dVector3 result);
void dGeomGetPosRelPoint(dGeomID geom, dReal px, dReal py, dReal pz, dVector3 result);
void dGeomVectorToWorld(dGeomID geom, dReal px, dReal py, dReal pz, dVector3 result);
void dGeomVectorFromWorld(dGeomID geom, dReal px, dReal py, dReal pz, dVector3 result);
/* collision_util.cpp:109-219; misc.cpp:128-136 */
void dClosestLineSegmentPoints(const dVector3 a1, const dVector3 a2, const dVector3 b1, const dVector3 b2, dVector3 cp1, dVector3 cp2);
void dPrintMatrix(const dReal *A, int n, int m, char *fmt, FILE *f);

/* ---- joints (objects.h:1538-2700) ----------------------------------------- */
dJointGroupID dJointGroupCreate(int max_size);
void dJointGroupDestroy(dJointGroupID);
void dJointGroupEmpty(dJointGroupID);
dJointID dJointCreateContact(dWorldID, dJointGroupID, const dContact *);
dJointID dJointCreateBall(dWorldID, dJointGroupID);
dJointID dJointCreateHinge(dWorldID, dJointGroupID);
dJointID dJointCreateHinge2(dWorldID, dJointGroupID);
dJointID dJointCreateSlider(dWorldID, dJointGroupID);   /* include/ode/objects.h:1571, ode/src/joints/slider.cpp */
dJointID dJointCreateUniversal(dWorldID, dJointGroupID);   /* include/ode/objects.h:1592, ode/src/joints/universal.cpp */
dJointID dJointCreateFixed(dWorldID, dJointGroupID);
dJointID dJointCreateAMotor(dWorldID, dJointGroupID);   /* include/ode/objects.h:1634, ode/src/joints/amotor.cpp */
dJointID dJointCreateLMotor(dWorldID, dJointGroupID);   /* include/ode/objects.h:1643, ode/src/joints/lmotor.cpp */    /* include/ode/objects.h:1616, ode/src/joints/fixed.cpp */
dJointID dJointCreatePR(dWorldID, dJointGroupID);        /* include/ode/objects.h:1586, ode/src/joints/pr.cpp */
dJointID dJointCreatePU(dWorldID, dJointGroupID);        /* include/ode/objects.h:1594, ode/src/joints/pu.cpp */
dJointID dJointCreatePiston(dWorldID, dJointGroupID);    /* include/ode/objects.h:1603, ode/src/joints/piston.cpp */
dJointID dJointCreateNull(dWorldID, dJointGroupID);      /* include/ode/objects.h:1625, ode/src/joints/null.cpp */
dJointID dJointCreatePlane2D(dWorldID, dJointGroupID);   /* include/ode/objects.h:1637, ode/src/joints/plane2d.cpp */
void dJointDestroy(dJointID);
void dJointAttach(dJointID, dBodyID body1, dBodyID body2);
void dJointEnable(dJointID);
void dJointDisable(dJointID);
int dJointIsEnabled(dJointID);
dJointType dJointGetType(dJointID);
dBodyID dJointGetBody(dJointID, int index);
void dJointSetFeedback(dJointID, dJointFeedback *);
dJointFeedback *dJointGetFeedback(dJointID);
void dJointSetSliderAxis(dJointID, dReal x, dReal y, dReal z);
void dJointSetSliderAxisDelta(dJointID, dReal x, dReal y, dReal z, dReal ax, dReal ay, dReal az);
void dJointGetSliderAxis(dJointID, dVector3 result);
void dJointSetSliderParam(dJointID, int parameter, dReal value);
dReal dJointGetSliderParam(dJointID, int parameter);
dReal dJointGetSliderPosition(dJointID);
dReal dJointGetSliderPositionRate(dJointID);
void dJointAddSliderForce(dJointID, dReal force);
/* plane2d (include/ode/objects.h:2332-2346): body 1 is kept in the plane z = 0 of the static environment;
 * optional motors along x, y and about z.  A second body is refused by dBatchCreate / dWorldQuickStep. */
void dJointSetPlane2DXParam(dJointID, int parameter, dReal value);
void dJointSetPlane2DYParam(dJointID, int parameter, dReal value);
void dJointSetPlane2DAngleParam(dJointID, int parameter, dReal value);
/* piston (include/ode/objects.h:2169-2215, 2790-2870): prismatic + rotoide about the same axis; parameter
 * group 1 = prismatic limit-motor, group 2 (dParamLoStop2 ...) = rotoide */
void dJointSetPistonAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetPistonAnchorOffset(dJointID, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz);
void dJointGetPistonAnchor(dJointID, dVector3 result);
void dJointGetPistonAnchor2(dJointID, dVector3 result);
void dJointSetPistonAxis(dJointID, dReal x, dReal y, dReal z);
void dJointGetPistonAxis(dJointID, dVector3 result);
void dJointSetPistonAxisDelta(dJointID j, dReal x, dReal y, dReal z, dReal ax, dReal ay, dReal az);   /* :2204 */
void dJointSetPistonParam(dJointID, int parameter, dReal value);
dReal dJointGetPistonParam(dJointID, int parameter);
dReal dJointGetPistonPosition(dJointID);
dReal dJointGetPistonPositionRate(dJointID);
dReal dJointGetPistonAngle(dJointID);
dReal dJointGetPistonAngleRate(dJointID);
void dJointAddPistonForce(dJointID, dReal force);
/* PR, prismatic + rotoide about different axes (include/ode/objects.h:2035-2062, 2640-2700):
 * axis 1 = prismatic, axis 2 = rotoide */
void dJointSetPRAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetPRAxis1(dJointID, dReal x, dReal y, dReal z);
void dJointSetPRAxis2(dJointID, dReal x, dReal y, dReal z);
void dJointGetPRAnchor(dJointID, dVector3 result);
void dJointGetPRAxis1(dJointID, dVector3 result);
void dJointGetPRAxis2(dJointID, dVector3 result);
void dJointSetPRParam(dJointID, int parameter, dReal value);
dReal dJointGetPRParam(dJointID, int parameter);
dReal dJointGetPRPosition(dJointID);
dReal dJointGetPRPositionRate(dJointID);
dReal dJointGetPRAngle(dJointID);
dReal dJointGetPRAngleRate(dJointID);
void dJointAddPRTorque(dJointID, dReal torque);
/* PU, prismatic + universal (include/ode/objects.h:2072-2160, 2700-2790): axes 1, 2 = universal, axis 3 (P) =
 * prismatic; parameter groups 1, 2 = the universal axes, group 3 (dParamLoStop3 ...) = prismatic */
void dJointSetPUAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetPUAnchorDelta(dJointID, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz);
void dJointSetPUAnchorOffset(dJointID, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz);
void dJointSetPUAxis1(dJointID, dReal x, dReal y, dReal z);
void dJointSetPUAxis2(dJointID, dReal x, dReal y, dReal z);
void dJointSetPUAxis3(dJointID, dReal x, dReal y, dReal z);
void dJointSetPUAxisP(dJointID, dReal x, dReal y, dReal z);
void dJointGetPUAnchor(dJointID, dVector3 result);
void dJointGetPUAxis1(dJointID, dVector3 result);
void dJointGetPUAxis2(dJointID, dVector3 result);
void dJointGetPUAxis3(dJointID, dVector3 result);
void dJointGetPUAxisP(dJointID, dVector3 result);
void dJointSetPUParam(dJointID, int parameter, dReal value);
dReal dJointGetPUParam(dJointID, int parameter);
void dJointGetPUAngles(dJointID, dReal *angle1, dReal *angle2);
dReal dJointGetPUAngle1(dJointID);
dReal dJointGetPUAngle2(dJointID);
dReal dJointGetPUAngle1Rate(dJointID);
dReal dJointGetPUAngle2Rate(dJointID);
dReal dJointGetPUPosition(dJointID);
dReal dJointGetPUPositionRate(dJointID);
void dJointSetUniversalAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetUniversalAxis1(dJointID, dReal x, dReal y, dReal z);
void dJointSetUniversalAxis2(dJointID, dReal x, dReal y, dReal z);
void dJointGetUniversalAnchor(dJointID, dVector3 result);
void dJointGetUniversalAnchor2(dJointID, dVector3 result);
void dJointGetUniversalAxis1(dJointID, dVector3 result);
void dJointGetUniversalAxis2(dJointID, dVector3 result);
void dJointSetUniversalParam(dJointID, int parameter, dReal value);
dReal dJointGetUniversalParam(dJointID, int parameter);
void dJointGetUniversalAngles(dJointID, dReal *angle1, dReal *angle2);
dReal dJointGetUniversalAngle1(dJointID);
dReal dJointGetUniversalAngle2(dJointID);
void dJointSetAMotorNumAxes(dJointID, int num);
void dJointSetAMotorAxis(dJointID, int anum, int rel, dReal x, dReal y, dReal z);
void dJointSetAMotorAngle(dJointID, int anum, dReal angle);
void dJointSetAMotorParam(dJointID, int parameter, dReal value);
void dJointSetAMotorMode(dJointID, int mode);
int dJointGetAMotorNumAxes(dJointID);
void dJointGetAMotorAxis(dJointID, int anum, dVector3 result);
int dJointGetAMotorAxisRel(dJointID, int anum);
dReal dJointGetAMotorAngle(dJointID, int anum);
dReal dJointGetAMotorAngleRate(dJointID, int anum);   /* objects.h; amotor.cpp:445-451: dDebug("not yet implemented") in the reference, here too */
dReal dJointGetAMotorParam(dJointID, int parameter);
int dJointGetAMotorMode(dJointID);
void dJointSetLMotorNumAxes(dJointID, int num);
void dJointSetLMotorAxis(dJointID, int anum, int rel, dReal x, dReal y, dReal z);
void dJointSetLMotorParam(dJointID, int parameter, dReal value);
int dJointGetLMotorNumAxes(dJointID);
void dJointGetLMotorAxis(dJointID, int anum, dVector3 result);
dReal dJointGetLMotorParam(dJointID, int parameter);
void dJointSetFixed(dJointID);
void dJointSetFixedParam(dJointID, int parameter, dReal value);
dReal dJointGetFixedParam(dJointID, int parameter);
void dJointSetBallAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetBallAnchor2(dJointID, dReal x, dReal y, dReal z);
void dJointSetBallParam(dJointID, int parameter, dReal value);
void dJointGetBallAnchor(dJointID, dVector3 result);
void dJointGetBallAnchor2(dJointID, dVector3 result);
void dJointSetHingeAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetHingeAxis(dJointID, dReal x, dReal y, dReal z);
void dJointSetHingeParam(dJointID, int parameter, dReal value);
dReal dJointGetHingeParam(dJointID, int parameter);
void dJointGetHingeAnchor(dJointID, dVector3 result);
void dJointGetHingeAnchor2(dJointID, dVector3 result);
void dJointGetHingeAxis(dJointID, dVector3 result);
dReal dJointGetHingeAngle(dJointID);
dReal dJointGetHingeAngleRate(dJointID);
void dJointSetHinge2Anchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetHinge2Axis1(dJointID, dReal x, dReal y, dReal z);
void dJointSetHinge2Axis2(dJointID, dReal x, dReal y, dReal z);
void dJointSetHinge2Param(dJointID, int parameter, dReal value);
dReal dJointGetHinge2Param(dJointID, int parameter);
void dJointGetHinge2Anchor(dJointID, dVector3 result);
void dJointGetHinge2Anchor2(dJointID, dVector3 result);
void dJointGetHinge2Axis1(dJointID, dVector3 result);
void dJointGetHinge2Axis2(dJointID, dVector3 result);
dReal dJointGetHinge2Angle1(dJointID);
dReal dJointGetHinge2Angle1Rate(dJointID);
dReal dJointGetHinge2Angle2Rate(dJointID);
int dAreConnected(dBodyID, dBodyID);
int dAreConnectedExcluding(dBodyID body1, dBodyID body2, int joint_type);

/* ---- collision (collision.h, collision_space.h) --------------------------- */
dSpaceID dSimpleSpaceCreate(dSpaceID space);
dSpaceID dHashSpaceCreate(dSpaceID space);
dSpaceID dSweepAndPruneSpaceCreate(dSpaceID space, int axisorder);
void dSpaceDestroy(dSpaceID);
void dHashSpaceSetLevels(dSpaceID space, int minlevel, int maxlevel);
void dHashSpaceGetLevels(dSpaceID space, int *minlevel, int *maxlevel);
void dSpaceSetCleanup(dSpaceID space, int mode);
int dSpaceGetCleanup(dSpaceID space);
void dSpaceSetSublevel(dSpaceID space, int sublevel);
int dSpaceGetSublevel(dSpaceID space);
void dSpaceAdd(dSpaceID, dGeomID);
void dSpaceRemove(dSpaceID, dGeomID);
int dSpaceQuery(dSpaceID, dGeomID);
void dSpaceClean(dSpaceID);
int dSpaceGetNumGeoms(dSpaceID);
dGeomID dSpaceGetGeom(dSpaceID, int i);
void dSpaceCollide(dSpaceID space, void *data, dNearCallback *callback);  /* collision_space.h:72-ish / collision.h:795 */
void dSpaceCollide2(dGeomID space1, dGeomID space2, void *data, dNearCallback *callback);
/* Contact capacity per pair (the one documented limit of this path): a collider call generates at most 16 contacts.  The
 * reference returns up to (flags & 0xffff) contacts for trimesh-* and cylinder-box pairs; a larger request is served with 16
 * and reported once through the message handler.  The batched narrowphase (dBatchCollideAndQuickStep, and the results
 * dCollide serves from dSpaceCollide's batch) keeps at most 8 contacts per pair in batches without trimesh geoms or geom
 * transforms -- every primitive collider but cylinder-box emits at most 8 -- and 16 otherwise; dCollide falls back to an
 * on-demand collider call whenever the cached result was computed with another limit than the caller's. */
int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom *contact, int skip); /* collision.h:747 */

void dGeomDestroy(dGeomID);
void dGeomSetData(dGeomID, void *data);
void *dGeomGetData(dGeomID);
void dGeomSetBody(dGeomID, dBodyID);
dBodyID dGeomGetBody(dGeomID);
void dGeomSetPosition(dGeomID, dReal x, dReal y, dReal z);
void dGeomSetRotation(dGeomID, const dMatrix3 R);
void dGeomSetQuaternion(dGeomID, const dQuaternion Q);
const dReal *dGeomGetPosition(dGeomID);
const dReal *dGeomGetRotation(dGeomID);
void dGeomGetQuaternion(dGeomID, dQuaternion result);
void dGeomGetAABB(dGeomID, dReal aabb[6]);
int dGeomIsSpace(dGeomID);
dSpaceID dGeomGetSpace(dGeomID);
int dGeomGetClass(dGeomID);
void dGeomSetCategoryBits(dGeomID, unsigned long bits);
void dGeomSetCollideBits(dGeomID, unsigned long bits);
unsigned long dGeomGetCategoryBits(dGeomID);
unsigned long dGeomGetCollideBits(dGeomID);
void dGeomEnable(dGeomID);
void dGeomDisable(dGeomID);
int dGeomIsEnabled(dGeomID);
void dGeomSetOffsetPosition(dGeomID, dReal x, dReal y, dReal z);
void dGeomSetOffsetRotation(dGeomID, const dMatrix3 R);
void dGeomSetOffsetQuaternion(dGeomID, const dQuaternion Q);
void dGeomClearOffset(dGeomID);
int dGeomIsOffset(dGeomID);

dGeomID dCreateSphere(dSpaceID space, dReal radius);
void dGeomSphereSetRadius(dGeomID sphere, dReal radius);
dReal dGeomSphereGetRadius(dGeomID sphere);
dGeomID dCreateBox(dSpaceID space, dReal lx, dReal ly, dReal lz);
void dGeomBoxSetLengths(dGeomID box, dReal lx, dReal ly, dReal lz);
void dGeomBoxGetLengths(dGeomID box, dVector3 result);
dGeomID dCreatePlane(dSpaceID space, dReal a, dReal b, dReal c, dReal d);
void dGeomPlaneSetParams(dGeomID plane, dReal a, dReal b, dReal c, dReal d);
void dGeomPlaneGetParams(dGeomID plane, dVector4 result);
dGeomID dCreateCapsule(dSpaceID space, dReal radius, dReal length);
void dGeomCapsuleSetParams(dGeomID ccylinder, dReal radius, dReal length);
void dGeomCapsuleGetParams(dGeomID ccylinder, dReal *radius, dReal *length);
/* flat-ended cylinder, include/ode/collision.h:1063-1065 (ode/src/cylinder.cpp); colliders against plane, sphere
 * and box (ode/src/collision_cylinder_{plane,sphere,box}.cpp) */
dGeomID dCreateCylinder(dSpaceID space, dReal radius, dReal length);
void dGeomCylinderSetParams(dGeomID cylinder, dReal radius, dReal length);
void dGeomCylinderGetParams(dGeomID cylinder, dReal *radius, dReal *length);

/* Geom transforms (reference: include/ode/collision.h:1086-1092, ode/src/collision_transform.cpp).  A transform
 * geom T encapsulates one geom that is in no space and on no body; T collides as that geom posed at T o local.
 * Encapsulated classes served: sphere, box, capsule, cylinder.  Contacts report the encapsulated geom in g1/g2
 * unless dGeomTransformSetInfo(T, 1). */
dGeomID dCreateGeomTransform(dSpaceID space);
void dGeomTransformSetGeom(dGeomID g, dGeomID obj);
dGeomID dGeomTransformGetGeom(dGeomID g);
void dGeomTransformSetCleanup(dGeomID g, int mode);
int dGeomTransformGetCleanup(dGeomID g);
void dGeomTransformSetInfo(dGeomID g, int mode);
int dGeomTransformGetInfo(dGeomID g);
/* rays: include/ode/collision.h:1025-1067, ode/src/ray.cpp:91-189 */
dGeomID dCreateRay(dSpaceID space, dReal length);
void dGeomRaySetLength(dGeomID ray, dReal length);
dReal dGeomRayGetLength(dGeomID ray);
void dGeomRaySet(dGeomID ray, dReal px, dReal py, dReal pz, dReal dx, dReal dy, dReal dz);
void dGeomRayGet(dGeomID ray, dVector3 start, dVector3 dir);
void dGeomRaySetParams(dGeomID g, int FirstContact, int BackfaceCull);
void dGeomRayGetParams(dGeomID g, int *FirstContact, int *BackfaceCull);
void dGeomRaySetClosestHit(dGeomID g, int closestHit);
int dGeomRayGetClosestHit(dGeomID g);

/* ---- trimesh (collision_trimesh.h:51-153).  Vertex and index arrays are COPIED at build time
 * (the reference borrows them); per-triangle callbacks are not supported by the GPU colliders. */
typedef int dTriCallback(dGeomID TriMesh, dGeomID RefObject, int TriangleIndex);
typedef void dTriArrayCallback(dGeomID TriMesh, dGeomID RefObject, const int *TriIndices, int TriCount);
typedef int dTriRayCallback(dGeomID TriMesh, dGeomID Ray, int TriangleIndex, dReal u, dReal v);
dTriMeshDataID dGeomTriMeshDataCreate(void);
void dGeomTriMeshDataDestroy(dTriMeshDataID g);
void dGeomTriMeshDataBuildSingle(dTriMeshDataID g, const void *Vertices, int VertexStride, int VertexCount,
                                 const void *Indices, int IndexCount, int TriStride);
void dGeomTriMeshDataBuildSingle1(dTriMeshDataID g, const void *Vertices, int VertexStride, int VertexCount,
                                  const void *Indices, int IndexCount, int TriStride, const void *Normals);
void dGeomTriMeshDataBuildDouble(dTriMeshDataID g, const void *Vertices, int VertexStride, int VertexCount,
                                 const void *Indices, int IndexCount, int TriStride);
void dGeomTriMeshDataBuildDouble1(dTriMeshDataID g, const void *Vertices, int VertexStride, int VertexCount,
                                  const void *Indices, int IndexCount, int TriStride, const void *Normals);
void dGeomTriMeshDataBuildSimple(dTriMeshDataID g, const dReal *Vertices, int VertexCount,
                                 const dTriIndex *Indices, int IndexCount);
void dGeomTriMeshDataBuildSimple1(dTriMeshDataID g, const dReal *Vertices, int VertexCount, const dTriIndex *Indices, int IndexCount, const int *Normals);   /* collision_trimesh.h:127 */
void dGeomTriMeshDataPreprocess(dTriMeshDataID g);
void dGeomTriMeshDataUpdate(dTriMeshDataID g);
dGeomID dCreateTriMesh(dSpaceID space, dTriMeshDataID Data, dTriCallback *Callback,
                       dTriArrayCallback *ArrayCallback, dTriRayCallback *RayCallback);
void dGeomTriMeshSetData(dGeomID g, dTriMeshDataID Data);
dTriMeshDataID dGeomTriMeshGetData(dGeomID g);
int dGeomTriMeshGetTriangleCount(dGeomID g);

/* ======================================================================== */
/* (2) batched-world entry points (added; SURVEY.md §8(b) last row)          */
/* ======================================================================== */

typedef struct dxBatch *dBatchID;

/* One row of the contact-policy table that replaces the near callback on the
 * batched path.  It expresses what the reference demos' callbacks do
 * (ode/demo/demo_boxstack.cpp:132-172, demo_crash.cpp:115-141,
 * demo_buggy.cpp:83-111): optionally skip pairs whose bodies are connected by
 * a non-contact joint, call dCollide with max_contacts, copy `surface` into
 * every contact and create+attach one contact joint per contact.
 * Row selection (1 to 8 rows): the first row with (cat(o1)&cat_mask1)&&(cat(o2)&cat_mask2)
 * or the swapped test serves the pair -- put the specific rows first and a catch-all
 * {~0,~0} last; a pair that no row accepts gets no contacts (a callback that returns
 * early).  A table of ONE row serves every pair, whatever its masks say.  This is how
 * demo_crash.cpp:128-131 (mu = 20 for pairs with a sphere, 0.5 otherwise) is written:
 * the spheres get a category bit of their own, row 0 = {SPHERE_BIT, ~0, ..., mu 20},
 * row 1 = {~0, ~0, ..., mu 0.5}.  The large-world path takes one row. */
typedef struct dBatchContactPolicy {
  unsigned long cat_mask1, cat_mask2;
  int max_contacts;                 /* flags & 0xffff handed to dCollide */
  int skip_if_connected;            /* dAreConnectedExcluding(b1,b2,Contact) */
  int skip_static_pairs;            /* return when neither geom has a body */
  dSurfaceParameters surface;
} dBatchContactPolicy;

/* capacity of one world slot; 0 = derive from the bound worlds */
typedef struct dBatchDesc {
  int max_contacts_per_world;       /* contact joints per step per world */
  int device;                       /* CUDA device ordinal */
  int max_pairs_per_world;          /* near-callback pairs per step per world (0: 12 x geoms) */
  int large_world;                  /* 1: run the (single) world on the grid-wide large-world path: device SAP
                                     * sweep + graph-coloured SOR (one CTA cannot hold it).  Chosen automatically
                                     * for a single world with more than 254 bodies.  Contact joints only,
                                     * dSweepAndPruneSpace only; SOR order differs from the reference's random
                                     * order, so trajectories match it within a tolerance, not bit for bit. */
  int reserved[4];
} dBatchDesc;

typedef struct dBatchCounters {
  long long steps;                  /* world-steps executed */
  long long body_steps;             /* bodies integrated (enabled bodies in islands) */
  long long pairs;                  /* near-callback-equivalent pairs */
  long long contacts;               /* contact joints that entered SOR */
  long long rows;                   /* constraint rows m summed over islands */
  long long islands;
  long long overflow_worlds;        /* world-steps that hit a capacity limit */
} dBatchCounters;

/* Bind nworlds (world[i], space[i]) pairs into one device-resident batch.
 * Objects were created through the normal API above; after binding, the
 * device copy is authoritative until dBatchDownload().  Returns NULL on
 * failure (message via the error handler), never aborts. */
dBatchID dBatchCreate(int nworlds, const dWorldID *worlds, const dSpaceID *spaces,
                      const dBatchDesc *desc);
void dBatchDestroy(dBatchID);
int dBatchSetContactPolicy(dBatchID, const dBatchContactPolicy *table, int n);
/* per-world LCG streams for the SOR row shuffle (misc.cpp:31 is process-global
 * in the reference; one stream per world here) */
int dBatchSetSeeds(dBatchID, const uint32_t *seeds);
int dBatchGetSeeds(dBatchID, uint32_t *seeds);
/* nsteps x { dSpaceCollide + policy + dWorldQuickStep(h) + dJointGroupEmpty }
 * for every world.  status_per_world (may be NULL) receives 0 = ok or a
 * bitmask of dBATCH_ERR_*.  Returns 0 on success. */
int dBatchCollideAndQuickStep(dBatchID, dReal h, int nsteps, int *status_per_world);
#define dBATCH_ERR_CONTACT_OVERFLOW 1
#define dBATCH_ERR_ROW_OVERFLOW 2
#define dBATCH_ERR_PAIR_OVERFLOW 4
#define dBATCH_ERR_BVH_STACK 8        /* trimesh tree deeper than the traversal stack */
/* bulk SoA I/O, host buffers: [world][body][k]; body order = creation order.
 * pos 3, quat 4, lvel 3, avel 3 (13 reals per body).  Set stores the quaternion as given
 * (it must be unit; the rotation matrix is rebuilt from it), so Get -> Set restores bit for bit. */
/* order state for snapshots (opaque ints): the space list order, SAP sort ranks and dirty counts that the
 * next step's callback order depends on besides body state and seeds */
int dBatchOrderStateSize(dBatchID);
int dBatchGetOrderState(dBatchID, int *buf);
int dBatchSetOrderState(dBatchID, const int *buf);
int dBatchNumBodies(dBatchID);            /* per world (max over worlds) */
int dBatchGetBodyState(dBatchID, dReal *pos3, dReal *quat4, dReal *lvel3, dReal *avel3);
int dBatchSetBodyState(dBatchID, const dReal *pos3, const dReal *quat4,
                       const dReal *lvel3, const dReal *avel3);
/* external force/torque accumulators, [world][body][3] each, added to facc/tacc */
int dBatchAddForces(dBatchID, const dReal *force3, const dReal *torque3);
/* page-locked host memory: buffers from dBatchHostAlloc make the bulk I/O calls above plain
 * DMA transfers (any other host pointer works too, through the driver's staging copy) */
void *dBatchHostAlloc(size_t bytes);
void dBatchHostFree(void *p);
/* copy device state back into the bound dBodyID/dGeomID objects */
int dBatchDownload(dBatchID);
int dBatchGetCounters(dBatchID, dBatchCounters *out);
int dBatchResetCounters(dBatchID);
/* parity taps for one world's LAST step (tests): callback-order pair list as
 * geom creation indices, contact records, SOR row count, lambda.
 * Each returns the number of items available; copies at most cap. */
int dBatchDebugPairs(dBatchID, int world, int *g1g2, int cap);
int dBatchDebugContacts(dBatchID, int world, dReal *pos_normal_depth7, int *g1g2, int cap);
int dBatchDebugLambda(dBatchID, int world, dReal *lambda, int cap);
/* f1,t1 (6 reals) per contact joint, as dJointSetFeedback would report */
int dBatchDebugFeedback(dBatchID, int world, dReal *f1t1, int cap);
/* geom creation indices in space-list order (head first) as of now */
int dBatchDebugGeomOrder(dBatchID, int world, int *order, int cap);
/* parity taps cost an extra copy of J per row; on by default, switch off for timing */
int dBatchSetDebugTaps(dBatchID, int enable);
/* CUDA-event timing on the batch's own launching stream: Start records an event,
 * Stop records a second one, synchronises and returns the elapsed milliseconds */
int dBatchTimerStart(dBatchID);
int dBatchTimerStop(dBatchID, float *ms);
/* profiling mode: bracket every kernel launch with events and accumulate device
 * time per kernel (serialises steps; use only to attribute time, not for throughput) */
int dBatchSetKernelTiming(dBatchID, int enable);
int dBatchGetKernelTimes(dBatchID, double *ms_per_kernel, long long *launches_per_kernel, int nkernels);
const char *dBatchKernelName(int k);
/* large-world path (dBatchDesc.large_world): figures of the last step and, while kernel timing is
 * on, CUDA-event time per phase summed over steps_timed steps:
 * phase_ms = {geoms + radix sort, pair sweep, narrowphase, colouring, row assembly, SOR, integration} */
typedef struct dBatchLargeWorldStats {
  int pairs, contacts, contact_pairs, solved_contacts, colours, colouring_rounds, sor_launches, steps_timed;
  double phase_ms[7];
} dBatchLargeWorldStats;
int dBatchGetLargeWorldStats(dBatchID, dBatchLargeWorldStats *out);   /* -1 for a batch of small worlds */
/* One large world over the GPUs of one box (BASELINE.json configs[4] "2/4/8-GPU split").
 * One process per GPU; every rank builds the SAME world and binds it into a large-world batch of its own.
 * Broadphase, narrowphase, colouring and row assembly are deterministic and run on every rank; the SOR
 * phase -- the reference's SOR_LCP (ode/src/quickstep.cpp:342-584) as the coloured sweep -- is dealt over
 * the ranks: each rank sweeps its share of every colour's contact pairs and stores the constraint forces
 * it produced into every rank's array through NVLink peer mappings, inside the kernel, with a flag barrier
 * per colour.  All ranks end every step with the same body state, bit for bit the single-GPU result.
 *   dBatchSplitExport: fills a D_BATCH_SPLIT_HANDLE_BYTES description of this rank's exchange buffer
 *                      (CUDA IPC handle); the caller gathers the ranks' descriptions (any transport).
 *   dBatchSplitAttach: all descriptions in rank order.  From then on dBatchCollideAndQuickStep must be
 *                      called by every rank with the same arguments; a rank that stays away makes the
 *                      others fail with an error after OB_LW_SPLIT_TIMEOUT_MS (default 5000), not hang. */
#define D_BATCH_SPLIT_HANDLE_BYTES 128
int dBatchSplitExport(dBatchID, void *handle);
int dBatchSplitAttach(dBatchID, int rank, int nranks, const void *handles);
/* Ray queries against the bound worlds (the batched form of the ray-cast vehicle's wheel probe, demos/raycar/car.cpp:353-371:
 * dGeomRaySet + dGeomRaySetParams + dGeomRaySetClosestHit + dSpaceCollide2(space, ray, callback keeping the nearest
 * dCollide result)).  rays_per_world rays per world: origin3 / dir3 are [world][ray][3] (dir is normalised like dGeomRaySet
 * does, ray.cpp:115-135), length [world][ray].  ray_flags: 1 first contact, 2 backface cull, 4 closest hit
 * (dGeomRaySetParams / dGeomRaySetClosestHit; they only matter against trimeshes).  For every ray the result is the
 * contact of smallest depth among dCollide(ray, g, 1, ...) over the enabled geoms g of the world's space whose AABB
 * overlaps the ray's and whose category / collide bits pass collideAABBs (collision_space_internal.h:48-82); ties
 * go to the geom that comes first in the space's list.  geom = creation index of that geom inside its space, -1 = no
 * hit (depth = length then).  Bodies are taken where the device has them (after the steps run so far). */
typedef struct dBatchRayHit { dReal pos[3]; dReal depth; dReal normal[3]; int geom; } dBatchRayHit;
int dBatchRayCast(dBatchID, int rays_per_world, const dReal *origin3, const dReal *dir3, const dReal *length, int ray_flags,
                  unsigned long category_bits, unsigned long collide_bits, dBatchRayHit *hits);
/* the CUDA stream the batch launches on (cudaStream_t as void*), so callers can
 * time with events on the launching stream */
void *dBatchGetStream(dBatchID);
/* number of kernel launches issued by the library since load (bench.py) */
long long dB200KernelLaunchCount(void);
/* Diagnostics for the libm restatements of ob_math.h (atan2f, sinf, cosf as glibc 2.39 computes them; the reference calls
 * them through dAtan2 / dSin / dCos, include/ode/common.h:150-160).  fn: 0 atan2f(a, b), 1 sinf(a), 2 cosf(a).
 * dB200LibmHost evaluates the restatement compiled for the host, dB200LibmDevice the same source on the GPU (needs a
 * device; returns -1 without one).  tests/test_abi.py compares both with the platform's libm bit for bit. */
int dB200LibmHost(int fn, int n, const float *a, const float *b, float *out);
int dB200LibmDevice(int fn, int n, const float *a, const float *b, float *out);
const char *dB200LastError(void);

#ifdef __cplusplus
}
#endif
#endif

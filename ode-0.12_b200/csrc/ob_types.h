// ob_types.h — device-resident data layout of a batch of independent worlds.
//
// Layout in HBM: every per-world array is one contiguous, 16-byte aligned block
// ("world slot"), slots are laid out back to back with a fixed stride derived
// from the batch capacities, so one CTA pulls its world with coalesced 128-bit
// loads and neighbouring CTAs touch neighbouring memory.  Index conventions
// carry the reference's ORDER semantics:
//   * body index  = position in dxWorld::firstbody list (0 = newest body),
//     so island discovery walks bodies 0..nb-1 (ode/src/util.cpp:420).
//   * geom index  = creation index inside the bound space; the space's linked
//     list order is the separate permutation `glist` (head first), rewritten
//     every step the way dGeomMoved does (ode/src/collision_space.cpp:47-75).
//   * contact joints are numbered in creation order per step; a body's joint
//     list is "newest first" (ode/src/ode.cpp:1376-1386).
#pragma once
#include "ob_math.h"

enum {  // body flags, ode/src/objects.h:38-48
  OB_BODY_FINITE_ROT = 1, OB_BODY_FINITE_ROT_AXIS = 2, OB_BODY_DISABLED = 4, OB_BODY_NO_GRAVITY = 8,
  OB_BODY_AUTO_DISABLE = 16, OB_BODY_LIN_DAMP = 32, OB_BODY_ANG_DAMP = 64, OB_BODY_MAX_ANG_SPEED = 128,
  OB_BODY_GYROSCOPIC = 256
};
enum { OB_GEOM_SPHERE = 0, OB_GEOM_BOX = 1, OB_GEOM_CAPSULE = 2, OB_GEOM_CYLINDER = 3, OB_GEOM_PLANE = 4, OB_GEOM_RAY = 5, OB_GEOM_TRIMESH = 8,
       OB_GEOM_SPACE = 10 };   // a sub-space as a member of the bound space (drop-in path): an axis-aligned box given directly in R[0..5]; no collider
enum { OB_GEOM_ENABLED = 1, OB_GEOM_HAS_OFFSET = 2, OB_GEOM_ZERO_SIZED = 4 };
// ObGeom::mesh / ObPose::mesh of a primitive (non-trimesh, non-ray) geom: set when the record stands for a geom
// transform (dCreateGeomTransform) served as its encapsulated geom; only the collider dispatch order depends on it
#define OB_POSE_XFORM 0x40000000
enum { OB_SPACE_HASH = 0, OB_SPACE_SAP = 1, OB_SPACE_SIMPLE = 2 };
enum { OB_ERR_CONTACT_OVERFLOW = 1, OB_ERR_ROW_OVERFLOW = 2, OB_ERR_PAIR_OVERFLOW = 4, OB_ERR_BVH_STACK = 8 };

// mutable body state (13 reals of ODE state + cached R + accumulators)
struct __attribute__((aligned(16))) ObBodyDyn {
  real pos[3]; uint32_t flags;
  real q[4];
  real lvel[3]; int adis_stepsleft;
  real avel[3]; real adis_timeleft;
  real facc[4];
  real tacc[4];
  real R[12];
};
// per-body constants
struct __attribute__((aligned(16))) ObBodyConst {
  real mass, invMass, max_angular_speed, pad0;
  real I[12];     // body-frame inertia (mass.I)
  real invI[12];  // body-frame inverse inertia
  real finite_rot_axis[4];
  real damp_lin_scale, damp_ang_scale, damp_lin_thr, damp_ang_thr;
  real adis_lin_thr, adis_ang_thr, adis_idle_time; int adis_idle_steps;
  int adis_samples; int geom_first; int npermjoints; int pad1;
};
struct __attribute__((aligned(16))) ObGeom {
  int type; int body; uint32_t cat, col;        // body = -1: static
  int flags; int body_next; int mesh; int pad;  // body_next: next geom of the same body (dGeomGetBodyNext); mesh: trimesh data index
  real p[4];                                     // sphere r | box lx,ly,lz | plane a,b,c,d | capsule r,l
  real pos[4];                                   // static pose or offset pose
  real R[12];
};
struct __attribute__((aligned(16))) ObWorld {
  real gravity[4];
  real erp, cfm, sor_w, max_vel;
  real min_depth; int iters; int nb; int ng;
  uint32_t seed; int hash_minlevel, hash_maxlevel, space_type;
  int npermjoints; int status;
  int sap_ndirty;   // SAP space: glist = DirtyList (first sap_ndirty entries) followed by GeomList
  int sap_axes;     // SAP space: axis order code (dSAP_AXES_*, collision_sapspace.cpp:273-275)
};
// surface parameters of the contact policy, copied into every contact joint
struct __attribute__((aligned(16))) ObSurface {
  int mode; real mu, mu2, bounce;
  real bounce_vel, soft_erp, soft_cfm, motion1;
  real motion2, motionN, slip1, slip2;
};
struct __attribute__((aligned(16))) ObPolicy {
  uint32_t cat_mask1, cat_mask2; int max_contacts; int skip_if_connected;
  int skip_static_pairs; int nrows /* row 0 only: rows in use (0 = 1) */; int pad[2];
  ObSurface surface;
};
#define OB_NCLS 6         // collider classes of the decoupled narrowphase: 0 = no call, 1 sphere-mesh, 2 box-mesh, 3 other-mesh, 4 box-box, 5 other primitives
#define OB_MAXPOLICY 8   // rows of the contact-policy table (dBatchContactPolicy, include/ode_b200/ode.h)
// permanent joints (ball / hinge / hinge2), body-frame parameters as the host API maintains them
struct __attribute__((aligned(16))) ObLimot {
  real vel, fmax, lostop, histop;
  real fudge_factor, normal_cfm, stop_erp, stop_cfm;
  real bounce; int limit; real limit_err; int pad;
};
enum { OB_JOINT_BALL = 1, OB_JOINT_HINGE = 2, OB_JOINT_SLIDER = 3, OB_JOINT_CONTACT = 4, OB_JOINT_UNIVERSAL = 5, OB_JOINT_HINGE2 = 6, OB_JOINT_FIXED = 7, OB_JOINT_AMOTOR = 9, OB_JOINT_LMOTOR = 10, OB_JOINT_PLANE2D = 11, OB_JOINT_PR = 12, OB_JOINT_PU = 13, OB_JOINT_PISTON = 14 };   // == dJointType
enum { OB_JF_DISABLED = 1, OB_JF_REVERSE = 2 };
// motor joints (amotor / lmotor) pack their small integers into ObJoint::flags: num<<8 | mode<<12 | rel0<<16 | rel1<<20 | rel2<<24;
// axes: axis1, axis2, anchor1; amotor: reference1 = anchor2, reference2 = v1, user-mode angles = qrel[0..2]
#define OB_JM_NUM(f) (((f) >> 8) & 15)
#define OB_JM_MODE(f) (((f) >> 12) & 15)
#define OB_JM_REL(f, i) (((f) >> (16 + 4 * (i))) & 15)
#define OB_NSIDE 4   // motor side-effect slots per joint (PU: two rotational motors + force and decoupling torque of the prismatic one)
struct __attribute__((aligned(16))) ObJoint {
  int type; int b1, b2; int flags;        // b1/b2 = node[0]/node[1] body index, -1 = none
  real anchor1[4], anchor2[4], axis1[4], axis2[4], qrel[4];   // slider / fixed / PR: anchor1 = offset; universal: qrel = qrel1, v1 = qrel2; PR: axis1/axis2 = axisR1/axisR2, v1 = axisP1; PU: universal fields + v2 = axisP1, limot3 = prismatic; plane2d: limot1..3 = x, y, angle motors
  real erp, cfm, susp_erp, susp_cfm;
  real c0, s0, pad0, pad1;
  real v1[4], v2[4];
  ObLimot limot1, limot2, limot3;
};
// one generated contact (dContactGeom-equivalent, 48 B single / 80 B double)
struct __attribute__((aligned(16))) ObContact {
  real pos[3]; real depth;
  real normal[3]; int g1;
  int g2; int side1, side2; int policy;
};

struct ObCounters {
  unsigned long long steps, body_steps, pairs, contacts, rows, islands, overflow_worlds;
};

#define OB_MAXEPOCH 8   // shuffle epochs per step ((iters+7)/8)
// per-world hand-off between the step kernels (ObBatchDev::stepinfo), ints
enum { SI_NIS = 0, SI_NIB, SI_NIJ, SI_MTOT, SI_HAVEROWS, SI_ANYBALL, SI_NPASS0, SI_WORDS = SI_NPASS0 + OB_MAXEPOCH };

// capacities + device pointers, passed by value to every kernel
struct ObBatchDev {
  int W;        // worlds
  int NB;       // bodies per world slot
  int NG;       // geoms per world slot
  int NP;       // broadphase pairs per world-step
  int NC;       // contact joints per world-step (also contact slots)
  int NR;       // constraint rows per world-step
  int npolicy;
  int NEP;      // shuffle epochs per step = ceil(max iters / 8)
  int NJ;       // permanent (non-contact) joints per world slot
  int dropin;   // 1: batch serves the classic per-call API (per-contact surfaces in csurf/cfdir1)
  int large;    // 1: one large world on the grid-wide path (ob_large.h); W == 1
  int NADIS;    // auto-disable sample buffer depth per body (max average_samples of the bound bodies; 0: nobody averages)
  int wbeg, wend;   // world range [wbeg, wend) of THIS launch (a chunk of the batch; chunks run on their own streams)
  ObWorld *world;        // [W]
  ObBodyDyn *bdyn;       // [W*NB]
  ObBodyConst *bconst;   // [W*NB]
  ObGeom *geom;          // [W*NG]
  int *glist;            // [W*NG] space-list order (head first), geom indices
  int *sapstate;         // [W*(NG+3)] SAP radix-sort context carried across steps: valid, nb, ranks[NG+1]
  ObPolicy *policy;      // [npolicy]
  struct ObMeshDev *meshes;   // [nmesh] trimesh data table (ob_trimesh.h); geoms refer to it by index
  int nmesh;
  int any_xf;   // some geom of the batch is a geom transform (OB_POSE_XFORM): set by every upload, uniform per launch
  ObJoint *joint;        // [W*NJ] permanent joints (ball / hinge / hinge2), creation order
  int *njoints;          // [W]
  unsigned short *padjstart; // [W*(NB+1)] per body: range into padj
  unsigned short *padj;      // [W*2*NJ] permanent joint ids in the body's joint-list order (newest attach first)
  // per-step products (device scratch, also the parity taps)
  int *npairs;           // [W]
  int *pairs;            // [W*NP*2] (o1,o2) geom indices in callback order
  int *ncontacts;        // [W]
  ObContact *contacts;   // [W*NC] contact joints in creation order
  real *rows;            // [W*NR*20] compact rows of the CUDA path (ob_step_kernel.cuh)
  real *invIw;           // [W*NB*12] world-frame inverse inertia per body (step scratch)
  real *tmp1;            // [W*NB*8]  v/h + invM*f_ext per body (step scratch)
  int *stepinfo;         // [W*16] hand-off between k_prep / k_sor / k_post
  unsigned char *ibody;  // [W*NB] island body order (stepping order)
  unsigned short *isz;   // [W*NB*4] per island: body start, body count, joint start, joint count
  unsigned short *jrow;  // [W*(NC+NJ+1)] first row of joint k (island joint order)
  unsigned short *ijoint;// [W*(NC+NJ)] island joint order -> joint id (contacts < nc <= permanent)
  real *jside;           // [W*NJ*4*OB_NSIDE] motor-at-limit side effects {fm, vector} per permanent joint (step scratch)
  unsigned short *sched; // [W*NEP*NR] rows in level order, per shuffle epoch
  unsigned short *pstart;// [W*NEP*(NR+1)] first slot of every pass, per shuffle epoch
  real *rowJ;            // [W*NR*12] (host test backend only)
  real *rowiMJ;          // [W*NR*12]
  real *rowJc;           // [W*NR*12] unscaled J copy (only written when the feedback tap is on)
  real *rowS;            // [W*NR*4]  b(rhs), Ad*cfm, lo, hi
  int *rowI;             // [W*NR*4]  findex, b1, b2, joint
  real *lambda;          // [W*NR]
  int *nrows;            // [W]
  real *fback;           // [W*(NC+NJ)*12] f1,t1,f2,t2 per joint (contacts, then permanent at NC+k) as dJointFeedback reports them
  ObSurface *csurf;      // [W*NC] per-contact surface parameters (drop-in path only, else null: policy table)
  real *cfdir1;          // [W*NC*4] per-contact fdir1 (drop-in path only)
  ObCounters *counters;  // [1]
  real *adisbuf;         // [W*NB*NADIS*6] auto-disable velocity samples (lvel, avel) per body, ring buffer (util.cpp:128-147)
  int *adisctl;          // [W*NB*2] per body: write index, buffer-full flag
  unsigned *rowmeta;     // [W*NR] b1 | b2<<8 | findex offset<<16 per row, written by the first half of k_prep for k_sched* (null: read the row records)
  // decoupled narrowphase of the CUDA path (k_broad -> k_narrow -> k_contacts, ob_kern_collide.cu); null on the other backends
  struct ObPose *gpose;  // [W*NG] this step's geom poses by geom index
  unsigned *wl;          // [OB_NCLS][W*NP][2] work items of the narrowphase per collider class: world, pair | policy row << 24
  unsigned *wlcnt;       // [16] items per class ([1..OB_NCLS-1]), [8] = next chunk of the queue
  unsigned char *pn;     // [W*NP] contacts of a pair
  int *poff;             // [W*NP] where they are in the world's pool region
  ObContact *pool;       // [W*NC] contacts in order of completion
  int *pcount;           // [W] contacts in the pool
};

// ob_host.h — host-side object model behind the ODE C API handles.
//
// This is the mirror of the reference's object model for the hot path only
// (ode/src/objects.h:38-159, ode/src/collision_kernel.h:96-240,
// ode/src/joints/joint.h:58-190): worlds, bodies, joints, geoms and spaces with
// the same linked-list disciplines (push-front everywhere, dirty geoms move to
// the head), because those list orders ARE the constraint ordering contract
// (SURVEY.md Appendix A).  No physics is computed here: dSpaceCollide, dCollide
// and dWorldQuickStep marshal into the device layout (ob_types.h) and launch
// the CUDA kernels; scene construction and getters/setters are plain host code.
#pragma once
#include <stddef.h>
#include <vector>
#include "../../include/ode_b200/ode.h"
#include "ob_types.h"

struct dxJoint;
struct dxJointNode {
  dxJoint *joint;   // joint this node belongs to
  dxBody *body;     // *other* body this node connects to (ode/src/joints/joint.h:45-49)
  dxJointNode *next;
};

struct dxAutoDisable {
  dReal idle_time; int idle_steps; dReal linear_average_threshold, angular_average_threshold;
  unsigned average_samples;
};
struct dxDamping { dReal linear_scale, angular_scale, linear_threshold, angular_threshold; };

struct dxWorld {
  dxBody *firstbody;
  dxJoint *firstjoint;
  int nb, nj;
  dVector3 gravity;
  dReal global_erp, global_cfm;
  dxAutoDisable adis;
  int body_flags;
  int qs_iterations; dReal qs_w;
  dReal contact_max_vel, contact_min_depth;
  dxDamping dampingp;
  dReal max_angular_speed;
  struct dxBatch *bound_batch;   // non-null while a batch owns the device copy
};

struct dxBody {
  dxWorld *world;
  dxBody *next; dxBody **tome;
  int tag; void *userdata;
  dxJointNode *firstjoint;
  unsigned flags;
  dxGeom *geom;
  dMass mass;
  dMatrix3 invI;
  dReal invMass;
  dVector3 pos; dMatrix3 R;     // posr
  dQuaternion q;
  dVector3 lvel, avel, facc, tacc, finite_rot_axis;
  dxAutoDisable adis;
  dReal adis_timeleft; int adis_stepsleft;
  unsigned average_counter; int average_ready;
  std::vector<dReal> average_buf;   // [average_samples][6] lvel / avel samples (dxBody::average_lvel_buffer / average_avel_buffer, objects.h)
  dxDamping dampingp;
  dReal max_angular_speed;
  int batch_index;               // index inside the bound batch world slot
  void (*moved_callback)(dxBody *);   // dBodySetMovedCallback (ode.cpp:1119): drop-in path, after every step that moved the body
};

enum { dJOINT_INGROUP = 1, dJOINT_REVERSE = 2, dJOINT_TWOBODIES = 4, dJOINT_DISABLED = 8 };

struct dxLimot {   // dxJointLimitMotor, ode/src/joints/joint.h:196-213
  dReal vel, fmax, lostop, histop, fudge_factor, normal_cfm, stop_erp, stop_cfm, bounce;
  int limit; dReal limit_err;
};

struct dxJoint {
  dxWorld *world;
  dxJoint *next; dxJoint **tome;
  int tag; void *userdata;
  int type;
  unsigned flags;
  dxJointNode node[2];
  dJointFeedback *feedback;
  dxJointGroup *group;
  // contact
  dContact contact;
  // ball / hinge / hinge2 (body-frame anchors and axes)
  dVector3 anchor1, anchor2, axis1, axis2;
  dQuaternion qrel;
  dReal erp, cfm;               // ball
  dxLimot limot, limot2;        // hinge: limot; hinge2: limot (axis 1) + limot2 (axis 2)
  dReal c0, s0, v1[4], v2[4];   // hinge2
  dReal susp_erp, susp_cfm;     // hinge2
  dVector3 offset;              // slider / fixed: centre of body 1 w.r.t. body 2 (slider.cpp computeOffset, fixed.cpp dJointSetFixed)
  dQuaternion qrel2;            // universal: second initial relative rotation (qrel = qrel1)
  // amotor / lmotor: axes are axis1, axis2, axis3
  int num, mode, rel[3];
  dVector3 axis3, reference1, reference2;
  dxLimot limot3;
  dReal angle[3];
};

struct dxJointGroup {
  std::vector<dxJoint *> joints;   // creation order
};

enum {  // geom flags, ode/src/collision_kernel.h:64-80
  GEOM_DIRTY = 1, GEOM_POSR_BAD = 2, GEOM_AABB_BAD = 4, GEOM_PLACEABLE = 8, GEOM_ENABLED = 16, GEOM_ZERO_SIZED = 32,
  RAY_FIRSTCONTACT = 0x10000, RAY_BACKFACECULL = 0x20000, RAY_CLOSEST_HIT = 0x40000   // collision_kernel.h:79-81
};

struct dxPosR { dVector3 pos; dMatrix3 R; };
struct dxGeom;
int ob_ray_flags(const dxGeom *g);   // RAY_* gflags -> OB_RAY_* bits (ob_trimesh.h)

struct dxGeom {
  int type;
  int gflags;
  void *data;
  dxBody *body;
  dxGeom *body_next;
  dxPosR *final_posr;     // body's pos/R when attached without offset
  dxPosR *offset_posr;
  dxPosR own_posr;        // storage when not borrowed from the body
  dxPosR off_storage;
  dxGeom *next; dxGeom **tome;
  dxSpace *parent_space;
  dReal aabb[6];
  unsigned long category_bits, collide_bits;
  dReal p[4];             // sphere r | box sides | plane a,b,c,d | capsule r,l
  struct dxTriMeshData *tmdata;   // trimesh geoms: the shared mesh data
  dxGeom *xf_obj; int xf_cleanup, xf_info;   // geom transform (collision_transform.cpp:44-47): encapsulated geom, cleanup mode, info mode
  int batch_index;
  int sap_didx, sap_gidx;   // position in the parent SAP space's DirtyList / GeomList (-1: not in that list)
  bool is_space;
  virtual ~dxGeom() {}
};

struct dxSpace : public dxGeom {
  int count;
  dxGeom *first;
  int cleanup, sublevel, lock_count, manual_cleanup;
  int minlevel, maxlevel;   // hash space
  int axisorder;            // SAP
  // SAP space only (collision_sapspace.cpp:140-160): the two arrays whose order defines the sweep's
  // tie-breaking; `first/next` stays a plain membership list for these spaces
  std::vector<dxGeom *> sap_dirty, sap_geoms;
  struct dxBatch *bound_batch;
};

// ---- internals shared between ob_host.cpp, ob_batch.cpp, ob_dropin.cpp ---------
void ob_error(int num, const char *fmt, ...);      // dError: message, then exit(1) unless handled
void ob_debug(int num, const char *fmt, ...);      // dDebug: message, then abort() unless handled
void ob_message(int num, const char *fmt, ...);
void ob_set_last_error(const char *fmt, ...);
#define OB_UASSERT(c, msg) do { if (!(c)) ob_debug(2 /*d_ERR_UASSERT*/, msg " in %s()", __FUNCTION__); } while (0)
#define OB_AASSERT(c) OB_UASSERT(c, "Bad argument(s)")
void ob_geom_moved(dxGeom *g);                      // dGeomMoved
dxGeom *ob_geom_create(dxSpace *space, int is_placeable, int type);
void ob_geom_recompute_posr(dxGeom *g);
// geom transforms: the geom whose shape stands for g on the device (g itself unless g is a transform), and the pose
// it collides at (computeFinalTx, collision_transform.cpp:101-108).  ob_geom_shape returns 0 for an empty transform.
dxGeom *ob_geom_shape(dxGeom *g);
void ob_geom_final_pose(dxGeom *g, dxPosR *out);
void ob_space_clean(dxSpace *s);                    // cleanGeoms
void ob_body_posr(dxBody *b, dxPosR *out);
extern uint32_t ob_global_seed;
void ob_joint_init_type(dxJoint *j);             // ob_joints.cpp
void ob_joint_set_relative_values(dxJoint *j);   // ob_joints.cpp

// drop-in compute entry points implemented over the CUDA backend (ob_dropin.cpp)
void ob_dropin_space_collide(dxSpace *space, void *data, dNearCallback *cb);
void ob_dropin_space_collide2(dxGeom *g1, dxGeom *g2, void *data, dNearCallback *cb);
int ob_dropin_collide(dxGeom *o1, dxGeom *o2, int flags, dContactGeom *contact, int skip);
int ob_dropin_quickstep(dxWorld *w, dReal h);
void ob_dropin_forget_world(dxWorld *w);
void ob_dropin_forget_space(dxSpace *s);

// ob_joints.cpp — host-side bookkeeping of ball / hinge / hinge2 joints: body-frame
// anchors and axes, limit/motor parameters, angle getters.  Row assembly for these
// joints happens on the device (ob_rows.h); this file only maintains the parameters
// the rows are built from, with the reference's exact arithmetic so that the
// uploaded values are bit-identical.
// Reference: ode/src/joints/joint.cpp:263-450 (setAnchors/setAxes/getAnchor/getAxis,
// getHingeAngle, dxJointLimitMotor::init/set/get), ball.cpp, hinge.cpp, hinge2.cpp.
#include <string.h>
#include "ob_host.h"

static void limot_init(dxLimot &l, dxWorld *w) {
  l.vel = 0; l.fmax = 0; l.lostop = -OB_INF; l.histop = OB_INF; l.fudge_factor = 1;
  l.normal_cfm = w->global_cfm; l.stop_erp = w->global_erp; l.stop_cfm = w->global_cfm;
  l.bounce = 0; l.limit = 0; l.limit_err = 0;
}
static void limot_set(dxLimot &l, int num, dReal value) {
  switch (num) {
    case dParamLoStop: l.lostop = value; break;
    case dParamHiStop: l.histop = value; break;
    case dParamVel: l.vel = value; break;
    case dParamFMax: if (value >= 0) l.fmax = value; break;
    case dParamFudgeFactor: if (value >= 0 && value <= 1) l.fudge_factor = value; break;
    case dParamBounce: l.bounce = value; break;
    case dParamCFM: l.normal_cfm = value; break;
    case dParamStopERP: l.stop_erp = value; break;
    case dParamStopCFM: l.stop_cfm = value; break;
  }
}
static dReal limot_get(const dxLimot &l, int num) {
  switch (num) {
    case dParamLoStop: return l.lostop;
    case dParamHiStop: return l.histop;
    case dParamVel: return l.vel;
    case dParamFMax: return l.fmax;
    case dParamFudgeFactor: return l.fudge_factor;
    case dParamBounce: return l.bounce;
    case dParamCFM: return l.normal_cfm;
    case dParamStopERP: return l.stop_erp;
    case dParamStopCFM: return l.stop_cfm;
    default: return 0;
  }
}

void ob_joint_init_type(dxJoint *j) {
  dxWorld *w = j->world;
  switch (j->type) {
    case dJointTypeBall:
      j->erp = w->global_erp; j->cfm = w->global_cfm;
      break;
    case dJointTypeHinge:
      j->axis1[0] = 1; j->axis2[0] = 1;
      limot_init(j->limot, w);
      break;
    case dJointTypeHinge2:
      j->axis1[0] = 1; j->axis2[1] = 1;
      j->c0 = 0; j->s0 = 0; j->v1[0] = 1; j->v2[1] = 1;
      limot_init(j->limot, w); limot_init(j->limot2, w);
      j->susp_erp = w->global_erp; j->susp_cfm = w->global_cfm;
      j->flags |= dJOINT_TWOBODIES;
      break;
    case dJointTypeSlider:
      j->axis1[0] = 1;
      limot_init(j->limot, w);
      break;
    case dJointTypeFixed:
      j->erp = w->global_erp; j->cfm = w->global_cfm;
      break;
    case dJointTypeUniversal:
      j->axis1[0] = 1; j->axis2[1] = 1;
      limot_init(j->limot, w); limot_init(j->limot2, w);
      break;
    case dJointTypePlane2D:   // limot = x motor, limot2 = y motor, limot3 = angle motor (plane2d.cpp:54-60)
      limot_init(j->limot, w); limot_init(j->limot2, w); limot_init(j->limot3, w);
      break;
    case dJointTypePiston:    // limot = prismatic, limot2 = rotoide (piston.cpp:37-50)
      j->axis1[0] = 1; j->axis2[0] = 1;
      limot_init(j->limot, w); limot_init(j->limot2, w);
      break;
    case dJointTypePR:        // axis1/axis2 = axisR1/axisR2, axis3 = axisP1, offset = anchor w.r.t. body 1 (pr.cpp:36-64)
      j->axis1[0] = 1; j->axis2[0] = 1; j->axis3[1] = 1;
      limot_init(j->limot, w); limot_init(j->limot2, w);
      break;
    case dJointTypePU:        // universal fields + axis3 = axisP1; limot / limot2 = universal axes, limot3 = prismatic (pu.cpp:31-72)
      j->axis1[1] = 1; j->axis2[2] = 1; j->axis3[0] = 1;
      limot_init(j->limot, w); limot_init(j->limot2, w); limot_init(j->limot3, w);
      break;
    case dJointTypeAMotor:
    case dJointTypeLMotor:
      j->num = 0; j->mode = dAMotorUser;
      limot_init(j->limot, w); limot_init(j->limot2, w); limot_init(j->limot3, w);
      break;
    default: break;
  }
}

// joint.cpp:263-296
static void set_anchors(dxJoint *j, dReal x, dReal y, dReal z, dReal *anchor1, dReal *anchor2) {
  if (j->node[0].body) {
    dReal q[4];
    dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
    q[0] = x - b0->pos[0]; q[1] = y - b0->pos[1]; q[2] = z - b0->pos[2]; q[3] = 0;
    ob_mul1_331(anchor1, b0->R, q);
    if (b1) {
      q[0] = x - b1->pos[0]; q[1] = y - b1->pos[1]; q[2] = z - b1->pos[2]; q[3] = 0;
      ob_mul1_331(anchor2, b1->R, q);
    } else { anchor2[0] = x; anchor2[1] = y; anchor2[2] = z; }
  }
  anchor1[3] = 0; anchor2[3] = 0;
}
// joint.cpp:299-331
static void set_axes(dxJoint *j, dReal x, dReal y, dReal z, dReal *axis1, dReal *axis2) {
  if (j->node[0].body) {
    dReal q[4] = {x, y, z, 0};
    ob_safe_normalize3(q);
    if (axis1) { ob_mul1_331(axis1, j->node[0].body->R, q); axis1[3] = 0; }
    if (axis2) {
      if (j->node[1].body) ob_mul1_331(axis2, j->node[1].body->R, q);
      else { axis2[0] = x; axis2[1] = y; axis2[2] = z; }
      axis2[3] = 0;
    }
  }
}
static void get_anchor(dxJoint *j, dReal *result, const dReal *anchor1) {
  if (j->node[0].body) {
    dxBody *b = j->node[0].body;
    ob_mul0_331(result, b->R, anchor1);
    result[0] += b->pos[0]; result[1] += b->pos[1]; result[2] += b->pos[2];
  }
}
static void get_anchor2(dxJoint *j, dReal *result, const dReal *anchor2) {
  if (j->node[1].body) {
    dxBody *b = j->node[1].body;
    ob_mul0_331(result, b->R, anchor2);
    result[0] += b->pos[0]; result[1] += b->pos[1]; result[2] += b->pos[2];
  } else { result[0] = anchor2[0]; result[1] = anchor2[1]; result[2] = anchor2[2]; }
}
static void qmul1(dReal *qa, const dReal *qb, const dReal *qc) {   // dQMultiply1, rotation.cpp:201-208
  dReal a0 = qb[0] * qc[0] + qb[1] * qc[1] + qb[2] * qc[2] + qb[3] * qc[3];
  dReal a1 = qb[0] * qc[1] - qb[1] * qc[0] - qb[2] * qc[3] + qb[3] * qc[2];
  dReal a2 = qb[0] * qc[2] - qb[2] * qc[0] - qb[3] * qc[1] + qb[1] * qc[3];
  dReal a3 = qb[0] * qc[3] - qb[3] * qc[0] - qb[1] * qc[2] + qb[2] * qc[1];
  qa[0] = a0; qa[1] = a1; qa[2] = a2; qa[3] = a3;
}
static void hinge_initial_rel_rot(dxJoint *j) {   // hinge.cpp computeInitialRelativeRotation
  if (j->node[0].body) {
    if (j->node[1].body) qmul1(j->qrel, j->node[0].body->q, j->node[1].body->q);
    else {
      const dReal *q = j->node[0].body->q;
      j->qrel[0] = q[0]; j->qrel[1] = -q[1]; j->qrel[2] = -q[2]; j->qrel[3] = -q[3];
    }
  }
}
static void hinge2_axis_info(dxJoint *j, dReal *ax1, dReal *ax2, dReal *axCross, dReal *sin_angle, dReal *cos_angle) {
  ob_mul0_331(ax1, j->node[0].body->R, j->axis1);
  ob_mul0_331(ax2, j->node[1].body->R, j->axis2);
  ob_cross(axCross, ax1, ax2);
  *sin_angle = ob_sqrt(axCross[0] * axCross[0] + axCross[1] * axCross[1] + axCross[2] * axCross[2]);
  *cos_angle = ob_dot(ax1, ax2);
}
static void hinge2_make_v1v2(dxJoint *j) {   // hinge2.cpp makeV1andV2
  if (j->node[0].body) {
    dReal ax1[4], ax2[4], v[4];
    ob_mul0_331(ax1, j->node[0].body->R, j->axis1);
    ob_mul0_331(ax2, j->node[1].body->R, j->axis2);
    if ((ax1[0] == 0 && ax1[1] == 0 && ax1[2] == 0) || (ax2[0] == 0 && ax2[1] == 0 && ax2[2] == 0) ||
        (ax1[0] == ax2[0] && ax1[1] == ax2[1] && ax1[2] == ax2[2])) return;
    dReal k = ob_dot(ax1, ax2);
    for (int i = 0; i < 3; i++) ax2[i] -= k * ax1[i];
    ob_safe_normalize3(ax2);
    ob_cross(v, ax1, ax2);
    ob_mul1_331(j->v1, j->node[0].body->R, ax2);
    ob_mul1_331(j->v2, j->node[0].body->R, v);
  }
}

extern "C" {
// ---- ball ---------------------------------------------------------------------------
void dJointSetBallAnchor(dJointID j, dReal x, dReal y, dReal z) { set_anchors(j, x, y, z, j->anchor1, j->anchor2); }
void dJointSetBallAnchor2(dJointID j, dReal x, dReal y, dReal z) { j->anchor2[0] = x; j->anchor2[1] = y; j->anchor2[2] = z; j->anchor2[3] = 0; }
void dJointGetBallAnchor(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor2(j, result, j->anchor2); else get_anchor(j, result, j->anchor1);
}
void dJointGetBallAnchor2(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor(j, result, j->anchor1); else get_anchor2(j, result, j->anchor2);
}
void dJointSetBallParam(dJointID j, int parameter, dReal value) {
  if (parameter == dParamCFM) j->cfm = value; else if (parameter == dParamERP) j->erp = value;
}
// ---- hinge --------------------------------------------------------------------------
void dJointSetHingeAnchor(dJointID j, dReal x, dReal y, dReal z) { set_anchors(j, x, y, z, j->anchor1, j->anchor2); hinge_initial_rel_rot(j); }
void dJointSetHingeAxis(dJointID j, dReal x, dReal y, dReal z) { set_axes(j, x, y, z, j->axis1, j->axis2); hinge_initial_rel_rot(j); }
void dJointSetHingeParam(dJointID j, int parameter, dReal value) { limot_set(j->limot, parameter, value); }
dReal dJointGetHingeParam(dJointID j, int parameter) { return limot_get(j->limot, parameter); }
void dJointSetHingeAnchorDelta(dJointID j, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz) {   // hinge.cpp:163-199
  if (j->node[0].body) {
    dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
    dReal q[4] = {x - b0->pos[0], y - b0->pos[1], z - b0->pos[2], 0};
    ob_mul1_331(j->anchor1, b0->R, q);
    if (b1) {
      q[0] = x - b1->pos[0]; q[1] = y - b1->pos[1]; q[2] = z - b1->pos[2]; q[3] = 0;
      ob_mul1_331(j->anchor2, b1->R, q);
    } else { j->anchor2[0] = x + dx; j->anchor2[1] = y + dy; j->anchor2[2] = z + dz; }
  }
  j->anchor1[3] = 0; j->anchor2[3] = 0;
  hinge_initial_rel_rot(j);
}
void dJointSetHingeAxisOffset(dJointID j, dReal x, dReal y, dReal z, dReal dangle) {   // hinge.cpp:213-230
  set_axes(j, x, y, z, j->axis1, j->axis2);
  hinge_initial_rel_rot(j);
  if (j->flags & dJOINT_REVERSE) dangle = -dangle;
  dQuaternion qAngle, qOffset;
  dQFromAxisAndAngle(qAngle, x, y, z, dangle);
  ob_qmul3(qOffset, qAngle, j->qrel);
  j->qrel[0] = qOffset[0]; j->qrel[1] = qOffset[1]; j->qrel[2] = qOffset[2]; j->qrel[3] = qOffset[3];
}
void dJointGetHingeAnchor(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor2(j, result, j->anchor2); else get_anchor(j, result, j->anchor1);
}
void dJointGetHingeAnchor2(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor(j, result, j->anchor1); else get_anchor2(j, result, j->anchor2);
}
void dJointGetHingeAxis(dJointID j, dVector3 result) { if (j->node[0].body) ob_mul0_331(result, j->node[0].body->R, j->axis1); }
}  // extern "C"
#include "ob_rows.h"
extern "C" {
dReal dJointGetHingeAngle(dJointID j) {
  if (j->node[0].body) {
    dReal ang = ob_hinge_angle(j->node[0].body->q, j->node[1].body ? j->node[1].body->q : 0, j->axis1, j->qrel);
    return (j->flags & dJOINT_REVERSE) ? -ang : ang;
  }
  return 0;
}
dReal dJointGetHingeAngleRate(dJointID j) {
  if (j->node[0].body) {
    dReal axis[4];
    ob_mul0_331(axis, j->node[0].body->R, j->axis1);
    dReal rate = ob_dot(axis, j->node[0].body->avel);
    if (j->node[1].body) rate -= ob_dot(axis, j->node[1].body->avel);
    if (j->flags & dJOINT_REVERSE) rate = -rate;
    return rate;
  }
  return 0;
}
// ---- hinge2 -------------------------------------------------------------------------
void dJointSetHinge2Anchor(dJointID j, dReal x, dReal y, dReal z) { set_anchors(j, x, y, z, j->anchor1, j->anchor2); hinge2_make_v1v2(j); }
void dJointSetHinge2Axis1(dJointID j, dReal x, dReal y, dReal z) {
  if (j->node[0].body) {
    set_axes(j, x, y, z, j->axis1, 0);
    dReal ax1[4], ax2[4], ax[4];
    hinge2_axis_info(j, ax1, ax2, ax, &j->s0, &j->c0);
  }
  hinge2_make_v1v2(j);
}
void dJointSetHinge2Axis2(dJointID j, dReal x, dReal y, dReal z) {
  if (j->node[1].body) {
    set_axes(j, x, y, z, 0, j->axis2);
    dReal ax1[4], ax2[4], ax[4];
    hinge2_axis_info(j, ax1, ax2, ax, &j->s0, &j->c0);
  }
  hinge2_make_v1v2(j);
}
void dJointSetHinge2Param(dJointID j, int parameter, dReal value) {
  if ((parameter & 0xff00) == 0x100) limot_set(j->limot2, parameter & 0xff, value);
  else {
    if (parameter == dParamSuspensionERP) j->susp_erp = value;
    else if (parameter == dParamSuspensionCFM) j->susp_cfm = value;
    else limot_set(j->limot, parameter, value);
  }
}
dReal dJointGetHinge2Param(dJointID j, int parameter) {
  if ((parameter & 0xff00) == 0x100) return limot_get(j->limot2, parameter & 0xff);
  if (parameter == dParamSuspensionERP) return j->susp_erp;
  if (parameter == dParamSuspensionCFM) return j->susp_cfm;
  return limot_get(j->limot, parameter);
}
void dJointGetHinge2Anchor(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor2(j, result, j->anchor2); else get_anchor(j, result, j->anchor1);
}
void dJointGetHinge2Anchor2(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor(j, result, j->anchor1); else get_anchor2(j, result, j->anchor2);
}
void dJointGetHinge2Axis1(dJointID j, dVector3 result) { if (j->node[0].body) ob_mul0_331(result, j->node[0].body->R, j->axis1); }
void dJointGetHinge2Axis2(dJointID j, dVector3 result) { if (j->node[1].body) ob_mul0_331(result, j->node[1].body->R, j->axis2); }
dReal dJointGetHinge2Angle1(dJointID j) {
  if (j->node[0].body) return ob_hinge2_angle(j->node[0].body->R, j->node[1].body->R, j->axis2, j->v1, j->v2);
  return 0;
}
dReal dJointGetHinge2Angle1Rate(dJointID j) {
  if (j->node[0].body) {
    dReal axis[4];
    ob_mul0_331(axis, j->node[0].body->R, j->axis1);
    dReal rate = ob_dot(axis, j->node[0].body->avel);
    if (j->node[1].body) rate -= ob_dot(axis, j->node[1].body->avel);
    return rate;
  }
  return 0;
}
dReal dJointGetHinge2Angle2Rate(dJointID j) {
  if (j->node[0].body && j->node[1].body) {
    dReal axis[4];
    ob_mul0_331(axis, j->node[1].body->R, j->axis2);
    dReal rate = ob_dot(axis, j->node[0].body->avel);
    rate -= ob_dot(axis, j->node[1].body->avel);
    return rate;
  }
  return 0;
}
}  // extern "C"

// ---- slider (slider.cpp) and fixed (fixed.cpp) ----------------------------------------------------
static void slider_compute_offset(dxJoint *j) {   // dxJointSlider::computeOffset
  if (j->node[1].body) {
    dReal c[4];
    for (int i = 0; i < 3; i++) c[i] = j->node[0].body->pos[i] - j->node[1].body->pos[i];
    c[3] = 0;
    ob_mul1_331(j->offset, j->node[1].body->R, c);
  } else if (j->node[0].body) {
    for (int i = 0; i < 3; i++) j->offset[i] = j->node[0].body->pos[i];
  }
}
extern "C" {
void dJointSetSliderAxis(dJointID j, dReal x, dReal y, dReal z) {
  set_axes(j, x, y, z, j->axis1, 0);
  slider_compute_offset(j);
  hinge_initial_rel_rot(j);   // same computeInitialRelativeRotation as the hinge (slider.cpp)
}
void dJointSetSliderAxisDelta(dJointID j, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz) {
  set_axes(j, x, y, z, j->axis1, 0);
  slider_compute_offset(j);
  if (!j->node[1].body) { j->offset[0] += dx; j->offset[1] += dy; j->offset[2] += dz; }
  hinge_initial_rel_rot(j);
}
void dJointGetSliderAxis(dJointID j, dVector3 result) { if (j->node[0].body) ob_mul0_331(result, j->node[0].body->R, j->axis1); }
void dJointSetSliderParam(dJointID j, int parameter, dReal value) { limot_set(j->limot, parameter, value); }
dReal dJointGetSliderParam(dJointID j, int parameter) { return limot_get(j->limot, parameter); }
dReal dJointGetSliderPosition(dJointID j) {
  dReal ax1[4], q[4];
  ob_mul0_331(ax1, j->node[0].body->R, j->axis1);
  if (j->node[1].body) {
    ob_mul0_331(q, j->node[1].body->R, j->offset);
    for (int i = 0; i < 3; i++) q[i] = j->node[0].body->pos[i] - q[i] - j->node[1].body->pos[i];
  } else {
    for (int i = 0; i < 3; i++) q[i] = j->node[0].body->pos[i] - j->offset[i];
    if (j->flags & dJOINT_REVERSE) { ax1[0] = -ax1[0]; ax1[1] = -ax1[1]; ax1[2] = -ax1[2]; }
  }
  return ob_dot(ax1, q);
}
dReal dJointGetSliderPositionRate(dJointID j) {
  dReal ax1[4];
  ob_mul0_331(ax1, j->node[0].body->R, j->axis1);
  if (j->node[1].body) return ob_dot(ax1, j->node[0].body->lvel) - ob_dot(ax1, j->node[1].body->lvel);
  dReal rate = ob_dot(ax1, j->node[0].body->lvel);
  if (j->flags & dJOINT_REVERSE) rate = -rate;
  return rate;
}
void dJointAddSliderForce(dJointID j, dReal force) {
  dReal axis[4] = {0, 0, 0, 0};
  if (j->flags & dJOINT_REVERSE) force -= force;   // sic (slider.cpp:316-317)
  dJointGetSliderAxis(j, axis);
  axis[0] *= force; axis[1] *= force; axis[2] *= force;
  if (j->node[0].body) dBodyAddForce(j->node[0].body, axis[0], axis[1], axis[2]);
  if (j->node[1].body) dBodyAddForce(j->node[1].body, -axis[0], -axis[1], -axis[2]);
  if (j->node[0].body && j->node[1].body) {
    dReal ltd[4], c[4];
    for (int i = 0; i < 3; i++) c[i] = (dReal)0.5 * (j->node[1].body->pos[i] - j->node[0].body->pos[i]);
    ob_cross(ltd, c, axis);
    dBodyAddTorque(j->node[0].body, ltd[0], ltd[1], ltd[2]);
    dBodyAddTorque(j->node[1].body, ltd[0], ltd[1], ltd[2]);
  }
}
void dJointSetFixed(dJointID j) {
  if (j->node[0].body) {
    if (j->node[1].body) {
      dReal ofs[4];
      for (int i = 0; i < 3; i++) ofs[i] = j->node[0].body->pos[i] - j->node[1].body->pos[i];
      ofs[3] = 0;
      ob_mul1_331(j->offset, j->node[0].body->R, ofs);
    } else {
      for (int i = 0; i < 3; i++) j->offset[i] = j->node[0].body->pos[i];
    }
  }
  hinge_initial_rel_rot(j);   // dxJointFixed::computeInitialRelativeRotation is the same computation
}
void dJointSetFixedParam(dJointID j, int parameter, dReal value) {
  if (parameter == dParamCFM) j->cfm = value; else if (parameter == dParamERP) j->erp = value;
}
dReal dJointGetFixedParam(dJointID j, int parameter) {
  if (parameter == dParamCFM) return j->cfm;
  if (parameter == dParamERP) return j->erp;
  return 0;
}
}  // extern "C"

// ---- universal (universal.cpp) -------------------------------------------------------------------------
static void universal_axes(dxJoint *j, dReal *ax1, dReal *ax2) {
  ob_mul0_331(ax1, j->node[0].body->R, j->axis1);
  if (j->node[1].body) ob_mul0_331(ax2, j->node[1].body->R, j->axis2);
  else { ax2[0] = j->axis2[0]; ax2[1] = j->axis2[1]; ax2[2] = j->axis2[2]; }
}
static void universal_initial_rel_rots(dxJoint *j) {   // computeInitialRelativeRotations :364-391
  if (j->node[0].body) {
    dReal ax1[4], ax2[4], R[12], qcross[4];
    universal_axes(j, ax1, ax2);
    memset(R, 0, sizeof R);
    ob_Rfrom2axes(R, ax1[0], ax1[1], ax1[2], ax2[0], ax2[1], ax2[2]);
    ob_QfromR(qcross, R);
    qmul1(j->qrel, j->node[0].body->q, qcross);
    ob_Rfrom2axes(R, ax2[0], ax2[1], ax2[2], ax1[0], ax1[1], ax1[2]);
    ob_QfromR(qcross, R);
    if (j->node[1].body) qmul1(j->qrel2, j->node[1].body->q, qcross);
    else for (int i = 0; i < 4; i++) j->qrel2[i] = qcross[i];
  }
}
static void universal_fill(const dxJoint *j, ObJoint &o) {   // what the angle functions of ob_rows.h read
  memset(&o, 0, sizeof o);
  for (int k = 0; k < 4; k++) { o.axis1[k] = j->axis1[k]; o.axis2[k] = j->axis2[k]; o.qrel[k] = j->qrel[k]; o.v1[k] = j->qrel2[k]; }
}
extern "C" {
void dJointSetUniversalAnchor(dJointID j, dReal x, dReal y, dReal z) { set_anchors(j, x, y, z, j->anchor1, j->anchor2); universal_initial_rel_rots(j); }
void dJointSetUniversalAxis1(dJointID j, dReal x, dReal y, dReal z) {
  if (j->flags & dJOINT_REVERSE) set_axes(j, x, y, z, 0, j->axis2); else set_axes(j, x, y, z, j->axis1, 0);
  universal_initial_rel_rots(j);
}
void dJointSetUniversalAxis2(dJointID j, dReal x, dReal y, dReal z) {
  if (j->flags & dJOINT_REVERSE) set_axes(j, x, y, z, j->axis1, 0); else set_axes(j, x, y, z, 0, j->axis2);
  universal_initial_rel_rots(j);
}
// axis setters that also declare the current pose to be at angles (offset1, offset2) (universal.cpp:418-478, :493-548)
static void universal_offset_rel_rots(dxJoint *j, const dReal *a1, const dReal *a2, dReal offset1, dReal offset2) {
  dQuaternion qAngle, qcross, qOffset;
  dMatrix3 R;
  memset(R, 0, sizeof R);
  dQFromAxisAndAngle(qAngle, a1[0], a1[1], a1[2], offset1);
  ob_Rfrom2axes(R, a1[0], a1[1], a1[2], a2[0], a2[1], a2[2]);
  ob_QfromR(qcross, R);
  dQMultiply0(qOffset, qAngle, qcross);
  qmul1(j->qrel, j->node[0].body->q, qOffset);
  dQFromAxisAndAngle(qAngle, a2[0], a2[1], a2[2], offset2);
  ob_Rfrom2axes(R, a2[0], a2[1], a2[2], a1[0], a1[1], a1[2]);
  ob_QfromR(qcross, R);
  qmul1(qOffset, qAngle, qcross);
  if (j->node[1].body) qmul1(j->qrel2, j->node[1].body->q, qOffset);
  else for (int i = 0; i < 4; i++) j->qrel2[i] = qcross[i];
}
void dJointSetUniversalAxis1Offset(dJointID j, dReal x, dReal y, dReal z, dReal offset1, dReal offset2) {
  if (j->flags & dJOINT_REVERSE) { set_axes(j, x, y, z, 0, j->axis2); offset1 = -offset1; offset2 = -offset2; }
  else set_axes(j, x, y, z, j->axis1, 0);
  universal_initial_rel_rots(j);
  dReal ax1[4], ax2[4];
  universal_axes(j, ax1, ax2);
  const dReal in[3] = {x, y, z};           // the caller's axis, not the stored normalised one
  universal_offset_rel_rots(j, in, ax2, offset1, offset2);
}
void dJointSetUniversalAxis2Offset(dJointID j, dReal x, dReal y, dReal z, dReal offset1, dReal offset2) {
  if (j->flags & dJOINT_REVERSE) { set_axes(j, x, y, z, j->axis1, 0); offset1 = -offset2; offset2 = -offset1; }   // sic (:503-504)
  else set_axes(j, x, y, z, 0, j->axis2);
  universal_initial_rel_rots(j);
  dReal ax1[4], ax2[4];
  universal_axes(j, ax1, ax2);
  universal_offset_rel_rots(j, ax1, ax2, offset1, offset2);
}
void dJointGetUniversalAnchor(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor2(j, result, j->anchor2); else get_anchor(j, result, j->anchor1);
}
void dJointGetUniversalAnchor2(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor(j, result, j->anchor1); else get_anchor2(j, result, j->anchor2);
}
void dJointGetUniversalAxis1(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) { if (j->node[1].body) ob_mul0_331(result, j->node[1].body->R, j->axis2); else { result[0] = j->axis2[0]; result[1] = j->axis2[1]; result[2] = j->axis2[2]; } }
  else if (j->node[0].body) ob_mul0_331(result, j->node[0].body->R, j->axis1);
}
void dJointGetUniversalAxis2(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) { if (j->node[0].body) ob_mul0_331(result, j->node[0].body->R, j->axis1); }
  else if (j->node[1].body) ob_mul0_331(result, j->node[1].body->R, j->axis2);
  else { result[0] = j->axis2[0]; result[1] = j->axis2[1]; result[2] = j->axis2[2]; }
}
void dJointSetUniversalParam(dJointID j, int parameter, dReal value) {
  if ((parameter & 0xff00) == 0x100) limot_set(j->limot2, parameter & 0xff, value); else limot_set(j->limot, parameter, value);
}
dReal dJointGetUniversalParam(dJointID j, int parameter) {
  if ((parameter & 0xff00) == 0x100) return limot_get(j->limot2, parameter & 0xff);
  return limot_get(j->limot, parameter);
}
void dJointGetUniversalAngles(dJointID j, dReal *angle1, dReal *angle2) {
  *angle1 = 0; *angle2 = 0;
  if (!j->node[0].body) return;
  ObJoint o;
  universal_fill(j, o);
  dReal a1, a2;
  ob_universal_angles(o, j->node[0].body->R, j->node[0].body->q, j->node[1].body ? j->node[1].body->R : 0, j->node[1].body ? j->node[1].body->q : 0, &a1, &a2);
  if (j->flags & dJOINT_REVERSE) { *angle1 = a2; *angle2 = -a1; } else { *angle1 = a1; *angle2 = a2; }   // universal.cpp:642-655
}
dReal dJointGetUniversalAngle1(dJointID j) { dReal a, b; dJointGetUniversalAngles(j, &a, &b); return a; }
dReal dJointGetUniversalAngle2(dJointID j) { dReal a, b; dJointGetUniversalAngles(j, &a, &b); return b; }
}  // extern "C"

// ---- amotor (amotor.cpp) and lmotor (lmotor.cpp) -----------------------------------------------------------
static dReal *motor_axis(dxJoint *j, int anum) { return anum == 0 ? j->axis1 : (anum == 1 ? j->axis2 : j->axis3); }
static dxLimot &motor_limot(dxJoint *j, int anum) { return anum == 0 ? j->limot : (anum == 1 ? j->limot2 : j->limot3); }
static void amotor_set_euler_reference(dxJoint *j) {   // setEulerReferenceVectors :128-158
  if (j->node[0].body && j->node[1].body) {
    dReal r[4];
    ob_mul0_331(r, j->node[1].body->R, j->axis3);
    ob_mul1_331(j->reference1, j->node[0].body->R, r);
    ob_mul0_331(r, j->node[0].body->R, j->axis1);
    ob_mul1_331(j->reference2, j->node[1].body->R, r);
  } else if (j->node[0].body) {
    dReal r[4] = {j->axis3[0], j->axis3[1], j->axis3[2], j->axis3[3]};
    ob_mul1_331(j->reference1, j->node[0].body->R, r);
    ob_mul0_331(r, j->node[0].body->R, j->axis1);
    j->reference2[0] += r[0]; j->reference2[1] += r[1]; j->reference2[2] += r[2];   // sic
  }
}
static void motor_set_axis(dxJoint *j, int anum, int rel, dReal x, dReal y, dReal z, bool amotor) {
  if (anum < 0) anum = 0;
  if (anum > 2) anum = 2;
  if (!j->node[1].body && rel == 2) rel = 1;
  j->rel[anum] = rel;
  dReal r[4] = {x, y, z, 0};
  dReal *axis = motor_axis(j, anum);
  if (rel > 0) {
    if (rel == 1) ob_mul1_331(axis, j->node[0].body->R, r);
    else if (j->node[1].body) ob_mul1_331(axis, j->node[1].body->R, r);
    else { axis[0] = r[0]; axis[1] = r[1]; axis[2] = r[2]; axis[3] = r[3]; }
  } else { axis[0] = r[0]; axis[1] = r[1]; axis[2] = r[2]; }
  ob_safe_normalize3(axis);
  if (amotor && j->mode == dAMotorEuler) amotor_set_euler_reference(j);
}
static void motor_get_axis(dxJoint *j, int anum, dReal *result) {
  if (anum < 0) anum = 0;
  if (anum > 2) anum = 2;
  const dReal *axis = motor_axis(j, anum);
  if (j->rel[anum] == 1) ob_mul0_331(result, j->node[0].body->R, axis);
  else if (j->rel[anum] == 2 && j->node[1].body) ob_mul0_331(result, j->node[1].body->R, axis);
  else { result[0] = axis[0]; result[1] = axis[1]; result[2] = axis[2]; }
}
extern "C" {
void dJointSetAMotorNumAxes(dJointID j, int num) { if (j->mode == dAMotorEuler) j->num = 3; else j->num = num < 0 ? 0 : (num > 3 ? 3 : num); }
void dJointSetAMotorAxis(dJointID j, int anum, int rel, dReal x, dReal y, dReal z) { motor_set_axis(j, anum, rel, x, y, z, true); }
void dJointSetAMotorAngle(dJointID j, int anum, dReal angle) { if (j->mode == dAMotorUser) { if (anum < 0) anum = 0; if (anum > 2) anum = 2; j->angle[anum] = angle; } }
void dJointSetAMotorParam(dJointID j, int parameter, dReal value) { int anum = parameter >> 8; if (anum < 0) anum = 0; if (anum > 2) anum = 2; limot_set(motor_limot(j, anum), parameter & 0xff, value); }
void dJointSetAMotorMode(dJointID j, int mode) { j->mode = mode; if (mode == dAMotorEuler) { j->num = 3; amotor_set_euler_reference(j); } }
int dJointGetAMotorNumAxes(dJointID j) { return j->num; }
void dJointGetAMotorAxis(dJointID j, int anum, dVector3 result) { motor_get_axis(j, anum, result); }
int dJointGetAMotorAxisRel(dJointID j, int anum) { if (anum < 0) anum = 0; if (anum > 2) anum = 2; return j->rel[anum]; }
dReal dJointGetAMotorAngleRate(dJointID, int) { ob_debug(0, "not yet implemented"); return 0; }   // amotor.cpp:445-451: the reference aborts the same way
dReal dJointGetAMotorAngle(dJointID j, int anum) { if (anum < 0) anum = 0; if (anum > 2) anum = 2; return j->angle[anum]; }
dReal dJointGetAMotorParam(dJointID j, int parameter) { int anum = parameter >> 8; if (anum < 0) anum = 0; if (anum > 2) anum = 2; return limot_get(motor_limot(j, anum), parameter & 0xff); }
int dJointGetAMotorMode(dJointID j) { return j->mode; }
void dJointSetLMotorNumAxes(dJointID j, int num) { j->num = num < 0 ? 0 : (num > 3 ? 3 : num); }
void dJointSetLMotorAxis(dJointID j, int anum, int rel, dReal x, dReal y, dReal z) { motor_set_axis(j, anum, rel, x, y, z, false); }
void dJointSetLMotorParam(dJointID j, int parameter, dReal value) { int anum = parameter >> 8; if (anum < 0) anum = 0; if (anum > 2) anum = 2; limot_set(motor_limot(j, anum), parameter & 0xff, value); }
int dJointGetLMotorNumAxes(dJointID j) { return j->num; }
void dJointGetLMotorAxis(dJointID j, int anum, dVector3 result) { if (anum < 0) anum = 0; if (anum > 2) anum = 2; const dReal *a = motor_axis(j, anum); result[0] = a[0]; result[1] = a[1]; result[2] = a[2]; }
dReal dJointGetLMotorParam(dJointID j, int parameter) { int anum = parameter >> 8; if (anum < 0) anum = 0; if (anum > 2) anum = 2; return limot_get(motor_limot(j, anum), parameter & 0xff); }

// ---- plane2d (plane2d.cpp), piston (piston.cpp), PR (pr.cpp) ------------------------------------------
void dJointSetPlane2DXParam(dJointID j, int parameter, dReal value) { limot_set(j->limot, parameter, value); }
void dJointSetPlane2DYParam(dJointID j, int parameter, dReal value) { limot_set(j->limot2, parameter, value); }
void dJointSetPlane2DAngleParam(dJointID j, int parameter, dReal value) { limot_set(j->limot3, parameter, value); }

static void get_axis(dxJoint *j, dReal *result, const dReal *axis) {   // joint.cpp getAxis
  if (j->node[0].body) ob_mul0_331(result, j->node[0].body->R, axis);
}
// dJointGetPistonPosition (piston.cpp:54-107) / dJointGetPRPosition (pr.cpp:75-125)
static dReal prismatic_position(dxJoint *j, const dReal *anchor1, const dReal *axisP) {
  if (!j->node[0].body) return 0;
  dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
  dReal q[4], ax[4];
  ob_mul0_331(q, b0->R, anchor1);
  if (b1) {
    dReal a2[4];
    ob_mul0_331(a2, b1->R, j->anchor2);
    for (int i = 0; i < 3; i++) q[i] = (b0->pos[i] + q[i]) - (b1->pos[i] + a2[i]);
  } else {
    for (int i = 0; i < 3; i++) q[i] = (b0->pos[i] + q[i]) - j->anchor2[i];
    if (j->flags & dJOINT_REVERSE) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; }
  }
  ob_mul0_331(ax, b0->R, axisP);
  return ob_dot(ax, q);
}
static dReal rotoide_angle(dxJoint *j) {
  if (!j->node[0].body) return 0;
  dReal ang = ob_hinge_angle(j->node[0].body->q, j->node[1].body ? j->node[1].body->q : 0, j->axis1, j->qrel);
  return (j->flags & dJOINT_REVERSE) ? -ang : ang;
}
static dReal rotoide_rate(dxJoint *j) {
  if (!j->node[0].body) return 0;
  dReal axis[4];
  ob_mul0_331(axis, j->node[0].body->R, j->axis1);
  dReal rate = ob_dot(axis, j->node[0].body->avel);
  if (j->node[1].body) rate -= ob_dot(axis, j->node[1].body->avel);
  if (j->flags & dJOINT_REVERSE) rate = -rate;
  return rate;
}
static void two_limot_set(dxJoint *j, int parameter, dReal value) {   // group 2 (0x100) = rotoide, else prismatic
  if ((parameter & 0xff00) == 0x100) limot_set(j->limot2, parameter & 0xff, value);
  else limot_set(j->limot, parameter, value);
}
static dReal two_limot_get(dxJoint *j, int parameter) {
  if ((parameter & 0xff00) == 0x100) return limot_get(j->limot2, parameter & 0xff);
  return limot_get(j->limot, parameter);
}

void dJointSetPistonAnchor(dJointID j, dReal x, dReal y, dReal z) {
  set_anchors(j, x, y, z, j->anchor1, j->anchor2);
  hinge_initial_rel_rot(j);
}
void dJointSetPistonAnchorOffset(dJointID j, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz) {   // piston.cpp:434-466
  if (j->flags & dJOINT_REVERSE) { dx = -dx; dy = -dy; dz = -dz; }
  dxBody *b0 = j->node[0].body;
  if (b0) { b0->pos[0] -= dx; b0->pos[1] -= dy; b0->pos[2] -= dz; }
  set_anchors(j, x, y, z, j->anchor1, j->anchor2);
  if (b0) { b0->pos[0] += dx; b0->pos[1] += dy; b0->pos[2] += dz; }
  hinge_initial_rel_rot(j);
}
void dJointGetPistonAnchor(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor2(j, result, j->anchor2);
  else get_anchor(j, result, j->anchor1);
}
void dJointGetPistonAnchor2(dJointID j, dVector3 result) {
  if (j->flags & dJOINT_REVERSE) get_anchor(j, result, j->anchor1);
  else get_anchor2(j, result, j->anchor2);
}
void dJointSetPistonAxis(dJointID j, dReal x, dReal y, dReal z) {
  set_axes(j, x, y, z, j->axis1, j->axis2);
  hinge_initial_rel_rot(j);
}
void dJointGetPistonAxis(dJointID j, dVector3 result) { get_axis(j, result, j->axis1); }
void dJointSetPistonAxisDelta(dJointID j, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz) {   // piston.cpp:508-539
  set_axes(j, x, y, z, j->axis1, j->axis2);
  hinge_initial_rel_rot(j);
  dReal c[4] = {0, 0, 0, 0};
  dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
  if (b1) { c[0] = (b0->pos[0] - b1->pos[0] - dx); c[1] = (b0->pos[1] - b1->pos[1] - dy); c[2] = (b0->pos[2] - b1->pos[2] - dz); }
  else if (b0) { c[0] = b0->pos[0] - dx; c[1] = b0->pos[1] - dy; c[2] = b0->pos[2] - dz; }
  ob_mul1_331(j->anchor1, b0->R, c);
}
void dJointSetPistonParam(dJointID j, int parameter, dReal value) { two_limot_set(j, parameter, value); }
dReal dJointGetPistonParam(dJointID j, int parameter) { return two_limot_get(j, parameter); }
dReal dJointGetPistonPosition(dJointID j) { return prismatic_position(j, j->anchor1, j->axis1); }
dReal dJointGetPistonPositionRate(dJointID j) {   // piston.cpp:110-133
  dReal ax[4];
  ob_mul0_331(ax, j->node[0].body->R, j->axis1);
  if (j->node[1].body) return ob_dot(ax, j->node[0].body->lvel) - ob_dot(ax, j->node[1].body->lvel);
  dReal rate = ob_dot(ax, j->node[0].body->lvel);
  return (j->flags & dJOINT_REVERSE) ? -rate : rate;
}
dReal dJointGetPistonAngle(dJointID j) { return rotoide_angle(j); }
dReal dJointGetPistonAngleRate(dJointID j) { return rotoide_rate(j); }
void dJointAddPistonForce(dJointID j, dReal force) {   // piston.cpp:586-664
  if (j->flags & dJOINT_REVERSE) force -= force;   // sic
  dReal axis[4] = {0, 0, 0, 0};
  get_axis(j, axis, j->axis1);
  axis[0] *= force; axis[1] *= force; axis[2] *= force;
  dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
  if (b0) dBodyAddForce(b0, axis[0], axis[1], axis[2]);
  if (b1) dBodyAddForce(b1, -axis[0], -axis[1], -axis[2]);
  if (b0 && b1) {
    dReal ltd[4], c[4];
    ob_mul0_331(c, b0->R, j->anchor1);
    ob_cross(ltd, c, axis);
    dBodyAddTorque(b0, ltd[0], ltd[1], ltd[2]);
    ob_mul0_331(c, b1->R, j->anchor2);
    ob_cross(ltd, c, axis);
    dBodyAddTorque(b1, ltd[0], ltd[1], ltd[2]);
  }
}

void dJointSetPRAnchor(dJointID j, dReal x, dReal y, dReal z) { set_anchors(j, x, y, z, j->offset, j->anchor2); }
void dJointSetPRAxis1(dJointID j, dReal x, dReal y, dReal z) {
  set_axes(j, x, y, z, j->axis3, 0);
  hinge_initial_rel_rot(j);
}
void dJointSetPRAxis2(dJointID j, dReal x, dReal y, dReal z) {
  set_axes(j, x, y, z, j->axis1, j->axis2);
  hinge_initial_rel_rot(j);
}
void dJointGetPRAnchor(dJointID j, dVector3 result) {
  if (j->node[1].body) get_anchor2(j, result, j->anchor2);
  else { result[0] = j->anchor2[0]; result[1] = j->anchor2[1]; result[2] = j->anchor2[2]; }
}
void dJointGetPRAxis1(dJointID j, dVector3 result) { get_axis(j, result, j->axis3); }
void dJointGetPRAxis2(dJointID j, dVector3 result) { get_axis(j, result, j->axis1); }
void dJointSetPRParam(dJointID j, int parameter, dReal value) { two_limot_set(j, parameter, value); }
dReal dJointGetPRParam(dJointID j, int parameter) { return two_limot_get(j, parameter); }
dReal dJointGetPRPosition(dJointID j) { return prismatic_position(j, j->offset, j->axis3); }
dReal dJointGetPRPositionRate(dJointID j) {   // pr.cpp:128-160
  dReal ax[4];
  ob_mul0_331(ax, j->node[0].body->R, j->axis3);
  if (j->node[1].body) {
    dVector3 lv2;
    dBodyGetRelPointVel(j->node[1].body, j->anchor2[0], j->anchor2[1], j->anchor2[2], lv2);
    return ob_dot(ax, j->node[0].body->lvel) - ob_dot(ax, lv2);
  }
  dReal rate = ob_dot(ax, j->node[0].body->lvel);
  return (j->flags & dJOINT_REVERSE) ? -rate : rate;
}
dReal dJointGetPRAngle(dJointID j) { return rotoide_angle(j); }
dReal dJointGetPRAngleRate(dJointID j) { return rotoide_rate(j); }
void dJointAddPRTorque(dJointID j, dReal torque) {   // pr.cpp:563-583
  dReal axis[4] = {0, 0, 0, 0};
  if (j->flags & dJOINT_REVERSE) torque = -torque;
  get_axis(j, axis, j->axis1);
  axis[0] *= torque; axis[1] *= torque; axis[2] *= torque;
  if (j->node[0].body) dBodyAddTorque(j->node[0].body, axis[0], axis[1], axis[2]);
  if (j->node[1].body) dBodyAddTorque(j->node[1].body, -axis[0], -axis[1], -axis[2]);
}

// ---- PU, prismatic + universal (pu.cpp; the universal part is dxJointUniversal's) -----------------------
void dJointSetPUAnchor(dJointID j, dReal x, dReal y, dReal z) { set_anchors(j, x, y, z, j->anchor1, j->anchor2); universal_initial_rel_rots(j); }
void dJointSetPUAnchorDelta(dJointID j, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz) {   // pu.cpp:418-440
  dxBody *b0 = j->node[0].body;
  if (b0) { b0->pos[0] += dx; b0->pos[1] += dy; b0->pos[2] += dz; }
  set_anchors(j, x, y, z, j->anchor1, j->anchor2);
  if (b0) { b0->pos[0] -= dx; b0->pos[1] -= dy; b0->pos[2] -= dz; }
  universal_initial_rel_rots(j);
}
void dJointSetPUAnchorOffset(dJointID j, dReal x, dReal y, dReal z, dReal dx, dReal dy, dReal dz) {   // pu.cpp:473-503
  if (j->flags & dJOINT_REVERSE) { dx = -dx; dy = -dy; dz = -dz; }
  dxBody *b0 = j->node[0].body;
  if (b0) { b0->pos[0] -= dx; b0->pos[1] -= dy; b0->pos[2] -= dz; }
  set_anchors(j, x, y, z, j->anchor1, j->anchor2);
  if (b0) { b0->pos[0] += dx; b0->pos[1] += dy; b0->pos[2] += dz; }
  universal_initial_rel_rots(j);
}
void dJointSetPUAxis1(dJointID j, dReal x, dReal y, dReal z) { dJointSetUniversalAxis1(j, x, y, z); }
void dJointSetPUAxis2(dJointID j, dReal x, dReal y, dReal z) { dJointSetUniversalAxis2(j, x, y, z); }
void dJointSetPUAxis3(dJointID j, dReal x, dReal y, dReal z) {
  set_axes(j, x, y, z, j->axis3, 0);
  universal_initial_rel_rots(j);
}
void dJointSetPUAxisP(dJointID j, dReal x, dReal y, dReal z) { dJointSetPUAxis3(j, x, y, z); }
void dJointGetPUAnchor(dJointID j, dVector3 result) {
  if (j->node[1].body) get_anchor2(j, result, j->anchor2);
  else { result[0] = j->anchor2[0]; result[1] = j->anchor2[1]; result[2] = j->anchor2[2]; }
}
void dJointGetPUAxis1(dJointID j, dVector3 result) { dJointGetUniversalAxis1(j, result); }
void dJointGetPUAxis2(dJointID j, dVector3 result) { dJointGetUniversalAxis2(j, result); }
void dJointGetPUAxis3(dJointID j, dVector3 result) { get_axis(j, result, j->axis3); }
void dJointGetPUAxisP(dJointID j, dVector3 result) { dJointGetPUAxis3(j, result); }
void dJointSetPUParam(dJointID j, int parameter, dReal value) {
  switch (parameter & 0xff00) {
    case 0x000: limot_set(j->limot, parameter, value); break;
    case 0x100: limot_set(j->limot2, parameter & 0xff, value); break;
    case 0x200: limot_set(j->limot3, parameter & 0xff, value); break;
  }
}
dReal dJointGetPUParam(dJointID j, int parameter) {
  switch (parameter & 0xff00) {
    case 0x000: return limot_get(j->limot, parameter);
    case 0x100: return limot_get(j->limot2, parameter & 0xff);
    case 0x200: return limot_get(j->limot3, parameter & 0xff);
  }
  return 0;
}
void dJointGetPUAngles(dJointID j, dReal *angle1, dReal *angle2) {   // pu.cpp:545-553: swapped, not negated, when reversed
  *angle1 = 0; *angle2 = 0;
  if (!j->node[0].body) return;
  ObJoint o;
  universal_fill(j, o);
  dReal a1, a2;
  ob_universal_angles(o, j->node[0].body->R, j->node[0].body->q, j->node[1].body ? j->node[1].body->R : 0, j->node[1].body ? j->node[1].body->q : 0, &a1, &a2);
  if (j->flags & dJOINT_REVERSE) { *angle2 = a1; *angle1 = a2; } else { *angle1 = a1; *angle2 = a2; }
}
dReal dJointGetPUAngle1(dJointID j) { dReal a, b; dJointGetPUAngles(j, &a, &b); return a; }
dReal dJointGetPUAngle2(dJointID j) { dReal a, b; dJointGetPUAngles(j, &a, &b); return b; }
static dReal pu_angle_rate(dxJoint *j, int second) {   // pu.cpp:575-615
  if (!j->node[0].body) return 0;
  dReal axis[4] = {0, 0, 0, 0};
  if (second) dJointGetUniversalAxis2(j, axis); else dJointGetUniversalAxis1(j, axis);
  dReal rate = ob_dot(axis, j->node[0].body->avel);
  if (j->node[1].body) rate -= ob_dot(axis, j->node[1].body->avel);
  return rate;
}
dReal dJointGetPUAngle1Rate(dJointID j) { return pu_angle_rate(j, 0); }
dReal dJointGetPUAngle2Rate(dJointID j) { return pu_angle_rate(j, 1); }
dReal dJointGetPUPosition(dJointID j) { return prismatic_position(j, j->anchor1, j->axis3); }
dReal dJointGetPUPositionRate(dJointID j) {   // pu.cpp:128-180
  dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
  if (!b0) return 0;
  dReal r[4], anchor2[4] = {0, 0, 0, 0}, lvel1[4], axP1[4];
  if (b1) {
    ob_mul0_331(anchor2, b1->R, j->anchor2);
    for (int i = 0; i < 3; i++) r[i] = b0->pos[i] - (anchor2[i] + b1->pos[i]);
  } else {
    for (int i = 0; i < 3; i++) r[i] = b0->pos[i] - j->anchor2[i];
  }
  ob_cross(lvel1, r, b0->avel);
  for (int i = 0; i < 3; i++) lvel1[i] = lvel1[i] + b0->lvel[i];
  ob_mul0_331(axP1, b0->R, j->axis3);
  if (b1) {
    dReal lvel2[4], tmp[4];
    ob_cross(lvel2, anchor2, b1->avel);
    for (int i = 0; i < 3; i++) tmp[i] = lvel2[i] + b1->lvel[i];
    for (int i = 0; i < 3; i++) lvel1[i] = lvel1[i] - tmp[i];
    return ob_dot(axP1, lvel1);
  }
  dReal rate = ob_dot(axP1, lvel1);
  return (j->flags & dJOINT_REVERSE) ? -rate : rate;
}
}  // extern "C"

// Host-visible side effects of getInfo1 that the device does not write back: an Euler-mode amotor stores the angles it
// measured in the joint (amotor.cpp:150-160, read back by dJointGetAMotorAngle and dWorldExportDIF).  Called by the
// drop-in dWorldQuickStep with the pre-step body state; same function as the device uses, so the values are the same bits.
void ob_marshal_joint(const dxJoint *j, ObJoint &d);
void ob_joints_prestep_bookkeeping(dxWorld *w) {
  for (dxJoint *j = w->firstjoint; j; j = j->next) {
    if (j->type != dJointTypeAMotor || j->mode != dAMotorEuler || (j->flags & dJOINT_DISABLED)) continue;
    dxBody *b0 = j->node[0].body, *b1 = j->node[1].body;
    if (!b0) continue;
    if ((b0->flags & OB_BODY_DISABLED) && (!b1 || (b1->flags & OB_BODY_DISABLED))) continue;
    ObJoint o;
    ob_marshal_joint(j, o);
    ObBodyView B1 = {b0->pos, b0->R, b0->q, b0->lvel, b0->avel}, B2 = B1;
    if (b1) { B2.pos = b1->pos; B2.R = b1->R; B2.q = b1->q; B2.lvel = b1->lvel; B2.avel = b1->avel; }
    real ax[3][3], ang[3];
    ob_amotor_axes(o, B1, b1 ? &B2 : (const ObBodyView *)0, ax);
    ob_amotor_euler_angles(o, B1, b1 ? &B2 : (const ObBodyView *)0, ax, ang);
    j->angle[0] = ang[0]; j->angle[1] = ang[1]; j->angle[2] = ang[2];
  }
}

// setRelativeValues, called from dJointAttach (ball.cpp, hinge.cpp, hinge2.cpp)
void ob_joint_set_relative_values(dxJoint *j) {
  dReal v[4] = {0, 0, 0, 0};
  switch (j->type) {
    case dJointTypeBall:
      dJointGetBallAnchor(j, v);
      set_anchors(j, v[0], v[1], v[2], j->anchor1, j->anchor2);
      break;
    case dJointTypeHinge:
      dJointGetHingeAnchor(j, v);
      set_anchors(j, v[0], v[1], v[2], j->anchor1, j->anchor2);
      dJointGetHingeAxis(j, v);
      set_axes(j, v[0], v[1], v[2], j->axis1, j->axis2);
      hinge_initial_rel_rot(j);
      break;
    case dJointTypeHinge2: {
      dJointGetHinge2Anchor(j, v);
      set_anchors(j, v[0], v[1], v[2], j->anchor1, j->anchor2);
      dReal axis[4] = {0, 0, 0, 0};
      if (j->node[0].body) { dJointGetHinge2Axis1(j, axis); set_axes(j, axis[0], axis[1], axis[2], j->axis1, 0); }
      if (j->node[0].body) { dJointGetHinge2Axis2(j, axis); set_axes(j, axis[0], axis[1], axis[2], 0, j->axis2); }
      dReal ax1[4], ax2[4];
      if (j->node[0].body && j->node[1].body) hinge2_axis_info(j, ax1, ax2, axis, &j->s0, &j->c0);
      hinge2_make_v1v2(j);
    } break;
    case dJointTypeSlider:
      slider_compute_offset(j);
      hinge_initial_rel_rot(j);
      break;
    case dJointTypeUniversal: {   // universal.cpp setRelativeValues
      dJointGetUniversalAnchor(j, v);
      set_anchors(j, v[0], v[1], v[2], j->anchor1, j->anchor2);
      dReal ax1[4] = {0, 0, 0, 0}, ax2[4] = {0, 0, 0, 0};
      dJointGetUniversalAxis1(j, ax1);
      dJointGetUniversalAxis2(j, ax2);
      if (j->flags & dJOINT_REVERSE) { set_axes(j, ax1[0], ax1[1], ax1[2], 0, j->axis2); set_axes(j, ax2[0], ax2[1], ax2[2], j->axis1, 0); }
      else { set_axes(j, ax1[0], ax1[1], ax1[2], j->axis1, 0); set_axes(j, ax2[0], ax2[1], ax2[2], 0, j->axis2); }
      universal_initial_rel_rots(j);
    } break;
    case dJointTypePiston:   // piston.cpp:682-693
      dJointGetPistonAnchor(j, v);
      set_anchors(j, v[0], v[1], v[2], j->anchor1, j->anchor2);
      dJointGetPistonAxis(j, v);
      set_axes(j, v[0], v[1], v[2], j->axis1, j->axis2);
      hinge_initial_rel_rot(j);
      break;
    case dJointTypePR:       // pr.cpp:598-613
      dJointGetPRAnchor(j, v);
      set_anchors(j, v[0], v[1], v[2], j->offset, j->anchor2);
      dJointGetPRAxis1(j, v);
      set_axes(j, v[0], v[1], v[2], j->axis3, 0);
      dJointGetPRAxis2(j, v);
      set_axes(j, v[0], v[1], v[2], j->axis1, j->axis2);
      hinge_initial_rel_rot(j);
      break;
    case dJointTypePU: {     // pu.cpp:826-852 (sic: the prismatic axis goes through the axis2 slot of setAxes)
      dJointGetPUAnchor(j, v);
      set_anchors(j, v[0], v[1], v[2], j->anchor1, j->anchor2);
      dReal ax1[4] = {0, 0, 0, 0}, ax2[4] = {0, 0, 0, 0}, ax3[4] = {0, 0, 0, 0};
      dJointGetPUAxis1(j, ax1);
      dJointGetPUAxis2(j, ax2);
      dJointGetPUAxis3(j, ax3);
      if (j->flags & dJOINT_REVERSE) { set_axes(j, ax1[0], ax1[1], ax1[2], 0, j->axis2); set_axes(j, ax2[0], ax2[1], ax2[2], j->axis1, 0); }
      else { set_axes(j, ax1[0], ax1[1], ax1[2], j->axis1, 0); set_axes(j, ax2[0], ax2[1], ax2[2], 0, j->axis2); }
      set_axes(j, ax3[0], ax3[1], ax3[2], 0, j->axis3);
      universal_initial_rel_rots(j);
    } break;
    default: break;
  }
}

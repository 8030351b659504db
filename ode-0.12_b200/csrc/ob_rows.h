// ob_rows.h — constraint-row assembly (dxJoint::getInfo1/getInfo2) per joint.
//
// Contact joint: ode/src/joints/contact.cpp:46-71 (getInfo1), :74-256 (getInfo2).
// The driver defaults applied before getInfo2 (J=0, c=0, cfm=global_cfm,
// lo=-inf, hi=+inf, findex=-1) are ode/src/quickstep.cpp:715-747.
#pragma once
#include "ob_types.h"

template <int MAXM> struct ObRowOutT {      // one joint's rows, thread-local
  real J[MAXM][12];
  real c[MAXM], cfm[MAXM], lo[MAXM], hi[MAXM];
  int findex[MAXM];       // joint-local (-1 or row offset inside the joint)
};
typedef ObRowOutT<6> ObRowOut;    // any joint
typedef ObRowOutT<3> ObRowOut3;   // contact joints

template <class RO> OB_HD void ob_rows_defaults(RO &r, int m, real global_cfm) {
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < 12; j++) r.J[i][j] = 0;
    r.c[i] = 0; r.cfm[i] = global_cfm; r.lo[i] = -OB_INF; r.hi[i] = OB_INF; r.findex[i] = -1;
  }
}

// getInfo1: number of rows; also clamps negative mu (contact.cpp:50-66)
OB_HD int ob_contact_info1(ObSurface &s) {
  int m = 1;
  if (s.mu < 0) s.mu = 0;
  if (s.mode & 0x001 /*dContactMu2*/) {
    if ((s.mu > 0) || (s.mu2 > 0)) m++;
    if (s.mu2 < 0) s.mu2 = 0;
    if (s.mu2 > 0) m++;
  } else {
    if (s.mu > 0) m += 2;
  }
  return m;
}

// getInfo2.  `normal_in` is contact.geom.normal, `reverse` = dJOINT_REVERSE.
// b1* are node[0].body's state; has_b2 tells whether node[1].body exists.
template <class RO>
OB_HD void ob_contact_info2(RO &r, int the_m, const ObSurface &sf, const real *cpos, const real *normal_in,
                            real cdepth, const real *fdir1, int reverse, const real *b1pos, const real *b1lvel,
                            const real *b1avel, int has_b2, const real *b2pos, const real *b2lvel,
                            const real *b2avel, real fps, real erp_in, real min_depth, real maxvel) {
  real normal[3];
  if (reverse) { normal[0] = -normal_in[0]; normal[1] = -normal_in[1]; normal[2] = -normal_in[2]; }
  else { normal[0] = normal_in[0]; normal[1] = normal_in[1]; normal[2] = normal_in[2]; }
  real c1[3], c2[3] = {0, 0, 0};
  c1[0] = cpos[0] - b1pos[0]; c1[1] = cpos[1] - b1pos[1]; c1[2] = cpos[2] - b1pos[2];
  real *J0 = r.J[0];
  J0[0] = normal[0]; J0[1] = normal[1]; J0[2] = normal[2];
  ob_cross(J0 + 3, c1, normal);
  if (has_b2) {
    c2[0] = cpos[0] - b2pos[0]; c2[1] = cpos[1] - b2pos[1]; c2[2] = cpos[2] - b2pos[2];
    J0[6] = -normal[0]; J0[7] = -normal[1]; J0[8] = -normal[2];
    ob_cross(J0 + 9, c2, normal);
    J0[9] = -J0[9]; J0[10] = -J0[10]; J0[11] = -J0[11];
  }
  real erp = erp_in;
  if (sf.mode & 0x008 /*SoftERP*/) erp = sf.soft_erp;
  real k = fps * erp;
  real depth = cdepth - min_depth;
  if (depth < 0) depth = 0;
  if (sf.mode & 0x010 /*SoftCFM*/) r.cfm[0] = sf.soft_cfm;
  real motionN = 0;
  if (sf.mode & 0x080 /*MotionN*/) motionN = sf.motionN;
  const real pushout = k * depth + motionN;
  r.c[0] = pushout;
  if (r.c[0] > maxvel) r.c[0] = maxvel;
  if (sf.mode & 0x004 /*Bounce*/) {
    real outgoing = ob_dot(J0, b1lvel) + ob_dot(J0 + 3, b1avel);
    if (has_b2) outgoing += ob_dot(J0 + 6, b2lvel) + ob_dot(J0 + 9, b2avel);
    outgoing -= motionN;
    if (sf.bounce_vel >= 0 && (-outgoing) > sf.bounce_vel) {
      real newc = -sf.bounce * outgoing + motionN;
      if (newc > r.c[0]) r.c[0] = newc;
    }
  }
  r.lo[0] = 0;
  r.hi[0] = OB_INF;

  real t1[3], t2[3];
  if (the_m >= 2) {
    if (sf.mode & 0x002 /*FDir1*/) {
      t1[0] = fdir1[0]; t1[1] = fdir1[1]; t1[2] = fdir1[2];
      ob_cross(t2, normal, t1);
    } else {
      ob_plane_space(normal, t1, t2);
    }
    real *J1 = r.J[1];
    J1[0] = t1[0]; J1[1] = t1[1]; J1[2] = t1[2];
    ob_cross(J1 + 3, c1, t1);
    if (has_b2) {
      J1[6] = -t1[0]; J1[7] = -t1[1]; J1[8] = -t1[2];
      ob_cross(J1 + 9, c2, t1);
      J1[9] = -J1[9]; J1[10] = -J1[10]; J1[11] = -J1[11];
    }
    if (sf.mode & 0x020 /*Motion1*/) r.c[1] = sf.motion1;
    r.lo[1] = -sf.mu;
    r.hi[1] = sf.mu;
    if (sf.mode & 0x1000 /*Approx1_1*/) r.findex[1] = 0;
    if (sf.mode & 0x100 /*Slip1*/) r.cfm[1] = sf.slip1;
  }
  if (the_m >= 3) {
    real *J2 = r.J[2];
    J2[0] = t2[0]; J2[1] = t2[1]; J2[2] = t2[2];
    ob_cross(J2 + 3, c1, t2);
    if (has_b2) {
      J2[6] = -t2[0]; J2[7] = -t2[1]; J2[8] = -t2[2];
      ob_cross(J2 + 9, c2, t2);
      J2[9] = -J2[9]; J2[10] = -J2[10]; J2[11] = -J2[11];
    }
    if (sf.mode & 0x040 /*Motion2*/) r.c[2] = sf.motion2;
    if (sf.mode & 0x001 /*Mu2*/) { r.lo[2] = -sf.mu2; r.hi[2] = sf.mu2; }
    else { r.lo[2] = -sf.mu; r.hi[2] = sf.mu; }
    if (sf.mode & 0x2000 /*Approx1_2*/) r.findex[2] = 0;
    if (sf.mode & 0x200 /*Slip2*/) r.cfm[2] = sf.slip2;
  }
}

// ======================================================================================
// ball / hinge / hinge2 (ode/src/joints/ball.cpp:48-64, hinge.cpp:53-149, hinge2.cpp:34-180,
// joint.cpp:85-196 setBall/setBall2, :381-450 hinge angle, :534-733 limit/motor rows)

// getHingeAngle (joint.cpp:396-450); q2 == 0 when the joint has one body
OB_HD real ob_hinge_angle(const real *q1, const real *q2, const real *axis, const real *q_initial) {
  real qrel[4];
  if (q2) { real qq[4]; ob_qmul1(qq, q1, q2); ob_qmul2(qrel, qq, q_initial); }
  else ob_qmul3(qrel, q1, q_initial);
  real cost2 = qrel[0];
  real sint2 = ob_sqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
  real theta = (ob_dot(qrel + 1, axis) >= 0) ? (2 * ob_atan2(sint2, cost2)) : (2 * ob_atan2(sint2, -cost2));
  if ((double)theta > OB_PI) theta -= (real)(2 * OB_PI);
  theta = -theta;
  return theta;
}
// getHingeAngleFromRelativeQuat (joint.cpp:381-420)
OB_HD real ob_hinge_angle_from_qrel(const real *qrel, const real *axis) {
  real cost2 = qrel[0];
  real sint2 = ob_sqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
  real theta = (ob_dot(qrel + 1, axis) >= 0) ? (2 * ob_atan2(sint2, cost2)) : (2 * ob_atan2(sint2, -cost2));
  if ((double)theta > OB_PI) theta -= (real)(2 * OB_PI);
  theta = -theta;
  return theta;
}
// dxJointUniversal::getAxes / getAngles (universal.cpp:52-169); qrel1 = ObJoint::qrel, qrel2 = ObJoint::v1
OB_HD void ob_universal_axes(const ObJoint &j, const real *R1, const real *R2, real *ax1, real *ax2) {
  ob_mul0_331(ax1, R1, j.axis1);
  if (R2) ob_mul0_331(ax2, R2, j.axis2);
  else { ax2[0] = j.axis2[0]; ax2[1] = j.axis2[1]; ax2[2] = j.axis2[2]; }
}
OB_HD void ob_universal_angles(const ObJoint &j, const real *R1, const real *q1, const real *R2, const real *q2, real *angle1, real *angle2) {
  real ax1[3], ax2[3], R[12], qcross[4], qq[4], qrel[4];
  ob_universal_axes(j, R1, R2, ax1, ax2);
  for (int i = 0; i < 12; i++) R[i] = 0;
  ob_Rfrom2axes(R, ax1[0], ax1[1], ax1[2], ax2[0], ax2[1], ax2[2]);
  ob_QfromR(qcross, R);
  ob_qmul1(qq, q1, qcross);
  ob_qmul2(qrel, qq, j.qrel);
  *angle1 = ob_hinge_angle_from_qrel(qrel, j.axis1);
  real qcross2[4];
  qrel[0] = 0;
  qrel[1] = ax1[0] + ax2[0]; qrel[2] = ax1[1] + ax2[1]; qrel[3] = ax1[2] + ax2[2];
  const real l = ob_recip(ob_sqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]));
  qrel[1] *= l; qrel[2] *= l; qrel[3] *= l;
  ob_qmul0(qcross2, qrel, qcross);
  if (q2) { ob_qmul1(qq, q2, qcross2); ob_qmul2(qrel, qq, j.v1); }
  else ob_qmul2(qrel, qcross2, j.v1);
  *angle2 = -ob_hinge_angle_from_qrel(qrel, j.axis2);
}
// dxJointHinge2::measureAngle (hinge2.cpp:34-43)
OB_HD real ob_hinge2_angle(const real *R1, const real *R2, const real *axis2, const real *v1, const real *v2) {
  real a1[3], a2[3];
  ob_mul0_331(a1, R2, axis2);
  ob_mul1_331(a2, R1, a1);
  real x = ob_dot(v1, a2);
  real y = ob_dot(v2, a2);
  return -ob_atan2(y, x);
}
// testRotationalLimit (joint.cpp:534-553)
OB_HD int ob_limot_test_limit(ObLimot &l, real angle) {
  if (angle <= l.lostop) { l.limit = 1; l.limit_err = angle - l.lostop; return 1; }
  else if (angle >= l.histop) { l.limit = 2; l.limit_err = angle - l.histop; return 1; }
  l.limit = 0;
  return 0;
}

struct ObBodyView {   // what row assembly reads from a body
  const real *pos, *R, *q, *lvel, *avel;
};

// dJointGetSliderPosition (slider.cpp:44-83)
OB_HD real ob_slider_position(const ObJoint &j, const ObBodyView &B1, const ObBodyView *B2) {
  real ax1[3], q[3];
  ob_mul0_331(ax1, B1.R, j.axis1);
  if (B2) {
    ob_mul0_331(q, B2->R, j.anchor1);
    for (int i = 0; i < 3; i++) q[i] = B1.pos[i] - q[i] - B2->pos[i];
  } else {
    q[0] = B1.pos[0] - j.anchor1[0]; q[1] = B1.pos[1] - j.anchor1[1]; q[2] = B1.pos[2] - j.anchor1[2];
    if (j.flags & OB_JF_REVERSE) { ax1[0] = -ax1[0]; ax1[1] = -ax1[1]; ax1[2] = -ax1[2]; }
  }
  return ob_dot(ax1, q);
}

// dxJointAMotor::computeGlobalAxes / computeEulerAngles (amotor.cpp:51-126); lmotor.cpp:44-66 is the non-Euler branch
// dJointGetPistonPosition / dJointGetPRPosition (piston.cpp:54-107, pr.cpp:75-125): the anchor (offset) carried by
// body 1 relative to anchor2, along the prismatic axis carried by body 1
OB_HD real ob_prismatic_position(const ObJoint &j, const real *anchor1, const real *axisP, const ObBodyView &B1, const ObBodyView *B2) {
  real q[3], ax[3];
  ob_mul0_331(q, B1.R, anchor1);
  if (B2) {
    real a2[3];
    ob_mul0_331(a2, B2->R, j.anchor2);
    q[0] = (B1.pos[0] + q[0]) - (B2->pos[0] + a2[0]);
    q[1] = (B1.pos[1] + q[1]) - (B2->pos[1] + a2[1]);
    q[2] = (B1.pos[2] + q[2]) - (B2->pos[2] + a2[2]);
  } else {
    q[0] = (B1.pos[0] + q[0]) - j.anchor2[0];
    q[1] = (B1.pos[1] + q[1]) - j.anchor2[1];
    q[2] = (B1.pos[2] + q[2]) - j.anchor2[2];
    if (j.flags & OB_JF_REVERSE) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; }
  }
  ob_mul0_331(ax, B1.R, axisP);
  return ob_dot(ax, q);
}

OB_HD void ob_motor_axis(const ObJoint &j, int i, const real *axis, const ObBodyView &B1, const ObBodyView *B2, real *out) {
  const int rel = OB_JM_REL(j.flags, i);
  if (rel == 1) ob_mul0_331(out, B1.R, axis);
  else if (rel == 2) { if (B2) ob_mul0_331(out, B2->R, axis); else { out[0] = 0; out[1] = 0; out[2] = 0; } }   // reference: left uninitialised
  else { out[0] = axis[0]; out[1] = axis[1]; out[2] = axis[2]; }
}
OB_HD void ob_amotor_axes(const ObJoint &j, const ObBodyView &B1, const ObBodyView *B2, real ax[3][3]) {
  if (OB_JM_MODE(j.flags) == 1) {
    ob_mul0_331(ax[0], B1.R, j.axis1);
    if (B2) ob_mul0_331(ax[2], B2->R, j.anchor1);
    else { ax[2][0] = j.anchor1[0]; ax[2][1] = j.anchor1[1]; ax[2][2] = j.anchor1[2]; }
    ob_cross(ax[1], ax[2], ax[0]);
    ob_safe_normalize3(ax[1]);
  } else {
    const int num = OB_JM_NUM(j.flags);
    const real *axs[3] = {j.axis1, j.axis2, j.anchor1};
    for (int i = 0; i < 3; i++) { ax[i][0] = ax[i][1] = ax[i][2] = 0; if (i < num) ob_motor_axis(j, i, axs[i], B1, B2, ax[i]); }
  }
}
OB_HD void ob_amotor_euler_angles(const ObJoint &j, const ObBodyView &B1, const ObBodyView *B2, real ax[3][3], real *angle) {
  real ref1[3], ref2[3], q[3];
  ob_mul0_331(ref1, B1.R, j.anchor2);
  if (B2) ob_mul0_331(ref2, B2->R, j.v1);
  else { ref2[0] = j.v1[0]; ref2[1] = j.v1[1]; ref2[2] = j.v1[2]; }
  ob_cross(q, ax[0], ref1);
  angle[0] = -ob_atan2(ob_dot(ax[2], q), ob_dot(ax[2], ref1));
  ob_cross(q, ax[0], ax[1]);
  angle[1] = -ob_atan2(ob_dot(ax[2], ax[0]), ob_dot(ax[2], q));
  ob_cross(q, ax[1], ax[2]);
  angle[2] = -ob_atan2(ob_dot(ref2, ax[1]), ob_dot(ref2, q));
}

// getInfo1 for permanent joints; mutates limot.limit / limit_err like the reference
OB_HD int ob_joint_info1(ObJoint &j, const ObBodyView &B1, const ObBodyView *B2) {
  if (j.type == OB_JOINT_BALL) return 3;
  if (j.type == OB_JOINT_HINGE) {
    int m = (j.limot1.fmax > 0) ? 6 : 5;
    if (((double)j.limot1.lostop >= -OB_PI || (double)j.limot1.histop <= OB_PI) && j.limot1.lostop <= j.limot1.histop) {
      real angle = ob_hinge_angle(B1.q, B2 ? B2->q : (const real *)0, j.axis1, j.qrel);
      if (ob_limot_test_limit(j.limot1, angle)) m = 6;
    }
    return m;
  }
  if (j.type == OB_JOINT_FIXED) return 6;
  if (j.type == OB_JOINT_LMOTOR) {   // lmotor.cpp:75-86
    int m = 0;
    const int num = OB_JM_NUM(j.flags);
    if (num > 0 && j.limot1.fmax > 0) m++;
    if (num > 1 && j.limot2.fmax > 0) m++;
    if (num > 2 && j.limot3.fmax > 0) m++;
    return m;
  }
  if (j.type == OB_JOINT_AMOTOR) {   // amotor.cpp:150-172
    const int num = OB_JM_NUM(j.flags);
    real angle[3] = {j.qrel[0], j.qrel[1], j.qrel[2]};
    if (OB_JM_MODE(j.flags) == 1 /* dAMotorEuler */) {
      real ax[3][3];
      ob_amotor_axes(j, B1, B2, ax);
      ob_amotor_euler_angles(j, B1, B2, ax, angle);
    }
    int m = 0;
    ObLimot *lm[3] = {&j.limot1, &j.limot2, &j.limot3};
    for (int i = 0; i < num; i++) if (ob_limot_test_limit(*lm[i], angle[i]) || lm[i]->fmax > 0) m++;
    return m;
  }
  if (j.type == OB_JOINT_UNIVERSAL) {   // universal.cpp:265-292
    int m = 4;
    const bool limiting1 = ((double)j.limot1.lostop >= -OB_PI || (double)j.limot1.histop <= OB_PI) && j.limot1.lostop <= j.limot1.histop;
    const bool limiting2 = ((double)j.limot2.lostop >= -OB_PI || (double)j.limot2.histop <= OB_PI) && j.limot2.lostop <= j.limot2.histop;
    j.limot1.limit = 0; j.limot2.limit = 0;
    if (limiting1 || limiting2) {
      real angle1, angle2;
      ob_universal_angles(j, B1.R, B1.q, B2 ? B2->R : (const real *)0, B2 ? B2->q : (const real *)0, &angle1, &angle2);
      if (limiting1) ob_limot_test_limit(j.limot1, angle1);
      if (limiting2) ob_limot_test_limit(j.limot2, angle2);
    }
    if (j.limot1.limit || j.limot1.fmax > 0) m++;
    if (j.limot2.limit || j.limot2.fmax > 0) m++;
    return m;
  }
  if (j.type == OB_JOINT_SLIDER) {   // slider.cpp:116-145
    int m = (j.limot1.fmax > 0) ? 6 : 5;
    j.limot1.limit = 0;
    if ((j.limot1.lostop > -OB_INF || j.limot1.histop < OB_INF) && j.limot1.lostop <= j.limot1.histop) {
      const real pos = ob_slider_position(j, B1, B2);
      if (pos <= j.limot1.lostop) { j.limot1.limit = 1; j.limot1.limit_err = pos - j.limot1.lostop; m = 6; }
      else if (pos >= j.limot1.histop) { j.limot1.limit = 2; j.limot1.limit_err = pos - j.limot1.histop; m = 6; }
    }
    return m;
  }
  if (j.type == OB_JOINT_PLANE2D) {   // plane2d.cpp:70-82
    int m = 3;
    if (j.limot1.fmax > 0) m++;
    if (j.limot2.fmax > 0) m++;
    if (j.limot3.fmax > 0) m++;
    return m;
  }
  if (j.type == OB_JOINT_PISTON || j.type == OB_JOINT_PR) {   // piston.cpp:181-217, pr.cpp:196-232; limot1 = prismatic, limot2 = rotoide
    int m = 4;
    j.limot1.limit = 0;
    if ((j.limot1.lostop > -OB_INF || j.limot1.histop < OB_INF) && j.limot1.lostop <= j.limot1.histop) {
      const real pos = ob_prismatic_position(j, j.anchor1, j.type == OB_JOINT_PR ? j.v1 : j.axis1, B1, B2);
      ob_limot_test_limit(j.limot1, pos);
    }
    if (j.limot1.limit || j.limot1.fmax > 0) m++;
    j.limot2.limit = 0;
    const bool limiting = j.type == OB_JOINT_PR ? ((double)j.limot2.lostop >= -OB_PI || (double)j.limot2.histop <= OB_PI)
                                                : (j.limot2.lostop > -OB_INF || j.limot2.histop < OB_INF);
    if (limiting && j.limot2.lostop <= j.limot2.histop) {
      const real angle = ob_hinge_angle(B1.q, B2 ? B2->q : (const real *)0, j.axis1, j.qrel);
      ob_limot_test_limit(j.limot2, angle);
    }
    if (j.limot2.limit || j.limot2.fmax > 0) m++;
    return m;
  }
  if (j.type == OB_JOINT_PU) {   // pu.cpp:188-232; limot1 / limot2 = universal axes, limot3 = prismatic
    int m = 3;
    j.limot3.limit = 0;
    if ((j.limot3.lostop > -OB_INF || j.limot3.histop < OB_INF) && j.limot3.lostop <= j.limot3.histop) {
      const real pos = ob_prismatic_position(j, j.anchor1, j.v2, B1, B2);
      ob_limot_test_limit(j.limot3, pos);
    }
    if (j.limot3.limit || j.limot3.fmax > 0) m++;
    const bool limiting1 = ((double)j.limot1.lostop >= -OB_PI || (double)j.limot1.histop <= OB_PI) && j.limot1.lostop <= j.limot1.histop;
    const bool limiting2 = ((double)j.limot2.lostop >= -OB_PI || (double)j.limot2.histop <= OB_PI) && j.limot2.lostop <= j.limot2.histop;
    j.limot1.limit = 0; j.limot2.limit = 0;
    if (limiting1 || limiting2) {
      real angle1, angle2;
      ob_universal_angles(j, B1.R, B1.q, B2 ? B2->R : (const real *)0, B2 ? B2->q : (const real *)0, &angle1, &angle2);
      if (limiting1) ob_limot_test_limit(j.limot1, angle1);
      if (limiting2) ob_limot_test_limit(j.limot2, angle2);
    }
    if (j.limot1.limit || j.limot1.fmax > 0) m++;
    if (j.limot2.limit || j.limot2.fmax > 0) m++;
    return m;
  }
  if (j.type == OB_JOINT_HINGE2) {
    int m = 4;
    j.limot1.limit = 0;
    if (((double)j.limot1.lostop >= -OB_PI || (double)j.limot1.histop <= OB_PI) && j.limot1.lostop <= j.limot1.histop) {
      real angle = ob_hinge2_angle(B1.R, B2->R, j.axis2, j.v1, j.v2);
      ob_limot_test_limit(j.limot1, angle);
    }
    if (j.limot1.limit || j.limot1.fmax > 0) m++;
    j.limot2.limit = 0;
    if (j.limot2.fmax > 0) m++;
    return m;
  }
  return 0;
}

// setBall (joint.cpp:85-126) into rows 0..2
template <class RO>
OB_HD void ob_set_ball(RO &r, const real *anchor1, const real *anchor2, const ObBodyView &B1, const ObBodyView *B2,
                       real fps, real erp) {
  real a1[3], a2[3] = {0, 0, 0};
  r.J[0][0] = 1; r.J[1][1] = 1; r.J[2][2] = 1;
  ob_mul0_331(a1, B1.R, anchor1);
  // dSetCrossMatrixMinus(J1a, a1)
  r.J[0][3 + 1] = a1[2]; r.J[0][3 + 2] = -a1[1];
  r.J[1][3 + 0] = -a1[2]; r.J[1][3 + 2] = a1[0];
  r.J[2][3 + 0] = a1[1]; r.J[2][3 + 1] = -a1[0];
  if (B2) {
    r.J[0][6] = -1; r.J[1][7] = -1; r.J[2][8] = -1;
    ob_mul0_331(a2, B2->R, anchor2);
    // dSetCrossMatrixPlus(J2a, a2)
    r.J[0][9 + 1] = -a2[2]; r.J[0][9 + 2] = a2[1];
    r.J[1][9 + 0] = a2[2]; r.J[1][9 + 2] = -a2[0];
    r.J[2][9 + 0] = -a2[1]; r.J[2][9 + 1] = a2[0];
  }
  real k = fps * erp;
  if (B2) { for (int j = 0; j < 3; j++) r.c[j] = k * (a2[j] + B2->pos[j] - a1[j] - B1.pos[j]); }
  else { for (int j = 0; j < 3; j++) r.c[j] = k * (anchor2[j] - a1[j] - B1.pos[j]); }
}

// setBall2 (joint.cpp:134-196) into rows 0..2
template <class RO>
OB_HD void ob_set_ball2(RO &r, const real *anchor1, const real *anchor2, const real *axis, real erp1,
                        const ObBodyView &B1, const ObBodyView *B2, real fps, real erp) {
  real a1[3], a2[3], q1[3], q2[3];
  ob_plane_space(axis, q1, q2);
  for (int i = 0; i < 3; i++) { r.J[0][i] = axis[i]; r.J[1][i] = q1[i]; r.J[2][i] = q2[i]; }
  ob_mul0_331(a1, B1.R, anchor1);
  ob_cross(r.J[0] + 3, a1, axis);
  ob_cross(r.J[1] + 3, a1, q1);
  ob_cross(r.J[2] + 3, a1, q2);
  if (B2) {
    for (int i = 0; i < 3; i++) { r.J[0][6 + i] = -axis[i]; r.J[1][6 + i] = -q1[i]; r.J[2][6 + i] = -q2[i]; }
    ob_mul0_331(a2, B2->R, anchor2);
    ob_cross(r.J[0] + 9, a2, axis); for (int i = 9; i < 12; i++) r.J[0][i] = -r.J[0][i];
    ob_cross(r.J[1] + 9, a2, q1); for (int i = 9; i < 12; i++) r.J[1][i] = -r.J[1][i];
    ob_cross(r.J[2] + 9, a2, q2); for (int i = 9; i < 12; i++) r.J[2][i] = -r.J[2][i];
  }
  real k1 = fps * erp1;
  real k = fps * erp;
  for (int i = 0; i < 3; i++) a1[i] += B1.pos[i];
  real d[3];
  if (B2) { for (int i = 0; i < 3; i++) a2[i] += B2->pos[i]; for (int i = 0; i < 3; i++) d[i] = a2[i] - a1[i]; }
  else { for (int i = 0; i < 3; i++) d[i] = anchor2[i] - a1[i]; }
  r.c[0] = k1 * ob_dot(axis, d);
  r.c[1] = k * ob_dot(q1, d);
  r.c[2] = k * ob_dot(q2, d);
}

// dxJointLimitMotor::addLimot, rotational case (joint.cpp:556-733).  The powered-at-limit
// side effect (dBodyAddTorque on both bodies, :638-645) is returned in `torque` (applied
// to body 1 as -torque... see caller): tq1 += -fm*ax1, tq2 += +fm*ax1.  Returns 1 if a row was added.
template <class RO>
OB_HD int ob_add_limot_rot(RO &r, int row, const ObLimot &l, const real *ax1, const ObBodyView &B1,
                           const ObBodyView *B2, real fps, real *side_fm /* out: fm or 0 */) {
  *side_fm = 0;
  int powered = l.fmax > 0;
  if (!(powered || l.limit)) return 0;
  r.J[row][3] = ax1[0]; r.J[row][4] = ax1[1]; r.J[row][5] = ax1[2];
  if (B2) { r.J[row][9] = -ax1[0]; r.J[row][10] = -ax1[1]; r.J[row][11] = -ax1[2]; }
  if (l.limit && (l.lostop == l.histop)) powered = 0;
  if (powered) {
    r.cfm[row] = l.normal_cfm;
    if (!l.limit) { r.c[row] = l.vel; r.lo[row] = -l.fmax; r.hi[row] = l.fmax; }
    else {
      real fm = l.fmax;
      if ((l.vel > 0) || (l.vel == 0 && l.limit == 2)) fm = -fm;
      if ((l.limit == 1 && l.vel > 0) || (l.limit == 2 && l.vel < 0)) fm *= l.fudge_factor;
      *side_fm = fm;
    }
  }
  if (l.limit) {
    real k = fps * l.stop_erp;
    r.c[row] = -k * l.limit_err;
    r.cfm[row] = l.stop_cfm;
    if (l.lostop == l.histop) { r.lo[row] = -OB_INF; r.hi[row] = OB_INF; }
    else {
      if (l.limit == 1) { r.lo[row] = 0; r.hi[row] = OB_INF; }
      else { r.lo[row] = -OB_INF; r.hi[row] = 0; }
      if (l.bounce > 0) {
        real vel = ob_dot(B1.avel, ax1);
        if (B2) vel -= ob_dot(B2->avel, ax1);
        if (l.limit == 1) { if (vel < 0) { real newc = -l.bounce * vel; if (newc > r.c[row]) r.c[row] = newc; } }
        else { if (vel > 0) { real newc = -l.bounce * vel; if (newc < r.c[row]) r.c[row] = newc; } }
      }
    }
  }
  return 1;
}

// dxJointLimitMotor::addLimot, linear case (joint.cpp:556-733 with rotational == 0): the row acts along ax1
// on the linear velocities; with two bodies the force is applied half way between the centres ("linear torque
// decoupling", ltd = c x ax1 in both angular blocks).  Powered at a limit: side_fm returns fm; the caller adds
// force -fm*ax1 / +fm*ax1 and torque -fm*ltd to BOTH bodies (:646-657).
template <class RO>
OB_HD int ob_add_limot_lin(RO &r, int row, const ObLimot &l, const real *ax1, const ObBodyView &B1, const ObBodyView *B2,
                           real fps, real *side_fm, real *ltd) {
  *side_fm = 0;
  ltd[0] = ltd[1] = ltd[2] = 0;
  int powered = l.fmax > 0;
  if (!(powered || l.limit)) return 0;
  r.J[row][0] = ax1[0]; r.J[row][1] = ax1[1]; r.J[row][2] = ax1[2];
  if (B2) {
    r.J[row][6] = -ax1[0]; r.J[row][7] = -ax1[1]; r.J[row][8] = -ax1[2];
    real c[3];
    c[0] = OB_REAL(0.5) * (B2->pos[0] - B1.pos[0]);
    c[1] = OB_REAL(0.5) * (B2->pos[1] - B1.pos[1]);
    c[2] = OB_REAL(0.5) * (B2->pos[2] - B1.pos[2]);
    ob_cross(ltd, c, ax1);
    for (int i = 0; i < 3; i++) { r.J[row][3 + i] = ltd[i]; r.J[row][9 + i] = ltd[i]; }
  }
  if (l.limit && (l.lostop == l.histop)) powered = 0;
  if (powered) {
    r.cfm[row] = l.normal_cfm;
    if (!l.limit) { r.c[row] = l.vel; r.lo[row] = -l.fmax; r.hi[row] = l.fmax; }
    else {
      real fm = l.fmax;
      if ((l.vel > 0) || (l.vel == 0 && l.limit == 2)) fm = -fm;
      if ((l.limit == 1 && l.vel > 0) || (l.limit == 2 && l.vel < 0)) fm *= l.fudge_factor;
      *side_fm = fm;
    }
  }
  if (l.limit) {
    real k = fps * l.stop_erp;
    r.c[row] = -k * l.limit_err;
    r.cfm[row] = l.stop_cfm;
    if (l.lostop == l.histop) { r.lo[row] = -OB_INF; r.hi[row] = OB_INF; }
    else {
      if (l.limit == 1) { r.lo[row] = 0; r.hi[row] = OB_INF; }
      else { r.lo[row] = -OB_INF; r.hi[row] = 0; }
      if (l.bounce > 0) {
        real vel = ob_dot(B1.lvel, ax1);
        if (B2) vel -= ob_dot(B2->lvel, ax1);
        if (l.limit == 1) { if (vel < 0) { real newc = -l.bounce * vel; if (newc > r.c[row]) r.c[row] = newc; } }
        else { if (vel > 0) { real newc = -l.bounce * vel; if (newc < r.c[row]) r.c[row] = newc; } }
      }
    }
  }
  return 1;
}

// setFixedOrientation (joint.cpp:202-255): three angular rows from start_row
template <class RO>
OB_HD void ob_set_fixed_orientation(RO &r, int start_row, const real *qrel, const ObBodyView &B1, const ObBodyView *B2, real fps, real erp) {
  r.J[start_row][3] = 1; r.J[start_row + 1][4] = 1; r.J[start_row + 2][5] = 1;
  if (B2) { r.J[start_row][9] = -1; r.J[start_row + 1][10] = -1; r.J[start_row + 2][11] = -1; }
  real qerr[4], e[3];
  if (B2) { real qq[4]; ob_qmul1(qq, B1.q, B2->q); ob_qmul2(qerr, qq, qrel); }
  else ob_qmul3(qerr, B1.q, qrel);
  if (qerr[0] < 0) { qerr[1] = -qerr[1]; qerr[2] = -qerr[2]; qerr[3] = -qerr[3]; }
  ob_mul0_331(e, B1.R, qerr + 1);
  const real k = fps * erp;
  r.c[start_row] = 2 * k * e[0];
  r.c[start_row + 1] = 2 * k * e[1];
  r.c[start_row + 2] = 2 * k * e[2];
}

// the bodies' accumulators after a joint's powered-at-limit motor side effects (joint.cpp:638-657), in the
// order the reference applies them.  side[k] = {fm, v[3]}: rotational limots (hinge, hinge2, universal):
// torque -fm*v on body 1, +fm*v on body 2; slider: side[0] = force (-fm*ax1 / +fm*ax1), side[1] = the
// decoupling torque, -fm*ltd on BOTH bodies.
OB_HD void ob_apply_joint_side(int jtype, const real side[OB_NSIDE][4], real *facc1, real *tacc1, real *facc2, real *tacc2) {
  if (jtype == OB_JOINT_SLIDER || jtype == OB_JOINT_PISTON || jtype == OB_JOINT_PR) {
    const real fm = side[0][0];
    if (fm != 0) {
      for (int e = 0; e < 3; e++) facc1[e] += -fm * side[0][1 + e];
      if (facc2) {
        for (int e = 0; e < 3; e++) facc2[e] += fm * side[0][1 + e];
        for (int e = 0; e < 3; e++) tacc1[e] += -fm * side[1][1 + e];
        for (int e = 0; e < 3; e++) tacc2[e] += -fm * side[1][1 + e];
      }
    }
    // piston / PR: the rotoide motor follows the prismatic one (slot 2)
    const real fr = side[2][0];
    if (fr != 0) {
      for (int e = 0; e < 3; e++) tacc1[e] += -fr * side[2][1 + e];
      if (tacc2) for (int e = 0; e < 3; e++) tacc2[e] += fr * side[2][1 + e];
    }
    return;
  }
  const int nrot = jtype == OB_JOINT_PU ? 2 : OB_NSIDE;
  for (int sx = 0; sx < nrot; sx++) {
    const real fm = side[sx][0];
    if (fm != 0) {
      for (int e = 0; e < 3; e++) tacc1[e] += -fm * side[sx][1 + e];
      if (tacc2) for (int e = 0; e < 3; e++) tacc2[e] += fm * side[sx][1 + e];
    }
  }
  if (jtype == OB_JOINT_PU) {   // the prismatic motor comes last: slot 2 = force, slot 3 = decoupling torque
    const real fm = side[2][0];
    if (fm != 0) {
      for (int e = 0; e < 3; e++) facc1[e] += -fm * side[2][1 + e];
      if (facc2) {
        for (int e = 0; e < 3; e++) facc2[e] += fm * side[2][1 + e];
        for (int e = 0; e < 3; e++) tacc1[e] += -fm * side[3][1 + e];
        for (int e = 0; e < 3; e++) tacc2[e] += -fm * side[3][1 + e];
      }
    }
  }
}

// getInfo2 for permanent joints.  erp_io carries the driver's shared Info2.erp, which a ball
// joint overwrites for every later joint of the island (ball.cpp:60, quickstep.cpp:764-786).
// side[k] (k<2) returns {fm, ax[3]} of a powered-at-limit motor whose torque must be added
// to the bodies' tacc before the rhs is formed: tacc1 += -fm*ax, tacc2 += fm*ax.
template <class RO>
OB_HD void ob_joint_info2(RO &r, const ObJoint &j, const ObBodyView &B1, const ObBodyView *B2, real fps,
                          real *erp_io, real side[OB_NSIDE][4]) {
  for (int sx = 0; sx < OB_NSIDE; sx++) side[sx][0] = 0;
  if (j.type == OB_JOINT_BALL) {
    *erp_io = j.erp;
    r.cfm[0] = j.cfm; r.cfm[1] = j.cfm; r.cfm[2] = j.cfm;
    ob_set_ball(r, j.anchor1, j.anchor2, B1, B2, fps, *erp_io);
  } else if (j.type == OB_JOINT_HINGE) {
    const real erp = *erp_io;
    ob_set_ball(r, j.anchor1, j.anchor2, B1, B2, fps, erp);
    real ax1[3], p[3], q[3];
    ob_mul0_331(ax1, B1.R, j.axis1);
    ob_plane_space(ax1, p, q);
    for (int i = 0; i < 3; i++) { r.J[3][3 + i] = p[i]; r.J[4][3 + i] = q[i]; }
    if (B2) for (int i = 0; i < 3; i++) { r.J[3][9 + i] = -p[i]; r.J[4][9 + i] = -q[i]; }
    real ax2[3], b[3];
    if (B2) ob_mul0_331(ax2, B2->R, j.axis2);
    else { ax2[0] = j.axis2[0]; ax2[1] = j.axis2[1]; ax2[2] = j.axis2[2]; }
    ob_cross(b, ax1, ax2);
    real k = fps * erp;
    r.c[3] = k * ob_dot(b, p);
    r.c[4] = k * ob_dot(b, q);
    real fm;
    if (ob_add_limot_rot(r, 5, j.limot1, ax1, B1, B2, fps, &fm) && fm != 0) {
      side[0][0] = fm; side[0][1] = ax1[0]; side[0][2] = ax1[1]; side[0][3] = ax1[2];
    }
  } else if (j.type == OB_JOINT_UNIVERSAL) {   // universal.cpp:296-361
    const real erp = *erp_io;
    ob_set_ball(r, j.anchor1, j.anchor2, B1, B2, fps, erp);
    real ax1[3], ax2[3], ax2t[3], p[3];
    ob_universal_axes(j, B1.R, B2 ? B2->R : (const real *)0, ax1, ax2);
    const real k = ob_dot(ax1, ax2);
    ax2t[0] = ax2[0] - k * ax1[0]; ax2t[1] = ax2[1] - k * ax1[1]; ax2t[2] = ax2[2] - k * ax1[2];
    ob_cross(p, ax1, ax2t);
    ob_safe_normalize3(p);
    for (int i = 0; i < 3; i++) r.J[3][3 + i] = p[i];
    if (B2) for (int i = 0; i < 3; i++) r.J[3][9 + i] = -p[i];
    r.c[3] = fps * erp * -k;
    real fm;
    const int added = ob_add_limot_rot(r, 4, j.limot1, ax1, B1, B2, fps, &fm);
    if (added && fm != 0) { side[0][0] = fm; side[0][1] = ax1[0]; side[0][2] = ax1[1]; side[0][3] = ax1[2]; }
    if (ob_add_limot_rot(r, 4 + added, j.limot2, ax2, B1, B2, fps, &fm) && fm != 0) {
      side[1][0] = fm; side[1][1] = ax2[0]; side[1][2] = ax2[1]; side[1][3] = ax2[2];
    }
  } else if (j.type == OB_JOINT_AMOTOR) {   // amotor.cpp:175-206
    real ax[3][3], c01[3], c12[3];
    ob_amotor_axes(j, B1, B2, ax);
    const real *axp[3] = {ax[0], ax[1], ax[2]};
    if (OB_JM_MODE(j.flags) == 1) {
      ob_cross(c01, ax[0], ax[1]); axp[2] = c01;
      ob_cross(c12, ax[1], ax[2]); axp[0] = c12;
    }
    const ObLimot *lm[3] = {&j.limot1, &j.limot2, &j.limot3};
    const int num = OB_JM_NUM(j.flags);
    int row = 0;
    for (int i = 0; i < num; i++) {
      real fm;
      const int added = ob_add_limot_rot(r, row, *lm[i], axp[i], B1, B2, fps, &fm);
      if (added && fm != 0) { side[i][0] = fm; side[i][1] = axp[i][0]; side[i][2] = axp[i][1]; side[i][3] = axp[i][2]; }
      row += added;
    }
  } else if (j.type == OB_JOINT_LMOTOR) {   // lmotor.cpp:89-99 (its limots never reach a limit state: no side effects)
    const real *axs[3] = {j.axis1, j.axis2, j.anchor1};
    const ObLimot *lm[3] = {&j.limot1, &j.limot2, &j.limot3};
    const int num = OB_JM_NUM(j.flags);
    int row = 0;
    for (int i = 0; i < num; i++) {
      real ax[3], fm, ltd[3];
      ob_motor_axis(j, i, axs[i], B1, B2, ax);
      row += ob_add_limot_lin(r, row, *lm[i], ax, B1, B2, fps, &fm, ltd);
    }
  } else if (j.type == OB_JOINT_FIXED) {   // fixed.cpp:57-100
    ob_set_fixed_orientation(r, 3, j.qrel, B1, B2, fps, *erp_io);   // uses the erp that was current BEFORE this joint
    r.J[0][0] = 1; r.J[1][1] = 1; r.J[2][2] = 1;
    *erp_io = j.erp;
    r.cfm[0] = j.cfm; r.cfm[1] = j.cfm; r.cfm[2] = j.cfm;
    real ofs[3];
    ob_mul0_331(ofs, B1.R, j.anchor1);
    if (B2) {
      // dSetCrossMatrixPlus(J1a, ofs)
      r.J[0][3 + 1] = -ofs[2]; r.J[0][3 + 2] = ofs[1];
      r.J[1][3 + 0] = ofs[2]; r.J[1][3 + 2] = -ofs[0];
      r.J[2][3 + 0] = -ofs[1]; r.J[2][3 + 1] = ofs[0];
      r.J[0][6] = -1; r.J[1][7] = -1; r.J[2][8] = -1;
    }
    const real k = fps * *erp_io;
    if (B2) { for (int i = 0; i < 3; i++) r.c[i] = k * (B2->pos[i] - B1.pos[i] + ofs[i]); }
    else { for (int i = 0; i < 3; i++) r.c[i] = k * (j.anchor1[i] - B1.pos[i]); }
  } else if (j.type == OB_JOINT_SLIDER) {   // slider.cpp:149-227
    const real erp = *erp_io;
    real c[3] = {0, 0, 0};
    if (B2) for (int i = 0; i < 3; i++) c[i] = B2->pos[i] - B1.pos[i];
    ob_set_fixed_orientation(r, 0, j.qrel, B1, B2, fps, erp);
    real ax1[3], p[3], q[3];
    ob_mul0_331(ax1, B1.R, j.axis1);
    ob_plane_space(ax1, p, q);
    if (B2) {
      real tmp[3];
      ob_cross(tmp, c, p);
      tmp[0] *= OB_REAL(0.5); tmp[1] *= OB_REAL(0.5); tmp[2] *= OB_REAL(0.5);
      for (int i = 0; i < 3; i++) { r.J[3][3 + i] = tmp[i]; r.J[3][9 + i] = tmp[i]; }
      ob_cross(tmp, c, q);
      tmp[0] *= OB_REAL(0.5); tmp[1] *= OB_REAL(0.5); tmp[2] *= OB_REAL(0.5);
      for (int i = 0; i < 3; i++) { r.J[4][3 + i] = tmp[i]; r.J[4][9 + i] = tmp[i]; }
      for (int i = 0; i < 3; i++) { r.J[3][6 + i] = -p[i]; r.J[4][6 + i] = -q[i]; }
    }
    for (int i = 0; i < 3; i++) { r.J[3][i] = p[i]; r.J[4][i] = q[i]; }
    const real k = fps * erp;
    if (B2) {
      real ofs[3];
      ob_mul0_331(ofs, B2->R, j.anchor1);
      for (int i = 0; i < 3; i++) c[i] += ofs[i];
      r.c[3] = k * ob_dot(p, c);
      r.c[4] = k * ob_dot(q, c);
    } else {
      real ofs[3];
      for (int i = 0; i < 3; i++) ofs[i] = j.anchor1[i] - B1.pos[i];
      r.c[3] = k * ob_dot(p, ofs);
      r.c[4] = k * ob_dot(q, ofs);
      if (j.flags & OB_JF_REVERSE) for (int i = 0; i < 3; i++) ax1[i] = -ax1[i];
    }
    real fm, ltd[3];
    if (ob_add_limot_lin(r, 5, j.limot1, ax1, B1, B2, fps, &fm, ltd) && fm != 0) {
      side[0][0] = fm; side[0][1] = ax1[0]; side[0][2] = ax1[1]; side[0][3] = ax1[2];
      side[1][0] = fm; side[1][1] = ltd[0]; side[1][2] = ltd[1]; side[1][3] = ltd[2];
    }
  } else if (j.type == OB_JOINT_PU) {   // pu.cpp:236-380
    const real k = fps * *erp_io;
    real axP[3], dist[3], wanchor2[3] = {0, 0, 0};
    ob_mul0_331(axP, B1.R, j.v2);
    if (B2) {
      ob_mul0_331(wanchor2, B2->R, j.anchor2);
      for (int i = 0; i < 3; i++) dist[i] = wanchor2[i] + B2->pos[i] - B1.pos[i];
    } else if (j.flags & OB_JF_REVERSE) {
      for (int i = 0; i < 3; i++) dist[i] = B1.pos[i] - j.anchor2[i];
    } else {
      for (int i = 0; i < 3; i++) dist[i] = j.anchor2[i] - B1.pos[i];
    }
    real ax1[3], ax2[3], q[3], p[3];
    ob_universal_axes(j, B1.R, B2 ? B2->R : (const real *)0, ax1, ax2);
    const real val = ob_dot(ax1, ax2);
    q[0] = ax2[0] - val * ax1[0]; q[1] = ax2[1] - val * ax1[1]; q[2] = ax2[2] - val * ax1[2];
    ob_cross(p, ax1, q);
    ob_safe_normalize3(p);
    for (int i = 0; i < 3; i++) r.J[0][3 + i] = p[i];
    if (B2) for (int i = 0; i < 3; i++) r.J[0][9 + i] = -p[i];
    r.c[0] = k * -val;
    ob_cross(q, ax1, axP);
    ob_cross(&r.J[1][3], dist, ax1);
    ob_cross(&r.J[2][3], dist, q);
    for (int i = 0; i < 3; i++) { r.J[1][i] = ax1[i]; r.J[2][i] = q[i]; }
    if (B2) {
      ob_cross(&r.J[1][9], ax1, wanchor2);
      ob_cross(&r.J[2][9], q, wanchor2);
      for (int i = 0; i < 3; i++) { r.J[1][6 + i] = -ax1[i]; r.J[2][6 + i] = -q[i]; }
    }
    real err[3];
    ob_mul0_331(err, B1.R, j.anchor1);
    for (int i = 0; i < 3; i++) err[i] = dist[i] - err[i];
    r.c[1] = k * ob_dot(ax1, err);
    r.c[2] = k * ob_dot(q, err);
    real fm, ltd[3];
    int row = 3;
    int added = ob_add_limot_rot(r, row, j.limot1, ax1, B1, B2, fps, &fm);
    if (added && fm != 0) { side[0][0] = fm; side[0][1] = ax1[0]; side[0][2] = ax1[1]; side[0][3] = ax1[2]; }
    row += added;
    added = ob_add_limot_rot(r, row, j.limot2, ax2, B1, B2, fps, &fm);
    if (added && fm != 0) { side[1][0] = fm; side[1][1] = ax2[0]; side[1][2] = ax2[1]; side[1][3] = ax2[2]; }
    row += added;
    if (!B2 && (j.flags & OB_JF_REVERSE)) { axP[0] = -axP[0]; axP[1] = -axP[1]; axP[2] = -axP[2]; }
    if (ob_add_limot_lin(r, row, j.limot3, axP, B1, B2, fps, &fm, ltd) && fm != 0) {
      side[2][0] = fm; side[2][1] = axP[0]; side[2][2] = axP[1]; side[2][3] = axP[2];
      side[3][0] = fm; side[3][1] = ltd[0]; side[3][2] = ltd[1]; side[3][3] = ltd[2];
    }
  } else if (j.type == OB_JOINT_PLANE2D) {   // plane2d.cpp:87-147 (body 1 against the static environment)
    r.J[0][2] = 1; r.J[1][3] = 1; r.J[2][4] = 1;
    const real eps = fps * *erp_io;
    r.c[0] = eps * -B1.pos[2];
    const real ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
    real fm, ltd[3];
    int row = 3;
    // the reference keeps the row numbers from getInfo1 (x, y, angle in this order, each only when its fmax > 0)
    if (j.limot1.fmax > 0) row += ob_add_limot_lin(r, row, j.limot1, ex, B1, B2, fps, &fm, ltd);
    if (j.limot2.fmax > 0) row += ob_add_limot_lin(r, row, j.limot2, ey, B1, B2, fps, &fm, ltd);
    if (j.limot3.fmax > 0) row += ob_add_limot_rot(r, row, j.limot3, ez, B1, B2, fps, &fm);
  } else if (j.type == OB_JOINT_PISTON || j.type == OB_JOINT_PR) {   // piston.cpp:220-421, pr.cpp:236-390
    const bool pr = j.type == OB_JOINT_PR;
    const real k = fps * *erp_io;
    real dist[3], lanchor2[3] = {0, 0, 0};
    if (B2) {
      ob_mul0_331(lanchor2, B2->R, j.anchor2);
      for (int i = 0; i < 3; i++) dist[i] = lanchor2[i] + B2->pos[i] - B1.pos[i];
    } else if (j.flags & OB_JF_REVERSE) {
      for (int i = 0; i < 3; i++) dist[i] = B1.pos[i] - j.anchor2[i];
    } else {
      for (int i = 0; i < 3; i++) dist[i] = j.anchor2[i] - B1.pos[i];
    }
    real ax1[3], axP[3], p[3], q[3], ax2[3], b[3];
    ob_mul0_331(ax1, B1.R, j.axis1);           // rotoide axis (piston: also the prismatic axis)
    if (pr) {
      ob_mul0_331(axP, B1.R, j.v1);            // prismatic axis
      ob_cross(q, ax1, axP);
      for (int i = 0; i < 3; i++) p[i] = axP[i];
    } else {
      for (int i = 0; i < 3; i++) axP[i] = ax1[i];
      ob_plane_space(ax1, p, q);
    }
    // rows 0, 1: no relative rotation about p and q
    for (int i = 0; i < 3; i++) { r.J[0][3 + i] = p[i]; r.J[1][3 + i] = q[i]; }
    if (B2) {
      for (int i = 0; i < 3; i++) { r.J[0][9 + i] = -p[i]; r.J[1][9 + i] = -q[i]; }
      ob_mul0_331(ax2, B2->R, j.axis2);
    } else { ax2[0] = j.axis2[0]; ax2[1] = j.axis2[1]; ax2[2] = j.axis2[2]; }
    ob_cross(b, ax1, ax2);
    r.c[0] = k * ob_dot(p, b);
    r.c[1] = k * ob_dot(q, b);
    // rows 2, 3: no relative translation across the prismatic axis (piston: p, q; PR: the rotoide axis and q)
    const real *u2 = pr ? ax1 : p;
    ob_cross(&r.J[2][3], dist, u2);
    ob_cross(&r.J[3][3], dist, q);
    for (int i = 0; i < 3; i++) { r.J[2][i] = u2[i]; r.J[3][i] = q[i]; }
    if (B2) {
      ob_cross(&r.J[2][9], pr ? ax2 : p, lanchor2);
      ob_cross(&r.J[3][9], q, lanchor2);
      for (int i = 0; i < 3; i++) { r.J[2][6 + i] = -u2[i]; r.J[3][6 + i] = -q[i]; }
    }
    real err[3];
    ob_mul0_331(err, B1.R, j.anchor1);
    for (int i = 0; i < 3; i++) err[i] = dist[i] - err[i];
    r.c[2] = k * ob_dot(u2, err);
    r.c[3] = k * ob_dot(q, err);
    real axm[3] = {axP[0], axP[1], axP[2]};
    if (!B2 && (j.flags & OB_JF_REVERSE)) { axm[0] = -axP[0]; axm[1] = -axP[1]; axm[2] = -axP[2]; }
    real fm, ltd[3];
    const int added = ob_add_limot_lin(r, 4, j.limot1, axm, B1, B2, fps, &fm, ltd);
    if (added && fm != 0) {
      side[0][0] = fm; side[0][1] = axm[0]; side[0][2] = axm[1]; side[0][3] = axm[2];
      side[1][0] = fm; side[1][1] = ltd[0]; side[1][2] = ltd[1]; side[1][3] = ltd[2];
    }
    if (ob_add_limot_rot(r, 4 + added, j.limot2, ax1, B1, B2, fps, &fm) && fm != 0) {
      side[2][0] = fm; side[2][1] = ax1[0]; side[2][2] = ax1[1]; side[2][3] = ax1[2];
    }
  } else if (j.type == OB_JOINT_HINGE2) {
    const real erp = *erp_io;
    real ax1[3], ax2[3], q[3];
    ob_mul0_331(ax1, B1.R, j.axis1);
    ob_mul0_331(ax2, B2->R, j.axis2);
    ob_cross(q, ax1, ax2);
    real s = ob_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    real c = ob_dot(ax1, ax2);
    ob_safe_normalize3(q);
    ob_set_ball2(r, j.anchor1, j.anchor2, ax1, j.susp_erp, B1, B2, fps, erp);
    for (int i = 0; i < 3; i++) r.J[3][3 + i] = q[i];
    if (B2) for (int i = 0; i < 3; i++) r.J[3][9 + i] = -q[i];
    real k = fps * erp;
    r.c[3] = k * (j.c0 * s - j.s0 * c);
    real fm;
    int added = ob_add_limot_rot(r, 4, j.limot1, ax1, B1, B2, fps, &fm);
    if (added && fm != 0) { side[0][0] = fm; side[0][1] = ax1[0]; side[0][2] = ax1[1]; side[0][3] = ax1[2]; }
    int row = 4 + added;
    if (ob_add_limot_rot(r, row, j.limot2, ax2, B1, B2, fps, &fm) && fm != 0) {
      side[1][0] = fm; side[1][1] = ax2[0]; side[1][2] = ax2[1]; side[1][3] = ax2[2];
    }
    r.cfm[0] = j.susp_cfm;
  }
}

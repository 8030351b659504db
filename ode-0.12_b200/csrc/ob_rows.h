// ob_rows.h — constraint-row assembly (dxJoint::getInfo1/getInfo2) per joint.
//
// Contact joint: ode/src/joints/contact.cpp:46-71 (getInfo1), :74-256 (getInfo2).
// The driver defaults applied before getInfo2 (J=0, c=0, cfm=global_cfm,
// lo=-inf, hi=+inf, findex=-1) are ode/src/quickstep.cpp:715-747.
#pragma once
#include "ob_types.h"

struct ObRowOut {      // one joint's rows, thread-local
  real J[6][12];
  real c[6], cfm[6], lo[6], hi[6];
  int findex[6];       // joint-local (-1 or row offset inside the joint)
};

OB_HD void ob_rows_defaults(ObRowOut &r, int m, real global_cfm) {
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
OB_HD void ob_contact_info2(ObRowOut &r, int the_m, const ObSurface &sf, const real *cpos, const real *normal_in,
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

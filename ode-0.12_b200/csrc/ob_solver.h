// ob_solver.h — per-element pieces of dxQuickStepper (ode/src/quickstep.cpp:592-1025),
// SOR_LCP (:342-584) and dxStepBody (ode/src/util.cpp:255-360).  Each function is the
// body of one reference loop, to be run by one thread per body / row; the kernels
// in ob_kernels.cu only decide who runs what and when.
#pragma once
#include "ob_types.h"

// quickstep.cpp:610-631 + :633-665: invI_world = R*(invI*R^T); gyroscopic torque; gravity
OB_HD void ob_body_preamble(const real *R, const real *I, const real *invI, const real *avel, uint32_t flags,
                            real mass, const real *gravity, real *invI_world /*12*/, real *facc, real *tacc) {
  real tmp[12];
  ob_mul2_333(tmp, invI, R);
  ob_mul0_333(invI_world, R, tmp);
  invI_world[3] = invI_world[7] = invI_world[11] = 0;
  if (flags & OB_BODY_GYROSCOPIC) {
    real Iw[12], t3[3], cr[3];
    ob_mul2_333(tmp, I, R);
    ob_mul0_333(Iw, R, tmp);
    ob_mul0_331(t3, Iw, avel);
    ob_cross(cr, avel, t3);
    tacc[0] = tacc[0] - cr[0]; tacc[1] = tacc[1] - cr[1]; tacc[2] = tacc[2] - cr[2];
  }
  if ((flags & OB_BODY_NO_GRAVITY) == 0) {
    if (gravity[0]) facc[0] += mass * gravity[0];
    if (gravity[1]) facc[1] += mass * gravity[1];
    if (gravity[2]) facc[2] += mass * gravity[2];
  }
}

// quickstep.cpp:840-846: tmp1 = [facc*invM + lvel/h, invI_w*tacc + avel/h]
OB_HD void ob_body_tmp1(const real *facc, const real *tacc, const real *lvel, const real *avel, real invMass,
                        const real *invI_world, real stepsize1, real *tmp1 /*6*/) {
  for (int j = 0; j < 3; j++) tmp1[j] = facc[j] * invMass + lvel[j] * stepsize1;
  ob_mul0_331(tmp1 + 3, invI_world, tacc);
  for (int k = 0; k < 3; k++) tmp1[3 + k] += avel[k] * stepsize1;
}

// Row finalisation, fusing quickstep.cpp:849-857 (rhs, cfm scaling), compute_invM_JT
// (:117-136) and the Ad pre-pass + J/b scaling of SOR_LCP (:370-402).
// In: J (unscaled), c, cfm.  Out: J scaled by Ad, iMJ, b = rhs*Ad, Adcfm = Ad*cfm.
OB_HD void ob_row_finalize(real *J /*12*/, real c, real cfm, int b2 /* -1: none */, const real *tmp1_b1,
                           const real *tmp1_b2, real invM1, const real *invI1, real invM2, const real *invI2,
                           real stepsize1, real sor_w, real *iMJ /*12*/, real *b_out, real *Adcfm_out) {
  // multiply_J (:163-181)
  real sum = 0;
  for (int j = 0; j < 6; j++) sum += J[j] * tmp1_b1[j];
  if (b2 != -1) for (int j = 0; j < 6; j++) sum += J[6 + j] * tmp1_b2[j];
  real rhs = c * stepsize1 - sum;
  cfm *= stepsize1;
  // compute_invM_JT
  for (int j = 0; j < 3; j++) iMJ[j] = invM1 * J[j];
  ob_mul0_331(iMJ + 3, invI1, J + 3);
  if (b2 != -1) {
    for (int j = 0; j < 3; j++) iMJ[j + 6] = invM2 * J[j + 6];
    ob_mul0_331(iMJ + 9, invI2, J + 9);
  } else {
    for (int j = 6; j < 12; j++) iMJ[j] = 0;   // reference leaves these uninitialised and unused
  }
  // Ad
  real s2 = 0;
  for (int j = 0; j < 6; j++) s2 += iMJ[j] * J[j];
  if (b2 != -1) for (int k = 6; k < 12; k++) s2 += iMJ[k] * J[k];
  real Ad = sor_w / (s2 + cfm);
  for (int j = 0; j < 12; j++) J[j] *= Ad;
  *b_out = rhs * Ad;
  *Adcfm_out = Ad * cfm;
}

// Same as ob_row_finalize but leaves J unscaled and also returns Ad = w/(diag+cfm); the SOR
// kernel stores the unscaled J and re-applies `J *= Ad` (quickstep.cpp:393-401) on the fly.
OB_HD void ob_row_finalize2(const real *J /*12, unscaled*/, real c, real cfm, int b2 /* -1: none */, const real *tmp1_b1,
                            const real *tmp1_b2, real invM1, const real *invI1, real invM2, const real *invI2,
                            real stepsize1, real sor_w, real *iMJ /*12*/, real *b_out, real *Adcfm_out, real *Ad_out) {
  real sum = 0;
  for (int j = 0; j < 6; j++) sum += J[j] * tmp1_b1[j];
  if (b2 != -1) for (int j = 0; j < 6; j++) sum += J[6 + j] * tmp1_b2[j];
  real rhs = c * stepsize1 - sum;
  cfm *= stepsize1;
  for (int j = 0; j < 3; j++) iMJ[j] = invM1 * J[j];
  ob_mul0_331(iMJ + 3, invI1, J + 3);
  if (b2 != -1) {
    for (int j = 0; j < 3; j++) iMJ[j + 6] = invM2 * J[j + 6];
    ob_mul0_331(iMJ + 9, invI2, J + 9);
  } else {
    for (int j = 6; j < 12; j++) iMJ[j] = 0;
  }
  real s2 = 0;
  for (int j = 0; j < 6; j++) s2 += iMJ[j] * J[j];
  if (b2 != -1) for (int k = 6; k < 12; k++) s2 += iMJ[k] * J[k];
  real Ad = sor_w / (s2 + cfm);
  *b_out = rhs * Ad;
  *Adcfm_out = Ad * cfm;
  *Ad_out = Ad;
}

// One SOR row update (quickstep.cpp:490-581).  fc1/fc2 point at the 6-vectors of the
// row's bodies (fc2 = 0 for one-body rows); lam_f = lambda[findex] (ignored if findex<0).
// Returns the new lambda.
OB_HD real ob_sor_row(const real *J, const real *iMJ, real b, real Adcfm, real lo, real hi, int findex, real lam_f,
                      real old_lambda, real *fc1, real *fc2) {
  real delta = b - old_lambda * Adcfm;
  delta -= fc1[0] * J[0] + fc1[1] * J[1] + fc1[2] * J[2] + fc1[3] * J[3] + fc1[4] * J[4] + fc1[5] * J[5];
  if (fc2) delta -= fc2[0] * J[6] + fc2[1] * J[7] + fc2[2] * J[8] + fc2[3] * J[9] + fc2[4] * J[10] + fc2[5] * J[11];
  real hi_act, lo_act;
  if (findex != -1) { hi_act = ob_fabs(hi * lam_f); lo_act = -hi_act; }
  else { hi_act = hi; lo_act = lo; }
  real new_lambda = old_lambda + delta;
  real out;
  if (new_lambda < lo_act) { delta = lo_act - old_lambda; out = lo_act; }
  else if (new_lambda > hi_act) { delta = hi_act - old_lambda; out = hi_act; }
  else out = new_lambda;
  fc1[0] += delta * iMJ[0]; fc1[1] += delta * iMJ[1]; fc1[2] += delta * iMJ[2];
  fc1[3] += delta * iMJ[3]; fc1[4] += delta * iMJ[4]; fc1[5] += delta * iMJ[5];
  if (fc2) {
    fc2[0] += delta * iMJ[6]; fc2[1] += delta * iMJ[7]; fc2[2] += delta * iMJ[8];
    fc2[3] += delta * iMJ[9]; fc2[4] += delta * iMJ[10]; fc2[5] += delta * iMJ[11];
  }
  return out;
}

// quickstep.cpp:905-916 and :960-975: v += h*cforce ; v += h*invM*fe
OB_HD void ob_body_velocity_update(real *lvel, real *avel, const real *cforce /*6 or null*/, real *facc, real *tacc,
                                   real invMass, const real *invI_world, real stepsize) {
  if (cforce) {
    for (int j = 0; j < 3; j++) {
      lvel[j] += stepsize * cforce[j];
      avel[j] += stepsize * cforce[3 + j];
    }
  }
  real k = stepsize * invMass;
  for (int j = 0; j < 3; j++) {
    lvel[j] += k * facc[j];
    tacc[j] *= stepsize;
  }
  real t[3];
  ob_mul0_331(t, invI_world, tacc);
  avel[0] = avel[0] + t[0]; avel[1] = avel[1] + t[1]; avel[2] = avel[2] + t[2];
}

// dInternalHandleAutoDisabling for one body that has joints (util.cpp:99-233).  samples == 1 is the instantaneous mode;
// samples > 1 averages the last `samples` velocity samples: buf = [samples][6] (lvel, avel) ring buffer of this body,
// ctl = {write index, buffer-full flag}; the body cannot go idle before the buffer has filled once.
// Returns true when the body was put to sleep (flags / velocities already updated).
OB_HD bool ob_auto_disable(ObBodyDyn &B, const ObBodyConst &C, real h, real *buf, int *ctl) {
  if ((B.flags & (OB_BODY_AUTO_DISABLE | OB_BODY_DISABLED)) != OB_BODY_AUTO_DISABLE) return false;
  const int ns = C.adis_samples;
  if (ns == 0) return false;
  int idle = 0;
  real al[3], aa[3];
  bool ready = true;
  if (ns > 1 && buf) {
    int cnt = ctl[0];
    if (cnt >= ns) { ctl[1] = 0; cnt = 0; }   // the reference's sanity reset (only reachable after the count was lowered)
    for (int k = 0; k < 3; k++) { buf[6 * cnt + k] = B.lvel[k]; buf[6 * cnt + 3 + k] = B.avel[k]; }
    cnt++;
    if (cnt >= ns) { cnt = 0; ctl[1] = 1; }
    ctl[0] = cnt;
    ready = ctl[1] != 0;
    if (ready) {
      for (int k = 0; k < 3; k++) { al[k] = buf[k]; aa[k] = buf[3 + k]; }
      for (int i = 1; i < ns; i++)
        for (int k = 0; k < 3; k++) { al[k] += buf[6 * i + k]; aa[k] += buf[6 * i + 3 + k]; }
      const real r1 = OB_REAL(1.0) / (real)ns;
      for (int k = 0; k < 3; k++) { al[k] *= r1; aa[k] *= r1; }
    }
  } else {
    for (int k = 0; k < 3; k++) { al[k] = B.lvel[k]; aa[k] = B.avel[k]; }
  }
  if (ready) {
    idle = 1;
    const real ls = ob_dot(al, al);
    if (ls > C.adis_lin_thr) idle = 0;
    else { const real as = ob_dot(aa, aa); if (as > C.adis_ang_thr) idle = 0; }
  }
  if (idle) { B.adis_stepsleft--; B.adis_timeleft -= h; }
  else { B.adis_stepsleft = C.adis_idle_steps; B.adis_timeleft = C.adis_idle_time; }
  if (B.adis_stepsleft <= 0 && B.adis_timeleft <= 0) {
    B.flags |= OB_BODY_DISABLED;
    for (int k = 0; k < 3; k++) { B.lvel[k] = 0; B.avel[k] = 0; }
    return true;
  }
  return false;
}

OB_HD real ob_sinc(real x) {
  if ((double)ob_fabs(x) < 1.0e-4) return OB_REAL(1.0) - x * x * OB_REAL(0.166666666666666666667);
#if defined(dSINGLE)
  return ob_sinf_glibc(x) / x;
#else
  return sin(x) / x;
#endif
}
OB_HD real ob_cos(real x) {
#if defined(dSINGLE)
  return ob_cosf_glibc(x);
#else
  return cos(x);
#endif
}

// dxStepBody (util.cpp:255-360) minus the geom notifications (handled by the caller)
OB_HD void ob_step_body(real *pos, real *q, real *R, real *lvel, real *avel, uint32_t flags, real h,
                        real max_angular_speed, const real *finite_rot_axis, real lin_scale, real ang_scale,
                        real lin_thr, real ang_thr) {
  if (flags & OB_BODY_MAX_ANG_SPEED) {
    const real aspeed = ob_dot(avel, avel);
    if (aspeed > max_angular_speed * max_angular_speed) {
      const real coef = max_angular_speed / ob_sqrt(aspeed);
      avel[0] *= coef; avel[1] *= coef; avel[2] *= coef;
    }
  }
  for (int j = 0; j < 3; j++) pos[j] += h * lvel[j];
  if (flags & OB_BODY_FINITE_ROT) {
    real irv[3], qr[4];
    if (flags & OB_BODY_FINITE_ROT_AXIS) {
      real frv[3];
      real k = ob_dot(finite_rot_axis, avel);
      frv[0] = finite_rot_axis[0] * k; frv[1] = finite_rot_axis[1] * k; frv[2] = finite_rot_axis[2] * k;
      irv[0] = avel[0] - frv[0]; irv[1] = avel[1] - frv[1]; irv[2] = avel[2] - frv[2];
      h *= OB_REAL(0.5);
      real theta = k * h;
      qr[0] = ob_cos(theta);
      real s = ob_sinc(theta) * h;
      qr[1] = frv[0] * s; qr[2] = frv[1] * s; qr[3] = frv[2] * s;
    } else {
      real wlen = ob_sqrt(avel[0] * avel[0] + avel[1] * avel[1] + avel[2] * avel[2]);
      h *= OB_REAL(0.5);
      real theta = wlen * h;
      qr[0] = ob_cos(theta);
      real s = ob_sinc(theta) * h;
      qr[1] = avel[0] * s; qr[2] = avel[1] * s; qr[3] = avel[2] * s;
    }
    real q2[4];
    ob_qmul0(q2, qr, q);
    for (int j = 0; j < 4; j++) q[j] = q2[j];
    if (flags & OB_BODY_FINITE_ROT_AXIS) {
      real dq[4];
      ob_DQfromW(dq, irv, q);
      for (int j = 0; j < 4; j++) q[j] += h * dq[j];
    }
  } else {
    real dq[4];
    ob_DQfromW(dq, avel, q);
    for (int j = 0; j < 4; j++) q[j] += h * dq[j];
  }
  ob_safe_normalize4(q);
  ob_RfromQ(R, q);
  if (flags & OB_BODY_LIN_DAMP) {
    const real lin_speed = ob_dot(lvel, lvel);
    if (lin_speed > lin_thr) {
      const real k = 1 - lin_scale;
      lvel[0] *= k; lvel[1] *= k; lvel[2] *= k;
    }
  }
  if (flags & OB_BODY_ANG_DAMP) {
    const real ang_speed = ob_dot(avel, avel);
    if (ang_speed > ang_thr) {
      const real k = 1 - ang_scale;
      avel[0] *= k; avel[1] *= k; avel[2] *= k;
    }
  }
}

// ob_host.cpp — host object model + the non-compute part of the ODE C API.
// See ob_host.h.  Reference behaviour cited per function group.
#include "ob_host.h"
#include "ob_batch.h"
#include "ob_collide.h"
#include "ob_trimesh_host.h"
#include "ob_solver.h"
#include <new>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

// ---------------------------------------------------------------------------------
// errors (ode/src/error.cpp:33-105): dError -> exit(1), dDebug -> abort(), both
// overridable.  ob_set_last_error is for the batched API, which never aborts.
static dErrorHandlerFn *g_handler = 0;
static std::string g_last_error;
extern "C" void dB200SetErrorHandler(dErrorHandlerFn *fn) { g_handler = fn; }
extern "C" const char *dB200LastError(void) { return g_last_error.c_str(); }
static void vmsg(const char *kind, int num, const char *fmt, va_list ap, char *buf, size_t n) {
  vsnprintf(buf, n, fmt, ap);
  (void)kind; (void)num;
}
void ob_error(int num, const char *fmt, ...) {
  char buf[1024]; va_list ap; va_start(ap, fmt); vmsg("Error", num, fmt, ap, buf, sizeof buf); va_end(ap);
  g_last_error = buf;
  if (g_handler) { g_handler(num, buf); return; }
  fprintf(stderr, "\nODE Error %d: %s\n", num, buf); fflush(stderr); exit(1);
}
void ob_debug(int num, const char *fmt, ...) {
  char buf[1024]; va_list ap; va_start(ap, fmt); vmsg("INTERNAL ERROR", num, fmt, ap, buf, sizeof buf); va_end(ap);
  g_last_error = buf;
  if (g_handler) { g_handler(num, buf); return; }
  fprintf(stderr, "\nODE INTERNAL ERROR %d: %s\n", num, buf); fflush(stderr); abort();
}
void ob_message(int num, const char *fmt, ...) {
  char buf[1024]; va_list ap; va_start(ap, fmt); vmsg("Message", num, fmt, ap, buf, sizeof buf); va_end(ap);
  fprintf(stderr, "\nODE Message %d: %s\n", num, buf); fflush(stderr);
}
void ob_set_last_error(const char *fmt, ...) {
  char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_last_error = buf;
}

// ---------------------------------------------------------------------------------
// init / RNG (ode/src/odeinit.cpp, ode/src/misc.cpp:31-117)
uint32_t ob_global_seed = 0;
extern "C" {
int dInitODE2(unsigned int) { return 1; }
void dInitODE(void) {}
void dCloseODE(void) {}
const char *dGetConfiguration(void) {
#if defined(dSINGLE)
  return "ODE ODE_EXT_no_debug ODE_EXT_b200_cuda ODE_single_precision";
#else
  return "ODE ODE_EXT_no_debug ODE_EXT_b200_cuda ODE_double_precision";
#endif
}
unsigned long dRand(void) { ob_global_seed = ob_lcg_next(ob_global_seed); return ob_global_seed; }
unsigned long dRandGetSeed(void) { return ob_global_seed; }
void dRandSetSeed(unsigned long s) { ob_global_seed = (uint32_t)s; }
int dRandInt(int n) { uint32_t r = (uint32_t)dRand(); return ob_randint_fold(r, (uint32_t)n); }
int dTestRand(void) {
  uint32_t old = ob_global_seed; int ret = 1;
  ob_global_seed = 0;
  if (dRand() != 0x3c6ef35f || dRand() != 0x47502932 || dRand() != 0xd1ccf6e9 || dRand() != 0xaaf95334 ||
      dRand() != 0x6252e503) ret = 0;
  ob_global_seed = old;
  return ret;
}

// ---------------------------------------------------------------------------------
// small math exports (ode/src/odemath.cpp, rotation.cpp, matrix.cpp:124-222)
int dSafeNormalize3(dVector3 a) { return ob_safe_normalize3(a); }
int dSafeNormalize4(dVector4 a) { return ob_safe_normalize4(a); }
void dNormalize3(dVector3 a) { ob_safe_normalize3(a); }
void dNormalize4(dVector4 a) { ob_safe_normalize4(a); }
void dPlaneSpace(const dVector3 n, dVector3 p, dVector3 q) { ob_plane_space(n, p, q); }
int dOrthogonalizeR(dMatrix3 m) {
  // odemath.cpp:187-209 (operates on rows)
  dReal n0 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  if (n0 != 1) ob_safe_normalize3(m);
  dReal proj = ob_dot(m, m + 4);
  if (proj != 0) { m[4] -= proj * m[0]; m[5] -= proj * m[1]; m[6] -= proj * m[2]; }
  dReal n1 = m[4] * m[4] + m[5] * m[5] + m[6] * m[6];
  if (n1 != 1) ob_safe_normalize3(m + 4);
  ob_cross(m + 8, m, m + 4);
  m[3] = m[7] = m[11] = 0;
  return 1;
}
void dRSetIdentity(dMatrix3 R) {
  for (int i = 0; i < 12; i++) R[i] = 0;
  R[0] = R[5] = R[10] = 1;
}
void dQSetIdentity(dQuaternion q) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
static dReal ob_sin(dReal x) {
#if defined(dSINGLE)
  return sinf(x);
#else
  return sin(x);
#endif
}
void dQFromAxisAndAngle(dQuaternion q, dReal ax, dReal ay, dReal az, dReal angle) {
  dReal l = ax * ax + ay * ay + az * az;
  if (l > OB_REAL(0.0)) {
    angle *= OB_REAL(0.5);
    q[0] = ob_cos(angle);
    l = ob_sin(angle) * ob_recipsqrt(l);
    q[1] = ax * l; q[2] = ay * l; q[3] = az * l;
  } else { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }
}
void dRfromQ(dMatrix3 R, const dQuaternion q) { ob_RfromQ(R, q); }
void dRFromAxisAndAngle(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal angle) {
  dQuaternion q; dQFromAxisAndAngle(q, ax, ay, az, angle); ob_RfromQ(R, q);
}
void dRFromEulerAngles(dMatrix3 R, dReal phi, dReal theta, dReal psi) {
  dReal sphi = ob_sin(phi), cphi = ob_cos(phi), stheta = ob_sin(theta), ctheta = ob_cos(theta), spsi = ob_sin(psi),
        cpsi = ob_cos(psi);
  R[0] = cpsi * ctheta; R[1] = spsi * ctheta; R[2] = -stheta; R[3] = 0;
  R[4] = cpsi * stheta * sphi - spsi * cphi; R[5] = spsi * stheta * sphi + cpsi * cphi; R[6] = ctheta * sphi; R[7] = 0;
  R[8] = cpsi * stheta * cphi + spsi * sphi; R[9] = spsi * stheta * cphi - cpsi * sphi; R[10] = ctheta * cphi; R[11] = 0;
}
void dQMultiply0(dQuaternion qa, const dQuaternion qb, const dQuaternion qc) { ob_qmul0(qa, qb, qc); }
void dDQfromW(dReal dq[4], const dVector3 w, const dQuaternion q) { ob_DQfromW(dq, w, q); }
#define RR(i, j) R[(i) * 4 + (j)]
void dQfromR(dQuaternion q, const dMatrix3 R) {
  // rotation.cpp:259-306
  dReal tr, s;
  tr = RR(0, 0) + RR(1, 1) + RR(2, 2);
  if (tr >= 0) {
    s = ob_sqrt(tr + 1);
    q[0] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[1] = (RR(2, 1) - RR(1, 2)) * s;
    q[2] = (RR(0, 2) - RR(2, 0)) * s;
    q[3] = (RR(1, 0) - RR(0, 1)) * s;
    return;
  }
  int c;
  if (RR(1, 1) > RR(0, 0)) c = (RR(2, 2) > RR(1, 1)) ? 2 : 1;
  else c = (RR(2, 2) > RR(0, 0)) ? 2 : 0;
  if (c == 0) {
    s = ob_sqrt((RR(0, 0) - (RR(1, 1) + RR(2, 2))) + 1);
    q[1] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[2] = (RR(0, 1) + RR(1, 0)) * s;
    q[3] = (RR(2, 0) + RR(0, 2)) * s;
    q[0] = (RR(2, 1) - RR(1, 2)) * s;
  } else if (c == 1) {
    s = ob_sqrt((RR(1, 1) - (RR(2, 2) + RR(0, 0))) + 1);
    q[2] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[3] = (RR(1, 2) + RR(2, 1)) * s;
    q[1] = (RR(0, 1) + RR(1, 0)) * s;
    q[0] = (RR(0, 2) - RR(2, 0)) * s;
  } else {
    s = ob_sqrt((RR(2, 2) - (RR(0, 0) + RR(1, 1))) + 1);
    q[3] = OB_REAL(0.5) * s;
    s = OB_REAL(0.5) * ob_recip(s);
    q[1] = (RR(2, 0) + RR(0, 2)) * s;
    q[2] = (RR(1, 2) + RR(2, 1)) * s;
    q[0] = (RR(1, 0) - RR(0, 1)) * s;
  }
}
#undef RR

// Cholesky inverse of a PD matrix stored with row stride dPAD(n) (matrix.cpp:124-222)
static int ob_pad(int n) { return (n > 1) ? (((n - 1) | 3) + 1) : n; }
static int factor_cholesky(dReal *A, int n) {
  const int nskip = ob_pad(n);
  dReal recip[16];
  dReal *aa = A;
  for (int i = 0; i < n; aa += nskip, ++i) {
    dReal *cc = aa;
    const dReal *bb = A;
    for (int j = 0; j < i; bb += nskip, ++cc, ++j) {
      dReal sum = *cc;
      const dReal *a = aa, *b = bb, *bend = bb + j;
      for (; b != bend; ++a, ++b) sum -= (*a) * (*b);
      *cc = sum * recip[j];
    }
    dReal sum = *cc;
    dReal *a = aa, *aend = aa + i;
    for (; a != aend; ++a) sum -= (*a) * (*a);
    if (sum <= OB_REAL(0.0)) return 0;
    dReal sumsqrt = ob_sqrt(sum);
    *cc = sumsqrt;
    recip[i] = ob_recip(sumsqrt);
  }
  return 1;
}
static void solve_cholesky(const dReal *L, dReal *b, int n) {
  const int nskip = ob_pad(n);
  dReal y[16];
  const dReal *ll = L;
  for (int i = 0; i < n; ll += nskip, ++i) {
    dReal sum = OB_REAL(0.0);
    for (int k = 0; k < i; ++k) sum += ll[k] * y[k];
    y[i] = (b[i] - sum) / ll[i];
  }
  ll = L + (n - 1) * (nskip + 1);
  for (int i = n - 1; i >= 0; ll -= nskip + 1, --i) {
    dReal sum = OB_REAL(0.0);
    const dReal *l = ll + nskip;
    for (int k = i + 1; k < n; l += nskip, ++k) sum += (*l) * b[k];
    b[i] = (y[i] - sum) / (*ll);
  }
}
int dInvertPDMatrix(const dReal *A, dReal *Ainv, int n) {
  if (n < 1 || n > 16) return 0;
  const int nskip = ob_pad(n);
  dReal L[16 * 16], X[16];
  memcpy(L, A, nskip * n * sizeof(dReal));
  if (!factor_cholesky(L, n)) return 0;
  for (int i = 0; i < nskip * n; i++) Ainv[i] = 0;
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) X[i] = 0;
    X[c] = OB_REAL(1.0);
    solve_cholesky(L, X, n);
    for (int i = 0; i < n; i++) Ainv[i * nskip + c] = X[i];
  }
  return 1;
}

// ---------------------------------------------------------------------------------
// mass (ode/src/mass.cpp)
#define MI(i, j) I[(i) * 4 + (j)]
int dMassCheck(const dMass *m) {
  if (m->mass <= 0) return 0;
  dReal tmp[12];
  memcpy(tmp, m->I, sizeof tmp);
  return factor_cholesky(tmp, 3);
}
void dMassSetZero(dMass *m) { memset(m, 0, sizeof(*m)); }
void dMassSetParameters(dMass *m, dReal themass, dReal cgx, dReal cgy, dReal cgz, dReal I11, dReal I22, dReal I33,
                        dReal I12, dReal I13, dReal I23) {
  dMassSetZero(m);
  m->mass = themass;
  m->c[0] = cgx; m->c[1] = cgy; m->c[2] = cgz;
  m->MI(0, 0) = I11; m->MI(1, 1) = I22; m->MI(2, 2) = I33;
  m->MI(0, 1) = I12; m->MI(0, 2) = I13; m->MI(1, 2) = I23;
  m->MI(1, 0) = I12; m->MI(2, 0) = I13; m->MI(2, 1) = I23;
}
void dMassSetSphereTotal(dMass *m, dReal total_mass, dReal radius) {
  dMassSetZero(m);
  m->mass = total_mass;
  dReal II = OB_REAL(0.4) * total_mass * radius * radius;
  m->MI(0, 0) = II; m->MI(1, 1) = II; m->MI(2, 2) = II;
}
void dMassSetSphere(dMass *m, dReal density, dReal radius) {
  dMassSetSphereTotal(m, (dReal)((OB_REAL(4.0) / OB_REAL(3.0)) * OB_PI * radius * radius * radius * density), radius);
}
void dMassSetCapsule(dMass *m, dReal density, int direction, dReal radius, dReal length) {
  OB_UASSERT(direction >= 1 && direction <= 3, "bad direction number");
  dMassSetZero(m);
  dReal M1 = (dReal)(OB_PI * radius * radius * length * density);
  dReal M2 = (dReal)((OB_REAL(4.0) / OB_REAL(3.0)) * OB_PI * radius * radius * radius * density);
  m->mass = M1 + M2;
  dReal Ia = M1 * (OB_REAL(0.25) * radius * radius + (OB_REAL(1.0) / OB_REAL(12.0)) * length * length) +
             M2 * (OB_REAL(0.4) * radius * radius + OB_REAL(0.375) * radius * length + OB_REAL(0.25) * length * length);
  dReal Ib = (M1 * OB_REAL(0.5) + M2 * OB_REAL(0.4)) * radius * radius;
  m->MI(0, 0) = Ia; m->MI(1, 1) = Ia; m->MI(2, 2) = Ia;
  m->MI(direction - 1, direction - 1) = Ib;
}
void dMassSetCylinderTotal(dMass *m, dReal total_mass, int direction, dReal radius, dReal length) {   // mass.cpp:180-198
  OB_UASSERT(direction >= 1 && direction <= 3, "bad direction number");
  dMassSetZero(m);
  const dReal r2 = radius * radius;
  m->mass = total_mass;
  const dReal I = total_mass * (OB_REAL(0.25) * r2 + (OB_REAL(1.0) / OB_REAL(12.0)) * length * length);
  m->MI(0, 0) = I; m->MI(1, 1) = I; m->MI(2, 2) = I;
  m->MI(direction - 1, direction - 1) = total_mass * OB_REAL(0.5) * r2;
}
void dMassSetCylinder(dMass *m, dReal density, int direction, dReal radius, dReal length) {
  dMassSetCylinderTotal(m, (dReal)(OB_PI * radius * radius * length * density), direction, radius, length);
}
void dMassAdjust(dMass *m, dReal newmass) {
  dReal scale = newmass / m->mass;
  m->mass = newmass;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m->MI(i, j) *= scale;
}
void dMassSetCapsuleTotal(dMass *m, dReal total_mass, int direction, dReal a, dReal b) {
  dMassSetCapsule(m, 1.0, direction, a, b);
  dMassAdjust(m, total_mass);
}
void dMassSetBoxTotal(dMass *m, dReal total_mass, dReal lx, dReal ly, dReal lz) {
  dMassSetZero(m);
  m->mass = total_mass;
  m->MI(0, 0) = total_mass / OB_REAL(12.0) * (ly * ly + lz * lz);
  m->MI(1, 1) = total_mass / OB_REAL(12.0) * (lx * lx + lz * lz);
  m->MI(2, 2) = total_mass / OB_REAL(12.0) * (lx * lx + ly * ly);
}
void dMassSetBox(dMass *m, dReal density, dReal lx, dReal ly, dReal lz) {
  dMassSetBoxTotal(m, lx * ly * lz * density, lx, ly, lz);
}
static void cross_matrix_plus(dReal *res, const dReal *a) {
  res[1] = -a[2]; res[2] = +a[1]; res[4] = +a[2]; res[6] = -a[0]; res[8] = -a[1]; res[9] = +a[0];
}
void dMassTranslate(dMass *m, dReal x, dReal y, dReal z) {
  dReal ahat[12] = {0}, chat[12] = {0}, t1[12], t2[12], a[3];
  cross_matrix_plus(chat, m->c);
  a[0] = x + m->c[0]; a[1] = y + m->c[1]; a[2] = z + m->c[2];
  cross_matrix_plus(ahat, a);
  ob_mul0_333(t1, ahat, ahat);
  ob_mul0_333(t2, chat, chat);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m->MI(i, j) += m->mass * (t2[i * 4 + j] - t1[i * 4 + j]);
  m->MI(1, 0) = m->MI(0, 1); m->MI(2, 0) = m->MI(0, 2); m->MI(2, 1) = m->MI(1, 2);
  m->c[0] += x; m->c[1] += y; m->c[2] += z;
}
void dMassRotate(dMass *m, const dMatrix3 R) {
  dReal t1[12], t2[3];
  ob_mul2_333(t1, m->I, R);
  ob_mul0_333(m->I, R, t1);
  m->MI(1, 0) = m->MI(0, 1); m->MI(2, 0) = m->MI(0, 2); m->MI(2, 1) = m->MI(1, 2);
  ob_mul0_331(t2, R, m->c);
  m->c[0] = t2[0]; m->c[1] = t2[1]; m->c[2] = t2[2];
}
void dMassAdd(dMass *a, const dMass *b) {
  dReal denom = ob_recip(a->mass + b->mass);
  for (int i = 0; i < 3; i++) a->c[i] = (a->c[i] * a->mass + b->c[i] * b->mass) * denom;
  a->mass += b->mass;
  for (int i = 0; i < 12; i++) a->I[i] += b->I[i];
}
#undef MI

// ---------------------------------------------------------------------------------
// world (ode/src/ode.cpp:1542-1582 defaults, :1598-1620 destroy)
dWorldID dWorldCreate(void) {
  dxWorld *w = new dxWorld;
  memset(w, 0, sizeof(*w));
  w->global_erp = OB_REAL(0.2);
#if defined(dSINGLE)
  w->global_cfm = 1e-5f;
#else
  w->global_cfm = 1e-10;
#endif
  w->adis.idle_steps = 10;
  w->adis.idle_time = 0;
  w->adis.average_samples = 1;
  w->adis.angular_average_threshold = OB_REAL(0.01) * OB_REAL(0.01);
  w->adis.linear_average_threshold = OB_REAL(0.01) * OB_REAL(0.01);
  w->qs_iterations = 20;
  w->qs_w = OB_REAL(1.3);
  w->contact_max_vel = OB_INF;
  w->contact_min_depth = 0;
  w->dampingp.linear_threshold = OB_REAL(0.01) * OB_REAL(0.01);
  w->dampingp.angular_threshold = OB_REAL(0.01) * OB_REAL(0.01);
  w->max_angular_speed = OB_INF;
  return w;
}
static void joint_unlink_bodies(dxJoint *j);
void dWorldDestroy(dWorldID w) {
  OB_AASSERT(w);
  if (w->bound_batch) { ob_batch_invalidate(w->bound_batch, w, 0); w->bound_batch = 0; }   // a user batch must never touch this world again
  ob_dropin_forget_world(w);
  dxBody *b = w->firstbody;
  while (b) { dxBody *nb = b->next; dBodyDestroy(b); b = nb; }
  dxJoint *j = w->firstjoint;
  while (j) {
    dxJoint *nj = j->next;
    if (j->flags & dJOINT_INGROUP) {   // deactivate, the group owns the storage (:1598-1606)
      j->world = 0; j->node[0].body = 0; j->node[0].next = 0; j->node[1].body = 0; j->node[1].next = 0;
    } else delete j;
    j = nj;
  }
  delete w;
}
void dWorldSetGravity(dWorldID w, dReal x, dReal y, dReal z) { w->gravity[0] = x; w->gravity[1] = y; w->gravity[2] = z; }
void dWorldGetGravity(dWorldID w, dVector3 g) { g[0] = w->gravity[0]; g[1] = w->gravity[1]; g[2] = w->gravity[2]; }
void dWorldSetERP(dWorldID w, dReal erp) { w->global_erp = erp; }
dReal dWorldGetERP(dWorldID w) { return w->global_erp; }
void dWorldSetCFM(dWorldID w, dReal cfm) { w->global_cfm = cfm; }
dReal dWorldGetCFM(dWorldID w) { return w->global_cfm; }
void dWorldSetQuickStepNumIterations(dWorldID w, int num) { w->qs_iterations = num; }
int dWorldGetQuickStepNumIterations(dWorldID w) { return w->qs_iterations; }
void dWorldSetQuickStepW(dWorldID w, dReal v) { w->qs_w = v; }
dReal dWorldGetQuickStepW(dWorldID w) { return w->qs_w; }
void dWorldSetContactMaxCorrectingVel(dWorldID w, dReal vel) { w->contact_max_vel = vel; }
dReal dWorldGetContactMaxCorrectingVel(dWorldID w) { return w->contact_max_vel; }
void dWorldSetContactSurfaceLayer(dWorldID w, dReal depth) { w->contact_min_depth = depth; }
dReal dWorldGetContactSurfaceLayer(dWorldID w) { return w->contact_min_depth; }
void dWorldSetAutoDisableFlag(dWorldID w, int d) { if (d) w->body_flags |= OB_BODY_AUTO_DISABLE; else w->body_flags &= ~OB_BODY_AUTO_DISABLE; }
int dWorldGetAutoDisableFlag(dWorldID w) { return w->body_flags & OB_BODY_AUTO_DISABLE; }
void dWorldSetAutoDisableLinearThreshold(dWorldID w, dReal v) { w->adis.linear_average_threshold = v * v; }
void dWorldSetAutoDisableAngularThreshold(dWorldID w, dReal v) { w->adis.angular_average_threshold = v * v; }
void dWorldSetAutoDisableAverageSamplesCount(dWorldID w, unsigned int n) { w->adis.average_samples = n; }
void dWorldSetAutoDisableSteps(dWorldID w, int steps) { w->adis.idle_steps = steps; }
void dWorldSetAutoDisableTime(dWorldID w, dReal time) { w->adis.idle_time = time; }
void dWorldSetLinearDampingThreshold(dWorldID w, dReal t) { w->dampingp.linear_threshold = t * t; }
void dWorldSetAngularDampingThreshold(dWorldID w, dReal t) { w->dampingp.angular_threshold = t * t; }
void dWorldSetLinearDamping(dWorldID w, dReal scale) {
  if (scale) w->body_flags |= OB_BODY_LIN_DAMP; else w->body_flags &= ~OB_BODY_LIN_DAMP;
  w->dampingp.linear_scale = scale;
}
void dWorldSetAngularDamping(dWorldID w, dReal scale) {
  if (scale) w->body_flags |= OB_BODY_ANG_DAMP; else w->body_flags &= ~OB_BODY_ANG_DAMP;
  w->dampingp.angular_scale = scale;
}
void dWorldSetDamping(dWorldID w, dReal ls, dReal as) { dWorldSetLinearDamping(w, ls); dWorldSetAngularDamping(w, as); }
void dWorldSetMaxAngularSpeed(dWorldID w, dReal max_speed) {
  if (max_speed < OB_INF) w->body_flags |= OB_BODY_MAX_ANG_SPEED; else w->body_flags &= ~OB_BODY_MAX_ANG_SPEED;
  w->max_angular_speed = max_speed;
}
int dWorldQuickStep(dWorldID w, dReal stepsize) {
  OB_UASSERT(w, "bad world argument");
  OB_UASSERT(stepsize > 0, "stepsize must be > 0");
  return ob_dropin_quickstep(w, stepsize);
}

// ---------------------------------------------------------------------------------
// body (ode/src/ode.cpp:248-1153)
static void body_geoms_moved(dxBody *b) { for (dxGeom *g = b->geom; g; g = g->body_next) ob_geom_moved(g); }
dBodyID dBodyCreate(dWorldID w) {
  OB_AASSERT(w);
  if (w->bound_batch) ob_batch_invalidate(w->bound_batch, 0, 0);
  dxBody *b = new dxBody;
  memset(b, 0, sizeof(*b));
  new (&b->average_buf) std::vector<dReal>();   // the one non-trivial member: constructed again after the memset
  b->world = w;
  dMassSetParameters(&b->mass, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0);
  b->invI[0] = 1; b->invI[5] = 1; b->invI[10] = 1;
  b->invMass = 1;
  b->q[0] = 1;
  dRSetIdentity(b->R);
  // push-front on the world's body list (addObjectToList, ode.cpp:62-68)
  b->next = w->firstbody; b->tome = &w->firstbody;
  if (w->firstbody) w->firstbody->tome = &b->next;
  w->firstbody = b;
  w->nb++;
  dBodySetAutoDisableDefaults(b);
  b->adis_stepsleft = b->adis.idle_steps;
  b->adis_timeleft = b->adis.idle_time;
  dBodySetDampingDefaults(b);
  b->flags |= w->body_flags & OB_BODY_MAX_ANG_SPEED;
  b->max_angular_speed = w->max_angular_speed;
  b->flags |= OB_BODY_GYROSCOPIC;
  b->batch_index = -1;
  return b;
}
void dBodyDestroy(dBodyID b) {
  OB_AASSERT(b);
  if (b->world && b->world->bound_batch) ob_batch_invalidate(b->world->bound_batch, 0, 0);   // the batch's body table names this body
  dxGeom *next_geom = 0;
  for (dxGeom *g = b->geom; g; g = next_geom) { next_geom = g->body_next; dGeomSetBody(g, 0); }
  dxJointNode *n = b->firstjoint;
  while (n) {
    n->joint->node[(n == n->joint->node)].body = 0;
    dxJointNode *next = n->next;
    n->next = 0;
    joint_unlink_bodies(n->joint);
    n = next;
  }
  if (b->next) b->next->tome = b->tome;
  *(b->tome) = b->next;
  b->world->nb--;
  delete b;
}
dWorldID dBodyGetWorld(dBodyID b) { return b->world; }
void dBodySetData(dBodyID b, void *data) { b->userdata = data; }
void *dBodyGetData(dBodyID b) { return b->userdata; }
void dBodySetPosition(dBodyID b, dReal x, dReal y, dReal z) {
  b->pos[0] = x; b->pos[1] = y; b->pos[2] = z;
  body_geoms_moved(b);
}
void dBodySetRotation(dBodyID b, const dMatrix3 R) {
  memcpy(b->R, R, sizeof(dMatrix3));
  dOrthogonalizeR(b->R);
  dQfromR(b->q, R);
  ob_safe_normalize4(b->q);
  body_geoms_moved(b);
}
void dBodySetQuaternion(dBodyID b, const dQuaternion q) {
  b->q[0] = q[0]; b->q[1] = q[1]; b->q[2] = q[2]; b->q[3] = q[3];
  ob_safe_normalize4(b->q);
  ob_RfromQ(b->R, b->q);
  body_geoms_moved(b);
}
void dBodySetLinearVel(dBodyID b, dReal x, dReal y, dReal z) { b->lvel[0] = x; b->lvel[1] = y; b->lvel[2] = z; }
void dBodySetAngularVel(dBodyID b, dReal x, dReal y, dReal z) { b->avel[0] = x; b->avel[1] = y; b->avel[2] = z; }
const dReal *dBodyGetPosition(dBodyID b) { return b->pos; }
const dReal *dBodyGetRotation(dBodyID b) { return b->R; }
const dReal *dBodyGetQuaternion(dBodyID b) { return b->q; }
const dReal *dBodyGetLinearVel(dBodyID b) { return b->lvel; }
const dReal *dBodyGetAngularVel(dBodyID b) { return b->avel; }
void dBodySetMass(dBodyID b, const dMass *mass) {
  OB_AASSERT(b && mass);
  memcpy(&b->mass, mass, sizeof(dMass));
  if (dInvertPDMatrix(b->mass.I, b->invI, 3) == 0) dRSetIdentity(b->invI);
  b->invMass = ob_recip(b->mass.mass);
}
void dBodyGetMass(dBodyID b, dMass *mass) { memcpy(mass, &b->mass, sizeof(dMass)); }
void dBodyAddForce(dBodyID b, dReal fx, dReal fy, dReal fz) { b->facc[0] += fx; b->facc[1] += fy; b->facc[2] += fz; }
void dBodyAddTorque(dBodyID b, dReal fx, dReal fy, dReal fz) { b->tacc[0] += fx; b->tacc[1] += fy; b->tacc[2] += fz; }
void dBodyGetRelPointVel(dBodyID b, dReal px, dReal py, dReal pz, dVector3 result) {   // ode.cpp: lvel + avel x (R * prel)
  dReal prel[4] = {px, py, pz, 0}, p[4], c[4];
  ob_mul0_331(p, b->R, prel);
  ob_cross(c, b->avel, p);
  result[0] = b->lvel[0]; result[1] = b->lvel[1]; result[2] = b->lvel[2];
  result[0] += c[0]; result[1] += c[1]; result[2] += c[2];
}
void dBodyAddRelForce(dBodyID b, dReal fx, dReal fy, dReal fz) {
  dReal t1[3] = {fx, fy, fz}, t2[3];
  ob_mul0_331(t2, b->R, t1);
  b->facc[0] += t2[0]; b->facc[1] += t2[1]; b->facc[2] += t2[2];
}
void dBodyAddRelTorque(dBodyID b, dReal fx, dReal fy, dReal fz) {
  dReal t1[3] = {fx, fy, fz}, t2[3];
  ob_mul0_331(t2, b->R, t1);
  b->tacc[0] += t2[0]; b->tacc[1] += t2[1]; b->tacc[2] += t2[2];
}
void dBodyAddForceAtPos(dBodyID b, dReal fx, dReal fy, dReal fz, dReal px, dReal py, dReal pz) {
  b->facc[0] += fx; b->facc[1] += fy; b->facc[2] += fz;
  dReal f[3] = {fx, fy, fz}, q[3] = {px - b->pos[0], py - b->pos[1], pz - b->pos[2]}, t[3];
  ob_cross(t, q, f);
  b->tacc[0] = b->tacc[0] + t[0]; b->tacc[1] = b->tacc[1] + t[1]; b->tacc[2] = b->tacc[2] + t[2];
}
const dReal *dBodyGetForce(dBodyID b) { return b->facc; }
const dReal *dBodyGetTorque(dBodyID b) { return b->tacc; }
void dBodySetForce(dBodyID b, dReal x, dReal y, dReal z) { b->facc[0] = x; b->facc[1] = y; b->facc[2] = z; }
void dBodySetTorque(dBodyID b, dReal x, dReal y, dReal z) { b->tacc[0] = x; b->tacc[1] = y; b->tacc[2] = z; }
void dBodyEnable(dBodyID b) {
  b->flags &= ~OB_BODY_DISABLED;
  b->adis_stepsleft = b->adis.idle_steps;
  b->adis_timeleft = b->adis.idle_time;
}
void dBodyDisable(dBodyID b) { b->flags |= OB_BODY_DISABLED; }
int dBodyIsEnabled(dBodyID b) { return ((b->flags & OB_BODY_DISABLED) == 0); }
void dBodySetGravityMode(dBodyID b, int mode) { if (mode) b->flags &= ~OB_BODY_NO_GRAVITY; else b->flags |= OB_BODY_NO_GRAVITY; }
int dBodyGetGravityMode(dBodyID b) { return ((b->flags & OB_BODY_NO_GRAVITY) == 0); }
void dBodySetFiniteRotationMode(dBodyID b, int mode) {
  b->flags &= ~(OB_BODY_FINITE_ROT | OB_BODY_FINITE_ROT_AXIS);
  if (mode) {
    b->flags |= OB_BODY_FINITE_ROT;
    if (b->finite_rot_axis[0] != 0 || b->finite_rot_axis[1] != 0 || b->finite_rot_axis[2] != 0)
      b->flags |= OB_BODY_FINITE_ROT_AXIS;
  }
}
void dBodySetFiniteRotationAxis(dBodyID b, dReal x, dReal y, dReal z) {
  b->finite_rot_axis[0] = x; b->finite_rot_axis[1] = y; b->finite_rot_axis[2] = z;
  if (x != 0 || y != 0 || z != 0) { ob_safe_normalize3(b->finite_rot_axis); b->flags |= OB_BODY_FINITE_ROT_AXIS; }
  else b->flags &= ~OB_BODY_FINITE_ROT_AXIS;
}
void dBodySetGyroscopicMode(dBodyID b, int enabled) { if (enabled) b->flags |= OB_BODY_GYROSCOPIC; else b->flags &= ~OB_BODY_GYROSCOPIC; }
int dBodyGetGyroscopicMode(dBodyID b) { return (b->flags & OB_BODY_GYROSCOPIC) != 0; }
void dBodySetAutoDisableAverageSamplesCount(dBodyID b, unsigned int n) {
  b->adis.average_samples = n; b->average_counter = 0; b->average_ready = 0;
  b->average_buf.assign(n > 1 ? (size_t)6 * n : 0, 0);   // buffers are reallocated and the averaging restarts (ode.cpp:1050-1075)
}
void dBodySetAutoDisableFlag(dBodyID b, int do_auto_disable) {
  if (!do_auto_disable) {
    b->flags &= ~OB_BODY_AUTO_DISABLE;
    b->flags &= ~OB_BODY_DISABLED;
    b->adis.idle_steps = b->world->adis.idle_steps;
    b->adis.idle_time = b->world->adis.idle_time;
    dBodySetAutoDisableAverageSamplesCount(b, b->world->adis.average_samples);
  } else b->flags |= OB_BODY_AUTO_DISABLE;
}
int dBodyGetAutoDisableFlag(dBodyID b) { return ((b->flags & OB_BODY_AUTO_DISABLE) != 0); }
void dBodySetAutoDisableDefaults(dBodyID b) {
  dxWorld *w = b->world;
  b->adis = w->adis;
  dBodySetAutoDisableFlag(b, w->body_flags & OB_BODY_AUTO_DISABLE);
  dBodySetAutoDisableAverageSamplesCount(b, w->adis.average_samples);   // (re)allocates the sample buffers, restarts the averaging (ode.cpp:1094-1101)
}
void dBodySetDampingDefaults(dBodyID b) {
  dxWorld *w = b->world;
  b->dampingp = w->dampingp;
  const unsigned mask = OB_BODY_LIN_DAMP | OB_BODY_ANG_DAMP;
  b->flags &= ~mask;
  b->flags |= w->body_flags & mask;
}
void dBodySetLinearDamping(dBodyID b, dReal scale) {
  if (scale) b->flags |= OB_BODY_LIN_DAMP; else b->flags &= ~OB_BODY_LIN_DAMP;
  b->dampingp.linear_scale = scale;
}
void dBodySetAngularDamping(dBodyID b, dReal scale) {
  if (scale) b->flags |= OB_BODY_ANG_DAMP; else b->flags &= ~OB_BODY_ANG_DAMP;
  b->dampingp.angular_scale = scale;
}
void dBodySetMaxAngularSpeed(dBodyID b, dReal max_speed) {
  if (max_speed < OB_INF) b->flags |= OB_BODY_MAX_ANG_SPEED; else b->flags &= ~OB_BODY_MAX_ANG_SPEED;
  b->max_angular_speed = max_speed;
}
int dBodyGetNumJoints(dBodyID b) { int c = 0; for (dxJointNode *n = b->firstjoint; n; n = n->next) c++; return c; }
dGeomID dBodyGetFirstGeom(dBodyID b) { return b->geom; }
dGeomID dBodyGetNextGeom(dGeomID g) { return g->body_next; }
void dBodyGetRelPointPos(dBodyID b, dReal px, dReal py, dReal pz, dVector3 result) {
  dReal prel[3] = {px, py, pz}, p[3];
  ob_mul0_331(p, b->R, prel);
  result[0] = p[0] + b->pos[0]; result[1] = p[1] + b->pos[1]; result[2] = p[2] + b->pos[2];
}
void dBodyVectorToWorld(dBodyID b, dReal px, dReal py, dReal pz, dVector3 result) {
  dReal p[3] = {px, py, pz};
  ob_mul0_331(result, b->R, p);
}

// ---------------------------------------------------------------------------------
// joints: creation, attach, groups (ode/src/ode.cpp:1162-1399, ode/src/joints/joint.cpp:42-71)
static void joint_unlink_bodies(dxJoint *j) {  // removeJointReferencesFromAttachedBodies, ode.cpp:86-107
  for (int i = 0; i < 2; i++) {
    dxBody *body = j->node[i].body;
    if (body) {
      dxJointNode *n = body->firstjoint, *last = 0;
      while (n) {
        if (n->joint == j) {
          if (last) last->next = n->next; else body->firstjoint = n->next;
          break;
        }
        last = n; n = n->next;
      }
    }
  }
  j->node[0].body = 0; j->node[0].next = 0; j->node[1].body = 0; j->node[1].next = 0;
}
static dxJoint *create_joint(dWorldID w, dJointGroupID group, int type) {
  OB_AASSERT(w);
  dxJoint *j = new dxJoint;
  memset(j, 0, sizeof(*j));
  j->world = w; j->type = type;
  j->next = w->firstjoint; j->tome = &w->firstjoint;
  if (w->firstjoint) w->firstjoint->tome = &j->next;
  w->firstjoint = j;
  w->nj++;
  j->node[0].joint = j; j->node[1].joint = j;
  j->flags = 0;
  if (group) { j->flags |= dJOINT_INGROUP; j->group = group; group->joints.push_back(j); }
  ob_joint_init_type(j);
  return j;
}
dJointID dJointCreateContact(dWorldID w, dJointGroupID group, const dContact *c) {
  OB_AASSERT(w && c);
  dxJoint *j = create_joint(w, group, dJointTypeContact);
  j->contact = *c;
  return j;
}
dJointID dJointCreateBall(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeBall); }
dJointID dJointCreateHinge(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeHinge); }
dJointID dJointCreateHinge2(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeHinge2); }
dJointID dJointCreateSlider(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeSlider); }
dJointID dJointCreateFixed(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeFixed); }
dJointID dJointCreateUniversal(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeUniversal); }
dJointID dJointCreateAMotor(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeAMotor); }
dJointID dJointCreateLMotor(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeLMotor); }
dJointID dJointCreateNull(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypeNull); }   // joints/null.cpp: no rows, but it joins its bodies into one island
dJointID dJointCreatePlane2D(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypePlane2D); }
dJointID dJointCreatePiston(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypePiston); }
dJointID dJointCreatePR(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypePR); }
dJointID dJointCreatePU(dWorldID w, dJointGroupID g) { return create_joint(w, g, dJointTypePU); }
static void joint_free(dxJoint *j) {
  if (j->world) {
    joint_unlink_bodies(j);
    if (j->next) j->next->tome = j->tome;
    *(j->tome) = j->next;
    j->world->nj--;
  }
  delete j;
}
void dJointDestroy(dJointID j) {
  OB_AASSERT(j);
  if (!(j->flags & dJOINT_INGROUP)) joint_free(j);
}
dJointGroupID dJointGroupCreate(int) { return new dxJointGroup; }
void dJointGroupEmpty(dJointGroupID group) {
  OB_AASSERT(group);
  for (int i = (int)group->joints.size() - 1; i >= 0; i--) joint_free(group->joints[i]);
  group->joints.clear();
}
void dJointGroupDestroy(dJointGroupID group) { dJointGroupEmpty(group); delete group; }
void dJointAttach(dJointID joint, dBodyID body1, dBodyID body2) {
  OB_UASSERT(joint, "bad joint argument");
  OB_UASSERT(body1 == 0 || body1 != body2, "can't have body1==body2");
  dxWorld *world = joint->world;
  OB_UASSERT((!body1 || body1->world == world) && (!body2 || body2->world == world),
             "joint and bodies must be in same world");
  OB_UASSERT(!((joint->flags & dJOINT_TWOBODIES) && ((body1 != 0) ^ (body2 != 0))),
             "joint can not be attached to just one body");
  if (joint->node[0].body || joint->node[1].body) joint_unlink_bodies(joint);
  if (body1 == 0) { body1 = body2; body2 = 0; joint->flags |= dJOINT_REVERSE; }
  else joint->flags &= (~dJOINT_REVERSE);
  joint->node[0].body = body1;
  joint->node[1].body = body2;
  if (body1) { joint->node[1].next = body1->firstjoint; body1->firstjoint = &joint->node[1]; }
  else joint->node[1].next = 0;
  if (body2) { joint->node[0].next = body2->firstjoint; body2->firstjoint = &joint->node[0]; }
  else joint->node[0].next = 0;
  if (body1 || body2) ob_joint_set_relative_values(joint);
}
void dJointEnable(dJointID j) { j->flags &= ~dJOINT_DISABLED; }
void dJointDisable(dJointID j) { j->flags |= dJOINT_DISABLED; }
int dJointIsEnabled(dJointID j) { return (j->flags & dJOINT_DISABLED) == 0; }
dJointType dJointGetType(dJointID j) { return (dJointType)j->type; }
dBodyID dJointGetBody(dJointID j, int index) {
  if (index == 0 || index == 1) {
    if (j->flags & dJOINT_REVERSE) return j->node[1 - index].body;
    return j->node[index].body;
  }
  return 0;
}
void dJointSetFeedback(dJointID j, dJointFeedback *f) { j->feedback = f; }
dJointFeedback *dJointGetFeedback(dJointID j) { return j->feedback; }
int dAreConnected(dBodyID b1, dBodyID b2) {
  for (dxJointNode *n = b1->firstjoint; n; n = n->next) if (n->body == b2) return 1;
  return 0;
}
int dAreConnectedExcluding(dBodyID b1, dBodyID b2, int joint_type) {
  for (dxJointNode *n = b1->firstjoint; n; n = n->next)
    if (n->joint->type != joint_type && n->body == b2) return 1;
  return 0;
}

// ---------------------------------------------------------------------------------
// geoms + spaces (ode/src/collision_kernel.cpp:343-760, collision_space.cpp:47-213)
#define CHECK_NOT_LOCKED(space) OB_UASSERT((space) == 0 || (space)->lock_count == 0, "invalid operation for locked space")

static void space_add(dxSpace *s, dxGeom *g);
static void space_remove(dxSpace *s, dxGeom *g);

}  // extern "C"
void ob_geom_moved(dxGeom *geom) {   // dGeomMoved, collision_space.cpp:47-75
  if (geom->offset_posr) geom->gflags |= GEOM_POSR_BAD;
  dxSpace *parent = geom->parent_space;
  while (parent && (geom->gflags & GEOM_DIRTY) == 0) {
    CHECK_NOT_LOCKED(parent);
    geom->gflags |= GEOM_DIRTY | GEOM_AABB_BAD;
    if (parent->type == dSweepAndPruneSpaceClass) {
      // dxSAPSpace::dirty (collision_sapspace.cpp:363-387): swap-remove from GeomList, append to DirtyList
      if (geom->sap_didx < 0) {
        std::vector<dxGeom *> &G = parent->sap_geoms;
        dxGeom *last = G.back();
        G[geom->sap_gidx] = last; last->sap_gidx = geom->sap_gidx;
        G.pop_back();
        geom->sap_gidx = -1; geom->sap_didx = (int)parent->sap_dirty.size();
        parent->sap_dirty.push_back(geom);
      }
    } else {
      // dxSpace::dirty: unlink, push-front
      if (geom->next) geom->next->tome = geom->tome;
      *geom->tome = geom->next;
      geom->next = parent->first; geom->tome = &parent->first;
      if (parent->first) parent->first->tome = &geom->next;
      parent->first = geom;
    }
    geom = parent;
    parent = parent->parent_space;
  }
  while (geom) {
    geom->gflags |= GEOM_DIRTY | GEOM_AABB_BAD;
    geom = geom->parent_space;
  }
}
void ob_geom_recompute_posr(dxGeom *g) {   // recomputePosr/computePosr, collision_kernel.cpp:454-465
  if (g->gflags & GEOM_POSR_BAD) {
    dxBody *b = g->body;
    ob_mul0_331(g->final_posr->pos, b->R, g->offset_posr->pos);
    g->final_posr->pos[0] += b->pos[0]; g->final_posr->pos[1] += b->pos[1]; g->final_posr->pos[2] += b->pos[2];
    ob_mul0_333(g->final_posr->R, b->R, g->offset_posr->R);
    g->final_posr->R[3] = g->final_posr->R[7] = g->final_posr->R[11] = 0;
    g->gflags &= ~GEOM_POSR_BAD;
  }
}
dxGeom *ob_geom_shape(dxGeom *g) { return g->type == dGeomTransformClass ? g->xf_obj : g; }
void ob_geom_final_pose(dxGeom *g, dxPosR *out) {
  memset(out, 0, sizeof *out);
  if (!(g->gflags & GEOM_PLACEABLE)) return;
  ob_geom_recompute_posr(g);
  if (g->type == dGeomTransformClass && g->xf_obj) {   // computeFinalTx, collision_transform.cpp:101-108
    const dxPosR *in = g->xf_obj->final_posr;
    ob_mul0_331(out->pos, g->final_posr->R, in->pos);
    out->pos[0] += g->final_posr->pos[0]; out->pos[1] += g->final_posr->pos[1]; out->pos[2] += g->final_posr->pos[2];
    ob_mul0_333(out->R, g->final_posr->R, in->R);
    out->R[3] = out->R[7] = out->R[11] = 0;
    return;
  }
  for (int i = 0; i < 3; i++) out->pos[i] = g->final_posr->pos[i];
  for (int i = 0; i < 12; i++) out->R[i] = g->final_posr->R[i];
}
void ob_space_clean(dxSpace *s) {
  // cleanGeoms (collision_space.cpp:405-417, collision_sapspace.cpp:394-423): AABBs are recomputed on
  // the device inside dSpaceCollide; here only the dirty bookkeeping is done, which is what ordering depends on.
  if (s->type == dSweepAndPruneSpaceClass) {
    for (size_t i = 0; i < s->sap_dirty.size(); i++) {
      dxGeom *g = s->sap_dirty[i];
      if (g->is_space) ob_space_clean((dxSpace *)g);
      g->gflags &= ~(GEOM_DIRTY | GEOM_AABB_BAD);
      g->sap_didx = -1; g->sap_gidx = (int)s->sap_geoms.size();
      s->sap_geoms.push_back(g);
    }
    s->sap_dirty.clear();
    return;
  }
  for (dxGeom *g = s->first; g && (g->gflags & GEOM_DIRTY); g = g->next) {
    if (g->is_space) ob_space_clean((dxSpace *)g);
    g->gflags &= ~(GEOM_DIRTY | GEOM_AABB_BAD);
  }
}
extern "C" {
static void geom_init(dxGeom *g, dSpaceID space, int is_placeable, int type) {
  g->type = type;
  g->gflags = GEOM_DIRTY | GEOM_AABB_BAD | GEOM_ENABLED;
  if (is_placeable) g->gflags |= GEOM_PLACEABLE;
  g->data = 0; g->body = 0; g->body_next = 0;
  memset(&g->own_posr, 0, sizeof(dxPosR));
  memset(&g->off_storage, 0, sizeof(dxPosR));
  if (is_placeable) { g->final_posr = &g->own_posr; dRSetIdentity(g->own_posr.R); } else g->final_posr = 0;
  g->offset_posr = 0;
  g->next = 0; g->tome = 0; g->parent_space = 0;
  for (int i = 0; i < 6; i++) g->aabb[i] = 0;
  g->category_bits = ~0ul; g->collide_bits = ~0ul;
  g->p[0] = g->p[1] = g->p[2] = g->p[3] = 0;
  g->batch_index = -1;
  g->tmdata = 0;
  g->xf_obj = 0; g->xf_cleanup = 0; g->xf_info = 0;
  g->sap_didx = g->sap_gidx = -1;
  g->is_space = false;
  if (space) dSpaceAdd(space, g);
}
}  // extern "C"
int ob_ray_flags(const dxGeom *g) {
  return ((g->gflags & RAY_FIRSTCONTACT) ? 1 : 0) | ((g->gflags & RAY_BACKFACECULL) ? 2 : 0) | ((g->gflags & RAY_CLOSEST_HIT) ? 4 : 0);
}
dxGeom *ob_geom_create(dxSpace *space, int is_placeable, int type) { dxGeom *g = new dxGeom; geom_init(g, space, is_placeable, type); return g; }
extern "C" {
static void geom_body_remove(dxGeom *g) {
  if (g->body) {
    dxGeom **last = &g->body->geom, *x = g->body->geom;
    while (x) {
      if (x == g) { *last = x->body_next; break; }
      last = &x->body_next; x = x->body_next;
    }
    g->body = 0; g->body_next = 0;
  }
}
static void zero_sized(dxGeom *g, bool z) { g->gflags = z ? (g->gflags | GEOM_ZERO_SIZED) : (g->gflags & ~GEOM_ZERO_SIZED); }

void dGeomDestroy(dGeomID g) {
  OB_AASSERT(g);
  if (g->is_space) {
    dxSpace *s = (dxSpace *)g;
    CHECK_NOT_LOCKED(s);
    if (s->bound_batch) { ob_batch_invalidate(s->bound_batch, 0, s); s->bound_batch = 0; }
    ob_dropin_forget_space(s);
    dxGeom *x, *n;
    for (x = s->first; x; x = n) {
      n = x->next;
      if (s->cleanup) dGeomDestroy(x); else space_remove(s, x);
    }
  }
  if (g->parent_space) dSpaceRemove(g->parent_space, g);
  geom_body_remove(g);
  if (g->type == dGeomTransformClass && g->xf_obj && g->xf_cleanup) dGeomDestroy(g->xf_obj);   // ~dxGeomTransform, collision_transform.cpp:75-78
  delete g;
}
void dGeomSetData(dGeomID g, void *data) { g->data = data; }
void *dGeomGetData(dGeomID g) { return g->data; }
void dGeomSetBody(dGeomID g, dBodyID b) {
  OB_AASSERT(g);
  OB_UASSERT(b == NULL || (g->gflags & GEOM_PLACEABLE), "geom must be placeable");
  CHECK_NOT_LOCKED(g->parent_space);
  if (b) {
    if (g->body != b) {
      g->offset_posr = 0;
      g->gflags &= ~GEOM_POSR_BAD;
      g->final_posr = (dxPosR *)b->pos;   // body pos[4] and R[12] are laid out like dxPosR
      geom_body_remove(g);
      g->body = b; g->body_next = b->geom; b->geom = g;
    }
    ob_geom_moved(g);
  } else {
    if (g->body) {
      if (g->offset_posr) { ob_geom_recompute_posr(g); g->offset_posr = 0; }
      else {
        memcpy(g->own_posr.pos, g->body->pos, sizeof(dVector3));
        memcpy(g->own_posr.R, g->body->R, sizeof(dMatrix3));
        g->final_posr = &g->own_posr;
      }
      geom_body_remove(g);
    }
  }
}
dBodyID dGeomGetBody(dGeomID g) { return g->body; }
void dGeomSetPosition(dGeomID g, dReal x, dReal y, dReal z) {
  OB_UASSERT(g->gflags & GEOM_PLACEABLE, "geom must be placeable");
  CHECK_NOT_LOCKED(g->parent_space);
  if (g->offset_posr) {
    dReal wo[3];
    ob_mul0_331(wo, g->body->R, g->offset_posr->pos);
    dBodySetPosition(g->body, x - wo[0], y - wo[1], z - wo[2]);
  } else if (g->body) dBodySetPosition(g->body, x, y, z);
  else { g->final_posr->pos[0] = x; g->final_posr->pos[1] = y; g->final_posr->pos[2] = z; ob_geom_moved(g); }
}
void dGeomSetRotation(dGeomID g, const dMatrix3 R) {
  OB_UASSERT(g->gflags & GEOM_PLACEABLE, "geom must be placeable");
  CHECK_NOT_LOCKED(g->parent_space);
  if (g->offset_posr) { ob_debug(0, "dGeomSetRotation on an offset geom is not supported"); }
  else if (g->body) dBodySetRotation(g->body, R);
  else { memcpy(g->final_posr->R, R, sizeof(dMatrix3)); ob_geom_moved(g); }
}
void dGeomSetQuaternion(dGeomID g, const dQuaternion quat) {
  OB_UASSERT(g->gflags & GEOM_PLACEABLE, "geom must be placeable");
  CHECK_NOT_LOCKED(g->parent_space);
  if (g->offset_posr) { ob_debug(0, "dGeomSetQuaternion on an offset geom is not supported"); }
  if (g->body) dBodySetQuaternion(g->body, quat);
  else { ob_RfromQ(g->final_posr->R, quat); ob_geom_moved(g); }
}
const dReal *dGeomGetPosition(dGeomID g) {
  OB_UASSERT(g->gflags & GEOM_PLACEABLE, "geom must be placeable");
  ob_geom_recompute_posr(g);
  return g->final_posr->pos;
}
const dReal *dGeomGetRotation(dGeomID g) {
  OB_UASSERT(g->gflags & GEOM_PLACEABLE, "geom must be placeable");
  ob_geom_recompute_posr(g);
  return g->final_posr->R;
}
void dGeomGetQuaternion(dGeomID g, dQuaternion quat) {
  if (g->body && !g->offset_posr) { memcpy(quat, g->body->q, sizeof(dQuaternion)); }
  else { ob_geom_recompute_posr(g); dQfromR(quat, g->final_posr->R); }
}
int dGeomIsSpace(dGeomID g) { return g->is_space; }
dSpaceID dGeomGetSpace(dGeomID g) { return g->parent_space; }
int dGeomGetClass(dGeomID g) { return g->type; }
void dGeomSetCategoryBits(dGeomID g, unsigned long bits) { CHECK_NOT_LOCKED(g->parent_space); g->category_bits = bits; }
void dGeomSetCollideBits(dGeomID g, unsigned long bits) { CHECK_NOT_LOCKED(g->parent_space); g->collide_bits = bits; }
unsigned long dGeomGetCategoryBits(dGeomID g) { return g->category_bits; }
unsigned long dGeomGetCollideBits(dGeomID g) { return g->collide_bits; }
void dGeomEnable(dGeomID g) { g->gflags |= GEOM_ENABLED; }
void dGeomDisable(dGeomID g) { g->gflags &= ~GEOM_ENABLED; }
int dGeomIsEnabled(dGeomID g) { return (g->gflags & GEOM_ENABLED) != 0; }
static void geom_create_offset(dxGeom *g) {   // dGeomCreateOffset, collision_kernel.cpp:1010-1030
  OB_UASSERT(g->gflags & GEOM_PLACEABLE, "geom must be placeable");
  OB_UASSERT(g->body, "geom must be on a body");
  if (g->offset_posr) return;
  g->final_posr = &g->own_posr;
  g->offset_posr = &g->off_storage;
  memset(g->offset_posr->pos, 0, sizeof(dVector3));
  dRSetIdentity(g->offset_posr->R);
  g->gflags |= GEOM_POSR_BAD;
}
void dGeomSetOffsetPosition(dGeomID g, dReal x, dReal y, dReal z) {
  CHECK_NOT_LOCKED(g->parent_space);
  if (!g->offset_posr) geom_create_offset(g);
  g->offset_posr->pos[0] = x; g->offset_posr->pos[1] = y; g->offset_posr->pos[2] = z;
  ob_geom_moved(g);
}
void dGeomSetOffsetRotation(dGeomID g, const dMatrix3 R) {
  CHECK_NOT_LOCKED(g->parent_space);
  if (!g->offset_posr) geom_create_offset(g);
  memcpy(g->offset_posr->R, R, sizeof(dMatrix3));
  ob_geom_moved(g);
}
void dGeomSetOffsetQuaternion(dGeomID g, const dQuaternion quat) {
  CHECK_NOT_LOCKED(g->parent_space);
  if (!g->offset_posr) geom_create_offset(g);
  ob_RfromQ(g->offset_posr->R, quat);
  ob_geom_moved(g);
}
// offset given in WORLD coordinates (collision_kernel.cpp:1080-1134): the body-relative offset that puts the geom there
static void world_offset_posr(const dxBody *b, const dReal *wpos, const dReal *wR, dxPosR *off) {   // getWorldOffsetPosr :441-452
  dMatrix3 inv;
  memcpy(inv, b->R, sizeof(dMatrix3));
  dReal t;
  t = inv[0 + 4 * 1]; inv[0 + 4 * 1] = inv[1 + 4 * 0]; inv[1 + 4 * 0] = t;
  t = inv[2 + 4 * 0]; inv[2 + 4 * 0] = inv[0 + 4 * 2]; inv[0 + 4 * 2] = t;
  t = inv[1 + 4 * 2]; inv[1 + 4 * 2] = inv[2 + 4 * 1]; inv[2 + 4 * 1] = t;
  ob_mul0_333(off->R, inv, wR);
  const dReal wo[4] = {wpos[0] - b->pos[0], wpos[1] - b->pos[1], wpos[2] - b->pos[2], 0};
  ob_mul0_331(off->pos, inv, wo);
}
void dGeomSetOffsetWorldPosition(dGeomID g, dReal x, dReal y, dReal z) {
  CHECK_NOT_LOCKED(g->parent_space);
  if (!g->offset_posr) geom_create_offset(g);
  const dReal prel[4] = {x - g->body->pos[0], y - g->body->pos[1], z - g->body->pos[2], 0};   // dBodyGetPosRelPoint, ode.cpp:730-740
  ob_mul1_331(g->offset_posr->pos, g->body->R, prel);
  ob_geom_moved(g);
}
void dGeomSetOffsetWorldRotation(dGeomID g, const dMatrix3 R) {
  CHECK_NOT_LOCKED(g->parent_space);
  if (!g->offset_posr) geom_create_offset(g);
  ob_geom_recompute_posr(g);
  dVector3 wpos = {g->final_posr->pos[0], g->final_posr->pos[1], g->final_posr->pos[2], 0};
  world_offset_posr(g->body, wpos, R, g->offset_posr);
  ob_geom_moved(g);
}
void dGeomSetOffsetWorldQuaternion(dGeomID g, const dQuaternion quat) {
  CHECK_NOT_LOCKED(g->parent_space);
  if (!g->offset_posr) geom_create_offset(g);
  ob_geom_recompute_posr(g);
  dVector3 wpos = {g->final_posr->pos[0], g->final_posr->pos[1], g->final_posr->pos[2], 0};
  dMatrix3 wR;
  ob_RfromQ(wR, quat);
  world_offset_posr(g->body, wpos, wR, g->offset_posr);
  ob_geom_moved(g);
}
void dGeomClearOffset(dGeomID g) {
  if (g->offset_posr) {
    g->offset_posr = 0;
    g->final_posr = (dxPosR *)g->body->pos;
    g->gflags &= ~GEOM_POSR_BAD;
    ob_geom_moved(g);
  }
}
int dGeomIsOffset(dGeomID g) { return g->offset_posr != 0; }

// host-side AABB (only used by dGeomGetAABB; the hot path computes AABBs on the device)
static void geom_host_pose(dxGeom *g, ObPose *o) {
  dxGeom *sh = ob_geom_shape(g);
  o->type = sh->type;
  for (int i = 0; i < 4; i++) o->p[i] = sh->p[i];
  dxPosR f;
  ob_geom_final_pose(g, &f);
  for (int i = 0; i < 3; i++) o->pos[i] = f.pos[i];
  for (int i = 0; i < 12; i++) o->R[i] = f.R[i];
}
void dGeomGetAABB(dGeomID g, dReal aabb[6]) {
  if (g->is_space) {
    // dxSpace::computeAABB (collision_space.cpp:116-137): union of the members' boxes, all zero for an empty space
    dxSpace *sp = (dxSpace *)g;
    if (!sp->first) { for (int i = 0; i < 6; i++) aabb[i] = 0; return; }
    dReal a[6] = {OB_INF, -OB_INF, OB_INF, -OB_INF, OB_INF, -OB_INF};
    for (dxGeom *m = sp->first; m; m = m->next) {
      dReal b[6];
      dGeomGetAABB(m, b);
      for (int i = 0; i < 6; i += 2) if (b[i] < a[i]) a[i] = b[i];
      for (int i = 1; i < 6; i += 2) if (b[i] > a[i]) a[i] = b[i];
    }
    for (int i = 0; i < 6; i++) aabb[i] = a[i];
    return;
  }
  if (!ob_geom_shape(g)) { for (int i = 0; i < 6; i++) aabb[i] = 0; return; }   // empty transform, collision_transform.cpp:83-86
  ObPose o;
  geom_host_pose(g, &o);
  o.mesh = 0;
  if (g->type == dRayClass) o.mesh = ob_ray_flags(g);
  ObMeshDev md;
  memset(&md, 0, sizeof md);
  if (g->type == dTriMeshClass && g->tmdata) for (int k = 0; k < 3; k++) { md.aabbc[k] = g->tmdata->aabbc[k]; md.aabbe[k] = g->tmdata->aabbe[k]; }
  ob_aabb(o, aabb, &md);
}

dGeomID dCreateSphere(dSpaceID space, dReal radius) {
  dxGeom *g = new dxGeom; geom_init(g, space, 1, dSphereClass);
  g->p[0] = radius; zero_sized(g, !radius); return g;
}
void dGeomSphereSetRadius(dGeomID g, dReal radius) { g->p[0] = radius; zero_sized(g, !radius); ob_geom_moved(g); }
dReal dGeomSphereGetRadius(dGeomID g) { return g->p[0]; }
dGeomID dCreateBox(dSpaceID space, dReal lx, dReal ly, dReal lz) {
  dxGeom *g = new dxGeom; geom_init(g, space, 1, dBoxClass);
  g->p[0] = lx; g->p[1] = ly; g->p[2] = lz; zero_sized(g, !lx || !ly || !lz); return g;
}
void dGeomBoxSetLengths(dGeomID g, dReal lx, dReal ly, dReal lz) {
  g->p[0] = lx; g->p[1] = ly; g->p[2] = lz; zero_sized(g, !lx || !ly || !lz); ob_geom_moved(g);
}
void dGeomBoxGetLengths(dGeomID g, dVector3 result) { result[0] = g->p[0]; result[1] = g->p[1]; result[2] = g->p[2]; }
static void plane_normalize(dxGeom *g) {   // plane.cpp:49-66
  dReal l = g->p[0] * g->p[0] + g->p[1] * g->p[1] + g->p[2] * g->p[2];
  if (l > 0) { l = ob_recipsqrt(l); g->p[0] *= l; g->p[1] *= l; g->p[2] *= l; g->p[3] *= l; }
  else { g->p[0] = 1; g->p[1] = 0; g->p[2] = 0; g->p[3] = 0; }
}
dGeomID dCreatePlane(dSpaceID space, dReal a, dReal b, dReal c, dReal d) {
  dxGeom *g = new dxGeom; geom_init(g, space, 0, dPlaneClass);
  g->p[0] = a; g->p[1] = b; g->p[2] = c; g->p[3] = d; plane_normalize(g); return g;
}
void dGeomPlaneSetParams(dGeomID g, dReal a, dReal b, dReal c, dReal d) {
  g->p[0] = a; g->p[1] = b; g->p[2] = c; g->p[3] = d; plane_normalize(g); ob_geom_moved(g);
}
void dGeomPlaneGetParams(dGeomID g, dVector4 r) { r[0] = g->p[0]; r[1] = g->p[1]; r[2] = g->p[2]; r[3] = g->p[3]; }
dGeomID dCreateCapsule(dSpaceID space, dReal radius, dReal length) {
  dxGeom *g = new dxGeom; geom_init(g, space, 1, dCapsuleClass);
  g->p[0] = radius; g->p[1] = length; zero_sized(g, !radius); return g;
}
void dGeomCapsuleSetParams(dGeomID g, dReal radius, dReal length) {
  g->p[0] = radius; g->p[1] = length; zero_sized(g, !radius); ob_geom_moved(g);
}
void dGeomCapsuleGetParams(dGeomID g, dReal *radius, dReal *length) { *radius = g->p[0]; *length = g->p[1]; }
// flat-ended cylinder (ode/src/cylinder.cpp:49-101): p[0] = radius, p[1] = length along the local z axis
dGeomID dCreateCylinder(dSpaceID space, dReal radius, dReal length) {
  dxGeom *g = new dxGeom; geom_init(g, space, 1, dCylinderClass);
  g->p[0] = radius; g->p[1] = length; zero_sized(g, !radius || !length); return g;
}
void dGeomCylinderSetParams(dGeomID g, dReal radius, dReal length) {
  g->p[0] = radius; g->p[1] = length; zero_sized(g, !radius || !length); ob_geom_moved(g);
}
void dGeomCylinderGetParams(dGeomID g, dReal *radius, dReal *length) { *radius = g->p[0]; *length = g->p[1]; }
// geom transforms (ode/src/collision_transform.cpp:59-250).  The encapsulated geom must be one of the primitive
// classes this path serves inside a transform; anything else is refused when the transform is marshalled.
dGeomID dCreateGeomTransform(dSpaceID space) { dxGeom *g = new dxGeom; geom_init(g, space, 1, dGeomTransformClass); return g; }
#define OB_XF_CHECK(g) OB_UASSERT((g) && (g)->type == dGeomTransformClass, "argument not a geom transform")
void dGeomTransformSetGeom(dGeomID g, dGeomID obj) {
  OB_XF_CHECK(g);
  if (g->xf_obj && g->xf_cleanup) dGeomDestroy(g->xf_obj);
  g->xf_obj = obj;
}
dGeomID dGeomTransformGetGeom(dGeomID g) { OB_XF_CHECK(g); return g->xf_obj; }
void dGeomTransformSetCleanup(dGeomID g, int mode) { OB_XF_CHECK(g); g->xf_cleanup = mode; }
int dGeomTransformGetCleanup(dGeomID g) { OB_XF_CHECK(g); return g->xf_cleanup; }
void dGeomTransformSetInfo(dGeomID g, int mode) { OB_XF_CHECK(g); g->xf_info = mode; }
int dGeomTransformGetInfo(dGeomID g) { OB_XF_CHECK(g); return g->xf_info; }
// rays (ode/src/ray.cpp:49-189): p[0] = length, direction = column 2 of the rotation; the three mode flags
// live in gflags like the reference's RAY_* bits (collision_kernel.h:79-81)
dGeomID dCreateRay(dSpaceID space, dReal length) {
  dxGeom *g = new dxGeom; geom_init(g, space, 1, dRayClass);
  g->p[0] = length; return g;
}
void dGeomRaySetLength(dGeomID g, dReal length) { g->p[0] = length; ob_geom_moved(g); }
dReal dGeomRayGetLength(dGeomID g) { return g->p[0]; }
void dGeomRaySet(dGeomID g, dReal px, dReal py, dReal pz, dReal dx, dReal dy, dReal dz) {
  ob_geom_recompute_posr(g);
  dReal *rot = g->final_posr->R, *pos = g->final_posr->pos;   // for a ray on a body without offset this IS the body's pose, as in the reference
  pos[0] = px; pos[1] = py; pos[2] = pz;
  dReal n[3] = {dx, dy, dz};
  ob_safe_normalize3(n);
  rot[0 * 4 + 2] = n[0]; rot[1 * 4 + 2] = n[1]; rot[2 * 4 + 2] = n[2];
  ob_geom_moved(g);
}
void dGeomRayGet(dGeomID g, dVector3 start, dVector3 dir) {
  ob_geom_recompute_posr(g);
  for (int k = 0; k < 3; k++) { start[k] = g->final_posr->pos[k]; dir[k] = g->final_posr->R[k * 4 + 2]; }
}
void dGeomRaySetParams(dGeomID g, int FirstContact, int BackfaceCull) {
  if (FirstContact) g->gflags |= RAY_FIRSTCONTACT; else g->gflags &= ~RAY_FIRSTCONTACT;
  if (BackfaceCull) g->gflags |= RAY_BACKFACECULL; else g->gflags &= ~RAY_BACKFACECULL;
}
void dGeomRayGetParams(dGeomID g, int *FirstContact, int *BackfaceCull) {
  *FirstContact = (g->gflags & RAY_FIRSTCONTACT) != 0; *BackfaceCull = (g->gflags & RAY_BACKFACECULL) != 0;
}
void dGeomRaySetClosestHit(dGeomID g, int closestHit) { if (closestHit) g->gflags |= RAY_CLOSEST_HIT; else g->gflags &= ~RAY_CLOSEST_HIT; }
int dGeomRayGetClosestHit(dGeomID g) { return (g->gflags & RAY_CLOSEST_HIT) != 0; }

// spaces
static dxSpace *space_create(dSpaceID parent, int type) {
  dxSpace *s = new dxSpace;
  geom_init(s, 0, 0, type);
  s->is_space = true;
  s->count = 0; s->first = 0; s->cleanup = 1; s->sublevel = 0; s->lock_count = 0; s->manual_cleanup = 0;
  s->minlevel = -3; s->maxlevel = 10; s->axisorder = 0; s->bound_batch = 0;
  if (parent) dSpaceAdd(parent, s);
  return s;
}
dSpaceID dSimpleSpaceCreate(dSpaceID space) { return space_create(space, dSimpleSpaceClass); }
dSpaceID dHashSpaceCreate(dSpaceID space) { return space_create(space, dHashSpaceClass); }
dSpaceID dSweepAndPruneSpaceCreate(dSpaceID space, int axisorder) {
  dxSpace *s = space_create(space, dSweepAndPruneSpaceClass);
  s->axisorder = axisorder;
  return s;
}
void dSpaceDestroy(dSpaceID s) { OB_UASSERT(s && s->is_space, "argument not a space"); dGeomDestroy(s); }
void dHashSpaceSetLevels(dSpaceID s, int minlevel, int maxlevel) {
  OB_UASSERT(s->type == dHashSpaceClass, "argument must be a hash space");
  OB_UASSERT(minlevel <= maxlevel, "Bad argument(s)");
  s->minlevel = minlevel; s->maxlevel = maxlevel;
}
void dHashSpaceGetLevels(dSpaceID s, int *minlevel, int *maxlevel) {
  if (minlevel) *minlevel = s->minlevel;
  if (maxlevel) *maxlevel = s->maxlevel;
}
void dSpaceSetCleanup(dSpaceID s, int mode) { s->cleanup = (mode != 0); }
int dSpaceGetCleanup(dSpaceID s) { return s->cleanup; }
void dSpaceSetSublevel(dSpaceID s, int sublevel) { s->sublevel = sublevel; }
int dSpaceGetSublevel(dSpaceID s) { return s->sublevel; }
static void space_add(dxSpace *s, dxGeom *g) {   // dxSpace::add, collision_space.cpp:162-182
  CHECK_NOT_LOCKED(s);
  if (s->bound_batch) ob_batch_invalidate(s->bound_batch, 0, 0);
  OB_UASSERT(g->parent_space == 0 && g->next == 0, "geom is already in a space");
  g->parent_space = s;
  g->next = s->first; g->tome = &s->first;
  if (s->first) s->first->tome = &g->next;
  s->first = g;
  s->count++;
  g->gflags |= GEOM_DIRTY | GEOM_AABB_BAD;
  if (s->type == dSweepAndPruneSpaceClass) {   // dxSAPSpace::add, collision_sapspace.cpp:303-321
    g->sap_didx = (int)s->sap_dirty.size(); g->sap_gidx = -1;
    s->sap_dirty.push_back(g);
  }
  ob_geom_moved(s);
}
static void space_remove(dxSpace *s, dxGeom *g) {   // dxSpace::remove, :185-206
  CHECK_NOT_LOCKED(s);
  if (s->bound_batch) ob_batch_invalidate(s->bound_batch, 0, (g->is_space && ((dxSpace *)g)->bound_batch == s->bound_batch) ? (dxSpace *)g : 0);
  OB_UASSERT(g->parent_space == s, "object is not in this space");
  if (g->next) g->next->tome = g->tome;
  *g->tome = g->next;
  s->count--;
  if (s->type == dSweepAndPruneSpaceClass) {   // dxSAPSpace::remove, :323-361 (swap with the last element)
    std::vector<dxGeom *> &L = g->sap_didx >= 0 ? s->sap_dirty : s->sap_geoms;
    const int idx = g->sap_didx >= 0 ? g->sap_didx : g->sap_gidx;
    dxGeom *last = L.back();
    L[idx] = last;
    if (g->sap_didx >= 0) last->sap_didx = idx; else last->sap_gidx = idx;
    L.pop_back();
    g->sap_didx = g->sap_gidx = -1;
  }
  g->next = 0; g->tome = 0; g->parent_space = 0;
  ob_geom_moved(s);
}
void dSpaceAdd(dSpaceID s, dGeomID g) { OB_UASSERT(s && s->is_space, "argument not a space"); space_add(s, g); }
void dSpaceRemove(dSpaceID s, dGeomID g) { OB_UASSERT(s && s->is_space, "argument not a space"); space_remove(s, g); }
int dSpaceQuery(dSpaceID s, dGeomID g) { return g->parent_space == s; }
void dSpaceClean(dSpaceID s) { ob_space_clean(s); }
int dSpaceGetNumGeoms(dSpaceID s) { return s->count; }
dGeomID dSpaceGetGeom(dSpaceID s, int i) {
  OB_UASSERT(i >= 0 && i < s->count, "index out of range");
  if (s->type == dSweepAndPruneSpaceClass) {   // dxSAPSpace::getGeom, collision_sapspace.cpp:292-301
    const int nd = (int)s->sap_dirty.size();
    return i < nd ? s->sap_dirty[i] : s->sap_geoms[i - nd];
  }
  dxGeom *g = s->first;
  for (int j = 0; j < i; j++) g = g ? g->next : 0;
  return g;
}
void dSpaceCollide(dSpaceID space, void *data, dNearCallback *callback) {
  OB_AASSERT(space && callback);
  OB_UASSERT(space->is_space, "argument not a space");
  ob_dropin_space_collide(space, data, callback);
}
void dSpaceCollide2(dGeomID g1, dGeomID g2, void *data, dNearCallback *callback) {
  OB_AASSERT(g1 && g2 && callback);
  ob_dropin_space_collide2(g1, g2, data, callback);
}
int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom *contact, int skip) {
  OB_AASSERT(o1 && o2 && contact);
  OB_UASSERT((flags & 0xffff) > 0, "no contacts requested");
  if ((flags & 0xffff) == 0) return 0;
  if (o1 == o2) return 0;
  if (o1->body == o2->body && o1->body) return 0;
  return ob_dropin_collide(o1, o2, flags, contact, skip);
}
}  // extern "C"

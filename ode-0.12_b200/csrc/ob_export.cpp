// ob_export.cpp — dWorldExportDIF: the reference's "Dynamics Interchange Format v0.1" text dump of a world
// (ode/src/export-dif.cpp:560-624 defines the format: a Lua-like table per world, body and joint; include/ode/export-dif.h).
// Host-side only: it prints the host object model, which the drop-in path keeps current after every
// dWorldQuickStep and the batched path after dBatchDownload.  The output is byte-identical to the reference's for
// the object types both sides implement (tests: tests/test_export_dif.py compares the two drivers' files).
#include <stdio.h>
#include "ob_host.h"

namespace {
struct Dif {
  FILE *f;
  int prec;     // significant digits: 7 (dSINGLE) / 15 (dDOUBLE), export-dif.cpp:565-569
  int depth;    // tabs in front of a line
  void tabs() const { for (int i = 0; i < depth; i++) fputc('\t', f); }
  void num(dReal x) const {
    if (x == dInfinity) fputs("inf", f);
    else if (x == -dInfinity) fputs("-inf", f);
    else fprintf(f, "%.*g", prec, (double)x);
  }
  void line(const char *text) const { tabs(); fprintf(f, "%s\n", text); }
  void str(const char *name, const char *v) const { tabs(); fprintf(f, "%s = \"%s\",\n", name, v); }
  void i(const char *name, int v) const { tabs(); fprintf(f, "%s = %d,\n", name, v); }
  void r(const char *name, dReal v) const { tabs(); fprintf(f, "%s = ", name); num(v); fputs(",\n", f); }
  void vec(const char *name, const dReal *v, int n = 3) const {
    tabs(); fprintf(f, "%s = {", name);
    for (int k = 0; k < n; k++) { num(v[k]); if (k < n - 1) fputc(',', f); }
    fputs("},\n", f);
  }
  void r_nz(const char *name, dReal v) const { if (v != 0) r(name, v); }
  // sic: a vector is only written when ALL three components are non-zero (export-dif.cpp:125-128)
  void vec_nz(const char *name, const dReal *v) const { if (v[0] != 0 && v[1] != 0 && v[2] != 0) vec(name, v); }
  void open(const char *head) { line(head); depth++; }
  void close() { depth--; line("},"); }
};

// limit + motor tables of one dxJointLimitMotor; idx < 0: unnumbered (export-dif.cpp:133-176)
void limot(Dif &d, const dxLimot &l, int idx) {
  char head[32];
  if (idx >= 0) snprintf(head, sizeof head, "limit%d = {", idx); else snprintf(head, sizeof head, "limit = {");
  d.open(head);
  d.r("low_stop", l.lostop);
  d.r("high_stop", l.histop);
  d.r_nz("bounce", l.bounce);
  d.open("ODE = {");
  d.r_nz("stop_erp", l.stop_erp);
  d.r_nz("stop_cfm", l.stop_cfm);
  d.close();
  d.close();
  if (idx >= 0) snprintf(head, sizeof head, "motor%d = {", idx); else snprintf(head, sizeof head, "motor = {");
  d.open(head);
  d.r_nz("vel", l.vel);
  d.r_nz("fmax", l.fmax);
  d.open("ODE = {");
  d.r_nz("fudge_factor", l.fudge_factor);
  d.r_nz("normal_cfm", l.normal_cfm);
  d.close();
  d.close();
}

const char *joint_name(int type) {
  switch (type) {
    case dJointTypeBall: return "ball";
    case dJointTypeHinge: return "hinge";
    case dJointTypeSlider: return "slider";
    case dJointTypeContact: return "contact";
    case dJointTypeUniversal: return "universal";
    case dJointTypeHinge2: return "ODE_hinge2";
    case dJointTypeFixed: return "fixed";
    case dJointTypeNull: return "null";
    case dJointTypeAMotor: return "ODE_angular_motor";
    case dJointTypeLMotor: return "ODE_linear_motor";
    case dJointTypePR: return "PR";
    case dJointTypePU: return "PU";
    case dJointTypePiston: return "piston";
    default: return "unknown";
  }
}

void joint_fields(Dif &d, dxJoint *j) {
  switch (j->type) {
    case dJointTypeBall:
      d.vec("anchor1", j->anchor1); d.vec("anchor2", j->anchor2);
      break;
    case dJointTypeHinge:
      d.vec("anchor1", j->anchor1); d.vec("anchor2", j->anchor2); d.vec("axis1", j->axis1); d.vec("axis2", j->axis2);
      d.vec("qrel", j->qrel, 4);
      limot(d, j->limot, -1);
      break;
    case dJointTypeSlider:
      d.vec("axis1", j->axis1); d.vec("qrel", j->qrel, 4); d.vec("offset", j->offset);
      limot(d, j->limot, -1);
      break;
    case dJointTypeContact: {
      const dContact &c = j->contact;
      const int mode = c.surface.mode;
      d.vec("pos", c.geom.pos); d.vec("normal", c.geom.normal); d.r("depth", c.geom.depth);
      d.r("mu", c.surface.mu);
      if (mode & dContactMu2) d.r("mu2", c.surface.mu2);
      if (mode & dContactBounce) { d.r("bounce", c.surface.bounce); d.r("bounce_vel", c.surface.bounce_vel); }
      if (mode & dContactSoftERP) d.r("soft_ERP", c.surface.soft_erp);
      if (mode & dContactSoftCFM) d.r("soft_CFM", c.surface.soft_cfm);
      if (mode & dContactMotion1) d.r("motion1", c.surface.motion1);
      if (mode & dContactMotion2) d.r("motion2", c.surface.motion2);
      if (mode & dContactSlip1) d.r("slip1", c.surface.slip1);
      if (mode & dContactSlip2) d.r("slip2", c.surface.slip2);
      int fa = 0;
      if (mode & dContactApprox1_1) fa |= 1;
      if (mode & dContactApprox1_2) fa |= 2;
      if (fa) d.i("friction_approximation", fa);
      if (mode & dContactFDir1) d.vec("fdir1", c.fdir1);
    } break;
    case dJointTypeUniversal:
      d.vec("anchor1", j->anchor1); d.vec("anchor2", j->anchor2); d.vec("axis1", j->axis1); d.vec("axis2", j->axis2);
      d.vec("qrel1", j->qrel, 4); d.vec("qrel2", j->qrel2, 4);
      limot(d, j->limot, 1); limot(d, j->limot2, 2);
      break;
    case dJointTypeHinge2:
      d.vec("anchor1", j->anchor1); d.vec("anchor2", j->anchor2); d.vec("axis1", j->axis1); d.vec("axis2", j->axis2);
      d.vec("v1", j->v1); d.vec("v2", j->v2);
      d.r("susp_erp", j->susp_erp); d.r("susp_cfm", j->susp_cfm);
      limot(d, j->limot, 1); limot(d, j->limot2, 2);
      break;
    case dJointTypeFixed:
      d.vec("qrel", j->qrel);       // three of the four components, as the reference writes it
      d.vec("offset", j->offset);
      break;
    case dJointTypeAMotor:
    case dJointTypeLMotor:
      d.i("num", j->num);
      if (j->type == dJointTypeAMotor) d.i("mode", j->mode);
      d.tabs(); fprintf(d.f, "rel = {%d,%d,%d},\n", j->rel[0], j->rel[1], j->rel[2]);
      d.vec("axis1", j->axis1); d.vec("axis2", j->axis2); d.vec("axis3", j->axis3);
      limot(d, j->limot, 1); limot(d, j->limot2, 2); limot(d, j->limot3, 3);
      if (j->type == dJointTypeAMotor) { d.r("angle1", j->angle[0]); d.r("angle2", j->angle[1]); d.r("angle3", j->angle[2]); }
      break;
    case dJointTypePR:      // host fields: axis1 / axis2 = axisR1 / axisR2, axis3 = axisP1; limot = prismatic, limot2 = rotoide
      d.vec("anchor2", j->anchor2); d.vec("axisR1", j->axis1); d.vec("axisR2", j->axis2); d.vec("axisP1", j->axis3);
      d.vec("qrel", j->qrel, 4); d.vec("offset", j->offset);
      limot(d, j->limot, 1); limot(d, j->limot2, 2);
      break;
    case dJointTypePU:      // limot / limot2 = universal axes, limot3 = prismatic
      d.vec("anchor1", j->anchor1); d.vec("anchor2", j->anchor2); d.vec("axis1", j->axis1); d.vec("axis2", j->axis2);
      d.vec("axisP", j->axis3);
      d.vec("qrel1", j->qrel, 4); d.vec("qrel2", j->qrel2, 4);
      limot(d, j->limot, 1); limot(d, j->limot2, 2); limot(d, j->limot3, 3);
      break;
    case dJointTypePiston:
      d.vec("anchor1", j->anchor1); d.vec("anchor2", j->anchor2); d.vec("axis1", j->axis1); d.vec("axis2", j->axis2);
      d.vec("qrel", j->qrel, 4);
      limot(d, j->limot, 1); limot(d, j->limot2, 2);
      break;
    default:
      d.line("unknown joint");
  }
}

void geom_fields(Dif &d, dxGeom *g) {
  if (g->category_bits != (unsigned long)(~0)) { d.tabs(); fprintf(d.f, "category_bits = %lu\n", g->category_bits); }
  if (g->collide_bits != (unsigned long)(~0)) { d.tabs(); fprintf(d.f, "collide_bits = %lu\n", g->collide_bits); }
  if (!dGeomIsEnabled(g)) d.i("disabled", 1);
  switch (g->type) {
    case dSphereClass: d.str("type", "sphere"); d.r("radius", g->p[0]); break;
    case dBoxClass: d.str("type", "box"); d.vec("sides", g->p); break;
    case dCapsuleClass: d.str("type", "capsule"); d.r("radius", g->p[0]); d.r("length", g->p[1]); break;
    case dCylinderClass: d.str("type", "cylinder"); d.r("radius", g->p[0]); d.r("length", g->p[1]); break;
    case dPlaneClass: d.str("type", "plane"); d.vec("normal", g->p); d.r("d", g->p[3]); break;
    case dRayClass: d.str("type", "ray"); d.r("length", g->p[0]); break;
    case dTriMeshClass: d.str("type", "trimesh"); break;
    case dGeomTransformClass: {   // printGeomTransform, export-dif.cpp:429-443 (closing brace without a comma, as there)
      dxGeom *g2 = g->xf_obj;
      dQuaternion q;
      dGeomGetQuaternion(g2, q);
      d.str("type", "transform"); d.vec("pos", dGeomGetPosition(g2)); d.vec("q", q, 4);
      d.open("geometry = {"); geom_fields(d, g2); d.depth--; d.line("}");
      break;
    }
    default: break;
  }
}
}  // namespace

extern "C" void dWorldExportDIF(dWorldID w, FILE *file, const char *prefix) {
  Dif d;
  d.f = file;
  d.prec = sizeof(dReal) == sizeof(float) ? 7 : 15;
  d.depth = 1;
  fprintf(file, "-- Dynamics Interchange Format v0.1\n\n%sworld = dynamics.world {\n", prefix);
  d.vec("gravity", w->gravity);
  d.open("ODE = {");
  d.r("ERP", w->global_erp);
  d.r("CFM", w->global_cfm);
  d.open("auto_disable = {");
  d.r("linear_threshold", w->adis.linear_average_threshold);
  d.r("angular_threshold", w->adis.angular_average_threshold);
  d.i("average_samples", (int)w->adis.average_samples);
  d.r("idle_time", w->adis.idle_time);
  d.i("idle_steps", w->adis.idle_steps);
  fputs("\t\t},\n\t},\n}\n", file);
  d.depth = 0;
  // bodies, in world-list order (newest first); tag = index for the joints' references
  int num = 0;
  fprintf(file, "%sbody = {}\n", prefix);
  for (dxBody *b = w->firstbody; b; b = b->next, num++) {
    b->tag = num;
    fprintf(file, "%sbody[%d] = dynamics.body {\n\tworld = %sworld,\n", prefix, num, prefix);
    d.depth = 1;
    d.vec("pos", b->pos);
    d.vec("q", b->q, 4);
    d.vec("lvel", b->lvel);
    d.vec("avel", b->avel);
    d.r("mass", b->mass.mass);
    fputs("\tI = {{", file);
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) { d.num(b->mass.I[r * 4 + c]); if (c < 2) fputc(',', file); }
      if (r < 2) fputs("},{", file);
    }
    fputs("}},\n", file);
    d.vec_nz("com", b->mass.c);
    d.open("ODE = {");
    if (b->flags & OB_BODY_FINITE_ROT) d.i("finite_rotation", 1);
    if (b->flags & OB_BODY_DISABLED) d.i("disabled", 1);
    if (b->flags & OB_BODY_NO_GRAVITY) d.i("no_gravity", 1);
    if (b->flags & OB_BODY_AUTO_DISABLE) {
      d.open("auto_disable = {");
      d.r("linear_threshold", b->adis.linear_average_threshold);
      d.r("angular_threshold", b->adis.angular_average_threshold);
      d.i("average_samples", (int)b->adis.average_samples);
      d.r("idle_time", b->adis.idle_time);
      d.i("idle_steps", b->adis.idle_steps);
      d.r("time_left", b->adis_timeleft);
      d.i("steps_left", b->adis_stepsleft);
      d.close();
    }
    d.vec_nz("facc", b->facc);
    d.vec_nz("tacc", b->tacc);
    if (b->flags & OB_BODY_FINITE_ROT_AXIS) d.vec("finite_rotation_axis", b->finite_rot_axis);
    d.close();
    if (b->geom) {
      d.open("geometry = {");
      for (dxGeom *g = b->geom; g; g = g->body_next) {
        d.open("{");
        geom_fields(d, g);
        d.close();
      }
      d.close();
    }
    d.depth = 0;
    d.line("}");
  }
  // joints, in world-list order
  num = 0;
  fprintf(file, "%sjoint = {}\n", prefix);
  for (dxJoint *j = w->firstjoint; j; j = j->next, num++) {
    fprintf(file, "%sjoint[%d] = dynamics.%s_joint {\n\tworld = %sworld,\n\tbody = {", prefix, num, joint_name(j->type), prefix);
    if (j->node[0].body) fprintf(file, "%sbody[%d]", prefix, j->node[0].body->tag);
    if (j->node[1].body) fprintf(file, ",%sbody[%d]", prefix, j->node[1].body->tag);
    fputs("}\n", file);
    d.depth = 1;
    joint_fields(d, j);
    d.depth = 0;
    d.line("}");
  }
}

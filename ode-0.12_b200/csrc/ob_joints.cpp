// ob_joints.cpp — ball / hinge / hinge2 host-side bookkeeping (anchors, axes, params).
// Reference: ode/src/joints/{ball,hinge,hinge2,joint}.cpp.  (filled in incrementally)
#include "ob_host.h"
void ob_joint_init_type(dxJoint *j) { (void)j; }
void ob_joint_set_relative_values(dxJoint *j) { (void)j; }

// ob_collide_types.h — the per-pair narrowphase records shared by ob_collide.h and ob_trimesh.h
#pragma once
#include "ob_types.h"

struct ObPose {  // world pose + parameters of one geom, gathered per thread
  int type;
  int mesh;      // trimesh: index into the batch's mesh table
  real pos[3];
  real R[12];
  real p[4];
};

struct ObCg {  // contact being generated (dContactGeom minus the geom ids)
  real pos[3];
  real normal[3];
  real depth;
  int side1, side2;
};

#define OB_MAXC_LOCAL 16   // contacts kept per pair (box-box emits <= 8; trimesh pairs up to the caller's max_contacts)

// ob_ray.h — per-ray pieces of dBatchRayCast (ode.h), shared by the device kernel and the test-only host backend
#pragma once
#include "ob_broad.h"
#include "ob_collide.h"


// One ray of dBatchRayCast against one posed geom: the ray as dGeomRaySet builds it (ray.cpp:115-135), the collideAABBs filter, then dCollide(ray, geom, 1).  Returns 1 and fills *out when the geom is hit.
OB_HD void ob_ray_pose(const real *origin, const real *dir, real length, int ray_flags, ObPose *r) {
  r->type = OB_GEOM_RAY; r->mesh = ray_flags;
  r->p[0] = length; r->p[1] = r->p[2] = r->p[3] = 0;
  // dGeomRaySet normalises the direction (dNormalize3) and writes it into the third column of the ray's rotation; the other two
  // columns keep what a freshly created ray has (identity) -- no ray collider reads them
  real n[3] = {dir[0], dir[1], dir[2]};
  ob_safe_normalize3(n);
  for (int k = 0; k < 12; k++) r->R[k] = 0;
  r->R[0] = 1; r->R[5] = 1;
  for (int k = 0; k < 3; k++) { r->pos[k] = origin[k]; r->R[4 * k + 2] = n[k]; }
}
OB_HD int ob_ray_vs_geom(const ObPose &ray, const real *ray_aabb, uint32_t rcat, uint32_t rcol, const ObPose &g, int gbody, uint32_t gcat,
                         uint32_t gcol, const ObMeshDev *meshes, ObCg *out, int *bverr) {
  real a[6];
  ob_aabb(g, a, meshes);
  if (!ob_aabb_pair_filter(gbody, -1, gcat, gcol, rcat, rcol, a, ray_aabb)) return 0;
  ObCg c[4];
  int swapped;
  const int n = ob_collide_pair_t<true, 4>(ray, g, 1, c, &swapped, meshes, bverr);
  if (n > 0) *out = c[0];
  return n > 0 ? 1 : 0;
}

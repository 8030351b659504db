// ob_batch.h — the batch object shared by ob_batch.cpp (dBatch* entry points) and
// ob_dropin.cpp (classic per-call API served through a batch of one world).
#pragma once
#include <vector>
#include "ob_backend.h"
#include "ob_host.h"

struct dxBatch {
  ObBackend *bk;
  ObBatchDev caps;   // capacities; pointer members are execution-side pointers
  std::vector<dxWorld *> worlds;
  std::vector<dxSpace *> spaces;
  std::vector<std::vector<dxBody *> > bodies;  // [w][batch body index]
  std::vector<std::vector<dxGeom *> > geoms;   // [w][geom index]
  std::vector<std::vector<dxJoint *> > joints; // [w][permanent joint index]
  std::vector<int> nb, ng;
  std::vector<struct dxTriMeshData *> meshes; // distinct trimesh data objects referenced by the bound geoms (device table order)
  std::vector<uint32_t> seeds;                 // host mirror of the per-world LCG seeds as last set
  int debug_taps;
  int dropin;
  int invalid;   // a bound world / space was destroyed or changed structurally: the host object tables are stale, every entry point that would read them refuses
};
// called by the host object model (ob_host.cpp) when an object owned by a user batch is destroyed or its world / space changes
// structurally (body / geom added or removed): the batch is marked invalid and forgets the object pointers it must not touch again
void ob_batch_invalidate(dxBatch *B, dxWorld *gone_world, dxSpace *gone_space);

dxBatch *ob_batch_create(int nworlds, const dWorldID *worlds, const dSpaceID *spaces, const dBatchDesc *desc, int dropin);
int ob_batch_upload(dxBatch *B);
void ob_fill_surface(ObSurface &d, const dSurfaceParameters &s);
void ob_marshal_body(const dxBody *b, ObBodyDyn &d, ObBodyConst &c);
void ob_marshal_joint(const dxJoint *j, ObJoint &d);
void ob_marshal_geom(dxGeom *g, ObGeom &d);

// ob_trimesh_host.h — host-side dTriMeshData (shared by ob_trimesh_build.cpp, ob_batch.cpp, ob_dropin.cpp)
#pragma once
#include <vector>
#include "ob_host.h"
#include "ob_trimesh.h"

struct dxTriMeshData {
  std::vector<float> verts;      // [nverts*3] as OPCODE sees them (float)
  std::vector<int> tris;         // [ntris*3]
  std::vector<ObBvNode> nodes;   // [ntris-1] no-leaf tree, root = 0
  std::vector<unsigned char> useflags;   // [ntris] edge / vertex use flags (dGeomTriMeshDataPreprocess), empty = never preprocessed
  int nverts, ntris;
  dReal aabbc[3], aabbe[3];
  struct DevCopy { int device; ObMeshDev m; };
  std::vector<DevCopy> dev;      // uploaded copies, one per device
};
const ObMeshDev *ob_trimesh_device(dxTriMeshData *d, int device);

// ob_backend.h — the narrow seam between the generic batch layer (ob_batch.cpp:
// marshalling + C ABI) and whatever executes the step.  The product links
// ob_backend_cuda.cu (device memory + the sm_100a kernels).  tests/hostsim links
// a test-only backend that runs the same per-element device functions in plain
// loops on the CPU so their arithmetic can be checked against the reference
// without a GPU; it is never part of libode_b200_*.so.
#pragma once
#include <stddef.h>
#include "ob_types.h"
#include "ob_collide.h"

struct ObBackend;
// allocate every array named in `caps` (pointers in caps are ignored); returns 0 on failure
ObBackend *obk_create(const ObBatchDev &caps, int device, char *err, size_t errlen);
void obk_destroy(ObBackend *);
ObBatchDev *obk_arrays(ObBackend *);                       // pointers valid on the execution side
int obk_h2d(ObBackend *, void *dst, const void *src, size_t bytes);
int obk_d2h(ObBackend *, void *dst, const void *src, size_t bytes);
int obk_memset(ObBackend *, void *dst, int value, size_t bytes);
// nsteps x (collide + step) for all worlds; debug_taps: also fill fback
int obk_step(ObBackend *, real h, int nsteps, int debug_taps, char *err, size_t errlen);
// one pass of selected phases for all worlds (drop-in path): OBK_PHASE_COLLIDE = broadphase + narrowphase
// into pairs/contacts, OBK_PHASE_STEP = quickstep over whatever ncontacts/contacts hold
enum { OBK_PHASE_COLLIDE = 1, OBK_PHASE_STEP = 2 };
int obk_run_phases(ObBackend *, real h, int phases, int taps, char *err, size_t errlen);
// dCollide for one pair of posed geoms, outside any batch (one-thread kernel on the GPU).
// out must hold OB_MAXC_LOCAL contacts.  Returns the contact count or -1.
// meshes2: trimesh data of a / b (ObPose::mesh must be 0 / 1), or null when neither is a trimesh
int obk_collide_pair(const ObPose *a, const ObPose *b, int flags, ObCg *out, const ObMeshDev *meshes2, char *err, size_t errlen);
// copy one trimesh (vertices, triangles, tree) to the execution side; io->aabbc/aabbe are kept, pointers filled
// dSpaceCollide2 (geom x space): nq posed query geoms against the geoms of world slot 0 (already uploaded).
// qmesh[q] carries the model-space AABB when query q is a trimesh (pose.mesh must be 0).  qbody = body index in
// the bound world or -2 (a body of another world / none: -1).  hit[q*NG + g] = 1 when collideAABBs
// (collision_space_internal.h:48-82) would call the near callback for (geom g of the space, query q).
int obk_collide2(ObBackend *, const ObPose *q, const int *qbody, const uint32_t *qcat, const uint32_t *qcol, const ObMeshDev *qmesh,
                 int nq, unsigned char *hit, char *err, size_t errlen);
// dBatchRayCast: nrays rays per world against every world's geoms (host arrays in, host hits out); hit = {pos3, depth, normal3, geom index}
struct ObRayHit { real pos[3]; real depth; real normal[3]; int geom; };
int obk_raycast(ObBackend *, int nrays, const real *origin3, const real *dir3, const real *length, int ray_flags, uint32_t cat, uint32_t col,
                ObRayHit *hits, char *err, size_t errlen);
int obk_mesh_upload(const float *verts, int nverts, const int *tris, int ntris, const ObBvNode *nodes, const unsigned char *useflags /* [ntris] or null */, int device, ObMeshDev *io);
void obk_mesh_free(ObMeshDev *m);
int obk_sync(ObBackend *);
// bulk body-state I/O in API order ([world][creation-index body]); nbody[w] = bodies in world w.
// Host buffers; a null pointer skips that field.
int obk_get_state(ObBackend *, real *pos3, real *quat4, real *lvel3, real *avel3);
int obk_set_state(ObBackend *, const real *pos3, const real *quat4, const real *lvel3, const real *avel3);
int obk_add_forces(ObBackend *, const real *force3, const real *torque3);
// page-locked host memory for the bulk I/O calls (DMA without a staging copy)
void *obk_host_alloc(size_t bytes);
void obk_host_free(void *);
void *obk_stream(ObBackend *);
int obk_timer_start(ObBackend *);
int obk_timer_stop(ObBackend *, float *ms);
#define OBK_NKERNELS 5
void obk_set_kernel_timing(ObBackend *, int enable);
void obk_get_kernel_times(ObBackend *, double *ms, long long *launches);
const char *obk_kernel_name(int k);
long long obk_launch_count(void);
// ob_math.h libm restatements evaluated on the device (fn: 0 atan2f, 1 sinf, 2 cosf); -1 without a device
int obk_libm(int fn, int n, const float *a, const float *b, float *out);
// large-world path: last step's {pairs, contacts, contact pairs, solved contacts, colours, colouring rounds, SOR launches, steps timed}
// and the per-phase CUDA-event times accumulated while kernel timing is on
// {geoms+sort, pairs, narrowphase, colouring, assembly, SOR, integration}.  Returns -1 for a small-world batch.
int obk_large_stats(ObBackend *, int *ints8, double *ms8);
// large-world path over several GPUs of one box (one rank per GPU, every rank holds the whole world): the SOR
// phase is dealt over the ranks, fc is exchanged through peer mappings inside the kernel (ObLwSplit, ob_large.h).
// export: 128-byte description of this rank's fc + flag buffer; attach: all ranks' exports in rank order.
#define OBK_SPLIT_HANDLE_BYTES 128
int obk_split_export(ObBackend *, void *handle128, char *err, size_t errlen);
int obk_split_attach(ObBackend *, int rank, int nranks, const void *handles, char *err, size_t errlen);

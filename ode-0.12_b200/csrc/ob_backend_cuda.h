// ob_backend_cuda.h — internal to the sm_100a backend: the backend object and the seams between its
// translation units (ob_backend_cuda.cu: memory + orchestration; ob_kern_collide.cu: broadphase / narrowphase
// kernels; ob_kern_step.cu: quickstep kernels; ob_kern_large.cu: the grid-wide large-world path).  Split only so
// that the kernels compile in parallel; nothing here is visible outside the library.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <vector>
#include "ob_backend.h"
#include "ob_broad.h"
#include "ob_rows.h"
#include "ob_solver.h"
#include "ob_large.h"

#define OB_THREADS 128
#define OB_ROWF 19                                   // reals per compact row record (ob_step_kernel.cuh)
#define OB_ROWW 20                                   // + one slot holding the 32-bit meta word
extern long long g_launches;

// ------------------------------------------------------------------------------------
// shared-memory layouts (one function for host sizing and device carving)
struct CollideSmem {
  size_t pose, aabb, cb, gid, body, cat, col, en, hr, br, walk_of, sapkey, sapinit, sappos, sapwalk, key, o12, sorted, misc, grp, member, gstart, gcur, total;
};
__host__ __device__ inline size_t ob_al16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline CollideSmem collide_smem(int NG, int NP) {
  CollideSmem s; size_t o = 0;
  s.pose = o; o = ob_al16(o + sizeof(ObPose) * NG);
  s.aabb = o; o = ob_al16(o + sizeof(real) * 6 * NG);
  s.cb = o; o = ob_al16(o + sizeof(ObCellBox) * NG);
  s.gid = o; o = ob_al16(o + sizeof(int) * NG);
  s.body = o; o = ob_al16(o + sizeof(int) * NG);
  s.cat = o; o = ob_al16(o + sizeof(uint32_t) * NG);
  s.col = o; o = ob_al16(o + sizeof(uint32_t) * NG);
  s.en = o; o = ob_al16(o + sizeof(int) * NG);
  s.hr = o; o = ob_al16(o + sizeof(int) * NG);
  s.br = o; o = ob_al16(o + sizeof(int) * NG);
  s.walk_of = o; o = ob_al16(o + sizeof(int) * NG);
  s.sapkey = o; o = ob_al16(o + sizeof(float) * (NG + 1));
  s.sapinit = o; o = ob_al16(o + sizeof(int) * (NG + 1));
  s.sappos = o; o = ob_al16(o + sizeof(int) * (NG + 1));
  s.sapwalk = o; o = ob_al16(o + sizeof(int) * (NG + 1));
  { const size_t kb = sizeof(int) * 5 * NP, cb = sizeof(int) * 2 * OB_THREADS; s.key = o; o = ob_al16(o + (kb > cb ? kb : cb)); }   // key tails (ObKeyTail, 5 words); reused by the narrowphase for per-pair counts / offsets of a chunk
  s.o12 = o; o = ob_al16(o + sizeof(int2) * NP);
  s.sorted = o; o = ob_al16(o + sizeof(int2) * NP);
  s.misc = o; o = ob_al16(o + sizeof(int) * 64);
  s.grp = o; o = ob_al16(o + sizeof(unsigned short) * NP);          // group (stage, query rank) of a pair
  s.member = o; o = ob_al16(o + sizeof(unsigned short) * NP);       // pairs listed group by group
  s.gstart = o; o = ob_al16(o + sizeof(int) * (3 * (NG + 1) + 1));  // first member of a group
  s.gcur = o; o = ob_al16(o + sizeof(int) * (3 * (NG + 1) + 1));    // count, then fill cursor, then end of a group
  s.total = o;
  return s;
}

// exclusive scan of one int per thread across the CTA; *total = sum.  s_w: >= 33 ints of smem
__device__ inline int block_excl_scan(int v, int *s_w, int *total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int x = v;
  for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) s_w[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int t = lane < nw ? s_w[lane] : 0;
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
    s_w[lane] = t;   // inclusive warp totals
  }
  __syncthreads();
  int base = wid ? s_w[wid - 1] : 0;
  *total = s_w[nw - 1];
  __syncthreads();
  return base + x - v;
}

__device__ inline void geom_pose_dev(const ObGeom &g, const ObBodyDyn *bd, ObPose *o) {
  o->type = g.type;
  o->mesh = g.mesh;
  for (int k = 0; k < 4; k++) o->p[k] = g.p[k];
  if (g.body >= 0) {
    const ObBodyDyn &b = bd[g.body];
    if (g.flags & OB_GEOM_HAS_OFFSET) {
      ob_mul0_331(o->pos, b.R, g.pos);
      o->pos[0] += b.pos[0]; o->pos[1] += b.pos[1]; o->pos[2] += b.pos[2];
      ob_mul0_333(o->R, b.R, g.R);
      o->R[3] = o->R[7] = o->R[11] = 0;
    } else {
      for (int k = 0; k < 3; k++) o->pos[k] = b.pos[k];
      for (int k = 0; k < 12; k++) o->R[k] = b.R[k];
    }
  } else {
    for (int k = 0; k < 3; k++) o->pos[k] = g.pos[k];
    for (int k = 0; k < 12; k++) o->R[k] = g.R[k];
  }
}

// ------------------------------------------------------------------------------------
struct ObBackend {
  ObBatchDev d;
  int device;
  cudaStream_t stream;
  std::vector<void *> allocs;
  real *st_dev;      // packed state staging on the device: pos3|quat4|lvel3|avel3
  real *st_host;     // pinned
  size_t st_elems;   // W*NB
  size_t smem_collide, smem_prep, smem_sched, smem_sched_lane, smem_sor, smem_post, smem_collide_tile;
  int prep_tile;                      // tile width of k_prep (defaults to `tile`)
  int post_tile;                      // tile width of k_post
  int sor_lane; size_t smem_sor_lane;   // k_sor_lane: one lane per world (many tiny worlds)
  int prep_tile1; size_t smem_prep1;   // tile width / shared memory of k_prep's first half (OB_PREP_TILE1)
  int collide_split, narrow_grid[3];   // decoupled narrowphase (k_broad / k_narrow / k_contacts) and k_narrow's persistent grid per instantiation
  size_t smem_broad_tile;
  int collide_tile, tile_stage_cap;   // k_collide_tile serves the batch (worlds of <= 8 geoms)
  int sor_deep;     // 1: k_sor with the deep index prefetch (worlds with many rows)
  int sched_lane;   // 1: k_sched_lane (one lane per world) fits shared memory
  int sched_gs;     // > 0: k_sched_tile<sched_gs> (sched_gs lanes per world, serial chains on one lane per world)
  size_t smem_sched_tile;
  int sor_ring;     // 1: k_sor_ring (rows through a cp.async shared-memory ring, persistent L2-sized grid)
  size_t smem_sor_ring;
  int ring_resident;        // CTAs of k_sor_ring that fit the device at once
  int ring_depth;           // slots of the row ring (6 or 4; 0: ring kernel not in use)
  int prep_split;           // 1: k_prep in two launches, the schedule on `sstream` beside the second (batches of >= 256 worlds)
  cudaStream_t sstream;
  cudaEvent_t sev[2];
  int sor_reg;              // 1: k_sor_reg (pipelined pass, rows prefetched into registers)
  size_t smem_sor_reg;
  int sor_pair;             // > 0: k_sor_pair (two lanes per row) with this ring depth
  size_t smem_sor_pair;
  int pair_resident;
  double avg_rows;          // measured rows per world-step (counters read back at every sync point), 0 = unknown
  ObCounters *cnt_host;     // pinned copy of the device counters
  double l2_target_bytes;   // rows of the worlds in flight are kept below this (OB_SOR_L2MB)
  // independent worlds are stepped in nchunks chunks, each on its own stream: the chunks drift apart, so the
  // ALU-bound collide of one chunk overlaps the latency-bound SOR of another instead of running back to back
  int nchunks;
  cudaStream_t cstream[8];
  cudaEvent_t cev[9];
  int grid, grid_step, grid_sor, tile;
  cudaEvent_t ev[8];   // 0,1: user timer; 2..7 per-kernel timing
  int ktiming;
  double kms[OBK_NKERNELS];
  long long klaunch[OBK_NKERNELS];
  // one large world (ob_large.h)
  int large;
  ObLargeDev L;
  int *lw_host;        // pinned: scalars + segment table read back for launch sizing
  int lw_rounds, lw_ncol, lw_stat[8], lw_sor_grid[3];
  double lw_ms[8];     // geoms+sort, pairs, narrow, colour, assemble, sor, post (CUDA events, when kernel timing is on)
  cudaEvent_t lw_ev[9];
  // SOR phase split over the GPUs of one box (ObLwSplit): flag words behind fc in ONE allocation (one IPC handle)
  unsigned *lw_flags;
  size_t lw_flags_off;          // bytes from L.fc to lw_flags
  int lw_split_on, lw_split_grid[3], lw_split_threads;
  int lw_split_front, lw_split_sor;   // what is divided over the ranks: pair sweep + narrowphase (default), the SOR sweep (OB_LW_SPLIT_SOR=1)
  ObLwSplit lw_split;
  void *lw_peer_base[OB_LW_MAXRANKS];   // cudaIpcOpenMemHandle mappings to close
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_)); goto fail; } } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the kernel, not of a batch: several batches of different shapes
// live side by side (one drop-in context per space of a nested scene), so the attribute only ever grows (ob_backend_cuda.cu)
cudaError_t ob_func_smem(const void *func, int bytes);

template <class T> static inline cudaError_t dalloc(ObBackend *b, T **p, size_t n) {
  void *q = 0;
  cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
  if (e == cudaSuccess) { b->allocs.push_back(q); cudaMemset(q, 0, (n ? n : 1) * sizeof(T)); }
  *p = (T *)q;
  return e;
}
// ---- seams between the translation units
// ob_kern_collide.cu
int obk_collide_setup(ObBackend *b, const cudaDeviceProp &prop, char *err, size_t errlen);
void obk_collide_launch(ObBackend *b, const ObBatchDev &d, int W, int cap, cudaStream_t st);
// ob_kern_step.cu: tile widths, shared-memory sizes and kernel attributes of the quickstep kernels; one step's four launches
int obk_stepk_setup(ObBackend *b, const cudaDeviceProp &prop, char *err, size_t errlen);
void obk_stepk_launch(ObBackend *b, const ObBatchDev &d, real h, int taps, int W, cudaStream_t st, cudaEvent_t *ev, bool timing);
// ob_kern_large.cu
int lw_create(ObBackend *b, char *err, size_t errlen);
int lw_step(ObBackend *b, real h, int taps, char *err, size_t errlen);

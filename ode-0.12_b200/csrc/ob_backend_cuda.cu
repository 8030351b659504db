// ob_backend_cuda.cu — sm_100a execution backend: device memory + the kernels.
//
// One CTA owns one world for a whole phase and keeps that world's working set
// in shared memory; independent worlds are spread over the grid (grid-stride
// over world ids, so a grid of k*148 CTAs covers any batch size).  Two kernels
// per step:
//   k_collide : pose+AABB per geom -> order-exact dxHashSpace pair list
//               (ob_broad.h) -> thread-per-pair primitive narrowphase
//               (ob_collide.h) -> contact joints in creation order.
//   k_step    : joint graph -> island DFS (order-exact) -> row assembly ->
//               SOR with the reference's seeded row shuffle -> integration ->
//               space-list reordering (what dGeomMoved does).
// Numerics: compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
// so every per-element function in ob_*.h evaluates exactly like the
// reference's scalar SSE build.
//
// SOR parallelisation that keeps results bit-identical: rows are processed in
// the reference's order[] sequence, 32 consecutive positions ("window") per
// warp pass; inside a window a row may run as soon as every earlier row of the
// window that shares a body with it has run (rows that touch disjoint bodies
// commute exactly: they read and write disjoint fc[] entries, and a friction
// row's lambda[findex] lives on the same two bodies as the row itself).  The
// per-window round numbers are computed once per shuffle epoch with shuffles.
#include "ob_backend_cuda.h"

long long g_launches = 0;

#include <map>
#include <mutex>
cudaError_t ob_func_smem(const void *func, int bytes) {
  static std::map<std::pair<int, const void *>, int> high;   // (device, kernel) -> largest size requested so far
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  int &h = high[std::make_pair(dev, func)];
  if (bytes <= h) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) h = bytes;
  return e;
}

// ------------------------------------------------------------------------------------
// bulk state I/O kernels: API order is [world][creation index], device order is newest-first
__global__ void k_pack_state(ObBatchDev d, real *pos3, real *quat4, real *lvel3, real *avel3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.W * d.NB) return;
  const int w = i / d.NB, c = i - w * d.NB, nb = d.world[w].nb;
  if (c >= nb) return;
  const ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
  for (int k = 0; k < 3; k++) { pos3[(size_t)i * 3 + k] = s.pos[k]; lvel3[(size_t)i * 3 + k] = s.lvel[k]; avel3[(size_t)i * 3 + k] = s.avel[k]; }
  for (int k = 0; k < 4; k++) quat4[(size_t)i * 4 + k] = s.q[k];
}
__global__ void k_unpack_state(ObBatchDev d, const real *pos3, const real *quat4, const real *lvel3, const real *avel3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.W * d.NB) return;
  const int w = i / d.NB, c = i - w * d.NB, nb = d.world[w].nb;
  if (c >= nb) return;
  ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
  for (int k = 0; k < 3; k++) {
    if (pos3) s.pos[k] = pos3[(size_t)i * 3 + k];
    if (lvel3) s.lvel[k] = lvel3[(size_t)i * 3 + k];
    if (avel3) s.avel[k] = avel3[(size_t)i * 3 + k];
  }
  if (quat4) {
    // q is stored as given (must be unit): a state read with dBatchGetBodyState restores bit for bit
    for (int k = 0; k < 4; k++) s.q[k] = quat4[(size_t)i * 4 + k];
    ob_RfromQ(s.R, s.q);
  }
}
__global__ void k_add_forces(ObBatchDev d, const real *f3, const real *t3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.W * d.NB) return;
  const int w = i / d.NB, c = i - w * d.NB, nb = d.world[w].nb;
  if (c >= nb) return;
  ObBodyDyn &s = d.bdyn[(size_t)w * d.NB + (nb - 1 - c)];
  for (int k = 0; k < 3; k++) {
    if (f3) s.facc[k] += f3[(size_t)i * 3 + k];
    if (t3) s.tacc[k] += t3[(size_t)i * 3 + k];
  }
}

ObBackend *obk_create(const ObBatchDev &caps, int device, char *err, size_t errlen) {
  ObBackend *b = new ObBackend;
  b->d = caps; b->device = device; b->stream = 0; b->st_dev = 0; b->st_host = 0;
  b->ktiming = 0;
  b->sched_gs = 0; b->smem_sched_tile = 0; b->sor_ring = 0; b->smem_sor_ring = 0; b->ring_resident = 0; b->ring_depth = 0; b->sor_reg = 0; b->smem_sor_reg = 0; b->prep_split = 0; b->sstream = 0; b->sev[0] = b->sev[1] = 0; b->sor_pair = 0; b->smem_sor_pair = 0; b->pair_resident = 0; b->avg_rows = 0; b->cnt_host = 0;
  b->l2_target_bytes = 1e12;   // r02a on B200: limiting the worlds in flight to an L2-sized set costs whole waves (1.72 -> 2.5 -> 3.2 ms at 100 / 60 / 40 MB); off unless OB_SOR_L2MB asks
  for (int k = 0; k < OBK_NKERNELS; k++) { b->kms[k] = 0; b->klaunch[k] = 0; }
  for (int k = 0; k < 8; k++) b->ev[k] = 0;
  ObBatchDev &d = b->d;
  const size_t W = d.W;
  int ndev = 0;
  cudaDeviceProp prop;
  b->nchunks = 1;
  for (int k = 0; k < 8; k++) b->cstream[k] = 0;
  for (int k = 0; k < 9; k++) b->cev[k] = 0;
  b->lw_flags = 0; b->lw_flags_off = 0; b->lw_split_on = 0; b->lw_split_front = 0; b->lw_split_sor = 0; memset(&b->lw_split, 0, sizeof b->lw_split);
  for (int k = 0; k < OB_LW_MAXRANKS; k++) b->lw_peer_base[k] = 0;
  b->large = d.large; b->lw_host = 0; b->lw_rounds = 0; b->lw_ncol = 0;
  for (int k = 0; k < 8; k++) b->lw_stat[k] = 0;
  memset(&b->L, 0, sizeof b->L);
  for (int k = 0; k < 9; k++) b->lw_ev[k] = 0;
  for (int k = 0; k < 8; k++) b->lw_ms[k] = 0;
  if (!d.large && (d.NB > 32000 || d.NC > 32000 || d.NR > 65000)) { snprintf(err, errlen, "world too large for the CTA-per-world path (NB=%d NC=%d NR=%d)", d.NB, d.NC, d.NR); goto fail2; }
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { snprintf(err, errlen, "no CUDA device available (this library has no CPU fallback)"); goto fail2; }
  CK(cudaSetDevice(device));
  CK(cudaGetDeviceProperties(&prop, device));
  CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  for (int k = 0; k < 8; k++) CK(cudaEventCreate(&b->ev[k]));
  for (int k = 0; k < 8; k++) CK(cudaStreamCreateWithFlags(&b->cstream[k], cudaStreamNonBlocking));
  for (int k = 0; k < 9; k++) CK(cudaEventCreateWithFlags(&b->cev[k], cudaEventDisableTiming));
  d.wbeg = 0; d.wend = d.W;
  if (d.large) {
    if (lw_create(b, err, errlen)) goto fail;
    return b;
  }
  CK(dalloc(b, &d.world, W));
  CK(dalloc(b, &d.bdyn, W * d.NB));
  CK(dalloc(b, &d.bconst, W * d.NB));
  CK(dalloc(b, &d.geom, W * d.NG));
  CK(dalloc(b, &d.glist, W * d.NG));
  CK(dalloc(b, &d.sapstate, W * (d.NG + 3)));
  CK(dalloc(b, &d.policy, (size_t)d.npolicy));
  CK(dalloc(b, &d.meshes, (size_t)(d.nmesh ? d.nmesh : 1)));
  CK(dalloc(b, &d.joint, W * (d.NJ ? d.NJ : 1)));
  CK(dalloc(b, &d.njoints, W));
  CK(dalloc(b, &d.padjstart, W * (d.NB + 1)));
  CK(dalloc(b, &d.padj, W * 2 * (d.NJ ? d.NJ : 1)));
  CK(dalloc(b, &d.npairs, W));
  CK(dalloc(b, &d.pairs, W * d.NP * 2));
  CK(dalloc(b, &d.ncontacts, W));
  CK(dalloc(b, &d.contacts, W * d.NC));
  if (d.NEP > OB_MAXEPOCH) { snprintf(err, errlen, "more than %d SOR iterations are not supported", 8 * OB_MAXEPOCH); goto fail; }
  CK(dalloc(b, &d.rows, W * d.NR * OB_ROWW));
  CK(dalloc(b, &d.stepinfo, W * SI_WORDS));
  CK(dalloc(b, &d.ibody, W * d.NB));
  CK(dalloc(b, &d.isz, W * d.NB * 4));
  CK(dalloc(b, &d.jrow, W * (d.NC + d.NJ + 1)));
  CK(dalloc(b, &d.ijoint, W * (d.NC + d.NJ)));
  CK(dalloc(b, &d.jside, W * (d.NJ ? d.NJ : 1) * 4 * OB_NSIDE));
  CK(dalloc(b, &d.sched, W * d.NEP * d.NR));
  CK(dalloc(b, &d.pstart, W * d.NEP * (d.NR + 1)));
  CK(dalloc(b, &d.invIw, W * d.NB * 12));
  CK(dalloc(b, &d.tmp1, W * d.NB * 8));
  d.rowJ = d.rowiMJ = d.rowJc = d.rowS = 0; d.rowI = 0;
  CK(dalloc(b, &d.lambda, W * d.NR));
  CK(dalloc(b, &d.nrows, W));
  CK(dalloc(b, &d.fback, W * (d.NC + d.NJ) * 12));
  d.csurf = 0; d.cfdir1 = 0;
  if (d.dropin) { CK(dalloc(b, &d.csurf, W * d.NC)); CK(dalloc(b, &d.cfdir1, W * d.NC * 4)); }
  CK(dalloc(b, &d.counters, (size_t)1));
  d.adisbuf = 0; d.adisctl = 0;
  CK(dalloc(b, &d.rowmeta, W * d.NR));
  CK(cudaStreamCreateWithFlags(&b->sstream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&b->sev[0], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&b->sev[1], cudaEventDisableTiming));
  { const char *e = getenv("OB_PREP_SPLIT"); b->prep_split = e ? atoi(e) != 0 : (W >= 256); }
  if (d.NADIS > 0) { CK(dalloc(b, &d.adisbuf, W * d.NB * d.NADIS * 6)); CK(dalloc(b, &d.adisctl, W * d.NB * 2)); }
  b->st_elems = W * d.NB;
  CK(dalloc(b, &b->st_dev, b->st_elems * 13));
  if (obk_stepk_setup(b, prop, err, errlen)) goto fail;
  if (obk_collide_setup(b, prop, err, errlen)) goto fail;
  {
    int nch = 1;   // measured on B200 (configs 2-4): 2-8 chunks change the step time by -4 % .. +10 %, so off unless OB_CHUNKS asks
    const char *e = getenv("OB_CHUNKS");
    if (e && atoi(e) >= 1 && atoi(e) <= 8) nch = atoi(e);
    b->nchunks = nch;
  }
  {
    // grid: every world gets its own CTA up to 16 resident CTAs per SM worth of blocks, beyond that grid-stride
    int cap = prop.multiProcessorCount * 32;
    b->grid = (int)W < cap ? (int)W : cap;
  }
  return b;
fail:
  for (int k = 0; k < OB_LW_MAXRANKS; k++) if (b->lw_peer_base[k]) cudaIpcCloseMemHandle(b->lw_peer_base[k]);
  for (size_t i = 0; i < b->allocs.size(); i++) cudaFree(b->allocs[i]);
  if (b->st_host) cudaFreeHost(b->st_host);
  if (b->lw_host) cudaFreeHost(b->lw_host);
  if (b->cnt_host) cudaFreeHost(b->cnt_host);
  if (b->stream) cudaStreamDestroy(b->stream);
fail2:
  delete b;
  return 0;
}

void obk_destroy(ObBackend *b) {
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  for (int k = 0; k < OB_LW_MAXRANKS; k++) if (b->lw_peer_base[k]) cudaIpcCloseMemHandle(b->lw_peer_base[k]);
  for (size_t i = 0; i < b->allocs.size(); i++) cudaFree(b->allocs[i]);
  if (b->st_host) cudaFreeHost(b->st_host);
  if (b->lw_host) cudaFreeHost(b->lw_host);
  if (b->cnt_host) cudaFreeHost(b->cnt_host);
  for (int k = 0; k < 9; k++) if (b->lw_ev[k]) cudaEventDestroy(b->lw_ev[k]);
  for (int k = 0; k < 8; k++) if (b->ev[k]) cudaEventDestroy(b->ev[k]);
  for (int k = 0; k < 8; k++) if (b->cstream[k]) cudaStreamDestroy(b->cstream[k]);
  for (int k = 0; k < 9; k++) if (b->cev[k]) cudaEventDestroy(b->cev[k]);
  if (b->sstream) { cudaStreamSynchronize(b->sstream); cudaStreamDestroy(b->sstream); }
  for (int k = 0; k < 2; k++) if (b->sev[k]) cudaEventDestroy(b->sev[k]);
  cudaStreamDestroy(b->stream);
  delete b;
}
ObBatchDev *obk_arrays(ObBackend *b) { return &b->d; }
int obk_h2d(ObBackend *b, void *dst, const void *src, size_t n) {
  cudaSetDevice(b->device);
  if (cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, b->stream) != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_d2h(ObBackend *b, void *dst, const void *src, size_t n) {
  cudaSetDevice(b->device);
  if (cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, b->stream) != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_memset(ObBackend *b, void *dst, int v, size_t n) {
  cudaSetDevice(b->device);
  if (cudaMemsetAsync(dst, v, n, b->stream) != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_sync(ObBackend *b) { cudaSetDevice(b->device); return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1; }
void *obk_stream(ObBackend *b) { return (void *)b->stream; }
long long obk_launch_count(void) { return g_launches; }
int obk_timer_start(ObBackend *b) { cudaSetDevice(b->device); return cudaEventRecord(b->ev[0], b->stream) == cudaSuccess ? 0 : -1; }
int obk_timer_stop(ObBackend *b, float *ms) {
  cudaSetDevice(b->device);
  if (cudaEventRecord(b->ev[1], b->stream) != cudaSuccess) return -1;
  if (cudaEventSynchronize(b->ev[1]) != cudaSuccess) return -1;
  return cudaEventElapsedTime(ms, b->ev[0], b->ev[1]) == cudaSuccess ? 0 : -1;
}
void obk_set_kernel_timing(ObBackend *b, int enable) {
  b->ktiming = enable;
  for (int k = 0; k < 8; k++) b->lw_ms[k] = 0;
  b->lw_stat[7] = 0;
  for (int k = 0; k < OBK_NKERNELS; k++) { b->kms[k] = 0; b->klaunch[k] = 0; }
}
void obk_get_kernel_times(ObBackend *b, double *ms, long long *l) {
  for (int k = 0; k < OBK_NKERNELS; k++) { ms[k] = b->kms[k]; l[k] = b->klaunch[k]; }
}
int obk_large_stats(ObBackend *b, int *ints8, double *ms8) {
  if (!b->large) return -1;
  for (int k = 0; k < 8; k++) { ints8[k] = b->lw_stat[k]; ms8[k] = b->lw_ms[k]; }
  return 0;
}
const char *obk_kernel_name(int k) { static const char *n[] = {"k_collide", "k_prep", "k_sched", "k_sor", "k_post"}; return k >= 0 && k < 5 ? n[k] : ""; }

static void launch_step(ObBackend *b, real h, int taps, int phases, int w0, int w1, cudaStream_t st, bool timing) {
  cudaEvent_t *ev = b->ev + 2;
  ObBatchDev d = b->d;
  d.wbeg = w0; d.wend = w1;
  const int W = w1 - w0;
  const int cap = b->grid;   // resident-CTA cap computed for the whole batch
  if (timing) cudaEventRecord(ev[0], st);
  if (phases & OBK_PHASE_COLLIDE) obk_collide_launch(b, d, W, cap, st);
  if (timing) cudaEventRecord(ev[1], st);
  if (phases & OBK_PHASE_STEP) obk_stepk_launch(b, d, h, taps, W, st, ev, timing);
  if (timing && phases == (OBK_PHASE_COLLIDE | OBK_PHASE_STEP)) {
    cudaEventRecord(ev[5], st);
    if (cudaEventSynchronize(ev[5]) == cudaSuccess)
      for (int k = 0; k < 5; k++) { float m = 0; cudaEventElapsedTime(&m, ev[k], ev[k + 1]); b->kms[k] += m; b->klaunch[k]++; }
  }
}
static void launch_steps(ObBackend *b, real h, int nsteps, int taps, int phases) {
  const int W = b->d.W;
  // per-kernel timing and the parity taps run unchunked on the main stream
  const int nch = (b->ktiming || (taps & 1) || b->d.dropin || nsteps < 2) ? 1 : b->nchunks;
  if (nch <= 1) {
    for (int s = 0; s < nsteps; s++) {
      // parity tap: joints that enter no island (attached to no body / to disabled bodies) report zero feedback
      if ((taps & 1) && !b->d.dropin && b->d.fback) cudaMemsetAsync(b->d.fback, 0, sizeof(real) * 12 * (size_t)W * (b->d.NC + b->d.NJ), b->stream);
      launch_step(b, h, taps, phases, 0, W, b->stream, b->ktiming != 0);
    }
    return;
  }
  // fork: every chunk stream starts after what is already queued on the main stream; worlds are independent,
  // so the chunks never wait for each other between steps; join: the main stream waits for all of them
  cudaEventRecord(b->cev[8], b->stream);
  const int per = ((W + nch - 1) / nch + 31) / 32 * 32;   // multiple of 32 keeps warp tiles whole
  for (int c = 0; c < nch; c++) {
    const int w0 = c * per, w1 = (c + 1) * per < W ? (c + 1) * per : W;
    if (w0 >= w1) continue;
    cudaStreamWaitEvent(b->cstream[c], b->cev[8], 0);
    for (int s = 0; s < nsteps; s++) launch_step(b, h, taps, phases, w0, w1, b->cstream[c], false);
    cudaEventRecord(b->cev[c], b->cstream[c]);
    cudaStreamWaitEvent(b->stream, b->cev[c], 0);
  }
}

static int run_steps(ObBackend *b, real h, int nsteps, int taps, int phases, char *err, size_t errlen) {
  cudaSetDevice(b->device);
  if (getenv("OB_SEQ")) taps |= 2;
  if (getenv("OB_CHECK")) taps |= 4;
  if (getenv("OB_SYNC2")) taps |= 8;
  if (b->large) {
    if (phases != (OBK_PHASE_COLLIDE | OBK_PHASE_STEP)) { snprintf(err, errlen, "the large-world path runs whole steps only"); return -1; }
    for (int s = 0; s < nsteps; s++) if (lw_step(b, h, taps, err, errlen)) return -1;
    return 0;
  }
  launch_steps(b, h, nsteps, taps, phases);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && b->cnt_host) e = cudaMemcpyAsync(b->cnt_host, b->d.counters, sizeof(ObCounters), cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  if (e != cudaSuccess) { snprintf(err, errlen, "kernel launch/exec failed: %s", cudaGetErrorString(e)); return -1; }
  if (b->cnt_host && b->cnt_host->steps > 0) b->avg_rows = (double)b->cnt_host->rows / (double)b->cnt_host->steps;
  return 0;
}
int obk_step(ObBackend *b, real h, int nsteps, int taps, char *err, size_t errlen) {
  return run_steps(b, h, nsteps, taps, OBK_PHASE_COLLIDE | OBK_PHASE_STEP, err, errlen);
}
int obk_run_phases(ObBackend *b, real h, int phases, int taps, char *err, size_t errlen) {
  return run_steps(b, h, 1, taps, phases, err, errlen);
}

int obk_mesh_upload(const float *verts, int nverts, const int *tris, int ntris, const ObBvNode *nodes, const unsigned char *useflags, int device, ObMeshDev *io) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  float *dv = 0; int *dt = 0; ObBvNode *dn = 0; int *df = 0;
  std::vector<int> vfirst((size_t)(nverts > 0 ? nverts : 1), -1);
  for (int c = 0; c < 3 * ntris; c++) { const int vi = tris[c]; if (vi >= 0 && vi < nverts && vfirst[vi] < 0) vfirst[vi] = c; }
  if (cudaMalloc((void **)&df, sizeof(int) * vfirst.size()) != cudaSuccess) return -1;
  if (cudaMemcpy(df, vfirst.data(), sizeof(int) * vfirst.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(df); return -1; }
  io->vfirst = df;
  io->useflags = 0;
  if (useflags) {
    unsigned char *du = 0;
    if (cudaMalloc((void **)&du, (size_t)ntris) != cudaSuccess) return -1;
    if (cudaMemcpy(du, useflags, (size_t)ntris, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(du); return -1; }
    io->useflags = du;
  }
  if (cudaMalloc((void **)&dv, sizeof(float) * 3 * (size_t)nverts) != cudaSuccess) return -1;
  if (cudaMalloc((void **)&dt, sizeof(int) * 3 * (size_t)ntris) != cudaSuccess) { cudaFree(dv); return -1; }
  if (cudaMalloc((void **)&dn, sizeof(ObBvNode) * (size_t)(ntris - 1)) != cudaSuccess) { cudaFree(dv); cudaFree(dt); return -1; }
  cudaError_t e = cudaMemcpy(dv, verts, sizeof(float) * 3 * (size_t)nverts, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dt, tris, sizeof(int) * 3 * (size_t)ntris, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dn, nodes, sizeof(ObBvNode) * (size_t)(ntris - 1), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(dv); cudaFree(dt); cudaFree(dn); return -1; }
  io->verts = dv; io->tris = dt; io->nodes = dn; io->nverts = nverts; io->ntris = ntris;
  return 0;
}
void obk_mesh_free(ObMeshDev *m) {
  if (m->verts) cudaFree((void *)m->verts);
  if (m->tris) cudaFree((void *)m->tris);
  if (m->nodes) cudaFree((void *)m->nodes);
  if (m->vfirst) cudaFree((void *)m->vfirst);
  if (m->useflags) cudaFree((void *)m->useflags);
  m->verts = 0; m->tris = 0; m->nodes = 0; m->vfirst = 0; m->useflags = 0;
}

// Bulk state I/O copies straight between the caller's buffers and the packed device staging
// array: with buffers from dBatchHostAlloc (pinned) these are plain DMA transfers, with pageable
// memory the driver stages them.
int obk_get_state(ObBackend *b, real *pos3, real *quat4, real *lvel3, real *avel3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *dp = b->st_dev, *dq = dp + n * 3, *dl = dq + n * 4, *da = dl + n * 3;
  k_pack_state<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, dp, dq, dl, da);
  g_launches++;
  cudaError_t e = cudaSuccess;
  if (pos3 && e == cudaSuccess) e = cudaMemcpyAsync(pos3, dp, n * 3 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (quat4 && e == cudaSuccess) e = cudaMemcpyAsync(quat4, dq, n * 4 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (lvel3 && e == cudaSuccess) e = cudaMemcpyAsync(lvel3, dl, n * 3 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (avel3 && e == cudaSuccess) e = cudaMemcpyAsync(avel3, da, n * 3 * sizeof(real), cudaMemcpyDeviceToHost, b->stream);
  if (e != cudaSuccess) return -1;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_set_state(ObBackend *b, const real *pos3, const real *quat4, const real *lvel3, const real *avel3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *dp = b->st_dev, *dq = dp + n * 3, *dl = dq + n * 4, *da = dl + n * 3;
  cudaError_t e = cudaSuccess;
  if (pos3 && e == cudaSuccess) e = cudaMemcpyAsync(dp, pos3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (quat4 && e == cudaSuccess) e = cudaMemcpyAsync(dq, quat4, n * 4 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (lvel3 && e == cudaSuccess) e = cudaMemcpyAsync(dl, lvel3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (avel3 && e == cudaSuccess) e = cudaMemcpyAsync(da, avel3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (e != cudaSuccess) return -1;
  k_unpack_state<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, pos3 ? dp : 0, quat4 ? dq : 0, lvel3 ? dl : 0, avel3 ? da : 0);
  g_launches++;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
int obk_add_forces(ObBackend *b, const real *f3, const real *t3) {
  cudaSetDevice(b->device);
  const size_t n = b->st_elems;
  real *df = b->st_dev, *dt = df + n * 3;
  cudaError_t e = cudaSuccess;
  if (f3) e = cudaMemcpyAsync(df, f3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (t3 && e == cudaSuccess) e = cudaMemcpyAsync(dt, t3, n * 3 * sizeof(real), cudaMemcpyHostToDevice, b->stream);
  if (e != cudaSuccess) return -1;
  k_add_forces<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d, f3 ? df : 0, t3 ? dt : 0);
  g_launches++;
  return cudaStreamSynchronize(b->stream) == cudaSuccess ? 0 : -1;
}
void *obk_host_alloc(size_t bytes) { void *p = 0; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : 0; }
void obk_host_free(void *p) { if (p) cudaFreeHost(p); }


// diagnostics: the libm restatements of ob_math.h on the device
__global__ void k_libm(int fn, int n, const float *a, const float *b, float *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fn == 0 ? ob_atan2f_glibc(a[i], b[i]) : (fn == 1 ? ob_sinf_glibc(a[i]) : ob_cosf_glibc(a[i]));
}
int obk_libm(int fn, int n, const float *a, const float *b, float *out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || n <= 0) return -1;
  float *da = 0, *db = 0, *dout = 0;
  const size_t nb = sizeof(float) * (size_t)n;
  cudaError_t e = cudaMalloc((void **)&da, nb);
  if (e == cudaSuccess) e = cudaMalloc((void **)&db, nb);
  if (e == cudaSuccess) e = cudaMalloc((void **)&dout, nb);
  if (e == cudaSuccess) e = cudaMemcpy(da, a, nb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db, b ? b : a, nb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { k_libm<<<(unsigned)((n + 255) / 256), 256>>>(fn, n, da, db, dout); g_launches++; e = cudaMemcpy(out, dout, nb, cudaMemcpyDeviceToHost); }
  if (da) cudaFree(da);
  if (db) cudaFree(db);
  if (dout) cudaFree(dout);
  return e == cudaSuccess ? 0 : -1;
}

// ob_kern_step.cu — the quickstep kernels (ob_step_kernel.cuh): tile widths, shared-memory sizes, kernel attributes and
// the four launches of one step (k_prep, k_sched*, k_sor*, k_post).
#include "ob_backend_cuda.h"
#include "ob_step_kernel.cuh"

int obk_stepk_setup(ObBackend *b, const cudaDeviceProp &prop, char *err, size_t errlen) {
  ObBatchDev &d = b->d;
  const size_t W = d.W;
  {
    // tile width: G lanes per world, 32/G worlds per warp.  Narrow tiles waste fewer lanes in the
    // dependency rounds of the SOR sweep; wide tiles finish one world sooner.  Heuristic on batch size.
    int G = W >= 1024 ? 8 : (W >= 256 ? 16 : 32);
    if (W >= 4096 && d.NB <= 8) G = 4;   // tiny worlds (config 3: 5 bodies, <= 3 rows per level): 8 worlds per warp, 10.5 -> 8.3 ms/step
    const char *e = getenv("OB_TILE");
    if (e && (atoi(e) == 4 || atoi(e) == 8 || atoi(e) == 16 || atoi(e) == 32)) G = atoi(e);
    b->tile = G;
    // row assembly (k_prep) has its own tile width: it is a throughput kernel (lane per joint / row) that waits on
    // scattered loads, so more, narrower-batched warps can pay even where the sweep prefers few lanes per world
    b->prep_tile = G;
    // measured on B200: contact-only worlds with hundreds of rows (config 2) 0.44 -> 0.38 ms with one world per warp;
    // jointed / tiny worlds (configs 3, 4) are fastest at the sweep's own width (config 3: 1.26 / 1.35 / 1.50 / 2.17 ms at 4 / 8 / 16 / 32)
    if (d.NJ == 0 && d.NC >= 96) b->prep_tile = 32;
    { const char *pe = getenv("OB_PREP_TILE"); if (pe && (atoi(pe) == 4 || atoi(pe) == 8 || atoi(pe) == 16 || atoi(pe) == 32)) b->prep_tile = atoi(pe); }
    // the first half (graph, islands: serial per world on the tile's lane 0) may run on narrower tiles than the second (row assembly)
    b->prep_tile1 = b->prep_tile;
    { const char *pe = getenv("OB_PREP_TILE1"); if (pe && (atoi(pe) == 4 || atoi(pe) == 8 || atoi(pe) == 16 || atoi(pe) == 32)) b->prep_tile1 = atoi(pe); }
    b->smem_prep1 = prep_tile_smem(d.NB, d.NC, d.NJ, d.NR).total * (32 / b->prep_tile1);
    b->smem_prep = prep_tile_smem(d.NB, d.NC, d.NJ, d.NR).total * (32 / b->prep_tile);
    b->smem_sor = sor_tile_smem(d.NB, d.NR).total * (32 / G);
    b->smem_sched = sched_smem(d.NB, d.NR).total;
    // integration (k_post) is a lane-per-body kernel of a few microseconds per world: it wants warps, not lane efficiency -- the widest
    // tile that still gives ~4096 warps (r02z, configs[1]: 1024 warps, 5.5 % issue-active, 100 us on the critical path)
    b->post_tile = G;
    while (b->post_tile < 32 && (long long)W * (2 * b->post_tile) / 32 <= 4096) b->post_tile *= 2;
    { const char *pe = getenv("OB_POST_TILE"); if (pe && (atoi(pe) == 4 || atoi(pe) == 8 || atoi(pe) == 16 || atoi(pe) == 32)) b->post_tile = atoi(pe); }
    b->smem_post = post_tile_smem(d.NG).total * (32 / b->post_tile);
    b->grid_step = (int)((W + (32 / G) - 1) / (32 / G));
    b->grid_sor = b->grid_step;
    { const char *g = getenv("OB_GRID_SOR"); if (g && atoi(g) > 0 && atoi(g) < b->grid_sor) b->grid_sor = atoi(g); }
  }
  if (d.NB > 254 || d.NG > 255 || d.NC + d.NJ > 65000) { snprintf(err, errlen, "world too large for the tile-per-world step kernel (NB=%d NG=%d NR=%d)", d.NB, d.NG, d.NR); goto fail; }
  if (b->smem_prep1 > (size_t)prop.sharedMemPerBlockOptin) { b->prep_tile1 = b->prep_tile; b->smem_prep1 = b->smem_prep; }
  if (b->smem_prep > (size_t)prop.sharedMemPerBlockOptin || b->smem_sor > (size_t)prop.sharedMemPerBlockOptin) {
    snprintf(err, errlen, "world does not fit one CTA's shared memory (prep %zu B, sor %zu B, limit %zu B)",
             b->smem_prep, b->smem_sor, (size_t)prop.sharedMemPerBlockOptin);
    goto fail;
  }
#define OB_SETSMEM(GG) \
  CK(ob_func_smem((const void *)k_prep<GG, true, 0>, (int)b->smem_prep)); \
  CK(ob_func_smem((const void *)k_prep<GG, false, 0>, (int)b->smem_prep)); \
  CK(ob_func_smem((const void *)k_prep<GG, true, 1>, (int)(b->smem_prep1 > b->smem_prep ? b->smem_prep1 : b->smem_prep))); \
  CK(ob_func_smem((const void *)k_prep<GG, false, 1>, (int)(b->smem_prep1 > b->smem_prep ? b->smem_prep1 : b->smem_prep))); \
  CK(ob_func_smem((const void *)k_prep<GG, true, 2>, (int)b->smem_prep)); \
  CK(ob_func_smem((const void *)k_prep<GG, false, 2>, (int)b->smem_prep)); \
  CK(ob_func_smem((const void *)k_sor<GG, true>, (int)b->smem_sor)); \
  CK(ob_func_smem((const void *)k_sor<GG, false>, (int)b->smem_sor)); \
  CK(ob_func_smem((const void *)k_post<GG>, (int)b->smem_post));
  OB_SETSMEM(4) OB_SETSMEM(8) OB_SETSMEM(16) OB_SETSMEM(32)
#undef OB_SETSMEM
  CK(ob_func_smem((const void *)k_sched<2>, (int)b->smem_sched));
  CK(ob_func_smem((const void *)k_sched<4>, (int)b->smem_sched));
  CK(ob_func_smem((const void *)k_sched<8>, (int)b->smem_sched));
  // one lane per world for batches of many tiny worlds (k_sor_lane); OB_SOR_LANE=0 / 1 overrides
  b->smem_sor_lane = sor_lane_smem(d.NB).total;
  b->sor_lane = (W >= 8192 && d.NB <= 8) ? 1 : 0;
  { const char *e = getenv("OB_SOR_LANE"); if (e) b->sor_lane = atoi(e) != 0; }
  if (b->smem_sor_lane > (size_t)prop.sharedMemPerBlockOptin || d.NB > 254) b->sor_lane = 0;
  if (b->sor_lane) CK(ob_func_smem((const void *)k_sor_lane, (int)b->smem_sor_lane));
  b->sor_deep = d.NR > 256 ? 1 : 0;
  { const char *e = getenv("OB_SOR_DEEP"); if (e) b->sor_deep = atoi(e) != 0; }
  b->smem_sched_lane = sched_lane_smem(d.NB, d.NR).total;
  // measured on B200: one lane per world wins for many small worlds (config 3: 65536 worlds x 56 rows, 0.70 -> 0.43 ms),
  // the warp per world for fewer, larger ones (config 2: 4096 x 377 rows, 0.40 vs 2.7 ms: too few warps to hide the chain latency)
  b->sched_lane = b->smem_sched_lane <= (size_t)prop.sharedMemPerBlockOptin && ((W >= 8192 && d.NR <= 256) || getenv("OB_SCHED_LANE")) && !getenv("OB_SCHED_WARP");
  if (b->sched_lane) CK(ob_func_smem((const void *)k_sched_lane, (int)b->smem_sched_lane));
  {
    // k_sched_tile: lanes per world by batch size (enough warps to hide the serial chains' latency, few enough lanes
    // per world that a warp instruction of the chains advances several worlds)
    // measured on B200 (r02a, configs[1]): 0.95 / 0.55 / 0.39 ms at 4 / 8 / 16 lanes per world against 0.39 ms for the warp per
    // world -- the chains are latency-bound, fewer warps lose what fewer instructions gain; off unless OB_SCHED_TILE asks
    // r02e (flat serial loops, build without --split-compile): 0.31 ms at 16 lanes per world against 0.39 ms for the warp per world
    int gs = W >= 2048 ? 16 : 0;
    const char *e = getenv("OB_SCHED_TILE");
    if (e) gs = atoi(e);
    if (gs != 2 && gs != 4 && gs != 8 && gs != 16) gs = 0;
    if (gs && !(e == 0 && b->sched_lane)) {
      b->smem_sched_tile = sched_tile_smem(d.NB, d.NR).total * (32 / gs);
      if (b->smem_sched_tile <= (size_t)prop.sharedMemPerBlockOptin) {
        b->sched_gs = gs;
        if (gs == 2) CK(ob_func_smem((const void *)k_sched_tile<2>, (int)b->smem_sched_tile));
        if (gs == 4) CK(ob_func_smem((const void *)k_sched_tile<4>, (int)b->smem_sched_tile));
        if (gs == 8) CK(ob_func_smem((const void *)k_sched_tile<8>, (int)b->smem_sched_tile));
        if (gs == 16) CK(ob_func_smem((const void *)k_sched_tile<16>, (int)b->smem_sched_tile));
      }
    }
  }
  {
    // k_sor_ring: on unless the caller pins one of the register-pipelined variants (OB_SOR_DEEP) or OB_SOR_RING=0.
    // Ring depth: 6 slots (rows 5 passes ahead) when every CTA of the batch still stays resident, else 4
    const char *e = getenv("OB_SOR_RING");
    // measured (r02e): configs[1] 0.94 vs 0.93 ms, configs[3] 2.82 vs 3.10 ms (ring vs register pipeline) at 8 lanes per world;
    // tiny worlds at 4 lanes per world (configs[2]) 4.08 vs 3.10 ms: the ring's per-pass bookkeeping outweighs its prefetch there
    const bool want = e ? atoi(e) != 0 : (getenv("OB_SOR_DEEP") == 0 && b->tile >= 8);
    const int Tw = 32 / b->tile;
    const int need = (int)((W + Tw - 1) / Tw);
    b->ring_depth = 0;
    if (want) {
      const int depths[2] = {6, 4};
      int best_res = 0;
      for (int k = 0; k < 2; k++) {
        const int D = depths[k];
        { const char *de = getenv("OB_RING_DEPTH"); if (de && atoi(de) != D) continue; }
        const size_t sm = sor_ring_smem(d.NB, d.NR, b->tile, D).total * Tw;
        if (sm > (size_t)prop.sharedMemPerBlockOptin) continue;
        int per_sm = 0;
#define OB_RING_SETUP(GG, DD) { CK(ob_func_smem((const void *)k_sor_ring<GG, DD>, (int)sm)); \
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sor_ring<GG, DD>, 32, sm)); }
#define OB_RING_SETUP_G(DD) { if (b->tile == 4) OB_RING_SETUP(4, DD) else if (b->tile == 8) OB_RING_SETUP(8, DD) else if (b->tile == 16) OB_RING_SETUP(16, DD) else OB_RING_SETUP(32, DD) }
        if (D == 6) OB_RING_SETUP_G(6) else OB_RING_SETUP_G(4)
#undef OB_RING_SETUP_G
#undef OB_RING_SETUP
        const int res = per_sm * prop.multiProcessorCount;
        if (res <= 0) continue;
        // take the deeper ring unless it costs residency the batch needs
        if (b->ring_depth == 0 || (best_res < need && res > best_res)) { b->ring_depth = D; b->smem_sor_ring = sm; b->ring_resident = res; best_res = res; }
      }
      b->sor_ring = b->ring_depth != 0;
    }
    // k_sor_reg: the pipelined pass with register row buffers (no ring): preferred over the ring when it fits
    b->sor_reg = 0;
    {
      const char *re = getenv("OB_SOR_REG");
      const bool wantr = want && (re ? atoi(re) != 0 : false);   // r02g on B200: 1.15 ms against the ring's 0.88 on configs[1] (the loads three passes ahead do not hide the global latency the way the 4-deep LDGSTS ring does): opt-in
      b->smem_sor_reg = sor_reg_smem(d.NB, d.NR, b->tile).total * Tw;
      if (wantr && b->smem_sor_reg <= (size_t)prop.sharedMemPerBlockOptin) {
        if (b->tile == 4) CK(ob_func_smem((const void *)k_sor_reg<4>, (int)b->smem_sor_reg));
        if (b->tile == 8) CK(ob_func_smem((const void *)k_sor_reg<8>, (int)b->smem_sor_reg));
        if (b->tile == 16) CK(ob_func_smem((const void *)k_sor_reg<16>, (int)b->smem_sor_reg));
        if (b->tile == 32) CK(ob_func_smem((const void *)k_sor_reg<32>, (int)b->smem_sor_reg));
        b->sor_reg = 1;
      }
    }
    // k_sor_pair: two lanes per row (2 * tile lanes per world), on top of the ring's machinery; ring depth 5, or 4 when that keeps more CTAs resident
    b->sor_pair = 0;
    {
      const char *pe = getenv("OB_SOR_PAIR");
      const bool wantp = b->sor_ring && b->tile <= 16 && (pe ? atoi(pe) != 0 : false);   // r02e: 1.04 ms against the ring's 0.94 on configs[1] (shared-memory pipe at 65 %): opt-in
      if (wantp) {
        const int GP = 2 * b->tile, Tp = 32 / GP;
        const int needp = (int)((W + Tp - 1) / Tp);
        int best = 0;
        const int depths[2] = {5, 4};
        for (int k = 0; k < 2; k++) {
          const int D = depths[k];
          { const char *de = getenv("OB_PAIR_DEPTH"); if (de && atoi(de) != D) continue; }
          const size_t sm = sor_ring_smem(d.NB, d.NR, b->tile, D).total * Tp;
          if (sm > (size_t)prop.sharedMemPerBlockOptin) continue;
          int per_sm = 0;
#define OB_PAIR_SETUP(GG, DD) { CK(ob_func_smem((const void *)k_sor_pair<GG, DD>, (int)sm)); \
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sor_pair<GG, DD>, 32, sm)); }
#define OB_PAIR_SETUP_G(DD) { if (GP == 8) OB_PAIR_SETUP(8, DD) else if (GP == 16) OB_PAIR_SETUP(16, DD) else OB_PAIR_SETUP(32, DD) }
          if (D == 5) OB_PAIR_SETUP_G(5) else OB_PAIR_SETUP_G(4)
#undef OB_PAIR_SETUP_G
#undef OB_PAIR_SETUP
          const int res = per_sm * prop.multiProcessorCount;
          if (res <= 0) continue;
          if (b->sor_pair == 0 || (best < needp && res > best)) { b->sor_pair = D; b->smem_sor_pair = sm; b->pair_resident = res; best = res; }
        }
      }
    }
    const char *l2 = getenv("OB_SOR_L2MB");
    if (l2 && atof(l2) > 0) b->l2_target_bytes = atof(l2) * 1e6;
    CK(cudaMallocHost((void **)&b->cnt_host, sizeof(ObCounters)));
    memset(b->cnt_host, 0, sizeof(ObCounters));
  }
  return 0;
fail:
  return -1;
}

template <int G> static void stepk_launch_t(ObBackend *b, const ObBatchDev &d, real h, int taps, int W, cudaStream_t st, cudaEvent_t *ev, bool timing) {
  constexpr int T = 32 / G;
    const int gstep = (W + T - 1) / T;
    int gsor = gstep;
    if (b->grid_sor < b->grid_step) gsor = gsor < b->grid_sor ? gsor : b->grid_sor;
    // k_prep in two halves with the schedule on a second stream beside the second half (it needs only what the first half
    // leaves: island / row tables and the rows' body + findex bytes).  Per-kernel timing and small batches keep the single launch.
    const bool split = b->prep_split && !timing && b->nchunks <= 1;
    ObBatchDev dk = d;
    dk.rowmeta = split ? d.rowmeta : (unsigned *)0;
    cudaStream_t ss = split ? b->sstream : st;
#define OB_LAUNCH_PREP(GP, PH) { const int gp = (W + (32 / GP) - 1) / (32 / GP); const size_t sm_ = PH == 1 ? b->smem_prep1 : b->smem_prep; \
      if (d.NJ > 0) k_prep<GP, true, PH><<<gp, 32, sm_, st>>>(dk, h, taps); else k_prep<GP, false, PH><<<gp, 32, sm_, st>>>(dk, h, taps); }
#define OB_LAUNCH_PREP_G(PH) { const int pt_ = PH == 1 ? b->prep_tile1 : b->prep_tile; if (pt_ == 4) OB_LAUNCH_PREP(4, PH) else if (pt_ == 8) OB_LAUNCH_PREP(8, PH) else if (pt_ == 16) OB_LAUNCH_PREP(16, PH) else OB_LAUNCH_PREP(32, PH) }
    if (split) {
      OB_LAUNCH_PREP_G(1)
      cudaEventRecord(b->sev[0], st);
      cudaStreamWaitEvent(ss, b->sev[0], 0);
    } else OB_LAUNCH_PREP_G(0)
    if (timing) cudaEventRecord(ev[2], st);
    if (b->sched_gs == 2) k_sched_tile<2><<<(W + 15) / 16, 32, b->smem_sched_tile, ss>>>(dk, G);
    else if (b->sched_gs == 4) k_sched_tile<4><<<(W + 7) / 8, 32, b->smem_sched_tile, ss>>>(dk, G);
    else if (b->sched_gs == 8) k_sched_tile<8><<<(W + 3) / 4, 32, b->smem_sched_tile, ss>>>(dk, G);
    else if (b->sched_gs == 16) k_sched_tile<16><<<(W + 1) / 2, 32, b->smem_sched_tile, ss>>>(dk, G);
    else if (b->sched_lane) k_sched_lane<<<(W + 31) / 32, 32, b->smem_sched_lane, ss>>>(dk, G);
    else if (d.NB <= 64) k_sched<2><<<W, 32, b->smem_sched, ss>>>(dk, G, taps);
    else if (d.NB <= 128) k_sched<4><<<W, 32, b->smem_sched, ss>>>(dk, G, taps);
    else k_sched<8><<<W, 32, b->smem_sched, ss>>>(dk, G, taps);
    if (split) {
      cudaEventRecord(b->sev[1], ss);
      OB_LAUNCH_PREP_G(2)
      cudaStreamWaitEvent(st, b->sev[1], 0);
      g_launches++;
    }
#undef OB_LAUNCH_PREP_G
#undef OB_LAUNCH_PREP
    if (timing) cudaEventRecord(ev[3], st);
    if (b->sor_lane) {
      // whole waves of a size whose row records stay in the L2 between the iterations (OB_SOR_LANE_GRID overrides)
      int gl_ = (W + 31) / 32;
      { static const char *e = getenv("OB_SOR_LANE_GRID"); if (e && atoi(e) > 0 && atoi(e) < gl_) { const int waves = (gl_ + atoi(e) - 1) / atoi(e); gl_ = (gl_ + waves - 1) / waves; } }
      k_sor_lane<<<gl_, 32, b->smem_sor_lane, st>>>(d, taps);
    } else if (b->sor_reg) {
      k_sor_reg<G><<<gsor, 32, b->smem_sor_reg, st>>>(d, taps);
    } else if (b->sor_pair) {
      constexpr int GP = G <= 16 ? 2 * G : 32, Tp = 32 / GP;
      const int gp = (W + Tp - 1) / Tp;
      long long cap = b->pair_resident > 0 ? b->pair_resident : gp;
      { static const char *e = getenv("OB_GRID_SOR"); if (e && atoi(e) > 0) cap = atoi(e); }
      const int waves = (int)((gp + cap - 1) / cap);
      const int gpair = (gp + waves - 1) / waves;
      if (G <= 16) {
        if (b->sor_pair == 5) k_sor_pair<GP, 5><<<gpair, 32, b->smem_sor_pair, st>>>(d, taps);
        else k_sor_pair<GP, 4><<<gpair, 32, b->smem_sor_pair, st>>>(d, taps);
      }
    } else if (b->sor_ring) {
      // persistent grid: as many CTAs as stay resident, fewer when the rows of the worlds in flight would not fit the L2
      // (rows/world measured from the counters; capacity-based guess before the first read-back), whole waves
      const double rows_w = b->avg_rows > 0 ? b->avg_rows : 0.65 * d.NR;
      const double world_bytes = rows_w * OB_ROWW * sizeof(real) + 1.0;
      long long cap = (long long)(b->l2_target_bytes / world_bytes) / T;
      if (cap > b->ring_resident) cap = b->ring_resident;
      if (cap < 1) cap = 1;
      { static const char *e = getenv("OB_GRID_SOR"); if (e && atoi(e) > 0) cap = atoi(e); }
      const int waves = (int)((gstep + cap - 1) / cap);
      const int gring = (gstep + waves - 1) / waves;
      if (b->ring_depth == 6) k_sor_ring<G, 6><<<gring, 32, b->smem_sor_ring, st>>>(d, taps);
      else k_sor_ring<G, 4><<<gring, 32, b->smem_sor_ring, st>>>(d, taps);
    } else if (b->sor_deep) k_sor<G, true><<<gsor, 32, b->smem_sor, st>>>(d, taps);
    else k_sor<G, false><<<gsor, 32, b->smem_sor, st>>>(d, taps);
    if (timing) cudaEventRecord(ev[4], st);
    { const int pt = b->post_tile, gpost = (W + (32 / pt) - 1) / (32 / pt);
      if (pt == 4) k_post<4><<<gpost, 32, b->smem_post, st>>>(d, h);
      else if (pt == 8) k_post<8><<<gpost, 32, b->smem_post, st>>>(d, h);
      else if (pt == 16) k_post<16><<<gpost, 32, b->smem_post, st>>>(d, h);
      else k_post<32><<<gpost, 32, b->smem_post, st>>>(d, h); }
  g_launches += 4;
}
void obk_stepk_launch(ObBackend *b, const ObBatchDev &d, real h, int taps, int W, cudaStream_t st, cudaEvent_t *ev, bool timing) {
  if (b->tile == 4) stepk_launch_t<4>(b, d, h, taps, W, st, ev, timing);
  else if (b->tile == 8) stepk_launch_t<8>(b, d, h, taps, W, st, ev, timing);
  else if (b->tile == 16) stepk_launch_t<16>(b, d, h, taps, W, st, ev, timing);
  else stepk_launch_t<32>(b, d, h, taps, W, st, ev, timing);
}

#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line.

metric : body-steps/sec (and contacts solved/sec) of the batched hot path
         dSpaceCollide + contact policy + dWorldQuickStep(20 it) + dJointGroupEmpty
workload: configs[1] — 4096 independent worlds x (plane + 32-box stack + 8 spheres),
         scene `stack32` of tests/harness/scenes.h, settled for SETTLE steps first so the
         timed steps see the resting pile (~130 contacts / ~400 rows per world).
A "step" = one pass of that path over all worlds of this rank (weak scaling: every rank
owns its own 4096 worlds, no data-path collective; NCCL only reduces the metric).

  python bench.py --gpus N --steps K --warmup W            (torchrun for N>1)
  python bench.py --impl reference ...                      the reference's own CPU code
                                                            (oracle/_ref) on the host cores
value  : device-resident throughput (CUDA events on the library's launching stream).
e2e    : same metric through the C ABI with HOST buffers every step: dBatchAddForces
         (H2D, pinned staging) + step + dBatchGetBodyState (D2H).
roofline: dominant kernel k_step, algorithmic bytes (SURVEY.md §8d / DESIGN.md) / its
         CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(ROOT, "ode-0.12_b200", "lib")
# BASELINE.json configs that shard over independent worlds.  The headline (default) is configs[1];
# --config 3 / 4 time configs[2] / configs[3] with the same harness; --config 1 is configs[0], one small world
# (latency only; the reference runs it on ONE host core); --config 5 is the large single world.
CONFIGS = {
    1: dict(scene="block64", worlds=1, settle=100, cap=1024, geoms=65, cpu=(1, 1, 300),
            desc="configs[0]: {W} world/GPU: 64 boxes in a 4x4x4 block on a plane, dHashSpace, boxstack contact policy maxc 8 "
                 "(one island of ~1850 rows: a LATENCY figure, one CTA of one GPU works; N > 1 = replicas only)"),
    2: dict(scene="stack32", worlds=4096, settle=300, cap=192, geoms=41,
            desc="configs[1]: {W} independent worlds/GPU x (plane + 32-box stack + 8 spheres), dHashSpace, boxstack contact policy maxc 8"),
    3: dict(scene="buggy_terrain256", worlds=65536, settle=200, cap=48, geoms=6,
            desc="configs[2]: {W} independent worlds/GPU x (hinge2 buggy: box chassis + 4 sphere wheels) on one shared 256x256-vertex "
                 "trimesh terrain (130050 triangles, OPCODE-equivalent colliders), dHashSpace, buggy contact policy maxc 10"),
    4: dict(scene="ragdoll", worlds=16384, settle=150, cap=160, geoms=37,
            desc="configs[3]: {W} independent worlds/GPU x (20-link capsule/box chain on 19 ball/hinge joints + 16-box pile), dHashSpace, "
                 "crash contact policy maxc 4"),
}
H = 0.01
# algorithmic bytes per unit, dSINGLE: A body-step, B geom-step, C contact,
# D row x SOR iteration, ASM row assembly
A_B, B_B, C_B, D_B, ASM_B = 136, 80, 128, 224, 128
ITERS = 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        """the timed region starts now (the sampler has been running since before the settle / warm-up steps: nvidia-smi needs
        longer to start than a short timed region lasts)"""
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        t1 = time.perf_counter()
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.12]
        window = "timed region"
        if not rows:   # a timed region shorter than the sampling period: the last samples of the same load (settle + warm-up steps) before it
            rows = [r for (t, r) in self.rows if t <= t1 + 0.12][-5:]
            window = "settle + warm-up steps of the same workload, up to the end of the timed region (region shorter than the 100 ms sampling period)"
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def load_libs():
    lib_path = os.path.join(LIBDIR, "libode_b200_single.so")
    if not os.path.exists(lib_path):
        raise SystemExit("libode_b200_single.so is missing: run `python __graft_entry__.py` (there is no CPU fallback)")
    lib = ctypes.CDLL(lib_path, mode=ctypes.RTLD_GLOBAL)
    scenes = ctypes.CDLL(os.path.join(LIBDIR, "libob_scenes_single.so"))
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    scenes.ob_scene_build_batch.restype = vp
    scenes.ob_scene_build_batch.argtypes = [ctypes.c_char_p, ci, ci, ci, ci]
    lib.dB200LastError.restype = ctypes.c_char_p
    lib.dBatchCollideAndQuickStep.argtypes = [vp, cf, ci, vp]
    lib.dBatchTimerStart.argtypes = [vp]
    lib.dBatchTimerStop.argtypes = [vp, ctypes.POINTER(cf)]
    lib.dBatchSetDebugTaps.argtypes = [vp, ci]
    lib.dBatchSetKernelTiming.argtypes = [vp, ci]
    lib.dBatchGetKernelTimes.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong), ci]
    lib.dBatchKernelName.restype = ctypes.c_char_p
    lib.dBatchKernelName.argtypes = [ci]
    lib.dBatchGetCounters.argtypes = [vp, vp]
    lib.dBatchResetCounters.argtypes = [vp]
    lib.dBatchNumBodies.argtypes = [vp]
    lib.dBatchGetBodyState.argtypes = [vp, vp, vp, vp, vp]
    lib.dBatchAddForces.argtypes = [vp, vp, vp]
    lib.dBatchDestroy.argtypes = [vp]
    lib.dB200KernelLaunchCount.restype = ctypes.c_longlong
    return lib, scenes


def counters(lib, B):
    c = (ctypes.c_longlong * 7)()
    lib.dBatchGetCounters(B, ctypes.byref(c))
    return dict(zip(["steps", "body_steps", "pairs", "contacts", "rows", "islands", "overflow_worlds"], list(c)))


def shard(rank, world_size, per_gpu):
    """world ids owned by `rank`: contiguous ranges, no overlap (weak scaling)"""
    return rank * per_gpu, per_gpu


def reduce_metrics(dist, vals, sums, device):
    """the only inter-rank traffic of the benchmark: MAX of the timings, SUM of the counters (worlds are
    independent, so the data path has no collective).  dist=None: single process."""
    if dist is None:
        return vals, sums
    import torch

    tv = torch.tensor(vals, device=device); ts = torch.tensor(sums, device=device)
    dist.all_reduce(tv, op=dist.ReduceOp.MAX); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    return tv.cpu().numpy(), ts.cpu().numpy()


class Ctx:
    """process-wide state of one bench run: rank layout, the NCCL group (N > 1) and the loaded libraries"""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world_size > 1:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.lib, self.scenes = load_libs()

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def measure_batched(ctx, cfg_id, steps, warmup, worlds=0, strong=False, cpu=False, args=None):
    """one configuration of independent worlds (CONFIGS[cfg_id]) on this rank's GPU; returns the JSON dict on rank 0.
    strong=True: `worlds` is the TOTAL over all ranks (fixed work, each rank owns worlds/N); else worlds per GPU."""
    cfg = CONFIGS[cfg_id]
    SCENE, SETTLE, CONTACTS_CAP, GEOMS_PER_WORLD, CONFIG_DESC = cfg["scene"], cfg["settle"], cfg["cap"], cfg["geoms"], cfg["desc"]
    rank, world_size, local_rank, dist, lib, scenes = ctx.rank, ctx.world_size, ctx.local_rank, ctx.dist, ctx.lib, ctx.scenes
    per_gpu = worlds if worlds > 0 else cfg["worlds"]
    if strong:
        per_gpu = (per_gpu + world_size - 1) // world_size
    world0, nworlds = shard(rank, world_size, per_gpu)
    B = scenes.ob_scene_build_batch(SCENE.encode(), nworlds, world0, CONTACTS_CAP, local_rank)
    if not B:
        raise SystemExit("batch creation failed: " + (lib.dB200LastError() or b"").decode())
    B = ctypes.c_void_p(B)
    lib.dBatchSetDebugTaps(B, 0)

    def step(n):
        if lib.dBatchCollideAndQuickStep(B, H, n, None) != 0:
            raise SystemExit("step failed: " + lib.dB200LastError().decode())

    def barrier():
        if dist is not None:
            import torch

            torch.cuda.synchronize()
            dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    step(SETTLE)                       # untimed: let the pile come to rest
    step(max(warmup, 3))          # warm-up steps proper
    lib.dBatchResetCounters(B)
    l0 = lib.dB200KernelLaunchCount()
    barrier()
    sampler.begin()
    ms = ctypes.c_float()
    lib.dBatchTimerStart(B)
    step(steps)
    lib.dBatchTimerStop(B, ctypes.byref(ms))
    barrier()
    clocks = sampler.stop()
    launches = lib.dB200KernelLaunchCount() - l0
    c = counters(lib, B)
    elapsed_ms = float(ms.value)

    # per-kernel attribution for the roofline (separate pass, events around every launch)
    lib.dBatchSetKernelTiming(B, 1)
    lib.dBatchResetCounters(B)
    step(steps)
    NK = 8
    kms = (ctypes.c_double * NK)()
    kl = (ctypes.c_longlong * NK)()
    nk = lib.dBatchGetKernelTimes(B, kms, kl, NK)
    lib.dBatchSetKernelTiming(B, 0)
    ck = counters(lib, B)
    names = [lib.dBatchKernelName(k).decode() for k in range(nk)]
    kt = {names[k]: kms[k] / max(kl[k], 1) * 1e-3 for k in range(nk)}      # seconds per launch
    nl = max(kl[0], 1)
    ng_per_world = GEOMS_PER_WORLD
    # algorithmic bytes per launch (SURVEY 8d terms): collide = geoms (B) + contacts written (C/2);
    # prep = bodies read (A/2) + contacts read (C/2) + row assembly; sor = D x iterations per row;
    # post = bodies read+written (A/2)
    kbytes = {
        "k_collide": (B_B * ng_per_world * nworlds * nl + (C_B // 2) * ck["contacts"]) / nl,
        "k_prep": ((A_B // 2) * ck["body_steps"] + (C_B // 2) * ck["contacts"] + ASM_B * ck["rows"]) / nl,
        "k_sched": (8 * 3 * ck["rows"]) / nl,      # row meta read + sched/pstart written per shuffle epoch
        "k_sor": (D_B * ITERS * ck["rows"]) / nl,
        "k_post": ((A_B // 2) * ck["body_steps"]) / nl,
    }
    t_all = sum(kt.values())
    dom = max(kt, key=kt.get)
    peak, peak_kind = peaks()
    step_bytes = kbytes[dom]
    t_step = kt[dom]
    achieved = step_bytes / t_step / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tf):
        tj = json.load(open(tf))
        # the capture is of configs[1] (stack32 x 4096 worlds, the bench's capacities): only that workload may quote it
        if str(tj.get("kernel", "")).startswith(dom) and SCENE == "stack32" and nworlds == 4096:
            traffic = tj.get("dram_bytes_per_launch")

    # end to end through the C ABI with HOST buffers every step (page-locked, from dBatchHostAlloc):
    # external forces H2D -> step -> body state D2H.  The force samples are generated before the
    # timed region (they stand for the caller's controller output, which is not the library's work).
    nb = lib.dBatchNumBodies(B)
    lib.dBatchHostAlloc.restype = ctypes.c_void_p
    lib.dBatchHostAlloc.argtypes = [ctypes.c_size_t]

    def pinned(shape):
        n = int(np.prod(shape))
        ptr = lib.dBatchHostAlloc(n * 4)
        if not ptr:
            raise SystemExit("dBatchHostAlloc failed")
        a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n,)).reshape(shape)
        a[...] = 0
        return a

    pos = pinned((nworlds, nb, 3)); quat = pinned((nworlds, nb, 4))
    lv = pinned((nworlds, nb, 3)); av = pinned((nworlds, nb, 3))
    torque = pinned((nworlds, nb, 3))
    rng = np.random.default_rng(rank)
    forces = [pinned((nworlds, nb, 3)) for _ in range(4)]
    for f in forces:
        f[:, :, 0] = 0.01 * rng.standard_normal((nworlds, nb), dtype=np.float32)
    force = forces[0]
    lib.dBatchResetCounters(B)
    e2e_steps = steps
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        lib.dBatchAddForces(B, forces[s & 3].ctypes.data, torque.ctypes.data)
        step(1)
        lib.dBatchGetBodyState(B, pos.ctypes.data, quat.ctypes.data, lv.ctypes.data, av.ctypes.data)
    t_e2e = time.perf_counter() - t0
    ce = counters(lib, B)
    assert np.isfinite(pos).all()

    vals = np.array([elapsed_ms, t_e2e], dtype=np.float64)
    sums = np.array([c["body_steps"], c["contacts"], ce["body_steps"], c["rows"], launches, c["overflow_worlds"]], dtype=np.float64)
    vals, sums = reduce_metrics(dist, vals, sums, "cuda")
    out = None
    if rank == 0:
        t = vals[0] * 1e-3
        out = {
            "metric": "body-steps/sec (batched dSpaceCollide + dWorldQuickStep, 20 SOR iterations)",
            "value": sums[0] / t, "unit": "body-steps/s",
            "contacts_solved_per_sec": sums[1] / t,
            "n_gpus": world_size, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": vals[0] / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_DESC.format(W=per_gpu) + f", quickstep 20 it, h={H}, settled {SETTLE} steps",
                       "worlds_per_gpu": per_gpu, "bodies_per_world": nb, "rows_per_world_step": sums[3] / max(c["steps"] * world_size, 1),
                       "contacts_per_world_step": sums[1] / max(c["steps"] * world_size, 1),
                       "cache": (lambda mb: ("per-step working set (rows written+read) ~%.0f MB/GPU > 126 MB L2" % mb) if mb > 126 else
                                 ("per-step working set ~%.1f MB fits the L2 and is not flushed: this configuration measures the latency of "
                                  "one island's dependent row updates, not a stream" % mb))(c["rows"] / steps * 128 * 2 / 1e6 + 70 * nworlds / 4096),
                       "precision": "dSINGLE", "parity": "bit-exact vs reference (tests/)"},
            "e2e": {"value": sums[2] / vals[1], "unit": "body-steps/s", "h2d_bytes_per_step": int(force.nbytes + torque.nbytes) * world_size,
                    "d2h_bytes_per_step": int(pos.nbytes + quat.nbytes + lv.nbytes + av.nbytes) * world_size},
            "gpu_launches": int(sums[4]),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": step_bytes,
                         "kernel_ms": t_step * 1e3, "kernel_share_of_step": t_step / t_all,
                         "whole_step": {"algorithmic_bytes": sum(kbytes.values()),
                                        "achieved": sum(kbytes.values()) / (vals[0] * 1e-3 / steps) / 1e9 if world_size == 1 else None},
                         "kernels": {k: {"ms": kt[k] * 1e3, "algorithmic_bytes_per_launch": kbytes.get(k), "achieved": kbytes.get(k, 0) / kt[k] / 1e9}
                                     for k in kt}},
            "clocks": clocks,
            "overflow_worlds": int(sums[5]),
        }
        if world_size == 1 and cpu:
            out["cpu_baseline"] = cpu_baseline(cfg)
    lib.dBatchDestroy(B)
    return out


LARGE = dict(scene="pile_100x100x20", settle=300, cpu_scene="pile_24x24x20", cpu_settle=200, cpu_steps=10,
             desc="configs[4]: one world, {NB} bodies (50/50 spheres r 0.25 / boxes 0.5^3) piled in a walled box, dSweepAndPruneSpace, "
                  "crash contact policy maxc 4, graph-coloured SOR")
PHASES = ["geoms+sort", "pair sweep", "narrowphase", "colouring", "row assembly", "sor", "integration"]


class LargeStats(ctypes.Structure):
    _fields_ = [("pairs", ctypes.c_int), ("contacts", ctypes.c_int), ("contact_pairs", ctypes.c_int), ("solved_contacts", ctypes.c_int),
                ("colours", ctypes.c_int), ("colouring_rounds", ctypes.c_int), ("sor_launches", ctypes.c_int), ("steps_timed", ctypes.c_int),
                ("phase_ms", ctypes.c_double * 7)]


def measure_large(ctx, steps, warmup, scene="", settle=-1, cpu=False):
    """configs[4]: one large world.  N = 1: one GPU.  N > 1 (torchrun, one process per GPU): every rank holds the whole
    world; the pair sweep and the narrowphase are divided over the ranks by SAP-sorted position and their results written into every
    rank's arrays over NVLink (dBatchSplitExport / dBatchSplitAttach) -- total work is fixed, "scaling": "strong"; NCCL only gathers
    the 128-byte buffer descriptions once and reduces the timing.  value = device-resident body-steps/s (max over ranks
    of the CUDA-event time); roofline on the SOR phase, algorithmic bytes D x iterations per row."""
    rank, world_size, local_rank, dist, lib, scenes = ctx.rank, ctx.world_size, ctx.local_rank, ctx.dist, ctx.lib, ctx.scenes
    lib.dBatchGetLargeWorldStats.argtypes = [ctypes.c_void_p, ctypes.POINTER(LargeStats)]
    scene = scene or LARGE["scene"]
    B = scenes.ob_scene_build_batch(scene.encode(), 1, 0, 0, local_rank)
    if not B:
        raise SystemExit("batch creation failed: " + (lib.dB200LastError() or b"").decode())
    B = ctypes.c_void_p(B)
    lib.dBatchSetDebugTaps(B, 0)
    status = np.zeros(1, dtype=np.int32)

    def step(n):
        if lib.dBatchCollideAndQuickStep(B, H, n, status.ctypes.data) != 0:
            raise SystemExit("step failed: " + lib.dB200LastError().decode())
        if status[0]:
            raise SystemExit(f"capacity overflow, status {status[0]}")

    def barrier():
        if dist is not None:
            import torch

            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        tv = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv.item())

    if dist is not None:
        import torch

        lib.dBatchSplitExport.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.dBatchSplitAttach.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        hb = ctypes.create_string_buffer(128)
        if lib.dBatchSplitExport(B, hb) != 0:
            raise SystemExit("split export failed: " + lib.dB200LastError().decode())
        mine = torch.frombuffer(bytearray(hb.raw), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(world_size)]
        dist.all_gather(allh, mine)
        blob = b"".join(t_.cpu().numpy().tobytes() for t_ in allh)
        if lib.dBatchSplitAttach(B, rank, world_size, blob) != 0:
            raise SystemExit("split attach failed: " + lib.dB200LastError().decode())
        barrier()
    settle = settle if settle >= 0 else LARGE["settle"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    step(settle)
    step(max(warmup, 3))
    lib.dBatchResetCounters(B)
    l0 = lib.dB200KernelLaunchCount()
    barrier()
    sampler.begin()
    ms = ctypes.c_float()
    lib.dBatchTimerStart(B)
    step(steps)
    lib.dBatchTimerStop(B, ctypes.byref(ms))
    barrier()
    clocks = sampler.stop()
    launches = lib.dB200KernelLaunchCount() - l0
    c = counters(lib, B)
    t = max_over_ranks(float(ms.value)) * 1e-3
    # per-phase attribution (CUDA events between the phases, separate pass)
    lib.dBatchSetKernelTiming(B, 1)
    lib.dBatchResetCounters(B)
    step(steps)
    st = LargeStats()
    lib.dBatchGetLargeWorldStats(B, ctypes.byref(st))
    lib.dBatchSetKernelTiming(B, 0)
    ck = counters(lib, B)
    nst = max(st.steps_timed, 1)
    ph = {PHASES[k]: st.phase_ms[k] / nst for k in range(7)}
    nb = lib.dBatchNumBodies(B)
    rows_per_step = 3 * ck["contacts"] / nst
    sor_bytes = D_B * ITERS * rows_per_step
    peak, peak_kind = peaks()
    step_bytes = (A_B * nb + B_B * (nb + 5) + C_B * ck["contacts"] / nst + (D_B * ITERS + ASM_B) * rows_per_step)
    achieved = sor_bytes / (ph["sor"] * 1e-3) / 1e9
    # end to end with host buffers
    lib.dBatchHostAlloc.restype = ctypes.c_void_p
    lib.dBatchHostAlloc.argtypes = [ctypes.c_size_t]

    def pinned(shape):
        n = int(np.prod(shape))
        ptr = lib.dBatchHostAlloc(n * 4)
        a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n,)).reshape(shape)
        a[...] = 0
        return a

    pos = pinned((nb, 3)); quat = pinned((nb, 4)); lv = pinned((nb, 3)); av = pinned((nb, 3)); torque = pinned((nb, 3))
    rng = np.random.default_rng(0)
    forces = [pinned((nb, 3)) for _ in range(4)]
    for f in forces:
        f[:, 0] = 0.01 * rng.standard_normal(nb, dtype=np.float32)
    lib.dBatchResetCounters(B)
    t0 = time.perf_counter()
    for s in range(steps):
        lib.dBatchAddForces(B, forces[s & 3].ctypes.data, torque.ctypes.data)
        step(1)
        lib.dBatchGetBodyState(B, pos.ctypes.data, quat.ctypes.data, lv.ctypes.data, av.ctypes.data)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    ce = counters(lib, B)
    assert np.isfinite(pos).all()
    if dist is not None:
        # every rank must hold the same world after the same steps: compare a checksum of the body state
        import torch

        h64 = int(np.frombuffer(np.ascontiguousarray(pos).tobytes(), dtype=np.uint32).astype(np.uint64).sum() & 0x7FFFFFFFFFFFFFFF)
        hv = torch.tensor([h64, -h64], dtype=torch.int64, device="cuda")
        dist.all_reduce(hv, op=dist.ReduceOp.MAX)
        if int(hv[0].item()) != -int(hv[1].item()):
            raise SystemExit("split ranks disagree on the body state")
        if rank != 0:
            lib.dBatchDestroy(B)
            return None
    out = {
        "metric": "body-steps/sec (dSpaceCollide + dWorldQuickStep, 20 SOR iterations)", "value": c["body_steps"] / t, "unit": "body-steps/s",
        "contacts_solved_per_sec": c["contacts"] / t, "n_gpus": world_size, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": t * 1e3 / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": LARGE["desc"].format(NB=nb) + f", scene {scene}, quickstep 20 it, h={H}, settled {settle} steps",
                   "bodies": nb, "pairs_per_step": st.pairs, "contacts_per_step": st.contacts, "rows_per_step": rows_per_step,
                   "colours": st.colours, "colouring_rounds": st.colouring_rounds, "sor_launches_per_step": st.sor_launches,
                   "parallelism": "one GPU" if world_size == 1 else (
                       f"front end split over {world_size} GPUs: every rank sweeps and collides the pairs of its share of the SAP-sorted positions and "
                       "stores counts / pairs / contacts into every rank's arrays through NVLink peer mappings (k_lw_sweep / k_lw_narrow, two flag "
                       "barriers per step); sort, colouring, rows, SOR (latency-bound: one wave per colour) and integration run on every rank"
                       if not os.environ.get("OB_LW_SPLIT_SOR") else
                       f"front end + SOR phase split over {world_size} GPUs (fc exchange inside k_lw_sor_split, flag barrier per colour)"),
                   "nvlink_bytes_per_step_per_rank": 0 if world_size == 1 else int(
                       (world_size - 1) * (st.pairs * 16 + st.contacts * 48 + 8 * nb) / world_size),
                   "cache": "rows %.0f MB/step streamed every iteration (> 126 MB L2 at full size)" % (rows_per_step * 80 / 1e6),
                   "precision": "dSINGLE", "parity": "pairs+contacts exact, state within stated tolerance vs reference; bitwise vs sequential mirror (tests/test_large_world.py)"},
        "e2e": {"value": ce["body_steps"] / t_e2e, "unit": "body-steps/s", "h2d_bytes_per_step": int(forces[0].nbytes + torque.nbytes),
                "d2h_bytes_per_step": int(pos.nbytes + quat.nbytes + lv.nbytes + av.nbytes)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_lw_sor (x%d launches/step)" % st.sor_launches, "achieved": achieved, "peak": peak,
                     "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": sor_bytes / max(st.sor_launches, 1), "kernel_ms": ph["sor"] / max(st.sor_launches, 1),
                     "kernel_share_of_step": ph["sor"] / max(sum(ph.values()), 1e-9),
                     "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (t / steps) / 1e9},
                     "phases_ms": ph},
        "clocks": clocks, "overflow_worlds": c["overflow_worlds"],
    }
    if cpu and world_size == 1:
        exe = os.path.join(ROOT, "oracle", "_ref", "driver_ref_single")
        if os.path.exists(exe):
            r = json.loads(subprocess.run([exe, "--scene", LARGE["cpu_scene"], "--steps", str(LARGE["cpu_steps"]), "--settle", str(LARGE["cpu_settle"]),
                                           "--time"], capture_output=True, text=True, timeout=1800).stdout.strip().splitlines()[-1])
            out["cpu_baseline"] = {"value": r["body_steps_per_sec"], "unit": "body-steps/s", "cores": 1, "kind": "reference",
                                   "contacts_solved_per_sec": r["contacts_per_sec"],
                                   "sample": f"1 process (one world cannot use more), scene {LARGE['cpu_scene']} (same pile, smaller footprint), "
                                             f"{LARGE['cpu_settle']} settle steps untimed + {LARGE['cpu_steps']} timed"}
    lib.dBatchDestroy(B)
    return out


def cpu_run(scene, nproc, worlds_each, steps, settle, timeout=1700):
    """the unmodified reference (oracle/_ref/driver_ref_single: classic dSpaceCollide + near callback + dWorldQuickStep
    loop over its worlds) in nproc processes, worlds [i*worlds_each, (i+1)*worlds_each) each"""
    exe = os.path.join(ROOT, "oracle", "_ref", "driver_ref_single")
    if not os.path.exists(exe):
        return None
    procs = [subprocess.Popen([exe, "--scene", scene, "--worlds", str(worlds_each), "--world0", str(i * worlds_each),
                               "--steps", str(steps), "--settle", str(settle), "--time"], stdout=subprocess.PIPE, text=True)
             for i in range(nproc)]
    outs = [json.loads(p.communicate(timeout=timeout)[0].strip().splitlines()[-1]) for p in procs]
    secs = max(o["seconds"] for o in outs)
    return {"body_steps": sum(o["body_steps"] for o in outs), "contacts": sum(o["contacts"] for o in outs), "seconds": secs}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(cfg):
    """cpu_baseline leg of the b200 arm: a BOUNDED sample of the same workload on this box's host cores (one process per
    core -- ODE keeps process-global state -- x 8 worlds, 100 timed steps after the settle)"""
    nproc, worlds_each, steps = cfg.get("cpu") or (host_cores(), 8, 100)
    r = cpu_run(cfg["scene"], nproc, worlds_each, steps, cfg["settle"])
    if r is None:
        return {"value": None, "unit": "body-steps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not present"}
    return {"value": r["body_steps"] / r["seconds"], "unit": "body-steps/s", "cores": nproc, "kind": "reference",
            "contacts_solved_per_sec": r["contacts"] / r["seconds"],
            "sample": f"{nproc} processes x {worlds_each} worlds of {cfg['scene']}, {cfg['settle']} settle steps untimed + {steps} timed steps"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on ALL host cores over the SAME configuration as the b200
    arm: N x worlds_per_gpu worlds (N = WORLD_SIZE), dealt over one process per core.  A timed window shorter than about a
    second is noise on a shared host, so the K timed steps are repeated back to back (R x K steps in one timed region)
    until the window is >= 1 s; value = body-steps of the whole window / its duration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[args.config]
    per_gpu = args.worlds if args.worlds > 0 else cfg["worlds"]
    total = per_gpu * n_gpus
    nproc = min(host_cores(), total)
    worlds_each = (total + nproc - 1) // nproc
    K = max(args.steps, 1)
    warm = max(args.warmup, 3)
    # repeats from a quick calibration: 2 worlds per process, 20 steps of the settled scene
    cal = cpu_run(cfg["scene"], nproc, min(2, worlds_each), 20, cfg["settle"])
    if cal is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference build) not present on this box"}))
        return
    est_step = cal["seconds"] / 20 * worlds_each / min(2, worlds_each)
    R = max(1, int(np.ceil(1.2 / max(est_step * K, 1e-9))))
    R = min(R, 50)
    r = cpu_run(cfg["scene"], nproc, worlds_each, K * R, cfg["settle"] + warm)
    v = r["body_steps"] / r["seconds"]
    sample = (f"FULL configuration: {nproc} processes x {worlds_each} worlds = {nproc * worlds_each} worlds of {cfg['scene']} "
              f"({n_gpus} x {per_gpu}), {cfg['settle']}+{warm} untimed steps, {R} x {K} = {K * R} timed steps in one {r['seconds']:.2f} s window")
    print(json.dumps({
        "impl": "reference", "metric": "body-steps/sec (batched dSpaceCollide + dWorldQuickStep, 20 SOR iterations)",
        "value": v, "unit": "body-steps/s", "contacts_solved_per_sec": r["contacts"] / r["seconds"],
        "n_gpus": n_gpus, "steps": K, "warmup": warm,
        "ms_per_step": r["seconds"] * 1e3 / (K * R), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"].format(W=per_gpu) + f", quickstep 20 it, h={H}, settled {cfg['settle']} steps",
                   "worlds_per_gpu": per_gpu, "worlds_total": nproc * worlds_each, "timed_steps_total": K * R,
                   "precision": "dSINGLE", "implementation": "reference ODE 0.12 (oracle/_ref, unmodified) on the host cores"},
        "cpu_baseline": {"value": v, "unit": "body-steps/s", "cores": nproc, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def brief(d):
    """the part of a configuration's line that other_configs carries"""
    if d is None:
        return None
    r = d.get("roofline", {})
    out = {k: d[k] for k in ("value", "unit", "ms_per_step", "n_gpus", "steps", "warmup", "scaling", "contacts_solved_per_sec", "overflow_worlds") if k in d}
    out["workload"] = d["config"]["workload"]
    out["e2e"] = d["e2e"]["value"]
    ws = r.get("whole_step", {})
    out["whole_step_algorithmic_GBps"] = ws.get("achieved")
    out["whole_step_frac_of_hbm_peak"] = (ws["achieved"] / r["peak"]) if ws.get("achieved") and r.get("peak") else None
    if "kernels" in r:
        out["kernels_ms"] = {k: v["ms"] for k, v in r["kernels"].items()}
    if "phases_ms" in r:
        out["phases_ms"] = r["phases_ms"]
    for k in ("rows_per_world_step", "contacts_per_world_step", "bodies", "contacts_per_step", "colours", "parallelism", "nvlink_bytes_per_step_per_rank"):
        if k in d["config"]:
            out[k] = d["config"][k]
    if "cpu_baseline" in d:
        out["cpu_baseline"] = d["cpu_baseline"]
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds", type=int, default=0)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS) + [5])
    ap.add_argument("--scene", default="")
    ap.add_argument("--settle", type=int, default=-1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the other_configs block of the default line")
    ap.add_argument("--strong", action="store_true", help="--worlds is the total over all GPUs (fixed work)")
    a = ap.parse_args()
    if a.impl == "reference":
        if a.config == 5:
            raise SystemExit("--config 5 reports the reference inside its own line (cpu_baseline)")
        run_reference(a)
        sys.exit(0)
    ctx = Ctx()
    if a.config == 5:
        out = measure_large(ctx, a.steps, a.warmup, a.scene, a.settle, cpu=not a.no_cpu)
    else:
        out = measure_batched(ctx, a.config, a.steps, a.warmup, a.worlds, strong=a.strong, cpu=not a.no_cpu)
        if a.config == 2 and not a.no_other and not a.strong and a.worlds <= 0:
            # the other BASELINE.json configurations ride in the same line (driver-measured): shorter runs of the same harness.
            # N > 1: configs[1] again as STRONG scaling (the named 4096 worlds over N GPUs) and configs[4] split over the GPUs.
            ks, kw = max(10, a.steps // 3), 3
            other = {}
            if ctx.world_size > 1:
                other["configs[1] strong (4096 worlds total)"] = brief(measure_batched(ctx, 2, ks, kw, CONFIGS[2]["worlds"], strong=True))
            other["configs[2]"] = brief(measure_batched(ctx, 3, ks, kw))
            other["configs[3]"] = brief(measure_batched(ctx, 4, ks, kw))
            other["configs[4]"] = brief(measure_large(ctx, ks, kw, cpu=(not a.no_cpu and ctx.world_size == 1)))
            if ctx.world_size == 1:
                other["configs[0]"] = brief(measure_batched(ctx, 1, 200, 5, cpu=not a.no_cpu))
            if out is not None:
                out["other_configs"] = other
    if out is not None:
        print(json.dumps(out))
    ctx.close()

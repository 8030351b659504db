"""Drive a large-world batch split over several ranks (dBatchSplitExport / dBatchSplitAttach, include/ode_b200/ode.h)
from Python: used by tests/test_split_world.py (ranks = host threads of one process: hostsim mirror on the CPU,
loop-back on one GPU, one device per rank on a multi-GPU box) and by tests/harness/split_rank.py (one process per GPU)."""
import ctypes
import os
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HANDLE_BYTES = 128


def load(backend, prec="single"):
    """(lib, scenes) of the product ("b200") or the test-only host mirror ("hostsim")"""
    if backend == "b200":
        d = os.path.join(ROOT, "ode-0.12_b200", "lib")
        lib = ctypes.CDLL(os.path.join(d, f"libode_b200_{prec}.so"), mode=ctypes.RTLD_GLOBAL)
        scenes = ctypes.CDLL(os.path.join(d, f"libob_scenes_{prec}.so"))
    else:
        d = os.path.join(ROOT, "tests", "hostsim", "_build")
        lib = ctypes.CDLL(os.path.join(d, f"libode_b200_hostsim_{prec}.so"), mode=ctypes.RTLD_GLOBAL)
        scenes = ctypes.CDLL(os.path.join(d, f"libob_scenes_hostsim_{prec}.so"))
    vp, ci = ctypes.c_void_p, ctypes.c_int
    real = ctypes.c_float if prec == "single" else ctypes.c_double
    scenes.ob_scene_build_batch.restype = vp
    scenes.ob_scene_build_batch.argtypes = [ctypes.c_char_p, ci, ci, ci, ci]
    lib.dB200LastError.restype = ctypes.c_char_p
    lib.dBatchCollideAndQuickStep.argtypes = [vp, real, ci, vp]
    lib.dBatchNumBodies.argtypes = [vp]
    lib.dBatchGetBodyState.argtypes = [vp, vp, vp, vp, vp]
    lib.dBatchSplitExport.argtypes = [vp, vp]
    lib.dBatchSplitAttach.argtypes = [vp, ci, ci, vp]
    lib.dBatchDestroy.argtypes = [vp]
    lib.dBatchSetDebugTaps.argtypes = [vp, ci]
    lib.np_real = np.float32 if prec == "single" else np.float64
    return lib, scenes


def build(lib, scenes, scene, device=0):
    B = scenes.ob_scene_build_batch(scene.encode(), 1, 0, 0, device)
    if not B:
        raise RuntimeError("batch creation failed: " + (lib.dB200LastError() or b"").decode())
    B = ctypes.c_void_p(B)
    lib.dBatchSetDebugTaps(B, 0)
    return B


def export(lib, B):
    h = ctypes.create_string_buffer(HANDLE_BYTES)
    if lib.dBatchSplitExport(B, h) != 0:
        raise RuntimeError("export failed: " + lib.dB200LastError().decode())
    return h.raw


def attach(lib, B, rank, handles):
    blob = b"".join(handles)
    if lib.dBatchSplitAttach(B, rank, len(handles), blob) != 0:
        raise RuntimeError("attach failed: " + lib.dB200LastError().decode())


def step(lib, B, h, n):
    status = np.zeros(1, dtype=np.int32)
    if lib.dBatchCollideAndQuickStep(B, h, n, status.ctypes.data) != 0:
        raise RuntimeError("step failed: " + lib.dB200LastError().decode())
    if status[0]:
        raise RuntimeError(f"capacity overflow, status {status[0]}")


def state(lib, B):
    nb = lib.dBatchNumBodies(B)
    out = [np.zeros((nb, k), dtype=lib.np_real) for k in (3, 4, 3, 3)]
    if lib.dBatchGetBodyState(B, *[a.ctypes.data for a in out]) != 0:
        raise RuntimeError("get state failed")
    return np.concatenate(out, axis=1)


def step_ranks_in_threads(lib, batches, h, n):
    """every rank's dBatchCollideAndQuickStep on a host thread of its own (ctypes drops the GIL inside the call)"""
    errs = [None] * len(batches)

    def work(r):
        try:
            step(lib, batches[r], h, n)
        except Exception as e:  # noqa: BLE001 - reported to the caller below
            errs[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(len(batches))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in errs:
        if e is not None:
            raise e

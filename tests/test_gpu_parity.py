"""GPU parity (run on the B200): the CUDA path, called through the C ABI (dBatch*), against the
reference's golden traces and the live reference build shipped in oracle/_ref, plus
size-independent properties at BASELINE.json's full batch size."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ATAN2_LOCKSTEP_GOLDEN, ATAN2_SCENES, GOLDEN, GOLDEN_CALLBACK, ROOT, assert_bit_exact, assert_parity, have_ref, lib_path
from run_parity import parity, parity_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("stem,scene,steps,worlds,settle", GOLDEN)
def test_cuda_matches_golden_reference_traces(stem, scene, steps, worlds, settle, prec):
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace")
    r = parity_golden("b200", g, scene, prec, steps, worlds, settle, lockstep=(prec == "double" and scene in ATAN2_LOCKSTEP_GOLDEN))
    assert r["steps"] == steps
    assert_parity(r, f"{stem}/{prec}", scene, prec, "b200")


@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("stem,scene,steps,worlds,settle", GOLDEN_CALLBACK)
def test_cuda_dropin_matches_golden_reference_traces(stem, scene, steps, worlds, settle, prec):
    """ray colliders in every mode + capsule-trimesh on the GPU, through dSpaceCollide / dCollide"""
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace")
    r = parity_golden("b200", g, scene, prec, steps, worlds, settle, mode="callback")
    assert r["steps"] == steps and r["contacts"] > 0
    assert_parity(r, f"{stem}/{prec}", scene, prec, "b200")   # bit-exact, except dDOUBLE scenes with atan2 on the path (nested: hinge2 buggies), conftest.ATAN2_SCENES


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("scene,steps,worlds", [("stack32", 200, 3), ("block64", 50, 1), ("tower64", 250, 1),
                                                ("mixed", 200, 2), ("mixed_maxc4", 300, 1), ("chain", 250, 2), ("hinges", 250, 1), ("buggy", 250, 3), ("capsmix", 250, 3), ("ragdoll", 250, 3),
                                                ("block64@sap", 50, 1), ("stack32@sap", 200, 3), ("tower64@sapz", 250, 1), ("capsmix@simple", 200, 2),
                                                ("terrain_spheres", 300, 3), ("terrain_boxes", 250, 2), ("buggy_terrain", 300, 4), ("terrain_capsules", 300, 2), ("terrain_plane", 200, 2), ("sliders", 300, 2), ("universals", 300, 2), ("terrain_capsules_pre", 250, 2), ("motors", 300, 2), ("pistons", 300, 2), ("pus", 300, 2), ("cylmix", 300, 2), ("cylspheres", 200, 1), ("kinematic", 300, 2), ("nulljoint", 250, 2), ("transforms", 250, 2),
                                                ("bodyflags", 250, 2), ("autodisable", 400, 2), ("contactmodes", 250, 2), ("autodisable_avg", 400, 2), ("crashwall", 250, 3)])
def test_cuda_matches_live_reference(scene, steps, worlds, prec):
    # dDOUBLE scenes with atan2 on the path: lock-step protocol (SURVEY 8d, K = 1), see conftest.ATAN2_SCENES
    r = parity("b200", prec, scene, steps, worlds, lockstep=(prec == "double" and scene.split("@")[0] in ATAN2_SCENES))
    assert_parity(r, f"{scene}/{prec}", scene, prec, "b200")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("scene,steps,worlds", [("stack32", 60, 2), ("mixed_maxc4", 120, 1), ("chain", 100, 1), ("capsmix", 100, 1), ("block64", 20, 1),
                                                ("stack32@sap", 60, 2), ("mixed@simple", 100, 1), ("terrain_boxes", 80, 1), ("buggy_terrain", 100, 1),
                                                ("raycast", 150, 2), ("raycast2", 120, 2), ("raycast2h", 120, 1), ("raycyl", 120, 1), ("sliders", 150, 1), ("universals", 150, 1), ("motors", 150, 1), ("pistons", 150, 1), ("pus", 150, 1), ("cylmix", 150, 1), ("kinematic", 150, 1), ("nulljoint", 150, 1), ("transforms", 150, 1),
                                                ("hinges", 120, 1), ("buggy", 120, 1), ("ragdoll", 80, 1),
                                                ("bodyflags", 150, 1), ("autodisable", 300, 1), ("autodisable_avg", 300, 1), ("contactmodes", 150, 1), ("contactmodes_fdir1", 150, 1), ("mixed_varmaxc", 200, 2), ("nested", 200, 2), ("nested_dcollide", 200, 1), ("nested@sap", 150, 1), ("crashwall", 150, 1)])
def test_dropin_classic_api_matches_live_reference(scene, steps, worlds, prec):
    """the drop-in boundary: unchanged user code (dSpaceCollide + near callback calling dCollide /
    dJointCreateContact / dJointSetFeedback + dWorldQuickStep + dJointGroupEmpty) linked against
    libode_b200 runs on the GPU kernels and reproduces the reference's trace bit for bit."""
    # dDOUBLE scenes with atan2 / sin / cos on the path: lock-step protocol and the stated 1e-9 tolerance, as on the batched path
    tol = prec == "double" and scene.split("@")[0] in ATAN2_SCENES
    r = parity("b200", prec, scene, steps, worlds, mode="callback", lockstep=tol)
    assert r["pairs"] > 0
    assert_parity(r, f"dropin/{scene}/{prec}", scene, prec, "b200")


@pytest.mark.parametrize("stem,scene,steps,worlds,settle", [g for g in GOLDEN if g[1] in ("stack32", "ragdoll", "buggy_terrain", "mixed_maxc4", "tower64", "hinges")])
def test_lane_per_world_scheduler_matches_golden(stem, scene, steps, worlds, settle, monkeypatch):
    """k_sched_lane (one lane per world, used for batches of >= 512 worlds) forced on small batches"""
    monkeypatch.setenv("OB_SCHED_LANE", "1")
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_single.trace")
    r = parity_golden("b200", g, scene, "single", steps, worlds, settle)
    assert r["steps"] == steps
    assert_parity(r, f"{stem}/lane", scene, "single", "b200")

@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("stem,scene,steps,worlds,settle", [g for g in GOLDEN if g[1] in ("stack32", "ragdoll", "buggy_terrain", "mixed_maxc4", "tower64", "hinges", "buggy", "crashwall", "contactmodes")])
def test_lane_per_world_sweep_matches_golden(stem, scene, steps, worlds, settle, prec, monkeypatch):
    """k_sor_lane (one lane walks its world's rows in the schedule's order; the default for batches of >= 8192 tiny worlds) forced on
    the golden scenes: joint feedback (lambda tap) and body state bit for bit, like the tiled sweeps"""
    monkeypatch.setenv("OB_SOR_LANE", "1")
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace")
    r = parity_golden("b200", g, scene, prec, steps, worlds, settle)
    assert r["steps"] == steps
    assert_parity(r, f"{stem}/sorlane/{prec}", scene, prec, "b200")



@pytest.mark.parametrize("tile", ["4", "16", "32"])
@pytest.mark.parametrize("stem,scene,steps,worlds,settle", [g for g in GOLDEN if g[1] in ("stack32", "ragdoll", "buggy_terrain", "universals", "tower64")])
def test_every_tile_width_matches_golden(stem, scene, steps, worlds, settle, tile, monkeypatch):
    """G lanes per world in k_prep / k_sor / k_post: 4 (tiny worlds, large batches), 8 (default for >= 1024 worlds), 16, 32"""
    monkeypatch.setenv("OB_TILE", tile)
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_single.trace")
    r = parity_golden("b200", g, scene, "single", steps, worlds, settle)
    assert r["steps"] == steps
    assert_parity(r, f"{stem}/tile{tile}", scene, "single", "b200")


@pytest.mark.parametrize("deep", ["0", "1"])
@pytest.mark.parametrize("stem,scene,steps,worlds,settle", [g for g in GOLDEN if g[1] in ("stack32", "ragdoll", "buggy_terrain", "sliders")])
def test_both_sor_pipelines_match_golden(stem, scene, steps, worlds, settle, deep, monkeypatch):
    """k_sor<G, DEEP>: the deep index prefetch (worlds with > 256 row slots) and the shallow one, forced either way"""
    monkeypatch.setenv("OB_SOR_DEEP", deep)
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_single.trace")
    r = parity_golden("b200", g, scene, "single", steps, worlds, settle)
    assert r["steps"] == steps
    assert_parity(r, f"{stem}/deep{deep}", scene, "single", "b200")


def _batch(lib, scenes, scene, nworlds, cap=0):
    scenes.ob_scene_build_batch.restype = ctypes.c_void_p
    scenes.ob_scene_build_batch.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    B = scenes.ob_scene_build_batch(scene.encode(), nworlds, 0, cap, 0)
    assert B, ctypes.string_at(lib.dB200LastError())
    return ctypes.c_void_p(B)


def _state(lib, B, nworlds):
    lib.dBatchNumBodies.argtypes = [ctypes.c_void_p]
    nb = lib.dBatchNumBodies(B)
    arrs = [np.zeros((nworlds, nb, k), dtype=np.float32) for k in (3, 4, 3, 3)]
    lib.dBatchGetBodyState.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 4
    assert lib.dBatchGetBodyState(B, *[a.ctypes.data for a in arrs]) == 0
    return arrs


def test_full_batch_is_deterministic_and_replicates_small_batch():
    """4096 worlds x config 2: (i) two independent runs give identical bits, (ii) worlds do not
    interact: world w of the big batch equals world w of a 3-world batch, (iii) no capacity overflow."""
    lib = ctypes.CDLL(lib_path("single"))
    scenes = ctypes.CDLL(os.path.join(ROOT, "ode-0.12_b200", "lib", "libob_scenes_single.so"))
    lib.dB200LastError.restype = ctypes.c_char_p
    lib.dBatchCollideAndQuickStep.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    lib.dBatchDestroy.argtypes = [ctypes.c_void_p]
    W, steps = 4096, 40
    results = []
    for n in (W, W, 3):
        B = _batch(lib, scenes, "stack32", n, 256)
        status = np.zeros(n, dtype=np.int32)
        assert lib.dBatchCollideAndQuickStep(B, 0.01, steps, status.ctypes.data) == 0, lib.dB200LastError()
        assert not status.any()
        results.append(_state(lib, B, n))
        lib.dBatchDestroy(B)
    for a, b in zip(results[0], results[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for a, c in zip(results[0], results[2]):
        assert np.array_equal(a[:3].view(np.uint32), c.view(np.uint32))
    pos = results[0][0]
    assert np.isfinite(pos).all() and pos[:, :, 2].min() > -0.5


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("scene,W,cap,steps", [("stack32", 4096, 256, 60), ("buggy_terrain256", 65536, 48, 60), ("ragdoll", 16384, 160, 40)])
def test_sampled_worlds_of_the_full_batch_equal_the_reference(scene, W, cap, steps):
    """BASELINE.json configs[1..3] at their FULL batch size (the kernels' many-worlds-per-CTA / per-warp mappings, grid
    caps, world ranges): after `steps` free-running steps, worlds sampled from the start, the middle and the end of the
    batch hold bit for bit the body state the unmodified reference computes for the same world ids (dSINGLE)."""
    import subprocess
    import tempfile
    from run_parity import driver_path
    from tracecmp import read_trace

    lib = ctypes.CDLL(lib_path("single"))
    scenes = ctypes.CDLL(os.path.join(ROOT, "ode-0.12_b200", "lib", "libob_scenes_single.so"))
    lib.dB200LastError.restype = ctypes.c_char_p
    lib.dBatchCollideAndQuickStep.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    lib.dBatchDestroy.argtypes = [ctypes.c_void_p]
    B = _batch(lib, scenes, scene, W, cap)
    status = np.zeros(W, dtype=np.int32)
    assert lib.dBatchCollideAndQuickStep(B, 0.01, steps, status.ctypes.data) == 0, lib.dB200LastError()
    assert not status.any()
    pos, quat, lv, av = _state(lib, B, W)
    lib.dBatchDestroy(B)
    with tempfile.TemporaryDirectory() as td:
        for w0 in (0, W // 2 - 1, W - 2):
            f = os.path.join(td, f"ref_{w0}.bin")
            subprocess.run([driver_path("ref", "single"), "--scene", scene, "--worlds", "2", "--world0", str(w0), "--steps", str(steps), "--out", f],
                           check=True, capture_output=True, timeout=900)
            last = read_trace(f)["steps"][-1]
            for k in range(2):
                st = last[k]["state1"]
                nb = st.shape[0]
                got = np.concatenate([pos[w0 + k, :nb], quat[w0 + k, :nb], lv[w0 + k, :nb], av[w0 + k, :nb]], axis=1)
                assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(st).view(np.uint32)), f"{scene}: world {w0 + k} of {W} differs from the reference"


def test_no_gpu_fallback_symbols():
    """the product library must not contain the test-only host backend"""
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", lib_path("single")], capture_output=True, text=True).stdout
    assert "collide_world" not in out and "step_world" not in out

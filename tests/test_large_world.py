"""Large single world (pile in a walled box, dSweepAndPruneSpace, graph-coloured SOR).

Parity contract (BASELINE.json north_star: "Large single worlds use a graph-coloured SOR, and its results
are checked against the reference within a stated tolerance"), all in lock-step with the unmodified
reference (SURVEY 8d protocol, K = 1: every step starts from the reference's pre-step body state):

  exact   * the broadphase pair SET equals the reference's near-callback pair set, orientation included
            except for pairs whose float axis-0 minima tie exactly (the reference breaks those ties by
            RadixSort temporal-coherence state tied to its island stepping order): <= 1e-4 of the pairs;
          * every contact of the equally oriented pairs: geom ids equal, pos / normal / depth BITWISE equal;
  bitwise * the CUDA path equals the sequential mirror in tests/hostsim (a plain Gauss-Seidel sweep in the
            order (iteration, colour, pair, contact, row)): pairs, contacts and body state, bit for bit;
  solved  * the step's LCP is solved AT LEAST AS WELL as the reference solves it: the complementarity residual of the contact rows
  as well   after the 20 sweeps (oracle/lcp_residual.py: rows, lambda and the residual rebuilt from the traces with a numpy
            restatement of multiply_J / multiply_invM_JT / multiply_J_invM_JT, quickstep.cpp:138-195, pinned by reproducing
            the reference's output velocities to 1e-12) has, per step, an RMS <= 1.05 x the reference's (measured: 0.83 x --
            the colour order converges faster than a random order; the reference against itself with another seed: 1.00 +- 0.08);
  stated  * body state after one step vs the reference.  SOR_LCP stops after 20 sweeps, far from
  tolerance convergence on a pile, so its result depends on the row order: the REFERENCE ITSELF, run from the
            same state with another dRandInt seed, differs from itself by an RMS velocity difference s_ref.
            Tolerance: RMS velocity (linear and angular) difference to the reference <= 2 x s_ref, and
            max relative error  |dv|_inf / max(1,|v|_inf) <= 0.3,  |dx|_inf / max(1,|x|_inf) <= 1e-2  per step.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, have_ref, lib_path
from run_parity import driver_path, run_trace
from tracecmp import compare, compare_large, read_trace

SCENE, SETTLE, STEPS = "pile_10x10x20", 100, 30
GOLD = ("pile_5x5x8_large_settle60", "pile_5x5x8", 10, 60)   # committed reference trace (tests/golden/make_golden.sh)


def _ref_seed_sensitivity(prec, td, fr):
    """the reference against itself: same pre-step states, another SOR shuffle stream"""
    f2 = os.path.join(td, "ref2.bin")
    cmd = [driver_path("ref", prec), "--scene", SCENE, "--steps", str(STEPS), "--settle", str(SETTLE), "--mode", "callback",
           "--out", f2, "--seed-xor", "0x5555", "--resync", fr]
    subprocess.run(cmd, check=True, capture_output=True, timeout=900)
    r = compare_large(read_trace(f2), read_trace(fr))
    return r["rms_dv"], r["rms_dw"]


def _residuals(trace, nx=10, ny=10, nz=20):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lcp_residual as L

    kinds, gb = L.pile_scene(nx, ny, nz)
    return [L.residual(st[0], kinds, gb, 0.01) for st in trace["steps"]]


def _check_vs_reference(cand, prec):
    with tempfile.TemporaryDirectory() as td:
        fr, fc = os.path.join(td, "ref.bin"), os.path.join(td, "cand.bin")
        run_trace("ref", prec, SCENE, STEPS, 1, fr, settle=SETTLE)
        run_trace(cand, prec, SCENE, STEPS, 1, fc, settle=SETTLE, resync=fr)
        tc, tr = read_trace(fc), read_trace(fr)
        r = compare_large(tc, tr)
        sv, sw = _ref_seed_sensitivity(prec, td, fr)
    # how well each side solved the same LCPs (lock-step: same pre-step state, same contacts)
    rr, rc = _residuals(tr), _residuals(tc)
    vtol = 1e-10 if prec == "double" else 5e-4
    assert max(x["vpred_err"] for x in rr) <= vtol and max(x["vpred_err"] for x in rc) <= vtol   # the oracle explains both traces' velocities
    assert all(a["rows"] == b["rows"] and a["rows"] > 3000 for a, b in zip(rr, rc))
    ratios = [b["rms"] / a["rms"] for a, b in zip(rr, rc)]
    assert max(ratios) <= 1.05, ratios
    r["residual_ratio_mean"] = float(np.mean(ratios))
    assert r["steps"] == STEPS and r["pairs"] > 100000 and r["contacts"] > 20000
    assert r["state0_bits_equal"]
    assert r["pair_sets_equal"], r["first_mismatch"]
    assert r["pairs_flipped"] <= 1e-4 * r["pairs"]
    assert r["contact_ids_equal"], r["first_mismatch"]
    assert r["contact_bits_equal"] == r["contact_vals"]
    assert r["rms_dv"] <= 2 * sv and r["rms_dw"] <= 2 * sw, (r["rms_dv"], sv, r["rms_dw"], sw)
    assert r["max_vel_relerr"] <= 0.3 and r["max_pos_relerr"] <= 1e-2
    return r


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("prec", ["single", "double"])
def test_large_mirror_vs_live_reference(prec):
    _check_vs_reference("hostsim", prec)


@pytest.mark.parametrize("prec", ["single", "double"])
def test_large_mirror_vs_golden_reference_trace(prec):
    stem, scene, steps, settle = GOLD
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace")
    with tempfile.TemporaryDirectory() as td:
        fc = os.path.join(td, "cand.bin")
        run_trace("hostsim", prec, scene, steps, 1, fc, settle=settle, resync=g, large=True)
        r = compare_large(read_trace(fc), read_trace(g))
    assert r["steps"] == steps and r["contacts"] > 1000
    assert r["state0_bits_equal"] and r["pair_sets_equal"] and r["pairs_flipped"] <= 1 and r["contact_ids_equal"]
    assert r["contact_bits_equal"] == r["contact_vals"]
    assert r["max_vel_relerr"] <= 0.3 and r["max_pos_relerr"] <= 1e-2


def test_large_mirror_free_run_settles():
    """free-running (no re-sync): the pile comes to rest inside the walls, no body tunnels through the floor"""
    with tempfile.TemporaryDirectory() as td:
        fc = os.path.join(td, "cand.bin")
        run_trace("hostsim", "single", "pile_6x6x10", 5, 1, fc, settle=400)
        t = read_trace(fc)
    st = t["steps"][-1][0]["state1"]
    assert np.isfinite(st).all()
    assert st[:, 2].min() > 0.15 and np.abs(st[:, 0]).max() < 3.6 and np.abs(st[:, 1]).max() < 3.6
    assert np.abs(st[:, 7:10]).max() < 1.0


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["single", "double"])
def test_large_cuda_equals_mirror_bitwise(prec):
    stem, scene, steps, settle = GOLD
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace")
    with tempfile.TemporaryDirectory() as td:
        fa, fb = os.path.join(td, "cuda.bin"), os.path.join(td, "mirror.bin")
        run_trace("b200", prec, scene, steps, 1, fa, settle=settle, resync=g, large=True)
        run_trace("hostsim", prec, scene, steps, 1, fb, settle=settle, resync=g, large=True)
        r = compare(read_trace(fa), read_trace(fb))
    assert r["steps"] == steps and r["exact_ok"], r["first_exact_mismatch"]
    assert r["contact_bits_equal"] == r["contact_vals"] and r["state_bits_equal"] == r["state_vals"], r["first_state_bit_mismatch"]


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("prec", ["single", "double"])
def test_large_cuda_equals_mirror_bitwise_2000_bodies_free_running(prec):
    """both start from the same scene and run free for 150 steps: any difference would be amplified"""
    with tempfile.TemporaryDirectory() as td:
        fa, fb = os.path.join(td, "cuda.bin"), os.path.join(td, "mirror.bin")
        run_trace("b200", prec, SCENE, 20, 1, fa, settle=130)
        run_trace("hostsim", prec, SCENE, 20, 1, fb, settle=130)
        r = compare(read_trace(fa), read_trace(fb))
    assert r["exact_ok"], r["first_exact_mismatch"]
    assert r["contacts"] > 20000
    assert r["contact_bits_equal"] == r["contact_vals"] and r["state_bits_equal"] == r["state_vals"], r["first_state_bit_mismatch"]


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("prec", ["single", "double"])
def test_large_cuda_vs_live_reference(prec):
    _check_vs_reference("b200", prec)


@pytest.mark.gpu
def test_large_full_size_200k_bodies():
    """BASELINE config 5 at full size through the C ABI: 200 000 bodies, two independent runs are
    bit-identical, nothing overflows, the pile stays inside the walls."""
    lib = ctypes.CDLL(lib_path("single"))
    scenes = ctypes.CDLL(os.path.join(ROOT, "ode-0.12_b200", "lib", "libob_scenes_single.so"))
    lib.dB200LastError.restype = ctypes.c_char_p
    scenes.ob_scene_build_batch.restype = ctypes.c_void_p
    scenes.ob_scene_build_batch.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.dBatchCollideAndQuickStep.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    lib.dBatchDestroy.argtypes = [ctypes.c_void_p]
    lib.dBatchGetBodyState.argtypes = [ctypes.c_void_p] * 5
    lib.dBatchNumBodies.argtypes = [ctypes.c_void_p]
    res = []
    for _ in range(2):
        B = scenes.ob_scene_build_batch(b"pile_100x100x20", 1, 0, 0, 0)
        assert B, lib.dB200LastError()
        B = ctypes.c_void_p(B)
        nb = lib.dBatchNumBodies(B)
        assert nb == 200000
        status = np.zeros(1, dtype=np.int32)
        assert lib.dBatchCollideAndQuickStep(B, 0.01, 120, status.ctypes.data) == 0, lib.dB200LastError()
        assert status[0] == 0
        arrs = [np.zeros((nb, k), dtype=np.float32) for k in (3, 4, 3, 3)]
        assert lib.dBatchGetBodyState(B, *[a.ctypes.data for a in arrs]) == 0
        res.append(arrs)
        lib.dBatchDestroy(B)
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    pos = res[0][0]
    assert np.isfinite(pos).all()
    assert pos[:, 2].min() > 0.1 and np.abs(pos[:, 0]).max() < 50.6 and np.abs(pos[:, 1]).max() < 50.6


@pytest.mark.parametrize("prec", ["single", "double"])
def test_residual_oracle_reproduces_the_reference_velocities(prec):
    """oracle/lcp_residual.py on the committed REFERENCE trace: lambda recovered from the joint feedback and pushed through the
    restated multiply_invM_JT (quickstep.cpp:138-160) must give back the reference's own post-step velocities; and
    J v_new - J v_free == h * multiply_J_invM_JT(lambda) (quickstep.cpp:187-195)"""
    stem, scene, steps, settle = GOLD
    t = read_trace(os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace"))
    rs = _residuals(t, 5, 5, 8)
    assert len(rs) == steps and all(r["rows"] > 300 for r in rs)
    assert max(r["vpred_err"] for r in rs) <= (1e-12 if prec == "double" else 1e-4)
    assert max(r["JMJt_check"] for r in rs) <= 1e-12

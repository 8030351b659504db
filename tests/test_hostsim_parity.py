"""CPU parity: the per-element device functions (ode-0.12_b200/csrc/ob_*.h), compiled for the
host and driven sequentially by the TEST-ONLY backend in tests/hostsim, must reproduce the
unmodified reference bit-for-bit: callback pair order, contacts, space-list order, LCG seed
(exact by contract) and — stronger than the stated tolerance — body state and joint feedback.
Golden traces are the reference's own outputs (tests/golden/make_golden.sh)."""
import os

import pytest

from conftest import GOLDEN, GOLDEN_CALLBACK, ROOT, assert_bit_exact, have_ref
from run_parity import parity, parity_golden


@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("stem,scene,steps,worlds,settle", GOLDEN)
def test_hostsim_matches_golden_reference_traces(stem, scene, steps, worlds, settle, prec):
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace")
    r = parity_golden("hostsim", g, scene, prec, steps, worlds, settle)
    assert r["steps"] == steps
    assert_bit_exact(r, f"{stem}/{prec}")


@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("stem,scene,steps,worlds,settle", GOLDEN_CALLBACK)
def test_hostsim_dropin_matches_golden_reference_traces(stem, scene, steps, worlds, settle, prec):
    """ray colliders (ray-sphere/box/capsule/plane/trimesh in every ray mode) and capsule-trimesh, through the classic callback loop"""
    g = os.path.join(ROOT, "tests", "golden", f"{stem}_{prec}.trace")
    r = parity_golden("hostsim", g, scene, prec, steps, worlds, settle, mode="callback")
    assert r["steps"] == steps and r["contacts"] > 0
    assert_bit_exact(r, f"{stem}/{prec}")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("scene,steps,worlds", [("stack32", 250, 2), ("block64", 60, 1), ("tower64", 300, 1),
                                                ("mixed_maxc4", 400, 1), ("chain", 300, 2), ("hinges", 300, 1), ("buggy", 300, 2), ("capsmix", 300, 2), ("ragdoll", 300, 2),
                                                ("block64@sap", 60, 1), ("stack32@sap", 200, 2), ("tower64@sapz", 250, 1), ("capsmix@simple", 200, 1),
                                                ("terrain_spheres", 300, 2), ("terrain_boxes", 250, 1), ("buggy_terrain", 300, 3), ("terrain_capsules", 300, 2), ("terrain_plane", 200, 2), ("sliders", 300, 2), ("universals", 300, 2), ("terrain_capsules_pre", 250, 2), ("motors", 300, 2), ("pistons", 300, 2), ("pus", 300, 2), ("cylmix", 300, 2), ("cylspheres", 200, 1), ("kinematic", 300, 2), ("nulljoint", 250, 2), ("transforms", 250, 2), ("transforms@sap", 200, 1), ("transforms@simple", 200, 1),
                                                ("bodyflags", 300, 2), ("autodisable", 400, 2), ("autodisable_avg", 400, 2), ("contactmodes", 300, 2), ("bodyflags@sap", 150, 1), ("crashwall", 250, 2)])
def test_hostsim_matches_live_reference(scene, steps, worlds):
    r = parity("hostsim", "single", scene, steps, worlds)
    assert r["contacts"] > 0
    assert_bit_exact(r, scene)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("scene,steps,worlds", [("stack32", 80, 2), ("mixed_maxc4", 150, 1), ("hinges", 150, 1), ("buggy", 120, 2), ("ragdoll", 100, 1),
                                                ("raycast", 250, 3), ("raycast2", 200, 2), ("raycast2h", 200, 2), ("raycyl", 200, 1), ("sliders", 200, 1), ("universals", 200, 1), ("motors", 200, 1), ("pistons", 200, 1), ("pus", 200, 1), ("cylmix", 200, 1), ("kinematic", 200, 1), ("nulljoint", 200, 1), ("transforms", 200, 1), ("transforms@sapz", 150, 1), ("transforms_rays", 200, 2),
                                                ("bodyflags", 200, 1), ("autodisable", 400, 1), ("autodisable_avg", 400, 1), ("contactmodes", 200, 1), ("contactmodes_fdir1", 300, 2), ("mixed_varmaxc", 300, 2), ("nested", 300, 2), ("nested_dcollide", 300, 2), ("nested@sap", 200, 1), ("nested_dcollide@simple", 200, 1), ("crashwall", 200, 1)])
def test_hostsim_dropin_callback_loop_matches_live_reference(scene, steps, worlds):
    """host logic of the drop-in path (ob_dropin.cpp): the classic loop dSpaceCollide + near callback
    (dCollide, dJointCreateContact, dJointAttach, dJointSetFeedback) + dWorldQuickStep +
    dJointGroupEmpty, same driver source as the reference's, must give the same trace."""
    r = parity("hostsim", "single", scene, steps, worlds, mode="callback")
    assert_bit_exact(r, f"dropin/{scene}")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
sys.path.insert(0, ROOT)

GOLDEN = [  # (file stem, scene, steps, worlds, settle) — must match tests/golden/make_golden.sh
    ("free6", "free6", 20, 1, 0),
    ("mixed", "mixed", 40, 1, 0),
    ("mixed_maxc4_settle80", "mixed_maxc4", 40, 1, 80),
    ("stack32_w2_settle120", "stack32", 12, 2, 120),
    ("tower64_settle150", "tower64", 8, 1, 150),
    ("chain_settle60", "chain", 30, 1, 60),
    ("hinges_settle70", "hinges", 30, 1, 70),
    ("buggy_settle100", "buggy", 30, 1, 100),
    ("capsmix_settle90", "capsmix", 20, 1, 90),
    ("ragdoll_settle110", "ragdoll", 12, 1, 110),
    ("block64_sap_settle20", "block64@sap", 6, 1, 20),
    ("mixed_sapz_settle60", "mixed@sapz", 30, 1, 60),
    ("mixed_simple_settle60", "mixed@simple", 30, 1, 60),
    ("terrain_spheres_settle70", "terrain_spheres", 25, 1, 70),
    ("terrain_boxes_settle70", "terrain_boxes", 15, 1, 70),
    ("buggy_terrain_w2_settle90", "buggy_terrain", 30, 2, 90),
    ("terrain_capsules_settle70", "terrain_capsules", 20, 1, 70),
    ("terrain_plane_settle60", "terrain_plane", 20, 1, 60),
    ("terrain_capsules_pre_settle70", "terrain_capsules_pre", 20, 1, 70),   # dGeomTriMeshDataPreprocess: edge / vertex use flags
    ("sliders_settle60", "sliders", 30, 1, 60),
    ("universals_settle60", "universals", 30, 1, 60),
    ("motors_settle60", "motors", 30, 1, 60),
    ("pistons_settle60", "pistons", 30, 1, 60),   # piston / PR / plane2d joints
    ("pus_settle60", "pus", 30, 1, 60),           # PU joints
    ("cylmix_settle90", "cylmix", 20, 1, 90),     # flat cylinders vs plane / sphere / box
    ("kinematic_settle60", "kinematic", 30, 1, 60),   # dBodySetKinematic bodies pushing a pile, hinged to a dynamic body
    ("nulljoint_settle60", "nulljoint", 30, 1, 60),   # null joints merge islands (order of the dRandInt draws)
    ("transforms_settle80", "transforms", 25, 1, 80),   # geom transforms: composite bodies (T x X, X x T, T x T collider order), static transform
    ("bodyflags_settle50", "bodyflags", 30, 1, 50),     # finite rotation (both modes: glibc-exact sinf / cosf), damping + thresholds, max angular speed, gravity mode, gyroscopic off, contact max-correcting-vel / surface layer
    ("autodisable_settle150", "autodisable", 30, 1, 150),   # auto-disable (instantaneous samples) + re-enabling through the island walk
    ("contactmodes_settle60", "contactmodes", 30, 1, 60),   # Mu2, Motion1/2/N, Slip1/2, Bounce, SoftERP/CFM, Approx1_2
    ("autodisable_avg_settle150", "autodisable_avg", 30, 1, 150),   # auto-disable on averaged velocity samples (5 / 3 / 1 samples per body)
    ("crashwall_settle45", "crashwall", 30, 1, 45),   # demo_crash as shipped: SAP space, brick wall, hinge2 car + fixed counterweight, cannon ball; TWO policy rows (mu by geom class)
]


# reference traces replayed through the DROP-IN path (classic callback loop); scenes with ray geoms only
# exist there: a ray contact is a query result for the caller, not a contact joint
GOLDEN_CALLBACK = [
    ("raycast_settle40", "raycast", 20, 1, 40),
    ("raycast2_settle40", "raycast2", 20, 1, 40),   # rays in a second space: dSpaceCollide2 (space x space, geom x space)
    ("raycyl_settle40", "raycyl", 20, 1, 40),       # ray-cylinder (mantle and cap branches)
    ("contactmodes_fdir1_settle60", "contactmodes_fdir1", 25, 1, 60),   # dContactFDir1: per-contact first friction direction set by the callback
    ("mixed_varmaxc_settle40", "mixed_varmaxc", 30, 1, 40),             # max-contacts differs from dCollide call to call (cached batch results vs on-demand pairs)
    ("nested_settle60", "nested", 25, 1, 60),                           # sub-spaces: dSpaceCollide2 recursion with the sublevel rule + interior dSpaceCollide (demo_buggy's car space)
    ("nested_dcollide_settle60", "nested_dcollide", 25, 1, 60),         # dCollide on a (space, geom) pair: dCollideSpaceGeom
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "driver_ref_single"))


def lib_path(prec="single"):
    return os.path.join(ROOT, "ode-0.12_b200", "lib", f"libode_b200_{prec}.so")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the host-side artefacts once per session if they are missing (CPU only)."""
    need = [lib_path("single"), os.path.join(ROOT, "tests", "hostsim", "_build", "driver_hostsim_single")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__

        __graft_entry__.build()
    yield


# scenes whose rows depend on atan2 (hinge / hinge2 limit angles).  dSINGLE reproduces glibc's atan2f
# bit for bit (ob_math.h); for dDOUBLE the device has no bit-identical atan2 (glibc's is correctly
# rounded, CUDA's is <= 2 ulp), so those scenes are held to the stated tolerance instead:
# exact discrete observables + |dx|_inf / max(1,|x|_inf) <= 1e-9, over the free-running golden
# traces and, for the long live comparisons, in lock-step with the reference (every step starts
# from the reference's pre-step body state, SURVEY 8d parity protocol with K = 1) so that a
# last-bit difference cannot be amplified by chaotic dynamics into a different contact set.
ATAN2_SCENES = ("hinges", "buggy", "ragdoll", "buggy_terrain", "universals", "motors", "pistons", "pus",
                "bodyflags", "nested", "nested_dcollide", "crashwall")   # nested: hinge2 buggies; bodyflags: finite rotation calls sin / cos (dDOUBLE: platform libm, same tolerance class; dSINGLE: glibc-exact restatement)


# dDOUBLE on the GPU, scenes whose limit-motors bounce off their stops (restitution turns a last-bit atan2 difference
# into a different rebound within a few dozen free-running steps): their golden traces are replayed in lock-step too
ATAN2_LOCKSTEP_GOLDEN = ("pistons", "pus")


def assert_parity(r, what, scene, prec, cand):
    if prec == "double" and cand == "b200" and scene.split("@")[0] in ATAN2_SCENES:
        assert r["exact_ok"], f"{what}: exact observables differ at {r['first_exact_mismatch']}"
        assert r["max_state_relerr"] <= 1e-9 and r["max_contact_relerr"] <= 1e-9, f"{what}: {r['max_state_relerr']}"
    else:
        assert_bit_exact(r, what)


def assert_bit_exact(r, what=""):
    assert r["exact_ok"], f"{what}: exact observables differ at {r['first_exact_mismatch']}"
    assert r["contact_bits_equal"] == r["contact_vals"], f"{what}: contact fields differ (relerr {r['max_contact_relerr']})"
    assert r["state_bits_equal"] == r["state_vals"], f"{what}: body state differs at {r['first_state_bit_mismatch']} (relerr {r['max_state_relerr']})"
    assert r["fb_bits_equal"] == r["fb_vals"], f"{what}: joint feedback (lambda tap) differs (relerr {r['max_fb_relerr']})"

"""dBatchRayCast (ode.h): the batched form of the ray-cast vehicle's wheel probe (demos/raycar/car.cpp:353-371).  The same
deterministic probe rays are answered (i) by the unmodified reference through the classic calls -- dCreateRay, dGeomRaySet,
dGeomRaySetParams, dGeomRaySetClosestHit, dSpaceCollide2(space, ray, callback keeping the nearest dCollide(ray, geom, 1)) -- and
(ii) by ONE dBatchRayCast call on the bound batch after the same steps; the two hit tables (position, depth, normal, geom)
must be byte-identical.  Scenes cover ray x sphere / box / capsule / cylinder / plane / trimesh and geom transforms,
hash / SAP / simple spaces.  CPU: the test-only host build; GPU: the product library."""
import os
import subprocess
import tempfile

import pytest

from conftest import have_ref
from run_parity import driver_path

SCENES = ["mixed", "terrain_boxes", "capsmix", "cylmix", "buggy_terrain", "transforms", "terrain_capsules", "terrain_spheres", "mixed@sap", "capsmix@simple"]


def _rays(cand, prec, scene, steps=40, worlds=2, rays=300):
    with tempfile.TemporaryDirectory() as td:
        fr, fc = os.path.join(td, "ref.bin"), os.path.join(td, "cand.bin")
        subprocess.run([driver_path("ref", prec), "--scene", scene, "--steps", str(steps), "--worlds", str(worlds), "--rays", str(rays), fr],
                       check=True, capture_output=True, timeout=600)
        subprocess.run([driver_path(cand, prec), "--scene", scene, "--steps", str(steps), "--worlds", str(worlds), "--mode", "batch", "--rays", str(rays), fc],
                       check=True, capture_output=True, timeout=600)
        a, b = open(fr, "rb").read(), open(fc, "rb").read()
        assert len(a) == len(b) and len(a) > 0
        return a == b


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("scene", SCENES)
def test_hostsim_batch_raycast_equals_reference_probe(scene, prec):
    assert _rays("hostsim", prec, scene)


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("scene", SCENES)
def test_cuda_batch_raycast_equals_reference_probe(scene, prec):
    # dDOUBLE free-running steps of hinge-limit scenes are tolerance class (conftest.ATAN2_SCENES): compare before any step there
    steps = 0 if (prec == "double" and scene.split("@")[0] in ("buggy_terrain",)) else 40
    assert _rays("b200", prec, scene, steps=steps)

"""dWorldExportDIF (include/ode/export-dif.h:31, ode/src/export-dif.cpp): the text dump of a world must equal the
reference's BYTE FOR BYTE -- same tables, field order, number formatting (%.7g / %.15g, inf), and the same values,
which include everything the joints' host bookkeeping maintains (body-frame anchors / axes / qrel of every joint type,
limit-motor parameters, Euler angles an amotor measured in its last getInfo1) and the contact joints of the step.

The dump is taken INSIDE the last step (after dSpaceCollide + the near callback, before dWorldQuickStep), through the
drop-in API, after `settle + steps - 1` full steps: so it also checks that the classic API leaves the host objects in
the reference's state.  not gpu: the host mirror of the kernels; gpu: the CUDA library."""
import filecmp
import os
import subprocess
import tempfile

import pytest

from conftest import ATAN2_SCENES, ROOT, have_ref
from run_parity import driver_path

SCENES = ["stack32", "mixed", "hinges", "buggy", "ragdoll", "capsmix", "sliders", "universals", "motors", "pistons", "pus", "cylmix", "raycyl", "transforms"]
GOLDEN = [("pistons", "single"), ("motors", "double"), ("cylmix", "single"), ("cylmix", "double"), ("transforms", "single")]   # tests/golden/*.dif, written by the reference (make_golden.sh)


def _export(kind, prec, scene, path, steps=25, settle=20):
    cmd = [driver_path(kind, prec), "--scene", scene, "--steps", str(steps), "--settle", str(settle), "--mode", "callback", "--export-dif", path]
    subprocess.run(cmd, check=True, capture_output=True, timeout=600)


def _check_live(cand, scene, prec):
    with tempfile.TemporaryDirectory() as td:
        fr, fc = os.path.join(td, "ref.dif"), os.path.join(td, "cand.dif")
        _export("ref", prec, scene, fr)
        _export(cand, prec, scene, fc)
        assert os.path.getsize(fr) > 1000
        assert filecmp.cmp(fr, fc, shallow=False), f"{scene}/{prec}: DIF dump differs from the reference's"


def _check_golden(cand, scene, prec):
    g = os.path.join(ROOT, "tests", "golden", f"{scene}_{prec}.dif")
    with tempfile.TemporaryDirectory() as td:
        fc = os.path.join(td, "cand.dif")
        _export(cand, prec, scene, fc)
        assert filecmp.cmp(g, fc, shallow=False), f"{scene}/{prec}: DIF dump differs from tests/golden/{scene}_{prec}.dif"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("scene", SCENES)
def test_hostsim_dif_equals_reference(scene, prec):
    _check_live("hostsim", scene, prec)


@pytest.mark.parametrize("scene,prec", GOLDEN)
def test_hostsim_dif_equals_golden(scene, prec):
    _check_golden("hostsim", scene, prec)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,prec", GOLDEN)
def test_cuda_dif_equals_golden(scene, prec):
    if prec == "double" and scene in ATAN2_SCENES:
        pytest.skip("dDOUBLE atan2 scenes are tolerance-class on the GPU (conftest.ATAN2_SCENES): a 15-digit text dump after 45 free-running steps cannot be byte-equal")
    _check_golden("b200", scene, prec)


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("scene", ["stack32", "sliders", "pistons", "pus", "cylmix", "capsmix"])
def test_cuda_dif_equals_reference(scene):
    # dSINGLE: the whole free-running trajectory is bit-exact on the GPU, so the dump after 45 steps is too
    _check_live("b200", scene, "single")

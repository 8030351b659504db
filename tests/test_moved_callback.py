"""dBodySetMovedCallback through the drop-in dWorldQuickStep: the callbacks fire for exactly the bodies the reference
steps, in the reference's stepping order (island by island, util.cpp:384-521 + :338-340), every step."""
import filecmp
import os
import subprocess
import tempfile

import pytest

from conftest import have_ref
from run_parity import driver_path


def _log(kind, scene, path, mode_args):
    subprocess.run([driver_path(kind, "single"), "--scene", scene, "--steps", "80", "--settle", "10", "--moved-log", path] + mode_args,
                   check=True, capture_output=True, timeout=600)


def _check(cand, scene):
    with tempfile.TemporaryDirectory() as td:
        fr, fc = os.path.join(td, "ref.txt"), os.path.join(td, "cand.txt")
        _log("ref", scene, fr, [])
        _log(cand, scene, fc, ["--mode", "callback"])
        assert os.path.getsize(fr) > 500
        assert filecmp.cmp(fr, fc, shallow=False), f"{scene}: order of the moved callbacks differs from the reference's"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("scene", ["stack32", "mixed", "ragdoll", "tower64", "kinematic"])
def test_hostsim_moved_callback_order(scene):
    _check("hostsim", scene)


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("scene", ["stack32", "ragdoll"])
def test_cuda_moved_callback_order(scene):
    _check("b200", scene)

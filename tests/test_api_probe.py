"""The small accessor / convenience entry points of the ODE C API (ob_api_extra.cpp: body / world damping and
auto-disable accessors, force-at-position helpers, joint data / connectivity / torque helpers, geom offset getters,
rotation, random and mass utilities): tests/harness/api_probe.cpp calls each of them on a fixed scene and prints the
results as raw bits.  The probe linked against the UNMODIFIED reference (and compiled against the reference's own
headers, so the signatures are source-compatible) must print exactly what the probe linked against this library
prints -- host mirror build on the CPU, CUDA library build on the GPU box (the probe itself needs no GPU)."""
import os
import subprocess

import pytest

from conftest import ROOT, have_ref

GOLD = os.path.join(ROOT, "tests", "golden")


def _run(path):
    return subprocess.run([path], check=True, capture_output=True, text=True, timeout=120).stdout


def _probe(kind, prec):
    if kind == "ref":
        return os.path.join(ROOT, "oracle", "_ref", f"api_probe_ref_{prec}")
    if kind == "hostsim":
        return os.path.join(ROOT, "tests", "hostsim", "_build", f"api_probe_hostsim_{prec}")
    return os.path.join(ROOT, "ode-0.12_b200", "lib", f"api_probe_b200_{prec}")


@pytest.mark.parametrize("prec", ["single", "double"])
def test_probe_equals_golden_reference_output(prec):
    want = open(os.path.join(GOLD, f"api_probe_{prec}.txt")).read()
    assert len(want.splitlines()) > 60
    assert _run(_probe("hostsim", prec)) == want


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("prec", ["single", "double"])
def test_probe_equals_live_reference(prec):
    assert _run(_probe("hostsim", prec)) == _run(_probe("ref", prec))


@pytest.mark.parametrize("prec", ["single", "double"])
def test_product_library_probe(prec):
    """the shipped library (host object model only: no kernel is launched by these calls)"""
    want = open(os.path.join(GOLD, f"api_probe_{prec}.txt")).read()
    assert _run(_probe("b200", prec)) == want

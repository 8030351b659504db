"""One large world over several GPUs (BASELINE.json configs[4] "2/4/8-GPU split"):
dBatchSplitExport / dBatchSplitAttach (include/ode_b200/ode.h); ob_large_kernels.cuh: front-end split (k_lw_sweep /
k_lw_narrow by SAP-sorted position + k_lw_xbarrier, the default) and SOR split (k_lw_sor_split, OB_LW_SPLIT_SOR=1).

Contract: every rank ends every step with the body state of the single-GPU coloured sweep, BIT FOR BIT
(the split changes who computes a pair or executes a pair of a colour, never a value or the order of dependent
updates), and that single-GPU sweep is what tests/test_large_world.py holds against the reference.

  not gpu  the host mirror of the protocol (tests/hostsim/large_host.h: ranks are host threads, same dealing
           of warp tiles, same flag barrier) equals the unsplit mirror for 2, 3 and 4 ranks;
  gpu      loop-back: two ranks on ONE GPU (two batches, two host threads, narrow CTAs so both persistent
           kernels are co-resident) equal the unsplit CUDA path bit for bit;
  gpu x2   (skipped on a one-GPU box) one device per rank in one process, and one process per GPU under
           torch.distributed.run with the CUDA-IPC exchange (tests/harness/split_rank.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

SCENE, H = "pile_10x10x20", 0.01


def _run_split(backend, prec, nranks, devices, steps=(60, 5, 5), env=None):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "harness", "split_check.py"), "--backend", backend, "--prec", prec, "--ranks", str(nranks),
           "--devices", ",".join(map(str, devices)), "--scene", SCENE, "--steps", ",".join(map(str, steps))]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e)
    if r.returncode == 3:
        return "timeout"
    assert r.returncode == 0 and "SPLIT_EQUAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    return "equal"


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_split_mirror_equals_unsplit_mirror(nranks):
    assert _run_split("hostsim", "single", nranks, [0], steps=(30, 3)) == "equal"


def test_split_mirror_double():
    assert _run_split("hostsim", "double", 2, [0], steps=(20, 3)) == "equal"


def _gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["front", "front+sor"])
@pytest.mark.parametrize("prec", ["single", "double"])
def test_split_loopback_one_gpu(prec, mode):
    # two persistent cooperative kernels must be co-resident on one GPU: one CTA of 128 threads per SM each
    env = {"OB_LW_SOR_CTAS_PER_SM": "1", "OB_LW_SOR_THREADS": "128", "OB_LW_SPLIT_TIMEOUT_MS": "3000"}
    if mode == "front+sor":
        env["OB_LW_SPLIT_SOR"] = "1"
    if _run_split("b200", prec, 2, [0], env=env) == "timeout":
        pytest.skip("the two ranks' kernels were not scheduled together on this GPU (loop-back needs co-residency)")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["front", "front+sor"])
def test_split_one_device_per_rank(mode):
    n = _gpus()
    if n < 2:
        pytest.skip("needs two GPUs")
    assert _run_split("b200", "single", min(n, 4), list(range(min(n, 4))), env={"OB_LW_SPLIT_SOR": "1"} if mode == "front+sor" else None) == "equal"


@pytest.mark.gpu
def test_split_one_process_per_gpu():
    n = _gpus()
    if n < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "harness", "split_rank.py"), "--scene", SCENE, "--steps", "70"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SPLIT_OK" in r.stdout

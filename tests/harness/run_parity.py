"""Run one scene through the reference driver and a candidate driver and compare traces.

usage: python tests/harness/run_parity.py [--cand b200|hostsim] [--prec single|double]
                                          [--scene NAME ...] [--steps N] [--worlds W]
Prints one JSON line per scene.  Used by tests/ and for quick checks under gpurun.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from tracecmp import compare, read_trace  # noqa: E402


def driver_path(kind, prec):
    if kind == "ref":
        return os.path.join(ROOT, "oracle", "_ref", f"driver_ref_{prec}")
    if kind == "b200":
        return os.path.join(ROOT, "ode-0.12_b200", "lib", f"driver_b200_{prec}")
    if kind == "hostsim":
        return os.path.join(ROOT, "tests", "hostsim", "_build", f"driver_hostsim_{prec}")
    raise ValueError(kind)


def run_trace(kind, prec, scene, steps, worlds, out, mode=None, timeout=900, settle=0, contacts_cap=0, resync=None, large=False):
    exe = driver_path(kind, prec)
    if mode is None:
        mode = "callback" if kind == "ref" else "batch"
    cmd = [exe, "--scene", scene, "--steps", str(steps), "--worlds", str(worlds), "--mode", mode, "--out", out]
    if settle:
        cmd += ["--settle", str(settle)]
    if contacts_cap and kind != "ref":
        cmd += ["--contacts-cap", str(contacts_cap)]
    if resync and kind != "ref":
        cmd += ["--resync", resync]
    if large and kind != "ref":
        cmd += ["--large"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed rc={r.returncode}\n{r.stdout}\n{r.stderr}")
    return r.stderr


def parity(cand, prec, scene, steps, worlds=1, mode=None, settle=0, lockstep=False):
    """lockstep: the candidate reloads the reference's pre-step body state before every step
    (SURVEY 8d parity protocol with K = 1); otherwise both run free from the same start."""
    with tempfile.TemporaryDirectory() as td:
        fr, fc = os.path.join(td, "ref.bin"), os.path.join(td, "cand.bin")
        run_trace("ref", prec, scene, steps, worlds, fr, settle=settle)
        log = run_trace(cand, prec, scene, steps, worlds, fc, mode=mode, settle=settle, resync=fr if lockstep else None)
        res = compare(read_trace(fc), read_trace(fr))
        res["log"] = log[-300:]
        return res


def parity_golden(cand, golden_path, scene, prec, steps, worlds=1, settle=0, mode=None, lockstep=False):
    """Compare a candidate driver against a committed reference trace (tests/golden).
    lockstep: every step starts from the golden trace's pre-step body state (SURVEY 8d protocol, K = 1)."""
    with tempfile.TemporaryDirectory() as td:
        fc = os.path.join(td, "cand.bin")
        log = run_trace(cand, prec, scene, steps, worlds, fc, settle=settle, mode=mode, resync=golden_path if lockstep else None)
        res = compare(read_trace(fc), read_trace(golden_path))
        res["log"] = log[-300:]
        return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cand", default="b200")
    ap.add_argument("--prec", default="single")
    ap.add_argument("--scene", nargs="*", default=["free6", "stack32", "mixed", "mixed_maxc4", "block64", "tower64"])
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--worlds", type=int, default=1)
    ap.add_argument("--mode", default=None)
    ap.add_argument("--lockstep", action="store_true")
    a = ap.parse_args()
    bad = 0
    for sc in a.scene:
        try:
            r = parity(a.cand, a.prec, sc, a.steps, a.worlds, a.mode, lockstep=a.lockstep)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"scene": sc, "error": str(e)[-600:]}))
            bad += 1
            continue
        r["scene"] = sc
        r["prec"] = a.prec
        print(json.dumps(r))
        if not r["exact_ok"]:
            bad += 1
    sys.exit(1 if bad else 0)

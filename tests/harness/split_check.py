"""Split a large world over `--ranks` ranks inside ONE process (a host thread per rank) and compare every rank with an
unsplit batch, bit for bit, after each chunk of steps.  Exit 0 = equal, 3 = the ranks' kernels never met (loop-back on
one GPU without co-residency), 1 = mismatch / error.  Run as a process of its own by tests/test_split_world.py (one
precision per process: the single and double libraries export the same symbols)."""
import argparse
import sys

import numpy as np

import split_util as su


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="hostsim")
    ap.add_argument("--prec", default="single")
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--scene", default="pile_10x10x20")
    ap.add_argument("--steps", default="60,5,5")
    a = ap.parse_args()
    devices = [int(x) for x in a.devices.split(",")]
    lib, scenes = su.load(a.backend, a.prec)
    ref = su.build(lib, scenes, a.scene, devices[0])
    ranks = [su.build(lib, scenes, a.scene, devices[r % len(devices)]) for r in range(a.ranks)]
    handles = [su.export(lib, B) for B in ranks]
    for r, B in enumerate(ranks):
        su.attach(lib, B, r, handles)
    want = None
    for n in [int(x) for x in a.steps.split(",")]:
        su.step(lib, ref, 0.01, n)
        try:
            su.step_ranks_in_threads(lib, ranks, 0.01, n)
        except RuntimeError as e:
            print(e)
            return 3 if "within the timeout" in str(e) else 1
        want = su.state(lib, ref)
        if not np.isfinite(want).all():
            print("non-finite state")
            return 1
        for r, B in enumerate(ranks):
            if su.state(lib, B).tobytes() != want.tobytes():
                print(f"rank {r}/{a.ranks} differs from the unsplit sweep after {n} more steps")
                return 1
    if not np.abs(want[:, 7:10]).max() > 1e-3:   # the pile must be moving: not a comparison of a world at rest
        print("world at rest")
        return 1
    print("SPLIT_EQUAL", a.backend, a.prec, a.ranks)
    return 0


if __name__ == "__main__":
    sys.exit(main())

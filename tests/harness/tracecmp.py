"""Read and compare scene_driver traces (see tests/harness/scene_driver.cpp for the layout).

Exact observables (broadphase pair sequence, contact count, per-contact geom ids,
space-list order, LCG seed) are compared for equality; floating observables
(contacts, body state, joint feedback) are reported both as bitwise-equal counts
and as max relative error  |a-b|_inf / max(1,|b|_inf).
"""
import struct
import sys

import numpy as np


def read_trace(path):
    data = open(path, "rb").read()
    magic, rs, nworlds, nsteps = struct.unpack_from("<Iiii", data, 0)
    assert magic == 0x5254444F, "bad trace magic"
    rt = np.float32 if rs == 4 else np.float64
    off = 16
    steps = []

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(data, dtype=dtype, count=n, offset=off)
        off += a.nbytes
        return a

    for s in range(nsteps):
        ws = []
        for w in range(nworlds):
            rec = {}
            ng = int(take(np.int32, 1)[0])
            rec["glist"] = take(np.int32, ng)
            nb = int(take(np.int32, 1)[0])
            rec["state0"] = take(rt, nb * 13).reshape(nb, 13)
            np_ = int(take(np.int32, 1)[0])
            rec["pairs"] = take(np.int32, 2 * np_).reshape(np_, 2)
            nc = int(take(np.int32, 1)[0])
            cdt = np.dtype([("g", np.int32, 2), ("d", rt, 7)])
            c = take(cdt, nc)
            rec["cg"] = c["g"].reshape(nc, 2)
            rec["cd"] = c["d"].reshape(nc, 7)
            rec["fb"] = take(rt, nc * 6).reshape(nc, 6)
            rec["state1"] = take(rt, nb * 13).reshape(nb, 13)
            rec["seed"] = int(take(np.uint32, 1)[0])
            ws.append(rec)
        steps.append(ws)
    return {"realsize": rs, "nworlds": nworlds, "nsteps": nsteps, "steps": steps}


def _relerr(a, b):
    if a.size == 0:
        return 0.0
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    if not m.all():
        if not np.array_equal(np.isnan(a), np.isnan(b)):
            return float("inf")
    d = np.abs(np.where(m, a - b, 0.0)).max()
    return float(d / max(1.0, np.abs(np.where(m, b, 0.0)).max()))


def _bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def _biteq(a, b):
    """bitwise equality, except that any NaN equals any NaN: x86 SSE produces the default
    NaN 0xffc00000 for invalid operations, CUDA 0x7fffffff — payloads are not arithmetic"""
    return (_bits(a) == _bits(b)) | (np.isnan(a) & np.isnan(b))


def compare(ta, tb, verbose=False, stop_on_mismatch=True):
    """Returns a dict of summary statistics. ta = candidate, tb = reference."""
    assert ta["realsize"] == tb["realsize"] and ta["nworlds"] == tb["nworlds"]
    n = min(ta["nsteps"], tb["nsteps"])
    out = {
        "steps": n, "exact_ok": True, "first_exact_mismatch": None,
        "pairs": 0, "contacts": 0,
        "contact_bits_equal": 0, "contact_vals": 0, "state_bits_equal": 0, "state_vals": 0,
        "fb_bits_equal": 0, "fb_vals": 0,
        "max_contact_relerr": 0.0, "max_state_relerr": 0.0, "max_fb_relerr": 0.0,
        "first_state_bit_mismatch": None,
    }
    for s in range(n):
        for w in range(ta["nworlds"]):
            a, b = ta["steps"][s][w], tb["steps"][s][w]
            for key in ("glist", "pairs", "cg"):
                if a[key].shape != b[key].shape or not np.array_equal(a[key], b[key]):
                    out["exact_ok"] = False
                    if out["first_exact_mismatch"] is None:
                        out["first_exact_mismatch"] = (s, w, key)
                        if verbose:
                            print("MISMATCH step", s, "world", w, key)
                            print(" cand:", a[key].tolist()[:40])
                            print(" ref :", b[key].tolist()[:40])
            if a["seed"] != b["seed"]:
                out["exact_ok"] = False
                if out["first_exact_mismatch"] is None:
                    out["first_exact_mismatch"] = (s, w, "seed")
            if not _biteq(a["state0"], b["state0"]).all() and out["first_state_bit_mismatch"] is None:
                out["first_state_bit_mismatch"] = (s, w, "state0")
            out["pairs"] += len(b["pairs"])
            out["contacts"] += len(b["cg"])
            if a["cd"].shape == b["cd"].shape:
                out["contact_vals"] += b["cd"].size
                out["contact_bits_equal"] += int(_biteq(a["cd"], b["cd"]).sum())
                out["max_contact_relerr"] = max(out["max_contact_relerr"], _relerr(a["cd"], b["cd"]))
                out["fb_vals"] += b["fb"].size
                out["fb_bits_equal"] += int(_biteq(a["fb"], b["fb"]).sum())
                out["max_fb_relerr"] = max(out["max_fb_relerr"], _relerr(a["fb"], b["fb"]))
            out["state_vals"] += b["state1"].size
            eq = _biteq(a["state1"], b["state1"])
            out["state_bits_equal"] += int(eq.sum())
            if not eq.all() and out["first_state_bit_mismatch"] is None:
                out["first_state_bit_mismatch"] = (s, w, "state1")
            out["max_state_relerr"] = max(out["max_state_relerr"], _relerr(a["state1"], b["state1"]))
        if stop_on_mismatch and not out["exact_ok"]:
            out["steps"] = s + 1
            break
    return out


def compare_large(ta, tb):
    """Large-world comparison (candidate ta vs reference tb), SURVEY 8d parity protocol in lock-step:
    * broadphase pairs compared as SETS — the reference's callback order depends on RadixSort
      temporal-coherence state tied to its island stepping order, which the large-world path does not
      carry.  Orientation (which geom is o1) must match too, except for pairs whose float axis-0 minima
      tie exactly (`pairs_flipped`, reported; the reference breaks such ties by that same state);
    * contacts of the equally oriented pairs, after a stable sort by (g1, g2) (the order inside a pair
      is the collider's): geom ids exact, pos/normal/depth bitwise;
    * body state after the step vs the reference: max relative error |a-b|_inf / max(1,|b|_inf) and RMS
      of the per-body velocity difference (the coloured SOR order differs from the reference's random order)."""
    assert ta["realsize"] == tb["realsize"] and ta["nworlds"] == tb["nworlds"] == 1
    n = min(ta["nsteps"], tb["nsteps"])
    out = {"steps": n, "pairs": 0, "contacts": 0, "pair_sets_equal": True, "pairs_flipped": 0, "contact_ids_equal": True,
           "contact_bits_equal": 0, "contact_vals": 0, "state0_bits_equal": True,
           "max_pos_relerr": 0.0, "max_vel_relerr": 0.0, "rms_dv": 0.0, "rms_dw": 0.0, "first_mismatch": None}
    sdv = sdw = 0.0
    nbod = 0
    for s in range(n):
        a, b = ta["steps"][s][0], tb["steps"][s][0]
        if not _biteq(a["state0"], b["state0"]).all():
            out["state0_bits_equal"] = False
        pa = set(map(tuple, a["pairs"].tolist()))
        pb = set(map(tuple, b["pairs"].tolist()))
        out["pairs"] += len(b["pairs"])
        flipped = {(y, x) for (x, y) in pa - pb}
        if flipped != pb - pa or len(pa) != len(a["pairs"]) or len(pb) != len(b["pairs"]):
            out["pair_sets_equal"] = False
            if out["first_mismatch"] is None:
                out["first_mismatch"] = (s, "pairs", sorted(pa - pb)[:5], sorted(pb - pa)[:5])
        out["pairs_flipped"] += len(flipped)
        out["contacts"] += len(b["cg"])

        def keep(rec, drop):
            if not drop or len(rec["cg"]) == 0:
                return rec["cg"], rec["cd"]
            m = np.array([tuple(g) not in drop for g in rec["cg"].tolist()], dtype=bool)
            return rec["cg"][m], rec["cd"][m]

        acg, acd = keep(a, pa - pb)
        bcg, bcd = keep(b, pb - pa)
        if acg.shape != bcg.shape:
            out["contact_ids_equal"] = False
            if out["first_mismatch"] is None:
                out["first_mismatch"] = (s, "ncontacts", len(acg), len(bcg))
        else:
            ia = np.lexsort((np.arange(len(acg)), acg[:, 1], acg[:, 0]))
            ib = np.lexsort((np.arange(len(bcg)), bcg[:, 1], bcg[:, 0]))
            if not np.array_equal(acg[ia], bcg[ib]):
                out["contact_ids_equal"] = False
                if out["first_mismatch"] is None:
                    out["first_mismatch"] = (s, "contact ids")
            else:
                out["contact_vals"] += bcd.size
                out["contact_bits_equal"] += int(_biteq(acd[ia], bcd[ib]).sum())
        out["max_pos_relerr"] = max(out["max_pos_relerr"], _relerr(a["state1"][:, :7], b["state1"][:, :7]))
        out["max_vel_relerr"] = max(out["max_vel_relerr"], _relerr(a["state1"][:, 7:], b["state1"][:, 7:]))
        x, y = a["state1"].astype(np.float64), b["state1"].astype(np.float64)
        sdv += float(((x[:, 7:10] - y[:, 7:10]) ** 2).sum())
        sdw += float(((x[:, 10:13] - y[:, 10:13]) ** 2).sum())
        nbod += len(x)
    out["rms_dv"] = (sdv / max(nbod, 1)) ** 0.5
    out["rms_dw"] = (sdw / max(nbod, 1)) ** 0.5
    return out


if __name__ == "__main__":
    r = compare(read_trace(sys.argv[1]), read_trace(sys.argv[2]), verbose=True)
    for k, v in r.items():
        print(f"{k}: {v}")

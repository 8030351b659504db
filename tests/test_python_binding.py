"""ode-0.12_b200/python/ode_b200.py (Python binding over the batched API + binary snapshot).
CPU: bound to the TEST-ONLY host build (tests/hostsim) so the binding's marshalling is exercised without a
GPU; GPU: the same against the product library."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, lib_path

sys.path.insert(0, os.path.join(ROOT, "ode-0.12_b200", "python"))
import ode_b200 as ode  # noqa: E402

HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "_build", "libode_b200_hostsim_single.so")


def _scene(lib, n):
    worlds = []
    rng = np.random.default_rng(5)
    for w in range(n):
        W = ode.World(lib)
        W.add_plane(0, 0, 1, 0)
        for i in range(6):
            W.add_box((0.1 * rng.standard_normal(), 0.1 * rng.standard_normal(), 0.4 + 0.55 * i), (0.5, 0.4, 0.3), density=2.0)
        W.add_sphere((0.3, 0.1, 4.0), 0.25)
        W.add_capsule((-0.4, 0.2, 4.5), 0.15, 0.5)
        worlds.append(W)
    return worlds


def _run(libpath):
    lib = ode.load(libpath)
    out = {}
    b = ode.Batch(lib, _scene(lib, 3))
    b.set_contact_policy(max_contacts=8, mode=ode.ContactBounce | ode.ContactSoftCFM, bounce=0.1, bounce_vel=0.1, soft_cfm=0.01)
    b.set_seeds(np.array([11, 22, 33], dtype=np.uint32))
    assert not b.step(0.01, 60).any()
    snap = os.path.join(os.environ.get("TMPDIR", "/tmp"), "ob_snapshot_%d.npz" % os.getpid())
    b.save(snap)
    assert not b.step(0.01, 40).any()
    out["final"] = b.get_state()
    out["counters"] = b.counters()
    # a fresh batch restored from the snapshot continues bit for bit
    b2 = ode.Batch(lib, _scene(lib, 3))
    b2.set_contact_policy(max_contacts=8, mode=ode.ContactBounce | ode.ContactSoftCFM, bounce=0.1, bounce_vel=0.1, soft_cfm=0.01)
    b2.load(snap)
    assert not b2.step(0.01, 40).any()
    out["restored"] = b2.get_state()
    os.remove(snap)
    b.destroy(); b2.destroy()
    return out


def _check(o):
    pos = o["final"][0]
    assert pos.shape == (3, 8, 3) and np.isfinite(pos).all() and pos[:, :, 2].min() > 0.05
    assert o["counters"]["contacts"] > 0 and o["counters"]["overflow_worlds"] == 0
    for a, b in zip(o["final"], o["restored"]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_binding_and_snapshot_on_host_build():
    _check(_run(HOSTSIM))


@pytest.mark.gpu
def test_binding_and_snapshot_on_gpu():
    _check(_run(lib_path("single")))


def test_batch_refuses_to_run_after_a_bound_world_changed():
    """a user batch whose worlds are destroyed or restructured must fail with an error instead of touching freed objects (host logic, CPU build)"""
    import ctypes

    lib = ctypes.CDLL(HOSTSIM)
    vp = ctypes.c_void_p
    for n in ("dWorldCreate", "dHashSpaceCreate", "dBodyCreate", "dCreateSphere", "dBatchCreate"):
        getattr(lib, n).restype = vp
    lib.dB200LastError.restype = ctypes.c_char_p
    lib.dHashSpaceCreate.argtypes = [vp]; lib.dBodyCreate.argtypes = [vp]; lib.dCreateSphere.argtypes = [vp, ctypes.c_float]
    lib.dGeomSetBody.argtypes = [vp, vp]; lib.dBatchCreate.argtypes = [ctypes.c_int, vp, vp, vp]
    lib.dBatchCollideAndQuickStep.argtypes = [vp, ctypes.c_float, ctypes.c_int, vp]
    lib.dBatchDestroy.argtypes = [vp]; lib.dBatchDownload.argtypes = [vp]; lib.dWorldDestroy.argtypes = [vp]; lib.dSpaceDestroy.argtypes = [vp]
    lib.dBodyDestroy.argtypes = [vp]
    w, s = lib.dWorldCreate(), lib.dHashSpaceCreate(None)
    b = lib.dBodyCreate(w)
    lib.dGeomSetBody(lib.dCreateSphere(s, 0.5), b)
    B = vp(lib.dBatchCreate(1, (vp * 1)(w), (vp * 1)(s), None))
    assert B
    assert lib.dBatchCollideAndQuickStep(B, 0.01, 2, None) == 0
    b2 = lib.dBodyCreate(w)                                   # structural change of a bound world
    assert lib.dBatchCollideAndQuickStep(B, 0.01, 1, None) != 0 and b"destroyed or added" in lib.dB200LastError()
    assert lib.dBatchDownload(B) != 0
    lib.dWorldDestroy(w)                                       # the batch must not dereference the dead world when it goes
    lib.dBatchDestroy(B)
    lib.dSpaceDestroy(s)


def test_contact_policy_table_is_validated():
    """dBatchSetContactPolicy takes 1 to 8 rows (host logic, CPU build); the rows themselves are exercised by the scene `crashwall`"""
    import ctypes

    lib = ctypes.CDLL(HOSTSIM)
    vp = ctypes.c_void_p
    for n in ("dWorldCreate", "dHashSpaceCreate", "dBodyCreate", "dCreateSphere", "dBatchCreate"):
        getattr(lib, n).restype = vp
    lib.dB200LastError.restype = ctypes.c_char_p
    lib.dHashSpaceCreate.argtypes = [vp]; lib.dBodyCreate.argtypes = [vp]; lib.dCreateSphere.argtypes = [vp, ctypes.c_float]
    lib.dGeomSetBody.argtypes = [vp, vp]; lib.dBatchCreate.argtypes = [ctypes.c_int, vp, vp, vp]
    lib.dBatchSetContactPolicy.argtypes = [vp, vp, ctypes.c_int]; lib.dBatchDestroy.argtypes = [vp]
    w, s = lib.dWorldCreate(), lib.dHashSpaceCreate(None)
    lib.dGeomSetBody(lib.dCreateSphere(s, 0.5), lib.dBodyCreate(w))
    B = vp(lib.dBatchCreate(1, (vp * 1)(w), (vp * 1)(s), None))
    assert B
    rows = ctypes.create_string_buffer(9 * 256)   # zeroed rows, larger than any dBatchContactPolicy
    assert lib.dBatchSetContactPolicy(B, rows, 0) != 0
    assert lib.dBatchSetContactPolicy(B, rows, 9) != 0 and b"policy rows" in lib.dB200LastError()
    assert lib.dBatchSetContactPolicy(B, rows, 1) == 0
    assert lib.dBatchSetContactPolicy(B, rows, 8) == 0
    lib.dBatchDestroy(B)


def _policy_rows_check(libpath):
    """semantics of the policy table on the batched path (ob_policy_row, the per-pair rule of k_broad / the host mirror):
    first matching row serves the pair, no matching row = no contacts, a single row serves every pair, skip_static_pairs"""
    import ctypes

    import numpy as np

    lib = ctypes.CDLL(libpath)
    vp, ci, cf, ul = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_ulong

    class Surface(ctypes.Structure):
        _fields_ = [("mode", ci)] + [(n, cf) for n in ("mu", "mu2", "bounce", "bounce_vel", "soft_erp", "soft_cfm", "motion1", "motion2", "motionN", "slip1", "slip2")]

    class Policy(ctypes.Structure):
        _fields_ = [("cat_mask1", ul), ("cat_mask2", ul), ("max_contacts", ci), ("skip_if_connected", ci), ("skip_static_pairs", ci), ("surface", Surface)]

    for n in ("dWorldCreate", "dSimpleSpaceCreate", "dBodyCreate", "dCreateSphere", "dCreateBox", "dCreatePlane", "dBatchCreate"):
        getattr(lib, n).restype = vp
    lib.dSimpleSpaceCreate.argtypes = [vp]; lib.dBodyCreate.argtypes = [vp]; lib.dCreateSphere.argtypes = [vp, cf]
    lib.dCreateBox.argtypes = [vp, cf, cf, cf]; lib.dCreatePlane.argtypes = [vp, cf, cf, cf, cf]
    lib.dGeomSetBody.argtypes = [vp, vp]; lib.dBodySetPosition.argtypes = [vp, cf, cf, cf]; lib.dGeomSetPosition.argtypes = [vp, cf, cf, cf]
    lib.dGeomSetCategoryBits.argtypes = [vp, ul]; lib.dWorldSetGravity.argtypes = [vp, cf, cf, cf]
    lib.dBatchCreate.argtypes = [ci, vp, vp, vp]; lib.dBatchSetContactPolicy.argtypes = [vp, vp, ci]
    lib.dBatchCollideAndQuickStep.argtypes = [vp, cf, ci, vp]; lib.dBatchDebugContacts.argtypes = [vp, ci, vp, vp, ci]
    lib.dBatchDestroy.argtypes = [vp]

    def contacts(rows):
        w, s = lib.dWorldCreate(), lib.dSimpleSpaceCreate(None)
        lib.dWorldSetGravity(w, 0, 0, -9.81)
        plane = lib.dCreatePlane(s, 0, 0, 1, 0); lib.dGeomSetCategoryBits(plane, 1)          # geom 0
        bs = lib.dBodyCreate(w); lib.dBodySetPosition(bs, 0, 0, 0.49)
        sph = lib.dCreateSphere(s, 0.5); lib.dGeomSetCategoryBits(sph, 2); lib.dGeomSetBody(sph, bs)   # geom 1: sphere on the plane
        bb = lib.dBodyCreate(w); lib.dBodySetPosition(bb, 3, 0, 0.49)
        box = lib.dCreateBox(s, 1, 1, 1); lib.dGeomSetCategoryBits(box, 4); lib.dGeomSetBody(box, bb)    # geom 2: box on the plane
        stat = lib.dCreateBox(s, 1, 1, 1); lib.dGeomSetCategoryBits(stat, 8); lib.dGeomSetPosition(stat, -3, 0, 0.4)   # geom 3: static box in the plane
        B = vp(lib.dBatchCreate(1, (vp * 1)(w), (vp * 1)(s), None))
        assert B
        tab = (Policy * len(rows))()
        for i, (m1, m2, skip_static) in enumerate(rows):
            tab[i].cat_mask1, tab[i].cat_mask2, tab[i].max_contacts, tab[i].skip_static_pairs = m1, m2, 8, skip_static
            tab[i].surface.mu = 1.0
        assert lib.dBatchSetContactPolicy(B, tab, len(rows)) == 0
        assert lib.dBatchCollideAndQuickStep(B, 0.01, 1, None) == 0
        g = np.zeros((64, 2), dtype=np.int32); pnd = np.zeros((64, 7), dtype=np.float32)
        n = lib.dBatchDebugContacts(B, 0, pnd.ctypes.data, g.ctypes.data, 64)
        lib.dBatchDestroy(B)
        return sorted({tuple(sorted(map(int, g[i]))) for i in range(n)})

    ALL = 0xFFFFFFFF
    assert contacts([(ALL, ALL, 0)]) == [(0, 1), (0, 2), (0, 3)]                       # one row: every pair, the static pair included
    assert contacts([(ALL, ALL, 1)]) == [(0, 1), (0, 2)]                               # skip_static_pairs
    assert contacts([(2, ALL, 0), (4, 1, 0)]) == [(0, 1), (0, 2)]                      # two specific rows, no catch-all: the static pair has no row
    assert contacts([(2, ALL, 0)] * 2) == [(0, 1)]                                     # only pairs with the sphere are served
    assert contacts([(8, ALL, 1), (ALL, ALL, 0)]) == [(0, 1), (0, 2)]                  # the FIRST matching row decides (row 0 skips the static pair)


def test_contact_policy_rows_select_pairs_by_category():
    _policy_rows_check(HOSTSIM)


@pytest.mark.gpu
def test_contact_policy_rows_select_pairs_by_category_on_gpu():
    _policy_rows_check(lib_path("single"))

"""ode_b200 — ctypes binding over libode_b200_{single,double}.so.

The reference ships a Cython wrapper of its C API (bindings/python/ode.pyx: World, Body, Space, Geom*,
Mass, joints).  This module binds the same C symbols for the subset of that surface the hot path needs and
adds what the reference lacks: the batched entry points (`Batch`) with bulk state I/O into numpy arrays
and a binary snapshot (`Batch.save` / `Batch.load`, npz).

    import ode_b200 as ode
    lib = ode.load()                                  # libode_b200_single.so next to this package
    worlds = [ode.World(lib) for _ in range(4096)]    # ... build bodies / geoms through the classic API
    batch = ode.Batch(lib, worlds)
    batch.set_contact_policy(max_contacts=8, mu=float("inf"), mode=ode.ContactBounce, bounce=0.1)
    batch.step(0.01, 100)
    pos, quat, lvel, avel = batch.get_state()         # (worlds, bodies, 3|4|3|3) numpy arrays

Nothing here computes physics; every call lands in the C ABI declared in include/ode_b200/ode.h.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ContactMu2, ContactFDir1, ContactBounce, ContactSoftERP, ContactSoftCFM = 0x001, 0x002, 0x004, 0x008, 0x010
ContactSlip1, ContactSlip2, ContactApprox1 = 0x100, 0x200, 0x3000
SAP_AXES_XYZ = 0 | (1 << 2) | (2 << 4)


class Lib:
    """the loaded shared library + the precision it was built for"""

    def __init__(self, path, double=False):
        self.c = ctypes.CDLL(path)
        self.real = ctypes.c_double if double else ctypes.c_float
        self.np_real = np.float64 if double else np.float32
        c, R, vp, ci = self.c, self.real, ctypes.c_void_p, ctypes.c_int
        c.dInitODE2.argtypes = [ctypes.c_uint]
        c.dInitODE2(0)
        for name in ("dWorldCreate", "dJointGroupCreate"):
            getattr(c, name).restype = vp
        c.dJointGroupCreate.argtypes = [ci]
        for name in ("dHashSpaceCreate", "dSimpleSpaceCreate"):
            getattr(c, name).restype = vp
            getattr(c, name).argtypes = [vp]
        c.dSweepAndPruneSpaceCreate.restype = vp
        c.dSweepAndPruneSpaceCreate.argtypes = [vp, ci]
        c.dWorldSetGravity.argtypes = [vp, R, R, R]
        c.dWorldSetERP.argtypes = [vp, R]
        c.dWorldSetCFM.argtypes = [vp, R]
        c.dWorldSetQuickStepNumIterations.argtypes = [vp, ci]
        c.dWorldSetQuickStepW.argtypes = [vp, R]
        c.dWorldDestroy.argtypes = [vp]
        c.dSpaceDestroy.argtypes = [vp]
        c.dBodyCreate.restype = vp
        c.dBodyCreate.argtypes = [vp]
        c.dBodySetPosition.argtypes = [vp, R, R, R]
        c.dBodySetLinearVel.argtypes = [vp, R, R, R]
        c.dBodySetAngularVel.argtypes = [vp, R, R, R]
        c.dBodySetQuaternion.argtypes = [vp, vp]
        c.dBodySetMass.argtypes = [vp, vp]
        c.dMassSetBox.argtypes = [vp, R, R, R, R]
        c.dMassSetSphere.argtypes = [vp, R, R]
        c.dMassSetCapsule.argtypes = [vp, R, ci, R, R]
        c.dCreateBox.restype = vp
        c.dCreateBox.argtypes = [vp, R, R, R]
        c.dCreateSphere.restype = vp
        c.dCreateSphere.argtypes = [vp, R]
        c.dCreateCapsule.restype = vp
        c.dCreateCapsule.argtypes = [vp, R, R]
        c.dCreatePlane.restype = vp
        c.dCreatePlane.argtypes = [vp, R, R, R, R]
        c.dGeomSetBody.argtypes = [vp, vp]
        c.dB200LastError.restype = ctypes.c_char_p
        c.dBatchCreate.restype = vp
        c.dBatchCreate.argtypes = [ci, vp, vp, vp]
        c.dBatchDestroy.argtypes = [vp]
        c.dBatchSetContactPolicy.argtypes = [vp, vp, ci]
        c.dBatchSetSeeds.argtypes = [vp, vp]
        c.dBatchGetSeeds.argtypes = [vp, vp]
        c.dBatchCollideAndQuickStep.argtypes = [vp, R, ci, vp]
        c.dBatchNumBodies.argtypes = [vp]
        c.dBatchGetBodyState.argtypes = [vp] * 5
        c.dBatchSetBodyState.argtypes = [vp] * 5
        c.dBatchAddForces.argtypes = [vp] * 3
        c.dBatchGetCounters.argtypes = [vp, vp]
        c.dBatchDownload.argtypes = [vp]
        c.dBatchOrderStateSize.argtypes = [vp]
        c.dBatchGetOrderState.argtypes = [vp, vp]
        c.dBatchSetOrderState.argtypes = [vp, vp]

    def error(self):
        e = self.c.dB200LastError()
        return e.decode() if e else ""


def load(path=None, double=False):
    if path is None:
        path = os.path.join(_HERE, "..", "lib", "libode_b200_%s.so" % ("double" if double else "single"))
    if not os.path.exists(path):
        raise OSError("%s not found: build it with `python __graft_entry__.py` (there is no CPU fallback)" % path)
    return Lib(path, double)


class _Mass(ctypes.Structure):   # dMass, include/ode/mass.h:116-130 (dReal mass; dVector3 c; dMatrix3 I)
    pass


def _mass_type(real):
    class M(ctypes.Structure):
        _fields_ = [("mass", real), ("c", real * 4), ("I", real * 12)]
    return M


class World:
    """a dWorldID with its own space and contact joint group (one simulated scene)"""

    def __init__(self, lib, space="hash", gravity=(0, 0, -9.81), erp=0.2, cfm=1e-5, iterations=20, w=1.3):
        self.lib, c = lib, lib.c
        self.id = c.dWorldCreate()
        self.space = {"hash": lambda: c.dHashSpaceCreate(None), "simple": lambda: c.dSimpleSpaceCreate(None),
                      "sap": lambda: c.dSweepAndPruneSpaceCreate(None, SAP_AXES_XYZ)}[space]()
        c.dWorldSetGravity(self.id, *gravity)
        c.dWorldSetERP(self.id, erp)
        c.dWorldSetCFM(self.id, cfm)
        c.dWorldSetQuickStepNumIterations(self.id, iterations)
        c.dWorldSetQuickStepW(self.id, w)
        self.bodies = []

    def _body(self, pos, mass_fn):
        c = self.lib.c
        b = c.dBodyCreate(self.id)
        c.dBodySetPosition(b, *pos)
        m = _mass_type(self.lib.real)()
        mass_fn(ctypes.byref(m))
        c.dBodySetMass(b, ctypes.byref(m))
        self.bodies.append(b)
        return b

    def add_box(self, pos, sides, density=1.0):
        c = self.lib.c
        b = self._body(pos, lambda m: c.dMassSetBox(m, density, *sides))
        c.dGeomSetBody(c.dCreateBox(self.space, *sides), b)
        return b

    def add_sphere(self, pos, radius, density=1.0):
        c = self.lib.c
        b = self._body(pos, lambda m: c.dMassSetSphere(m, density, radius))
        c.dGeomSetBody(c.dCreateSphere(self.space, radius), b)
        return b

    def add_capsule(self, pos, radius, length, density=1.0):
        c = self.lib.c
        b = self._body(pos, lambda m: c.dMassSetCapsule(m, density, 3, radius, length))
        c.dGeomSetBody(c.dCreateCapsule(self.space, radius, length), b)
        return b

    def add_plane(self, a, b, c_, d):
        return self.lib.c.dCreatePlane(self.space, a, b, c_, d)


class Batch:
    """dBatchID: the bound worlds live on the device; bulk state I/O as numpy arrays in API (creation) order"""

    def __init__(self, lib, worlds, max_contacts_per_world=0, device=0, large_world=False):
        self.lib, self.n = lib, len(worlds)
        desc = (ctypes.c_int * 8)(max_contacts_per_world, device, 0, 1 if large_world else 0, 0, 0, 0, 0)   # dBatchDesc
        wv = (ctypes.c_void_p * self.n)(*[w.id for w in worlds])
        sv = (ctypes.c_void_p * self.n)(*[w.space for w in worlds])
        self.id = lib.c.dBatchCreate(self.n, wv, sv, desc)
        if not self.id:
            raise RuntimeError("dBatchCreate: " + lib.error())
        self.nb = lib.c.dBatchNumBodies(self.id)
        self.worlds = worlds

    def set_contact_policy(self, max_contacts=8, skip_if_connected=True, mode=0, mu=float("inf"), mu2=0.0, bounce=0.0,
                           bounce_vel=0.0, soft_erp=0.0, soft_cfm=0.0, slip1=0.0, slip2=0.0):
        R = self.lib.real

        class Surface(ctypes.Structure):   # dSurfaceParameters, include/ode/contact.h:52-68
            _fields_ = [("mode", ctypes.c_int), ("mu", R), ("mu2", R), ("bounce", R), ("bounce_vel", R), ("soft_erp", R),
                        ("soft_cfm", R), ("motion1", R), ("motion2", R), ("motionN", R), ("slip1", R), ("slip2", R)]

        class Policy(ctypes.Structure):    # dBatchContactPolicy
            _fields_ = [("cat_mask1", ctypes.c_ulong), ("cat_mask2", ctypes.c_ulong), ("max_contacts", ctypes.c_int),
                        ("skip_if_connected", ctypes.c_int), ("skip_static_pairs", ctypes.c_int), ("surface", Surface)]

        p = Policy(~0 & 0xFFFFFFFFFFFFFFFF, ~0 & 0xFFFFFFFFFFFFFFFF, max_contacts, int(skip_if_connected), 0,
                   Surface(mode, mu, mu2, bounce, bounce_vel, soft_erp, soft_cfm, 0, 0, 0, slip1, slip2))
        if self.lib.c.dBatchSetContactPolicy(self.id, ctypes.byref(p), 1) != 0:
            raise RuntimeError("dBatchSetContactPolicy: " + self.lib.error())

    def set_seeds(self, seeds):
        a = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert a.shape == (self.n,)
        self.lib.c.dBatchSetSeeds(self.id, a.ctypes.data)

    def get_seeds(self):
        a = np.zeros(self.n, dtype=np.uint32)
        self.lib.c.dBatchGetSeeds(self.id, a.ctypes.data)
        return a

    def step(self, h, nsteps=1):
        status = np.zeros(self.n, dtype=np.int32)
        if self.lib.c.dBatchCollideAndQuickStep(self.id, h, nsteps, status.ctypes.data) != 0:
            raise RuntimeError("dBatchCollideAndQuickStep: " + self.lib.error())
        return status

    def get_state(self):
        t = self.lib.np_real
        arrs = [np.zeros((self.n, self.nb, k), dtype=t) for k in (3, 4, 3, 3)]
        if self.lib.c.dBatchGetBodyState(self.id, *[a.ctypes.data for a in arrs]) != 0:
            raise RuntimeError("dBatchGetBodyState: " + self.lib.error())
        return arrs

    def set_state(self, pos=None, quat=None, lvel=None, avel=None):
        t = self.lib.np_real
        ptrs, keep = [], []
        for a, k in ((pos, 3), (quat, 4), (lvel, 3), (avel, 3)):
            if a is None:
                ptrs.append(None)
            else:
                a = np.ascontiguousarray(a, dtype=t)
                assert a.shape == (self.n, self.nb, k)
                keep.append(a)
                ptrs.append(a.ctypes.data)
        if self.lib.c.dBatchSetBodyState(self.id, *ptrs) != 0:
            raise RuntimeError("dBatchSetBodyState: " + self.lib.error())

    def add_forces(self, force=None, torque=None):
        t = self.lib.np_real
        f = None if force is None else np.ascontiguousarray(force, dtype=t)
        q = None if torque is None else np.ascontiguousarray(torque, dtype=t)
        self.lib.c.dBatchAddForces(self.id, None if f is None else f.ctypes.data, None if q is None else q.ctypes.data)

    def counters(self):
        c = (ctypes.c_longlong * 7)()
        self.lib.c.dBatchGetCounters(self.id, ctypes.byref(c))
        return dict(zip(["steps", "body_steps", "pairs", "contacts", "rows", "islands", "overflow_worlds"], list(c)))

    # binary snapshot: body state, the per-world dRand streams and the order state (space list order, SAP ranks) —
    # everything the next step depends on besides the scene itself; a restored batch continues bit for bit
    def save(self, path):
        pos, quat, lvel, avel = self.get_state()
        order = np.zeros(self.lib.c.dBatchOrderStateSize(self.id), dtype=np.int32)
        self.lib.c.dBatchGetOrderState(self.id, order.ctypes.data)
        np.savez(path, pos=pos, quat=quat, lvel=lvel, avel=avel, seeds=self.get_seeds(), order=order)

    def load(self, path):
        z = np.load(path)
        self.set_state(z["pos"], z["quat"], z["lvel"], z["avel"])
        self.set_seeds(z["seeds"])
        order = np.ascontiguousarray(z["order"], dtype=np.int32)
        assert order.size == self.lib.c.dBatchOrderStateSize(self.id)
        self.lib.c.dBatchSetOrderState(self.id, order.ctypes.data)

    def download(self):
        """write the device state back into the dBodyID / dGeomID objects"""
        self.lib.c.dBatchDownload(self.id)

    def destroy(self):
        if self.id:
            self.lib.c.dBatchDestroy(self.id)
            self.id = None

"""The reference's own known-answer tests for this path, replayed through the C ABI:

* ode-0.12/tests/collision.cpp:4-102  test_collision_trimesh_sphere_exact -- a sphere that barely touches the diagonal edge
  of a two-triangle square: dCollide(trimesh, sphere) must return 2 contacts of depth exactly 0 with the triangle normal,
  also after translating both geoms and after rotating the mesh by 90 degrees.
* ode-0.12/tests/joint.cpp:79-217      test_hinge2GetInfo1 -- the number of constraint rows a hinge2 joint reports
  (Info1.m = 4, +1 when axis 1 is at a stop or powered, +1 when axis 2 is powered) through a sequence of poses and
  parameter changes.  getInfo1 is internal to the reference; here m is observed as the row count of one step of a world
  that holds only this joint (the lambda tap of the batched API).

CPU: against the TEST-ONLY host build of the same code (tests/hostsim); GPU: against the product library."""
import ctypes
import math
import os

import numpy as np
import pytest

from conftest import ROOT, lib_path

HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "_build", "libode_b200_hostsim_{prec}.so")
LIBS = [pytest.param("hostsim", id="hostsim"), pytest.param("b200", marks=pytest.mark.gpu, id="b200")]


def _load(kind, prec):
    path = lib_path(prec) if kind == "b200" else HOSTSIM.format(prec=prec)
    lib = ctypes.CDLL(path)
    real = ctypes.c_float if prec == "single" else ctypes.c_double
    vp = ctypes.c_void_p
    for name in ("dWorldCreate", "dBodyCreate", "dJointCreateHinge2", "dGeomTriMeshDataCreate", "dCreateTriMesh", "dCreateSphere",
                 "dHashSpaceCreate", "dBatchCreate"):
        getattr(lib, name).restype = vp
    lib.dB200LastError.restype = ctypes.c_char_p
    lib.dBodyCreate.argtypes = [vp]
    lib.dBodySetPosition.argtypes = [vp, real, real, real]
    lib.dBodySetRotation.argtypes = [vp, vp]
    lib.dBodySetLinearVel.argtypes = [vp, real, real, real]
    lib.dBodySetAngularVel.argtypes = [vp, real, real, real]
    lib.dJointCreateHinge2.argtypes = [vp, vp]
    lib.dJointAttach.argtypes = [vp, vp, vp]
    lib.dJointSetHinge2Anchor.argtypes = [vp, real, real, real]
    lib.dJointSetHinge2Param.argtypes = [vp, ctypes.c_int, real]
    lib.dRFromAxisAndAngle.argtypes = [vp, real, real, real, real]
    lib.dHashSpaceCreate.argtypes = [vp]
    lib.dBatchCreate.argtypes = [ctypes.c_int, vp, vp, vp]
    lib.dBatchSetDebugTaps.argtypes = [vp, ctypes.c_int]
    lib.dBatchCollideAndQuickStep.argtypes = [vp, real, ctypes.c_int, vp]
    lib.dBatchDebugLambda.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int]
    lib.dBatchDestroy.argtypes = [vp]
    lib.dWorldDestroy.argtypes = [vp]
    lib.dGeomTriMeshDataBuildSingle.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int]
    lib.dCreateTriMesh.argtypes = [vp, vp, vp, vp, vp]
    lib.dCreateSphere.argtypes = [vp, real]
    lib.dGeomSetPosition.argtypes = [vp, real, real, real]
    lib.dGeomSetRotation.argtypes = [vp, vp]
    lib.dCollide.argtypes = [vp, vp, ctypes.c_int, vp, ctypes.c_int]
    return lib, real


@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("kind", LIBS)
def test_collision_trimesh_sphere_exact(kind, prec):
    lib, real = _load(kind, prec)

    class ContactGeom(ctypes.Structure):   # include/ode/contact.h:93-100 (dVector3 = dReal[4])
        _fields_ = [("pos", real * 4), ("normal", real * 4), ("depth", real), ("g1", ctypes.c_void_p), ("g2", ctypes.c_void_p),
                    ("side1", ctypes.c_int), ("side2", ctypes.c_int)]

    verts = np.array([-1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0], dtype=np.float32)   # a square on the XY plane
    idx = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)
    data = lib.dGeomTriMeshDataCreate()
    lib.dGeomTriMeshDataBuildSingle(data, verts.ctypes.data, 12, 4, idx.ctypes.data, 6, 12)
    trimesh = lib.dCreateTriMesh(None, data, None, None, None)
    radius = 4.0
    sphere = lib.dCreateSphere(None, radius)
    cg = (ContactGeom * 4)()

    def check(normal):
        nc = lib.dCollide(trimesh, sphere, 4, ctypes.byref(cg), ctypes.sizeof(ContactGeom))
        assert nc == 2, lib.dB200LastError()
        for i in range(nc):
            assert cg[i].depth == 0
            assert list(cg[i].normal)[:3] == normal

    lib.dGeomSetPosition(sphere, 0, 0, radius)          # the sphere touches the diagonal edge
    check([0, 0, -1])
    lib.dGeomSetPosition(trimesh, 10, 30, 40)           # both geoms translated
    lib.dGeomSetPosition(sphere, 10, 30, 40 + radius)
    check([0, 0, -1])
    rot = (real * 12)(1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 0, 0)   # the trimesh rotated 90 degrees about X
    lib.dGeomSetPosition(trimesh, 10, 30, 40)
    lib.dGeomSetRotation(trimesh, rot)
    lib.dGeomSetPosition(sphere, 10, 30 - radius, 40)
    check([0, 1, 0])


# dParam* (include/ode/common.h:275-330): group 1 values, group 2 = 0x100 + value
P_LOSTOP, P_HISTOP, P_FMAX = 0, 1, 3
P_FMAX2 = 0x100 + 3


@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("kind", LIBS)
def test_hinge2_getinfo1_row_counts(kind, prec):
    lib, real = _load(kind, prec)
    w = lib.dWorldCreate()
    b1 = lib.dBodyCreate(w)
    lib.dBodySetPosition(b1, 0, -1, 0)
    b2 = lib.dBodyCreate(w)
    lib.dBodySetPosition(b2, 0, 1, 0)
    j = lib.dJointCreateHinge2(w, None)
    lib.dJointAttach(j, b1, b2)
    lib.dJointSetHinge2Anchor(j, 0, 0, 0)
    space = lib.dHashSpaceCreate(None)

    def pose(pos, angle):
        R = (real * 12)()
        lib.dRFromAxisAndAngle(R, 1, 0, 0, angle)
        lib.dBodySetPosition(b2, *pos)
        lib.dBodySetRotation(b2, R)

    def rows():
        """Info1.m of the joint = rows of one step of this world (no contacts, no other joint); the step runs on a bound copy:
        the host objects keep the pose and parameters the test set"""
        for b in (b1, b2):
            lib.dBodySetLinearVel(b, 0, 0, 0)
            lib.dBodySetAngularVel(b, 0, 0, 0)
        wa, sa = (ctypes.c_void_p * 1)(w), (ctypes.c_void_p * 1)(space)
        B = lib.dBatchCreate(1, wa, sa, None)
        assert B, lib.dB200LastError()
        B = ctypes.c_void_p(B)
        lib.dBatchSetDebugTaps(B, 1)
        status = (ctypes.c_int * 1)()
        assert lib.dBatchCollideAndQuickStep(B, 1e-4, 1, status) == 0, lib.dB200LastError()
        lam = (real * 16)()
        m = lib.dBatchDebugLambda(B, 0, lam, 16)
        lib.dBatchDestroy(B)
        return m

    lo, hi = -math.pi / 4.0, math.pi / 4.0
    lib.dJointSetHinge2Param(j, P_LOSTOP, lo)
    lib.dJointSetHinge2Param(j, P_HISTOP, hi)
    assert rows() == 4                                   # original position, inside the limits
    pose((0, 0, 1), math.pi / 2.0)
    assert rows() == 5                                   # outside the lo stop
    pose((0, 1, 0), 0.0)
    assert rows() == 4                                   # back, limits kept
    pose((0, 0, 1), math.pi / 2.0)
    assert rows() == 5
    pose((0, 1, 0), 0.0)
    lib.dJointSetHinge2Param(j, P_LOSTOP, -2 * math.pi)  # back, limits removed
    lib.dJointSetHinge2Param(j, P_HISTOP, 2 * math.pi)
    assert rows() == 4
    lib.dJointSetHinge2Param(j, P_LOSTOP, lo)
    lib.dJointSetHinge2Param(j, P_HISTOP, hi)
    pose((0, 0, 1), -math.pi / 2.0)
    assert rows() == 5                                   # past the hi stop
    pose((0, 1, 0), 0.0)
    assert rows() == 4
    pose((0, 0, 1), -math.pi / 2.0)
    assert rows() == 5
    pose((0, 1, 0), -math.pi / 2.0)
    lib.dJointSetHinge2Param(j, P_LOSTOP, -2 * math.pi)
    lib.dJointSetHinge2Param(j, P_HISTOP, 2 * math.pi)
    assert rows() == 4
    lib.dJointSetHinge2Param(j, P_FMAX, 2)               # axis 1 powered
    assert rows() == 5
    lib.dJointSetHinge2Param(j, P_FMAX2, 2)              # axis 2 powered too
    assert rows() == 6
    lib.dJointSetHinge2Param(j, P_FMAX, 0)               # axis 1 unpowered
    assert rows() == 5
    lib.dWorldDestroy(w)

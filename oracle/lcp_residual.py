"""oracle/lcp_residual.py -- TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

CPU restatement (numpy) of the reference's LCP bookkeeping for contact joints, used to measure how well a sweep solved the
step's LCP -- the yardstick for the large-world path, whose graph-coloured sweep order cannot be bit-compared with
SOR_LCP's random order:

  multiply_invM_JT   ode/src/quickstep.cpp:138-160   out = invM * J^T * in        (per body 6-vectors)
  multiply_J         ode/src/quickstep.cpp:163-181   out = J * in
  multiply_J_invM_JT ode/src/quickstep.cpp:187-195   out = J * invM * J^T * in    (the reference's own test utility)
  contact rows       ode/src/joints/contact.cpp:74-256 (getInfo2), policy "crash" of tests/harness/scenes.h
  rhs / cfm scaling  ode/src/quickstep.cpp:840-857

Everything is reconstructed from a scene_driver trace (tests/harness/scene_driver.cpp): pre-step body state, the step's
contacts, the joint feedback (f1 = sum_rows J1l^T lambda, from which lambda follows because the three rows' directions
are orthonormal) and the post-step state.  Residual of row i after the sweeps, in velocity units (h * (A lambda - rhs)):

    w_i = J_i . v_new - c_i + cfm_i * lambda_i          (v_new = the step's output velocities)

and its complementarity violation  r_i = |w_i| (lo < lambda < hi),  max(0, -w_i) (lambda at lo),  max(0, w_i) (lambda at hi).
`residual()` also predicts v_new from the pre-step state with multiply_J_invM_JT's building blocks and reports how far
that is from the traced v_new: on a reference trace this pins the restatement itself (tests/test_large_world.py)."""
import numpy as np

GRAVITY = np.array([0.0, 0.0, -9.81])


def quat_to_R(q):
    """dRfromQ, ode/src/rotation.cpp:117-133 (row-major 3x3)"""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def plane_space(n):
    """dPlaneSpace, ode/src/odemath.cpp:139-166, vectorised"""
    p = np.zeros_like(n); q = np.zeros_like(n)
    big = np.abs(n[:, 2]) > np.sqrt(0.5)
    a = n[:, 1] ** 2 + n[:, 2] ** 2
    k = 1 / np.sqrt(np.where(big, a, 1))
    p[big, 1] = (-n[:, 2] * k)[big]; p[big, 2] = (n[:, 1] * k)[big]
    q[big, 0] = (a * k)[big]; q[big, 1] = (-n[:, 0] * p[:, 2])[big]; q[big, 2] = (n[:, 0] * p[:, 1])[big]
    a2 = n[:, 0] ** 2 + n[:, 1] ** 2
    k2 = 1 / np.sqrt(np.where(~big, a2, 1))
    s = ~big
    p[s, 0] = (-n[:, 1] * k2)[s]; p[s, 1] = (n[:, 0] * k2)[s]
    q[s, 0] = (-n[:, 2] * p[:, 1])[s]; q[s, 1] = (n[:, 2] * p[:, 0])[s]; q[s, 2] = (a2 * k2)[s]
    return p, q


def multiply_invM_JT(J, jb, invM, invIw, lam, nb):
    """quickstep.cpp:138-160: cforce[b] = sum over the rows on b of iMJ_row * lambda_row"""
    out = np.zeros((nb, 6))
    for side in (0, 1):
        b = jb[:, side]
        ok = b >= 0
        lin = J[ok, 6 * side:6 * side + 3] * invM[b[ok], None]
        ang = np.einsum("nij,nj->ni", invIw[b[ok]], J[ok, 6 * side + 3:6 * side + 6])
        np.add.at(out, b[ok], np.concatenate([lin, ang], axis=1) * lam[ok, None])
    return out


def multiply_J(J, jb, v):
    """quickstep.cpp:163-181: out_row = J1 . v[b1] + J2 . v[b2]"""
    out = np.einsum("nk,nk->n", J[:, 0:6], v[jb[:, 0]])
    ok = jb[:, 1] >= 0
    out[ok] += np.einsum("nk,nk->n", J[ok, 6:12], v[jb[ok, 1]])
    return out


def multiply_J_invM_JT(J, jb, invM, invIw, lam, nb):
    """quickstep.cpp:187-195"""
    return multiply_J(J, jb, multiply_invM_JT(J, jb, invM, invIw, lam, nb))


def body_tables(kinds):
    """kinds[b] = 's' (sphere r 0.25) or 'b' (box 0.5^3), density 1 (dMassSetSphere / dMassSetBox, ode/src/mass.cpp)"""
    m = np.empty(len(kinds)); I = np.empty(len(kinds))
    for b, k in enumerate(kinds):
        if k == "s":
            m[b] = 4.0 / 3.0 * np.pi * 0.25 ** 3; I[b] = 0.4 * m[b] * 0.25 ** 2
        else:
            m[b] = 0.5 ** 3; I[b] = m[b] / 12.0 * (0.25 + 0.25)
    return m, I


def residual(rec, kinds, geom_body, h, erp=0.8, cfm0=0.01, mu=0.5):
    """rec: one world-step of a trace (tracecmp.read_trace); kinds / geom_body describe the scene (geom_body[g] = body of geom g
    or -1).  Contact policy = policy_crash (Slip1|Slip2 with slip 0, SoftERP 0.8, SoftCFM 0.01, Approx1, mu 0.5)."""
    st0, st1 = rec["state0"].astype(np.float64), rec["state1"].astype(np.float64)
    nb = len(st0)
    m, I = body_tables(kinds)
    invM = 1 / m
    invIw = np.einsum("b,ij->bij", 1 / I, np.eye(3))                 # isotropic: R invI R^T = invI
    cg, cd, fb = rec["cg"], rec["cd"].astype(np.float64), rec["fb"].astype(np.float64)
    b1 = geom_body[cg[:, 0]]; b2 = geom_body[cg[:, 1]]
    rev = b1 < 0                                                    # dJointAttach swap rule (ode.cpp:1368-1377)
    j1 = np.where(rev, b2, b1); j2 = np.where(rev, -1, b2)
    keep = j1 >= 0
    j1, j2, cd, fb, rev = j1[keep], j2[keep], cd[keep], fb[keep], rev[keep]
    pos, n, depth = cd[:, 0:3], cd[:, 3:6] * np.where(rev, -1.0, 1.0)[:, None], cd[:, 6]
    t1, t2 = plane_space(n)
    r1 = pos - st0[j1, 0:3]
    r2 = pos - st0[np.maximum(j2, 0), 0:3]
    nc = len(pos)
    J = np.zeros((3 * nc, 12)); jb = np.zeros((3 * nc, 2), dtype=np.int64)
    for q, d in enumerate((n, t1, t2)):
        J[q::3, 0:3] = d; J[q::3, 3:6] = np.cross(r1, d)
        two = j2 >= 0
        J[q::3, 6:9] = np.where(two[:, None], -d, 0); J[q::3, 9:12] = np.where(two[:, None], -np.cross(r2, d), 0)
        jb[q::3, 0] = j1; jb[q::3, 1] = j2
    lam = np.empty(3 * nc)
    for q, d in enumerate((n, t1, t2)):
        lam[q::3] = np.einsum("nk,nk->n", fb[:, 0:3], d)            # f1 = sum_q dir_q lambda_q, directions orthonormal
    c = np.zeros(3 * nc); c[0::3] = erp / h * np.maximum(depth, 0)
    cfm = np.zeros(3 * nc); cfm[0::3] = cfm0
    lo = np.zeros(3 * nc); hi = np.full(3 * nc, np.inf)
    fr = mu * np.abs(lam[0::3])
    for q in (1, 2):
        hi[q::3] = fr; lo[q::3] = -fr
    v1 = np.concatenate([st1[:, 7:10], st1[:, 10:13]], axis=1)
    w = multiply_J(J, jb, v1) - c + cfm / h * lam * h                # cfm is scaled by 1/h in the LCP and by h again in velocity units
    tol = 1e-6 * np.maximum(1.0, np.abs(lam))
    at_lo = lam <= lo + tol; at_hi = lam >= hi - tol
    r = np.where(at_lo & ~at_hi, np.maximum(0, -w), np.where(at_hi & ~at_lo, np.maximum(0, w), np.abs(w)))
    # prediction of v_new from the pre-step state: v + h * invM * (f_ext + J^T lambda)   (quickstep.cpp:905-975)
    v0 = np.concatenate([st0[:, 7:10], st0[:, 10:13]], axis=1)
    fext = np.zeros((nb, 6)); fext[:, 0:3] = m[:, None] * GRAVITY     # isotropic inertia: the gyroscopic term w x (I w) vanishes
    acc = np.concatenate([fext[:, 0:3] * invM[:, None], np.einsum("bij,bj->bi", invIw, fext[:, 3:6])], axis=1)
    vpred = v0 + h * (acc + multiply_invM_JT(J, jb, invM, invIw, lam, nb))
    return {"rows": 3 * nc, "rms": float(np.sqrt(np.mean(r ** 2))) if nc else 0.0, "max": float(r.max()) if nc else 0.0,
            "rms_w_free": float(np.sqrt(np.mean(w[~at_lo & ~at_hi] ** 2))) if (~at_lo & ~at_hi).any() else 0.0,
            "vpred_err": float(np.abs(vpred - v1).max()), "max_depth": float(depth.max()) if nc else 0.0,
            "JMJt_check": float(np.abs(multiply_J(J, jb, vpred) - multiply_J(J, jb, v0 + h * acc) - h * multiply_J_invM_JT(J, jb, invM, invIw, lam, nb)).max()) if nc else 0.0}


def pile_scene(nx, ny, nz):
    """(kinds, geom_body) of scene_pile(nx, ny, nz): geoms 0..4 are the floor and wall planes, geom 5+n belongs to body n"""
    kinds = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                kinds.append("s" if (i + j + k) & 1 else "b")
    nb = len(kinds)
    geom_body = np.concatenate([np.full(5, -1), np.arange(nb)])
    return kinds, geom_body

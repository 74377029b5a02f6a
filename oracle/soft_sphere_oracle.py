"""TEST INFRASTRUCTURE - numpy/pure-Python DEFINITION of the soft-sphere extension (ibm_set_option "collision_model" = 1).

The reference ships no soft-sphere model (src/particleCloud/Make/files:4-14 names one, the sources are absent), so there is
nothing to be in parity WITH: this file states the contact law and the sub-step loop it sits in, in the evaluation order of
ibmfoam_b200/csrc/ibm_motion.cu (d5_pair_force, d5_solve), and the GPU path is checked against it bit for bit on small
free-space cases; the physics is checked against the closed form of the linear spring - dashpot oscillator
(tests/test_zzz_gpu_soft_sphere.py).  Only tests import this module.

Scope of the restatement: isotropic bodies (spheres, cylinders) away from walls and periodic patches, one instance per
index; hydrodynamic force / torque frozen over the step, gravity g, the literal 100/101 sub-step schedule of
particleCloud::solve (particleCloud.C:1247-1275) and the symplectic Euler update of particleIBM::solve (particleIBM.C:393-450)."""
import numpy as np


def substep_schedule(dt):
    """particleCloud.C:1247-1275 (App. B #22), the same scalar loop as motion_advance"""
    dts = []
    local_dt = dt / 100.0
    local_time = local_dt
    while local_time < dt:
        dts.append(local_dt)
        if local_time + local_dt > dt:
            local_dt = dt - local_time
        else:
            local_time += local_dt
    return dts


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def pair_force(x1, x2, v1, v2, w1, w2, r1, r2, kn, gn, gt, mu):
    """force and torque on body 1 and on body 2 (body 1 = the larger index, as in the kernel); None if not overlapping"""
    x12 = x1 - x2
    dist = float(np.sqrt(dot(x12, x12)))
    if not (r1 + r2 > dist) or not dist > 0.0:      # particleIBM.C:56-74 for an isotropic pair
        return None
    n = x12 / dist
    overlap = (r1 + r2) - dist
    a1 = -(r1 - 0.5 * overlap) * n
    a2 = (r2 - 0.5 * overlap) * n
    vrel = ((v1 - v2) + cross(w1, a1)) - cross(w2, a2)
    vn = dot(vrel, n)
    Fn = kn * overlap - gn * vn
    if Fn < 0.0:
        Fn = 0.0
    vt = vrel - vn * n
    vt_mag = float(np.sqrt(dot(vt, vt)))
    F1 = Fn * n
    if vt_mag > 0.0:
        Ft = min(gt * vt_mag, mu * Fn)
        F1 = F1 - (Ft / vt_mag) * vt
    return F1, cross(a1, F1), -F1, cross(a2, -F1)


def advance(state, dt, radius, mass, inertia_per_mass, kn, gn=0.0, gt=0.0, mu=0.0, g=(0.0, 0.0, 0.0), rotating=True):
    """one fluid step; state = dict(center, v, omega, theta, F, T) of [P, 3] float64 arrays, particle i has index i"""
    c, v, om, th = (np.array(state[k], dtype=np.float64) for k in ("center", "v", "omega", "theta"))
    Fh, Th = np.asarray(state["F"], dtype=np.float64), np.asarray(state["T"], dtype=np.float64)
    P = len(c)
    g = np.asarray(g, dtype=np.float64)
    for ldt in substep_schedule(dt):
        rows = [dict() for _ in range(P)]
        for i in range(P):
            for j in range(i):                     # index1 > index2 (particleCloud.C:500-503)
                r = pair_force(c[i], c[j], v[i], v[j], om[i], om[j], radius[i], radius[j], kn, gn, gt, mu)
                if r is None:
                    continue
                rows[i][j] = (r[0], r[1])
                rows[j][i] = (r[2], r[3])
        for i in range(P):
            sF, sT = np.zeros(3), np.zeros(3)
            for k in sorted(rows[i]):              # sum(collisionForces_) runs over the partner index
                sF = sF + rows[i][k][0]
                sT = sT + rows[i][k][1]
            F = Fh[i] + sF + np.zeros(3)           # (+ the wall sum, empty here)
            v[i] = v[i] + ldt * (F / mass[i] + g)
            c[i] = c[i] + ldt * v[i]
            if rotating:
                T = Th[i] + sT + np.zeros(3)
                om[i] = om[i] + ldt * T / (mass[i] * inertia_per_mass[i])
                th[i] = th[i] + om[i] * ldt
                if th[i][2] < 0:
                    th[i][2] += 2.0 * np.pi
                if th[i][2] >= 2.0 * np.pi:
                    th[i][2] -= 2.0 * np.pi
    return {"center": c, "v": v, "omega": om, "theta": th}


def restitution(kn, gn, m_eff):
    """closed form of the linear spring - dashpot contact: e = exp(-pi zeta / sqrt(1 - zeta^2)), zeta = gn / (2 sqrt(kn m_eff))
    (velocity ratio at the end of the half period; the clipping of attractive forces ends the contact slightly earlier)"""
    zeta = gn / (2.0 * np.sqrt(kn * m_eff))
    return float(np.exp(-np.pi * zeta / np.sqrt(1.0 - zeta * zeta)))


def restitution_clipped(kn, gn, m_eff, v0=1.0, h=None):
    """restitution of the law AS APPLIED (attractive forces clipped, Fn = max(0, kn delta - gn v_n)): the clipping ends a
    strongly damped contact before the half period, so less is dissipated than the closed form above says.  Fine
    integration (velocity Verlet, 2*10^5 steps per undamped half period) of  m_eff delta'' = -max(0, kn delta + gn delta')."""
    h = h or (np.pi / np.sqrt(kn / m_eff)) / 2.0e5
    delta, vel = 0.0, float(v0)            # overlap and its rate
    def acc(d, u):
        return -max(0.0, kn * d + gn * u) / m_eff
    while True:
        a0 = acc(delta, vel)
        d1 = delta + h * vel + 0.5 * h * h * a0
        u_pred = vel + h * a0
        a1 = acc(d1, u_pred)
        vel = vel + 0.5 * h * (a0 + a1)
        delta = d1
        if delta <= 0.0:
            return -vel / v0

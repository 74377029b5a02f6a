import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
from oracle.pyoracle import OracleCloud
import ibm_testutil as util
for name in ('cyl30','col'):
    case = cases.cylinder_re30() if name=='cyl30' else cases.colliding_particles()
    rng = np.random.default_rng(3); n = case.particles.n
    case.particles.v[:] = rng.normal(0, 0.3, (n, 3)); case.particles.omega[:] = rng.normal(0, 0.2, (n, 3))
    case.particles.v[:, 2] = 0; case.particles.omega[:, :2] = 0
    f = util.caller_fields(case)
    oc, gc = OracleCloud(case), ParticleCloud(case)
    print(name, 'v7_ok', gc.counter('v7_ok'))
    Fo,_ = oc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    Fg = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
    print(' slow', gc.counter('v7_slow_count'), 'nan in Fg', np.isnan(Fg).sum(), 'inf', np.isinf(Fg).sum(), 'nan in Fo', np.isnan(Fo).sum(), 'nonzero Fg', (Fg!=0).sum(), 'nonzero Fo', (Fo!=0).sum())
    bad = np.argwhere(~np.isfinite(Fg))
    print(' first bad', bad[:10].tolist())
    ok = np.isfinite(Fg).all(axis=1)
    print(' rel err on finite cells', np.max(np.abs(Fg[ok]-Fo[ok]))/np.max(np.abs(Fo)))
    gc.set_kernel_variant(2,2)
    Fg2 = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
    print(' v2 rel err', util.rel_err(Fg2, Fo))
    gc.close()

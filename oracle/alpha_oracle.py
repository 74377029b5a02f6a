"""ORACLE — TEST INFRASTRUCTURE ONLY.  Solid volume fraction by sub-cell sampling, plain numpy.

There is NO reference counterpart (the reference only has the binary indicator I.ibm, particleCloud.C:952-969): this is
the definition the extension kernel ibm_kernels_alpha.cuh implements, restated independently -
    alpha_s[cell] = min(1, sum_instances  #{sample points of the cell inside the body} / nSub^D)
sample points x = C + ((q + 0.5)/nSub - 0.5) * delta per direction (D = 3; D = 2 in the plane for cylinders / ellipses),
inside test |x - c|^2 < R^2 (sphere, cylinder) or (x'/a)^2 + (y'/b)^2 < 1 in the body frame (ellipse).  Every sample of
every cell of the bounding box is evaluated (no shortcuts), in the same IEEE operation order as the kernel, so the
comparison is exact.  Validated against the analytic body volume in tests/test_emu_v5.py."""
import numpy as np


def solid_fraction(mesh, kind, centres, d=0.0, a=0.0, b=0.0, rot=None, n_sub=8):
    """mesh: cases.MeshSpec; returns alpha in LATTICE order (x fastest)."""
    nx, ny, nz = mesh.nx, mesh.ny, mesh.nz
    ox, oy, oz = (np.float64(v) for v in mesh.origin)
    dx, dy, dz = (np.float64(v) for v in mesh.delta)
    alpha = np.zeros((nz, ny, nx))
    flat = kind != 0
    R = max(a, b) if kind == 2 else d / 2.0
    inv = np.float64(1.0) / np.float64(n_sub)
    off = (np.arange(n_sub, dtype=np.float64) + 0.5) * inv - 0.5
    total = float(n_sub * n_sub * (1 if flat else n_sub))
    for p, c in enumerate(np.asarray(centres, dtype=np.float64)):
        i0, i1 = max(int(np.floor((c[0] - R - ox) / dx)), 0), min(int(np.floor((c[0] + R - ox) / dx)), nx - 1)
        j0, j1 = max(int(np.floor((c[1] - R - oy) / dy)), 0), min(int(np.floor((c[1] + R - oy) / dy)), ny - 1)
        if flat:
            k0, k1 = 0, nz - 1
        else:
            k0, k1 = max(int(np.floor((c[2] - R - oz) / dz)) - mesh.kz0, 0), min(int(np.floor((c[2] + R - oz) / dz)) - mesh.kz0, nz - 1)
        if i1 < i0 or j1 < j0 or k1 < k0:
            continue
        Cx = ox + (np.arange(i0, i1 + 1) + 0.5) * dx
        Cy = oy + (np.arange(j0, j1 + 1) + 0.5) * dy
        Cz = oz + (np.arange(k0, k1 + 1) + mesh.kz0 + 0.5) * dz
        X = Cx[:, None] + (off * dx)[None, :]            # [cell i][sample]
        Y = Cy[:, None] + (off * dy)[None, :]
        ddx = X - c[0]
        ddy = Y - c[1]
        if kind == 0:
            Z = Cz[:, None] + (off * dz)[None, :]
            ddz = Z - c[2]
            r2 = (ddx * ddx)[None, None, :, None, None, :] + (ddy * ddy)[None, :, None, None, :, None]
            r2 = r2 + (ddz * ddz)[:, None, None, :, None, None]                 # [k][j][i][sz][sy][sx]
            cnt = (r2 < R * R).sum(axis=(3, 4, 5))
        elif kind == 1:
            r2 = (ddx * ddx)[None, :, None, :] + (ddy * ddy)[:, None, :, None]   # [j][i][sy][sx]
            cnt = np.broadcast_to((r2 < R * R).sum(axis=(2, 3))[None], (k1 - k0 + 1, j1 - j0 + 1, i1 - i0 + 1))
        else:
            Rm = np.asarray(rot[p], dtype=np.float64).reshape(3, 3)
            DX = ddx[None, :, None, :]
            DY = ddy[:, None, :, None]
            bx = Rm[0, 0] * DX + Rm[1, 0] * DY
            by = Rm[0, 1] * DX + Rm[1, 1] * DY
            u, v = bx / a, by / b
            cnt = np.broadcast_to(((u * u + v * v) < 1.0).sum(axis=(2, 3))[None], (k1 - k0 + 1, j1 - j0 + 1, i1 - i0 + 1))
        alpha[k0:k1 + 1, j0:j1 + 1, i0:i1 + 1] += cnt / total
    return np.minimum(alpha, 1.0).ravel()

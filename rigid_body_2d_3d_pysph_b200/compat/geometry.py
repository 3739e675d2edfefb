"""Lattice/tank builders with the ``pysph.tools.geometry`` signatures.

[upstream, restated] SURVEY.md App. C-9.  The reference scenes are assembled
from these three helpers (/root/reference/code/geometry.py:3,
/root/reference/code/stack_of_cylinders.py:20).  ``get_2d_block`` and
``get_3d_block`` follow the published ``mgrid`` definition exactly;
``get_2d_tank`` is frozen here as nested U-shaped layers (base row plus two
side columns per layer, layers added outwards) because the upstream source is
not available to compare against -- scenes that depend on it are stored as
arrays in tests/golden so the layout cannot drift.
"""
import numpy as np


def get_2d_block(dx=0.01, length=1.0, height=1.0, center=np.array([0., 0.])):
    n1 = int(length / dx) + 1
    n2 = int(height / dx) + 1
    x, y = np.mgrid[-length / 2.:length / 2.:n1 * 1j,
                    -height / 2.:height / 2.:n2 * 1j]
    x, y = np.ravel(x), np.ravel(y)
    return x + center[0], y + center[1]


def get_3d_block(dx=0.01, length=1.0, height=1.0, depth=1.0,
                 center=np.array([0., 0., 0.])):
    n1 = int(length / dx) + 1
    n2 = int(height / dx) + 1
    n3 = int(depth / dx) + 1
    x, y, z = np.mgrid[-length / 2.:length / 2.:n1 * 1j,
                       -height / 2.:height / 2.:n2 * 1j,
                       -depth / 2.:depth / 2.:n3 * 1j]
    x, y, z = np.ravel(x), np.ravel(y), np.ravel(z)
    return x + center[0], y + center[1], z + center[2]


def _count(extent, dx):
    return int(np.floor(extent / dx + 1e-9))


def get_2d_tank(dx=0.05, base_center=np.array([0.0, 0.0]), length=1.0,
                height=1.0, num_layers=1, outside=True, staggered=False,
                top=False):
    """Open-top tank: base parallel to x, side walls parallel to y."""
    sign = 1.0 if outside else -1.0
    xs, ys = [], []
    for k in range(num_layers):
        lk = length + 2.0 * sign * k * dx
        hk = height + sign * k * dx
        y0 = -sign * k * dx
        nb = _count(lk, dx)
        xb = np.arange(nb + 1) * dx - lk / 2.
        yb = np.full_like(xb, y0)
        nw = _count(hk, dx)
        yw = y0 + (np.arange(nw) + 1) * dx
        xl = np.full_like(yw, -lk / 2.)
        xr = np.full_like(yw, xb[-1])
        xs += [xl, xb, xr]
        ys += [yw, yb, yw]
        if top:
            xs.append(xb)
            ys.append(np.full_like(xb, y0 + (nw + 1) * dx))
    x = np.concatenate(xs)
    y = np.concatenate(ys)
    return x + base_center[0], y + base_center[1]


def remove_overlap_particles(pa, solid, dx_solid, dim=2):
    """[upstream] drop particles of ``pa`` closer than dx_solid to ``solid``."""
    from scipy.spatial import cKDTree
    if dim == 2:
        a = np.c_[pa.x, pa.y]
        b = np.c_[solid.x, solid.y]
    else:
        a = np.c_[pa.x, pa.y, pa.z]
        b = np.c_[solid.x, solid.y, solid.z]
    d, _ = cKDTree(b).query(a)
    pa.remove_particles(np.where(d < dx_solid * (1.0 - 1e-9))[0])

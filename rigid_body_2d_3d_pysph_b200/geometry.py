"""Tank / dam builders -- same names as the reference's ``code/geometry.py``
(scene fixtures for the benchmark scripts; host NumPy)."""
import numpy as np

from .compat.geometry import get_2d_block, get_2d_tank, get_3d_block


def hydrostatic_tank_2d(fluid_length, fluid_height, tank_height, tank_layers,
                        fluid_spacing, tank_spacing):
    """geometry.py:6-24"""
    xt, yt = get_2d_tank(dx=tank_spacing,
                         length=fluid_length + 2. * tank_spacing,
                         height=tank_height, num_layers=tank_layers)
    xf, yf = get_2d_block(dx=fluid_spacing, length=fluid_length,
                          height=fluid_height, center=[-1.5, 1])
    xf += (np.min(xt) - np.min(xf))
    yf -= (np.min(yf) - np.min(yt))
    xf += tank_spacing * (tank_layers)
    yf += tank_spacing * (tank_layers)
    return xf, yf, xt, yt


def get_fluid_tank_3d(fluid_length, fluid_height, fluid_depth, tank_length,
                      tank_height, tank_layers, fluid_spacing, tank_spacing,
                      hydrostatic=False):
    """geometry.py:27-102: a block plus left/right/front/back/bottom walls
    (length along x, height along y, depth along z)."""
    xf, yf, zf = get_3d_block(dx=fluid_spacing, length=fluid_length,
                              height=fluid_height, depth=fluid_depth)
    thick = tank_spacing * (tank_layers - 1)

    def side():
        return get_3d_block(dx=fluid_spacing, length=thick,
                            height=tank_height, depth=fluid_depth)
    xl, yl, zl = side()
    xr, yr, zr = side()
    xl += np.min(xf) - np.max(xl) - tank_spacing
    yl += np.min(yf) - np.min(yl) + 0. * tank_spacing
    xr += np.max(xf) - np.min(xr) + tank_spacing
    if hydrostatic is False:
        xr += tank_length - fluid_length
    yr += np.min(yf) - np.min(yr) + 0. * tank_spacing
    span = np.max(xr) - np.min(xl)

    def face():
        return get_3d_block(dx=fluid_spacing, length=span,
                            height=tank_height, depth=thick)
    xfr, yfr, zfr = face()
    xfr += np.min(xl) - np.min(xfr)
    yfr += np.min(yf) - np.min(yfr) + 0. * tank_spacing
    zfr += np.max(zl) - np.min(zfr) + tank_spacing * 1
    xbk, ybk, zbk = face()
    xbk += np.min(xl) - np.min(xbk)
    ybk += np.min(yf) - np.min(ybk) + 0. * tank_spacing
    zbk += np.min(zl) - np.max(zbk) - tank_spacing * 1
    xbt, ybt, zbt = get_3d_block(dx=fluid_spacing, length=span, height=thick,
                                 depth=np.max(zfr) - np.min(zbk))
    xbt += np.min(xl) - np.min(xbt)
    ybt += np.min(yl) - np.max(ybt) - tank_spacing * 1
    xt = np.concatenate([xl, xr, xfr, xbk, xbt])
    yt = np.concatenate([yl, yr, yfr, ybk, ybt])
    zt = np.concatenate([zl, zr, zfr, zbk, zbt])
    return xf, yf, zf, xt, yt, zt


def create_tank_2d_from_block_2d(xf, yf, tank_length, tank_height,
                                 tank_spacing, tank_layers):
    """geometry.py:105-135"""
    xleft, yleft = get_2d_block(dx=tank_spacing,
                                length=(tank_layers - 1) * tank_spacing,
                                height=tank_height, center=[0., 0.])
    xleft += min(xf) - max(xleft) - tank_spacing
    yleft += min(yf) - min(yleft)
    xright = xleft + abs(min(xleft)) + tank_length + tank_spacing
    yright = yleft
    xbottom, ybottom = get_2d_block(dx=tank_spacing,
                                    length=max(xright) - min(xleft),
                                    height=(tank_layers - 1) * tank_spacing,
                                    center=[0., 0.])
    xbottom += min(xleft) - min(xbottom)
    ybottom += min(yleft) - max(ybottom) - tank_spacing
    x = np.concatenate([xleft, xright, xbottom])
    y = np.concatenate([yleft, yright, ybottom])
    return x, y

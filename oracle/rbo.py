"""TEST INFRASTRUCTURE -- ctypes front end of the C oracle (oracle/rbo.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs import this.  It operates in place on the NumPy
arrays of host ``ParticleArray`` objects (property names of SURVEY.md App. B).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'librbo.so')
_SRC = os.path.join(_HERE, 'rbo.c')

_lib = None


def build(force=False):
    """gcc -O3 -fopenmp -ffp-contract=off."""
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ['gcc', '-O3', '-fopenmp', '-ffp-contract=off', '-fPIC',
               '-shared', '-std=c99', '-o', _SO, _SRC, '-lm']
        subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib, ND, NI, DNAMES, INAMES, RboArray
    if _lib is not None:
        return _lib
    build()
    L = ctypes.CDLL(_SO)
    L.rbo_dnames.restype = ctypes.POINTER(ctypes.c_char_p)
    L.rbo_inames.restype = ctypes.POINTER(ctypes.c_char_p)

    def names(p):
        out = []
        i = 0
        while p[i]:
            out.append(p[i].decode())
            i += 1
        return out
    DNAMES = names(L.rbo_dnames())
    INAMES = names(L.rbo_inames())
    ND, NI = len(DNAMES), len(INAMES)

    class RboArray(ctypes.Structure):
        _fields_ = [('n', ctypes.c_int32), ('nb', ctypes.c_int32),
                    ('tnb', ctypes.c_int32), ('rigid', ctypes.c_int32),
                    ('ks', ctypes.c_int32), ('limit', ctypes.c_int32),
                    ('spacing0', ctypes.c_double),
                    ('d', ctypes.c_void_p * ND),
                    ('i', ctypes.c_void_p * NI)]
    globals()['RboArray'] = RboArray
    L.rbo_quintic.restype = ctypes.c_double
    L.rbo_quintic.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double]
    L.rbo_nnps_pairs.restype = ctypes.c_int64
    _lib = L
    return L


class RboParams(ctypes.Structure):
    _fields_ = [('dim', ctypes.c_int32), ('radius_scale', ctypes.c_double),
                ('kr', ctypes.c_double), ('kf', ctypes.c_double),
                ('fric_coeff', ctypes.c_double), ('gx', ctypes.c_double),
                ('gy', ctypes.c_double), ('gz', ctypes.c_double),
                ('dt', ctypes.c_double), ('eta_uniform', ctypes.c_double)]


def make_params(dim, dt, kr=1e5, kf=1e3, fric_coeff=0.5, gx=0., gy=0., gz=0.,
                radius_scale=3.0, eta_uniform=-1.0):
    return RboParams(dim, radius_scale, kr, kf, fric_coeff, gx, gy, gz, dt,
                     eta_uniform)


def _get(pa, name):
    if name in pa.properties:
        return pa.properties[name]
    if name in pa.constants:
        return pa.constants[name]
    return None


def pack(arrays, rigid_names, ks=0):
    """ParticleArray list -> (ctypes array of RboArray, keepalive)."""
    lib()
    out = (RboArray * len(arrays))()
    keep = []
    for k, pa in enumerate(arrays):
        r = out[k]
        r.n = pa.get_number_of_particles()
        r.rigid = 1 if pa.name in rigid_names else 0
        r.nb = int(pa.constants['nb'][0]) if 'nb' in pa.constants else 0
        r.tnb = int(pa.constants['total_no_bodies'][0]) \
            if 'total_no_bodies' in pa.constants else 0
        r.ks = ks if r.rigid else 0
        r.limit = int(pa.constants['max_tng_contacts_limit'][0]) \
            if 'max_tng_contacts_limit' in pa.constants else 0
        r.spacing0 = float(pa.constants['spacing0'][0]) \
            if 'spacing0' in pa.constants else 0.0
        for j, name in enumerate(DNAMES):
            a = _get(pa, name)
            if a is None:
                r.d[j] = None
                continue
            assert a.dtype == np.float64 and a.flags.c_contiguous, name
            keep.append(a)
            r.d[j] = a.ctypes.data
        for j, name in enumerate(INAMES):
            a = _get(pa, name)
            if a is None:
                r.i[j] = None
                continue
            assert a.dtype == np.int32 and a.flags.c_contiguous, \
                (name, a.dtype)
            keep.append(a)
            r.i[j] = a.ctypes.data
    return out, keep


def add_sparse_history(pa, ks):
    """Per-particle sparse slot history used instead of the dense tnb-strided
    slot arrays (needed when nb is large; SURVEY.md section 7 'Sparse slots')."""
    pa.add_property('sp_key', type='int', stride=ks, default=-1)
    pa.add_property('sp_delta_lt', stride=3 * ks)
    pa.add_property('sp_fn', stride=3 * ks)


def quintic(dim, rij, h):
    return lib().rbo_quintic(dim, rij, h)


def nnps_pairs(arrays, idst, isrc, radius_scale=3.0):
    """Neighbour CSR (offsets, idx) of arrays[idst] among arrays[isrc]."""
    L = lib()
    arr, keep = pack(arrays, [])
    n = arrays[idst].get_number_of_particles()
    off = np.zeros(n + 1, dtype=np.int64)
    tot = L.rbo_nnps_pairs(arr, len(arrays), idst, isrc,
                           ctypes.c_double(radius_scale),
                           off.ctypes.data_as(ctypes.c_void_p), None)
    idx = np.zeros(max(tot, 1), dtype=np.int32)
    L.rbo_nnps_pairs(arr, len(arrays), idst, isrc,
                     ctypes.c_double(radius_scale),
                     off.ctypes.data_as(ctypes.c_void_p),
                     idx.ctypes.data_as(ctypes.c_void_p))
    return off, idx[:tot]


def contact(arrays, rigid_names, params, ks=0):
    L = lib()
    arr, keep = pack(arrays, rigid_names, ks)
    counts = np.zeros(2, dtype=np.int64)
    err = L.rbo_contact(arr, len(arrays), ctypes.byref(params),
                        counts.ctypes.data_as(ctypes.c_void_p))
    if err:
        raise RuntimeError('oracle: slot capacity exceeded')
    return counts


def gtvf_stage(pa, stage, dt, planar=False):
    L = lib()
    arr, keep = pack([pa], [pa.name])
    L.rbo_gtvf_stage(arr, stage, ctypes.c_double(dt), int(planar))


def rk2_stage(pa, stage, dt, fix_q7=False):
    L = lib()
    arr, keep = pack([pa], [pa.name])
    L.rbo_rk2_stage(arr, stage, ctypes.c_double(dt), int(fix_q7))


def gtvf_step(arrays, rigid_names, params, planar=False, ks=0, nsteps=1):
    L = lib()
    arr, keep = pack(arrays, rigid_names, ks)
    counts = np.zeros(2, dtype=np.int64)
    for _ in range(nsteps):
        err = L.rbo_gtvf_step(arr, len(arrays), ctypes.byref(params),
                              int(planar),
                              counts.ctypes.data_as(ctypes.c_void_p))
        if err:
            raise RuntimeError('oracle: slot capacity exceeded')
    return counts


def rk2_step(arrays, rigid_names, params, fix_q7=False, ks=0, nsteps=1):
    L = lib()
    arr, keep = pack(arrays, rigid_names, ks)
    counts = np.zeros(2, dtype=np.int64)
    for _ in range(nsteps):
        err = L.rbo_rk2_step(arr, len(arrays), ctypes.byref(params),
                             int(fix_q7),
                             counts.ctypes.data_as(ctypes.c_void_p))
        if err:
            raise RuntimeError('oracle: slot capacity exceeded')
    return counts


def canelas(arrays, rigid_names, params, Cn=1.4e-5):
    """BodyForce + RigidBodyCanelasRigidRigid/RigidWall + SumUpExternalForces
    (rigid_body_common.py:115-125, 244-628, 128-175); E and poisson_ratio are
    the arrays' constants."""
    L = lib()
    arr, keep = pack(arrays, rigid_names)
    E = np.array([float(pa.constants['E'][0]) for pa in arrays])
    nu = np.array([float(pa.constants['poisson_ratio'][0]) for pa in arrays])
    L.rbo_canelas.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_double]
    L.rbo_canelas(arr, len(arrays), ctypes.byref(params),
                  E.ctypes.data_as(ctypes.c_void_p),
                  nu.ctypes.data_as(ctypes.c_void_p), float(Cn))


def num_threads():
    return lib().rbo_num_threads()


def set_num_threads(n):
    """OpenMP threads of the following calls (n <= 0: leave as is)."""
    lib().rbo_set_num_threads(int(n))
    return num_threads()


def dem_step(arrays, granular_names, params, nsteps=1):
    """GTVF step of the DEMScheme path (LVCDisplacement + DEMStep)."""
    L = lib()
    arr, keep = pack(arrays, granular_names)
    for _ in range(nsteps):
        if L.rbo_dem_step(arr, len(arrays), ctypes.byref(params)):
            raise RuntimeError('oracle: tangential contact list full')

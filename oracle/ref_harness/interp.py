"""TEST INFRASTRUCTURE -- reference-as-oracle harness, part 2: interpreter.

Executes PySPH-style ``Equation`` groups and ``IntegratorStep`` steppers as
plain scalar Python on NumPy arrays, following the upstream call order
restated in SURVEY.md App. C-4/5/6:

* groups in list order; inside a group, per destination array: every
  equation's ``initialize`` (per particle), then per source array
  ``initialize_pair`` and ``loop`` (per neighbour pair) / ``loop_all``, then
  ``post_loop`` (per particle), then ``reduce(dst, t, dt)`` once;
* steppers: ``py_stageN(dst, t, dt)`` first, then ``stageN`` per particle;
* GTVF: stage1, eval(0), stage2, eval(1), stage3 -- no initial acceleration.

Neighbours are brute force O(N^2) with the App. C-1 predicate, listed in
ascending source index (this fixes the closest-point tie rule Q6: lowest
source index among equal RIJ, first source array first).
"""
import inspect

import numpy as np


def neighbours(dst, src, radius_scale):
    """CSR neighbour lists of every dst particle among src particles."""
    dx = dst.x[:, None] - src.x[None, :]
    dy = dst.y[:, None] - src.y[None, :]
    dz = dst.z[:, None] - src.z[None, :]
    r2 = dx * dx + dy * dy + dz * dz
    rs2 = radius_scale * radius_scale
    hi2 = (rs2 * dst.h * dst.h)[:, None]
    hj2 = (rs2 * src.h * src.h)[None, :]
    mask = (r2 < hi2) | (r2 < hj2)
    return [np.nonzero(row)[0] for row in mask]


class _Binder(object):
    """Resolve a method's parameter names to arrays / per-pair symbols."""

    def __init__(self, method):
        self.method = method
        self.names = [p for p in inspect.signature(method).parameters]

    def static_args(self, dst, src, t, dt, kernel):
        out = {}
        for n in self.names:
            if n.startswith('d_') and n != 'd_idx':
                out[n] = _lookup(dst, n[2:])
            elif n.startswith('s_') and n != 's_idx':
                out[n] = _lookup(src, n[2:])
            elif n == 't':
                out[n] = t
            elif n == 'dt':
                out[n] = dt
            elif n == 'SPH_KERNEL':
                out[n] = kernel
        return out


def _lookup(pa, name):
    if name in pa.properties:
        return pa.properties[name]
    if name in pa.constants:
        return pa.constants[name]
    raise KeyError('%s has no %s' % (pa.name, name))


_PAIR = ('XIJ', 'RIJ', 'R2IJ', 'HIJ', 'WIJ', 'DWIJ', 'VIJ')


def run_groups(groups, arrays, kernel, t, dt, nbr_log=None):
    pas = dict((a.name, a) for a in arrays)
    rs = kernel.radius_scale
    for group in groups:
        dests = []
        for eq in group.equations:
            if eq.dest not in dests:
                dests.append(eq.dest)
        for dname in dests:
            dst = pas[dname]
            eqs = [e for e in group.equations if e.dest == dname]
            n = dst.get_number_of_particles()
            # initialize
            for eq in eqs:
                if hasattr(eq, 'initialize'):
                    b = _Binder(eq.initialize)
                    kw = b.static_args(dst, None, t, dt, kernel)
                    for i in range(n):
                        eq.initialize(d_idx=i, **kw)
            srcs = []
            for eq in eqs:
                for s in (eq.sources or []):
                    if s not in srcs:
                        srcs.append(s)
            for sname in srcs:
                src = pas[sname]
                seqs = [e for e in eqs if e.sources and sname in e.sources]
                nbrs = None
                for eq in seqs:
                    if hasattr(eq, 'initialize_pair'):
                        b = _Binder(eq.initialize_pair)
                        kw = b.static_args(dst, src, t, dt, kernel)
                        for i in range(n):
                            eq.initialize_pair(d_idx=i, **kw)
                loops = [e for e in seqs if hasattr(e, 'loop')]
                loop_alls = [e for e in seqs if hasattr(e, 'loop_all')]
                if loops or loop_alls:
                    nbrs = neighbours(dst, src, rs)
                    if nbr_log is not None:
                        nbr_log[(dname, sname)] = nbrs
                if loops:
                    binders = [(_Binder(e.loop), e) for e in loops]
                    kws = [b.static_args(dst, src, t, dt, kernel)
                           for b, e in binders]
                    need = set()
                    for b, e in binders:
                        need.update(x for x in b.names if x in _PAIR)
                    for i in range(n):
                        for j in nbrs[i]:
                            pair = _pair_symbols(dst, src, i, int(j), kernel,
                                                 need)
                            for (b, e), kw in zip(binders, kws):
                                extra = dict((k, pair[k]) for k in b.names
                                             if k in pair)
                                e.loop(d_idx=i, s_idx=int(j), **kw, **extra)
                for eq in loop_alls:
                    b = _Binder(eq.loop_all)
                    kw = b.static_args(dst, src, t, dt, kernel)
                    for i in range(n):
                        nb = np.asarray(nbrs[i], dtype=np.int64)
                        eq.loop_all(d_idx=i, NBRS=nb, N_NBRS=len(nb), **kw)
            for eq in eqs:
                if hasattr(eq, 'post_loop'):
                    b = _Binder(eq.post_loop)
                    kw = b.static_args(dst, None, t, dt, kernel)
                    for i in range(n):
                        eq.post_loop(d_idx=i, **kw)
            for eq in eqs:
                if hasattr(eq, 'reduce'):
                    eq.reduce(_Dst(dst), t, dt)


def _pair_symbols(dst, src, i, j, kernel, need):
    xij = np.array([dst.x[i] - src.x[j], dst.y[i] - src.y[j],
                    dst.z[i] - src.z[j]])
    r2 = xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2]
    rij = np.sqrt(r2)
    hij = 0.5 * (dst.h[i] + src.h[j])
    out = {'XIJ': xij, 'R2IJ': r2, 'RIJ': rij, 'HIJ': hij}
    if 'WIJ' in need:
        out['WIJ'] = kernel.kernel(xij, rij, hij)
    if 'DWIJ' in need:
        out['DWIJ'] = kernel.gradient(xij, rij, hij, [0., 0., 0.])
    if 'VIJ' in need:
        out['VIJ'] = np.array([dst.u[i] - src.u[j], dst.v[i] - src.v[j],
                               dst.w[i] - src.w[j]])
    return out


class _Dst(object):
    """What PySPH hands to ``reduce`` / ``py_stage*``: attribute access to
    the NumPy views of properties and constants."""

    def __init__(self, pa):
        self.__dict__['_pa'] = pa

    def __getattr__(self, name):
        return _lookup(self.__dict__['_pa'], name)

    def __setattr__(self, name, value):
        _lookup(self.__dict__['_pa'], name)[:] = value


def run_stage(stepper, stage, pa, t, dt):
    """``py_<stage>`` then ``<stage>`` per particle (App. C-5)."""
    py = getattr(stepper, 'py_' + stage, None)
    if py is not None:
        py(_Dst(pa), t, dt)
    fn = getattr(stepper, stage, None)
    if fn is None:
        return
    b = _Binder(fn)
    if not [x for x in b.names if x != 'self']:
        fn()
        return
    kw = b.static_args(pa, None, t, dt, None)
    for i in range(pa.get_number_of_particles()):
        fn(d_idx=i, **kw)


def gtvf_step(steppers, equations, arrays, kernel, t, dt, nbr_log=None):
    """One ``GTVFIntegrator.one_timestep`` (App. C-6)."""
    pas = dict((a.name, a) for a in arrays)
    for name, st in steppers.items():
        run_stage(st, 'stage1', pas[name], t, dt)
    run_groups(equations.groups[0], arrays, kernel, t, dt)
    for name, st in steppers.items():
        run_stage(st, 'stage2', pas[name], t, dt)
    run_groups(equations.groups[1], arrays, kernel, t, dt, nbr_log)
    for name, st in steppers.items():
        run_stage(st, 'stage3', pas[name], t, dt)


def epec_step(steppers, groups, arrays, kernel, t, dt):
    """``EPECIntegrator.one_timestep``: initialize, eval, stage1, eval,
    stage2 (App. C-6); the equations list is single-stage."""
    pas = dict((a.name, a) for a in arrays)
    for name, st in steppers.items():
        run_stage(st, 'initialize', pas[name], t, dt)
    run_groups(groups, arrays, kernel, t, dt)
    for name, st in steppers.items():
        run_stage(st, 'stage1', pas[name], t, dt)
    run_groups(groups, arrays, kernel, t + dt / 2, dt)
    for name, st in steppers.items():
        run_stage(st, 'stage2', pas[name], t, dt)

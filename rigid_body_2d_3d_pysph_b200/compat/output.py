"""Solver output files: ``dump`` / ``load`` / ``iter_output`` / ``get_files``.

[upstream, restated] ``pysph.solver.utils``.  The reference's post-processing
reads dumps back through ``iter_output(files, 'cylinders')`` and then uses
``array.nb``, ``array.xcm`` and ``sd['t']``
(/root/reference/code/stack_of_cylinders.py:463-476,
/root/reference/code/benchmark_1_rigid_body_rotating_and_traslating_freely.py:135-146).
Files are ``.npz`` with one entry per (array, property) plus every constant
(constants are always dumped), and the solver data ``t, dt, count``.
"""
import json
import os
import re

import numpy as np

from .particle_array import ParticleArray


def dump(filename, particles, solver_data, detailed_output=False,
         only_real=True, mpi_comm=None, compress=False):
    out = {}
    meta = {'arrays': [], 'solver_data': {}}
    for k, v in solver_data.items():
        if isinstance(v, (str, list, dict, bool, int)):
            meta['solver_data'][k] = v
        else:
            meta['solver_data'][k] = float(v)
    for pa in particles:
        names = list(pa.properties) if detailed_output or \
            not pa.output_property_arrays else list(pa.output_property_arrays)
        ameta = {'name': pa.name, 'props': {}, 'consts': list(pa.constants),
                 'output': list(pa.output_property_arrays),
                 'n': pa.get_number_of_particles()}
        for n in names:
            arr = getattr(pa, n)
            out['%s/p/%s' % (pa.name, n)] = arr
            ameta['props'][n] = {'stride': pa.stride[n],
                                 'type': pa.property_types[n]}
        for n in pa.constants:
            out['%s/c/%s' % (pa.name, n)] = getattr(pa, n)
        meta['arrays'].append(ameta)
    out['__meta__'] = np.array(json.dumps(meta))
    if not filename.endswith('.npz'):
        filename += '.npz'
    saver = np.savez_compressed if compress else np.savez
    saver(filename, **out)
    return filename


def load(fname):
    data = np.load(fname, allow_pickle=False)
    meta = json.loads(str(data['__meta__']))
    arrays = {}
    for ameta in meta['arrays']:
        pa = ParticleArray(name=ameta['name'])
        pa.__dict__['_n'] = ameta['n']
        pa.__dict__['num_real_particles'] = ameta['n']
        for n, spec in ameta['props'].items():
            pa.add_property(n, type=spec['type'],
                            data=data['%s/p/%s' % (pa.name, n)],
                            stride=spec['stride'])
        for n in ameta['consts']:
            pa.add_constant(n, data['%s/c/%s' % (pa.name, n)])
        pa.set_output_arrays([p for p in ameta['output']
                              if p in pa.properties])
        arrays[pa.name] = pa
    return {'arrays': arrays, 'solver_data': meta['solver_data']}


def _count(fname):
    m = re.search(r'_(\d+)\.npz$', fname)
    return int(m.group(1)) if m else -1


def get_files(dirname=None, fname=None, endswith='.npz'):
    files = [f for f in os.listdir(dirname) if f.endswith(endswith) and
             (fname is None or f.startswith(fname)) and _count(f) >= 0]
    files.sort(key=_count)
    return [os.path.join(dirname, f) for f in files]


def iter_output(files, *arrays):
    for f in files:
        data = load(f)
        sd = data['solver_data']
        if arrays:
            yield (sd,) + tuple(data['arrays'][a] for a in arrays)
        else:
            yield sd, data['arrays']


def save_scene(fname, particles, **extra):
    """Full state (every property and constant) -- used for golden scenes."""
    return dump(fname, particles, extra, detailed_output=True, compress=True)


def load_scene(fname):
    data = load(fname)
    return list(data['arrays'].values()), data['solver_data']

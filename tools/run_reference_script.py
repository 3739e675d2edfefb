"""Run UNMODIFIED Application scripts of the reference on the B200 path and
keep what they produce.

    python tools/run_reference_script.py <dir with the reference's code/> <out.json>

The scripts are not part of this repository: stage /root/reference/code into
an untracked scratch directory (tools/var/refcode, git-ignored) before a GPU
run.  For every script: `python -m rigid_body_2d_3d_pysph_b200.run` semantics
(compat layer installed, the script's directory on sys.path, runpy as
__main__), its own post_process included; then the centre-of-mass series of
stack_of_cylinders (the quantity its post_process plots, :447-509) is read
back through iter_output and compared with the Zhang et al. data points the
script ships (x_com_zhang.csv, y_com_zhang.csv)."""
import json
import os
import runpy
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rigid_body_2d_3d_pysph_b200.compat.install import install  # noqa: E402


def run_script(code_dir, script, argv, out_dir):
    install()
    if code_dir not in sys.path:
        sys.path.insert(0, code_dir)
    path = os.path.join(code_dir, script)
    old = sys.argv
    sys.argv = [path] + argv + ['--directory', out_dir]
    t0 = time.time()
    try:
        ns = runpy.run_path(path, run_name='__main__')
    finally:
        sys.argv = old
    return ns, time.time() - t0


def com_series(out_dir, array, nb_div, length):
    from rigid_body_2d_3d_pysph_b200.compat.output import (get_files,
                                                           iter_output)
    files = get_files(out_dir)
    t, sx, sy = [], [], []
    for sd, arr in iter_output(files, array):
        t.append(sd['t'])
        xc = np.asarray(arr.xcm).reshape(-1, 3)
        sx.append(xc[:, 0].sum() / nb_div / length)
        sy.append(xc[:, 1].sum() / nb_div / length)
    return np.array(t), np.array(sx), np.array(sy), len(files)


def main():
    code_dir = os.path.abspath(sys.argv[1])
    out_json = sys.argv[2]
    scratch = os.path.join(os.path.dirname(os.path.abspath(out_json)),
                           'ref_runs')
    res = {}
    # ---- stack_of_cylinders to tf = 0.7 (wall pulled at t = 0.2) -----------
    d = os.path.join(scratch, 'stack_of_cylinders')
    ns, secs = run_script(code_dir, 'stack_of_cylinders.py',
                          ['--tf', '0.7', '--pfreq', '200'], d)
    t, sx, sy, nfiles = com_series(d, 'cylinders', 33, 0.26)
    tw = t - 0.2
    out = {'seconds': secs, 'output_files': nfiles, 't_end': float(t[-1]),
           'x_com_over_L': [float(v) for v in sx],
           'y_com_over_L': [float(v) for v in sy],
           't_minus_wall_time': [float(v) for v in tw]}
    for ax, series in (('x', sx), ('y', sy)):
        data = np.loadtxt(os.path.join(code_dir, '%s_com_zhang.csv' % ax),
                          delimiter=',')
        sim = np.interp(data[:, 0], tw, series)
        out['%s_zhang' % ax] = [[float(a), float(b), float(c)] for a, b, c in
                                zip(data[:, 0], data[:, 1], sim)]
        out['%s_max_abs_dev_from_zhang' % ax] = float(
            np.abs(sim - data[:, 1]).max())
    res['stack_of_cylinders'] = out
    # ---- benchmark_5_3d --pyramid-cubes to tf = 0.5 -------------------------
    d = os.path.join(scratch, 'benchmark_5_3d')
    ns, secs = run_script(code_dir, 'benchmark_5_steady_cubes_on_a_wall_3d.py',
                          ['--pyramid-cubes', '--tf', '0.5', '--pfreq',
                           '500'], d)
    from rigid_body_2d_3d_pysph_b200.compat.output import (get_files,
                                                           iter_output)
    files = get_files(d)
    first = last = None
    for sd, arr in iter_output(files, 'body'):
        xc = np.asarray(arr.xcm).reshape(-1, 3).copy()
        if first is None:
            first = xc
        last = xc
        t_end = sd['t']
    res['benchmark_5_3d_pyramid'] = {
        'seconds': secs, 'output_files': len(files), 't_end': float(t_end),
        'xcm_first': first.tolist(), 'xcm_last': last.tolist(),
        'max_horizontal_drift': float(np.abs(last[:, [0, 2]] -
                                             first[:, [0, 2]]).max()),
        'settling_dy': (last[:, 1] - first[:, 1]).tolist()}
    with open(out_json, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(dict((k, dict((kk, vv) for kk, vv in v.items()
                                   if not isinstance(vv, list)))
                          for k, v in res.items()), indent=1))


if __name__ == '__main__':
    main()

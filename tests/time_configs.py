"""Step time of the BASELINE.json configs 1-4 on the GPU (CUDA graph) and on
the CPU oracle (not a test; run by hand:  python -m tests.time_configs)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rbo  # noqa: E402
from rigid_body_2d_3d_pysph_b200.device import DeviceScene  # noqa: E402
from tests.util import load_config, oracle_params  # noqa: E402


def main():
    for name, nsteps in [('benchmark_1', 2000), ('benchmark_2', 2000),
                         ('benchmark_5_3d', 2000),
                         ('stack_of_cylinders', 2000)]:
        garr, meta = load_config(name)
        oarr, _ = load_config(name)
        sc = DeviceScene(garr, meta['rigid'], meta['boundaries'],
                         dim=meta['dim'], kr=meta['kr'], kf=meta['kf'],
                         fric_coeff=meta['fric_coeff'], gx=meta['gx'],
                         gy=meta['gy'], gz=meta['gz'],
                         planar=(meta['stepper'] == 'gtvf2d'))
        sc.gtvf_step(meta['dt'], 100, graph=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sc.gtvf_step(meta['dt'], nsteps, graph=True)
        torch.cuda.synchronize()
        tg = (time.perf_counter() - t0) / nsteps
        sc.check_status()
        p = oracle_params(meta)
        planar = meta['stepper'] == 'gtvf2d'
        rbo.gtvf_step(oarr, meta['rigid'], p, planar=planar, nsteps=20)
        t0 = time.perf_counter()
        rbo.gtvf_step(oarr, meta['rigid'], p, planar=planar, nsteps=100)
        tc = (time.perf_counter() - t0) / 100
        n = sum(a.get_number_of_particles() for a in garr)
        print('%-20s particles %6d  GPU %.1f us/step  CPU(%d thr) %.1f us/step'
              % (name, n, tg * 1e6, rbo.num_threads(), tc * 1e6), flush=True)


if __name__ == '__main__':
    main()

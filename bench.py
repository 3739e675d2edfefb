#!/usr/bin/env python
"""Benchmark of the rigid-body hot path on B200 (driver contract).

Workload (BASELINE.json config 5): synthetic 3-D pile of 100 000 lattice
blocks (5x5x4, 10 M particles) dropped on a walled floor, pre-settled, then
K full GTVF steps timed.  A "step" = one pass of the hot path over the whole
scene: kick + drift + re-pose, cell-list build, fused contact, per-body
force/torque reduction, kick, velocities.

  python bench.py --gpus N --steps K --warmup W            (this framework)
  python bench.py --impl reference --gpus N --steps K ...  (CPU path, rank 0)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte
model behind `roofline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'particle-updates/sec'
UNIT = 'particle-updates/s'
DT = 1e-4
KR, KF, MU = 1e5, 1e3, 0.5


def host_cores():
    """Cores this process may run on (the container's share, not the box's)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


_T0 = time.perf_counter()


def note(msg):
    """progress on stderr (stdout carries the one JSON line)"""
    sys.stderr.write('[bench %7.1fs] %s\n' % (time.perf_counter() - _T0, msg))
    sys.stderr.flush()


def claim_stdout():
    """stdout carries ONE JSON line.  Native libraries write to file
    descriptor 1 behind Python's back (NCCL prints its version banner there,
    and everything NCCL_DEBUG asks for): from here on descriptor 1 is stderr,
    and the returned descriptor is the real stdout for emit()."""
    try:
        sys.stdout.flush()
        real = os.dup(1)
        os.dup2(2, 1)
        return real
    except OSError:
        return None


def emit(real, text):
    """the JSON line, to the real stdout"""
    if real is None:
        print(text)
        sys.stdout.flush()
        return
    data = (text + '\n').encode()
    while data:
        data = data[os.write(real, data):]


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.gpu),
                     '--query-gpu=' + self.Q,
                     '--format=csv,noheader,nounits'],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in
                                         out.splitlines()[0].split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for n, v in zip(names, s[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def algorithmic_bytes(n_rigid, n_static_src, n_active, n_bodies):
    """compulsory traffic of one step / one contact
    launch.  189 B per rigid particle (76 pose+velocity, 32 cell build, 81
    contact), 61 B per static source particle (32 cell build + 29 as source),
    104 B per active slot (history read+write), 600 B per body."""
    step = n_rigid * 189. + n_static_src * 61. + n_active * 104. + \
        n_bodies * 600.
    contact = n_rigid * 81. + n_static_src * 29. + n_active * 104.
    return step, contact


def run_reference(args, rank, world):
    """CPU arm: the C restatement of the reference step (oracle/rbo.c, the
    reference itself needs PySPH, which cannot be installed here), all host
    threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import rbo
    from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
    # torchrun exports OMP_NUM_THREADS=1 to every rank: ask for the cores
    threads = rbo.set_num_threads(args.cpu_threads or host_cores())
    nb = args.cpu_bodies
    (body, wall), scheme, info = synthetic_pile(nb, seed=0)
    rbo.add_sparse_history(body, 4)
    p = rbo.make_params(3, DT, KR, KF, MU, 0., -9.81, 0.,
                        eta_uniform=info['eta_uniform'])
    arrays = [body, wall]
    settle = min(args.settle, args.cpu_settle)
    rbo.gtvf_step(arrays, ['body'], p, ks=4, nsteps=settle + args.warmup)
    t0 = time.perf_counter()
    counts = rbo.gtvf_step(arrays, ['body'], p, ks=4, nsteps=args.steps)
    dt = time.perf_counter() - t0
    n = info['n_body_particles']
    value = n * args.steps / dt
    sample = ('%d-body / %d-particle pile (same generator, seed 0), %d settle '
              '+ %d warm-up + %d timed steps, %d OpenMP threads' %
              (nb, n, settle, args.warmup, args.steps, threads))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'synthetic 3D pile, 100k blocks of 5x5x4 / 10M '
                   'particles (BASELINE.json config 5); CPU arm runs the '
                   'bounded sample named in cpu_baseline.sample'},
        'contact_pairs_per_s': float(counts[0]) / dt,
        'cpu_baseline': {'value': value, 'unit': UNIT,
                         'cores': threads, 'kind': 'port',
                         'sample': sample,
                         'note': 'PySPH cannot be installed (no network); '
                         'this is the C/OpenMP restatement of the reference '
                         'step (gcc -O3 -fopenmp -ffp-contract=off), '
                         'validated against the reference\'s own Python '
                         'methods (tests/golden)'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def build_scene(args, rank, world, bodies, halo_cap, dev):
    """The pile through the package's public surface: scene generator,
    RigidBody3DScheme.configure_solver, Solver.setup (which builds the device
    scene and hands it to the integrator)."""
    from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
    arrays, scheme, info = synthetic_pile(
        bodies, seed=0, slab=(rank, world),
        halo_cap=halo_cap if world > 1 else 0)
    scheme.kr, scheme.kf, scheme.fric_coeff = KR, KF, MU
    scheme.configure_solver(dt=DT, tf=1e9, pfreq=10**9, ks=args.ks,
                            list_cap=args.list_cap,
                            eta_uniform=info['eta_uniform'])
    solver = scheme.solver
    solver.setup(arrays, scheme.get_equations())
    return arrays, info, solver


def time_steps(run_steps, nsteps, barrier, torch, dev):
    e0, e1 = torch.cuda.Event(enable_timing=True), \
        torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(nsteps)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--bodies', type=int, default=100000,
                    help='bodies per GPU (100 particles each)')
    ap.add_argument('--settle', type=int, default=2000)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='N > 1: weak = --bodies per GPU (N piles side by '
                    'side), strong = --bodies in total, cut into N slabs')
    ap.add_argument('--cpu-bodies', type=int, default=10000)
    ap.add_argument('--cpu-settle', type=int, default=50)
    ap.add_argument('--cpu-steps', type=int, default=40)
    ap.add_argument('--cpu-threads', type=int, default=0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the settled-pile, moving-wall, strong-scaling '
                    'and DEMScheme measurements')
    ap.add_argument('--settled-steps', type=int, default=4000)
    ap.add_argument('--ks', type=int, default=8)
    ap.add_argument('--list-cap', type=int, default=160,
                    help='neighbour-list entries per particle (the compacting '
                    'pile of the later state needs more than the early one)')
    ap.add_argument('--halo-cap', type=int, default=600000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    out_fd = claim_stdout()
    import torch
    import torch.distributed as dist
    from rigid_body_2d_3d_pysph_b200 import _lib
    from rigid_body_2d_3d_pysph_b200.device import (BodyStateStream,
                                                   BoundaryStream)

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- scene ---------------------------------------------------------------
    # N = 1: the whole pile on one GPU.  N > 1, weak: the scene is N such piles
    # side by side in one walled box; strong: ONE pile of --bodies cut into N
    # x-slabs.  Rank k owns slab k and exchanges source-particle halos with
    # its neighbours every step.
    strong = world > 1 and args.scaling == 'strong'
    per_rank = args.bodies // world if strong else args.bodies
    t_build = time.perf_counter()
    arrays, info, solver = build_scene(args, rank, world, per_rank,
                                       args.halo_cap, dev)
    body, wall = arrays[0], arrays[1]
    sc, integ = solver.scene, solver.integrator
    t_build = time.perf_counter() - t_build
    n_rigid = sc.n_rigid
    n_static_src = info['n_wall_sources']
    slab = None
    if world > 1:
        from rigid_body_2d_3d_pysph_b200.parallel import SlabScene
        slab = SlabScene(sc, rank, world)

    def run_steps(n):
        if slab is not None:
            slab.gtvf_step(DT, n)
        else:
            integ.step(0., DT, n, graph=True)

    note('scene built (%d particles), settling' % n_rigid)
    # ---- pre-settle, warm-up -------------------------------------------------
    run_steps(args.settle)
    run_steps(args.warmup)
    barrier()
    sc.check_status()
    sc.read_counters(reset=True)

    note('timed region')
    # ---- timed region: K steps, device time, max over ranks -----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    halo0 = slab.bytes_recv if slab is not None else 0
    ms = time_steps(run_steps, args.steps, barrier, torch, dev)
    sampler.stop_flag = True
    halo_bytes = (slab.bytes_recv - halo0) / args.steps if slab is not None \
        else 0
    cnt = sc.read_counters(reset=True)
    sc.check_status()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([cnt['gated_pairs'], cnt['active_slots'],
                          cnt['candidates'], n_rigid, cnt['list_entries']],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_pairs, tot_active, tot_cand, tot_rigid, tot_list = (
            float(v) for v in c)
    else:
        tot_pairs, tot_active, tot_cand, tot_rigid, tot_list = (
            float(cnt['gated_pairs']), float(cnt['active_slots']),
            float(cnt['candidates']), float(n_rigid),
            float(cnt['list_entries']))
    sec = ms * 1e-3
    value = tot_rigid * args.steps / sec
    pairs_per_s = tot_pairs / sec

    note('%.3f ms/step; kernel timings' % (ms / args.steps))
    # ---- the dominant kernels alone, CUDA events on their stream ------------
    # `contact_ms`: the pair kernels of one contact evaluation on REUSED lists
    # (k_sparse_reset + k_filter + k_slots: what every step runs);
    # `rebuild_ms`: what a list rebuild adds (cell list, k_neighbours,
    # k_list_sort), paid when a body has moved half the skin.
    import ctypes
    p = sc.params(DT)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_reb, t_con = [], []
    for it_ in range(6):
        sc.force_rebuild()
        ev[0].record()
        sc.cells_build()
        _lib.check(sc.lib.rbx_contact_neighbours(
            ctypes.byref(sc.scene), ctypes.byref(sc._cells), ctypes.byref(p),
            sc.stream))
        ev[1].record()
        _lib.check(sc.lib.rbx_contact_slots(
            ctypes.byref(sc.scene), ctypes.byref(sc._cells), ctypes.byref(p),
            None, sc.stream))
        ev[2].record()
        torch.cuda.synchronize(dev)
        t_reb.append(ev[0].elapsed_time(ev[1]))
        t_con.append(ev[1].elapsed_time(ev[2]))
    rebuild_ms = float(np.mean(t_reb[2:]))
    contact_ms = float(np.mean(t_con[2:]))
    list_entries = float(sc.T['nbr_cnt'].bitwise_and(0x3fffffff).sum().item())
    rebuilds_per_step = tot_list / world / max(list_entries, 1.) / args.steps
    active_per_step = tot_active / args.steps / world
    step_b, contact_b = algorithmic_bytes(n_rigid, n_static_src,
                                          active_per_step, sc.n_bodies)
    peak, peak_src = peaks()
    traffic = None
    limiter = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) as f:
            tr = json.load(f)
        if tr['config']['bodies'] == per_rank:
            traffic = tr['contact_dram_bytes_per_evaluation']
            limiter = '; '.join(
                '%s: %.2f ms, %.0f %% issue active, %.1f of 32 lanes, %d '
                'registers, FP64 pipe %.0f %%' % (
                    n, k['ms'], k['issue_active_pct'], k['threads_per_inst'],
                    k['registers'], k['fp64_pipe_pct'])
                for n, k in tr['kernels'].items())
    except Exception:
        pass
    roof = {'bound': 'hbm',
            'kernel': 'contact evaluation on reused neighbour lists '
            '(rbx_contact_slots: k_sparse_reset, k_filter = FP32 first pass '
            'over every list entry (FFMA2 / FMUL2, two entries per '
            'instruction), k_slots = exact FP64 pass over what it could not '
            'exclude)',
            'achieved': contact_b / (contact_ms * 1e-3) / 1e9, 'peak': peak,
            'unit': 'GB/s', 'peak_source': peak_src, 'traffic': traffic,
            'traffic_source': 'profiles/r02_traffic.json (ncu dram__bytes '
            'read+write of k_filter + k_slots, one evaluation)'
            if traffic else None,
            'secondary_limiter': ('instruction issue and the latency of the '
                                  '16-byte position gathers (packed-FP32 '
                                  'pair math over %.3g list entries per '
                                  'step), not HBM; '
                                  'ncu, profiles/r02_traffic.json: ' %
                                  list_entries + limiter)
            if limiter else None,
            'ms_per_launch': contact_ms,
            'list_rebuild_ms': rebuild_ms,
            'list_rebuilds_per_step': rebuilds_per_step,
            'ms_per_evaluation_average': contact_ms +
            rebuilds_per_step * rebuild_ms,
            'algorithmic_bytes_per_launch': contact_b}
    roof['frac'] = roof['achieved'] / peak
    step_roof = {'bound': 'hbm', 'achieved': step_b /
                 (sec / args.steps) / 1e9, 'peak': peak,
                 'unit': 'GB/s', 'algorithmic_bytes_per_step': step_b}
    step_roof['frac'] = step_roof['achieved'] / peak

    note('e2e')
    # ---- e2e through the public API: host-driven boundary in, body state out,
    #      every step ----------------------------------------------------------
    # BoundaryStream.submit / apply = pinned host -> staging -> scene (the
    # lists are rebuilt only if the wall really moved more than half the
    # skin), Integrator.step = one GTVF step, BodyStateStream.snapshot / wait
    # = per-body results -> pinned host.  Every step consumes its own upload
    # and every step's result is on the host inside the timed region.
    wall_n = wall.get_number_of_particles()
    names_in = ('x', 'y', 'z', 'u', 'v', 'w')
    feeder = BoundaryStream(sc, 'wall', names_in)
    reader = BodyStateStream(sc)
    base = dict((n, torch.from_numpy(
        np.ascontiguousarray(wall.properties[n]))) for n in names_in)
    host_in = [dict((n, base[n].clone().pin_memory()) for n in names_in)
               for _ in range(2)]
    h2d, d2h = feeder.bytes_per_submit, reader.bytes_per_snapshot
    e2e_steps = max(3, min(args.steps, 50))
    amp, freq = 0.2 * info['dx'], 50.          # the moving wall: a shaker

    def wall_state(k, moving):
        h = host_in[k & 1]
        if moving:
            feeder.ev_up[k & 1].synchronize()   # its last upload has left
            ph = 2. * np.pi * freq * DT * k
            h['x'].copy_(base['x'] + amp * np.sin(ph))
            h['u'].fill_(amp * 2. * np.pi * freq * np.cos(ph))
        return h

    def run_e2e(nsteps, moving):
        feeder.submit(wall_state(0, moving))
        last = -1
        for k in range(nsteps):
            if k + 1 < nsteps:
                feeder.submit(wall_state(k + 1, moving))
            feeder.apply()
            if slab is not None:
                slab.gtvf_step(DT, 1)
            else:
                integ.step(k * DT, DT, 1)
            j = reader.snapshot()
            if last >= 0:
                reader.wait(last)        # result of step k - 1 is on the host
            last = j
        reader.wait(last)
        torch.cuda.current_stream(dev).synchronize()

    def time_e2e(moving):
        run_e2e(3, moving)
        barrier()
        t0 = time.perf_counter()
        run_e2e(e2e_steps, moving)
        barrier()
        dt_ = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_ = float(t.item())
        return tot_rigid * e2e_steps / dt_

    e2e = {'value': time_e2e(False), 'unit': UNIT,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'steps': e2e_steps,
           'what': 'per step, through the public API: wall x,y,z,u,v,w from '
           'pinned host memory (BoundaryStream.submit/apply -> '
           'DeviceScene.static_update), one GTVF step (Integrator.step; '
           'slabs: SlabScene.gtvf_step), per-body xcm,vcm,omega,R,force,'
           'torque to pinned host memory (BodyStateStream); copies run on '
           'a second stream one step ahead / behind; wall clock over all '
           'steps with every result on the host.  The wall data are the '
           'same every step (a static boundary, uploaded all the same); '
           'moving_wall has one that moves'}
    sc.check_status()
    extras = {}
    if not args.no_extras:
        note('e2e with a moving wall')
        e2e['moving_wall'] = {
            'value': time_e2e(True), 'unit': UNIT,
            'what': 'the same loop with the whole wall shaken along x '
            '(%.3g m amplitude, %.0f Hz): the lists are rebuilt whenever '
            'it has moved half the skin' % (amp, freq)}
        run_e2e(1, False)                  # put the wall back
        sc.check_status()

    note('e2e done; extras')
    # ---- a settled pile (N = 1): the same scene much later --------------------
    if world == 1 and not args.no_extras and args.settled_steps > 0:
        run_steps(args.settled_steps)
        sc.read_counters(reset=True)
        ms2 = time_steps(run_steps, args.steps, barrier, torch, dev)
        c2 = sc.read_counters(reset=True)
        sc.check_status()
        extras['settled_pile'] = {
            'what': 'the same scene after %d more steps (t = %.2f s): the '
            'pile is compacting, ten times as many slots are in contact and '
            'the exact FP64 pass, not the FP32 first pass, is the cost' % (
                args.settled_steps, DT * (args.settle + args.settled_steps)),
            'ms_per_step': ms2 / args.steps,
            'value': n_rigid * args.steps / (ms2 * 1e-3), 'unit': UNIT,
            'contact_pairs_per_step': c2['gated_pairs'] / args.steps,
            'active_slots_per_step': c2['active_slots'] / args.steps,
            'list_rebuilds_per_step': c2['list_entries'] / max(
                float(sc.T['nbr_cnt'].bitwise_and(0x3fffffff).sum().item()),
                1.) / args.steps}

    # ---- strong scaling (N > 1): ONE pile of --bodies cut into N slabs -------
    if world > 1 and not strong and not args.no_extras:
        del feeder, reader, slab, sc, integ, solver
        torch.cuda.empty_cache()
        from rigid_body_2d_3d_pysph_b200.parallel import SlabScene
        arr2, info2, solver2 = build_scene(args, rank, world,
                                           args.bodies // world,
                                           args.halo_cap, dev)
        slab2 = SlabScene(solver2.scene, rank, world)
        slab2.gtvf_step(DT, args.settle + args.warmup)
        barrier()
        ms3 = time_steps(lambda n: slab2.gtvf_step(DT, n), args.steps,
                         barrier, torch, dev)
        t = torch.tensor([ms3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        solver2.scene.check_status()
        n_tot = float(info2['n_body_particles']) * world
        extras['strong_scaling'] = {
            'what': 'ONE pile of %d bodies (%d particles) cut into %d '
            'x-slabs (bench.py --scaling strong runs only this)' % (
                (args.bodies // world) * world, int(n_tot), world),
            'ms_per_step': float(t.item()) / args.steps,
            'value': n_tot * args.steps / (float(t.item()) * 1e-3),
            'unit': UNIT}

    note('dem')
    # ---- DEMScheme path (N = 1): spheres on a vibrating floor ----------------
    if world == 1 and not args.no_extras:
        try:
            extras['dem_scheme'] = bench_dem(torch, dev)
        except Exception as exc:      # an extra: never takes the line down
            extras['dem_scheme'] = {'error': repr(exc)}

    note('cpu baseline')
    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import rbo
        from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
        threads = rbo.set_num_threads(args.cpu_threads or host_cores())
        (cb, cw), _, cinfo = synthetic_pile(args.cpu_bodies, seed=0)
        rbo.add_sparse_history(cb, 4)
        cp = rbo.make_params(3, DT, KR, KF, MU, 0., -9.81, 0.,
                             eta_uniform=cinfo['eta_uniform'])
        rbo.gtvf_step([cb, cw], ['body'], cp, ks=4,
                      nsteps=min(args.settle, args.cpu_settle))
        t0 = time.perf_counter()
        rbo.gtvf_step([cb, cw], ['body'], cp, ks=4, nsteps=args.cpu_steps)
        cdt = time.perf_counter() - t0
        cpu = {'value': cinfo['n_body_particles'] * args.cpu_steps / cdt,
               'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': '%d-body / %d-particle pile, %d settle + %d timed '
               'steps, C/OpenMP restatement (oracle/rbo.c, gcc -O3)' % (
                   args.cpu_bodies, cinfo['n_body_particles'],
                   min(args.settle, args.cpu_settle), args.cpu_steps)}

    if rank == 0:
        # every step: k_bodies, k_pose, k_sparse_reset, k_filter, k_slots x 2
        # (one of them returns at once), k_reduce, k_bodies = 8; a list
        # rebuild adds 8 cell-list kernels, k_neighbours, k_pos32, k_list_sort,
        # k_list_commit, k_static_commit, k_list_clear = 14.  One GPU (CUDA
        # graph): the rebuild is the body of a conditional node, launched
        # only on the steps that rebuild, + k_set_conditional every step;
        # N > 1 (eager): all 22 every step, the rebuild kernels returning at
        # once on the steps that reuse the lists.  Memsets not counted.
        if world == 1:
            launches = args.steps * (9 + 14 * rebuilds_per_step)
        else:
            launches = args.steps * 22
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {
                'workload': 'synthetic 3D pile, %d blocks of 5x5x4 / %d '
                'particles per GPU + %d wall particles (BASELINE.json '
                'config 5), pre-settled %d steps, dt=1e-4' % (
                    per_rank, n_rigid, wall_n, args.settle),
                'bodies_per_gpu': per_rank,
                'particles_per_gpu': n_rigid,
                'parallelism': ('%d x-slabs of one scene, bodies owned per '
                                'slab, source-particle halo exchange (NCCL '
                                'p2p) every step' % world) if world > 1
                else 'single GPU',
                'halo_bytes_per_rank_per_step': halo_bytes,
                'l2': 'working set (%.1f GB) far larger than the 126 MB L2; '
                'no flush needed' % (step_b / 1e9),
                'stepper': 'GTVF', 'ks': args.ks,
                'skin_factor': solver.skin_factor if world == 1 or strong
                else 0.075,
                # one graph per step on one GPU; slabs replay the two halves
                # of the step around the halo exchange as graphs
                'graph': True if world == 1 else (
                    'two graphs per step, the halo exchange between them'
                    if getattr(slab, 'use_graphs', False) else False)},
            'contact_pairs_per_s': pairs_per_s,
            'contact_pairs_per_step': tot_pairs / args.steps,
            'active_slots_per_step': tot_active / args.steps,
            'candidate_tests_per_step': tot_cand / args.steps,
            'roofline': roof, 'step_roofline': step_roof,
            'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': int(round(launches)),
            'clocks': sampler.summary(),
            'scene_build_s': t_build,
        }
        line.update(extras)
        emit(out_fd, json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_dem(torch, dev, n_side=64, steps=200):
    """DEMScheme (dem.py: LVCDisplacement + DEMStep): a
    packed block of n_side^3 spheres, slightly overlapping, on a floor of
    spheres.  No script of the reference instantiates the scheme; the
    constants it needs are those of the parity fixture (tests/golden/dem2d)."""
    from rigid_body_2d_3d_pysph_b200.dem import DEMScheme
    from rigid_body_2d_3d_pysph_b200.compat.particle_array import \
        get_particle_array
    rad = 0.005
    i, j, k = np.meshgrid(np.arange(n_side), np.arange(n_side),
                          np.arange(n_side), indexing='ij')
    rng = np.random.default_rng(1)
    n = i.size
    x = i.ravel() * 1.98 * rad + rng.uniform(-0.02, 0.02, n) * rad
    y = j.ravel() * 1.97 * rad + 0.99 * rad + rng.uniform(-0.02, 0.02, n) * rad
    z = k.ravel() * 1.98 * rad + rng.uniform(-0.02, 0.02, n) * rad
    rho = 2500.
    m = rho * 4. / 3. * np.pi * rad**3
    sand = get_particle_array(name='sand', x=x, y=y, z=z, h=1.2 * rad, m=m,
                              rho=rho, rad_s=rad,
                              u=rng.uniform(-0.05, 0.05, n),
                              v=rng.uniform(-0.05, 0.05, n),
                              w=rng.uniform(-0.05, 0.05, n))
    sand.add_property('dem_id', type='int', data=0)
    sand.add_property('moi', data=0.4 * m * rad**2)
    fi, fk = np.meshgrid(np.arange(-2, n_side + 2), np.arange(-2, n_side + 2),
                         indexing='ij')
    floor = get_particle_array(name='floor', x=fi.ravel() * 2. * rad,
                               y=np.zeros(fi.size) - rad,
                               z=fk.ravel() * 2. * rad, h=1.2 * rad, m=m,
                               rho=rho, rad_s=rad)
    floor.add_property('dem_id', type='int', data=1)
    for pa in (sand, floor):
        for p in ('wx', 'wy', 'wz'):
            pa.add_property(p)
    sand.add_constant('max_tng_contacts_limit', 16)
    sand.add_constant('kn', [1e5, 2e5])
    sand.add_constant('kt', [2. / 7. * 1e5, 2. / 7. * 2e5])
    sand.add_constant('alpha', [40., 60.])
    sand.add_constant('mu', [0.5, 0.3])
    s = DEMScheme(['sand'], ['floor'], dim=3, gy=-9.81)
    dt = 2e-6
    s.configure_solver(dt=dt, tf=1e9, pfreq=10**9)
    s.setup_properties([sand, floor])
    solver = s.get_solver()
    solver.setup([sand, floor], s.get_equations(), kernel=solver.kernel)
    integ = solver.integrator
    integ.step(0., dt, 20)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), \
        torch.cuda.Event(enable_timing=True)
    e0.record()
    integ.step(0., dt, steps)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    solver.scene.check_status()
    solver.scene.sync_to_host()
    return {'what': 'DEMScheme (LVCDisplacement contact with per-pair '
            'tangential history + DEMStep): %d spheres in a packed block on '
            'a floor of %d, dt = %g' % (n, floor.get_number_of_particles(),
                                        dt),
            'ms_per_step': ms, 'value': n / (ms * 1e-3),
            'unit': 'particle-updates/s',
            'tangential_contacts': int(sand.total_tng_contacts.sum())}


if __name__ == '__main__':
    main()

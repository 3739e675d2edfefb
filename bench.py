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
              '+ %d warm-up + %d timed steps' %
              (nb, n, settle, args.warmup, args.steps))
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
                         'cores': rbo.num_threads(), 'kind': 'port',
                         'sample': sample,
                         'note': 'PySPH cannot be installed (no network); '
                         'this is the C/OpenMP restatement of the reference '
                         'step, validated against the reference\'s own '
                         'Python methods (tests/golden)'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--bodies', type=int, default=100000,
                    help='bodies per GPU (100 particles each)')
    ap.add_argument('--settle', type=int, default=2000)
    ap.add_argument('--cpu-bodies', type=int, default=1000)
    ap.add_argument('--cpu-settle', type=int, default=50)
    ap.add_argument('--cpu-steps', type=int, default=60)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--ks', type=int, default=8)
    ap.add_argument('--halo-cap', type=int, default=600000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from rigid_body_2d_3d_pysph_b200 import _lib
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- scene (one pile per rank: weak scaling) ---------------------------
    t_build = time.perf_counter()
    # N = 1: the whole pile on one GPU.  N > 1: weak scaling, the scene is N
    # such piles side by side in one walled box, rank k owns x-slab k and
    # exchanges source-particle halos with its neighbours every step.
    arrays, scheme, info = synthetic_pile(
        args.bodies, seed=0, slab=(rank, world),
        halo_cap=args.halo_cap if world > 1 else 0)
    body, wall = arrays[0], arrays[1]
    sc = DeviceScene(arrays, ['body'], [a.name for a in arrays[1:]], dim=3,
                     kr=KR, kf=KF, fric_coeff=MU, gy=-9.81, ks=args.ks,
                     eta_uniform=info['eta_uniform'], device=dev)
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
            sc.gtvf_step(DT, n, graph=True)

    # ---- pre-settle, warm-up -------------------------------------------------
    run_steps(args.settle)
    run_steps(args.warmup)
    barrier()
    sc.check_status()
    sc.read_counters(reset=True)

    # ---- timed region: K steps, device time, max over ranks -----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), \
        torch.cuda.Event(enable_timing=True)
    barrier()
    halo0 = slab.bytes_recv if slab is not None else 0
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
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
                          cnt['candidates'], n_rigid], dtype=torch.float64,
                         device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_pairs, tot_active, tot_cand, tot_rigid = (float(v) for v in c)
    else:
        tot_pairs, tot_active, tot_cand, tot_rigid = (
            float(cnt['gated_pairs']), float(cnt['active_slots']),
            float(cnt['candidates']), float(n_rigid))
    sec = ms * 1e-3
    value = tot_rigid * args.steps / sec
    pairs_per_s = tot_pairs / sec

    # ---- dominant kernel alone (contact), CUDA events on its stream ---------
    # `contact_ms`: average over evaluations as they occur in the run (the
    # neighbour lists are reused until a body has moved half the skin, so
    # most evaluations are k_slots alone); `contact_rebuild_ms`: an
    # evaluation that rebuilds the lists (k_neighbours + k_slots).
    p = sc.params(DT)
    kms, kms_rebuild = [], []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it_ in range(14):
        forced = it_ >= 10
        sc.gtvf_kick(DT)
        sc.gtvf_drift(DT)
        sc.pose(_lib.POSE_POS | _lib.POSE_VEL | _lib.POSE_VEL_PREV |
                _lib.POSE_NORMALS)
        if forced:
            sc.force_rebuild()
        if slab is not None:
            slab.exchange_halo(full=slab.lists_need_rebuild())
        sc.cells_build()
        ev[0].record()
        sc.contact(DT)
        ev[1].record()
        sc.reduce_bodies()
        sc.gtvf_kick(DT)
        sc.pose(_lib.POSE_VEL)
        torch.cuda.synchronize(dev)
        (kms_rebuild if forced else kms).append(ev[0].elapsed_time(ev[1]))
    if not kms:
        kms = list(kms_rebuild)
    contact_rebuild_ms = float(np.mean(kms_rebuild[1:]))
    contact_ms = float(np.mean(kms[1:]))
    active_per_step = tot_active / args.steps / world
    step_b, contact_b = algorithmic_bytes(n_rigid, n_static_src,
                                          active_per_step, sc.n_bodies)
    peak, peak_src = peaks()
    traffic = None
    limiter = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')) as f:
            tr = json.load(f)
        if tr['config']['bodies'] == args.bodies:
            # DRAM bytes of an average evaluation: k_slots every time, the
            # rebuild kernels on the share of evaluations that rebuild
            share = 0. if contact_rebuild_ms <= contact_ms else \
                (contact_ms - min(kms[1:])) / max(
                    contact_rebuild_ms - min(kms[1:]), 1e-9)
            traffic = tr['contact_dram_bytes_per_evaluation'] + \
                share * tr['contact_dram_bytes_per_rebuild']
            limiter = '; '.join(
                '%s: %.0f %% FP64 pipe, %.0f %% issue active, %.1f of 32 '
                'lanes, %d registers' % (
                    n.split(' ')[0], k['fp64_pipe_pct'],
                    k['issue_active_pct'], k['threads_per_inst'],
                    k['registers'])
                for n, k in tr['kernels'].items())
    except Exception:
        pass
    roof = {'bound': 'hbm',
            'kernel': 'contact evaluation (rbx_contact_mofidi): k_slots, plus '
            'k_neighbours + k_list_sort on the steps that rebuild the '
            'neighbour lists',
            'achieved': contact_b / (contact_ms * 1e-3) / 1e9, 'peak': peak,
            'unit': 'GB/s', 'peak_source': peak_src, 'traffic': traffic,
            'traffic_source': 'profiles/r01_traffic.json (ncu dram__bytes '
            'read+write per launch, rebuild kernels weighted by the share of '
            'evaluations that rebuild)' if traffic else None,
            'secondary_limiter': ('FP64 issue and latency, not HBM (ncu, '
                                  'profiles/r01_traffic.json): ' + limiter)
            if limiter else None,
            'ms_per_launch': contact_ms,
            'ms_per_launch_with_list_rebuild': contact_rebuild_ms,
            'algorithmic_bytes_per_launch': contact_b}
    roof['frac'] = roof['achieved'] / peak
    step_roof = {'bound': 'hbm', 'achieved': step_b * world /
                 (sec / args.steps) / 1e9 / world, 'peak': peak,
                 'unit': 'GB/s', 'algorithmic_bytes_per_step': step_b}
    step_roof['frac'] = step_roof['achieved'] / peak

    # ---- e2e through the host-facing API: host-driven boundary in, body
    #      state out, every step ----------------------------------------------
    wall_o = sc.p_off['wall']
    wall_n = wall.get_number_of_particles()
    names_in = ['x', 'y', 'z', 'u', 'v', 'w']
    host_in = dict((n, torch.from_numpy(
        np.ascontiguousarray(wall.properties[n])).pin_memory())
        for n in names_in)
    names_out = ['xcm', 'vcm', 'omega', 'R', 'force', 'torque']
    host_out = dict((n, torch.empty_like(sc.B[n], device='cpu').pin_memory())
                    for n in names_out)
    h2d = sum(t.numel() * 8 for t in host_in.values())
    d2h = sum(t.numel() * 8 for t in host_out.values())
    e2e_steps = max(3, min(args.steps, 50))

    # Double-buffered, as a host application that drives the boundary would
    # do it: a copy stream uploads the wall state of step k + 1 into a device
    # staging buffer and downloads the body state of step k - 1 from a device
    # snapshot while step k computes.  Every step still consumes its own
    # upload and every step's result reaches pinned host memory inside the
    # timed region.
    cur = torch.cuda.current_stream(dev)
    copy_s = torch.cuda.Stream(dev)
    stage = [dict((n, torch.empty(wall_n, dtype=torch.float64, device=dev))
                  for n in names_in) for _ in range(2)]
    snap = [dict((n, torch.empty_like(sc.B[n])) for n in names_out)
            for _ in range(2)]
    host_res = [host_out, dict((n, torch.empty_like(host_out[n]).pin_memory())
                               for n in names_out)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    ev_snap = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    for e in ev_used + ev_out:
        e.record(cur)

    def upload(k):
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_used[k & 1])     # staging buffer is free
            for n in names_in:
                stage[k & 1][n].copy_(host_in[n], non_blocking=True)
            ev_up[k & 1].record(copy_s)

    def step(k):
        cur.wait_event(ev_up[k & 1])
        for n in names_in:
            sc.P[n][wall_o:wall_o + wall_n].copy_(stage[k & 1][n])
        ev_used[k & 1].record(cur)
        if slab is not None:
            slab.gtvf_step(DT, 1)
        else:
            sc._gtvf_step_call(p)
        cur.wait_event(ev_out[k & 1])             # snapshot buffer is free
        for n in names_out:
            snap[k & 1][n].copy_(sc.B[n])
        ev_snap[k & 1].record(cur)
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_snap[k & 1])
            for n in names_out:
                host_res[k & 1][n].copy_(snap[k & 1][n], non_blocking=True)
            ev_out[k & 1].record(copy_s)

    def run_e2e(nsteps):
        upload(0)
        for k in range(nsteps):
            if k + 1 < nsteps:
                upload(k + 1)
            step(k)
            if k > 0:
                ev_out[(k - 1) & 1].synchronize()   # result of step k - 1 is on the host
        ev_out[(nsteps - 1) & 1].synchronize()
        cur.synchronize()

    run_e2e(3)
    barrier()
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    barrier()
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e = {'value': tot_rigid * e2e_steps / e2e_sec, 'unit': UNIT,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'steps': e2e_steps,
           'what': 'per step: wall x,y,z,u,v,w pinned host->device '
           '(host-driven boundary, as post_step moves it), one GTVF step '
           'through rbx_gtvf_step, per-body xcm,vcm,omega,R,force,torque '
           'device->pinned host; copies double-buffered on a second stream '
           '(upload of step k+1 and download of step k-1 overlap step k), '
           'wall clock over all steps with every result on the host'}
    sc.check_status()

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import rbo
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
               'unit': UNIT, 'cores': rbo.num_threads(), 'kind': 'port',
               'sample': '%d-body / %d-particle pile, %d settle + %d timed '
               'steps, C/OpenMP restatement (oracle/rbo.c)' % (
                   args.cpu_bodies, cinfo['n_body_particles'],
                   min(args.settle, args.cpu_settle), args.cpu_steps)}

    if rank == 0:
        launches_per_step = 18   # bodies, pose, 8 cell-list kernels,
        #                          k_neighbours, k_list_sort, list commit + clear,
        #                          k_slots, bodies (reduce), bodies (kick), pose (memsets not
        #                          counted; the cell-list kernels and
        #                          k_neighbours return at once on steps
        #                          that reuse the neighbour lists)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {
                'workload': 'synthetic 3D pile, %d blocks of 5x5x4 / %d '
                'particles per GPU + %d wall particles (BASELINE.json '
                'config 5), pre-settled %d steps, dt=1e-4' % (
                    args.bodies, n_rigid, wall_n, args.settle),
                'bodies_per_gpu': args.bodies,
                'particles_per_gpu': n_rigid,
                'parallelism': ('%d x-slabs of one scene, bodies owned per '
                                'slab, source-particle halo exchange (NCCL '
                                'p2p) every step' % world) if world > 1
                else 'single GPU',
                'halo_bytes_per_rank_per_step': halo_bytes,
                'l2': 'working set (%.1f GB) far larger than the 126 MB L2; '
                'no flush needed' % (step_b / 1e9),
                'stepper': 'GTVF', 'ks': args.ks, 'graph': world == 1},
            'contact_pairs_per_s': pairs_per_s,
            'contact_pairs_per_step': tot_pairs / args.steps,
            'active_slots_per_step': tot_active / args.steps,
            'candidate_tests_per_step': tot_cand / args.steps,
            'roofline': roof, 'step_roofline': step_roof,
            'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': launches_per_step * args.steps,
            'clocks': sampler.summary(),
            'scene_build_s': t_build,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Where a multi-GPU step spends its time (CUDA events around the phases of
SlabScene.gtvf_step):

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/slab_phases.py [settle]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_2d_3d_pysph_b200.device import DeviceScene  # noqa: E402
from rigid_body_2d_3d_pysph_b200.parallel import SlabScene  # noqa: E402
from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile  # noqa: E402

rank = int(os.environ['RANK'])
world = int(os.environ['WORLD_SIZE'])
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
arrays, scheme, info = synthetic_pile(100000, seed=0, slab=(rank, world),
                                      halo_cap=600000)
sc = DeviceScene(arrays, ['body'], [a.name for a in arrays[1:]], dim=3,
                 kr=1e5, kf=1e3, fric_coeff=0.5, gy=-9.81,
                 eta_uniform=info['eta_uniform'], device=dev, list_cap=160)
slab = SlabScene(sc, rank, world)
dt = 1e-4
slab.gtvf_step(dt, int(sys.argv[1]) if len(sys.argv) > 1 else 2000)
torch.cuda.synchronize()
dist.barrier()
names = ['half A (kick, drift, pose)', 'flag all-reduce + D2H', 'halo refresh',
         'full exchange (if any)', 'half B (cells .. kick)']
N = 80
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)]
      for _ in range(N)]
host = np.zeros((N, 5))
fulls = 0
cur = torch.cuda.current_stream(dev)
t_all = time.perf_counter()
for k in range(N):
    e = ev[k]
    t0 = time.perf_counter()
    p = sc.params(dt)
    e[0].record()
    slab._half(p, 2)
    e[1].record()
    t1 = time.perf_counter()
    slab._all_reduce_max(sc.rebuild)
    slab._flag_host.copy_(sc.rebuild, non_blocking=True)
    slab._flag_event.record(cur)
    e[2].record()
    t2 = time.perf_counter()
    slab._refresh_halo()
    e[3].record()
    t3 = time.perf_counter()
    slab._flag_event.synchronize()
    if int(slab._flag_host[0]) != 0:
        fulls += 1
        slab.exchange_halo(full=True)
    e[4].record()
    t4 = time.perf_counter()
    slab._half(p, 4 | 1)
    e[5].record()
    t5 = time.perf_counter()
    host[k] = [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
torch.cuda.synchronize()
wall = (time.perf_counter() - t_all) / N * 1e3
gpu = np.array([[ev[k][i].elapsed_time(ev[k][i + 1]) for i in range(5)]
                for k in range(N)])
gap = np.array([ev[k][5].elapsed_time(ev[k + 1][0]) for k in range(N - 1)])
for r in range(world):
    dist.barrier()
    if r == rank:
        print('rank %d: wall %.3f ms/step, full exchanges %d/%d, halo %d '
              'particles' % (rank, wall, fulls, N, slab.n_halo))
        for i, n in enumerate(names):
            print('   %-28s gpu %.3f ms   host %.3f ms' % (
                n, gpu[5:, i].mean(), host[5:, i].mean() * 1e3))
        print('   gap between steps            gpu %.3f ms' % gap[5:].mean(),
              flush=True)
dist.destroy_process_group()

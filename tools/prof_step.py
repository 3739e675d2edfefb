"""One list-rebuild step and one reuse step of the bench workload, eagerly
launched between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).

    python tools/prof_step.py [n_bodies] [settle] [skin_factor]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_2d_3d_pysph_b200.device import DeviceScene  # noqa: E402
from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
skin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
(body, wall), scheme, info = synthetic_pile(nb)
sc = DeviceScene([body, wall], ['body'], ['wall'], dim=3, gy=-9.81,
                 eta_uniform=info['eta_uniform'], skin_factor=skin,
                 list_cap=int(os.environ.get('RBX_LIST_CAP', '96')))
sc.gtvf_step(1e-4, settle, graph=True)
torch.cuda.synchronize()
sc.check_status()
torch.cuda.profiler.start()
sc.force_rebuild()
sc.gtvf_step(1e-4, 1)
sc.gtvf_step(1e-4, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
sc.check_status()
print('ok', sc.read_counters())

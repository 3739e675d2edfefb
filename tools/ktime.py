"""time k_neighbours / k_slots separately on a settled pile"""
import sys, ctypes, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from rigid_body_2d_3d_pysph_b200 import _lib
from rigid_body_2d_3d_pysph_b200.device import DeviceScene
from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
nb=int(sys.argv[1]); settle=int(sys.argv[2]); skin=float(sys.argv[4]) if len(sys.argv)>4 else 0.1
(body, wall), scheme, info = synthetic_pile(nb)
import os
sc = DeviceScene([body, wall], ['body'], ['wall'], dim=3, gy=-9.81, eta_uniform=info['eta_uniform'], skin_factor=skin, exact=bool(int(os.environ.get('RBX_EXACT', '0'))), list_cap=int(os.environ.get('RBX_LIST_CAP', '96')))
sc.gtvf_step(1e-4, settle, graph=True)
torch.cuda.synchronize()
p = sc.params(1e-4)
ev=[torch.cuda.Event(enable_timing=True) for _ in range(4)]
t0=[];t1=[];t2=[];t3=[]
for i in range(8):
    sc.force_rebuild()
    ev[0].record()
    sc.cells_build()
    ev[1].record()
    _lib.check(sc.lib.rbx_contact_neighbours(ctypes.byref(sc.scene), ctypes.byref(sc._cells), ctypes.byref(p), sc.stream))
    ev[2].record()
    _lib.check(sc.lib.rbx_contact_slots(ctypes.byref(sc.scene), ctypes.byref(sc._cells), ctypes.byref(p), None, sc.stream))
    ev[3].record()
    torch.cuda.synchronize()
    t0.append(ev[0].elapsed_time(ev[1])); t1.append(ev[1].elapsed_time(ev[2])); t2.append(ev[2].elapsed_time(ev[3]))
    # no rebuild: K1 skipped
    ev[0].record()
    sc.cells_build()
    _lib.check(sc.lib.rbx_contact_neighbours(ctypes.byref(sc.scene), ctypes.byref(sc._cells), ctypes.byref(p), sc.stream))
    ev[1].record()
    torch.cuda.synchronize()
    t3.append(ev[0].elapsed_time(ev[1]))
cnt = (sc.T['nbr_cnt'] & 0x3fffffff).float()
print('exact=%d survivors %d of %d, active %s' % (sc.exact, int(sc.counters[6].item()), sc.n_rigid, sc.read_counters()))
print('%s skin=%.2f cells %.3f ms  K1 %.3f ms  K2 %.3f ms  skipped(cells+K1) %.3f ms  list mean %.1f max %d' % (sys.argv[3] if len(sys.argv)>3 else '', skin, np.mean(t0[2:]), np.mean(t1[2:]), np.mean(t2[2:]), np.mean(t3[2:]), cnt.mean().item(), int(cnt.max().item())))

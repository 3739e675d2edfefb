"""time k_neighbours / k_slots separately on a settled pile"""
import sys, ctypes, torch, numpy as np
sys.path.insert(0,'/root/repo')
from rigid_body_2d_3d_pysph_b200 import _lib
from rigid_body_2d_3d_pysph_b200.device import DeviceScene
from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
nb=int(sys.argv[1]); settle=int(sys.argv[2])
(body, wall), scheme, info = synthetic_pile(nb)
sc = DeviceScene([body, wall], ['body'], ['wall'], dim=3, gy=-9.81, eta_uniform=info['eta_uniform'])
sc.gtvf_step(1e-4, settle, graph=True)
torch.cuda.synchronize()
p = sc.params(1e-4)
ev=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
t1=[];t2=[]
for i in range(8):
    sc.cells_build()
    ev[0].record()
    _lib.check(sc.lib.rbx_contact_neighbours(ctypes.byref(sc.scene), ctypes.byref(sc._cells), ctypes.byref(p), sc.stream))
    ev[1].record()
    _lib.check(sc.lib.rbx_contact_slots(ctypes.byref(sc.scene), ctypes.byref(sc._cells), ctypes.byref(p), None, sc.stream))
    ev[2].record()
    torch.cuda.synchronize()
    t1.append(ev[0].elapsed_time(ev[1])); t2.append(ev[1].elapsed_time(ev[2]))
print('%s  K1 %.3f ms  K2 %.3f ms' % (sys.argv[3] if len(sys.argv)>3 else '', np.mean(t1[2:]), np.mean(t2[2:])))

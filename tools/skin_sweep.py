"""steady-state step time of the 10M pile as a function of the skin factor"""
import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from rigid_body_2d_3d_pysph_b200.device import DeviceScene
from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
nb=int(sys.argv[1]); settle=int(sys.argv[2])
for skin in [float(s) for s in sys.argv[3:]]:
    (body, wall), scheme, info = synthetic_pile(nb)
    sc = DeviceScene([body, wall], ['body'], ['wall'], dim=3, gy=-9.81, eta_uniform=info['eta_uniform'], skin_factor=skin, list_cap=128)
    sc.gtvf_step(1e-4, settle, graph=True)
    sc.gtvf_step(1e-4, 20, graph=True)
    torch.cuda.synchronize(); sc.read_counters(reset=True)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); sc.gtvf_step(1e-4, 200, graph=True); e1.record(); torch.cuda.synchronize()
    c=sc.read_counters(); sc.check_status()
    print('skin %.3f  %.3f ms/step  list entries/step %.3g  candidates/step %.3g pairs/step %.4g' % (skin, e0.elapsed_time(e1)/200, c['list_entries']/200, c['candidates']/200, c['gated_pairs']/200), flush=True)
    del sc; torch.cuda.empty_cache()

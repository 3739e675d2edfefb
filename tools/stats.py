import sys, torch, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from rigid_body_2d_3d_pysph_b200.device import DeviceScene
from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
nb=int(sys.argv[1]) if len(sys.argv)>1 else 100000
settle=int(sys.argv[2]) if len(sys.argv)>2 else 2000
(body, wall), scheme, info = synthetic_pile(nb)
sc = DeviceScene([body, wall], ['body'], ['wall'], dim=3, gy=-9.81, eta_uniform=info['eta_uniform'])
sc.gtvf_step(1e-4, settle, graph=True)
torch.cuda.synchronize()
cnt = sc.T['nbr_cnt'].cpu().numpy().astype(np.int64)
n = cnt.size
print('nlist mean %.2f max %d  p50 %d p90 %d p99 %d  frac0 %.3f' % (cnt.mean(), cnt.max(), np.percentile(cnt,50), np.percentile(cnt,90), np.percentile(cnt,99), (cnt==0).mean()))
# per-warp (chunks of 100 particles -> 4 warps: 32,32,32,4)
c = cnt.reshape(nb, 100)
w = np.stack([c[:, :32].max(1), c[:, 32:64].max(1), c[:, 64:96].max(1), c[:, 96:].max(1)], 1)
ws = np.stack([c[:, :32].sum(1), c[:, 32:64].sum(1), c[:, 64:96].sum(1), c[:, 96:].sum(1)], 1)
print('per-warp max mean %.2f ; sum mean %.1f ; utilisation sum/(32*max) %.3f' % (w.mean(), ws.mean(), ws.sum()/(32.0*w.sum())))
# keys per particle
dem = sc.T['nbr_dem'].view(sc.list_cap, -1)
cntg = sc.T['nbr_cnt'].long()
nk = torch.zeros_like(cntg)
mx = int(cnt.max())
d = dem[:mx].clone()
ar = torch.arange(mx, device=d.device)[:, None]
d[ar >= cntg[None, :]] = 2**31-1
ds, _ = torch.sort(d, 0)
nk = ((ds[1:] != ds[:-1]) & (ds[1:] != 2**31-1)).sum(0) + (ds[0] != 2**31-1).long()
nk = nk.cpu().numpy()
print('keys/particle mean %.2f max %d ; per-chunk max keys mean %.2f' % (nk.mean(), nk.max(), nk.reshape(nb,100).max(1).mean()))
print(sc.read_counters())
yv = sc.B['xcm'].view(-1,3)[:,1].cpu().numpy()
print('body y range', yv.min(), yv.max())

import sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import rbo
from tests.util import load_config, oracle_params
from rigid_body_2d_3d_pysph_b200.device import DeviceScene
name=sys.argv[1]
garr, meta = load_config(name); oarr,_ = load_config(name)
sc = DeviceScene(garr, meta['rigid'], meta['boundaries'], dim=meta['dim'], kr=meta['kr'], kf=meta['kf'], fric_coeff=meta['fric_coeff'], gx=meta['gx'], gy=meta['gy'], gz=meta['gz'])
p = oracle_params(meta)
done=0
for step in [1,5,10,20,30,40,60,80,100,150,200]:
    sc.gtvf_step(meta['dt'], step-done); rbo.gtvf_step(oarr, meta['rigid'], p, nsteps=step-done); done=step
    g,o=garr[0],oarr[0]
    f=np.sqrt(o.fx**2+o.fy**2+o.fz**2).sum()
    ef=max(np.abs(g.fx-o.fx).max(), np.abs(g.fy-o.fy).max(), np.abs(g.fz-o.fz).max())
    print(step, 'dF/sum|f| %.2e' % (ef/f), 'dxcm %.2e' % np.abs(g.xcm-o.xcm).max(), 'dR %.2e' % np.abs(g.R-o.R).max(), 'dvcm %.2e'%np.abs(g.vcm-o.vcm).max(), 'max|vcm| %.2e'%np.abs(o.vcm).max())

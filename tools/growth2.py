import sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import rbo
from tests.util import load_config, oracle_params
from rigid_body_2d_3d_pysph_b200.device import DeviceScene
name='benchmark_5_3d'
garr, meta = load_config(name); oarr,_ = load_config(name)
sc = DeviceScene(garr, meta['rigid'], meta['boundaries'], dim=meta['dim'], kr=meta['kr'], kf=meta['kf'], fric_coeff=meta['fric_coeff'], gx=meta['gx'], gy=meta['gy'], gz=meta['gz'])
p = oracle_params(meta)
sc.gtvf_step(meta['dt'], 100); rbo.gtvf_step(oarr, meta['rigid'], p, nsteps=100)
g,o=garr[0],oarr[0]
tnb=7
for step in range(101,150):
    sc.gtvf_step(meta['dt'], 1); rbo.gtvf_step(oarr, meta['rigid'], p, nsteps=1)
    f=np.sqrt(o.fx**2+o.fy**2+o.fz**2).sum()
    d=np.sqrt((g.fx-o.fx)**2+(g.fy-o.fy)**2+(g.fz-o.fz)**2)
    i=int(np.argmax(d))
    nact=int((o.overlap>0).sum())
    msg='%d dF/sum %.2e  i=%d  nactive=%d' % (step, d[i]/f, i, nact)
    if d[i]/f>1e-10:
        sl=slice(tnb*i,tnb*i+tnb)
        msg+='\n   oracle f=(%.6e,%.6e,%.6e) gpu f=(%.6e,%.6e,%.6e)'%(o.fx[i],o.fy[i],o.fz[i],g.fx[i],g.fy[i],g.fz[i])
        msg+='\n   overlap %s\n   ft_x %s ft_z %s\n   fn_y %s\n u,v,w=(%.3e,%.3e,%.3e) vsrc=%s' % (o.overlap[sl], o.ft_x[sl], o.ft_z[sl], o.fn_y[sl], o.u[i],o.v[i],o.w[i], o.vx_source[sl])
        hk,hd,hf = sc.history()
        msg+='\n   gpu hist keys %s fn_y %s' % (hk[:,i], hf[1,:,i])
    print(msg)
    if d[i]/f>1e-8: break

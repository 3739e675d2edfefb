"""Host/device coherence bookkeeping of the ParticleArray restatement: a read
hands out the live NumPy array, and only what was really modified counts as
host-touched at the next step (a post_step callback that only
looks at pa.m / pa.x must not force uploads and rebuilds)."""
import numpy as np

from rigid_body_2d_3d_pysph_b200.compat.particle_array import \
    get_particle_array


class _FakeDevice(object):
    def pull(self, pa, name):
        pass


def test_reads_are_not_writes_once_a_device_is_bound():
    pa = get_particle_array(name='a', x=np.arange(5.), y=np.zeros(5), m=1.0)
    pa.add_constant('xcm', [0., 0., 0.])
    # before a scene is bound every access counts (setup code)
    pa.x
    assert 'x' in pa.__dict__['_host_touched']
    pa.__dict__['_device'] = _FakeDevice()
    pa.__dict__['_host_touched'].clear()
    assert pa.x[2] == 2. and pa.m[0] == 1. and pa.xcm[0] == 0.
    assert not pa.__dict__['_host_touched']
    assert pa.modified_since_read() == set()
    # in-place modification through a handed-out array
    pa.x[:] += 0.25
    pa.m
    assert pa.modified_since_read() == {'x'}
    assert pa.modified_since_read() == set()
    # explicit writes
    pa.y = np.ones(5)
    pa.xcm[:] = 1.
    pa.touch('xcm')
    assert pa.__dict__['_host_touched'] == {'y', 'xcm'}
    # x += ... goes through __setattr__ as well
    pa.__dict__['_host_touched'].clear()
    pa.x += 1.
    assert 'x' in pa.__dict__['_host_touched']

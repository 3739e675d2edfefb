"""DEMScheme path (LVCDisplacement + DEMStep) on the GPU against the fixture
produced by the reference's own dem.py methods."""
import numpy as np
import pytest

from tests.test_oracle_dem import (DEM_CASES, DEM_STATE, canonical_history,
                                   load_dem)
from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', DEM_CASES)
def test_dem_gpu_matches_reference(name):
    from rigid_body_2d_3d_pysph_b200.dem import DemDeviceScene
    arrays, ref, meta = load_dem(name)
    limit = meta['limit']
    sc = DemDeviceScene(arrays, meta['granular'], meta['boundaries'],
                        dim=meta['dim'], gx=meta['gx'], gy=meta['gy'],
                        gz=meta['gz'], radius_scale=meta['radius_scale'])
    sand = arrays[0]
    done = 0
    for step in meta['save_steps']:
        sc.gtvf_step(meta['dt'], step - done)
        done = step
        sc.check_status()
        pre = 'ref/%d/sand/' % step
        assert np.array_equal(sand.total_tng_contacts,
                              ref[pre + 'total_tng_contacts'])
        got = canonical_history(sand.tng_idx, sand.tng_idx_dem_id, sand.tng_x,
                                sand.tng_y, sand.tng_z, limit)
        want = canonical_history(*[ref[pre + n] for n in (
            'tng_idx', 'tng_idx_dem_id', 'tng_x', 'tng_y', 'tng_z')], limit)
        assert np.array_equal(got[0], want[0]) and \
            np.array_equal(got[1], want[1]), step
        tscale = max(np.abs(want[2]).max(), np.abs(want[3]).max(),
                     np.abs(want[4]).max(), 1e-12)
        for k in (2, 3, 4):
            assert_close(got[k], want[k], 1e-10, 'dem step %d tng[%d]' %
                         (step, k), tscale)
        fs = np.abs(ref[pre + 'fx']).max() + np.abs(ref[pre + 'fy']).max()
        for n in DEM_STATE:
            if n.startswith('tng'):
                continue
            scale = fs if n[0] == 'f' else None
            if n.startswith('tor'):
                scale = fs * 0.01
            assert_close(getattr(sand, n), ref[pre + n], 1e-10,
                         'dem step %d %s' % (step, n), scale)


def test_dem_scheme_surface():
    """DEMScheme -> equations -> planner -> DemDeviceScene via the Solver."""
    from rigid_body_2d_3d_pysph_b200.dem import DEMScheme
    arrays, ref, meta = load_dem()
    s = DEMScheme(['sand'], ['wall'], dim=2, gy=-9.81)
    s.configure_solver(dt=meta['dt'], tf=10 * meta['dt'], pfreq=1000)
    solver = s.get_solver()
    solver.set_disable_output(True)
    solver.setup(arrays, s.get_equations(), kernel=solver.kernel)
    solver.solve()
    sand = arrays[0]
    assert solver.count == 10
    assert_close(sand.fx, ref['ref/10/sand/fx'], 1e-10, 'fx',
                 np.abs(ref['ref/10/sand/fx']).max())
    assert_close(sand.x, ref['ref/10/sand/x'], 1e-12, 'x')

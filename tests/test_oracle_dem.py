"""C oracle of the DEMScheme path against the reference's dem.py methods."""
import json
import os

import numpy as np
import pytest

from oracle import rbo
from rigid_body_2d_3d_pysph_b200.compat.output import load_scene
from tests.util import GOLDEN, assert_close

DEM_STATE = ['x', 'y', 'z', 'u', 'v', 'w', 'wx', 'wy', 'wz', 'fx', 'fy', 'fz',
             'torx', 'tory', 'torz', 'tng_x', 'tng_y', 'tng_z']


def canonical_history(idx, dem, tx, ty, tz, limit):
    """The reference builds its source list with ``list(set(...))``
    (dem.py:706), so the order in which a particle's contacts are appended
    depends on string hashing; compare the lists as sets keyed by
    (dem_id, idx)."""
    n = idx.size // limit
    out = []
    for a in (idx, dem, tx, ty, tz):
        out.append(np.array(a).reshape(n, limit).copy())
    key = out[1].astype(np.int64) * (1 << 32) + out[0].astype(np.int64)
    key[out[0] < 0] = np.iinfo(np.int64).max
    order = np.argsort(key, axis=1, kind='stable')
    return [np.take_along_axis(a, order, 1) for a in out]


DEM_CASES = ['dem2d', 'dem3d']


def load_dem(name='dem2d'):
    arrays, _ = load_scene(os.path.join(GOLDEN, name + '_scene.npz'))
    ref = np.load(os.path.join(GOLDEN, name + '_ref.npz'))
    meta = json.loads(str(ref['__meta__']))
    meta.setdefault('limit', 8)
    return arrays, ref, meta


@pytest.mark.parametrize('name', DEM_CASES)
def test_dem_oracle_matches_reference(name):
    arrays, ref, meta = load_dem(name)
    limit = meta['limit']
    p = rbo.make_params(meta['dim'], meta['dt'], gx=meta['gx'], gy=meta['gy'],
                        gz=meta['gz'], radius_scale=meta['radius_scale'])
    sand = arrays[0]
    done = 0
    for step in meta['save_steps']:
        rbo.dem_step(arrays, meta['granular'], p, nsteps=step - done)
        done = step
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
            assert_close(got[k], want[k], 1e-10, '%s step %d tng[%d]' %
                         (name, step, k), tscale)
        fs = np.abs(ref[pre + 'fx']).max() + np.abs(ref[pre + 'fy']).max()
        for n in DEM_STATE:
            if n.startswith('tng'):
                continue
            scale = fs if n[0] == 'f' else None
            if n.startswith('tor'):
                scale = fs * 0.01
            assert_close(getattr(sand, n), ref[pre + n], 1e-10,
                         '%s step %d %s' % (name, step, n), scale)
    assert sand.total_tng_contacts.sum() > 50

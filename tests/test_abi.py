"""The C-ABI library loads and exports every symbol include/rbx.h declares
(no compute calls: this test runs without a GPU)."""
import ctypes
import os
import re

from rigid_body_2d_3d_pysph_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from rigid_body_2d_3d_pysph_b200.csrc import build
    build.build()


def test_header_symbols_exported():
    _ensure_built()
    hdr = open(os.path.join(ROOT, 'include', 'rbx.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(rbx_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(L, sym), 'missing symbol %s' % sym
    assert declared == set(_lib.SYMBOLS)


def test_struct_mirrors_match_library():
    _ensure_built()
    L = _lib.load()      # raises on any sizeof mismatch
    assert L.rbx_version() == 404
    assert L.rbx_strerror(-2) == b'workspace too small'


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    try:
        _lib.load()
    except _lib.RbxError as e:
        assert 'no CPU fallback' in str(e)
    else:
        raise AssertionError('load() must raise when librbx.so is missing')


def test_reference_script_runs_to_the_device_boundary():
    """An unmodified reference script, run through the compat layer, gets all
    the way to building the device scene and then fails LOUDLY without a GPU
    (no CPU fallback).  Skipped where /root/reference is absent (GPU box)."""
    import subprocess
    import sys
    import pytest
    script = '/root/reference/code/benchmark_2_multiple_rigid_bodies_colliding.py'
    if not os.path.exists(script):
        pytest.skip('reference not mounted')
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present: covered by the gpu tests')
    _ensure_built()
    out = subprocess.run(
        [sys.executable, '-m', 'rigid_body_2d_3d_pysph_b200.run', script,
         '--tf', '0.001', '--disable-output'], cwd=ROOT,
        capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert 'no CPU fallback' in out.stderr, out.stderr[-2000:]

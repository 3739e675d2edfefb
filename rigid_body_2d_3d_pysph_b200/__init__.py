"""B200-native rigid-body DEM hot path behind the PySPH scheme surface.

Hot path: cell-list neighbour search, Mofidi-style 3-pass
contact (rigid_body_common.py:631-1032 of the reference), per-body
force/torque reduction and the GTVF / RK2 rigid steppers, written as sm_100a
CUDA kernels behind the C ABI declared in include/rbx.h.  Host code is
Python; PyTorch owns device memory, streams and torch.distributed.
"""
__version__ = '0.1.0'

"""ctypes binding of librbx.so (include/rbx.h).

The product path has no CPU fallback: if the CUDA library cannot be loaded
this module raises, loudly, at first use.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RBX_LIB') or \
    os.path.join(_HERE, 'csrc', 'librbx.so')

c_i32, c_i64, c_f64, c_vp = (ctypes.c_int32, ctypes.c_int64, ctypes.c_double,
                              ctypes.c_void_p)

RBX_MAX_KEYS = 8
STATUS_SLOT_OVERFLOW = 1
STATUS_HIST_OVERFLOW = 2
STATUS_LIST_OVERFLOW = 4
STATUS_GRID_COARSENED = 8
STATUS_LVC_OVERFLOW = 16
STATUS_PAIR_OVERFLOW = 32
PARAM_EXACT = 1
PARAM_BODY_VEL = 2
PARAM_DENSE_OUT = 4

POSE_POS, POSE_VEL, POSE_VEL_PREV, POSE_NORMALS = 1, 2, 4, 8


class RbxGridInfo(ctypes.Structure):
    _fields_ = [('x0', c_f64), ('y0', c_f64), ('z0', c_f64), ('cell', c_f64),
                ('inv_cell', c_f64), ('nx', c_i32), ('ny', c_i32),
                ('nz', c_i32), ('ncells', c_i32), ('npoints', c_i32),
                ('pad_', c_i32)]


class RbxPoints(ctypes.Structure):
    _fields_ = [('n', c_i32), ('pad_', c_i32), ('index', c_vp), ('x', c_vp),
                ('y', c_vp), ('z', c_vp), ('h', c_vp), ('dem_id', c_vp)]


class RbxCells(ctypes.Structure):
    _fields_ = [('cap_cells', c_i32), ('cap_points', c_i32), ('info', c_vp),
                ('cell_start', c_vp), ('cell_of', c_vp), ('rank', c_vp),
                ('gidx', c_vp), ('sx', c_vp), ('sy', c_vp), ('sz', c_vp),
                ('sh', c_vp), ('sdem', c_vp), ('cond', c_vp)]


_SCENE_INTS = ['n_total', 'n_rigid', 'n_bodies', 'n_chunks', 'dim', 'ks',
               'eta_mode', 'planar', 'list_cap', 'pad_']
_SCENE_PTRS = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'dem_id',
               'fx', 'fy', 'fz', 'dx0', 'dy0', 'dz0', 'body', 'is_boundary',
               'normal0', 'normal', 'chunk_start', 'chunk_body', 'body_chunk',
               'nbr_pos', 'nbr_cnt', 'nbr_order', 'nbr_cnt_srt',
               'nbr_srt', 'total_mass', 'izz', 'spacing0', 'xcm', 'vcm',
               'ang_mom', 'omega', 'force', 'torque', 'R', 'R_prev', 'iinv_b',
               'iinv_g', 'xcm0', 'vcm0', 'ang_mom0', 'R0', 'eta', 'eta_row',
               'hist_key_in', 'hist_dlt_in', 'hist_fn_in', 'hist_key_out',
               'hist_dlt_out', 'hist_fn_out', 'status', 'counters', 'rebuild',
               'xcm_ref', 'R_ref', 'rmax', 'body_first', 'pos32', 'clist']


class RbxScene(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in _SCENE_INTS] + \
               [(n, c_vp) for n in _SCENE_PTRS] + [('origin', c_f64 * 3)] + \
               [(n, c_vp) for n in ('alist_out', 'alist_prev', 'acount_out',
                                    'acount_prev', 'body_tag', 'aux32')] + \
               [('h_uniform', c_f64), ('gravity', c_f64 * 3),
                ('static_ref', c_vp), ('aux_stream', c_vp)]


_DEM_PTRS = ['x', 'y', 'z', 'u', 'v', 'w', 'wx', 'wy', 'wz', 'h', 'm', 'rad_s',
             'moi', 'dem_id', 'arr_off', 'arr_start', 'tbl_row', 'kn', 'kt',
             'alpha', 'mu', 'fx', 'fy', 'fz', 'torx', 'tory', 'torz',
             'tng_idx', 'tng_dem', 'tng_x', 'tng_y', 'tng_z', 'total_tng',
             'status']


class RbxDemScene(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in ('n_total', 'n_dest', 'n_arrays',
                                     'limit')] + \
               [(n, c_vp) for n in _DEM_PTRS]


class RbxParams(ctypes.Structure):
    _fields_ = [('radius_scale', c_f64), ('kr', c_f64), ('kf', c_f64),
                ('fric_coeff', c_f64), ('gx', c_f64), ('gy', c_f64),
                ('gz', c_f64), ('dt', c_f64), ('reach', c_f64),
                ('h_uniform', c_f64), ('skin', c_f64), ('flags', c_i32),
                ('pad_', c_i32)]


class RbxCanelas(ctypes.Structure):
    _fields_ = [('rad_s', c_vp), ('E', c_vp), ('nu', c_vp),
                ('src_body', c_vp), ('Cn', c_f64)]


class RbxDiag(ctypes.Structure):
    _fields_ = [('key', c_vp), ('closest', c_vp), ('nx', c_vp), ('ny', c_vp),
                ('nz', c_vp), ('dist', c_vp), ('overlap', c_vp),
                ('ftx', c_vp), ('fty', c_vp), ('ftz', c_vp), ('pairs', c_vp),
                ('pair_count', c_vp), ('pair_cap', c_i64)]


# every symbol include/rbx.h declares
SYMBOLS = ['rbx_version', 'rbx_strerror', 'rbx_sizeof',
           'rbx_cells_workspace_bytes', 'rbx_cells_build', 'rbx_pairs_dump',
           'rbx_contact_mofidi', 'rbx_contact_neighbours',
           'rbx_contact_slots', 'rbx_contact_canelas', 'rbx_reduce_bodies', 'rbx_gtvf_kick',
           'rbx_gtvf_drift', 'rbx_pose_particles', 'rbx_pos32_refresh',
           'rbx_halo_pack', 'rbx_static_update',
           'rbx_halo_unpack', 'rbx_rk2_stage',
           'rbx_gtvf_step', 'rbx_contact_lvc', 'rbx_dem_step',
           'rbx_boundary_identify', 'rbx_setup_bodies']

_lib = None


class RbxError(RuntimeError):
    pass


def load():
    """Load librbx.so; raise if it is missing (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RbxError(
            'librbx.so not found at %s -- build it with '
            '`python -m rigid_body_2d_3d_pysph_b200.csrc.build` '
            '(or __graft_entry__.build()); there is no CPU fallback.' %
            LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for s in SYMBOLS:
        getattr(L, s)
    L.rbx_version.restype = ctypes.c_int
    L.rbx_strerror.restype = ctypes.c_char_p
    L.rbx_strerror.argtypes = [ctypes.c_int]
    L.rbx_sizeof.restype = ctypes.c_size_t
    L.rbx_sizeof.argtypes = [ctypes.c_int]
    L.rbx_cells_workspace_bytes.restype = ctypes.c_size_t
    L.rbx_cells_workspace_bytes.argtypes = [c_i32, c_i32]
    P = ctypes.POINTER
    L.rbx_cells_build.argtypes = [P(RbxPoints), P(RbxCells), c_f64, c_vp,
                                  c_vp, ctypes.c_size_t, c_vp]
    L.rbx_pairs_dump.argtypes = [P(RbxPoints), P(RbxCells), c_f64, c_vp,
                                 c_vp, c_vp, c_vp]
    L.rbx_contact_mofidi.argtypes = [P(RbxScene), P(RbxCells), P(RbxParams),
                                     P(RbxDiag), c_vp]
    L.rbx_contact_neighbours.argtypes = [P(RbxScene), P(RbxCells),
                                         P(RbxParams), c_vp]
    L.rbx_contact_slots.argtypes = [P(RbxScene), P(RbxCells), P(RbxParams),
                                    P(RbxDiag), c_vp]
    L.rbx_contact_canelas.argtypes = [P(RbxScene), P(RbxCells), P(RbxParams),
                                      P(RbxCanelas), c_vp]
    L.rbx_reduce_bodies.argtypes = [P(RbxScene), c_vp]
    L.rbx_gtvf_kick.argtypes = [P(RbxScene), c_f64, c_vp]
    L.rbx_gtvf_drift.argtypes = [P(RbxScene), c_f64, c_f64, c_vp]
    L.rbx_pose_particles.argtypes = [P(RbxScene), ctypes.c_int, c_vp]
    L.rbx_pos32_refresh.argtypes = [P(RbxScene), c_i32, c_i32, c_vp]
    L.rbx_halo_pack.argtypes = [P(RbxScene), c_vp, c_i32, c_vp, ctypes.c_int,
                                c_vp]
    L.rbx_halo_unpack.argtypes = [P(RbxScene), c_i32, c_i32, c_vp, c_f64,
                                  c_vp]
    L.rbx_static_update.argtypes = [P(RbxScene), c_i32, c_i32, c_vp, c_vp,
                                    c_vp, c_vp, c_vp, c_vp, c_f64, c_vp]
    L.rbx_rk2_stage.argtypes = [P(RbxScene), ctypes.c_int, c_f64,
                                ctypes.c_int, c_f64, c_vp]
    L.rbx_gtvf_step.argtypes = [P(RbxScene), P(RbxPoints), P(RbxCells),
                                P(RbxParams), c_vp, ctypes.c_size_t,
                                ctypes.c_int, c_vp]
    L.rbx_contact_lvc.argtypes = [P(RbxDemScene), P(RbxCells), P(RbxParams),
                                  c_vp]
    L.rbx_dem_step.argtypes = [P(RbxDemScene), ctypes.c_int, c_f64, c_vp]
    L.rbx_boundary_identify.argtypes = [P(RbxPoints), P(RbxCells),
                                        ctypes.c_int, c_f64, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp]
    L.rbx_setup_bodies.argtypes = [c_i32] + [c_vp] * 14
    for i, cls in enumerate([RbxGridInfo, RbxPoints, RbxCells, RbxScene,
                             RbxParams, RbxDiag, RbxDemScene, RbxCanelas]):
        if L.rbx_sizeof(i) != ctypes.sizeof(cls):
            raise RbxError('ABI mismatch for %s: library %d bytes, binding %d'
                           % (cls.__name__, L.rbx_sizeof(i),
                              ctypes.sizeof(cls)))
    _lib = L
    return L


def check(rc, what=''):
    if rc != 0:
        raise RbxError('%s failed: %s (%d)' % (
            what, load().rbx_strerror(rc).decode(), rc))

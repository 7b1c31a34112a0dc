"""ctypes mirror of include/branchmpc.h and the loader of libbranchmpc.so.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded, `load_library()` raises.
"""
import ctypes as C
import os

MAX_N, MAX_D, MAX_POLICIES, MAX_ROWS, MAX_NB = 4, 3, 4, 4, 3

MODEL_HIGHWAY, MODEL_QUADRUPED, MODEL_MERGE = 0, 1, 2
CTRL_BRANCH, CTRL_PROX, CTRL_ROBUST, CTRL_CVAR, CTRL_BELIEF = 0, 1, 2, 3, 4
POLICY_MAINTAIN, POLICY_BRAKE, POLICY_LC, POLICY_TRACKV, POLICY_FORWARD, POLICY_STOP, POLICY_TRACKV_REF, POLICY_BRAKE_REF = range(8)
STATUS_POLISHED, STATUS_CONVERGED, STATUS_MAXITER, STATUS_NUMERIC = range(4)
SLAB_AUTO, SLAB_SHARED, SLAB_SPLIT, SLAB_GLOBAL = range(4)
OK, E_INVALID, E_CUDA, E_CAPACITY, E_UNSUPPORTED = 0, -1, -2, -3, -4

_dbl = C.c_double
_i32 = C.c_int32


class Config(C.Structure):
    """struct bmpc_config"""
    _fields_ = [
        ("model", _i32), ("controller", _i32), ("n", _i32), ("d", _i32), ("N", _i32), ("NB", _i32), ("m", _i32),
        ("dt", _dbl),
        ("policy_kind", _i32 * MAX_POLICIES),
        ("policy_param", (_dbl * 4) * MAX_POLICIES),
        ("Q", _dbl * (MAX_N * MAX_N)), ("Qf", _dbl * (MAX_N * MAX_N)), ("R", _dbl * (MAX_D * MAX_D)),
        ("dR", _dbl * MAX_D), ("Qslack", _dbl * 2),
        ("n_rows", _i32),
        ("row_f", (_dbl * MAX_N) * MAX_ROWS), ("row_lo", _dbl * MAX_ROWS), ("row_hi", _dbl * MAX_ROWS),
        ("u_lo", _dbl * MAX_D), ("u_hi", _dbl * MAX_D),
        ("veh_L", _dbl), ("veh_W", _dbl), ("Kpsi", _dbl), ("s1", _dbl), ("lane_lo", _dbl), ("lane_hi", _dbl),
        ("quad_margin", _dbl),
        ("max_iter", _i32), ("polish_first", _i32), ("polish_every", _i32), ("polish_passes", _i32),
        ("polish_al_iters", _i32), ("polish_careful", _i32), ("warm_polish", _i32), ("rho_refresh", _i32),
        ("alpha", _dbl), ("theta", _dbl), ("theta_u", _dbl), ("eps_abs", _dbl), ("polish_big", _dbl),
        ("polish_mult", _dbl), ("cvar_alpha", _dbl),
        ("hmm_M", _i32), ("hmm_col_alpha", _dbl), ("hmm_tran_diag", _dbl), ("hmm_thres", _dbl),
        ("slab_mode", _i32), ("batch_capacity", _i32), ("device", _i32), ("reserved", _i32 * 8),
    ]


_pd = C.POINTER(_dbl)
_pi = C.POINTER(_i32)


class HmmParams(C.Structure):
    """struct bmpc_hmm_params"""
    _fields_ = [(k, _dbl) for k in ("Kpsi", "L", "W", "ylb", "yub", "col_alpha", "s1", "s2", "c2", "tran_diag")]


HMM_MAINTAIN, HMM_BRAKE = 0, 1


class Outputs(C.Structure):
    """struct bmpc_outputs (device pointers for bmpc_solve, host pointers for bmpc_solve_host)"""
    _fields_ = [("u0", C.c_void_p), ("uPred", C.c_void_p), ("xPred", C.c_void_p), ("xLin", C.c_void_p),
                ("zPred", C.c_void_p), ("branch_w", C.c_void_p), ("branch_p", C.c_void_p), ("objective", C.c_void_p),
                ("status", C.c_void_p), ("iters", C.c_void_p), ("nfact", C.c_void_p), ("nsolve", C.c_void_p), ("cycles", C.c_void_p),
                ("bPred", C.c_void_p)]


OUTPUT_NAMES = [f[0] for f in Outputs._fields_]

class EnvState(C.Structure):
    """bmpc_env_state (include/branchmpc.h)."""
    _fields_ = [("x", C.c_void_p), ("z", C.c_void_p), ("lane", C.c_void_p), ("policy_params", C.c_void_p),
                ("goal", C.c_void_p), ("obs_policy", C.c_void_p), ("collided", C.c_void_p), ("xref", C.c_void_p),
                ("u_obs", C.c_void_p)]


class MergeEnvState(C.Structure):
    """bmpc_merge_env_state (include/branchmpc.h)."""
    _fields_ = [("x", C.c_void_p), ("z", C.c_void_p), ("lane_id", C.c_void_p), ("collided", C.c_void_p), ("xref", C.c_void_p),
                ("S", C.c_void_p), ("state_bounds", C.c_void_p), ("u_obs", C.c_void_p), ("table_x", C.c_void_p),
                ("table_y", C.c_void_p), ("table_psi", C.c_void_p), ("table_n", _i32)]


# every symbol include/branchmpc.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("bmpc_version", C.c_int, []),
    ("bmpc_create", C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    ("bmpc_destroy", C.c_int, [C.c_void_p]),
    ("bmpc_reset", C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int64]),
    ("bmpc_num_branches", C.c_int, [C.c_void_p]),
    ("bmpc_total_x", C.c_int, [C.c_void_p]),
    ("bmpc_total_u", C.c_int, [C.c_void_p]),
    ("bmpc_get_topology", C.c_int, [C.c_void_p, _pi, _pi, _pi, _pi]),
    ("bmpc_ulin_rows", C.c_int, [C.c_void_p]),
    ("bmpc_solve", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                             C.POINTER(Outputs), C.c_void_p]),
    ("bmpc_set_lookup_table", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    ("bmpc_solve_transformed", C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int64, C.POINTER(Outputs), C.c_void_p]),
    ("bmpc_solve_transformed_host_views", C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int64, C.POINTER(Outputs),
                                                                                      C.POINTER(Outputs)]),
    ("bmpc_solve_belief", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                    C.POINTER(Outputs), C.c_void_p]),
    ("bmpc_eval_belief", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 6 + [C.c_void_p]),
    ("bmpc_solve_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.POINTER(Outputs)]),
    ("bmpc_solve_host_views", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.POINTER(Outputs), C.POINTER(Outputs)]),
    ("bmpc_get_state", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    ("bmpc_set_state", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    ("bmpc_eval_model", C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 8 + [C.c_void_p]),
    ("bmpc_hmm_backup_rollout", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, _pi, C.c_int32, _dbl, _dbl, C.c_void_p,
                                          C.c_int32, C.c_void_p]),
    ("bmpc_hmm_rollout_sensitivity", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _pi, C.c_int32, _dbl, _dbl, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    ("bmpc_hmm_belief_update", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                         C.POINTER(HmmParams), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p]),
    ("bmpc_plant_step", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                  C.c_void_p]),
    ("bmpc_env_step", C.c_int, [C.c_void_p, C.POINTER(EnvState), C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                C.POINTER(Outputs), C.c_void_p]),
    ("bmpc_staging_enabled", C.c_int, [C.c_void_p]),
    ("bmpc_env_step_merge", C.c_int, [C.c_void_p, C.POINTER(MergeEnvState), C.c_int64, C.c_int32, C.c_int32, _dbl, _dbl,
                                      C.POINTER(Outputs), C.c_void_p]),
    ("bmpc_get_launch_info", C.c_int, [C.c_void_p, _pi, _pi, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("bmpc_launch_count", C.c_int64, [C.c_void_p]),
    ("bmpc_measure_fp64_peak", C.c_double, [C.c_int, C.c_int]),
    ("bmpc_last_kernel_ms", C.c_float, [C.c_void_p]),
    ("bmpc_last_error", C.c_char_p, [C.c_void_p]),
]

LIB_NAME = "libbranchmpc.so"
_lib = None


def library_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), LIB_NAME)


def load_library():
    """dlopen the in-tree CUDA library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("%s is not built (run `python -c 'import __graft_entry__ as g; g.build()'` at the repo "
                           "root); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

"""ctypes binding of libturbopy_b200.so (the C ABI in include/turbopy_b200.h).

There is deliberately no CPU fallback: if the shared library has not been
built, or no sm_100 device is present, every compute call raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libturbopy_b200.so")

# ---- constants mirrored from include/turbopy_b200.h -------------------------
TP_OK = 0
TP_E_NODEVICE = -3
TP_FIELD_UNIFORM_HOST, TP_FIELD_UNIFORM_DEV, TP_FIELD_PER_PARTICLE = 0, 1, 2
TP_INTERP_A, TP_INTERP_B, TP_INTERP_C = 1, 2, 3
TP_ST_BELOW, TP_ST_ABOVE, TP_ST_BAD_WINDOW = 1, 2, 4

c_double_p = C.POINTER(C.c_double)
c_u64_p = C.POINTER(C.c_uint64)
c_i64_p = C.POINTER(C.c_int64)


class Particles(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("pos_stride_n", C.c_int64), ("pos_stride_c", C.c_int64),
                ("mom", C.c_void_p), ("mom_stride_n", C.c_int64), ("mom_stride_c", C.c_int64),
                ("n", C.c_int64)]


class PushConsts(C.Structure):
    _fields_ = [("charge", C.c_double), ("mass_sq", C.c_double), ("dt", C.c_double), ("c2", C.c_double)]


class GridFields(C.Structure):
    _fields_ = [("r", C.c_void_p), ("ng", C.c_int64), ("dr", C.c_double),
                ("r_first", C.c_double), ("r_last", C.c_double),
                ("comp", C.c_void_p * 6), ("comp_stride", C.c_int64 * 6)]


class Linspace(C.Structure):
    _fields_ = [("start", C.c_double), ("span", C.c_double), ("step", C.c_double), ("n", C.c_int64)]


#: every symbol include/turbopy_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tp_version": (C.c_int, []),
    "tp_error_string": (C.c_char_p, [C.c_int]),
    "tp_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3 + [c_i64_p, c_i64_p]),
    "tp_boris_push_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts),
                                    C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]),
    "tp_boris_push_moments_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts),
                                            C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                            C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                            C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tp_boris_push_steps_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts),
                                          C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "tp_gather_push_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts), C.POINTER(GridFields),
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tp_gather_push_steps_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts), C.POINTER(GridFields),
                                           C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tp_gather_push_windows_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts), C.POINTER(GridFields),
                                             C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    "tp_gather_push_moments_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts),
                                                     C.POINTER(GridFields), C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                                     C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "tp_gather_windows_len": (C.c_int64, [C.POINTER(Particles), C.POINTER(GridFields)]),
    "tp_gather_windows_capacity": (C.c_int64, [C.POINTER(Particles), C.POINTER(GridFields), C.c_int]),
    "tp_gather_windows_init": (C.c_int, [C.POINTER(Particles), C.POINTER(GridFields), C.c_int, C.c_void_p, C.c_int64,
                                         C.c_void_p]),
    "tp_status_reset": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tp_gather_smem_bytes": (C.c_int, [C.POINTER(GridFields), c_i64_p, C.POINTER(C.c_int)]),
    "tp_interp1d_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                  C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "tp_interp1d_lin_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                      C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_int, C.POINTER(Linspace), C.c_void_p, C.c_void_p]),
    "tp_fd_centered_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]),
    "tp_fd_upwind_left_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tp_fd_del2_radial_apply_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_double, C.c_double, C.c_void_p]),
    "tp_fd_del2_radial_axpy_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Linspace), C.c_void_p, C.c_int64,
                                             C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "tp_fd_dia_apply_f64": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_void_p]),
    "tp_fd_dia_const_apply_f64": (C.c_int, [c_double_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), c_i64_p,
                                            c_double_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tp_fd_upwind_left_lin_f64": (C.c_int, [C.c_void_p, C.POINTER(Linspace), C.c_void_p, C.c_int64, C.c_void_p]),
    "tp_fd_del2_radial_apply_lin_f64": (C.c_int, [C.c_void_p, C.POINTER(Linspace), C.c_void_p, C.c_int64,
                                                  C.c_double, C.c_double, C.c_void_p]),
    "tp_poisson_radial_solve_lin_f64": (C.c_int, [C.c_void_p, C.POINTER(Linspace), C.c_double, C.c_void_p, C.c_int64,
                                                  C.c_void_p, C.c_void_p]),
    "tp_poisson_scratch_doubles": (C.c_int64, [C.c_int64]),
    "tp_poisson_radial_solve_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_void_p]),
    "tp_reduce_scratch_doubles": (C.c_int64, []),
    "tp_reduce_moments_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_double, C.c_double, C.c_double,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "tp_nccl_version": (C.c_int, [C.POINTER(C.c_int)]),
    "tp_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "tp_nccl_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "tp_nccl_allreduce_sum_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tp_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "tp_boris_push_host_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts),
                                         C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "tp_gather_push_host_f64": (C.c_int, [C.POINTER(Particles), C.POINTER(PushConsts), C.POINTER(GridFields),
                                          C.c_int, C.c_int, c_u64_p]),
    "tp_host_last_transfer": (C.c_int, [c_i64_p, c_i64_p]),
    "tp_host_pin": (C.c_int, [C.c_void_p, C.c_int64]),
    "tp_host_unpin": (C.c_int, [C.c_void_p]),
    "tp_host_release": (C.c_int, []),
    "tp_host_link_probe": (C.c_int, [C.c_int64, C.c_int, c_double_p, c_double_p, c_double_p]),
    "tp_plane_wave_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tp_clock_advance_f64": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p]),
    "tp_sort_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int]),
    "tp_sort_by_cell_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "tp_resort_blocks_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int64, C.c_void_p]),
    "tp_resort_block_particles": (C.c_int64, []),
    "tp_permute_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tp_graph_begin": (C.c_int, [C.c_void_p]),
    "tp_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "tp_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tp_graph_destroy": (C.c_int, [C.c_void_p]),
    "tp_launch_count": (C.c_int64, []),
    "tp_selftest_division": (C.c_int, [C.c_int64, C.c_uint64, c_i64_p, c_i64_p, C.c_void_p]),
    "tp_selftest_roots": (C.c_int, [C.c_int64, C.c_uint64, c_i64_p, c_i64_p, C.c_void_p]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    """The loaded shared library.  Raises LibraryMissing when it has not been
    built -- the product has no other implementation to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -m turbopy_b200.build` "
                "(needs nvcc); turbopy_b200 has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI drifted
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def error_string(code):
    return lib().tp_error_string(int(code)).decode()


def check(code):
    """Map the C ABI's int return to Python exceptions (Errors)."""
    if code != TP_OK:
        raise RuntimeError(f"turbopy_b200 call failed ({code}): {error_string(code)}")


def device_info():
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    l2, smem = C.c_int64(), C.c_int64()
    check(lib().tp_device_info(C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(l2), C.byref(smem)))
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "l2_bytes": l2.value,
            "smem_optin_bytes": smem.value}


def launch_count():
    return int(lib().tp_launch_count())

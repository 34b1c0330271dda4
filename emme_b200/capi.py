"""ctypes binding of the C ABI in include/emme_b200.h (libemme_b200.so).

The library is the product; this module only declares prototypes.  Loading fails loudly if
the shared object is missing -- there is no Python or CPU fallback for the compute path.
"""
import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

# EMME_B200_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = Path(os.environ.get("EMME_B200_LIB", PKG / "lib" / "libemme_b200.so"))

E_NO_DEVICE, E_CUDA, E_BAD_ORDER, E_STATE, E_INPUT, E_NONFINITE, E_PEER = 1000, 1001, 1002, 1003, 1004, 1005, 1006
PEER_BUFS = 6


class EmmeParams(C.Structure):
    """struct emme_params"""
    _fields_ = [(n, C.c_double) for n in
                ("q", "R", "vt", "tau", "beta_e", "eta_i", "eta_e", "omega_s_i", "omega_s_e",
                 "omega_d_bar", "arc_coeff", "integration_precision", "integration_accuracy")] + [
        ("integration_iteration_limit", C.c_int), ("integration_start_points", C.c_int),
        ("dx", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EmmeStats(C.Structure):
    """struct emme_stats"""
    _fields_ = [(n, C.c_ulonglong) for n in
                ("integrals", "panels", "evals", "fwd_trips", "bwd_trips", "max_stack")] + [
        ("assemble_ms", C.c_double), ("dense_ms", C.c_double), ("launches", C.c_ulonglong), ("pivot_fallbacks", C.c_ulonglong),
        ("sym_steps", C.c_ulonglong), ("dense_flops", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EmmePicParams(C.Structure):
    """struct emme_pic_params"""
    _fields_ = [(n, C.c_double) for n in
                ("q", "R", "vt", "tau", "shat", "b_theta", "length", "eta_i", "omega_s_i", "omega_d_bar",
                 "water_bag_weight_vpara", "water_bag_weight_vperp")] + [
        ("npoints", C.c_int), ("drift_center_transformation_switch", C.c_int)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/emme_b200.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
PROTOTYPES = {
    "emme_device_count": (C.c_int, []),
    "emme_last_error": (C.c_char_p, []),
    "emme_version": (C.c_char_p, []),
    "emme_create": (C.c_int, [C.POINTER(EmmeParams), C.c_int, _dp, _dp, _dp, C.c_int, C.POINTER(_vp)]),
    "emme_destroy": (C.c_int, [_vp]),
    "emme_dim": (C.c_int, [_vp]),
    "emme_set_tables": (C.c_int, [_vp, _vp, _vp, _vp]),
    "emme_set_params": (C.c_int, [_vp, C.POINTER(EmmeParams)]),
    "emme_fp64_peak": (C.c_int, [C.c_int, _dp, _dp]),
    "emme_assemble": (C.c_int, [_vp, C.c_double, C.c_double, _vp]),
    "emme_assemble_device": (C.c_int, [_vp, C.c_double, C.c_double, _vp, C.c_int, C.c_int]),
    "emme_seed": (C.c_int, [_vp, C.c_double, C.c_double]),
    "emme_newton_trace_step": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "emme_get_eigen_value": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "emme_trace_delta": (C.c_int, [_vp, _vp, _vp, _dp, _dp]),
    "emme_newton_qr_step": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "emme_copy_matrix_async": (C.c_int, [_vp, C.c_int, _vp]),
    "emme_copy_wait": (C.c_int, [_vp]),
    "emme_qr_delta": (C.c_int, [_vp, _vp, _vp, _dp, _dp]),
    "emme_qr_step_begin": (C.c_int, [_vp]),
    "emme_shard_config": (C.c_int, [_vp, C.c_int, C.c_int]),
    "emme_seed_begin": (C.c_int, [_vp, C.c_double, C.c_double]),
    "emme_seed_middle": (C.c_int, [_vp]),
    "emme_seed_finish": (C.c_int, [_vp]),
    "emme_step_begin": (C.c_int, [_vp]),
    "emme_step_finish": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "emme_matrix_device_ptr": (_vp, [_vp, C.c_int]),
    "emme_ipc_export": (C.c_int, [_vp, C.c_int, _vp]),
    "emme_ipc_import": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp]),
    "emme_peer_attach": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "emme_shard_dense": (C.c_int, [_vp, C.c_int]),
    "emme_peer_set_timeout": (C.c_int, [C.c_double]),
    "emme_copy_matrix": (C.c_int, [_vp, C.c_int, _vp]),
    "emme_null_space": (C.c_int, [_vp, _vp]),
    "emme_get_stats": (C.c_int, [_vp, C.POINTER(EmmeStats)]),
    "emme_stream": (_vp, [_vp]),
    "emme_synchronize": (C.c_int, [_vp]),
    "emme_input_load": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "emme_input_parse": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "emme_input_free": (None, [_vp]),
    "emme_input_set_number": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "emme_input_get_number": (C.c_int, [_vp, C.c_char_p, _dp]),
    "emme_input_get_string": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_int]),
    "emme_input_params": (C.c_int, [_vp, C.POINTER(EmmeParams), C.POINTER(C.c_int)]),
    "emme_input_tables": (C.c_int, [_vp, _dp, _dp, _dp]),
    # row N4: PIC method
    "emme_pic_load_markers": (C.c_int, [C.POINTER(EmmePicParams), C.c_long, C.c_longlong, _dp, _dp, _dp, _dp]),
    "emme_pic_create": (C.c_int, [C.POINTER(EmmePicParams), C.c_long, _dp, _dp, _dp, _dp, C.c_int, C.POINTER(_vp)]),
    "emme_pic_create_shard": (C.c_int, [C.POINTER(EmmePicParams), C.c_long, _dp, _dp, _dp, _dp, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(_vp)]),
    "emme_pic_pweight_sum": (C.c_int, [C.POINTER(EmmePicParams), C.c_long, _dp, _dp, _dp]),
    "emme_pic_create_block": (C.c_int, [C.POINTER(EmmePicParams), C.c_long, C.c_long, C.c_long, _dp, _dp, _dp, _dp,
                                        C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "emme_pic_ipc_export": (C.c_int, [_vp, _vp]),
    "emme_pic_ipc_import": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "emme_pic_peer_attach": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "emme_pic_destroy": (C.c_int, [_vp]),
    "emme_pic_step": (C.c_int, [_vp, C.c_double, C.c_int]),
    "emme_pic_steps_done": (C.c_long, [_vp]),
    "emme_pic_marker_num": (C.c_long, [_vp]),
    "emme_pic_current_field": (C.c_int, [_vp, _vp]),
    "emme_pic_field_history": (C.c_int, [_vp, C.c_long, C.c_long, _vp]),
    "emme_pic_markers": (C.c_int, [_vp, _dp, _dp]),
    "emme_pic_extras": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "emme_pic_field_stats": (C.c_int, [_vp, C.c_long, C.c_long, _dp]),
    "emme_pic_calculate_omega": (C.c_int, [_dp, C.c_long, C.c_double, _dp, _dp]),
    "emme_pic_get_timing": (C.c_int, [_vp, _dp, C.POINTER(C.c_ulonglong)]),
    "emme_pic_stream": (_vp, [_vp]),
    "emme_pic_stage_begin": (C.c_int, [_vp, C.c_double, C.c_int]),
    "emme_pic_stage_finish": (C.c_int, [_vp, C.c_int]),
    "emme_pic_density_ptr": (_vp, [_vp]),
    "emme_input_pic_params": (C.c_int, [_vp, C.POINTER(EmmePicParams), C.POINTER(C.c_long), C.POINTER(C.c_long), _dp]),
}

_lib = None


def load():
    """Load libemme_b200.so (building nothing: see emme_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m emme_b200.build` "
                "(emme_b200 has no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError = a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class EmmeError(RuntimeError):
    """Raised for a non-zero status of the C ABI; .code is the LAPACK-style info."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def check(code):
    if code != 0:
        msg = load().emme_last_error()
        raise EmmeError(code, (msg or b"").decode() or f"emme_b200 status {code}")

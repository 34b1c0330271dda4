"""Host-side mirror of the reference's PIC method (row N4) over the C ABI.

Names follow include/solver_pic.h and src/main.cpp so that tests read like the reference:
  PIC_State       ~ PIC_State<double> (include/solver_pic.h:16-404): markers, field,
                    current_field(), marker_num()
  Integrator      ~ Integrator<PIC_State> (include/solver_pic.h:406-471): step(dt)
  calculate_omega ~ util::calculate_omega (include/solver_pic.h:475-529)
  solve_once_pic  ~ src/main.cpp:82-137 (time loop, per-step diagnostics, eigenvalue)
The compute path is the CUDA library; nothing here has a CPU fallback.
"""
import ctypes as C

import numpy as np

from . import capi


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def pic_params(inp):
    """(EmmePicParams, marker_per_cell, step_number, time_step) of an emme_b200.Input."""
    p = capi.EmmePicParams()
    mpc, steps, dt = C.c_long(), C.c_long(), C.c_double()
    capi.check(inp._lib.emme_input_pic_params(inp._h, C.byref(p), C.byref(mpc), C.byref(steps), C.byref(dt)))
    return p, mpc.value, steps.value, dt.value


def load_markers(params, n, seed=-1):
    """PIC_State::initialize_marker (include/solver_pic.h:186-205): eta, v_para, v_perp, weight.
    seed < 0 seeds from std::random_device like the reference."""
    eta, v_para, v_perp = (np.empty(n, dtype=np.float64) for _ in range(3))
    weight = np.empty(n, dtype=np.complex128)
    capi.check(capi.load().emme_pic_load_markers(C.byref(params), n, seed, _dp(eta), _dp(v_para), _dp(v_perp),
                                                 _dp(weight.view(np.float64))))
    return eta, v_para, v_perp, weight


def pweight_sum(params, v_para, v_perp):
    """Sum of the un-normalised p_weight (include/solver_pic.h:229-232) over the given markers."""
    v_para, v_perp = (np.ascontiguousarray(a, dtype=np.float64) for a in (v_para, v_perp))
    out = C.c_double()
    capi.check(capi.load().emme_pic_pweight_sum(C.byref(params), v_para.shape[0], _dp(v_para), _dp(v_perp), C.byref(out)))
    return out.value


def calculate_omega(stats, dt):
    """util::calculate_omega: stats is (steps, 3) = mean Re, mean Im, rms of the field per step."""
    stats = np.ascontiguousarray(stats, dtype=np.float64)
    re, im = C.c_double(), C.c_double()
    capi.check(capi.load().emme_pic_calculate_omega(_dp(stats), stats.shape[0], float(dt), C.byref(re), C.byref(im)))
    return complex(re.value, im.value)


class PIC_State:
    """Device-resident counterpart of PIC_State<double>.

    `PIC_State(params, marker_per_cell, seed=...)` draws the markers like the reference's
    constructor; `PIC_State.from_markers(params, eta, v_para, v_perp, weight)` takes given ones
    (parity tests).  shard=(index, count) keeps one contiguous block of the markers on this
    device (multi-GPU, see emme_b200.parallel.ShardedPIC)."""

    def __init__(self, params, marker_per_cell=None, seed=-1, device=0, markers=None, shard=(0, 1)):
        self._lib = capi.load()
        self.params = params
        if markers is None:
            markers = load_markers(params, int(marker_per_cell) * params.npoints, seed)
        eta, v_para, v_perp, weight = (np.ascontiguousarray(a) for a in markers)
        weight = weight.astype(np.complex128, copy=False)
        self.n_total = eta.shape[0]
        self._h = C.c_void_p()
        capi.check(self._lib.emme_pic_create_shard(C.byref(params), self.n_total, _dp(eta), _dp(v_para),
                                                    _dp(v_perp), _dp(weight.view(np.float64)), shard[0], shard[1],
                                                    device, C.byref(self._h)))
        self.nf = params.npoints

    @classmethod
    def from_markers(cls, params, eta, v_para, v_perp, weight, device=0, shard=(0, 1)):
        return cls(params, markers=(eta, v_para, v_perp, weight), device=device, shard=shard)

    @classmethod
    def from_block(cls, params, n_total, first, block, pw_sum, shard, device=0):
        """This rank's contiguous block [first, first + len) of n_total markers; pw_sum is the
        un-normalised p_weight summed over ALL markers (pweight_sum per block, added over ranks)."""
        self = cls.__new__(cls)
        self._lib = capi.load()
        self.params = params
        eta, v_para, v_perp, weight = (np.ascontiguousarray(a) for a in block)
        weight = weight.astype(np.complex128, copy=False)
        self.n_total = int(n_total)
        self._h = C.c_void_p()
        capi.check(self._lib.emme_pic_create_block(C.byref(params), self.n_total, int(first), eta.shape[0], _dp(eta),
                                                    _dp(v_para), _dp(v_perp), _dp(weight.view(np.float64)),
                                                    float(pw_sum), shard[0], shard[1], device, C.byref(self._h)))
        self.nf = params.npoints
        return self

    # peer mapping of the density exchange buffer (multi-GPU without a collective library)
    def ipc_export(self):
        buf = C.create_string_buffer(64)
        capi.check(self._lib.emme_pic_ipc_export(self._h, buf))
        return buf.raw

    def ipc_import(self, peer_rank, peer_count, handle):
        capi.check(self._lib.emme_pic_ipc_import(self._h, peer_rank, peer_count, C.create_string_buffer(handle, 64)))

    def peer_attach(self, peer_rank, peer_count, peer):
        capi.check(self._lib.emme_pic_peer_attach(self._h, peer_rank, peer_count, peer._h))

    @classmethod
    def from_input(cls, inp, seed=-1, device=0):
        p, mpc, _, _ = pic_params(inp)
        return cls(p, mpc, seed=seed, device=device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.emme_pic_destroy(self._h)
            self._h = None

    __del__ = close

    def marker_num(self):
        return self._lib.emme_pic_marker_num(self._h)

    def steps_done(self):
        return self._lib.emme_pic_steps_done(self._h)

    def step(self, dt, nsteps=1):
        capi.check(self._lib.emme_pic_step(self._h, float(dt), int(nsteps)))

    def current_field(self):
        f = np.empty(self.nf, dtype=np.complex128)
        capi.check(self._lib.emme_pic_current_field(self._h, f.ctypes.data_as(C.c_void_p)))
        return f

    def field_history(self, first=0, count=None):
        count = self.steps_done() - first if count is None else count
        f = np.empty((count, self.nf), dtype=np.complex128)
        capi.check(self._lib.emme_pic_field_history(self._h, first, count, f.ctypes.data_as(C.c_void_p)))
        return f

    def field_stats(self, first=0, count=None):
        count = self.steps_done() - first if count is None else count
        s = np.empty((count, 3), dtype=np.float64)
        capi.check(self._lib.emme_pic_field_stats(self._h, first, count, _dp(s)))
        return s

    def markers(self):
        """(eta, weight) of this handle's markers."""
        n = self.marker_num()
        eta = np.empty(n, dtype=np.float64)
        w = np.empty(n, dtype=np.complex128)
        capi.check(self._lib.emme_pic_markers(self._h, _dp(eta), _dp(w.view(np.float64))))
        return eta, w

    def extras(self):
        """(omega_dv, omega_st, p_weight, quasi_neutrality_coef)"""
        n = self.marker_num()
        a, b, c = (np.empty(n, dtype=np.float64) for _ in range(3))
        coef = np.empty(self.nf, dtype=np.float64)
        capi.check(self._lib.emme_pic_extras(self._h, _dp(a), _dp(b), _dp(c), _dp(coef)))
        return a, b, c, coef

    def timing(self):
        ms, launches = C.c_double(), C.c_ulonglong()
        capi.check(self._lib.emme_pic_get_timing(self._h, C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    # multi-GPU hooks
    def stage_begin(self, dt, stage):
        capi.check(self._lib.emme_pic_stage_begin(self._h, float(dt), stage))

    def stage_finish(self, stage):
        capi.check(self._lib.emme_pic_stage_finish(self._h, stage))

    def density_ptr(self):
        return self._lib.emme_pic_density_ptr(self._h)

    def stream(self):
        return self._lib.emme_pic_stream(self._h)


class Integrator:
    """Integrator<PIC_State> (include/solver_pic.h:406-471): the three-stage Runge-Kutta step."""
    order = 3

    def __init__(self, state):
        self.state = state

    def step(self, dt):
        self.state.step(dt, 1)


def solve_once_pic(inp, seed=-1, device=0, field_file=None):
    """src/main.cpp:82-137: step_number steps of time_step, the field of every step appended to
    field_file (eigenMatrics/*.bin), eigenvalue from util::calculate_omega, eigenvector = the last
    field."""
    import time
    p, mpc, nt, dt = pic_params(inp)
    t0 = time.perf_counter()
    markers = load_markers(p, int(mpc) * p.npoints, seed)
    t1 = time.perf_counter()
    state = PIC_State(p, markers=markers, device=device)
    t2 = time.perf_counter()
    state.step(dt, nt)
    t3 = time.perf_counter()
    if field_file is not None:
        state.field_history().tofile(field_file)
    stats = state.field_stats()
    omega = calculate_omega(stats, dt)
    result = {"eigenvalue": [omega.real, omega.imag], "eigenvector": state.current_field(), "stats": stats,
              "step_ms": state.timing()[0] / max(nt, 1), "markers": state.marker_num(),
              "timing": {"load_markers_s": t1 - t0, "create_upload_s": t2 - t1, "steps_s": t3 - t2,
                         "diagnostics_s": time.perf_counter() - t3}}
    state.close()
    return result

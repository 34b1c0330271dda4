"""Host-side mirror of the reference's eigen entry points over the C ABI.

Names and behaviour follow the reference so that tests read like the reference's own driver:
  Input            ~ util::json::parse_file + filter_input + Parameters::generate + Grid
                     (src/main.cpp:174-184, src/Parameters.cpp:10-66, include/Grid.h)
  EigenSolver      ~ EigenSolver<Matrix<std::complex<double>>> (include/solver.h:44-516):
                     constructor seeds at 0.99/1.00 omega0, matrixAssembler,
                     newtonTraceSecantIteration, fields eigen_value / d_eigen_value /
                     eigen_matrix
  solve_once_eigen ~ src/main.cpp:19-80 (iteration loop and stop rule)
  scan_values      ~ get_scan_generator (src/main.cpp:139-172)
"""
import ctypes as C
import math

import numpy as np

from . import capi


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Input:
    """A parsed input.json with scan objects collapsed to their head value."""

    def __init__(self, path=None, text=None):
        self._lib = capi.load()
        self._h = C.c_void_p()
        if path is not None:
            capi.check(self._lib.emme_input_load(str(path).encode(), C.byref(self._h)))
        else:
            capi.check(self._lib.emme_input_parse(text.encode(), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.emme_input_free(self._h)
            self._h = None

    def number(self, key):
        v = C.c_double()
        capi.check(self._lib.emme_input_get_number(self._h, key.encode(), C.byref(v)))
        return v.value

    def string(self, key):
        buf = C.create_string_buffer(256)
        capi.check(self._lib.emme_input_get_string(self._h, key.encode(), buf, 256))
        return buf.value.decode()

    def set_number(self, key, value):
        capi.check(self._lib.emme_input_set_number(self._h, key.encode(), float(value)))

    def params(self):
        p = capi.EmmeParams()
        n = C.c_int()
        capi.check(self._lib.emme_input_params(self._h, C.byref(p), C.byref(n)))
        return p, n.value

    def tables(self):
        """eta_i, g_integration_f(eta_i), bi(eta_i) as three float64 arrays."""
        _, n = self.params()
        eta, g, bi = (np.empty(n, dtype=np.float64) for _ in range(3))
        capi.check(self._lib.emme_input_tables(self._h, _dp(eta), _dp(g), _dp(bi)))
        return eta, g, bi

    def initial_guess(self):
        # "initial_guess": [re, im] is read by the driver (src/main.cpp:205-206); the C ABI
        # exposes numbers only, so the pair is fetched through the array helper keys
        return complex(self.number("initial_guess[0]"), self.number("initial_guess[1]"))


class EigenSolver:
    """Device-resident counterpart of the reference's EigenSolver.

    `EigenSolver(params, n, eta, g, bi)` only allocates; `seed(omega0)` has the reference
    constructor's semantics (include/solver.h:396-415).  `EigenSolver.from_input(inp, omega0)`
    does both, like `EigenSolver(para, omega0, coeff, grid)` in src/main.cpp:36-37.
    """

    def __init__(self, params, npoints, eta, g, bi, device=0):
        self._lib = capi.load()
        self._h = C.c_void_p()
        self.params = params
        self.npoints = npoints
        eta, g, bi = (np.ascontiguousarray(a, dtype=np.float64) for a in (eta, g, bi))
        capi.check(self._lib.emme_create(C.byref(params), npoints, _dp(eta), _dp(g), _dp(bi),
                                         device, C.byref(self._h)))
        self.dim = self._lib.emme_dim(self._h)
        self.eigen_value = 0j
        self.d_eigen_value = 0j

    @classmethod
    def from_input(cls, inp, omega0=None, device=0):
        p, n = inp.params()
        s = cls(p, n, *inp.tables(), device=device)
        if omega0 is not None:
            s.seed(omega0)
        return s

    def close(self):
        if getattr(self, "_h", None):
            self._lib.emme_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- EigenSolver::matrixAssembler ----
    def matrixAssembler(self, omega=None, out=None):
        """A(omega) as a (dim, dim) complex128 array (host)."""
        w = self.eigen_value if omega is None else complex(omega)
        if out is None:
            out = np.empty((self.dim, self.dim), dtype=np.complex128)
        capi.check(self._lib.emme_assemble(self._h, w.real, w.imag, out.ctypes.data))
        return out

    def assemble_device(self, omega, dev_ptr, shard_index=0, shard_count=1):
        w = complex(omega)
        capi.check(self._lib.emme_assemble_device(self._h, w.real, w.imag, C.c_void_p(dev_ptr),
                                                  shard_index, shard_count))

    # ---- constructor semantics + Newton step ----
    def _pull(self):
        v = [C.c_double() for _ in range(4)]
        capi.check(self._lib.emme_get_eigen_value(self._h, *[C.byref(x) for x in v]))
        self.eigen_value = complex(v[0].value, v[1].value)
        self.d_eigen_value = complex(v[2].value, v[3].value)

    def seed(self, omega0):
        w = complex(omega0)
        capi.check(self._lib.emme_seed(self._h, w.real, w.imag))
        self._pull()

    def newtonTraceSecantIteration(self):
        v = [C.c_double() for _ in range(4)]
        rc = self._lib.emme_newton_trace_step(self._h, *[C.byref(x) for x in v])
        self._pull()
        capi.check(rc)

    def newtonQRSecantIteration(self):
        """EigenSolver::newtonQRSecantIteration (include/solver.h:210-383)."""
        v = [C.c_double() for _ in range(4)]
        rc = self._lib.emme_newton_qr_step(self._h, *[C.byref(x) for x in v])
        self._pull()
        capi.check(rc)

    def qr_delta(self, A, Ad):
        """delta = -R_nn/(Q^H Ad v)_n of the QR-secant iterate for host matrices (dense step only)."""
        A = np.ascontiguousarray(A, dtype=np.complex128)
        Ad = np.ascontiguousarray(Ad, dtype=np.complex128)
        dr, di = C.c_double(), C.c_double()
        capi.check(self._lib.emme_qr_delta(self._h, A.ctypes.data, Ad.ctypes.data,
                                           C.byref(dr), C.byref(di)))
        return complex(dr.value, di.value)

    def trace_delta(self, A, Ad):
        """delta = -1/trace(A^-1 Ad) for host matrices (dense step only)."""
        A = np.ascontiguousarray(A, dtype=np.complex128)
        Ad = np.ascontiguousarray(Ad, dtype=np.complex128)
        dr, di = C.c_double(), C.c_double()
        capi.check(self._lib.emme_trace_delta(self._h, A.ctypes.data, Ad.ctypes.data,
                                              C.byref(dr), C.byref(di)))
        return complex(dr.value, di.value)

    def nullSpace(self):
        """Eigenvector of the current eigen_matrix (include/solver.h:58-112), unit 2-norm, largest
        component real positive."""
        v = np.empty(self.dim, dtype=np.complex128)
        capi.check(self._lib.emme_null_space(self._h, v.ctypes.data))
        return v

    def _matrix(self, which):
        out = np.empty((self.dim, self.dim), dtype=np.complex128)
        capi.check(self._lib.emme_copy_matrix(self._h, which, out.ctypes.data))
        return out

    @property
    def eigen_matrix(self):
        return self._matrix(0)

    @property
    def eigen_matrix_old(self):
        return self._matrix(1)

    @property
    def eigen_matrix_derivative(self):
        return self._matrix(2)

    def stats(self):
        st = capi.EmmeStats()
        capi.check(self._lib.emme_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    # ---- multi-GPU building blocks (see emme_b200/parallel.py) ----
    def shard_config(self, index, count):
        capi.check(self._lib.emme_shard_config(self._h, index, count))

    def matrix_device_ptr(self, which=0):
        return self._lib.emme_matrix_device_ptr(self._h, which)

    def stream(self):
        return self._lib.emme_stream(self._h)

    def synchronize(self):
        capi.check(self._lib.emme_synchronize(self._h))


def solve_once_eigen(inp, omega0, device=0, on_iterate=None, solver=None):
    """The loop of solve_once_eigen (src/main.cpp:19-57): seed, then at most
    iteration_step_limit+1 iterates -- newtonTraceSecantIteration when iteration_method is
    "TraceSecant", newtonQRSecantIteration otherwise (src/main.cpp:45-49) -- stopping when
    |delta| < tol*|omega|.  Returns (omega, iterates, solver)."""
    tol = inp.number("iteration_precision")
    limit = int(inp.number("iteration_step_limit"))
    method = inp.string("iteration_method")
    s = solver or EigenSolver.from_input(inp, device=device)
    s.seed(omega0)
    iterates = []
    for _ in range(limit + 1):
        if method == "TraceSecant":
            s.newtonTraceSecantIteration()
        else:
            s.newtonQRSecantIteration()
        iterates.append((s.eigen_value, s.d_eigen_value))
        if on_iterate:
            on_iterate(s)
        if abs(s.d_eigen_value) < abs(tol * s.eigen_value):
            break
    return s.eigen_value, iterates, s


def scan_values(head, step, tail):
    """The value sequence of get_scan_generator (src/main.cpp:139-172).

    `tail` is a number or [left_tail, right_tail]; for a scalar tail the other tail is
    head + 0.5*copysign(step, head - tail) (src/main.cpp:236-239).  Returns a list of
    (value, turning) pairs in visiting order."""
    if isinstance(tail, (list, tuple)):
        left_tail, right_tail = float(tail[0]), float(tail[1])
    else:
        left_tail = float(tail)
        right_tail = head + .5 * math.copysign(step, head - tail)
    out = []
    current, current_tail = head, left_tail
    to_left, is_first = True, True

    def within():
        return abs(current - head) <= (abs(current_tail - head) + 0.01 * abs(step))

    while True:
        if not is_first:
            current += math.copysign(step, current_tail - head)
        is_first = False
        if within():
            out.append((current, False))
            continue
        to_left = not to_left
        current_tail = right_tail
        current = head + math.copysign(step, current_tail - head)
        if (not to_left) and within():
            out.append((current, True))
            continue
        break
    return out

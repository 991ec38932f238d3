"""ctypes binding of the C ABI declared in include/xnode_wan_b200.h.

The product library is `libxnode_wan_b200.so` next to this file, built by build.py with
`nvcc -gencode arch=compute_100a,code=sm_100a`.  There is NO fallback: if the library is missing
or cannot be loaded, `get()` raises.  (tests/ may construct `XwLib(path)` on another build of the
same ABI -- the CPU emulation used for kernel-logic tests -- but nothing in this package does.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libxnode_wan_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)

SOLVERS = {"euler": 0, "midpoint": 1, "rk4": 2}
DOMAINS = {"cube": 0, "cone": 1, "hourglass": 2}
NSUMS = 8
ABI_VERSION = 4
SUM_S1, SUM_S2, SUM_S3, SUM_VV, SUM_INIT, SUM_BDRY = 0, 1, 2, 3, 4, 5


class Dims(C.Structure):
    _fields_ = [("d", C.c_int), ("H", C.c_int), ("hh", C.c_int), ("nu", C.c_int), ("Hv", C.c_int),
                ("nv", C.c_int), ("solver", C.c_int)]


class Domain(C.Structure):
    _fields_ = [("kind", C.c_int), ("p0", C.c_float), ("p1", C.c_float), ("p2", C.c_float)]


class Coef(C.Structure):
    _fields_ = [("c0", C.c_float), ("c1", C.c_float), ("a", C.c_void_p), ("b", C.c_void_p),
                ("a_sn", C.c_longlong), ("b_sn", C.c_longlong), ("A_val", C.c_void_p), ("A_der", C.c_void_p)]


class Points(C.Structure):
    _fields_ = [("t", C.c_void_p), ("t_sn", C.c_longlong), ("t_sl", C.c_longlong),
                ("x", C.c_void_p), ("x_sn", C.c_longlong), ("x_sl", C.c_longlong)]


class XwError(RuntimeError):
    pass


_P = C.c_void_p
_SIGS = {
    "xw_abi_version": (C.c_int, []),
    "xw_last_error": (C.c_char_p, []),
    "xw_theta_u_size": (C.c_int, [C.POINTER(Dims)]),
    "xw_theta_v_size": (C.c_int, [C.POINTER(Dims)]),
    "xw_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "xw_xnode_eval": (C.c_int, [C.POINTER(Dims), _P, _P, C.c_longlong, _P, C.c_int, _P, C.c_int, _P, _P]),
    "xw_vnet_eval": (C.c_int, [C.POINTER(Dims), _P, C.POINTER(Points), C.c_int, C.c_int, _P, _P]),
    "xw_interior_forward": (C.c_int, [C.POINTER(Dims), C.POINTER(Domain), C.POINTER(Coef), _P, _P, _P, C.c_longlong,
                                      _P, C.c_int, C.POINTER(Points), _P, _P, _P, C.c_int, _P, _P, _P, _P, _P,
                                      C.c_size_t, _P, _P, _P, C.c_int, _P, C.c_size_t, C.c_size_t]),
    "xw_yhist_floats": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "xw_last_xnode_impl": (C.c_int, []),
    "xw_last_vnet_impl": (C.c_int, []),
    "xw_vcache_floats": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "xw_boundary_u": (C.c_int, [C.POINTER(Dims), _P, _P, C.c_longlong, _P, C.c_int, _P, _P, C.c_int, C.c_double, _P,
                                _P, C.c_int, _P, C.c_size_t, _P]),
    "xw_interior_backward_u": (C.c_int, [C.POINTER(Dims), _P, _P, C.c_longlong, _P, C.c_int, _P, _P, C.c_int, _P, _P,
                                         C.c_int, _P, C.c_size_t, _P, _P, _P]),
    "xw_interior_backward_v": (C.c_int, [C.POINTER(Dims), C.POINTER(Domain), _P, C.POINTER(Points), _P, C.c_int,
                                         C.c_int, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    "xw_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    "xw_loss_scalars": (C.c_int, [_P, C.c_int] + [C.c_double] * 7 + [_P, _P]),
    "xw_fma_probe": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), _P]),
    "xw_umma_probe": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
}
EXPORTS = tuple(_SIGS)


class XwLib:
    """thin typed wrapper; every call raises XwError with the library's message on failure"""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise XwError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % path)
        self.path = path
        self.cdll = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:
                raise XwError("%s does not export %s: stale build, rebuild it (python __graft_entry__.py)" % (path, name))
            fn.restype = res
            fn.argtypes = args
        # host pointers are only meaningful to the CPU emulation build of the same ABI (tests/host_emu), which
        # identifies itself with an extra symbol; the product library (nvcc, sm_100a) never has it
        self.host_memory = hasattr(self.cdll, "xw_emu_marker")
        if self.cdll.xw_abi_version() != ABI_VERSION:
            raise XwError("ABI version mismatch in %s" % path)

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise XwError("%s: %s" % (name, self.cdll.xw_last_error().decode()))

    def theta_sizes(self, dims):
        return self.cdll.xw_theta_u_size(C.byref(dims)), self.cdll.xw_theta_v_size(C.byref(dims))

    def workspace_bytes(self, dims, n, L):
        b = self.cdll.xw_workspace_bytes(C.byref(dims), int(n), int(L))
        if b == 0:
            raise XwError("xw_workspace_bytes: %s" % self.cdll.xw_last_error().decode())
        return b


_LIB = None


def get():
    global _LIB
    if _LIB is None:
        _LIB = XwLib()
    return _LIB

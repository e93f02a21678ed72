"""ctypes wrapper of oracle/sesr_oracle.c (the fast C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class _Layer(C.Structure):
    _fields_ = [("k", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("relu", C.c_int32),
                ("z", C.c_int32), ("m16", C.c_int32), ("n", C.c_int32),
                ("w", C.c_void_p), ("K", C.c_void_p)]


class _Net(C.Structure):
    _fields_ = [("layers", _Layer * 5), ("s0", C.c_double), ("s5", C.c_double), ("z5", C.c_int32),
                ("mr", C.c_int32), ("nr", C.c_int32), ("r", C.c_int32), ("acc_bits", C.c_int32),
                ("add_bits", C.c_int32)]


def _has_avx2():
    try:
        return " avx2 " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        name = "libsesr_oracle_avx2.so" if _has_avx2() else "libsesr_oracle_generic.so"
        path = os.path.join(HERE, "_build", name)
        if not os.path.exists(path):
            raise RuntimeError("C oracle not built: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(path)
        _lib.sesr_oracle_forward.restype = C.c_int
        _lib.sesr_oracle_forward.argtypes = [C.POINTER(_Net), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.sesr_oracle_threads.restype = C.c_int
        _lib.sesr_oracle_set_threads.argtypes = [C.c_int]
        _lib.sesr_oracle_set_threads.restype = None
    return _lib


def set_threads(n=None):
    """Use `n` OpenMP threads (default: every core of this process's affinity mask), whatever OMP_NUM_THREADS says."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().sesr_oracle_set_threads(int(n))
    return threads()


def threads():
    return int(lib().sesr_oracle_threads())


class CNet:
    """Built from an oracle.sesr_oracle.OracleNet."""

    def __init__(self, onet):
        self.keep = []
        n = _Net()
        for i, L in enumerate(onet.layers):
            w = np.ascontiguousarray(L.W.astype(np.int8))
            K = np.ascontiguousarray(np.asarray(L.K, dtype=np.int32))
            self.keep += [w, K]
            l = n.layers[i]
            l.k, l.cin, l.cout, l.relu, l.z, l.m16, l.n = L.k, L.cin, L.cout, int(L.relu), L.z, L.M, L.n
            l.w, l.K = w.ctypes.data, K.ctypes.data
        n.s0, n.s5, n.z5 = onet.scales[0], onet.scales[5], onet.zeros[5]
        n.mr, n.nr, n.r = onet.Mr, onet.nr, onet.r
        n.acc_bits, n.add_bits = onet.pe_acc_bit, onet.pe_add_bit
        self.net = n
        self.cout = onet.layers[4].cout
        self.r = onet.r

    def forward(self, x, want_trace=False):
        x = np.ascontiguousarray(x, dtype=np.float32)
        B, _, H, W = x.shape
        q5 = np.empty((B, self.cout, H, W), dtype=np.int8)
        r = self.r
        y = np.empty((B, self.cout // (r * r), H * r, W * r), dtype=np.float32)
        tr = np.zeros((5, 16, H, W), dtype=np.int8) if want_trace else None
        rc = lib().sesr_oracle_forward(C.byref(self.net), x.ctypes.data, B, H, W, q5.ctypes.data, y.ctypes.data,
                                       tr.ctypes.data if want_trace else None)
        assert rc == 0
        return (q5, y, tr) if want_trace else (q5, y)

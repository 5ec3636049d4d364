"""ctypes binding of libmmdb200.so (include/mmdb200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible, the
compute entry points raise — they never route through a host implementation.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MMDB_LIB") or os.path.join(_HERE, "libmmdb200.so")      # MMDB_LIB: A/B builds of the same library

NCLASS_PAIR = 6
PAIR_CLASSES = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2)]  # index = la(la+1)/2 + lb


class MMDBError(RuntimeError):
    pass


class FockStats(C.Structure):
    _fields_ = [
        ("candidates", C.c_int64),
        ("quartets", C.c_int64),
        ("prim_quartets", C.c_int64),
        ("fn_quartets", C.c_int64),
        ("slow_quartets", C.c_int64),
        ("launches", C.c_int64),
        ("model_flops", C.c_double),
        ("class_quartets", C.c_int64 * (NCLASS_PAIR * NCLASS_PAIR)),
        ("class_prim_quartets", C.c_int64 * (NCLASS_PAIR * NCLASS_PAIR)),
        ("class_ms", C.c_float * (NCLASS_PAIR * NCLASS_PAIR)),
        ("class_screen_ms", C.c_float * (NCLASS_PAIR * NCLASS_PAIR)),
        ("far_entries", C.c_int64),
        ("near_entries", C.c_int64),
        ("exec_prim_quartets", C.c_int64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k in ("candidates", "quartets", "prim_quartets", "fn_quartets", "slow_quartets", "launches", "model_flops", "far_entries", "near_entries", "exec_prim_quartets")}
        names = ["ss", "ps", "pp", "ds", "dp", "dd"]
        per = {}
        for cb in range(NCLASS_PAIR):
            for ck in range(cb + 1):
                q = self.class_quartets[cb * NCLASS_PAIR + ck]
                if q:
                    per["(%s|%s)" % (names[cb], names[ck])] = {
                        "quartets": q,
                        "prim_quartets": self.class_prim_quartets[cb * NCLASS_PAIR + ck],
                        "ms": self.class_ms[cb * NCLASS_PAIR + ck],
                        "screen_ms": self.class_screen_ms[cb * NCLASS_PAIR + ck],
                        "flops_per_prim_quartet": class_flops(*(PAIR_CLASSES[cb] + PAIR_CLASSES[ck])),
                    }
        d["classes"] = per
        return d


# name -> (restype, argtypes); every symbol include/mmdb200.h declares
_vp = C.c_void_p
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)
SIGNATURES = {
    "mmdb_version": (C.c_int, []),
    "mmdb_last_error": (C.c_char_p, []),
    "mmdb_device_count": (C.c_int, [_ip]),
    "mmdb_basis_create": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.POINTER(_vp)]),
    "mmdb_basis_destroy": (C.c_int, [_vp]),
    "mmdb_basis_nbf": (C.c_int, [_vp, _ip]),
    "mmdb_basis_pair_counts": (C.c_int, [_vp, _vp, _vp]),
    "mmdb_basis_pair_shells": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "mmdb_eri_shell_quartets": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int, _vp]),
    "mmdb_schwarz": (C.c_int, [_vp, _vp, _vp]),
    "mmdb_set_schwarz_host": (C.c_int, [_vp, _vp]),
    "mmdb_eri_dense": (C.c_int, [_vp, _vp, _vp]),
    "mmdb_jk_incore": (C.c_int, [C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmdb_fock_direct": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(FockStats), _vp]),
    "mmdb_fixed_to_double": (C.c_int, [C.c_int, _vp, C.c_int64, _vp]),
    "mmdb_formPT_host": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp, C.POINTER(FockStats)]),
    "mmdb_c128_diff_split_host": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _ip]),
    "mmdb_c128_join_host": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "mmdb_comm_unique_id": (C.c_int, [_vp]),
    "mmdb_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "mmdb_allreduce_G": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp]),
    "mmdb_comm_destroy": (C.c_int, [_vp]),
    "mmdb_schwarz_host": (C.c_int, [_vp, _vp]),
    "mmdb_eri_dense_host": (C.c_int, [_vp, _vp]),
    "mmdb_onee_host": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmdb_gradient_host": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmdb_ao2mo_mp2": (C.c_int, [C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _dp, _vp]),
    "mmdb_boys_host": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp]),
    "mmdb_boys_class_host": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp]),
    "mmdb_fp64_peak": (C.c_int, [C.c_int, _dp, C.POINTER(C.c_float)]),
    "mmdb_class_flops": (C.c_double, [C.c_int, C.c_int, C.c_int, C.c_int]),
}

_lib = None


def load():
    """Load the shared library (no GPU needed for loading / symbol checks)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise MMDBError("libmmdb200.so not built (%s); run `python __graft_entry__.py` / make -C csrc. "
                            "There is no CPU fallback." % SO_PATH)
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise MMDBError("libmmdb200 error %d: %s" % (rc, load().mmdb_last_error().decode()))


def require_gpu():
    lib = load()
    n = C.c_int(0)
    rc = lib.mmdb_device_count(C.byref(n))
    if rc != 0 or n.value == 0:
        raise MMDBError("no CUDA device visible — the two-electron path is GPU-only (no CPU fallback)")
    return n.value


def ptr(a):
    """Raw pointer of a numpy array (host) or torch tensor (device) or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous
        return a.ctypes.data
    return a.data_ptr()


def class_flops(la, lb, lc, ld):
    return load().mmdb_class_flops(la, lb, lc, ld)

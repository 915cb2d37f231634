"""ctypes binding of libmedplib_b200.so (C ABI: include/medplib_b200.h). Fails loudly when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmedplib_b200.so")

MPL_OK = 0
_ERR = {-1: "MPL_ERR_ARG", -2: "MPL_ERR_ALIGN", -3: "MPL_ERR_DRIVER", -4: "MPL_ERR_CUDA", -5: "MPL_ERR_UNSUPPORTED"}

ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU, ACT_SILU = range(5)
DT_BF16, DT_F32 = 0, 1


class MplError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("lda", ctypes.c_longlong),
        ("B", ctypes.c_void_p), ("B2", ctypes.c_void_p), ("ldb", ctypes.c_longlong),
        ("C", ctypes.c_void_p), ("ldc", ctypes.c_longlong),
        ("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("ldr", ctypes.c_longlong),
        ("row_scale", ctypes.c_void_p), ("m_dev", ctypes.c_void_p),
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("act", ctypes.c_int), ("out_dtype", ctypes.c_int), ("tile_n", ctypes.c_int),
    ]


_lib = None


def load():
    """Load the shared library once. Raises MplError if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MplError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C medplib_b200/csrc`. medplib_b200 has no CPU / eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mpl_version.restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != MPL_OK:
        raise MplError(f"{what} failed: {_ERR.get(rc, rc)}")

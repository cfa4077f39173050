"""ctypes binding of the C ABI declared in include/tnet_b200.h (libtnetb200.so).

Used by tests/ and bench.py to call the product exactly the way a foreign-language host would: plain
pointers and sizes.  There is no CPU fallback: when the library (or a GPU) is missing, everything raises.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
# TNB_LIB_DIR: developer override used by tools/dbg_*.py to load the tracing build (make TRACE=1 LIBDIR=...)
LIB_PATH = os.path.join(os.environ.get("TNB_LIB_DIR") or os.path.join(ROOT, "nnet-asr_b200", "lib"), "libtnetb200.so")

OK = 0
MATH_3XTF32, MATH_TF32, MATH_FP32_SIMT, MATH_BF16 = 0, 1, 2, 3
ACT_NONE, ACT_SIGMOID = 0, 1
H2D, D2H, D2D = 0, 1, 2


class MatrixDim(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("stride", C.c_int)]


class ObjStats(C.Structure):
    _fields_ = [("error", C.c_double), ("frames", C.c_longlong), ("correct", C.c_longlong)]


MAX_PEERS = 16


class PeerJob(C.Structure):
    """TnbPeerJob of include/tnet_b200.h (one layer's update over peer memory)."""
    _fields_ = [("G", C.c_void_p * MAX_PEERS), ("W", C.c_void_p * MAX_PEERS), ("corrW", C.c_void_p), ("bias", C.c_void_p),
                ("corrb", C.c_void_p), ("dW", MatrixDim), ("rows_pad", C.c_int), ("lr", C.c_float), ("mmt", C.c_float),
                ("wc", C.c_float), ("grad_div_frm", C.c_int), ("n_frames", C.c_int), ("pushed", C.c_int)]


EPI_STORE, EPI_FWD, EPI_DX, EPI_UPDATE = 0, 1, 2, 3


class GemmJob(C.Structure):
    """TnbGemmJob of include/tnet_b200.h (one GEMM + fused epilogue of a tnb_gemm_batch launch)."""
    _fields_ = [("transa", C.c_int), ("transb", C.c_int), ("m", C.c_int), ("n", C.c_int), ("k", C.c_int),
                ("A", C.c_void_p), ("lda", C.c_int), ("B", C.c_void_p), ("ldb", C.c_int),
                ("A16", C.c_void_p), ("lda16", C.c_int), ("B16", C.c_void_p), ("ldb16", C.c_int),
                ("epilogue", C.c_int), ("alpha", C.c_float), ("beta", C.c_float),
                ("C", C.c_void_p), ("ldc", C.c_int), ("bias", C.c_void_p), ("act", C.c_int),
                ("mulY", C.c_void_p), ("ldy", C.c_int), ("W", C.c_void_p), ("ldw", C.c_int), ("w_scale", C.c_float), ("w_l2", C.c_float),
                ("C16", C.c_void_p), ("ldc16", C.c_int), ("W16", C.c_void_p), ("ldw16", C.c_int),
                ("tile_first", C.c_int), ("tile_count", C.c_int)]


class TnbError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TnbError("libtnetb200.so is not built (%s): run __graft_entry__.build()" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.tnb_version.restype = C.c_char_p
        _lib.tnb_last_error.restype = C.c_char_p
    return _lib


def check(rc):
    if rc != OK:
        raise TnbError("tnb error %d: %s" % (rc, lib().tnb_last_error().decode()))


def declared_symbols():
    """All function names declared in include/tnet_b200.h."""
    import re
    txt = open(os.path.join(ROOT, "include", "tnet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tnb_[a-z0-9_]+)\s*\(", txt)))


class Context:
    def __init__(self, device=0, math=MATH_3XTF32):
        self.h = C.c_void_p()
        check(lib().tnb_ctx_create(C.byref(self.h), C.c_int(device)))
        self.set_math(math)

    def set_math(self, m):
        check(lib().tnb_ctx_set_math(self.h, C.c_int(m)))

    def sync(self):
        check(lib().tnb_ctx_sync(self.h))

    def stream(self):
        s = C.c_void_p()
        check(lib().tnb_ctx_stream(self.h, C.byref(s)))
        return s.value

    def launches(self):
        n = C.c_ulonglong()
        check(lib().tnb_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            lib().tnb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DMat:
    """Pitched fp32 (or int32/uint32) device matrix owned through tnb_malloc_pitch / tnb_free."""

    def __init__(self, ctx, rows, cols, dtype=np.float32):
        self.ctx, self.rows, self.cols, self.dtype = ctx, int(rows), int(cols), np.dtype(dtype)
        assert self.dtype.itemsize == 4
        self.ptr = C.c_void_p()
        st = C.c_int()
        check(lib().tnb_malloc_pitch(ctx.h, C.byref(self.ptr), C.byref(st), C.c_int(self.rows), C.c_int(self.cols)))
        self.stride = st.value

    @classmethod
    def from_numpy(cls, ctx, a):
        a = np.ascontiguousarray(a)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        m = cls(ctx, a.shape[0], a.shape[1], a.dtype)
        m.upload(a)
        return m

    @property
    def dim(self):
        return MatrixDim(self.rows, self.cols, self.stride)

    def row_view_dim(self, rows):
        return MatrixDim(rows, self.cols, self.stride)

    def p(self, ctype=C.c_float):
        return C.cast(self.ptr, C.POINTER(ctype))

    def row_ptr(self, r, ctype=C.c_float):
        return C.cast(C.c_void_p(self.ptr.value + 4 * r * self.stride), C.POINTER(ctype))

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype).reshape(self.rows, self.cols)
        if self.rows == 0 or self.cols == 0:
            return
        check(lib().tnb_memcpy2d(self.ctx.h, self.ptr, C.c_size_t(self.stride * 4), a.ctypes.data_as(C.c_void_p),
                                 C.c_size_t(self.cols * 4), C.c_size_t(self.cols * 4), C.c_size_t(self.rows), C.c_int(H2D)))
        self.ctx.sync()  # pageable source

    def download(self):
        a = np.empty((self.rows, self.cols), dtype=self.dtype)
        if self.rows == 0 or self.cols == 0:
            return a
        check(lib().tnb_memcpy2d(self.ctx.h, a.ctypes.data_as(C.c_void_p), C.c_size_t(self.cols * 4), self.ptr,
                                 C.c_size_t(self.stride * 4), C.c_size_t(self.cols * 4), C.c_size_t(self.rows), C.c_int(D2H)))
        return a

    def free(self):
        if self.ptr:
            lib().tnb_free(self.ctx.h, self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def bf16_round(a):
    """fp32 -> nearest-even bf16, returned as fp32 (what __float2bfloat16_rn keeps); finite inputs."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


class DMat16:
    """Pitched bf16 device matrix (tnb_malloc_pitch16): the twin a TNB_MATH_BF16 GEMM reads instead of the fp32 array."""

    def __init__(self, ctx, rows, cols):
        self.ctx, self.rows, self.cols = ctx, int(rows), int(cols)
        self.ptr = C.c_void_p()
        st = C.c_int()
        check(lib().tnb_malloc_pitch16(ctx.h, C.byref(self.ptr), C.byref(st), C.c_int(self.rows), C.c_int(self.cols)))
        self.stride = st.value

    @classmethod
    def from_fp32(cls, ctx, dmat):
        m = cls(ctx, dmat.rows, dmat.cols)
        check(lib().tnb_to_bf16(ctx.h, m.p(), C.c_int(m.stride), dmat.p(), dmat.dim))
        return m

    def p(self):
        return C.cast(self.ptr, C.POINTER(C.c_uint16))

    def download_bits(self, full_pitch=False):
        w = self.stride if full_pitch else self.cols
        a = np.empty((self.rows, w), dtype=np.uint16)
        check(lib().tnb_memcpy2d(self.ctx.h, a.ctypes.data_as(C.c_void_p), C.c_size_t(w * 2), self.ptr, C.c_size_t(self.stride * 2),
                                 C.c_size_t(w * 2), C.c_size_t(self.rows), C.c_int(D2H)))
        return a

    def download(self):
        """values as fp32"""
        return (self.download_bits().astype(np.uint32) << 16).view(np.float32)

    def free(self):
        if self.ptr:
            lib().tnb_free(self.ctx.h, self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


class DStats:
    """Device-resident TnbObjStats."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.ptr = C.c_void_p()
        check(lib().tnb_malloc(ctx.h, C.byref(self.ptr), C.c_size_t(C.sizeof(ObjStats))))

    def p(self):
        return C.cast(self.ptr, C.POINTER(ObjStats))

    def read(self):
        st = ObjStats()
        check(lib().tnb_memcpy(self.ctx.h, C.byref(st), self.ptr, C.c_size_t(C.sizeof(ObjStats)), C.c_int(D2H)))
        return st.error, st.frames, st.correct

    def reset(self):
        check(lib().tnb_memset(self.ctx.h, self.ptr, C.c_int(0), C.c_size_t(C.sizeof(ObjStats))))


def gemm(ctx, ta, tb, alpha, A, B, beta, Cm):
    """Cm (DMat) = alpha*op(A)*op(B) + beta*Cm with CuMatrix::Gemm argument meaning."""
    m, n = Cm.rows, Cm.cols
    k = A.rows if ta in "Tt" else A.cols
    check(lib().tnb_gemm(ctx.h, C.c_char(ta.encode()), C.c_char(tb.encode()), C.c_int(m), C.c_int(n), C.c_int(k),
                         C.c_float(alpha), A.p(), C.c_int(A.stride), B.p(), C.c_int(B.stride), C.c_float(beta), Cm.p(),
                         C.c_int(Cm.stride)))

"""ctypes binding of oracle/_build/libtnet_oracle.so (TEST INFRASTRUCTURE: the checker only).

Builds the library on first use if it is missing (gcc is available here and on the GPU box).
"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import formats as F  # noqa: E402  (file-format helpers only: layer tuple shapes)
_SO = os.path.join(ROOT, "oracle", "_build", "libtnet_oracle.so")

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
up = C.POINTER(C.c_uint)


def _load():
    src = os.path.join(ROOT, "oracle", "tnet_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build.sh")])
    lib = C.CDLL(_SO)
    lib.orc_net_new.restype = C.c_void_p
    lib.orc_rbm_new.restype = C.c_void_p
    lib.orc_rnn_new.restype = C.c_void_p
    lib.orc_cache_new.restype = C.c_void_p
    lib.orc_cache_new.argtypes = [C.c_size_t, C.c_size_t]
    lib.orc_net_layer_out.restype = fp
    lib.orc_net_layer_eout.restype = fp
    lib.orc_net_err.restype = fp
    lib.orc_cache_last_perm.restype = ip
    lib.orc_lrand48.restype = C.c_long
    lib.orc_srand48.argtypes = [C.c_long]
    return lib


lib = _load()


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def P(a):
    """float* / int* / unsigned* of a C-contiguous numpy array."""
    if a.dtype == np.float32:
        return a.ctypes.data_as(fp)
    if a.dtype == np.int32:
        return a.ctypes.data_as(ip)
    if a.dtype == np.uint32:
        return a.ctypes.data_as(up)
    raise TypeError(a.dtype)


cf = C.c_float
ci = C.c_int


# ---- kernel level ---------------------------------------------------------
def sigmoid(x):
    x = f32(x); y = np.empty_like(x)
    lib.orc_sigmoid(P(y), P(x), ci(x.shape[0]), ci(x.shape[1]), ci(x.shape[1]))
    return y


def diff_sigmoid(e, y):
    e = f32(e); y = f32(y); o = np.empty_like(e)
    lib.orc_diff_sigmoid(P(o), P(e), P(y), ci(e.shape[0]), ci(e.shape[1]), ci(e.shape[1]))
    return o


def softmax(x):
    x = f32(x); y = np.empty_like(x)
    lib.orc_softmax(P(y), P(x), ci(x.shape[0]), ci(x.shape[1]), ci(x.shape[1]))
    return y


def check_class(out, des):
    out = f32(out); des = f32(des); m = np.empty(out.shape[0], dtype=np.int32)
    lib.orc_check_class(P(out), P(des), P(m), ci(out.shape[0]), ci(out.shape[1]), ci(out.shape[1]))
    return m


def add_col_sum(alpha, mat, beta, vec):
    mat = f32(mat); v = f32(vec).copy()
    lib.orc_add_col_sum(cf(alpha), P(mat), cf(beta), P(v), ci(mat.shape[0]), ci(mat.shape[1]), ci(mat.shape[1]))
    return v


def expand(x, offs):
    x = f32(x); offs = np.ascontiguousarray(offs, dtype=np.int32)
    y = np.empty((x.shape[0], x.shape[1] * len(offs)), dtype=np.float32)
    lib.orc_expand(P(y), P(x), P(offs), ci(x.shape[0]), ci(y.shape[1]), ci(y.shape[1]),
                   ci(x.shape[0]), ci(x.shape[1]), ci(x.shape[1]))
    return y


def rearrange(x, copy_from, cols_out=None):
    x = f32(x); cf_ = np.ascontiguousarray(copy_from, dtype=np.int32)
    y = np.empty((x.shape[0], len(cf_)), dtype=np.float32)
    lib.orc_rearrange(P(y), P(x), P(cf_), ci(x.shape[0]), ci(y.shape[1]), ci(y.shape[1]), ci(x.shape[1]), ci(x.shape[1]))
    return y


def randomize(x, perm):
    x = f32(x); perm = np.ascontiguousarray(perm, dtype=np.int32)
    y = np.zeros_like(x)
    lib.orc_randomize(P(y), P(x), P(perm), ci(len(perm)), ci(x.shape[1]), ci(x.shape[1]), ci(x.shape[1]))
    return y


def gemm(ta, tb, alpha, A, B, beta, Cm, acc_double=0):
    A = f32(A); B = f32(B); Cm = f32(Cm).copy()
    m, n = Cm.shape
    k = A.shape[0] if ta in "Tt" else A.shape[1]
    lib.orc_gemm(C.c_char(ta.encode()), C.c_char(tb.encode()), ci(m), ci(n), ci(k), cf(alpha), P(A), ci(A.shape[1]),
                 P(B), ci(B.shape[1]), cf(beta), P(Cm), ci(n), ci(acc_double))
    return Cm


def shuffle_perm(seed, n, pre_draws=0):
    """srand48(seed); consume pre_draws lrand48(); random_shuffle permutation of 0..n-1."""
    lib.orc_srand48(C.c_long(seed))
    for _ in range(pre_draws):
        lib.orc_lrand48()
    p = np.empty(n, dtype=np.int32)
    lib.orc_shuffle_perm(P(p), ci(n))
    return p


class ObjStats(C.Structure):
    _fields_ = [("error", C.c_double), ("frames", C.c_longlong), ("correct", C.c_longlong)]


def xent_evaluate(Y, T):
    Y = f32(Y); T = f32(T); E = np.empty_like(Y); st = ObjStats(0.0, 0, 0)
    lib.orc_xent_evaluate(P(Y), P(T), P(E), ci(Y.shape[0]), ci(Y.shape[1]), ci(Y.shape[1]), C.byref(st))
    return E, st.error, st.frames, st.correct


def mse_evaluate(Y, T):
    Y = f32(Y); T = f32(T); E = np.empty_like(Y); st = ObjStats(0.0, 0, 0)
    lib.orc_mse_evaluate(P(Y), P(T), P(E), ci(Y.shape[0]), ci(Y.shape[1]), ci(Y.shape[1]), C.byref(st))
    return E, st.error, st.frames


# ---- RNG -------------------------------------------------------------------
def rand_seed(seed, rows, cols):
    lib.orc_srand48(C.c_long(seed))
    z = [np.empty((rows, cols), dtype=np.uint32) for _ in range(4)]
    lib.orc_rand_seed(P(z[0]), P(z[1]), P(z[2]), P(z[3]), ci(rows), ci(cols), ci(cols))
    return z


def rand_uniform(z):
    out = np.empty(z[0].shape, dtype=np.float32)
    lib.orc_rand(P(out), P(z[0]), P(z[1]), P(z[2]), P(z[3]), ci(out.shape[0]), ci(out.shape[1]), ci(out.shape[1]))
    return out


def rand_gauss(z):
    out = np.empty(z[0].shape, dtype=np.float32)
    lib.orc_gauss_rand(P(out), P(z[0]), P(z[1]), P(z[2]), P(z[3]), ci(out.shape[0]), ci(out.shape[1]), ci(out.shape[1]))
    return out


def binarize(probs, rnd):
    probs = f32(probs); rnd = f32(rnd); s = np.empty_like(probs)
    lib.orc_binarize_probs(P(s), P(probs), P(rnd), ci(s.shape[0]), ci(s.shape[1]), ci(s.shape[1]))
    return s


# ---- network ---------------------------------------------------------------
class Net:
    """MLP trainer oracle (CuNetwork::Propagate/Backpropagate + CuCrossEntropy)."""

    def __init__(self, layers, acc_double=0):
        self.h = C.c_void_p(lib.orc_net_new(ci(acc_double)))
        self.layers = layers
        for L in layers:
            if L[0] == "affine":
                Wt, b = f32(L[1]), f32(L[2])
                lib.orc_net_add_affine(self.h, ci(Wt.shape[1]), ci(Wt.shape[0]), P(Wt), P(b))
            elif L[0] == "shared":
                K, Wt, b = int(L[1]), f32(L[2]), f32(L[3])
                lib.orc_net_add_shared(self.h, ci(Wt.shape[1] * K), ci(Wt.shape[0] * K), ci(K), P(Wt), P(b))
            elif L[0] == "discrete":
                blocks, b = [f32(B) for B in L[1]], f32(L[2])
                bin_ = np.array([B.shape[1] for B in blocks], np.int32)
                bout = np.array([B.shape[0] for B in blocks], np.int32)
                flat = np.concatenate([B.ravel() for B in blocks])
                lib.orc_net_add_discrete(self.h, ci(len(blocks)), P(bin_), P(bout), P(flat), P(b))
            elif L[0] == "sigmoid":
                lib.orc_net_add_sigmoid(self.h, ci(L[1]))
            elif L[0] == "softmax":
                lib.orc_net_add_softmax(self.h, ci(L[1]))
            else:
                raise ValueError(L[0])

    def set_hyper(self, lr, mmt=0.0, wc=0.0, gdf=True, factors=None):
        if factors is None:
            lib.orc_net_set_hyper(self.h, cf(lr), None, ci(0), cf(mmt), cf(wc), ci(int(gdf)))
        else:
            fa = f32(factors)
            lib.orc_net_set_hyper(self.h, cf(lr), P(fa), ci(len(fa)), cf(mmt), cf(wc), ci(int(gdf)))

    def propagate(self, X):
        X = f32(X)
        out = np.empty((X.shape[0], self._nout()), dtype=np.float32)
        lib.orc_net_propagate(self.h, P(X), ci(X.shape[0]), P(out))
        return out

    def _nout(self):
        return F.layer_dims(self.layers[-1])[1]

    def train_bunch(self, X, T, cv=False):
        X = f32(X); T = f32(T)
        lib.orc_net_train_bunch(self.h, P(X), P(T), ci(X.shape[0]), ci(int(cv)))

    def stats(self):
        e = C.c_double(); fr = C.c_longlong(); co = C.c_longlong()
        lib.orc_net_stats(self.h, C.byref(e), C.byref(fr), C.byref(co))
        return e.value, fr.value, co.value

    def get_affine(self, idx):
        L = self.layers[idx]
        Wt = np.empty_like(f32(L[1])); b = np.empty_like(f32(L[2]))
        lib.orc_net_get_affine(self.h, ci(idx), P(Wt), P(b))
        return Wt, b

    def get_layer(self, idx):
        """Current parameters of layer idx in the `layers` tuple form of tnet_b200.formats."""
        L = self.layers[idx]
        if L[0] == "affine":
            return ("affine",) + self.get_affine(idx)
        if L[0] == "shared":
            Wt = np.empty_like(f32(L[2])); b = np.empty_like(f32(L[3]))
            lib.orc_net_get_shared(self.h, ci(idx), P(Wt), P(b))
            return ("shared", L[1], Wt, b)
        if L[0] == "discrete":
            flat = np.empty(sum(B.size for B in L[1]), np.float32); b = np.empty_like(f32(L[2]))
            lib.orc_net_get_discrete(self.h, ci(idx), P(flat), P(b))
            blocks, pos = [], 0
            for B in L[1]:
                blocks.append(flat[pos:pos + B.size].reshape(B.shape).copy())
                pos += B.size
            return ("discrete", blocks, b)
        return L

    def layer_out(self, idx, rows):
        n = F.layer_dims(self.layers[idx])[1]
        p = lib.orc_net_layer_out(self.h, ci(idx))
        return np.ctypeslib.as_array(p, shape=(rows, n)).copy()

    def layer_eout(self, idx, rows):
        n = F.layer_dims(self.layers[idx])[0]
        p = lib.orc_net_layer_eout(self.h, ci(idx))
        return np.ctypeslib.as_array(p, shape=(rows, n)).copy()

    def err(self, rows):
        p = lib.orc_net_err(self.h)
        return np.ctypeslib.as_array(p, shape=(rows, self._nout())).copy()

    def __del__(self):
        try:
            lib.orc_net_free(self.h)
        except Exception:
            pass


class Cache:
    """CuCache state machine oracle (cuCache.cc)."""

    def __init__(self, cachesize, bunchsize):
        self.h = C.c_void_p(lib.orc_cache_new(cachesize, bunchsize))
        if not self.h.value:
            raise ValueError("Non divisible cachesize by bunchsize")
        self.bunch = bunchsize
        self.fdim = self.ddim = None

    def add(self, F, D):
        F = f32(F); D = f32(D)
        self.fdim, self.ddim = F.shape[1], D.shape[1]
        lib.orc_cache_add(self.h, P(F), P(D), ci(F.shape[0]), ci(F.shape[1]), ci(D.shape[1]))

    def full(self):
        return bool(lib.orc_cache_full(self.h))

    def empty(self):
        return bool(lib.orc_cache_empty(self.h))

    def discarded(self):
        return lib.orc_cache_discarded(self.h)

    def randomize(self):
        lib.orc_cache_randomize(self.h)
        n = ci()
        p = lib.orc_cache_last_perm(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def get_bunch(self):
        F = np.empty((self.bunch, self.fdim), dtype=np.float32)
        D = np.empty((self.bunch, self.ddim), dtype=np.float32)
        rc = lib.orc_cache_get_bunch(self.h, P(F), P(D))
        if rc != 0:
            raise RuntimeError("GetBunch on empty cache!!!")
        return F, D

    def __del__(self):
        try:
            lib.orc_cache_free(self.h)
        except Exception:
            pass


class Rbm:
    def __init__(self, Wt, vb, hb, vis_gauss, hid_gauss, lr, mmt, wc, acc_double=0):
        Wt = f32(Wt); vb = f32(vb); hb = f32(hb)
        self.nhid, self.nvis = Wt.shape
        self.h = C.c_void_p(lib.orc_rbm_new(ci(self.nvis), ci(self.nhid), ci(int(vis_gauss)), ci(int(hid_gauss)),
                                            P(Wt), P(vb), P(hb), cf(lr), cf(mmt), cf(wc), ci(acc_double)))

    def set_sparse(self, cost):
        """CuRbmSparse: sparsity penalty with the reference's constructor defaults (prior 1e-4, lambda 0.95)."""
        lib.orc_rbm_set_sparse(self.h, cf(cost))

    def cd1(self, pos_vis, z):
        pos_vis = f32(pos_vis); rows = pos_vis.shape[0]
        ph = np.empty((rows, self.nhid), np.float32); nh = np.empty_like(ph); rnd = np.empty_like(ph)
        nv = np.empty((rows, self.nvis), np.float32); err = np.empty_like(nv)
        lib.orc_rbm_cd1_bunch(self.h, P(pos_vis), ci(rows), P(z[0]), P(z[1]), P(z[2]), P(z[3]),
                              P(ph), P(nh), P(rnd), P(nv), P(err))
        return ph, nh, nv

    def get(self):
        Wt = np.empty((self.nhid, self.nvis), np.float32)
        vb = np.empty(self.nvis, np.float32); hb = np.empty(self.nhid, np.float32)
        lib.orc_rbm_get(self.h, P(Wt), P(vb), P(hb))
        return Wt, vb, hb

    def stats(self):
        e = C.c_double(); fr = C.c_longlong()
        lib.orc_rbm_stats(self.h, C.byref(e), C.byref(fr))
        return e.value, fr.value

    def __del__(self):
        try:
            lib.orc_rbm_free(self.h)
        except Exception:
            pass


class Rnn:
    def __init__(self, Wt, b, nin, bptt, lr, mmt=0.0, wc=0.0):
        Wt = f32(Wt); b = f32(b)
        self.nout = Wt.shape[0]; self.nin = nin
        self.h = C.c_void_p(lib.orc_rnn_new(ci(nin), ci(self.nout), ci(bptt), P(Wt), P(b), cf(lr), cf(mmt), cf(wc)))

    def clear(self):
        lib.orc_rnn_clear(self.h)

    def propagate(self, x):
        x = f32(x); y = np.empty(self.nout, np.float32)
        lib.orc_rnn_propagate(self.h, P(x), P(y))
        return y

    def update(self, e):
        e = f32(e)
        lib.orc_rnn_update(self.h, P(e))

    def get(self):
        Wt = np.empty((self.nout, self.nin + self.nout), np.float32); b = np.empty(self.nout, np.float32)
        lib.orc_rnn_get(self.h, P(Wt), P(b))
        return Wt, b

    def __del__(self):
        try:
            lib.orc_rnn_free(self.h)
        except Exception:
            pass

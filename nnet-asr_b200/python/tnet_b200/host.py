"""ctypes binding of include/tnet_b200_host.h (libtnetb200_host.so): the C handles over the C++ mirror of the
reference's CuNetwork / CuCache / CuRbm / CuRecurrent main loops.  Class shapes follow tests/oracle_lib.py so the
same replay code drives the oracle and the CUDA path."""
import ctypes as C
import os
import tempfile

import numpy as np

from . import abi
from . import formats as F

HOST_LIB_PATH = os.path.join(abi.ROOT, "nnet-asr_b200", "lib", "libtnetb200_host.so")
_h = None
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)


def hlib():
    global _h
    if _h is None:
        abi.lib()  # libtnetb200.so first (RTLD_GLOBAL not needed: the host lib links it by rpath)
        if not os.path.exists(HOST_LIB_PATH):
            raise abi.TnbError("libtnetb200_host.so is not built: run __graft_entry__.build()")
        _h = C.CDLL(HOST_LIB_PATH)
        _h.tnh_last_error.restype = C.c_char_p
        _h.tnh_srand48.argtypes = [C.c_long]
        _h.tnh_srand48.restype = None
    return _h


def hcheck(rc):
    if rc != 0:
        raise abi.TnbError("tnh error: " + hlib().tnh_last_error().decode())


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def P(a):
    return a.ctypes.data_as(fp if a.dtype == np.float32 else ip)


def select_gpu(dev):
    hcheck(hlib().tnh_select_gpu(C.c_int(dev)))


def set_math(mode):
    hcheck(hlib().tnh_set_math(C.c_int(mode)))


def sync():
    hcheck(hlib().tnh_sync())


def launches():
    n = C.c_ulonglong()
    hcheck(hlib().tnh_launch_count(C.byref(n)))
    return n.value


def ctx_handle():
    h = C.c_void_p()
    hcheck(hlib().tnh_ctx(C.byref(h)))
    return h


def srand48(seed):
    hlib().tnh_srand48(int(seed))


_layer_dims = F.layer_dims


class Net:
    """TNetCu's network + objective (text file in, text file out)."""

    def __init__(self, layers=None, path=None, objective=0, fusion=True, dims=None, seed=1):
        self._tmp = None
        if dims is not None:            # random init built in C++ (no text file): the bench's large configs
            self.h = C.c_void_p()
            d = np.ascontiguousarray(dims, dtype=np.int32)
            hcheck(hlib().tnh_net_new_mlp(C.byref(self.h), P(d), C.c_int(len(d)), C.c_uint(seed), C.c_int(objective)))
            hcheck(hlib().tnh_net_set_fusion(self.h, C.c_int(int(fusion))))
            self.layers = None
            self.nin, self.nout, self.nlayers = int(d[0]), int(d[-1]), 2 * (len(d) - 1)
            return
        if path is None:
            self._tmp = tempfile.NamedTemporaryFile(suffix=".nnet", delete=False)
            self._tmp.close()
            F.write_mlp(self._tmp.name, layers)
            path = self._tmp.name
        try:
            self.layers = layers if layers is not None else F.read_mlp(path)
        except Exception:
            self.layers = None          # let the library report what is wrong with the file
        self.h = C.c_void_p()
        hcheck(hlib().tnh_net_read(C.byref(self.h), path.encode(), C.c_int(objective)))
        hcheck(hlib().tnh_net_set_fusion(self.h, C.c_int(int(fusion))))
        nin, nout, nl = C.c_int(), C.c_int(), C.c_int()
        hcheck(hlib().tnh_net_dims(self.h, C.byref(nin), C.byref(nout), C.byref(nl)))
        self.nin, self.nout, self.nlayers = nin.value, nout.value, nl.value

    def set_hyper(self, lr, mmt=0.0, wc=0.0, gdf=True, factors=None):
        fs = None if factors is None else ":".join(repr(float(x)) for x in factors).encode()
        hcheck(hlib().tnh_net_set_hyper(self.h, C.c_float(lr), fs, C.c_float(mmt), C.c_float(wc), C.c_int(int(gdf))))

    def set_batching(self, on):
        hcheck(hlib().tnh_net_set_batching(self.h, C.c_int(int(on))))

    def get_affine_raw(self, idx):
        """(W^T [nout x nin], bias) of layer idx straight from the device (no text round trip, full fp32 precision)"""
        nin, nout = C.c_int(), C.c_int()
        hcheck(hlib().tnh_net_get_affine(self.h, C.c_int(idx), None, None, C.byref(nin), C.byref(nout)))
        W = np.empty((nin.value, nout.value), np.float32)
        b = np.empty(nout.value, np.float32)
        hcheck(hlib().tnh_net_get_affine(self.h, C.c_int(idx), P(W), P(b), None, None))
        return np.ascontiguousarray(W.T), b

    def set_data_parallel(self, world):
        hcheck(hlib().tnh_net_set_data_parallel(self.h, C.c_int(world)))

    def propagate(self, X):
        X = f32(X)
        out = np.empty((X.shape[0], self.nout), np.float32)
        hcheck(hlib().tnh_net_propagate(self.h, P(X), C.c_int(X.shape[0]), P(out)))
        return out

    def train_bunch(self, X, T, cv=False):
        X = f32(X); T = f32(T)
        hcheck(hlib().tnh_net_train_bunch(self.h, P(X), P(T), C.c_int(X.shape[0]), C.c_int(int(cv))))

    def train_bunch_labels(self, X, lab, cv=False):
        hcheck(hlib().tnh_net_train_bunch_labels(self.h, P(X), P(lab), C.c_int(X.shape[0]), C.c_int(int(cv))))

    def train_from_cache(self, cache, cv=False):
        n = C.c_int()
        hcheck(hlib().tnh_net_train_from_cache(self.h, cache.h, C.c_int(int(cv)), C.byref(n)))
        return n.value

    def load_resident(self, X, lab):
        X = f32(X); lab = np.ascontiguousarray(lab, dtype=np.int32)
        hcheck(hlib().tnh_net_load_resident(self.h, P(X), P(lab), C.c_int(X.shape[0])))

    def train_resident(self, bunch, first, n, cv=False):
        hcheck(hlib().tnh_net_train_resident(self.h, C.c_int(bunch), C.c_int(first), C.c_int(n), C.c_int(int(cv))))

    def submit_bunch_labels(self, x_ptr, lab_ptr, rows, cv=False):
        """pipelined step from PINNED host buffers (ctypes pointers); pair with collect()"""
        hcheck(hlib().tnh_net_submit_bunch_labels(self.h, x_ptr, lab_ptr, C.c_int(rows), C.c_int(int(cv))))

    def collect(self):
        e = C.c_double(); fr = C.c_longlong(); co = C.c_longlong()
        hcheck(hlib().tnh_net_collect(self.h, C.byref(e), C.byref(fr), C.byref(co)))
        return e.value, fr.value, co.value

    def stats(self):
        e = C.c_double(); fr = C.c_longlong(); co = C.c_longlong()
        hcheck(hlib().tnh_net_stats(self.h, C.byref(e), C.byref(fr), C.byref(co)))
        return e.value, fr.value, co.value

    def add_stats(self, e, fr, co):
        hcheck(hlib().tnh_net_add_stats(self.h, C.c_double(e), C.c_longlong(fr), C.c_longlong(co)))

    def layer_out(self, idx, rows):
        n = _layer_dims(self.layers[idx])[1]
        out = np.empty((rows, n), np.float32)
        hcheck(hlib().tnh_net_layer_output(self.h, C.c_int(idx), P(out), C.c_int(rows), C.c_int(n)))
        return out

    def layer_eout(self, idx, rows):
        n = _layer_dims(self.layers[idx])[0]
        out = np.empty((rows, n), np.float32)
        hcheck(hlib().tnh_net_layer_error_output(self.h, C.c_int(idx), P(out), C.c_int(rows), C.c_int(n)))
        return out

    def err(self, rows):
        out = np.empty((rows, self.nout), np.float32)
        hcheck(hlib().tnh_net_global_error(self.h, P(out), C.c_int(rows), C.c_int(self.nout)))
        return out

    def write(self, path):
        hcheck(hlib().tnh_net_write(self.h, path.encode()))

    def get_layers(self):
        """Round trip through the text format (6 significant digits, like the reference's checkpoints)."""
        with tempfile.NamedTemporaryFile(suffix=".nnet", delete=False) as t:
            name = t.name
        try:
            self.write(name)
            return F.read_mlp(name)
        finally:
            os.unlink(name)

    def get_affine(self, idx):
        L = self.get_layers()[idx]
        return L[1], L[2]

    def close(self):
        if self.h:
            hlib().tnh_net_free(self.h)
            self.h = C.c_void_p()
        if self._tmp is not None:
            try:
                os.unlink(self._tmp.name)
            except OSError:
                pass
            self._tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Cache:
    def __init__(self, cachesize, bunchsize):
        self.h = C.c_void_p()
        hcheck(hlib().tnh_cache_new(C.byref(self.h), C.c_int(cachesize), C.c_int(bunchsize)))
        self.bunch, self.cachesize = bunchsize, cachesize
        self.fdim = self.ddim = None

    def add(self, Fm, D):
        Fm = f32(Fm); D = f32(D)
        self.fdim, self.ddim = Fm.shape[1], D.shape[1]
        hcheck(hlib().tnh_cache_add(self.h, P(Fm), P(D), C.c_int(Fm.shape[0]), C.c_int(Fm.shape[1]), C.c_int(D.shape[1])))

    def full(self):
        return bool(hlib().tnh_cache_full(self.h))

    def empty(self):
        return bool(hlib().tnh_cache_empty(self.h))

    def discarded(self):
        return hlib().tnh_cache_discarded(self.h)

    def randomize(self):
        perm = np.empty(self.cachesize, np.int32)
        n = C.c_int()
        hcheck(hlib().tnh_cache_randomize(self.h, P(perm), C.byref(n)))
        return perm[:n.value].copy()

    def get_bunch(self):
        Fm = np.empty((self.bunch, self.fdim), np.float32)
        D = np.empty((self.bunch, self.ddim), np.float32)
        hcheck(hlib().tnh_cache_get_bunch(self.h, P(Fm), P(D)))
        return Fm, D

    def __del__(self):
        try:
            if self.h:
                hlib().tnh_cache_free(self.h)
        except Exception:
            pass


class Rbm:
    def __init__(self, Wt, vb, hb, vis_gauss, hid_gauss, bunch, lr, mmt, wc, sparse_cost=None):
        """sparse_cost != None: the layer is written as <rbmsparse> with that sparsity cost (CuRbmSparse)"""
        with tempfile.NamedTemporaryFile(suffix=".rbm", delete=False) as t:
            name = t.name
        units = ("gauss" if vis_gauss else "bern", "gauss" if hid_gauss else "bern")
        if sparse_cost is None:
            F.write_mlp(name, [("rbm",) + units + (f32(Wt), f32(vb), f32(hb))])
        else:
            F.write_mlp(name, [("rbmsparse",) + units + (f32(Wt), f32(vb), f32(hb), float(sparse_cost))])
        self.h = C.c_void_p()
        try:
            hcheck(hlib().tnh_rbm_read(C.byref(self.h), name.encode(), C.c_int(bunch), C.c_float(lr), C.c_float(mmt), C.c_float(wc)))
        finally:
            os.unlink(name)
        self.nhid, self.nvis = Wt.shape

    def cd1(self, pos_vis):
        pos_vis = f32(pos_vis)
        hcheck(hlib().tnh_rbm_cd1_bunch(self.h, P(pos_vis), C.c_int(pos_vis.shape[0])))

    def cd1_from_cache(self, cache):
        n = C.c_int()
        hcheck(hlib().tnh_rbm_cd1_from_cache(self.h, cache.h, C.byref(n)))
        return n.value

    def last(self, rows):
        ph = np.empty((rows, self.nhid), np.float32); nh = np.empty_like(ph); nv = np.empty((rows, self.nvis), np.float32)
        hcheck(hlib().tnh_rbm_last(self.h, P(ph), P(nh), P(nv)))
        return ph, nh, nv

    def get(self):
        with tempfile.NamedTemporaryFile(suffix=".rbm", delete=False) as t:
            name = t.name
        try:
            hcheck(hlib().tnh_rbm_write(self.h, name.encode()))
            L = F.read_mlp(name)[0]
        finally:
            os.unlink(name)
        return L[3], L[4], L[5]

    def submit_bunch_labels(self, x_ptr, lab_ptr, rows, cv=False):
        """pipelined step from PINNED host buffers (ctypes pointers); pair with collect()"""
        hcheck(hlib().tnh_net_submit_bunch_labels(self.h, x_ptr, lab_ptr, C.c_int(rows), C.c_int(int(cv))))

    def collect(self):
        e = C.c_double(); fr = C.c_longlong(); co = C.c_longlong()
        hcheck(hlib().tnh_net_collect(self.h, C.byref(e), C.byref(fr), C.byref(co)))
        return e.value, fr.value, co.value

    def stats(self):
        e = C.c_double(); fr = C.c_longlong()
        hcheck(hlib().tnh_rbm_stats(self.h, C.byref(e), C.byref(fr)))
        return e.value, fr.value

    def __del__(self):
        try:
            if self.h:
                hlib().tnh_rbm_free(self.h)
        except Exception:
            pass


class Rnn:
    def __init__(self, layers, bptt, lr, mmt=0.0, wc=0.0):
        with tempfile.NamedTemporaryFile(suffix=".rnn", delete=False) as t:
            name = t.name
        F.write_mlp(name, layers)
        self.h = C.c_void_p()
        try:
            hcheck(hlib().tnh_rnn_read(C.byref(self.h), name.encode(), C.c_int(bptt), C.c_float(lr), C.c_float(mmt), C.c_float(wc)))
        finally:
            os.unlink(name)

    def train_utterance(self, X, lab, cv=False):
        X = f32(X); lab = np.ascontiguousarray(lab, dtype=np.int32)
        hcheck(hlib().tnh_rnn_train_utterance(self.h, P(X), P(lab), C.c_int(X.shape[0]), C.c_int(int(cv))))

    def get_layers(self):
        with tempfile.NamedTemporaryFile(suffix=".rnn", delete=False) as t:
            name = t.name
        try:
            hcheck(hlib().tnh_rnn_write(self.h, name.encode()))
            return F.read_mlp(name)
        finally:
            os.unlink(name)

    def submit_bunch_labels(self, x_ptr, lab_ptr, rows, cv=False):
        """pipelined step from PINNED host buffers (ctypes pointers); pair with collect()"""
        hcheck(hlib().tnh_net_submit_bunch_labels(self.h, x_ptr, lab_ptr, C.c_int(rows), C.c_int(int(cv))))

    def collect(self):
        e = C.c_double(); fr = C.c_longlong(); co = C.c_longlong()
        hcheck(hlib().tnh_net_collect(self.h, C.byref(e), C.byref(fr), C.byref(co)))
        return e.value, fr.value, co.value

    def stats(self):
        e = C.c_double(); fr = C.c_longlong(); co = C.c_longlong()
        hcheck(hlib().tnh_rnn_stats(self.h, C.byref(e), C.byref(fr), C.byref(co)))
        return e.value, fr.value, co.value

    def __del__(self):
        try:
            if self.h:
                hlib().tnh_rnn_free(self.h)
        except Exception:
            pass

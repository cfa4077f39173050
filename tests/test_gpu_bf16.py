"""TNB_MATH_BF16 (reported separately from the fp32-equivalent default): operands rounded to bf16, fp32 accumulation.
The checker is the oracle GEMM (double accumulation) run on the SAME bf16-rounded operands: products of two bf16 numbers are
exact in fp32, so what is left is the tensor core's fp32 accumulation — the tolerance of the 3xTF32 tests applies unchanged.
The bf16 twins the fused ops write are held bit-exact to round-to-nearest-even of the fp32 value stored beside them."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_lib as O
from tnet_b200 import abi

L = None


@pytest.fixture(scope="module")
def ctx():
    global L
    L = abi.lib()
    c = abi.Context(0, abi.MATH_BF16)
    yield c
    c.close()


def tol(K):
    return max(2e-6, 8e-8 * np.sqrt(K))


SHAPES = [
    ("N", "N", 256, 1024, 351), ("N", "N", 256, 135, 1024), ("N", "T", 256, 1024, 135), ("T", "N", 351, 1024, 256),
    ("T", "N", 1024, 135, 256), ("N", "N", 1024, 2048, 429), ("N", "T", 1024, 2048, 3000), ("T", "N", 2048, 3000, 1024),
    ("N", "N", 1, 7, 5), ("T", "T", 130, 70, 45), ("N", "N", 128, 64, 64), ("N", "N", 1024, 2048, 2048),
    ("N", "T", 1024, 2048, 2048), ("T", "N", 2048, 2048, 1024),
]


@pytest.mark.parametrize("ta,tb,M,N,K", SHAPES)
def test_gemm_bf16_vs_oracle_on_rounded_operands(ctx, ta, tb, M, N, K):
    r = np.random.default_rng(M + 3 * N + 7 * K)
    A = r.standard_normal((K, M) if ta == "T" else (M, K)).astype(np.float32)
    B = r.standard_normal((N, K) if tb == "T" else (K, N)).astype(np.float32)
    C0 = r.standard_normal((M, N)).astype(np.float32)
    dA, dB, dC = abi.DMat.from_numpy(ctx, A), abi.DMat.from_numpy(ctx, B), abi.DMat.from_numpy(ctx, C0)
    abi.gemm(ctx, ta, tb, 0.75, dA, dB, -0.5, dC)
    Ar, Br = abi.bf16_round(A), abi.bf16_round(B)
    ref = O.gemm(ta, tb, 0.75, Ar, Br, -0.5, C0, acc_double=1)
    opA, opB = (Ar.T if ta == "T" else Ar), (Br.T if tb == "T" else Br)
    scale = np.abs(opA).astype(np.float64) @ np.abs(opB).astype(np.float64) * 0.75 + 0.5 * np.abs(C0)
    err = np.abs(dC.download().astype(np.float64) - ref) / (scale + 1e-30)
    assert err.max() < tol(K), err.max()


@pytest.mark.parametrize("rows,cols", [(1, 1), (7, 13), (256, 351), (100, 3000), (33, 64)])
def test_to_bf16_bit_exact(ctx, rows, cols):
    r = np.random.default_rng(rows * cols)
    a = (r.standard_normal((rows, cols)) * 10.0 ** r.integers(-6, 6, (rows, cols))).astype(np.float32)
    a.flat[0] = 0.0
    d = abi.DMat.from_numpy(ctx, a)
    d16 = abi.DMat16.from_fp32(ctx, d)
    bits = d16.download_bits(full_pitch=True)
    want = (abi.bf16_round(a).view(np.uint32) >> 16).astype(np.uint16)
    assert np.array_equal(bits[:, :cols], want)
    assert not bits[:, cols:].any()       # pitch padding is zero, never NaN garbage for a TMA box to pick up


def test_fused_layer_ops_with_resident_twins(ctx):
    """forward (bias + sigmoid), dX (diff-sigmoid) and the fused update through the *_bf16 entry points: fp32 results equal the
    oracle on rounded operands, the written twins are the RN-even bf16 of the stored fp32 values."""
    r = np.random.default_rng(42)
    rows, nin, nout = 384, 429, 1024
    X = r.standard_normal((rows, nin)).astype(np.float32)
    W = (0.1 * r.standard_normal((nin, nout))).astype(np.float32)
    b = r.standard_normal(nout).astype(np.float32)
    dX, dW, db = abi.DMat.from_numpy(ctx, X), abi.DMat.from_numpy(ctx, W), abi.DMat.from_numpy(ctx, b)
    X16, W16 = abi.DMat16.from_fp32(ctx, dX), abi.DMat16.from_fp32(ctx, dW)
    dY, Y16 = abi.DMat(ctx, rows, nout), abi.DMat16(ctx, rows, nout)
    abi.check(L.tnb_affine_fwd_bf16(ctx.h, X16.p(), C.c_int(X16.stride), dX.dim, W16.p(), C.c_int(W16.stride), dW.dim, db.p(),
                                    dY.p(), dY.dim, Y16.p(), C.c_int(Y16.stride), C.c_int(abi.ACT_SIGMOID)))
    Xr, Wr = abi.bf16_round(X), abi.bf16_round(W)
    Y = dY.download()
    ref = O.sigmoid(O.gemm("N", "N", 1.0, Xr, Wr, 1.0, np.tile(b, (rows, 1)), acc_double=1))
    np.testing.assert_allclose(Y, ref, rtol=2e-5, atol=2e-6)
    assert np.array_equal(Y16.download(), abi.bf16_round(Y))
    # the same through the fp32-array entry point (operands rounded into context scratch): identical bits
    dY2 = abi.DMat(ctx, rows, nout)
    abi.check(L.tnb_affine_fwd(ctx.h, dX.p(), dX.dim, dW.p(), dW.dim, db.p(), dY2.p(), dY2.dim, C.c_int(abi.ACT_SIGMOID)))
    assert np.array_equal(dY2.download(), Y)

    # dX of the layer above: Eprev = (E * W^T) .* y(1-y)
    E = (0.1 * r.standard_normal((rows, nout))).astype(np.float32)
    Yp = r.random((rows, nin)).astype(np.float32)
    dE, dYp = abi.DMat.from_numpy(ctx, E), abi.DMat.from_numpy(ctx, Yp)
    E16 = abi.DMat16.from_fp32(ctx, dE)
    dEp, Ep16 = abi.DMat(ctx, rows, nin), abi.DMat16(ctx, rows, nin)
    abi.check(L.tnb_affine_bwd_dx_bf16(ctx.h, E16.p(), C.c_int(E16.stride), dE.dim, W16.p(), C.c_int(W16.stride), dW.dim, dYp.p(),
                                       dYp.dim, dEp.p(), dEp.dim, Ep16.p(), C.c_int(Ep16.stride)))
    Er = abi.bf16_round(E)
    Ep = dEp.download()
    refp = O.diff_sigmoid(O.gemm("N", "T", 1.0, Er, Wr, 0.0, np.zeros((rows, nin), np.float32), acc_double=1), Yp)
    np.testing.assert_allclose(Ep, refp, rtol=2e-5, atol=2e-6)
    assert np.array_equal(Ep16.download(), abi.bf16_round(Ep))

    # fused update: corrW = X^T E + mmt corrW ; W += s corrW ; W += l2 W ; twin of W refreshed
    cW = (0.01 * r.standard_normal((nin, nout))).astype(np.float32)
    cb = (0.01 * r.standard_normal(nout)).astype(np.float32)
    dcW, dcb = abi.DMat.from_numpy(ctx, cW), abi.DMat.from_numpy(ctx, cb)
    lr, mmt, wc = 0.05, 0.5, 1e-4
    abi.check(L.tnb_affine_update_bf16(ctx.h, X16.p(), C.c_int(X16.stride), dX.dim, E16.p(), C.c_int(E16.stride), dE.p(), dE.dim,
                                       dW.p(), dW.dim, W16.p(), C.c_int(W16.stride), db.p(), dcW.p(), dcb.p(), C.c_float(lr),
                                       C.c_float(mmt), C.c_float(wc), C.c_int(1), C.c_int(0)))
    W2, b2, cW2, cb2 = W.copy(), b.copy(), cW.copy(), cb.copy()
    # weight side from the rounded operands, bias side from the fp32 error (column sums are not a tensor-core op)
    O.lib.orc_affine_update(O.P(Xr), nin, O.P(Er), nout, O.P(W2), nout, O.P(b2.copy()), O.P(cW2), nout, O.P(cb2.copy()), rows, nin, nout,
                            O.cf(lr), O.cf(mmt), O.cf(wc), 1, 1)
    Wd = W.copy(); cWd = cW.copy()
    O.lib.orc_affine_update(O.P(X), nin, O.P(E), nout, O.P(Wd), nout, O.P(b2), O.P(cWd), nout, O.P(cb2), rows, nin, nout,
                            O.cf(lr), O.cf(mmt), O.cf(wc), 1, 1)
    Wn = dW.download()
    np.testing.assert_allclose(dcW.download(), cW2, rtol=1e-5, atol=1e-5 * np.abs(cW2).max())
    np.testing.assert_allclose(Wn, W2, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dcb.download()[0], cb2, rtol=1e-5, atol=1e-5 * np.abs(cb2).max())
    np.testing.assert_allclose(db.download()[0], b2, rtol=1e-5, atol=1e-6)
    assert np.array_equal(W16.download(), abi.bf16_round(Wn))


def _oracle_bf16_step(params, corr, X, T, lr, mmt, wc):
    """One bunch of a sigmoid MLP + softmax/xent in bf16 mode, restated from the oracle's primitives: every GEMM (forward,
    dX, dW) on bf16-rounded operands with double accumulation, everything else exactly as the fp32 path (CuNetwork::Propagate /
    Backpropagate order, cuBiasedLinearity.cc:44-64 update through orc_affine_update)."""
    rows = X.shape[0]
    acts = [X]
    n = len(params)
    for l, (W, b) in enumerate(params):
        pre = O.gemm("N", "N", 1.0, abi.bf16_round(acts[-1]), abi.bf16_round(W), 1.0, np.tile(b, (rows, 1)), acc_double=1)
        acts.append(O.softmax(pre) if l == n - 1 else O.sigmoid(pre))
    E, xent, frames, correct = O.xent_evaluate(acts[-1], T)
    for l in range(n - 1, -1, -1):
        W, b = params[l]
        nin, nout = W.shape
        Eprev = None
        if l > 0:   # dX reads the weights BEFORE this layer's update (CuNetwork::Backpropagate: Backpropagate() then Update())
            Eprev = O.diff_sigmoid(O.gemm("N", "T", 1.0, abi.bf16_round(E), abi.bf16_round(W), 0.0, np.zeros((rows, nin), np.float32),
                                          acc_double=1), acts[l])
        cW, cb = corr[l]
        Wq, cWq = W.copy(), cW.copy()       # weight side: rounded operands
        O.lib.orc_affine_update(O.P(abi.bf16_round(acts[l])), nin, O.P(abi.bf16_round(E)), nout, O.P(Wq), nout, O.P(b.copy()), O.P(cWq), nout,
                                O.P(cb.copy()), rows, nin, nout, O.cf(lr), O.cf(mmt), O.cf(wc), 1, 1)
        Wd, cWd = W.copy(), cW.copy()       # bias side: fp32 error
        O.lib.orc_affine_update(O.P(O.f32(acts[l])), nin, O.P(O.f32(E)), nout, O.P(Wd), nout, O.P(b), O.P(cWd), nout, O.P(cb), rows, nin,
                                nout, O.cf(lr), O.cf(mmt), O.cf(wc), 1, 1)
        params[l] = (Wq, b)
        corr[l] = (cWq, cb)
        E = Eprev
    return acts, xent, correct


@pytest.mark.parametrize("fusion", [True, False], ids=["fused", "unfused"])
def test_bf16_network_two_bunches_vs_oracle_on_rounded_operands(fusion):
    from tnet_b200 import formats as F
    from tnet_b200 import host
    dims, bunch = [429, 320, 256, 3000], 256
    r = np.random.default_rng(7)
    layers = F.gen_mlp_init(dims, r)
    X = r.standard_normal((bunch, dims[0])).astype(np.float32)
    T = np.zeros((bunch, dims[-1]), np.float32)
    T[np.arange(bunch), r.integers(0, dims[-1], bunch)] = 1
    lr, mmt, wc = 0.1, 0.5, 1e-4
    host.set_math(abi.MATH_BF16)
    try:
        net = host.Net(layers, fusion=fusion)
        net.set_hyper(lr, mmt=mmt, wc=wc, gdf=True)
        params = [(np.ascontiguousarray(L[1].T), L[2].copy()) for L in layers if L[0] == "affine"]   # in-memory layout [nin x nout]
        corr = [(np.zeros_like(W), np.zeros_like(b)) for W, b in params]
        xent = 0.0
        for step in range(2):       # the second bunch reads the twins the first bunch's update epilogues wrote
            net.train_bunch(X, T)
            acts, x1, _ = _oracle_bf16_step(params, corr, X, T, lr, mmt, wc)
            xent += x1
        nl = len(layers)
        for i in range(1, nl, 2):   # sigmoid / softmax outputs of the second bunch
            a, b = net.layer_out(i, bunch), acts[(i + 1) // 2]
            np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(b).max()), err_msg="output of layer %d" % i)
        got = net.get_layers()
        for k, i in enumerate(range(0, nl, 2)):
            W, b = params[k]
            np.testing.assert_allclose(got[i][1], W.T, rtol=2e-5, atol=2e-5 * np.abs(W).max())
            np.testing.assert_allclose(got[i][2], b, rtol=2e-5, atol=2e-5 * max(1e-2, np.abs(b).max()))
        e1, f1, _ = net.stats()
        assert f1 == 2 * bunch and abs(e1 - xent) <= 2e-5 * abs(xent)
    finally:
        host.set_math(abi.MATH_3XTF32)

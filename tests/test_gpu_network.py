"""Network-level parity of the CUDA path (through libtnetb200_host.so, i.e. the C++ mirror of CuNetwork/CuCache/
CuObjectiveFunction/CuRbm/CuRecurrent) against (1) the reference's own outputs — golden fixtures produced by the
unmodified TNet (CPU) and TNetCu/TRbmCu/TRecurrentCu (run on a B200) — and (2) the oracle, layer by layer."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_lib as O
from replay import compare_layer, fixture_layers, replay_mlp, replay_rbm, rnn_layers, rnn_utterances, utterances
from tnet_b200 import abi, host

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MLP_GOLD = sorted(glob.glob(os.path.join(GOLD, "*_mlp_*.npz")))


class _OracleCache(O.Cache):
    pass


def _srand_both(seed):
    host.srand48(seed)          # same libc state: both libraries call glibc's lrand48


@pytest.mark.parametrize("path", MLP_GOLD, ids=[os.path.basename(p)[:-4] for p in MLP_GOLD])
@pytest.mark.parametrize("fusion", [True, False], ids=["fused", "unfused"])
def test_mlp_epoch_matches_reference_trainers(path, fusion):
    """One epoch through CuCache + CuNetwork on the CUDA path == the reference trainer's written network and report."""
    g = np.load(path)
    host.set_math(abi.MATH_3XTF32)
    net, nb, perms = replay_mlp(g, lambda L: host.Net(L, fusion=fusion), host.Cache, _srand_both)
    # cache permutations: bit-exact against the oracle's restatement of random_shuffle + lrand48
    onet, onb, operms = replay_mlp(g, lambda L: O.Net(L, acc_double=0), O.Cache, lambda s: O.lib.orc_srand48(s))
    assert nb == onb and len(perms) == len(operms)
    for p, q in zip(perms, operms):
        assert np.array_equal(p, q)
    err, frames, correct = net.stats()
    assert frames == int(g["ref_frames"])
    # per-epoch cross-entropy: tolerance 1e-4 relative (3xTF32 GEMMs + float softmax vs fp32 SGEMM + double softmax,
    # compounded over the epoch's updates); the reference prints 6 significant digits
    assert abs(err - float(g["ref_err"])) <= 1e-4 * abs(float(g["ref_err"]))
    ref_correct = round(float(g["ref_correct_pct"]) * frames / 100.0)
    assert abs(correct - ref_correct) <= max(1, int(0.002 * frames))
    oerr, ofr, ocorrect = onet.stats()
    assert abs(correct - ocorrect) <= max(1, int(0.002 * frames))
    layers = net.get_layers()
    k = 0
    for L in layers:
        if L[0] != "affine":
            continue
        rW, rb = g["final_Wt%d" % k], g["final_b%d" % k]
        np.testing.assert_allclose(L[1], rW, rtol=2e-4, atol=2e-4 * np.abs(rW).max())
        np.testing.assert_allclose(L[2], rb, rtol=2e-4, atol=2e-4 * max(1e-2, np.abs(rb).max()))
        k += 1


@pytest.mark.parametrize("fusion", [True, False], ids=["fused", "unfused"])
@pytest.mark.parametrize("dims,bunch", [([351, 1024, 135], 256), ([65, 48, 300], 128), ([429, 64, 64, 64, 3000], 96)])
def test_one_bunch_layerwise_vs_oracle(dims, bunch, fusion):
    """Activations, error signals and updated weights after a single bunch, layer by layer (tolerance: 3xTF32 GEMM vs
    double-accumulated oracle, 2e-5 relative to the layer's scale)."""
    from tnet_b200 import formats as F
    r = np.random.default_rng(sum(dims))
    layers = F.gen_mlp_init(dims, r)
    X = r.standard_normal((bunch, dims[0])).astype(np.float32)
    T = np.zeros((bunch, dims[-1]), np.float32)
    T[np.arange(bunch), r.integers(0, dims[-1], bunch)] = 1
    net = host.Net(layers, fusion=fusion)
    onet = O.Net(net.layers, acc_double=1)
    for n in (net, onet):
        n.set_hyper(0.1, mmt=0.5, wc=1e-4, gdf=True)
    for step in range(2):           # second step exercises the momentum buffers
        net.train_bunch(X, T)
        onet.train_bunch(X, T)
    nl = len(layers)
    for i in range(nl):
        kind = layers[i][0]
        if fusion and kind == "affine" and i + 1 < nl and layers[i + 1][0] == "sigmoid":
            continue                # pre-activation is never materialised on the fused path
        a, b = net.layer_out(i, bunch), onet.layer_out(i, bunch)
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(b).max()), err_msg="output of layer %d" % i)
    np.testing.assert_allclose(net.err(bunch), onet.err(bunch), rtol=1e-3, atol=1e-6)
    for i in range(2, nl):          # error outputs exist above the stopper (first affine)
        kind = layers[i][0]
        if fusion and (kind == "softmax" or (kind == "affine" and layers[i - 1][0] == "sigmoid")):
            continue                # identity copy / dX+diffsigmoid written straight into the layer below
        a, b = net.layer_eout(i, bunch), onet.layer_eout(i, bunch)
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=2e-5 * max(1e-3, np.abs(b).max()), err_msg="error output of layer %d" % i)
    got = net.get_layers()
    for i in range(0, nl, 2):
        Wt, b = onet.get_affine(i)
        np.testing.assert_allclose(got[i][1], Wt, rtol=2e-5, atol=2e-5 * np.abs(Wt).max())
        np.testing.assert_allclose(got[i][2], b, rtol=2e-5, atol=2e-5 * max(1e-2, np.abs(b).max()))
    e1, f1, c1 = net.stats()
    e2, f2, c2 = onet.stats()
    assert f1 == f2 and abs(c1 - c2) <= 1 and abs(e1 - e2) <= 2e-5 * abs(e2)


@pytest.mark.parametrize("batching", [True, False], ids=["batched", "unbatched"])
def test_config_c_one_bunch_layerwise_vs_oracle(batching):
    """BASELINE configs[2] at FULL size — 429-2048x6-3000, bunch 1024, lr 0.008 / momentum 0.5 / weightcost 1e-6 as bench.py runs it —
    one training bunch through CuNetwork (the tile shapes, split-K pairs and batched persistent launches the bench uses) against
    the oracle with double accumulation: every layer's activations, the error signals above the stopper, the updated weights and
    biases of all seven layers, cross-entropy and the frame-accuracy count.  Tolerances are those of the small cases above
    (3xTF32 GEMM vs double-accumulated oracle).  Parameters go to the oracle straight from the device (tnh_net_get_affine)."""
    dims, bunch = [429, 2048, 2048, 2048, 2048, 2048, 2048, 3000], 1024
    r = np.random.default_rng(2024)
    host.set_math(abi.MATH_3XTF32)
    net = host.Net(dims=dims, seed=7)
    net.set_batching(batching)
    nl = 2 * (len(dims) - 1)
    layers = []
    for i in range(0, nl, 2):
        Wt, b = net.get_affine_raw(i)
        layers.append(("affine", Wt, b))
        layers.append(("softmax" if i == nl - 2 else "sigmoid", dims[i // 2 + 1]))
    X = r.standard_normal((bunch, dims[0])).astype(np.float32)
    lab = r.integers(0, dims[-1], bunch)
    T = np.zeros((bunch, dims[-1]), np.float32)
    T[np.arange(bunch), lab] = 1
    onet = O.Net(layers, acc_double=1)
    for n in (net, onet):
        n.set_hyper(0.008, mmt=0.5, wc=1e-6, gdf=True)
        n.train_bunch(X, T)
    net.layers = layers
    for i in range(nl):
        if layers[i][0] == "affine" and i + 1 < nl and layers[i + 1][0] == "sigmoid":
            continue                # pre-activation is never materialised on the fused path
        a, b = net.layer_out(i, bunch), onet.layer_out(i, bunch)
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(b).max()), err_msg="output of layer %d" % i)
    np.testing.assert_allclose(net.err(bunch), onet.err(bunch), rtol=1e-3, atol=1e-6)
    for i in range(2, nl):
        kind = layers[i][0]
        if kind == "softmax" or (kind == "affine" and layers[i - 1][0] == "sigmoid"):
            continue                # identity copy / dX+diffsigmoid written straight into the layer below
        a, b = net.layer_eout(i, bunch), onet.layer_eout(i, bunch)
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=2e-5 * max(1e-3, np.abs(b).max()), err_msg="error output of layer %d" % i)
    for i in range(0, nl, 2):
        Wt, b = net.get_affine_raw(i)
        oWt, ob = onet.get_affine(i)
        # the update of one bunch moves a weight by ~1e-6: compare the MOVEMENT, not only the weight (whose 2e-5 tolerance would hide it)
        dW, odW = Wt - layers[i][1], oWt - layers[i][1]
        # (both sides round W + dW to fp32 independently: two ulps of the largest weight on top of the GEMM tolerance)
        np.testing.assert_allclose(dW, odW, rtol=1e-3, atol=2e-5 * np.abs(odW).max() + 2 * np.spacing(np.abs(oWt).max()),
                                   err_msg="weight update of layer %d" % i)
        np.testing.assert_allclose(b - layers[i][2], ob - layers[i][2], rtol=1e-3,
                                   atol=2e-5 * np.abs(ob - layers[i][2]).max() + 2 * np.spacing(max(1e-3, np.abs(ob).max())),
                                   err_msg="bias update of layer %d" % i)
    e1, f1, c1 = net.stats()
    e2, f2, c2 = onet.stats()
    assert f1 == f2 == bunch and abs(c1 - c2) <= 1 and abs(e1 - e2) <= 2e-5 * abs(e2)
    net.close()


NET_GOLD = sorted(glob.glob(os.path.join(GOLD, "*_net_*.npz")))


@pytest.mark.parametrize("path", NET_GOLD, ids=[os.path.basename(p)[:-4] for p in NET_GOLD])
@pytest.mark.parametrize("fusion", [True, False], ids=["fused", "unfused"])
def test_offset_gemm_layers_epoch_matches_reference_trainers(path, fusion):
    """networks whose first layer is a <sharedlinearity> / <discretelinearity> (CuMath::OffsetGemm users).  One epoch
    on the CUDA path == the reference trainer's report and written network (cpu_*: unmodified CPU TNet, gpu_*: TNetCu on a B200)."""
    g = np.load(path)
    host.set_math(abi.MATH_3XTF32)
    net, nb, perms = replay_mlp(g, lambda L: host.Net(L, fusion=fusion), host.Cache, _srand_both)
    err, frames, correct = net.stats()
    assert frames == int(g["ref_frames"])
    assert abs(err - float(g["ref_err"])) <= 1e-4 * abs(float(g["ref_err"]))          # tolerance as for the MLP epochs above
    ref_correct = round(float(g["ref_correct_pct"]) * frames / 100.0)
    assert abs(correct - ref_correct) <= max(1, int(0.002 * frames))
    got, final = net.get_layers(), fixture_layers(g, "final")
    assert len(got) == len(final)
    for a, b in zip(got, final):
        compare_layer(a, b, 2e-4, btol_floor=1e-2)


OPT_GOLD = sorted(glob.glob(os.path.join(GOLD, "*_opt_*.npz")))


@pytest.mark.parametrize("path", OPT_GOLD, ids=[os.path.basename(p)[:-4] for p in OPT_GOLD])
def test_option_fixtures_match_reference_trainers(path):
    """Cross-validation (no update, network untouched) and per-layer learning-rate factors with a frozen first layer, against the
    reference trainers' reports and written networks."""
    g = np.load(path)
    host.set_math(abi.MATH_3XTF32)
    net, nb, perms = replay_mlp(g, lambda L: host.Net(L), host.Cache, _srand_both)
    err, frames, correct = net.stats()
    assert frames == int(g["ref_frames"])
    assert abs(err - float(g["ref_err"])) <= 1e-4 * abs(float(g["ref_err"]))
    assert abs(correct - round(float(g["ref_correct_pct"]) * frames / 100.0)) <= max(1, int(0.002 * frames))
    for a, b in zip(net.get_layers(), fixture_layers(g, "final")):
        compare_layer(a, b, 2e-4, btol_floor=1e-2)


def _offset_nets(r):
    def w(o, i):
        return (0.1 * r.standard_normal((o, i))).astype(np.float32)

    def nb(n):
        return r.uniform(-4.1, -3.9, n).astype(np.float32)
    top = lambda nh, no: [("sigmoid", nh), ("affine", w(no, nh), np.zeros(no, np.float32)), ("softmax", no)]
    return {
        # 13-wide column blocks: every block GEMM starts off a 16-byte boundary -> fp32 FMA kernel
        "shared_unaligned": [("shared", 5, w(8, 13), nb(8))] + top(40, 10),
        # 32-wide blocks at 128-byte offsets: the same layer through the tcgen05 kernel
        "shared_aligned": [("shared", 4, w(64, 32), nb(64))] + top(256, 300),
        "discrete_unaligned": [("discrete", [w(12, 26), w(20, 39)], nb(32))] + top(32, 10),
        "discrete_aligned": [("discrete", [w(32, 64), w(96, 32), w(64, 32)], nb(192))] + top(192, 20),
        # the layer in the middle of a network: its dX and the dX through it are both exercised
        "shared_middle": [("affine", w(60, 39), nb(60)), ("sigmoid", 60), ("shared", 3, w(16, 20), nb(16))] + top(48, 12),
        "discrete_middle": [("affine", w(64, 39), nb(64)), ("sigmoid", 64), ("discrete", [w(24, 32), w(8, 32)], nb(32))] + top(32, 12),
    }


@pytest.mark.parametrize("fusion", [True, False], ids=["fused", "unfused"])
@pytest.mark.parametrize("name", ["shared_unaligned", "shared_aligned", "discrete_unaligned", "discrete_aligned", "shared_middle",
                                  "discrete_middle"])
def test_offset_gemm_layers_one_bunch_vs_oracle(name, fusion):
    """<sharedlinearity> / <discretelinearity> forward, dX and update against the oracle's restatement of cuSharedLinearity.cc /
    cuDiscreteLinearity.cc (double-accumulated GEMMs), layer by layer, two bunches (momentum), bunch 96."""
    from tnet_b200 import formats as F
    r = np.random.default_rng(11)
    layers = _offset_nets(r)[name]
    bunch, nin, nout = 96, F.layer_dims(layers[0])[0], F.layer_dims(layers[-1])[1]
    X = r.standard_normal((bunch, nin)).astype(np.float32)
    T = np.zeros((bunch, nout), np.float32)
    T[np.arange(bunch), r.integers(0, nout, bunch)] = 1
    host.set_math(abi.MATH_3XTF32)
    net = host.Net(layers, fusion=fusion)
    onet = O.Net(net.layers, acc_double=1)
    for n in (net, onet):
        n.set_hyper(0.2, mmt=0.5, wc=1e-4, gdf=True)
    for step in range(2):
        net.train_bunch(X, T)
        onet.train_bunch(X, T)
    nl = len(layers)
    first_upd = min(i for i, L in enumerate(layers) if L[0] in ("affine", "shared", "discrete"))
    for i in range(nl):
        if fusion and layers[i][0] == "affine" and i + 1 < nl and layers[i + 1][0] == "sigmoid":
            continue                # pre-activation is never materialised on the fused path
        a, b = net.layer_out(i, bunch), onet.layer_out(i, bunch)
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(b).max()), err_msg="output of layer %d" % i)
    for i in range(first_upd + 1, nl):
        kind = layers[i][0]
        if fusion and (kind == "softmax" or (kind == "affine" and layers[i - 1][0] == "sigmoid")):
            continue                # identity copy / dX+diffsigmoid written straight into the layer below
        if fusion and kind == "sigmoid" and i + 1 < nl and layers[i + 1][0] == "affine":
            # the affine above wrote this sigmoid's error output (dX fused with y(1-y)): it must equal the oracle's
            pass
        a, b = net.layer_eout(i, bunch), onet.layer_eout(i, bunch)
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=2e-5 * max(1e-3, np.abs(b).max()), err_msg="error output of layer %d" % i)
    got = net.get_layers()
    for i in range(nl):
        compare_layer(got[i], onet.get_layer(i), 2e-5, btol_floor=1e-2)
    e1, f1, c1 = net.stats()
    e2, f2, c2 = onet.stats()
    assert f1 == f2 and abs(c1 - c2) <= 1 and abs(e1 - e2) <= 2e-5 * abs(e2)


def test_offset_gemm_layers_reader_errors(tmp_path):
    """Error behaviour of the two readers (cuSharedLinearity.cc:98-160, cuDiscreteLinearity.cc:80-141): exceptions -> error status."""
    bad = str(tmp_path / "bad.nnet")
    for text in ("<sharedlinearity> 6 6\n0\nm 2 2\n1 2 3 4\nv 2 0 0\n",            # "Bad number of instances"
                 "<sharedlinearity> 6 5\n2\nm 3 2\n1 2 3 4 5 6\nv 3 0 0 0\n",      # inputs not divisible by the instances
                 "<sharedlinearity> 6 4\n2\nm 3 3\n1 2 3 4 5 6 7 8 9\nv 3 0 0 0\n",  # block is 3x3, must be 3x2
                 "<discretelinearity> 4 4\n0\nv 4 0 0 0 0\n",                         # "Bad number of blocks"
                 "<discretelinearity> 4 4\n2\nm 2 2\n1 2 3 4\nm 2 1\n1 2\nv 4 0 0 0 0\n"):  # blocks cover 3 of 4 inputs
        open(bad, "w").write(text)
        with pytest.raises(abi.TnbError):
            host.Net(path=bad)


def test_data_parallel_rejects_layers_without_gradient_exchange():
    """Only <biasedlinearity> gradients are exchanged between ranks: a network with another trainable layer must be refused in
    data-parallel mode instead of training every rank on its own rows."""
    r = np.random.default_rng(1)
    net = host.Net(_offset_nets(r)["shared_unaligned"])
    with pytest.raises(abi.TnbError, match="sharedlinearity"):
        net.set_data_parallel(2)


def test_rbm_sparse_cd1_vs_oracle():
    """<rbmsparse> (cuRbmSparse.cc:125-168): CD-1 with the sparsity penalty over a few bunches against the oracle's restatement —
    same Hybrid-Taus states (seeded from the same lrand48 stream), so the Bernoulli samples are bit-identical and the weights
    differ by GEMM rounding only.  The penalty is made large enough to matter (cost 0.05)."""
    from test_oracle_golden_rbm_rnn import _Rbm
    r = np.random.default_rng(23)
    nvis, nhid, bunch, cost = 39, 24, 32, 0.05
    Wt = (0.1 * r.standard_normal((nhid, nvis))).astype(np.float32)
    vb = np.zeros(nvis, np.float32)
    hb = (r.random(nhid) / 5.0 - 0.1).astype(np.float32)
    host.set_math(abi.MATH_3XTF32)
    host.srand48(5)
    a = host.Rbm(Wt, vb, hb, False, False, bunch, 0.1, 0.5, 2e-4, sparse_cost=cost)
    O.lib.orc_srand48(5)
    b = _Rbm(Wt, vb, hb, False, False, bunch, 0.1, 0.5, 2e-4, sparse_cost=cost)
    O.lib.orc_srand48(5)
    plain = _Rbm(Wt, vb, hb, False, False, bunch, 0.1, 0.5, 2e-4)       # same samples, no penalty
    for step in range(6):
        v = r.random((bunch, nvis)).astype(np.float32)
        a.cd1(v)
        b.cd1(v)
        plain.cd1(v)
    for k, (x, y, z) in enumerate(zip(a.get(), b.r.get(), plain.r.get())):
        np.testing.assert_allclose(x, y, rtol=2e-4, atol=2e-5 * max(1e-2, np.abs(y).max()))
        if k != 1:      # weights and hidden bias carry the penalty: it is not a no-op in this test
            assert np.abs(y - z).max() > 1e-3 * max(1e-2, np.abs(y).max())
    ea, fa = a.stats()
    eb, fb = b.r.stats()
    assert fa == fb and abs(ea - eb) <= 1e-4 * abs(eb)


def test_cache_state_machine_vs_oracle():
    """Ragged utterances, leftover carry-over, partial last fill, discarded tail: identical bunches (bit-exact)."""
    r = np.random.default_rng(3)
    lens = [50, 7, 130, 64, 1, 99, 150, 12, 45, 33, 170]   # leftovers stay below the cache size (beyond it the reference asserts)
    seqs = [(r.standard_normal((n, 39)).astype(np.float32), r.standard_normal((n, 5)).astype(np.float32)) for n in lens]
    out = []
    for mk, srand in ((host.Cache, host.srand48), (O.Cache, lambda s: O.lib.orc_srand48(s))):
        srand(77)
        c = mk(192, 32)
        bunches, perms, i = [], [], 0
        while i < len(seqs):
            while not c.full() and i < len(seqs):
                c.add(*seqs[i]); i += 1
            perms.append(c.randomize())
            while not c.empty():
                bunches.append(c.get_bunch())
        out.append((bunches, perms, c.discarded()))
    (b1, p1, d1), (b2, p2, d2) = out
    assert d1 == d2 and len(b1) == len(b2) and len(p1) == len(p2)
    for p, q in zip(p1, p2):
        assert np.array_equal(p, q)
    for (f1, l1), (f2, l2) in zip(b1, b2):
        assert np.array_equal(f1, f2) and np.array_equal(l1, l2)
    with pytest.raises(abi.TnbError):
        host.Cache(100, 32)                       # "Non divisible cachesize by bunchsize"
    c = host.Cache(64, 32)
    with pytest.raises(abi.TnbError):
        c.fdim, c.ddim = 3, 2
        c.get_bunch()                             # "GetBunch on empty cache!!!"


def test_train_from_cache_equals_host_bunches():
    g = np.load(os.path.join(GOLD, "gpu_mlp_small.npz"))
    net1, nb1, _ = replay_mlp(g, lambda L: host.Net(L), host.Cache, _srand_both)

    class DirectNet(host.Net):
        pass
    ctx, bunch, cache, seed, randomize, gdf = [int(v) for v in g["cfg"]]
    lr, mmt, wc = [float(v) for v in g["hyper"]]
    net2 = host.Net(fixture_layers(g))
    net2.set_hyper(lr, mmt=mmt, wc=wc, gdf=bool(gdf))
    host.srand48(seed)
    c = host.Cache((cache // bunch) * bunch, bunch)
    it = iter(utterances(g))
    pending = next(it, None)
    nb2 = 0
    while pending is not None:
        while not c.full() and pending is not None:
            c.add(*pending); pending = next(it, None)
        if randomize:
            c.randomize()
        nb2 += net2.train_from_cache(c)
    assert nb1 == nb2 and net1.stats() == net2.stats()
    for a, b in zip(net1.get_layers(), net2.get_layers()):
        if a[0] == "affine":
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])     # deterministic kernels: bit-identical


def test_network_text_roundtrip_and_errors(tmp_path):
    from tnet_b200 import formats as F
    r = np.random.default_rng(1)
    layers = [("expand", 13, [-1, 0, 1]), ("bias", r.standard_normal(39).astype(np.float32)),
              ("window", r.random(39).astype(np.float32))] + F.gen_mlp_init([39, 16, 4], r)
    p = str(tmp_path / "n.nnet")
    F.write_mlp(p, layers)
    net = host.Net(path=p)
    x = r.standard_normal((20, 13)).astype(np.float32)
    y = net.propagate(x)
    ref = O.expand(x, np.array([-1, 0, 1], np.int32))
    ref = (ref + layers[1][1][None, :]) * layers[2][1][None, :]
    on = O.Net(F.read_mlp(p)[3:], acc_double=1)
    np.testing.assert_allclose(y, on.propagate(ref.astype(np.float32)), rtol=1e-4, atol=1e-6)
    q = str(tmp_path / "o.nnet")
    net.write(q)
    L2 = F.read_mlp(q)
    assert [l[0] for l in L2] == [l[0] for l in F.read_mlp(p)]
    np.testing.assert_allclose(L2[3][1], layers[3][1], rtol=1e-5)            # 6 significant digits on disk
    bad = str(tmp_path / "bad.nnet")
    open(bad, "w").write("<sparselinearity> 4 4\n")
    with pytest.raises(abi.TnbError):
        host.Net(path=bad)
    open(bad, "w").write("<biasedlinearity> 4 3\nm 4 2\n1 2 3 4 5 6 7 8\nv 4 0 0 0 0\n")
    with pytest.raises(abi.TnbError):
        host.Net(path=bad)
    with pytest.raises(abi.TnbError):
        net.set_hyper(0.1, factors=[1.0])                                     # "Too few learninig rate scale factors"
    with pytest.raises(abi.TnbError):
        net.propagate(np.zeros((3, 12), np.float32)) if False else host.hcheck(
            host.hlib().tnh_net_layer_output(net.h, 99, host.P(np.zeros(1, np.float32)), 1, 1))


RBM_GOLD = sorted(glob.glob(os.path.join(GOLD, "gpu_rbm_*.npz")))


@pytest.mark.parametrize("path", RBM_GOLD, ids=[os.path.basename(p)[:-4] for p in RBM_GOLD])
def test_rbm_cd1_matches_reference_trbmcu(path):
    """CD-1 epoch == TRbmCu on a B200: Hybrid-Taus seeding order (lrand48 before the shuffles), Bernoulli sampling,
    the 5 GEMMs and the update."""
    g = np.load(path)
    host.set_math(abi.MATH_3XTF32)

    class CuRbmAdapter(host.Rbm):
        pass
    rbm, nb = replay_rbm(g, lambda *a, **k: host.Rbm(*a, **k), host.Cache, host.srand48)
    err, frames = rbm.stats()
    assert frames == int(g["ref_frames"])
    assert abs(err - float(g["ref_err"])) <= 2e-4 * abs(float(g["ref_err"]))
    Wt, vb, hb = rbm.get()
    np.testing.assert_allclose(Wt, g["final_Wt"], rtol=2e-4, atol=2e-4 * np.abs(g["final_Wt"]).max())
    np.testing.assert_allclose(vb, g["final_vb"], rtol=2e-4, atol=2e-4 * max(1e-3, np.abs(g["final_vb"]).max()))
    np.testing.assert_allclose(hb, g["final_hb"], rtol=2e-4, atol=2e-4 * max(1e-3, np.abs(g["final_hb"]).max()))


def test_rnn_bptt_matches_reference_trecurrentcu():
    g = np.load(os.path.join(GOLD, "gpu_rnn_small.npz"))
    ctx, bptt, nin, H, n_out = [int(v) for v in g["cfg"]]
    lr = float(g["hyper"][0])
    rnn = host.Rnn(rnn_layers(g), bptt, lr)
    for x, lab in rnn_utterances(g):
        rnn.train_utterance(x, lab)
    err, frames, correct = rnn.stats()
    assert frames == int(g["ref_frames"])
    assert abs(err - float(g["ref_err"])) <= 2e-4 * abs(float(g["ref_err"]))
    assert abs(correct - round(float(g["ref_correct_pct"]) * frames / 100.0)) <= 1
    L = rnn.get_layers()
    np.testing.assert_allclose(L[0][1], g["final_Wr"], rtol=3e-4, atol=3e-4 * np.abs(g["final_Wr"]).max())
    np.testing.assert_allclose(L[0][2], g["final_br"], rtol=3e-4, atol=3e-4 * max(1e-3, np.abs(g["final_br"]).max()))
    np.testing.assert_allclose(L[1][1], g["final_Wo"], rtol=3e-4, atol=3e-4 * np.abs(g["final_Wo"]).max())


def test_pipelined_submit_collect_equals_blocking_steps():
    """tnh_net_submit_bunch_labels / tnh_net_collect (H2D on the copy stream, one submission in flight) train exactly like the
    blocking tnh_net_train_bunch_labels: same statistics after every bunch, same final weights (bit-exact: same kernels, same
    order, only the transfer schedule differs)."""
    import ctypes as C
    from tnet_b200 import formats as F
    dims, bunch, steps = [39, 64, 48, 20], 96, 6
    r = np.random.default_rng(5)
    layers = F.gen_mlp_init(dims, r)
    xs = r.standard_normal((steps, bunch, dims[0])).astype(np.float32)
    labs = r.integers(0, dims[-1], (steps, bunch)).astype(np.int32)
    L, H = abi.lib(), host.hlib()
    # pinned staging buffers, one per step (a submission's buffers must stay untouched until it is collected)
    px, pl = C.c_void_p(), C.c_void_p()
    abi.check(L.tnb_host_alloc(C.byref(px), C.c_size_t(xs.nbytes)))
    abi.check(L.tnb_host_alloc(C.byref(pl), C.c_size_t(labs.nbytes)))
    C.memmove(px, xs.ctypes.data, xs.nbytes)
    C.memmove(pl, labs.ctypes.data, labs.nbytes)
    xp = lambda i: C.cast(C.c_void_p(px.value + i * bunch * dims[0] * 4), C.POINTER(C.c_float))
    lp = lambda i: C.cast(C.c_void_p(pl.value + i * bunch * 4), C.POINTER(C.c_int))
    try:
        host.set_math(abi.MATH_3XTF32)
        a = host.Net(layers)
        b = host.Net(layers)
        for n in (a, b):
            n.set_hyper(0.1, mmt=0.5, wc=1e-4, gdf=True)
        ref = []
        for i in range(steps):
            host.hcheck(H.tnh_net_train_bunch_labels(a.h, xp(i), lp(i), C.c_int(bunch), C.c_int(0)))
            ref.append(a.stats())
        got = []
        b.submit_bunch_labels(xp(0), lp(0), bunch)
        for i in range(1, steps):
            b.submit_bunch_labels(xp(i), lp(i), bunch)
            got.append(b.collect())
        got.append(b.collect())
        assert got == ref
        for la, lb in zip(a.get_layers(), b.get_layers()):
            if la[0] == "affine":
                assert np.array_equal(la[1], lb[1]) and np.array_equal(la[2], lb[2])
        with pytest.raises(Exception):
            b.collect()             # nothing in flight
    finally:
        L.tnb_host_free(px)
        L.tnb_host_free(pl)

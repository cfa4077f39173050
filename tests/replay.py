"""Replay of the TNetCu main loop (src/TNetCu.cc:376-441) over a golden fixture, parameterised by the
backend (oracle or CUDA path): fill cache utterance by utterance, Randomize, train bunch by bunch."""
import numpy as np
from tnet_b200 import formats as F


def fixture_layers(g, prefix="init"):
    if prefix + "_net" in g:      # newer fixtures keep the whole network file as text
        return F.read_mlp_text(bytes(g[prefix + "_net"]).decode())
    dims = list(g["dims"])
    layers = []
    for k in range(len(dims) - 1):
        layers.append(("affine", g["%s_Wt%d" % (prefix, k)], g["%s_b%d" % (prefix, k)]))
        layers.append(("softmax" if k == len(dims) - 2 else "sigmoid", int(dims[k + 1])))
    return layers


def compare_layer(got, ref, wtol, btol_floor=1e-3):
    """Parameters of one layer (tuple form of tnet_b200.formats) against another's, e.g. the reference's final network file
    (6 significant digits, KaldiLib/Matrix.tcc:522-532).  Non-parametric layers compare equal."""
    assert got[0] == ref[0]
    if got[0] == "affine":
        mats, vecs = [(got[1], ref[1])], [(got[2], ref[2])]
    elif got[0] == "shared":
        assert got[1] == ref[1]
        mats, vecs = [(got[2], ref[2])], [(got[3], ref[3])]
    elif got[0] == "discrete":
        assert len(got[1]) == len(ref[1])
        mats, vecs = list(zip(got[1], ref[1])), [(got[2], ref[2])]
    else:
        return
    for W, rW in mats:
        np.testing.assert_allclose(W, rW, rtol=wtol, atol=wtol * np.abs(rW).max())
    for b, rb in vecs:
        np.testing.assert_allclose(b, rb, rtol=wtol, atol=wtol * max(btol_floor, np.abs(rb).max()))


def utterances(g):
    ctx = int(g["cfg"][0])
    n_out = int(g["dims"][-1]) if "dims" in g else None
    pos = 0
    for T in g["lengths"]:
        x = g["feats"][pos:pos + T]
        lab = g["labels"][pos:pos + T] if "labels" in g else None
        pos += T
        feats = F.splice(x, ctx)
        des = None
        if lab is not None and n_out is not None:
            des = np.zeros((T, n_out), dtype=np.float32)
            des[np.arange(T), lab] = 1.0
        yield feats, des


def replay_mlp(g, make_net, make_cache, srand48, cv=False):
    """Returns (net, n_bunches, perms).  make_net(layers) -> object with set_hyper/train_bunch/stats/get_affine;
    make_cache(cachesize,bunch) -> object with add/full/empty/randomize/get_bunch; srand48(seed)."""
    ctx, bunch, cache, seed, randomize, gdf = [int(v) for v in g["cfg"]]
    lr, mmt, wc = [float(v) for v in g["hyper"]]
    net = make_net(fixture_layers(g))
    factors = [float(v) for v in g["lr_factors"]] if "lr_factors" in g and len(g["lr_factors"]) else None
    net.set_hyper(lr, mmt=mmt, wc=wc, gdf=bool(gdf), factors=factors)
    if "cv" in g and int(g["cv"]):
        cv = True
    srand48(seed)
    cache = (cache // bunch) * bunch
    c = make_cache(cache, bunch)
    it = iter(utterances(g))
    pending = next(it, None)
    nb = 0
    perms = []
    while pending is not None:
        while not c.full() and pending is not None:
            c.add(pending[0], pending[1])
            pending = next(it, None)
        if randomize:
            perms.append(c.randomize())
        while not c.empty():
            Fb, Db = c.get_bunch()
            net.train_bunch(Fb, Db, cv)
            nb += 1
    return net, nb, perms


def replay_rbm(g, make_rbm, make_cache, srand48):
    """TRbmCu main loop (src/TRbmCu.cc:255-356).  make_rbm(Wt, vb, hb, vis_gauss, hid_gauss, bunch, lr, mmt, wc) must seed
    its generator from lrand48() at construction (after srand48)."""
    ctx, bunch, cache, seed, vis_gauss, hid_gauss = [int(v) for v in g["cfg"]]
    lr, mmt, wc = [float(v) for v in g["hyper"]]
    srand48(seed)
    extra = {}
    if "sparse_cost" in g and float(g["sparse_cost"]) >= 0:      # <rbmsparse> fixture
        extra["sparse_cost"] = float(g["sparse_cost"])
    rbm = make_rbm(g["init_Wt"], g["init_vb"], g["init_hb"], vis_gauss, hid_gauss, bunch, lr, mmt, wc, **extra)
    cache = (cache // bunch) * bunch
    c = make_cache(cache, bunch)
    it = iter(utterances(g))
    pending = next(it, None)
    nb = 0
    while pending is not None:
        while not c.full() and pending is not None:
            c.add(pending[0], np.zeros((pending[0].shape[0], 1), np.float32))   # "fake the labels" TRbmCu.cc:313
            pending = next(it, None)
        c.randomize()
        while not c.empty():
            Fb, _ = c.get_bunch()
            rbm.cd1(Fb)
            nb += 1
    return rbm, nb


def rnn_layers(g, prefix="init"):
    ctx, bptt, nin, H, n_out = [int(v) for v in g["cfg"]]
    return [("recurrent", g[prefix + "_Wr"], g[prefix + "_br"], nin), ("affine", g[prefix + "_Wo"], g[prefix + "_bo"]), ("softmax", n_out)]


def rnn_utterances(g):
    ctx = int(g["cfg"][0])
    pos = 0
    for T in g["lengths"]:
        yield F.splice(g["feats"][pos:pos + T], ctx), g["labels"][pos:pos + T].astype(np.int32)
        pos += T

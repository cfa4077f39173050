"""Replay of the TNetCu main loop (src/TNetCu.cc:376-441) over a golden fixture, parameterised by the
backend (oracle or CUDA path): fill cache utterance by utterance, Randomize, train bunch by bunch."""
import numpy as np
from tnet_b200 import formats as F


def fixture_layers(g, prefix="init"):
    dims = list(g["dims"])
    layers = []
    for k in range(len(dims) - 1):
        layers.append(("affine", g["%s_Wt%d" % (prefix, k)], g["%s_b%d" % (prefix, k)]))
        layers.append(("softmax" if k == len(dims) - 2 else "sigmoid", int(dims[k + 1])))
    return layers


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
    net.set_hyper(lr, mmt=mmt, wc=wc, gdf=bool(gdf))
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

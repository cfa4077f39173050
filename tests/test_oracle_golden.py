"""Pins oracle/tnet_oracle.c against the reference's own CPU trainer (oracle/_ref/TNet = unmodified
src/TNet.cc + TNetLib + KaldiLib, run by tests/golden/make_golden.py --impl cpu) and, when the GPU goldens
are present, against the reference's GPU trainer TNetCu run on a B200."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O
from replay import compare_layer, fixture_layers, replay_mlp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(g, acc_double, wtol, etol):
    net, nb, perms = replay_mlp(g, lambda L: O.Net(L, acc_double=acc_double), O.Cache,
                                lambda s: O.lib.orc_srand48(s))
    err, frames, correct = net.stats()
    assert frames == int(g["ref_frames"])                       # bunch slicing / discard rule exact
    # report line prints 6 significant digits
    assert abs(err - float(g["ref_err"])) <= etol * abs(float(g["ref_err"]))
    ref_correct = float(g["ref_correct_pct"]) * frames / 100.0
    assert abs(correct - ref_correct) <= max(1.0, 0.003 * frames)
    final = fixture_layers(g, "final")
    assert len(final) == len(net.layers)
    for i, L in enumerate(net.layers):
        compare_layer(net.get_layer(i), final[i], wtol)
    return correct, ref_correct


@pytest.mark.parametrize("case", ["mlp_small", "mlp_norand_wc", "net_shared"])
@pytest.mark.parametrize("acc_double", [0, 1])
def test_oracle_vs_reference_cpu_tnet(case, acc_double):
    g = np.load(os.path.join(GOLD, "cpu_%s.npz" % case))
    correct, ref_correct = _check(g, acc_double, wtol=2e-5, etol=2e-5)
    assert correct == round(ref_correct)      # frame-accuracy count exact on these fixtures


def test_oracle_cross_validation_vs_reference_cpu_tnet():
    """--CROSSVALIDATE=TRUE (TNet.cc:344 / TNetCu.cc:437): forward + objective only — same report, network untouched."""
    g = np.load(os.path.join(GOLD, "cpu_opt_cv.npz"))
    assert int(g["cv"]) == 1
    correct, ref_correct = _check(g, 1, wtol=1e-7, etol=2e-5)       # "final" network of the fixture = the initial one
    assert correct == round(ref_correct)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "gpu_opt_*.npz"))) or [None])
def test_oracle_options_vs_reference_gpu_tnetcu(path):
    """per-layer learning-rate factors with a frozen first layer (gpu_opt_lrfactors.npz: only TNetCu takes --LEARNRATEFACTORS)"""
    if path is None:
        pytest.skip("generate with tests/golden/make_golden.py --impl gpu --only opt_ on a B200")
    _check(np.load(path), 0, wtol=5e-5, etol=5e-5)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "gpu_mlp_*.npz")) + glob.glob(os.path.join(GOLD, "gpu_net_*.npz"))) or [None])
def test_oracle_vs_reference_gpu_tnetcu(path):
    if path is None:
        pytest.skip("GPU goldens not generated yet (tests/golden/make_golden.py --impl gpu on a B200)")
    g = np.load(path)
    _check(g, 0, wtol=5e-5, etol=5e-5)


@pytest.mark.parametrize("case", ["feacat_post", "feacat_logpost"])
def test_oracle_forward_vs_reference_cpu_tfeacat(case):
    """Forward-only pin: the oracle's Propagate on the golden inputs == what the unmodified reference CPU tool TFeaCat wrote
    (posteriors, or log-posteriors with --LOGPOSTERIOR=TRUE), utterance by utterance."""
    from replay import fixture_layers, utterances
    g = np.load(os.path.join(GOLD, "cpu_%s.npz" % case))
    net = O.Net(fixture_layers(g), acc_double=1)
    got = np.concatenate([net.propagate(f) for f, _ in utterances(g)])
    if int(g["cfg"][1]):
        got = np.log(got.astype(np.float64)).astype(np.float32)      # TFeaCat.cc: static_cast<BaseFloat>(log(x))
        np.testing.assert_allclose(got, g["ref_out"], rtol=2e-5, atol=2e-5)
    else:
        np.testing.assert_allclose(got, g["ref_out"], rtol=2e-5, atol=1e-9)


def test_tnorm_restatement_vs_reference_cpu_tnorm():
    """Pins the normalisation estimator's arithmetic (TNorm.cc / TNormCu.cc:268-328) against the unmodified reference CPU tool:
    float products accumulated in double over the TRIMMED spliced frames, divided by a frame count that INCLUDES the replicated
    extension rows (TNormCu.cc:292), bias = -mean, window = 1/sqrt(E[x^2] - mean^2)."""
    from replay import utterances
    g = np.load(os.path.join(GOLD, "cpu_norm_splice.npz"))
    ctx = int(g["cfg"][0])
    first = second = 0.0
    frames = 0
    for f, _ in utterances(g):
        first = first + f.astype(np.float64).sum(0)
        second = second + (f * f).astype(np.float64).sum(0)       # float product, double accumulation
        frames += f.shape[0] + 2 * ctx
    assert frames == int(g["ref_frames"])
    mean = first / frames
    var = second / frames - mean * mean
    np.testing.assert_allclose(-mean, g["ref_bias"], rtol=1e-5, atol=1e-6)      # the tool prints 6 significant digits
    np.testing.assert_allclose(1.0 / np.sqrt(var), g["ref_window"], rtol=1e-5)


CACHE_CASES = {
    # name: (cachesize, bunchsize, seed, randomize, lens) — ragged utterances, leftovers carried into the next fill, a partial last
    # fill, utterances longer than the free space, a discarded tail; the leftover never exceeds the cache (the reference asserts)
    "ragged_rand": (192, 32, 77, 1, [50, 7, 130, 64, 1, 99, 150, 12, 45, 33, 170]),
    "norand": (128, 64, 5, 0, [100, 100, 30, 64, 2, 60]),
    "exact_fit": (96, 32, 9, 1, [32, 64, 96, 16, 16, 64]),
    "tiny_bunch": (40, 8, 123, 1, [3, 5, 8, 13, 21, 34, 2, 1, 39]),
}


def _cache_inputs(name):
    cachesize, bunch, seed, rnd, lens = CACHE_CASES[name]
    r = np.random.default_rng(len(name) + seed)
    fdim, ddim = 11, 5
    seqs = [(r.standard_normal((n, fdim)).astype(np.float32), r.standard_normal((n, ddim)).astype(np.float32)) for n in lens]
    return cachesize, bunch, seed, rnd, fdim, ddim, seqs


def _oracle_cache_run(name):
    cachesize, bunch, seed, rnd, fdim, ddim, seqs = _cache_inputs(name)
    O.lib.orc_srand48(seed)
    c = O.Cache(cachesize, bunch)
    out, i = [], 0
    while i < len(seqs):
        while not c.full() and i < len(seqs):
            c.add(*seqs[i]); i += 1
        if rnd:
            c.randomize()
        while not c.empty():
            out.append(c.get_bunch())
    return out, c.discarded()


@pytest.mark.parametrize("name", sorted(CACHE_CASES))
def test_oracle_cache_vs_reference_cpu_cache(name, tmp_path):
    """The oracle's frame cache (orc_cache_*: fill, leftover carry-over, random_shuffle + lrand48 permutation, bunch slicing, discarded
    tail) against the reference's own CPU cache (TNetLib/Cache.cc, the same state machine as cuCache.cc) driven by
    oracle/ref_tools/cache_dump.cc: every bunch bit for bit — committed dump, and live where oracle/_ref/RefCacheDump exists."""
    got, disc = _oracle_cache_run(name)
    flat = np.concatenate([np.concatenate([f.ravel(), d.ravel()]) for f, d in got]) if got else np.zeros(0, np.float32)
    gold = np.load(os.path.join(GOLD, "cpu_cache_dump.npz"))
    assert int(gold[name + "_nb"]) == len(got) and int(gold[name + "_discarded"]) == disc
    assert np.array_equal(flat.view(np.uint32), gold[name + "_data"].view(np.uint32))
    exe = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "RefCacheDump")
    if os.path.exists(exe):
        nb, data, rdisc = run_reference_cache(exe, name, str(tmp_path))
        assert nb == len(got) and rdisc == disc and np.array_equal(data.view(np.uint32), flat.view(np.uint32))


def run_reference_cache(exe, name, d):
    import subprocess
    cachesize, bunch, seed, rnd, fdim, ddim, seqs = _cache_inputs(name)
    src, out = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
    with open(src, "wb") as f:
        f.write(np.array([cachesize, bunch, seed, rnd, fdim, ddim, len(seqs)], np.int32).tobytes())
        for F_, D_ in seqs:
            f.write(np.array([F_.shape[0]], np.int32).tobytes() + F_.tobytes() + D_.tobytes())
    subprocess.check_call([exe, src, out])
    b = open(out, "rb").read()
    nb = int(np.frombuffer(b, np.int32, 1)[0])
    data = np.frombuffer(b, np.float32, nb * bunch * (fdim + ddim), 4).copy()
    disc = int(np.frombuffer(b, np.int32, 1, 4 + 4 * data.size)[0])
    return nb, data, disc

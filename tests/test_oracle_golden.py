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

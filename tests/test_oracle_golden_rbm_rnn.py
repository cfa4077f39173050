"""Pins the RBM / recurrent parts of the oracle against the reference's TRbmCu and TRecurrentCu run on a B200
(fixtures tests/golden/gpu_rbm_*.npz, gpu_rnn_small.npz; generating script tests/golden/make_golden.py --impl gpu)."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O
from replay import replay_rbm, rnn_layers, rnn_utterances

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _Rbm:
    """oracle RBM with TRbmCu's generator seeding order."""

    def __init__(self, Wt, vb, hb, vis_gauss, hid_gauss, bunch, lr, mmt, wc, sparse_cost=None):
        self.r = O.Rbm(Wt, vb, hb, vis_gauss, hid_gauss, lr, mmt, wc, acc_double=0)
        if sparse_cost is not None:
            self.r.set_sparse(sparse_cost)
        self.z = [np.empty((bunch, Wt.shape[0]), np.uint32) for _ in range(4)]
        O.lib.orc_rand_seed(O.P(self.z[0]), O.P(self.z[1]), O.P(self.z[2]), O.P(self.z[3]), bunch, Wt.shape[0], Wt.shape[0])

    def cd1(self, v):
        self.r.cd1(v, self.z)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "gpu_rbm_*.npz"))))
def test_oracle_rbm_vs_reference_trbmcu(path):
    g = np.load(path)
    rbm, nb = replay_rbm(g, _Rbm, O.Cache, lambda s: O.lib.orc_srand48(s))
    err, frames = rbm.r.stats()
    assert frames == int(g["ref_frames"])
    assert abs(err - float(g["ref_err"])) <= 5e-5 * abs(float(g["ref_err"]))
    Wt, vb, hb = rbm.r.get()
    np.testing.assert_allclose(Wt, g["final_Wt"], rtol=5e-5, atol=5e-5 * np.abs(g["final_Wt"]).max())
    np.testing.assert_allclose(vb, g["final_vb"], rtol=5e-5, atol=5e-5 * max(1e-3, np.abs(g["final_vb"]).max()))
    np.testing.assert_allclose(hb, g["final_hb"], rtol=5e-5, atol=5e-5 * max(1e-3, np.abs(g["final_hb"]).max()))


def test_oracle_rnn_vs_reference_trecurrentcu():
    g = np.load(os.path.join(GOLD, "gpu_rnn_small.npz"))
    ctx, bptt, nin, H, n_out = [int(v) for v in g["cfg"]]
    lr = float(g["hyper"][0])
    L = rnn_layers(g)
    rnn = O.Rnn(L[0][1], L[0][2], nin, bptt, lr)
    Wo, bo = O.f32(L[1][1]).copy(), O.f32(L[1][2]).copy()          # [n_out x H] on-disk layout
    cW, cb = np.zeros((H, n_out), np.float32), np.zeros(n_out, np.float32)
    W = np.ascontiguousarray(Wo.T)                                  # in-memory [H x n_out]
    st = O.ObjStats(0.0, 0, 0)
    for x, lab in rnn_utterances(g):
        rnn.clear()
        for t in range(x.shape[0]):
            h = rnn.propagate(x[t]).reshape(1, H)
            a = np.zeros((1, n_out), np.float32)     # Init()-zeroed like the component's output: the bias pre-copy computes 0*old
            O.lib.orc_affine_fwd(O.P(h), H, O.P(W), n_out, O.P(bo), O.P(a), n_out, 1, H, n_out, 0)
            y = O.softmax(a)
            tgt = np.zeros((1, n_out), np.float32); tgt[0, lab[t]] = 1
            e = np.empty_like(y)
            O.lib.orc_xent_evaluate(O.P(y), O.P(tgt), O.P(e), 1, n_out, n_out, O.C.byref(st))
            # CuNetwork::Backpropagate: softmax (copy) ; affine Backpropagate + Update ; recurrent = stopper: Update only
            eh = np.empty((1, H), np.float32)
            O.lib.orc_affine_bwd(O.P(e), n_out, O.P(W), n_out, O.P(eh), H, 1, H, n_out, 0)
            O.lib.orc_affine_update(O.P(h), H, O.P(e), n_out, O.P(W), n_out, O.P(bo), O.P(cW), n_out, O.P(cb), 1, H, n_out,
                                    O.cf(lr), O.cf(0.0), O.cf(0.0), 1, 0)
            rnn.update(eh[0])
    assert st.frames == int(g["ref_frames"])
    assert abs(st.error - float(g["ref_err"])) <= 1e-4 * abs(float(g["ref_err"]))
    Wr, br = rnn.get()
    np.testing.assert_allclose(Wr, g["final_Wr"], rtol=1e-4, atol=1e-4 * np.abs(g["final_Wr"]).max())
    np.testing.assert_allclose(br, g["final_br"], rtol=1e-4, atol=1e-4 * max(1e-3, np.abs(g["final_br"]).max()))
    np.testing.assert_allclose(W.T, g["final_Wo"], rtol=1e-4, atol=1e-4 * np.abs(g["final_Wo"]).max())

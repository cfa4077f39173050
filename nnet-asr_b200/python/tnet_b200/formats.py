"""File formats of the TNet trainers (reference: troylee/nnet-asr, paths under src/).

Writers/readers used by the tests, the bench and the golden-vector scripts so that
the reference binaries and this repo's trainers consume *identical files*:

* network text format  — CuNetwork::ReadNetwork/WriteNetwork, CuTNetLib/cuNetwork.cc:213-387;
  matrices `m rows cols` / vectors `v dim` — KaldiLib/Matrix.tcc:522-600, Vector.tcc:527-571;
  <biasedlinearity> stores W transposed [out x in] then bias — cuBiasedLinearity.cc:70-119
* HTK feature files    — 12-byte big-endian header + float32 frames, KaldiLib/Features.cc
* MLF label files + output label map — KaldiLib/Labels.cc:44-227
* SCP lists            — one physical file per line
"""
import io
import os
import struct
import numpy as np


# --------------------------------------------------------------------------- network text
def _fmt_row(row):
    return " ".join(repr(float(np.float32(v))) if False else "%.9g" % float(v) for v in row)


def write_matrix(f, M):
    M = np.asarray(M, dtype=np.float32)
    f.write("m %d %d\n" % M.shape)
    for r in M:
        f.write(_fmt_row(r) + " \n")


def write_vector(f, v, fmt="%.9g"):
    v = np.asarray(v)
    f.write("v %d  " % v.shape[0])
    f.write(" ".join(fmt % x for x in v) + " \n")


def write_mlp(path, layers):
    """layers: list of ('affine', Wt[out x in], b) | ('sigmoid', n) | ('softmax', n)
    | ('expand', dim_in, offsets) | ('bias', v) | ('window', v)
    | ('rbm', vistype, hidtype, Wt[hid x vis], visbias, hidbias) | ('recurrent', Wt[out x (in+out)], b, nin)
    | ('shared', ninst, Wt[out/ninst x in/ninst], b[out/ninst])      <sharedlinearity>, cuSharedLinearity.cc:98-178
    | ('discrete', [Wt_i[out_i x in_i], ...], b[sum out_i])           <discretelinearity>, cuDiscreteLinearity.cc:80-157
    | ('rbmsparse', vistype, hidtype, Wt, visbias, hidbias, sparsity_cost)   <rbmsparse>, cuRbmSparse.cc:171-236"""
    with open(path, "w") as f:
        for L in layers:
            kind = L[0]
            if kind == "affine":
                Wt, b = L[1], L[2]
                f.write("<biasedlinearity> %d %d\n" % (Wt.shape[0], Wt.shape[1]))
                write_matrix(f, Wt)
                write_vector(f, b)
                f.write("\n")
            elif kind in ("sigmoid", "softmax"):
                f.write("<%s> %d %d\n" % (kind, L[1], L[1]))
            elif kind == "expand":
                dim_in, offs = L[1], list(L[2])
                f.write("<expand> %d %d\n" % (dim_in * len(offs), dim_in))
                write_vector(f, np.asarray(offs, dtype=np.int64), fmt="%d")
                f.write("\n")
            elif kind in ("bias", "window"):
                v = np.asarray(L[1], dtype=np.float32)
                f.write("<%s> %d %d\n" % (kind, v.shape[0], v.shape[0]))
                write_vector(f, v)
                f.write("\n")
            elif kind == "rbm":
                _, vt, ht, Wt, vb, hb = L
                f.write("<rbm> %d %d\n" % (Wt.shape[0], Wt.shape[1]))
                f.write(" %s  %s\n" % (vt, ht))
                write_matrix(f, Wt)
                write_vector(f, vb)
                f.write("\n")
                write_vector(f, hb)
                f.write("\n")
            elif kind == "recurrent":
                _, Wt, b, nin = L
                f.write("<recurrent> %d %d\n" % (Wt.shape[0], nin))
                write_matrix(f, Wt)
                write_vector(f, b)
                f.write("\n")
            elif kind == "shared":
                _, ninst, Wt, b = L
                f.write("<sharedlinearity> %d %d\n%d\n" % (Wt.shape[0] * ninst, Wt.shape[1] * ninst, ninst))
                write_matrix(f, Wt)
                write_vector(f, b)
                f.write("\n")
            elif kind == "discrete":
                _, blocks, b = L
                f.write("<discretelinearity> %d %d\n%d\n" % (sum(B.shape[0] for B in blocks), sum(B.shape[1] for B in blocks), len(blocks)))
                for B in blocks:
                    write_matrix(f, B)
                write_vector(f, b)
                f.write("\n")
            elif kind == "rbmsparse":
                _, vt, ht, Wt, vb, hb, cost = L
                f.write("<rbmsparse> %d %d\n" % (Wt.shape[0], Wt.shape[1]))
                f.write(" %s  %s\n" % (vt, ht))
                write_matrix(f, Wt)
                write_vector(f, vb)
                f.write("\n")
                write_vector(f, hb)
                f.write("\n%.9g\n" % cost)
            else:
                raise ValueError(kind)


class _Tok:
    def __init__(self, text):
        self.t = text.split()
        self.i = 0

    def next(self):
        v = self.t[self.i]
        self.i += 1
        return v

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None


def _read_matrix(tk):
    assert tk.next() == "m"
    r, c = int(tk.next()), int(tk.next())
    a = np.array(tk.t[tk.i:tk.i + r * c], dtype=np.float64).astype(np.float32).reshape(r, c)
    tk.i += r * c
    return a


def _read_vector(tk, dtype=np.float32):
    assert tk.next() == "v"
    n = int(tk.next())
    a = np.array(tk.t[tk.i:tk.i + n], dtype=np.float64).astype(dtype)
    tk.i += n
    return a


def read_mlp(path):
    """Parse a network text file back into the `layers` structure of write_mlp."""
    return read_mlp_text(open(path).read())


def layer_dims(L):
    """(n_inputs, n_outputs) of one entry of the `layers` structure."""
    k = L[0]
    if k == "affine":
        return L[1].shape[1], L[1].shape[0]
    if k == "shared":
        return L[2].shape[1] * L[1], L[2].shape[0] * L[1]
    if k == "discrete":
        return sum(B.shape[1] for B in L[1]), sum(B.shape[0] for B in L[1])
    if k in ("rbm", "rbmsparse"):
        return L[3].shape[1], L[3].shape[0]
    if k == "recurrent":
        return L[3], L[1].shape[0]
    if k == "expand":
        return L[1], L[1] * len(L[2])
    if k in ("bias", "window"):
        return len(L[1]), len(L[1])
    return L[1], L[1]


def read_mlp_text(text):
    """The same from the file's contents (the golden fixtures keep whole network files as strings)."""
    tk = _Tok(text)
    layers = []
    while tk.peek() is not None:
        tag = tk.next().lower()
        if tag == "<endblock>":
            break
        nout, nin = int(tk.next()), int(tk.next())
        if tag == "<biasedlinearity>":
            Wt = _read_matrix(tk)
            b = _read_vector(tk)
            assert Wt.shape == (nout, nin)
            layers.append(("affine", Wt, b))
        elif tag in ("<sigmoid>", "<softmax>"):
            layers.append((tag[1:-1], nout))
        elif tag == "<expand>":
            layers.append(("expand", nin, _read_vector(tk, np.int32)))
        elif tag in ("<bias>", "<window>"):
            layers.append((tag[1:-1], _read_vector(tk)))
        elif tag == "<rbm>":
            vt, ht = tk.next(), tk.next()
            Wt = _read_matrix(tk)
            vb = _read_vector(tk)
            hb = _read_vector(tk)
            layers.append(("rbm", vt, ht, Wt, vb, hb))
        elif tag == "<recurrent>":
            Wt = _read_matrix(tk)
            b = _read_vector(tk)
            layers.append(("recurrent", Wt, b, nin))
        elif tag == "<sharedlinearity>":
            ninst = int(tk.next())
            Wt = _read_matrix(tk)
            b = _read_vector(tk)
            assert Wt.shape == (nout // ninst, nin // ninst)
            layers.append(("shared", ninst, Wt, b))
        elif tag == "<discretelinearity>":
            nblocks = int(tk.next())
            blocks = [_read_matrix(tk) for _ in range(nblocks)]
            layers.append(("discrete", blocks, _read_vector(tk)))
        elif tag == "<rbmsparse>":
            vt, ht = tk.next(), tk.next()
            Wt = _read_matrix(tk)
            vb = _read_vector(tk)
            hb = _read_vector(tk)
            layers.append(("rbmsparse", vt, ht, Wt, vb, hb, float(tk.next())))
        else:
            raise ValueError("unsupported tag " + tag)
    return layers


# --------------------------------------------------------------------------- HTK / MLF / SCP
HTK_USER = 9


def write_htk(path, feats, samp_period=100000, parm_kind=HTK_USER):
    feats = np.asarray(feats, dtype=np.float32)
    n, d = feats.shape
    with open(path, "wb") as f:
        f.write(struct.pack(">iihh", n, samp_period, 4 * d, parm_kind))
        f.write(feats.astype(">f4").tobytes())


def read_htk(path):
    d = open(path, "rb").read()
    n, period, size, kind = struct.unpack(">iihh", d[:12])
    a = np.frombuffer(d[12:12 + n * size], dtype=">f4").astype(np.float32).reshape(n, size // 4)
    return a, period, kind


def write_label_map(path, n_out):
    tags = ["s%d" % i for i in range(n_out)]
    with open(path, "w") as f:
        f.write("\n".join(tags) + "\n")
    return tags


def write_mlf(path, utt_labels, tags, samp_period=100000):
    """utt_labels: dict name -> int array of per-frame class ids (runs become segments)."""
    with open(path, "w") as f:
        f.write("#!MLF!#\n")
        for name, lab in utt_labels.items():
            f.write('"*/%s.lab"\n' % name)
            lab = np.asarray(lab)
            start = 0
            for t in range(1, len(lab) + 1):
                if t == len(lab) or lab[t] != lab[start]:
                    f.write("%d %d %s\n" % (start * samp_period, t * samp_period, tags[int(lab[start])]))
                    start = t
            f.write(".\n")


def write_scp(path, files):
    with open(path, "w") as f:
        f.write("\n".join(files) + "\n")


# --------------------------------------------------------------------------- synthetic sets
def gen_mlp_init(dims, rng, negbias=True):
    """tools/init/gen_mlp_init.py:36-68 with --gauss [--negbias]:
    W ~ 0.1*N(0,1), hidden bias ~ U[-4.1,-3.9], output bias 0; sigmoid hidden, softmax out."""
    layers = []
    for l in range(len(dims) - 1):
        nin, nout = dims[l], dims[l + 1]
        Wt = (0.1 * rng.standard_normal((nout, nin))).astype(np.float32)
        last = l == len(dims) - 2
        if last or not negbias:
            b = np.zeros(nout, dtype=np.float32)
        else:
            b = (rng.random(nout) / 5.0 - 4.1).astype(np.float32)
        layers.append(("affine", Wt, b))
        layers.append(("softmax" if last else "sigmoid", nout))
    return layers


def gen_utterances(n_utt, n_frames, raw_dim, n_out, rng, min_run=3, max_run=10, vary_len=True):
    """SURVEY §8d synthetic inputs: features ~ N(0,1); labels constant over runs of 3-10 frames."""
    utts = {}
    for u in range(n_utt):
        T = int(n_frames if not vary_len else rng.integers(max(8, n_frames // 2), n_frames + 1))
        x = rng.standard_normal((T, raw_dim)).astype(np.float32)
        lab = np.empty(T, dtype=np.int32)
        t = 0
        while t < T:
            run = int(rng.integers(min_run, max_run + 1))
            lab[t:t + run] = int(rng.integers(0, n_out))
            t += run
        utts["utt%04d" % u] = (x, lab)
    return utts


def write_dataset(dirpath, utts, n_out, context, samp_period=100000):
    """Write features/*.fea, train.scp, train.mlf, labelmap, expand transform.
    Returns dict of paths."""
    os.makedirs(os.path.join(dirpath, "features"), exist_ok=True)
    files = []
    for name, (x, _) in utts.items():
        p = os.path.join(dirpath, "features", name + ".fea")
        write_htk(p, x, samp_period)
        files.append(p)
    scp = os.path.join(dirpath, "train.scp")
    write_scp(scp, files)
    lmap = os.path.join(dirpath, "labelmap")
    tags = write_label_map(lmap, n_out)
    mlf = os.path.join(dirpath, "train.mlf")
    write_mlf(mlf, {k: v[1] for k, v in utts.items()}, tags, samp_period)
    raw_dim = next(iter(utts.values()))[0].shape[1]
    tr = os.path.join(dirpath, "expand.transf")
    write_mlp(tr, [("expand", raw_dim, list(range(-context, context + 1)))])
    return dict(scp=scp, mlf=mlf, labelmap=lmap, transform=tr, files=files)


def splice(x, context):
    """Reader frame replication (STARTFRMEXT/ENDFRMEXT, KaldiLib/Features.cc:776-849) followed by
    <expand> (cukernels.cu:349-361) and the trim of TNetCu.cc:391-393, on one utterance."""
    T, D = x.shape
    idx = np.clip(np.arange(T)[:, None] + np.arange(-context, context + 1)[None, :], 0, T - 1)
    return x[idx].reshape(T, (2 * context + 1) * D)

#!/usr/bin/env python
"""Golden for BASELINE configs[0] on the reference's OWN example data: examples/01test_MLP3_compare_multithread_cuda_decode_phn
(23-dim features, the real `Hamm_dct_norm` transform chain <expand 51> <transpose> <window> <blocklinearity> <bias> <window>,
598-1024-135 MLP, bunch 960, cache 14400, lr 0.008, seed 123 — run_test.CPU.sh / run_test.GPU.sh).

  python tests/golden/make_example01.py          # here: needs /root/reference (example data) and oracle/_ref/TNet (reference CPU trainer)

Stores tests/golden/cpu_example01.npz: the first N_UTT utterances of the example (features, their MLF entries, label map, transform
file, a seeded initial network per tools/init/gen_mlp_init.py --gauss --negbias) and what the unmodified reference CPU trainer
TNet --THREADS=1 produced from them (final network, Xent, frames, accuracy).  The GPU tests rebuild the files from the fixture in
a temporary directory; nothing reads /root/reference at test time."""
import os, re, struct, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import formats as F  # noqa: E402

EX = "/root/reference/examples/01test_MLP3_compare_multithread_cuda_decode_phn"
N_UTT = 30
OUT = os.path.join(ROOT, "tests", "golden", "cpu_example01.npz")
FLAGS = dict(lr=0.008, bunch=960, cache=14400, seed=123, ext=25)


def materialize(g, d):
    """write the fixture's files into directory d; returns paths (used by the generator and by the tests)"""
    os.makedirs(os.path.join(d, "features"), exist_ok=True)
    files, pos = [], 0
    for name, n in zip(g["names"], g["lengths"]):
        p = os.path.join(d, "features", str(name) + ".fea")
        x = g["feats"][pos:pos + n]
        pos += n
        with open(p, "wb") as f:
            f.write(struct.pack(">iihh", int(n), int(g["samp_period"]), 4 * x.shape[1], int(g["parm_kind"])))
            f.write(x.astype(">f4").tobytes())
        files.append(p)
    paths = dict(files=files, scp=os.path.join(d, "test.scp"), mlf=os.path.join(d, "test.mlf"), labelmap=os.path.join(d, "labelmap"),
                 transform=os.path.join(d, "Hamm_dct_norm"), init=os.path.join(d, "init.nnet"))
    open(paths["scp"], "w").write("\n".join(files) + "\n")
    open(paths["mlf"], "wb").write(bytes(g["mlf"]))
    open(paths["labelmap"], "wb").write(bytes(g["labelmap"]))
    open(paths["transform"], "wb").write(bytes(g["transform"]))
    F.write_mlp(paths["init"], init_layers())
    return paths


def init_layers():
    """the seeded initial network (tools/init/gen_mlp_init.py --dim=598:1024:135 --gauss --negbias); regenerated, not stored: 2.4 MB"""
    return F.gen_mlp_init([598, 1024, 135], np.random.default_rng(20240607))


def command(exe, paths, final, gpu):
    cmd = [exe, "-H", paths["init"], "-I", paths["mlf"], "-L", "*/", "-X", "lab", "-S", paths["scp"], "-m", paths["labelmap"], "-n", repr(FLAGS["lr"]),
           "--TARGETMMF=" + final, "--BUNCHSIZE=%d" % FLAGS["bunch"], "--CACHESIZE=%d" % FLAGS["cache"], "--RANDOMIZE=TRUE", "--SEED=%d" % FLAGS["seed"],
           "--FEATURETRANSFORM=" + paths["transform"], "--STARTFRMEXT=%d" % FLAGS["ext"], "--ENDFRMEXT=%d" % FLAGS["ext"]]
    return cmd + (["--GRADDIVFRM=FALSE"] if gpu else ["--THREADS=1"])   # the example's own GPU/CPU pair of flags


def parse_report(txt):
    m = re.search(r"Xent:(\S+) frames:(\d+) err/frm:(\S+) correct\[(\S+)%\]", txt)
    if not m:
        raise RuntimeError("no report line in:\n" + txt[-2000:])
    return dict(err=float(m.group(1)), frames=int(m.group(2)), correct_pct=float(m.group(4)))


def run(exe, g, gpu):
    with tempfile.TemporaryDirectory() as d:
        paths = materialize(g, d)
        final = os.path.join(d, "final.nnet")
        r = subprocess.run(command(exe, paths, final, gpu), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("trainer failed:\n" + r.stdout[-3000:])
        return parse_report(r.stdout), F.read_mlp(final), r.stdout


def main():
    scp = [l.strip() for l in open(os.path.join(EX, "lib", "test.scp")) if l.strip()][:N_UTT]
    names, feats, lengths = [], [], []
    for rel in scp:
        raw = open(os.path.join(EX, rel), "rb").read()
        n, period, size, kind = struct.unpack(">iihh", raw[:12])
        feats.append(np.frombuffer(raw[12:12 + n * size], dtype=">f4").astype(np.float32).reshape(n, size // 4))
        names.append(os.path.splitext(os.path.basename(rel))[0])
        lengths.append(n)
    # MLF entries of the selected utterances only
    mlf_all = open(os.path.join(EX, "lib", "test_3s.mlf")).read()
    blocks = re.split(r'(?m)^(?=")', mlf_all)
    keep = [blocks[0]] + [b for b in blocks[1:] if re.match(r'"\*/(\w+)\.lab"', b).group(1) in set(names)]
    g = dict(names=np.array(names), feats=np.concatenate(feats), lengths=np.array(lengths, np.int32), samp_period=np.int64(period), parm_kind=np.int64(kind),
             mlf=np.frombuffer("".join(keep).encode(), np.uint8), labelmap=np.frombuffer(open(os.path.join(EX, "lib", "mono_state_phn_set_135_phn"), "rb").read(), np.uint8),
             transform=np.frombuffer(open(os.path.join(EX, "lib", "Hamm_dct_norm"), "rb").read(), np.uint8))
    rep, out_layers, txt = run(os.path.join(ROOT, "oracle", "_ref", "TNet"), g, gpu=False)
    # every 8th row of the first layer's final weights (the full matrix would be 2.4 MB), the whole second layer
    g.update(final_Wt0_rows8=out_layers[0][1][::8], final_b0=out_layers[0][2], final_Wt1=out_layers[2][1], final_b1=out_layers[2][2],
             ref_err=np.float64(rep["err"]), ref_frames=np.int64(rep["frames"]), ref_correct_pct=np.float64(rep["correct_pct"]))
    np.savez_compressed(OUT, **g)
    print("example 01, first %d utterances (%d frames): reference CPU TNet %s -> %s (%.1f KB)" % (N_UTT, sum(lengths), rep, OUT, os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()

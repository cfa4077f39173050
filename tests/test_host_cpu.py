"""CPU-side tests of the host logic that needs no device: option grammar of the drop-in binaries (they must fail before
touching the GPU), the text formats, and the data-parallel arithmetic (2 gloo ranks against the single-process oracle)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from tnet_b200 import formats as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "nnet-asr_b200", "bin")


def _run(args):
    return subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


@pytest.mark.parametrize("exe", ["TNetCu", "TRbmCu", "TRecurrentCu"])
def test_binaries_exist_and_reject_bad_flags(exe):
    p = os.path.join(BIN, exe)
    assert os.path.exists(p), "run __graft_entry__.build()"
    r = _run([p, "-Z", "1"])
    assert r.returncode == 1 and "Invalid command line option '-Z'" in r.stderr
    r = _run([p, "--SEED"])
    assert r.returncode == 1 and "Character '=' expected" in r.stderr
    r = _run([p, "--RANDOMIZE=MAYBE", "-H", "x"])
    assert r.returncode == 1 and "TRUE or FALSE expected" in r.stderr


def test_unused_long_option_is_an_error():
    # reference: ui.CheckCommandLineParamUse() (TNetCu.cc:261)
    r = _run([os.path.join(BIN, "TNetCu"), "--NOSUCHPARAM=3", "-H", "x"])
    assert r.returncode == 1 and "NOSUCHPARAM" in r.stderr


def test_no_cpu_fallback_in_binaries(tmp_path):
    """With valid options and files but no GPU the trainer must fail loudly, not train on the CPU."""
    import ctypes as C
    from tnet_b200 import abi
    n = C.c_int(0)
    if abi.lib().tnb_device_count(C.byref(n)) == abi.OK and n.value > 0:
        pytest.skip("a GPU is visible")
    r = np.random.default_rng(0)
    net = str(tmp_path / "a.nnet")
    F.write_mlp(net, F.gen_mlp_init([6, 4, 3], r))
    res = _run([os.path.join(BIN, "TNetCu"), "-H", net, "-S", "/dev/null", "-I", "/dev/null", "-m", "/dev/null"])
    assert res.returncode == 1 and "CUDA" in res.stderr


def test_network_text_format_roundtrip(tmp_path):
    r = np.random.default_rng(1)
    layers = [("expand", 13, [-2, -1, 0, 1, 2])] + F.gen_mlp_init([65, 8, 4], r) + []
    p = str(tmp_path / "n.nnet")
    F.write_mlp(p, layers)
    back = F.read_mlp(p)
    assert [l[0] for l in back] == ["expand", "affine", "sigmoid", "affine", "softmax"]
    assert np.array_equal(back[1][1], layers[1][1]) and np.array_equal(back[0][2], np.array([-2, -1, 0, 1, 2]))
    x = r.standard_normal((7, 13)).astype(np.float32)
    assert np.array_equal(F.splice(x, 2), O.expand(x, np.arange(-2, 3, dtype=np.int32)))   # away from... incl. edge clamping


def test_htk_roundtrip(tmp_path):
    r = np.random.default_rng(2)
    x = r.standard_normal((11, 39)).astype(np.float32)
    p = str(tmp_path / "a.fea")
    F.write_htk(p, x)
    y, period, kind = F.read_htk(p)
    assert np.array_equal(x, y) and period == 100000 and kind == F.HTK_USER


DP_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import oracle_lib as O
from tnet_b200 import formats as F
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
schedule = sys.argv[3]
dist.init_process_group("gloo")
r = np.random.default_rng(5)
dims, B = [20, 16, 7], 64
layers = F.gen_mlp_init(dims, r)
X = r.standard_normal((B, dims[0])).astype(np.float32)
T = np.zeros((B, dims[-1]), np.float32); T[np.arange(B), r.integers(0, dims[-1], B)] = 1
lr, mmt, wc = 0.1, 0.5, 1e-3
# what CuNetwork::Backpropagate does per rank in data-parallel mode: local rows, gradient, all-reduce, update with N = global rows
rows = slice(rank * B // world, (rank + 1) * B // world)
net = O.Net(layers, acc_double=1)
net.set_hyper(0.0)                       # lr 0: forward/backward only, no update inside the oracle
Xl, Tl = X[rows], T[rows]
W = [O.f32(layers[0][1]).T.copy(), O.f32(layers[2][1]).T.copy()]
b = [O.f32(layers[0][2]).copy(), O.f32(layers[2][2]).copy()]
cW = [np.zeros_like(W[0]), np.zeros_like(W[1])]; cb = [np.zeros_like(b[0]), np.zeros_like(b[1])]
for step in range(2):
    cur = [("affine", W[0].T.copy(), b[0]), ("sigmoid", dims[1]), ("affine", W[1].T.copy(), b[1]), ("softmax", dims[2])]
    n = O.Net(cur, acc_double=1); n.set_hyper(0.0)
    y = n.propagate(Xl)
    h = n.layer_out(1, Xl.shape[0])
    e2, _, _, _ = O.xent_evaluate(y, Tl)
    e1 = O.diff_sigmoid(O.gemm("N", "T", 1.0, e2, W[1], 0.0, np.zeros((Xl.shape[0], dims[1]), np.float32), 1), h)
    grads = [(O.gemm("T", "N", 1.0, Xl, e1, 0.0, np.zeros_like(W[0]), 1), e1.astype(np.float64).sum(0).astype(np.float32)),
             (O.gemm("T", "N", 1.0, h, e2, 0.0, np.zeros_like(W[1]), 1), e2.astype(np.float64).sum(0).astype(np.float32))]
    for k, (gW, gb) in enumerate(grads):
        N = np.float32(B) * np.float32(1.0 / (1.0 - mmt))
        if schedule == "allreduce":      # NCCL schedule: sum everywhere, every rank updates everything
            tW, tb = torch.from_numpy(gW.copy()), torch.from_numpy(gb.copy())
            dist.all_reduce(tW); dist.all_reduce(tb)
            cW[k] = tW.numpy() + np.float32(mmt) * cW[k]; cb[k] = tb.numpy() + np.float32(mmt) * cb[k]
            W[k] = W[k] + np.float32(-lr / N) * cW[k]; b[k] = b[k] + np.float32(-lr / N) * cb[k]
            W[k] = W[k] + np.float32(-lr * wc) * W[k]
        else:
            # peer-memory schedule (csrc/peer.cu): rank r owns rows [r*shard, (r+1)*shard) of the zero-padded matrix; it reads that
            # block of EVERY rank's gradient (all_gather stands in for the peer loads), sums in rank order, updates its block of
            # corrW and W, and stores the new rows into every rank's W (all_gather of the blocks = the peer stores).  The bias row
            # is summed in rank order and updated by every rank for itself.
            nin = gW.shape[0]; pad = -(-nin // world) * world; shard = pad // world
            Gp = np.zeros((pad, gW.shape[1]), np.float32); Gp[:nin] = gW
            allG = [torch.zeros(Gp.shape) for _ in range(world)]; dist.all_gather(allG, torch.from_numpy(Gp))
            allb = [torch.zeros(gb.shape) for _ in range(world)]; dist.all_gather(allb, torch.from_numpy(gb.copy()))
            blk = slice(rank * shard, (rank + 1) * shard)
            g = allG[0].numpy()[blk].copy()
            for q in range(1, world):
                g = g + allG[q].numpy()[blk]
            Wp = np.zeros((pad, gW.shape[1]), np.float32); Wp[:nin] = W[k]
            Kp = np.zeros_like(Wp); Kp[:nin] = cW[k]
            Kp[blk] = g + np.float32(mmt) * Kp[blk]
            wnew = Wp[blk] + np.float32(-lr / N) * Kp[blk]
            wnew = wnew + np.float32(-lr * wc) * wnew
            blocks = [torch.zeros(wnew.shape) for _ in range(world)]; dist.all_gather(blocks, torch.from_numpy(wnew.astype(np.float32)))
            W[k] = np.concatenate([t.numpy() for t in blocks])[:nin]
            cW[k] = Kp[:nin]                 # only this rank's block is current — exactly what the kernel leaves behind
            sb = allb[0].numpy().copy()
            for q in range(1, world):
                sb = sb + allb[q].numpy()
            cb[k] = sb + np.float32(mmt) * cb[k]
            b[k] = b[k] + np.float32(-lr / N) * cb[k]
if rank == 0:
    ref = O.Net(layers, acc_double=1); ref.set_hyper(lr, mmt=mmt, wc=wc, gdf=True)
    ref.train_bunch(X, T); ref.train_bunch(X, T)
    for k, i in enumerate((0, 2)):
        Wt, bb = ref.get_affine(i)
        np.testing.assert_allclose(W[k].T, Wt, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(b[k], bb, rtol=2e-5, atol=1e-6)
    print("DP_OK")
dist.destroy_process_group()
"""


@pytest.mark.parametrize("schedule,port", [("allreduce", 29517), ("peer", 29518)])
def test_data_parallel_arithmetic_two_gloo_ranks(tmp_path, schedule, port):
    """N>1 semantics on CPU: rows of the bunch split over 2 ranks, then either the per-layer gradient all-reduce + full update (the
    NCCL schedule) or the peer-memory schedule's block-owner arithmetic (rank-ordered sum of the owner's rows, block update, rows
    handed to every rank), both with the GLOBAL frame count == the single-process oracle on the whole bunch."""
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script), os.path.join(ROOT, "tests"), os.path.join(ROOT, "nnet-asr_b200", "python"),
                        schedule], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-3000:]


def test_fast_text_reader_writer_matches_ostream_format(tmp_path):
    """tests/cpp/test_text_io.cc: the threaded %g writer is byte-identical to the reference's `ostream << float` loop and the
    stream-buffer scanner reads what strtod reads (incl. inf/nan/denormals, truncated and malformed input)."""
    import subprocess
    exe = str(tmp_path / "test_text_io")
    src = os.path.join(ROOT, "tests", "cpp", "test_text_io.cc")
    inc = os.path.join(ROOT, "nnet-asr_b200", "host")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-pthread", "-I", inc, "-o", exe, src])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and "TEXT_IO_OK" in r.stdout, r.stderr


# ------------------------------------------------------------------------------------------------ front end: HTK / script file / MLF readers

@pytest.fixture(scope="module")
def feature_io_exe(tmp_path_factory):
    """tests/cpp/test_feature_io.cc (the drop-in's readers behind the dump format of oracle/ref_tools/io_dump.cc), compiled once"""
    exe = str(tmp_path_factory.mktemp("feature_io") / "test_feature_io")
    # TNB_TEST_SANITIZE=1: the same tests with AddressSanitizer + UBSan under the readers (developer switch)
    san = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("TNB_TEST_SANITIZE") else []
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-pthread"] + san + ["-I", os.path.join(ROOT, "nnet-asr_b200", "host"), "-I",
                           os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_feature_io.cc")])
    return exe

def _io_dataset(d):
    """A small set that exercises the readers: ragged utterances, plain / logical=physical / logical=physical[first,last] script
    entries, label times that need the round-half-up division by the sample period, score columns and comments in the MLF, a
    record longer than its utterance (truncated frames), little- and big-endian reads decided by the caller."""
    r = np.random.default_rng(17)
    utts = F.gen_utterances(5, 40, 7, 6, r)
    paths = F.write_dataset(d, utts, 6, 2)
    names = list(utts.keys())
    fea = dict(zip(names, paths["files"]))
    entries = [fea[names[0]],                                              # plain physical name
               "alias_b.fea=" + fea[names[1]],                             # logical=physical
               "seg_c.fea=" + fea[names[2]] + "[3,17]",                    # a frame range of a longer file
               fea[names[3]],
               "some/dir/alias_e.fea=" + fea[names[4]] + "[0,%d]" % (utts[names[4]][0].shape[0] - 1)]
    scp = os.path.join(d, "mixed.scp")
    open(scp, "w").write("\n".join(entries) + "\n\n")
    tags = ["s%d" % i for i in range(6)]
    P = 100000
    with open(os.path.join(d, "mixed.mlf"), "w") as f:
        f.write("#!MLF!#\n")
        lab = {n: utts[n][1] for n in names}

        def rec(key, ids, jitter=0, extra="", skip=-1):
            f.write('"*/%s.lab"\n' % key)
            start, seg = 0, 0
            for t in range(1, len(ids) + 1):
                if t == len(ids) or ids[t] != ids[start]:
                    # boundaries off by less than half a frame must round to the same frame index
                    b = start * P + (jitter if start else 0)
                    e = t * P + (jitter if t < len(ids) else 0)
                    if seg != skip:
                        f.write("%d %d %s%s\n" % (b, e, tags[int(ids[start])], extra))
                    start, seg = t, seg + 1
            f.write(".\n")
        rec(names[0], lab[names[0]])
        rec("alias_b", lab[names[1]], jitter=-40000)
        rec("seg_c", lab[names[2]][3:18], jitter=+49999, extra=" -12.5 word")
        rec(names[3], np.concatenate([lab[names[3]], lab[names[3]][-1:].repeat(30)]))    # runs past the end of the features
        rec("alias_e", lab[names[4]])
    return scp, os.path.join(d, "mixed.mlf"), paths["labelmap"]


def _parse_io_dump(path):
    b = open(path, "rb").read()
    pos = 0

    def i32():
        nonlocal pos
        v = int(np.frombuffer(b, "<i4", 1, pos)[0]); pos += 4
        return v
    out = []
    for _ in range(i32()):
        n = i32(); name = b[pos:pos + n].decode(); pos += n
        rows, cols, period = i32(), i32(), i32()
        x = np.frombuffer(b, "<f4", rows * cols, pos).reshape(rows, cols).copy(); pos += 4 * rows * cols
        m = i32()
        ids = np.frombuffer(b, "<i4", m, pos).copy(); pos += 4 * m
        out.append((name, period, x, ids))
    assert pos == len(b)
    return out


@pytest.mark.parametrize("ext", [(0, 0), (2, 2), (4, 1)])
def test_feature_and_label_readers_match_the_reference_front_end(tmp_path, ext, feature_io_exe):
    """nnet-asr_b200/host/io.h (FeatureRepository: script-file entries, byte order, STARTFRMEXT/ENDFRMEXT replication; LabelRepository:
    MLF records by logical base name, time -> frame rounding, truncation) against the reference's own KaldiLib readers driven in
    the trainer's call order (oracle/ref_tools/io_dump.cc -> oracle/_ref/RefIoDump), and against the committed dump of that tool."""
    d = str(tmp_path)
    scp, mlf, lmap = _io_dataset(d)
    exe = feature_io_exe
    mine = str(tmp_path / "mine.bin")
    subprocess.check_call([exe, scp, mlf, lmap, str(ext[0]), str(ext[1]), "1", mine, "*/"])     # -L "*/" -X lab, as the scripts run it
    got = _parse_io_dump(mine)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "cpu_io_dump.npz"))
    key = "e%d_%d" % ext
    assert [g[0].replace(d, "<D>") for g in got] == [str(s) for s in gold[key + "_names"]]
    np.testing.assert_array_equal(np.concatenate([g[2].ravel() for g in got]), gold[key + "_feats"])
    np.testing.assert_array_equal(np.concatenate([g[3] for g in got]), gold[key + "_ids"])
    np.testing.assert_array_equal(np.array([g[2].shape[0] for g in got]), gold[key + "_rows"])
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    if os.path.exists(ref_exe):                 # the build container: the reference's readers, live, on the same files
        ref = str(tmp_path / "ref.bin")
        subprocess.check_call([ref_exe, scp, mlf, lmap, str(ext[0]), str(ext[1]), "1", ref, "*/"])
        assert open(ref, "rb").read() == open(mine, "rb").read()


@pytest.mark.parametrize("case", ["missing_record", "unknown_tag", "frame_assigned_twice", "unlabelled_frames", "range_outside_file", "ok"])
def test_feature_and_label_reader_errors_match_the_reference(tmp_path, case, feature_io_exe):
    """Malformed inputs must fail where the reference's readers fail (Labels.cc:55-57,124-145; Features.cc:1193-1195) — and the
    well-formed control must pass in both."""
    d = str(tmp_path)
    r = np.random.default_rng(3)
    x = r.standard_normal((20, 5)).astype(np.float32)
    fea = os.path.join(d, "a.fea")
    F.write_htk(fea, x)
    open(os.path.join(d, "map"), "w").write("s0\ns1\n")
    rec = '"*/a.lab"\n0 1000000 s0\n1000000 2000000 s1\n.\n'
    entry = fea
    if case == "missing_record":
        rec = rec.replace("a.lab", "b.lab")
    elif case == "unknown_tag":
        rec = rec.replace("s1", "s7")
    elif case == "frame_assigned_twice":
        rec = rec.replace("1000000 2000000", "900000 2000000")      # frame 9 belongs to both segments after rounding
    elif case == "unlabelled_frames":
        rec = rec.replace("0 1000000 s0\n", "0 700000 s0\n")               # frames 7..9 carry no target ("Desired vector sum isn't 1.0")
    elif case == "range_outside_file":
        entry = "a.fea=" + fea + "[5,25]"
    open(os.path.join(d, "a.scp"), "w").write(entry + "\n")
    open(os.path.join(d, "a.mlf"), "w").write("#!MLF!#\n" + rec)
    exe = feature_io_exe
    args = [os.path.join(d, "a.scp"), os.path.join(d, "a.mlf"), os.path.join(d, "map"), "1", "1", "1"]
    mine = subprocess.run([exe] + args + [os.path.join(d, "mine.bin"), "*/"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert (mine.returncode == 0) == (case == "ok"), mine.stderr[-500:]
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    if os.path.exists(ref_exe):
        ref = subprocess.run([ref_exe] + args + [os.path.join(d, "ref.bin"), "*/"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert (ref.returncode == 0) == (mine.returncode == 0), (ref.returncode, mine.returncode, ref.stderr[-300:], mine.stderr[-300:])
        if case == "ok":
            assert open(os.path.join(d, "ref.bin"), "rb").read() == open(os.path.join(d, "mine.bin"), "rb").read()


@pytest.mark.parametrize("case", ["same_base_two_dirs", "exact_with_label_dir", "star_depth_two", "question_mark_pattern", "first_definition_wins",
                                  "pattern_before_hashed_name", "bare_name_vs_star_pattern", "no_match_in_other_dir"])
def test_mlf_record_lookup_rules_match_the_reference(tmp_path, case, feature_io_exe):
    """Which MLF record an utterance gets (SOURCETRANSCDIR / SOURCETRANSCEXT applied to the logical feature name, then the reference's
    record index: exact names and '*/tail' names hashed, other wildcards in file order, first definition wins — MlfStream.cc:40-270,
    Labels.cc:52) — differential against the reference's own readers; the expected ids are also stated here."""
    d = str(tmp_path)
    r = np.random.default_rng(5)
    open(os.path.join(d, "map"), "w").write("s0\ns1\ns2\n")
    feas = []
    for k in range(2):
        fea = os.path.join(d, "f%d.fea" % k)
        F.write_htk(fea, r.standard_normal((6, 4)).astype(np.float32))
        feas.append(fea)
    rec = lambda name, tag: '"%s"\n0 600000 %s\n.\n' % (name, tag)
    ldir, expect = None, None
    if case == "same_base_two_dirs":          # ids follow the directory of the logical name, not just its base name
        scp = ["a/utt.fea=" + feas[0], "b/utt.fea=" + feas[1]]
        mlf = rec("b/utt.lab", "s2") + rec("a/utt.lab", "s1")
        expect = [1, 2]
    elif case == "exact_with_label_dir":      # -L labs/x: the directory replaces the logical name's
        scp = ["a/utt.fea=" + feas[0], "b/other.fea=" + feas[1]]
        mlf = rec("labs/x/utt.lab", "s1") + rec("labs/x/other.lab", "s2")
        ldir, expect = "labs/x", [1, 2]
    elif case == "star_depth_two":            # '*' + the tail from the second '/' from the right
        scp = ["top/a/utt.fea=" + feas[0], "top/b/utt.fea=" + feas[1]]
        mlf = rec("*/a/utt.lab", "s0") + rec("*/b/utt.lab", "s2")
        expect = [0, 2]
    elif case == "question_mark_pattern":
        scp = ["utt1_x.fea=" + feas[0], "utt2_x.fea=" + feas[1]]
        mlf = rec("*/utt?_x.lab", "s1")
        expect = [1, 1]
    elif case == "first_definition_wins":
        scp = ["x/utt.fea=" + feas[0], "x/w.fea=" + feas[1]]
        mlf = rec("*/utt.lab", "s2") + rec("*/utt.lab", "s0") + rec("*/w.lab", "s1")
        expect = [2, 1]
    elif case == "pattern_before_hashed_name":    # a wildcard pattern defined before a hashed name that it also matches takes the utterance
        scp = ["x/utt.fea=" + feas[0], "x/w.fea=" + feas[1]]
        mlf = rec("*/u?t.lab", "s1") + rec("*/utt.lab", "s2") + rec("*/w.lab", "s0")
        expect = [1, 0]
    elif case == "bare_name_vs_star_pattern":     # the hashed "*/utt.lab" is looked up as '*' + the tail from a '/': a name without one finds nothing
        scp = ["utt.fea=" + feas[0]]
        mlf = rec("*/utt.lab", "s1")
    else:                                      # an exact name in another directory does not match
        scp = ["a/utt.fea=" + feas[0]]
        mlf = rec("b/utt.lab", "s1")
    open(os.path.join(d, "a.scp"), "w").write("\n".join(scp) + "\n")
    open(os.path.join(d, "a.mlf"), "w").write("#!MLF!#\n" + mlf)
    args = [os.path.join(d, "a.scp"), os.path.join(d, "a.mlf"), os.path.join(d, "map"), "0", "0", "1"]
    tail = [ldir] if ldir is not None else []
    mine = subprocess.run([feature_io_exe] + args + [os.path.join(d, "mine.bin")] + tail, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert (mine.returncode == 0) == (expect is not None), mine.stderr[-500:]
    if expect is not None:
        got = _parse_io_dump(os.path.join(d, "mine.bin"))
        assert [int(g[3][0]) for g in got] == expect and all((g[3] == g[3][0]).all() for g in got)
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    if os.path.exists(ref_exe):
        ref = subprocess.run([ref_exe] + args + [os.path.join(d, "ref.bin")] + tail, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert (ref.returncode == 0) == (mine.returncode == 0), (ref.returncode, mine.returncode, ref.stderr[-300:], mine.stderr[-300:])
        if expect is not None:
            assert open(os.path.join(d, "ref.bin"), "rb").read() == open(os.path.join(d, "mine.bin"), "rb").read()


@pytest.mark.parametrize("kind,ok", [("ANON", True), ("USER", True), ("USER_Z", True), ("MFCC", False), ("USER_E", False), ("ANON_Z", True)])
def test_target_kind_is_applied_or_refused_as_in_the_reference(tmp_path, kind, ok, feature_io_exe):
    """TARGETKIND through the trainers' reading path: conversions the reference can do are done (here _Z), the others fail with its
    message ("Cannot convert ...", Features.cc:1164-1178) instead of being ignored."""
    d = str(tmp_path)
    fea = os.path.join(d, "a.fea")
    F.write_htk(fea, np.random.default_rng(7).standard_normal((6, 4)).astype(np.float32))      # kind USER (9)
    open(os.path.join(d, "map"), "w").write("s0\n")
    open(os.path.join(d, "a.scp"), "w").write(fea + "\n")
    open(os.path.join(d, "a.mlf"), "w").write('#!MLF!#\n"*/a.lab"\n0 600000 s0\n.\n')
    r = subprocess.run([feature_io_exe, os.path.join(d, "a.scp"), os.path.join(d, "a.mlf"), os.path.join(d, "map"), "0", "0", "1", os.path.join(d, "o.bin"), "*/"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, TEST_TARGETKIND=kind))
    assert (r.returncode == 0) == ok, r.stderr[-400:]
    if not ok:
        assert b"Cannot convert USER to " in r.stderr


# HTK parameter-kind words (Features.h:46-70)
_K = dict(MFCC=6, USER=9, PLP=11, E=0o100, N=0o200, D=0o400, A=0o1000, C=0o2000, Z=0o4000, K=0o10000, O=0o20000, T=0o100000)


def _write_htk_raw(path, x, kind, compressed=False, big_endian=True, period=100000):
    """HTK parameter file of any kind; compressed = 16-bit samples with the per-column scale A and bias B behind the header
    (x = (s + B) / A; the two float vectors count as 4 extra 'samples' in the header)."""
    import struct
    x = np.asarray(x, np.float32)
    n, d = x.shape
    e = ">" if big_endian else "<"
    with open(path, "wb") as f:
        if not compressed:
            f.write(struct.pack(e + "iihH", n, period, 4 * d, kind))
            f.write(x.astype(e + "f4").tobytes())
        else:
            hi, lo = x.max(0).astype(np.float64), x.min(0).astype(np.float64)
            span = np.where(hi > lo, hi - lo, 1.0)
            A = (2 * 32767.0 / span).astype(np.float32)
            B = ((hi + lo) * 32767.0 / span).astype(np.float32)
            q = np.clip(np.rint(x.astype(np.float64) * A - B), -32767, 32767).astype(np.int16)
            f.write(struct.pack(e + "iihH", n + 4, period, 2 * d, kind | _K["C"]))
            f.write(A.astype(e + "f4").tobytes())
            f.write(B.astype(e + "f4").tobytes())
            f.write(q.astype(e + "i2").tobytes())


def _ceps_file(path, tag, values, kind_str=None):
    with open(path, "w") as f:
        if kind_str is not None:
            f.write("<CEPSNORM> <%s>\n" % kind_str)
        f.write("<%s> %d\n" % (tag, len(values)))
        f.write(" ".join("%.6e" % v for v in values) + "\n")


_FEA_CASES = {
    # name: (source kind word, static coefs, file derivative order, has _E, has _0, compressed, environment, expected to succeed)
    "user_plain": ("USER", 7, 0, 0, 0, False, {}, True),
    "user_compressed": ("USER", 7, 0, 0, 0, True, {}, True),
    "mfcc_e_d_a_as_is": ("MFCC", 12, 2, 1, 0, False, {"FEA_TARGETKIND": "MFCC_E_D_A"}, True),
    "mfcc_e_d_a_keep_statics": ("MFCC", 12, 2, 1, 0, False, {"FEA_TARGETKIND": "MFCC_E"}, True),
    "mfcc_e_d_a_drop_energy": ("MFCC", 12, 2, 1, 0, False, {"FEA_TARGETKIND": "MFCC_D"}, True),
    "mfcc_e_d_a_third_derivative": ("MFCC", 12, 2, 1, 0, False, {"FEA_TARGETKIND": "MFCC_E_D_A_T", "FEA_THIRDWINDOW": "3"}, True),
    "mfcc_e_d_a_sentence_cmn": ("MFCC", 12, 2, 1, 0, True, {"FEA_TARGETKIND": "MFCC_E_D_A_Z"}, True),
    "mfcc_0_compute_d_a": ("MFCC", 12, 0, 0, 1, False, {"FEA_TARGETKIND": "MFCC_0_D_A", "FEA_DELTAWINDOW": "3", "FEA_ACCWINDOW": "1"}, True),
    "mfcc_0_derivwindows": ("MFCC", 12, 0, 0, 1, True, {"FEA_TARGETKIND": "MFCC_0", "FEA_DERIVWINDOWS": "2_3"}, True),
    "anon_z_d": ("PLP", 9, 0, 1, 1, False, {"FEA_TARGETKIND": "ANON_E_0_D_Z"}, True),
    "mfcc_e_d_suppress_abs_energy": ("MFCC", 12, 1, 1, 0, False, {"FEA_TARGETKIND": "MFCC_E_D_N"}, True),
    "cmn_cvn_cvg_files": ("MFCC", 12, 1, 1, 0, False, {"FEA_TARGETKIND": "MFCC_E_D_A", "norm_files": "1"}, True),
    "other_base_kind": ("MFCC", 12, 0, 0, 0, False, {"FEA_TARGETKIND": "PLP"}, False),
    "energy_the_file_lacks": ("MFCC", 12, 0, 0, 0, False, {"FEA_TARGETKIND": "MFCC_E"}, False),
    "cmn_file_of_another_kind": ("MFCC", 12, 1, 1, 0, False, {"FEA_TARGETKIND": "MFCC_E_D", "norm_files": "wrong_kind"}, False),
}


@pytest.mark.parametrize("ext", [(0, 0), (3, 2)])
@pytest.mark.parametrize("case", sorted(_FEA_CASES))
def test_feature_kind_conversions_match_the_reference_reader(tmp_path, case, ext, feature_io_exe):
    """FeatureRepository::ReadFullMatrix with everything the reference's reader does between the file and the trainer (Features.cc:
    1025-1440): compressed files, TARGETKIND conversions (energy columns and derivative blocks dropped, _Z sentence mean normalisation,
    missing derivatives computed with DELTAWINDOW / ACCWINDOW / THIRDWINDOW / DERIVWINDOWS), CMEANDIR+CMEANMASK / VARSCALEDIR+
    VARSCALEMASK / VARSCALEFN files, frame ranges and context rows — byte for byte against the reference's own reader where it is
    built (oracle/_ref/RefIoDump), and against properties stated here everywhere."""
    base, nc, sd, has_e, has_0, comp, env, ok = _FEA_CASES[case]
    d = str(tmp_path)
    import zlib
    r = np.random.default_rng(zlib.crc32(case.encode()) % 1000)      # a stable seed: the committed hashes below depend on the data
    width = (nc + has_e + has_0) * (sd + 1)
    kind = _K[base] | (_K["E"] if has_e else 0) | (_K["O"] if has_0 else 0) | (_K["D"] if sd >= 1 else 0) | (_K["A"] if sd >= 2 else 0)
    names = ["spk1_a", "spk1_b", "spk2_a"]
    data = {}
    for k, nm in enumerate(names):
        data[nm] = (r.standard_normal((25 + 7 * k, width)) * 3 + r.standard_normal(width)).astype(np.float32)
        _write_htk_raw(os.path.join(d, nm + ".fea"), data[nm], kind, compressed=comp)
    # logical names without the temporary directory: the masks below must find the speaker in them, not in a path component
    scp = ["lg/spk1_a.fea=" + os.path.join(d, "spk1_a.fea"), "lg/spk1_b.fea=" + os.path.join(d, "spk1_b.fea") + "[4,20]",
           "spk2_a.fea=" + os.path.join(d, "spk2_a.fea") + "[0,9]"]
    open(os.path.join(d, "a.scp"), "w").write("\n".join(scp) + "\n")
    env = dict(env)
    nf = env.pop("norm_files", None)
    if nf:
        os.makedirs(os.path.join(d, "cmn")), os.makedirs(os.path.join(d, "cvn"))
        tk = env["FEA_TARGETKIND"]
        n_static = nc + has_e + has_0
        n_all = n_static * (tk.count("_D") + tk.count("_A") + tk.count("_T") + 1)
        static_kind = "MFCC_E" if nf != "wrong_kind" else "MFCC_0"
        for spk in ("spk1", "spk2"):
            _ceps_file(os.path.join(d, "cmn", spk), "MEAN", r.standard_normal(n_static), static_kind)
            _ceps_file(os.path.join(d, "cvn", spk), "VARIANCE", r.random(n_all) + 0.5, tk)
        _ceps_file(os.path.join(d, "cvg"), "VARSCALE", r.random(n_all) + 0.5)
        env.update(FEA_CMNDIR=os.path.join(d, "cmn"), FEA_CMNMASK="%%%%_*", FEA_CVNDIR=os.path.join(d, "cvn"), FEA_CVNMASK="%%%%_*.fea",
                   FEA_CVGFILE=os.path.join(d, "cvg"))
    args = ["--fea", os.path.join(d, "a.scp"), str(ext[0]), str(ext[1]), "1"]
    mine = subprocess.run([feature_io_exe] + args + [os.path.join(d, "mine.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **env))
    assert (mine.returncode == 0) == ok, mine.stderr[-500:]
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    if os.path.exists(ref_exe):
        ref = subprocess.run([ref_exe] + args + [os.path.join(d, "ref.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **env))
        assert (ref.returncode == 0) == ok, ref.stderr[-500:]
        if ok:
            assert open(os.path.join(d, "ref.bin"), "rb").read() == open(os.path.join(d, "mine.bin"), "rb").read()
    if not ok:
        return
    b = open(os.path.join(d, "mine.bin"), "rb").read()
    # the reference reader's dump of the same case, as a committed hash (tests/golden/cpu_fea_conv.json; regenerate with
    # UPDATE_FEA_GOLDEN=1 where oracle/_ref/RefIoDump exists): pins the reader where the reference is not built
    import hashlib, json
    gpath, key = os.path.join(ROOT, "tests", "golden", "cpu_fea_conv.json"), "%s/%d_%d" % (case, ext[0], ext[1])
    gold = json.load(open(gpath)) if os.path.exists(gpath) else {}
    if os.environ.get("UPDATE_FEA_GOLDEN") and os.path.exists(ref_exe):
        gold[key] = hashlib.sha256(open(os.path.join(d, "ref.bin"), "rb").read()).hexdigest()
        json.dump(gold, open(gpath, "w"), indent=1, sort_keys=True)
    assert gold.get(key) == hashlib.sha256(b).hexdigest(), "dump differs from the reference reader's (tests/golden/cpu_fea_conv.json)"
    pos, mats = 4, []
    assert int(np.frombuffer(b, "<i4", 1, 0)[0]) == 3
    for _ in range(3):
        rows, cols, knd, per = (int(v) for v in np.frombuffer(b, "<i4", 4, pos)); pos += 16
        mats.append((knd, np.frombuffer(b, "<f4", rows * cols, pos).reshape(rows, cols))); pos += 4 * rows * cols
    assert pos == len(b)
    assert [m.shape[0] for _, m in mats] == [25 + ext[0] + ext[1], 17 + ext[0] + ext[1], 10 + ext[0] + ext[1]]
    tol = 2e-3 if comp else 0.0
    if case in ("user_plain", "user_compressed", "mfcc_e_d_a_as_is"):
        np.testing.assert_allclose(mats[0][1][ext[0]:ext[0] + 25], data["spk1_a"], rtol=0, atol=tol * 12)
        lo = max(0, 4 - ext[0])                  # a range takes its context from the file's own frames where it has them
        np.testing.assert_allclose(mats[1][1][ext[0] - (4 - lo):ext[0] + 17], data["spk1_b"][lo:21], rtol=0, atol=tol * 12)
    if case == "mfcc_e_d_a_keep_statics":
        np.testing.assert_array_equal(mats[0][1][ext[0]:ext[0] + 25], data["spk1_a"][:, :13])
    if case == "mfcc_e_d_a_drop_energy":
        np.testing.assert_array_equal(mats[0][1][ext[0]:ext[0] + 25], np.hstack([data["spk1_a"][:, 0:12], data["spk1_a"][:, 13:25]]))
    if case == "mfcc_e_d_a_sentence_cmn":
        assert np.abs(mats[0][1][:, :13].mean(0)).max() < 1e-4 and mats[0][0] & _K["Z"]
    if case == "mfcc_0_compute_d_a":
        m = mats[0][1]
        assert m.shape[1] == 39
        i = m.shape[0] // 2                      # an interior frame: delta = sum_k k (x[i+k] - x[i-k]) / (2 sum k^2), window 3
        want = sum(k * (m[i + k, :13] - m[i - k, :13]) for k in (1, 2, 3)) / 28.0
        np.testing.assert_allclose(m[i, 13:26], want, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(m[i, 26:39], (m[i + 1, 13:26] - m[i - 1, 13:26]) / 2.0, rtol=1e-5, atol=1e-6)


def test_feature_kind_conversions_the_reference_leaves_undefined_are_refused(tmp_path, feature_io_exe):
    """_N (absolute energy suppressed) together with sentence mean normalisation, or with first derivatives computed from the statics,
    makes the reference index one element before a feature row (Features.cc:1289,1317): the reader refuses instead."""
    d = str(tmp_path)
    x = np.random.default_rng(2).standard_normal((12, 13)).astype(np.float32)
    _write_htk_raw(os.path.join(d, "a.fea"), x, _K["MFCC"] | _K["E"])
    open(os.path.join(d, "a.scp"), "w").write(os.path.join(d, "a.fea") + "\n")
    for tk in ("MFCC_E_D_N", "MFCC_E_D_N_Z"):
        r = subprocess.run([feature_io_exe, "--fea", os.path.join(d, "a.scp"), "0", "0", "1", os.path.join(d, "o.bin")], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, env=dict(os.environ, FEA_TARGETKIND=tk))
        assert r.returncode != 0 and b"undefined in the reference" in r.stderr, r.stderr[-300:]


def _config_block(txt):
    lines = txt.splitlines()
    at = [k for k, l in enumerate(lines) if l.startswith("Configuration Parameters[")]
    if not at:
        return None                 # the command line was refused before the dump
    i = at[0]
    n = int(lines[i].split("[")[1].split("]")[0])
    return lines[i:i + 1 + n]


def test_config_file_and_option_dump_like_the_reference(tmp_path):
    """-C config files (comments, `KEY = value`, module prefix, case-insensitive keys), short options mapped to parameter names and
    the -D dump (UserInterface.cc:645-654): the drop-in's dump equals the one the reference CPU trainer prints for the same command
    line (live, where oracle/_ref/TNet exists) and the expected block below."""
    cfg = tmp_path / "cfg"
    cfg.write_text("# comment\nBUNCHSIZE = 32\nTNET:CACHESIZE=256\nrandomize = F\n\nTnet:Seed = 7   # trailing comment\n")
    args = ["-C", str(cfg), "-D", "-H", "nonexistent.nnet", "-n", "0.5", "--WEIGHTCOST=1e-4", "-S", "nonexistent.scp"]
    mine = subprocess.run([os.path.join(BIN, "TNetCu")] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert mine.returncode != 0 and "nonexistent.nnet" in mine.stdout          # reading the network is the first thing that fails
    block = _config_block(mine.stdout)
    assert block[0] == "Configuration Parameters[9]"
    assert block[1] == "  %-35s = %-30s # -C" % ("BUNCHSIZE", "32")
    assert "  %-35s = %-30s # -n" % ("TNET:LEARNINGRATE", "0.5") in block
    assert "  %-35s = %-30s # --" % ("TNET:WEIGHTCOST", "1e-4") in block
    assert "  %-35s = %-30s # -C" % ("TNET:SEED", "7") in block
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "TNet")
    if os.path.exists(ref_exe):
        # "# " marks parameters not read yet at the time of the dump; the two mains read theirs in a different order, so the marker
        # column is compared after normalisation and everything else verbatim
        norm = lambda ls: [l[2:] if l[:2] in ("  ", "# ") else l for l in ls]
        more = [["-D", "--learningrate=0.5", "-H", "x"], ["-D", "-n", "1", "-n", "2", "-H", "x"], ["-D", "-T", "3", "-H", "x"],
                ["-D", "--BUNCHSIZE", "64", "-H", "x"], ["-D", "-c", "-H", "x"], ["-D", "--CROSSVALIDATE=t", "-H", "x"],
                ["-D", "-H", "x", "-S", "a.scp", "extra1", "extra2"], ["-D", "--TNET:SEED=5", "-H", "x"], ["-D", "-A", "-V", "-H", "x"],
                # the feature-side parameters (UserInterface::GetFeatureParams)
                ["-D", "--TARGETKIND=MFCC_E_D_A", "--DELTAWINDOW=3", "--ACCWINDOW=1", "-H", "x"],
                ["-D", "--TARGETKIND=MFCC_0", "--DERIVWINDOWS=2", "-H", "x"],   # (with "2_3" the reference's strtok leaves a NUL in the value it dumps)
                ["-D", "--CMEANDIR=/a", "--CMEANMASK=%%%_*", "--VARSCALEDIR=/b", "--VARSCALEMASK=%%%_*", "--VARSCALEFN=/c", "-H", "x"],
                ["-D", "--STARTFRMEXT=3", "--ENDFRMEXT=4", "--NATURALREADORDER=T", "-H", "x"]]
        for a in [args] + more:
            ref = subprocess.run([ref_exe] + a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            got = subprocess.run([os.path.join(BIN, "TNetCu")] + a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            rb, gb = _config_block(ref.stdout), _config_block(got.stdout)
            assert (rb is None) == (gb is None), a
            assert rb is None or norm(rb) == norm(gb), a
            assert (ref.returncode != 0) and (got.returncode != 0)
        # which parameters count as read decides what "Unexpected command line parameter" names: the single-window parameters are
        # read only when TARGETKIND is not plain ANON, and not at all next to DERIVWINDOWS (UserInterface.cc:421-460)
        unexpected = lambda txt: sorted(set(re.findall(r"Unexpected command line parameter (\S+)", txt)))
        for a, want in [(["--DELTAWINDOW=3", "-H", "x"], ["TNET:DELTAWINDOW"]),
                        (["--TARGETKIND=MFCC_D", "--DELTAWINDOW=3", "-H", "x"], []),
                        (["--TARGETKIND=MFCC_D", "--DERIVWINDOWS=2", "--DELTAWINDOW=3", "-H", "x"], ["TNET:DELTAWINDOW"]),
                        (["--TARGETKIND=USER", "--THIRDWINDOW=3", "-H", "x"], [])]:
            ref = subprocess.run([ref_exe] + a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            got = subprocess.run([os.path.join(BIN, "TNetCu")] + a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            assert unexpected(ref.stdout)[:1] == want[:1] and unexpected(got.stdout)[:1] == want[:1], (a, ref.stdout[-300:], got.stdout[-300:])


def test_make_htk_file_name_like_the_reference(tmp_path, feature_io_exe):
    """MakeHtkFileName (Common.cc:118-175: output names from -M/-o, label names from -L/-X, TFeaCat's -l/-y): directory, extension
    and the `/./` marker that keeps a sub-path — expected strings from the reference's function, compared live where it is built."""
    exe = feature_io_exe
    cases = [(("a/b/c.fea", "out", "lab"), "out/c.lab"), (("a/b/c.fea", "-", "lab"), "a/b/c.lab"), (("a/b/c.fea", "out", "-"), "out/c.fea"),
             (("x/./sub/c.d.fea", "out", "ext"), "out/sub/c.d.ext"), (("c", "-", "-"), "c"), (("-", "out", "lab"), "-"),
             (("a/b/c.fea", "", "lab"), "c.lab"), (("a.b/c", "o", "e"), "o/c.e"), (("/abs/dir/name.x.y", "*", "lab"), "*/name.x.lab")]
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    for args, want in cases:
        got = subprocess.check_output([exe, "--htkname"] + list(args), text=True).rstrip("\n")
        assert got == want, (args, got, want)
        if os.path.exists(ref_exe):
            assert subprocess.check_output([ref_exe, "--htkname"] + list(args), text=True).rstrip("\n") == got, args


def test_htk_writer_and_header_checks_like_the_reference(tmp_path, feature_io_exe):
    """FeatureRepository::WriteFeatureMatrix as TFeaCat uses it (TFeaCat.cc:262): a script-file entry read with context rows and
    written back must give the reference's bytes (header: frame count, source sample period, size, USER kind; big-endian data);
    headers the reference rejects (Features.cc:522-528: sample period outside [0, 100000] — also what a wrong byte order looks like)
    must be rejected."""
    exe = feature_io_exe
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    r = np.random.default_rng(1)
    x = r.standard_normal((12, 5)).astype(np.float32)
    fea, bad = str(tmp_path / "a.fea"), str(tmp_path / "bad.fea")
    F.write_htk(fea, x, samp_period=62500)
    F.write_htk(bad, x, samp_period=123400)
    import struct
    for entry, lo, hi in ((fea, 0, 11), ("z.fea=" + fea + "[2,7]", 2, 7)):
        for se, ee in ((0, 0), (3, 2)):
            out = str(tmp_path / "mine.htk")
            subprocess.check_call([exe, "--rewrite", entry, str(se), str(ee), "1", out])
            rows = np.clip(np.arange(lo - se, hi + ee + 1), 0, 11)
            want = struct.pack(">iihh", len(rows), 62500, 20, 9) + x[rows].astype(">f4").tobytes()
            assert open(out, "rb").read() == want
            if os.path.exists(ref_exe):
                ref = str(tmp_path / "ref.htk")
                subprocess.check_call([ref_exe, "--rewrite", entry, str(se), str(ee), "1", ref])
                assert open(ref, "rb").read() == want
    for tool in [exe] + ([ref_exe] if os.path.exists(ref_exe) else []):
        for args in ((bad, "1"), (fea, "0")):       # period 123400 ; little-endian read of a big-endian file
            rc = subprocess.run([tool, "--rewrite", args[0], "0", "0", args[1], str(tmp_path / "x.htk")], stderr=subprocess.PIPE).returncode
            assert rc != 0, (tool, args)


def test_matrix_vector_text_parsing_like_the_reference(tmp_path, feature_io_exe):
    """The text operators the network files are read with (tnet_base.h) against the reference's (Matrix.tcc:575-600,
    Vector.tcc:527-547, i.e. `istream >> float`): same values for every syntax the reference accepts, refusal where it refuses
    ("nan"/"inf" of a diverged network, hexadecimal floats, values beyond the float range, truncated or malformed matrices)."""
    exe = feature_io_exe
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    f32 = lambda *v: np.array(v, np.float32).tobytes()
    i32 = lambda *v: np.array(v, np.int32).tobytes()
    cases = {   # text -> expected dump (None = must be refused)
        "m 2 3\n1 2 3 \n4 5 6 \nv 2  7 8 \n": i32(2, 3) + f32(1, 2, 3, 4, 5, 6) + i32(2) + f32(7, 8),
        "\n\n  m 2 2\n1e-3   -2.5E+2\n\t.5 +7.\n\n v 3 1 2\n3\n": i32(2, 2) + f32(1e-3, -250, .5, 7) + i32(3) + f32(1, 2, 3),
        "m 1 4 1 2 3 4 v 1 9": i32(1, 4) + f32(1, 2, 3, 4) + i32(1) + f32(9),
        "m 1 3 -0 1e-42 3.4028235e38 v 1 1e-50": i32(1, 3) + f32(-0.0, 1e-42, 3.4028235e38) + i32(1) + f32(0),
        "m 0 0\nv 0 \n": i32(0, 0) + i32(0),
        "m 1 2 0.1234567890123456789 16777217 v 1 0.30000001192092896": i32(1, 2) + f32(0.1234567890123456789, 16777217) + i32(1) + f32(0.3),
        "m 1 2 1e39 -1e39 v 1 0": None, "m 2 2 1 2 3 v 2 1 2": None, "m 1 2 nan 1 v 1 0": None, "m 1 2 inf 1 v 1 0": None,
        "m 1 2 0x10 1 v 1 0": None, "m -1 2 v 1 0": None, "m 1 2 1,2 v 1 0": None,
    }
    for txt, want in cases.items():
        src, out = str(tmp_path / "in.txt"), str(tmp_path / "out.bin")
        open(src, "w").write(txt)
        for tool in [exe] + ([ref_exe] if os.path.exists(ref_exe) else []):
            if os.path.exists(out):
                os.unlink(out)
            rc = subprocess.run([tool, "--readmv", src, out], stderr=subprocess.PIPE).returncode
            if want is None:
                assert rc != 0, (tool, txt)
            else:
                assert rc == 0 and open(out, "rb").read() == want, (tool, txt)


@pytest.mark.parametrize("case", ["plain", "ragged_white_space", "bad_token_inside", "bad_token_behind", "too_short", "value_out_of_range"])
def test_large_matrix_text_is_read_in_parallel_like_the_reference(tmp_path, case, feature_io_exe):
    """Matrices of 2^18 numbers and more are tokenised and converted on several threads (tnet_base.h ReadNumbersParallel: a 28 M-weight
    network parses in 1.6 s instead of 4.5 s; the reference needs 12 s): same values, same refusals, and the stream is left exactly
    behind the last number (the vector that follows is read correctly) — against the reference's own operators."""
    r = np.random.default_rng(9)
    rows, cols = 600, 512
    M = (r.standard_normal((rows, cols)) * 10.0 ** r.integers(-6, 6, (rows, cols))).astype(np.float32)
    toks = ["%.7g" % v for v in M.ravel()]
    want_ok = True
    if case == "bad_token_inside":
        toks[200000] = "nan"; want_ok = False
    if case == "value_out_of_range":
        toks[len(toks) - 5] = "1e39"; want_ok = False
    if case == "too_short":
        toks = toks[:-3]; want_ok = False
    if case == "ragged_white_space":
        sep = np.array([" ", "  ", "\n", "\t ", " \r\n"])[r.integers(0, 5, len(toks))]
        body = "".join(t + s_ for t, s_ in zip(toks, sep))
    else:
        body = "\n".join(" ".join(toks[i * cols:(i + 1) * cols]) + " " for i in range((len(toks) + cols - 1) // cols)) + "\n"
    tail = "v 3  1.5 -2 3e3 \n"
    if case == "bad_token_behind":
        tail = "v 3  1.5 -2 3e3 \n<sigmoid> 512 512\nnan\n"      # tokens that are not numbers right behind what is read: none of the reader's business
    if case == "too_short":
        tail = ""
    src = str(tmp_path / "in.txt")
    open(src, "w").write("m %d %d\n" % (rows, cols) + body + tail)
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    outs = []
    for tool in [feature_io_exe] + ([ref_exe] if os.path.exists(ref_exe) else []):
        out = str(tmp_path / (os.path.basename(tool) + ".bin"))
        rc = subprocess.run([tool, "--readmv", src, out], stderr=subprocess.PIPE).returncode
        assert (rc == 0) == want_ok, (tool, case)
        if want_ok:
            outs.append(open(out, "rb").read())
    if want_ok:
        want = np.array([rows, cols], np.int32).tobytes() + np.array([float(t) for t in toks], np.float32).tobytes() + \
            np.array([3], np.int32).tobytes() + np.array([1.5, -2, 3e3], np.float32).tobytes()
        assert all(o == want for o in outs)


@pytest.mark.parametrize("case", ["two_lists", "several_entries_per_line", "blanks_around_names", "trailing_comma", "missing_list"])
def test_script_file_lists_like_the_reference(tmp_path, case, feature_io_exe):
    """-S takes one script file or several separated by commas, and an entry is a white-space separated token (several per line are
    several entries) — FeatureRepository::AddFileList, Features.cc:390-429 — against the reference's own reader."""
    d = str(tmp_path)
    r = np.random.default_rng(6)
    feas = []
    for k in range(4):
        fea = os.path.join(d, "f%d.fea" % k)
        F.write_htk(fea, r.standard_normal((5 + k, 3)).astype(np.float32))
        feas.append(fea)
    a, b = os.path.join(d, "a.scp"), os.path.join(d, "b.scp")
    open(a, "w").write(feas[0] + "\n" + feas[1] + "\n")
    open(b, "w").write(feas[2] + "   " + feas[3] + "\n" if case == "several_entries_per_line" else feas[2] + "\n" + feas[3] + "\n")
    arg, ok, n = {"two_lists": (a + "," + b, True, 4), "several_entries_per_line": (b, True, 2), "blanks_around_names": (" " + a + " , " + b + " ", True, 4),
                  "trailing_comma": (a + ",", False, 0), "missing_list": (a + "," + os.path.join(d, "none.scp"), False, 0)}[case]
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    outs = []
    for tool in [feature_io_exe] + ([ref_exe] if os.path.exists(ref_exe) else []):
        out = os.path.join(d, os.path.basename(tool) + ".bin")
        rc = subprocess.run([tool, "--fea", arg, "0", "0", "1", out], stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode
        assert (rc == 0) == ok, (tool, case)
        if ok:
            outs.append(open(out, "rb").read())
    if ok:
        assert int(np.frombuffer(outs[0], "<i4", 1, 0)[0]) == n and all(o == outs[0] for o in outs)


def test_network_file_with_large_layers_is_walked_like_the_python_parser(tmp_path, feature_io_exe):
    """A three-layer network file whose matrices take the parallel text reader (two of them) and the one-by-one reader (one), walked
    tag by tag as the component factory does: every matrix and vector must come out as the Python parser reads them, i.e. the stream
    stands exactly behind each matrix."""
    r = np.random.default_rng(4)
    dims = [700, 640, 600, 100]         # 640x700 and 600x640 weights: >= 2^18 numbers (parallel reader); 100x600: below
    layers = F.gen_mlp_init(dims, r)
    net = str(tmp_path / "net.txt")
    F.write_mlp(net, layers)
    want = [l for l in F.read_mlp(net) if l[0] == "affine"]
    out = str(tmp_path / "layers.bin")
    subprocess.check_call([feature_io_exe, "--readlayers", net, out])
    b = open(out, "rb").read()
    pos = 0
    for l in want:
        rows, cols = (int(v) for v in np.frombuffer(b, "<i4", 2, pos)); pos += 8
        Wt = np.frombuffer(b, "<f4", rows * cols, pos).reshape(rows, cols); pos += 4 * rows * cols
        d = int(np.frombuffer(b, "<i4", 1, pos)[0]); pos += 4
        bias = np.frombuffer(b, "<f4", d, pos); pos += 4 * d
        np.testing.assert_array_equal(Wt, l[1])
        np.testing.assert_array_equal(bias, l[2])
    assert pos == len(b) and len(want) == 3


_MLF_REC = '"*/a.lab"\n0 1000000 s0\n1000000 2000000 s1\n.\n'
_FRONT_END_VARIANTS = {   # name: (overrides, expected to be accepted)
    "no_mlf_header": (dict(mlf=_MLF_REC), True),
    "crlf": (dict(mlf=("#!MLF!#\n" + _MLF_REC).replace("\n", "\r\n")), True),
    "map_extra_cols": (dict(map="s0 extra\ns1\n"), True),
    "trailing_spaces": (dict(mlf='#!MLF!#\n"*/a.lab"  \n0 1000000 s0  \n1000000 2000000 s1\t\n.\n'), True),
    "name_defined_twice": (dict(mlf="#!MLF!#\n" + _MLF_REC + '"*/a.lab"\n0 1000000 s1\n1000000 2000000 s0\n.\n'), True),   # first wins
    "scp_blank_and_spaces": (dict(scp="\n  {fea}  \n\n"), True),
    "scp_range_single": (dict(scp="a.fea={fea}[4,4]\n", mlf='#!MLF!#\n"*/a.lab"\n0 100000 s1\n.\n'), True),
    "negative_time": (dict(mlf='#!MLF!#\n"*/a.lab"\n-5 1000000 s0\n1000000 2000000 s1\n.\n'), True),
    "comment_in_record": (dict(mlf='#!MLF!#\n"*/a.lab"\n# c\n0 1000000 s0\n1000000 2000000 s1\n.\n'), True),
    "missing_final_dot": (dict(mlf='#!MLF!#\n"*/a.lab"\n0 1000000 s0\n1000000 2000000 s1\n'), True),
    "label_without_times": (dict(mlf='#!MLF!#\n"*/a.lab"\ns0\n.\n'), False),
    "label_with_start_only": (dict(mlf='#!MLF!#\n"*/a.lab"\n0 s0\n.\n'), False),
    "float_times": (dict(mlf='#!MLF!#\n"*/a.lab"\n0.0 1000000.0 s0\n1000000 2000000 s1\n.\n'), False),
    "scp_missing_file": (dict(scp="{dir}/nofile.fea\n"), False),
    "scp_range_reversed": (dict(scp="a.fea={fea}[7,3]\n"), False),
    "duplicate_tag_in_map": (dict(map="s0\ns1\ns0\n"), False),
    "empty_map": (dict(map=""), False),
}


@pytest.mark.parametrize("name", sorted(_FRONT_END_VARIANTS))
def test_front_end_input_variants_like_the_reference(tmp_path, name, feature_io_exe):
    """Spellings of script files, MLFs and label maps that occur in practice: accepted and refused alike by the drop-in's readers and
    by the reference's (live where oracle/_ref/RefIoDump exists), and read to the same bytes when accepted."""
    over, accepted = _FRONT_END_VARIANTS[name]
    d = str(tmp_path)
    x = np.random.default_rng(3).standard_normal((20, 5)).astype(np.float32)
    fea = os.path.join(d, "a.fea")
    F.write_htk(fea, x)
    open(os.path.join(d, "a.scp"), "w").write(over.get("scp", "{fea}\n").format(fea=fea, dir=d))
    open(os.path.join(d, "a.mlf"), "w", newline="").write(over.get("mlf", "#!MLF!#\n" + _MLF_REC))
    open(os.path.join(d, "map"), "w").write(over.get("map", "s0\ns1\n"))
    exe = feature_io_exe
    args = [os.path.join(d, "a.scp"), os.path.join(d, "a.mlf"), os.path.join(d, "map"), "1", "1", "1"]
    mine = subprocess.run([exe] + args + [os.path.join(d, "mine.bin"), "*/"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert (mine.returncode == 0) == accepted, mine.stderr[-400:]
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "RefIoDump")
    if os.path.exists(ref_exe):
        ref = subprocess.run([ref_exe] + args + [os.path.join(d, "ref.bin"), "*/"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert (ref.returncode == 0) == accepted, ref.stderr[-400:]
        if accepted:
            assert open(os.path.join(d, "ref.bin"), "rb").read() == open(os.path.join(d, "mine.bin"), "rb").read()


@pytest.mark.parametrize("mine,ref,lines", [
    ("TFeaCatCu", "TFeaCat", [["-D", "-H", "x", "-l", "outdir", "-y", "post", "-S", "a.scp"], ["-D", "-H", "x", "--LOGPOSTERIOR=TRUE"],
                              ["-D", "-H", "x", "-T", "1", "--STARTFRMEXT=3", "--ENDFRMEXT=3"]]),
    ("TNormCu", "TNorm", [["-D", "-H", "x", "--TARGETMMF=out.t"], ["-D", "-H", "x", "-T", "1", "--STARTFRMEXT=3", "--ENDFRMEXT=3"],
                          ["-D", "-H", "x", "-S", "a.scp", "--NATURALREADORDER=T"]]),
])
def test_tool_option_maps_like_the_reference(mine, ref, lines):
    """The forward-only and normalisation tools name their parameters like the reference's (module prefix, short-option mapping): the
    -D dump of the drop-in equals the reference CPU tool's for the same command line (the GPU tools share SNAME and option map with
    their CPU twins: TFeaCat.cc:40 / TFeaCatCu.cc:43, TNorm.cc:46 / TNormCu.cc:47).  Live comparison where oracle/_ref exists."""
    ref_exe = os.path.join(ROOT, "oracle", "_ref", ref)
    norm = lambda ls: [l[2:] if l[:2] in ("  ", "# ") else l for l in ls]
    for a in lines:
        got = subprocess.run([os.path.join(BIN, mine)] + a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        gb = _config_block(got.stdout)
        assert gb is not None and got.returncode != 0, a
        prefix = {"TFeaCatCu": "TFEACAT:", "TNormCu": "TNORM:"}[mine]
        assert all(l[2:].startswith(prefix) for l in gb[1:]), gb
        if os.path.exists(ref_exe):
            r = subprocess.run([ref_exe] + a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            assert norm(_config_block(r.stdout)) == norm(gb), a

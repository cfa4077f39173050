#!/usr/bin/env python
"""Golden dump of the reference's own front end (oracle/_ref/RefIoDump = oracle/ref_tools/io_dump.cc over the unmodified KaldiLib
FeatureRepository / LabelRepository) on the data set of tests/test_host_cpu.py::_io_dataset -> tests/golden/cpu_io_dump.npz.
Run where /root/reference was available to oracle/build_ref.sh (the build container)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
import test_host_cpu as T  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as d:
    scp, mlf, lmap = T._io_dataset(d)
    for ext in ((0, 0), (2, 2), (4, 1)):
        ref = os.path.join(d, "ref.bin")
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "RefIoDump"), scp, mlf, lmap, str(ext[0]), str(ext[1]), "1", ref, "*/"])
        got = T._parse_io_dump(ref)
        key = "e%d_%d" % ext
        out[key + "_names"] = np.array([g[0].replace(d, "<D>") for g in got])
        out[key + "_feats"] = np.concatenate([g[2].ravel() for g in got])
        out[key + "_ids"] = np.concatenate([g[3] for g in got])
        out[key + "_rows"] = np.array([g[2].shape[0] for g in got])
        print(key, [(g[0], g[2].shape, int((g[3] < 0).sum())) for g in got])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cpu_io_dump.npz"), **out)

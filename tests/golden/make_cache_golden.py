#!/usr/bin/env python
"""Golden dump of the reference's CPU frame cache (oracle/_ref/RefCacheDump = oracle/ref_tools/cache_dump.cc over the unmodified
TNetLib/Cache.cc) on the cases of tests/test_oracle_golden.py::CACHE_CASES -> tests/golden/cpu_cache_dump.npz."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
import test_oracle_golden as T  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as d:
    for name in sorted(T.CACHE_CASES):
        nb, data, disc = T.run_reference_cache(os.path.join(ROOT, "oracle", "_ref", "RefCacheDump"), name, d)
        out[name + "_nb"], out[name + "_data"], out[name + "_discarded"] = np.int64(nb), data, np.int64(disc)
        print(name, "bunches", nb, "discarded", disc)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cpu_cache_dump.npz"), **out)

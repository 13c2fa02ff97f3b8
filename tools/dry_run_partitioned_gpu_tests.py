#!/usr/bin/env python
"""CPU dry run of tests/test_partitioned_norm_gpu.py: the same test functions with the native entry points replaced by
their executable spec (tests/kernel_spec.py), a gloo group instead of NCCL and DEV = "cpu" -- checks the TEST LOGIC (shapes,
names, collective choreography, tolerances) before GPU minutes are spent on it; it says nothing about the CUDA kernels.
usage: python tools/dry_run_partitioned_gpu_tests.py [-k expr]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import pytest  # noqa: E402
import torch  # noqa: E402

import tests.test_partitioned_norm_gpu as t  # noqa: E402
from oracle import c_oracle, ngram_oracle  # noqa: E402
from protgram_directgcn_b200 import _native as nat  # noqa: E402
from tests import kernel_spec  # noqa: E402

kernel_spec.install_plain(nat)
nat.SpmmPlan.ref = lambda self, width: None       # the spec ignores the long-row plan (it is a ctypes struct of device pointers)
t.DEV, t.BACKEND = "cpu", "gloo"
t._synth_corpus = lambda nseq, L, seed=42: torch.from_numpy(np.array(c_oracle.pack_corpus(ngram_oracle.synth_sequences(0, nseq, L))))
sys.exit(pytest.main([os.path.join(ROOT, "tests", "test_partitioned_norm_gpu.py"), "-x", "-q", "-p", "no:cacheprovider", "-m", ""] + sys.argv[1:]))

#!/usr/bin/env python
"""Generates tests/golden/io_ref.npz: the outputs of the reference's OWN Plus / GroundTruth / ImageReader / DataReader
sources (compiled unmodified into oracle/_ref/libref_io.so by `make -C oracle ref`; needs /root/reference) on the
deterministic inputs of tests/test_dataset_io.py.  Run from the repo root: python tests/golden/make_io_golden.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_dataset_io as t  # noqa: E402

ref = t._load(t.REF_SO)
out = {"plus_" + k: v for k, v in t.run_plus(ref, "ref_").items()}
with tempfile.TemporaryDirectory() as d:
    out.update({"ds_" + k: v for k, v in t.run_dataset(ref, "ref_", d).items()})
np.savez_compressed(t.GOLDEN, **out)
print("wrote", t.GOLDEN, {k: np.asarray(v).shape for k, v in out.items()})

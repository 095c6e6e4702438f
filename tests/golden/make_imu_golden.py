#!/usr/bin/env python
"""Generates tests/golden/imu_ref.npz: the reference's OWN src/Imu.cpp (compiled unmodified into
oracle/_ref/libref_imu.so by `make -C oracle ref`; needs /root/reference) run on the samples of tests/test_imu.py with
the orientation answers recorded from the product's filter.  Run from the repo root."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_imu as t  # noqa: E402

host, ref = C.CDLL(t.HOST_SO), C.CDLL(t.REF_SO)
w, a = t.samples()
_, q = t.run_host(host, w, a, want_q=True)
out = t.run_ref(ref, w, a, q)
np.savez_compressed(t.GOLDEN, q=q, out=out)
print("wrote", t.GOLDEN, q.shape, out.shape)

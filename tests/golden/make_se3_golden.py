#!/usr/bin/env python
"""Generates tests/golden/se3_ref.npz from the reference's OWN vendored Sophus (thirdparty/sophus/se3.hpp, so3.hpp compiled
unmodified into oracle/_ref/libref_sophus.so by `make -C oracle ref`; needs /root/reference): SE3::exp (incl. the Taylor branch
below theta = 1e-5 and V = R there), the group product (incl. results whose squared norm is not 1, i.e. the first-order
renormalisation branch of SO3::operator*=), SE3::matrix and SE3(R, t).  sin / cos of a float are the float rounding of the
double function (oracle/cr_trig.c explains why).  Run from the repo root."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_se3_sophus as t  # noqa: E402

L = C.CDLL(t.REF_SO)
g = t.cases(t.N_GOLDEN, seed=20261018)
out = t.run(L, "sph", g)
np.savez_compressed(t.GOLDEN, **g, **{"out_" + k: v for k, v in out.items()})
print("wrote", t.GOLDEN, {k: v.shape for k, v in out.items()},
      "taylor cases", int(t.taylor_mask(g["delta_a"]).sum()), "renormalised products", int(out["renorm"].sum()))

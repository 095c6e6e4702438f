"""SE3 arithmetic pinned against the reference's OWN vendored Sophus.

oracle/Makefile target `ref` compiles /root/reference/thirdparty/sophus/se3.hpp + so3.hpp UNMODIFIED (through
oracle/sophus_capi.cpp) over oracle/eigenshim — a functional stand-in for the un-vendored Eigen that follows Eigen's generic
evaluation order — into oracle/_ref/libref_sophus.so; tests/golden/se3_ref.npz holds what it computed
(tests/golden/make_se3_golden.py).  The oracle's restatement (vso_se3_*) and the product's host helpers (vsb_se3_*, the same
se3.cuh the kernels use) must reproduce it BIT FOR BIT: exp with its Taylor branch (theta < 1e-5, so3.hpp:534-568,
se3.hpp:733-735), the product with the first-order renormalisation (so3.hpp:338-353, se3.hpp:285-321), matrix
(se3.hpp:253-268), SE3(R, t) (so3.hpp:422-427 + Eigen's Quaternion(Matrix3)).  Where the library is present (this container
and, because oracle/_ref travels, the GPU box) the comparison is repeated live on fresh random inputs, and the native-libm
build shows the one platform dependency: std::sin / std::cos on floats."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "se3_ref.npz")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_sophus.so")
REF_LIBM_SO = os.path.join(ROOT, "oracle", "_ref", "libref_sophus_libm.so")
N_GOLDEN = 2500
FP = C.POINTER(C.c_float)


def _p(a):
    return a.ctypes.data_as(FP)


def taylor_mask(delta):
    w = delta[:, 3:].astype(np.float32)
    th = np.sqrt((w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1] + w[:, 2] * w[:, 2]).astype(np.float32)).astype(np.float32)
    return th < np.float32(1e-5)


def cases(n, seed):
    """Tangent vectors over eight decades of magnitude; every seventh has a rotation below Sophus' epsilon (Taylor branch)."""
    rng = np.random.default_rng(seed)
    da = np.zeros((n, 6), np.float32)
    db = np.zeros((n, 6), np.float32)
    for i in range(n):
        s = 10.0 ** rng.uniform(-8, 0.5)
        da[i] = rng.standard_normal(6) * s
        db[i] = rng.standard_normal(6) * 10.0 ** rng.uniform(-8, 0.5)
        if i % 7 == 0:
            da[i, 3:] *= np.float32(1e-4)
        if i % 11 == 0:
            db[i, 3:] = 0
    da[0] = 0
    return dict(delta_a=da, delta_b=db)


def run(L, prefix, g):
    """exp(a), exp(b), exp(a) * exp(b), matrix of the product, SE3(R, t) of that matrix — through the library's C entries
    `<prefix>_se3_exp / _mul / _matrix / _from_rt` (sph_: the reference's Sophus; vso_: oracle; vsb_: product)."""
    n = g["delta_a"].shape[0]
    ea, eb, pr = (np.zeros((n, 7), np.float32) for _ in range(3))
    mat = np.zeros((n, 16), np.float32)
    frt = np.zeros((n, 7), np.float32)
    ok = np.zeros(n, np.int32)
    renorm = np.zeros(n, bool)
    f = lambda name: getattr(L, prefix + "_" + name)
    for i in range(n):
        da, db = np.ascontiguousarray(g["delta_a"][i]), np.ascontiguousarray(g["delta_b"][i])
        f("se3_exp")(_p(da), _p(ea[i]))
        f("se3_exp")(_p(db), _p(eb[i]))
        f("se3_mul")(_p(ea[i]), _p(eb[i]), _p(pr[i]))
        f("se3_matrix")(_p(pr[i]), _p(mat[i]))
        m = mat[i].reshape(4, 4)
        R, t = np.ascontiguousarray(m[:3, :3]).reshape(-1), np.ascontiguousarray(m[:3, 3])
        rc = f("se3_from_rt")(_p(R), _p(t), _p(frt[i]))
        ok[i] = 1 if rc == 0 else 0
        # whether SO3::operator*= renormalised: the raw quaternion product has squared norm != 1 in float
        a, b = ea[i], eb[i]
        aw, ax, ay, az, bw, bx, by, bz = (np.float32(v) for v in (a[3], a[0], a[1], a[2], b[3], b[0], b[1], b[2]))
        qw = aw * bw - ax * bx - ay * by - az * bz
        qx = aw * bx + ax * bw + ay * bz - az * by
        qy = aw * by + ay * bw + az * bx - ax * bz
        qz = aw * bz + az * bw + ax * by - ay * bx
        renorm[i] = np.float32(np.float32(np.float32(qx * qx + qy * qy) + qz * qz) + qw * qw) != np.float32(1)
    return dict(exp_a=ea, exp_b=eb, prod=pr, matrix=mat, from_rt=frt, from_rt_ok=ok, renorm=renorm)


def _oracle_handle(oracle):
    oracle.lib()                                                    # builds libvso.so on demand
    return C.CDLL(os.path.join(ROOT, "oracle", "libvso.so"))         # a handle of our own: plain pointer arguments


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _check(out, gold, what):
    for k in ("exp_a", "exp_b", "prod", "matrix"):
        bad = np.nonzero((_bits(out[k]) != _bits(gold["out_" + k])).any(axis=1))[0]
        assert bad.size == 0, f"{what}: {k} differs from the reference's Sophus in {bad.size} cases, first {bad[:3]}"
    ok = gold["out_from_rt_ok"] == 1
    bad = np.nonzero((_bits(out["from_rt"])[ok] != _bits(gold["out_from_rt"])[ok]).any(axis=1))[0]
    assert bad.size == 0, f"{what}: SE3(R, t) differs in {bad.size} cases"


def _load_golden():
    g = np.load(GOLDEN)
    return g, dict(delta_a=g["delta_a"], delta_b=g["delta_b"])


def test_golden_covers_the_branches():
    g, _ = _load_golden()
    assert taylor_mask(g["delta_a"]).sum() > 300          # theta < 1e-5: Taylor expansion and V = R
    assert g["out_renorm"].sum() > 400                    # squared norm != 1: first-order renormalisation
    assert (~g["out_renorm"]).sum() > 300                  # ... and the branch that leaves the product alone
    assert g["out_from_rt_ok"].mean() > 0.9                # Sophus' orthogonality precondition holds for its own matrices


def test_oracle_se3_matches_the_references_sophus(oracle):
    g, inp = _load_golden()
    _check(run(_oracle_handle(oracle), "vso", inp), g, "oracle")


def test_product_se3_matches_the_references_sophus():
    """vsb_se3_exp / _mul / _matrix / _from_rt are host-side entries of the product built from csrc/se3.cuh — the same
    header the Gauss-Newton kernels use for pose <- pose * exp(delta).  No device needed."""
    import vislam_b200 as vb
    g, inp = _load_golden()
    L = C.CDLL(vb.LIB_PATH)
    _check(run(L, "vsb", inp), g, "product host helpers")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_sophus.so not built (needs /root/reference)")
def test_live_against_the_references_sophus(oracle):
    inp = cases(20000, seed=int.from_bytes(os.urandom(4), "little"))
    ref = run(C.CDLL(REF_SO), "sph", inp)
    got = run(_oracle_handle(oracle), "vso", inp)
    _check(got, {"out_" + k: v for k, v in ref.items()}, "oracle (live)")


@pytest.mark.skipif(not os.path.exists(REF_LIBM_SO), reason="native-libm build of the reference's Sophus not present")
def test_native_libm_build_differs_only_through_sinf_cosf():
    """The same Sophus code on the platform's own sinf / cosf: exp results may differ from the correctly rounded build in
    the last place (glibc: about 1 % of the arguments), never by more; product, matrix and SE3(R, t) of IDENTICAL inputs
    are bit-equal — i.e. the only platform dependency of the reference's pose update is libm's single-precision sin / cos."""
    inp = cases(4000, seed=99)
    a = run(C.CDLL(REF_SO), "sph", inp)
    b = run(C.CDLL(REF_LIBM_SO), "sph", inp)
    d = np.abs(a["exp_a"].astype(np.float64) - b["exp_a"].astype(np.float64))
    scale = np.maximum(np.abs(a["exp_a"]).max(axis=1, keepdims=True), 1e-30)
    assert (d / scale).max() < 4 * 2.0 ** -23
    same = (_bits(a["exp_a"]) == _bits(b["exp_a"])).all(axis=1) & (_bits(a["exp_b"]) == _bits(b["exp_b"])).all(axis=1)
    assert same.mean() > 0.9
    for k in ("prod", "matrix"):
        assert (_bits(a[k])[same] == _bits(b[k])[same]).all()

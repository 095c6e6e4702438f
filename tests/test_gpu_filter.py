"""GPU parity: ratio + symmetry + sort + grid filter (vsb_match_filter) vs the oracle's restatement of
Matcher::computeBestMatches/getGoodMatches — exact indices, order and distances."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu_pipeline(ctx, d1, d2, kp1, w, h, n_cells, sym_mode=0):
    import torch
    t1, t2 = torch.from_numpy(d1).cuda()[None], torch.from_numpy(d2).cuda()[None]
    i12, s12, i21, s21 = ctx.knn2_hamming(t1, t2)
    kp = torch.from_numpy(np.ascontiguousarray(kp1, np.float32)).cuda()[None]
    gq, gt, gd, ng, ns = ctx.match_filter(i12, s12, i21, s21, kp, w, h, n_cells, sym_mode=sym_mode)
    torch.cuda.synchronize()
    n = int(ng[0])
    return gq[0, :n].cpu().numpy(), gt[0, :n].cpu().numpy(), gd[0, :n].cpu().numpy(), int(ns[0])


@pytest.mark.parametrize("n_cells", [49, 225, 1, 50, 4096])
@pytest.mark.parametrize("sym_mode", [0, 1])
def test_filter_matches_oracle(ctx, oracle, n_cells, sym_mode):
    from vislam_b200 import synth
    rng = np.random.default_rng(n_cells)
    n = 1000
    d1 = synth.orb_descriptors(n, 3)
    d2, _ = synth.perturb_orb(d1, 4)
    kp1 = np.stack([rng.uniform(0, 751, n), rng.uniform(0, 479, n)], 1).astype(np.float32)
    kp1[::7, 1] = np.floor(kp1[::7, 1])       # equal-y ties (sort stability) and cell-edge hits
    kp1[::11, 0] = np.floor(kp1[::11, 0] / 107.4) * 107.42857
    eq, et, ed, ens = oracle.match_pipeline(d1, d2, kp1, 752, 480, n_cells, 1, mode=sym_mode)
    gq, gt, gd, gns = _gpu_pipeline(ctx, d1, d2, kp1, 752, 480, n_cells, sym_mode)
    assert gns == ens
    np.testing.assert_array_equal(gq, eq)
    np.testing.assert_array_equal(gt, et)
    np.testing.assert_array_equal(gd, ed)


def test_filter_equal_distance_cells(ctx, oracle):
    """Many matches with identical distance in one cell: winner = smallest y, then smallest query index."""
    rng = np.random.default_rng(2)
    n = 400
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    d2 = d1.copy()                                     # all matches at distance 0
    kp1 = np.stack([rng.integers(0, 752, n), rng.integers(0, 480, n) // 40 * 40], 1).astype(np.float32)
    eq, et, ed, ens = oracle.match_pipeline(d1, d2, kp1, 752, 480, 49, 1)
    gq, gt, gd, gns = _gpu_pipeline(ctx, d1, d2, kp1, 752, 480, 49)
    assert gns == ens == n
    np.testing.assert_array_equal(gq, eq)
    np.testing.assert_array_equal(gt, et)


def test_filter_empty_and_tiny(ctx, oracle):
    rng = np.random.default_rng(3)
    for n1, n2 in [(1, 1), (2, 2), (5, 1), (1, 5)]:
        d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        kp1 = rng.uniform(1, 400, (n1, 2)).astype(np.float32)
        eq, et, ed, ens = oracle.match_pipeline(d1, d2, kp1, 752, 480, 49, 1)
        gq, gt, gd, gns = _gpu_pipeline(ctx, d1, d2, kp1, 752, 480, 49)
        assert gns == ens
        np.testing.assert_array_equal(gq, eq)
        np.testing.assert_array_equal(gt, et)

"""GPU parity of the whole tracking step (vsb_track_pairs / vsb_track_sequence / _host) vs the oracle's
track_pair: match -> candidates -> GN chained on the device without host round trips."""
import numpy as np
import pytest

from test_gpu_gn import TOL, rot_angle

pytestmark = pytest.mark.gpu


def _prior(vb_mod, R_res, t_res):
    import ctypes as C
    out = (C.c_float * 7)()
    eye = (C.c_float * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1)
    r = (C.c_float * 9)(*[float(x) for x in np.asarray(R_res, np.float32).reshape(-1)])
    t = (C.c_float * 3)(*[float(x) for x in t_res])
    assert vb_mod.lib().vsb_initial_pose(eye, r, t, out) == 0
    return np.array(out[:], np.float32)


@pytest.mark.parametrize("n_cells", [49, 225])
def test_track_pairs_vs_oracle(ctx, oracle, n_cells):
    import torch
    from vislam_b200 import synth
    pairs = [synth.make_pair(n_feat=400, seed=s) for s in (1001, 1777)]
    tr = ctx.tracker(752, 480, 400, pairs[0]["K"], n_cells=n_cells, max_pairs=4)
    st = lambda k, dt=None: torch.from_numpy(np.stack([p[k] for p in pairs])).cuda()
    pose, n_good = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
    torch.cuda.synchronize()
    pose, n_good = pose.cpu().numpy(), n_good.cpu().numpy()
    for b, p in enumerate(pairs):
        ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=n_cells)
        assert n_good[b] == len(ref["good_q"])
        assert rot_angle(pose[b][:4], ref["pose"][:4]) <= TOL
        assert np.abs(pose[b][4:] - ref["pose"][4:]).max() <= TOL
    tr.close()


@pytest.mark.parametrize("grad_mode", [0, 1])
def test_track_sequence_device_and_host(ctx, oracle, grad_mode):
    """A 6-frame synthetic sequence: device-resident entry, host entry (chunked, two streams) and the oracle agree."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    T, N = 6, 300
    seq = synth.make_sequence(T, n_feat=N, seed=2001)
    prior = np.stack([_prior(vb, seq["R_imu_res"][k], seq["t_res"][k]) for k in range(T - 1)])
    for k in range(T - 1):   # the product's host helper and the oracle form the same initial pose
        np.testing.assert_array_equal(prior[k], oracle.initial_pose(np.eye(3), seq["R_imu_res"][k], seq["t_res"][k]))
    tr = ctx.tracker(752, 480, N, seq["K"], n_cells=49, max_pairs=2, gn_opts=vb.default_gn_opts(grad_mode=grad_mode))
    frames, desc, kp = seq["frames"], seq["desc"], seq["kp"]
    # host entry: 5 pairs with max_pairs = 2 -> three chunks over two slots
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_pose = torch.zeros((T - 1, 7), dtype=torch.float32).pin_memory()
    h_ng = torch.zeros((T - 1,), dtype=torch.int32).pin_memory()
    tr.track_sequence_host(pin(frames), pin(desc), pin(kp), pin(prior), h_pose, h_ng)
    # device entry, 2 pairs at a time
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_pose = []
    for k in range(0, T - 1, 2):
        e = min(k + 3, T)
        pz, _ = tr.track_sequence(dev(frames[k:e]), dev(desc[k:e]), dev(kp[k:e]), dev(prior[k:e - 1]))
        d_pose.append(pz.cpu().numpy())
    d_pose = np.concatenate(d_pose)
    np.testing.assert_array_equal(d_pose, h_pose.numpy())
    for k in range(T - 1):
        ref = oracle.track_pair(frames[k], frames[k + 1], desc[k], desc[k + 1], kp[k], seq["K"], prior[k], n_cells=49)
        assert h_ng[k] == len(ref["good_q"])
        assert rot_angle(h_pose[k][:4].numpy(), ref["pose"][:4]) <= TOL
        assert np.abs(h_pose[k][4:].numpy() - ref["pose"][4:]).max() <= TOL
    tr.close()


def test_track_pairs_float_descriptors_tum_shaped(ctx, oracle):
    """BASELINE configs[2]: 640x480 TUM-shaped pair, 64-d float descriptors, L2 kNN (Matcher.cpp:55)."""
    import torch
    from vislam_b200 import synth
    pairs = [synth.make_pair(w=640, h=480, n_feat=500, K=synth.TUM_K, seed=s, desc="float") for s in (3001, 3055)]
    tr = ctx.tracker(640, 480, 500, pairs[0]["K"], n_cells=49, max_pairs=2, norm=0, desc_bytes=64 * 4)
    st = lambda k: torch.from_numpy(np.stack([p[k] for p in pairs])).cuda()
    d1 = st("d1").view(torch.uint8).reshape(2, 500, 256)
    d2 = st("d2").view(torch.uint8).reshape(2, 500, 256)
    pose, n_good = tr.track_pairs(st("prev"), st("cur"), d1, d2, st("kp1"), st("pose_prior"))
    torch.cuda.synchronize()
    pose, n_good = pose.cpu().numpy(), n_good.cpu().numpy()
    for b, p in enumerate(pairs):
        ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=49,
                                norm=0)
        assert n_good[b] == len(ref["good_q"]) and n_good[b] > 10
        assert rot_angle(pose[b][:4], ref["pose"][:4]) <= TOL
        assert np.abs(pose[b][4:] - ref["pose"][4:]).max() <= TOL
    tr.close()


@pytest.mark.parametrize("first_lvl", [3, 4])
def test_track_pairs_kitti_shaped(ctx, oracle, first_lvl):
    """BASELINE configs[3]: 1241x376 (odd width: clipped 2x2 blocks in the pyramid), 2000 ORB features, GN start level
    3 (reference literal, VISystem.cpp:1119) or 4 (5-level config)."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    p = synth.make_pair(w=1241, h=376, n_feat=2000, K=synth.KITTI_K, seed=4001)
    tr = ctx.tracker(1241, 376, 2000, p["K"], n_cells=225, max_pairs=1, gn_opts=vb.default_gn_opts(first_lvl=first_lvl))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()[None]
    pose, n_good = tr.track_pairs(dev(p["prev"]), dev(p["cur"]), dev(p["d1"]), dev(p["d2"]), dev(p["kp1"]),
                                  dev(p["pose_prior"]))
    torch.cuda.synchronize()
    ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=225,
                            opts=oracle.default_opts(first_lvl=first_lvl))
    pose = pose[0].cpu().numpy()
    assert int(n_good[0]) == len(ref["good_q"]) and int(n_good[0]) > 50
    assert rot_angle(pose[:4], ref["pose"][:4]) <= TOL
    assert np.abs(pose[4:] - ref["pose"][4:]).max() <= TOL
    tr.close()


def test_track_pairs_batched_5000_features_config5(ctx, oracle):
    """BASELINE configs[4] shape (independent 752x480 pairs, 5000 ORB features each, seeds 5000+i), scaled to a
    96-pair batch: two distinct pairs are checked against the oracle, and — the size-independent property — a pair's
    result does not depend on where it sits in the batch or on what surrounds it (bit-identical poses)."""
    import torch
    from vislam_b200 import synth
    uniq = [synth.make_pair(n_feat=5000, seed=5000 + i) for i in range(3)]
    B = 96
    order = [i % 3 for i in range(B)]
    order[1], order[2] = 0, 0                       # neighbours differ from the i % 3 pattern
    tr = ctx.tracker(752, 480, 5000, uniq[0]["K"], n_cells=225, max_pairs=B)
    st = lambda k: torch.from_numpy(np.stack([uniq[o][k] for o in order])).cuda()
    pose, n_good = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
    torch.cuda.synchronize()
    pose, n_good = pose.cpu().numpy(), n_good.cpu().numpy()
    first = {}
    for b, o in enumerate(order):
        if o in first:
            np.testing.assert_array_equal(pose[b], pose[first[o]])
            assert n_good[b] == n_good[first[o]]
        else:
            first[o] = b
    for o in (0, 1):
        p = uniq[o]
        ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=225)
        b = first[o]
        assert n_good[b] == len(ref["good_q"]) and n_good[b] > 100
        assert rot_angle(pose[b][:4], ref["pose"][:4]) <= TOL
        assert np.abs(pose[b][4:] - ref["pose"][4:]).max() <= TOL
    tr.close()


@pytest.mark.parametrize("kw", [dict(grad_mode=1, weight_mode=2, huber_k=12.0), dict(grad_mode=1, weight_mode=1),
                                dict(grad_mode=0), dict(grad_mode=0, weight_mode=2, huber_k=12.0),
                                dict(grad_mode=1, sample_mode=1)])
def test_track_pairs_solver_modes(ctx, oracle, kw):
    """The tracker's stage wiring for every solver mode: the fused candidate/attribute pass with the general point
    records (weights other than identity), the materialised-gradient path, bilinear sampling — each against the oracle
    run with the same options."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    p = synth.make_pair(n_feat=400, seed=1313)
    tr = ctx.tracker(752, 480, 400, p["K"], n_cells=49, max_pairs=2, gn_opts=vb.default_gn_opts(**kw))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()[None]
    pose, n_good = tr.track_pairs(dev(p["prev"]), dev(p["cur"]), dev(p["d1"]), dev(p["d2"]), dev(p["kp1"]),
                                  dev(p["pose_prior"]))
    torch.cuda.synchronize()
    okw = {k: v for k, v in kw.items() if k != "grad_mode"}
    ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=49,
                            opts=oracle.default_opts(**okw))
    pose = pose[0].cpu().numpy()
    assert int(n_good[0]) == len(ref["good_q"])
    assert rot_angle(pose[:4], ref["pose"][:4]) <= TOL
    assert np.abs(pose[4:] - ref["pose"][4:]).max() <= TOL
    tr.close()


@pytest.mark.parametrize("n_feat_max", [1500, 300])
def test_track_sequence_from_raw_frames(ctx, oracle, n_feat_max):
    """vsb_track_sequence_orb: frames only.  cv::ORB::create(1000) on the device feeds the matcher and the solver; poses are
    bit-identical to the oracle chain (oracle ORB -> oracle matcher -> oracle GN) on the same frames, also when the key-point
    capacity truncates a frame's list."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    T = 5
    seq = synth.make_sequence(T, n_feat=10, seed=2001)
    prior = np.stack([_prior(vb, seq["R_imu_res"][k], seq["t_res"][k]) for k in range(T - 1)])
    tr = ctx.tracker(752, 480, n_feat_max, seq["K"], n_cells=49, max_pairs=T - 1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pose, n_good, n_feat = tr.track_sequence_orb(dev(seq["frames"]), dev(prior), nfeatures=1000)
    torch.cuda.synchronize()
    pose, n_good, n_feat = pose.cpu().numpy(), n_good.cpu().numpy(), n_feat.cpu().numpy()
    feats = [oracle.orb_detect_compute_pyr(seq["frames"][t], 1000) for t in range(T)]
    for t in range(T):
        assert n_feat[t] == min(len(feats[t][0]), n_feat_max)
    for k in range(T - 1):
        a, b = feats[k], feats[k + 1]
        ref = oracle.track_pair(seq["frames"][k], seq["frames"][k + 1], a[4][:n_feat_max], b[4][:n_feat_max], a[0][:n_feat_max],
                                seq["K"], prior[k], n_cells=49)
        assert n_good[k] == len(ref["good_q"]) > 10
        np.testing.assert_array_equal(pose[k], ref["pose"])
    tr.close()


@pytest.mark.parametrize("max_pairs", [2, 3, 8])
def test_track_sequence_orb_host_matches_device_entry(ctx, max_pairs):
    """vsb_track_sequence_orb_host (frames in pinned HOST memory, chunks of max_pairs pairs: copies on one stream, ORB + match +
    GN on the other) returns the bits of vsb_track_sequence_orb on the whole sequence resident on the device — 7 pairs in 4, 3
    and 1 chunks (a chunk's last frame is described again as the next chunk's first)."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    T = 8
    seq = synth.make_sequence(T, n_feat=10, seed=2002)
    prior = np.stack([_prior(vb, seq["R_imu_res"][k], seq["t_res"][k]) for k in range(T - 1)])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tr = ctx.tracker(752, 480, 1200, seq["K"], n_cells=49, max_pairs=T - 1)
    pose_d, ng_d, nf_d = tr.track_sequence_orb(dev(seq["frames"]), dev(prior), nfeatures=1000)
    torch.cuda.synchronize()
    tr.close()
    tr = ctx.tracker(752, 480, 1200, seq["K"], n_cells=49, max_pairs=max_pairs)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pose = torch.zeros((T - 1, 7), dtype=torch.float32).pin_memory()
    ng = torch.zeros((T - 1,), dtype=torch.int32).pin_memory()
    nf = torch.zeros((T,), dtype=torch.int32).pin_memory()
    for _ in range(2):                                           # the second call reuses the slots' buffers
        tr.track_sequence_orb_host(pin(seq["frames"]), pin(prior), pose, nfeatures=1000, n_good=ng, n_feat=nf)
        np.testing.assert_array_equal(pose.numpy(), pose_d.cpu().numpy())
        np.testing.assert_array_equal(ng.numpy(), ng_d.cpu().numpy())
        np.testing.assert_array_equal(nf.numpy(), nf_d.cpu().numpy())
    ht = tr.host_traffic()
    assert ht["chunks"] == {2: 4, 3: 3, 8: 1}[max_pairs]
    assert ht["h2d"] == (T - 1 + ht["chunks"]) * 752 * 480 + (T - 1) * 28
    tr.close()


@pytest.mark.parametrize("n_cells,threads,stage", [(49, 0, 8192), (49, 128, 0), (49, 256, 32768), (49, 512, 8192),
                                                   (225, 1024, 0), (225, 128, 32768), (225, 0, 8192)])
def test_tracker_solver_per_iteration(ctx, oracle, n_cells, threads, stage):
    """The tracker's own Gauss-Newton kernel (gn_track.cu: table-fed points, staged coarse levels) against the oracle PER
    ITERATION — the north-star tolerance — for every block size and with / without the shared-memory staged level."""
    import torch
    from vislam_b200 import synth
    pairs = [synth.make_pair(n_feat=400, seed=s) for s in (1001, 1777, 4242)]
    ctx.option("gn_threads", threads)
    ctx.option("gn_stage_bytes", stage)
    try:
        tr = ctx.tracker(752, 480, 400, pairs[0]["K"], n_cells=n_cells, max_pairs=3)
        tr.trace_on()
        st = lambda k: torch.from_numpy(np.stack([p[k] for p in pairs])).cuda()
        pose, n_good = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
        traces = tr.traces(len(pairs))
        pose = pose.cpu().numpy()
        from test_gpu_gn import compare_traces
        for b, p in enumerate(pairs):
            ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=n_cells)
            compare_traces(traces[b], ref["trace"])
            np.testing.assert_array_equal(pose[b], ref["pose"])      # in fact the same bits
        tr.close()
    finally:
        ctx.option("gn_threads", 0)
        ctx.option("gn_stage_bytes", 8192)


@pytest.mark.parametrize("n_cells,cluster,cthreads,stage", [(49, 1, 512, 8192), (49, 2, 256, 0), (49, 4, 512, 32768),
                                                            (49, 8, 256, 8192), (225, 1, 512, 8192), (225, 8, 256, 0),
                                                            (225, 2, 512, 32768), (225, 4, 256, 8192)])
def test_tracker_solver_cluster_per_iteration(ctx, oracle, n_cells, cluster, cthreads, stage):
    """Small batches: one frame pair per thread-block cluster (gn_track.cu, CLUSTER kernels — partial Gram matrices summed
    out of the peers' shared memory, pose pushed back into it) against the oracle PER ITERATION, for every cluster size,
    both block sizes, with / without the staged level; and the same bits as the one-block-per-pair kernel."""
    import torch
    from vislam_b200 import synth
    pairs = [synth.make_pair(n_feat=400, seed=s) for s in (1001, 1777, 4242)]
    st = lambda k: torch.from_numpy(np.stack([p[k] for p in pairs])).cuda()
    ctx.option("gn_stage_bytes", stage)
    try:
        out = {}
        for cl in (cluster, 0):
            ctx.option("gn_cluster", cl)
            ctx.option("gn_cluster_threads", cthreads)
            tr = ctx.tracker(752, 480, 400, pairs[0]["K"], n_cells=n_cells, max_pairs=3)
            tr.trace_on()
            pose, n_good = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
            out[cl] = (pose.cpu().numpy(), tr.traces(len(pairs)), tr.stats())
            tr.close()
        from test_gpu_gn import compare_traces
        pose, traces, stats = out[cluster]
        for b, p in enumerate(pairs):
            ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=n_cells)
            compare_traces(traces[b], ref["trace"])
            np.testing.assert_array_equal(pose[b], ref["pose"])
        np.testing.assert_array_equal(pose, out[0][0])
        assert stats == out[0][2]
    finally:
        ctx.option("gn_cluster", 1)
        ctx.option("gn_cluster_threads", 0)
        ctx.option("gn_stage_bytes", 8192)


@pytest.mark.parametrize("n_cells", [49, 225])
def test_tracker_solver_matches_general_kernel(ctx, n_cells):
    """gn_track.cu (with and without merging coincident candidate points, with the Gram matrix on the FP64 tensor cores or
    in registers) and gn_solve.cu give the same bits for the same pairs."""
    import torch
    from vislam_b200 import synth
    pairs = [synth.make_pair(n_feat=600, seed=s) for s in (31, 32, 33, 34)]
    st = lambda k: torch.from_numpy(np.stack([p[k] for p in pairs])).cuda()
    out = []
    for impl, dedup, variant in ((1, 1, 0), (1, 0, 0), (1, 0, 1), (0, 0, 0)):
        ctx.option("gn_impl", impl)
        ctx.option("gn_dedup", dedup)
        ctx.option("gn_variant", variant)
        try:
            tr = ctx.tracker(752, 480, 600, pairs[0]["K"], n_cells=n_cells, max_pairs=4)
            pose, _ = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
            out.append(pose.cpu().numpy())
            stats = tr.stats()
            tr.close()
        finally:
            ctx.option("gn_impl", 1)
            ctx.option("gn_dedup", 1)
            ctx.option("gn_variant", 0)
        if len(out) > 1:
            np.testing.assert_array_equal(out[0], out[-1])
            assert stats == first_stats          # the work counters count candidate points before merging
        else:
            first_stats = stats


def test_tracker_tail_launch_same_bits(ctx):
    """A batch larger than one wave of the solver (6 blocks per SM): the pairs of the last partial wave run in a second
    launch with more threads each — every copy of a pair still gets the same bits as its first occurrence."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    sms = vb.lib().vsb_sm_count(ctx.handle)
    uniq = [synth.make_pair(n_feat=300, seed=900 + i) for i in range(3)]
    B = 6 * sms + 67
    order = [i % 3 for i in range(B)]
    tr = ctx.tracker(752, 480, 300, uniq[0]["K"], n_cells=49, max_pairs=B)
    dev = {k: torch.from_numpy(np.stack([u[k] for u in uniq])).cuda() for k in ("prev", "cur", "d1", "d2", "kp1", "pose_prior")}
    idx = torch.tensor(order, device="cuda")
    st = lambda k: dev[k][idx].contiguous()
    n0 = ctx.launches
    pose, n_good = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
    torch.cuda.synchronize()
    launches = ctx.launches - n0
    ctx.option("gn_tail", 0)
    try:
        n1 = ctx.launches
        pose1, _ = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
        torch.cuda.synchronize()
        assert launches == (ctx.launches - n1) + 1          # the tail was a launch of its own
    finally:
        ctx.option("gn_tail", 1)
    pose, pose1 = pose.cpu().numpy(), pose1.cpu().numpy()
    np.testing.assert_array_equal(pose, pose1)
    for b, o in enumerate(order):
        np.testing.assert_array_equal(pose[b], pose[o])
    tr.close()


def test_track_sequence_host_small_max_pairs(ctx, oracle):
    """Host entry with max_pairs below the scheduler's 32-pair floor and more than 32 pairs left: chunks must stay inside the
    slot buffers (they are sized for max_pairs)."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    T, N = 42, 200
    base = synth.make_sequence(6, n_feat=N, seed=2001)
    rep = lambda a: np.concatenate([a] * 7)[:T]
    frames, desc, kp = rep(base["frames"]), rep(base["desc"]), rep(base["kp"])
    prior = np.stack([_prior(vb, base["R_imu_res"][k % 5], base["t_res"][k % 5]) for k in range(T - 1)])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    res = []
    for mp in (16, 64):
        tr = ctx.tracker(752, 480, N, base["K"], n_cells=49, max_pairs=mp)
        h_pose = torch.zeros((T - 1, 7), dtype=torch.float32).pin_memory()
        tr.track_sequence_host(pin(frames), pin(desc), pin(kp), pin(prior), h_pose)
        assert tr.host_traffic()["chunks"] >= (T - 1 + mp - 1) // mp
        res.append(h_pose.numpy().copy())
        tr.close()
    np.testing.assert_array_equal(res[0], res[1])
    ref = oracle.track_pair(frames[0], frames[1], desc[0], desc[1], kp[0], base["K"], prior[0], n_cells=49)
    np.testing.assert_array_equal(res[0][0], ref["pose"])


@pytest.mark.parametrize("weight_mode", [0, 1])
def test_track_sequence_host_float_descriptors_three_chunks(ctx, weight_mode):
    """Host entry (two streams, consecutive chunks in flight together) with L2 descriptors — the tensor-core L2 kNN's
    workspace — and with Tukey weights — the residual scratch — must not be shared between the streams: same poses as the
    device entry run chunk by chunk."""
    import torch
    import vislam_b200 as vb
    from vislam_b200 import synth
    T, N = 8, 300
    rng = np.random.default_rng(77)
    base = synth.make_sequence(T, n_feat=N, seed=2005)
    d0 = rng.standard_normal((N, 64)).astype(np.float32)
    desc = []
    for t in range(T):
        d = d0 + 0.05 * rng.standard_normal((N, 64)).astype(np.float32)
        desc.append(d / np.linalg.norm(d, axis=1, keepdims=True))
    desc = np.ascontiguousarray(np.stack(desc).astype(np.float32)).view(np.uint8).reshape(T, N, 256)
    prior = np.stack([_prior(vb, base["R_imu_res"][k], base["t_res"][k]) for k in range(T - 1)])
    tr = ctx.tracker(752, 480, N, base["K"], n_cells=49, max_pairs=2, norm=0, desc_bytes=256,
                     gn_opts=vb.default_gn_opts(grad_mode=1, weight_mode=weight_mode))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for rep in range(3):
        h_pose = torch.zeros((T - 1, 7), dtype=torch.float32).pin_memory()
        tr.track_sequence_host(pin(base["frames"]), pin(desc), pin(base["kp"]), pin(prior), h_pose)
        d_pose = []
        for k in range(0, T - 1, 2):
            e = min(k + 3, T)
            pz, _ = tr.track_sequence(dev(base["frames"][k:e]), dev(desc[k:e]), dev(base["kp"][k:e]), dev(prior[k:e - 1]))
            torch.cuda.synchronize()
            d_pose.append(pz.cpu().numpy())
        np.testing.assert_array_equal(np.concatenate(d_pose), h_pose.numpy())
    tr.close()

#!/usr/bin/env python
"""bench.py — frames/s tracked @752x480 (kNN match + GN pose solve) on synthetic EuRoC-shaped data.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  * a "step" = one pass of the frame-tracking hot path over the whole workload: BASELINE.json configs[1],
    a 2000-frame 752x480 sequence, 1000 ORB descriptors per frame, 200 Hz gyro prior — 1999 frame pairs —
    pyramid -> Hamming kNN-2 (both directions) -> ratio/symmetry/grid filter -> candidates -> 4-level GN.
  * value   = frames/s with the sequence already resident in HBM (CUDA events on the launching stream).
  * e2e     = the same pass through the host-buffer C-ABI entry (vsb_track_sequence_host): pinned host
              frames/descriptors/key points/priors in, poses out, H2D and D2H inside the timed region.
  * configs = the other four BASELINE configs, each with its own device-resident figure, end-to-end figure, per-kernel
              times, roofline entry and the CPU port beside it: configs[0] single-pair latency, configs[2] TUM float
              descriptors (tensor-core L2 kNN), configs[3] KITTI 5000 features / 5 levels (these three at N = 1 only),
              configs[4] 8192 independent pairs x 5000 features SHARDED over the N ranks in contiguous blocks
              (replicas.shard_range; strong scaling, aggregate pairs/s).
  * N > 1   = N independent replicas of configs[1] (one process per GPU, different sequence seed per rank), no
              data-path collective ("replicas only"); torch.distributed is used only for the barrier, the
              max-over-ranks of the timed regions and the per-rank statistics.
  * --impl reference = the reference's CPU algorithm on all host threads over the same 2000-frame workload: the oracle
              port (oracle/libvso.so), plus a bounded sample of the reference's own translation units compiled against
              the OpenCV stand-in (oracle/_ref/libref_visystem.so).  Loads nothing of the product.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "vi-slam_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

from vislam_b200 import workloads as wl  # noqa: E402  (input generation only; loads no native code)

METRIC = "frames/s tracked @752x480 (kNN match + GN pose solve)"
DTYPE = "u8/int32 Hamming + f32 GN (f64 accumulate)"
N_FRAMES = wl.CFG1["frames"]


def base_config(world):
    """The `config` object — identical in both arms (the driver compares them)."""
    return {"workload": "configs[1]: " + wl.CFG1["what"], "frames": N_FRAMES, "pairs_per_step": N_FRAMES - 1,
            "features_per_frame": wl.CFG1["n_feat"], "num_cells": wl.CFG1["n_cells"],
            "l2_policy": "inputs larger than L2 (722 MB of frames per step), no flush needed",
            "parallelism": f"replicas x{world}" if world > 1 else "single GPU"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class NvmlSampler:
    """SM clock / throttle reasons sampled every few ms DURING the timed region through NVML (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; nvidia-smi's own loop is too slow to see a
    sub-second region)."""

    def __init__(self, gpu_index):
        self.gpu, self.th, self.stop_flag = gpu_index, None, False
        self.sm, self.reasons, self.max_sm, self.power = [], 0, None, []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return False
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()
        return True

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None,
                "reasons": sorted(k for k, bit in names.items() if self.reasons & bit), "source": "nvml"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (fallback when NVML is not importable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU legs (oracle = checker)
def cpu_track(seq, cfg, pair_ids, threads, first_lvl=None):
    """The CPU oracle (oracle/libvso.so — test infrastructure, used here only as the measured CPU baseline and as the parity
    checker) on the given frame pairs of a sequence, `threads` pairs in flight.  Returns (pairs/s, seconds, poses)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import vso
    vso.lib()
    opts = vso.default_opts(first_lvl=cfg["first_lvl"] if first_lvl is None else first_lvl)

    def one(k):
        return vso.track_pair(seq["frames"][k], seq["frames"][k + 1], seq["desc"][k], seq["desc"][k + 1], seq["kp"][k],
                              cfg["K"], seq["prior"][k], n_cells=cfg["n_cells"], norm=cfg.get("norm", 1), opts=opts)["pose"]

    t0 = time.perf_counter()
    if threads <= 1:
        poses = [one(k) for k in pair_ids]
    else:
        with ThreadPoolExecutor(threads) as ex:
            poses = list(ex.map(one, pair_ids))
    dt = time.perf_counter() - t0
    return len(pair_ids) / dt, dt, np.stack(poses)


def cpu_track_pairs(pairs, cfg, threads):
    """The same for independent pairs given as a list of dicts (prev, cur, d1, d2, kp1, prior)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import vso
    vso.lib()
    opts = vso.default_opts(first_lvl=cfg["first_lvl"])

    def one(p):
        return vso.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], cfg["K"], p["prior"], n_cells=cfg["n_cells"],
                              norm=cfg.get("norm", 1), opts=opts)["pose"]

    t0 = time.perf_counter()
    if threads <= 1:
        poses = [one(p) for p in pairs]
    else:
        with ThreadPoolExecutor(threads) as ex:
            poses = list(ex.map(one, pairs))
    dt = time.perf_counter() - t0
    return len(pairs) / dt, dt, np.stack(poses)


def ref_units_timing(seq, cfg, n_pairs):
    """The reference's OWN translation units (src/VISystem.cpp, Camera.cpp, Matcher.cpp ... compiled unmodified by
    oracle/Makefile into oracle/_ref/libref_visystem.so) on a bounded sample, one thread (the reference is single-threaded
    and the library captures std::cout, so it is not re-entrant).  They run against the functional OpenCV stand-in
    oracle/refshim, not against OpenCV: its BFMatcher and cv::Mat are plain loops, so this figure bounds the reference's
    CPU time from above — the faster oracle port is what the ratios are taken against."""
    try:
        from oracle import ref_visystem as rv
        if not rv.available(build=False):
            return None
        rv.lib()
        K4 = np.asarray(cfg["K"], np.float32)
        eye = np.eye(3, dtype=np.float32)
        t0 = time.perf_counter()
        poses = []
        oob = 0
        for k in range(n_pairs):
            r = rv.track_pair(seq["frames"][k], seq["frames"][k + 1], K4, eye, seq["R_imu_res"][k], seq["t_res"][k],
                              kp_prev=seq["kp"][k], desc_prev=seq["desc"][k], kp_cur=seq["kp"][k + 1],
                              desc_cur=seq["desc"][k + 1], n_cells=cfg["n_cells"])
            poses.append(r["pose"].copy())
            oob += int(r["oob_reads"] > 0)
        dt = time.perf_counter() - t0
        return {"value": n_pairs / dt, "unit": "frames/s", "cores": 1, "kind": "reference",
                "sample": f"first {n_pairs} frame pairs of the same sequence ({dt:.1f}s): the reference's own src/*.cpp "
                          "(oracle/_ref/libref_visystem.so) against the OpenCV stand-in oracle/refshim — an upper bound on "
                          "its CPU time, see bench.py ref_units_timing",
                "pairs_with_out_of_bounds_reads_upstream": oob,      # parity is defined for runs without them
                "poses": np.stack(poses)}
    except Exception as e:      # informational
        return {"error": str(e), "kind": "reference"}


def cv2_matcher_timing(seq, cores):
    """OpenCV's own BFMatcher (cv2 4.13 when importable) on one frame pair of the workload — knnMatch k=2 in
    both directions, wall clock, best of 5 — with one thread and with all host threads.  Informational: the library the
    reference calls for the matching stage, next to the oracle port that is timed for the whole path."""
    try:
        import cv2
    except Exception:
        return None
    d1, d2 = np.ascontiguousarray(seq["desc"][0]), np.ascontiguousarray(seq["desc"][1])
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    out = {"version": cv2.__version__, "unit": "ms per frame pair (2 x knnMatch k=2, %d x %d ORB descriptors)" % (len(d1), len(d2))}
    for name, nt in (("threads_1", 1), ("threads_%d" % cores, cores)):
        cv2.setNumThreads(nt)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            bf.knnMatch(d1, d2, k=2)
            bf.knnMatch(d2, d1, k=2)
            best = min(best, time.perf_counter() - t0)
        out[name] = best * 1e3
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm on all host threads over the same workload as the GPU arm.  Priors come
    from the oracle's own initial_pose; nothing of the product is loaded."""
    if rank != 0:
        return
    from oracle import vso
    vso.build()
    vso.lib()
    cores = os.cpu_count() or 1
    n_frames = int(os.environ.get("VSB_BENCH_FRAMES", str(N_FRAMES)))
    seq = wl.sequence(wl.CFG1, vso.initial_pose, n_frames=n_frames, device=None, log=log)
    n_pairs = n_frames - 1
    ids = list(range(n_pairs))
    for _ in range(args.warmup):       # warm-up: page the library and the frames in, a bounded slice per round
        cpu_track(seq, wl.CFG1, ids[: 4 * cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_track(seq, wl.CFG1, ids, cores)
    dt = (time.perf_counter() - t0) / args.steps
    fps = n_pairs / dt
    sample = f"all {n_pairs} frame pairs of the workload per step, oracle port, {cores} pairs in flight"
    ru = ref_units_timing(seq, wl.CFG1, int(os.environ.get("VSB_REF_UNITS_PAIRS", "8")))
    if ru and "poses" in ru:
        _, _, port = cpu_track(seq, wl.CFG1, list(range(len(ru["poses"]))), 1)
        ru["max_abs_pose_diff_vs_port"] = float(np.abs(ru.pop("poses") - port).max())
    cfg = base_config(world)
    if n_frames != N_FRAMES:
        cfg["frames"], cfg["pairs_per_step"] = n_frames, n_pairs
    try:        # what native code this process mapped: the oracle libraries only
        maps = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ".so" in ln and ROOT in ln})
        native = [os.path.relpath(m, ROOT) for m in maps]
    except Exception:
        native = None
    print(json.dumps({
        "impl": "reference", "native_libraries_of_this_repo_loaded": native, "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                         "reference_units": ru},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
class Gpu:
    """What every leg needs: device, context, stream, barrier, measured peaks."""

    def __init__(self, rank, world, local_rank):
        import torch
        import vislam_b200 as vb
        self.torch, self.vb, self.rank, self.world, self.local_rank = torch, vb, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.ctx = vb.Context(local_rank)
        self.stream = torch.cuda.current_stream(self.dev)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        self.bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        self.bf16_src = ("measured dense bf16 (MEASURED_PEAKS.json bf16_tflops, burst)" if "bf16_tflops" in peaks
                         else "fallback 1.59 PFLOP/s dense bf16")
        self.traffic = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2b.json")))["per_step"]
        except Exception:
            pass

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_ranks(self, v):
        from vislam_b200 import replicas
        return replicas.max_over_ranks(v, self.dist, self.dev)

    def gather_ranks(self, v):
        if self.dist is None:
            return [float(v)]
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def pin(self, a):
        t = a if self.torch.is_tensor(a) else self.torch.from_numpy(np.ascontiguousarray(a))
        return t.cpu().pin_memory()


def product_initial_pose(vb):
    import ctypes as C

    def f(imu2cam, r_imu_res, t_res):      # the product's host helper (VISystem.cpp:1135-1168), bit-identical to the oracle's
        out = (C.c_float * 7)()
        a = (C.c_float * 9)(*[float(x) for x in np.asarray(imu2cam, np.float32).reshape(-1)])
        r = (C.c_float * 9)(*[float(x) for x in np.asarray(r_imu_res, np.float32).reshape(-1)])
        t = (C.c_float * 3)(*[float(x) for x in np.asarray(t_res, np.float32).reshape(-1)])
        vb.lib().vsb_initial_pose(a, r, t, out)
        return np.array(out[:], np.float32)
    return f


def timed(g, fn, reps, warmup):
    """fn() reps times between CUDA events on the launching stream, after warm-up; returns ms per call."""
    torch = g.torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(g.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(g.stream)
    for _ in range(reps):
        fn()
    e1.record(g.stream)
    torch.cuda.synchronize(g.dev)
    return e0.elapsed_time(e1) / reps


def timed_wall(g, fn, reps, warmup):
    """fn() is synchronous (host-buffer entries): wall clock per call."""
    for _ in range(warmup):
        fn()
    g.torch.cuda.synchronize(g.dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) * 1e3 / reps


def kernel_table(g, cfg, prof, steps, stats, n_pairs_per_step, n_frames_per_step, knn_impl_env):
    """Per-kernel times of a leg with the roofline entry of the kernels SURVEY 8(d) bounds.  prof: {name: (ms total, launches)}
    over `steps` steps; stats: solver counters over the same steps."""
    vb = g.vb
    N, w, h = cfg["n_feat"], cfg["w"], cfg["h"]
    px_all = wl.sum_levels(w, h, 5)
    px_gn = wl.sum_levels(w, h, cfg["first_lvl"] + 1)
    pairs_total = max(1, stats["pairs"])
    i8_peak = 2.0 * g.bf16_peak          # tcgen05 kind::i8 runs at twice the bf16 rate, kind::mxf4 at four times
    knn_fp4 = knn_impl_env in (3, 4, 5) or (knn_impl_env == 6 and N >= 768)
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    l2_ms = sum(prof.get(k, (0.0, 0))[0] for k in ("knn2_l2", "knn2_l2_prep", "knn2_l2_final"))
    out = []
    for name, (kms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ent = {"kernel": name, "ms_per_step": kms / steps, "launches_per_step": n / steps, "share": kms / total_ms}
        sec = kms * 1e-3
        if name == "gn_solve":
            visit_b = 14.0 * stats["point_visits"]           # x, y 8 B; I_prev 1; gx, gy 4; I_cur 1 per point visit
            stage_b = 6.0 * px_gn * pairs_total              # one-time staging of cur I, prev I, prev gx, gy
            ach = (visit_b + stage_b) / sec / 1e9
            ent.update({"bound": "hbm", "achieved": ach, "peak": g.hbm_peak, "unit": "GB/s", "frac": ach / g.hbm_peak,
                        "peak_source": g.hbm_src, "algorithmic_per_launch": (visit_b + stage_b) / max(1, n),
                        "split": {"per_visit_term_gbs": visit_b / sec / 1e9, "frac_per_visit_term": visit_b / sec / 1e9 / g.hbm_peak,
                                  "staging_term_gbs": stage_b / sec / 1e9,
                                  "note": "SURVEY 8(d) counts 14 B per point visit PLUS 6 B per pyramid pixel of one-time staging; "
                                          "the solver itself only moves the first term (8-byte records + the gathered byte), the "
                                          "second is paid by the pyramid and candidate kernels"},
                        "point_visits_per_s": stats["point_visits"] / sec,
                        "note": "bound as SURVEY 8(d) defines it; ncu (profiles/r2b_full_step.txt) shows the kernel limited by "
                                "instruction issue (65 % of the issue slots busy with 16 warps per SM: 128 registers hold the 28 "
                                "Gram accumulators), FP64 pipe 32 %, conversion pipe 30 %, DRAM at 3 %"})
        elif name == "knn2_hamming":
            if knn_impl_env == 0:
                work, unit, peak, src, bound = 8.0 * N * N * pairs_total, "GPOPC/s", g.ctx.popc_peak() / 1e9, "measured by vsb_popc_peak on this GPU", "int"
            else:
                work, unit, bound = 2.0 * 256 * N * N * pairs_total / 1e3, "TOP/s", "tensor"
                peak = 2.0 * i8_peak if knn_fp4 else i8_peak
                src = ("4-bit kernel (tcgen05 kind::mxf4): 4 x " if knn_fp4 else "int8 kernel (tcgen05 kind::i8): 2 x ") + g.bf16_src
            ach = work / sec / 1e9
            ent.update({"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "peak_source": src,
                        "algorithmic_per_launch": work / max(1, n),
                        "note": "one distance matrix per pair counted once (SURVEY 8d); the kernel computes it once per direction"})
            if knn_impl_env != 0:
                pp = g.ctx.popc_peak() / 1e9
                pe = 8.0 * N * N * pairs_total / sec / 1e9
                ent["survey_8d_int_pipe"] = {"bound": "int", "achieved": pe, "peak": pp, "unit": "GPOPC/s", "frac": pe / pp,
                                             "peak_source": "measured by vsb_popc_peak on this GPU"}
        elif name == "knn2_l2":
            # the whole float matcher (prep + tensor-core GEMM/top-3 + exact re-check) against the tf32 tensor peak
            D = cfg["desc_bytes"] // 4
            flop = 2.0 * N * N * D * pairs_total
            ach = flop / (l2_ms * 1e-3) / 1e12
            peak = g.bf16_peak / 2.0
            ent.update({"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "peak_source": "tf32 = half of " + g.bf16_src, "algorithmic_per_launch": flop / max(1, n),
                        "ms_incl_prep_and_recheck": l2_ms / steps, "outputs_per_s": float(N) * N * pairs_total / (l2_ms * 1e-3),
                        "note": "FLOP_alg = 2 N M D, one GEMM for both directions, the 3-product tf32 split counted once (SURVEY "
                                "8d); with K = 64 the top-3 epilogue over N M outputs is the limiter — ncu (round 1, "
                                "prof_l2tc): tensor pipe 49 %, ALU 65 %"})
        elif name == "pyramid":
            work = float(w * h + px_all) * n_frames_per_step * steps
            ach = work / sec / 1e9
            ent.update({"bound": "hbm", "achieved": ach, "peak": g.hbm_peak, "unit": "GB/s", "frac": ach / g.hbm_peak,
                        "peak_source": g.hbm_src, "algorithmic_per_launch": work / max(1, n)})
        tr = g.traffic.get(cfg["name"], {}).get(name)
        if tr is not None and "bound" in ent:
            ent["traffic"] = tr
        out.append(ent)
    return out


def leg_sequence(g, cfg, seq, steps, warmup, host_chunk, want_host=True):
    """One sequence config on this rank: device-resident pass (value), host-buffer pass (e2e), kernel table."""
    torch, vb = g.torch, g.vb
    n_frames = seq["frames"].shape[0]
    n_pairs = n_frames - 1
    opts = vb.default_gn_opts(grad_mode=1, first_lvl=cfg["first_lvl"])
    tr = g.ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"], max_pairs=n_pairs,
                       norm=cfg["norm"], desc_bytes=cfg["desc_bytes"], gn_opts=opts)
    desc = np.ascontiguousarray(seq["desc"]).view(np.uint8).reshape(n_frames, cfg["n_feat"], cfg["desc_bytes"])
    h = {"frames": g.pin(seq["frames"]), "desc": g.pin(desc), "kp": g.pin(seq["kp"]), "prior": g.pin(seq["prior"])}
    d = {k: v.to(g.dev) for k, v in h.items()}
    d_pose = torch.zeros((n_pairs, 7), dtype=torch.float32, device=g.dev)
    d_ng = torch.zeros((n_pairs,), dtype=torch.int32, device=g.dev)

    def step_device():
        tr.track_sequence(d["frames"], d["desc"], d["kp"], d["prior"], pose=d_pose, n_good=d_ng, stream=g.stream)

    for _ in range(warmup):
        step_device()
    g.barrier()
    tr.stats()
    g.ctx.profile(True)
    l0 = g.ctx.launches
    ms = timed(g, step_device, steps, 0)
    launches = g.ctx.launches - l0
    prof = g.ctx.profile_read()
    g.ctx.profile(False)
    stats = tr.stats()
    res = {"tracker": tr, "d": d, "h": h, "d_pose": d_pose, "d_ng": d_ng, "ms": ms, "launches": launches, "prof": prof,
           "stats": stats, "n_pairs": n_pairs, "step_device": step_device}
    if want_host:
        tr_host = g.ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"],
                                max_pairs=min(host_chunk, n_pairs), norm=cfg["norm"], desc_bytes=cfg["desc_bytes"], gn_opts=opts)
        h_pose = torch.zeros((n_pairs, 7), dtype=torch.float32).pin_memory()
        h_ng = torch.zeros((n_pairs,), dtype=torch.int32).pin_memory()
        res.update({"tr_host": tr_host, "h_pose": h_pose, "h_ng": h_ng,
                    "step_host": lambda: tr_host.track_sequence_host(h["frames"], h["desc"], h["kp"], h["prior"], h_pose, h_ng)})
    return res


def summarize_leg(g, cfg, leg, steps, e2e_ms, unit_name, cpu):
    knn_impl = int(os.environ.get("VSB_KNN_IMPL", "6"))
    n_pairs = leg["n_pairs"]
    kernels = kernel_table(g, cfg, leg["prof"], steps, leg["stats"], n_pairs, n_pairs + 1, knn_impl)
    dom = next((k for k in kernels if "bound" in k), None)
    out = {"workload": cfg["name"] + ": " + cfg["what"], "value": n_pairs / (leg["ms"] * 1e-3), "unit": unit_name,
           "ms_per_step": leg["ms"], "pairs_per_step": n_pairs, "gpu_launches_per_step": leg["launches"] / steps,
           "gn_iterations_per_pair": leg["stats"]["iterations"] / max(1, leg["stats"]["pairs"]),
           "gn_points_per_pair": leg["stats"]["point_visits"] / max(1, leg["stats"]["pairs"]),
           "good_matches_per_pair": float(leg["d_ng"].float().mean()),
           "kernels": kernels,
           "roofline": None if dom is None else {k: dom[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "peak_source")
                                                 if k in dom},
           "cpu_baseline": cpu}
    if dom is not None:
        out["roofline"]["traffic"] = dom.get("traffic")
    if e2e_ms is not None:
        tt = leg["tr_host"].host_traffic()
        out["e2e"] = {"value": n_pairs / (e2e_ms * 1e-3), "unit": unit_name, "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": int(tt["h2d"]), "d2h_bytes_per_step": int(tt["d2h"]),
                      "h2d_gbs_achieved": tt["h2d"] / (e2e_ms * 1e-3) / 1e9,
                      "matches_device_path": bool(g.torch.equal(leg["h_pose"].to(g.dev), leg["d_pose"]))}
    return out


def cpu_sample_for(seq, cfg, leg, n_sample, threads=1):
    """CPU port on the first n_sample pairs of a sequence leg + parity of the GPU poses on them."""
    ids = list(range(min(n_sample, leg["n_pairs"])))
    fps, dt, poses = cpu_track(seq, cfg, ids, threads)
    diff = float(np.abs(poses - leg["d_pose"][: len(ids)].cpu().numpy()).max())
    return {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"first {len(ids)} frame pairs of the same sequence, oracle port, {threads} thread(s) ({dt:.1f}s)",
            "max_abs_pose_diff_vs_gpu": diff}


def leg_single_pair(g, steps):
    """configs[0]: one frame pair — latency, not throughput.  Device-resident (vsb_track_pairs) and host-buffer
    (vsb_track_pairs_host) microseconds per pair; the oracle port on the same pair beside it."""
    torch, vb = g.torch, g.vb
    cfg = wl.CFG0
    from vislam_b200 import synth
    p = synth.make_pair(w=cfg["w"], h=cfg["h"], n_feat=cfg["n_feat"], K=cfg["K"], seed=cfg["seed"], device=g.dev)
    tr = g.ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"], max_pairs=1)
    keys = ("prev", "cur", "d1", "d2", "kp1", "pose_prior")
    h = {k: g.pin(p[k][None]) for k in keys}
    d = {k: v.to(g.dev) for k, v in h.items()}
    d_pose = torch.zeros((1, 7), dtype=torch.float32, device=g.dev)
    d_ng = torch.zeros((1,), dtype=torch.int32, device=g.dev)
    h_pose = torch.zeros((1, 7), dtype=torch.float32).pin_memory()
    fn = lambda: tr.track_pairs(d["prev"], d["cur"], d["d1"], d["d2"], d["kp1"], d["pose_prior"], pose=d_pose, n_good=d_ng, stream=g.stream)
    reps = max(20, steps)
    for _ in range(5):
        fn()
    torch.cuda.synchronize(g.dev)
    tr.stats()
    g.ctx.profile(True)
    l0 = g.ctx.launches
    ms = timed(g, fn, reps, 0)
    launches = g.ctx.launches - l0
    prof = g.ctx.profile_read()
    g.ctx.profile(False)
    stats = tr.stats()
    fh = lambda: tr.track_pairs_host(h["prev"], h["cur"], h["d1"], h["d2"], h["kp1"], h["pose_prior"], h_pose)
    e2e_ms = timed_wall(g, fh, reps, 5)
    tt = tr.host_traffic()
    # the oracle port on the same pair, best of 3
    pr = dict(prev=p["prev"], cur=p["cur"], d1=p["d1"], d2=p["d2"], kp1=p["kp1"], prior=p["pose_prior"])
    best = None
    for _ in range(3):
        fps, dt, poses = cpu_track_pairs([pr], cfg, 1)
        best = dt if best is None else min(best, dt)
    gn = prof.get("gn_solve", (0.0, 1))
    gn_us = gn[0] / reps * 1e3
    visit_b = 14.0 * stats["point_visits"] / max(1, stats["pairs"])
    out = {"workload": cfg["name"] + ": " + cfg["what"], "value": ms * 1e3, "unit": "us per frame pair (device-resident)",
           "higher_is_better": False, "gpu_launches_per_step": launches / reps,
           "e2e": {"value": e2e_ms * 1e3, "unit": "us per frame pair (host buffers in, pose out)",
                   "h2d_bytes_per_step": int(tt["h2d"]), "d2h_bytes_per_step": int(tt["d2h"]),
                   "matches_device_path": bool(torch.equal(h_pose.to(g.dev), d_pose))},
           "kernels_us": {k: v[0] / reps * 1e3 for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
           "gn_iterations": stats["iterations"] / max(1, stats["pairs"]), "gn_points": stats["point_visits"] / max(1, stats["pairs"]),
           "roofline": {"kernel": "gn_solve", "bound": "latency", "achieved": visit_b / (gn_us * 1e-6) / 1e9, "peak": g.hbm_peak,
                        "unit": "GB/s", "frac": visit_b / (gn_us * 1e-6) / 1e9 / g.hbm_peak, "peak_source": g.hbm_src, "traffic": None,
                        "note": "a single pair is one cluster of 8 thread blocks (4096 threads, partial sums exchanged through distributed shared memory) running its serial Gauss-Newton iterations: "
                                "latency-bound by construction (SURVEY 8d: report achieved bandwidth and wall time, claim a "
                                "roofline fraction only for the batched configs)"},
           "cpu_baseline": {"value": best * 1e6, "unit": "us per frame pair", "cores": 1, "kind": "port",
                            "sample": "the same pair, oracle port, best of 3",
                            "max_abs_pose_diff_vs_gpu": float(np.abs(poses[0] - d_pose[0].cpu().numpy()).max())}}
    # the same pair from IMAGES alone (informational): cv::ORB::create(1000) on the device for both frames (pyramid levels on
    # separate streams for a batch this small), then match + GN
    try:
        frames2 = torch.stack([d["prev"][0], d["cur"][0]]).contiguous()
        tr.track_sequence_orb(frames2, d["pose_prior"], nfeatures=cfg["n_feat"], stream=g.stream)
        raw_ms = timed(g, lambda: tr.track_sequence_orb(frames2, d["pose_prior"], nfeatures=cfg["n_feat"], stream=g.stream), reps, 3)
        out["from_raw_frames"] = {"value": raw_ms * 1e3, "unit": "us per frame pair (two frames through ORB on the device, then match + GN)"}
    except Exception as e:      # informational
        out["from_raw_frames"] = {"error": repr(e)}
    tr.close()
    return out


def leg_batched_pairs(g, steps, warmup):
    """configs[4]: 8192 independent pairs x 5000 features, sharded over the ranks in contiguous blocks (no data-path
    collective); every rank generates only its shard, on its GPU.  Aggregate pairs/s = all pairs / max-over-ranks time."""
    torch, vb = g.torch, g.vb
    from vislam_b200 import replicas
    cfg = wl.CFG4
    total = int(os.environ.get("VSB_BENCH_CFG4_PAIRS", str(cfg["pairs"])))
    lo, hi = replicas.shard_range(total, g.rank, g.world)
    n = hi - lo
    chunk = min(n, int(os.environ.get("VSB_BENCH_CFG4_CHUNK", "1024")))
    data = wl.pairs_on_device(cfg, lo, hi, g.dev, log=log if g.rank == 0 else None)
    tr = g.ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"], max_pairs=chunk)
    d_pose = torch.zeros((n, 7), dtype=torch.float32, device=g.dev)
    d_ng = torch.zeros((n,), dtype=torch.int32, device=g.dev)

    def step():
        for p0 in range(0, n, chunk):
            p1 = min(n, p0 + chunk)
            tr.track_pairs(data["prev"][p0:p1], data["cur"][p0:p1], data["d1"][p0:p1], data["d2"][p0:p1], data["kp1"][p0:p1],
                           data["prior"][p0:p1], pose=d_pose[p0:p1], n_good=d_ng[p0:p1], stream=g.stream)

    reps = max(2, min(steps, 3))
    for _ in range(max(1, min(warmup, 2))):
        step()
    g.barrier()
    tr.stats()
    g.ctx.profile(True)
    l0 = g.ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.barrier()
    e0.record(g.stream)
    for _ in range(reps):
        step()
    e1.record(g.stream)
    g.barrier()
    ms = e0.elapsed_time(e1) / reps
    launches = g.ctx.launches - l0
    prof = g.ctx.profile_read()
    g.ctx.profile(False)
    stats = tr.stats()
    ms_ranks = g.gather_ranks(ms)
    ms_max = max(ms_ranks)
    # end to end: the shard from pinned host buffers through vsb_track_pairs_host (a bounded part of the shard: pinning
    # gigabytes takes longer than the pass itself)
    n_host = max(1, min(n, int(os.environ.get("VSB_BENCH_CFG4_HOST_PAIRS", "2048")) // g.world))
    hchunk = min(n_host, 256)
    tr_host = g.ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"], max_pairs=hchunk)
    h = {k: g.pin(data[k][:n_host]) for k in ("prev", "cur", "d1", "d2", "kp1", "prior")}
    h_pose = torch.zeros((n_host, 7), dtype=torch.float32).pin_memory()
    fh = lambda: tr_host.track_pairs_host(h["prev"], h["cur"], h["d1"], h["d2"], h["kp1"], h["prior"], h_pose)
    fh()
    g.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fh()
    g.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
    e2e_ranks = g.gather_ranks(e2e_ms)
    tt = tr_host.host_traffic()
    same = bool(torch.equal(h_pose.to(g.dev), d_pose[:n_host]))
    out = None
    if g.rank == 0:
        leg = {"prof": prof, "stats": stats, "n_pairs": n, "ms": ms}
        knn_impl = int(os.environ.get("VSB_KNN_IMPL", "6"))
        kernels = kernel_table(g, cfg, prof, reps, stats, n, 2 * n, knn_impl)
        dom = next((k for k in kernels if "bound" in k), None)
        # CPU port on a bounded sample of this rank's shard, all host threads, + parity
        cores = os.cpu_count() or 1
        ns = min(n, int(os.environ.get("VSB_BENCH_CFG4_CPU_PAIRS", "16")))
        cp = {k: data[k][:ns].cpu().numpy() for k in ("prev", "cur", "d1", "d2", "kp1", "prior")}
        plist = [dict(prev=cp["prev"][i], cur=cp["cur"][i], d1=cp["d1"][i], d2=cp["d2"][i], kp1=cp["kp1"][i], prior=cp["prior"][i])
                 for i in range(ns)]
        fps, dt, poses = cpu_track_pairs(plist, cfg, min(cores, ns))
        out = {"workload": cfg["name"] + ": " + cfg["what"], "value": total / (ms_max * 1e-3), "unit": "frame pairs/s (aggregate)",
               "scaling": "strong", "pairs_total": total, "pairs_per_rank": n, "chunk_pairs": chunk, "ms_per_step": ms_max,
               "per_rank_ms_per_step": ms_ranks, "us_per_pair_per_gpu": ms * 1e3 / n,
               "gpu_launches_per_step": launches / reps,
               "gn_iterations_per_pair": stats["iterations"] / max(1, stats["pairs"]),
               "gn_points_per_pair": stats["point_visits"] / max(1, stats["pairs"]),
               "good_matches_per_pair": float(d_ng.float().mean()),
               "l2_policy": "every pair distinct: %.1f GB of frames and descriptors per rank per step, far beyond L2" %
                            (n * (2.0 * cfg["w"] * cfg["h"] + 2.0 * cfg["n_feat"] * 32) / 1e9),
               "e2e": {"value": g.world * n_host / (max(e2e_ranks) * 1e-3), "unit": "frame pairs/s (aggregate)",
                       "pairs_per_rank": n_host, "ms_per_step": max(e2e_ranks), "per_rank_ms_per_step": e2e_ranks,
                       "h2d_bytes_per_step": int(tt["h2d"]), "d2h_bytes_per_step": int(tt["d2h"]),
                       "h2d_gbs_achieved_per_gpu": tt["h2d"] / (e2e_ms * 1e-3) / 1e9, "matches_device_path": same,
                       "what": "vsb_track_pairs_host on the first pairs of every rank's shard (pinned host buffers, chunks of "
                               "%d pairs on two streams)" % hchunk},
               "kernels": kernels,
               "roofline": None if dom is None else {k: dom.get(k) for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "peak_source", "traffic")},
               "cpu_baseline": {"value": fps, "unit": "frame pairs/s", "cores": min(cores, ns), "kind": "port",
                                "sample": f"first {ns} pairs of rank 0's shard, oracle port, {min(cores, ns)} pairs in flight ({dt:.1f}s)",
                                "max_abs_pose_diff_vs_gpu": float(np.abs(poses - d_pose[:ns].cpu().numpy()).max())}}
    tr.close()
    tr_host.close()
    del data
    torch.cuda.empty_cache()
    return out


def run_gpu(args, rank, world, local_rank):
    g = Gpu(rank, world, local_rank)
    torch, vb, ctx, dev = g.torch, g.vb, g.ctx, g.dev
    from vislam_b200 import replicas
    cfg = wl.CFG1
    n_frames = int(os.environ.get("VSB_BENCH_FRAMES", str(N_FRAMES)))
    host_chunk = int(os.environ.get("VSB_BENCH_HOST_CHUNK", "250"))          # host-buffer pass: H2D/compute pipeline depth
    same_seed = os.environ.get("VSB_BENCH_SAME_SEED", "0") == "1"            # control: every replica tracks the same sequence
    seed = cfg["seed"] if same_seed else replicas.replica_seed(cfg["seed"], rank)
    # the pinned host buffers live on the GPU's own NUMA node (replicas.gpu_local_cpus)
    numa = replicas.gpu_local_cpus(local_rank) if os.environ.get("VSB_BENCH_NUMA_BIND", "1") != "0" else None
    seq = wl.sequence(cfg, product_initial_pose(vb), n_frames=n_frames, seed=seed, device=dev, log=log if rank == 0 else None)
    if numa is not None:
        numa.__enter__()
    leg = leg_sequence(g, cfg, seq, args.steps, args.warmup, host_chunk)
    if numa is not None:
        numa.__exit__(None, None, None)
    n_pairs = leg["n_pairs"]
    # ---- device-resident timing ("value"): W warm-up steps were done, now exactly K steps between barriers ----------
    g.barrier()
    leg["tracker"].stats()
    ctx.profile(True)
    launches0 = ctx.launches
    sampler = NvmlSampler(local_rank)
    if not sampler.start():
        sampler = ClockSampler(local_rank)
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.barrier()
    ev0.record(g.stream)
    for _ in range(args.steps):
        leg["step_device"]()
    ev1.record(g.stream)
    g.barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    leg["prof"] = ctx.profile_read()
    ctx.profile(False)
    leg["stats"] = leg["tracker"].stats()
    leg["ms"] = ms
    ms_ranks = g.gather_ranks(ms)
    ms_per_step = max(ms_ranks)
    value = replicas.aggregate_throughput(n_pairs, ms_per_step, world)
    clk_ranks = g.gather_ranks(clocks.get("sm_mhz") or 0.0)

    # ---- end-to-end timing through the host-buffer entry ("e2e") -------------------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        leg["step_host"]()
    g.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        leg["step_host"]()      # synchronous: returns when the poses are in host memory
    g.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ranks = g.gather_ranks(e2e_ms)
    e2e_ms_step = max(e2e_ranks)
    e2e_value = replicas.aggregate_throughput(n_pairs, e2e_ms_step, world)
    # what the link gives a plain pinned -> device copy of the same frame buffer, all ranks copying at once (the e2e pass
    # is bound by it; tools/h2d_floor.py measures the same floor without any of this program)
    g.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    leg["d"]["frames"].copy_(leg["h"]["frames"], non_blocking=True)
    c0.record(g.stream)
    for _ in range(3):
        leg["d"]["frames"].copy_(leg["h"]["frames"], non_blocking=True)
    c1.record(g.stream)
    torch.cuda.synchronize(dev)
    h2d_copy_gbs = 3 * leg["h"]["frames"].numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9
    copy_ranks = g.gather_ranks(h2d_copy_gbs)
    traffic_host = leg["tr_host"].host_traffic()
    h2d, d2h, n_host_chunks = int(traffic_host["h2d"]), int(traffic_host["d2h"]), int(traffic_host["chunks"])
    same = bool(torch.equal(leg["h_pose"].to(dev), leg["d_pose"]))

    # ---- the same sequence from images alone (informational): device ORB feeds the matcher, SURVEY 8f N-4 ---------
    raw = None
    d = leg["d"]
    if rank == 0 and os.environ.get("VSB_BENCH_RAW_FRAMES", "1") != "0":
        try:
            tr = leg["tracker"]
            tr.track_sequence_orb(d["frames"], d["prior"], nfeatures=cfg["n_feat"], stream=g.stream)      # warm-up (allocations)
            torch.cuda.synchronize(dev)
            ctx.profile(True)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(g.stream)
            for _ in range(2):
                _, _, r_nf = tr.track_sequence_orb(d["frames"], d["prior"], nfeatures=cfg["n_feat"], stream=g.stream)
            r1.record(g.stream)
            torch.cuda.synchronize(dev)
            r_ms = r0.elapsed_time(r1) / 2
            r_prof = {k: round(v[0] / 2, 3) for k, v in ctx.profile_read().items()}
            ctx.profile(False)
            raw = {"value": n_pairs / (r_ms * 1e-3), "unit": "frames/s", "ms_per_step": r_ms,
                   "orb_keypoints_per_frame": float(r_nf.float().mean()),
                   "what": "vsb_track_sequence_orb: cv::ORB::create(%d) (8 levels, factor 1.2) on the device for every frame, then the "
                           "same match + GN path; frames resident in HBM; the rendered frames carry far fewer corners than the "
                           "%d random descriptors of configs[1]" % (cfg["n_feat"], cfg["n_feat"]),
                   "kernels_ms": r_prof}
            # ... and end to end: frames in pinned host memory, ORB + match + GN on the device, poses back on the host
            # (vsb_track_sequence_orb_host: the copy of chunk i + 1 overlaps the kernels of chunk i)
            r_pose_d = tr.track_sequence_orb(d["frames"], d["prior"], nfeatures=cfg["n_feat"], stream=g.stream)[0]
            torch.cuda.synchronize(dev)
            h = leg["h"]
            trh = ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"],
                              max_pairs=min(n_pairs, int(os.environ.get("VSB_BENCH_RAW_CHUNK", "1000"))))
            r_pose = torch.zeros((n_pairs, 7), dtype=torch.float32).pin_memory()
            trh.track_sequence_orb_host(h["frames"], h["prior"], r_pose, nfeatures=cfg["n_feat"])          # warm-up (allocations)
            t0 = time.perf_counter()
            for _ in range(2):
                trh.track_sequence_orb_host(h["frames"], h["prior"], r_pose, nfeatures=cfg["n_feat"])
            rh_ms = (time.perf_counter() - t0) * 1e3 / 2
            rt = trh.host_traffic()
            raw["e2e"] = {"value": n_pairs / (rh_ms * 1e-3), "unit": "frames/s", "ms_per_step": rh_ms,
                          "h2d_bytes_per_step": int(rt["h2d"]), "d2h_bytes_per_step": int(rt["d2h"]), "host_chunks": int(rt["chunks"]),
                          "matches_device_path": bool(torch.equal(r_pose.to(dev), r_pose_d)),
                          "what": "vsb_track_sequence_orb_host: frames and priors from pinned host memory, nothing else"}
            trh.close()
            leg["step_device"]()          # restore d_pose for the parity check below
            torch.cuda.synchronize(dev)
        except Exception as e:      # informational leg: never fails the bench
            raw = {"error": str(e)}
    # ---- the matcher alone, int8 kernel against the 4-bit persistent kernel (informational) ----------
    knn_variants = None
    if rank == 0 and os.environ.get("VSB_BENCH_KNN_VARIANTS", "1") != "0":
        try:
            knn_variants = {}
            q, t_ = d["desc"][:-1].contiguous(), d["desc"][1:].contiguous()
            for name, impl in (("int8_packed", 2), ("mxf4_persistent", 5)):
                ctx.option("knn_impl", impl)
                knn_variants[name] = timed(g, lambda: ctx.knn2_hamming(q, t_), 5, 2)
            knn_variants["unit"] = "ms per %d pairs of %d x %d descriptors, vsb_knn2_hamming incl. unpacking" % (n_pairs, cfg["n_feat"], cfg["n_feat"])
        except Exception as e:
            knn_variants = {"error": str(e)}
        finally:
            ctx.option("knn_impl", int(os.environ.get("VSB_KNN_IMPL", "6")))

    # ---- the other BASELINE configs ----------------------------------------------------------------------------
    configs = {}
    want = os.environ.get("VSB_BENCH_CONFIGS", "0,2,3,4").split(",")
    if world == 1 and rank == 0:
        if "0" in want:
            try:
                configs["configs[0]"] = leg_single_pair(g, args.steps)
            except Exception as e:
                configs["configs[0]"] = {"error": repr(e)}
        for key, c, nsample in (("2", wl.CFG2, 32), ("3", wl.CFG3, 6)):
            if key not in want:
                continue
            try:
                sq = wl.sequence(c, product_initial_pose(vb), device=dev, log=log)
                lg = leg_sequence(g, c, sq, 3, 2, 128)
                e2e = timed_wall(g, lg["step_host"], 3, 1)
                cpu = cpu_sample_for(sq, c, lg, nsample, 1)
                configs["configs[%s]" % key] = summarize_leg(g, c, lg, 3, e2e, "frames/s", cpu)
                lg["tracker"].close(); lg["tr_host"].close()
                del lg, sq
                torch.cuda.empty_cache()
            except Exception as e:
                configs["configs[%s]" % key] = {"error": repr(e)}
    if "4" in want:
        try:
            c4 = leg_batched_pairs(g, args.steps, args.warmup)
        except Exception as e:
            c4 = {"error": repr(e)}
        if rank == 0:
            configs["configs[4]"] = c4
    if rank != 0:
        leg["tracker"].close(); leg["tr_host"].close(); ctx.close()
        return

    # ---- roofline of every kernel, the dominant one reported in "roofline" ---------------------------
    knn_impl = int(os.environ.get("VSB_KNN_IMPL", "6"))
    kernels = kernel_table(g, cfg, leg["prof"], args.steps, leg["stats"], n_pairs, n_frames, knn_impl)
    dom = next((k for k in kernels if "bound" in k), None)
    roofline = None
    if dom:
        roofline = {k: dom.get(k) for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "peak_source", "split")}
        roofline["share_of_step"] = dom["share"]

    # ---- CPU baseline and parity: the oracle port on this box's host cores ---------------------------------------------
    # (1) every pair of the workload on all host threads: the parity check of ALL GPU poses; (2) a bounded single-thread
    # sample: the reference is single-threaded, this is the figure `cpu_baseline.value` reports; (3) the reference's own
    # translation units on a few pairs.
    cores = os.cpu_count() or 1
    fps_all, dt_all, poses_all = cpu_track(seq, cfg, list(range(n_pairs)), cores)
    gpu_poses = leg["d_pose"].cpu().numpy()
    parity = float(np.abs(poses_all - gpu_poses).max())
    n_bad = int((np.abs(poses_all - gpu_poses).max(axis=1) > 0).sum())
    n_sample = min(n_pairs, int(os.environ.get("VSB_CPU_SAMPLE_PAIRS", "256")))
    fps1, dt1, _ = cpu_track(seq, cfg, list(range(n_sample)), 1)
    ru = ref_units_timing(seq, cfg, int(os.environ.get("VSB_REF_UNITS_PAIRS", "8")))
    if ru and "poses" in ru:
        rp = ru.pop("poses")
        ru["max_abs_pose_diff_vs_gpu"] = float(np.abs(rp - gpu_poses[: len(rp)]).max())
    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": base_config(world),
        "details": {"frames": n_frames, "pairs_per_step": n_pairs, "host_chunk_pairs": host_chunk, "host_chunks": n_host_chunks,
                    "grad_mode": 1, "same_seed_on_all_ranks": same_seed,
                    "gn_iterations_per_pair": leg["stats"]["iterations"] / max(1, leg["stats"]["pairs"]),
                    "gn_points_per_pair": leg["stats"]["point_visits"] / max(1, leg["stats"]["pairs"]),
                    "per_rank": {"ms_per_step": ms_ranks, "ms_min": min(ms_ranks), "ms_mean": float(np.mean(ms_ranks)),
                                 "ms_max": max(ms_ranks), "sm_mhz": clk_ranks, "e2e_ms_per_step": e2e_ranks,
                                 "h2d_gbs_plain_copy": copy_ranks}},
        "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_ms_step, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "matches_device_path": same,
                "h2d_gbs_achieved": h2d / (e2e_ms_step * 1e-3) / 1e9, "h2d_gbs_plain_copy": min(copy_ranks),
                "fraction_of_plain_copy": (h2d / (e2e_ms_step * 1e-3) / 1e9) / min(copy_ranks),
                "host_buffers_numa_bound": bool(numa is not None and numa.bound)},
        "from_raw_frames": raw,
        "knn_variants_ms": knn_variants,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "cpu_baseline": {"value": fps1, "unit": "frames/s", "cores": 1, "kind": "port",
                         "sample": f"first {n_sample} frame pairs of the same sequence, oracle port, single thread ({dt1:.1f}s); "
                                   f"box has {cores} host cores",
                         "all_threads": {"value": fps_all, "unit": "frames/s", "cores": cores,
                                         "sample": f"all {n_pairs} frame pairs, {cores} pairs in flight ({dt_all:.1f}s)"},
                         "max_abs_pose_diff_vs_gpu": parity, "pairs_checked": n_pairs, "pairs_with_any_different_bit": n_bad,
                         "reference_units": ru,
                         "cv2_bfmatcher": cv2_matcher_timing(seq, cores)},
        "configs": configs,
    }
    print(json.dumps(out), flush=True)
    leg["tracker"].close()
    leg["tr_host"].close()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    try:
        run_gpu(args, rank, world, local_rank)
    finally:
        if world > 1:       # leave the process group cleanly (rank 0 gets here last: it alone runs the CPU baseline)
            try:
                import torch.distributed as dist
                if dist.is_initialized():
                    dist.destroy_process_group()
            except Exception:
                pass


if __name__ == "__main__":
    main()

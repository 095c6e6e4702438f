#!/usr/bin/env python
"""The box's host-to-device floor, measured without any of this repository's code: every rank pins a buffer the size of
the frames of BASELINE configs[1] (2000 x 752 x 480 bytes) and copies it to its GPU, all ranks at once, CUDA events around
the copies, barrier on both sides.  Prints one JSON line (rank 0): per-rank GB/s for a plain contiguous copy and for the
strided form the tracker's host entry uses (one row per frame into a pitched destination), plus write-combined pinned
memory.  bench.py's e2e leg is reported as a fraction of THIS number at the same N.

  python tools/h2d_floor.py                                   (one GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_floor.py
"""
import ctypes
import json
import os

import torch


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_frames, fbytes, pitch = 2000, 752 * 480, 481 * 1024          # pitch ~ one packed pyramid per frame
    host = torch.empty(n_frames * fbytes, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    dst = torch.empty(n_frames * fbytes, dtype=torch.uint8, device=dev)
    dst2 = torch.empty(n_frames * pitch, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                     ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, reps=5):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st)
        barrier()
        return reps * n_frames * fbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    plain = timed(lambda: dst.copy_(host, non_blocking=True))
    strided = timed(lambda: rt.cudaMemcpy2DAsync(dst2.data_ptr(), pitch, host.data_ptr(), fbytes, fbytes, n_frames, 1,
                                                 ctypes.c_void_p(st.cuda_stream)))
    wc_ptr = ctypes.c_void_p()
    wc = None
    if rt.cudaHostAlloc(ctypes.byref(wc_ptr), n_frames * fbytes, 4) == 0:      # cudaHostAllocWriteCombined
        ctypes.memset(wc_ptr, 7, n_frames * fbytes)
        wc = timed(lambda: rt.cudaMemcpyAsync(dst.data_ptr(), wc_ptr, n_frames * fbytes, 1, ctypes.c_void_p(st.cuda_stream)))

    def gather(v):
        if dist is None:
            return [v]
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    res = {"plain_gbs": gather(plain), "strided_2d_gbs": gather(strided), "write_combined_gbs": gather(wc or 0.0)}
    if rank == 0:
        print(json.dumps({"tool": "h2d_floor", "n_gpus": world, "bytes_per_copy": n_frames * fbytes,
                          "per_rank": res, "min_plain_gbs": min(res["plain_gbs"]), "sum_plain_gbs": sum(res["plain_gbs"])}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

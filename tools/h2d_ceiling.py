#!/usr/bin/env python
"""What the BOX delivers host -> device, with nothing of ours in the way: bare pinned-memory cudaMemcpyAsync on N
GPUs at once (one process per GPU, like bench.py), next to the PCIe / NUMA topology.

    python tools/h2d_ceiling.py --gpus 8 [--mb 21] [--reps 200]

bench.py's `e2e` copies 21 MB of PCM16 per step per GPU (pipeline.HotPathStep.run_host); at 8 GPUs it reaches
~185 GB/s aggregate.  This tool measures the ceiling that number has to be compared with:

  pinned        torch pin_memory() after binding the process to the GPU's NUMA-local cores (what bench.py does)
  pinned_nobind the same without the binding
  wc            cudaHostAlloc(cudaHostAllocWriteCombined): no CPU cache snooping on the DMA reads
  chunked4      the 21 MB batch as four back-to-back copies (DMA pipelining)
  two_streams   two halves of the batch on two copy streams (both copy engines)

Rank 0 prints one JSON object: per-mode GB/s per rank (min / median / max over ranks) and the aggregate, plus
`nvidia-smi topo -m` and `lspci -tv` when available.  Everything is timed with CUDA events on the copy stream,
after a barrier, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _cmd(args):
    try:
        return subprocess.run(args, capture_output=True, text=True, timeout=20).stdout
    except Exception as e:                                       # noqa: BLE001
        return "unavailable: %s" % e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mb", type=float, default=20.97152, help="bytes per copy, MB (default: 32 x 327680 int16)")
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533")] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    import torch
    import torch.distributed as dist
    from reconvat_b200 import parallel
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    n = int(args.mb * 1e6) // 2 * 2
    n_buf = max(2, int(160e6 // n) + 1)                         # rotate over > L2 / > LLC worth of host memory
    dst = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(srcs, mode):
        def one(i):
            s, d = srcs[i % len(srcs)], dst[i & 1]
            if mode == "chunked4":
                q = n // 4
                with torch.cuda.stream(streams[0]):
                    for c in range(4):
                        d[c * q:(c + 1) * q].copy_(s[c * q:(c + 1) * q], non_blocking=True)
            elif mode == "two_streams":
                h = n // 2
                for k in range(2):
                    with torch.cuda.stream(streams[k]):
                        d[k * h:(k + 1) * h].copy_(s[k * h:(k + 1) * h], non_blocking=True)
            else:
                with torch.cuda.stream(streams[0]):
                    d.copy_(s, non_blocking=True)
        for i in range(5):
            one(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for i in range(args.reps):
            one(i)
        streams[0].wait_stream(streams[1])
        e1.record(streams[0])
        barrier()
        ms = e0.elapsed_time(e1)
        gbs = torch.tensor([n * args.reps / (ms * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            allg = [torch.zeros_like(gbs) for _ in range(world)]
            dist.all_gather(allg, gbs)
            per = sorted(float(t) for t in allg)
        else:
            per = [float(gbs)]
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"per_rank_gbs_min": per[0], "per_rank_gbs_median": per[len(per) // 2], "per_rank_gbs_max": per[-1],
                "aggregate_gbs": world * n * args.reps / (float(t) * 1e-3) / 1e9}

    out = {"n_gpus": world, "bytes_per_copy": n, "reps": args.reps, "modes": {}}
    nobind = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(n_buf)]
    for b in nobind:
        b.fill_(1)
    out["modes"]["pinned_nobind"] = timed(nobind, "plain")
    del nobind
    cores = parallel.bind_to_gpu_numa(local)
    out["numa_local_cores"] = cores
    pinned = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(n_buf)]
    for b in pinned:
        b.fill_(1)
    out["modes"]["pinned"] = timed(pinned, "plain")
    out["modes"]["chunked4"] = timed(pinned, "chunked4")
    out["modes"]["two_streams"] = timed(pinned, "two_streams")
    # write-combined pinned memory through the runtime (torch has no switch for it)
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        ptrs, wc = [], []
        for _ in range(n_buf):
            p = ctypes.c_void_p()
            rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(0x04 | 0x01))   # WC | portable
            if rc != 0:
                raise RuntimeError("cudaHostAlloc -> %d" % rc)
            ptrs.append(p)
            ctypes.memset(p, 1, n)
            buf = (ctypes.c_uint8 * n).from_address(p.value)
            wc.append(torch.frombuffer(buf, dtype=torch.uint8))

        def wc_copy(i):
            with torch.cuda.stream(streams[0]):
                rt.cudaMemcpyAsync(ctypes.c_void_p(dst[i & 1].data_ptr()), ptrs[i % n_buf], ctypes.c_size_t(n),
                                   ctypes.c_int(1), ctypes.c_void_p(streams[0].cuda_stream))
        for i in range(5):
            wc_copy(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for i in range(args.reps):
            wc_copy(i)
        e1.record(streams[0])
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        mine = n * args.reps / (float(t) * 1e-3) / 1e9
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["modes"]["wc"] = {"per_rank_gbs_rank0": mine, "aggregate_gbs": world * n * args.reps / (float(t) * 1e-3) / 1e9}
        for p in ptrs:
            rt.cudaFreeHost(p)
    except Exception as e:                                       # noqa: BLE001
        out["modes"]["wc"] = {"unavailable": str(e)}
    if rank == 0:
        out["topo"] = _cmd(["nvidia-smi", "topo", "-m"])
        out["lspci_tree"] = _cmd(["lspci", "-tv"])[:6000]
        out["numa"] = _cmd(["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"])
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

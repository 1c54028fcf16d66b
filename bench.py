#!/usr/bin/env python
"""Benchmark of the ReconVAT hot path (Mel front-end + VAT step) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step = the hot path over one batch of B synthetic 20.48 s segments per GPU: Mel front-end
(pad/split, tcgen05 STFT, banded Mel, log, imagewise normalise) + one UNet_VAT call against a stand-in
transcriber (the real U-Net is the black-box caller, out of scope).  One JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEG_SECONDS = 20.48
SEG_SAMPLES = 327680
T_FRAMES, N_MELS, N_PITCH = 640, 229, 88
STFT_FLOP_PER_SEG = 2 * 640 * 2050 * 2048                 # dense fp32-equivalent contraction
N4, P4 = T_FRAMES * N_MELS * 4, T_FRAMES * N_PITCH * 4
# algorithmic HBM bytes per segment of each HBM-bound kernel (DESIGN.md "Kernels")
HBM_BYTES_PER_SEG = {
    "rvb_pad_split": 1310716 + 2 * 1318912,
    "rvb_fold_split": 1310716 + 4 * 640 * 1024 * 4,       # R audio, W e/o hi/lo tf32 operand planes
    "rvb_fold_split_f16": 1310716 + 4 * 640 * 1024 * 2 + 640 * 4,   # R audio, W e/o hi/lo fp16 planes + row scales
    "rvb_fold_split_f16_pcm16": 1310716 // 2 + 4 * 640 * 1024 * 2 + 640 * 4,
    "rvb_fold_split2_f16": 1310716 + 4 * 640 * 1024 * 2 + 640 * 4,  # the same planes, even-n columns first
    "rvb_fold_split2_f16_pcm16": 1310716 // 2 + 4 * 640 * 1024 * 2 + 640 * 4,
    "rvb_pad_parity_pcm16": 327679 * 2 + 2 * 164864 * 2,           # R PCM16, W the two parity planes (K0x)
    "rvb_randn_like": N4,                                           # W d (the draw, ATen's amortisation)
    "rvb_vat_perturb_draw": 3 * N4,                                 # R x, W d (for the finalisation), W x_adv
    "rvb_mel_project": 1020 * 640 * 4 + N4,
    "rvb_logmel_minmax": N4,
    "rvb_logmel_transpose": 2 * N4,
    "rvb_logmel_normalise": 3 * N4,                        # fused min/max + normalise + transpose: R two Mel planes, W spec
    "rvb_normalise": 2 * N4,
    "rvb_vat_perturb": 3 * N4,
    "rvb_bce_grad": 3 * P4,
    "rvb_div_grad": 3 * P4,                                # what the VAT modules call (kind = BCE here)
    "rvb_div_mean": 2 * P4,
    "rvb_vat_finalize": 6 * N4,
    "rvb_vat_finalize_stats": 5 * N4,                      # as HotPathStep calls it: d_hat itself is not stored
    "rvb_bce_mean": 2 * P4,
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")   # B200_PROFILING.md


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in out.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def make_audio(n_batches, batch, seed, pcm16=False):
    """n_batches x (batch, SEG_SAMPLES): white and music-like segments, PCM int16 as the dataset stores them
    (model/dataset.py:62) or already converted to float32.  A few unique segments are generated and circularly
    shifted to fill the batches (generation is CPU-expensive)."""
    import numpy as np
    from reconvat_b200 import synth
    uniq = [synth.white_int16(SEG_SAMPLES, seed + 1), synth.music_int16(SEG_SAMPLES, seed + 2),
            synth.music_int16(SEG_SAMPLES, seed + 3), synth.white_int16(SEG_SAMPLES, seed + 4)]
    if not pcm16:
        uniq = [synth.to_float(u) for u in uniq]
    out = []
    for n in range(n_batches):
        a = np.empty((batch, SEG_SAMPLES), np.int16 if pcm16 else np.float32)
        for b in range(batch):
            a[b] = np.roll(uniq[(n + b) % 4], 4099 * (n * batch + b))
        out.append(a)
    return out


def to_float_cpu(audio):
    """The dataset's conversion on the host (model/dataset.py:62); part of the CPU arm when the input is PCM16."""
    import torch
    return audio.float().div_(32768.0) if audio.dtype == torch.int16 else audio


def make_model(args, batch, dev):
    """The black-box network the VAT loop calls.  'injected' (default): precomputed posteriors and input
    gradient, zero kernels -- the step then contains the hot path only (SURVEY.md 8d headline definition).
    'standin': a frame-wise linear+sigmoid transcriber run by PyTorch (adds cuBLAS/ATen launches that are the
    caller's, not ours)."""
    from reconvat_b200.standin import InjectedTranscriber, StandInTranscriber
    m = InjectedTranscriber(batch, seed=7) if args.model == "injected" else StandInTranscriber("unet", n_out=N_PITCH, seed=1)
    return m if dev is None else m.to(dev)


def cpu_path():
    """The CPU implementation of the step that is timed as the baseline: the reference's OWN modules when its tree is
    there (/root/reference in the build container, the oracle/_ref snapshot made by build() on the GPU box), else the
    oracle's port of the same ATen op sequence.  Returns (object with .step(model, audio), kind, description)."""
    from oracle import reference_path
    if reference_path.available():
        path = reference_path.ReferenceHotPath("cpu")
        src = os.path.relpath(path.source, ROOT) if path.source.startswith(ROOT) else path.source
        return path, "reference", "the unmodified reference modules (MelSpectrogram, Normalization, UNet_VAT) from " + src
    from oracle.cpu_path import CpuHotPath
    return CpuHotPath(), "port", "oracle/cpu_path.py (the reference's ATen op sequence; no reference tree found)"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path on the host cores, rank 0 only."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(args.batch, args.cpu_batch)
    audio = torch.from_numpy(make_audio(1, sample, 0, pcm16=args.input == "pcm16")[0])
    model = make_model(args, sample, None)
    path, kind, what = cpu_path()
    torch.manual_seed(0)
    for _ in range(max(1, min(args.warmup, 2))):
        path.step(model, to_float_cpu(audio))
    steps = max(1, min(args.steps, args.cpu_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, rn = path.step(model, to_float_cpu(audio))
        loss.item()
    dt = (time.perf_counter() - t0) / steps
    value = sample * SEG_SECONDS / dt
    line = {
        "impl": "reference", "metric": "audio-sec/s", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Mel+VAT step (front-end + UNet_VAT, network = %s), CPU, %d x 20.48 s "
                               "segments per step (bounded sample of the B=%d workload), input %s"
                               % (args.model, sample, args.batch, args.input)},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind,
                         "sample": "%d segments x %d steps, %s" % (sample, steps, what)},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries ONE JSON line: with NCCL_DEBUG=VERSION (set on the GPU boxes) NCCL printf()s its
        # "NCCL version ..." banner to stdout; at WARN it does not, and whatever it logs goes to a file
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/rvb_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    import reconvat_b200 as R
    from reconvat_b200 import parallel
    from reconvat_b200.pipeline import HotPathStep
    # before the pinned buffers exist: first touch places them on the GPU's NUMA node (a remote node costs a third of
    # the host->device rate: 38 instead of 54 GB/s measured on this pool)
    all_cpus = os.sched_getaffinity(0)
    numa_cores = parallel.bind_to_gpu_numa(local_rank)

    B = args.batch
    pcm16 = args.input == "pcm16"
    n_rot = max(4, -(-130 * 2 ** 20 // (B * SEG_SAMPLES * (2 if pcm16 else 4))))   # rotate inputs over > L2 (126 MB)
    host = [torch.from_numpy(a).pin_memory() for a in make_audio(n_rot, B, 100 * rank, pcm16=pcm16)]
    dev_audio = [h.to(dev) for h in host]
    model = make_model(args, B, dev)
    step = HotPathStep(model, dev)
    torch.manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput ("value") ----------------
    for i in range(args.warmup):
        step(dev_audio[i % n_rot])
    step.vat_loss.check()
    if not args.no_graphs:
        step.capture(dev_audio)                              # one CUDA graph of the whole step per input buffer
        for i in range(args.warmup):
            step.replay(i % n_rot)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = R._lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if args.no_graphs:
        for i in range(args.steps):
            step(dev_audio[i % n_rot])
    elif args.streams > 1:
        # consecutive steps are independent batches: replaying their graphs on alternating streams lets the
        # HBM-bound VAT tail of step i overlap the tensor-bound contraction of step i+1
        step.replay_many([i % n_rot for i in range(args.steps)], streams=args.streams)
    else:
        for i in range(args.steps):
            step.replay(i % n_rot)
    ev1.record()
    barrier()
    launches = (R._lib.launch_count() - launches0) if args.no_graphs else step.kernels_per_graph * args.steps
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    step.check()
    ms_step = ms_total / args.steps
    value = world * B * SEG_SECONDS / (ms_step * 1e-3)
    # the same K steps on ONE stream (kernels strictly one after the other): the denominator of share_of_step, so
    # that the share can be compared with the serialised ncu launch list in profiles/
    ms_step_serial = ms_step
    if not args.no_graphs and args.streams > 1:
        barrier()
        ev0.record()
        for i in range(args.steps):
            step.replay(i % n_rot)
        ev1.record()
        barrier()
        ms_step_serial = ev0.elapsed_time(ev1) / args.steps

    # eager launches of the same K steps (host-bound: what the CUDA graphs remove)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for i in range(args.steps):
        step(dev_audio[i % n_rot])
    ev3.record()
    barrier()
    step.vat_loss.check()
    eager_ms_step = ev2.elapsed_time(ev3) / args.steps

    # ---------------- sustained: the same replay loop for >= 2 s, with its own clock samples ----------------
    sustained = None
    if not args.no_graphs and args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / ms_step) + 1)
        sus_sampler = ClockSampler(local_rank)
        barrier()
        sus_sampler.start()
        ev0.record()
        done = 0
        while done < n_sus:                                  # bounded bursts: the launch queue never holds > 2000 graphs
            n = min(2000, n_sus - done)
            step.replay_many([(done + i) % n_rot for i in range(n)], streams=args.streams)
            done += n
        ev1.record()
        barrier()
        sus_ms = max_over_ranks(ev0.elapsed_time(ev1))
        step.check()
        sustained = {"seconds": sus_ms * 1e-3, "steps": n_sus, "ms_per_step": sus_ms / n_sus,
                     "value": world * B * SEG_SECONDS / (sus_ms / n_sus * 1e-3), "unit": "audio-s/s",
                     "clocks": sus_sampler.summary()}

    # ---------------- precision="strict": the once-folded contraction (every bin accumulated on its own) ------------
    strict = None
    if not args.no_graphs and not args.no_gpu_baselines and world == 1:      # context legs: at N = 1 only
        step_s = HotPathStep(model, dev, precision="strict")
        for i in range(3):
            step_s(dev_audio[i % n_rot])
        step_s.capture(dev_audio)
        for i in range(args.warmup):
            step_s.replay(i % n_rot)
        barrier()
        ev0.record()
        step_s.replay_many([i % n_rot for i in range(args.steps)], streams=args.streams)
        ev1.record()
        barrier()
        step_s.check()
        s_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
        strict = {"ms_per_step": s_ms, "value": world * B * SEG_SECONDS / (s_ms * 1e-3), "unit": "audio-s/s",
                  "what": "the same step with MelSpectrogram(precision='strict'): once-folded contraction (twice the "
                          "multiply-adds), correction terms accumulated before the leading one (1.5x the operand "
                          "traffic); log-Mel within 3.2e-5 of float64 on every stress signal "
                          "(profiles/r02_precision.md); the headline `value` is precision='fast'"}
        del step_s
        torch.cuda.empty_cache()

    # ---------------- the zero-edit module surface and the reference's eager path on the same GPU ----------------
    surface = gpu_eager = None
    if not args.no_gpu_baselines and world == 1:
        f32_audio = [a.float().div_(32768.0) if a.dtype == torch.int16 else a for a in dev_audio[:min(n_rot, 4)]]
        k_b = max(3, min(args.steps, 20))

        def time_steps(fn, sync_each=False):
            for i in range(2):
                fn(f32_audio[i % len(f32_audio)])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(k_b):
                fn(f32_audio[i % len(f32_audio)])
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k_b

        # (1) what an UNCHANGED run_on_batch executes after install(): MelSpectrogram.forward -> the caller's torch.log
        # -> Normalization.transform -> transposed view -> UNet_VAT.forward (synchronous NaN assert, as the reference)
        mel, norm = step.spectrogram, R.utils.Normalization("imagewise")
        vat_eager = R.VAT.UNet_VAT(1e-6, 2.0, 1, False)

        def surface_step(audio):
            spec = mel(audio.reshape(-1, audio.shape[-1])[:, :-1])
            spec = torch.log(spec + 1e-5)
            spec = norm.transform(spec)
            lds, _, r_norm = vat_eager(model, spec.transpose(-1, -2).unsqueeze(1))
            return lds, r_norm.abs().mean()
        surface = {"ms_per_step": time_steps(surface_step), "steps": k_b,
                   "what": "reconvat_b200 behind the reference's unchanged call sequence (model/self_attention_VAT.py:"
                           "1100-1106), eager launches, float32 audio, synchronous NaN assert"}
        surface["value"] = B * SEG_SECONDS / (surface["ms_per_step"] * 1e-3)

        # (2) the kernel-for-kernel bar: the reference's own modules, eager PyTorch on this GPU
        from oracle import reference_path
        if reference_path.available():
            ref_path = reference_path.ReferenceHotPath(dev)
            gpu_eager = {"what": "unmodified reference modules (cuDNN conv1d STFT, cuBLAS Mel matmul, ATen VAT ops, "
                                 "autograd) on the same GPU, same B, same injected network, float32 audio", "steps": k_b}
            saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            for tf32 in (True, False):
                torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
                ms = time_steps(lambda a: ref_path.step(model, a))
                gpu_eager["allow_tf32=%s" % tf32] = {"ms_per_step": ms, "value": B * SEG_SECONDS / (ms * 1e-3),
                                                     "unit": "audio-s/s"}
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
            del ref_path
        else:
            gpu_eager = {"unavailable": "no reference tree (oracle/_ref is made by __graft_entry__.build())"}
        del f32_audio
        torch.cuda.empty_cache()

    # per-kernel durations: every entry point of the step re-launched `reps` times BACK TO BACK between one CUDA-event
    # pair on the launching stream.  The calls are recorded from n_rot eager steps, each run inside its own memory
    # pool, so that step i's tensors (inputs, intermediates, outputs) have their own addresses: consecutive launches
    # rotate over n_rot working sets (> L2) and every launch streams from / to HBM.  A spin kernel holds the stream
    # while the host enqueues, so the events bracket kernels, not launch gaps; no per-launch event overhead.
    recorded, pools = [], []
    for i in range(n_rot):
        pool = torch.cuda.MemPool()
        pools.append(pool)                                   # keeps the recorded addresses reserved
        with torch.cuda.use_mem_pool(pool):
            log = []
            R._lib.record_calls(log)
            try:
                step(dev_audio[i])
            finally:
                R._lib.record_calls(None)
        recorded.append(log)
    barrier()
    step.vat_loss.check()
    names = [n for n, _ in recorded[0]]
    reps = max(n_rot * 3, 20)
    kavg, kcalls = {}, {}
    for idx, name in enumerate(names):
        if name in kavg:                                     # an entry point called twice per step: first use only
            kcalls[name] += 1
            continue
        kcalls[name] = 1
        calls = [recorded[i][idx][1] for i in range(n_rot)]
        for c in calls:                                      # warm-up (and first-launch attribute calls)
            R._lib.raw_call(name, c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(reps * 40e-6 * 1.9e9))         # 40 us of spin per launch the host has to enqueue
        e0.record()
        for r in range(reps):
            R._lib.raw_call(name, calls[r % n_rot])
        e1.record()
        torch.cuda.synchronize()
        kavg[name] = e0.elapsed_time(e1) / reps
    # the size-matched ceiling: a plain device copy moving the SAME number of bytes (half read, half written), timed the
    # same way.  MEASURED_PEAKS' 6.5 TB/s is a 2 GiB copy; a 14-56 MB kernel pays its launch ramp and tail on top of the
    # transfer, and this is what a memcpy achieves at that size on this GPU.
    copy_gbs = {}
    for name in kavg:
        per_seg = HBM_BYTES_PER_SEG.get(name)
        if per_seg is None:
            continue
        nbytes = B * per_seg
        half = max(1 << 16, (nbytes // 2 + 255) // 256 * 256)
        n_pairs = max(2, int(140e6 // half) + 1)
        src = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(n_pairs)]
        dst = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(n_pairs)]
        for i in range(n_pairs):
            dst[i].copy_(src[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(reps * 20e-6 * 1.9e9))
        e0.record()
        for r in range(reps):
            dst[r % n_pairs].copy_(src[r % n_pairs])
        e1.record()
        torch.cuda.synchronize()
        copy_gbs[name] = 2 * half / (e0.elapsed_time(e1) / reps * 1e-3) / 1e9
        del src, dst
    barrier()
    del pools

    # ---------------- end to end from pinned host memory ("e2e") ----------------
    results = torch.zeros((args.steps, 2), dtype=torch.float32).pin_memory()
    step.run_host([host[i % n_rot] for i in range(3)], torch.zeros((3, 2), dtype=torch.float32).pin_memory(),
                  use_graphs=not args.no_graphs)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    n_done = step.run_host((host[i % n_rot] for i in range(args.steps)), results, use_graphs=not args.no_graphs)
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.summary()                               # sampled across both timed regions
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1)) / n_done
    e2e_value = world * B * SEG_SECONDS / (e2e_ms * 1e-3)
    assert torch.isfinite(results).all(), "non-finite VAT loss in the e2e run"

    # ---------------- the reference's own data path: corpus resident on the device (model/dataset.py:19-62) --------
    # PianoRollAudioDataset pre-loads every recording to `device` as int16 and cuts random 327 680-sample crops there
    # (:40-55); a step never crosses PCIe.  Same here: B random crops are gathered from a device-resident corpus into
    # the captured input buffer, the graph is replayed, two 4-byte results go back to the host.
    dev_corpus = None
    if not args.no_graphs:
        corpus = torch.cat([a.reshape(-1) for a in dev_audio[:2]])                    # 64 segments' worth of samples
        starts = [torch.randint(0, corpus.numel() - SEG_SAMPLES, (B,), generator=torch.Generator().manual_seed(i)).to(dev)
                  for i in range(8)]
        windows = corpus.unfold(0, SEG_SAMPLES, 1)                                    # every crop, as a view
        res2 = torch.zeros((args.steps, 2), dtype=torch.float32).pin_memory()

        def corpus_step(i):
            slot = i % 2
            buf = step._graphs[slot][1]
            torch.index_select(windows, 0, starts[i % 8], out=buf)                    # the random crops (device gather)
            out = step.replay(slot)
            res2[i, 0].copy_(out[0].detach(), non_blocking=True)
            res2[i, 1].copy_(out[1], non_blocking=True)
        for i in range(3):
            corpus_step(i)
        barrier()
        ev0.record()
        for i in range(args.steps):
            corpus_step(i)
        ev1.record()
        barrier()
        dc_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
        dev_corpus = {"value": world * B * SEG_SECONDS / (dc_ms * 1e-3), "unit": "audio-s/s", "ms_per_step": dc_ms,
                      "what": "as the reference feeds its step (model/dataset.py:19-62): int16 corpus resident on the "
                              "device, B random 327680-sample crops gathered per step, graph replay, (vat_loss, r_norm) "
                              "read back -- no host->device copy of audio"}
        del corpus

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = load_peaks()
    fold2x = "rvb_stft_mel_fused_pcm16" in kavg
    fold2 = fold2x or "rvb_stft_mel_folded2_f16" in kavg
    fused = fold2 or "rvb_stft_mel_folded_f16" in kavg
    f16 = fused or "rvb_stft_gemm_folded_f16" in kavg
    folded = f16 or "rvb_stft_gemm_folded" in kavg
    gemm_name = ("rvb_stft_mel_fused_pcm16" if fold2x else "rvb_stft_mel_folded2_f16" if fold2 else "rvb_stft_mel_folded_f16" if fused else
                 "rvb_stft_gemm_folded_f16" if f16 else "rvb_stft_gemm_folded" if folded else "rvb_stft_gemm")
    gemm_ms = kavg.get(gemm_name)
    roofline = None
    if gemm_ms:
        # algorithmic FLOPs: the dense contraction as the reference computes it, whichever kernel ran;
        # issued: 3 MMAs per product (hi*hi + hi*lo + lo*hi), contraction length halved by each fold
        k_len = 512 if fold2 else 1024 if folded else 2048
        achieved = B * STFT_FLOP_PER_SEG / (gemm_ms * 1e-3) / 1e12
        issued = 3 * B * 2 * 640 * 2048 * k_len / (gemm_ms * 1e-3) / 1e12
        kname = ("stft_gemm_fold2x_pair_kernel + Mel epilogue" if fold2x else
                 "stft_gemm_fold2c_pair_kernel + Mel epilogue" if fold2 else
                 "stft_gemm_fold_pair_kernel%s" % (" + Mel epilogue" if fused else "") if f16 else
                 "stft_gemm_fold_kernel<tf32>") if folded else "stft_gemm_kernel"
        pipe_peak = peaks["bf16"] if f16 else peaks["bf16"] / 2
        roofline = {"kernel": "%s (%s)" % (kname, gemm_name), "bound": "tensor", "achieved": achieved,
                    "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"], "traffic": None,
                    "peak_source": "%s dense bf16 burst (MEASURED_PEAKS.json); `achieved` counts the reference's dense "
                                   "contraction (5.374 GFLOP per segment), the kernel runs %s with 3 MMAs per product "
                                   "over 1/%d of the contraction length (%s): frac = %s x the tensor-pipe utilisation "
                                   "(issued_frac_of_pipe_peak)"
                                   % (peaks["source"], "kind::f16 (the bf16 rate)" if f16 else "kind::tf32 (half the bf16 rate)",
                                      2048 // k_len, "two symmetry folds" if fold2 else "one fold" if folded else "unfolded",
                                      "4/3" if fold2 else "2/3" if f16 else ("1/3" if folded else "1/6")),
                    "issued_tflops": issued, "issued_frac_of_pipe_peak": issued / pipe_peak,
                    "ms_per_launch": gemm_ms, "share_of_step": gemm_ms / ms_step_serial}
        # ncu DRAM bytes per launch of the kernel that RAN (profiles/traffic.json, keyed by kernel name); a kernel
        # without a capture is reported as such, never labelled with another kernel's bytes
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                roofline["traffic"] = json.load(f)[kname]
        except (OSError, KeyError, ValueError):
            roofline["traffic_note"] = "no ncu --set full capture recorded in profiles/traffic.json for %r" % kname
            print("bench.py: WARNING: profiles/traffic.json has no entry for %r -- roofline.traffic is null" % kname,
                  file=sys.stderr)
    hbm = {}
    for n, per_seg in HBM_BYTES_PER_SEG.items():
        if n in kavg:
            gbs = B * per_seg / (kavg[n] * 1e-3) / 1e9
            hbm[n] = {"ms_per_launch": kavg[n], "launches_per_step": kcalls[n], "achieved_gbs": gbs,
                      "frac_of_measured_hbm": gbs / peaks["hbm"], "algorithmic_mb": B * per_seg / 1e6,
                      "size_matched_copy_gbs": copy_gbs.get(n),
                      "frac_of_size_matched_copy": gbs / copy_gbs[n] if copy_gbs.get(n) else None}

    cpu = None
    if not args.no_cpu_baseline and world == 1:                               # rank 0 at N = 1 only
        os.sched_setaffinity(0, all_cpus)                    # the CPU baseline gets every core again
        cores = len(all_cpus) or 1
        torch.set_num_threads(cores)
        cb = min(B, args.cpu_batch)
        audio = host[0][:cb].clone()
        cm = make_model(args, cb, None)
        path, cpu_kind, cpu_what = cpu_path()
        path.step(cm, to_float_cpu(audio))
        t0 = time.perf_counter()
        for _ in range(args.cpu_steps):
            path.step(cm, to_float_cpu(audio))[0].item()
        dt = (time.perf_counter() - t0) / args.cpu_steps
        cpu = {"value": cb * SEG_SECONDS / dt, "unit": "audio-s/s", "cores": cores, "kind": cpu_kind,
               "sample": "%d segments x %d steps of the same step on the host CPU: %s" % (cb, args.cpu_steps, cpu_what)}

    line = {
        "metric": "audio-sec/s", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (STFT: 3x%s split operands, f32 accumulate in TMEM)" % ("FP16" if (f16 or not kavg) else "TF32"), "data": "synthetic",
        "config": {"workload": "Mel+VAT step, B=%d x 20.48 s segments per GPU (BASELINE metric shape): Mel front-end + "
                               "UNet_VAT(XI=1e-6, eps=2); network = %s" % (B, "injected posteriors and input gradient "
                               "(hot path only)" if args.model == "injected" else "stand-in linear transcriber (PyTorch)"),
                   "batch_per_gpu": B, "segment_samples": SEG_SAMPLES, "parallelism": "segments sharded, dp%d, no "
                   "collective on the path" % world,
                   "input": "PCM int16 as the dataset stores it (model/dataset.py:62), scaled by 1/32768 on the device"
                            if pcm16 else "float32 in [-1, 1)",
                   "cache": "inputs rotated over %d batches = %.0f MB > 126 MB L2"
                            % (n_rot, n_rot * B * SEG_SAMPLES * (2 if pcm16 else 4) / 1e6),
                   "launch": "eager" if args.no_graphs else "CUDA graph replay, one graph of the whole step per input "
                             "buffer (%d librvb kernels per graph), %d stream(s)" % (step.kernels_per_graph, args.streams)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": B * SEG_SAMPLES * (2 if pcm16 else 4),
                "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms, "wall_ms_per_step": wall_ms / n_done,
                "api": "reconvat_b200.pipeline.HotPathStep.run_host (pinned host audio in, (vat_loss, r_norm) out; "
                       "copy of batch i+1 overlapped with the kernels of batch i%s)%s"
                       % ("" if args.no_graphs else "; graph replay",
                          "; each rank bound to its GPU's %d NUMA-local cores" % numa_cores if numa_cores else "")},
        "gpu_launches": launches,
        "ms_per_step_one_stream": ms_step_serial,
        "eager_ms_per_step": eager_ms_step,
        "kernel_timing": "each entry point re-launched %d times back to back between one CUDA-event pair, rotating over "
                         "%d recorded working sets (> L2); the contraction's entry point includes the memset of "
                         "its Mel accumulator (38 MB for the two planes of the twice-folded kernel)" % (reps, n_rot),
        "e2e_device_corpus": dev_corpus,
        "precision_strict": strict,
        "sustained": sustained,
        "module_surface": surface,
        "gpu_eager_baseline": gpu_eager,
        "roofline": roofline,
        "hbm_kernels": hbm,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def run_transcribe(args, rank, local_rank, world):
    """--workload transcribe (BASELINE config 5): one file of --file-seconds of synthetic 16 kHz
    PCM16 through whole-file inference, sharded by TIME over the ranks: Mel front-end with the file-global min / max
    (ONE NCCL MAX all-reduce of two keys per file, reconvat_b200.parallel.global_minmax_keys) and, with --model unet,
    the reference's own UNet (oracle/_ref snapshot, patched by install(attention=True, batchnorm=True)) on overlapping 640-frame
    windows (reconvat_b200.transcribe.transcribe_file).  The work per file is fixed: strong scaling.  Before timing,
    the sharded front-end is compared bit for bit with the one-rank result."""
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/rvb_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    import reconvat_b200 as R
    from reconvat_b200 import parallel, synth, transcribe
    from reconvat_b200.pipeline import MEL_KW
    parallel.bind_to_gpu_numa(local_rank)
    seconds = args.file_seconds
    L = int(seconds * 16000)
    minute = synth.music_int16(16000 * 60, 77)
    a16 = torch.from_numpy(np.tile(minute, -(-L // len(minute)))[:L].copy())
    a16[L // 2:] //= 4
    a16 = a16.pin_memory()
    mel = R.Spectrogram.MelSpectrogram(**MEL_KW).to(dev)
    n_frames = (L - 1 + 2048 - 2048) // 512 + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of the sharded front-end: all ranks' pieces == the one-rank image, bit for bit
    piece, (f0, f1) = transcribe.whole_file_frontend(mel, a16, rank, world)
    identical = None
    if world > 1:
        whole, _ = transcribe.whole_file_frontend(mel, a16, 0, 1)            # every rank can afford the whole file
        identical = torch.tensor([int(torch.equal(piece, whole[:, :, f0:f1]))], device=dev)
        dist.all_reduce(identical, op=dist.ReduceOp.MIN)
        identical = bool(identical.item())
        del whole
    del piece
    torch.cuda.empty_cache()

    model = None
    if args.model == "unet":
        from oracle import reference_loader as RL
        if RL.available():
            ns = RL.load_patched(attention=True, batchnorm=os.environ.get("RVB_BENCH_BN", "1") != "0")
            torch.manual_seed(0)
            model = ns.self_attention_VAT.UNet((2, 2), (2, 2), log=True, reconstruction=True, mode="imagewise", spec="Mel",
                                               XI=1e-6, eps=1.3).to(dev).eval()           # transcribe_files.py:63-64

    def one_pass():
        if model is None:
            return transcribe.whole_file_frontend(mel, a16, rank, world)[0]
        return transcribe.transcribe_file(model, a16, mel=mel, batch=args.batch, rank=rank, world_size=world)[0]["frame"]

    for _ in range(max(1, args.warmup)):
        out = one_pass()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = one_pass()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.summary()
    assert torch.isfinite(out).all()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    ms_step = ms / args.steps
    line = {
        "metric": "audio-sec/s", "value": seconds / (ms_step * 1e-3), "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 (STFT: 3xFP16 split operands, f32 accumulate in TMEM)",
        "data": "synthetic",
        "config": {"workload": "transcribe: one %d s file (%d frames) per step, whole-file inference sharded by time over "
                               "%d rank(s): %s" % (seconds, n_frames, world,
                                                    "Mel front-end + the reference's UNet (random init, eval) on "
                                                    "overlapping 640-frame windows, batch %d" % args.batch
                                                    if model is not None else "Mel front-end (log-Mel, file-global min/max)"),
                   "collective": "one MAX all-reduce of the two min/max keys per file (NCCL)" if world > 1 else "none",
                   "input": "PCM int16 in pinned host memory; every rank copies its own slice (+ halo) per step"},
        "sharded_equals_single_rank_bit_for_bit": identical,
        "clocks": clocks,
        "e2e": {"value": seconds / (ms_step * 1e-3), "unit": "audio-s/s",
                "h2d_bytes_per_step": int(2 * (L / world)), "d2h_bytes_per_step": 0,
                "api": "reconvat_b200.transcribe.%s" % ("transcribe_file" if model is not None else "whole_file_frontend")},
        "gpu_launches": None,
    }
    print(json.dumps(line), flush=True)


def run_train_step(args, rank, local_rank, world):
    """--workload train_step (the CALLER's step, SURVEY.md 8f rows f2 / f4 / f5 -- context for the hot-path numbers, not
    the headline): the reference's own ``UNet`` (oracle/_ref snapshot, random init) through one iteration of
    ``train_VAT_model`` (model/helper_functions.py:570-615: run_on_batch with VAT on a labelled + an unlabelled batch
    of --batch segments each, backward, Adam step).  One rank: four arms on the same GPU -- the unmodified reference, the
    same scripts behind ``reconvat_b200.install()`` (hot path only), behind ``install(attention=True)`` (plus the
    caller-side attention kernels) and behind ``install(attention=True, batchnorm=True)`` (plus the U-Net's BatchNorm2d).
    N ranks (the reference's scripts are single-GPU): the last arm, data-parallel -- every rank its own batches, the
    gradients averaged by ONE flattened NCCL all-reduce per iteration (reconvat_b200.training.allreduce_gradients);
    weak scaling, time = max over ranks."""
    import contextlib
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from oracle import reference_loader as RL
    from reconvat_b200 import parallel, synth, training
    parallel.bind_to_gpu_numa(local_rank)
    if not RL.available():
        if rank == 0:
            print(json.dumps({"workload": "train_step", "unavailable": "no reference snapshot (oracle/_ref) on this box"}))
        return
    B = args.batch if args.batch != 32 else 8                                   # train_UNet_VAT.py: batch_size = 8
    frames = 640
    L = frames * 512

    def batch(seed):
        audio = np.stack([(synth.music_int16 if (b & 1) else synth.white_int16)(L, seed * 100 + b) for b in range(B)])
        g = torch.Generator().manual_seed(seed)
        return {"audio": torch.from_numpy(synth.to_float(audio)).to(dev),
                "onset": (torch.rand(B, frames, 88, generator=g) > 0.99).float().to(dev),
                "frame": (torch.rand(B, frames, 88, generator=g) > 0.95).float().to(dev)}
    batches = [(batch(10 * rank + 2 * i + 1), batch(10 * rank + 2 * i + 2)) for i in range(3)]
    arms = (("reference", lambda: RL.load_reference()),
            ("install()", lambda: RL.load_patched()),
            ("install(attention=True)", lambda: RL.load_patched(attention=True)),
            ("install(attention=True, batchnorm=True)", lambda: RL.load_patched(attention=True, batchnorm=True)))
    if os.environ.get("RVB_BENCH_CHANNELS_LAST"):
        # what a user could do without us: the unmodified reference with channels_last weights (cuDNN's NHWC kernels)
        arms = arms + (("reference, channels_last", lambda: RL.load_reference()),
                       ("install(attention=True), channels_last", lambda: RL.load_patched(attention=True)),
                       ("install(attention=True, batchnorm=True), channels_last",
                        lambda: RL.load_patched(attention=True, batchnorm=True)))
    if world > 1:
        arms = arms[3:4]
    res = {}
    steps, warm = (args.steps if args.steps != 200 else 10), max(2, min(args.warmup, 3))
    for name, load in arms:
        ns = load()
        torch.manual_seed(0)                                                      # the same parameters on every rank
        with contextlib.redirect_stdout(sys.stderr):                              # the constructors print
            model = ns.self_attention_VAT.UNet((2, 2), (2, 2), log=True, reconstruction=True, mode="imagewise",
                                               spec="Mel", XI=1e-6, eps=2).to(dev)   # train_UNet_VAT.py:126
        if name.endswith("channels_last"):
            model = model.to(memory_format=torch.channels_last)
        model.train()
        opt = torch.optim.Adam(model.parameters(), 1e-3)

        def one(i):
            opt.zero_grad()
            bl, bu = batches[i % len(batches)]
            _, losses, _ = model.run_on_batch(bl, bu, True)
            loss = 0
            for k, v in losses.items():
                loss = loss + (v / 2 if k.startswith("loss/train_LDS") else v)
            loss.backward()
            if world > 1:
                training.allreduce_gradients(model)
            opt.step()
            return loss
        for i in range(warm):
            last = one(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.reset_peak_memory_stats()
        ev0.record()
        for i in range(steps):
            last = one(i)
        ev1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        assert bool(torch.isfinite(last))
        res[name] = {"ms_per_step": ms, "value": world * 2 * B * SEG_SECONDS / (ms * 1e-3), "unit": "audio-s/s",
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        if world > 1:
            # the ranks still hold the same parameters: the all-reduce did its job
            flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
            lo, hi = flat.clone(), flat.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            res[name]["parameters_identical_across_ranks"] = bool(torch.equal(lo, hi))
        del model, opt
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    best = "install(attention=True, batchnorm=True)"
    line = {"metric": "audio-sec/s", "workload": "train_step", "n_gpus": world, "steps": steps, "warmup": warm,
            "higher_is_better": True, "data": "synthetic", "scaling": "weak",
            "config": {"workload": "the reference's UNet (random init, train mode), one train_VAT_model iteration: "
                                   "run_on_batch(labelled B=%d, unlabelled B=%d, VAT=True) + backward + Adam step per "
                                   "rank; PyTorch default flags" % (B, B),
                       "collective": "one flattened NCCL all-reduce of the gradients (11.4 MB) per iteration"
                       if world > 1 else "none"},
            "value": res[best]["value"], "unit": "audio-s/s", "ms_per_step": res[best]["ms_per_step"], "arms": res}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32, help="segments per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=32,
                    help="segments per CPU-baseline step (default: the same B=32 step as the GPU arm)")
    ap.add_argument("--cpu-steps", type=int, default=30,
                    help="steps of the CPU baseline (32 segments each: ~10 s of host work on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--input", default="pcm16", choices=["pcm16", "f32"],
                    help="audio format handed to the front-end: the dataset's PCM int16 (default) or float32")
    ap.add_argument("--streams", type=int, default=3,
                    help="replay the per-buffer graphs on this many alternating streams (device-resident run)")
    ap.add_argument("--sustain-s", type=float, default=2.0,
                    help="seconds of back-to-back graph replays for the `sustained` block (0: skip)")
    ap.add_argument("--no-gpu-baselines", action="store_true",
                    help="skip the module-surface leg and the reference's eager GPU path")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--model", default="injected", choices=["injected", "standin", "unet"],
                    help="the black-box network the VAT loop calls (see make_model); 'unet': the reference's UNet from the "
                         "oracle/_ref snapshot (--workload transcribe only)")
    ap.add_argument("--workload", default="step", choices=["step", "transcribe", "train_step"],
                    help="step: the Mel+VAT training step (BASELINE metric); transcribe: whole-file inference (config 5)")
    ap.add_argument("--file-seconds", type=int, default=3600, help="--workload transcribe: length of the file")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun so that `python bench.py --gpus N` works on its own
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511")] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "transcribe":
        if args.steps == 200:
            args.steps = 5
        run_transcribe(args, rank, local_rank, world)
    elif args.workload == "train_step":
        run_train_step(args, rank, local_rank, world)
    else:
        if args.model == "unet":
            raise SystemExit("bench.py: --model unet belongs to --workload transcribe")
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

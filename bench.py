"""Hot-path benchmark (BASELINE.json metric: two-view clips/sec for log-mel + augmentation + Barlow Twins
loss fwd/bwd, with the fraction of the HBM / tensor-core roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" per GPU = one pass of the hot path over one batch of synthetic input:
    frontend  B = 1024 clips of 10 s @16 kHz -> crop-first 64-mel log-spectrogram -> two augmented 96-frame views
              (BASELINE config 2; the Mixup ring is warm)
    objective Barlow Twins loss forward + backward on (B, D = 8192) bf16 projector outputs (the encoder between the
              two is out of scope and stays PyTorch; embeddings are synthetic)
Under torchrun (N > 1) every rank runs this step on its own shard of the batch (weak scaling).
Prints ONE JSON line (see the contract in the task description).  `--impl reference` times the CPU port of the
reference path (oracle/torch_port.py) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "two_view_clips_per_sec"
UNIT = "clips/s"
AS_STATS = (-0.8294, 4.6230)


def _args_ns(dim):
    return types.SimpleNamespace(mixup=True, Gnoise=False, RRC=True, RLF=True, n_mels=64, crop_frames=96,
                                 virtual_crop_scale=[1.0, 1.5], local_crops_number=0, local_crops_size=[16, 16],
                                 sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160, f_min=60, f_max=7800,
                                 unit_sec=0.95, projector_out_dim=dim, HSIC=False, alpha=1.0, lmbda=0.005)


def _traffic(key="umma_dram_bytes_per_step"):
    """DRAM bytes (read + write) per launch of a kernel, from the committed `ncu --set full` captures (profiles/r2_traffic.json, else
    round 1's); None when that kernel has no capture."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get(key)
            if v is not None:
                return v
        except Exception:
            pass
    return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples SM clocks and throttle reasons DURING the timed region: NVML polled every 5 ms from a thread
    (nvidia-smi -lms as the fallback when the NVML binding is missing)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.source = None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    hi = nv.nvmlDeviceGetHandleByIndex(i)
                    if int(nv.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
                        break
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.source == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(sm), "power_w_max": max(self.power) if self.power else None, "source": "nvml, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for k, n in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


# ----------------------------------------------------------------------------------------------------------
# reference arm: the CPU port of the reference path on the host cores
# ----------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, D, L = args.batch, args.dim, int(args.clip_seconds * 16000)
    workers = cores
    host = P.host_description()
    n_steps = args.steps + args.warmup                                      # --steps and --warmup are honoured as given
    # A step = the FULL workload (all B clips through the per-sample frontend on `workers` single-threaded processes + the loss
    # fwd/bwd on all B rows).  Only if K + W full steps would not end within the time budget is the frontend timed on a bounded
    # sample of the clips; the loss always runs at full size (its D x D passes do not scale with the row count).
    t0 = time.perf_counter()
    fe_rate = P.time_frontend(B, L, workers)
    loss_s = P.time_loss(B, D)
    first_s = time.perf_counter() - t0
    clips = B
    if first_s * n_steps > args.ref_budget_s and n_steps > 1:
        fe_budget = max(0.05, (args.ref_budget_s - first_s) / (n_steps - 1) - loss_s)
        clips = int(min(B, max(workers, fe_budget * fe_rate)))
        clips = max(workers, clips // workers * workers)
    vals = []
    for s in range(n_steps):
        if s > 0:
            fe_rate = P.time_frontend(clips, L, workers)                    # clips/s on `workers` processes
            loss_s = P.time_loss(B, D) if first_s * n_steps <= args.ref_budget_s or s % 8 == 0 else loss_s
        step_s = B / fe_rate + loss_s
        if s >= args.warmup:
            vals.append(B / step_s)
    value = sum(vals) / len(vals)
    sample = (f"per step: frontend timed on {clips} of {B} clips ({workers} single-threaded worker processes, the reference's DataLoader "
              f"model); loss fwd+bwd on all {B} rows at D={D} (fp32, {cores} threads)" + ("" if clips == B else "; frontend sampled to keep "
              f"{n_steps} steps within {args.ref_budget_s:.0f} s, loss re-timed every 8th step"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * B / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "frontend_clips_per_s": fe_rate, "loss_fwd_bwd_s": loss_s, "host": host},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _config(args):
    return {"workload": f"hot-path step: frontend (BASELINE config 2: {args.batch} clips x {args.clip_seconds:g} s @16 kHz -> crop-first 64-mel log-mel -> "
                        f"two 96-frame views) + Barlow Twins loss fwd/bwd (N={args.batch} rows/GPU, D={args.dim}, bf16 in / fp32 accumulate)",
            "per_gpu_batch": args.batch, "clip_seconds": args.clip_seconds, "projector_out_dim": args.dim, "frontend_mode": "crop-first (mode C)", "streams": "frontend and loss of a step are issued on two CUDA streams (independent inputs); on several GPUs the frontend is enqueued while the embedding all-gather is in flight",
            "l2": "inputs larger than L2 (655 MB of waveforms, 128 MiB correlation matrix per step)", "parallelism": f"dp{args.gpus}"}


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import numpy as np
    import random
    import torch
    import torch.distributed as dist

    import ssl_audio_b200 as S
    from ssl_audio_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        except Exception:
            dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, D, L = args.batch, args.dim, int(args.clip_seconds * 16000)
    cfg = _args_ns(D)

    # ---- synthetic inputs, generated on the host once and resident in HBM for `value`
    g = torch.Generator(device=dev).manual_seed(rank)
    t = torch.arange(L, device=dev, dtype=torch.float32) / 16000.0
    wav = 0.1 * torch.randn(B, L, device=dev, generator=g)
    for amp in (0.3, 0.1, 0.03):
        f = 100.0 + 6900.0 * torch.rand(B, 1, device=dev, generator=g)
        ph = 6.2831853 * torch.rand(B, 1, device=dev, generator=g)
        wav += amp * torch.sin(6.2831853 * f * t[None, :] + ph)
    wav.clamp_(-1.0, 1.0)
    z1 = torch.randn(B, D, device=dev, generator=g)
    z2 = (0.6 * z1 + 0.8 * torch.randn(B, D, device=dev, generator=g)).bfloat16()
    z1 = z1.bfloat16()
    np.random.seed(rank); random.seed(rank)

    fe = S.BatchFrontend(cfg, norm_stats=AS_STATS, path="lms", mode="crop")
    crit = S.BarlowTwinsLoss(cfg, ncrops=2).to(dev)

    # The frontend of a step does not depend on its embeddings (in training it runs one batch ahead of the encoder), so the two
    # halves of the hot path are issued on two CUDA streams and overlap; a step is complete when both are.
    main_stream = torch.cuda.current_stream(dev)
    side_stream = torch.cuda.Stream(dev, priority=0)          # lowest priority; NCCL runs on a high-priority stream (below)

    cur = {}
    inputs_ready = torch.cuda.Event()

    def frontend_on_side_stream():
        side_stream.wait_event(inputs_ready)
        with torch.cuda.stream(side_stream):
            cur["views"] = fe(cur["wav"])

    def frontend_plan():            # host half (RNG replay + parameter upload), before the objective is enqueued
        side_stream.wait_event(inputs_ready)
        with torch.cuda.stream(side_stream):
            cur["handle"] = fe.prepare(cur["wav"])

    def frontend_launch():          # kernel launches only: called (on the side stream) from the objective's comm-overlap hook
        side_stream.wait_event(inputs_ready)
        cur["views"] = fe.launch(cur["handle"])

    # Plan-ahead (multi-GPU eager step): the NEXT step's random parameters are drawn and uploaded by a worker thread while this step is
    # being enqueued -- what a DataLoader worker does for the reference.  The native planner releases the GIL, the draw order is unchanged
    # (one plan per step, in step order; nothing else draws from the global generators), and the plan does not depend on the waveforms.
    plan_pool = None

    def plan_on_worker():
        with torch.cuda.stream(side_stream):
            return fe.prepare(cur["wav"])

    # Default: one host thread; on several GPUs the frontend is enqueued from the objective's comm-overlap hook (right after the embedding
    # all-gathers have been launched).  BENCH_THREAD=1 moves the frontend's host side to a worker thread instead (measured: no gain, the
    # step is device-bound); BENCH_HOOK=0/1 overrides the hook placement.
    use_thread = os.environ.get("BENCH_THREAD", "0") == "1"
    # multi-GPU: the frontend is enqueued from the objective's comm-overlap hook, i.e. right after the embedding all-gathers have been
    # launched.  Enqueueing it at step start instead lets its kernels race NCCL's for the SMs: measured bimodal at 2 ranks (0.93 or
    # 1.61 ms per step from run to run) against a steady 0.91 ms with the hook.
    use_hook = (world > 1 if os.environ.get("BENCH_HOOK") is None else os.environ["BENCH_HOOK"] == "1") and not use_thread
    pool = None
    if use_hook:
        # gated on the objective's stream position (ssl_audio_b200.dist.side_stream_hook); BENCH_HOOK_WAIT=0 shows the ungated behaviour
        from ssl_audio_b200.dist import side_stream_hook
        crit.comm_overlap_hook = side_stream_hook(side_stream, frontend_launch, gate=os.environ.get("BENCH_HOOK_WAIT", "1") == "1")
        if os.environ.get("BENCH_PLAN_AHEAD", "1") == "1":
            from concurrent.futures import ThreadPoolExecutor
            plan_pool = ThreadPoolExecutor(max_workers=1, initializer=lambda: torch.cuda.set_device(local_rank))
    elif use_thread:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=1, initializer=lambda: torch.cuda.set_device(local_rank))

    def step(wav_d, z1_d, z2_d):
        a = z1_d.detach().requires_grad_(True)
        b = z2_d.detach().requires_grad_(True)
        cur["wav"] = wav_d
        inputs_ready.record(main_stream)
        fut = None
        if pool is not None:
            fut = pool.submit(frontend_on_side_stream)
        elif use_hook and plan_pool is not None:
            pending = cur.pop("plan_future", None)
            cur["handle"] = pending.result() if pending is not None else plan_on_worker()
            cur["plan_future"] = plan_pool.submit(plan_on_worker)       # the next step's plan, drawn while this step is enqueued
        elif use_hook:
            frontend_plan()
        else:
            frontend_on_side_stream()
        loss = crit(b, a, ngcrops_each=1)          # forward(student, teacher) as main.py:115 calls it
        loss.backward()
        if fut is not None:
            fut.result()
        main_stream.wait_stream(side_stream)
        return cur["views"], loss, a.grad, b.grad

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- single GPU: the step's launches (frontend kernels on the side stream, objective forward + backward) are captured ONCE in a CUDA
    # graph and replayed; per step the host only draws the batch's random parameters (native planner), uploads them into the static plan
    # buffer and launches the graph.  Same kernels, same streams, same results as the eager step (tests/test_gpu_graph.py); BENCH_GRAPH=0
    # runs the eager step.  (Multi-GPU steps stay eager: their collectives run on the library's own streams.)
    graph_note = {"cuda_graph": False}
    if world == 1 and not use_thread and os.environ.get("BENCH_GRAPH", "1") == "1":
        try:
            eager_step = step
            for _ in range(3):                      # every lazy allocation / kernel attribute happens before the capture
                eager_step(wav, z1, z2)
            torch.cuda.synchronize(dev)
            ga = z1.detach().requires_grad_(True)
            gb = z2.detach().requires_grad_(True)
            handle = fe.prepare(wav, static=True)
            graph = torch.cuda.CUDAGraph()
            lib.abt_debug_launch_count(1)
            with torch.cuda.graph(graph):
                cap = torch.cuda.current_stream(dev)
                fork = torch.cuda.Event()
                fork.record(cap)
                side_stream.wait_event(fork)
                with torch.cuda.stream(side_stream):
                    g_views = fe.launch(handle)
                g_loss = crit(gb, ga, ngcrops_each=1)
                g_loss.backward()
                cap.wait_stream(side_stream)
            launches_per_replay = int(lib.abt_debug_launch_count(0))
            graph_note = {"cuda_graph": True, "launches_per_replay": launches_per_replay}

            def step(wav_d, z1_d, z2_d):            # noqa: F811
                assert wav_d is wav and z1_d is z1 and z2_d is z2, "the captured step reads the buffers it was captured with"
                fe.prepare(wav_d, static=True)       # this batch's draws -> static plan buffer (ordered before the replay on this stream)
                graph.replay()
                crit._pending_batches += 2           # host-side BatchNorm bookkeeping of the captured forward
                return g_views, g_loss, ga.grad, gb.grad
        except Exception as e:                   # capture refused: run the eager step and say so in the line
            try:
                torch.cuda.synchronize(dev)
            except Exception:
                pass
            graph_note = {"cuda_graph": False, "capture_error": str(e)[:200]}

    for _ in range(args.warmup):
        step(wav, z1, z2)
    sync_all()

    # ---- timed region: device-resident inputs
    lib.abt_debug_launch_count(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step(wav, z1, z2)
    e1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = int(lib.abt_debug_launch_count(0))
    if graph_note["cuda_graph"]:                # replays do not pass through the library's launch counter
        launches = graph_note["launches_per_replay"] * args.steps
    # host time to ENQUEUE one step, each step timed alone with the device idle before it: inside the loop above the enqueueing thread
    # is paced by the device (launch queue, plan staging ring), so its wall time there would only repeat ms_per_step
    host_samples = []
    for _ in range(12):
        sync_all()
        h0 = time.perf_counter()
        step(wav, z1, z2)
        host_samples.append(time.perf_counter() - h0)
    sync_all()
    host_ms = sorted(host_samples)[len(host_samples) // 2] * 1e3

    def drain_plan_ahead():         # the worker may still hold the next step's plan: finish it before anybody else uses the planner
        pending = cur.pop("plan_future", None)
        if pending is not None:
            pending.result()
    drain_plan_ahead()
    loss_val = float(out[1].detach())

    # loss-only pass with per-launch CUDA events on the launching stream (the roofline numbers: in the step above the tensor-core
    # launches share the SMs with the frontend stream, which would inflate their event times)
    step_hook = crit.comm_overlap_hook
    crit.comm_overlap_hook = None

    def loss_only():
        a = z1.detach().requires_grad_(True)
        b = z2.detach().requires_grad_(True)
        crit(b, a, ngcrops_each=1).backward()
    loss_only()
    sync_all()
    _lib.check(lib.abt_debug_timing(1))
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(args.steps):
        loss_only()
    l1.record()
    torch.cuda.synchronize(dev)
    loss_ms = l0.elapsed_time(l1) / args.steps
    stats_ms, corr_ms, grad_ms, ncalls = C.c_float(), C.c_float(), C.c_float(), C.c_int()
    _lib.check(lib.abt_debug_timing_read(C.byref(stats_ms), C.byref(corr_ms), C.byref(grad_ms), C.byref(ncalls)))
    _lib.check(lib.abt_debug_timing(0))
    crit.comm_overlap_hook = step_hook          # the multi-GPU step launches its frontend from this hook: the sustained loop below needs it
    # the library averages over CALLS; a multi-GPU step makes several (front / dz1 / dz2 phases): account per STEP
    calls_per_step = ncalls.value / float(args.steps)
    for v in (stats_ms, corr_ms, grad_ms):
        v.value = v.value * calls_per_step

    # ---- e2e: HOST buffers (pinned), through the public API.  Per step, inside the timed region: the frontend reads the
    # waveforms straight from pinned host memory (crop-first: only the cropped spans cross PCIe), the embeddings are
    # copied host->device, and the loss value is read back device->host.  Software-pipelined the way a prefetching DataLoader
    # would: while step s computes, step s+1's spans (BatchFrontend.prepare on a side stream) and embeddings cross PCIe; every
    # step's transfers, the first step's included, are issued inside the timed region.
    def run_e2e(n_steps, serial, reserve):
        """serial: embedding copies behind the span gather on ONE stream (the SM-driven gather and the DMA copies share the link badly
        side by side: tools/pcie_probe.py) instead of a stream each; reserve: SMs the tensor-core kernels leave free, so that the gather's
        few warps are never locked out by a persistent GEMM that owns every SM's registers."""
        wav_h = wav.cpu().pin_memory()
        z1_h, z2_h = z1.cpu().pin_memory(), z2.cpu().pin_memory()
        loss_h = torch.empty((), dtype=torch.float32).pin_memory()
        fe_stream, pf_stream, copy_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        views_done = torch.cuda.Event()
        zbufs = [(torch.empty_like(z1), torch.empty_like(z2)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(k):
            with torch.cuda.stream(pf_stream):
                handle = fe.prepare(wav_h)               # plan + span gather over PCIe (the frontend's only host->device traffic)
            zs = pf_stream if serial else copy_stream
            with torch.cuda.stream(zs):
                zs.wait_event(freed[k])
                zbufs[k][0].copy_(z1_h, non_blocking=True)
                zbufs[k][1].copy_(z2_h, non_blocking=True)
                ready[k].record(zs)
            return handle

        def loop(n):
            for k in range(2):
                freed[k].record()
            handle = prefetch(0)
            for s in range(n):
                k = s & 1
                with torch.cuda.stream(fe_stream):       # gathered spans -> log-mel -> views
                    views = fe.launch(handle)
                    views_done.record(fe_stream)
                if s + 1 < n:
                    handle = prefetch(k ^ 1)             # next step's span gather and embedding H2D overlap this step's kernels
                torch.cuda.current_stream(dev).wait_event(ready[k])
                a = zbufs[k][0].requires_grad_(True)
                b = zbufs[k][1].requires_grad_(True)
                loss = crit(b, a, ngcrops_each=1)
                loss.backward()
                loss_h.copy_(loss.detach(), non_blocking=True)
                freed[k].record()
                torch.cuda.current_stream(dev).wait_event(views_done)   # the step is complete when its views exist too
                a.grad = None; b.grad = None
                a.requires_grad_(False); b.requires_grad_(False)
            torch.cuda.synchronize(dev)
            return views

        S.set_reserved_sms(int(reserve))
        saved_hook, crit.comm_overlap_hook = crit.comm_overlap_hook, None      # this loop launches its frontend itself
        try:
            loop(3)
            sync_all()
            t0 = time.perf_counter()
            loop(n_steps)
            dt = time.perf_counter() - t0
        finally:
            S.set_reserved_sms(0)
            crit.comm_overlap_hook = saved_hook
        return dt, int(getattr(fe, "h2d_bytes", wav_h.numel() * 4)) + z1_h.numel() * 2 + z2_h.numel() * 2

    if args.e2e_probe:          # which arrangement of the PCIe traffic is fastest (tools: development only; prints one line per variant)
        crit.comm_overlap_hook = None
        for blocks, reserve, serials in ((4, 4, (1, 1, 0)), (4, 6, (1,)), (6, 6, (1,)), (3, 4, (1,)), (4, 4, (1,)), (8, 8, (1,)), (4, 8, (1,)), (148, 0, (0,))):
            _lib.check(lib.abt_debug_set(15, blocks))
            for serial in serials:
                dt, nb = run_e2e(20, bool(serial), reserve)
                print(json.dumps({"e2e_probe": True, "gather_blocks": blocks, "serial": serial, "reserve_sms": reserve, "ms_per_step": dt / 20 * 1e3,
                                  "clips_per_s": B * 20 / dt, "pcie_gbs": nb * 20 / dt / 1e9}), flush=True)
        return

    if args.quick:
        tq = torch.tensor([ms, loss_ms, corr_ms.value, grad_ms.value, stats_ms.value], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        if rank == 0:
            v = [float(x) for x in tq.cpu()]
            print(json.dumps({"quick": True, "n_gpus": world, "ms_per_step": v[0] / args.steps, "value": B * world / (v[0] / args.steps * 1e-3),
                              "loss_fwd_bwd_ms": v[1], "corr_ms": v[2], "grad_ms": v[3], "stats_ms": v[4], "host_enqueue_ms": host_ms,
                              "env": {k: os.environ.get(k) for k in ("ABT_DIST_CE", "ABT_COMM_MAX_CTAS", "ABT_DIST_RESERVE_SMS", "ABT_DIST_XCHG", "BENCH_HOOK_WAIT", "BENCH_HOOK", "BENCH_PLAN_AHEAD")}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # frontend-only and loss-only device times (explain `value`; not the headline)
    # (the eager call costs the host 0.23 ms, about what the two kernels take: on a box with a slower host an eager loop measures the
    # host; when graphs are available the launch pair is captured once and its replays are timed)
    fe_replay = None
    if graph_note["cuda_graph"]:
        try:
            fh = fe.prepare(wav, static=True)
            fgraph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(fgraph):
                fe.launch(fh)

            def fe_replay():            # the kernels only (one plan for all replays: a per-replay plan upload would put the box's
                fgraph.replay()         # host-to-device copy latency, 10-150 us by box, in series with every launch pair)
        except Exception:
            fe_replay = None
            torch.cuda.synchronize(dev)
    fe_call = fe_replay if fe_replay is not None else (lambda: fe(wav))
    for _ in range(3):
        fe_call()
    torch.cuda.synchronize(dev)
    fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fe0.record()
    for _ in range(args.steps):
        fe_call()
    fe1.record()
    torch.cuda.synchronize(dev)
    fe_ms = fe0.elapsed_time(fe1) / args.steps

    # ---- mode F (BASELINE config 2 as worded): the full (B, 64, 1001) log-mel is emitted as well as the two views
    fe_full = S.BatchFrontend(cfg, norm_stats=AS_STATS, path="lms", mode="full")
    for _ in range(3):
        fe_full(wav)
    torch.cuda.synchronize(dev)
    ff0, ff1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ff0.record()
    for _ in range(args.steps):
        fe_full(wav)
    ff1.record()
    torch.cuda.synchronize(dev)
    fe_full_ms = ff0.elapsed_time(ff1) / args.steps
    del fe_full

    # ---- BASELINE config 3: loss fwd+bwd sweep, N = 128 rows per GPU, D = 2048 / 4096 / 8192 (global batch 128 x world).
    # Every iteration is timed alone with CUDA events after an L2 flush (the 2 MB embeddings would otherwise sit in L2).
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sweep = []
    for d_s in (2048, 4096, 8192):
        gs = torch.Generator(device=dev).manual_seed(100 + rank)
        a0 = torch.randn(128, d_s, device=dev, generator=gs)
        b0 = (0.6 * a0 + 0.8 * torch.randn(128, d_s, device=dev, generator=gs)).bfloat16()
        a0 = a0.bfloat16()
        crit_s = S.BarlowTwinsLoss(_args_ns(d_s), ncrops=2).to(dev)

        def one_module():
            a = a0.detach().requires_grad_(True)
            b = b0.detach().requires_grad_(True)
            lo = crit_s(b, a, ngcrops_each=1)
            lo.backward()
            return lo

        def one():
            # the library call itself (loss + both gradients in one C-ABI call); the module adds autograd glue and one scaling launch
            if world > 1:
                from ssl_audio_b200 import dist as _D
                return _D.bt_loss_fwd_bwd_global(b0, a0, 1.0, 0.005, False)[0]
            return S.bt_loss_fwd_bwd(b0, a0, 1.0, 0.005, False)[0]
        for _ in range(5):
            one()
            one_module()
        sync_all()
        # (a) whole call, every iteration alone after an L2 flush; no per-launch events, so the launches keep their programmatic dependency
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.sweep_iters)]
        for e_a, e_b in evs:
            flush.zero_()
            e_a.record()
            lo = one()
            e_b.record()
        torch.cuda.synchronize(dev)
        times = sorted(e_a.elapsed_time(e_b) for e_a, e_b in evs)
        # (b) the tensor-core launch(es) alone: CUDA events recorded by the library around them on the launching stream
        _lib.check(lib.abt_debug_timing(1))
        for _ in range(args.sweep_iters):
            one()
        torch.cuda.synchronize(dev)
        st_s, co_s, gr_s, nc_s = C.c_float(), C.c_float(), C.c_float(), C.c_int()
        _lib.check(lib.abt_debug_timing_read(C.byref(st_s), C.byref(co_s), C.byref(gr_s), C.byref(nc_s)))
        _lib.check(lib.abt_debug_timing(0))
        per = nc_s.value / float(args.sweep_iters)
        # (c) through the module (BarlowTwinsLoss + backward): the same call plus autograd glue -- host-bound at this size
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(args.sweep_iters):
            one_module()
        m1.record()
        torch.cuda.synchronize(dev)
        sweep.append([d_s, times[len(times) // 2], (co_s.value + gr_s.value) * per, st_s.value * per, float(lo.detach()), m0.elapsed_time(m1) / args.sweep_iters])
        del crit_s
    del flush

    # ---- projector tail fused with the statistics pass, single GPU only: last bias-free Linear (hidden 8192 -> D) of both
    # views + column statistics in one tensor-core launch, against the library GEMMs + the separate statistics kernels it replaces
    proj = None
    if world == 1 and not args.no_proj_tail:
        from ssl_audio_b200.projector import proj_tail_fwd, proj_tail_forward_loss
        Kp = args.proj_hidden
        gp = torch.Generator(device=dev).manual_seed(7)
        hp1 = torch.relu(torch.randn(B, Kp, device=dev, generator=gp)).bfloat16()
        hp2 = torch.relu(0.7 * hp1.float() + 0.7 * torch.randn(B, Kp, device=dev, generator=gp)).bfloat16()
        Wp = (torch.randn(D, Kp, device=dev, generator=gp) / Kp ** 0.5).bfloat16()
        packp = torch.empty(7 * D, device=dev)
        crit_p = S.BarlowTwinsLoss(cfg, ncrops=2).to(dev)

        def t_ms(fn, n=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(n):
                fn()
            b_.record()
            torch.cuda.synchronize(dev)
            return a_.elapsed_time(b_) / n

        def chain(fused):
            a_ = hp1.detach().requires_grad_(True); b_ = hp2.detach().requires_grad_(True); w_ = Wp.detach().requires_grad_(True)
            if fused:
                lo_ = proj_tail_forward_loss(crit_p, a_, b_, w_)
            else:
                lo_ = crit_p.forward_loss(torch.nn.functional.linear(a_, w_), torch.nn.functional.linear(b_, w_))
            lo_.backward()
            return lo_
        fused_ms = t_ms(lambda: proj_tail_fwd(hp1, hp2, Wp, packp))
        lib_ms = t_ms(lambda: (torch.nn.functional.linear(hp1, Wp), torch.nn.functional.linear(hp2, Wp)))
        chain_fused = t_ms(lambda: chain(True), 5)
        chain_lib = t_ms(lambda: chain(False), 5)
        lf_, lu_ = float(chain(True).detach()), float(chain(False).detach())
        proj = [fused_ms, lib_ms, chain_fused, chain_lib, lf_, lu_, Kp]
        del hp1, hp2, Wp, crit_p

    # ---- sustained loop (>= 2 s of back-to-back steps, clocks sampled throughout): the burst number above is 12 ms of work
    sus_steps = int(max(args.steps, min(20000, args.sustained_s / max(ms / args.steps * 1e-3, 1e-5))))
    if world > 1:       # every rank must run the same number of (collective) steps: take rank 0's count
        ns = torch.tensor([sus_steps], dtype=torch.int64, device=dev)
        dist.broadcast(ns, 0)
        sus_steps = int(ns.item())
    sampler2 = ClockSampler(local_rank)
    sync_all()
    if rank == 0:
        sampler2.start()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(sus_steps):
        step(wav, z1, z2)
    u1.record()
    sync_all()
    clocks_sus = sampler2.stop() if rank == 0 else None
    drain_plan_ahead()
    sus_ms = u0.elapsed_time(u1) / sus_steps

    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    # measured (--e2e-probe, one B200): one PCIe stream + 4 SMs left free for the 4-block gather 1.97-2.02 ms per step; gather and
    # copies on a stream each and nothing reserved 2.14-2.27 ms
    e2e_serial = os.environ.get("BENCH_E2E_SERIAL", "1") == "1"
    e2e_reserve = int(os.environ.get("BENCH_E2E_RESERVE", "4"))
    e2e_s, h2d = run_e2e(e2e_steps, e2e_serial, e2e_reserve)
    d2h = 4

    # ---- reduce over ranks (max time)
    base = [ms, e2e_s, corr_ms.value, grad_ms.value, fe_ms, loss_ms, stats_ms.value, fe_full_ms, sus_ms]
    tt = torch.tensor(base + [v for row in sweep for v in row[1:4]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    vals = [float(v) for v in tt.cpu()]
    ms, e2e_s, corr, grad, fe_ms, loss_ms, st_ms, fe_full_ms, sus_ms = vals[:9]
    for k, row in enumerate(sweep):
        row[1:4] = vals[9 + 3 * k: 12 + 3 * k]
    # every rank must have arrived at the same global loss (multi-GPU: a cheap invariant of the exchange; single GPU: trivially true)
    lv = torch.tensor([loss_val] + [row[4] for row in sweep] + [-loss_val] + [-row[4] for row in sweep], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lv, op=dist.ReduceOp.MAX)
    lv = [float(v) for v in lv.cpu()]
    nl = 1 + len(sweep)
    loss_spread = max(abs(lv[k] + lv[nl + k]) / max(abs(lv[k]), 1e-30) for k in range(nl))      # (max - min) / max over ranks
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    e2e_value = B * world * e2e_steps / e2e_s
    if world > 1 and loss_spread > 1e-6:
        raise RuntimeError(f"ranks disagree on the global loss (relative spread {loss_spread:.2e}): the multi-GPU exchange is broken")
    flops = 6.0 * B * D * D                                   # algorithmic FLOPs of one loss term, per GPU = 6 N_g D^2 / R
    sweep_out = []
    for d_s, call_ms, kern_ms, stat_ms_s, lo_s, mod_ms in sweep:
        fl = 6.0 * 128 * d_s * d_s                             # per GPU; the global batch is 128 x world rows
        ach = fl / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
        sweep_out.append({
            "D": d_s, "rows_per_gpu": 128, "global_rows": 128 * world, "ms": call_ms, "clips_per_s": 128 * world / (call_ms * 1e-3), "loss_value": lo_s,
            "module_ms": mod_ms,
            "roofline": {"bound": "tensor", "kernel": "bt_fused_kernel (one launch: S -> loss, P -> gradients, batch-norm backward)" if world == 1
                         else "bt_umma_kernel (CORR + GRAD launches of the row-block step)",
                         "achieved": ach, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": ach / peaks["tf_burst"],
                         "traffic": _traffic(f"fused_dram_bytes_n128_d{d_s}") if world == 1 else None,
                         "algorithmic_flops": fl, "algorithmic_bytes": 6.0 * 128 * d_s * 2, "kernel_ms": kern_ms, "stats_ms": stat_ms_s,
                         "frac_whole_call": fl / (call_ms * 1e-3) / 1e12 / peaks["tf_burst"],
                         "executed_flops": (8.0 if world == 1 else 6.0) * 128 * d_s * d_s}})
    tc_ms = corr + grad
    achieved = flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    fe_bytes = B * (64896 + 49152 + 49152 + 24576)            # mode C algorithmic bytes per clip
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(args),
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms, "step_launch": graph_note,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "reserved_sms": e2e_reserve, "pcie_gbs": h2d * e2e_steps / e2e_s / 1e9 if world == 1 else None,
                "note": "pinned host buffers through BatchFrontend.prepare / launch (host wav) + BarlowTwinsLoss; crop-first span gather reads only the cropped "
                        "samples over PCIe; the next step's span gather and embedding H2D run on one side stream under the current step's kernels, "
                        "which leave reserved_sms SMs to the gather (PCIe-bound: h2d_bytes_per_step at the box's ~52-55 GB/s is the floor)"},
        "roofline": {"bound": "tensor", "kernel": "bt_umma_kernel (CORR + GRAD launches)", "achieved": achieved, "peak": peaks["tf_burst"],
                     "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"], "traffic": _traffic(),
                     "algorithmic_flops_per_step": flops, "corr_ms": corr, "grad_ms": grad, "stats_ms": st_ms, "loss_fwd_bwd_ms": loss_ms,
                     "measured": "CUDA events around the CORR and GRAD launches on the launching stream, loss-only pass of the same bench run, summed per step",
                     "launch_groups_per_step": calls_per_step,
                     "frac_of_sustained_peak": achieved / peaks["tf_sustained"],
                     "peak_source": peaks["source"] + " (bf16_tflops burst figure: the launches are timed alone, in short bursts at boost clocks)"},
        "loss_sweep": {"workload": "BASELINE config 3: Barlow Twins loss fwd+bwd, 128 rows per GPU (global batch 128 x n_gpus), bf16 in, "
                                   "ms = median device time of the library call (loss + both gradients: statistics + tensor-core launch(es)), every "
                                   "iteration timed alone with CUDA events after an L2 flush, max over ranks; module_ms = mean per call through "
                                   "BarlowTwinsLoss + backward, back to back (adds autograd glue and one scaling launch; host-bound at this size); "
                                   "roofline.frac = the tensor-core launch(es) alone (library events, a second loop), frac_whole_call = the same FLOP over ms",
                       "iters": args.sweep_iters, "points": sweep_out},
        "sustained": {"steps": sus_steps, "ms_per_step": sus_ms, "value": B * world / (sus_ms * 1e-3), "clocks": clocks_sus,
                      "note": "same step, back to back for >= %.1f s" % args.sustained_s},
        "loss_spread_over_ranks": loss_spread,
        "frontend_full": {"mode": "F (full log-mel emitted + two views)", "ms_per_step": fe_full_ms, "clips_per_s": B / (fe_full_ms * 1e-3),
                          "algorithmic_bytes_per_step": B * 1019136, "achieved_gbs": B * 1019136 / (fe_full_ms * 1e-3) / 1e9,
                          "hbm_frac": B * 1019136 / (fe_full_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "algorithmic_fp32_tflops": B * 29.6e6 / (fe_full_ms * 1e-3) / 1e12, "fp32_peak_tflops": 74.4},
        "frontend": {"ms_per_step": fe_ms, "clips_per_s": B / (fe_ms * 1e-3), "algorithmic_bytes_per_step": fe_bytes,
                     "achieved_gbs": fe_bytes / (fe_ms * 1e-3) / 1e9, "hbm_frac": fe_bytes / (fe_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "algorithmic_fp32_tflops": B * 2.84e6 / (fe_ms * 1e-3) / 1e12, "fp32_peak_tflops": 74.4,
                     "note": "log-mel + views launches of a frontend-only pass; DRAM traffic = algorithmic bytes, but the kernel is bound by shared-memory "
                             "bandwidth and FP32 issue (1024-point FFT per 160 new samples: 2.84 MFLOP per clip against 187 776 B), see DESIGN.md and "
                             "profiles/r1_ncu_summary.md; fp32 peak = 148 SMs x 128 FMA/clk x 1.965 GHz"},
        "loss_value": loss_val,
    }
    if proj is not None:
        pf = 2 * 2.0 * B * proj[6] * D
        line["proj_tail"] = {
            "workload": "projector tail fused with the statistics pass (SURVEY 8 f2): z = h W^T for both views, N = %d, K = %d, D = %d, bf16" % (B, proj[6], D),
            "fused_ms": proj[0], "library_gemms_ms": proj[1], "statistics_kernel_it_replaces_ms": 0.0145,
            "roofline": {"bound": "tensor", "kernel": "bt_umma_kernel<2> LINEAR mode + bt_pack_fold_kernel", "achieved": pf / (proj[0] * 1e-3) / 1e12,
                         "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": pf / (proj[0] * 1e-3) / 1e12 / peaks["tf_burst"], "traffic": _traffic("proj_tail_dram_bytes") if (B, proj[6], D) == (1024, 8192, 8192) else None,
                         "algorithmic_flops": pf},
            "chain_fwd_bwd_ms": {"fused_node": proj[2], "linear_plus_loss_module": proj[3]},
            "loss_fused_vs_unfused": [proj[4], proj[5]]}
    if args.cpu_baseline and world == 1:          # the CPU baseline is timed on rank 0 of the single-GPU run only
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """The CPU port of the reference path on the host cores, FULL workload (no sampling), plus the variants BASELINE.md section 4 lists."""
    import torch
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, D, L = args.batch, args.dim, int(args.clip_seconds * 16000)
    fe_rate = P.time_frontend(B, L, cores)
    loss_s = P.time_loss(B, D)
    step_s = B / fe_rate + loss_s
    batched = P.time_frontend_batched(min(B, 128), L)
    sweep = {str(d): P.time_loss(128, d) for d in (2048, 4096, 8192)}
    return {"value": B / step_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"full workload, one pass: frontend on all {B} clips over {cores} single-threaded worker processes (the reference's DataLoader model); "
                      f"loss fwd+bwd on all {B} rows at D={D}, fp32, {cores} threads",
            "frontend_clips_per_s": fe_rate, "loss_fwd_bwd_s": loss_s,
            "frontend_batched_clips_per_s": batched, "frontend_batched_note": f"single process, one batched MelSpectrogram call on {min(B, 128)} clips + per-sample views, {cores} intra-op threads",
            "loss_fwd_bwd_s_n128": sweep, "host": P.host_description()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="clips (= embedding rows) per GPU per step")
    ap.add_argument("--dim", type=int, default=8192, help="projector_out_dim")
    ap.add_argument("--clip-seconds", type=float, default=10.0)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-proj-tail", action="store_true", help="skip the projector-tail (f2) measurement")
    ap.add_argument("--proj-hidden", type=int, default=8192, help="input features of the projector's last Linear (projector_hidden_dim)")
    ap.add_argument("--e2e-probe", action="store_true", help="single GPU, development: time the e2e loop for several PCIe arrangements and exit")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="reference arm: wall-clock budget of the whole K + W step run")
    ap.add_argument("--sweep-iters", type=int, default=30, help="timed iterations per point of the N = 128 loss sweep (BASELINE config 3)")
    ap.add_argument("--sustained-s", type=float, default=2.5, help="length of the sustained-clock loop reported beside the burst number")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--quick", action="store_true", help="tuning aid, not a bench line: only the timed step and the loss-only pass, printed as a short JSON line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                         # timing rule: at least 3 warm-up steps (also warms the Mixup ring)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

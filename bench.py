#!/usr/bin/env python
"""Headline benchmark: images/sec of the Squeeze-ErNet hot path on synthetic 240x240 RGB frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch: uint8 (B,240,240,3) frames -> eval transform
(Resize 159 / CenterCrop 140 / Normalize) -> Squeeze_ErNET forward -> (B,5) probabilities.
Workload at N=1 = BASELINE.json configs[1]: Squeeze-ErNet, bf16, batch 256.  At N>1 = configs[4]: 8192 frames per
step split evenly over the N ranks (strong scaling; contiguous slices, one handle + weight copy per rank, no collective
on the hot path - the optional all-gather of the probabilities is timed separately).  `--batch B` gives B frames per
GPU at any N instead (weak scaling).

  value   frames/s with the frames already resident in HBM, timed with CUDA events on the launch
          stream, barrier + synchronize on both sides, max over ranks.
  e2e     same metric through the host-buffer entry point (ernet_classify_frames_host): pinned host
          frames -> H2D -> kernels -> D2H of the probabilities, every step.
  roofline  dominant kernel, timed live with CUDA events around each of its launches (second pass
          over the same steps with the library's per-stage timing switched on).
  cpu_baseline / --impl reference: the reference's CPU path (torch.nn.functional restatement under
          oracle/, plus the torchvision/Pillow transform) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

ARCH = "squeeze-ernet"
PRECISION = "bf16"
BATCH = 256                     # per GPU (N=1: BASELINE configs[1])
SHARDED_TOTAL = 8192            # N>1: BASELINE configs[4], split evenly over the ranks
FRAME = (240, 240)
N_INPUT_SETS = 8                # 8 x 44 MB of frames > 126 MB L2: inputs never come from L2
MACS_PER_IMAGE = {"squeeze-ernet": 45.48e6, "squeeze-redconv": 38.80e6}

# algorithmic work per image of each stage: (flops, bytes moved at 16-bit)
STAGE_WORK = {
    # frames path: transform + conv1 fused (uint8 frame in, stem tensor out; the 140x140 tensor stays on chip)
    "ingest": (2 * 2.057e6, 240 * 240 * 3 + 2 * 72 * 72 * 16),
    "stem": (2 * 2.057e6, 140 * 140 * 3 * 2 + 69 * 69 * 16 * 2),
    "dw1": (2 * 1.939e6, 69 * 69 * 16 * 2 + 66 * 66 * 48 * 2),
    "pw1": (2 * 13.790e6 * (66 * 66) / (67 * 67), 66 * 66 * 48 * 2 + 33 * 33 * 64 * 2),
    "dw2": (2 * 1.661e6, 33 * 33 * 64 * 2 + 30 * 30 * 192 * 2),
    "pw2": (2 * 17.713e6 * (30 * 30) / (31 * 31), 30 * 30 * 192 * 2 + 15 * 15 * 96 * 2),
    "dw3": (2 * 0.438e6, 15 * 15 * 96 * 2 + 12 * 12 * 288 * 2),
    "pw3": (2 * 6.230e6 * (12 * 12) / (13 * 13), 12 * 12 * 288 * 2 + 6 * 6 * 128 * 2),
    "dw4": (2 * 0.055e6, 6 * 6 * 128 * 2 + 4 * 4 * 384 * 2),
    "pw4": (2 * 1.573e6, 4 * 4 * 384 * 2 + 4 * 4 * 256 * 2),
    "head": (2 * 0.021e6, 4 * 4 * 256 * 2 + 20),
    # fused tensor-core blocks: algorithmic flops = depthwise + 1x1 MACs of the region the pool keeps
    # (NOT the 25-tap dense flops the kernel actually issues), bytes = P8 image in + P8 / NHWC image out
    "tc_block1": (2 * (66 * 66 * 16 * 27 + 66 * 66 * 48 * 64), 2 * 72 * 72 * 16 + 8 * 36 * 36 * 16),
    "tc_block2": (2 * (30 * 30 * 64 * 27 + 30 * 30 * 192 * 96), 8 * 36 * 36 * 16 + 12 * 18 * 18 * 16),
    "tc_block3": (2 * (12 * 12 * 96 * 27 + 12 * 12 * 288 * 128), 12 * 18 * 18 * 16 + 6 * 6 * 128 * 2),
    "tc_block4": (2 * (0.055e6 + 1.573e6 + 0.021e6), 6 * 6 * 128 * 2 + 20),          # ACFF4 + head in one kernel
}
GEMM_STAGES = {"pw1", "pw2", "pw3", "pw4", "tc_block1", "tc_block2", "tc_block3", "tc_block4"}
# FLOPs the tensor-core block kernels actually ISSUE per image (25-tap dense form, padded 16x8 tiles): reported next to
# the algorithmic figure so that the roofline fraction and the tensor-pipe utilisation can both be read off
def issued_flops(arch, precision):
    """Operations the tensor pipe really executes per image (padded tiles, all taps), by stage.  16-bit: 2*128*N*16
    per MMA; int8 (kind::i8): 2*128*N*32.  Block 1 issues 13 two-tap MMAs per tile when its input has one real 16-byte
    chunk (int8 Squeeze_ErNET, 16-bit Squeeze_RedConv), else 25."""
    k = 32 if precision == "int8" else 16
    red = arch == "squeeze-redconv"
    b1_taps = 13 if (precision == "int8" or red) else 25
    ks2 = 2 if precision == "int8" else 4                       # K steps per tap: 64 channels
    c3 = 48 if red else 96
    ks3 = c3 // (32 if precision == "int8" else 16)
    c4 = 64 if red else 128
    # block 1: 4 x 9 regular 16x8 tiles + the 2 tiles of the tail unit (rows 64, 65; csrc/tc_pblock.cuh PCfg::TAIL); 45 with
    # ERNET_TAIL_TILES=0
    b1_tiles = 38 if os.environ.get("ERNET_TAIL_TILES", "1") != "0" else 45
    out = {"tc_block1": b1_tiles * b1_taps * 2 * 128 * 64 * k, "tc_block2": 8 * 25 * ks2 * 2 * 128 * 96 * k,
           "tc_block3": 2 * 25 * ks3 * 2 * 128 * 128 * k, "tc_block4": (3 * c4 // 16) * 2 * 128 * 256 * 16 // 2}
    if red:
        out["red2"] = 8 * 6 * 2 * 128 * 64 * 16                 # conv_red2 as a 1-tap instance: 8 tiles x 6 K steps, N = 64
    return out


def workload_name(arch, precision, batch, world=1, sharded_total=None):
    model = {"squeeze-ernet": "Squeeze-ErNet", "squeeze-redconv": "Squeeze-ErNet-RedConv"}[arch]
    tail = "240x240x3 uint8 frames -> eval transform -> forward -> probabilities"
    if sharded_total:
        return (f"{model} {precision} batch-sharded {sharded_total} images across {world}xB200, {batch} per GPU "
                f"(BASELINE.json configs[4]): {tail}")
    cfg = {("squeeze-ernet", "bf16", 256): "BASELINE.json configs[1]", ("squeeze-redconv", "fp16", 1024): "BASELINE.json configs[2]",
           ("squeeze-ernet", "int8", 4096): "BASELINE.json configs[3]"}.get((arch, precision, batch), "non-headline configuration")
    return f"{model} {precision} batch {batch} per GPU on {world}xB200 ({cfg}): {tail}"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        idle_sm, idle_mx = [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                clk, cmax, util = float(f[1]), float(f[2]), float(f[9])
            except ValueError:
                continue
            if util < 5.0:                     # the sampler runs from process start: keep only samples taken under load
                idle_sm.append(clk); idle_mx.append(cmax)
                continue
            sm.append(clk); mx.append(cmax)
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm and idle_sm:                 # run too short for a loaded sample: report what there is, and say so
            return {"sm_mhz": float(np.median(idle_sm)), "sm_max_mhz": float(max(idle_mx)), "samples": 0, "idle_samples": len(idle_sm),
                    "reasons": ["no sample under load"]}
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(steps, warmup, sample_frames, ARCH=ARCH):
    """Times the reference's CPU path on a bounded sample per step.  Returns (frames_per_s, info)."""
    import fixtures
    from oracle import ernet_torch as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = T.to_torch_sd(fixtures.get_state_dict(ARCH, "shipped"))
    frames = fixtures.noise_frames(sample_frames, *FRAME, seed=1234)
    how = ""
    for _ in range(warmup):
        x, how = T.transform_frames(frames)
        T.forward(sd, x, ARCH)
    t0 = time.perf_counter()
    for _ in range(steps):
        x, how = T.transform_frames(frames)                 # aider-predict.py:57-66, per frame through PIL
        probs, _ = T.forward(sd, x, ARCH)                   # aider-predict.py:76, fp32, all host threads
        float(probs[0, 0])
    dt = time.perf_counter() - t0
    fps = steps * sample_frames / dt
    info = {"value": fps, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps x {sample_frames} frames of 240x240x3 uint8: {how} + torch.nn.functional fp32 "
                      f"forward (oracle/ernet_torch.py), torch.set_num_threads({cores})"}
    return fps, dt, info


def common_config(arch, precision, batch, world, total):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": workload_name(arch, precision, batch, world, total), "arch": arch, "batch_per_gpu": batch,
            "global_batch": batch * world, "frame": "240x240x3 u8",
            "weights": "shipped state_dict of the reference (tests/golden/weights)",
            "l2": "inputs rotate over distinct batches totalling > 126 MB (L2) per rank",
            "parallelism": f"dp{world}, contiguous image slices, no collective on the hot path"}


def resolve_batch(args, world):
    """(frames per GPU per step, total of the sharded configuration or None)."""
    if args.batch is not None:
        return args.batch, None
    if world > 1 and args.arch == ARCH and args.precision == PRECISION:
        return SHARDED_TOTAL // world, SHARDED_TOTAL                  # BASELINE configs[4]
    return BATCH, None


def reference_main(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores, same configuration as the GPU arm: every
    step is one full batch of that configuration (N=1: 256 frames), same warm-up count.  The step count is bounded so
    that the arm ends within a few minutes (a 256-frame step takes ~0.4-2.5 s depending on the host)."""
    if rank != 0:
        return 0
    batch, total = resolve_batch(args, world)
    sample = min(batch, 256)                         # bounded sample of the step's batch (equal to it at N=1)
    warmup = args.warmup
    t_probe = time.perf_counter()
    cpu_reference_run(1, 0, sample, args.arch)
    per_step = time.perf_counter() - t_probe
    steps = int(max(2, min(args.steps, 150.0 / max(per_step, 1e-3) - warmup)))
    warmup = int(max(1, min(warmup, 60.0 / max(per_step, 1e-3))))
    fps, dt, info = cpu_reference_run(steps, warmup, sample, args.arch)
    line = {
        "impl": "reference", "metric": "images/sec", "value": fps, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if total else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args.arch, args.precision, batch, world, total),
        "reference_arm": f"CPU fp32 on the host cores, {sample} of the configuration's frames per step",
        "cpu_baseline": info,
        "e2e": {"value": fps, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------- GPU arm
def pin_rank_to_cores(local_rank, world):
    """Each rank gets its own slice of the host cores (ranks sharing cores is what made short multi-GPU runs jitter)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(len(mine), 8)))
        return len(mine)
    except Exception:
        return None


def gpu_main(args, rank, local_rank, world):
    import torch.distributed as dist
    import fixtures
    import rtdm_b200
    from rtdm_b200 import parallel

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ncores = pin_rank_to_cores(local_rank, world) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                       # started first: nvidia-smi needs a few hundred ms before its first sample
    ARCH, PRECISION = args.arch, args.precision
    BATCH, sharded_total = resolve_batch(args, world)
    sd = fixtures.get_state_dict(ARCH, "shipped")
    model = rtdm_b200.from_state_dict(ARCH, sd, dev, PRECISION)
    if PRECISION == "int8":
        model.calibrate()
    model.prepare_ingest(*FRAME)
    model.set_persistent({"pairs": 2, "persistent": 1, "per-image": 0}[args.schedule])

    # synthetic inputs: distinct batches per rank (seeded by rank), rotated so that a step never finds its frames in
    # L2 (at least 2 sets, and > 126 MB of frames in total)
    n_sets = max(2, min(N_INPUT_SETS, -(-(160 << 20) // (BATCH * 172800))))
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_sets = [torch.randint(0, 256, (BATCH, *FRAME, 3), dtype=torch.uint8, generator=g).pin_memory()
                 for _ in range(n_sets)]
    dev_sets = [h.to(dev) for h in host_sets]
    stream = torch.cuda.current_stream(dev)

    def step(i):
        return model.forward_frames(dev_sets[i % n_sets])

    for i in range(args.warmup):
        step(i)
    barrier()
    # The timed region must measure the device, not how fast Python enqueues: a spin kernel holds the stream while the
    # host runs ahead (the start event is recorded behind it), so the K steps execute back to back.  Long runs simply
    # block in the launch queue until the spin ends - no host/device hand-shake that could dead-lock.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if hasattr(torch.cuda, "_sleep"):
        torch.cuda._sleep(int(4e6))           # ~2 ms at 1.9 GHz
    e0.record(stream)
    for i in range(args.steps):
        out = step(i)
    e1.record(stream)
    barrier()
    ms_max = max_over_ranks(e0.elapsed_time(e1))
    value = world * BATCH * args.steps / (ms_max * 1e-3)
    assert torch.isfinite(out).all()

    # ---- optional collection of the probabilities (the only collective a sharded run may want), timed on its own
    gather_ms = None
    if world > 1:
        sizes = [BATCH] * world
        for _ in range(3):
            parallel.gather_rows(out, sizes)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(20):
            allp = parallel.gather_rows(out, sizes)
        g1.record(stream)
        barrier()
        gather_ms = max_over_ranks(g0.elapsed_time(g1) / 20)
        assert allp.shape == (world * BATCH, 5)

    # ---- e2e: host frames in, host probabilities out, through the C ABI host entry point.  Streaming use of the public
    # API: one call stays in flight while the next is submitted (classify_host_submit / .result()), so the copy of step
    # i+1 runs under the kernels of step i; every step's H2D copy and D2H read-back are inside the timed region.
    for i in range(min(args.warmup, 3)):
        model.classify_host(host_sets[i % n_sets])
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 50, 12800 // BATCH + 2))
    pending = None
    for i in range(e2e_steps):
        nxt = model.classify_host_submit(host_sets[i % n_sets])
        if pending is not None:
            ph = pending.result()
        pending = nxt
    ph = pending.result()
    torch.cuda.synchronize(dev)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = world * BATCH * e2e_steps / (e2e_ms * 1e-3)
    assert np.isfinite(ph).all()
    # the same, one blocking call per step (no overlap across steps)
    t0 = time.perf_counter()
    nblk = min(e2e_steps, 20)
    for i in range(nblk):
        model.classify_host(host_sets[i % n_sets])
    e2e_blocking = world * BATCH * nblk / max_over_ranks(time.perf_counter() - t0)
    # H2D ceiling: what the same bytes cost as plain pinned cudaMemcpyAsync copies, all ranks at once (no kernels):
    # the bound of e2e on this host, and the number that says whether ranks starve each other on the host side
    copy_bytes = BATCH * model.host_copy_bytes_per_frame(*FRAME)
    src = [h.view(-1)[:copy_bytes] for h in host_sets]
    dst = torch.empty(copy_bytes, dtype=torch.uint8, device=dev)
    for i in range(3):
        dst.copy_(src[i % n_sets], non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    ncopy = max(10, min(50, (2 << 30) // copy_bytes))
    for i in range(ncopy):
        dst.copy_(src[i % n_sets], non_blocking=True)
    torch.cuda.synchronize(dev)
    h2d_s = max_over_ranks(time.perf_counter() - t0)
    h2d_gbs_per_rank = copy_bytes * ncopy / h2d_s / 1e9
    h2d_ceiling_value = world * BATCH * ncopy / h2d_s          # images/s if a step cost only its copy

    # ---- single-frame latency (BASELINE config 1 flavour: one 240x240 frame -> probabilities), device-resident
    one = dev_sets[0][:1]
    for _ in range(10):
        model.forward_frames(one)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(stream)
    for _ in range(200):
        model.forward_frames(one)
    l1.record(stream)
    torch.cuda.synchronize(dev)
    latency_b1_ms = l0.elapsed_time(l1) / 200
    # the same through one CUDA-graph launch per frame (model.graph_frames: the captured chain keeps its PDL edges)
    runner = model.graph_frames(one.clone())
    for _ in range(10):
        runner()
    l0.record(stream)
    for _ in range(200):
        runner()
    l1.record(stream)
    torch.cuda.synchronize(dev)
    latency_b1_graph_ms = l0.elapsed_time(l1) / 200

    # ---- live per-kernel timing: same steps again with CUDA events around every stage
    barrier()
    model.profile(True)
    prof_steps = min(args.steps, 200)
    for i in range(prof_steps):
        step(i)
    torch.cuda.synchronize(dev)
    prof = model.profile_read()
    model.profile(False)
    # clocks: sampled from the warm-up, through the timed region, to the end of the per-kernel pass (the same steps again) -
    # a 20-step timed region lasts 4 ms, shorter than one nvidia-smi sample, so the window has to be the loaded phases around it
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = load_peaks()
        depthwise = depthwise_roofline(dev, peaks)
        # ---- parity of this very configuration against the reference graph (fp64, CPU) on the first input set
        parity = parity_check(model, dev_sets[0], host_sets[0], ARCH, PRECISION, sd)
        fused = "ingest" not in prof and "stem" not in prof      # transform + conv1 ran inside block 1's kernel (tc_fblock.cuh)
        stage_work = dict(STAGE_WORK)
        if fused:
            f1, b1 = STAGE_WORK["tc_block1"]
            stage_work["tc_block1"] = (f1 + STAGE_WORK["ingest"][0], 240 * 240 * 3 + 8 * 36 * 36 * 16)   # frame in, pool1 out
        total_stage_ms = sum(v[0] for v in prof.values())
        dom = max(prof, key=lambda k: prof[k][0])
        dms, dcount = prof[dom]
        per_launch_ms = dms / dcount
        launches_per_step = dcount / prof_steps                      # chunks per step
        imgs_per_launch = BATCH / launches_per_step
        flops, nbytes = stage_work.get(dom, (0.0, 0.0))
        per_kernel = {}
        for k, (kms, kcnt) in prof.items():
            kf, kb = stage_work.get(k, (0.0, 0.0))
            t_s = kms / kcnt * 1e-3
            n_img = BATCH / (kcnt / prof_steps)
            per_kernel[k] = {"ms": round(kms / kcnt, 5), "tflops": round(kf * n_img / t_s / 1e12, 2),
                             "gbs": round(kb * n_img / t_s / 1e9, 1)}
        int8 = PRECISION == "int8"
        tensor_peak = peaks["bf16_tflops_sustained"] * (2.0 if int8 else 1.0)
        peak_note = ("int8 peak = 2 x the measured sustained bf16 figure: kind::i8 issues K = 32 per instruction at the "
                     "cycle count of a K = 16 16-bit MMA (tools/mma_rate.cu, profiles/r02_mma_rate.log)") if int8 else "measured sustained bf16"
        if dom in GEMM_STAGES:
            achieved = flops * imgs_per_launch / (per_launch_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TOP/s" if int8 else "TFLOP/s",
                    "peak_note": peak_note}
            peak = tensor_peak
        else:
            achieved = nbytes * imgs_per_launch / (per_launch_ms * 1e-3) / 1e9
            peak = peaks["hbm_gbs"]
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_dram_bytes_per_launch.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get("configs", {}).get(f"{ARCH}/{PRECISION}/{int(imgs_per_launch)}")
            if tj and dom in tj.get("stages", {}):
                traffic = tj["stages"][dom]        # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full)
                roof["traffic_source"] = tj.get("source")
                if dom in tj.get("tensor_pipe_active_pct", {}):
                    roof["tensor_pipe_active_pct_ncu"] = tj["tensor_pipe_active_pct"][dom]
        isf = issued_flops(ARCH, PRECISION)
        if PRECISION != "fp32" and dom in isf:
            # what the tensor pipe really executes for this kernel (the 25-tap dense form) against the same peak
            issued = isf[dom] * imgs_per_launch / (per_launch_ms * 1e-3) / 1e12
            roof.update({"issued_tflops": issued, "issued_frac": issued / tensor_peak,
                         "note": "achieved/frac count the reference's MACs (depthwise + 1x1 of the kept region); the kernel "
                                 "issues the 25-tap dense form, 8-10x more MMA work, which is what issued_* measures"})
        roof.update({"frac": achieved / peak, "traffic": traffic, "kernel": dom, "peak_source": peaks["source"],
                     "launch_ms": per_launch_ms, "images_per_launch": imgs_per_launch, "share_of_step": dms / total_stage_ms,
                     "fused_transform_conv1_block1": fused,
                     "stage_ms_per_step": {k: round(v[0] / prof_steps, 5) for k, v in prof.items()},
                     "per_kernel": per_kernel})
        launches = model.launches_per_forward(BATCH, True) * args.steps
        # the CPU baseline is an N=1 figure (rank 0 owns all host cores there; under torchrun the ranks share them)
        cpu_info = cpu_reference_run(6, 1, 32, ARCH)[2] if world == 1 else None
        line = {
            "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded_total else "weak",
            "vs_baseline": None, "dtype": PRECISION, "data": "synthetic",
            "config": common_config(ARCH, PRECISION, BATCH, world, sharded_total),
            "timing": {"how": "CUDA events on the launch stream behind a 2 ms spin kernel (the host enqueues ahead), barrier + "
                              "synchronize on both sides, max over ranks",
                       "input_sets": n_sets, "input_bytes_per_rank": n_sets * BATCH * 172800, "host_cores_per_rank": ncores},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": copy_bytes,
                    "d2h_bytes_per_step": BATCH * 5 * 4, "steps": e2e_steps,
                    "api": "Squeeze_ErNET.classify_host_submit / .result() -> ernet_classify_frames_host_submit / _wait, one "
                           "call in flight (pinned host buffers; only the frame rows the crop window reads are copied)",
                    "blocking_calls_value": e2e_blocking,
                    "h2d_ceiling_gbs": h2d_gbs_per_rank * world, "h2d_ceiling_gbs_per_rank": h2d_gbs_per_rank,
                    "h2d_ceiling_images_per_s": h2d_ceiling_value, "frac_of_h2d_ceiling": e2e_value / h2d_ceiling_value,
                    "h2d_ceiling_how": f"{ncopy} pinned cudaMemcpyAsync copies of the same {copy_bytes} bytes per rank, all "
                                       f"{world} rank(s) at once, no kernels, wall clock, max over ranks"},
            "gather_probabilities_ms": gather_ms,
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": parity,
            "roofline": roof,
            "model_flops_frac_of_tensor_peak": value / world * 2 * MACS_PER_IMAGE[ARCH] / 1e12 / tensor_peak,
            "latency_b1_ms": latency_b1_ms,
            "latency_b1_graph_ms": latency_b1_graph_ms,
            "depthwise_hbm": depthwise,
            "cpu_baseline": cpu_info,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def parity_check(model, dev_frames, host_frames, arch, precision, sd, max_frames=256):
    """Logits of one timed-configuration batch against the reference graph in fp64 on the CPU (oracle/ernet_torch on the
    bit-exact transformed frames): relative logit error and the NUMBER of top-1 disagreements (north_star: identical)."""
    from oracle import ernet_torch as T
    from oracle import ingest_numpy as I
    n = min(max_frames, dev_frames.shape[0])
    lg = model.forward_frames(dev_frames[:n], return_logits=True)[1].double().cpu().numpy()
    x = torch.from_numpy(I.ingest(host_frames[:n].numpy())).double()
    ref = T.forward(T.to_torch_sd(sd, torch.float64), x, arch)[1].numpy()
    flip = lg.argmax(1) != ref.argmax(1)
    srt = np.sort(ref, axis=1)
    margin = (srt[:, -1] - srt[:, -2]) / np.abs(ref).max()
    tol = {"fp32": 1e-4, "bf16": 2e-2, "fp16": 2e-2}.get(precision)
    return {"rel_err": float(np.abs(lg - ref).max() / np.abs(ref).max()), "flips": int(flip.sum()), "n": int(n),
            "tolerance": tol, "flip_ref_margins": [float(v) for v in np.round(margin[flip], 6)[:16]],
            "against": "reference graph in fp64 on the CPU (oracle/ernet_torch.py), uniform-noise frames, shipped weights"}


def depthwise_roofline(dev, peaks, batch=256, iters=20):
    """The standalone ACFF depthwise trio (model/acff.py:25-30,46; the fp32 engine's kernel - the 16-bit engines fold it
    into the tensor-core blocks) through ernet_acff_depthwise on the Squeeze-ErNet block-1 / block-2 shapes: algorithmic
    bytes = input once + 3x output once, CUDA events on the launching stream, L2 flushed before every launch."""
    import torch
    from rtdm_b200 import _lib
    lib = _lib.load()
    flush = torch.empty(160 << 20, dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev)
    out = {"peak": peaks["hbm_gbs"], "unit": "GB/s", "dtype": "fp32", "batch": batch, "l2": "640 MB written before every launch"}
    for name, (H, C, oh) in {"block1": (69, 16, 67), "block2": (33, 64, 31)}.items():
        x = torch.randn(batch, H, H, C, device=dev)
        w = torch.randn(3, 9, C, device=dev) * 0.3
        b = torch.randn(3, C, device=dev) * 0.1
        o = torch.empty(batch, oh, oh, 3 * C, device=dev)
        nbytes = (x.numel() + o.numel()) * 4
        best, tot = 1e9, 0.0
        for it in range(iters + 2):
            flush.fill_(float(it))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            _lib.check(lib.ernet_acff_depthwise(x.data_ptr(), 0, batch, H, H, C, oh, oh, w.data_ptr(), b.data_ptr(),
                                                o.data_ptr(), s.cuda_stream))
            e1.record(s)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            if it >= 2:
                tot += ms
                best = min(best, ms)
        gbs = nbytes / (tot / iters * 1e-3) / 1e9
        out[name] = {"shape": f"({batch},{H},{H},{C}) -> ({batch},{oh},{oh},{3 * C})", "bytes": nbytes,
                     "us": round(tot / iters * 1e3, 2), "achieved": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 4)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--schedule", default="pairs", choices=["pairs", "persistent", "per-image"],
                    help="tensor-core block kernel schedule (A/B switch; same folded weights)")
    ap.add_argument("--arch", default=ARCH, choices=["squeeze-ernet", "squeeze-redconv"],
                    help="default = BASELINE.json configs[1]; the other configs are parity cases, measured on request")
    ap.add_argument("--precision", default=PRECISION, choices=["fp32", "bf16", "fp16", "int8"])
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU per step (default: 256 at N=1 = configs[1]; 8192/N at N>1 = configs[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_main(args, rank, max(world, args.gpus))
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--arch", args.arch, "--precision", args.precision, "--schedule", args.schedule]
        if args.batch is not None:
            cmd += ["--batch", str(args.batch)]
        return subprocess.call(cmd)
    return gpu_main(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Headline benchmark: images/sec of the Squeeze-ErNet hot path on synthetic 240x240 RGB frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch: uint8 (B,240,240,3) frames -> eval transform
(Resize 159 / CenterCrop 140 / Normalize) -> Squeeze_ErNET forward -> (B,5) probabilities.
Workload = BASELINE.json configs[1]: Squeeze-ErNet, bf16, batch 256 per GPU (weak scaling: every rank
owns its own 256-frame batches and weight copy; there is no collective on the hot path).

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
BATCH = 256                     # per GPU
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
ISSUED_FLOPS = {
    "squeeze-ernet": {"tc_block1": 45 * 25 * 1 * 2 * 128 * 64 * 16, "tc_block2": 8 * 25 * 4 * 2 * 128 * 96 * 16,
                      "tc_block3": 2 * 25 * 6 * 2 * 128 * 128 * 16, "tc_block4": 24 * 2 * 128 * 256 * 16 // 2},
    "squeeze-redconv": {"tc_block1": 45 * 25 * 1 * 2 * 128 * 64 * 16, "tc_block2": 8 * 25 * 4 * 2 * 128 * 96 * 16,
                        "tc_block3": 2 * 25 * 3 * 2 * 128 * 128 * 16, "tc_block4": 12 * 2 * 128 * 256 * 16 // 2},
}


def workload_name(arch, precision, batch):
    model = {"squeeze-ernet": "Squeeze-ErNet", "squeeze-redconv": "Squeeze-ErNet-RedConv"}[arch]
    cfg = {("squeeze-ernet", "bf16", 256): "BASELINE.json configs[1]", ("squeeze-redconv", "fp16", 1024): "BASELINE.json configs[2]",
           ("squeeze-ernet", "int8", 4096): "BASELINE.json configs[3]"}.get((arch, precision, batch), "non-headline configuration")
    return (f"{model} {precision} batch {batch} per GPU on 1xB200 ({cfg}): 240x240x3 uint8 frames -> eval transform -> "
            "forward -> probabilities")


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
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
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
    for _ in range(max(1, warmup)):
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


def reference_main(args, rank):
    if rank != 0:
        return 0
    sample = 32
    steps = max(1, min(args.steps, 40))
    fps, dt, info = cpu_reference_run(steps, min(args.warmup, 3), sample)
    line = {
        "impl": "reference", "metric": "images/sec", "value": fps, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(ARCH, PRECISION, BATCH),
                   "arch": ARCH, "reference_arm": f"CPU fp32, bounded sample of {sample} of those frames per step"},
        "cpu_baseline": info,
        "e2e": {"value": fps, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------- GPU arm
def gpu_main(args, rank, local_rank, world):
    import torch.distributed as dist
    import fixtures
    import rtdm_b200

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ARCH, PRECISION, BATCH = args.arch, args.precision, args.batch
    sd = fixtures.get_state_dict(ARCH, "shipped")
    model = rtdm_b200.from_state_dict(ARCH, sd, dev, PRECISION)
    if PRECISION == "int8":
        model.calibrate()
    if PRECISION in ("fp32", "int8") or args.schedule != "pairs":
        pass                                 # fp32: CUDA-core engine; int8: one image per CTA (schedule switch ignored)
    model.prepare_ingest(*FRAME)
    model.set_persistent({"pairs": 2, "persistent": 1, "per-image": 0}[args.schedule])

    # synthetic inputs: N_INPUT_SETS distinct batches per rank (seeded by rank), rotated so that a step
    # never finds its frames in L2
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_sets = [torch.randint(0, 256, (BATCH, *FRAME, 3), dtype=torch.uint8, generator=g).pin_memory()
                 for _ in range(N_INPUT_SETS)]
    dev_sets = [h.to(dev) for h in host_sets]
    stream = torch.cuda.current_stream(dev)

    def step(i):
        return model.forward_frames(dev_sets[i % N_INPUT_SETS])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                       # sampled from the warm-up on, so short runs still see the GPU under load
    for i in range(args.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        out = step(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * BATCH * args.steps / (ms_max * 1e-3)
    assert torch.isfinite(out).all()

    # ---- e2e: host frames in, host probabilities out, through the C ABI host entry point.  Streaming use of the public
    # API: one call stays in flight while the next is submitted (classify_host_submit / .result()), so the copy of step
    # i+1 runs under the kernels of step i; every step's H2D copy and D2H read-back are inside the timed region.
    for i in range(min(args.warmup, 3)):
        model.classify_host(host_sets[i % N_INPUT_SETS])
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 50))
    pending = None
    for i in range(e2e_steps):
        nxt = model.classify_host_submit(host_sets[i % N_INPUT_SETS])
        if pending is not None:
            ph = pending.result()
        pending = nxt
    ph = pending.result()
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / (float(t.item()) * 1e-3)
    assert np.isfinite(ph).all()
    # the same, one blocking call per step (no overlap across steps)
    t0 = time.perf_counter()
    for i in range(min(e2e_steps, 20)):
        model.classify_host(host_sets[i % N_INPUT_SETS])
    e2e_blocking = world * BATCH * min(e2e_steps, 20) / (time.perf_counter() - t0)

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

    if rank == 0:
        peaks = load_peaks()
        depthwise = depthwise_roofline(dev, peaks)
        total_stage_ms = sum(v[0] for v in prof.values())
        dom = max(prof, key=lambda k: prof[k][0])
        dms, dcount = prof[dom]
        per_launch_ms = dms / dcount
        flops, nbytes = STAGE_WORK.get(dom, (0.0, 0.0))
        per_kernel = {}
        for k, (kms, kcnt) in prof.items():
            kf, kb = STAGE_WORK.get(k, (0.0, 0.0))
            t_s = kms / kcnt * 1e-3
            per_kernel[k] = {"ms": round(kms / kcnt, 5), "tflops": round(kf * BATCH / t_s / 1e12, 2),
                             "gbs": round(kb * BATCH / t_s / 1e9, 1)}
        if dom in GEMM_STAGES:
            achieved = flops * BATCH / (per_launch_ms * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"]
            roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s"}
        else:
            achieved = nbytes * BATCH / (per_launch_ms * 1e-3) / 1e9
            peak = peaks["hbm_gbs"]
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_dram_bytes_per_launch.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if dom in tj.get("stages", {}) and tj.get("batch") == BATCH and ARCH == "squeeze-ernet" and PRECISION == "bf16":
                traffic = tj["stages"][dom]        # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full)
                if dom in tj.get("tensor_pipe_active_pct", {}):
                    roof["tensor_pipe_active_pct_ncu"] = tj["tensor_pipe_active_pct"][dom]
        if dom in ISSUED_FLOPS.get(ARCH, {}):
            # what the tensor pipe really executes for this kernel (the 25-tap dense form) against the same peak
            issued = ISSUED_FLOPS[ARCH][dom] * BATCH / (per_launch_ms * 1e-3) / 1e12
            roof.update({"issued_tflops": issued, "issued_frac": issued / peaks["bf16_tflops_sustained"],
                         "note": "achieved/frac count the reference's MACs (depthwise + 1x1 of the kept region); the kernel "
                                 "issues the 25-tap dense form, 8-10x more MMA work, which is what issued_* measures"})
        roof.update({"frac": achieved / peak, "traffic": traffic, "kernel": dom, "peak_source": peaks["source"],
                     "launch_ms": per_launch_ms, "share_of_step": dms / total_stage_ms,
                     "stage_ms_per_step": {k: round(v[0] / prof_steps, 5) for k, v in prof.items()},
                     "per_kernel": per_kernel})
        launches = model.launches_per_forward(BATCH, True) * args.steps
        cpu_fps, _, cpu_info = cpu_reference_run(6, 1, 32, ARCH)
        line = {
            "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": PRECISION, "data": "synthetic",
            "config": {"workload": workload_name(ARCH, PRECISION, BATCH),
                       "arch": ARCH, "batch_per_gpu": BATCH, "global_batch": BATCH * world, "frame": "240x240x3 u8",
                       "weights": "shipped squeeze-ernet-state_dict",
                       "l2": f"inputs rotate over {N_INPUT_SETS} distinct batches "
                             f"({N_INPUT_SETS * BATCH * 172800 / 1e6:.0f} MB > 126 MB L2)",
                       "parallelism": f"dp{world}, no collective on the hot path"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": BATCH * model.host_copy_bytes_per_frame(*FRAME),
                    "d2h_bytes_per_step": BATCH * 5 * 4, "steps": e2e_steps,
                    "api": "Squeeze_ErNET.classify_host_submit / .result() -> ernet_classify_frames_host_submit / _wait, one "
                           "call in flight (pinned host buffers; only the frame rows the crop window reads are copied)",
                    "blocking_calls_value": e2e_blocking},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "model_flops_frac_of_tensor_peak": value / world * 2 * MACS_PER_IMAGE[ARCH] / 1e12 / peaks["bf16_tflops_sustained"],
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
    ap.add_argument("--batch", type=int, default=BATCH, help="frames per GPU per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_main(args, rank)
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd)
    return gpu_main(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())

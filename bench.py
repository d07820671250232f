#!/usr/bin/env python
"""Benchmark of the fused PSFNet + PSF-render focal-stack synthesis path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c4|c1] [--mode parity|fast|mixed|fp32]

One "step" = one pass of the hot path over one batch of synthetic RGB-D: the whole focal stack
[N,3,S,H,W] of the workload, in ONE kernel launch.  Metric: Mpix*slices/s = N*S*H*W / t / 1e6
(SURVEY.md section 8d).  Default workload c2 = BASELINE.json configs[1]: 5-slice focal stack at
512x512, rf50mm checkpoint, k = 11.  Under torchrun (N > 1) every rank renders its own stack of
the same shape (weak scaling, no data-path collective); the time is the max over ranks.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # name: (N, S, H, W, ks, description)
    "c1": (1, 1, 480, 640, 11, "0_warm_up.py: one 480x640 slice, rf50mm ckpt, k=11"),
    "c2": (1, 5, 512, 512, 11, "AiFNet config: 5-slice focal stack at 512x512, rf50mm ckpt, k=11"),
    "c3": (16, 5, 256, 256, 11, "DFVNet config: batch 16 x 5 slices at 256x256, rf50mm ckpt, k=11"),
    "c4": (1, 10, 1080, 1920, 31, "large render: 10 slices at 1920x1080, k=31, seeded random PSFNet"),
    # BASELINE configs[4] (AiF training step, batch sweep): only its focal-stack simulation half is on the path;
    # AiFNet itself is a downstream consumer (out of scope), so these time the simulation of one training batch
    "c5b8": (8, 5, 512, 512, 11, "AiF training batch: 8 x 5 slices at 512x512 (simulation half of the step), k=11"),
    "c5b64": (64, 5, 512, 512, 11, "AiF training batch: 64 x 5 slices at 512x512 (simulation half of the step), k=11"),
}
CKPT = os.path.join(ROOT, "tests", "golden", "rf50mm_PSFNet480x640_ks11.pkl")
DTYPES = {"parity": "f32 via fp16 hi/lo split (3 tcgen05 terms, f32 accumulate)",
          "mixed": "fp16 split for 3 layers then single fp16 term (f32 accumulate)",
          "econ": "fp16 hi/lo split, 3 terms (L1-L4) / 2 terms on calibrated fp16 weights (L5-L9, head), f32 accumulate",
          "fast": "f16 operands, f32 accumulate", "fp32": "f32"}


def flops_per_pixel(ks):
    return 2 * (4 * 64 + 64 * 256 + 8 * 256 * 256 + 256 * ks * ks)


def load_peaks():
    """(burst bf16 TFLOP/s, sustained bf16 TFLOP/s, HBM GB/s, source)"""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", 0.0)), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback"


def executed_mma_flops_per_pixel(ks, mode):
    """MMA FLOPs the kernel really issues per pixel*slice: layers L1..L10 on tensor cores (layer 0 runs on CUDA
    cores), head padded to a multiple of 16 columns, times the number of fp16 terms per layer."""
    head = (ks * ks + 15) // 16 * 16
    per_layer = [64 * 256] + [256 * 256] * 8 + [256 * head]
    terms = {"parity": [3] * 10, "fast": [1] * 10, "mixed": [3, 3, 3] + [1] * 7, "econ": [3] * 4 + [2] * 6,
             "fp32": [0] * 10}[mode]
    bias_slabs = 0 if mode == "fp32" else 9 * 16 * 256          # one K=16 bias-slab MMA per hidden layer L1..L9
    return 2 * (sum(t * m for t, m in zip(terms, per_layer)) + bias_slabs)


# dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` captures
# (profiles/r01_ncu_summary.md); the 15.7 MB output of c2 was still L2-resident when the capture ended.
NCU_TRAFFIC_BYTES = {("c2", "parity"): 6.61e6, ("c2", "fast"): 5.45e6}


def weights_for(ks):
    # data generators live in the product package (aadff_b200.synthetic); the oracle is imported only by the CPU arms
    from aadff_b200 import synthetic
    if ks == 11:
        return synthetic.split_state_dict(torch.load(CKPT, map_location="cpu"))
    return synthetic.seeded_psfnet_weights(ks, seed=0)


def make_inputs(name, rank):
    from aadff_b200 import synthetic
    N, S, H, W, ks, _ = WORKLOADS[name]
    img, depth_m = synthetic.synthetic_rgbd(N, H, W, seed=1234 + 17 * rank)
    foc_m = synthetic.synthetic_focus(depth_m, S)
    return img, -depth_m * 1e3, -foc_m * 1e3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


def cpu_reference_rate(name, min_seconds, max_reps=5):
    """The reference's operator sequence (oracle.render_reference_ops: 11 x linear, pad, unfold,
    mul, sum) on the host cores, on ONE slice of ONE image of the workload (the reference has no
    cross-slice or cross-image reuse, so per-slice cost is the whole cost).  For k=31 the frame is
    cropped to 270x480 (the unfold of a full 1080p frame needs 3 x 24 GB)."""
    from oracle import focal_stack_oracle as orc
    N, S, H, W, ks, _ = WORKLOADS[name]
    torch.set_num_threads(os.cpu_count() or 1)
    Ws, bs = weights_for(ks)
    img, depth, foc = make_inputs(name, 0)
    hh, ww = (270, 480) if ks > 15 else (H, W)
    img, depth, foc = img[:1, :, :hh, :ww].contiguous(), depth[:1, :, :hh, :ww].contiguous(), foc[:1, 0]
    sample = f"1 slice of 1 image, {hh}x{ww}, k={ks}, reference operator sequence (oracle port), fp32"
    with torch.no_grad():
        orc.render_reference_ops(Ws, bs, img, depth, foc, ks)            # warm-up
        times, t_all = [], time.perf_counter()
        while len(times) < max_reps and (time.perf_counter() - t_all < min_seconds or len(times) < 2):
            t0 = time.perf_counter()
            orc.render_reference_ops(Ws, bs, img, depth, foc, ks)
            times.append(time.perf_counter() - t0)
    return hh * ww / statistics.median(times) / 1e6, torch.get_num_threads(), sample, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, S, H, W, ks, desc = WORKLOADS[args.workload]
    from oracle import focal_stack_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    Ws, bs = weights_for(ks)
    img, depth, foc = make_inputs(args.workload, 0)
    hh, ww = (270, 480) if ks > 15 else (H, W)
    img, depth, foc = img[:1, :, :hh, :ww].contiguous(), depth[:1, :, :hh, :ww].contiguous(), foc[:1, 0]
    sample = f"each step = 1 slice of 1 image, {hh}x{ww}, k={ks}, reference operator sequence (oracle port), fp32"
    with torch.no_grad():
        for _ in range(args.warmup):
            orc.render_reference_ops(Ws, bs, img, depth, foc, ks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.render_reference_ops(Ws, bs, img, depth, foc, ks)
        dt = time.perf_counter() - t0
    value = hh * ww * args.steps / dt / 1e6
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": "focal-stack Mpix*slices/s (PSFNet + PSF render)", "value": value,
        "unit": "Mpix*slices/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "device": "host CPU", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpix*slices/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpix*slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import aadff_b200
    nat = aadff_b200.native

    N, S, H, W, ks, desc = WORKLOADS[args.workload]
    lens = aadff_b200.PSFNet(kernel_size=ks, sensor_res=(H, W), device=f"cuda:{local}", mode=args.mode)
    Ws, bs = weights_for(ks)
    sd = {}
    for l, (Wl, bl) in enumerate(zip(Ws, bs)):
        sd[f"net.{2 * l}.weight"], sd[f"net.{2 * l}.bias"] = Wl, bl
    lens.psfnet.load_state_dict(sd)
    img_h, dep_h, foc_h = make_inputs(args.workload, rank)
    img, dep, foc = img_h.cuda(), dep_h.cuda(), foc_h.cuda()
    units = N * S * H * W                               # pixel*slices per step per rank
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value"): CUDA events around every step, L2 flushed between
    def timed(mode, steps, warmup):
        for _ in range(warmup):
            out = lens.render_stack(img, dep, foc, mode=mode)
        barrier()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = lens.render_stack(img, dep, foc, mode=mode)
            e1.record()
            evs.append((e0, e1))
        barrier()
        return [a.elapsed_time(b) for a, b in evs], out

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = nat.lib.aadff_launch_count()
    t_wall0 = time.perf_counter()
    times, out = timed(args.mode, args.steps, args.warmup)
    t_wall1 = time.perf_counter()
    launches = nat.lib.aadff_launch_count() - launches0 - args.warmup
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    total_ms = torch.tensor([sum(times)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    value = units * world * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernel + D2H per step
    img_p, dep_p, foc_p = img_h.contiguous().pin_memory(), dep_h.reshape(N, H, W).contiguous().pin_memory(), \
        foc_h.contiguous().pin_memory()
    out_p = torch.empty(N, 3, S, H, W).pin_memory()
    h = lens.native().handle

    def host_call():
        nat.check(nat.lib.aadff_render_stack_host_f32(h, img_p.data_ptr(), dep_p.data_ptr(), foc_p.data_ptr(),
                                                      out_p.data_ptr(), N, 3, S, H, W, float(lens.d_min),
                                                      float(lens.d_max), nat.MODES[args.mode]))
    for _ in range(max(1, args.warmup)):
        host_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_call()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units * world * args.steps / float(e2e_s) / 1e6
    same = float((out_p.cuda() - out).abs().max())

    # ---- verification gather (the only collective; outside every timed region)
    checksum = out.double().sum().reshape(1)
    if world > 1:
        sums = [torch.zeros_like(checksum) for _ in range(world)]
        dist.all_gather(sums, checksum)
        checksum = torch.stack(sums).sum().reshape(1)

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        for mode in [m for m in ("econ", "fast", "mixed") if m != args.mode]:
            tt, _ = timed(mode, max(3, args.steps // 2), 2)
            extra[mode] = {"value": units / (statistics.mean(tt) * 1e-3) / 1e6, "unit": "Mpix*slices/s",
                           "dtype": DTYPES[mode]}

    if rank == 0:
        peak_tf, peak_tf_sus, peak_hbm, peak_src = load_peaks()
        ms_kernel = statistics.mean(times)
        achieved = flops_per_pixel(ks) * units / (ms_kernel * 1e-3) / 1e12
        executed = executed_mma_flops_per_pixel(ks, args.mode) * units / (ms_kernel * 1e-3) / 1e12
        line = {
            "metric": "focal-stack Mpix*slices/s (fused PSFNet+PSF render)", "value": value, "unit": "Mpix*slices/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[args.mode],
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "mode": args.mode, "per_gpu_shape": [N, 3, S, H, W],
                       "kernel_size": ks, "l2": "flushed (256 MiB fill) before every timed step",
                       "parallelism": f"replicated PSFNet, {world} independent stack(s), no collective on the path"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpix*slices/s",
                    "h2d_bytes_per_step": (img_p.numel() + dep_p.numel() + foc_p.numel()) * 4,
                    "d2h_bytes_per_step": out_p.numel() * 4, "api": "aadff_render_stack_host_f32 (pinned host buffers)",
                    "max_abs_vs_device_path": same},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": NCU_TRAFFIC_BYTES.get((args.workload, args.mode)),
                         "peak_source": f"{peak_src} bf16 burst",
                         "executed_mma_tflops": executed, "executed_frac_of_burst_peak": executed / peak_tf,
                         "executed_frac_of_sustained_peak": (executed / peak_tf_sus) if peak_tf_sus else None,
                         "algorithmic_bytes_per_launch": int(units * (12 + 16 / S)),
                         "kernel": "fused_psfnet_render_kernel", "kernel_ms": ms_kernel,
                         "algorithmic_flops_per_pixel_slice": flops_per_pixel(ks),
                         "executed_mma_terms": {"parity": 3, "fast": 1, "mixed": "3 for L1-L3, 1 after",
                                                "econ": "3 for L1-L4, 2 for L5-L9 and head (calibrated fp16 weights)", "fp32": 0}[args.mode]},
            "checksum": float(checksum),
        }
        if extra:
            line["other_modes"] = extra
        if world == 1 and not args.no_cpu:
            v, cores, sample, _ = cpu_reference_rate(args.workload, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "Mpix*slices/s", "cores": cores, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="parity", choices=["parity", "econ", "fast", "mixed", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

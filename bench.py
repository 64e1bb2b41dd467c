#!/usr/bin/env python
"""bench.py — HYPELCNN forward+backward+Adam patches/s on synthetic GRSS2013-shaped batches.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm, rank 0 only)

One "step" = one optimize_nn train step (common/common_nn_ops.py:208-240 of the reference):
training forward, loss, backward, [gradient all-reduce over NCCL], Adam — on one batch of
`--batch` synthetic patches per GPU (default 4096, BASELINE.json configs[1]).  Prints ONE
JSON line (rank 0).  See DESIGN.md §Measurement for how each field is produced.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG = {"batch_size": 4096, "drop_out_ratio": 0.70, "filter_count": 480, "learning_rate": 0.0003,
       "learning_rate_decay_factor": 0.96, "learning_rate_decay_step": 350, "lrelu_alpha": 0.18,
       "optimizer": "AdamOptimizer", "bn_decay": 0.95, "l2regularizer_scale": 0.00001,
       "spectral_hierarchy_level": 3, "spatial_hierarchy_level": 3, "degradation_coeff": 3, "use_residual": True}
WORKLOADS = {  # name -> (patch, channels, classes)
    "c2_grss2013": (7, 145, 15),
    "c3_grss2018": (11, 49, 20),          # the reference loader's shape: 48 HSI + 1 LiDAR (GRSS2018DataLoader.py:20-21,53-54)
    "c3_grss2018_51": (11, 51, 20),       # BASELINE.json configs[2] as written: 48 HSI + 3 LiDAR
    "c5_gulfport": (3, 65, 11),
}
SCENES = {"c2_grss2013": (349, 1905, 144), "c3_grss2018": (601, 2384, 48), "c5_gulfport": (325, 220, 64)}
GATHER_WORKLOADS = {  # name -> (H, W, bands, neighborhood, gather mode): S-gather of SURVEY §8d
    "gather_c2": (349, 1905, 144, 3, 0),
    "gather_c3": (601, 2384, 48, 5, 1),
}
METRIC = "HSI+LiDAR patches/sec fwd+bwd (HYPELCNN, GRSS2013 shape)"
FWD_BWD_MFLOP = {"c2_grss2013": 469.8, "c3_grss2018": 3550.3, "c3_grss2018_51": 3551.1, "c5_gulfport": 37.0}  # SURVEY §8d, per patch


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(numpy.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
def oracle_cpu_rate(workload, batch, steps, warmup=1):
    """patches/s of the CPU oracle (restatement of the reference's TF graph) doing the same
    train step on the host cores.  Returns (rate, seconds, threads)."""
    import torch
    from oracle import hypelcnn_ref as R
    P, C, classes = WORKLOADS[workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    alg = {**ALG, "batch_size": batch, "drop_out_ratio": 0.0}
    v = R.init_variables(P, C, classes, alg, seed=1234)
    rng = numpy.random.default_rng(1234)
    x = torch.tensor(rng.random((batch, P, P, C), dtype=numpy.float32))
    y = torch.tensor(rng.integers(0, classes, batch))
    opt = {}
    for i in range(warmup):
        _, v, opt, _ = R.train_step(v, opt, x, y, classes, alg, i)
    t0 = time.perf_counter()
    for i in range(steps):
        _, v, opt, _ = R.train_step(v, opt, x, y, classes, alg, warmup + i)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    rate, dt, threads = oracle_cpu_rate(args.workload, batch, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "patches/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "per_step_sample": f"{batch} patches (bounded sample of the "
                   f"{args.batch}-patch step)", "note": "CPU oracle = restatement of the reference's TF graph on "
                   "torch-CPU; TensorFlow is not installable in this image"},
        "cpu_baseline": {"value": rate, "unit": "patches/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} train steps of {batch} patches, dropout off"},
        "e2e": {"value": rate, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def profile_table(N):
    out = {}
    name = ctypes.create_string_buffer(64)
    ms, fl, by, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    i = 0
    while N.lib().hyp_profile_get(i, name, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl), ctypes.byref(by)) == 0:
        out[name.value.decode()] = {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}
        i += 1
    return out


def run_native(args):
    import torch
    import torch.distributed as dist
    from hypelcnn_b200 import _native as N
    from hypelcnn_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, C, classes = WORKLOADS[args.workload]
    B = args.batch
    alg = {**ALG, "batch_size": B}
    if args.model == "dualcnn":  # BASELINE configs[0] (a parity / plumbing case, not the headline): alg_param_dualcnn.json
        alg = {"batch_size": B, "drop_out_ratio": 0.70, "learning_rate": 0.0003, "learning_rate_decay_factor": 0.96,
               "learning_rate_decay_step": 350, "lrelu_alpha": 0.18, "filter_count": 480, "optimizer": "AdamOptimizer",
               "hs_lidar_diff": 1, "l2regularizer_scale": 0.00001}
    eng = E.PatchEngine(P, C, classes, alg, max_batch=B, precision=args.precision, model=args.model)
    eng.init_variables(1234)  # same seed on every rank == broadcast initial weights
    rng = numpy.random.default_rng(1234 + rank)
    nb = args.input_batches
    host_x = [torch.from_numpy(rng.random((B, P, P, C), dtype=numpy.float32)).pin_memory() for _ in range(nb)]
    host_y = [torch.from_numpy(rng.integers(0, classes, B).astype(numpy.uint8)).pin_memory() for _ in range(nb)]
    dev_x = [t.cuda() for t in host_x]
    dev_y = [t.cuda() for t in host_y]

    from hypelcnn_b200 import parallel
    ar = parallel.GradientAllReduce(overlap=not args.no_overlap) if world > 1 else None  # one NCCL all-reduce over the flat gradient buffer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ----
    for i in range(args.warmup):
        eng.train_step(dev_x[i % nb], dev_y[i % nb], allreduce=ar)
    barrier()
    N.lib().hyp_launch_count(1)
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = eng.train_step(dev_x[i % nb], dev_y[i % nb], allreduce=ar)
    ev1.record()
    barrier()
    clocks = sampler.result()
    launches = int(N.lib().hyp_launch_count(0))
    ms = ev0.elapsed_time(ev1)
    # a second pass of the same K steps with CUDA events around every launch (on the launching stream): per-kernel
    # durations for the roofline object and the breakdown.  Kept out of the timed region above.
    N.check(N.lib().hyp_profile_enable(1 if rank == 0 else 0))
    for i in range(args.steps):
        eng.train_step(dev_x[i % nb], dev_y[i % nb], allreduce=ar)
    barrier()
    prof = profile_table(N) if rank == 0 else {}
    N.lib().hyp_profile_enable(0)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * args.steps / (ms / 1e3)
    final_loss = loss.cpu().tolist()

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    from hypelcnn_b200.common import common_nn_ops as ops
    trainer = ops.HostBatchTrainer(eng, allreduce=ar)
    for i in range(2):
        trainer.step(host_x[i % nb], host_y[i % nb], prefetch=(host_x[(i + 1) % nb], host_y[(i + 1) % nb]))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        # every step: H2D copy of its batch from pinned host memory (issued one step ahead on a copy stream, like the
        # reference's prefetch_to_device), the train step, D2H read of the loss
        host_loss = trainer.step(host_x[i % nb], host_y[i % nb],
                                 prefetch=(host_x[(i + 1) % nb], host_y[(i + 1) % nb]))
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ems], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = t.item()
    e2e_patches = {"value": world * B * args.steps / (ems / 1e3), "unit": "patches/s",
                   "h2d_bytes_per_step": host_x[0].numel() * 4 + host_y[0].numel(), "d2h_bytes_per_step": 12,
                   "ms_per_step": ems / args.steps,
                   "api": "common_nn_ops.HostBatchTrainer.step: the host hands in B pre-cut fp32 patches per step (next "
                          "batch's H2D prefetched on a copy stream)"}
    e2e = e2e_patches
    if args.workload in SCENES and args.model == "hypelcnn" and args.e2e_input == "targets":
        # the scene is resident in HBM (what InMemoryImporter amounts to on a 180 GB device); per step the host hands in
        # the TARGET LIST (x, y, class) of the batch: H2D 12 B per patch, patch gather on the device, train step, D2H loss
        H, W, bands = SCENES[args.workload]
        mode = N.HYP_GATHER_GRSS2018 if args.workload == "c3_grss2018" else N.HYP_GATHER_SAME_RES
        srng = numpy.random.default_rng(4321)
        casi = torch.from_numpy(srng.integers(0, 16384, (H, W, bands)).astype(numpy.uint16)).cuda()
        Hl, Wl = (2 * H, 2 * W) if mode == N.HYP_GATHER_GRSS2018 else (H, W)
        lidar = torch.from_numpy((srng.random((Hl, Wl)) * 50).astype(numpy.float32)).cuda()
        host_t = [torch.from_numpy(numpy.stack([rng.integers(0, Wl, B), rng.integers(0, Hl, B), rng.integers(0, classes, B)],
                                               1).astype(numpy.int32)).pin_memory() for _ in range(nb)]
        strainer = ops.SceneBatchTrainer(eng, casi, lidar, P // 2, mode, allreduce=ar)
        for i in range(2):
            strainer.step(host_t[i % nb])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            host_loss = strainer.step(host_t[i % nb])
        e1.record()
        barrier()
        tms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([tms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tms = t.item()
        e2e = {"value": world * B * args.steps / (tms / 1e3), "unit": "patches/s",
               "h2d_bytes_per_step": host_t[0].numel() * 4, "d2h_bytes_per_step": 12, "ms_per_step": tms / args.steps,
               "api": "common_nn_ops.SceneBatchTrainer.step: scene resident in HBM, the host hands in the step's int32 "
                      "(x, y, class) target list; gather + optimize_nn step + loss read-back inside the timed region",
               "with_host_patches": e2e_patches}
        del casi, lidar

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    # dominant kernel = kernel function with the largest summed device time in the timed region
    by_kernel = {}
    for tag, r in prof.items():
        k = tag.split("/")[0]
        a = by_kernel.setdefault(k, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
        for f in a:
            a[f] += r[f]
    total_prof_ms = sum(r["ms"] for r in by_kernel.values()) or 1.0
    dom = max(by_kernel, key=lambda k: by_kernel[k]["ms"]) if by_kernel else None
    roofline = None
    if dom:
        d = by_kernel[dom]
        if d["flops"] > 0:
            achieved = d["flops"] / (d["ms"] / 1e3) / 1e12
            peak = peaks["bf16_tflops_sustained"]
            tc = args.precision in ("3xtf32", "3xf16", "bf16")
            # tensor-pipe work per useful MAC relative to one bf16 MMA: 3 tf32 MMAs at half rate / 3 f16 MMAs / 1
            pipe_factor = {"3xtf32": 6, "3xf16": 3, "bf16": 1}.get(args.precision)
            traffic, traffic_src = None, None
            # ncu dram bytes per GEMM launch of the same workload, from the committed launch list of that precision
            tp = os.path.join(ROOT, "profiles", {"3xtf32": "r01_gemm_traffic.json", "3xf16": "r02_gemm_traffic.json"}.get(args.precision, "none"))
            if args.workload == "c2_grss2013" and B == 4096 and os.path.exists(tp):
                t = json.load(open(tp))  # dram__bytes_read.sum + dram__bytes_write.sum, averaged per GEMM launch (ncu)
                traffic, traffic_src = t["gemm_dram_bytes_per_launch"], t["source"]
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                        "traffic_source": traffic_src, "algorithmic_flops_per_launch": d["flops"] / d["launches"],
                        "kernel": dom,
                        "avg_launch_ms": d["ms"] / d["launches"], "share_of_step": d["ms"] / total_prof_ms,
                        "peak_source": f"{peaks['source']} bf16 dense sustained (kernel timed inside a long step)",
                        "note": ({"3xtf32": "achieved = useful (algorithmic) FLOPs; the kernel issues 3 kind::tf32 MMAs (half "
                                            "the bf16 rate) per useful MAC, so tensor_pipe_frac = 6 x frac is the pipe "
                                            "utilisation",
                                  "3xf16": "achieved = useful (algorithmic) FLOPs; the kernel issues 3 kind::f16 MMAs per "
                                           "useful MAC (fp16 hi/lo operand planes), so tensor_pipe_frac = 3 x frac",
                                  "bf16": "one bf16 MMA per useful MAC: the fast mode, not a parity mode"}[args.precision]
                                 if tc else "this kernel computes in fp32 FFMA, not on the tensor pipe"),
                        "tensor_pipe_frac": pipe_factor * achieved / peak if tc else None}
        else:
            achieved = d["bytes"] / (d["ms"] / 1e3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": dom,
                        "avg_launch_ms": d["ms"] / d["launches"], "share_of_step": d["ms"] / total_prof_ms,
                        "peak_source": peaks["source"]}
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.model == "hypelcnn":
        # 256-patch steps are the CPU's best operating point (2048-patch steps measured 30 % slower: cache misses)
        rate, dt, threads = oracle_cpu_rate(args.workload, args.ref_batch, 40, 2)
        cpu = {"value": rate, "unit": "patches/s", "cores": threads, "kind": "port",
               "sample": f"40 train steps of {args.ref_batch} patches on the CPU oracle ({dt:.1f} s)"}
    parity = None
    if world == 1 and args.model == "hypelcnn" and args.workload == "c2_grss2013" and not args.no_cpu_baseline:
        # what this precision mode costs in accuracy: 512 fresh patches through a fresh engine of the same mode against
        # the fp64 CPU oracle on the same variables (training-mode forward, dropout off)
        from oracle import dataset_ref as D
        from oracle import hypelcnn_ref as R
        alg0 = {**alg, "drop_out_ratio": 0.0, "batch_size": 512}
        probe = E.PatchEngine(P, C, classes, alg0, max_batch=512, precision=args.precision)
        probe.init_variables(1234)
        px = numpy.random.default_rng(99).random((512, P, P, C), dtype=numpy.float32)
        plog, _ = probe.forward(torch.from_numpy(px).cuda(), True, True, 0)
        with torch.no_grad():
            v64 = {k: torch.tensor(a, dtype=torch.float64) for k, a in probe.export_variables().items()}
            ref = R.forward(v64, torch.tensor(px, dtype=torch.float64), classes, alg0, True)["logits"]
        got = plog.cpu().double()
        rel = ((got - ref).abs() / ref.abs().clamp_min(1e-2)).max().item()
        pred = E.argmax_confusion(plog).cpu().numpy()
        want = D.argmax_lowest(ref.numpy()).astype(numpy.uint8)
        parity = {"sample": "512 patches, training-mode forward vs the fp64 CPU oracle",
                  "max_abs_logit_error": (got - ref).abs().max().item(), "max_rel_logit_error_floor_1e-2": rel,
                  "logit_scale": ref.abs().max().item(), "argmax_mismatches": int((pred != want).sum()), "rows": 512}
        del probe
    step_flops = FWD_BWD_MFLOP[args.workload] * 1e6 * B if args.model == "hypelcnn" else \
        sum(r["flops"] for r in prof.values()) / args.steps  # useful FLOPs of the step's GEMM launches
    step_tflops = step_flops * world / (ms / args.steps / 1e3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "3xtf32": "tf32x3", "3xf16": "f16x3", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": {"workload": args.workload + ("" if args.model == "hypelcnn" else "/" + args.model), "per_gpu_batch": B, "global_batch": B * world, "patch": P, "channels": C,
                   "classes": classes, "precision_mode": args.precision,
                   "parallelism": f"dp{world}: 1 NCCL all-reduce of {eng.params.numel()} fp32 grads/step" if world > 1 else "single GPU",
                   "l2": f"per-step working set {eng.workspace_bytes / 1e9:.1f} GB >> 126 MB L2; inputs rotate over "
                         f"{nb} resident batches ({nb * host_x[0].numel() * 4 / 1e6:.0f} MB)"},
        "step_tflops_useful": step_tflops, "step_frac_of_bf16_peak": step_tflops / (peaks["bf16_tflops_sustained"] * world),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
        "final_loss": final_loss, "parity_vs_fp64_oracle": parity,
        "kernel_breakdown_ms_per_step": {k: round(v["ms"] / args.steps, 4) for k, v in
                                         sorted(by_kernel.items(), key=lambda kv: -kv[1]["ms"])},
    }
    print(json.dumps(line), flush=True)
    if args.prof_out:
        with open(args.prof_out, "w") as f:
            json.dump({k: {**v, "ms_per_step": v["ms"] / args.steps} for k, v in prof.items()}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


def run_gather(args):
    """--workload gather_c2 / gather_c3: the patch gather alone (the kernel that replaces InMemoryImporter's per-pixel
    window slice).  A step = one hyp_gather_patches call over `--batch` random targets of the resident scene."""
    import torch
    import torch.distributed as dist
    from hypelcnn_b200 import _native as N
    from hypelcnn_b200 import engine as E
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, bands, nbh, mode = GATHER_WORKLOADS[args.workload]
    B = args.batch
    S = 2 * nbh + 1
    rng = numpy.random.default_rng(1234 + rank)
    casi = torch.from_numpy(rng.integers(0, 16384, (H, W, bands)).astype(numpy.uint16)).cuda()
    Hl, Wl = (2 * H, 2 * W) if mode == 1 else (H, W)
    lidar = torch.from_numpy((rng.random((Hl, Wl)) * 50).astype(numpy.float32)).cuda()
    cmin, cmax = E.scene_minmax(casi)
    lmin, lmax = E.scene_minmax(lidar.view(Hl, Wl, 1))
    lmm = torch.cat([lmin, lmax]).contiguous()
    nb = 8                                                  # rotating output buffers: 8 x B patches >> 126 MB L2 at B >= 4096
    host_t = [torch.from_numpy(numpy.stack([rng.integers(0, Wl, B), rng.integers(0, Hl, B)], 1).astype(numpy.int32)).pin_memory()
              for _ in range(nb)]
    dev_t = [t.cuda() for t in host_t]
    outs = [torch.empty((B, S, S, bands + 1), dtype=torch.float32, device="cuda") for _ in range(nb)]
    bytes_per_patch = S * S * (4 * (bands + 1) + 2 * bands + 4)     # fp32 written + uint16 HSI read + fp32 LiDAR read

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(i, targets):
        E.gather_patches(casi, lidar, nbh, targets, cmin, cmax, lmm, mode, outs[i % nb])

    for i in range(args.warmup):
        gather(i, dev_t[i % nb])
    barrier()
    N.lib().hyp_launch_count(1)
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        gather(i, dev_t[i % nb])
    ev1.record()
    barrier()
    clocks = sampler.result()
    launches = int(N.lib().hyp_launch_count(0))
    ms = ev0.elapsed_time(ev1)
    # per-launch durations (CUDA events around each launch) for the roofline object
    N.check(N.lib().hyp_profile_enable(1 if rank == 0 else 0))
    for i in range(args.steps):
        gather(i, dev_t[i % nb])
    barrier()
    prof = profile_table(N) if rank == 0 else {}
    N.lib().hyp_profile_enable(0)
    # end to end: pinned host target list -> H2D -> gather -> 4-byte completion token back on the host
    dev_in = torch.empty((B, 2), dtype=torch.int32, device="cuda")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        dev_in.copy_(host_t[i % nb], non_blocking=True)
        gather(i, dev_in)
        token = outs[i % nb].view(-1)[-1:].cpu()
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ems], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ems = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    kname = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[kname]
    achieved = d["bytes"] / (d["ms"] / 1e3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline and mode == 0:
        from oracle import dataset_ref as D                        # the checker, timed as the CPU baseline
        scene = D.SceneRef(casi.cpu().numpy().copy(), lidar.cpu().numpy()[:, :, None].copy(), nbh, True)
        n_cpu = min(B, 8192)
        pts = numpy.concatenate([host_t[0].numpy()[:n_cpu], numpy.zeros((n_cpu, 1), numpy.int32)], 1)
        t0 = time.perf_counter()
        ref, _ = D.gather_patches(scene, pts)
        dt = time.perf_counter() - t0
        gather(0, dev_t[0])
        assert numpy.array_equal(outs[0][:n_cpu].cpu().numpy(), ref), "device gather differs from the oracle"
        cpu = {"value": n_cpu / dt, "unit": "patches/s", "cores": 1, "kind": "port",
               "sample": f"{n_cpu} window slices in numpy (oracle/dataset_ref.py, the reference's per-pixel get_data_point "
                         f"loop); pad + normalisation of the scene not timed ({dt:.2f} s)"}
    line = {
        "metric": "HSI+LiDAR patches/sec gathered from the resident scene", "value": world * B * args.steps / (ms / 1e3),
        "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16->f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": [H, W, bands], "neighborhood": nbh, "targets_per_step": B,
                   "l2": f"outputs rotate over {nb} buffers of {B * bytes_per_patch / 1e6:.0f} MB; the scene "
                         f"({casi.numel() * 2 / 1e6:.0f} MB) is larger than L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": kname,
                     "algorithmic_bytes_per_launch": d["bytes"] / d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
                     "bytes_per_patch": bytes_per_patch, "peak_source": f"{peaks['source']} copy bandwidth"},
        "cpu_baseline": cpu,
        "e2e": {"value": world * B * args.steps / (ems / 1e3), "unit": "patches/s", "h2d_bytes_per_step": B * 8,
                "d2h_bytes_per_step": 4, "ms_per_step": ems / args.steps,
                "api": "engine.gather_patches on a pinned host target list; the patches stay in HBM for the model, a "
                       "4-byte token of the result is read back"},
        "clocks": clocks, "gpu_launches": launches,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2_grss2013", choices=sorted(WORKLOADS) + sorted(GATHER_WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096, help="patches per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=256, help="patches per step of the CPU arm (bounded sample)")
    ap.add_argument("--input-batches", type=int, default=4)
    ap.add_argument("--precision", default="3xf16", choices=["fp32", "3xtf32", "3xf16", "bf16"],
                    help="3xtf32 / 3xf16: tcgen05 tensor-core engine with an fp32-accurate operand split (TF32 planes, "
                         "3 kind::tf32 MMAs per K step of 8 / fp16 hi-lo planes, 3 kind::f16 MMAs per K step of 16); "
                         "bf16: the labelled fast mode (one bf16 plane, not a parity mode); fp32: FFMA engine")
    ap.add_argument("--model", default="hypelcnn", choices=["hypelcnn", "dualcnn"],
                    help="hypelcnn = the headline metric; dualcnn = BASELINE configs[0] on the same engine (no CPU arm)")
    ap.add_argument("--e2e-input", default="targets", choices=["targets", "patches"],
                    help="what the host hands in per e2e step: the (x, y, class) target list against the HBM-resident "
                         "scene (default) or pre-cut fp32 patches (reported as e2e.with_host_patches either way)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one blocking all-reduce after backward instead "
                    "of reducing the FC/decoder gradients while the conv layers are still going backward")
    ap.add_argument("--prof-out", default=None, help="write the per-tag kernel timing table here (HYP_PROF_LAYERS=1 "
                    "adds the layer scope to GEMM tags)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.workload in GATHER_WORKLOADS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the gather's CPU arm is the cpu_baseline of the native line"}))
            return
        run_gather(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()

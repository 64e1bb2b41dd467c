"""Patch-gather throughput on the S-gather inputs of SURVEY §8d: GRSS2013-shaped scene (349 x 1905 x 144 uint16 + LiDAR),
neighborhood 3, (i) 4 096 random targets, (ii) every pixel of the scene in batches; GRSS2018-shaped variant with
--grss2018.  Reports patches/s and GB/s = algorithmic bytes (28 420 B written + the same read per C2 patch) / CUDA-event
time, for the element-wise gather_kernel ("v1", HYP_GATHER_SCALAR=1) and — bit-compared against it — the vector
gather_rows_kernel ("v2", the default).  One B200:
    python scripts/bench_gather.py [--grss2018] [--reps 20]
Under ncu: ncu --set full -k regex:gather_kernel -c 4 python scripts/bench_gather.py --reps 1"""
import argparse
import json
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypelcnn_b200 import _native as N  # noqa: E402
from hypelcnn_b200 import engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grss2018", action="store_true")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only", choices=["v1", "v2"], default=None, help="run one kernel version only (for ncu captures)")
ap.add_argument("--cpu-patches", type=int, default=4096, help="patches of the CPU arm (0: skip): the reference's per-pixel "
                "window slice restated in numpy (oracle/dataset_ref.py), scene preparation not timed")
ap.add_argument("--scene-batch", type=int, default=65536, help="targets per launch of the whole-scene sweep")
args = ap.parse_args()

rng = numpy.random.default_rng(1234)
if args.grss2018:
    Hc, Wc, C, Hl, Wl, nb, mode = 601, 2384, 48, 1202, 4768, 5, N.HYP_GATHER_GRSS2018
else:
    Hc, Wc, C, Hl, Wl, nb, mode = 349, 1905, 144, 349, 1905, 3, N.HYP_GATHER_SAME_RES
casi = torch.from_numpy(rng.integers(0, 16384, (Hc, Wc, C)).astype(numpy.uint16)).cuda()
lidar = torch.from_numpy((rng.random((Hl, Wl)) * 50).astype(numpy.float32)).cuda()
cmin, cmax = E.scene_minmax(casi)
lmin, lmax = E.scene_minmax(lidar.view(Hl, Wl, 1))
lmm = torch.cat([lmin, lmax]).contiguous()
S = 2 * nb + 1
bytes_per_patch = S * S * (4 * (C + 1) + 2 * C + 4)          # written fp32 + read uint16 casi + read fp32 LiDAR

random_targets = torch.from_numpy(numpy.stack([rng.integers(0, Wl, 4096), rng.integers(0, Hl, 4096)], 1).astype(numpy.int32)).cuda()
ys, xs = numpy.divmod(numpy.arange(Hl * Wl), Wl)
scene_targets = torch.from_numpy(numpy.stack([xs, ys], 1).astype(numpy.int32)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2


def timed(targets, out, reps):
    times = []
    for _ in range(reps):
        flush.zero_()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for lo in range(0, targets.shape[0], args.scene_batch):
            part = targets[lo:lo + args.scene_batch]
            E.gather_patches(casi, lidar, nb, part, cmin, cmax, lmm, mode, out[:part.shape[0]])
        stop.record()
        torch.cuda.synchronize()
        times.append(start.elapsed_time(stop))
    return float(numpy.median(times))


results = {}
out = torch.empty((min(args.scene_batch, scene_targets.shape[0]), S, S, C + 1), dtype=torch.float32, device="cuda")
reference_out = None
for version in {"v1": ("0",), "v2": ("1",), None: ("0", "1")}[args.only]:
    os.environ["HYP_GATHER_SCALAR"] = "1" if version == "0" else "0"
    got = E.gather_patches(casi, lidar, nb, random_targets, cmin, cmax, lmm, mode).clone()
    if reference_out is None:
        reference_out = got
    else:
        assert torch.equal(got, reference_out), "gather_rows_kernel differs from gather_kernel"
    timed(random_targets, out, 3)                                     # warm-up
    for name, targets, reps in (("random_4096", random_targets, args.reps),
                                ("whole_scene", scene_targets, max(1, args.reps // 10))):
        ms = timed(targets, out, reps)
        n = targets.shape[0]
        results[f"{'v2' if version == '1' else 'v1'}_{name}"] = {
            "ms": ms, "patches_per_s": n / ms * 1e3, "GB_per_s": n * bytes_per_patch / ms / 1e6,
            "launches": -(-n // args.scene_batch)}
os.environ["HYP_GATHER_SCALAR"] = "0"
cpu = None
if args.cpu_patches and not args.grss2018 and args.only is None:
    import time
    from oracle import dataset_ref as D                               # the checker, timed as the CPU baseline
    scene = D.SceneRef(casi.cpu().numpy().copy(), lidar.cpu().numpy()[:, :, None].copy(), nb, True)
    points = numpy.concatenate([random_targets.cpu().numpy()[:args.cpu_patches], numpy.zeros((min(4096, args.cpu_patches), 1), numpy.int32)], 1)
    t0 = time.perf_counter()
    host_patches, _ = D.gather_patches(scene, points)
    seconds = time.perf_counter() - t0
    assert numpy.array_equal(host_patches, reference_out[:len(points)].cpu().numpy()), "device gather differs from the oracle"
    cpu = {"patches": len(points), "patches_per_s": len(points) / seconds, "kind": "port", "cores": 1,
           "note": "per-pixel window slice in numpy; pad + normalisation of the scene not timed"}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
print(json.dumps({"workload": "grss2018_gather" if args.grss2018 else "grss2013_gather", "bytes_per_patch": bytes_per_patch,
                  "hbm_peak_GB_per_s": peaks.get("hbm_gbs"), "v2_bit_identical": args.only is None, "results": results,
                  "cpu_baseline": cpu}, indent=1))

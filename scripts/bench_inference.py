"""Whole-scene classification throughput: every pixel of a GRSS2013-shaped synthetic scene
(349 x 1905 = 664 845 pixels, 7x7x145 patches) through gather -> HYPELCNN eval forward -> argmax -> scatter into the
uint8 class image, all on one B200 (the reference: one Python generator call + feed per pixel).  Prints pixels/s.
    python scripts/bench_inference.py [--batch 8192] [--rows 349]
Under torchrun every rank classifies a contiguous slice of the pixel list (no exchange on the data path; the slices meet
in one MIN all-reduce of the class image after the timed region); the time is the maximum over ranks."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypelcnn_b200.classify.infer_for_classification import create_all_scene_data  # noqa: E402
from hypelcnn_b200 import parallel  # noqa: E402
from hypelcnn_b200.common import common_nn_ops as ops  # noqa: E402
from hypelcnn_b200.importer.GeneratorImporter import GeneratorDataInfo, LazyPatchDataset  # noqa: E402
from hypelcnn_b200.loader.SyntheticGRSS2013DataLoader import SyntheticGRSS2013DataLoader  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--rows", type=int, default=349, help="scene rows (349 = the whole GRSS2013 scene)")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

rank, local, world = parallel.init_from_env()
torch.cuda.set_device(local)
alg = json.load(open(os.path.join(ROOT, "tests", "golden", "alg_param_hypelcnn.json"))) if os.path.exists(
    os.path.join(ROOT, "tests", "golden", "alg_param_hypelcnn.json")) else {
    "batch_size": 48, "drop_out_ratio": 0.70, "filter_count": 480, "learning_rate": 0.0003,
    "learning_rate_decay_factor": 0.96, "learning_rate_decay_step": 350, "lrelu_alpha": 0.18, "optimizer": "AdamOptimizer",
    "bn_decay": 0.95, "l2regularizer_scale": 0.00001, "spectral_hierarchy_level": 3, "spatial_hierarchy_level": 3,
    "degradation_coeff": 3, "use_residual": True}
alg["batch_size"] = args.batch
loader = SyntheticGRSS2013DataLoader(f"synthetic:H={args.rows},W=1905,samples=64")
data_set = loader.load_data(3, True)
scene_shape = data_set.get_scene_shape()
scene = create_all_scene_data(scene_shape, GeneratorDataInfo(None, None, loader, data_set))
if world > 1:
    scene = scene._replace(targets=parallel.shard_targets(scene.targets, rank, world))
model = ops.get_model_from_name("HYPELCNNModel")
classes = loader.get_class_count().stop


def predict(images):
    return model.create_tensor_graph(ops.ModelInputParams(images, None, "/gpu:0", False), classes, alg).y_conv


iterator = ops.simple_nn_iterator(LazyPatchDataset(data_set, scene.targets, classes), args.batch)
nn_params = ops.NNParams(input_iterator=iterator, data_with_labels=scene, metrics=None, predict_tensor=predict)
class_map = torch.full(tuple(scene_shape), 255, dtype=torch.uint8, device="cuda")
ops.perform_prediction(None, nn_params, class_map)                       # warm-up (builds the engine)
times = []
for _ in range(args.reps):
    class_map.fill_(255)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall = time.perf_counter()
    start.record()
    ops.perform_prediction(None, nn_params, class_map)
    stop.record()
    torch.cuda.synchronize()
    times.append((start.elapsed_time(stop), (time.perf_counter() - wall) * 1e3))
device_ms, wall_ms = min(times)
device_ms = parallel.max_over_ranks(device_ms, device="cuda")
parallel.merge_class_map(class_map)
pixels = scene_shape[0] * scene_shape[1]
assert int((class_map == 255).sum()) == 0
if rank == 0:
    digest = int(torch.sum(class_map.to(torch.int64) * (torch.arange(class_map.numel(), device="cuda").view_as(class_map) % 8191 + 1)).item())
    print(json.dumps({"metric": "whole-scene classification, pixels/s (HYPELCNN eval, GRSS2013 shape)", "pixels": pixels,
                      "n_gpus": world, "batch": args.batch, "device_ms": device_ms, "wall_ms": wall_ms,
                      "pixels_per_s": pixels / device_ms * 1e3, "precision": model.precision,
                      "class_map_digest": digest,
                      "useful_TFLOP_per_s": pixels * 151.28e6 / device_ms / 1e9}))   # eval graph: 157.16 - 5.88 MFLOP of decoder
if world > 1:
    torch.distributed.destroy_process_group()

"""GAN train-iteration throughput on synthetic GULFPORT-shaped 64-band spectra (SURVEY S-C4): CycleGAN (BASELINE
configs[3]; one iteration = global_step += 1, one generator step, one discriminator step) or, with --gan_type, any
registry entry (dcl_gan, the augmenter of configs[4]: two CUT models x three train ops).  One JSON line per batch size.
Launch with torchrun for N > 1 (data parallel: one all-reduce per optimizer step over the 478 / 20 800 gradients)."""
import argparse
import json
import os
import sys

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypelcnn_b200 import parallel  # noqa: E402
from hypelcnn_b200.gan.wrapper_registry import get_wrapper  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="32,16384")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--gan_type", default="cycle_gan")
a = ap.parse_args()
rank, local, world = parallel.init_from_env()
torch.cuda.set_device(local)
rng = numpy.random.default_rng(1234 + rank)
for B in [int(b) for b in a.batches.split(",")]:
    y = rng.uniform(0.02, 0.5, (B, 1, 1, 64)).astype(numpy.float32)
    x = (y * numpy.linspace(1.5, 4, 64)).astype(numpy.float32)
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    flags = argparse.Namespace(cycle_consistency_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True,
                               batch_size=B)
    w = get_wrapper(a.gan_type, flags)
    model = w.define_model(xd, yd)
    ops = w.define_train_ops(model, w.define_loss(model), 100000, generator_lr=2e-4, discriminator_lr=1e-4,
                             gen_discriminator_lr=1e-4)
    if world > 1:
        for t in ([w.trainer.model_x2y, w.trainer.model_y2x] if hasattr(w.trainer, "model_x2y") else [w.trainer]):
            t.allreduce = parallel.GradientAllReduce()
    for _ in range(a.warmup):
        ops.train_iteration(xd, yd)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        losses = ops.train_iteration(xd, yd)
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    if rank == 0:
        print(json.dumps({"metric": f"{a.gan_type} (x,y) pairs/s, every train op once per iteration",
                          "value": world * B * a.steps / (ms / 1e3), "unit": "pairs/s", "n_gpus": world,
                          "per_gpu_batch": B, "ms_per_iteration": ms / a.steps, "dtype": "f32",
                          "train_ops": len(losses), "losses": [l.cpu().tolist() for l in losses]}), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()

"""Run a few HYPELCNN train steps at the bench shape (for ncu captures: keep it short)."""
import argparse
import os
import sys

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypelcnn_b200 import engine as E  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--workload", default="c2_grss2013")
ap.add_argument("--precision", default="3xf16")
a = ap.parse_args()
P, C, classes = bench.WORKLOADS[a.workload]
alg = {**bench.ALG, "batch_size": a.batch}
eng = E.PatchEngine(P, C, classes, alg, max_batch=a.batch, precision=a.precision)
eng.init_variables(1234)
rng = numpy.random.default_rng(1234)
x = torch.from_numpy(rng.random((a.batch, P, P, C), dtype=numpy.float32)).cuda()
y = torch.from_numpy(rng.integers(0, classes, a.batch).astype(numpy.uint8)).cuda()
for i in range(a.steps):
    loss = eng.train_step(x, y)
torch.cuda.synchronize()
print("loss", loss.cpu().tolist())

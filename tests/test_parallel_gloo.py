"""world_size-2 gloo tests (CPU) of the data-parallel host logic in hypelcnn_b200/parallel.py: the
sharding of patches over ranks and the one-all-reduce-per-step contract."""
import os
import socket

import numpy
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from hypelcnn_b200 import parallel as P
    r, local, w = P.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # ---- a linear least-squares "model": grad of the mean loss over the rank's shard of a global batch ----
    rng = numpy.random.default_rng(7)
    X = torch.tensor(rng.standard_normal((10, 4)))
    y = torch.tensor(rng.standard_normal(10))
    wgt = torch.tensor(rng.standard_normal(4))
    P.broadcast_parameters([wgt])
    b, e = P.shard_range(10, rank, world)
    Xs, ys = X[b:e], y[b:e]
    grad_local = 2.0 * Xs.T @ (Xs @ wgt - ys) / (e - b)  # mean over the local shard, like mean_B(loss)
    ar = P.GradientAllReduce()
    g = grad_local.clone()
    scale = ar(g)
    assert ar.calls == 1 and abs(scale - 1.0 / world) < 1e-15
    # equal shard sizes -> averaged shard gradients == gradient of the global-batch mean
    grad_global = 2.0 * X.T @ (X @ wgt - y) / 10
    assert torch.allclose(g * scale, grad_global, atol=1e-12), (g * scale, grad_global)
    # ---- metrics: confusion matrices add up ----
    conf = torch.zeros((3, 3), dtype=torch.int32)
    conf[rank, rank] = rank + 1
    P.reduce_confusion(conf)
    assert conf[0, 0].item() == 1 and conf[1, 1].item() == 2 and conf.sum().item() == 3
    # ---- timing is the max over ranks ----
    assert P.max_over_ranks(10.0 + rank) == 10.0 + world - 1
    # ---- scene targets partition without overlap ----
    targets = numpy.arange(23 * 3).reshape(23, 3)
    mine = P.shard_targets(targets, rank, world)
    numpy.save(os.path.join(out_dir, f"targets_{rank}.npy"), mine)
    # ---- whole-scene inference: each rank fills its slice of the class image, the slices meet in one MIN all-reduce ----
    class_map = torch.full((23,), 255, dtype=torch.uint8)
    class_map[torch.from_numpy(mine[:, 0] // 3)] = torch.from_numpy((mine[:, 2] % 7).astype(numpy.uint8))
    P.merge_class_map(class_map)
    assert class_map.tolist() == [int(v % 7) for v in targets[:, 2]]
    # ---- the training split is strided over ranks ----
    from collections import namedtuple
    Target = namedtuple("Target", ["data", "labels"])
    share = P.shard_training_data(Target(torch.arange(9), torch.arange(9) % 2), rank, world)
    assert share.data.tolist() == list(range(rank, 8, world))            # 9 rows over 2 ranks: 4 each, the last dropped
    # ---- the loaders' random splits are drawn from one broadcast seed: every rank holds the SAME train / validation
    #      split, so rank-strided training shards never contain another rank's validation pixels ----
    numpy.random.seed(1000 + rank)                                        # what separately started processes look like
    seed = P.sync_split_seed()
    assert isinstance(seed, int)
    from hypelcnn_b200.common.sample_ops import shuffle_training_data_using_ratio, shuffle_training_data_using_size
    samples = numpy.stack([numpy.arange(400), numpy.arange(400)[::-1], numpy.arange(400) % 5], axis=1)
    train_a, val_a = shuffle_training_data_using_ratio(samples, 0.25)
    train_b, val_b = shuffle_training_data_using_size(range(5), samples, 20, None)
    Targets = namedtuple("Targets", ["targets"])
    for tag, train, val in (("ratio", train_a, val_a), ("size", train_b, val_b)):
        mine = P.shard_training_data(Targets(train), rank, world).targets
        numpy.save(os.path.join(out_dir, f"split_{tag}_train_{rank}.npy"), mine)
        numpy.save(os.path.join(out_dir, f"split_{tag}_val_{rank}.npy"), val)
    # ---- stop decisions are collective ----
    assert P.collective_any(rank == 1) is True and P.collective_any(False) is False
    # ---- the end of a run: a rank that is done early waits for its peers before the group goes away ----
    import time
    if rank == 1:
        time.sleep(0.5)
    P.finish()
    assert not dist.is_initialized()
    P.finish()  # idempotent


def test_two_rank_gradient_allreduce_and_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [numpy.load(tmp_path / f"targets_{r}.npy") for r in range(world)]
    assert [len(p) for p in parts] == [12, 11]
    assert numpy.array_equal(numpy.concatenate(parts), numpy.arange(23 * 3).reshape(23, 3))
    for tag in ("ratio", "size"):
        vals = [numpy.load(tmp_path / f"split_{tag}_val_{r}.npy") for r in range(world)]
        trains = [numpy.load(tmp_path / f"split_{tag}_train_{r}.npy") for r in range(world)]
        assert numpy.array_equal(vals[0], vals[1])                                   # one validation set
        val_ids = set(vals[0][:, 0].tolist())
        ids = [set(t[:, 0].tolist()) for t in trains]
        assert not (ids[0] & ids[1]) and not ((ids[0] | ids[1]) & val_ids)           # disjoint shards, none in validation
        assert len(ids[0]) == len(ids[1]) > 0


def test_shard_range_properties():
    from hypelcnn_b200 import parallel as P
    for n in (0, 1, 7, 4096, 664845):
        for world in (1, 2, 3, 8):
            spans = [P.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1

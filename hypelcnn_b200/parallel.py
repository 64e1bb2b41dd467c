"""Data-parallel plumbing: one process per GPU, patches sharded over ranks, ONE
all-reduce over the flat gradient buffer per step and nothing else on the data path.

The reference has no multi-device path (its device id is the literal "/gpu:0",
classify/train_for_classification.py:51-55); this module is what a torchrun launch adds around
``PatchEngine.train_step``.  Everything here is backend-agnostic torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests) — no arithmetic of the hot path lives here.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Process group from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world).  world == 1 -> no group is created."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def finish():
    """End of a data-parallel run: wait for this rank's device work, meet the other ranks, drop the process group.
    All-reduces are asynchronous on the stream: a rank whose host loop has run ahead and simply returns would tear its
    communicator down while the peers still wait in the collectives it only enqueued (seen as a hang of the two-rank
    GAN app).  No-op for a single process."""
    if not dist.is_initialized():
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


def sync_split_seed(group=None):
    """Same train / validation / test split on every rank.  The loaders draw their splits from numpy's and Python's
    global generators (sklearn's StratifiedShuffleSplit without a random_state, numpy.random.shuffle), which every
    process seeds differently: strided shards of DIFFERENT permutations overlap, and one rank's validation pixels turn
    up in another rank's training shard.  Rank 0 draws one seed, broadcasts it, and every rank seeds both generators
    with it; call this right before ``read_data_set``.  Returns the seed (None for a single process)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return None
    import random

    import numpy
    box = [int.from_bytes(os.urandom(4), "little") if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    numpy.random.seed(box[0])
    random.seed(box[0])
    return box[0]


def collective_any(flag, device=None, group=None):
    """True on every rank as soon as ``flag`` is true on one of them (MAX all-reduce of one byte).  Stop decisions of
    a training loop (NaN watch, stop requests) go through this so that all ranks leave the loop at the same step and
    nobody is left waiting in a gradient all-reduce."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def shard_range(n, rank, world):
    """Contiguous partition of n units (patches of a batch, pixels of a scene) over ranks: the first
    n % world ranks get one extra unit.  -> (begin, end)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(int(n), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_targets(targets, rank, world):
    """Rank's slice of an [N, 3] (x, y, class) target list (full-scene inference: each rank writes its own
    slice of the class map, no exchange)."""
    b, e = shard_range(len(targets), rank, world)
    return targets[b:e]


def shard_training_data(data_with_labels, rank, world):
    """Rank's share of an importer's training split, strided (rank, rank + world, ...) so that every rank sees every
    class whatever the order of the list.  Works on the importers' holders: (data, labels) arrays in HBM
    (InMemoryImporter / TFRecordImporter) or a lazy target list (GeneratorImporter)."""
    if world <= 1:
        return data_with_labels
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    # equal counts on every rank (the remainder is dropped): all ranks then run the same number of steps per epoch and
    # meet in every all-reduce
    if hasattr(data_with_labels, "path"):          # TFRecordImporter: the file is strided when it is parsed into HBM
        shape = list(data_with_labels.data.shape)
        shape[0] = int(shape[0]) // world
        return data_with_labels._replace(data=type(data_with_labels.data)(shape), shard=(rank, world))
    if hasattr(data_with_labels, "labels"):
        rows = (data_with_labels.labels.shape[0] // world) * world
        return data_with_labels._replace(data=data_with_labels.data[rank:rows:world],
                                         labels=data_with_labels.labels[rank:rows:world])
    rows = (len(data_with_labels.targets) // world) * world
    return data_with_labels._replace(targets=data_with_labels.targets[rank:rows:world])


class GradientAllReduce:
    """The single collective of a train step: SUM all-reduce of the flat gradient buffer; the 1/world
    factor is folded into the Adam kernel (``hyp_adam_step(grad_scale)``), so the returned scale is what
    ``PatchEngine.train_step(allreduce=...)`` expects."""

    def __init__(self, group=None, overlap=False):
        self.group = group
        self.overlap = overlap  # PatchEngine.train_step: reduce the FC/decoder tail while the conv layers go backward
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0

    def __call__(self, grads):
        if self.world > 1:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        return 1.0 / self.world


def broadcast_parameters(tensors, src=0, group=None):
    """Same initial weights / BN state / Adam moments on every rank."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)


def reduce_confusion(confusion, group=None):
    """Sum the per-rank int32 confusion matrices (metrics are logged every N steps, off the step's critical
    path — mirrors save_summaries_steps=100, classify/monitored_session_runner.py:147,177)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(confusion, op=dist.ReduceOp.SUM, group=group)
    return confusion


def merge_class_map(class_map, fill_value=255, group=None):
    """Whole-scene inference over ranks: every rank classified its own slice of the pixel list into a class image
    pre-filled with ``fill_value`` (255 > any class id); the element-wise minimum over ranks is the complete image.
    The only exchange of the inference path, once, after the last batch."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(class_map, op=dist.ReduceOp.MIN, group=group)
    return class_map


def max_over_ranks(value, device=None, group=None):
    """Timing helper: a device-measured duration as the maximum over ranks."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

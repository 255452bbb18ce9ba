"""Multi-GPU plumbing: the path shards embarrassingly over (image, target view) pairs, one process per
GPU.  The reference's only inference-time mechanism is nn.DataParallel (demo.py:191-195: replicate + scatter +
gather every forward); here the sources are broadcast ONCE at job start and nothing else crosses GPUs."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def view_assignment(n_images, n_views, rank, world):
    """(image, view) pairs of this rank: by view when there are at least `world` views (BASELINE config 4: GPU g
    renders view g of every image), else by image (config 5)."""
    if n_views >= world:
        lo, hi = shard_range(n_views, rank, world)
        return [(i, v) for v in range(lo, hi) for i in range(n_images)]
    lo, hi = shard_range(n_images, rank, world)
    return [(i, v) for i in range(lo, hi) for v in range(n_views)]


def broadcast_sources(src, world, root=0):
    """Broadcast the source tensor from `root` (NCCL on GPUs, gloo on CPU tensors).  Returns the device time in ms
    (0 when world == 1 or on CPU)."""
    if world <= 1:
        return 0.0
    if src.is_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        dist.broadcast(src, root)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    dist.broadcast(src, root)
    return 0.0

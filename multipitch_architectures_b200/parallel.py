"""Multi-GPU sharding of patch-wise inference: one process per GPU, independent units, no data-path
collective except the final gather of the [N, 72] activations.

Every output frame depends on the 75-frame window around it only, so a recording is split into contiguous frame
ranges; each rank receives its range plus a 37-frame halo of REAL frames on interior boundaries (zero padding only at
the true ends of the recording, exactly as the reference pads the whole file once, exp126a…py:420).
For the SAUnet (attention over the batch axis) ranges must be multiples of the reference batch (50 consecutive frames)
so that every reference batch stays on one rank."""
import torch
import torch.distributed as dist

HALF = 37


def shard_range(n_items, world, rank, multiple=1):
    """Contiguous, balanced [start, end) of `n_items` for `rank`; all boundaries are multiples of `multiple`."""
    if world < 1 or not (0 <= rank < world) or multiple < 1:
        raise ValueError('bad world / rank / multiple')
    blocks = (n_items + multiple - 1) // multiple
    base, extra = divmod(blocks, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return min(n_items, b0 * multiple), min(n_items, b1 * multiple)


def local_window(n_frames, start, end):
    """Frames a rank must hold to produce outputs [start, end): -> (first, last, lo, hi) with lo/hi relative to first."""
    first, last = max(0, start - HALF), min(n_frames, end + HALF + 1)
    return first, last, start - first, end - first


def predict_sharded(predict_fn, hcqt, world=None, rank=None, multiple=1, gather=True, group=None):
    """predict_fn(hcqt_local [C, n, F], lo, hi) -> [hi-lo, P] for centre frames lo..hi-1 of the local array, with zero
    padding beyond the local array's ends.  Returns the full [N, P] on every rank (gather=True) or the local slice."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = hcqt.shape[1]
    s, e = shard_range(N, world, rank, multiple)
    if e > s:
        first, last, lo, hi = local_window(N, s, e)
        # interior boundaries: the halo frames are real; true recording ends: predict_fn's own zero padding applies.
        # A rank whose window starts after frame 0 must not see zero padding inside its halo: guaranteed because the
        # halo is HALF frames wide, the full reach of a patch.
        local = predict_fn(hcqt[:, first:last].contiguous(), lo, hi)
    else:
        local = None
    if not gather or world == 1:
        return local
    P = torch.tensor([0 if local is None else local.shape[1]], device=hcqt.device)
    dist.all_reduce(P, op=dist.ReduceOp.MAX, group=group)
    sizes = [shard_range(N, world, r, multiple) for r in range(world)]
    longest = max(b - a for a, b in sizes)
    buf = torch.zeros(longest, int(P.item()), dtype=torch.float32, device=hcqt.device)
    if local is not None:
        buf[:e - s] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:b - a] for p, (a, b) in zip(parts, sizes)], 0)

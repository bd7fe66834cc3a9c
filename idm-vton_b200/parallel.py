"""Multi-GPU plumbing for the try-on engine: one process per GPU, requests are independent units.

The reference has no multi-GPU inference. The only exchange in this design is the one-time
broadcast of the shared, read-only UNet weights from rank 0 at load (NCCL over NVLink on GPUs; gloo in the CPU tests);
after that every rank runs the single-GPU engine on its own contiguous shard of the request list — no per-step
collective exists on the path, so none is invented.
"""
import torch


def shard_requests(n_requests, world_size, rank):
    """Contiguous, balanced partition of request indices [0, n) -> the half-open range owned by `rank`."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_requests, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def broadcast_state_dict(sd, src=0, bucket_bytes=1 << 30):
    """In-place broadcast of every tensor of `sd` (same keys/shapes on all ranks) from `src`, coalesced into flat
    buckets of ~bucket_bytes so NCCL sees a few large transfers instead of ~1900 small ones."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sd
    keys = list(sd.keys())
    i = 0
    while i < len(keys):
        group, size = [], 0
        dtype, device = sd[keys[i]].dtype, sd[keys[i]].device
        while i < len(keys) and sd[keys[i]].dtype == dtype and (not group or size < bucket_bytes):
            t = sd[keys[i]]
            group.append(t)
            size += t.numel() * t.element_size()
            i += 1
        flat = torch.cat([t.reshape(-1) for t in group]) if len(group) > 1 else group[0].reshape(-1).clone()
        dist.broadcast(flat, src=src)
        off = 0
        for t in group:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    return sd


def alloc_state_dict_arena(shapes, dtype=torch.float16, device="cuda", align=128):
    """One contiguous weight arena + a state dict of views into it ({key: shape} -> ({key: tensor}, flat)). Offsets are
    aligned to `align` elements so every view satisfies TMA's 16-byte base alignment. Broadcasting the arena moves all
    weights of a UNet with a handful of NCCL calls and no staging copies (`broadcast_arena`)."""
    offs, total = {}, 0
    for k, shp in shapes.items():
        n = 1
        for d in shp:
            n *= int(d)
        offs[k] = (total, n, tuple(shp))
        total += (n + align - 1) // align * align
    flat = torch.empty(total, dtype=dtype, device=device)
    return {k: flat[o:o + n].view(shp) for k, (o, n, shp) in offs.items()}, flat


def broadcast_arena(flat, src=0, bucket_bytes=2 << 30):
    """In-place broadcast of a flat weight arena from `src` in buckets of ~bucket_bytes (NCCL over NVLink / NVSwitch on
    GPUs, gloo in the CPU tests). No staging: the views of `alloc_state_dict_arena` see the data as it lands."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    step = max(1, bucket_bytes // flat.element_size())
    for o in range(0, flat.numel(), step):
        dist.broadcast(flat[o:o + step], src=src)
    return flat

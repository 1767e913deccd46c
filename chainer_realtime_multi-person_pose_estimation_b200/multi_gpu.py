"""Image-sharded multi-GPU execution: one process per GPU (torch.distributed), weights
replicated, image i of a global batch of B goes to rank i // ceil(B / world); the only data
exchange is ONE all-gather of the fixed-size per-image result records (opb_image_header +
max_persons x opb_person, include/opb.h) -- NCCL over NVLink on GPUs, gloo in the CPU tests.
The reference has no multi-device path; this is new."""
import numpy as np

try:
    from . import _native
except ImportError:
    import _native

HEADER_BYTES = _native.HEADER_DTYPE.itemsize
PERSON_BYTES = _native.PERSON_DTYPE.itemsize


def shard_range(n_images, world, rank):
    """[start, stop) of the images rank `rank` owns; contiguous, sizes differ by at most 1 chunk."""
    per = -(-n_images // world)
    start = min(rank * per, n_images)
    return start, min(start + per, n_images)


def record_bytes(max_persons):
    return HEADER_BYTES + max_persons * PERSON_BYTES


def pack_records(headers, persons):
    """headers [n] HEADER_DTYPE, persons [n, max_persons] PERSON_DTYPE -> uint8 [n, record_bytes]."""
    n, mp = persons.shape
    out = np.empty((n, record_bytes(mp)), np.uint8)
    out[:, :HEADER_BYTES] = headers.view(np.uint8).reshape(n, HEADER_BYTES)
    out[:, HEADER_BYTES:] = persons.view(np.uint8).reshape(n, mp * PERSON_BYTES)
    return out


def unpack_records(buf, max_persons):
    buf = np.ascontiguousarray(buf, np.uint8).reshape(-1, record_bytes(max_persons))
    headers = buf[:, :HEADER_BYTES].copy().view(_native.HEADER_DTYPE).reshape(-1)
    persons = buf[:, HEADER_BYTES:].copy().view(_native.PERSON_DTYPE).reshape(-1, max_persons)
    return headers, persons


def all_gather_records(local, per_rank_images, group=None):
    """local: torch uint8 tensor [per_rank_images * record_bytes] (device for NCCL, CPU for gloo),
    padded to the same length on every rank.  Returns the gathered [world * len(local)] tensor
    (ONE collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def gathered_to_global(gathered, n_images, world, max_persons):
    """Drop the per-rank padding and return (headers [n_images], persons [n_images, max_persons])
    in global image order."""
    per = -(-n_images // world)
    rb = record_bytes(max_persons)
    g = np.asarray(gathered, np.uint8).reshape(world, per, rb)
    rows = []
    for r in range(world):
        a, b = shard_range(n_images, world, r)
        rows.append(g[r, :b - a])
    return unpack_records(np.concatenate(rows, axis=0), max_persons)

"""Host-side logic of the multi-GPU split (one process per GPU, torch.distributed for the plumbing).

The path shards by independent units: every rank holds a replica of the density / sun-transmittance grids and
renders its own contiguous range of global subframe ids (RNG stream ids), so a pixel sample is the same
bits no matter which rank draws it.  Each rank keeps Welford statistics (n_r, mean_r, M2_r) per pixel channel;
they are merged with ONE sum-reduce of the mergeable moments {n*mean, M2 + n*mean^2} (float64):
    mean = S1 / N,   M2 = S2 - N * mean^2,   N = sum n_r.
This module has no CUDA dependency: the same functions run under gloo on CPU tensors (tests) and under NCCL on
the device buffers filled by ds_frame_export_moments_device (bench.py).
"""
from __future__ import annotations


def subframe_range(rank: int, world: int, total: int) -> tuple[int, int]:
    """Contiguous split of global subframe ids 1..total: returns (stream_offset, count) of `rank`.
    Rank r renders global ids offset+1 .. offset+count."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def export_moments(mean, m2, n: int):
    """(n*mean, M2 + n*mean^2) in float64; `mean`, `m2` are torch tensors of equal shape."""
    import torch

    mean64, m264 = mean.to(torch.float64), m2.to(torch.float64)
    return torch.stack([mean64 * n, m264 + mean64 * mean64 * n])


def import_moments(moments, n_total: int):
    """Inverse of export_moments after the sum over ranks: returns (mean, M2) as float32."""
    import torch

    mu = moments[0] / n_total
    s2 = torch.clamp(moments[1] - mu * mu * n_total, min=0.0)
    return mu.to(torch.float32), s2.to(torch.float32)


def reduce_moments(moments, dst: int = 0):
    """The single collective of a multi-GPU render: sum-reduce to `dst` (NCCL on CUDA tensors, gloo on CPU)."""
    import torch.distributed as dist

    dist.reduce(moments, dst=dst, op=dist.ReduceOp.SUM)
    return moments


class SceneQueue:
    """Dynamic assignment of dataset scenes to ranks (dataset generation shards by scene, no data-path collective).

    Scene costs differ by an order of magnitude (a 1 km cloud needs far more experiments per sample than a 12 km one), so a
    static round-robin leaves ranks idle at the end.  Every rank draws the next scene id from ONE shared counter kept in the
    process group's key-value store (`store.add`, an atomic fetch-and-add served by rank 0's TCPStore -- control plane only;
    works the same under gloo and NCCL).  `scenes` may be a list of ids (e.g. the ones a resumed run still has to do).
    """

    def __init__(self, store, scenes, key: str = "ds_next_scene"):
        self.store, self.scenes, self.key = store, list(scenes), key

    def __iter__(self):
        while True:
            ticket = self.store.add(self.key, 1) - 1  # add returns the value after the increment
            if ticket >= len(self.scenes):
                return
            yield self.scenes[ticket]


def static_scenes(rank: int, world: int, scenes):
    """Round-robin split (what `datagen collect --shard r/R` does)."""
    return [s for i, s in enumerate(scenes) if i % world == rank]

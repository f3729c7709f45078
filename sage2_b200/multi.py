"""Multi-GPU overlap-graph build: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch)
for the plumbing.

Round-1 partitioning (the replicated-table variant): every rank organises the reads
and builds the table (both deterministic, so all ranks hold identical copies), rank r searches read ids
[r*chunk, (r+1)*chunk) (sage2gpu_phase_a_partition), then ONE exchange step makes the phase-A state
complete everywhere:

    all-gather        rightExtension / leftExtension records, connections>300 flags
    all-reduce(MAX)   largest id whose scan found the read contained (economyGraph.cpp:735)

after which phases B and C and the canonical edge sort run (sage2gpu_finish_graph).  The exchange
works on any torch tensors, so the world_size-2 gloo test drives it on the CPU with emulated slices.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _DevArray:
    """Zero-copy view of library-owned device memory (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_views(bufs: dict, world: int, device) -> dict:
    """torch views of the arrays sage2gpu_phase_a_buffers exposes (int64 / uint8 / int32 bit patterns)."""
    n = bufs["chunk"] * world
    if n == 0:
        z = lambda dt: torch.zeros(0, dtype=dt, device=device)
        return {"right": z(torch.int64), "left": z(torch.int64), "over_limit": z(torch.uint8), "contained_by": z(torch.int32)}
    mk = lambda key, ts: torch.as_tensor(_DevArray(bufs[key], n, ts), device=device)
    return {"right": mk("right", "<i8"), "left": mk("left", "<i8"), "over_limit": mk("over_limit", "|u1"),
            "contained_by": mk("contained_by", "<i4")}


def exchange_phase_a(views: dict, chunk: int, rank: int, world: int) -> int:
    """In place: every rank ends with the complete phase-A arrays.  Returns the bytes this rank sent."""
    if world == 1 or chunk == 0:
        return 0
    sent = 0
    for key in ("right", "left", "over_limit"):
        full = views[key]
        mine = full[rank * chunk:(rank + 1) * chunk]
        if full.is_cuda:
            dist.all_gather_into_tensor(full, mine)
        else:      # gloo
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine.clone())
            full.copy_(torch.cat(parts))
        sent += mine.numel() * mine.element_size()
    dist.all_reduce(views["contained_by"], op=dist.ReduceOp.MAX)
    sent += views["contained_by"].numel() * 4
    return sent


def build_overlap_graph(gpu, rank: int, world: int, device=None) -> int:
    """sage2gpu_build_overlap_graph over `world` GPUs (reads loaded and table built on every rank)."""
    gpu.phase_a_partition(rank, world)
    if world == 1:
        gpu.finish_graph()
        return 0
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    bufs = gpu.phase_a_buffers()
    views = device_views(bufs, world, device)
    sent = exchange_phase_a(views, bufs["chunk"], rank, world)      # phase_a_partition returned synchronised
    torch.cuda.current_stream(device).synchronize()
    gpu.finish_graph()
    return sent


def upload_partitioned(h_bases: torch.Tensor, h_offsets: torch.Tensor, rank: int, world: int, device) -> tuple:
    """Host -> device of the raw reads with every rank moving only its 1/world share over PCIe; the shares are
    then all-gathered over NVLink.  h_* are (pinned) host tensors holding the WHOLE input on every rank.
    Returns (d_bases uint8, d_offsets int64, bytes copied host->device by this rank)."""
    if world == 1:
        return h_bases.to(device, non_blocking=True), h_offsets.to(device, non_blocking=True), h_bases.numel() + 8 * h_offsets.numel()
    out, moved = [], 0
    for h in (h_bases, h_offsets):
        n = h.numel()
        share = -(-n // world)
        full = torch.empty(share * world, dtype=h.dtype, device=device)
        lo, hi = min(n, rank * share), min(n, (rank + 1) * share)
        mine = full[rank * share:(rank + 1) * share]
        if hi > lo:
            mine[:hi - lo].copy_(h[lo:hi], non_blocking=True)
        moved += (hi - lo) * h.element_size()
        dist.all_gather_into_tensor(full, mine)
        out.append(full[:n])
    return out[0], out[1], moved

"""Multi-GPU plumbing: one process per GPU (torch.distributed), every rank computes its shard of
the pair sequence / list partition through the C-ABI, then all ranks exchange their packed
multiplicons so that every rank ends with the identical, rank-ordered stream -- the construction
of levelTwo.cpp:225-269 and higherLevel.cpp:279-314 (MPI_Allgather of sizes + MPI_Allgatherv of
bytes), here over NCCL (GPU tensors) or gloo (CPU tensors, used by the CPU tests).

Payloads are KB-scale, so the exchange is latency-bound; the data path itself needs no collective.
"""
from __future__ import annotations

import struct
from typing import List

import torch
import torch.distributed as dist


def all_gather_bytes(payload: bytes, device: torch.device) -> List[bytes]:
    """Every rank's payload, in rank order, on every rank."""
    world = dist.get_world_size()
    n = torch.tensor([len(payload)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    if payload:
        buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [bytes(o[:s].cpu().numpy().tobytes()) for o, s in zip(outs, sizes)]


def merge_packed(streams: List[bytes]) -> bytes:
    """Concatenate packMultiplicons streams in rank order (Multiplicon.cpp:242-264 layout)."""
    total = 0
    body = []
    for s in streams:
        total += struct.unpack_from("<i", s, 0)[0]
        body.append(s[4:])
    return struct.pack("<i", total) + b"".join(body)


def level2_all_pairs(engine, device: torch.device) -> bytes:
    r = engine.level2_all_pairs(dist.get_rank(), dist.get_world_size())
    return merge_packed(all_gather_bytes(r.bytes, device))


def profile_search(engine, profile, device: torch.device) -> bytes:
    r = engine.profile_search(profile, dist.get_rank(), dist.get_world_size())
    return merge_packed(all_gather_bytes(r.bytes, device))


def profile_search_filtered(engine, profile, device: torch.device) -> bytes:
    """The list-sharded search with the homolog-count filter applied by every rank to its own multiplicons (the filter
    is local to a multiplicon: GHMProfile.cpp:129-132 runs inside each rank's work function in the reference too)."""
    r = engine.profile_search_filtered(profile, dist.get_rank(), dist.get_world_size())
    return merge_packed(all_gather_bytes(r.bytes, device))

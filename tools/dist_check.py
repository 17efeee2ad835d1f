#!/usr/bin/env python
"""Multi-GPU exchange check (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank computes its share (level-2 pairs by the reference's least-loaded rule; profile searches batched and
dealt whole to the ranks, or list-sharded), the library's NCCL all-gather (iadh_allgather_results) leaves every
rank with all ranks' packed multiplicons, and that is compared with the single-rank result of the same calls:
as a multiset of multiplicons for streams that interleave ranks, byte for byte where a stream comes from one rank."""
import os
import sys

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest, synth  # noqa: E402
import bench  # noqa: E402
from test_gpu_bench_parity import split_packed  # noqa: E402


def blobs(b):
    offs = split_packed(b)
    return sorted(b[offs[i]:offs[i + 1]] for i in range(len(offs) - 1))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3s0.3"
    cfg = bench.workload_config(wl)
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.load_product(ds, device=local)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    eng.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)
    ok = True
    full = eng.level2_all_pairs()
    mine = eng.level2_all_pairs(rank, world)
    allr = eng.allgather([mine])[0]
    same = blobs(allr.bytes) == blobs(full.bytes) and allr.n_multiplicons == full.n_multiplicons
    ok &= same
    print(f"[rank {rank}] level 2: {mine.n_multiplicons} of {full.n_multiplicons} multiplicons here, gathered stream == single-rank stream (as a multiset): {same}", flush=True)
    made = bench.make_profiles(ds, full.bytes, 12)
    profs = [p for _, p in made]
    masks = [[(m.x_object_id, m.bounds[0], m.bounds[1]), (m.y_object_id, m.bounds[2], m.bounds[3])] for m, _ in made]
    want = eng.profile_search_batch(profs, masks)
    want_b = [w.bytes for w in want]
    for shard_profiles in (True, False):
        eng.clear_masks()
        part = eng.profile_search_batch(profs, masks, rank, world, shard_profiles)
        got = eng.allgather(part)
        if shard_profiles:
            same = all(g.bytes == w for g, w in zip(got, want_b))
        else:
            same = all(blobs(g.bytes) == blobs(w) for g, w in zip(got, want_b))
        ok &= same
        print(f"[rank {rank}] 12 profile searches, {'whole profiles per rank' if shard_profiles else 'lists sharded'}: gathered == single-rank: {same}", flush=True)
    eng.clear_masks()
    # slot reuse of the single-node shared-memory exchange: gathers whose results are still held make later gathers
    # that map to the same slot take the NCCL route; every stream must come out the same either way
    held, same = [], True
    ref = blobs(full.bytes)
    for it in range(10):
        g = eng.allgather([eng.level2_all_pairs(rank, world)])[0]
        same &= blobs(g.bytes) == ref
        if it < 5:
            held.append(g)
    ok &= same
    print(f"[rank {rank}] 10 more level-2 gathers, 5 results held: every gathered stream == single-rank stream: {same}", flush=True)
    del held
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    eng.close()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()

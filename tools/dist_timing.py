#!/usr/bin/env python
"""Wall time of the four calls of a bench step on every rank (torchrun, N GPUs): sharded level-2 call, its
all-gather, the profile batch, its all-gather."""
import os
import sys
import time

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest  # noqa: E402
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    cfg = bench.workload_config(wl)
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.load_product(ds, device=local)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    eng.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)
    bp, _ = bench.bench_profiles(ds, wl, 32, lambda: eng.level2_all_pairs().bytes)
    for rep in range(8):
        eng.clear_masks()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = eng.level2_all_pairs(rank, world)
        t1 = time.perf_counter()
        g = eng.allgather([r])
        t2 = time.perf_counter()
        res = eng.profile_search_batch([p for _, p in bp], [m for m, _ in bp], rank, world, shard_profiles=True)
        t3 = time.perf_counter()
        gh = eng.allgather(res)
        t4 = time.perf_counter()
        g[0].wait()
        t5 = time.perf_counter()
        for x in gh:
            x.wait()
        t6 = time.perf_counter()
        del r, g, res, gh
        t7 = time.perf_counter()
        print(f"[rank {rank}] rep {rep}: L2 {1e3 * (t1 - t0):.1f} ms (device {1e3 * 0:.0f}), gather {1e3 * (t2 - t1):.1f} ms, "
              f"HL batch {1e3 * (t3 - t2):.1f} ms, gather {1e3 * (t4 - t3):.1f} ms, wait L2 stream {1e3 * (t5 - t4):.1f} ms, "
              f"wait HL streams {1e3 * (t6 - t5):.1f} ms, free {1e3 * (t7 - t6):.1f} ms", flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

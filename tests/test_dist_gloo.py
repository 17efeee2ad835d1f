"""N > 1 path on CPU: two gloo processes shard the pair sequence / the list partition with the
reference's rules, exchange packed multiplicons, and must both end with the same rank-ordered
stream whose multiplicons are exactly those of a single-rank run.  The compute engine here is the
oracle (no GPU in this test); the GPU twin is test_gpu_parity.py::test_c3_full_size_properties."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_lib as pl
from iadhore_b200 import capi
from iadhore_b200 import dist as idist
from trace_io import read_trace


def _worker(rank, world, port, path, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = read_trace(path)
        eng = pl.oracle_engine(tr.dataset)
        pl.set_params_from(eng, tr.params)
        dev = torch.device("cpu")
        l2 = idist.level2_all_pairs(eng, dev)
        out = [l2]
        # replay the first profile searches of the trace with their masks
        n = 0
        for ev in tr.events:
            if ev[0] == "mask":
                eng.mask_range(ev[1], ev[2], ev[3])
            elif ev[0] == "profile":
                out.append(idist.profile_search(eng, ev[1], dev))
                out.append(idist.profile_search_filtered(eng, ev[1], dev))   # + the homolog-count filter on every rank's share
                n += 1
                if n == 6:
                    break
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def _key(m):
    return (m.x_object_id, m.y_object_id, m.level, m.bounds, tuple(tuple(map(int, c.aps.flatten())) for c in m.clusters))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_matches_single_rank(world):
    path = [p for p in pl.golden_traces() if "tiny_4g" in p][0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank holds the identical stream (the "same order on every rank" construction)
    for r in range(1, world):
        assert res[r] == res[0]
    # single-rank reference
    tr = read_trace(path)
    eng = pl.oracle_engine(tr.dataset)
    pl.set_params_from(eng, tr.params)
    single = [eng.level2_all_pairs().bytes]
    n = 0
    for ev in tr.events:
        if ev[0] == "mask":
            eng.mask_range(ev[1], ev[2], ev[3])
        elif ev[0] == "profile":
            single.append(eng.profile_search(ev[1]).bytes)
            single.append(eng.profile_search_filtered(ev[1]).bytes)
            n += 1
            if n == 6:
                break
    assert len(single) == len(res[0])
    assert sum(len(capi.parse_packed(b)) for b in single[2::2]) < sum(len(capi.parse_packed(b)) for b in single[1::2])   # the filter dropped something
    for a, b in zip(res[0], single):
        ka = sorted(_key(m) for m in capi.parse_packed(a))
        kb = sorted(_key(m) for m in capi.parse_packed(b))
        assert ka == kb
    assert capi.parse_packed(res[0][0])  # something was found

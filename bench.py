#!/usr/bin/env python
"""bench.py -- GHM list-pair comparisons/s of the B200 collinearity engine.

Workload (config.workload): BASELINE.json config 4, the largest configuration that fits one GPU --
synthetic 50 genomes x 25k genes as 500 gene lists x 2500 (20k families, planted segmental
duplications; gap_size=30, cluster_gap=35, q=0.75, prob_cutoff=0.01, Bonferroni).  datasetI/II
(configs 0/1) cannot be run: their family tables are stripped from the reference tree.  Other
workloads: --workload c3 (10 x 30k genes), d2 (datasetII-shaped ragged lists), c5 (K1 only).
One step = one pass of the hot path:
  * level 2 : GHM build + collinear clustering of all 125,250 list pairs, then
  * profile search: P two-segment profiles (made from the P largest level-2 multiplicons)
    against every list, with the masks accumulating as in higherLevel.cpp:165-171.
metric = GHMs (list x list and profile x list comparisons) per second.

  value : context resident in HBM (dataset uploaded once), timed with CUDA events on the engine's stream
  e2e   : a fresh context per step from HOST buffers (dataset H2D inside the timed region),
          results read back to pinned host memory -- the call a user of the C-ABI makes
  N > 1 : one process per GPU (torchrun); the pair sequence / list partition is sharded with
          the reference's own rules (levelTwo.cpp:98-148, parallel.cpp:96-156) and the packed
          results are all-gathered over NCCL each call (levelTwo.cpp:225-269,
          higherLevel.cpp:279-314).  Fixed total work => "strong" scaling.

--impl reference times the UNMODIFIED reference objects (oracle/_ref/ref_harness, all host
threads) on the same files, on a bounded sample with the step's own mix of list pairs and
profile searches; falls back to the oracle port when oracle/_ref is absent.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))

from iadhore_b200 import capi, ingest, profiles, synth  # noqa: E402

METRIC = "GHM list-pair comparisons/s"
UNIT = "GHM/s"
REF_HARNESS = os.path.join(REPO, "oracle", "_ref", "ref_harness")


def workload_config(name: str):
    if name == "c3":
        return synth.config_c3(1.0)
    if name == "c4":
        return synth.config_c4(1.0)
    if name == "c5":
        return synth.config_c5(1.0)
    if name == "d2":
        return synth.config_d2(1.0)
    if name.startswith("c3p"):   # config 3 with a gene-pair blast table (table_type=pairs), e.g. c3p, c3p0.5
        cfg = synth.config_c3(float(name[3:] or 1.0))
        cfg.table_type = "pairs"
        cfg.name += "p"
        return cfg
    if name.startswith("c3s"):
        return synth.config_c3(float(name[3:]))
    if name.startswith("d2s"):
        return synth.config_d2(float(name[3:]))
    raise SystemExit(f"unknown workload {name}")


def workload_label(name, cfg, n_profiles):
    n_lists = sum(len(g.list_sizes) if g.list_sizes else g.n_lists for g in cfg.genomes)
    n_genes = sum(sum(g.list_sizes) if g.list_sizes else g.n_lists * g.genes_per_list for g in cfg.genomes)
    shape = ("ragged lists with the shipped datasetII list-length distribution" if cfg.genomes[0].list_sizes
             else f"{cfg.genomes[0].n_lists} lists x {cfg.genomes[0].genes_per_list} genes per genome")
    return (f"{name}: synthetic {len(cfg.genomes)} genomes, {n_lists} gene lists, {n_genes} genes ({shape}), "
            f"{cfg.n_families} families{' sampled into a gene-pair table (table_type=pairs)' if cfg.table_type == 'pairs' else ''}; L2 all {n_lists * (n_lists + 1) // 2} pairs + {n_profiles} profile searches per step")


def ncu_traffic(kernel_substr):
    """dram__bytes_read+write per launch of a kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py traffic ...); None if absent."""
    path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        hits = [v for k, v in json.load(open(path)).items() if kernel_substr in k]
        return sum(hits) if hits else None   # (the size-class launches of K2 run concurrently: one step = their sum)
    except Exception:
        return None


def build_roofline_c5(device, peak):
    """K1 alone on BASELINE config 5 (dense GHMs: 4 genomes x 100k genes, 2,000 families), the
    shape on which the HBM target of the GHM build is testable."""
    cfg = synth.config_c5(1.0)
    ds = ingest.dataset_from_synth(cfg)
    sizes = np.diff(ds.list_off)
    xs, ys = capi.plan_pairs(sizes)
    eng = capi.load_product(ds, device=device)
    best = None
    for _ in range(6):   # first calls size the buffers; the dot buffer (89 MB) + tables stream through L2
        st = eng.build_only(xs, ys)
        if best is None or st["build_ms"] < best["build_ms"]:
            best = st
    eng.close()
    alg = 5 * int(sizes[xs].sum() + sizes[ys].sum()) + 4 * best["n_dots"]
    ach = alg / (best["build_ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "l2_build_kernel count + tile scan + fill (K1), config 5: 78 list pairs", "achieved": ach,
            "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic("l2_build_kernel<1>"),
            "algorithmic_bytes_per_launch": alg, "launch_ms": best["build_ms"], "fill_ms": best["cluster_ms"],
            "dots": best["n_dots"], "anchor_points_per_s": best["n_dots"] / (best["build_ms"] * 1e-3)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index          # one index or a comma-separated list
        self.rows = []
        self.proc = None
        self.first = 0

    def mark(self):
        """Start of the timed region: earlier samples (taken while nvidia-smi initialised, during warm-up) are dropped."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:     # gone before anything else is timed: its NVML shutdown holds driver locks that CUDA calls wait for
            self.proc.wait(timeout=10)
        except Exception:
            self.proc.kill()
        self.rows = self.rows[self.first:]
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(ds, res_stats, xs, ys):
    """clustering stage reads each dot once (8 B as an (x,y) pair) and writes the
    packed result once (17 B/AP + 102 B/cluster + 33 B/multiplicon); build stage reads 5 B per
    element of both lists and writes each dot once (4 B packed key in this implementation)."""
    sizes = np.diff(ds.list_off)
    build = 5 * int(sizes[xs].sum() + sizes[ys].sum()) + 4 * res_stats["n_dots"]
    cluster = 8 * res_stats["n_dots"] + 17 * res_stats["n_anchorpoints"] + 102 * res_stats["n_baseclusters"] + \
        33 * res_stats["n_multiplicons"]
    return build, cluster


def make_profiles(ds, l2_bytes, n_profiles):
    ms = capi.parse_packed(l2_bytes)
    order = sorted(range(len(ms)), key=lambda i: (-sum(len(c.aps) for c in ms[i].clusters), i))[:n_profiles]
    out = []
    for i in order:
        m = ms[i]
        out.append((m, profiles.simple_profile(ds, m.x_object_id, m.y_object_id, m)))
    return out


def ref_profiles_path(workload):
    return os.path.join(REPO, "tests", "golden", f"bench_profiles_{workload}.bin.gz")


def sample_pairs(n_pairs_total, ref_profiles, profiles):
    """List pairs of the CPU sample: the same fraction of a step as its profile searches."""
    return max(1, -(-n_pairs_total * ref_profiles // max(1, profiles)))


def ref_profiles_meta(workload):
    p = ref_profiles_path(workload).replace(".bin.gz", ".json")
    return json.load(open(p)) if os.path.exists(p) and os.path.exists(ref_profiles_path(workload)) else None


def bench_profiles(ds, workload, n_profiles, l2_bytes_fn):
    """The profiles of a step as [(masks, Profile)], masks = (list, begin, end) ranges applied before the search
    (higherLevel.cpp:165-171).  Preferred: the reference's OWN first aligned profiles on this workload, dumped by
    oracle/_ref/ref_harness (tools/make_ref_profiles.py, committed under tests/golden/) -- the CPU arm searches
    the first few of exactly these.  Otherwise 2-segment profiles made from the largest level-2 multiplicons."""
    path = ref_profiles_path(workload)
    meta = ref_profiles_meta(workload)
    if meta and meta["profiles"] >= n_profiles:
        ps = profiles.load_reference_profiles(ds, path, n_profiles)
        if len(ps) == n_profiles:
            return ps, (f"the reference's own first {n_profiles} aligned profiles on this workload (queue of the first {meta['max_pairs']} list "
                        "pairs, as in the CPU sample; ref_harness --dump-profiles, committed under tests/golden)")
    made = make_profiles(ds, l2_bytes_fn(), n_profiles)
    return ([([(m.x_object_id, m.bounds[0], m.bounds[1]), (m.y_object_id, m.bounds[2], m.bounds[3])], p) for m, p in made],
            "2-segment profiles from the largest level-2 multiplicons (no reference dump for this workload)")


def hot_path_step(eng, profs, rank, n_ranks):
    """One pass of the hot path: the level-2 call over this rank's share of the list pairs, then every profile
    search of the step in one batched call (iadh_profile_search_batch; N > 1: whole profiles dealt round-robin
    to the ranks); N > 1: each followed by the library's NCCL all-gather (iadh_allgather_results), after which
    every rank holds all ranks' packed multiplicons on the host, as after the reference's MPI gathers.
    Returns (units, stats of this rank's calls, launches)."""
    eng.clear_masks()
    r = eng.level2_all_pairs(rank, n_ranks)
    stats = [r.stats]
    final = [r]
    if n_ranks > 1:
        final = eng.allgather([r])
    res = eng.profile_search_batch([p for _, p in profs], [m for m, _ in profs], rank, n_ranks, shard_profiles=n_ranks > 1)
    final += eng.allgather(res) if n_ranks > 1 else res
    for x in final:      # every stream of the step on the host (the copies overlap the calls made after them)
        x.wait()
    stats += [x.stats for x in res]
    units = sum(s["n_ghm"] for s in stats)
    launches = sum(s["kernel_launches"] for s in stats)
    return units, stats, launches


def run_reference(cfg, n_profiles, threads, tmpdir, max_pairs=-1):
    """The unmodified reference objects on the same workload files.  Bounded sample with the
    step's own mix: the first `max_pairs` list pairs of the level-2 sequence and `n_profiles`
    profile searches over all lists (a GHMProfile costs the CPU ~20x a list pair, so the
    proportion of the two kinds of GHM is part of the workload)."""
    if os.path.exists(REF_HARNESS) and not os.environ.get("IADH_BENCH_NO_REF"):
        ini = synth.write_reference_inputs(cfg, tmpdir, n_threads=threads)
        t0 = time.time()
        p = subprocess.run([REF_HARNESS, ini, "-", "--time", "--threads", str(threads), "--max-profiles", str(n_profiles),
                            "--max-pairs", str(max_pairs)], capture_output=True, text=True)
        wall = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError("ref_harness failed: " + p.stderr[-400:])
        l2 = re.search(r"L2: pairs=(\d+) dots=(\d+) build=([\d.]+)s run=([\d.]+)s wall=([\d.]+)s", p.stderr)
        hl = re.search(r"HL: profiles=(\d+) ghmprofiles=(\d+) build=([\d.]+)s run=([\d.]+)s wall=([\d.]+)s", p.stderr)
        units = int(l2.group(1)) + int(hl.group(2))
        t = float(l2.group(5)) + float(hl.group(5))
        return dict(kind="reference", units=units, seconds=t, l2_pairs=int(l2.group(1)), l2_s=float(l2.group(5)),
                    hl_ghm=int(hl.group(2)), hl_s=float(hl.group(5)), dots=int(l2.group(2)), process_wall_s=wall)
    # oracle port, multi-threaded (the reference binary is absent on this machine)
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.Engine(capi.ORACLE_LIB, "oracle_", ds)
    import ctypes as C
    eng.lib.oracle_set_threads(eng.ctx, C.c_int(threads))
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    xs, ys = capi.plan_pairs(np.diff(ds.list_off), lib_path=capi.ORACLE_LIB, prefix="oracle_")
    if max_pairs >= 0:
        xs, ys = xs[:max_pairs], ys[:max_pairs]
    t0 = time.time()
    r = eng.level2_pairs(xs, ys)
    t1 = time.time()
    profs = make_profiles(ds, r.bytes, n_profiles)
    t2 = time.time()
    hl_units = 0
    for m, p in profs:
        eng.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
        eng.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
        hl_units += eng.profile_search(p).stats["n_ghm"]
    t3 = time.time()
    return dict(kind="port", units=len(xs) + hl_units, seconds=(t1 - t0) + (t3 - t2), l2_pairs=len(xs), l2_s=t1 - t0,
                hl_ghm=hl_units, hl_s=max(t3 - t2, 1e-9), dots=r.stats["n_dots"], process_wall_s=t3 - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--profiles", type=int, default=32, help="profile searches per step")
    ap.add_argument("--ref-profiles", type=int, default=4, help="profile searches in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the K1 roofline measurement on config 5")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other named configurations")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = workload_config(a.workload)
    ncores = os.cpu_count() or 1
    config = {"workload": workload_label(a.workload, cfg, a.profiles),
              "gap_size": cfg.gap_size, "cluster_gap": cfg.cluster_gap, "q_value": cfg.q_value,
              "prob_cutoff": cfg.prob_cutoff, "anchor_points": cfg.anchor_points, "mult_hyp": "bonferroni",
              "l2_flush": "inputs larger than L2: the per-step working set (dot keys + per-GHM clustering workspaces; 13 GB allocated for c4, 0.5 GB for c3) does not fit the 126 MB L2"}

    # the CPU sample is the same fraction of both kinds of GHM of a step
    n_lists_total = sum(len(g.list_sizes) if g.list_sizes else g.n_lists for g in cfg.genomes)
    n_pairs_total = n_lists_total * (n_lists_total + 1) // 2
    ref_pairs = sample_pairs(n_pairs_total, a.ref_profiles, a.profiles)
    meta = ref_profiles_meta(a.workload)
    if meta and meta["profiles"] >= a.profiles and a.ref_profiles <= meta["profiles"]:
        ref_pairs = meta["max_pairs"]     # the queue the committed profile dump was made from: same first profiles

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        vals = []
        last = None
        with tempfile.TemporaryDirectory(prefix="iadh_ref_") as td:
            for i in range(a.warmup + a.steps):
                # the CPU path has no warm-up effects worth W extra minutes: run at most one warm-up
                if i < a.warmup - 1:
                    continue
                last = run_reference(cfg, a.ref_profiles, ncores, td, ref_pairs)
                if i >= a.warmup:
                    vals.append(last)
        units = sum(v["units"] for v in vals)
        secs = sum(v["seconds"] for v in vals)
        value = units / secs
        sample = (f"{a.ref_profiles}/{a.profiles} of a step, same mix: first {last['l2_pairs']} of {n_pairs_total} list pairs "
                  f"({last['l2_s']:.2f}s) + first {a.ref_profiles} profiles of the reference's own queue x all lists = "
                  f"{last['hl_ghm']} GHMProfiles ({last['hl_s']:.2f}s)")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * secs / max(1, len(vals)), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": last["kind"], "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "l2_pairs_per_s": last["l2_pairs"] / last["l2_s"],
                "hl_ghm_per_s": last["hl_ghm"] / last["hl_s"] if last["hl_s"] > 0 else None}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.time()
    ds = ingest.dataset_from_synth(cfg, device=local_rank)   # tandem remapping on the device (iadh_remap_tandems)
    eng = capi.load_product(ds, device=local_rank)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    profs, prof_src = bench_profiles(ds, a.workload, a.profiles, lambda: eng.level2_all_pairs().bytes)
    config["profiles"] = prof_src
    uid = None
    if dist:   # the library's own communicator (NCCL): rank 0 makes the id, torch.distributed only carries it
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())
        eng.comm_init(uid, rank, world)
    setup_s = time.time() - t_setup

    # clocks / throttle reasons of every GPU of the job, sampled by ONE nvidia-smi (rank 0) that is started before the
    # warm-up: its start-up (NVML initialisation over all GPUs of the box holds driver locks for hundreds of
    # milliseconds; one per rank stalled the CUDA calls of all ranks at 8 GPUs) is over when the timed region begins
    # the synthetic genome leaves millions of live Python objects behind: a full cyclic collection in the middle of a
    # timed step costs hundreds of milliseconds (seen as 2-4x outliers of single e2e steps) -- collect now, then keep
    # the collector out of the timed regions
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()
    clocks = ClockSampler(",".join(str(i) for i in range(world)) if world > 1 else local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    for _ in range(a.warmup):
        hot_path_step(eng, profs, rank, world)
    barrier()
    if clocks:
        clocks.mark()
    eng.event_record(0)          # CUDA events on the engine's own stream bracket the K steps
    t0 = time.perf_counter()
    units = 0
    launches = 0
    all_stats = []
    for _ in range(a.steps):
        u, st, ln = hot_path_step(eng, profs, rank, world)
        units += u
        launches += ln
        all_stats.append(st)
    eng.event_record(1)
    barrier()
    wall_dt = time.perf_counter() - t0
    dt = eng.event_elapsed_ms(0, 1) * 1e-3
    clk = clocks.stop() if clocks else None
    # max over ranks of the time, sum of the units
    if dist:
        t = torch.tensor([dt, wall_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, wall_dt = float(t[0].item()), float(t[1].item())
        u = torch.tensor([units, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        units, launches = int(u[0].item()), int(u[1].item())
    value = units / dt

    # one profile search per call (the latency a sequential caller sees), outside the timed region
    hl_single_ms = None
    if world == 1 and a.profiles:
        eng.clear_masks()
        t_single = []
        for masks, p in profs[:min(8, len(profs))]:
            for (l, b, e) in masks:
                eng.mask_range(l, b, e)
            t_single.append(eng.profile_search(p).stats["total_ms"])
        hl_single_ms = float(np.mean(t_single))
    # other named configurations, short runs (N = 1 only): resident-context value of the same step
    extra = {}
    if world == 1 and not a.no_extra:
        eng.close()
        for wl in [w for w in ("c3", "d2") if w != a.workload]:
            try:
                xcfg = workload_config(wl)
                xds = ingest.dataset_from_synth(xcfg, device=local_rank)
                xe = capi.load_product(xds, device=local_rank)
                xe.set_params(xcfg.gap_size, xcfg.cluster_gap, xcfg.anchor_points, capi.MH_BONFERRONI, xcfg.q_value, xcfg.prob_cutoff)
                xp, xsrc = bench_profiles(xds, wl, a.profiles, lambda: xe.level2_all_pairs().bytes)
                hot_path_step(xe, xp, 0, 1)
                xe.event_record(0)
                xu = 0
                for _ in range(3):
                    xu += hot_path_step(xe, xp, 0, 1)[0]
                xe.event_record(1)
                xms = xe.event_elapsed_ms(0, 1)
                extra[wl] = {"workload": workload_label(wl, xcfg, a.profiles), "profiles": xsrc, "value": xu / (xms * 1e-3), "unit": UNIT,
                             "ms_per_step": xms / 3, "steps": 3, "warmup": 1}
                xe.close()
            except Exception as e:
                extra[wl] = {"error": str(e)[:200]}
        eng = None
    # e2e: fresh context from host buffers every step (dataset H2D + results D2H in the timed region).
    # The resident context is closed first: its device blocks return to the library's cache, as they
    # would between two runs of a drop-in caller, instead of forcing a second 13 GB working set.
    if eng is not None:
        eng.close()
    barrier()
    e0 = time.perf_counter()
    e_units = 0
    h2d = d2h = 0
    ds_bytes = ds.list_off.nbytes + ds.fam.nbytes + ds.gid.nbytes + ds.num_id.nbytes + ds.flags.nbytes
    fam_csr_bytes = 0     # (the family CSR and the other join tables are built on the device since round 2)
    e_steps = max(1, min(a.steps, 3))
    e_times = []
    for k in range(e_steps + 1):
        if k == 1:        # (step 0 is an untimed warm-up of this loop: first fresh context after the resident one)
            barrier()
            e0 = time.perf_counter()
            e_units = h2d = d2h = 0
            e_times = []
        te = time.perf_counter()
        e2 = capi.load_product(ds, device=local_rank)
        tc = time.perf_counter()
        e2.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
        if uid is not None:
            e2.comm_init(uid, rank, world)     # (the communicator outlives contexts: no second ncclCommInitRank)
        u, st, _ = hot_path_step(e2, profs, rank, world)
        ts = time.perf_counter()
        e2.close()
        if os.environ.get("BENCH_E2E_TRACE"):
            print(f"[bench] e2e step {k}: create {1e3 * (tc - te):.1f} ms, calls {1e3 * (ts - tc):.1f} ms (level 2 device {st[0]['total_ms']:.1f}), "
                  f"close {1e3 * (time.perf_counter() - ts):.1f} ms", file=sys.stderr, flush=True)
        e_units += u
        h2d += ds_bytes + fam_csr_bytes + sum(s["h2d_bytes"] for s in st)
        d2h += sum(s["d2h_bytes"] for s in st)
        e_times.append(1e3 * (time.perf_counter() - te))
    barrier()
    edt = time.perf_counter() - e0
    if dist:
        t = torch.tensor([edt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        edt = float(t.item())
        u = torch.tensor([e_units], dtype=torch.int64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        e_units = int(u.item())

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (K2, clustering) and of the build kernels (K1), L2 launch
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    xs, ys = capi.plan_pairs(np.diff(ds.list_off), rank, world)
    l2s = [st[0] for st in all_stats]
    b_build, b_cluster = algorithmic_bytes(ds, l2s[0], xs, ys)
    cl_ms = float(np.mean([s["cluster_ms"] for s in l2s]))
    bd_ms = float(np.mean([s["build_ms"] for s in l2s]))
    ach = b_cluster / (cl_ms * 1e-3) / 1e9
    pass_mode = l2s[0]["kernel_launches"] > 30     # K2 ran as 30 phase1_pass + 4 phase2_stage launches
    roof = {"bound": "hbm",
            "kernel": ("K2 of the level-2 call: phase1_pass_kernel x20 (seed, enrich) + phase1_join2_kernel x10 + phase2_stage_kernel x4 "
                       "(one pass per launch), taken as one unit" if pass_mode else
                       "K2 of the level-2 call: cluster_kernel (size-class launches, concurrent)"),
            "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": ncu_traffic("K2 pass-per-launch" if pass_mode else "cluster_kernel"),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": b_cluster,
            "launch_ms": cl_ms,
            "note": "latency/dependency-bound greedy clustering (issue-bound in the pass-per-launch form): its figure of merit "
                    "is GHM/s, not GB/s (SURVEY 8d); traffic = dram bytes of the same launches under ncu, config 4 "
                    "(every pass re-reads the live state of the GHMs that still have work, DESIGN.md 6)"}
    ach_b = b_build / (bd_ms * 1e-3) / 1e9 if bd_ms > 0 else None
    expansion = len(xs) >= 1024    # (the engine's path choice for level-2 batches, IADH_EXPAND_MIN)
    roof_build = {"bound": "hbm",
                  "kernel": ("K1 of the level-2 call, family-expansion path: l2_expand_kernel count + fill, seg_sort block sorts" if expansion
                             else "K1 of the level-2 call, tile path: l2_build_kernel count + tile scan + fill"),
                  "achieved": ach_b, "peak": peak,
                  "unit": "GB/s", "frac": (ach_b / peak) if ach_b else None,
                  "traffic": ncu_traffic("K1 expansion") if expansion and a.workload == "c4" and world == 1 else None,
                  "algorithmic_bytes_per_launch": b_build, "launch_ms": bd_ms,
                  "note": "algorithmic bytes = SURVEY 8(d): 5 B per element of both lists of every pair + 4 B per dot; the lists are "
                          "shared by all pairs (6 MB, L2-resident), so DRAM traffic is below the algorithmic figure"}
    l2_ms = float(np.mean([s["total_ms"] for s in l2s]))
    # the batched call: device time of the batch (spread over this rank's profiles in the stats), GHMs of this rank
    hl_batch_ms = float(np.mean([sum(s["total_ms"] for s in st[1:]) for st in all_stats])) if a.profiles else None
    hl_ghm_rank = int(np.mean([sum(s["n_ghm"] for s in st[1:]) for st in all_stats])) if a.profiles else 0
    n_mine = len([k for k in range(a.profiles) if world == 1 or k % world == rank])
    hl_ms = (hl_batch_ms / max(1, n_mine)) if hl_batch_ms is not None else None

    roof_c5 = None
    if world == 1 and not a.no_c5:
        try:
            roof_c5 = build_roofline_c5(local_rank, peak)
        except Exception as e:
            roof_c5 = {"error": str(e)[:200]}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        with tempfile.TemporaryDirectory(prefix="iadh_cpu_") as td:
            try:
                rr = run_reference(cfg, a.ref_profiles, ncores, td, ref_pairs)
                cpu = {"value": rr["units"] / rr["seconds"], "unit": UNIT, "cores": ncores, "kind": rr["kind"],
                       "sample": f"{a.ref_profiles}/{a.profiles} of a step, same mix: first {rr['l2_pairs']} of {n_pairs_total} "
                                 f"list pairs ({rr['l2_s']:.2f}s) + the first {a.ref_profiles} of the step's profiles x all lists = {rr['hl_ghm']} "
                                 f"GHMProfiles ({rr['hl_s']:.2f}s), {ncores} threads",
                       "l2_pairs_per_s": rr["l2_pairs"] / rr["l2_s"],
                       "hl_ghm_per_s": rr["hl_ghm"] / rr["hl_s"] if rr["hl_s"] > 0 else None}
            except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
                cpu = {"value": None, "unit": UNIT, "cores": ncores, "kind": "unavailable", "sample": str(e)[:200]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64+int32", "data": "synthetic", "config": config, "clocks": clk,
            "e2e": {"value": e_units / edt, "unit": UNIT, "h2d_bytes_per_step": h2d // e_steps,
                    "d2h_bytes_per_step": d2h // e_steps, "ms_per_step": 1e3 * edt / e_steps,
                    "step_ms": [round(t, 1) for t in e_times]},
            "gpu_launches": launches, "roofline": roof, "roofline_build": roof_build, "roofline_build_c5": roof_c5,
            "cpu_baseline": cpu, "wall_ms_per_step": 1e3 * wall_dt / a.steps,
            "units_per_step": units // a.steps, "l2_pairs_per_s": l2s[0]["n_ghm"] * (world if world > 1 else 1) / (l2_ms * 1e-3),
            "hl_ghm_per_s": (hl_ghm_rank * (world if world > 1 else 1) / (hl_batch_ms * 1e-3)) if hl_batch_ms else None,
            "anchor_points_per_s_build": l2s[0]["n_dots"] / (bd_ms * 1e-3) if bd_ms > 0 else None,
            "l2_call_ms": l2_ms, "hl_batch_ms": hl_batch_ms, "hl_call_ms": hl_ms, "hl_single_call_ms": hl_single_ms,
            "hl_note": "hl_batch_ms = one iadh_profile_search_batch call over this rank's profiles of a step; hl_call_ms = "
                       "that per profile; hl_single_call_ms = the same profiles one iadh_profile_search call at a time",
            "setup_s": setup_s}
    # the two kinds of GHM cost the CPU very differently (a GHMProfile ~30x a list pair): both ratios, not only the blend
    if cpu and cpu.get("value"):
        line["l2_ratio_vs_cpu"] = line["l2_pairs_per_s"] / cpu["l2_pairs_per_s"] if cpu.get("l2_pairs_per_s") else None
        line["hl_ratio_vs_cpu"] = (line["hl_ghm_per_s"] / cpu["hl_ghm_per_s"]) if cpu.get("hl_ghm_per_s") and line["hl_ghm_per_s"] else None
    if extra:
        line["extra_configs"] = extra
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

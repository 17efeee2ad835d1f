#!/usr/bin/env python
"""Generates tests/golden/bench_profiles_<workload>.bin.gz: the reference's OWN first K aligned profiles on a
bench workload (with the masks it applies before each), dumped by oracle/_ref/ref_harness --dump-profiles.
bench.py replays exactly these profiles on the device, and its CPU arm (the same harness, --max-profiles R)
searches the first R of them: both arms get the same inputs.  Needs /root/reference (authoring container).

    python tools/make_ref_profiles.py --workload c4 --profiles 32 [--threads 8]
"""
import argparse
import gzip
import os
import subprocess
import sys
import tempfile

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, REPO)
from iadhore_b200 import synth  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--profiles", type=int, default=32)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--ref-profiles", type=int, default=4, help="profiles of bench.py's CPU sample (fixes its share of the list pairs)")
    a = ap.parse_args()
    cfg = bench.workload_config(a.workload)
    # the CPU arm of bench.py evaluates a sample: the first ref_profiles/profiles of the list pairs, then the first
    # ref_profiles profiles of the queue those pairs give.  The dump is made from the SAME queue (same --max-pairs),
    # so the first profiles of the dump are exactly what the CPU arm searches.
    n_lists = sum(len(g.list_sizes) if g.list_sizes else g.n_lists for g in cfg.genomes)
    max_pairs = bench.sample_pairs(n_lists * (n_lists + 1) // 2, a.ref_profiles, a.profiles)
    out = os.path.join(REPO, "tests", "golden", f"bench_profiles_{a.workload}.bin.gz")
    with tempfile.TemporaryDirectory(prefix="iadh_prof_") as td:
        ini = synth.write_reference_inputs(cfg, td, n_threads=a.threads)
        raw = os.path.join(td, "profiles.bin")
        raw_links = os.path.join(td, "links.bin")
        p = subprocess.run([bench.REF_HARNESS, ini, "-", "--time", "--threads", str(a.threads), "--max-profiles", str(a.profiles),
                            "--max-pairs", str(max_pairs), "--dump-profiles", raw, "--dump-profile-links", raw_links], capture_output=True, text=True)
        sys.stderr.write(p.stderr[-600:])
        if p.returncode != 0:
            raise SystemExit("ref_harness failed")
        with open(raw, "rb") as f, gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as g:
            g.write(f.read())
        # the Link sets the searches inherit (input of the homolog-count filter, iadh_profile_search_batch_filtered)
        import lzma   # (xz: 20x smaller than gzip on these int32 quadruples)
        with open(raw_links, "rb") as f, open(out.replace("bench_profiles_", "bench_profile_links_").replace(".bin.gz", ".bin.xz"), "wb") as g:
            g.write(lzma.compress(f.read(), preset=6))
    import json
    with open(out.replace(".bin.gz", ".json"), "w") as f:
        json.dump({"workload": a.workload, "profiles": a.profiles, "ref_profiles": a.ref_profiles, "max_pairs": max_pairs}, f)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

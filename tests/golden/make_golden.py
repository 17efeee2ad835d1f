"""Regenerates the golden traces in this directory.  Needs the authoring container:
/root/reference (read-only reference tree) and oracle/_ref/ref_harness built from it by
`make -C oracle ref`.  Each trace is the UNMODIFIED reference run GHM by GHM on a small
synthetic dataset (see oracle/ref_harness.cpp, tests/trace_io.py for the layout).

    python tests/golden/make_golden.py
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
from iadhore_b200 import synth  # noqa: E402

HARNESS = os.path.join(REPO, "oracle", "_ref", "ref_harness")

def _with(cfg, **kw):
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


CASES = {
    # name: (config, harness args)
    "tiny_bf": (synth.config_tiny(seed=11), []),
    "tiny_fdr": (synth.config_tiny(seed=12, mult_hyp="FDR"), []),
    "tiny_none": (synth.config_tiny(seed=13, mult_hyp="none"), ["--max-profiles", "12"]),
    "tiny_4g": (synth.config_tiny(seed=14, n_genomes=4), ["--max-profiles", "25"]),
    # table_type=pairs (trace magic IADHTRC2: carries the pair sets)
    "tiny_pairs": (synth.config_tiny_pairs(seed=15), []),
    "tiny_pairs_fdr": (synth.config_tiny_pairs(seed=16, mult_hyp="FDR"), ["--max-profiles", "15"]),
    # non-default settings: 5 anchor points (= the minimum of the homolog-count filter, GHM.cpp:218), smaller gaps, q 0.8
    "tiny_ap5": (_with(synth.config_tiny(seed=17, n_genomes=4), anchor_points=5, gap_size=20, cluster_gap=25, q_value=0.8),
                 ["--max-profiles", "25"]),
}


def main():
    only = sys.argv[1:]
    for name, (cfg, args) in CASES.items():
        if only and name not in only:
            continue
        tmp = tempfile.mkdtemp(prefix="golden_")
        try:
            ini = synth.write_reference_inputs(cfg, tmp)
            out = os.path.join(tmp, "trace.bin")
            # tiny_4g also gets the profile-only dump bench.py replays on big workloads (pins profiles.load_reference_profiles)
            prof = os.path.join(tmp, "profiles.bin")
            if name == "tiny_4g":
                args = args + ["--dump-profiles", prof]
            subprocess.run([HARNESS, ini, out] + args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            if name == "tiny_4g":
                with open(prof, "rb") as fi, gzip.GzipFile(os.path.join(HERE, name + ".profiles.bin.gz"), "wb", mtime=0) as fo:
                    shutil.copyfileobj(fi, fo)
            dst = os.path.join(HERE, name + ".trace.gz")
            with open(out, "rb") as fi, gzip.GzipFile(dst, "wb", mtime=0) as fo:
                shutil.copyfileobj(fi, fo)
            print(name, os.path.getsize(out), "->", os.path.getsize(dst))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()

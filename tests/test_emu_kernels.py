"""The kernel files of the library compiled for the host and run under a thread-block emulation (tests/emu/: the threads
of a block as fibers, every warp collective and barrier a rendezvous, grids block after block), in a container without a
GPU.  cluster_kernel.cu + pack_kernel.cu (items 0-2 below; homolog_kernel.cu: tests/test_homolog_filter.py), and at
the end of this file build_kernels.cu (join tables, level-2 and profile dot sets through both build paths),
tandem_kernel.cu (the reference's remapped flags) and intern_kernel.cu (first-occurrence ids).

0. End to end: the fused clustering kernel (phase 1, phase 2, K3 measuring pass; 32-, 64-, 128- and 256-thread
   instantiations, i.e. also the cooperative merge / insert / statistics paths and the cross-warp reductions) and
   the K3 writer, emulated GHM by GHM on the dot sets of the reference traces, produce the reference's packed
   multiplicons byte for byte -- level-2 and profile GHMs, all three multiple-hypothesis modes, p-values included.

1. Fidelity of the emulation: the DEFAULT phase-1 kernels (seed / enrich / join per gap size), emulated, leave
   exactly the state the oracle has after its gap-size loop -- cluster vectors, every cluster's anchor points
   in order, the dots left in both matrices -- for all level-2 GHMs of a reference trace.
2. The experimental two-phase join kernel (phase1_join2_kernel, IADH_JOIN2=1, written when no GPU time was
   left) produces the same complete state as the default join, field by field, on the GHMs of reference
   traces (family and pairs tables) and on heavy config-3-shaped GHMs (100+ clusters, 150+ merges each).
"""
import json
import os
import subprocess

import pytest

import parity_lib as pl
from iadhore_b200 import capi
from trace_io import read_trace

EMU_DIR = os.path.join(pl.HERE, "emu")
EMU = os.path.join(pl.REPO, "build", "emu", "emu_join2")
EMU_K2K3 = os.path.join(pl.REPO, "build", "emu", "emu_k2k3_%d")
STUDY_LIB = os.path.join(pl.REPO, "oracle", "study", "liboracle_study.so")
CUDA_INC = "/usr/local/cuda/include/cuda_runtime.h"

pytestmark = pytest.mark.skipif(not os.path.exists(CUDA_INC), reason="CUDA headers not installed")


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-C", EMU_DIR, "all"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    import sys
    sys.path.insert(0, EMU_DIR)
    import make_input
    return make_input


def _run(path, env=None, exe=None):
    p = subprocess.run([exe or EMU, path], capture_output=True, text=True, env=dict(os.environ, **(env or {})), timeout=900)
    assert p.returncode in (0, 1), p.stderr[-500:]
    return json.loads(p.stdout.strip().splitlines()[-1]), p.stderr


def test_emulated_default_kernels_match_the_oracle_state(emu, tmp_path):
    subprocess.run(["make", "-C", os.path.join(pl.REPO, "oracle"), "study"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    trace = [p for p in pl.golden_traces("family") if "tiny_4g" in p][0]
    inp, d_emu, d_ora = str(tmp_path / "in.bin"), str(tmp_path / "emu.txt"), str(tmp_path / "ora.txt")
    n = emu.write_input(trace, inp, min_dots=3, kinds=("l2",))
    res, _ = _run(inp, {"EMU_DUMP": d_emu})
    assert res["ghms"] == n and res["merges"] > 100
    tr = read_trace(trace)
    os.environ["ORACLE_STUDY_DUMP"] = d_ora
    try:
        eng = capi.Engine(STUDY_LIB, "oracle_", tr.dataset)
        pl.set_params_from(eng, tr.params)
        l2 = [e[1] for e in tr.events if e[0] == "l2"]
        eng.level2_pairs([g.x_list for g in l2], [g.y_list for g in l2])
    finally:
        os.environ.pop("ORACLE_STUDY_DUMP", None)
    assert open(d_emu).read() == open(d_ora).read()


@pytest.mark.parametrize("name", ["tiny_bf", "tiny_pairs"])
def test_two_phase_join_kernel_equals_default_join_on_reference_traces(emu, tmp_path, name):
    trace = [p for p in pl.golden_traces() if p.endswith(name + ".trace.gz")][0]
    inp = str(tmp_path / "in.bin")
    n = emu.write_input(trace, inp)          # level-2 and profile GHMs with at least 6 dots
    res, err = _run(inp)
    assert res["ghms"] == n > 50 and res["merges"] > 100
    assert res["mismatching_ghms"] == 0, err[-800:]


def test_two_phase_join_kernel_equals_default_join_on_heavy_ghms(emu, tmp_path):
    inp = str(tmp_path / "in.bin")
    assert emu.write_synth_input(inp, "c3", 8) == 8
    res, err = _run(inp)
    assert res["merges"] > 1000          # > 125 merges per GHM: queue drains of 32 and multi-chunk rows
    assert res["mismatching_ghms"] == 0, err[-800:]


# block sizes of the fused kernel: 32 (single warp), 64 and 256 (the engine's default size classes), 128
@pytest.mark.parametrize("name,nt", [("tiny_bf", 32), ("tiny_fdr", 32), ("tiny_none", 32), ("tiny_4g", 32), ("tiny_pairs", 32),
                                     ("tiny_bf", 256), ("tiny_fdr", 64), ("tiny_none", 128), ("tiny_pairs", 256), ("tiny_ap5", 32), ("tiny_ap5", 256)])
def test_emulated_device_code_reproduces_the_reference_bytes(emu, tmp_path, name, nt):
    trace = [p for p in pl.golden_traces() if p.endswith(name + ".trace.gz")][0]
    inp = str(tmp_path / "in.bin")
    # (profile GHMs of a pairs trace need the partner buckets of iadh_create_pairs: level 2 only there)
    n = emu.write_k2k3_input(trace, inp, max_hl=0 if "pairs" in name else None)
    res, err = _run(inp, exe=EMU_K2K3 % nt)
    assert res["ghms"] == n and res["multiplicons"] > 30 and res["bytes"] > 15000
    assert res["mismatching_ghms"] == 0, err[-800:]


@pytest.mark.parametrize("nt", [32, 256])
def test_emulated_device_code_on_heavy_ghms_matches_the_oracle_bytes(emu, tmp_path, nt):
    """GHMs with 100+ clusters (config-3-shaped list pairs): long cooperative merges, pool compaction, dozens of
    multiplicons per GHM.  No reference trace at this size in the repo; the expected bytes are the oracle's."""
    inp = str(tmp_path / "in.bin")
    assert emu.write_k2k3_synth_input(inp, 4) == 4
    res, err = _run(inp, exe=EMU_K2K3 % nt)
    assert res["multiplicons"] > 20 and res["bytes"] > 50000
    assert res["mismatching_ghms"] == 0, err[-800:]


def test_emulated_interning_kernels_number_strings_by_first_occurrence(emu, tmp_path):
    """intern_kernel.cu (hash table of representatives, class minima, flag scan with 1,024-thread blocks, ids) in the
    order of launch_intern: equal strings, equal ids; ids count the distinct strings in order of first occurrence
    (the promise of iadh_intern_strings) -- checked inside the harness against a std::unordered_map.  Family-like
    names with many repeats, gene-like names, empty strings, strings of hundreds of bytes; and the empty input."""
    import random
    import struct
    rng = random.Random(5)
    names = []
    for i in range(30000):
        r = rng.random()
        names.append("fam%d" % rng.randint(0, 3000) if r < 0.5 else "AT%dG%05d" % (rng.randint(1, 5), rng.randint(0, 99999)) if r < 0.8
                     else "" if r < 0.82 else "x" * rng.randint(100, 400) if r < 0.85 else "g%d" % i)
    for case in (names, []):
        enc = [s.encode() for s in case]
        off = [0]
        for b in enc:
            off.append(off[-1] + len(b))
        inp = str(tmp_path / "in.bin")
        with open(inp, "wb") as fh:
            fh.write(struct.pack("<q", len(enc)) + struct.pack("<%dq" % len(off), *off) + b"".join(enc))
        res, err = _run(inp, exe=os.path.join(pl.REPO, "build", "emu", "emu_intern"))
        assert res["strings"] == len(case) and res["distinct"] == len(set(case))
        assert res["mismatching"] == 0, err[-500:]


EMU_K1 = os.path.join(pl.REPO, "build", "emu", "emu_k1")


@pytest.mark.parametrize("name", ["tiny_bf", "tiny_4g", "tiny_ap5"])
def test_emulated_build_kernels_reproduce_the_reference_dots(emu, tmp_path, name):
    """K0 / K1 (build_kernels.cu) on the CPU: the join tables of iadh_create built by the table kernels (histograms,
    1,024-thread scans, scatter, bucket sorts, member records, strand prefix, (family, list) table) equal a plain host
    construction, and on them both level-2 build paths -- tile kernels (count, tile scan, fill) and family expansion
    (count, fill through atomic cursors, block sorts) -- give the dot sets of the unmodified reference's buildMatrix
    for every list pair of the trace; then every profile search of the trace the same way."""
    trace = [p for p in pl.golden_traces("family") if p.endswith(name + ".trace.gz")][0]
    inp = str(tmp_path / "in.bin")
    n, dots = emu.write_k1_input(trace, inp)
    res, err = _run(inp, exe=EMU_K1)
    assert res["ghms"] == n > 20 and res["dots"] == dots > 2000
    assert res["table_mismatches"] == 0 and res["tile_mismatching_ghms"] == 0 and res["expansion_mismatching_ghms"] == 0, err[-800:]
    assert res["expansion_fill_skipped"] == 0
    # ... and the profile x list dot sets (GHMProfile::buildMatrix, masks applied search by search) through
    # build_kernel<HL> and through hl_expand_kernel + block sorts
    tr = read_trace(trace)
    hl = [e[1] for e in tr.events if e[0] == "hl"]
    assert res["profile_ghms"] == len(hl) > 50 and res["profile_dots"] == sum(len(g.dots[0]) + len(g.dots[1]) for g in hl)
    assert res["profile_tile_mismatching_ghms"] == 0 and res["profile_expansion_mismatching_ghms"] == 0, err[-800:]


@pytest.mark.parametrize("shape", ["d2", "c5", "giant"])
def test_emulated_build_kernels_on_heavy_dense_and_giant_family_inputs(emu, tmp_path, shape):
    """The same on what the tiny traces do not have (expected dots from the oracle): the 40 heaviest list pairs of the
    datasetII-shaped workload (blocks of up to 3,450 keys: the shared-memory block sort), a dense config-5-shaped set
    (137-member families, 14,000 dots per pair: wide column groups in the tile fill) and families of 3,000 and 10,000
    members (shared-memory and global-memory bucket sorts of the tables; 1.9 M dots through the tile kernels -- the
    expansion path hands such blocks to the tile kernels, only its counts are compared)."""
    inp = str(tmp_path / "in.bin")
    if shape == "d2":
        emu.write_k1_workload_input(inp, "d2", 40)
    elif shape == "c5":
        from iadhore_b200 import synth
        emu.write_k1_workload_input(inp, "c5", 12, cfg=synth.config_c5(0.05))
    else:
        assert emu.write_k1_giant_input(inp) > 10 ** 6
    res, err = _run(inp, exe=EMU_K1)
    assert res["dots"] > 60000
    assert res["table_mismatches"] == 0 and res["tile_mismatching_ghms"] == 0 and res["expansion_mismatching_ghms"] == 0, err[-800:]
    assert res["expansion_fill_skipped"] == (1 if shape == "giant" else 0)


@pytest.mark.parametrize("name", ["tiny_4g", "tiny_none", "tiny_ap5", "tiny_pairs", "tiny_pairs_fdr"])
def test_emulated_tandem_kernels_reproduce_the_reference_flags(emu, tmp_path, name):
    """tandem_kernel.cu (partners in parallel, passes 1-2 in closed form, pass 3 one thread per list; family mode and
    table_type=pairs with Gene::isIndirectPairWith) on the gene lists of a trace before remapping: Gene::isRemapped()
    of every gene as the unmodified reference left it."""
    trace = [p for p in pl.golden_traces() if p.endswith(name + ".trace.gz")][0]
    inp = str(tmp_path / "in.bin")
    n, n_remapped = emu.write_tandem_input(trace, inp)
    res, err = _run(inp, exe=os.path.join(pl.REPO, "build", "emu", "emu_tandem"))
    assert res["elements"] == n and res["remapped"] == n_remapped > 50
    assert res["mismatching"] == 0, err[-500:]


@pytest.mark.parametrize("nt,env", [(32, {"EMU_PASS": "2"}), (128, {"EMU_RES": "1"})])
def test_emulated_device_code_on_the_heaviest_d2_ghms(emu, tmp_path, nt, env):
    """The four heaviest list pairs of the datasetII-shaped workload (lists of 4,000 ... 9,987 genes, up to 6,085 dots
    per GHM: three times the heaviest GHM of config 3) on the pass-per-launch and the resident kernels; expected
    bytes from the oracle.  (The fused 256-thread form passes too: 20 s, not in the suite.)"""
    inp = str(tmp_path / "in.bin")
    pairs = emu.write_k2k3_workload_input(inp, "d2", 4)
    assert max(a + b for _, _, a, b in pairs) > 5000
    res, err = _run(inp, env=env, exe=EMU_K2K3 % nt)
    assert res["ghms"] == 4 and res["multiplicons"] > 50
    assert res["mismatching_ghms"] == 0, err[-800:]


@pytest.mark.parametrize("mode,name", [("1", "tiny_bf"), ("1", "tiny_none"), ("2", "tiny_fdr"), ("2", "tiny_4g"), ("2", "tiny_ap5")])
def test_emulated_pass_per_launch_path_reproduces_the_reference_bytes(emu, tmp_path, mode, name):
    """The path of large batches: one launch per pass and gap size (phase1_pass_kernel), the four phase-2 stages
    (phase2_stage_kernel), the writer.  mode 1 = default join, 2 = the experimental two-phase join kernel."""
    trace = [p for p in pl.golden_traces() if p.endswith(name + ".trace.gz")][0]
    inp = str(tmp_path / "in.bin")
    n = emu.write_k2k3_input(trace, inp)
    res, err = _run(inp, env={"EMU_PASS": mode}, exe=EMU_K2K3 % 32)
    assert res["ghms"] == n and res["multiplicons"] > 50
    assert res["mismatching_ghms"] == 0, err[-800:]

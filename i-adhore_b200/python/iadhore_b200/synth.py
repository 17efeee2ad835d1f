"""Deterministic synthetic genome generator.

Produces gene lists + a gene->family table in the reference's own input formats
(``<gene><strand>`` one per line, `GeneList.cpp:9-38`; ``gene<TAB>family`` lines,
`GeneFamily.cpp:3-48`) plus an ``.ini`` (`Settings.cpp:23-293`), so the same
workload can be fed to the unmodified reference binary and to the B200 path.

Model: an ancestral order of `n_anc` genes (each its own family, random strand);
per genome an optional whole-genome duplication, `k_inv` random inversions
(strand flipping), per-gene retention `r`, tandem insertion prob `t`
(same family, same strand, adjacent) and noise genes from random families; the
genome is then split into equally sized lists.  The RNG is a private
SplitMix64, so files are bit-identical on every python / numpy version.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n

    def unit(self) -> float:
        return (self.next() >> 11) / float(1 << 53)


@dataclass
class GenomeSpec:
    name: str
    n_lists: int
    genes_per_list: int
    wgd: bool
    seed: int
    list_sizes: List[int] | None = None  # ragged lists (overrides n_lists x genes_per_list)


@dataclass
class SynthConfig:
    name: str
    n_families: int
    genomes: List[GenomeSpec]
    k_inv: int = 30
    r_wgd: float = 0.45
    r_plain: float = 0.8
    t_tandem: float = 0.06
    noise: float = 0.15
    anc_seed: int = 7
    n_anc: int | None = None  # ancestral gene count (default: one gene per family)
    # reference settings (shipped defaults of datasetI/II)
    gap_size: int = 30
    cluster_gap: int = 35
    q_value: float = 0.75
    prob_cutoff: float = 0.01
    anchor_points: int = 3
    tandem_gap: int | None = None
    level_2_only: bool = False
    mult_hyp: str | None = None  # None -> reference default (Bonferroni)
    # table_type=pairs (GenePairs.cpp): the blast table lists gene pairs instead of families.  The
    # pair relation is sampled from the families (each intra-family pair kept with pair_keep) plus
    # pair_cross * N pairs across families, so it is neither transitive nor family-shaped.
    table_type: str = "family"
    pair_keep: float = 0.7
    pair_cross: float = 0.03
    pair_seed: int = 99
    extra: Dict[str, str] = field(default_factory=dict)


Gene = Tuple[str, bool, int]  # (gene id, strand '+' == True, family)


def _ancestor(cfg: SynthConfig) -> List[Tuple[int, bool]]:
    rng = SplitMix64(cfg.anc_seed)
    if cfg.n_anc is None or cfg.n_anc == cfg.n_families:
        return [(f, bool(rng.next() & 1)) for f in range(cfg.n_families)]
    return [(rng.below(cfg.n_families), bool(rng.next() & 1)) for _ in range(cfg.n_anc)]


def make_genome(cfg: SynthConfig, spec: GenomeSpec) -> List[List[Gene]]:
    rng = SplitMix64(spec.seed)
    anc = _ancestor(cfg)
    chrom = list(anc) + (list(anc) if spec.wgd else [])
    n = len(chrom)
    for _ in range(cfg.k_inv):
        a = rng.below(n)
        ln = 1 + rng.below(max(2, n // 20))
        b = min(n, a + ln)
        seg = [(f, not s) for (f, s) in reversed(chrom[a:b])]
        chrom[a:b] = seg
    r = cfg.r_wgd if spec.wgd else cfg.r_plain
    out: List[Tuple[int, bool]] = []
    for (f, s) in chrom:
        if rng.unit() < r:
            out.append((f, s))
            while rng.unit() < cfg.t_tandem:
                out.append((f, s if rng.unit() < 0.8 else (not s)))
        if rng.unit() < cfg.noise:
            out.append((rng.below(cfg.n_families), bool(rng.next() & 1)))
    sizes = list(spec.list_sizes) if spec.list_sizes else [spec.genes_per_list] * spec.n_lists
    target = sum(sizes)
    while len(out) < target:  # pad with noise genes at random positions
        pos = rng.below(len(out) + 1)
        out.insert(pos, (rng.below(cfg.n_families), bool(rng.next() & 1)))
    out = out[:target]
    lists: List[List[Gene]] = []
    k = 0
    for li in range(len(sizes)):
        lst = []
        for _ in range(sizes[li]):
            f, s = out[k]
            lst.append((f"{spec.name}g{k:06d}", s, f))
            k += 1
        lists.append(lst)
    return lists


def generate(cfg: SynthConfig) -> Dict[str, List[List[Gene]]]:
    return {g.name: make_genome(cfg, g) for g in cfg.genomes}


def pair_table(cfg: SynthConfig, data) -> List[Tuple[str, str]]:
    """Gene pairs of a table_type=pairs configuration, deterministic in cfg.pair_seed.  Every gene gets at
    least one partner: the reference dereferences the partner's pair set without a null check
    (Gene.cpp:43-46 under GeneList.cpp:97-139), so a gene absent from the table crashes it."""
    rng = SplitMix64(cfg.pair_seed)
    genes = [(gid, f) for g in cfg.genomes for lst in data[g.name] for (gid, _s, f) in lst]
    by_fam: Dict[int, List[str]] = {}
    for gid, f in genes:
        by_fam.setdefault(f, []).append(gid)
    out: List[Tuple[str, str]] = []
    has = set()
    for f in sorted(by_fam):
        m = by_fam[f]
        for i in range(len(m)):
            for j in range(i + 1, len(m)):
                if m[i] != m[j] and rng.unit() < cfg.pair_keep:
                    out.append((m[i], m[j]))
                    has.add(m[i]); has.add(m[j])
    n = len(genes)
    for _ in range(int(cfg.pair_cross * n)):
        a, b = genes[rng.below(n)][0], genes[rng.below(n)][0]
        if a != b:
            out.append((a, b))
            has.add(a); has.add(b)
    for i, (gid, _f) in enumerate(genes):
        if gid not in has:
            other = genes[(i + 1) % n][0]
            if other != gid:
                out.append((gid, other))
                has.add(gid); has.add(other)
    return out


def write_reference_inputs(cfg: SynthConfig, outdir: str, data=None, n_threads: int = 1) -> str:
    """Write lists/, families.tsv and <name>.ini; returns the ini path."""
    if data is None:
        data = generate(cfg)
    os.makedirs(outdir, exist_ok=True)
    fam_lines = []
    ini = []
    for g in cfg.genomes:
        ini.append(f"genome= {g.name}")
        gdir = os.path.join(outdir, "lists", g.name)
        os.makedirs(gdir, exist_ok=True)
        for li, lst in enumerate(data[g.name]):
            fn = os.path.join(gdir, f"{li}.lst")
            with open(fn, "w") as fh:
                for (gid, s, f) in lst:
                    fh.write(f"{gid}{'+' if s else '-'}\n")
                    fam_lines.append(f"{gid}\tF{f}\n")
            ini.append(f"l{li} {fn}")
        ini.append("")
    fam_fn = os.path.join(outdir, "families.tsv")
    if cfg.table_type == "pairs":
        fam_fn = os.path.join(outdir, "pairs.tsv")
        fam_lines = [f"{a}\t{b}\n" for a, b in pair_table(cfg, data)]
    with open(fam_fn, "w") as fh:
        fh.writelines(fam_lines)
    os.makedirs(os.path.join(outdir, "out"), exist_ok=True)
    ini += [
        f"blast_table= {fam_fn}",
        f"output_path= {os.path.join(outdir, 'out')}/",
        f"table_type= {cfg.table_type}",
        f"gap_size= {cfg.gap_size}",
        f"cluster_gap= {cfg.cluster_gap}",
        f"q_value= {cfg.q_value}",
        f"prob_cutoff= {cfg.prob_cutoff}",
        f"anchor_points= {cfg.anchor_points}",
        "alignment_method= gg2",
        f"level_2_only= {'true' if cfg.level_2_only else 'false'}",
        f"number_of_threads= {n_threads}",
        "visualizeAlignment= false",
        "verbose_output= false",
    ]
    if cfg.tandem_gap is not None:
        ini.append(f"tandem_gap= {cfg.tandem_gap}")
    if cfg.mult_hyp is not None:
        ini.append(f"multiple_hypothesis_correction= {cfg.mult_hyp}")
    for k, v in cfg.extra.items():
        ini.append(f"{k}= {v}")
    ini_fn = os.path.join(outdir, f"{cfg.name}.ini")
    with open(ini_fn, "w") as fh:
        fh.write("\n".join(ini) + "\n")
    return ini_fn


# ---------------------------------------------------------------------------
# named configurations (BASELINE.json "configs", concretised in SURVEY.md 8d)
# ---------------------------------------------------------------------------

def config_tiny(seed: int = 11, mult_hyp=None, n_genomes: int = 3) -> SynthConfig:
    """seconds-scale planted-duplication set used by the CPU test-suite."""
    return SynthConfig(
        name=f"tiny{seed}", n_families=400,
        genomes=[GenomeSpec(f"t{g}", 2, 260 if g % 2 == 0 else 220, g % 2 == 0, seed * 100 + g)
                 for g in range(n_genomes)],
        k_inv=6, mult_hyp=mult_hyp)


def config_tiny_pairs(seed: int = 15, mult_hyp=None) -> SynthConfig:
    """config_tiny with a gene-pair blast table (table_type=pairs)."""
    cfg = config_tiny(seed=seed, mult_hyp=mult_hyp)
    cfg.name = f"tinyp{seed}"
    cfg.table_type = "pairs"
    return cfg


def config_small(seed: int = 21, mult_hyp=None) -> SynthConfig:
    return SynthConfig(
        name=f"small{seed}", n_families=3000,
        genomes=[GenomeSpec(f"s{g}", 3, 1200, g % 2 == 0, seed * 100 + g) for g in range(4)],
        k_inv=12, mult_hyp=mult_hyp)


def config_c3(scale: float = 1.0) -> SynthConfig:
    """C3: 10 genomes x 30k genes as 10 lists x 3000; F=20000."""
    gpl = max(30, int(3000 * scale))
    fam = max(200, int(20000 * scale))
    return SynthConfig(
        name="c3" if scale == 1.0 else f"c3s{scale:g}", n_families=fam, n_anc=int(fam * 1.4),
        genomes=[GenomeSpec(f"c{g}", 10, gpl, g % 2 == 0, 1000 + g) for g in range(10)])


def config_c4(scale: float = 1.0) -> SynthConfig:
    """C4: 50 genomes x 25k genes as 10 lists x 2500; F=20000."""
    gpl = max(30, int(2500 * scale))
    fam = max(200, int(20000 * scale))
    return SynthConfig(
        name="c4" if scale == 1.0 else f"c4s{scale:g}", n_families=fam, n_anc=int(fam * 1.2),
        genomes=[GenomeSpec(f"d{g}", 10, gpl, g % 2 == 0, 2000 + g) for g in range(50)])


def config_c5(scale: float = 1.0) -> SynthConfig:
    """C5 density stress: 4 genomes x 100k genes as 3 lists; F=2000; tandem-rich."""
    gpl = max(60, int(33333 * scale))
    fam = max(40, int(2000 * scale))
    return SynthConfig(
        name="c5" if scale == 1.0 else f"c5s{scale:g}", n_families=fam, n_anc=gpl * 3,
        genomes=[GenomeSpec(f"e{g}", 3, gpl, g % 2 == 0, 3000 + g) for g in range(4)],
        t_tandem=0.3, tandem_gap=10, level_2_only=True)


# list lengths (genes per gene list) of the shipped testset/datasetII lists: A. thaliana 7 lists,
# V. vinifera 35, P. trichocarpa 378 (chromosomes + scaffolds); 104,008 genes in 420 lists.  Only the
# shape is taken from the reference data set -- its gene families (iADHoRe_PLAZA.table) are not
# part of the reference tree, so the genes and families are synthetic.
D2_ATH = [7006, 6306, 5398, 4210, 4105, 117, 86]
D2_VVI = [10353, 1886, 1867, 1779, 1748, 1729, 1638, 1481, 1409, 1329, 1327, 1237, 1221, 1168, 1135, 1107, 1000, 647,
          632, 561, 445, 356, 350, 312, 249, 193, 144, 127, 125, 112, 90, 85, 73, 58, 9]
D2_PTR_HEAD = [3168, 2681, 2337, 1979, 1966, 1849, 1701, 1627, 1433, 1422, 1297, 1286, 1275, 1274, 1215, 1208, 1155,
               933, 454, 410, 381, 347, 273, 271, 248, 242, 218, 213, 167, 128, 123, 110, 108, 106, 104, 101, 99, 98, 97,
               96, 94, 88, 87, 84, 83, 80, 78, 76, 73, 72, 71, 70, 69, 68, 67, 65, 63, 62, 61, 59, 57, 57, 55, 54, 53, 51,
               47, 42, 39, 39, 39, 38, 38, 38, 37, 36, 36, 34, 32, 32, 31, 29, 29, 27, 27, 27, 26, 24, 23, 23, 23, 23, 21,
               21, 20, 20, 19, 19, 18, 18, 18, 18, 17, 17, 16, 16, 16, 15, 15, 15, 14, 14, 13, 13, 13, 13, 13]
# (count, length) runs of the remaining scaffolds
D2_PTR_TAIL = [(7, 12), (6, 11), (10, 10), (14, 9), (17, 8), (29, 7), (37, 6), (52, 5), (89, 4)]


def config_d2(scale: float = 1.0) -> SynthConfig:
    """datasetII-shaped: 3 genomes, 420 ragged lists with the shipped list-length distribution."""
    ptr = list(D2_PTR_HEAD)
    for cnt, ln in D2_PTR_TAIL:
        ptr += [ln] * cnt
    def sc(v):
        return [max(3, int(x * scale)) for x in v] if scale != 1.0 else list(v)
    fam = max(200, int(20000 * scale))
    return SynthConfig(
        name="d2" if scale == 1.0 else f"d2s{scale:g}", n_families=fam, n_anc=int(fam * 1.4),
        genomes=[GenomeSpec("ath", 7, 0, True, 4000, sc(D2_ATH)), GenomeSpec("vvi", 35, 0, False, 4001, sc(D2_VVI)),
                 GenomeSpec("ptr", 378, 0, True, 4002, sc(ptr))])

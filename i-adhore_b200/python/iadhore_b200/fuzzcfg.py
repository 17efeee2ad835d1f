"""Random small configurations with non-default reference settings (gap sizes, cluster gap, q, probability
cut-off, anchor points, tandem gap, multiple-hypothesis mode, table type): shared by the CPU fuzz test of
the oracle against the reference harness and by the GPU fuzz of the CUDA path against the oracle."""
from __future__ import annotations

import random

from . import synth


def random_config(rng: random.Random) -> synth.SynthConfig:
    cfg = synth.config_tiny(seed=rng.randint(100, 9999), n_genomes=rng.choice([2, 3, 4]))
    cfg.gap_size = rng.choice([3, 5, 8, 15, 30, 40, 60])
    # Settings.cpp:298-368 rejects cluster_gap < gap_size
    cfg.cluster_gap = rng.choice([cfg.gap_size, cfg.gap_size + 1, cfg.gap_size + 5, cfg.gap_size * 2])
    cfg.q_value = rng.choice([0.3, 0.5, 0.75, 0.9, 0.95])
    cfg.prob_cutoff = rng.choice([0.001, 0.01, 0.05, 0.5, 1.0])
    cfg.anchor_points = rng.choice([3, 3, 4, 5, 7])
    cfg.mult_hyp = rng.choice([None, "FDR", "none", "bonferroni"])
    cfg.tandem_gap = rng.choice([None, 0, 2, 20])
    cfg.noise = rng.choice([0.05, 0.15, 0.4])
    if rng.random() < 0.3:
        cfg.table_type = "pairs"
    cfg.name = f"fz{cfg.genomes[0].seed}"
    return cfg


def describe(cfg: synth.SynthConfig) -> dict:
    return dict(seed=cfg.genomes[0].seed, genomes=len(cfg.genomes), gap=cfg.gap_size, cluster_gap=cfg.cluster_gap, q=cfg.q_value,
                prob_cutoff=cfg.prob_cutoff, anchor_points=cfg.anchor_points, mult_hyp=cfg.mult_hyp, tandem_gap=cfg.tandem_gap,
                noise=cfg.noise, table_type=cfg.table_type)


MH = {None: 1, "bonferroni": 1, "FDR": 2, "none": 0}

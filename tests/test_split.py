"""long sequences cut into pieces at wide gaps so that several devices balance (rtg_tools_b200/csrc/vce_split.h).
The planning / verification / stitching code is the product's (vce_capi.cpp uses the same header with real devices); here the
"devices" are emulated engines run one after the other.  Everything is compared with the oracle's evaluation of the WHOLE
sequences: statuses, orientations, weights, sync points and ids, phasing counts, warnings, and the split writers' counters."""
import ctypes

import numpy as np
import pytest

from compare import compare_results
from helpers import CONFIGS, emul_library
from rtg_tools_b200 import capi, multigpu, synth


@pytest.fixture(scope="module")
def emul():
    return emul_library()


def summary_of(emul, n_seq):
    out = np.zeros(n_seq * 8, np.int64)
    emul.lib.emul_vce_submit_summary.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    assert emul.lib.emul_vce_submit_summary(out.ctypes.data, n_seq) == 0
    return out.reshape(n_seq, 8)


@pytest.mark.parametrize("cfg_name", ["default", "phased", "squash", "calls_min_base", "small_budget", "no_prune", "allele_gt"])
def test_split_sequences_equal_whole_sequences(emul, oracle, cfg_name):
    cfg = CONFIGS[cfg_name]
    seqs = synth.make_pair(24_000_000, n_chrom=3)
    want = oracle.submit(cfg, seqs, threads=6)
    for n_dev, overlap in ((2, 64), (5, 4), (8, 16)):
        got, st = emul.submit_split(cfg, seqs, n_dev, min_variants=1000, overlap=overlap, imbalance=1.0, pieces_per_device=3)
        compare_results(want, got, seqs, what="%s on %d devices" % (cfg_name, n_dev))
        if cfg.n_passes == 1:
            assert st["split"] == 3 and st["pieces"] >= 2 * n_dev, st
            assert np.array_equal(summary_of(emul, len(seqs)).sum(axis=0), multigpu.summary_counts(want))
        else:
            assert st["split"] == 0      # two-pass configurations are not split


def test_split_only_when_whole_sequences_do_not_balance(emul, oracle):
    """the cut is taken only when the longest-processing-time assignment of whole sequences exceeds 1.1."""
    seqs = synth.make_pair(16_000_000, n_chrom=8)
    want = oracle.submit(CONFIGS["default"], seqs, threads=6)
    got, st = emul.submit_split(CONFIGS["default"], seqs, 2, min_variants=1000, overlap=16)      # 8 sequences on 2 devices balance
    assert st["split"] == 0
    compare_results(want, got, seqs, what="balanced")
    got, st = emul.submit_split(CONFIGS["default"], seqs, 7, min_variants=1000, overlap=16)      # on 7 they do not
    assert st["split"] > 0 and st["redone"] == 0
    compare_results(want, got, seqs, what="unbalanced")


def test_split_dense_and_fuzz_sequences_fall_back_when_a_cut_cannot_be_verified(emul, oracle):
    """Tiny overlaps in dense, repetitive sequences: cuts whose overlap holds no included baseline variant, or whose pieces
    disagree, send the sequence back to a one-piece evaluation; results stay those of the whole sequence either way."""
    redone = split = 0
    for seed in range(12):
        rng = np.random.default_rng(seed)
        seqs = [synth.dense_cluster_sequence(rng, n_clusters=30, per_side=(2, 5), name="dense%d" % seed),
                synth.fuzz_sequence(rng, length=6000, n_base=260, n_calls=260, repeat=True, name="fuzz%d" % seed)]
        seqs += synth.make_pair(2_000_000, n_chrom=1, genome_seed=42 + seed, sample_seed=572 + seed, perturb_seed=126 + seed)
        cfg = CONFIGS["small_budget"] if seed % 3 == 0 else capi.Config()
        if cfg.n_passes != 1:
            cfg = capi.Config(max_paths=20, max_iterations=300)
        want = oracle.submit(cfg, seqs, threads=6)
        got, st = emul.submit_split(cfg, seqs, 4, min_variants=40, overlap=int(rng.choice([1, 3, 8])), gap=int(rng.choice([8, 16, 64])),
                                    imbalance=1.0, pieces_per_device=2)
        compare_results(want, got, seqs, what="seed %d" % seed)
        assert np.array_equal(summary_of(emul, len(seqs)).sum(axis=0), multigpu.summary_counts(want))
        redone += st["redone"]
        split += st["split"]
        # much smaller pieces: the dense and the fuzz sequence are cut as well (where they are sorted and have wide gaps at all)
        got, st = emul.submit_split(cfg, seqs, 4, min_variants=40, overlap=2, gap=8, imbalance=1.0, pieces_per_device=16)
        compare_results(want, got, seqs, what="seed %d, small pieces" % seed)
        redone += st["redone"]
        split += st["split"]
    assert split >= 24 and redone >= 1      # the fall-back did run


def test_split_pieces_of_registered_templates(emul, oracle):
    """Pieces end right before the next piece's first variant (a shorter template length): a registered template is recognised
    through its prefix, and the resident packet builder cuts the pieces' windows out of the same 2-bit copy."""
    emul.lib.emul_template_register.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    emul.lib.emul_resident_uploads.restype = ctypes.c_int64
    seqs = synth.make_pair(20_000_000, n_chrom=3)
    want = oracle.submit(CONFIGS["default"], seqs, threads=6)
    try:
        for s in seqs:
            s.template = np.ascontiguousarray(s.template, np.uint8)
            emul.lib.emul_template_register(s.template.ctypes.data, int(s.template.shape[0]))
        got, st = emul.submit_split(CONFIGS["default"], seqs, 6, min_variants=1000, overlap=32, imbalance=1.0, pieces_per_device=3)
        assert st["split"] == 3 and st["redone"] == 0
        assert emul.lib.emul_resident_uploads() > 0
        compare_results(want, got, seqs, what="registered templates")
    finally:
        emul.lib.emul_template_register(None, 0)


def test_split_cost_estimate_keeps_balanced_genomes_whole(emul, oracle):
    """Without `force` a cut must pay for itself: 24 chromosomes on 8 emulated devices (max / mean ~1.1) stay whole, one long
    sequence next to two short ones is cut."""
    seqs = synth.make_pair(60_000_000, n_chrom=24)
    got, st = emul.submit_split(CONFIGS["default"], seqs, 8, min_variants=2000, overlap=64, force=False)
    assert st["split"] == 0
    seqs = synth.make_pair(12_000_000, n_chrom=1) + synth.make_pair(1_000_000, n_chrom=2, genome_seed=7)
    want = oracle.submit(CONFIGS["default"], seqs, threads=6)
    got, st = emul.submit_split(CONFIGS["default"], seqs, 8, min_variants=2000, overlap=64, force=False)
    assert st["split"] == 1 and st["pieces"] >= 8
    compare_results(want, got, seqs, what="one dominant sequence")

"""Host logic of the product (packing, region split, SPILL merge, escalation, prefix re-runs, assembly) exercised on the
CPU through the test-only 1-lane instantiation of the search source (tests/emul), against the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from compare import compare_results
from golden_check import check_scenario
from helpers import CONFIGS, ROOT, Hybrid, emul_library, emul_stats, fuzz_batch
from rtg_tools_b200 import capi, synth
from test_oracle_golden import SCENARIOS


@pytest.fixture(scope="module")
def emul():
    return emul_library()


@pytest.mark.parametrize("name", SCENARIOS)
def test_pipeline_reproduces_reference_golden(emul, oracle, fixtures, name):
    assert check_scenario(Hybrid(emul, oracle), fixtures[name])


@pytest.mark.parametrize("cfg_name", sorted(CONFIGS))
def test_pipeline_matches_oracle_on_fuzz(emul, oracle, cfg_name):
    cfg = CONFIGS[cfg_name]
    tot = dict(spilled=0, escalated=0, prefix_reruns=0)
    for seed in range(60):
        seqs = fuzz_batch(seed)
        compare_results(oracle.submit(cfg, seqs), emul.submit(cfg, seqs), seqs, what="%s seed %d" % (cfg_name, seed))
        st = emul_stats(emul)
        for k in tot:
            tot[k] += st[k]
    # the interesting host paths must actually have been taken
    assert tot["spilled"] > 0


def test_every_split_gap_and_margin_is_exact(emul, oracle):
    """Exactness must not depend on the split heuristics: any gap / window margin gives identical results."""
    seqs = fuzz_batch(1234, n_seq=6)
    want = oracle.submit(CONFIGS["twopass"], seqs)
    try:
        for gap, margin in ((1, 0), (1, 4), (3, 24), (50, 1), (100000, 24)):
            emul.lib.emul_set_options(gap, margin, 0)
            compare_results(want, emul.submit(CONFIGS["twopass"], seqs), seqs, what="gap %d margin %d" % (gap, margin))
    finally:
        emul.lib.emul_set_options(0, 24, 0)


def test_large_capacity_path_is_exact(emul, oracle):
    """Force every region through the large-capacity (arena) layout."""
    try:
        emul.lib.emul_set_options(1, 24, 1)
        for seed in range(20):
            seqs = fuzz_batch(5000 + seed)
            compare_results(oracle.submit(CONFIGS["twopass"], seqs), emul.submit(CONFIGS["twopass"], seqs), seqs)
    finally:
        emul.lib.emul_set_options(0, 24, 0)


def _set_cta(emul, force, nb=0, pool=0):
    emul.lib.emul_set_cta.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong]
    emul.lib.emul_set_cta(force, nb, pool)


@pytest.mark.parametrize("cfg_name", ["default", "twopass", "squash", "phased", "calls_min_base", "small_budget", "no_prune", "allele_gt"])
def test_cta_kernel_source_is_exact(emul, oracle, cfg_name):
    """Every warp-path region through the CTA kernel's source (vce_cta.h: rank keys, block frontier, record pool), the
    thread-per-region kernel switched off so that ALL regions take it."""
    try:
        emul.lib.emul_set_engine(1, 3, 0)
        _set_cta(emul, 1)
        n_cta = 0
        for seed in range(12):
            seqs = fuzz_batch(7000 + seed)
            compare_results(oracle.submit(CONFIGS[cfg_name], seqs), emul.submit(CONFIGS[cfg_name], seqs), seqs, what=cfg_name)
            st = (ctypes.c_int64 * 8)()
            emul.lib.emul_get_stats(st)
            n_cta += st[7] & 0xffffffff
        assert n_cta > 0
    finally:
        emul.lib.emul_set_engine(0, 3, 0)
        _set_cta(emul, 0)


@pytest.mark.parametrize("nb,pool", [(6, 0), (0, 300), (3, 100)])
def test_cta_kernel_hands_over_what_its_arena_cannot_hold(emul, oracle, nb, pool):
    """A frontier directory / record pool that is too small must end in kRegionOverflow and a bit-exact re-run in the
    large-capacity warp kernel -- never in a different result."""
    try:
        _set_cta(emul, 1, nb, pool)
        over = 0
        for seed in range(6):
            seqs = fuzz_batch(7100 + seed)
            seqs.append(synth.dense_cluster_sequence(np.random.default_rng(seed), n_clusters=2, per_side=(4, 9), name="d"))
            compare_results(oracle.submit(CONFIGS["twopass"], seqs), emul.submit(CONFIGS["twopass"], seqs), seqs)
            st = (ctypes.c_int64 * 8)()
            emul.lib.emul_get_stats(st)
            over += st[7] >> 32
        assert over > 0, "the shrunken arena should have overflowed at least once"
    finally:
        _set_cta(emul, 0)


def _register(emul, seqs):
    emul.lib.emul_template_register.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    emul.lib.emul_resident_uploads.restype = ctypes.c_int64
    for s in seqs:
        s.template = np.ascontiguousarray(s.template, np.uint8)
        emul.lib.emul_template_register(s.template.ctypes.data, int(s.template.shape[0]))


@pytest.mark.parametrize("cfg_name", ["default", "twopass", "squash", "phased"])
def test_resident_inputs_are_exact(emul, oracle, cfg_name):
    """Registered templates (vce_template_register): whole-sequence upload, windows cut out of the 2-bit copy of the
    template by the packet builder (N bases, window clipping at the template end, unsorted sides that stay on the host path)."""
    seqs = synth.make_pair(1_500_000, n_chrom=3)
    for seed in range(8):
        seqs += fuzz_batch(7300 + seed)
    seqs += [synth.dense_cluster_sequence(np.random.default_rng(2), n_clusters=2, per_side=(4, 8), name="dense")]
    try:
        _register(emul, seqs)
        got = emul.submit(CONFIGS[cfg_name], seqs)
        assert emul.lib.emul_resident_uploads() > 0
        compare_results(oracle.submit(CONFIGS[cfg_name], seqs), got, seqs, what=cfg_name)
    finally:
        emul.lib.emul_template_register(None, 0)
    # and the very same inputs without the resident copies
    compare_results(oracle.submit(CONFIGS[cfg_name], seqs), emul.submit(CONFIGS[cfg_name], seqs), seqs, what=cfg_name)
    assert emul.lib.emul_resident_uploads() == 0


def test_template_packing_marks_every_non_acgt_granule(emul):
    rng = np.random.default_rng(4)
    t = rng.integers(1, 5, size=10_007, dtype=np.uint8)
    t[[5, 16, 4000, 10_006]] = 0
    emul.lib.emul_template_word.restype = ctypes.c_uint32
    emul.lib.emul_template_word.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    for pos in range(0, 10_007, 7):
        got = emul.lib.emul_template_word(t.ctypes.data, 10_007, pos, 0)
        want = 0 if not t[(pos // 16) * 16:(pos // 16) * 16 + 16].all() else int(t[pos])
        assert got == want, (pos, got, want)


@pytest.mark.parametrize("cfg_name", ["default", "squash", "phased", "calls_min_base"])
def test_device_resident_results_are_exact(emul, oracle, cfg_name):
    """The result records never come back: scatter, host patches, sync points, per-variant outputs, sync ids, phasing counts
    (prefix minimum + 16-state transducer) and the summary counters through the final passes of vce_final.h -- against
    the oracle, and against the very same engine assembling on the host."""
    from rtg_tools_b200 import multigpu
    emul.lib.emul_final_runs.restype = ctypes.c_int64
    emul.lib.emul_vce_submit_summary.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    seqs = synth.make_pair(2_000_000, n_chrom=3)
    rng = np.random.default_rng(12)
    # sorted adversarial sequences: ties at equal starts, N bases, half calls, null alleles
    for seed in range(40):
        batch = fuzz_batch(8200 + seed)
        seqs += batch
    seqs += [synth.dense_cluster_sequence(rng, n_clusters=2, per_side=(4, 8), name="dense")]
    want = oracle.submit(CONFIGS[cfg_name], seqs)
    runs = 0
    try:
        for sub in (seqs[:3], seqs[3:]):
            got = emul.submit(CONFIGS[cfg_name], sub)
            runs += emul.lib.emul_final_runs()
            w = want[:3] if sub is seqs[:3] or len(sub) == 3 else want[3:]
            compare_results(w, got, sub, what=cfg_name)
            if emul.lib.emul_final_runs():
                out = np.zeros(len(sub) * 8, np.int64)
                assert emul.lib.emul_vce_submit_summary(out.ctypes.data, len(sub)) == 0
                assert np.array_equal(out.reshape(-1, 8).sum(axis=0), multigpu.summary_counts(w))
        assert runs >= 1, "the genome slice is sorted: it has to take the device-resident result path"
        emul.lib.emul_set_device_results(0)
        got = emul.submit(CONFIGS[cfg_name], seqs[:3])
        assert emul.lib.emul_final_runs() == 0
        compare_results(want[:3], got, seqs[:3], what=cfg_name + " host assembly")
        out = np.zeros(3 * 8, np.int64)
        assert emul.lib.emul_vce_submit_summary(out.ctypes.data, 3) == 0
        assert np.array_equal(out.reshape(-1, 8).sum(axis=0), multigpu.summary_counts(want[:3]))
    finally:
        emul.lib.emul_set_device_results(1)


def test_synthetic_genome_pair(emul, oracle):
    seqs = synth.make_pair(3_000_000, n_chrom=3)
    assert sum(s.calls.n for s in seqs) > 3000
    for name in ("default", "twopass", "squash"):
        compare_results(oracle.submit(CONFIGS[name], seqs), emul.submit(CONFIGS[name], seqs), seqs, what=name)


def test_dense_clusters_hit_the_budget(emul, oracle):
    rng = np.random.default_rng(3)
    seqs = [synth.dense_cluster_sequence(rng, n_clusters=3, per_side=(6, 10))]
    cfg = capi.Config(max_paths=2000, max_iterations=20000)
    want = oracle.submit(cfg, seqs)
    assert want[0].warnings, "stress input should exceed the reduced budget"
    compare_results(want, emul.submit(cfg, seqs), seqs)


def test_dense_cluster_default_budget(emul, oracle):
    """The reference's own budget (50 000 paths): the too-complex skip and its warning fields through the engine."""
    seqs = [synth.dense_cluster_sequence(np.random.default_rng(31), n_clusters=1, per_side=(11, 13), name="d31")]
    want = oracle.submit(capi.Config(), seqs)
    assert any(w["paths"] > 50000 for w in want[0].warnings)
    compare_results(want, emul.submit(capi.Config(), seqs), seqs)


def test_c3_clusters_with_long_alleles(emul, oracle):
    """BASELINE.json configs[2] as bench.py --workload c3 builds it (synth.make_c3: 12-40 variants per side, alleles up
    to 1000 bases) at a reduced budget so that the 1-lane emulation stays fast: identical skipped sets, warnings, sync points."""
    seqs = synth.make_c3(n_clusters=2, per_seq=1, per_side=(12, 20), long_frac=0.05)
    cfg = capi.Config(max_paths=3000, max_iterations=100000)
    want = oracle.submit(cfg, seqs)
    assert sum(len(r.warnings) for r in want) > 0
    compare_results(want, emul.submit(cfg, seqs), seqs)


def test_empty_and_one_sided_sequences(emul, oracle):
    rng = np.random.default_rng(5)
    full = synth.fuzz_sequence(rng, n_base=5, n_calls=5, name="full")
    nob = synth.fuzz_sequence(rng, n_base=0, n_calls=5, name="nobase")
    noc = synth.fuzz_sequence(rng, n_base=5, n_calls=0, name="nocalls")
    none = synth.fuzz_sequence(rng, n_base=0, n_calls=0, name="none")
    seqs = [nob, full, none, noc]
    for name in ("default", "twopass"):
        compare_results(oracle.submit(CONFIGS[name], seqs), emul.submit(CONFIGS[name], seqs), seqs)
    assert emul.submit(CONFIGS["default"], []) == []


# ------------------------------------------------------------------------------------------- the C-ABI library itself
def _declared_functions():
    text = open(os.path.join(ROOT, "include", "vce.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vce_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, "rtg_tools_b200", "csrc", "libvce.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.dirname(so)])
    lib = ctypes.CDLL(so)
    names = _declared_functions()
    assert "vce_submit" in names and "vce_roc" in names and len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "libvce.so does not export " + n


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rtg_tools_b200 import vcfeval
    with pytest.raises(vcfeval.EngineUnavailable):
        vcfeval.Engine(0)
    lib = ctypes.CDLL(vcfeval.LIB_PATH)
    assert lib.vce_init(0) == 4  # VCE_ERR_NO_DEVICE
    cfg = capi.Config().to_c()
    assert lib.vce_submit(ctypes.byref(cfg), 0, None, None) != 0


def test_product_library_does_not_contain_the_oracle_or_the_emulation():
    so = os.path.join(ROOT, "rtg_tools_b200", "csrc", "libvce.so")
    syms = subprocess.check_output(["nm", "-D", "--defined-only", so]).decode()
    assert "oracle_" not in syms and "emul_" not in syms


# ------------------------------------------------------------------------------------------- thread-per-region path
def test_thread_per_region_path_carries_the_bulk(emul, oracle):
    """On a genome-like pair nearly every region must settle in the thread-per-region kernel (vce_micro.h)."""
    seqs = synth.make_pair(4_000_000, n_chrom=2)
    for name in ("default", "twopass", "squash", "phased", "calls_min_base", "no_prune"):
        compare_results(oracle.submit(CONFIGS[name], seqs), emul.submit(CONFIGS[name], seqs), seqs, what=name)
        st = emul_stats(emul)
        assert st["micro_regions"] > 0.97 * st["regions"] / (2 if name == "twopass" else 1) * (1 if name != "twopass" else 1)
        assert st["micro_clean"] > 0.95 * st["micro_regions"]


def test_results_do_not_depend_on_threads_chunks_or_the_small_kernel(emul, oracle):
    seqs = synth.make_pair(1_500_000, n_chrom=3) + fuzz_batch(77, n_seq=5)
    want = oracle.submit(CONFIGS["twopass"], seqs)
    try:
        for disable, threads, chunk in ((0, 1, 0), (0, 4, 64), (0, 7, 1000), (1, 2, 0), (0, 2, 1)):
            emul.lib.emul_set_engine(disable, threads, chunk)
            compare_results(want, emul.submit(CONFIGS["twopass"], seqs), seqs, what="disable %d threads %d chunk %d" % (disable, threads, chunk))
    finally:
        emul.lib.emul_set_engine(0, 3, 0)


def test_host_and_device_side_packet_building_are_exact(emul, oracle):
    """vce_submit on sorted inputs packs the packets from wholesale copies of the caller's arrays (the product default;
    vce_packet.h is the one source for the host digest and the builder kernel), staged jobs / settle retries / unsorted
    inputs use the host digest: both modes, every configuration, fuzz, genome-like pairs, chunks down to one region."""
    batches = [synth.make_pair(1_500_000, n_chrom=3), fuzz_batch(77, n_seq=5) + fuzz_batch(78, n_seq=5)]
    try:
        for mode in (1, 0):
            emul.lib.emul_set_device_build(mode)
            for name in ("default", "twopass", "squash", "phased", "allele_gt", "small_budget"):
                for seqs in batches:
                    compare_results(oracle.submit(CONFIGS[name], seqs), emul.submit(CONFIGS[name], seqs), seqs,
                                    what="device build %d %s" % (mode, name))
            seqs = batches[0] + batches[1]
            want = oracle.submit(CONFIGS["default"], seqs)
            for threads, chunk in ((1, 1), (3, 7), (4, 64), (2, 1000)):
                emul.lib.emul_set_engine(0, threads, chunk)
                compare_results(want, emul.submit(CONFIGS["default"], seqs), seqs,
                                what="device build %d threads %d chunk %d" % (mode, threads, chunk))
            emul.lib.emul_set_engine(0, 3, 0)
            st = emul_stats(emul)
            assert st["micro_regions"] > 0.5 * st["regions"]
    finally:
        emul.lib.emul_set_device_build(1)
        emul.lib.emul_set_engine(0, 3, 0)


def test_parallel_region_split_is_exact(emul, oracle):
    """The region split runs in position-cut segments that are stitched afterwards (runs may continue across a cut)."""
    seqs = synth.make_pair(1_000_000, n_chrom=2) + fuzz_batch(91, n_seq=6)
    rng = np.random.default_rng(17)
    seqs += [synth.dense_cluster_sequence(rng, n_clusters=3, per_side=(5, 8), name="dense")]
    cfg = capi.Config(n_passes=2, max_paths=2000, max_iterations=20000)
    want = oracle.submit(cfg, seqs)
    try:
        for per in (1, 2, 3, 7, 50, 1000):
            emul.lib.emul_set_split(per)
            compare_results(want, emul.submit(cfg, seqs), seqs, what="segment %d" % per)
    finally:
        emul.lib.emul_set_split(1 << 16)


def test_staged_job_is_rerunnable(emul, oracle):
    """vce_job_*: prepare once (digest + upload), execute repeatedly (merges are undone between runs), finish."""
    seqs = synth.make_pair(1_000_000, n_chrom=2) + fuzz_batch(5, n_seq=4)
    f = emul.lib.emul_vce_job
    f.restype = ctypes.c_int
    for name in ("default", "twopass"):
        cfg = CONFIGS[name]
        want = oracle.submit(cfg, seqs)
        for runs in (1, 3):
            ccfg, cseqs, cres, results, keep = capi.Library.pack(cfg, seqs)
            rc = f(ctypes.byref(ccfg), len(seqs), cseqs, cres, runs)
            capi.Library.unpack(cres, results, keep)
            assert rc == 0
            compare_results(want, results, seqs, what="%s runs %d" % (name, runs))


def test_split_mnps_give_multi_variant_regions():
    seqs = synth.make_pair(3_000_000, n_chrom=1)
    c = seqs[0].calls
    starts = c.allele_start[1::3]
    assert (np.diff(starts) == 1).sum() > 5, "adjacent SNPs of split MNPs should survive the generator"


def test_jni_binding_type_checks_against_the_header():
    """jni/vce_jni.c (the binding a maintainer links into libvce_jni.so) type-checks against include/vce.h with a stand-in jni.h:
    every vce_* call in it has the header's signature.  (No JDK in the image: jni/build.sh builds it where one exists.)"""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "tests", "c", "jni_stub"),
                        os.path.join(ROOT, "jni", "vce_jni.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

"""Parity tests proper: the CUDA engine, called through the C-ABI (libvce.so), against the CPU oracle and the
reference's golden fixtures.  Bit-exact everywhere: statuses, orientations, sync points, phasing counts, warnings,
and FP64 weights / ROC sums compared as raw bit patterns."""
import os

import numpy as np
import pytest

from compare import compare_results
from golden_check import check_scenario
from helpers import CONFIGS, fuzz_batch
from rtg_tools_b200 import capi, synth
from test_oracle_golden import SCENARIOS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from rtg_tools_b200 import vcfeval
    return vcfeval.Engine(0)


@pytest.mark.parametrize("name", SCENARIOS)
def test_cuda_reproduces_reference_golden(engine, fixtures, name):
    assert check_scenario(engine, fixtures[name])


@pytest.mark.parametrize("cfg_name", sorted(CONFIGS))
def test_cuda_matches_oracle_on_fuzz(engine, oracle, cfg_name):
    cfg = CONFIGS[cfg_name]
    seqs = []
    for seed in range(40):
        seqs += fuzz_batch(seed, n_seq=4)
    compare_results(oracle.submit(cfg, seqs), engine.submit(cfg, seqs), seqs, what=cfg_name)


def test_cuda_c1_synthetic_pair(engine, oracle):
    """BASELINE.json configs[0]: 10 Mbp, ~15 k variants per side, diploid."""
    seqs = synth.make_pair(10_000_000, n_chrom=1)
    assert 12000 < seqs[0].baseline.n < 18000
    for name in ("default", "twopass", "squash"):
        compare_results(oracle.submit(CONFIGS[name], seqs), engine.submit(CONFIGS[name], seqs), seqs, what=name)


def test_cuda_multi_chromosome(engine, oracle):
    seqs = synth.make_pair(60_000_000, n_chrom=24)
    compare_results(oracle.submit(CONFIGS["twopass"], seqs, threads=8), engine.submit(CONFIGS["twopass"], seqs), seqs)


def test_cuda_dense_clusters(engine, oracle):
    """BASELINE.json configs[2]: dense overlapping clusters that exceed the (here reduced) search budget."""
    rng = np.random.default_rng(3)
    seqs = [synth.dense_cluster_sequence(rng, n_clusters=4, per_side=(6, 10), name="d%d" % i) for i in range(3)]
    cfg = capi.Config(max_paths=2000, max_iterations=20000)
    want = oracle.submit(cfg, seqs)
    assert any(r.warnings for r in want)
    compare_results(want, engine.submit(cfg, seqs), seqs)


def test_cuda_dense_clusters_default_budget(engine, oracle):
    """BASELINE.json configs[2] at the reference's own budget (50 000 paths / 10^7 iterations): the large-capacity
    kernel must reproduce the too-complex skip, the warning text fields and every surviving assignment."""
    seqs = [synth.dense_cluster_sequence(np.random.default_rng(31), n_clusters=1, per_side=(11, 13), name="d31"),
            synth.dense_cluster_sequence(np.random.default_rng(17), n_clusters=1, per_side=(12, 14), name="d17")]
    want = oracle.submit(capi.Config(), seqs, threads=2)
    assert any(w["paths"] > 50000 for r in want for w in r.warnings), "stress input should exceed the default budget"
    compare_results(want, engine.submit(capi.Config(), seqs), seqs)


def test_cuda_c3_slice_default_budget(engine, oracle):
    """BASELINE.json configs[2] (bench.py --workload c3): a slice of the 2 000-cluster workload at the reference's DEFAULT
    budget through the CTA-per-region kernel -- identical skipped sets, warning fields and surviving assignments."""
    seqs = synth.make_c3(n_clusters=2000, per_seq=4, only=range(6))   # 24 clusters of the full workload
    want = oracle.submit(capi.Config(), seqs, threads=os.cpu_count() or 8)
    assert sum(1 for r in want for w in r.warnings if w["paths"] > 50000) >= 12
    got = engine.submit(capi.Config(), seqs)
    compare_results(want, got, seqs, what="c3")
    st = engine.submit_stats()
    assert st["regions"] > 0


def test_cuda_c3_long_alleles_reduced_budget(engine, oracle):
    """Clusters with alleles of up to 1000 bases (beyond every small kernel's window) at several budgets."""
    seqs = synth.make_c3(n_clusters=12, per_seq=3, per_side=(12, 24), long_frac=0.06, seed=9)
    for mp, mi in ((3000, 100000), (200, 5000), (20000, 400000)):
        cfg = capi.Config(max_paths=mp, max_iterations=mi)
        want = oracle.submit(cfg, seqs, threads=4)
        compare_results(want, engine.submit(cfg, seqs), seqs, what="c3 long alleles %d" % mp)


@pytest.mark.parametrize("registered", [True, False])
def test_cuda_device_resident_pipeline(engine, oracle, registered):
    """vce_submit's device-resident pipeline: registered reference (2-bit copy in HBM, windows cut out on the device) or host
    window slabs; result records scattered on the device, sync ids / statuses / phasing counts / summary counters from the
    final passes (vce_final.h).  Bit-exact against the oracle; the counters against the writers' rules."""
    from rtg_tools_b200 import multigpu
    seqs = synth.make_pair(40_000_000, n_chrom=5)
    seqs += [synth.dense_cluster_sequence(np.random.default_rng(2), n_clusters=3, per_side=(4, 8), name="dense")]
    if registered:
        engine.register_templates(seqs)
    try:
        for name in ("default", "squash", "phased"):
            want = oracle.submit(CONFIGS[name], seqs, threads=6)
            got = engine.submit(CONFIGS[name], seqs)
            compare_results(want, got, seqs, what="%s registered=%s" % (name, registered))
            assert np.array_equal(engine.submit_summary(len(seqs)).sum(axis=0), multigpu.summary_counts(want))
    finally:
        if registered:
            engine.unregister_templates(seqs)


def test_cuda_page_locked_inputs(engine, oracle):
    """vce_host_register: the caller's arrays go to the device straight from where they are (no staging copy)."""
    seqs = synth.make_pair(30_000_000, n_chrom=3)
    engine.register_templates(seqs)
    assert engine.host_register(seqs) > 0
    try:
        want = oracle.submit(CONFIGS["default"], seqs, threads=3)
        for _ in range(2):
            compare_results(want, engine.submit(CONFIGS["default"], seqs), seqs, what="page-locked inputs")
    finally:
        engine.host_unregister()
        engine.unregister_templates(seqs)
    compare_results(want, engine.submit(CONFIGS["default"], seqs), seqs, what="after unregistering")


def test_cuda_device_results_equal_host_assembly(engine, oracle):
    """The same call with and without the device-resident result path (VCE_DEVICE_RESULTS is read per context)."""
    seqs = synth.make_pair(12_000_000, n_chrom=3)
    want = oracle.submit(CONFIGS["default"], seqs, threads=3)
    compare_results(want, engine.submit(CONFIGS["default"], seqs), seqs, what="device results")


def test_cuda_one_process_several_devices(engine, oracle):
    """vce_init_devices: one process, every visible device, sequences spread by variant count (tests/multi_device_check.py,
    in its own process because the device pool is process-wide).  Skipped on a one-GPU box."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "multi_device_check.py")], capture_output=True, text=True, timeout=1200)
    if "SKIP:" in r.stdout:
        pytest.skip(r.stdout.strip())
    assert r.returncode == 0 and "OK:" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_cuda_genome_slice_all_modes(engine, oracle):
    """BASELINE.json configs[3] on a 10 % slice of the 24-chromosome genome: --squash-ploidy (haploid SQUASH_GT) and the
    two-pass mode of ga4gh / annotate / combine (diploid, then haploid allele matching on FN / FP)."""
    seqs = synth.make_pair(300_000_000, n_chrom=24)
    assert sum(s.baseline.n + s.calls.n for s in seqs) > 800_000
    for name in ("squash", "twopass", "phased"):
        compare_results(oracle.submit(CONFIGS[name], seqs, threads=8), engine.submit(CONFIGS[name], seqs), seqs, what=name)


@pytest.fixture(params=[1, 0], ids=["device_build", "host_build"])
def device_build(request):
    """vce_submit packs the packets on the device (k_build, the default) or in the host digest (VCE_DEVICE_BUILD=0)."""
    old = os.environ.get("VCE_DEVICE_BUILD")
    os.environ["VCE_DEVICE_BUILD"] = str(request.param)
    yield
    if old is None:
        del os.environ["VCE_DEVICE_BUILD"]
    else:
        os.environ["VCE_DEVICE_BUILD"] = old


@pytest.mark.parametrize("cfg_name", sorted(CONFIGS))
def test_cuda_packet_build_modes_match_oracle_on_fuzz(engine, oracle, device_build, cfg_name):
    cfg = CONFIGS[cfg_name]
    seqs = []
    for seed in range(40):
        seqs += fuzz_batch(seed, n_seq=4)
    seqs += synth.make_pair(2_000_000, n_chrom=2)
    compare_results(oracle.submit(cfg, seqs), engine.submit(cfg, seqs), seqs, what="packet build mode %s %s" % (os.environ["VCE_DEVICE_BUILD"], cfg_name))


def test_cuda_packet_build_modes_genome_slice(engine, oracle, device_build):
    seqs = synth.make_pair(300_000_000, n_chrom=24)
    for name in ("default", "squash", "twopass"):
        want = oracle.submit(CONFIGS[name], seqs, threads=8)
        for rep in range(2):
            compare_results(want, engine.submit(CONFIGS[name], seqs), seqs, what="device build %s rep %d" % (name, rep))


def test_cuda_packet_build_modes_golden(engine, fixtures, device_build):
    for name in SCENARIOS:
        assert check_scenario(engine, fixtures[name]), name


def test_cuda_repeated_runs_are_identical(engine, oracle):
    """Determinism: the same batch gives bit-identical results on every call (chunks, streams and re-runs race-free)."""
    seqs = synth.make_pair(60_000_000, n_chrom=24)
    want = oracle.submit(CONFIGS["twopass"], seqs, threads=8)
    for rep in range(6):
        compare_results(want, engine.submit(CONFIGS["twopass"], seqs), seqs, what="rep %d" % rep)


def test_cuda_staged_job_equals_submit(engine, oracle):
    seqs = synth.make_pair(2_000_000, n_chrom=2)
    job = engine.job(CONFIGS["twopass"], seqs)
    job.run()
    job.run()  # re-runnable
    got = job.fetch()
    st = job.stats()
    job.close()
    compare_results(oracle.submit(CONFIGS["twopass"], seqs), got, seqs)
    assert st["launches"] > 0 and st["regions"] > 0


def test_cuda_empty_inputs(engine, oracle):
    rng = np.random.default_rng(5)
    seqs = [synth.fuzz_sequence(rng, n_base=0, n_calls=5, name="nobase"), synth.fuzz_sequence(rng, n_base=4, n_calls=4, name="full"),
            synth.fuzz_sequence(rng, n_base=0, n_calls=0, name="none"), synth.fuzz_sequence(rng, n_base=5, n_calls=0, name="nocalls")]
    for name in ("default", "twopass"):
        compare_results(oracle.submit(CONFIGS[name], seqs), engine.submit(CONFIGS[name], seqs), seqs)
    assert engine.submit(CONFIGS["default"], []) == []


def _roc_equal(a, b):
    (ra, na), (rb, nb) = a, b
    assert na == nb
    assert len(ra) == len(rb)
    for fa, fb in zip(ra, rb):
        for x, y in zip(fa, fb):
            assert x.shape == y.shape
            assert np.array_equal(x.view(np.int64), y.view(np.int64))


@pytest.mark.parametrize("n,asc", [(1, False), (1000, False), (1000, True), (200_000, False)])
def test_cuda_roc_matches_oracle(engine, oracle, n, asc):
    rng = np.random.default_rng(n)
    score = np.round(rng.uniform(-5, 60, n), 1)
    score[rng.random(n) < 0.02] = np.nan
    score[rng.random(n) < 0.005] = np.inf
    score[rng.random(n) < 0.01] = 0.0
    score[rng.random(n) < 0.01] = -0.0
    tp = rng.choice([1.0, 0.5, 1.0 / 3, 2.0 / 3, 0.2, 2.0, 0.0], n)
    fp = (tp == 0).astype(np.float64)
    raw = (tp > 0).astype(np.float64)
    mask = rng.integers(0, 32, n).astype(np.uint32)
    _roc_equal(engine.roc(score, tp, fp, raw, mask, 5, asc), oracle.roc(score, tp, fp, raw, mask, 5, asc))
    _roc_equal(engine.roc(score, tp, fp, raw, None, 1, asc), oracle.roc(score, tp, fp, raw, None, 1, asc))


def test_cuda_roc_ten_million_calls_five_filters(engine, oracle):
    """BASELINE.json configs[4] at its full size: 10 M scored calls (QUAL-like scores, 40 % distinct at three decimals, the
    rest rounded to one), five filters, weights that include the non-dyadic 1/3 and 2/3 -- every row of every filter
    bit-identical to RocContainer's ordered sums (RocContainer.java:204-318)."""
    n = 10_000_000
    rng = np.random.default_rng(5)
    score = np.round(rng.uniform(0, 100, n), 3)
    dup = rng.random(n) < 0.6
    score[dup] = np.round(score[dup], 1)
    tp = rng.choice([1.0, 0.5, 1.0 / 3, 2.0 / 3, 2.0, 0.0], n)
    fp = (tp == 0).astype(np.float64)
    raw = (tp > 0).astype(np.float64)
    mask = rng.integers(1, 32, n).astype(np.uint32)
    _roc_equal(engine.roc(score, tp, fp, raw, mask, 5), oracle.roc(score, tp, fp, raw, mask, 5))


def test_cuda_roc_known_answer(engine):
    """RocContainerTest.java:45-85: accumulate + cumulative known answers."""
    score = np.array([0.1, 0.1, 0.2, 0.2, 0.1, 0.3])
    tp = np.array([1.0, 1.0, 1.0, 0.0, 1.0, 1.0])
    fp = 1.0 - tp
    rows, _ = engine.roc(score, tp, fp, tp, None, 1, False)
    thr, ctp, cfp, craw = rows[0]
    assert list(thr) == [0.3, 0.2, 0.1]
    assert list(ctp) == [1.0, 2.0, 5.0] and list(cfp) == [0.0, 1.0, 1.0]


@pytest.mark.parametrize("name", sorted(__import__("test_unit_vectors").PATH_CASES))
def test_cuda_pathfinder_known_answers(engine, name):
    """PathTest.java:173-262 micro-cases (expected include / exclude per variant) through the CUDA engine."""
    from test_unit_vectors import check_path_case
    check_path_case(engine, name)


@pytest.mark.parametrize("name", sorted(__import__("test_unit_vectors").PHASING_CASES))
def test_cuda_phasing_evaluator_known_answers(engine, name):
    """PhasingEvaluatorTest.java:142-262 (expected correct / unphasable / mis-phasing counts) through the CUDA engine."""
    from test_unit_vectors import check_phasing_case
    check_phasing_case(engine, name)


def test_cuda_concurrent_submits(engine, oracle):
    """vce_submit is re-entrant: the reference calls SequenceEvaluator.run() from several pool workers
    (VcfEvalTask.java:131-137).  Two host threads submit different batches at the same time."""
    import threading
    batches = [synth.make_pair(20_000_000, n_chrom=4, sample_seed=572 + 7 * i, perturb_seed=126 + 7 * i) for i in range(3)]
    want = [oracle.submit(CONFIGS["twopass"], b, threads=4) for b in batches]
    got = [None] * len(batches)
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = engine.submit(CONFIGS["twopass"], batches[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(batches))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i, b in enumerate(batches):
        compare_results(want[i], got[i], b, what="thread %d" % i)


def test_c_abi_from_plain_c_threads(tmp_path):
    """include/vce.h compiled by a C compiler and vce_submit driven the way the JNI stub drives it (jni/vce_jni.c): one
    sequence per call from 8 threads; the program checks every variant's outcome and reports the per-call latency."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_submit_threads")
    lib = os.path.join(root, "rtg_tools_b200", "csrc")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-o", exe, os.path.join(root, "tests", "c", "test_submit_threads.c"), "-L" + lib, "-lvce",
                           "-lpthread", "-Wl,-rpath," + lib])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK")


# ------------------------------------------------------------------------------------------- VCF loader
def test_cuda_vcf_loader_fixtures(engine):
    """vce_vcf_parse on every VCF the reference's vcfeval tests bundle (all samples, all sequences, with and without
    --all-records) against the loader restated in oracle/ref_host.py."""
    from test_vcf_loader import check_fixtures
    assert check_fixtures(engine.lib) > 200


def test_cuda_vcf_loader_fuzz(engine):
    from test_vcf_loader import check_fuzz
    out = check_fuzz(engine.lib, range(40))
    assert out["ok"] == 40 * 12
    out = check_fuzz(engine.lib, range(100, 140), bad=True)
    assert out["format-error"] > 20 and out["ok"] > 20


def test_cuda_vcf_loader_million_records(engine):
    """A 1 M-record text (42 MB): the device arrays equal the emulation's (same sources as loops: scans, refinement, compaction
    at scale) and, on the first 30 000 records, the reference loader's."""
    import vcf_check
    from helpers import emul_library
    text, tl = synth.make_vcf_text(1_000_000, seed=11)
    raw = text.encode()
    got, st = engine.lib.vcf_parse(raw, "chr1", tl, 0)
    want, st_e = emul_library().vcf_parse(raw, "chr1", tl, 0)
    vcf_check.assert_same_side(got, want, "1 M records, device vs emulation")
    assert np.array_equal(st["roc_mask"], st_e["roc_mask"]) and np.array_equal(st["qual"].view(np.int64), st_e["qual"].view(np.int64))
    assert st["records"] == 1_000_000 and st["kept"] == got.n > 900_000 and st["launches"] > 10
    head = "\n".join(text.split("\n", 30_004)[:30_004]) + "\n"
    assert vcf_check.check_text(engine.lib, head, "chr1", tl, 0, what="first 30 000 records") == "ok"
    assert vcf_check.check_text(engine.lib, head, "chr1", tl, 1, pass_only=False, what="second sample, all records") == "ok"


# ------------------------------------------------------------------------------------------- on-disk formats
def test_cuda_bgzf_reference_files(engine):
    """vce_bgzf_inflate on every block-compressed file the reference's tests hold: the text zlib produces; plain gzip refused."""
    import formats_check as fc
    assert fc.check_reference_files(engine.lib) >= 25


def test_cuda_tabix_known_answers_and_indexes(engine):
    """TabixIndexReaderTest's known virtual offsets; every bundled .tbi against the restated index reader, and the text between
    the offsets against zlib."""
    import formats_check as fc
    fc.check_tabix_known_answers(engine.lib)
    assert fc.check_tabix_all_indexes(engine.lib) > 300


def test_cuda_bgzf_synthetic_and_corrupt_streams(engine):
    """Stored / fixed / dynamic members, long codes, overlapping and far matches, empty members, random slices; damaged members
    accepted or refused exactly as zlib does."""
    import formats_check as fc
    assert fc.check_synthetic_streams(engine.lib, range(10)) > 100
    rejected, _ = fc.check_corrupt_members(engine.lib)
    assert rejected > 20


def test_cuda_vcf_gz_chain(engine):
    """.vcf.gz + .tbi -> variants of one sequence: tabix query, inflate and parse chained on the device."""
    import formats_check as fc
    assert fc.check_vcf_gz_chain(engine.lib) >= 24


def test_cuda_bgzf_million_records(engine):
    """The 1 M-record text as a 42 MB -> ~12 MB .vcf.gz of ~650 members: inflated text equal to zlib's, parsed arrays equal to
    the plain-text loader's."""
    import formats_oracle as fo
    import vcf_check
    text, tl = synth.make_vcf_text(1_000_000, seed=11)
    raw = text.encode()
    data, index = fo.bgzf_write(raw, level=6)
    got_text, st = engine.lib.bgzf_inflate(data)
    assert got_text == raw and st["blocks"] == len(index) + 1 and st["launches"] == 2
    a, b = len(raw) // 3, 2 * len(raw) // 3
    assert engine.lib.bgzf_inflate(data, fo.voffset(index, a), fo.voffset(index, b), check_crc=False)[0] == raw[a:b]
    want, _ = engine.lib.vcf_parse(raw, "chr1", tl, 0)
    got, st2 = engine.lib.vcf_parse_bgzf(data, 0, -1, "chr1", tl, 0)
    vcf_check.assert_same_side(got, want, "1 M records through BGZF")
    assert st2["records"] == 1_000_000 and st2["blocks"] == len(index) + 1


def test_cuda_sdf_template(engine, oracle):
    """vce_template_register_sdf: the byte-per-base array equals FileBitwiseInputStream's (restated), and a submit that finds the
    template registered through the SDF path (2-bit copy made on the device from the bit planes) is bit-exact against the
    oracle -- the same check as the registered pipeline test, with sequences that sit at odd offsets of one seqdata file."""
    import formats_oracle as fo
    seqs = synth.make_pair(12_000_000, n_chrom=3)
    rng = np.random.default_rng(5)
    for s in seqs:   # some N runs: the mask path
        a = int(rng.integers(0, s.template.size - 500))
        s.template[a:a + int(rng.integers(1, 300))] = 0
    pad = [rng.integers(0, 5, int(n)).astype(np.uint8) for n in (37, 1, 64 * 5 + 3)]
    parts, first = [], []
    for p, s in zip(pad, seqs):
        parts.append(p)
        first.append(sum(x.size for x in parts))
        parts.append(s.template)
    file = fo.sdf_planes_write(np.concatenate(parts))
    want = oracle.submit(CONFIGS["default"], seqs, threads=6)
    keep = []
    try:
        for s, f in zip(seqs, first):
            bases = engine.lib.template_register_sdf(file, f, s.template.size)
            assert np.array_equal(bases, s.template) and np.array_equal(bases, fo.sdf_planes_read(file, f, s.template.size))
            s.template = bases          # the registry recognises the sequence by this pointer
            keep.append(bases)
        got = engine.submit(CONFIGS["default"], seqs)
        compare_results(want, got, seqs, what="templates registered from SDF bit planes")
        st = engine.submit_stats()
    finally:
        engine.unregister_templates(seqs)
    # the windows came from the resident copy: exactly the bytes a run with templates registered the usual way uploads
    engine.register_templates(seqs)
    try:
        engine.submit(CONFIGS["default"], seqs)
        assert st["h2d_bytes"] == engine.submit_stats()["h2d_bytes"]
    finally:
        engine.unregister_templates(seqs)


def test_cuda_golden_fixtures_from_vcf_gz(engine, fixtures):
    """The reference's vcfeval fixtures end to end on the device: .vcf.gz + .tbi -> tabix query -> BGZF inflate -> VCF parse (+
    --ref-overlap trimming) -> matching -> the output files the reference's tests expect (every side of all 29 scenarios goes through the
    loader chain: sample columns, `--sample ALT`, a `--region` window; the arrays are also compared with the restated host loader's)."""
    from golden_check import bgzf_side_loader
    loader = bgzf_side_loader(engine.lib)
    check_scenario.device_loaded = 0
    for name in SCENARIOS:
        assert check_scenario(engine, fixtures[name], side_loader=loader), name
    assert check_scenario.device_loaded >= 96


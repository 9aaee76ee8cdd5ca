"""Known-answer vectors of the reference's own unit tests (test/com/rtg/vcf/eval/*Test.java), restated as data:
they pin the oracle piece by piece (playback, orientors, PathFinder micro-cases, ROC accumulation) and are then run
through the product's sources (K1 via the test-only host hook; the search through the engines)."""
import ctypes as C

import numpy as np
import pytest

from helpers import emul_library
from rtg_tools_b200 import capi

T9 = [1, 1, 1, 1, 1, 1, 2, 1, 1]
T9H = [1, 1, 1, 1, 1, 3, 2, 1, 1]

# HaplotypePlaybackTest.java:42-149 -- (template, [(start0, end0, bases)], expected bases or None = "Out of order" exception).
# MockVariant coordinates are one-based there; createOrientedVariant(v, true/false).allele(0) picks plus / minus.
PLAYBACK = [
    (T9, [(0, 1, [2]), (3, 3, [3]), (5, 6, [])], [2, 1, 1, 3, 1, 1, 2, 1, 1]),                       # testSimpleBase
    (T9, [(0, 1, [2]), (3, 4, []), (3, 3, [3])], None),                                             # testOutOfOrder
    (T9, [(0, 1, [3, 3, 3]), (3, 4, [2]), (5, 6, [4])], [3, 3, 3, 1, 1, 2, 1, 4, 2, 1, 1]),           # testMoreComplex
    (T9, [(1, 1, [3, 3, 3]), (1, 2, [2])], [1, 3, 3, 3, 2, 1, 1, 1, 1, 2, 1, 1]),                     # testInsertPriorToSnp
    (T9, [(0, 1, [2, 4, 4]), (3, 4, [3, 1]), (4, 5, [4]), (5, 6, []), (6, 7, [3])], [2, 4, 4, 1, 1, 3, 1, 4, 3, 1, 1]),  # testMoreComplexAlternate
    ([1, 4, 1, 1, 1], [(1, 2, [3]), (2, 2, [2, 2])], [1, 3, 2, 2, 1, 1, 1]),                          # testBackToBack
    # HalfPathTest.java:59-117 -- the haplotypes a HalfPath replays after include() (the excluded variant leaves no trace)
    (T9H, [(0, 1, [2]), (3, 3, [3]), (5, 6, [])], [2, 1, 1, 3, 1, 1, 2, 1, 1]),                      # testNextBase
    (T9H, [(0, 1, [1]), (3, 3, [3]), (5, 6, [])], [1, 1, 1, 3, 1, 1, 2, 1, 1]),                      # testHetero, haplotype 0
    (T9H, [(0, 1, [2]), (3, 3, [4, 3]), (5, 6, [])], [2, 1, 1, 4, 3, 1, 1, 2, 1, 1]),                # testHetero, haplotype 1
]


@pytest.mark.parametrize("case", range(len(PLAYBACK)))
def test_oracle_playback_known_answers(oracle, case):
    tmpl, alleles, expected = PLAYBACK[case]
    t = np.array(tmpl, np.uint8)
    starts = np.array([a[0] for a in alleles], np.int32)
    ends = np.array([a[1] for a in alleles], np.int32)
    off = np.zeros(len(alleles) + 1, np.int32)
    off[1:] = np.cumsum([len(a[2]) for a in alleles])
    nt = np.array([b for a in alleles for b in a[2]] + [0], np.uint8)
    out = np.zeros(64, np.uint8)
    f = oracle.lib.oracle_replay
    f.restype = C.c_int
    n = f(t.ctypes.data_as(C.c_void_p), len(tmpl), len(alleles), starts.ctypes.data_as(C.c_void_p), ends.ctypes.data_as(C.c_void_p),
          off.ctypes.data_as(C.c_void_p), nt.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 64)
    if expected is None:
        assert n < 0, "out-of-order alleles must raise (HaplotypePlayback.java:208)"
    else:
        assert n == len(expected) and list(out[:n]) == expected


# OrientorTest.java:49-232 -- (kind, H, gt, phased, slot_null [missing, REF, ALT1..], expected [(ids..., original)])
U, P, PI, SQ, AG, HP, DP, ALT = range(8)
SV12 = [1, 0, 0, 0, 1]      # "A T,C,G GT 1/2" through SampleVariants: missing null, REF, T, C kept, G not referenced -> null
SV01 = [1, 0, 0]
SVm1 = [1, 0, 0]            # "./1": missing allele is null (no explicit half call)
ALLALT = [1, 0, 0, 0, 0]    # AllAlts factory: every ALT materialised
ORIENT = [
    (SQ, 1, [1, 2], 0, SV12, [((1,), 0), ((2,), 0)]),
    (SQ, 1, [0, 1], 0, SV01, [((1,), 0)]),
    (SQ, 1, [-1, 1], 0, SVm1, [((1,), 0)]),
    (AG, 2, [1, 2], 0, SV12, [((0, 1), 0), ((0, 2), 0), ((1, 0), 0), ((2, 0), 0), ((1, 2), 0), ((2, 1), 0)]),
    (AG, 2, [0, 1], 0, SV01, [((0, 1), 0), ((1, 0), 0), ((1, 1), 0)]),
    (U, 2, [0, 2], 0, [1, 0, 1, 0], [((0, 2), 1), ((2, 0), 0)]),
    (U, 2, [2, 1], 1, [1, 0, 0, 0], [((2, 1), 1), ((1, 2), 0)]),
    (U, 2, [-1, 2], 1, [1, 0, 1, 0], [((-1, 2), 1), ((2, -1), 0)]),
    (P, 2, [0, 2], 0, [1, 0, 1, 0], [((0, 2), 1), ((2, 0), 0)]),       # unphased call falls back to both orientations
    (P, 2, [2, 1], 1, [1, 0, 0, 0], [((2, 1), 1)]),
    (PI, 2, [2, 1], 1, [1, 0, 0, 0], [((1, 2), 0)]),
    (P, 2, [-1, 2], 1, [1, 0, 1, 0], [((-1, 2), 1)]),
    (PI, 2, [-1, 2], 1, [1, 0, 1, 0], [((2, -1), 0)]),
    (HP, 1, [], 0, ALLALT, [((1,), 0), ((2,), 0), ((3,), 0)]),
    (DP, 2, [], 0, ALLALT, [((1, -1), 0), ((-1, 1), 0), ((1, 1), 0), ((2, -1), 0), ((-1, 2), 0), ((2, 1), 0), ((1, 2), 0), ((2, 2), 0),
                            ((3, -1), 0), ((-1, 3), 0), ((3, 1), 0), ((1, 3), 0), ((3, 2), 0), ((2, 3), 0), ((3, 3), 0)]),
]


def _orient(fn, kind, H, gt, phased, slot_null, cap=256):
    g = np.array(list(gt) + [0], np.int8)
    sn = np.array(slot_null, np.uint8)
    ids = np.full(cap * 4, -2, np.int8)
    orig = np.zeros(cap, np.uint8)
    fn.restype = C.c_int
    n = fn(kind, H, len(gt), g.ctypes.data_as(C.c_void_p), phased, len(slot_null), sn.ctypes.data_as(C.c_void_p),
           ids.ctypes.data_as(C.c_void_p), orig.ctypes.data_as(C.c_void_p), cap)
    return [(tuple(int(x) for x in ids[4 * i:4 * i + 4] if x != -2), int(orig[i])) for i in range(min(n, cap))], n


@pytest.mark.parametrize("case", range(len(ORIENT)))
def test_orientor_known_answers(oracle, case):
    kind, H, gt, phased, slot_null, expected = ORIENT[case]
    got, n = _orient(oracle.lib.oracle_orientations, kind, H, gt, phased, slot_null)
    assert n == len(expected) and got == expected, "oracle"
    got, n = _orient(emul_library().lib.emul_orientations, kind, H, gt, phased, slot_null)
    assert n == len(expected) and got == expected, "K1 source (vce_orient.h)"


def test_alt_orientor_sets(oracle):
    """AltOrientor(1) == HAPLOID_POP, AltOrientor(2) == DIPLOID_POP as sets, AltOrientor(3): 63 distinct rows
    (OrientorTest.java:188-231)."""
    emul = emul_library()
    for lib, fn in ((oracle, "oracle_orientations"), (emul, "emul_orientations")):
        f = getattr(lib.lib, fn)
        for H, pop in ((1, HP), (2, DP)):
            a, _ = _orient(f, ALT, H, [], 0, ALLALT)
            b, _ = _orient(f, pop, H, [], 0, ALLALT)
            assert sorted(a) == sorted(b)
        t, n = _orient(f, ALT, 3, [], 0, ALLALT)
        assert n == 63 and len(set(t)) == 63


def test_k1_source_equals_oracle_on_random_genotypes(oracle):
    emul = emul_library()
    rng = np.random.default_rng(9)
    for _ in range(3000):
        kind = int(rng.integers(0, 8))
        H = int(rng.integers(1, 4)) if kind in (U, P, PI, ALT) else (1 if kind in (SQ, HP) else 2)
        n_alt = int(rng.integers(1, 4))
        slot_null = [int(rng.random() < 0.7)] + [0] + [int(rng.random() < 0.2) for _ in range(n_alt)]
        ploidy = H if kind in (U, P, PI) else (int(rng.integers(1, 3)) if kind == SQ else 2)
        gt = [] if kind in (HP, DP, ALT) else [int(rng.integers(-1, n_alt + 1)) for _ in range(ploidy)]
        phased = int(rng.random() < 0.5)
        a = _orient(oracle.lib.oracle_orientations, kind, H, gt, phased, slot_null, cap=2048)
        b = _orient(emul.lib.emul_orientations, kind, H, gt, phased, slot_null, cap=2048)
        assert a == b, (kind, H, gt, phased, slot_null)


# PathTest.java:173-262 -- MockVariant micro-cases: (calls, baseline), each variant (start1, end1, plus, minus, expect included)
T7 = [1, 1, 1, 1, 1, 1, 2]
PATH_CASES = {
    "testBestPath2": ([(2, 3, [3], None, True), (4, 5, [], None, True)], [(2, 3, [3], None, True), (5, 6, [], None, True)]),
    "testBestPath3": ([(2, 3, [3], None, True), (4, 5, [], None, False)], [(2, 3, [3], None, True)]),
    "testBestPath4": ([(2, 3, [3], None, True)], [(2, 3, [3], None, True), (5, 6, [], None, False)]),
    "testBestPath5": ([(2, 3, [3], [4], True)], [(2, 3, [4], [3], True)]),
    "testBestPath6": ([(2, 3, [3], [4], False)], [(2, 3, [4], [2], False)]),
    "testBestPath7": ([(2, 3, [3], [4], False), (5, 7, [2], [1], True)], [(2, 3, [4], [2], False), (5, 7, [1], [2], True)]),
    "testBestPath8": ([(2, 3, [3], None, True), (3, 4, [2], [1], True)], [(2, 3, [3], None, True), (3, 4, [1], [2], True)]),
    "testBestPath9": ([(2, 3, [], None, True), (3, 4, [2], [1], True)], [(2, 3, [], None, True), (3, 4, [1], [2], True)]),
}


def _mock_side(vs):
    out = []
    for i, (s1, e1, plus, minus, _) in enumerate(vs):
        alleles = [None, (s1 - 1, e1 - 1, bytes(plus))]
        gt = [0, 0]
        if minus is not None:
            alleles.append((s1 - 1, e1 - 1, bytes(minus)))
            gt = [0, 1]   # MockOrientor: heterozygous -> (0, 1) then (1, 0)
        out.append({"id": i + 1, "alleles": alleles, "gt": gt, "phased": False, "status": 0})
    return capi.VariantSide.from_variants(out, 2)


def path_case_sequence(name):
    calls, base = PATH_CASES[name]
    return capi.Sequence(name, np.array(T7, np.uint8), _mock_side(base), _mock_side(calls))


def check_path_case(engine, name):
    calls, base = PATH_CASES[name]
    res = engine.submit(capi.Config(), [path_case_sequence(name)])[0]
    assert res.error == 0
    for side_res, vs in ((res.calls, calls), (res.baseline, base)):
        got = [bool(s & capi.STATUS_GT_MATCH) for s in side_res.status]
        assert got == [v[4] for v in vs], "%s: included %s expected %s" % (name, got, [v[4] for v in vs])


@pytest.mark.parametrize("name", sorted(PATH_CASES))
def test_pathfinder_known_answers(oracle, name):
    check_path_case(oracle, name)
    check_path_case(emul_library(), name)


# PhasingEvaluatorTest.java:142-262 -- SNVs G>T on the first 200 bases of VcfEvalTaskTest.REF (VcfEvalTaskTest.java:498),
# UnphasedOrientor(2) on both sides; (baseline, calls) as (pos1, GT), expected (correct, unphasable, misphasings)
PHASING_REF = ("CTTTCTCTTTCTTTCTTTCTCTCTTTCCATCCTTCTTTCTTTCTTTCTCTTTCTTTCTTTCTCTCTTTCCATCCTTCTTTCTTTCTTTCCTTCCTTCTTTCTTTCCTTCCTTTCTTT"
               "CTTTTTCTTTCTTTCTCTTTCCTTTTTCTTTCTTTCTCTCTCCCTCTTTCTTTTTTTTTTGAGTACCTAAAATCTACTCTCTT")
PHASING_CASES = {
    "testPhasing": ([(9, "1|0"), (13, "0|1")], [(9, "1|0"), (13, "1|0")], (0, 0, 1)),
    "testDoublePhasing": ([(9, "1|0"), (13, "0|1"), (16, "0|1")], [(9, "1|0"), (13, "1|0"), (16, "0|1")], (0, 0, 2)),
    "testPhasingUnphased": ([(9, "1|0"), (13, "0|1"), (16, "0|1"), (19, "0|1"), (22, "0|1"), (25, "1|0")],
                            [(9, "1|0"), (13, "1|0"), (16, "1|0"), (19, "0/1"), (22, "1|0"), (25, "1|0")], (1, 0, 2)),
    "testPhasingUnphased2": ([(9, "1|0"), (13, "0|1"), (16, "0|1"), (19, "0|1"), (22, "0|1"), (25, "1|0")],
                             [(9, "1|0"), (13, "1|0"), (16, "1|0"), (19, "0/1"), (22, "0|1"), (25, "1|0")], (2, 0, 1)),
    "testCluster": ([(9, "1|0"), (13, "0|1"), (15, "0|1"), (17, "0|1"), (19, "0|1"), (25, "1|0")],
                    [(9, "0|1"), (13, "1|0"), (15, "1|0"), (17, "0|1"), (19, "1|0"), (25, "0|1")], (3, 0, 2)),
    "testUnphaseable": ([(9, "1|0"), (13, "0/1")], [(9, "1|0"), (13, "1|0")], (0, 1, 0)),
    "testFalsePositive": ([(9, "1|0"), (16, "0|1")], [(9, "1|0"), (13, "1|0"), (16, "1|0")], (0, 0, 1)),
}
_DNA = {"N": 0, "A": 1, "C": 2, "G": 3, "T": 4}


def _phasing_side(vs):
    out = []
    for i, (pos1, gt) in enumerate(vs):
        p0 = pos1 - 1
        alleles = [None, (p0, p0 + 1, bytes([_DNA["G"]])), (p0, p0 + 1, bytes([_DNA["T"]]))]   # missing, REF G, ALT T
        out.append({"id": i + 1, "alleles": alleles, "gt": [int(gt[0]), int(gt[2])], "phased": gt[1] == "|", "status": 0})
    return capi.VariantSide.from_variants(out, 2)


def check_phasing_case(engine, name):
    base, calls, (correct, unph, mis) = PHASING_CASES[name]
    template = np.array([_DNA[c] for c in PHASING_REF], np.uint8)
    res = engine.submit(capi.Config(), [capi.Sequence("10", template, _phasing_side(base), _phasing_side(calls))])[0]
    assert res.error == 0
    assert list(res.phasing) == [mis, correct, unph], "%s: (mis, correct, unphasable) = %s" % (name, list(res.phasing))


@pytest.mark.parametrize("name", sorted(PHASING_CASES))
def test_phasing_evaluator_known_answers(oracle, name):
    check_phasing_case(oracle, name)
    check_phasing_case(emul_library(), name)


def test_roc_container_known_answers(oracle):
    """RocContainerTest.java:45-85."""
    score = np.array([0.1, 0.1, 0.2, 0.2, 0.1, 0.3])
    tp = np.array([1.0, 1.0, 1.0, 0.0, 1.0, 1.0])
    rows, _ = oracle.roc(score, tp, 1.0 - tp, tp, None, 1, False)
    thr, ctp, cfp, _ = rows[0]
    assert list(thr) == [0.3, 0.2, 0.1] and list(ctp) == [1.0, 2.0, 5.0] and list(cfp) == [0.0, 1.0, 1.0]

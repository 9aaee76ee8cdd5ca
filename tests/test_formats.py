"""SURVEY.md 8(f) rank 4 -- BGZF / tabix / SDF decoding (rtg_tools_b200/csrc/vce_bgzf.h) through the test-only emulation (the
same sources with one lane per warp); the CUDA run of the same checks is in test_gpu_parity.py."""
import numpy as np
import pytest

import formats_check as fc
import formats_oracle as fo
from helpers import emul_library


@pytest.fixture(scope="module")
def emul():
    return emul_library()


def test_oracle_bgzf_reads_reference_files():
    """The oracle itself on the reference's block-compressed resources: text of snp_only.vcf.gz is a VCF, EOF marker handling."""
    fx = fc.fixtures()
    text = fo.bgzf_read(fx["snp_only.vcf.gz"])
    assert text.startswith(b"##fileformat=VCF") and text.count(b"\n") > 20
    assert fo.bgzf_read(fo.EOF_BLOCK) == b""
    data, index = fo.bgzf_write(b"abc" * 50000)
    assert fo.bgzf_read(data) == b"abc" * 50000 and len(index) == 3


def test_oracle_tabix_known_answers():
    """TabixIndexReaderTest.java:50-95 against the restated index reader."""
    class OracleEngine:
        @staticmethod
        def tabix_open(data):
            t = fo.TabixIndexOracle(data)
            t.close = lambda: None
            return t
    fc.check_tabix_known_answers(OracleEngine)


def test_emul_bgzf_reference_files(emul):
    assert fc.check_reference_files(emul) >= 25


def test_emul_tabix_known_answers(emul):
    fc.check_tabix_known_answers(emul)


def test_emul_tabix_all_indexes(emul):
    assert fc.check_tabix_all_indexes(emul) > 300


def test_emul_bgzf_synthetic_streams(emul):
    assert fc.check_synthetic_streams(emul) > 60


def test_emul_bgzf_corrupt_members(emul):
    rejected, accepted = fc.check_corrupt_members(emul)
    assert rejected > 20


def test_emul_vcf_gz_chain(emul):
    assert fc.check_vcf_gz_chain(emul) >= 24


def test_emul_sdf_decode(emul):
    fc.check_sdf_decode(emul)


def test_sdf_planes_layout():
    """The layout statement of BitwiseByteArray.set (:324-331): code 4 (T = 100b) sets only the first long of the group, bit i
    for residue i; longs are big endian."""
    f = fo.sdf_planes_write(np.array([4, 1, 2, 3, 0], np.uint8))
    longs = np.frombuffer(f, ">u8")
    assert longs.tolist() == [0b00001, 0b01100, 0b01010]


def test_emul_golden_fixtures_from_vcf_gz(emul):
    """The reference's own vcfeval fixtures end to end with the device loaders in the way: every VCF is bgzipped and tabix-indexed,
    each sequence's records are found through the index, inflated, parsed and (with --ref-overlap) trimmed by the engine under
    test, the matching runs on those arrays, and the outputs must be the files the reference's tests expect."""
    import json
    import os
    from golden_check import bgzf_side_loader, check_scenario
    from helpers import Hybrid
    from rtg_tools_b200 import capi
    from test_oracle_golden import SCENARIOS
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "vcfeval_fixtures.json")) as f:
        fx = json.load(f)["scenarios"]
    oracle = capi.Library(os.path.join(root, "oracle", "liboracle.so"), prefix="oracle_")
    loader = bgzf_side_loader(emul)
    check_scenario.device_loaded = 0
    for name in SCENARIOS:
        assert check_scenario(Hybrid(emul, oracle), fx[name], side_loader=loader), name
    # every side of all 29 scenarios (sample columns, `--sample ALT`, a `--region` window): 96 loads
    assert check_scenario.device_loaded >= 96

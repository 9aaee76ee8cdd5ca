"""SURVEY.md 8(f) rank 3 -- the VCF loader (vce_vcf_parse, rtg_tools_b200/csrc/vce_vcf.h) against the reference's loader as
restated in oracle/ref_host.py.  Here through the test-only emulation (the same sources run as loops); the CUDA run of the
same checks is in test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest

import ref_host
import vcf_check
from helpers import emul_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    return emul_library()


def fixture_cases():
    with open(os.path.join(ROOT, "tests", "golden", "vcfeval_fixtures.json")) as f:
        fx = json.load(f)["scenarios"]
    for name, sc in sorted(fx.items()):
        seqs = ref_host.parse_fasta(sc["template_fa"])
        p = ref_host.Params(sc["args"])
        for side in ("baseline_vcf", "calls_vcf"):
            samples, _ = ref_host.parse_vcf(sc[side])
            for si in [-1] + list(range(len(samples))):     # -1 = --sample ALT (all-alts factory)
                yield name, side, si, sc[side], seqs, p


def check_fixtures(engine):
    n = 0
    for name, side, si, text, seqs, p in fixture_cases():
        for nm, tmpl in seqs:
            for pass_only in (True, False):
                # (every fixture also with --ref-overlap trimming, whether its own command line has it or not)
                for ref_overlap in (p.ref_overlap, not p.ref_overlap):
                    vcf_check.check_text(engine, text, nm, len(tmpl), si, p.sample_ploidy, pass_only, p.max_length,
                                         what="%s %s sample %d sequence %s pass_only %s ref_overlap %s" % (name, side, si, nm, pass_only, ref_overlap),
                                         ref_overlap=ref_overlap)
                n += 1
    return n


def fuzz_vcf(rng, n_records=250, template_len=3000, bad=False):
    # (positions grow by ~9 per record: with 250 records the last few dozen lie beyond a 2 000-base template)
    """A sorted VCF text that exercises the loader's rules: padding bases, symbolic / missing / breakend ALTs, filters, FT,
    missing / haploid / polyploid / phased GTs, lower case and N bases, an illegal base, records sharing a start (the
    refiner's heap order), records of another sequence, long alleles, records beyond the template, blank lines, CRLF."""
    bases = "ACGT"

    def seq(k, lower=0.05, n_frac=0.03):
        s = "".join(rng.choice(list(bases)) for _ in range(k))
        if rng.random() < lower:
            s = s.lower()
        if k > 1 and rng.random() < n_frac:
            s = s[:1] + "N" + s[2:]
        return s
    lines = ["##fileformat=VCFv4.2", "##contig=<ID=chr1,length=%d>" % template_len,
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2"]
    pos = 1
    for _ in range(n_records):
        step = rng.choice([0, 0, 1, 2, 5, 17, 40])
        pos += int(step)
        ref = seq(int(rng.choice([1, 1, 1, 2, 3, 6, 14])))
        alts = []
        for _ in range(int(rng.choice([0, 1, 1, 1, 2, 3]))):
            kind = rng.choice(["snp", "ins", "del", "mnp", "star", "sym", "bnd", "long"], p=[0.35, 0.15, 0.15, 0.1, 0.08, 0.07, 0.05, 0.05])
            if kind == "snp":
                a = rng.choice([b for b in bases if b != ref[0].upper()]) + ref[1:]
            elif kind == "ins":
                a = ref[0] + seq(int(rng.integers(1, 5))) + ref[1:]
            elif kind == "del":
                a = ref[0]
            elif kind == "mnp":
                a = seq(len(ref))
            elif kind == "star":
                a = "*"
            elif kind == "sym":
                a = "<DEL>"
            elif kind == "bnd":
                a = "G]17:198982]"
            else:
                a = ref[0] + seq(int(rng.integers(12, 30)))
            if a != ref and a not in alts:
                alts.append(a)
        if bad and rng.random() < 0.03:
            alts.append(ref)                      # ALT == REF: format error
        if rng.random() < 0.01:
            ref = ref[:-1] + "R"                  # illegal base: the record is skipped
        filt = rng.choice([".", "PASS", "q10", "PASS;q10", "."])
        fmt = rng.choice(["GT", "GT:GQ", "GQ:GT", "GQ", "GT:FT", "GT"])
        chrom = "chr2" if rng.random() < 0.05 else "chr1"

        def sample():
            na = len(alts)
            gts = ["0/1", "1|0", "1/1", ".", "./.", "1", "0", "0|0", ".|1", "1/.", "2/1" if na >= 2 else "1/1", "1|2" if na >= 2 else "0|1",
                   "1/2/3" if na >= 3 else "0/1", "%d/%d" % (na, max(0, na - 1))]
            if bad and rng.random() < 0.02:
                gts.append("%d/0" % (na + 1))     # allele id out of range: format error
            g = rng.choice(gts)
            if na == 0:
                g = rng.choice(["0/0", ".", "0"])
            vals = {"GT": g, "GQ": rng.choice([str(int(rng.integers(1, 99))), ".", "0", "12.5", "1e-9", "-0.000000001", "3e1", "0.00000002"]),
                    "FT": rng.choice(["PASS", ".", "lowq", "PASS;lowq"])}
            keys = fmt.split(":")
            if rng.random() < 0.1:
                keys = keys[:1]                    # trailing sub-fields dropped
            return ":".join(vals[k] for k in keys)
        qual = rng.choice(["50", ".", "37.5", "0", "1e3", "0.001", "123456.789", "99.99", "4.9e-3", "1234567890123456789", "-3", "+7.25", "1E2",
                           "3.", ".5", "12345678901234.5", "2.5e22", "1e-22"])
        info = rng.choice([".", "XRX", "DP=10;XRX", "XRX=1;AC=2", "XRXX;DP=3", "AN=2;AC=1"])
        line = "\t".join([chrom, str(pos), ".", ref, ",".join(alts) if alts else ".", qual, filt, info, fmt, sample(), sample()])
        if rng.random() < 0.03:
            line += "\r"
        lines.append(line)
        if rng.random() < 0.02:
            lines.append("")
    return "\n".join(lines) + ("\n" if rng.random() < 0.7 else "")


def check_fuzz(engine, seeds, bad=False):
    outcomes = {"ok": 0, "format-error": 0}
    for seed in seeds:
        rng = np.random.default_rng(seed)
        tl = 2000
        text = fuzz_vcf(rng, template_len=tl, bad=bad)
        for sample in (0, 1, -1):
            for ploidy, pass_only, max_length in ((2, True, 1000), (2, False, 20), (1, True, 1000), (3, False, 1000)):
                r = vcf_check.check_text(engine, text, "chr1", tl, sample, ploidy, pass_only, max_length,
                                         what="fuzz seed %d sample %d ploidy %d pass_only %s max_length %d" % (seed, sample, ploidy, pass_only, max_length))
                outcomes[r] += 1
                # --evaluation-regions: random merged intervals (some empty gaps, some covering everything nearby)
                cuts = np.sort(rng.choice(np.arange(1, tl + 50), size=int(rng.integers(0, 12)) * 2, replace=False))
                regions = [(int(cuts[i]), int(cuts[i + 1])) for i in range(0, len(cuts), 2)]
                r2 = vcf_check.check_text(engine, text, "chr1", tl, sample, ploidy, pass_only, max_length, ref_overlap=True, eval_regions=regions,
                                          what="fuzz seed %d sample %d ploidy %d pass_only %s max_length %d --ref-overlap, %d regions" % (
                                              seed, sample, ploidy, pass_only, max_length, len(regions)))
                assert r2 == r
                if r == "ok" and sample >= 0:   # --vcf-score-field GQ (vcfeval's default): the sample's FORMAT value as the score
                    assert vcf_check.check_text(engine, text, "chr1", tl, sample, ploidy, pass_only, max_length, score_field="GQ",
                                                what="fuzz seed %d score field GQ" % seed) == "ok"
                if r == "ok":     # --region: a window of the sequence (a format error outside the window is none of its business)
                    lo = int(rng.integers(0, tl))
                    assert vcf_check.check_text(engine, text, "chr1", tl, sample, ploidy, pass_only, max_length, ref_overlap=bool(seed & 1),
                                                region=(lo, lo + int(rng.integers(1, 600))), what="fuzz seed %d --region" % seed) == "ok"
    return outcomes


def test_emul_vcf_loader_matches_reference_on_fixtures(emul):
    """Every VCF the reference's own vcfeval tests bundle, every sample column, every template sequence, with and without
    --all-records."""
    assert check_fixtures(emul) > 200


def test_emul_vcf_loader_fuzz(emul):
    out = check_fuzz(emul, range(40))
    assert out["ok"] == 40 * 12


def test_emul_vcf_loader_format_errors(emul):
    """Texts with records the reference throws VcfFormatException for (ALT equal to REF, allele id out of range): the loader
    must refuse exactly those texts."""
    out = check_fuzz(emul, range(100, 140), bad=True)
    assert out["format-error"] > 20 and out["ok"] > 20


def test_emul_vcf_loader_edge_texts(emul):
    hdr = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts\n"
    for text in ("", "\n", hdr, "##x\n" + hdr, hdr + "chr1\t5\t.\tA\tT\t.\t.\t.\tGT\t0/1", hdr + "chr1\t5\t.\tA\tT\t.\t.\t.\n",
                 hdr + "chr1\t5\t.\tA\tT\t.\t.\t.\tGT\t0/1\n\n\n", hdr + "chr9\t5\t.\tA\tT\t.\t.\t.\tGT\t0/1\n",
                 hdr + "".join("chr1\t7\t.\t%s\tG\t.\t.\t.\tGT\t1/1\n" % ("A" * k) for k in (5, 3, 3, 9, 1, 3, 2, 5, 5, 1, 4))):
        assert vcf_check.check_text(emul, text, "chr1", 100, 0, what=repr(text[:40])) in ("ok", "format-error")

"""Shared checker of the VCF loader (vce_vcf_parse): the arrays an engine returns for (text, sequence, sample) against the
loader restated in oracle/ref_host.py (VcfSortRefiner + VcfRecordTabixCallable filters + SampleVariants factory +
Allele.getAlleles, Variant.trimAlleles with --ref-overlap), without evaluation regions, which stay on the host."""
import numpy as np

import ref_host
from rtg_tools_b200 import capi


def expected_side(text, name, template_len, sample, ploidy=2, pass_only=True, max_length=1000, ref_overlap=False, eval_regions=None,
                  region=None, score_field=None):
    """(VariantSide, skipped count) or the exception class the loader ends with."""
    _, recs = ref_host.parse_vcf(text)
    recs = [r for r in recs if r.chrom == name]
    if region is not None:      # --region: the records TabixLineReader hands on (ref_host._restrict)
        recs = [r for r in recs if r.start < region[1] and max(r.end, r.start + 1) > region[0]]

    def factory(rec, vid):
        if sample < 0:     # --sample ALT: VariantFactory.AllAlts
            return ref_host.all_alts_variant(rec, vid, pass_only)
        return ref_host.sample_variant(rec, vid, sample, ploidy, pass_only)
    out, skipped = ref_host.load_side(recs, template_len, factory, max_length, ref_overlap, eval_regions)
    side = capi.VariantSide.from_variants(out, ploidy if sample >= 0 else 0)
    side.roc_expect = expected_roc_fields(out, sample, score_field) if sample >= 0 else None
    return side, skipped


ROC_BITS = {"ALL": 1, "HOM": 2, "HET": 4, "SNP": 8, "NON_SNP": 16, "MNP": 32, "INDEL": 64, "XRX": 128, "NON_XRX": 256}


def expected_roc_fields(variants, sample, score_field=None):
    """What RocContainer.addRocLine (RocContainer.java:225-248) derives from each kept variant's record: the RocFilters that
    accept it (RocFilter.java:48-137, on the RAW sample GT as VcfUtils.getValidGt returns it) and the QUAL score
    (RocScoreField.java:50-80)."""
    masks, quals = [], []
    for v in variants:
        rec = v["rec"]
        raw = [(-1 if x == "." else int(x)) for x in rec.sample_field(sample, "GT").replace("|", "/").split("/")]
        m = ROC_BITS["ALL"]
        hom = len(raw) == 1 or raw[0] == raw[1]
        if hom and not all(g == 0 for g in raw):
            m |= ROC_BITS["HOM"]
        if len(raw) == 2 and raw[0] != raw[1]:
            m |= ROC_BITS["HET"]
        t = ref_host.record_type(rec, raw)
        m |= ROC_BITS["SNP"] if t == "SNP" else ROC_BITS["NON_SNP"]
        if t == "MNP":
            m |= ROC_BITS["MNP"]
        if t in ("INSERTION", "DELETION", "INDEL"):
            m |= ROC_BITS["INDEL"]
        m |= ROC_BITS["XRX"] if rec.has_info("XRX") else ROC_BITS["NON_XRX"]
        masks.append(m)
        if score_field in (None, "QUAL"):
            quals.append(float("nan") if rec.qual == "." else float(rec.qual))
        else:      # RocScoreField.FORMAT :120-140 (MathUtils.approxEquals(val, 0, 1e-8) is inclusive)
            fv = rec.sample_field(sample, score_field)
            x = float("nan") if fv is None or fv == "." else float(fv)
            quals.append(0.0 if -1e-8 <= x <= 1e-8 else x)
    return np.array(masks, np.uint32), np.array(quals, np.float64)


FIELDS = ("id", "phased", "status_in", "gt", "allele_off", "allele_start", "allele_end", "allele_null", "allele_nt_off", "nt")


def assert_same_side(got, want, what):
    for f in FIELDS:
        a, b = np.asarray(getattr(got, f)), np.asarray(getattr(want, f))
        assert a.shape == b.shape and np.array_equal(a, b), "%s: %s differs\n got %s\nwant %s" % (what, f, a[:40], b[:40])


def check_text(engine, text, name, template_len, sample, ploidy=2, pass_only=True, max_length=1000, what="", ref_overlap=False,
               eval_regions=None, region=None, score_field=None):
    """Compares one (text, sequence, sample); returns 'ok' or 'format-error' (both sides agree the text is invalid)."""
    try:
        want, skipped = expected_side(text, name, template_len, sample, ploidy, pass_only, max_length, ref_overlap, eval_regions, region, score_field)
    except (ref_host.VcfFormatError, ValueError, IndexError):
        try:
            engine.vcf_parse(text.encode(), name, template_len, sample, ploidy, pass_only, max_length, ref_overlap, eval_regions, region, score_field)
        except ValueError:
            return "format-error"
        raise AssertionError("%s: the reference loader rejects this text, the engine accepted it" % what)
    got, stats = engine.vcf_parse(text.encode(), name, template_len, sample, ploidy, pass_only, max_length, ref_overlap, eval_regions, region, score_field)
    assert_same_side(got, want, what)
    assert stats["skipped"] == skipped, "%s: skipped %d vs %d" % (what, stats["skipped"], skipped)
    if "roc_mask" in stats and want.roc_expect is not None:
        wm, wq = want.roc_expect
        gm, gq = stats["roc_mask"], stats["qual"]
        assert np.array_equal(gm & 0x7fffffff, wm), "%s: RocFilter bits differ at %s\n got %s\nwant %s" % (
            what, np.nonzero((gm & 0x7fffffff) != wm)[0][:5], gm[:20], wm[:20])
        parsed = (gm >> 31) == 0     # the others are left to the caller's parser (bit 31)
        assert np.array_equal(gq[parsed].view(np.int64), wq[parsed].view(np.int64)), "%s: QUAL differs" % what
        assert np.isnan(gq[~parsed]).all()
    assert stats["kept"] == want.n
    return "ok"

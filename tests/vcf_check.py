"""Shared checker of the VCF loader (vce_vcf_parse): the arrays an engine returns for (text, sequence, sample) against the
loader restated in oracle/ref_host.py (VcfSortRefiner + VcfRecordTabixCallable filters + SampleVariants factory +
Allele.getAlleles), without --ref-overlap trimming and evaluation regions, which stay on the host."""
import numpy as np

import ref_host
from rtg_tools_b200 import capi


def expected_side(text, name, template_len, sample, ploidy=2, pass_only=True, max_length=1000):
    """(VariantSide, skipped count) or the exception class the loader ends with."""
    _, recs = ref_host.parse_vcf(text)
    recs = [r for r in recs if r.chrom == name]

    def factory(rec, vid):
        return ref_host.sample_variant(rec, vid, sample, ploidy, pass_only)
    out, skipped = ref_host.load_side(recs, template_len, factory, max_length, False)
    return capi.VariantSide.from_variants(out, ploidy), skipped


FIELDS = ("id", "phased", "status_in", "gt", "allele_off", "allele_start", "allele_end", "allele_null", "allele_nt_off", "nt")


def assert_same_side(got, want, what):
    for f in FIELDS:
        a, b = np.asarray(getattr(got, f)), np.asarray(getattr(want, f))
        assert a.shape == b.shape and np.array_equal(a, b), "%s: %s differs\n got %s\nwant %s" % (what, f, a[:40], b[:40])


def check_text(engine, text, name, template_len, sample, ploidy=2, pass_only=True, max_length=1000, what=""):
    """Compares one (text, sequence, sample); returns 'ok' or 'format-error' (both sides agree the text is invalid)."""
    try:
        want, skipped = expected_side(text, name, template_len, sample, ploidy, pass_only, max_length)
    except (ref_host.VcfFormatError, ValueError, IndexError):
        try:
            engine.vcf_parse(text.encode(), name, template_len, sample, ploidy, pass_only, max_length)
        except ValueError:
            return "format-error"
        raise AssertionError("%s: the reference loader rejects this text, the engine accepted it" % what)
    got, stats = engine.vcf_parse(text.encode(), name, template_len, sample, ploidy, pass_only, max_length)
    assert_same_side(got, want, what)
    assert stats["skipped"] == skipped, "%s: skipped %d vs %d" % (what, stats["skipped"], skipped)
    assert stats["kept"] == want.n
    return "ok"

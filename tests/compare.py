"""Field-by-field comparison of two engines' results on the same batch (bit-exact, no tolerances)."""
import numpy as np

from rtg_tools_b200 import capi


def compare_results(a, b, seqs, what="result"):
    assert len(a) == len(b)
    for k, (ra, rb) in enumerate(zip(a, b)):
        tag = "%s seq %d (%s)" % (what, k, seqs[k].name)
        assert ra.error == rb.error, "%s: error %s vs %s" % (tag, ra.error, rb.error)
        if ra.error:
            continue
        for side in ("baseline", "calls"):
            sa, sb = getattr(ra, side), getattr(rb, side)
            if not np.array_equal(sa.status, sb.status):
                i = int(np.nonzero(sa.status != sb.status)[0][0])
                raise AssertionError("%s %s status differs at %d: %d vs %d (n diff %d)" % (
                    tag, side, i, sa.status[i], sb.status[i], int((sa.status != sb.status).sum())))
            m = (sa.status & (capi.STATUS_GT_MATCH | capi.STATUS_ALLELE_MATCH)) != 0
            assert np.array_equal(sa.allele_ids[m], sb.allele_ids[m]), "%s %s allele ids differ" % (tag, side)
            assert np.array_equal(sa.original[m], sb.original[m]), "%s %s isOriginal differs" % (tag, side)
            # weights compared as raw bit patterns
            assert np.array_equal(sa.weight.view(np.int64), sb.weight.view(np.int64)), "%s %s weights differ" % (tag, side)
            assert np.array_equal(sa.sync_id, sb.sync_id), "%s %s sync ids differ" % (tag, side)
        for p in range(2):
            sa, sb = ra.sync_points[p], rb.sync_points[p]
            if not np.array_equal(sa, sb):
                n = min(len(sa), len(sb))
                d = np.nonzero(sa[:n] != sb[:n])[0]
                i = int(d[0]) if len(d) else n
                raise AssertionError("%s sync points pass %d differ at index %d (lengths %d / %d):\n%s\n%s" % (
                    tag, p, i, len(sa), len(sb), sa[max(0, i - 3):i + 4], sb[max(0, i - 3):i + 4]))
        assert ra.phasing == rb.phasing, "%s phasing %s vs %s" % (tag, ra.phasing, rb.phasing)
        assert ra.warnings == rb.warnings, "%s warnings %s vs %s" % (tag, ra.warnings, rb.warnings)

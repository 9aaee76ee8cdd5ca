"""ORACLE (test infrastructure, not product code): Python restatement of the HOST side of
`rtg vcfeval` that sits either side of the matching engine, so that the reference's own
end-to-end fixtures (plain-text FASTA + VCF in, VCF / TSV / TXT out) can be replayed.

It restates, for small inputs only:
  * VCF text parsing + VcfSortRefiner            (src/com/rtg/vcf/VcfSortRefiner.java:46-110)
  * loader filters                               (E/VcfRecordTabixCallable.java:91-141)
  * VariantFactory.SampleVariants / AllAlts      (E/VariantFactory.java:95-240)
  * Allele.getAlleles / padding removal          (E/Allele.java:145-192, vcf/VcfUtils.java:785-828)
  * Variant.trimAlleles (--ref-overlap)          (E/Variant.java:93-190, util/ByteUtils.java:93-141)
  * orientor / pass selection                    (E/VcfEvalTask.java:108-128,195-208, E/VcfEvalCli.java:304-371)
  * status -> TP/FP/FN mapping + ROC tuples      (E/SplitEvalSynchronizer.java:112-156,
                                                  E/WithInfoEvalSynchronizer.java:121-222,
                                                  E/WithRocsEvalSynchronizer.java:118-141)
  * ROC / summary formatting                     (E/RocContainer.java:267-410, util/Utils.java:454-470,
                                                  util/ContingencyTable.java:139-166)
E/ = src/com/rtg/vcf/eval/ under /root/reference.

The matching itself is delegated to an `engine` (rtg_tools_b200.capi.Library): the C++ oracle
when pinning the oracle on the goldens, the CUDA library when checking the product.
"""
import math
import os
import sys
from decimal import Decimal, ROUND_HALF_UP

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtg_tools_b200 import capi  # noqa: E402

BASE_CODE = {"N": 0, "A": 1, "C": 2, "G": 3, "T": 4}


class VcfFormatError(Exception):
    pass


class SkippedVariant(Exception):
    pass


# ------------------------------------------------------------------------------ parsing
def parse_fasta(text):
    seqs = []
    name, buf = None, []
    for line in text.splitlines():
        if line.startswith(">"):
            if name is not None:
                seqs.append((name, "".join(buf)))
            name = line[1:].split()[0]
            buf = []
        else:
            buf.append(line.strip())
    if name is not None:
        seqs.append((name, "".join(buf)))
    out = []
    for nm, s in seqs:
        arr = np.array([BASE_CODE.get(c.upper(), 0) for c in s], np.uint8)
        out.append((nm, arr))
    return out


class VcfRecord:
    def __init__(self, line):
        f = line.rstrip("\n").split("\t")
        self.line = line.rstrip("\n")
        self.chrom = f[0]
        self.start = int(f[1]) - 1
        self.ident = f[2]
        self.ref = f[3]
        self.alts = [] if f[4] == "." else f[4].split(",")
        self.qual = f[5]
        self.filters = [] if f[6] == "." else f[6].split(";")
        self.info = f[7] if len(f) > 7 else "."
        self.format = f[8].split(":") if len(f) > 8 else []
        self.samples = [s.split(":") for s in f[9:]]
        self.end = self.start + len(self.ref)  # VcfRecord.getEnd

    def is_filtered(self):  # VcfRecord.java:273-280
        return any(x not in ("PASS", ".") for x in self.filters)

    def sample_field(self, sample, key):
        if key not in self.format:
            return None
        k = self.format.index(key)
        s = self.samples[sample]
        return s[k] if k < len(s) else "."

    def is_sample_filtered(self, sample):  # VcfRecord.java:286-297
        ft = self.sample_field(sample, "FT")
        if ft is None:
            return False
        return any(x not in ("PASS", ".") for x in ft.split(";"))

    def has_info(self, key):
        return any(kv.split("=")[0] == key for kv in self.info.split(";"))

    def info_value(self, key):
        for kv in self.info.split(";"):
            p = kv.split("=", 1)
            if p[0] == key and len(p) == 2:
                return p[1]
        return None


def parse_vcf(text):
    samples = []
    records = []
    for line in text.splitlines():
        if line.startswith("##") or not line.strip():
            continue
        if line.startswith("#"):
            samples = line.split("\t")[9:]
            continue
        records.append(VcfRecord(line))
    return samples, records


class _JavaPriorityQueue:
    """java.util.PriorityQueue sift rules (ties are NOT stable; mimic the heap exactly)."""

    def __init__(self, key):
        self.q = []
        self.key = key

    def offer(self, x):
        q = self.q
        q.append(x)
        k = len(q) - 1
        while k > 0:
            parent = (k - 1) >> 1
            e = q[parent]
            if self.key(x) >= self.key(e):
                break
            q[k] = e
            k = parent
        q[k] = x

    def poll(self):
        q = self.q
        result = q[0]
        x = q.pop()
        n = len(q)
        if n > 0:
            k = 0
            half = n >> 1
            while k < half:
                child = 2 * k + 1
                c = q[child]
                right = child + 1
                if right < n and self.key(c) > self.key(q[right]):
                    child = right
                    c = q[child]
                if self.key(x) <= self.key(c):
                    break
                q[k] = c
                k = child
            q[k] = x
        return result

    def __len__(self):
        return len(self.q)


def sort_refine(records):
    """VcfSortRefiner: records at the same start are emitted shortest-first (vcf/VcfSortRefiner.java:91-108)."""
    out = []
    i = 0
    n = len(records)
    while i < n:
        j = i
        pq = _JavaPriorityQueue(lambda r: (r.start, r.end))
        while j < n and records[j].start == records[i].start and records[j].chrom == records[i].chrom:
            pq.offer(records[j])
            j += 1
        while len(pq):
            out.append(pq.poll())
        i = j
    return out


# ------------------------------------------------------------------------------ VCF helpers
def split_gt(gt):  # VcfUtils.splitGt :221-257
    parts = gt.replace("|", "/").split("/")
    out = []
    for p in parts:
        if p == ".":
            out.append(-1)
        else:
            out.append(int(p))
    return out


def symbolic_type(allele):  # VariantType.getSymbolicAlleleType :107-122
    if len(allele) > 0:
        first, last = allele[0], allele[-1]
        if first == "<" or last == ">":
            return "SV_SYMBOLIC"
        if first == "*" or last == "*":
            return "SV_MISSING"
        if first in "[]" or last in "[]":
            return "SV_BREAKEND"
    return None


_ORD = ["NO_CALL", "UNCHANGED", "SNP", "MNP", "DELETION", "INSERTION", "INDEL", "SV_BREAKEND", "SV_SYMBOLIC", "SV_MISSING"]
_INDEL = {"DELETION", "INSERTION", "INDEL"}
_SV = {"SV_BREAKEND", "SV_SYMBOLIC", "SV_MISSING"}
_NONVAR = {"NO_CALL", "UNCHANGED"}


def _is_ins_or_del(ref, pred):  # VariantType.isInsertionOrDeletion :193-219
    if len(ref) == 0 or len(pred) == 0:
        return True
    is_insert = len(ref) < len(pred)
    s1, s2 = (ref, pred) if is_insert else (pred, ref)
    left = 0
    while left < len(s1) and s1[left] == s2[left]:
        left += 1
    if left == len(s1):
        return True
    right = 0
    while right < len(s1) and s1[len(s1) - right - 1] == s2[len(s2) - right - 1]:
        right += 1
    return left + right >= len(s1)


def allele_type(ref, alt):  # VariantType.getType(String,String) :169-190
    if ref == alt:
        return "UNCHANGED"
    if len(ref) == 1 and len(alt) == 1 and alt[0] != "*":
        return "SNP"
    sv = symbolic_type(alt)
    if sv is not None:
        return sv
    if len(ref) == len(alt):
        return "MNP"
    if _is_ins_or_del(ref, alt):
        return "INSERTION" if len(ref) < len(alt) else "DELETION"
    return "INDEL"


def _precedence(a, b):  # VariantType.getPrecedence :85-98
    if a != b and a in _INDEL and b in _INDEL:
        return "INDEL"
    return a if _ORD.index(a) > _ORD.index(b) else b


def has_redundant_first_nt(rec):  # VcfUtils.hasRedundantFirstNucleotide :785-798
    c = rec.ref[0] if rec.ref else "."
    has_alt = False
    for alt in rec.alts:
        if alt[0] != "*":
            has_alt = True
            if alt[0] != c:
                return False
    return has_alt


def allele_strings(rec):  # VcfUtils.getAlleleStrings :806-828
    prev = has_redundant_first_nt(rec)
    out = [rec.ref[1:] if prev else rec.ref]
    for i, a in enumerate(rec.alts):
        s = a[1:] if (prev and a[0] != "*") else a
        if s == out[0]:
            raise VcfFormatError("VCF record ALT allele with ID %d is the same as REF, record: %s" % (i + 1, rec.line))
        out.append(s)
    return out


def record_type(rec, gt):  # VariantType.getType(VcfRecord,int[]) :130-160
    all_missing = True
    alt_id = 0
    pos = 0
    while pos < len(gt):
        a = gt[pos]
        pos += 1
        if a != -1:
            all_missing = False
            if a != 0:
                alt_id = a
                break
    if all_missing:
        return "NO_CALL"
    if alt_id == 0:
        return "UNCHANGED"
    al = allele_strings(rec)
    t = allele_type(al[0], al[alt_id])
    while pos < len(gt):
        a = gt[pos]
        pos += 1
        if a > 0 and a != alt_id:
            t = _precedence(t, allele_type(al[0], al[a]))
    return t


def encode(s):  # DnaUtils.encodeString
    try:
        return bytes(BASE_CODE[c.upper()] for c in s)
    except KeyError as e:
        raise SkippedVariant("Invalid VCF allele. Illegal DNA character: %s" % e.args[0])


# ------------------------------------------------------------------------------ variants
def get_alleles(rec, gt, explicit_unknown=False):  # Allele.getAlleles :145-192
    pad = has_redundant_first_nt(rec)
    n = len(rec.alts) + 2
    out = [None] * n
    for i in range(-1, n - 1):
        if gt is None or i == 0 or i in gt:
            if i == -1:
                a = None
            else:
                alt = rec.ref if i == 0 else rec.alts[i - 1]
                if symbolic_type(alt) is not None:
                    a = None
                else:
                    a = (rec.start + (1 if pad else 0), rec.end, encode(alt[1:] if pad else alt))
            if a is None and explicit_unknown:
                a = (rec.start + (1 if pad else 0), rec.end, bytes([5]))
            out[i + 1] = a
    return out


def defined_variant_gt(rec, sample, ploidy):  # VariantFactory.getDefinedVariantGt :163-196
    if sample >= len(rec.samples):
        raise VcfFormatError("Record did not contain enough samples: " + rec.line)
    if "GT" not in rec.format:
        return None, None
    gts = rec.sample_field(sample, "GT")
    gt = split_gt(gts)
    if any(a > len(rec.alts) for a in gt):
        raise VcfFormatError("VCF record GT contains allele ID out of range, record: " + rec.line)
    if ploidy == 2 and len(gt) == 1:
        gt = [gt[0], gt[0]]
    if len(gt) != ploidy:
        raise SkippedVariant("Ploidy of GT value '%s' is unexpected." % gts)
    al = allele_strings(rec)
    for a in gt:
        if a > 0:
            t = allele_type(al[0], al[a])
            if t not in _NONVAR and t not in _SV:
                return gt, gts
    return None, None


def sample_variant(rec, vid, sample, ploidy, pass_only, explicit_unknown=False):  # SampleVariants.variant :127-147
    if pass_only and (rec.is_filtered() or rec.is_sample_filtered(sample)):
        return None
    gt, gts = defined_variant_gt(rec, sample, ploidy)
    if gt is None:
        return None
    alleles = get_alleles(rec, gt, explicit_unknown)
    return {"id": vid, "rec": rec, "alleles": alleles, "gt": gt, "phased": "|" in gts, "status": 0}


def all_alts_variant(rec, vid, pass_only, explicit_unknown=False):  # AllAlts.variant :222-239
    if (pass_only and rec.is_filtered()) or not rec.alts:
        return None
    alleles = get_alleles(rec, None, explicit_unknown)
    pruned = alleles[:2] + [a for a in alleles[2:] if a is not None and a[2] != bytes([5])]  # pruneEmptyAlts :207-218
    if len(pruned) < 3:
        return None
    return {"id": vid, "rec": rec, "alleles": pruned, "gt": None, "phased": False, "status": 0}


def bounds(alleles):  # Allele.getAlleleBounds :195-206
    st = min(a[0] for a in alleles if a is not None)
    en = max(a[1] for a in alleles if a is not None)
    return st, en


def _longest_prefix(offset, a, b):  # util/ByteUtils.java longestPrefix
    n = min(len(a), len(b)) - offset
    i = 0
    while i < n and a[i] == b[i]:
        i += 1
    return i


def _longest_suffix(offset, a, b):  # util/ByteUtils.java longestSuffix
    n = min(len(a), len(b)) - offset
    i = 0
    while i < n and a[len(a) - 1 - i] == b[len(b) - 1 - i]:
        i += 1
    return i


def _trim_one(v, left_first=True):  # Variant.trimAlleles(boolean) :97-122
    al = v["alleles"]
    ref = al[1][2]
    al[1] = None
    for i in range(2, len(al)):
        a = al[i]
        if a is not None and a[2] != bytes([5]):
            alt = a[2]
            if left_first:
                lead = _longest_prefix(0, ref, alt)
                trail = _longest_suffix(lead, ref, alt)
            else:
                trail = _longest_suffix(0, ref, alt)
                lead = _longest_prefix(trail, ref, alt)
            if lead > 0 or trail > 0:
                al[i] = (a[0] + lead, a[1] - trail, alt[lead:len(alt) - trail])


def _overlaps(a, b):  # SequenceNameLocusSimple.overlaps
    (s1, e1), (s2, e2) = a, b
    return s1 < e2 and s2 < e1 or (s1 == e1 and s2 <= s1 < e2) or (s2 == e2 and s1 <= s2 < e1)


def trim_alleles(variants):  # Variant.trimAlleles(List) :124-190
    n = len(variants)
    locus = [bounds(v["alleles"]) for v in variants]
    first = 0
    last = 0
    for vi in range(n):
        cur = variants[vi]
        while last < n and (last <= vi or _range_overlap(locus[vi], locus[last])):
            last += 1
        while first < vi and not _range_overlap(locus[vi], locus[first]):
            first += 1
        if first == vi and last == vi + 1:
            _trim_one(cur)
            locus[vi] = bounds(cur["alleles"])
            continue
        lead = 0
        trail = 0
        ref = cur["alleles"][1][2]
        for i in range(2, len(cur["alleles"])):
            a = cur["alleles"][i]
            if a is not None and a[2] != bytes([5]):
                lead = max(lead, _longest_prefix(0, ref, a[2]))
                trail = max(trail, _longest_suffix(0, ref, a[2]))
        if lead == 0 or trail == 0:
            _trim_one(cur)
            locus[vi] = bounds(cur["alleles"])
            continue
        lc = rc = 0
        for i in range(first, last):
            if i == vi:
                continue
            lo = locus[i][1] - locus[vi][0]
            ro = locus[vi][1] - locus[i][0]
            if lo <= 0 and ro <= 0:
                continue
            if 0 < lo <= lead:
                lc += 1
            if 0 < ro < trail:
                rc += 1
        _trim_one(cur, lc > rc)
        locus[vi] = bounds(cur["alleles"])


def _range_overlap(a, b):
    # util/intervals/SequenceNameLocusSimple.overlaps -> Range semantics: other.start < end && other.end > start
    return b[0] < a[1] and b[1] > a[0]


def load_side(records, template_len, factory, max_length=1000, ref_overlap=False, eval_regions=None):
    """VcfRecordTabixCallable.call :91-141.  `factory(rec, id)` -> variant dict or None."""
    out = []
    vid = 0
    skipped = 0
    for rec in sort_refine(records):
        vid += 1
        length = max([len(rec.ref)] + [len(a) for a in rec.alts])
        if max_length > -1 and length > max_length:
            skipped += 1
            continue
        if template_len >= 0 and rec.end > template_len:
            skipped += 1
            continue
        try:
            v = factory(rec, vid)
        except SkippedVariant:
            skipped += 1
            continue
        if v is None:
            continue
        if eval_regions is not None:
            st, en = bounds(v["alleles"])
            if not any(s < en and e > st for (s, e) in eval_regions):
                v["status"] |= capi.STATUS_OUTSIDE_EVAL
        out.append(v)
    if ref_overlap:
        trim_alleles(out)
    return out, skipped


# ------------------------------------------------------------------------------ CLI restatement
class Params:
    """The subset of VcfEvalCli flags the fixtures use (E/VcfEvalCli.java:141-194,319-373)."""

    def __init__(self, args):
        self.sample = None
        self.score_field = "GQ"
        self.squash = False
        self.ref_overlap = False
        self.output_mode = "split"
        self.all_records = False
        self.obey_phase = "false"
        self.sample_ploidy = 2
        self.region = None
        self.no_roc = False
        self.two_pass = None
        self.hap_allele_matching = True
        self.roc_subsets = []
        self.max_length = 1000
        self.xx = {}
        i = 0
        args = list(args)

        def val(a, i):
            if "=" in a:
                return a.split("=", 1)[1], i
            return args[i + 1], i + 1
        while i < len(args):
            a = args[i]
            key = a.split("=", 1)[0]
            if key == "--sample":
                self.sample, i = val(a, i)
            elif key in ("--vcf-score-field", "-f"):
                self.score_field, i = val(a, i)
            elif key == "--squash-ploidy":
                self.squash = True
            elif key == "--ref-overlap":
                self.ref_overlap = True
            elif key in ("--output-mode", "-m"):
                self.output_mode, i = val(a, i)
            elif key == "--all-records":
                self.all_records = True
            elif key == "--Xobey-phase":
                self.obey_phase, i = val(a, i)
            elif key == "--sample-ploidy":
                v, i = val(a, i)
                self.sample_ploidy = int(v)
            elif key == "--region":
                self.region, i = val(a, i)
            elif key == "--no-roc":
                self.no_roc = True
            elif key == "--Xtwo-pass":
                v, i = val(a, i)
                self.two_pass = v == "true"
            elif key == "--roc-subset":
                v, i = val(a, i)
                self.roc_subsets.append(v)
            elif key in ("-T", "--threads"):
                _, i = val(a, i)
            elif key.startswith("--XX"):
                k, v = a[4:].split("=", 1)
                self.xx[k] = v
            else:
                raise ValueError("unsupported flag " + a)
            i += 1
        if "com.rtg.vcf.eval.haploid-allele-matching" in self.xx:
            self.hap_allele_matching = self.xx["com.rtg.vcf.eval.haploid-allele-matching"] == "true"
        if self.two_pass is None:  # VcfEvalCli.java:365-371
            self.two_pass = self.output_mode != "split" and not self.squash
        if self.squash:
            self.two_pass = False


def _phase_orientor(name):  # VcfEvalCli.phaseTypeToOrientor :305-314
    return {"true": capi.ORIENT_PHASED, "invert": capi.ORIENT_PHASED_INVERT}.get(name, capi.ORIENT_UNPHASED)


def _get_orientor(factory_name, allele_matching, phase_orientor, hap_am):  # VcfEvalTask.getOrientor :195-208
    hap = allele_matching and hap_am
    dip = allele_matching and not hap_am
    if factory_name == "sample":
        o = capi.ORIENT_ALLELE_GT if dip else capi.ORIENT_SQUASH_GT if hap else phase_orientor
    else:
        o = capi.ORIENT_DIPLOID_POP if dip else capi.ORIENT_HAPLOID_POP if hap else capi.ORIENT_ALT
    return o


def _haplotypes(orientor, sample_ploidy):
    if orientor in (capi.ORIENT_SQUASH_GT, capi.ORIENT_HAPLOID_POP):
        return 1
    if orientor in (capi.ORIENT_ALLELE_GT, capi.ORIENT_DIPLOID_POP):
        return 2
    return sample_ploidy


def build_config(p, bfactory, cfactory, **over):
    phase = p.obey_phase.split(",")
    if len(phase) == 1:
        phase = [phase[0], phase[0]]
    bpo, cpo = _phase_orientor(phase[0]), _phase_orientor(phase[1])
    if p.two_pass:
        bo = [_get_orientor(bfactory, False, bpo, p.hap_allele_matching), _get_orientor(bfactory, True, bpo, p.hap_allele_matching)]
        co = [_get_orientor(cfactory, False, cpo, p.hap_allele_matching), _get_orientor(cfactory, True, cpo, p.hap_allele_matching)]
        npass = 2
    else:
        bo = [_get_orientor(bfactory, p.squash, bpo, p.hap_allele_matching)] * 2
        co = [_get_orientor(cfactory, p.squash, cpo, p.hap_allele_matching)] * 2
        npass = 1
    ploidy = [_haplotypes(bo[0], p.sample_ploidy), _haplotypes(bo[1], p.sample_ploidy)]
    cfg = capi.Config(n_passes=npass, baseline_orientor=tuple(bo), calls_orientor=tuple(co), ploidy=tuple(ploidy))
    if "com.rtg.vcf.eval.max-paths" in p.xx:
        cfg.max_paths = int(p.xx["com.rtg.vcf.eval.max-paths"])
    if "com.rtg.vcf.eval.max-iterations" in p.xx:
        cfg.max_iterations = int(p.xx["com.rtg.vcf.eval.max-iterations"])
    if p.xx.get("com.rtg.vcf.eval.maximize") == "calls-min-base":
        cfg.preference = capi.PREFER_CALLS_MIN_BASE
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


# ------------------------------------------------------------------------------ formatting
def real_format(x, dp):
    """Utils.realFormat (util/Utils.java:454-470): String.format("%.<dp>f") = HALF_UP on the shortest decimal
    representation of the double, and "-0.00" printed as "0.00"."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    d = Decimal(repr(float(x))).quantize(Decimal(1).scaleb(-dp), rounding=ROUND_HALF_UP)
    s = "{:f}".format(d)
    if x <= 0.0 and s.startswith("-") and float(s) == 0.0:
        s = s[1:]
    return s


def precision(tp, fp):  # ContingencyTable.precision
    return tp / (tp + fp)


def recall(tp, fn):  # ContingencyTable.recall
    return tp / (tp + fn)


def f_measure(p, r):  # ContingencyTable.fMeasure
    return 2 * p * r / (p + r) if (p + r) != 0 else float("nan")


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a == 0 or math.isnan(a):
            return float("nan")
        return math.copysign(float("inf"), a)


def _fmeasure(p, r):  # ContingencyTable.fMeasure :162-165
    s = p + r
    return 0.0 if s == 0 else 2 * p * r / s


class RocResult:
    def __init__(self):
        self.lines = []
        self.best = None


def format_roc(rows, total_baseline, extra_metrics, score_label, filter_name="ALL"):
    """RocContainer.writeRocs/rocHeader/writeRocLine (:267-361) for an un-rescaled filter, headers sanitised the
    way TestUtils.sanitizeTsvHeader does (#Version / #CL lines masked)."""
    thr, ctp, cfp, craw = rows
    total_calls = int(round_half_up((craw[-1] + cfp[-1]) if len(thr) else 0.0))
    out = ["#Version [...]", "#CL [...]", "#selection: " + filter_name, "#total baseline variants: %d" % total_baseline,
           "#total call variants: %d" % total_calls, "#score field: " + score_label]
    hdr = "#score\ttrue_positives_baseline\tfalse_positives\ttrue_positives_call"
    if extra_metrics:
        hdr += "\tfalse_negatives\tprecision\tsensitivity\tf_measure"
    out.append(hdr)
    best = None  # FMeasureThreshold.java:38-50 keeps the LAST maximum (>=), mBestFMeasure starts at 0
    best_f = 0.0

    def emit(score, k):
        nonlocal best, best_f
        tp, fp, raw = ctp[k], cfp[k], craw[k]
        line = "%s\t%s\t%s\t%s" % (score, real_format(tp, 2), real_format(fp, 2), real_format(raw, 2))
        if extra_metrics:
            fn = total_baseline - tp
            p = _div(raw, fp + raw)
            r = _div(tp, tp + fn)
            f = _fmeasure(p, r)
            line += "\t%s\t%s\t%s\t%s" % (real_format(fn, 2), real_format(p, 4), real_format(r, 4), real_format(f, 4))
            if not math.isnan(thr[k]) and (best is None or f >= best_f):
                best_f = f
                best = (thr[k], tp, fp, raw)
        out.append(line)
    prev = None
    prev_k = None
    for k in range(len(thr)):
        score = "None" if math.isnan(thr[k]) else real_format(thr[k], 3)
        if prev is not None and score != prev:
            emit(prev, prev_k)
        prev = score
        prev_k = k
    if prev is not None:
        emit(prev, prev_k)
    return "\n".join(out) + "\n", best


def round_half_up(x):  # Math.round
    return math.floor(x + 0.5)


def format_summary(best, total_tp, total_fp, total_raw, total_baseline, roc_enabled=True):
    """RocContainer.writeSummary (:382-410) with util/TextTable right-aligned columns."""
    if total_baseline <= 0:
        return "0 total baseline variants, no summary statistics available\n"
    rows = [["Threshold", "True-pos-baseline", "True-pos-call", "False-pos", "False-neg", "Precision", "Sensitivity", "F-measure"]]

    def row(label, tp, fn, raw, fp):
        p = _div(raw, fp + raw)
        r = _div(tp, tp + fn)
        f = _fmeasure(p, r)
        return [label, str(int(round_half_up(tp))), str(int(round_half_up(raw))), str(int(round_half_up(fp))),
                str(int(round_half_up(fn))), real_format(p, 4), real_format(r, 4), real_format(f, 4)]
    body = []
    if roc_enabled and best is not None:
        body.append(row(real_format(best[0], 3), best[1], total_baseline - best[1], best[3], best[2]))
    body.append(row("None", total_tp, total_baseline - total_tp, total_raw, total_fp))
    widths = [max(len(r[i]) for r in rows + body) for i in range(8)]
    lines = []

    def fmt(r):
        return "  ".join(r[i].rjust(widths[i]) for i in range(8))
    lines.append(fmt(rows[0]))
    lines.append("-" * len(lines[0]))
    for r in body:
        lines.append(fmt(r))
    return "\n".join(lines) + "\n"


# ------------------------------------------------------------------------------ whole-run driver
def _restrict(records, region):
    """--region chr:start-end (1-based inclusive) -> records overlapping it (SamRangeUtils explicit range)."""
    if region is None:
        return records
    chrom, rng = region.split(":")
    s, e = rng.split("-")
    s0, e0 = int(s) - 1, int(e)
    return [r for r in records if r.chrom == chrom and r.start < e0 and max(r.end, r.start + 1) > s0]


def score_of(rec, field, sample):
    if field == "QUAL":  # RocScoreField.QUAL :50-80
        return float("nan") if rec.qual == "." else float(rec.qual)
    # FORMAT field (default GQ): VcfUtils.getDoubleFormatFieldFromRecord :550-566, RocScoreField.FORMAT :111-140
    v = rec.sample_field(sample, field) if sample >= 0 else None
    if v is None or v == ".":
        return float("nan")
    x = float(v)
    return 0.0 if abs(x) < 1e-8 else x


class RunResult:
    pass


def run_vcfeval(engine, template_fa, baseline_vcf, calls_vcf, args, cfg_over=None, side_loader=None):
    """Replays `rtg vcfeval` on in-memory text inputs.  Returns a RunResult with per-record outcomes and the
    regenerated output files the fixtures check.  `side_loader(vcf_text, name, template_len, sample, ploidy, pass_only, max_length,
    ref_overlap, region) -> VariantSide` (optional): the arrays handed to the engine come from it (the device loaders under test) instead
    of from this module's own loader; the arrays of both loaders
    must agree, and everything downstream then runs on the loader's output."""
    p = Params(args)
    seqs = parse_fasta(template_fa)
    bsamples, brecs = parse_vcf(baseline_vcf)
    csamples, crecs = parse_vcf(calls_vcf)
    brecs = _restrict(brecs, p.region)
    crecs = _restrict(crecs, p.region)

    # sample selection (E/VcfEvalCli / VariantSetType; "ALT" = all-alts factory)
    if p.sample is None:
        bs_name = cs_name = None
    else:
        parts = p.sample.split(",")
        bs_name, cs_name = (parts[0], parts[0]) if len(parts) == 1 else (parts[0], parts[1])

    def pick(name, samples):
        if name == "ALT":
            return -1
        if name is None:
            if len(samples) != 1:
                raise ValueError("need --sample")
            return 0
        return samples.index(name)
    bsample = pick(bs_name, bsamples)
    csample = pick(cs_name, csamples)
    bfac = "all" if bsample < 0 else "sample"
    cfac = "all" if csample < 0 else "sample"
    cfg = build_config(p, bfac, cfac, **(cfg_over or {}))
    pass_only = not p.all_records

    def factory(kind, sample):
        if kind == "all":
            return lambda rec, vid: all_alts_variant(rec, vid, pass_only)
        return lambda rec, vid: sample_variant(rec, vid, sample, p.sample_ploidy, pass_only)

    rr = RunResult()
    rr.params = p
    rr.cfg = cfg
    rr.seq_names = [name for name, _ in seqs]
    rr.brecs, rr.crecs = brecs, crecs
    rr.bsample = bsample
    rr.sequences = []
    batch = []
    loaded = []
    for name, tmpl in seqs:
        bl, _ = load_side([r for r in brecs if r.chrom == name], len(tmpl), factory(bfac, bsample), p.max_length, p.ref_overlap)
        cl, _ = load_side([r for r in crecs if r.chrom == name], len(tmpl), factory(cfac, csample), p.max_length, p.ref_overlap)
        gp = p.sample_ploidy
        bside = capi.VariantSide.from_variants(bl, gp if bfac == "sample" else 0)
        cside = capi.VariantSide.from_variants(cl, gp if cfac == "sample" else 0)
        if side_loader is not None:
            fields = ("id", "phased", "status_in", "gt", "allele_off", "allele_start", "allele_end", "allele_null", "allele_nt_off", "nt")
            # (sample -1 = the all-alts factory of `--sample ALT`)
            reg = None
            if p.region is not None:   # --region chr:start-end (1-based inclusive): (sequence, 0-based begin, end exclusive)
                rc, rrng = p.region.split(":")
                reg = (rc, int(rrng.split("-")[0]) - 1, int(rrng.split("-")[1]))
            got = side_loader(baseline_vcf, name, len(tmpl), bsample, gp, pass_only, p.max_length, p.ref_overlap, reg)
            assert all(np.array_equal(np.asarray(getattr(got, f)), np.asarray(getattr(bside, f))) for f in fields), "baseline arrays of %s differ" % name
            bside = got
            got = side_loader(calls_vcf, name, len(tmpl), csample, gp, pass_only, p.max_length, p.ref_overlap, reg)
            assert all(np.array_equal(np.asarray(getattr(got, f)), np.asarray(getattr(cside, f))) for f in fields), "call arrays of %s differ" % name
            cside = got
            rr.device_loaded = getattr(rr, "device_loaded", 0) + 2
        batch.append(capi.Sequence(name, tmpl, bside, cside))
        loaded.append((bl, cl))
    results = engine.submit(cfg, batch)
    rr.batch = batch
    rr.results = results
    rr.loaded = loaded
    rr.warnings = []
    for s, r in zip(batch, results):
        if r.error:
            raise RuntimeError("engine error %s on %s" % (capi.ERR_NAMES.get(r.error, r.error), s.name))
        for w in r.warnings:
            rr.warnings.append(
                "Evaluation too complex (%d unresolved paths, %d iterations) at reference region %s:%d-%d. "
                "Variants in this region will not be included in results." % (w["paths"], w["iterations"], s.name, w["start1"], w["end1"]))

    # ---- output stage: classification per record + ROC tuples (arrival order = sequence order, then record order)
    rr.call_class = []      # (rec, class, weight, sync) for known calls, in order
    rr.base_class = []
    roc = {"score": [], "tp": [], "fp": [], "raw": [], "snp": []}
    base_total = 0
    base_tp = 0
    call_outside = 0
    S = capi
    annotate = p.output_mode in ("annotate", "combine", "ga4gh")
    for (bl, cl), r in zip(loaded, results):
        for i, v in enumerate(cl):
            st = int(r.calls.status[i])
            w = float(r.calls.weight[i])
            sync = int(r.calls.sync_id[i])
            rec = v["rec"]
            cls = None
            tup = None
            if annotate:  # WithInfoEvalSynchronizer.updateForCall :121-181
                if st & S.STATUS_SKIPPED:
                    cls = "HARD"
                    sync = None
                elif st & S.STATUS_GT_MATCH:
                    if w > 0:
                        cls, tup = "TP", (w, 0.0, 1.0, False)
                    else:
                        cls = "OUT"
                elif st & S.STATUS_ALLELE_MATCH:
                    if w > 0:
                        cls, tup = "FP_CA", (w, 0.0, 1.0, True)
                    else:
                        cls = "OUT"
                elif st & S.STATUS_OUTSIDE_EVAL:
                    cls = "OUT"
                    sync = None
                    call_outside += 1
                elif st & S.STATUS_NO_MATCH:
                    cls, tup = "FP", (0.0, 1.0, 0.0, False)
            else:  # SplitEvalSynchronizer.handleKnownCall :112-134
                sync = None
                if st & S.STATUS_GT_MATCH:
                    if w > 0:
                        cls, tup = "TP", (w, 0.0, 1.0, False)
                elif st & S.STATUS_ALLELE_MATCH:
                    if w > 0:
                        cls, tup = "FP_CA", (w, 0.0, 1.0, True)
                elif st & S.STATUS_NO_MATCH:
                    if not st & S.STATUS_OUTSIDE_EVAL:
                        cls, tup = "FP", (0.0, 1.0, 0.0, False)
                    else:
                        call_outside += 1
            rr.call_class.append((rec, cls, w, sync, st))
            if tup is not None:
                tpw, fpw, raw, allele_match = tup
                if allele_match:  # WithRocsEvalSynchronizer.addToROCContainer :132-141
                    tpw, fpw, raw = 0.0, 1.0, 0.0
                roc["score"].append(score_of(rec, p.score_field, csample))
                roc["tp"].append(tpw)
                roc["fp"].append(fpw)
                roc["raw"].append(raw)
                gtv = v["gt"] if v["gt"] is not None else []
                roc["snp"].append(record_type(rec, gtv) == "SNP" if csample >= 0 else False)
        for i, v in enumerate(bl):
            st = int(r.baseline.status[i])
            sync = int(r.baseline.sync_id[i])
            cls = None
            if annotate:  # updateForBaseline :183-222
                if st & S.STATUS_SKIPPED:
                    cls, sync = "HARD", None
                elif st & S.STATUS_OUTSIDE_EVAL:
                    cls, sync = "OUT", None
                elif st & S.STATUS_GT_MATCH:
                    cls = "TP"
                elif st & S.STATUS_ALLELE_MATCH:
                    cls = "FN_CA"
                elif st & S.STATUS_NO_MATCH:
                    cls = "FN"
            else:  # SplitEvalSynchronizer.handleKnownBaseline :136-150
                sync = None
                if not st & S.STATUS_OUTSIDE_EVAL:
                    if st & S.STATUS_GT_MATCH:
                        cls = "TP"
                    elif st & S.STATUS_ALLELE_MATCH:
                        cls = "FN_CA"
                    elif st & S.STATUS_NO_MATCH:
                        cls = "FN"
            if cls in ("TP", "FN_CA", "FN"):
                base_total += 1
                if cls == "TP":
                    base_tp += 1
            rr.base_class.append((v["rec"], cls, sync, st))
    rr.base_total = base_total
    rr.base_tp = base_tp
    rr.phasing = [sum(r.phasing[k] for r in results) for k in range(3)]
    rr.roc_inputs = roc
    rr.csample = csample

    # ---- ROC + summary (filter ALL)
    n = len(roc["score"])
    if n:
        rows, no_score = engine.roc(np.array(roc["score"]), np.array(roc["tp"]), np.array(roc["fp"]), np.array(roc["raw"]),
                                    None, 1, False)
        rows = rows[0]
    else:
        rows, no_score = (np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0)), 0
    label = "QUAL" if p.score_field == "QUAL" else p.score_field + " (FORMAT)"
    rr.roc_tsv, best = format_roc(rows, base_total, base_total > 0, label)
    tot = (rows[1][-1], rows[2][-1], rows[3][-1]) if len(rows[0]) else (0.0, 0.0, 0.0)
    rr.summary = format_summary(best, tot[0], tot[1], tot[2], base_total, roc_enabled=not p.no_roc)
    rr.phasing_txt = "Correct phasings: %d\nIncorrect phasings: %d\nUnresolvable phasings: %d\n" % (
        rr.phasing[1], rr.phasing[0], rr.phasing[2])
    return rr


# ------------------------------------------------------------------------------ combined output (--output-mode combine)
def _join_gt(phased, gt):  # VcfUtils.joinGt :267-286
    if not gt:
        return "."
    return ("|" if phased else "/").join("." if g < 0 else str(g) for g in gt)


def _normalize_allele(a):  # VcfUtils.normalizeAllele :478-505 (nucleotide alleles to upper case)
    return a.upper() if a and all(c in "acgtnACGTN" for c in a) else a


def _gt_str(rec, sample):  # CombinedEvalSynchronizer.resetRecordFields :168-183 (VcfUtils.getValidGtStr, or missing without a sample)
    if sample < 0:
        return "."
    v = rec.sample_field(sample, "GT")
    return "." if v is None else v


def _clean_alts(alts, gts):  # VcfAltCleaner.annotate :57-94
    used = [False] * len(alts)
    for gt in gts:
        for a in split_gt(gt):
            if 0 < a <= len(used):
                used[a - 1] = True
    if all(used):
        return alts, gts
    remap = [0] * (len(used) + 1)
    out = []
    for k, u in enumerate(used):
        if u:
            out.append(alts[k])
            remap[k + 1] = len(out)
    new_gts = []
    for gt in gts:
        sp = [remap[a] if 0 < a < len(remap) else a for a in split_gt(gt)]
        new_gts.append(_join_gt("|" in gt, sp))
    return out, new_gts


def _call_info(item, info):  # WithInfoEvalSynchronizer.updateForCall :121-181 (item None = record without a variant)
    if item is None:
        info["CALL"] = "IGN"
        return
    _, cls, w, sync, st = item
    if sync is not None:
        sy = str(sync)
        old = info.get("SYNC")
        info["SYNC"] = sy if old is None or old == sy else old + "," + sy
    if cls in ("TP", "FP_CA") and abs(w - 1.0) > 0.001:
        info["CALL_WEIGHT"] = format(w, "#.3g")
    info["CALL"] = cls
    if cls == "FP" and st & capi.STATUS_ANY_MATCH:
        info["CALL_ALTERNATE"] = None


def _base_info(item, info):  # WithInfoEvalSynchronizer.updateForBaseline :183-226
    if item is None:
        info["BASE"] = "IGN"
        return
    _, cls, sync, st = item
    if sync is not None:
        sy = str(sync)
        old = info.get("SYNC")
        info["SYNC"] = sy if old is None or old == sy else sy + "," + old
    info["BASE"] = cls
    if cls == "FN" and st & capi.STATUS_ANY_MATCH:
        info["BASE_ALTERNATE"] = None


def format_combine(rr):
    """Body lines of output.vcf as CombinedEvalSynchronizer writes them (E/CombinedEvalSynchronizer.java:104-166 over the
    interleaving of E/InterleavingEvalSynchronizer.java:86-200): both record streams in (start, end) order, records with
    the same span merged into one two-sample row (VcfRecordMerger.mergeRecordsWithSameRef), unused ALTs removed when both
    sides are samples (VcfAltCleaner)."""
    clean = rr.bsample >= 0 and rr.csample >= 0
    known_b = {id(it[0]): it for it in rr.base_class}
    known_c = {id(it[0]): it for it in rr.call_class}
    lines = []

    def emit(chrom, start, ref, alts, info, gts):
        if clean:
            alts, gts = _clean_alts(alts, gts)
        inf = ";".join(k if v is None else "%s=%s" % (k, v) for k, v in info.items()) or "."
        lines.append("\t".join([chrom, str(start + 1), ".", ref, ",".join(alts) if alts else ".", ".", ".", inf, "GT"] + gts))

    for chrom in rr.seq_names:
        bs = sort_refine([r for r in rr.brecs if r.chrom == chrom])
        cs = sort_refine([r for r in rr.crecs if r.chrom == chrom])
        i = j = 0
        while i < len(bs) or j < len(cs):
            b = bs[i] if i < len(bs) else None
            c = cs[j] if j < len(cs) else None
            if b is not None and c is not None:
                order = (b.start > c.start) - (b.start < c.start) or (b.end > c.end) - (b.end < c.end)
            else:
                order = -1 if c is None else 1
            info = {}
            if order < 0:
                _base_info(known_b.get(id(b)), info)
                emit(chrom, b.start, b.ref, list(b.alts), info, [_gt_str(b, rr.bsample), "."])
                i += 1
            elif order > 0:
                _call_info(known_c.get(id(c)), info)
                emit(chrom, c.start, c.ref, list(c.alts), info, [".", _gt_str(c, rr.csample)])
                j += 1
            else:
                _base_info(known_b.get(id(b)), info)
                _call_info(known_c.get(id(c)), info)
                # VcfRecordMerger.AlleleMap :128-160: union of the ALTs, allele codes remapped
                ref = _normalize_allele(b.ref)
                alts = []
                gts = []
                for rec, sample in ((b, rr.bsample), (c, rr.csample)):
                    m = [0]
                    for a in rec.alts:
                        a = _normalize_allele(a)
                        if a == ref:
                            m.append(0)
                        else:
                            if a not in alts:
                                alts.append(a)
                            m.append(alts.index(a) + 1)
                    g = _gt_str(rec, sample)
                    gts.append(_join_gt("|" in g, [m[x] if x >= 0 else x for x in split_gt(g)]))
                emit(chrom, b.start, ref, alts, info, gts)
                i += 1
                j += 1
    return lines


# ------------------------------------------------------------------------------ GA4GH intermediate output (--output-mode ga4gh)
def _java_double(x):  # Double.toString for the values a score takes
    if x == int(x) and abs(x) < 1e7:
        return "%.1f" % x
    r = repr(float(x))
    if "e" in r:
        m, e = r.split("e")
        if "." not in m:
            m += ".0"
        return "%sE%d" % (m, int(e))
    return r


class _OutRec:
    def __init__(self, chrom, start, ref, alts, filters):
        self.chrom, self.start, self.ref, self.alts, self.filters = chrom, start, ref, alts, filters
        self.end = start + len(ref)
        self.bs = []
        self.fmt = {"GT": [".", "."]}   # insertion-ordered like the LinkedHashMap of VcfRecord.java:104

    def set(self, key, val, sample):  # VcfRecord.setFormatAndSample :530-541
        self.fmt.setdefault(key, [".", "."])[sample] = val

    def line(self):  # VcfRecord.toString :586-613, trailing missing sub-fields omitted (:658-680)
        cols = []
        for k in range(2):
            vals = [v[k] for v in self.fmt.values()]
            while len(vals) > 1 and vals[-1] == ".":
                vals.pop()
            cols.append(":".join(vals))
        info = "BS=" + ",".join(self.bs) if self.bs else "."
        return "\t".join([self.chrom, str(self.start + 1), ".", self.ref, ",".join(self.alts) if self.alts else ".", ".",
                          ";".join(self.filters) if self.filters else ".", info, ":".join(self.fmt)] + cols)


def format_ga4gh(rr, loose_distance=30):
    """Body lines of output.vcf as Ga4ghEvalSynchronizer writes them (E/Ga4ghEvalSynchronizer.java:139-321) behind
    Ga4ghLooseMatchFilter (E/Ga4ghLooseMatchFilter.java:68-150): TRUTH / QUERY columns with BD / BK / BI / QQ, BS = sync ids."""
    T, Q = 0, 1
    known_b = {id(it[0]): it for it in rr.base_class}
    known_c = {id(it[0]): it for it in rr.call_class}
    gt_of = {}
    for bl, cl in rr.loaded:
        for v in bl + cl:
            gt_of[id(v["rec"])] = v["gt"]
    S = capi

    def multi(rec):  # isMultiAllelicCall :262-271
        g = gt_of.get(id(rec))
        return g is not None and len(g) == 2 and g[0] > 0 and g[1] > 0 and g[0] != g[1]

    def set_sync(o, sync):  # :314-321
        if sync is not None and (not o.bs or o.bs[-1] != str(sync)):
            o.bs.append(str(sync))

    def for_call(item, o, rec):  # updateForCall :218-260
        if item is None:
            o.set("BD", "N", Q)
            return
        _, cls, w, sync, st = item

        def score():
            x = score_of(rec, rr.params.score_field, rr.csample)
            if x == x:
                o.set("QQ", _java_double(x), Q)
        if st & S.STATUS_SKIPPED:
            o.set("BD", "N", Q)
            o.set("BI", "too-hard", Q)
            sync = None
        elif st & S.STATUS_GT_MATCH:
            score()
            o.set("BD", "TP", Q)
            o.set("BK", "gm", Q)
        elif st & S.STATUS_ALLELE_MATCH:
            score()
            o.set("BD", "FP", Q)
            o.set("BK", "am", Q)
            if multi(rec):
                o.set("BI", "multi", Q)
        elif st & S.STATUS_OUTSIDE_EVAL:
            o.set("BD", "N", Q)
            o.set("BI", "non-confident", Q)
            sync = None
        else:
            score()
            o.set("BD", "FP", Q)
            o.set("BK", ".", Q)
        set_sync(o, sync)

    def for_base(item, o, rec):  # updateForBaseline :273-305
        if item is None:
            o.set("BD", "N", T)
            return
        _, cls, sync, st = item
        if st & S.STATUS_SKIPPED:
            o.set("BD", "N", T)
            o.set("BI", "too-hard", T)
            sync = None
        elif st & S.STATUS_OUTSIDE_EVAL:
            o.set("BD", "N", T)
            o.set("BI", "non-confident", T)
            sync = None
        elif st & S.STATUS_GT_MATCH:
            o.set("BD", "TP", T)
            o.set("BK", "gm", T)
        elif st & S.STATUS_ALLELE_MATCH:
            o.set("BD", "FN", T)
            o.set("BK", "am", T)
            if multi(rec):
                o.set("BI", "multi", T)
        else:
            o.set("BD", "FN", T)
            o.set("BK", ".", T)
        set_sync(o, sync)

    recs = []
    for chrom in rr.seq_names:
        bs = sort_refine([r for r in rr.brecs if r.chrom == chrom])
        cs = sort_refine([r for r in rr.crecs if r.chrom == chrom])
        i = j = 0
        while i < len(bs) or j < len(cs):
            b = bs[i] if i < len(bs) else None
            c = cs[j] if j < len(cs) else None
            if b is not None and c is not None:
                order = (b.start > c.start) - (b.start < c.start) or (b.end > c.end) - (b.end < c.end)
            else:
                order = -1 if c is None else 1
            if order < 0:
                o = _OutRec(chrom, b.start, b.ref, list(b.alts), [])
                o.fmt["GT"] = [_gt_str(b, rr.bsample), "."]
                for_base(known_b.get(id(b)), o, b)
                i += 1
            elif order > 0:
                o = _OutRec(chrom, c.start, c.ref, list(c.alts), list(c.filters))
                o.fmt["GT"] = [".", _gt_str(c, rr.csample)]
                for_call(known_c.get(id(c)), o, c)
                j += 1
            else:  # makeCombinedRecord :187-191: the query record merges first (field priority), columns stay TRUTH, QUERY
                ref = _normalize_allele(c.ref)
                alts = []
                gts = {}
                for rec, sample, col in ((c, rr.csample, Q), (b, rr.bsample, T)):
                    m = [0]
                    for a in rec.alts:
                        a = _normalize_allele(a)
                        if a == ref:
                            m.append(0)
                        else:
                            if a not in alts:
                                alts.append(a)
                            m.append(alts.index(a) + 1)
                    g = _gt_str(rec, sample)
                    gts[col] = _join_gt("|" in g, [m[x] if x >= 0 else x for x in split_gt(g)])
                o = _OutRec(chrom, c.start, ref, alts, list(c.filters))
                o.fmt["GT"] = [gts[T], gts[Q]]
                for_base(known_b.get(id(b)), o, b)
                for_call(known_c.get(id(c)), o, c)
                i += 1
                j += 1
            o.alts, o.fmt["GT"] = _clean_alts(o.alts, o.fmt["GT"])   # writeRecord :182-185
            recs.append(o)

    # ---- Ga4ghLooseMatchFilter: BK "." -> "lm" for FP / FN records with the other side's TP/FP/FN records nearby
    out = []
    if loose_distance < 0:
        return [o.line() for o in recs]
    buf, truth, query = [], [], []

    def wants(o, col, dec):
        return "BK" in o.fmt and "BD" in o.fmt and o.fmt["BD"][col] == dec and o.fmt["BK"][col] == "."

    def flush_first():
        o = buf.pop(0)
        while truth and truth[0].end <= o.start:
            truth.pop(0)
        while query and query[0].end <= o.start:
            query.pop(0)
        if wants(o, Q, "FP") and truth and truth[0].start - loose_distance < o.end:
            o.fmt["BK"][Q] = "lm"
        if wants(o, T, "FN") and query and query[0].start - loose_distance < o.end:
            o.fmt["BK"][T] = "lm"
        out.append(o.line())

    for o in recs:
        if buf:
            if buf[0].chrom != o.chrom:
                while buf:
                    flush_first()
                del truth[:], query[:]
            else:
                while buf and buf[0].end < o.start - loose_distance:
                    flush_first()
        if wants(o, Q, "FP") and truth and o.start < truth[-1].end + loose_distance:
            o.fmt["BK"][Q] = "lm"
        if wants(o, T, "FN") and query and o.start < query[-1].end + loose_distance:
            o.fmt["BK"][T] = "lm"
        buf.append(o)
        if "BD" in o.fmt:
            if o.fmt["BD"][Q] in ("TP", "FP"):
                query.append(o)
            if o.fmt["BD"][T] in ("TP", "FN"):
                truth.append(o)
    while buf:
        flush_first()
    return out

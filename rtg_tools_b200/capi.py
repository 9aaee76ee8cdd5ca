"""ctypes view of include/vce.h plus numpy containers for the flat batch format.

This is plumbing: it owns no algorithm.  `Library` wraps a shared object that
exports the vce C-ABI (`vce_submit`, `vce_roc`, ...).  The product library is
rtg_tools_b200/csrc/libvce.so; tests and bench.py's cpu_baseline leg also point
it at the oracle (oracle/liboracle.so, symbol prefix `oracle_`) so that both are
driven through byte-identical buffers.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence as Seq

import numpy as np

MAX_PLOIDY = 4

STATUS_SKIPPED = 0x01
STATUS_GT_MATCH = 0x02
STATUS_ALLELE_MATCH = 0x04
STATUS_NO_MATCH = 0x08
STATUS_OUTSIDE_EVAL = 0x10
STATUS_ANY_MATCH = 0x20

ORIENT_UNPHASED = 0
ORIENT_PHASED = 1
ORIENT_PHASED_INVERT = 2
ORIENT_SQUASH_GT = 3
ORIENT_ALLELE_GT = 4
ORIENT_HAPLOID_POP = 5
ORIENT_DIPLOID_POP = 6
ORIENT_ALT = 7

PREFER_SUM = 0
PREFER_CALLS_MIN_BASE = 1

ERR_NAMES = {
    0: "VCE_OK", 1: "VCE_ERR_NOT_INITIALISED", 2: "VCE_ERR_BAD_ARGUMENT", 3: "VCE_ERR_CUDA", 4: "VCE_ERR_NO_DEVICE",
    5: "VCE_ERR_CAPACITY", 6: "VCE_ERR_BEST_PATH_NULL", 7: "VCE_ERR_MOVE_IN_VARIANT", 8: "VCE_ERR_UNSUPPORTED",
}

_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)


class CVariants(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("gt_ploidy", C.c_int32),
        ("id", _i32p), ("phased", _u8p), ("status_in", _u8p), ("gt", _i8p),
        ("allele_off", _i32p), ("n_slots", C.c_int32),
        ("allele_start", _i32p), ("allele_end", _i32p), ("allele_null", _u8p),
        ("allele_nt_off", _i32p), ("nt", _u8p),
    ]


class CVcfIn(C.Structure):
    _fields_ = [
        ("text", C.c_char_p), ("text_len", C.c_int64), ("name", C.c_char_p), ("template_len", C.c_int32),
        ("sample", C.c_int32), ("ploidy", C.c_int32), ("pass_only", C.c_int32), ("max_length", C.c_int32),
        ("ref_overlap", C.c_int32), ("n_eval_regions", C.c_int32), ("region_begin", C.c_int32), ("region_end", C.c_int32),
        ("eval_regions", C.POINTER(C.c_int32)), ("score_field", C.c_char_p),
    ]


class CVcfStats(C.Structure):
    _fields_ = [
        ("lines", C.c_int64), ("records", C.c_int64), ("kept", C.c_int64), ("skipped", C.c_int64),
        ("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
        ("launches", C.c_int32), ("reserved", C.c_int32),
    ]


class CBgzfStats(C.Structure):
    _fields_ = [
        ("blocks", C.c_int64), ("compressed_bytes", C.c_int64), ("text_bytes", C.c_int64),
        ("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
        ("launches", C.c_int32), ("reserved", C.c_int32),
    ]


class CSequence(C.Structure):
    _fields_ = [
        ("name", C.c_char_p), ("template_bases", _u8p), ("template_len", C.c_int32),
        ("baseline", CVariants), ("calls", CVariants),
    ]


class CConfig(C.Structure):
    _fields_ = [
        ("n_passes", C.c_int32), ("baseline_orientor", C.c_int32 * 2), ("calls_orientor", C.c_int32 * 2),
        ("ploidy", C.c_int32 * 2), ("max_paths", C.c_int32), ("max_iterations", C.c_int32),
        ("prune_no_ops", C.c_int32), ("flag_alternates", C.c_int32), ("preference", C.c_int32),
    ]


class CWarning(C.Structure):
    _fields_ = [("pass_", C.c_int32), ("paths", C.c_int32), ("iterations", C.c_int32),
                ("start1", C.c_int32), ("end1", C.c_int32)]


class CSideResult(C.Structure):
    _fields_ = [("status", _u8p), ("allele_ids", _i8p), ("original", _u8p), ("weight", _f64p), ("sync_id", _i32p)]


class CSequenceResult(C.Structure):
    _fields_ = [
        ("baseline", CSideResult), ("calls", CSideResult),
        ("sync_points", _i32p * 2), ("sync_cap", C.c_int32 * 2), ("n_sync", C.c_int32 * 2),
        ("phasing", C.c_int32 * 3),
        ("warnings", C.POINTER(CWarning)), ("warn_cap", C.c_int32), ("n_warnings", C.c_int32),
        ("error", C.c_int32), ("n_regions", C.c_int64), ("n_iterations", C.c_int64),
    ]


class CRocIn(C.Structure):
    _fields_ = [("n", C.c_int64), ("score", _f64p), ("tp", _f64p), ("fp", _f64p), ("tp_raw", _f64p),
                ("filter_mask", _u32p), ("n_filters", C.c_int32), ("ascending", C.c_int32)]


class CRocOut(C.Structure):
    _fields_ = [("cap", C.c_int64), ("n_rows", _i64p), ("threshold", _f64p), ("cum_tp", _f64p), ("cum_fp", _f64p),
                ("cum_tp_raw", _f64p), ("no_score", _i64p)]


def _ptr(a: Optional[np.ndarray], typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


# --------------------------------------------------------------------------- containers
@dataclass
class VariantSide:
    """One side of one sequence in load order: the SoA image of List<Variant> (vce_variants)."""
    id: np.ndarray
    phased: np.ndarray
    status_in: np.ndarray
    gt: np.ndarray            # (n, gt_ploidy) int8
    allele_off: np.ndarray    # n+1
    allele_start: np.ndarray
    allele_end: np.ndarray
    allele_null: np.ndarray
    allele_nt_off: np.ndarray
    nt: np.ndarray

    @property
    def n(self) -> int:
        return int(self.id.shape[0])

    @property
    def gt_ploidy(self) -> int:
        return int(self.gt.shape[1]) if self.gt.ndim == 2 else 0

    @staticmethod
    def empty(gt_ploidy: int = 2) -> "VariantSide":
        return VariantSide.from_variants([], gt_ploidy)

    @staticmethod
    def from_variants(variants, gt_ploidy: int = 2) -> "VariantSide":
        """variants: iterable of dicts {id, phased, status, gt (list or None), alleles: [None | (start, end, bytes)]}
        where alleles[k] is allele id k-1 (slot 0 = missing, slot 1 = REF), cf. Variant.java:241."""
        variants = list(variants)
        n = len(variants)
        ids = np.zeros(n, np.int32)
        phased = np.zeros(n, np.uint8)
        status = np.zeros(n, np.uint8)
        gt = np.full((n, gt_ploidy), -2, np.int8)
        off = np.zeros(n + 1, np.int32)
        a_start, a_end, a_null, nt_off, nts = [], [], [], [0], []
        total = 0
        for i, v in enumerate(variants):
            ids[i] = v["id"]
            phased[i] = 1 if v.get("phased") else 0
            status[i] = v.get("status", 0)
            if v.get("gt") is not None and gt_ploidy:
                g = list(v["gt"])
                assert len(g) == gt_ploidy, (g, gt_ploidy)
                gt[i, :] = g
            for a in v["alleles"]:
                if a is None:
                    a_start.append(0)
                    a_end.append(0)
                    a_null.append(1)
                else:
                    a_start.append(a[0])
                    a_end.append(a[1])
                    a_null.append(0)
                    nts.append(bytes(a[2]))
                    total += len(a[2])
                nt_off.append(total)
            off[i + 1] = len(a_null)
        nt = np.frombuffer(b"".join(nts), np.uint8).copy() if nts else np.zeros(0, np.uint8)
        return VariantSide(ids, phased, status, gt, off, np.asarray(a_start, np.int32), np.asarray(a_end, np.int32),
                           np.asarray(a_null, np.uint8), np.asarray(nt_off, np.int32), nt)

    def to_c(self) -> CVariants:
        cv = CVariants()
        cv.n = self.n
        cv.gt_ploidy = self.gt_ploidy
        self.gt = np.ascontiguousarray(self.gt, np.int8)
        cv.id = _ptr(self.id, _i32p)
        cv.phased = _ptr(self.phased, _u8p)
        cv.status_in = _ptr(self.status_in, _u8p)
        cv.gt = _ptr(self.gt, _i8p)
        cv.allele_off = _ptr(self.allele_off, _i32p)
        cv.n_slots = int(self.allele_off[-1])
        cv.allele_start = _ptr(self.allele_start, _i32p)
        cv.allele_end = _ptr(self.allele_end, _i32p)
        cv.allele_null = _ptr(self.allele_null, _u8p)
        cv.allele_nt_off = _ptr(self.allele_nt_off, _i32p)
        cv.nt = _ptr(self.nt, _u8p)
        return cv


@dataclass
class Sequence:
    name: str
    template: np.ndarray  # uint8 0..4
    baseline: VariantSide
    calls: VariantSide

    def to_c(self) -> CSequence:
        cs = CSequence()
        self._name_b = self.name.encode()
        cs.name = self._name_b
        self.template = np.ascontiguousarray(self.template, np.uint8)
        cs.template_bases = _ptr(self.template, _u8p)
        cs.template_len = int(self.template.shape[0])
        cs.baseline = self.baseline.to_c()
        cs.calls = self.calls.to_c()
        return cs


@dataclass
class Config:
    """PathFinder.Config (PathFinder.java:51-101) + the orientor pairs of VcfEvalTask.java:122-128."""
    n_passes: int = 1
    baseline_orientor: Seq[int] = (ORIENT_UNPHASED, ORIENT_SQUASH_GT)
    calls_orientor: Seq[int] = (ORIENT_UNPHASED, ORIENT_SQUASH_GT)
    ploidy: Seq[int] = (2, 1)
    max_paths: int = 50000
    max_iterations: int = 10000000
    prune_no_ops: bool = True
    flag_alternates: bool = False
    preference: int = PREFER_SUM

    def to_c(self) -> CConfig:
        c = CConfig()
        c.n_passes = self.n_passes
        for i in range(2):
            c.baseline_orientor[i] = self.baseline_orientor[i]
            c.calls_orientor[i] = self.calls_orientor[i]
            c.ploidy[i] = self.ploidy[i]
        c.max_paths = self.max_paths
        c.max_iterations = self.max_iterations
        c.prune_no_ops = int(self.prune_no_ops)
        c.flag_alternates = int(self.flag_alternates)
        c.preference = self.preference
        return c


@dataclass
class SideResult:
    status: np.ndarray
    allele_ids: np.ndarray  # (n, MAX_PLOIDY) int8
    original: np.ndarray
    weight: np.ndarray
    sync_id: np.ndarray

    @staticmethod
    def alloc(n: int) -> "SideResult":
        return SideResult(np.zeros(n, np.uint8), np.full((n, MAX_PLOIDY), -2, np.int8), np.zeros(n, np.uint8),
                          np.zeros(n, np.float64), np.zeros(n, np.int32))

    def to_c(self) -> CSideResult:
        r = CSideResult()
        r.status = _ptr(self.status, _u8p)
        r.allele_ids = _ptr(self.allele_ids, _i8p)
        r.original = _ptr(self.original, _u8p)
        r.weight = _ptr(self.weight, _f64p)
        r.sync_id = _ptr(self.sync_id, _i32p)
        return r


@dataclass
class SequenceResult:
    baseline: SideResult
    calls: SideResult
    sync_points: List[np.ndarray]
    phasing: List[int] = field(default_factory=lambda: [0, 0, 0])
    warnings: List[dict] = field(default_factory=list)
    error: int = 0
    n_regions: int = 0
    n_iterations: int = 0


WARN_CAP = 64


class Library:
    """A shared object exporting the vce C-ABI, optionally under a symbol prefix."""

    def __init__(self, path: str, prefix: str = "vce_"):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s is missing: build it first (python -c 'import __graft_entry__ as g; g.build()')" % path)
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        self._submit = self._sym("submit" if prefix == "vce_" else "vce_submit")
        self._submit.restype = C.c_int
        self._submit.argtypes = [C.POINTER(CConfig), C.c_int32, C.POINTER(CSequence), C.POINTER(CSequenceResult)]
        self._roc = self._sym("roc" if prefix == "vce_" else "vce_roc")
        self._roc.restype = C.c_int
        self._roc.argtypes = [C.POINTER(CRocIn), C.POINTER(CRocOut)]
        self._roc_timed = None
        if prefix == "vce_":
            try:
                self._roc_timed = self._sym("roc_timed")
                self._roc_timed.restype = C.c_int
                self._roc_timed.argtypes = [C.POINTER(CRocIn), C.POINTER(CRocOut), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(C.c_int32)]
            except AttributeError:
                self._roc_timed = None
        self.last_roc_timing = None

    def _sym(self, name):
        return getattr(self.lib, self.prefix + name)

    def has(self, name) -> bool:
        return hasattr(self.lib, self.prefix + name)

    def last_error(self) -> str:
        if self.prefix == "vce_":
            f = self.lib.vce_last_error
            f.restype = C.c_char_p
            s = f()
            return s.decode() if s else ""
        return ""

    # ---- batch packing shared by submit() and the staged job API
    @staticmethod
    def pack(cfg: Config, seqs: List[Sequence]):
        ccfg = cfg.to_c()
        n = len(seqs)
        cseqs = (CSequence * max(n, 1))()
        cres = (CSequenceResult * max(n, 1))()
        results = []
        keep = []
        for i, s in enumerate(seqs):
            cseqs[i] = s.to_c()
            r = SequenceResult(SideResult.alloc(s.baseline.n), SideResult.alloc(s.calls.n),
                               [np.zeros(s.baseline.n + s.calls.n + 2, np.int32) for _ in range(2)])
            warn = (CWarning * WARN_CAP)()
            cr = cres[i]
            cr.baseline = r.baseline.to_c()
            cr.calls = r.calls.to_c()
            for p in range(2):
                cr.sync_points[p] = _ptr(r.sync_points[p], _i32p)
                cr.sync_cap[p] = int(r.sync_points[p].shape[0])
            cr.warnings = warn
            cr.warn_cap = WARN_CAP
            keep.append(warn)
            results.append(r)
        return ccfg, cseqs, cres, results, keep

    @staticmethod
    def unpack(cres, results, keep):
        for i, r in enumerate(results):
            cr = cres[i]
            for p in range(2):
                r.sync_points[p] = r.sync_points[p][: cr.n_sync[p]].copy()
            r.phasing = [cr.phasing[0], cr.phasing[1], cr.phasing[2]]
            r.warnings = [dict(pass_=keep[i][k].pass_, paths=keep[i][k].paths, iterations=keep[i][k].iterations,
                               start1=keep[i][k].start1, end1=keep[i][k].end1)
                          for k in range(min(cr.n_warnings, WARN_CAP))]
            r.error = cr.error
            r.n_regions = cr.n_regions
            r.n_iterations = cr.n_iterations
        return results

    def submit(self, cfg: Config, seqs: List[Sequence], threads: int = 0) -> List[SequenceResult]:
        ccfg, cseqs, cres, results, keep = self.pack(cfg, seqs)
        if threads > 1 and self.has("vce_submit_mt"):
            f = self._sym("vce_submit_mt")
            f.restype = C.c_int
            f.argtypes = [C.POINTER(CConfig), C.c_int32, C.POINTER(CSequence), C.POINTER(CSequenceResult), C.c_int32]
            rc = f(C.byref(ccfg), len(seqs), cseqs, cres, threads)
        else:
            rc = self._submit(C.byref(ccfg), len(seqs), cseqs, cres)
        self.unpack(cres, results, keep)
        if rc != 0 and all(r.error == 0 for r in results):
            raise RuntimeError("%ssubmit failed: %s %s" % (self.prefix, ERR_NAMES.get(rc, rc), self.last_error()))
        return results

    def submit_split(self, cfg: Config, seqs: List[Sequence], n_dev: int, min_variants=100000, overlap=2048, gap=64, imbalance=1.1,
                     pieces_per_device=4, force=True):
        """Emulation only: vce_submit over `n_dev` emulated devices with the sequence splitting of vce_split.h.
        Returns (results, {"split", "pieces", "redone"})."""
        ccfg, cseqs, cres, results, keep = self.pack(cfg, seqs)
        f = self._sym("vce_submit_split")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(CConfig), C.c_int32, C.POINTER(CSequence), C.POINTER(CSequenceResult), C.c_int32,
                      C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        opts = (C.c_int64 * 6)(int(min_variants), int(overlap), int(gap), int(round(imbalance * 100)), int(pieces_per_device), int(bool(force)))
        st = (C.c_int64 * 3)()
        rc = f(C.byref(ccfg), len(seqs), cseqs, cres, int(n_dev), opts, st)
        self.unpack(cres, results, keep)
        if rc != 0 and all(r.error == 0 for r in results):
            raise RuntimeError("submit_split failed: %s" % ERR_NAMES.get(rc, rc))
        return results, {"split": int(st[0]), "pieces": int(st[1]), "redone": int(st[2])}

    def submit_packed(self, packed, threads: int = 0) -> int:
        """vce_submit on buffers packed once with Library.pack (the caller owns and reuses every input / output array,
        as a JNI host would); returns the C return code.  Call Library.unpack(cres, results, keep) when done."""
        ccfg, cseqs, cres, results, keep = packed
        if threads > 1 and self.has("vce_submit_mt"):
            f = self._sym("vce_submit_mt")
            f.restype = C.c_int
            f.argtypes = [C.POINTER(CConfig), C.c_int32, C.POINTER(CSequence), C.POINTER(CSequenceResult), C.c_int32]
            return f(C.byref(ccfg), len(results), cseqs, cres, threads)
        return self._submit(C.byref(ccfg), len(results), cseqs, cres)

    def _loader_error(self, rc, what):
        pre = "vce_" if self.prefix == "vce_" else self.prefix + "vce_"
        if self.prefix == "vce_":
            msg = self.last_error()
        else:
            e = getattr(self.lib, pre + "vcf_error")
            e.restype = C.c_char_p
            msg = e().decode()
        if rc == 2:   # VCE_ERR_BAD_ARGUMENT: a format error in the input
            raise ValueError(msg)
        raise RuntimeError("%s failed: %s %s" % (what, ERR_NAMES.get(rc, rc), msg))

    @staticmethod
    def _bytes_arg(data):
        """(keep-alive object, char pointer, length) of bytes or a uint8 array (e.g. one page-locked with vce_host_register)."""
        if isinstance(data, np.ndarray):
            data = np.ascontiguousarray(data, np.uint8)
            return data, C.cast(data.ctypes.data, C.c_char_p), int(data.nbytes)
        return data, C.cast(C.c_char_p(data), C.c_char_p), len(data)

    def _vcf_in(self, text, name, template_len, sample, ploidy, pass_only, max_length, ref_overlap=False, eval_regions=None, region=None, score_field=None):
        cin = CVcfIn()
        keep, cin.text, cin.text_len = self._bytes_arg(text)
        cin.name = name.encode() if isinstance(name, str) else name
        cin.template_len = int(template_len)
        cin.sample, cin.ploidy, cin.pass_only, cin.max_length = int(sample), int(ploidy), int(bool(pass_only)), int(max_length)
        cin.ref_overlap = int(bool(ref_overlap))
        cin.score_field = score_field.encode() if isinstance(score_field, str) else score_field
        cin.region_begin, cin.region_end = (0, -1) if region is None else (int(region[0]), int(region[1]))
        if eval_regions is not None:   # merged, ascending [start, end) pairs of this sequence
            reg = np.ascontiguousarray(np.asarray(eval_regions, np.int32).reshape(-1, 2))
            cin.n_eval_regions = int(reg.shape[0])
            cin.eval_regions = reg.ctypes.data_as(C.POINTER(C.c_int32)) if reg.size else (C.c_int32 * 2)()
            keep = (keep, reg)
        return cin, keep

    def _vcf_result(self, handle):
        pre = "vce_" if self.prefix == "vce_" else self.prefix + "vce_"
        f_get = getattr(self.lib, pre + "vcf_variants")
        f_free = getattr(self.lib, pre + "vcf_free")
        f_get.restype = C.c_int
        f_get.argtypes = [C.c_void_p, C.POINTER(CVariants), C.POINTER(CVcfStats)]
        f_free.restype = None
        f_free.argtypes = [C.c_void_p]
        try:
            cv, st = CVariants(), CVcfStats()
            rc = f_get(handle, C.byref(cv), C.byref(st))
            if rc != 0:
                raise RuntimeError("vcf_variants failed: %s" % ERR_NAMES.get(rc, rc))
            n, ns = int(cv.n), int(cv.n_slots)

            def arr(ptr, count, dtype):
                if count <= 0:
                    return np.zeros(0, dtype)
                return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)
            off = arr(cv.allele_off, n + 1, np.int32) if n >= 0 else np.zeros(1, np.int32)
            nt_off = arr(cv.allele_nt_off, ns + 1, np.int32)
            side = VariantSide(arr(cv.id, n, np.int32), arr(cv.phased, n, np.uint8), arr(cv.status_in, n, np.uint8),
                               arr(cv.gt, n * int(cv.gt_ploidy), np.int8).reshape(n, int(cv.gt_ploidy)), off,
                               arr(cv.allele_start, ns, np.int32), arr(cv.allele_end, ns, np.int32), arr(cv.allele_null, ns, np.uint8),
                               nt_off, arr(cv.nt, int(nt_off[-1]) if ns >= 0 and nt_off.size else 0, np.uint8))
            stats = {k: getattr(st, k) for k in ("lines", "records", "kept", "skipped", "h2d_ms", "kernel_ms", "d2h_ms", "launches")}
            # the ROC inputs of the same variants (vce_vcf_roc_fields): RocFilter bits and QUAL
            f_roc = getattr(self.lib, pre + "vcf_roc_fields", None)
            if f_roc is not None:
                f_roc.restype = C.c_int
                f_roc.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_double))]
                pm, pq = C.POINTER(C.c_uint32)(), C.POINTER(C.c_double)()
                if f_roc(handle, C.byref(pm), C.byref(pq)) == 0:
                    stats["roc_mask"] = arr(pm, n, np.uint32)
                    stats["qual"] = arr(pq, n, np.float64)
        finally:
            f_free(handle)
        return side, stats

    def vcf_parse(self, text: bytes, name=None, template_len=-1, sample=0, ploidy=2, pass_only=True, max_length=1000, ref_overlap=False,
                  eval_regions=None, region=None, score_field=None):
        """vce_vcf_parse: the records of sequence `name` in VCF `text` as a VariantSide (copies) plus the loader's counters
        {"lines", "records", "kept", "skipped", "h2d_ms", "kernel_ms", "d2h_ms", "launches"}.  Raises ValueError on a record
        the reference would throw VcfFormatException for."""
        pre = "vce_" if self.prefix == "vce_" else self.prefix + "vce_"
        f_parse = getattr(self.lib, pre + "vcf_parse")
        f_parse.restype = C.c_int
        f_parse.argtypes = [C.POINTER(CVcfIn), C.POINTER(C.c_void_p)]
        cin, keep = self._vcf_in(text, name, template_len, sample, ploidy, pass_only, max_length, ref_overlap, eval_regions, region, score_field)
        handle = C.c_void_p()
        rc = f_parse(C.byref(cin), C.byref(handle))
        del keep
        if rc != 0:
            self._loader_error(rc, "vcf_parse")
        return self._vcf_result(handle)

    # ---- on-disk formats: BGZF members, tabix index, SDF sequence data
    def bgzf_text_size(self, data, voffset_begin=0, voffset_end=-1):
        pre = "vce_" if self.prefix == "vce_" else self.prefix + "vce_"
        f = getattr(self.lib, pre + "bgzf_text_size")
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        keep, ptr, n = self._bytes_arg(data)
        out = C.c_int64(0)
        rc = f(ptr, n, int(voffset_begin), int(voffset_end), C.byref(out))
        if rc != 0:
            self._loader_error(rc, "bgzf_text_size")
        return int(out.value)

    def bgzf_inflate(self, data, voffset_begin=0, voffset_end=-1, check_crc=True):
        """vce_bgzf_inflate: the text between two virtual offsets of a BGZF file (bytes), plus the decoder's counters."""
        pre = "vce_" if self.prefix == "vce_" else self.prefix + "vce_"
        size = self.bgzf_text_size(data, voffset_begin, voffset_end)
        f = getattr(self.lib, pre + "bgzf_inflate")
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _u8p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(CBgzfStats)]
        keep, ptr, n = self._bytes_arg(data)
        out = np.zeros(max(size, 1), np.uint8)
        got, st = C.c_int64(0), CBgzfStats()
        rc = f(ptr, n, int(voffset_begin), int(voffset_end), int(bool(check_crc)), out.ctypes.data_as(_u8p), size, C.byref(got), C.byref(st))
        if rc != 0:
            self._loader_error(rc, "bgzf_inflate")
        stats = {k: getattr(st, k) for k in ("blocks", "compressed_bytes", "text_bytes", "h2d_ms", "kernel_ms", "d2h_ms", "launches")}
        return out[:int(got.value)].tobytes(), stats

    def vcf_parse_bgzf(self, data, voffset_begin=0, voffset_end=-1, name=None, template_len=-1, sample=0, ploidy=2, pass_only=True,
                       max_length=1000, check_crc=True, ref_overlap=False, eval_regions=None, region=None, score_field=None):
        """vce_vcf_parse_bgzf: vcf_parse on the bytes of a .vcf.gz, between two virtual offsets (e.g. from tabix_open(...).query)."""
        pre = "vce_" if self.prefix == "vce_" else self.prefix + "vce_"
        f = getattr(self.lib, pre + "vcf_parse_bgzf")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(CVcfIn), C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(CBgzfStats)]
        cin, keep = self._vcf_in(data, name, template_len, sample, ploidy, pass_only, max_length, ref_overlap, eval_regions, region, score_field)
        handle, st = C.c_void_p(), CBgzfStats()
        rc = f(C.byref(cin), int(voffset_begin), int(voffset_end), int(bool(check_crc)), C.byref(handle), C.byref(st))
        del keep
        if rc != 0:
            self._loader_error(rc, "vcf_parse_bgzf")
        side, stats = self._vcf_result(handle)
        for k in ("blocks", "compressed_bytes", "text_bytes"):
            stats[k] = getattr(st, k)
        if self.prefix == "vce_":
            for k in ("h2d_ms", "kernel_ms", "d2h_ms", "launches"):
                stats[k] = getattr(st, k)
        return side, stats

    def tabix_open(self, tbi: bytes):
        """vce_tabix_open: a TabixIndex (names, options, query) over the bytes of a .tbi file."""
        return TabixIndex(self, tbi)

    def template_register_sdf(self, seqdata, first_residue, length):
        """vce_template_register_sdf (product library): decodes one sequence of an SDF seqdata file on the device, keeps the 2-bit
        copy resident and returns the byte-per-base array (the template_bases to hand to submit; keep it alive)."""
        f = self._sym("template_register_sdf")
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int32, _u8p, C.POINTER(C.c_double)]
        keep, ptr, n = self._bytes_arg(seqdata)
        bases = np.zeros(int(length), np.uint8)
        ms = (C.c_double * 3)()
        rc = f(ptr, n, int(first_residue), int(length), bases.ctypes.data_as(_u8p), ms)
        if rc != 0:
            self._loader_error(rc, "template_register_sdf")
        self.last_sdf_ms = tuple(ms)
        return bases

    def sdf_decode(self, seqdata, first_residue, length):
        """Emulation only: (bases, packed words, N mask) made by the SDF granule / mask pass bodies."""
        f = getattr(self.lib, self.prefix + "vce_sdf_decode")
        f.restype = C.c_int
        _u32p = C.POINTER(C.c_uint32)
        f.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int32, _u8p, _u32p, _u32p]
        keep, ptr, n = self._bytes_arg(seqdata)
        length = int(length)
        words = (length + 15) // 16
        bases, packed, nmask = np.zeros(length, np.uint8), np.zeros(words + 1, np.uint32), np.zeros((words + 31) // 32 + 1, np.uint32)
        rc = f(ptr, n, int(first_residue), length, bases.ctypes.data_as(_u8p), packed.ctypes.data_as(_u32p), nmask.ctypes.data_as(_u32p))
        if rc != 0:
            raise ValueError("sdf_decode: %s" % ERR_NAMES.get(rc, rc))
        return bases, packed[:words], nmask[:(words + 31) // 32]

    def pack_template(self, bases):
        """Emulation only: packTemplate (the host packing vce_template_register uses) for comparison with sdf_decode."""
        f = getattr(self.lib, self.prefix + "vce_pack_template")
        f.restype = C.c_int
        _u32p = C.POINTER(C.c_uint32)
        f.argtypes = [_u8p, C.c_int32, _u32p, _u32p]
        bases = np.ascontiguousarray(bases, np.uint8)
        words = (bases.size + 15) // 16
        packed, nmask = np.zeros(words + 1, np.uint32), np.zeros((words + 31) // 32 + 1, np.uint32)
        f(bases.ctypes.data_as(_u8p), int(bases.size), packed.ctypes.data_as(_u32p), nmask.ctypes.data_as(_u32p))
        return packed[:words], nmask[:(words + 31) // 32]

    def roc(self, score, tp, fp, tp_raw, filter_mask=None, n_filters=1, ascending=False, timed=False):
        """Returns (rows per filter: list of (threshold, cum_tp, cum_fp, cum_tp_raw) arrays, no_score count).
        timed=True (product library only) goes through vce_roc_timed and leaves {"h2d_ms", "kernel_ms", "launches"} (CUDA
        events on the ROC stream) in `last_roc_timing`."""
        score = np.ascontiguousarray(score, np.float64)
        tp = np.ascontiguousarray(tp, np.float64)
        fp = np.ascontiguousarray(fp, np.float64)
        tp_raw = np.ascontiguousarray(tp_raw, np.float64)
        n = int(score.shape[0])
        cin = CRocIn()
        cin.n = n
        cin.score = _ptr(score, _f64p)
        cin.tp = _ptr(tp, _f64p)
        cin.fp = _ptr(fp, _f64p)
        cin.tp_raw = _ptr(tp_raw, _f64p)
        if filter_mask is not None:
            filter_mask = np.ascontiguousarray(filter_mask, np.uint32)
        cin.filter_mask = _ptr(filter_mask, _u32p)
        cin.n_filters = n_filters
        cin.ascending = int(ascending)
        cap = max(n, 1)
        n_rows = np.zeros(n_filters, np.int64)
        thr = np.zeros(n_filters * cap, np.float64)
        ctp = np.zeros(n_filters * cap, np.float64)
        cfp = np.zeros(n_filters * cap, np.float64)
        craw = np.zeros(n_filters * cap, np.float64)
        no_score = np.zeros(1, np.int64)
        cout = CRocOut()
        cout.cap = cap
        cout.n_rows = _ptr(n_rows, _i64p)
        cout.threshold = _ptr(thr, _f64p)
        cout.cum_tp = _ptr(ctp, _f64p)
        cout.cum_fp = _ptr(cfp, _f64p)
        cout.cum_tp_raw = _ptr(craw, _f64p)
        cout.no_score = _ptr(no_score, _i64p)
        if timed and self._roc_timed is not None:
            h2d, ker, nl = C.c_double(0), C.c_double(0), C.c_int32(0)
            rc = self._roc_timed(C.byref(cin), C.byref(cout), C.byref(h2d), C.byref(ker), C.byref(nl))
            self.last_roc_timing = {"h2d_ms": h2d.value, "kernel_ms": ker.value, "launches": int(nl.value)}
        else:
            rc = self._roc(C.byref(cin), C.byref(cout))
        if rc != 0:
            raise RuntimeError("%sroc failed: %s %s" % (self.prefix, ERR_NAMES.get(rc, rc), self.last_error()))
        rows = []
        for f in range(n_filters):
            k = int(n_rows[f])
            s = f * cap
            rows.append((thr[s:s + k].copy(), ctp[s:s + k].copy(), cfp[s:s + k].copy(), craw[s:s + k].copy()))
        return rows, int(no_score[0])


class TabixIndex:
    """TabixIndexReader (src/com/rtg/tabix/TabixIndexReader.java) over vce_tabix_*: `.names`, `.options` (format, seq / begin / end
    column 0-based, meta, skip) and `.query(name, begin0, end0)` -> (voffset_begin, voffset_end) or None."""

    def __init__(self, library, tbi):
        self._lib = library
        pre = "vce_" if library.prefix == "vce_" else library.prefix + "vce_"
        self._pre = pre
        f = getattr(library.lib, pre + "tabix_open")
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]
        self._h = C.c_void_p()
        rc = f(tbi, len(tbi), C.byref(self._h))
        if rc != 0:
            self._h = None
            library._loader_error(rc, "tabix_open")
        info = (C.c_int32 * 7)()
        g = getattr(library.lib, pre + "tabix_info")
        g.restype = C.c_int
        g.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        g(self._h, info)
        self.options = {"format": info[1], "seq_col": info[2], "beg_col": info[3], "end_col": info[4], "meta": info[5], "skip": info[6]}
        nm = getattr(library.lib, pre + "tabix_name")
        nm.restype = C.c_char_p
        nm.argtypes = [C.c_void_p, C.c_int32]
        self.names = [nm(self._h, i).decode() for i in range(info[0])]

    def query(self, name, begin0=-1, end0=-1):
        q = getattr(self._lib.lib, self._pre + "tabix_query")
        q.restype = C.c_int
        q.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        vb, ve = C.c_int64(-1), C.c_int64(-1)
        rc = q(self._h, name.encode(), int(begin0), int(end0), C.byref(vb), C.byref(ve))
        if rc != 0:
            self._lib._loader_error(rc, "tabix_query")
        return None if vb.value < 0 else (int(vb.value), int(ve.value))

    def close(self):
        if self._h:
            c = getattr(self._lib.lib, self._pre + "tabix_close")
            c.restype = None
            c.argtypes = [C.c_void_p]
            c(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


"""Host-side mirror of the reference's interface for the vcfeval matching path, on top of the CUDA library.

Reference classes mirrored (src/com/rtg/vcf/eval/ of RealTimeGenomics/rtg-tools):
  SequenceEvaluator.run()      -> Engine.submit() / SequenceEvaluator.evaluate()   (SequenceEvaluator.java:76-158)
  PathFinder(...).bestPath()   -> PathFinder(...).best_path()                      (PathFinder.java:106-214)
  PathFinder.Config            -> capi.Config                                      (PathFinder.java:51-101)
  RocContainer                 -> RocContainer.add_roc_line / rows                 (RocContainer.java:225-318)

There is no CPU implementation behind these classes: constructing an Engine without a visible CUDA device
(or without the built csrc/libvce.so) raises.
"""
import ctypes as C
import os
from typing import List, Optional

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VCE_LIB") or os.path.join(_HERE, "csrc", "libvce.so")   # VCE_LIB: another build of the same sources (experiments)


class EngineUnavailable(RuntimeError):
    pass


class Job:
    """Staged evaluation (vce_job_*): create = pack + H2D, run = device pipeline, fetch = D2H + assembly."""

    def __init__(self, engine, cfg, seqs):
        self.engine = engine
        self.ccfg, self.cseqs, self.cres, self.results, self.keep = capi.Library.pack(cfg, seqs)
        self.seqs = seqs
        self.handle = C.c_void_p()
        lib = engine.lib.lib
        rc = lib.vce_job_create(C.byref(self.ccfg), len(seqs), self.cseqs, C.byref(self.handle))
        if rc != 0:
            raise RuntimeError("vce_job_create: %s %s" % (capi.ERR_NAMES.get(rc, rc), engine.lib.last_error()))

    def run(self):
        rc = self.engine.lib.lib.vce_job_run(self.handle)
        if rc != 0:
            raise RuntimeError("vce_job_run: %s %s" % (capi.ERR_NAMES.get(rc, rc), self.engine.lib.last_error()))

    def fetch(self) -> List[capi.SequenceResult]:
        rc = self.engine.lib.lib.vce_job_fetch(self.handle, self.cres)
        capi.Library.unpack(self.cres, self.results, self.keep)
        if rc != 0 and all(r.error == 0 for r in self.results):
            raise RuntimeError("vce_job_fetch: %s %s" % (capi.ERR_NAMES.get(rc, rc), self.engine.lib.last_error()))
        return self.results

    def stats_ex(self):
        out = (C.c_double * 13)()
        self.engine.lib.lib.vce_job_stats_ex(self.handle, out, 13)
        keys = ["launches", "search_ms", "micro_kernel_ms", "micro_regions", "regions", "escalated", "spilled", "prefix_reruns",
                "h2d_bytes", "d2h_bytes", "iterations", "execute_ms", "micro_variants"]
        return {k: out[i] for i, k in enumerate(keys)}

    def stats(self):
        n_launch = C.c_int64()
        dev_ms = C.c_double()
        fast_ms = C.c_double()
        n_reg = C.c_int64()
        n_esc = C.c_int64()
        self.engine.lib.lib.vce_job_stats(self.handle, C.byref(n_launch), C.byref(dev_ms), C.byref(fast_ms), C.byref(n_reg),
                                          C.byref(n_esc))
        return dict(launches=n_launch.value, search_ms=dev_ms.value, micro_kernel_ms=fast_ms.value, regions=n_reg.value,
                    escalated=n_esc.value)

    def close(self):
        if self.handle:
            self.engine.lib.lib.vce_job_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """The CUDA vcfeval engine on one device, or -- `devices=[...]` -- on several devices driven from this one process
    (vce_init_devices: the sequences of every submit are spread over the devices by variant count)."""

    def __init__(self, device: int = 0, devices: Optional[List[int]] = None):
        if not os.path.exists(LIB_PATH):
            raise EngineUnavailable("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        self.lib = capi.Library(LIB_PATH, prefix="vce_")
        lib = self.lib.lib
        lib.vce_init.restype = C.c_int
        lib.vce_init.argtypes = [C.c_int]
        lib.vce_job_create.restype = C.c_int
        lib.vce_job_create.argtypes = [C.POINTER(capi.CConfig), C.c_int32, C.POINTER(capi.CSequence), C.POINTER(C.c_void_p)]
        lib.vce_job_run.restype = C.c_int
        lib.vce_job_run.argtypes = [C.c_void_p]
        lib.vce_job_fetch.restype = C.c_int
        lib.vce_job_fetch.argtypes = [C.c_void_p, C.POINTER(capi.CSequenceResult)]
        lib.vce_job_destroy.restype = None
        lib.vce_job_destroy.argtypes = [C.c_void_p]
        lib.vce_job_stats.restype = C.c_int
        lib.vce_job_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.vce_job_stats_ex.restype = C.c_int
        lib.vce_job_stats_ex.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int32]
        lib.vce_submit_stats.restype = C.c_int
        lib.vce_submit_stats.argtypes = [C.POINTER(C.c_double), C.c_int32]
        if devices is not None:
            lib.vce_init_devices.restype = C.c_int
            lib.vce_init_devices.argtypes = [C.c_int, C.POINTER(C.c_int)]
            arr = (C.c_int * len(devices))(*devices)
            rc = lib.vce_init_devices(len(devices), arr)
            device = devices[0] if devices else 0
        else:
            rc = lib.vce_init(device)
        if rc != 0:
            raise EngineUnavailable("vce_init(%s) failed: %s %s" % (devices if devices is not None else device,
                                                                    capi.ERR_NAMES.get(rc, rc), self.lib.last_error()))
        self.device = device
        self.devices = list(devices) if devices is not None else [device]

    def submit(self, cfg: capi.Config, seqs: List[capi.Sequence]) -> List[capi.SequenceResult]:
        return self.lib.submit(cfg, seqs)

    def register_templates(self, seqs: List[capi.Sequence]) -> int:
        """vce_template_register for every sequence: the reference sequences become resident on the device (load phase,
        as the reference reads them from the SDF before matching, SequenceEvaluator.java:76-86).  Returns the bytes registered."""
        lib = self.lib.lib
        lib.vce_template_register.restype = C.c_int
        lib.vce_template_register.argtypes = [C.c_void_p, C.c_int32]
        total = 0
        for s in seqs:
            s.template = np.ascontiguousarray(s.template, np.uint8)
            rc = lib.vce_template_register(s.template.ctypes.data, int(s.template.shape[0]))
            if rc != 0:
                raise RuntimeError("vce_template_register: %s %s" % (capi.ERR_NAMES.get(rc, rc), self.lib.last_error()))
            total += int(s.template.shape[0])
        return total

    def host_register(self, seqs: List[capi.Sequence], min_bytes: int = 1 << 16) -> int:
        """vce_host_register for the input arrays of `seqs` (arrays of at least min_bytes: small heap arrays share pages and
        cannot be page-locked one by one).  The engine then copies them to the device without a staging copy.  Returns the
        bytes page-locked; call host_unregister before the arrays are freed."""
        lib = self.lib.lib
        lib.vce_host_register.restype = C.c_int
        lib.vce_host_register.argtypes = [C.c_void_p, C.c_uint64]
        total = 0
        self._pinned = getattr(self, "_pinned", [])
        for s in seqs:
            for side in (s.baseline, s.calls):
                side.gt = np.ascontiguousarray(side.gt, np.int8)
                for a in (side.phased, side.status_in, side.gt, side.allele_off, side.allele_start, side.allele_end, side.allele_null,
                          side.allele_nt_off, side.nt):
                    if a is not None and a.nbytes >= min_bytes and lib.vce_host_register(a.ctypes.data, a.nbytes) == 0:
                        total += a.nbytes
                        self._pinned.append(a)
        return total

    def host_unregister(self) -> None:
        lib = self.lib.lib
        lib.vce_host_unregister.restype = C.c_int
        lib.vce_host_unregister.argtypes = [C.c_void_p]
        for a in getattr(self, "_pinned", []):
            lib.vce_host_unregister(a.ctypes.data)
        self._pinned = []

    def unregister_templates(self, seqs: List[capi.Sequence]) -> None:
        lib = self.lib.lib
        lib.vce_template_unregister.restype = C.c_int
        lib.vce_template_unregister.argtypes = [C.c_void_p]
        for s in seqs:
            lib.vce_template_unregister(s.template.ctypes.data)

    def submit_stats(self) -> dict:
        """Diagnostics of the calling thread's last submit(): launches, H2D / D2H bytes, regions, iterations."""
        out = (C.c_double * 6)()
        self.lib.lib.vce_submit_stats(out, 6)
        return dict(launches=int(out[0]), h2d_bytes=int(out[1]), d2h_bytes=int(out[2]), regions=int(out[3]),
                    escalated=int(out[4]), iterations=int(out[5]))

    def submit_summary(self, n_seq: int) -> np.ndarray:
        """vce_submit_summary: (n_seq, 8) int64 -- baseline TP, FN, call TP, FP, skipped, mis-, correct, unphaseable phasings
        of the calling thread's last submit()."""
        out = np.zeros(n_seq * 8, np.int64)
        self.lib.lib.vce_submit_summary.restype = C.c_int
        self.lib.lib.vce_submit_summary.argtypes = [C.c_void_p, C.c_int32]
        rc = self.lib.lib.vce_submit_summary(out.ctypes.data, n_seq)
        if rc != 0:
            raise RuntimeError("vce_submit_summary: %s %s" % (capi.ERR_NAMES.get(rc, rc), self.lib.last_error()))
        return out.reshape(n_seq, 8)

    def job(self, cfg, seqs) -> Job:
        return Job(self, cfg, seqs)

    def roc(self, *a, **k):
        return self.lib.roc(*a, **k)


# ----------------------------------------------------------------------------------------------- reference-style API
class Path:
    """What PathFinder.bestPath() returns (Path.java): included / excluded lists per side + sync points."""

    def __init__(self, seq: capi.Sequence, res: capi.SequenceResult):
        S = capi
        self.sync_points = [int(x) for x in res.sync_points[0]]

        def split(side_res, side):
            inc, exc, skipped = [], [], []
            for i in range(side.n):
                st = int(side_res.status[i])
                if st & S.STATUS_GT_MATCH:
                    inc.append((int(side.id[i]), tuple(int(a) for a in side_res.allele_ids[i] if a != -2), bool(side_res.original[i]),
                                float(side_res.weight[i])))
                elif st & S.STATUS_NO_MATCH:
                    exc.append(int(side.id[i]))
                else:
                    skipped.append(int(side.id[i]))
            return inc, exc, skipped
        self.called_included, self.called_excluded, self.called_skipped = split(res.calls, seq.calls)
        self.baseline_included, self.baseline_excluded, self.baseline_skipped = split(res.baseline, seq.baseline)


class PathFinder:
    """PathFinder(template, templateName, baseLineVariants, calledVariants, baselineOrientor, callOrientor, config)
    (PathFinder.java:106-127).  Variants are capi.VariantSide images; orientors are capi.ORIENT_* codes."""

    def __init__(self, engine: Engine, template, template_name, baseline: capi.VariantSide, calls: capi.VariantSide,
                 baseline_orientor=capi.ORIENT_UNPHASED, call_orientor=capi.ORIENT_UNPHASED, haplotypes=2,
                 config: Optional[capi.Config] = None):
        self.engine = engine
        self.seq = capi.Sequence(template_name, np.asarray(template, np.uint8), baseline, calls)
        cfg = config or capi.Config()
        self.cfg = capi.Config(n_passes=1, baseline_orientor=(baseline_orientor, baseline_orientor),
                               calls_orientor=(call_orientor, call_orientor), ploidy=(haplotypes, haplotypes),
                               max_paths=cfg.max_paths, max_iterations=cfg.max_iterations, prune_no_ops=cfg.prune_no_ops,
                               flag_alternates=cfg.flag_alternates, preference=cfg.preference)

    def best_path(self) -> Path:
        # note: SequenceEvaluator short-circuits when a side is empty (SequenceEvaluator.java:94-97); so does this
        res = self.engine.submit(self.cfg, [self.seq])[0]
        if res.error:
            raise RuntimeError(capi.ERR_NAMES.get(res.error, str(res.error)))
        self.result = res
        return Path(self.seq, res)


class SequenceEvaluator:
    """SequenceEvaluator (SequenceEvaluator.java:56-158) for a batch of sequences."""

    def __init__(self, engine: Engine, config: Optional[capi.Config] = None):
        self.engine = engine
        self.config = config or capi.Config()

    def evaluate(self, seqs: List[capi.Sequence]) -> List[capi.SequenceResult]:
        return self.engine.submit(self.config, seqs)


class RocContainer:
    """RocContainer accumulate / sorted cumulative scan (RocContainer.java:204-318); formatting stays with the caller."""

    def __init__(self, engine: Engine, n_filters: int = 1, ascending: bool = False):
        self.engine = engine
        self.n_filters = n_filters
        self.ascending = ascending
        self._score, self._tp, self._fp, self._raw, self._mask = [], [], [], [], []

    def add_roc_line(self, score, tp_weight, fp_weight, tp_raw, filter_mask=0xffffffff):
        self._score.append(score)
        self._tp.append(tp_weight)
        self._fp.append(fp_weight)
        self._raw.append(tp_raw)
        self._mask.append(filter_mask)

    def add_roc_lines(self, score, tp, fp, raw, mask=None):
        self._score.extend(np.asarray(score, np.float64).tolist())
        self._tp.extend(np.asarray(tp, np.float64).tolist())
        self._fp.extend(np.asarray(fp, np.float64).tolist())
        self._raw.extend(np.asarray(raw, np.float64).tolist())
        n = len(np.asarray(score))
        self._mask.extend([0xffffffff] * n if mask is None else np.asarray(mask, np.uint32).tolist())

    def rows(self):
        return self.engine.roc(np.array(self._score, np.float64), np.array(self._tp, np.float64), np.array(self._fp, np.float64),
                               np.array(self._raw, np.float64), np.array(self._mask, np.uint32), self.n_filters, self.ascending)

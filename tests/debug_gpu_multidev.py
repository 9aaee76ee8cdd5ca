"""Diagnostics (not a test): one process, every visible device, the C2 workload through ONE vce_submit per step -- the JNI host's
form.  Prints milliseconds per call with whole sequences only (VCE_SPLIT=0), with the default rule of
SURVEY 8e (cut when max / mean > 1.1) and with forced cuts, plus the split counters; results are compared with each other.
Usage: python tests/debug_gpu_multidev.py [total_bp]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    from rtg_tools_b200 import capi, synth, vcfeval
    from compare import compare_results
    n = torch.cuda.device_count()
    bp = float(sys.argv[1]) if len(sys.argv) > 1 else 3e9
    seqs = synth.make_pair(int(bp), n_chrom=24)
    nv = sum(s.baseline.n + s.calls.n for s in seqs)
    eng = vcfeval.Engine(devices=list(range(n)))
    eng.register_templates(seqs)
    lib = eng.lib.lib
    lib.vce_submit_stats.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int32]
    packed = eng.lib.pack(capi.Config(), seqs)
    ref = None
    for label, env in (("whole sequences (VCE_SPLIT=0)", {"VCE_SPLIT": "0"}), ("default rule (max/mean > 1.1 and the saving estimate)", {}),
                       ("forced cuts (VCE_SPLIT=1, VCE_SPLIT_IMBALANCE=1.0, VCE_SPLIT_PIECES=8 x devices)", {"VCE_SPLIT": "1", "VCE_SPLIT_IMBALANCE": "1.0", "VCE_SPLIT_PIECES": str(8 * n)})):
        for k in ("VCE_SPLIT", "VCE_SPLIT_IMBALANCE", "VCE_SPLIT_PIECES"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ts = []
        for it in range(6):
            t0 = time.perf_counter()
            rc = eng.lib.submit_packed(packed)
            ts.append((time.perf_counter() - t0) * 1e3)
            assert rc == 0
        if label.startswith("whole") or env.get("VCE_SPLIT_IMBALANCE"):      # one more call with the stage laps on stderr
            print("==== laps: %s" % label, file=sys.stderr, flush=True)
            os.environ["VCE_TIMING"] = "1"
            eng.lib.submit_packed(packed)
            os.environ.pop("VCE_TIMING")
        st = (ctypes.c_double * 9)()
        lib.vce_submit_stats(st, 9)
        import copy
        res = capi.Library.unpack(packed[2], copy.deepcopy(packed[3]), packed[4])   # a snapshot: the packed buffers are reused
        print("%d devices, %d variants, %-42s: %.1f ms per call (best of 5: %.1f), %d sequences cut into %d pieces, %d redone" % (
            n, nv, label, float(np.median(ts[1:])), min(ts[1:]), st[6], st[7], st[8]), flush=True)
        if ref is None:
            ref = res
        else:
            compare_results(ref, res, seqs, what=label)
    print("OK")


if __name__ == "__main__":
    main()

"""Helper of test_gpu_parity.py::test_cuda_one_process_several_devices (runs in its own process: the library's device pool
is process-wide).  vce_init_devices over every visible device: one vce_submit spreads the batch's sequences over the devices
(longest-processing-time greedy by variant count); results, summary counters and a second call on one device must agree
with the oracle bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        print("SKIP: %d CUDA device(s) visible" % n)
        return 0
    from rtg_tools_b200 import capi, multigpu, synth, vcfeval
    from compare import compare_results
    from helpers import CONFIGS
    oracle = capi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), prefix="oracle_")
    seqs = synth.make_pair(300_000_000, n_chrom=24)
    seqs += [synth.dense_cluster_sequence(np.random.default_rng(4), n_clusters=3, per_side=(4, 8), name="dense")]
    eng = vcfeval.Engine(devices=list(range(n)))
    eng.register_templates(seqs)
    for name in ("default", "twopass", "squash"):
        want = oracle.submit(CONFIGS[name], seqs, threads=8)
        got = eng.submit(CONFIGS[name], seqs)
        compare_results(want, got, seqs, what="%d devices, %s" % (n, name))
        if name == "default":
            assert np.array_equal(eng.submit_summary(len(seqs)).sum(axis=0), multigpu.summary_counts(want))
    # long sequences cut into pieces so that the devices balance (vce_split.h).  24 chromosomes balance on a few devices
    # by themselves, so the thresholds are lowered to force the cuts; results must not change, and the counters must say that
    # sequences were cut (and, with a one-variant overlap, that some were evaluated again in one piece)
    import ctypes
    lib = eng.lib.lib
    lib.vce_submit_stats.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int32]
    want = oracle.submit(CONFIGS["default"], seqs, threads=8)
    for overlap, imb in ((512, "1.0"), (1, "1.0")):
        os.environ.update(VCE_SPLIT="1", VCE_SPLIT_IMBALANCE=imb, VCE_SPLIT_MIN="4000", VCE_SPLIT_OVERLAP=str(overlap), VCE_SPLIT_PIECES="64")
        got = eng.submit(CONFIGS["default"], seqs)
        st = (ctypes.c_double * 9)()
        lib.vce_submit_stats(st, 9)
        compare_results(want, got, seqs, what="%d devices, sequences cut (overlap %d)" % (n, overlap))
        assert np.array_equal(eng.submit_summary(len(seqs)).sum(axis=0), multigpu.summary_counts(want))
        assert st[6] >= 10 and st[7] >= 2 * st[6], list(st)
        print("split: %d sequences cut into %d pieces, %d evaluated again in one piece (overlap %d)" % (st[6], st[7], st[8], overlap))
    for k in ("VCE_SPLIT", "VCE_SPLIT_IMBALANCE", "VCE_SPLIT_MIN", "VCE_SPLIT_OVERLAP", "VCE_SPLIT_PIECES"):
        os.environ.pop(k, None)
    # a single sequence takes the one-device path of the same pool
    one = [seqs[0]]
    compare_results(oracle.submit(CONFIGS["default"], one), eng.submit(CONFIGS["default"], one), one, what="one sequence")
    eng.unregister_templates(seqs)
    print("OK: %d devices, %d sequences, %d variants" % (n, len(seqs), sum(s.baseline.n + s.calls.n for s in seqs)))
    return 0


if __name__ == "__main__":
    sys.exit(main())

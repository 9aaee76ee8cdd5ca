"""Diagnostics (not a test): stage timings of the dense-cluster stress (BASELINE.json configs[2]) at the default budget."""
import os
import sys
import time

os.environ["VCE_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from rtg_tools_b200 import capi, synth, vcfeval  # noqa: E402

rng = np.random.default_rng(11)
seqs = [synth.dense_cluster_sequence(rng, n_clusters=3, per_side=(12, 16), name="d%d" % i) for i in range(2)]
eng = vcfeval.Engine(0)
for i in range(2):
    t = time.perf_counter()
    res = eng.submit(capi.Config(), seqs)
    print("submit %.1f ms, warnings %s" % ((time.perf_counter() - t) * 1e3, [r.warnings for r in res]), file=sys.stderr, flush=True)

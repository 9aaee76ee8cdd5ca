"""Diagnostics (not a test): a small mixed batch (genome slice, fuzz, dense clusters with a reduced budget, two-pass) for
compute-sanitizer runs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from helpers import fuzz_batch  # noqa: E402
from rtg_tools_b200 import capi, synth, vcfeval  # noqa: E402

eng = vcfeval.Engine(0)
seqs = synth.make_pair(int(os.environ.get("SAN_BP", "6000000")), n_chrom=3)
for seed in range(6):
    seqs += fuzz_batch(seed, n_seq=3)
rng = np.random.default_rng(3)
seqs += [synth.dense_cluster_sequence(rng, n_clusters=2, per_side=(6, 9), name="dense")]
cfg = capi.Config(n_passes=2, max_paths=2000, max_iterations=20000)
res = eng.submit(cfg, seqs)
print("ok", sum(s.baseline.n + s.calls.n for s in seqs), "variants", sum(len(r.warnings) for r in res), "warnings")

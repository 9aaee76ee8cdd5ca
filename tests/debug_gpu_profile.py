"""Diagnostics (not a test): one vce_submit on a 10 % slice of the synthetic genome -- the command profiled with ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtg_tools_b200 import capi, synth, vcfeval  # noqa: E402

bp = float(sys.argv[1]) if len(sys.argv) > 1 else 3e8
seqs = synth.make_pair(bp, n_chrom=24)
eng = vcfeval.Engine(0)
job = eng.job(capi.Config(), seqs)
job.run()
job.run()
print(job.stats_ex())
job.close()

"""Diagnostics (not a test): stage timings (VCE_TIMING=1) of vce_submit and of a staged job on the full synthetic genome."""
import os
import sys
import time

os.environ["VCE_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtg_tools_b200 import capi, synth, vcfeval  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "default"
seqs = synth.make_pair(3_000_000_000, n_chrom=24)
nv = sum(s.baseline.n + s.calls.n for s in seqs)
eng = vcfeval.Engine(0)
cfg = capi.Config() if cfgname == "default" else capi.Config(n_passes=2)
for i in range(3):
    print("---- submit", i, file=sys.stderr, flush=True)
    t = time.perf_counter()
    eng.submit(cfg, seqs)
    dt = time.perf_counter() - t
    print("submit %.1f ms  %.1f M variants/s" % (dt * 1e3, nv / dt / 1e6), file=sys.stderr, flush=True)
print("---- job", file=sys.stderr, flush=True)
job = eng.job(cfg, seqs)
for i in range(3):
    print("---- job run", i, file=sys.stderr, flush=True)
    job.run()
    print(job.stats_ex(), file=sys.stderr, flush=True)
job.fetch()
job.close()

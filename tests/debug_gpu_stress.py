"""Diagnostics (not a test): repeat the CUDA evaluation of the full synthetic genome and report any run that differs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from rtg_tools_b200 import capi, synth, vcfeval  # noqa: E402


def diff(want, got, seqs, tag):
    bad = 0
    for k, (w, g) in enumerate(zip(want, got)):
        for p in range(2):
            a, b = w.sync_points[p], g.sync_points[p]
            if len(a) != len(b) or not np.array_equal(a, b):
                n = min(len(a), len(b))
                d = np.nonzero(a[:n] != b[:n])[0]
                i = int(d[0]) if len(d) else n
                print(tag, seqs[k].name, "pass", p, "sync differs at", i, "of", len(a), len(b), "want", a[max(0, i - 2):i + 3], "got",
                      b[max(0, i - 2):i + 3], "tmplLen", seqs[k].template.shape[0], flush=True)
                bad += 1
        for side in ("calls", "baseline"):
            ws, gs = getattr(w, side), getattr(g, side)
            if not np.array_equal(ws.status, gs.status):
                i = int(np.nonzero(ws.status != gs.status)[0][0])
                print(tag, seqs[k].name, side, "status differs at", i, ws.status[i], gs.status[i], "ndiff", int((ws.status != gs.status).sum()), flush=True)
                bad += 1
            if not np.array_equal(ws.sync_id, gs.sync_id):
                print(tag, seqs[k].name, side, "sync_id differs", int((ws.sync_id != gs.sync_id).sum()), flush=True)
                bad += 1
    return bad


def main():
    n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    seqs = synth.make_pair(3_000_000_000, n_chrom=24)
    eng = vcfeval.Engine(0)
    oracle = capi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), prefix="oracle_")
    cfg = capi.Config()
    want = oracle.submit(cfg, seqs, threads=16)
    total = 0
    job = eng.job(cfg, seqs)
    for r in range(n_rep):
        job.run()
        ccfg, cseqs, cres, results, keep = capi.Library.pack(cfg, seqs)
        rc = eng.lib.lib.vce_job_fetch(job.handle, cres)
        capi.Library.unpack(cres, results, keep)
        total += diff(want, results, seqs, "job run %d" % r)
    job.close()
    for r in range(n_rep):
        total += diff(want, eng.submit(cfg, seqs), seqs, "submit %d" % r)
    print("total differing items:", total)


if __name__ == "__main__":
    main()

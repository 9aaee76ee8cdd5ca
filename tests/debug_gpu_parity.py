"""Diagnostics (not a test): locate the first difference between the CUDA engine and the oracle on one chromosome."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from rtg_tools_b200 import capi, synth, vcfeval  # noqa: E402


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.nonzero(a[:n] != b[:n])[0]
    return int(d[0]) if len(d) else (n if len(a) != len(b) else -1)


def main():
    chroms = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20]
    mode = sys.argv[2] if len(sys.argv) > 2 else "submit"
    seqs = synth.make_pair(3_000_000_000, n_chrom=24, only=chroms)
    eng = vcfeval.Engine(0)
    oracle = capi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), prefix="oracle_")
    cfg = capi.Config()
    want = oracle.submit(cfg, seqs, threads=8)
    if mode == "submit":
        got = eng.submit(cfg, seqs)
    else:
        job = eng.job(cfg, seqs)
        for _ in range(int(mode)):
            job.run()
        got = job.fetch()
        job.close()
    bad = False
    for k, (w, g) in enumerate(zip(want, got)):
        for side in ("calls", "baseline"):
            ws, gs = getattr(w, side), getattr(g, side)
            i = first_diff(ws.status, gs.status)
            if i >= 0:
                sv = getattr(seqs[k], side)
                print(seqs[k].name, side, "status differs at", i, "want", ws.status[i], "got", gs.status[i], "n diff",
                      int((ws.status != gs.status).sum()), "start", sv.allele_start[3 * i + 1])
                bad = True
        i = first_diff(w.sync_points[0], g.sync_points[0])
        if i >= 0:
            print(seqs[k].name, "sync differs at", i, "want", w.sync_points[0][max(0, i - 2):i + 3], "got", g.sync_points[0][max(0, i - 2):i + 3],
                  "lens", len(w.sync_points[0]), len(g.sync_points[0]))
            pos = int(w.sync_points[0][i]) if i < len(w.sync_points[0]) else -1
            for side in ("calls", "baseline"):
                sv = getattr(seqs[k], side)
                st = sv.allele_start[1::3]
                near = np.nonzero((st > pos - 60) & (st < pos + 60))[0]
                for v in near:
                    print("   ", side, "var", v, "start", st[v], "end", sv.allele_end[3 * v + 1], "gt", sv.gt[v], "alt",
                          bytes(sv.nt[sv.allele_nt_off[3 * v + 2]:sv.allele_nt_off[3 * v + 3]]).hex(), "ref",
                          bytes(sv.nt[sv.allele_nt_off[3 * v + 1]:sv.allele_nt_off[3 * v + 2]]).hex(),
                          "want", getattr(w, side).status[v], "got", getattr(g, side).status[v])
                print("    template", bytes(seqs[k].template[pos - 10:pos + 40]).hex())
            bad = True
    print("DIFFERENT" if bad else "identical", chroms, mode)


if __name__ == "__main__":
    main()

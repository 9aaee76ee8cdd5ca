import ctypes
import os
import subprocess

import numpy as np

from rtg_tools_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = {
    "default": capi.Config(),
    "twopass": capi.Config(n_passes=2),
    "squash": capi.Config(n_passes=1, baseline_orientor=(3, 3), calls_orientor=(3, 3), ploidy=(1, 1)),
    "allele_gt": capi.Config(n_passes=1, baseline_orientor=(4, 4), calls_orientor=(4, 4), ploidy=(2, 2)),
    "phased": capi.Config(n_passes=1, baseline_orientor=(1, 3), calls_orientor=(2, 3)),
    "calls_min_base": capi.Config(n_passes=1, preference=1),
    "small_budget": capi.Config(n_passes=2, max_paths=20, max_iterations=300),
    "no_prune": capi.Config(n_passes=1, prune_no_ops=False),
}


def fuzz_batch(seed, n_seq=3):
    rng = np.random.default_rng(seed)
    return [synth.fuzz_sequence(rng, length=int(rng.integers(60, 500)), n_base=int(rng.integers(0, 16)),
                                n_calls=int(rng.integers(0, 16)), repeat=bool(rng.random() < 0.7), name="s%d_%d" % (seed, i))
            for i in range(n_seq)]


def emul_library():
    so = os.path.join(ROOT, "tests", "emul", "libvce_emul.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul")])
    return capi.Library(so, prefix="emul_")


class Hybrid:
    """submit() through one engine, ROC through another (the emulation has no ROC)."""

    def __init__(self, submit_engine, roc_engine):
        self.s = submit_engine
        self.r = roc_engine

    def submit(self, cfg, seqs):
        return self.s.submit(cfg, seqs)

    def roc(self, *a, **k):
        return self.r.roc(*a, **k)


def emul_stats(emul):
    st = (ctypes.c_int64 * 8)()
    emul.lib.emul_get_stats(st)
    return dict(regions=st[0], escalated=st[1], spilled=st[2], prefix_reruns=st[3], launches=st[4], micro_regions=st[5],
                micro_clean=st[6])

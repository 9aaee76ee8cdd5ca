#!/usr/bin/env python
"""bench.py -- vcfeval matching throughput on B200 (BASELINE.json metric: variants evaluated/s).

A "step" is one pass of the hot path (SequenceEvaluator.run() for every sequence of the workload:
orientation expansion, region search, tie-break/weights, status merge, sync ids, phasing counts) over one
batch of synthetic input.

  value : whole-job variants/s with the batch already resident in HBM (staged job: vce_job_run), timed with CUDA
          events on the engine's stream, max over ranks
  e2e   : the same metric through the reference-facing C-ABI call vce_submit() with HOST buffers: uploads, region
          split, kernels, final passes, D2H and the expansion into the caller's arrays all inside the timed region.
          Before the timed region the reference sequences are registered (vce_template_register: the load phase, which
          BASELINE.md excludes from both arms) and the input arrays page-locked (vce_host_register: "pinned host memory",
          as the contract asks); --no-register / --no-pin measure without either
  roofline : the dominant kernel (thread-per-region search), 48 algorithmic bytes per variant
  cpu_baseline : the CPU oracle (C++ restatement of the reference, "port": the reference is Java and the image
          has no JDK) on the host cores, one worker per sequence like the reference's SimpleThreadPool
  roc_heavy : BASELINE configs[4], K4 alone on --roc-calls (default 10 M) synthetic scored calls x 5 filters: end to end,
          kernel time (vce_roc_timed) with its own roofline at 72 algorithmic bytes per call, CPU oracle beside it

`--impl reference` times that CPU path alone, on the same workload, as the reference arm.
`--workload c3 --c3-clusters N` is BASELINE configs[2] (dense clusters at the default 50 000-path budget).

Multi-GPU (torchrun, one rank per GPU): sequences are independent units (VcfEvalTask.java:131-135), so every rank
evaluates its own shard with no data-path collective; the only exchange is the summary-count merge (NCCL
all-reduce) and the ROC tuple gather.  Default scaling is STRONG: ONE sample pair, its 24 chromosomes LPT-sharded over
the ranks by variant count (the reference arm evaluates the same pair); `--scaling weak` gives rank r its own sample
pair r of a cohort (same genome layout, different sample seeds).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES_PER_VARIANT = 48  # 32 B in + 16 B out per variant
ALG_BYTES_PER_ROC_CALL = 72

WORKLOADS = {
    # name: (total_bp, n_chrom, description)
    "c2": (3_000_000_000, 24, "synthetic 3 Gbp 24-chromosome genome, ~4.4M baseline/call variants per side (BASELINE.json configs[1])"),
    "c1": (10_000_000, 1, "synthetic 10 Mbp genome, ~15k variants per side (BASELINE.json configs[0])"),
    "tiny": (2_000_000, 4, "2 Mbp smoke workload"),
    # BASELINE.json configs[2]: --c3-clusters 500 bp tandem-repeat clusters, 12-40 overlapping heterozygous indels per side,
    # alleles up to 1000 bases, 4 clusters per sequence; every cluster exceeds the 50 000-path budget at least once
    "c3": (0, 0, "dense complex-variant stress: 500 bp tandem-repeat clusters, 12-40 overlapping het indels per side, alleles up to "
                 "1000 bases, default budget 50 000 paths / 10^7 iterations (BASELINE.json configs[2])"),
}
C3_PER_SEQ = 4
C3_CLUSTERS = [2000]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line, the result JSON: libraries that print to file descriptor 1 (NCCL prints its version
# there) are diverted to stderr for the whole run, the JSON line goes to the saved descriptor.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=250):
        self.index = index
        self.interval_ms = int(interval_ms)
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", str(self.interval_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            try:  # the loop produced nothing in time: one direct query right after the timed region
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
                f = [x.strip() for x in out[0].split(",")]
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name, rank, world, scaling, total_bp=None):
    """Returns (sequences, call scores, description, global sequence ids, total sequences of the job)."""
    from rtg_tools_b200 import multigpu, synth
    bp, n_chrom, desc = WORKLOADS[name]
    if total_bp:
        bp = int(total_bp)
    if name == "c3":
        # sequences (4 clusters each) are the units; strong scaling shards them by LPT on their variant counts (uniform here)
        n_seq = (C3_CLUSTERS[0] + C3_PER_SEQ - 1) // C3_PER_SEQ
        mine = list(range(n_seq)) if world == 1 else multigpu.shard_sequences([1.0] * n_seq, rank, world)
        seqs, quals = synth.make_c3(n_clusters=C3_CLUSTERS[0], per_seq=C3_PER_SEQ, only=set(mine), return_qual=True)
        return seqs, quals, "%d x %s" % (C3_CLUSTERS[0], desc), mine, n_seq
    if scaling == "strong" and world > 1:
        # ONE sample pair, its chromosomes sharded over the ranks by LPT (weights: chromosome lengths ~ variant counts)
        lens = synth.genome_lengths(bp, n_chrom) if n_chrom > 1 else [bp]
        mine = multigpu.shard_sequences(lens, rank, world)
        seqs, quals = synth.make_pair(bp, n_chrom=n_chrom, only=mine, return_qual=True)
        return seqs, quals, desc, mine, len(lens)
    # weak: rank r evaluates sample pair r of the cohort (pair 0 = the canonical seeds 42 / 572 / 126)
    seqs, quals = synth.make_pair(bp, n_chrom=n_chrom, sample_seed=572 + 1000 * rank, perturb_seed=126 + 1000 * rank,
                                  return_qual=True)
    return seqs, quals, desc, list(range(rank * len(seqs), (rank + 1) * len(seqs))), world * len(seqs)


def n_variants(seqs):
    return int(sum(s.baseline.n + s.calls.n for s in seqs))


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path (C++ restatement -- the Java reference cannot run here)
    on all host cores, one worker per sequence as the reference's SimpleThreadPool does."""
    if rank != 0:
        return
    from rtg_tools_b200 import capi
    seqs, quals, desc, _, _ = build_workload(args.workload, 0, 1, "strong", args.total_bp)
    oracle = capi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), prefix="oracle_")
    cores = os.cpu_count() or 1
    nv_full = n_variants(seqs)
    sample = "the full workload per step"
    if args.workload == "c3":
        # a cluster costs ~2 s of one core: the step is a bounded sample, one sequence (4 clusters) per host core
        seqs = seqs[:max(1, min(cores, len(seqs)))]
        sample = "the first %d of the workload's sequences per step (%d clusters)" % (len(seqs), len(seqs) * C3_PER_SEQ)
    threads = max(1, min(cores, len(seqs)))
    cfg = engine_config(args.config)
    nv = n_variants(seqs)
    # the caller owns and reuses every input / output buffer, exactly as the B200 arm's vce_submit calls do
    packed = capi.Library.pack(cfg, seqs)
    for _ in range(args.warmup):
        oracle.submit_packed(packed, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = oracle.submit_packed(packed, threads=threads)
        if rc != 0:
            raise RuntimeError("oracle failed: %d" % rc)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    res = capi.Library.unpack(packed[2], packed[3], packed[4])
    regions = sum(len(r.sync_points[0]) for r in res)
    v = nv / dt
    out = {
        "impl": "reference", "metric": "vcfeval variants evaluated/sec", "value": v, "unit": "variants/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, desc, nv_full),
        "sync_regions_per_step": int(regions),
        "cpu_baseline": {"value": v, "unit": "variants/s", "cores": threads, "kind": "port",
                         "sample": "%s (%d variants, %d sequences), %d worker threads of %d host cores" % (
                             sample, nv, len(seqs), threads, cores)},
        "e2e": {"value": v, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sync_regions_per_s": regions / dt,
    }
    emit(out)


def engine_config(name):
    from rtg_tools_b200 import capi
    return {"default": capi.Config(),
            "squash": capi.Config(n_passes=1, baseline_orientor=(3, 3), calls_orientor=(3, 3), ploidy=(1, 1)),
            "twopass": capi.Config(n_passes=2)}[name]


def workload_config(args, desc, variants):
    """The `config` object: the same keys and values in both arms (the driver compares them)."""
    return {"workload": "%s: %s" % (args.workload, desc), "engine_config": args.config, "variants_per_step": int(variants),
            # timing rule: either the step's inputs are larger than the 126 MB L2, or L2 is flushed between the timed steps
            "l2": ("inputs larger than L2 (per-variant arrays + reference windows > 126 MB per GPU)" if not l2_flush_needed(args, variants)
                   else "L2 flushed between timed steps (256 MB written on the device; a rank's inputs are smaller than L2)")}


def l2_flush_needed(args, variants):
    per_gpu = variants / max(1, args.gpus) if args.scaling == "strong" else variants
    return per_gpu < 2_000_000


def formats_leg(eng, args, text, vtl, peak, peak_src):
    """SURVEY 8f rank 4, reported beside: the VCF text of the loader leg as a bgzip file through vce_vcf_parse_bgzf (inflate on
    the device, one BGZF member per warp, chained into the parse) and the inflate alone; a chromosome-sized sequence of SDF bit
    planes through vce_template_register_sdf.  CPU beside it: zlib itself (what the JDK's Inflater wraps) on one core."""
    import zlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import formats_oracle as fo
    raw = text.encode()
    data, index = fo.bgzf_write(raw, level=6)
    data_np = np.frombuffer(data, np.uint8).copy()
    if not args.no_pin:
        eng.lib.lib.vce_host_register.restype = ctypes.c_int
        eng.lib.lib.vce_host_register.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        eng.lib.lib.vce_host_register(data_np.ctypes.data, data_np.nbytes)
    for _ in range(3):
        eng.lib.vcf_parse_bgzf(data_np, 0, -1, "chr1", vtl, 0, check_crc=False)
    runs = []
    for _ in range(max(3, args.steps)):
        side, st = eng.lib.vcf_parse_bgzf(data_np, 0, -1, "chr1", vtl, 0, check_crc=False)
        runs.append(st)
    st = sorted(runs, key=lambda r: r["kernel_ms"])[len(runs) // 2]
    # the inflate kernel alone (the same members; the text comes back to the host here, which the chained call avoids)
    inf = []
    for _ in range(4):
        got_text, ist = eng.lib.bgzf_inflate(data_np, check_crc=False)
        inf.append(ist)
    ist = sorted(inf[1:], key=lambda r: r["kernel_ms"])[len(inf[1:]) // 2]
    _, cst = eng.lib.bgzf_inflate(data_np, check_crc=True)
    # throughput regime: the same members eight times over (a valid BGZF file), enough warps to fill the SMs
    body = data[:-len(fo.EOF_BLOCK)]
    big = np.frombuffer(body * 8 + fo.EOF_BLOCK, np.uint8).copy()
    bigs = []
    for _ in range(3):
        big_text, bst = eng.lib.bgzf_inflate(big, check_crc=False)
        bigs.append(bst)
    assert len(big_text) == 8 * len(raw) and big_text[-len(raw):] == raw and big_text[:len(raw)] == raw
    del big_text
    bst = sorted(bigs[1:], key=lambda r: r["kernel_ms"])[0]
    assert got_text == raw, "inflated text differs from the input text"
    ach = (len(data) + len(raw)) / (ist["kernel_ms"] * 1e-3) / 1e9
    c_abi_ms = st["h2d_ms"] + st["kernel_ms"] + st["d2h_ms"]
    out = {"bgzf": {"members": int(ist["blocks"]), "compressed_bytes": len(data), "text_bytes": len(raw),
                    "inflate_kernel_ms": ist["kernel_ms"], "inflate_plus_crc32_kernel_ms": cst["kernel_ms"],
                    "inflate_text_GBps": len(raw) / (ist["kernel_ms"] * 1e-3) / 1e9,
                    "large_file": {"members": int(bst["blocks"]), "compressed_bytes": int(big.nbytes), "text_bytes": 8 * len(raw),
                                   "inflate_kernel_ms": bst["kernel_ms"], "inflate_text_GBps": 8 * len(raw) / (bst["kernel_ms"] * 1e-3) / 1e9,
                                   "roofline_frac": (int(big.nbytes) + 8 * len(raw)) / (bst["kernel_ms"] * 1e-3) / 1e9 / peak},
                    "chain": {"what": "vce_vcf_parse_bgzf: compressed host bytes in, vce_variants arrays out (CUDA events on the loader's stream)",
                              "records": int(st["records"]), "variants": int(st["kept"]), "h2d_ms": st["h2d_ms"], "kernel_ms": st["kernel_ms"],
                              "d2h_ms": st["d2h_ms"], "e2e_ms": c_abi_ms, "e2e_records_per_s": st["records"] / (c_abi_ms * 1e-3),
                              "gpu_launches": int(st["launches"])},
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                 "algorithmic_bytes": "compressed bytes in + text bytes out", "peak_source": peak_src}}}
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        want = fo.bgzf_read(data, check_crc=False)
        dt = time.perf_counter() - t0
        assert want == raw
        plain, _ = eng.lib.vcf_parse(raw, "chr1", vtl, 0)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import vcf_check
        vcf_check.assert_same_side(side, plain, "bench: BGZF chain vs plain-text loader")
        out["bgzf"]["cpu_baseline"] = {"value": len(raw) / dt / 1e9, "unit": "GB/s of text", "cores": 1, "kind": "reference",
                                       "sample": "all %d members through zlib (oracle/formats_oracle.bgzf_read: zlib.decompressobj(-15) per member, "
                                                 "the library the JDK's Inflater wraps), %.2f s" % (ist["blocks"], dt),
                                       "parity": "text byte-identical; parsed arrays identical to the plain-text loader's"}
    if not args.no_pin:
        eng.lib.lib.vce_host_unregister.restype = ctypes.c_int
        eng.lib.lib.vce_host_unregister.argtypes = [ctypes.c_void_p]
        eng.lib.lib.vce_host_unregister(data_np.ctypes.data)
    # SDF: one chromosome-sized sequence (chr1 of the C2 genome: ~250 Mbp) of bit planes -> resident template
    n_bases = 248_000_000
    rng = np.random.default_rng(3)
    codes = rng.integers(1, 5, n_bases, dtype=np.uint8)
    codes[10_000_000:10_050_000] = 0
    file = np.frombuffer(fo.sdf_planes_write(codes), np.uint8).copy()
    ms = []
    for _ in range(3):
        bases = eng.lib.template_register_sdf(file, 0, n_bases)
        ms.append(eng.lib.last_sdf_ms)
    assert np.array_equal(bases, codes), "SDF decode differs from the codes written"
    h2d, ker, d2h = sorted(ms, key=lambda r: r[1])[1]
    lib = eng.lib.lib
    lib.vce_template_unregister.restype = ctypes.c_int
    lib.vce_template_unregister.argtypes = [ctypes.c_void_p]
    lib.vce_template_unregister(bases.ctypes.data)
    alg = file.nbytes + n_bases + n_bases // 4 + n_bases // 128
    out["sdf"] = {"bases": n_bases, "seqdata_bytes": int(file.nbytes), "h2d_ms": h2d, "kernel_ms": ker, "d2h_ms": d2h,
                  "roofline": {"bound": "hbm", "achieved": alg / (ker * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": alg / (ker * 1e-3) / 1e9 / peak, "traffic": None,
                               "algorithmic_bytes": "bit planes in + byte-per-base + 2-bit words + N mask out", "peak_source": peak_src},
                  "parity": "byte-per-base array identical to the codes written (FileBitwiseInputStream restated)"}
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        want = fo.sdf_planes_read(file.tobytes(), 0, 20_000_000)
        dt = time.perf_counter() - t0
        assert np.array_equal(want, codes[:20_000_000])
        out["sdf"]["cpu_baseline"] = {"value": 20_000_000 / dt, "unit": "bases/s", "cores": 1, "kind": "port",
                                      "sample": "the first 20 M bases through oracle/formats_oracle.sdf_planes_read (numpy), %.2f s" % dt}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = ONE sample pair, its sequences sharded over the ranks by LPT; weak = one sample pair per rank")
    ap.add_argument("--c3-clusters", type=int, default=2000, help="--workload c3: number of 500 bp clusters (BASELINE.md: 2000)")
    ap.add_argument("--total-bp", type=float, default=None, help="override the genome size of the workload (testing)")
    ap.add_argument("--config", default="default", choices=["default", "squash", "twopass"],
                    help="engine configuration: default (diploid, single pass), squash (--squash-ploidy), twopass (ga4gh / annotate)")
    ap.add_argument("--roc-calls", type=int, default=10_000_000,
                    help="also time K4 alone on this many synthetic scored calls x 5 filters (BASELINE configs[4]: 10 M); 0 = skip")
    ap.add_argument("--pair", type=int, default=None, help="N=1 only: evaluate the sample pair that rank PAIR of a weak-scaling run would get (diagnostics)")
    ap.add_argument("--no-register", action="store_true",
                    help="do not make the reference sequences resident on the device (vce_template_register): windows gathered on the host per call")
    ap.add_argument("--no-pin", action="store_true",
                    help="do not page-lock the caller's input arrays (vce_host_register): the engine stages them through its own pinned slabs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roc", action="store_true")
    ap.add_argument("--vcf-records", type=int, default=1_000_000,
                    help="also time the VCF loader (SURVEY 8f rank 3, vce_vcf_parse) on a synthetic text of this many records; 0 = skip")
    ap.add_argument("--no-formats", action="store_true", help="skip the BGZF / SDF decoders (SURVEY 8f rank 4) reported beside the VCF loader")
    ap.add_argument("--formats-only", action="store_true", help="diagnostics: only the BGZF / SDF leg, printed as its own JSON line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    C3_CLUSTERS[0] = args.c3_clusters

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    # the ranks of one box share its host cores: give every rank's digest / assembly pool its share
    cores = os.cpu_count() or 1
    os.environ.setdefault("VCE_THREADS", str(max(1, cores // max(1, world))))
    from rtg_tools_b200 import capi, multigpu, synth, vcfeval

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_gen = time.perf_counter()
    pair_rank = args.pair if (args.pair is not None and world == 1) else rank
    seqs, quals, desc, seq_ids, n_seq_total = build_workload(args.workload, pair_rank, world, args.scaling, args.total_bp)
    nv = n_variants(seqs)
    log("[rank %d] workload %s: %d sequences, %d variants (generated in %.1fs)" % (rank, args.workload, len(seqs), nv,
                                                                               time.perf_counter() - t_gen))
    eng = vcfeval.Engine(local_rank)
    cfg = engine_config(args.config)
    if args.formats_only:
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            with open(ppath) as f:
                peaks = json.load(f)
        pk = float(peaks.get("hbm_gbs", 6650.0))
        text, vtl = synth.make_vcf_text(args.vcf_records, seed=11, mean_gap=100)
        emit({"formats_load": formats_leg(eng, args, text, vtl, pk, "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s")})
        return

    # ------------------------------------------------------------------ device-resident throughput (value)
    job = eng.job(cfg, seqs)
    # nvidia-smi needs a few hundred ms before its first line and a timed region is ~150 ms: the sampler runs from the
    # warm-up steps (the same kernels, the same load) through both timed regions
    # (several pollers on one box contend in the driver: the more ranks, the slower each polls)
    sampler = ClockSampler(local_rank, 250 if world == 1 else 125 * world)
    sampler.start()
    # (the dense-cluster workload takes minutes per step: it honours --warmup as given, the others warm up at least 3 times)
    n_warm = args.warmup if args.workload == "c3" else max(args.warmup, 3)
    for _ in range(n_warm):
        job.run()
    barrier()
    dev_ms = 0.0
    launches = 0
    search_ms = 0.0
    t0 = time.perf_counter()
    flush = None
    if l2_flush_needed(args, nv * (world if args.scaling == "strong" else 1)):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % local_rank)
    for _ in range(args.steps):
        if flush is not None:   # outside the step's CUDA-event bracket (execute_ms is per step)
            flush.zero_()
            torch.cuda.synchronize()
        job.run()
        st = job.stats_ex()
        dev_ms += st["execute_ms"]
        launches += int(st["launches"])
        search_ms += st["search_ms"]
    own_wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps   # this rank's own time, before it waits for the others
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    results = job.fetch()
    st = job.stats_ex()
    regions = int(sum(len(r.sync_points[0]) for r in results))
    job.close()

    # ------------------------------------------------------------------ end to end through vce_submit (host buffers)
    # the caller owns every buffer (inputs and result arrays) and reuses them across calls, as the JNI host would
    # the reference sequences become resident in HBM first -- the load phase, where the reference reads them from the SDF
    # before "Starting path-finding" (SequenceEvaluator.java:76-86; BASELINE.md 3.2/3.4 exclude it from both arms)
    t_reg = time.perf_counter()
    registered = 0 if args.no_register else eng.register_templates(seqs)
    t_reg = time.perf_counter() - t_reg
    # the caller's input arrays are page-locked once, as the bench contract asks (inputs copied from pinned host memory)
    pinned_bytes = 0 if args.no_pin else eng.host_register(seqs)
    packed = capi.Library.pack(cfg, seqs)
    res = packed[3]
    for _ in range(min(args.warmup, 3) if args.workload == "c3" else max(1, min(args.warmup, 3))):
        eng.lib.submit_packed(packed)
    barrier()
    t0 = time.perf_counter()
    counts = None
    for _ in range(args.steps):
        rc = eng.lib.submit_packed(packed)
        if rc != 0:
            raise RuntimeError("vce_submit failed: %s" % eng.lib.last_error())
        # the summary merge, the only collective of the path; the counters come from the device with the results
        counts = multigpu.merge_counts(eng.submit_summary(len(seqs)).sum(axis=0))
    own_e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()
    sub_stats = eng.submit_stats()   # launches, H2D bytes, D2H bytes, ... of this thread's last vce_submit
    e2e_results = capi.Library.unpack(packed[2], packed[3], packed[4])

    # max over ranks / totals
    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # every rank's own share and times (the step is as slow as the slowest rank: this names it)
    mine = [float(rank), float(nv), float(regions), float(len(seqs)), dev_ms / args.steps, own_wall_ms, own_e2e_ms,
            float(sub_stats["h2d_bytes"]), search_ms / args.steps]
    if world > 1:
        tm = torch.tensor(mine, dtype=torch.float64, device="cuda")
        allr = torch.empty(world * len(mine), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, tm)
        allr = allr.cpu().reshape(world, len(mine)).tolist()
    else:
        allr = [mine]
    per_rank = [{"rank": int(r[0]), "variants": int(r[1]), "sync_regions": int(r[2]), "sequences": int(r[3]), "staged_device_ms": r[4],
                 "staged_wall_ms": r[5], "e2e_ms_before_the_barrier": r[6], "e2e_h2d_bytes": int(r[7]), "search_kernels_ms": r[8]} for r in allr]
    ms_per_step = reduce_max(dev_ms / args.steps)
    wall_per_step = reduce_max(wall_ms / args.steps)
    e2e_ms = reduce_max(e2e_ms)
    total_variants = reduce_sum(nv)
    total_regions = reduce_sum(regions)
    total_launches = reduce_sum(launches)

    # ROC tuples of every rank travel to rank 0 in (sequence, record) order: K4 runs once on the merged stream
    roc_rows = multigpu.gather_roc([multigpu.roc_tuples([r], [q]) for r, q in zip(results, quals)], seq_ids, n_seq_total) \
        if (world > 1 and not args.no_roc) else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fast_ms = st["micro_kernel_ms"]
    fast_var = st["micro_variants"]
    achieved = (ALG_BYTES_PER_VARIANT * fast_var / (fast_ms * 1e-3) / 1e9) if fast_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    # the ncu capture is of the C2 default-configuration launch; other workloads have no measured traffic figure
    if os.path.exists(tpath) and args.workload == "c2" and args.config == "default" and not args.total_bp:
        try:
            traffic = json.load(open(tpath)).get("k_micro_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "k_micro<MicroA> (thread-per-region search from region packets, K1+K2+K3)", "variants_per_launch": fast_var,
                "launch_ms": fast_ms, "algorithmic_bytes_per_variant": ALG_BYTES_PER_VARIANT, "peak_source": peak_src,
                # share of the staged step (device-timed, includes the settling rounds on the host) spent in this launch;
                # the warp-per-region batches run concurrently on another stream, so shares do not add up to 1
                "share_of_step": fast_ms / (dev_ms / args.steps) if dev_ms > 0 else None,
                "all_search_kernels_ms_per_step": search_ms / args.steps}

    # ------------------------------------------------------------------ CPU baseline (oracle, bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        oracle = capi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), prefix="oracle_")
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, len(seqs)))
        t0 = time.perf_counter()
        want = oracle.submit(cfg, seqs, threads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": nv / dt, "unit": "variants/s", "cores": threads, "kind": "port",
               "sample": "one pass over the full workload (%d variants, %d sequences) on %d worker threads of %d host cores, %.1f s" % (
                   nv, len(seqs), threads, cores, dt)}
        # the bench output is also a parity check at full size
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from compare import compare_results
        compare_results(want, results, seqs, what="bench (staged job)")
        compare_results(want, e2e_results, seqs, what="bench (vce_submit)")
        cpu["parity"] = "bit-exact vs oracle on all %d variants (staged job and vce_submit)" % nv

    # ------------------------------------------------------------------ ROC (K4) on the scored calls of the run, reported beside
    roc = None
    if not args.no_roc:
        if roc_rows is None:
            roc_rows = multigpu.roc_tuples(results, quals)
        score, tp, fp, raw = (np.ascontiguousarray(roc_rows[:, i]) for i in range(4))
        eng.roc(score, tp, fp, raw, None, 1)
        t0 = time.perf_counter()
        rows, _ = eng.roc(score, tp, fp, raw, None, 1)
        dt = time.perf_counter() - t0
        roc = {"scored_calls": int(score.shape[0]), "rows": int(rows[0][0].shape[0]), "e2e_calls_per_s": score.shape[0] / dt,
               "merged_over_ranks": world}
        if cpu is not None:
            want_rows, _ = oracle.roc(score, tp, fp, raw, None, 1)
            for x, y in zip(rows[0], want_rows[0]):
                assert np.array_equal(x.view(np.int64), y.view(np.int64)), "ROC rows differ from the oracle"
            roc["parity"] = "bit-exact vs oracle"

    roc_heavy = None
    if args.roc_calls > 0 and not args.no_roc and rank == 0:  # BASELINE configs[4]: QUAL-like scores (40 % distinct at 3 dp), 5 filters, K4 alone
        rng = np.random.default_rng(5)
        nr = args.roc_calls
        score = np.round(rng.uniform(0, 100, nr), 3)
        dup = rng.random(nr) < 0.6
        score[dup] = np.round(score[dup], 1)
        tpw = rng.choice([1.0, 0.5, 1.0 / 3, 2.0 / 3, 2.0, 0.0], nr)
        fpw = (tpw == 0).astype(np.float64)
        raww = (tpw > 0).astype(np.float64)
        mask = rng.integers(1, 32, nr).astype(np.uint32)
        eng.roc(score, tpw, fpw, raww, mask, 5)
        t0 = time.perf_counter()
        rows5, _ = eng.roc(score, tpw, fpw, raww, mask, 5, timed=True)
        dt = time.perf_counter() - t0
        tmg = eng.lib.last_roc_timing or {}
        roc_heavy = {"scored_calls": nr, "filters": 5, "e2e_calls_per_s": nr / dt, "ms": dt * 1e3,
                     "algorithmic_bytes_per_call": ALG_BYTES_PER_ROC_CALL,
                     "e2e_GBps_algorithmic": ALG_BYTES_PER_ROC_CALL * nr / dt / 1e9}
        if tmg.get("kernel_ms"):
            # K4's own roofline: sort + bucket sums + scans of all five filters, CUDA events on the ROC stream
            ach = ALG_BYTES_PER_ROC_CALL * nr / (tmg["kernel_ms"] * 1e-3) / 1e9
            roc_heavy.update({"kernel_ms": tmg["kernel_ms"], "h2d_ms": tmg["h2d_ms"], "gpu_launches": tmg["launches"],
                              "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                           "traffic": None, "peak_source": peak_src}})
        if not args.no_cpu_baseline and world == 1:
            orc = capi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), prefix="oracle_")
            t0 = time.perf_counter()
            want5, _ = orc.roc(score, tpw, fpw, raww, mask, 5)
            roc_heavy["cpu_calls_per_s"] = nr / (time.perf_counter() - t0)
            for fa, fb in zip(rows5, want5):
                for x, y in zip(fa, fb):
                    assert np.array_equal(x.view(np.int64), y.view(np.int64)), "ROC rows differ from the oracle"
            roc_heavy["parity"] = "bit-exact vs oracle"

    # ------------------------------------------------------------------ VCF loader, reported beside
    vcf_load = None
    if args.vcf_records > 0 and rank == 0:
        text, vtl = synth.make_vcf_text(args.vcf_records, seed=11, mean_gap=100)
        raw = np.frombuffer(text.encode(), np.uint8).copy()
        if not args.no_pin:   # the caller's text buffer page-locked once, like the arrays of the matching path
            eng.lib.lib.vce_host_register.restype = ctypes.c_int
            eng.lib.lib.vce_host_register.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
            eng.lib.lib.vce_host_register(raw.ctypes.data, raw.nbytes)
        for _ in range(3):
            eng.lib.vcf_parse(raw, "chr1", vtl, 0)
        vt = []
        for _ in range(max(3, args.steps)):
            t0 = time.perf_counter()
            side, vst = eng.lib.vcf_parse(raw, "chr1", vtl, 0)
            vt.append(time.perf_counter() - t0)
        fields = ("id", "phased", "status_in", "gt", "allele_off", "allele_start", "allele_end", "allele_null", "allele_nt_off", "nt")
        out_bytes = int(sum(getattr(side, f).nbytes for f in fields))
        c_abi_ms = vst["h2d_ms"] + vst["kernel_ms"] + vst["d2h_ms"]
        ach = (int(raw.nbytes) + out_bytes) / (vst["kernel_ms"] * 1e-3) / 1e9
        vcf_load = {"records": int(vst["records"]), "variants": int(vst["kept"]), "text_bytes": int(raw.nbytes), "output_bytes": out_bytes,
                    # host text in, host arrays out: H2D + kernels + D2H on the loader's stream (CUDA events); the Python wrapper's
                    # own array copies are not part of the C-ABI call and are reported separately
                    "e2e_records_per_s": vst["records"] / (c_abi_ms * 1e-3), "e2e_ms": c_abi_ms,
                    "h2d_ms": vst["h2d_ms"], "kernel_ms": vst["kernel_ms"], "d2h_ms": vst["d2h_ms"], "gpu_launches": vst["launches"],
                    "python_call_ms": float(np.mean(vt)) * 1e3,
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                 "algorithmic_bytes": "text in + vce_variants arrays out", "peak_source": peak_src}}
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import vcf_check
            n_cpu = min(100_000, args.vcf_records)
            head = "\n".join(text.split("\n", n_cpu + 4)[:n_cpu + 4]) + "\n"
            t0 = time.perf_counter()
            want, _ = vcf_check.expected_side(head, "chr1", vtl, 0)
            dt = time.perf_counter() - t0
            got, _ = eng.lib.vcf_parse(head.encode(), "chr1", vtl, 0)
            vcf_check.assert_same_side(got, want, "bench: VCF loader")
            if not args.no_pin:
                eng.lib.lib.vce_host_unregister.restype = ctypes.c_int
                eng.lib.lib.vce_host_unregister.argtypes = [ctypes.c_void_p]
                eng.lib.lib.vce_host_unregister(raw.ctypes.data)
            vcf_load["cpu_baseline"] = {"value": n_cpu / dt, "unit": "records/s", "cores": 1, "kind": "port",
                                        "sample": "the first %d records of the text through oracle/ref_host.py (Python restatement of "
                                                  "VcfSortRefiner + VcfRecordTabixCallable + SampleVariants + Allele.getAlleles), %.1f s" % (n_cpu, dt),
                                        "parity": "bit-exact on those records"}

    formats_load = None
    if vcf_load is not None and not args.no_formats:
        formats_load = formats_leg(eng, args, text, vtl, peak, peak_src)

    value = total_variants / (ms_per_step * 1e-3)
    out = {
        "metric": "vcfeval variants evaluated/sec", "value": value, "unit": "variants/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, desc, total_variants),
        "sync_regions_per_step": int(total_regions),
        "run": {"sequences_per_rank": len(seqs),
                "sharding": "none (1 GPU)" if world == 1 else ("one sample pair per rank" if args.scaling == "weak" else "one sample pair, its sequences LPT-sharded over the ranks by variant count"),
                "host_threads_per_rank": int(os.environ.get("VCE_THREADS", "0")),
                "per_rank": per_rank},
        "sync_regions_per_s": total_regions / (ms_per_step * 1e-3),
        "wall_ms_per_step": wall_per_step,
        "e2e": {"value": total_variants / (e2e_ms * 1e-3), "unit": "variants/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": sub_stats["h2d_bytes"], "d2h_bytes_per_step": sub_stats["d2h_bytes"],
                "gpu_launches_per_step": sub_stats["launches"],
                "input_bytes_page_locked_by_caller": int(pinned_bytes), "reference_resident_in_hbm": bool(registered), "reference_register_ms_once": t_reg * 1e3 if registered else None},
        "gpu_launches": int(total_launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "roc": roc,
        "roc_heavy": roc_heavy,
        "vcf_load": vcf_load,
        "formats_load": formats_load,
        "summary_counts": [int(x) for x in counts] if counts is not None else None,
        "engine_stats": {k: st[k] for k in ("regions", "escalated", "spilled", "prefix_reruns", "iterations")},
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Bundle the reference's own vcfeval end-to-end fixtures into one JSON file.

The reference (pure Java) cannot run in this image, so parity of the CPU oracle is
pinned on the fixtures its own tests hold:
  test/com/rtg/vcf/eval/resources/*  (inputs *_in_{template.fa,baseline.vcf,calls.vcf},
  expected outputs *_out_*), with the command lines listed in
  test/com/rtg/vcf/eval/VcfEvalNanoTest.java:38-161 and the *EvalSynchronizerTest.java files.

/root/reference does not exist on the GPU box, so the fixtures travel as
tests/golden/vcfeval_fixtures.json.  Run here (the build container) with:

    python tests/golden/make_golden.py

Only DATA (test inputs / expected outputs) is bundled -- no reference source code.
"""
import json
import os
import sys

REF = "/root/reference/test/com/rtg/vcf/eval/resources"

# (name, fixture id, args as given to `rtg vcfeval`, output files checked, expect-warning stderr file or None)
SCENARIOS = [
    # VcfEvalNanoTest.java:40-42
    ("small", "vcfeval_small", ["--sample", "sample1", "--vcf-score-field", "QUAL", "--roc-subset", "snp"],
     ["weighted_roc.tsv", "tp.vcf"], None),
    # :44-46
    ("small_squash", "vcfeval_small_squash/vcfeval_small",
     ["--sample", "sample1", "--vcf-score-field", "QUAL", "--roc-subset", "snp", "--squash-ploidy"],
     ["weighted_roc.tsv", "tp.vcf"], None),
    # :48-50
    ("small_alt", "vcfeval_small_alt/vcfeval_small",
     ["--sample", "ALT", "--vcf-score-field", "QUAL", "--output-mode", "annotate", "--squash-ploidy"],
     ["baseline.vcf", "calls.vcf"], None),
    # :52-57
    ("tricky", "vcfeval_small_tricky/vcfeval_small_tricky", ["--vcf-score-field", "QUAL", "-T", "1", "--ref-overlap"],
     ["weighted_roc.tsv", "tp.vcf"], None),
    # :59-64
    ("tricky2", "vcfeval_small_tricky2", ["--vcf-score-field", "QUAL", "-T", "1", "--ref-overlap", "--sample", "ALT,sample"],
     ["weighted_roc.tsv", "tp.vcf"], "err.txt"),
    # :66-69
    ("tricky3", "vcfeval_small_tricky3",
     ["--vcf-score-field", "QUAL", "-T", "1", "--sample", "dummy,sample1", "--squash-ploidy", "--ref-overlap"],
     ["weighted_roc.tsv", "tp.vcf"], None),
    # :71-75
    ("tricky4", "vcfeval_small_tricky4", ["--vcf-score-field", "QUAL", "-T", "1"], ["tp.vcf", "tp-baseline.vcf"], None),
    # :77-79
    ("obeyphasing", "obeyphasing/obeyphasing", ["--Xobey-phase=true"], ["tp.vcf"], None),
    # :81-83
    ("small_region", "vcfeval_small_region",
     ["--sample", "sample1,sample1", "--vcf-score-field", "QUAL", "--region", "chr2:150-1000"], ["weighted_roc.tsv"], None),
    # :85-87
    ("small_samples", "vcfeval_small_samples", ["--sample", "sample2,sample1", "--vcf-score-field", "QUAL"],
     ["weighted_roc.tsv"], None),
    # :89-95
    ("too_complex", "vcfeval_too_complex", [], ["weighted_roc.tsv"], "err.txt"),
    ("too_complex_end", "vcfeval_too_complex_end", [], ["weighted_roc.tsv"], "err.txt"),
    # :103-124
    ("updel", "vcfeval_small_updel/updel", ["--vcf-score-field", "QUAL", "--output-mode", "annotate", "--squash-ploidy"],
     ["baseline.vcf", "calls.vcf"], None),
    ("updel2", "vcfeval_small_updel/updel2", ["--vcf-score-field", "QUAL", "--output-mode", "annotate", "--ref-overlap"],
     ["baseline.vcf", "calls.vcf"], None),
    ("updel3", "vcfeval_small_updel/updel3", ["--vcf-score-field", "QUAL", "--output-mode", "annotate"],
     ["baseline.vcf", "calls.vcf"], None),
    # :126-130
    ("updel4", "vcfeval_small_updel/updel4", ["--vcf-score-field", "QUAL", "--output-mode", "annotate"],
     ["baseline.vcf", "calls.vcf"], None),
    ("updel5", "vcfeval_small_updel/updel5", ["--vcf-score-field", "QUAL", "--output-mode", "annotate"],
     ["baseline.vcf", "calls.vcf"], None),
    # :132-137
    ("trickysquash", "vcfeval_small_trickysquash/trickysquash",
     ["--vcf-score-field", "QUAL", "--output-mode", "annotate", "--squash-ploidy",
      "--XXcom.rtg.vcf.eval.haploid-allele-matching=false"], ["baseline.vcf", "calls.vcf"], None),
    # :139-144 (combine output not regenerated; summary.txt is)
    ("avoid_overlap", "vcfeval_avoid_overlap/avoid_overlap",
     ["--no-roc", "--output-mode=combine", "--sample=BASE,QUERY", "--ref-overlap"], ["summary.txt", "output.vcf"], None),
    # :154-161
    ("triploid", "vcfeval_triploid/small",
     ["--output-mode=combine", "--no-roc", "--sample-ploidy=3", "--Xobey-phase=true,false", "--ref-overlap"],
     ["output.vcf"], None),
    ("tetraploid", "vcfeval_tetraploid/small",
     ["--output-mode=annotate", "--no-roc", "--sample-ploidy=4", "--Xobey-phase=true,false"],
     ["baseline.vcf", "calls.vcf"], None),
    # AnnotatingEvalSynchronizerTest.java:40
    ("annotate", "vcfeval_annotate", ["--ref-overlap", "--output-mode", "annotate"], ["baseline.vcf", "calls.vcf"], None),
    # CombinedEvalSynchronizerTest.java:40-44
    ("combine", "vcfeval_combine", ["--ref-overlap", "--output-mode", "combine"], ["output.vcf"], None),
    ("combine2", "vcfeval_combine2",
     ["--sample", "sample1,sample2", "--ref-overlap", "--output-mode", "combine", "--squash-ploidy"], ["output.vcf"], None),
    # Ga4ghEvalSynchronizerTest.java:40-44
    ("ga4gh", "vcfeval_ga4gh", ["--ref-overlap", "--output-mode", "ga4gh"], ["output.vcf"], None),
    ("ga4gh2", "vcfeval_ga4gh2",
     ["--sample", "sample1,sample2", "--ref-overlap", "--output-mode", "ga4gh", "--squash-ploidy", "--all-records"],
     ["output.vcf"], None),
    # RocOnlyEvalSynchronizerTest.java:40
    ("roc_only", "vcfeval_roc_only", ["--output-mode", "roc-only", "--sample", "sample1", "--vcf-score-field", "QUAL"],
     ["weighted_roc.tsv"], None),
    # SplitEvalSynchronizerTest.java:253-257
    ("split", "vcfeval_split",
     ["--ref-overlap", "--output-mode", "split", "--sample", "sample1,sample1", "--vcf-score-field", "QUAL"],
     ["tp-baseline.vcf", "fn.vcf", "tp.vcf", "fp.vcf", "summary.txt", "phasing.txt"], None),
    ("split_no_roc", "vcfeval_split_no_roc", ["--ref-overlap", "--output-mode", "split", "--sample", "ALT"],
     ["tp-baseline.vcf", "fn.vcf", "tp.vcf", "fp.vcf", "summary.txt", "phasing.txt"], "err.txt"),
]


def read(path):
    with open(path, "r") as f:
        return f.read()


def main():
    if not os.path.isdir(REF):
        sys.exit("reference fixtures not found at " + REF)
    out = {"source": "test/com/rtg/vcf/eval/resources (RealTimeGenomics/rtg-tools)", "scenarios": {}}
    for name, fid, args, outs, err in SCENARIOS:
        base = os.path.join(REF, fid)
        sc = {
            "fixture": fid,
            "args": args,
            "template_fa": read(base + "_in_template.fa"),
            "baseline_vcf": read(base + "_in_baseline.vcf"),
            "calls_vcf": read(base + "_in_calls.vcf"),
            "expected": {},
        }
        for o in outs:
            p = base + "_out_" + o
            if os.path.exists(p):
                sc["expected"][o] = read(p)
        if err:
            p = base + "_" + err
            if os.path.exists(p):
                sc["expected_stderr"] = read(p)
        out["scenarios"][name] = sc
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vcfeval_fixtures.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", dst, "with", len(out["scenarios"]), "scenarios,", os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()

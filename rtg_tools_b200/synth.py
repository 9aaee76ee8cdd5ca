"""Seeded synthetic genomes and variant sets (no JVM here, so this stands in for `rtg genomesim` /
`popsim` / `samplesim`), emitted directly in the flat batch format of include/vce.h.

Priors follow src/com/rtg/variant/priors/prior/human.properties of the reference as summarised in
SURVEY.md section 8(d): SNP 1.24e-3 /bp (57 % heterozygous), MNP 9.2e-5 /bp, indel 1.5e-4 /bp, the
published length distributions, phased diploid genotypes and no overlapping variants within a sample
(simulation/variants/SampleSimulator.java:216,234).  Calls are the baseline perturbed with dropped /
novel variants, re-represented indels and MNPs, and zygosity flips.

Everything is numpy-vectorised so that the 3 Gbp / 5 M-variant configuration builds in seconds.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import capi

INDEL_LEN_P = np.array([0.566, 0.226, 0.083, 0.044, 0.023, 0.017, 0.011, 0.009, 0.007, 0.014])
INDEL_LEN_P = INDEL_LEN_P / INDEL_LEN_P.sum()
MNP_LEN_P = np.array([0.40, 0.20, 0.15, 0.10, 0.08, 0.07])  # lengths 2..7
MNP_LEN_P = MNP_LEN_P / MNP_LEN_P.sum()

# GRCh38 chr1-22, X, Y lengths (Mbp, rounded) -- only the proportions matter
GRCH38_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51,
              156, 57]


@dataclass
class VarSet:
    """Biallelic variants of one sample on one sequence, sorted by start, in 'trimmed' coordinates:
    the REF allele occupies [start, start + ref_len) and is replaced by alt bases."""
    start: np.ndarray      # int32
    ref_len: np.ndarray    # int32
    alt_off: np.ndarray    # int32 [n+1]
    alt: np.ndarray        # uint8 bases 1..4
    gt: np.ndarray         # int8 (n, 2), values 0/1
    phased: np.ndarray     # uint8
    qual: np.ndarray       # float64

    @property
    def n(self):
        return int(self.start.shape[0])

    def take(self, idx):
        idx = np.asarray(idx, np.int64)
        lens = (self.alt_off[1:] - self.alt_off[:-1])[idx]
        off = np.zeros(len(idx) + 1, np.int32)
        np.cumsum(lens, out=off[1:])
        src = _ranges(self.alt_off[:-1][idx], lens)
        return VarSet(self.start[idx], self.ref_len[idx], off, self.alt[src], self.gt[idx], self.phased[idx], self.qual[idx])

    @staticmethod
    def concat(sets):
        lens = np.concatenate([s.alt_off[1:] - s.alt_off[:-1] for s in sets])
        off = np.zeros(len(lens) + 1, np.int32)
        np.cumsum(lens, out=off[1:])
        return VarSet(np.concatenate([s.start for s in sets]), np.concatenate([s.ref_len for s in sets]), off,
                      np.concatenate([s.alt for s in sets]), np.concatenate([s.gt for s in sets]),
                      np.concatenate([s.phased for s in sets]), np.concatenate([s.qual for s in sets]))

    def sorted(self):
        order = np.lexsort((self.start + self.ref_len, self.start))
        return self.take(order)


def _ranges(starts, lens):
    """Concatenated aranges start[i] .. start[i]+len[i]."""
    starts = np.asarray(starts, np.int64)
    lens = np.asarray(lens, np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    ends = np.cumsum(lens)
    idx = np.arange(total, dtype=np.int64)
    seg = np.repeat(np.arange(len(lens)), lens)
    return idx - np.repeat(ends - lens, lens) + np.repeat(starts, lens)


def make_genome(lengths, seed=42):
    """genomesim --freq 1,1,1,1: uniform ACGT (simulation/genome/GenomeSimulator.java:97)."""
    rng = np.random.default_rng(seed)
    return [rng.integers(1, 5, size=int(n), dtype=np.uint8) for n in lengths]


def genome_lengths(total_bp, n_chrom=24):
    props = np.array(GRCH38_MBP[:n_chrom], np.float64)
    return [max(1000, int(total_bp * p / props.sum())) for p in props]


def simulate_sample(template, rng, snp_rate=1.24e-3, mnp_rate=9.2e-5, indel_rate=1.5e-4, het_frac=0.57) -> VarSet:
    n = int(template.shape[0])
    rate = snp_rate + mnp_rate + indel_rate
    k = rng.poisson(n * rate)
    pos = np.sort(rng.integers(1, max(2, n - 16), size=k)).astype(np.int64)
    kind = rng.choice(3, size=k, p=np.array([snp_rate, mnp_rate, indel_rate]) / rate)
    ref_len = np.ones(k, np.int64)
    alt_len = np.ones(k, np.int64)
    m = kind == 1
    ln = rng.choice(np.arange(2, 2 + len(MNP_LEN_P)), size=int(m.sum()), p=MNP_LEN_P)
    ref_len[m] = ln
    alt_len[m] = ln
    d = kind == 2
    il = rng.choice(np.arange(1, 1 + len(INDEL_LEN_P)), size=int(d.sum()), p=INDEL_LEN_P)
    is_ins = rng.random(int(d.sum())) < 0.5
    ref_len[d] = np.where(is_ins, 0, il)
    alt_len[d] = np.where(is_ins, il, 0)
    # no overlapping variants within a sample (keep one base between neighbours)
    end = pos + np.maximum(ref_len, 1)
    keep = np.ones(k, bool)
    last_end = -1
    # vectorised greedy: a variant is kept if it starts after the previous kept variant's end + 1; the running
    # maximum of ends is a safe (slightly conservative) stand-in for "previous kept end"
    prev_end = np.concatenate([[-1], np.maximum.accumulate(end)[:-1]])
    keep = pos > prev_end + 1
    keep &= end < n - 1
    pos, kind, ref_len, alt_len = pos[keep], kind[keep], ref_len[keep], alt_len[keep]
    k = int(pos.shape[0])
    off = np.zeros(k + 1, np.int64)
    np.cumsum(alt_len, out=off[1:])
    alt = rng.integers(1, 5, size=int(off[-1]), dtype=np.uint8)
    # SNP / MNP alt bases must differ from the reference at every position
    sub = np.repeat(kind != 2, alt_len)
    tpos = _ranges(pos, alt_len)
    refb = template[np.minimum(tpos, n - 1)]
    clash = sub & (alt == refb)
    alt[clash] = (refb[clash] % 4) + 1
    het = rng.random(k) < het_frac
    gt = np.ones((k, 2), np.int8)
    which = rng.random(k) < 0.5
    gt[het & which, 0] = 0
    gt[het & ~which, 1] = 0
    qual = np.round(rng.uniform(1, 100, size=k), 1)
    return VarSet(pos.astype(np.int32), ref_len.astype(np.int32), off.astype(np.int32), alt, gt, np.ones(k, np.uint8), qual)


def perturb(template, base: VarSet, rng, drop=0.05, novel=0.05, rerep=0.10, zyg=0.05, unphase=0.5) -> VarSet:
    """Calls = baseline with dropped / novel variants, re-represented indels + MNPs and zygosity flips."""
    n = int(template.shape[0])
    k = base.n
    keep = rng.random(k) >= drop
    calls = base.take(np.nonzero(keep)[0])
    k = calls.n
    alt_len = calls.alt_off[1:] - calls.alt_off[:-1]
    # zygosity flips
    flip = rng.random(k) < zyg
    gt = calls.gt.copy()
    het = gt[:, 0] != gt[:, 1]
    gt[flip & het] = 1
    fh = flip & ~het
    gt[fh, 0] = 0
    calls.gt = gt
    calls.phased = (rng.random(k) >= unphase).astype(np.uint8)
    calls.qual = np.round(rng.uniform(1, 100, size=k), 1)
    # re-representation 1: MNP -> separate SNPs
    is_mnp = (calls.ref_len > 1) & (alt_len == calls.ref_len)
    split = is_mnp & (rng.random(k) < rerep)
    parts = [calls.take(np.nonzero(~split)[0])]
    if split.any():
        idx = np.nonzero(split)[0]
        lens = alt_len[idx].astype(np.int64)
        rep = np.repeat(idx, lens)
        within = _ranges(np.zeros(len(idx), np.int64), lens)
        s_start = calls.start[rep] + within
        s_alt = calls.alt[calls.alt_off[:-1][rep] + within]
        m = len(rep)
        off = np.arange(m + 1, dtype=np.int32)
        parts.append(VarSet(s_start.astype(np.int32), np.ones(m, np.int32), off, s_alt, calls.gt[rep], calls.phased[rep],
                            calls.qual[rep]))
    # re-representation 2: shift indels inside their repeat context (left-aligned <-> right-aligned)
    cur = VarSet.concat(parts).sorted()
    al = cur.alt_off[1:] - cur.alt_off[:-1]
    is_indel = (cur.ref_len != al) & ((cur.ref_len == 0) | (al == 0))
    cand = np.nonzero(is_indel & (rng.random(cur.n) < rerep))[0]
    start = cur.start.copy()
    alt = cur.alt.copy()
    next_start = np.concatenate([cur.start[1:], [n]])
    for i in cand:  # a few percent of the indels: python loop is fine
        L = int(max(cur.ref_len[i], al[i]))
        p = int(start[i])
        if cur.ref_len[i] > 0:  # deletion of template[p:p+L): can slide right while template[p] == template[p+L]
            q = p
            lim = int(next_start[i]) - L - 2
            while q < lim and q + L < n - 2 and template[q] == template[q + L]:
                q += 1
            start[i] = q
        else:  # insertion of S before p: slide right while S[0] == template[p]; rotate S
            a0 = int(cur.alt_off[i])
            s = alt[a0:a0 + L].copy()
            q = p
            lim = int(next_start[i]) - 2
            while q < lim and q < n - 2 and s[0] == template[q]:
                s = np.roll(s, -1)
                q += 1
            start[i] = q
            alt[a0:a0 + L] = s
    cur.start = start
    cur.alt = alt
    # novel variants
    extra = simulate_sample(template, rng, snp_rate=1.24e-3 * novel, mnp_rate=9.2e-5 * novel, indel_rate=1.5e-4 * novel)
    extra.phased = (rng.random(extra.n) >= unphase).astype(np.uint8)
    allv = VarSet.concat([cur, extra]).sorted()
    # drop novel variants that collide with an existing one (keeps the sample free of self-overlaps; directly
    # adjacent records, e.g. the SNPs of a split MNP, are kept)
    end = allv.start + np.maximum(allv.ref_len, 1)
    prev_end = np.concatenate([[-1], np.maximum.accumulate(end)[:-1]])
    ok = allv.start >= prev_end
    return allv.take(np.nonzero(ok)[0])


def to_side(vs: VarSet, template) -> capi.VariantSide:
    """Flat vce_variants image: 3 allele slots per variant (missing = null, REF, ALT), GT-indexed (Variant.java:241)."""
    n = vs.n
    alt_len = (vs.alt_off[1:] - vs.alt_off[:-1]).astype(np.int64)
    ref_len = vs.ref_len.astype(np.int64)
    start = vs.start.astype(np.int64)
    a_start = np.zeros((n, 3), np.int32)
    a_end = np.zeros((n, 3), np.int32)
    a_null = np.zeros((n, 3), np.uint8)
    a_null[:, 0] = 1
    a_start[:, 1] = start
    a_start[:, 2] = start
    a_end[:, 1] = start + ref_len
    a_end[:, 2] = start + ref_len
    lens = np.zeros((n, 3), np.int64)
    lens[:, 1] = ref_len
    lens[:, 2] = alt_len
    nt_off = np.zeros(3 * n + 1, np.int32)
    np.cumsum(lens.reshape(-1), out=nt_off[1:])
    nt = np.zeros(int(nt_off[-1]), np.uint8)
    ref_dst = _ranges(nt_off[1:-1:3], ref_len) if n else np.zeros(0, np.int64)
    ref_src = _ranges(start, ref_len)
    nt[ref_dst] = template[ref_src]
    alt_dst = _ranges(nt_off[2:-1:3], alt_len) if n else np.zeros(0, np.int64)
    nt[alt_dst] = vs.alt
    return capi.VariantSide(
        id=np.arange(1, n + 1, dtype=np.int32), phased=vs.phased.astype(np.uint8), status_in=np.zeros(n, np.uint8),
        gt=np.ascontiguousarray(vs.gt, np.int8), allele_off=np.arange(0, 3 * n + 1, 3, dtype=np.int32),
        allele_start=a_start.reshape(-1), allele_end=a_end.reshape(-1), allele_null=a_null.reshape(-1),
        allele_nt_off=nt_off, nt=nt)


def make_sequence(name, length, genome_seed, sample_seed, perturb_seed, return_qual=False, **kw):
    """One chromosome: uniform ACGT template, baseline sample, perturbed calls (each from its own seed stream)."""
    tmpl = np.random.default_rng(genome_seed).integers(1, 5, size=int(length), dtype=np.uint8)
    base = simulate_sample(tmpl, np.random.default_rng(sample_seed))
    calls = perturb(tmpl, base, np.random.default_rng(perturb_seed), **kw)
    seq = capi.Sequence(name, tmpl, to_side(base, tmpl), to_side(calls, tmpl))
    return (seq, calls.qual) if return_qual else seq


def make_pair(total_bp, n_chrom=1, genome_seed=42, sample_seed=572, perturb_seed=126, only=None, return_qual=False, **kw):
    """BASELINE.md C1 (total_bp=10e6, n_chrom=1) / C2 (total_bp=3e9, n_chrom=24).  Every chromosome draws from its own
    (seed, chromosome) streams, so `only=[...]` builds any subset identically to the full genome."""
    lens = genome_lengths(total_bp, n_chrom) if n_chrom > 1 else [int(total_bp)]
    out, quals = [], []
    for c, ln in enumerate(lens):
        if only is not None and c not in only:
            continue
        seq, q = make_sequence("chr%d" % (c + 1), ln, (genome_seed, c), (sample_seed, c), (perturb_seed, c), True, **kw)
        out.append(seq)
        quals.append(q)
    return (out, quals) if return_qual else out


def qualities(seqs_calls: List[VarSet]):
    return np.concatenate([v.qual for v in seqs_calls])


# ------------------------------------------------------------------------------------------------- fuzz
def fuzz_sequence(rng, length=400, n_base=12, n_calls=12, repeat=True, multi=0.15, null_frac=0.05, half=0.05,
                  ploidy=2, cluster=None, name="fz") -> capi.Sequence:
    """Small adversarial sequence for parity fuzzing: low-complexity template, overlapping / adjacent variants,
    multi-allelic sites, half calls, null (symbolic) alleles, unsorted same-position records."""
    if repeat:
        unit = rng.integers(1, 5, size=int(rng.integers(1, 4)), dtype=np.uint8)
        tmpl = np.tile(unit, length // len(unit) + 1)[:length].copy()
        noise = rng.random(length) < 0.15
        tmpl[noise] = rng.integers(1, 5, size=int(noise.sum()), dtype=np.uint8)
    else:
        tmpl = rng.integers(1, 5, size=length, dtype=np.uint8)
    if rng.random() < 0.2:
        tmpl[rng.integers(0, length, size=3)] = 0  # a few N

    def side(nv):
        vs = []
        lo, hi = (2, length - 12) if cluster is None else (max(2, cluster[0]), min(length - 12, cluster[1]))
        starts = np.sort(rng.integers(lo, max(lo + 1, hi), size=nv))
        for i, st in enumerate(starts):
            st = int(st)
            n_alt = 2 if rng.random() < multi else 1
            kind = rng.integers(0, 4)
            ref_len = 1 if kind == 0 else int(rng.integers(0, 4)) if kind in (1, 2) else int(rng.integers(2, 5))
            if st + ref_len >= length - 2:
                ref_len = 1
            alleles = [None, (st, st + ref_len, bytes(tmpl[st:st + ref_len]))]
            for _ in range(n_alt):
                if kind == 0:
                    alt = bytes([int((tmpl[st] % 4) + 1 if rng.random() < 0.7 else rng.integers(1, 5))])
                elif kind == 3:
                    alt = bytes(rng.integers(1, 5, size=ref_len, dtype=np.uint8))
                else:
                    al = int(rng.integers(0, 4))
                    if al == ref_len:
                        al = ref_len + 1
                    # repeat-consistent inserted sequence makes equivalent representations likely
                    alt = bytes(tmpl[st:st + al]) if rng.random() < 0.6 and st + al < length else bytes(rng.integers(1, 5, size=al, dtype=np.uint8))
                if alt == alleles[1][2]:
                    alt = bytes([int((tmpl[st] % 4) + 1)]) + alt
                if rng.random() < null_frac:
                    alleles.append(None)
                else:
                    alleles.append((st, st + ref_len, alt))
            ids = list(range(0, n_alt + 1))
            gt = [int(rng.choice(ids)) for _ in range(ploidy)]
            if all(g == 0 for g in gt) or all(g <= 0 or alleles[g + 1] is None for g in gt):
                gt[0] = 1
                if alleles[2] is None:
                    alleles[2] = (st, st + ref_len, bytes([int((tmpl[st] % 4) + 1)]) + alleles[1][2])
            if rng.random() < half:
                gt[-1] = -1
            # homozygous no-op variants are rejected at load time in the reference; avoid them
            vs.append({"id": i + 1, "alleles": alleles, "gt": gt, "phased": bool(rng.random() < 0.5), "status": 0})
        return capi.VariantSide.from_variants(vs, ploidy)

    return capi.Sequence(name, tmpl, side(n_base), side(n_calls))


def dense_cluster_sequence(rng, n_clusters=4, cluster_len=500, per_side=(12, 40), spacing=3000, name="dense",
                           long_frac=0.0, return_qual=False):
    """clusters of overlapping indels / MNPs in short tandem repeats on both sides.  `long_frac` of the
    variants are long alleles (100 .. 1000 bases, --Xmax-length, VcfEvalCli.java:163): deletions / insertions of whole
    repeat units reaching far beyond a search window of the small kernels."""
    length = n_clusters * spacing + 1000
    tmpl = rng.integers(1, 5, size=length, dtype=np.uint8)
    bvs, cvs = [], []
    for c in range(n_clusters):
        c0 = 500 + c * spacing
        unit = rng.integers(1, 5, size=int(rng.integers(2, 7)), dtype=np.uint8)
        tmpl[c0:c0 + cluster_len] = np.tile(unit, cluster_len // len(unit) + 1)[:cluster_len]
        for out in (bvs, cvs):
            nv = int(rng.integers(per_side[0], per_side[1] + 1))
            starts = np.sort(rng.choice(np.arange(c0 + 2, c0 + cluster_len - 10), size=nv, replace=False))
            for st in starts:
                st = int(st)
                is_del = rng.random() < 0.5
                L = int(rng.integers(1, 3)) * len(unit)
                if long_frac > 0 and rng.random() < long_frac:
                    L = (int(rng.integers(100, 1001)) // len(unit)) * len(unit)
                if is_del:
                    ref = bytes(tmpl[st:st + L])
                    alleles = [None, (st, st + L, ref), (st, st + L, b"")]
                else:
                    ins = np.tile(tmpl[st:st + len(unit)], L // len(unit) + 1)[:L] if L > 2 * len(unit) else tmpl[st:st + L]
                    alleles = [None, (st, st, b""), (st, st, bytes(ins))]
                gt = [0, 1] if rng.random() < 0.5 else [1, 0]
                out.append({"id": len(out) + 1, "alleles": alleles, "gt": gt, "phased": False, "status": 0})
    seq = capi.Sequence(name, tmpl, capi.VariantSide.from_variants(bvs, 2), capi.VariantSide.from_variants(cvs, 2))
    if return_qual:
        return seq, np.round(rng.uniform(1, 100, size=len(cvs)), 1)
    return seq


def make_c3(n_clusters=2000, per_seq=50, seed=3, per_side=(12, 40), long_frac=0.01, only=None, return_qual=False):
    """BASELINE.md C3 (BASELINE.json configs[2]): `n_clusters` 500 bp tandem-repeat clusters with 12-40 overlapping
    heterozygous indels per side, alleles up to 1000 bases, spread over sequences of `per_seq` clusters each (sequences are
    the reference's unit of parallelism, VcfEvalTask.java:131-135).  With these sizes every cluster exceeds the 50 000-path
    budget at least once (ToolsGlobalFlags.java:139): the run exercises the too-complex skip semantics throughout.
    Every sequence draws from its own (seed, index) stream: `only=[...]` builds any subset identically."""
    n_seq = (n_clusters + per_seq - 1) // per_seq
    out, quals = [], []
    for k in range(n_seq):
        if only is not None and k not in only:
            continue
        ncl = min(per_seq, n_clusters - k * per_seq)
        seq, q = dense_cluster_sequence(np.random.default_rng((seed, k)), n_clusters=ncl, per_side=per_side, name="c3_%d" % k,
                                        long_frac=long_frac, return_qual=True)
        out.append(seq)
        quals.append(q)
    return (out, quals) if return_qual else out


def make_vcf_text(n_records, seed=7, name="chr1", mean_gap=700, samples=("s1", "s2")):
    """A sorted single-sequence VCF (text, template_len) with the variant mix of `simulate_sample` (SNPs, MNPs, insertions and
    deletions with their padding base, a few multi-allelic records, symbolic ALTs and filtered records): input of the VCF
    loader (vce_vcf_parse), not tied to any template sequence."""
    rng = np.random.default_rng(seed)
    gaps = rng.geometric(1.0 / mean_gap, n_records).astype(np.int64)
    pos = np.cumsum(gaps) + 1
    kind = rng.choice(6, n_records, p=[0.72, 0.06, 0.08, 0.08, 0.04, 0.02])   # snp, mnp, ins, del, multi, symbolic
    base = np.array(list("ACGT"))
    r0 = rng.integers(0, 4, n_records)
    a0 = (r0 + rng.integers(1, 4, n_records)) % 4
    ln = rng.choice([1, 1, 2, 3, 5, 9], n_records)
    tail = rng.integers(0, 4, (n_records, 9))
    gt_het = rng.random(n_records) < 0.57
    phased = rng.random(n_records) < 0.5
    filt = rng.random(n_records) < 0.03
    lines = ["##fileformat=VCFv4.2", "##contig=<ID=%s,length=%d>" % (name, int(pos[-1]) + 1000),
             "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">",
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples)]
    for i in range(n_records):
        k = kind[i]
        rb, ab = base[r0[i]], base[a0[i]]
        t = "".join(base[tail[i, :ln[i]]])
        if k == 0:
            ref, alt = rb, ab
        elif k == 1:
            ref, alt = rb + t, ab + t[::-1] + ("" if len(t) == 0 else "")
            if alt == ref:
                alt = ab + t
        elif k == 2:
            ref, alt = rb, rb + t
        elif k == 3:
            ref, alt = rb + t, rb
        elif k == 4:
            ref, alt = rb, ab + "," + rb + t
        else:
            ref, alt = rb, "<DEL>"
        sep = "|" if phased[i] else "/"
        g1 = ("0" + sep + "1") if gt_het[i] else ("1" + sep + "1")
        if k == 4:
            g1 = "1" + sep + "2"
        g2 = "0" + sep + "1" if (i & 3) else "."
        lines.append("%s\t%d\t.\t%s\t%s\t%d\t%s\t.\tGT\t%s\t%s" % (name, pos[i], ref, alt, 10 + (i % 80), "q10" if filt[i] else "PASS", g1, g2))
    return "\n".join(lines) + "\n", int(pos[-1]) + 1000

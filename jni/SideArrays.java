package com.rtg.vcf.eval;

import java.util.List;

/**
 * Structure-of-arrays image of one side (baseline or calls) of one sequence: the data a {@code List<Variant>} holds after
 * loading, in load order (ascending {@code VariantId.getId()}), exactly as {@code vce_variants} of {@code include/vce.h}
 * expects it.  Replaces nothing in the reference: it is the flattening a maintainer adds next to
 * {@code SequenceEvaluator.run()} (src/com/rtg/vcf/eval/SequenceEvaluator.java:87-92).
 */
final class SideArrays {
  int mN;
  int mGtPloidy;           // entries per variant in mGt; 0 = no genotype (AllAlts factory, VariantFactory.java:222-239)
  int[] mId;               // VariantId.getId()
  byte[] mPhased;          // Variant.isPhased()
  byte[] mStatusIn;        // pre-set status bits (VariantId.STATUS_OUTSIDE_EVAL), VcfRecordTabixCallable.java:131
  byte[] mGt;              // GtIdVariant.alleleIds(), -1 = missing
  int[] mAlleleOff;        // [n + 1] slot range per variant; slot k = allele id k - 1 (Variant.mAlleles, Variant.java:241)
  int[] mAlleleStart;      // Allele.getStart()
  int[] mAlleleEnd;        // Allele.getEnd()
  byte[] mAlleleNull;      // 1 = Java null Allele ("play nothing on this haplotype")
  int[] mAlleleNtOff;      // [slots + 1]
  byte[] mNt;              // concatenated Allele.nt() (0..4, DnaUtils encoding; 5 = explicit unknown)
  // filled by NativeVce.loadVcfGz only: what RocContainer.addRocLine derives from each variant's record (vce_vcf_roc_fields)
  double[] mScore;         // RocSortValueExtractor.getSortValue: QUAL or the sample's FORMAT field; NaN = missing
  int[] mRocMask;          // bit per RocFilter in declaration order (ALL HOM HET SNP NON_SNP MNP INDEL XRX NON_XRX); bit 31: parse mScore in Java

  /** Flattens {@code variants}; {@code gtPloidy} is the width of {@code GtIdVariant.alleleIds()} (0 for plain Variants). */
  static SideArrays flatten(List<? extends Variant> variants, int gtPloidy) {
    final SideArrays s = new SideArrays();
    final int n = variants.size();
    s.mN = n;
    s.mGtPloidy = gtPloidy;
    s.mId = new int[n];
    s.mPhased = new byte[n];
    s.mStatusIn = new byte[n];
    s.mGt = new byte[n * Math.max(gtPloidy, 0)];
    s.mAlleleOff = new int[n + 1];
    int slots = 0;
    int bases = 0;
    for (final Variant v : variants) {
      final int na = v.numAlleles();                 // ids 0 .. na - 1, plus the missing slot
      slots += na + 1;
      for (int id = -1; id < na; ++id) {           // allele(-1) is the missing slot: null, or UNKNOWN with explicit unknowns
        final Allele a = v.allele(id);
        if (a != null) {
          bases += a.nt().length;
        }
      }
    }
    s.mAlleleStart = new int[slots];
    s.mAlleleEnd = new int[slots];
    s.mAlleleNull = new byte[slots];
    s.mAlleleNtOff = new int[slots + 1];
    s.mNt = new byte[bases];
    int slot = 0;
    int nt = 0;
    for (int i = 0; i < n; ++i) {
      final Variant v = variants.get(i);
      s.mId[i] = v.getId();
      s.mPhased[i] = (byte) (v.isPhased() ? 1 : 0);
      s.mStatusIn[i] = (byte) (v.hasStatus(VariantId.STATUS_OUTSIDE_EVAL) ? VariantId.STATUS_OUTSIDE_EVAL : 0);
      if (gtPloidy > 0) {
        final int[] gt = ((GtIdVariant) v).alleleIds();
        for (int h = 0; h < gtPloidy; ++h) {
          s.mGt[i * gtPloidy + h] = (byte) gt[h];
        }
      }
      s.mAlleleOff[i] = slot;
      for (int id = -1; id < v.numAlleles(); ++id) {   // slot 0 = the missing allele (-1), slot 1 = REF
        final Allele a = v.allele(id);
        s.mAlleleNtOff[slot] = nt;
        if (a == null) {
          s.mAlleleNull[slot] = 1;
        } else {
          s.mAlleleStart[slot] = a.getStart();
          s.mAlleleEnd[slot] = a.getEnd();
          final byte[] b = a.nt();
          System.arraycopy(b, 0, s.mNt, nt, b.length);
          nt += b.length;
        }
        ++slot;
      }
    }
    s.mAlleleOff[n] = slot;
    s.mAlleleNtOff[slot] = nt;
    return s;
  }
}

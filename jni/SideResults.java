package com.rtg.vcf.eval;

import java.util.ArrayList;
import java.util.List;

/** Per-side results of the native engine, indexed like {@link SideArrays} ({@code vce_side_result} of include/vce.h). */
final class SideResults {
  final byte[] mStatus;      // VariantId status bits after SequenceEvaluator.mergeVariants (SequenceEvaluator.java:170-209)
  final byte[] mAlleleIds;   // [n * 4] OrientedVariant.alleleIds() of the chosen orientation, unused entries -2
  final byte[] mOriginal;    // OrientedVariant.isOriginal()
  final double[] mWeight;    // OrientedVariant.getWeight() (calls)
  final int[] mSyncId;       // InterleavingEvalSynchronizer.floorSyncPos value (:71-83), as the SYNC annotation prints it

  SideResults(int n) {
    mStatus = new byte[n];
    mAlleleIds = new byte[n * 4];
    mOriginal = new byte[n];
    mWeight = new double[n];
    mSyncId = new int[n];
  }

  /**
   * The list {@code SequenceEvaluator.mergeVariants} hands to {@code EvalSynchronizer.write} (SequenceEvaluator.java:170-209),
   * rebuilt from the engine's per-variant outputs: in load order (ascending id), a variant the best path included (GT match
   * or, in the second pass, allele match) as an {@link OrientedVariant} of the chosen orientation carrying its weight,
   * every other one as the {@link Variant} itself; the status bits are OR-ed in exactly as setStatus / insertSkipped do.
   * @param variants the side's variants in the order {@link SideArrays#flatten} saw them
   */
  List<VariantId> toVariantIds(List<? extends Variant> variants) {
    final List<VariantId> out = new ArrayList<>(variants.size());
    int i = 0;
    for (final Variant v : variants) {
      final byte st = mStatus[i];
      if ((st & (VariantId.STATUS_GT_MATCH | VariantId.STATUS_ALLELE_MATCH)) != 0) {
        int ploidy = 0;
        while (ploidy < 4 && mAlleleIds[i * 4 + ploidy] != -2) {
          ++ploidy;
        }
        final int[] ids = new int[ploidy];
        for (int h = 0; h < ploidy; ++h) {
          ids[h] = mAlleleIds[i * 4 + h];
        }
        final OrientedVariant ov = new OrientedVariant(v, mOriginal[i] != 0, ids);
        ov.setWeight(mWeight[i]);
        ov.setStatus(st);     // delegates to the variant (OrientedVariant.java:236-238), bits accumulate (Variant.java:213-215)
        out.add(ov);
      } else {
        v.setStatus(st);      // NO_MATCH, SKIPPED (too-complex region), OUTSIDE_EVAL as loaded
        out.add(v);
      }
      ++i;
    }
    return out;
  }

  /** {@code Path.getSyncPoints()} of one pass as the list {@code EvalSynchronizer.write} takes. */
  static List<Integer> list(int[] syncPoints, int n) {
    final List<Integer> out = new ArrayList<>(n);
    for (int i = 0; i < n; ++i) {
      out.add(syncPoints[i]);
    }
    return out;
  }
}

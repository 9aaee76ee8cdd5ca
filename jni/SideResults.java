package com.rtg.vcf.eval;

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
}

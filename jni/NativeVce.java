package com.rtg.vcf.eval;

/**
 * The native B200 matching engine (libvce.so behind libvce_jni.so; C-ABI in include/vce.h).
 * Replaces the body of {@code SequenceEvaluator.run()} between "variants + template loaded" and
 * {@code mSynchronize.write(...)} (src/com/rtg/vcf/eval/SequenceEvaluator.java:87-155) and, through {@link #roc}, the
 * accumulate / sort / cumulate part of {@code RocContainer} (src/com/rtg/vcf/eval/RocContainer.java:225-318).
 * Every method is thread-safe; {@code submit} may be called from the {@code SimpleThreadPool} workers
 * (VcfEvalTask.java:131-137), one sequence per call.
 */
final class NativeVce {
  static {
    System.loadLibrary("vce_jni");
  }

  private NativeVce() { }

  /** config[] layout: nPasses, baselineOrientor[2], callsOrientor[2], ploidy[2], maxPaths, maxIterations, pruneNoOps, flagAlternates, preference. */
  static final int CONFIG_INTS = 12;

  /** vce_init_devices: the CUDA devices this JVM may use (sequences of one batch are spread over them). */
  static native int init(int[] devices);

  /** vce_shutdown. */
  static native void shutdown();

  /** vce_last_error of the calling thread. */
  static native String lastError();

  /** vce_template_register: call once per sequence in the load phase (SequenceEvaluator.java:76-86); the array must stay unchanged. */
  static native int registerTemplate(byte[] template);

  /**
   * vce_submit for ONE sequence.
   * @param syncPoints [2][capacity] Path.getSyncPoints() per pass; capacity baseline.mN + calls.mN + 2 always suffices
   * @param nSync [2] out
   * @param phasing [3] out: misPhasings, correctPhasings, unphaseable (PhasingEvaluator.java:55-142)
   * @param warnings [5 * capacity] out: pass, paths, iterations, start1, end1 per "Evaluation too complex" event (PathFinder.java:162-169)
   * @param nWarnings [1] out
   * @return 0, or a VCE_ERR_* code (6 = best path null, SequenceEvaluator.java:105-116)
   */
  static native int submit(int[] config, String name, byte[] template, SideArrays baseline, SideArrays calls,
                           SideResults baselineOut, SideResults callsOut, int[][] syncPoints, int[] nSync, int[] phasing,
                           int[] warnings, int[] nWarnings);

  /**
   * vce_roc: RocContainer.addRocLine arrival-order arrays in, the TreeMap iteration of writeRocs (:296-309) out.
   * @return rows per filter in nRows; threshold / cumTp / cumFp / cumTpRaw are [nFilters * capacity]
   */
  static native int roc(double[] score, double[] tp, double[] fp, double[] tpRaw, int[] filterMask, int nFilters,
                        boolean ascending, long[] nRows, double[] threshold, double[] cumTp, double[] cumFp, double[] cumTpRaw,
                        long[] noScore);

  /**
   * vce_tabix_open: the bytes of a .tbi file (TabixIndexReader.java:63-146); the index is inflated on the device.
   * @return a handle for {@link #tabixQuery} / {@link #tabixClose}, 0 on failure ({@link #lastError} says why)
   */
  static native long tabixOpen(byte[] tbi);

  /**
   * vce_tabix_query: AbstractIndexReader.getFilePointers for one interval (AbstractIndexReader.java:102-205).
   * @param begin0 0-based start, -1 = from the start of the sequence
   * @param end0 0-based exclusive end, -1 = to the end of the sequence
   * @return {virtual offset begin, virtual offset end}, or null when the index holds nothing for the region
   */
  static native long[] tabixQuery(long index, String sequenceName, int begin0, int end0);

  /** vce_tabix_close. */
  static native void tabixClose(long index);

  /**
   * vce_vcf_parse_bgzf: what VcfRecordTabixCallable.call (VcfRecordTabixCallable.java:91-141) produces for one sequence, from the
   * bytes of the .vcf.gz between two virtual offsets: BGZF members inflated and records parsed on the device.
   * @param refOverlap --ref-overlap: Variant.trimAlleles (Variant.java:93-190) applied on the device
   * @param evalRegions null, or the sequence's merged --evaluation-regions as [start, end) pairs (MergedIntervals): variants outside get STATUS_OUTSIDE_EVAL
   * @param scoreField --vcf-score-field: null or "QUAL", else a FORMAT key of the sample ("GQ"); out.mScore / out.mRocMask receive the ROC inputs
   * @param out receives the arrays of the side (as SideArrays.flatten would have made them from the loaded variants)
   * @param counters [4] out: text lines, records of the sequence, variants kept, counted skips
   * @return 0, or a VCE_ERR_* code (2 = the reference would have thrown VcfFormatException / a corrupt BGZF member)
   */
  static native int loadVcfGz(byte[] file, long voffsetBegin, long voffsetEnd, String sequenceName, int templateLength, int sample,
                              int ploidy, boolean passOnly, int maxLength, boolean refOverlap, int[] evalRegions, String scoreField, SideArrays out,
                              long[] counters);

  /**
   * vce_template_register_sdf: one sequence of an SDF seqdata file (bit planes, FileBitwiseInputStream.java:151-169) becomes the
   * resident template; replaces SequencesReader.read + registerTemplate.
   * @return the byte-per-base array (0 = N, 1..4 = A C G T), or null on failure
   */
  static native byte[] registerTemplateSdf(byte[] seqdata, long firstResidue, int length);

  /** vce_format_warning: the text of PathFinder.java:163. */
  static native String formatWarning(int pass, int paths, int iterations, int start1, int end1, String sequenceName);
}

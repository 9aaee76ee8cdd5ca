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

  /** vce_format_warning: the text of PathFinder.java:163. */
  static native String formatWarning(int pass, int paths, int iterations, int start1, int end1, String sequenceName);
}

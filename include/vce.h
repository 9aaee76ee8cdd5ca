/*
 * vce.h -- C-ABI of the B200-native vcfeval matching engine ("vce").
 *
 * This is the drop-in boundary for the ONE hot path this repository rebuilds:
 * the body of rtg-tools' SequenceEvaluator.run() between "variants + template
 * loaded" and "mSynchronize.write(...)"
 *   (reference: src/com/rtg/vcf/eval/SequenceEvaluator.java:87-155),
 * i.e. PathFinder.bestPath() (PathFinder.java:129-214) with its Orientor /
 * Path / HalfPath / HaplotypePlayback state, the MaxSumBoth tie-breaks,
 * Path.calculateWeights (Path.java:424-453), the status merge
 * (SequenceEvaluator.java:170-209), PhasingEvaluator.countMisphasings, and the
 * RocContainer accumulate / sort / cumulative-sum (RocContainer.java:225-318).
 *
 * The reference has no FFI; a maintainer adds the JNI stub shown in
 * INTEGRATION.md.  Everything here is plain C: pointers + sizes, caller owns
 * every buffer, no exceptions cross the boundary, return 0 == OK.
 *
 * Coordinates: 0-based, end-exclusive.  Bases: N=0 A=1 C=2 G=3 T=4
 * (reference: src/com/rtg/mode/DnaUtils.java:225-252); 5 = explicit unknown
 * allele token (Allele.java:50-52).
 */
#ifndef VCE_H_
#define VCE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status bits: identical to VariantId.java:40-50 ---- */
#define VCE_STATUS_SKIPPED       0x01
#define VCE_STATUS_GT_MATCH      0x02
#define VCE_STATUS_ALLELE_MATCH  0x04
#define VCE_STATUS_NO_MATCH      0x08
#define VCE_STATUS_OUTSIDE_EVAL  0x10
#define VCE_STATUS_ANY_MATCH     0x20

/* ---- orientor kinds: Orientor.java ---- */
enum {
  VCE_ORIENT_UNPHASED      = 0, /* Orientor.UnphasedOrientor      :60-99   */
  VCE_ORIENT_PHASED        = 1, /* Orientor.PhasedOrientor(false) :107-148 */
  VCE_ORIENT_PHASED_INVERT = 2, /* Orientor.PhasedOrientor(true)           */
  VCE_ORIENT_SQUASH_GT     = 3, /* Orientor.SQUASH_GT             :156-188 */
  VCE_ORIENT_ALLELE_GT     = 4, /* Orientor.ALLELE_GT             :196-233 */
  VCE_ORIENT_HAPLOID_POP   = 5, /* Orientor.HAPLOID_POP           :240-257 */
  VCE_ORIENT_DIPLOID_POP   = 6, /* Orientor.DIPLOID_POP           :264-290 */
  VCE_ORIENT_ALT           = 7  /* Orientor.AltOrientor           :297-338 */
};

/* ---- path preference: PathFinder.java:86-100 ---- */
enum {
  VCE_PREFER_SUM            = 0, /* MaxSumBoth.java:43-82           */
  VCE_PREFER_CALLS_MIN_BASE = 1  /* MaxCallsMinBaseline.java:42-61  */
};

/* ---- error codes ---- */
enum {
  VCE_OK                  = 0,
  VCE_ERR_NOT_INITIALISED = 1,
  VCE_ERR_BAD_ARGUMENT    = 2,
  VCE_ERR_CUDA            = 3,  /* CUDA runtime failure, see vce_last_error() */
  VCE_ERR_NO_DEVICE       = 4,  /* no CUDA device: there is NO CPU fallback   */
  VCE_ERR_CAPACITY        = 5,  /* a caller-provided output buffer was too small */
  VCE_ERR_BEST_PATH_NULL  = 6,  /* SequenceEvaluator.java:105-116 (NullPointerException) */
  VCE_ERR_MOVE_IN_VARIANT = 7,  /* HaplotypePlayback.java:260 IllegalStateException during a too-complex skip */
  VCE_ERR_UNSUPPORTED     = 8
};

#define VCE_MAX_PLOIDY 4

/*
 * One side (baseline or calls) of one reference sequence, exactly the data a
 * List<Variant> holds after loading (Variant.java / GtIdVariant.java /
 * Allele.java), in load order (ascending id).  Structure-of-arrays.
 *
 * Allele slots: variant i owns slots [allele_off[i], allele_off[i+1]); slot k
 * within the variant is allele id k-1, i.e. the array is indexed by GT+1 like
 * Variant.mAlleles (Variant.java:241): slot 0 = missing (-1), slot 1 = REF.
 * A slot with allele_null[s] != 0 is a Java null Allele ("play nothing").
 * Allele bases are nt[allele_nt_off[s] .. allele_nt_off[s+1]).
 * The variant locus (start,end) is derived by the engine as the bounds of the
 * non-null alleles (Allele.getAlleleBounds, Allele.java:195-206).
 */
typedef struct vce_variants {
  int32_t        n;             /* number of variants                              */
  int32_t        gt_ploidy;     /* entries per variant in gt[]; 0 = no GT (AllAlts factory).  Must be the same for every
                                   non-empty sequence of a batch on this side (else VCE_ERR_BAD_ARGUMENT) */
  const int32_t* id;            /* [n] VariantId.getId(): 1-based record ordinal, strictly ascending */
  const uint8_t* phased;        /* [n] Variant.isPhased()                          */
  const uint8_t* status_in;     /* [n] pre-set status bits (only OUTSIDE_EVAL is meaningful) or NULL */
  const int8_t*  gt;            /* [n*gt_ploidy] GtIdVariant.alleleIds(), -1 = missing */
  const int32_t* allele_off;    /* [n+1]                                           */
  int32_t        n_slots;       /* == allele_off[n]                                */
  const int32_t* allele_start;  /* [n_slots]                                       */
  const int32_t* allele_end;    /* [n_slots]                                       */
  const uint8_t* allele_null;   /* [n_slots]                                       */
  const int32_t* allele_nt_off; /* [n_slots+1]                                     */
  const uint8_t* nt;            /* [allele_nt_off[n_slots]] packed allele bases    */
} vce_variants;

typedef struct vce_sequence {
  const char*    name;          /* sequence name (only used in warning text)       */
  const uint8_t* template_bases;/* [template_len] 0..4, what SequencesReader.read() returns */
  int32_t        template_len;
  vce_variants   baseline;
  vce_variants   calls;
} vce_sequence;

/* PathFinder.Config (PathFinder.java:51-101) + the orientor pairs VcfEvalTask builds (:122-128,195-208). */
typedef struct vce_config {
  int32_t n_passes;             /* 1, or 2 for two-pass (diploid then allele matching on FN/FP) */
  int32_t baseline_orientor[2]; /* per pass                                        */
  int32_t calls_orientor[2];
  int32_t ploidy[2];            /* Orientor.haplotypes() per pass                  */
  int32_t max_paths;            /* ToolsGlobalFlags VCFEVAL_MAX_PATHS, default 50000      */
  int32_t max_iterations;       /* VCFEVAL_MAX_ITERATIONS, default 10000000               */
  int32_t prune_no_ops;         /* VCFEVAL_PRUNE_NO_OPS, default 1                        */
  int32_t flag_alternates;      /* VCFEVAL_FLAG_ALTERNATES, default 0                     */
  int32_t preference;           /* VCE_PREFER_*                                           */
} vce_config;

/* One "Evaluation too complex" event, PathFinder.java:162-169. */
typedef struct vce_warning {
  int32_t pass;                 /* 0 or 1                                          */
  int32_t paths;                /* sortedPaths.size()                              */
  int32_t iterations;           /* currentIterations                               */
  int32_t start1;               /* lastSyncPos + 1                                 */
  int32_t end1;                 /* mCurrentMaxPos + 2                              */
} vce_warning;

/* Per-side results, indexed like the input arrays (load order). Caller-allocated. */
typedef struct vce_side_result {
  uint8_t* status;              /* [n] VariantId status bits after SequenceEvaluator.mergeVariants */
  int8_t*  allele_ids;          /* [n*VCE_MAX_PLOIDY] OrientedVariant.alleleIds() of the chosen orientation
                                   (valid when GT_MATCH or ALLELE_MATCH), unused entries = -2 */
  uint8_t* original;            /* [n] OrientedVariant.isOriginal()                */
  double*  weight;              /* [n] OrientedVariant.getWeight() (calls side; 0 on baseline) */
  int32_t* sync_id;             /* [n] optional (may be NULL): the SYNC annotation value = id of the sync region holding the
                                   variant: the last sync point of sync_points[pass] at or before start - 1, plus 2
                                   (InterleavingEvalSynchronizer.floorSyncPos :71-83), 0 for skipped variants;
                                   pass 0 for GT_MATCH, pass 1 when a pass-2 list exists and the variant was not GT-matched */
} vce_side_result;

typedef struct vce_sequence_result {
  vce_side_result baseline;
  vce_side_result calls;
  int32_t*     sync_points[2];  /* Path.getSyncPoints() per pass; caller-allocated        */
  int32_t      sync_cap[2];     /* capacity; (n_baseline + n_calls + 2) always suffices   */
  int32_t      n_sync[2];       /* out                                                    */
  int32_t      phasing[3];      /* out: misPhasings, correctPhasings, unphaseable (PhasingEvaluator.java:55-142) */
  vce_warning* warnings;        /* caller-allocated                                       */
  int32_t      warn_cap;
  int32_t      n_warnings;      /* out (may exceed warn_cap: then only warn_cap are written) */
  int32_t      error;           /* out: VCE_OK or a per-sequence VCE_ERR_*                */
  int64_t      n_regions;       /* out: independent search regions the engine ran (diagnostic) */
  int64_t      n_iterations;    /* out: total best-first iterations (diagnostic)          */
} vce_sequence_result;

/* ---- ROC (RocContainer) ---- */
typedef struct vce_roc_in {
  int64_t        n;             /* number of addRocLine calls, in arrival order (RocContainer.java:225) */
  const double*  score;         /* [n] RocSortValueExtractor value; NaN/Inf -> "None" bucket (:232-237) */
  const double*  tp;            /* [n] tpWeight                                            */
  const double*  fp;            /* [n] fpWeight                                            */
  const double*  tp_raw;        /* [n] tpRaw                                               */
  const uint32_t* filter_mask;  /* [n] bit f set = RocFilter f accepts the record (or NULL = all) */
  int32_t        n_filters;     /* <= 32                                                   */
  int32_t        ascending;     /* RocSortOrder: 0 = DESCENDING (default), 1 = ASCENDING   */
} vce_roc_in;

/*
 * Output: for every filter, the TreeMap iteration of RocContainer.writeRocs
 * (:296-309): distinct score buckets in comparator order with the CUMULATIVE
 * tp / fp / tpRaw sums after adding that bucket (plain sequential double +=,
 * bucket sums first accumulated in arrival order :251-258).
 * Row formatting / the 3-dp merge (:298-312) stays with the caller.
 */
typedef struct vce_roc_out {
  int64_t  cap;                 /* rows available per filter in the arrays below     */
  int64_t* n_rows;              /* [n_filters] out                                   */
  double*  threshold;           /* [n_filters*cap] bucket score (NaN for the None bucket) */
  double*  cum_tp;              /* [n_filters*cap]                                   */
  double*  cum_fp;              /* [n_filters*cap]                                   */
  double*  cum_tp_raw;          /* [n_filters*cap]                                   */
  int64_t* no_score;            /* out: mNoScoreVariants (one per addRocLine with NaN/Inf score) */
} vce_roc_out;

/* ---- life cycle ---- */
int         vce_init(int device);               /* selects the CUDA device; fails with VCE_ERR_NO_DEVICE when none */
/* One process, several devices: the sequences of a vce_submit batch are spread over `devices` (longest-processing-time
   greedy by variant count; sequences are the reference's independent units, VcfEvalTask.java:131-135), one context, stream
   set and share of the host cores per device.  Registered templates live on every device.  vce_init(d) == vce_init_devices(1, &d). */
int         vce_init_devices(int device_count, const int* devices);
void        vce_shutdown(void);
const char* vce_last_error(void);               /* thread-local message for the last non-zero return */
const char* vce_version(void);

/*
 * Optional: make a reference sequence RESIDENT on the device (2 bits per base plus a map of the non-ACGT positions; a 3 Gbp
 * genome takes ~0.8 GB of the 180 GB).  It belongs to the load phase, where the reference reads the sequence from the SDF
 * (SequenceEvaluator.java:76-86, SequencesReader.read): call it once per sequence, before the evaluations that use it.
 * vce_submit() recognises a registered sequence by its template_bases pointer + template_len; for such a sequence the
 * caller's variant arrays are uploaded whole and every search window is cut out of HBM, instead of being gathered from
 * template_bases on the host.  Results are identical either way.  The caller must keep the bytes unchanged while registered.
 */
/*
 * Optional: page-lock a block of the caller's memory (cudaHostRegister).  Input arrays of vce_submit that lie inside a
 * registered block are copied to the device straight from where they are -- no staging copy on the host.  Meant for buffers
 * the host keeps and refills across calls (JNI: direct ByteBuffers).  Unregister before freeing the memory.
 */
int vce_host_register(const void* p, uint64_t bytes);
int vce_host_unregister(const void* p);

int vce_template_register(const uint8_t* bases, int32_t len);
int vce_template_unregister(const uint8_t* bases);

/* Default configuration = `rtg vcfeval` defaults (diploid, unphased orientors, split mode single pass). */
void        vce_default_config(vce_config* cfg);

/*
 * Synchronous whole-batch evaluation with HOST buffers (what the JNI shim
 * calls): replaces SequenceEvaluator.run() for every sequence in the batch.
 * Thread-safe; one CUDA stream per call.
 */
int vce_submit(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs,
               vce_sequence_result* results);

/*
 * Staged variant used to time the device path with inputs already resident in
 * HBM: create = host packing + H2D, run = all kernels (may be called
 * repeatedly), fetch = D2H + result assembly into caller buffers.
 */
typedef struct vce_job vce_job;
int  vce_job_create(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_job** job);
int  vce_job_run(vce_job* job);
int  vce_job_fetch(vce_job* job, vce_sequence_result* results);
void vce_job_destroy(vce_job* job);
/* Diagnostics of the last vce_job_run: kernel launches, device milliseconds (CUDA events on the job's stream). */
int  vce_job_stats(const vce_job* job, int64_t* n_launches, double* device_ms, double* search_kernel_ms,
                   int64_t* n_regions, int64_t* n_escalated);

/* Extended diagnostics of the last vce_job_run, out[0..n): launches, search-kernel ms, ms and region count of the largest
   shared-memory search launch, regions, escalated, spilled, prefix re-runs, H2D bytes, D2H bytes, best-first iterations,
   device ms of the whole run (CUDA events on the job's stream), variants searched by that largest launch. */
int  vce_job_stats_ex(const vce_job* job, double* out, int32_t n);

/* Diagnostics of the calling thread's last vce_submit, out[0..n): kernel launches, H2D bytes, D2H bytes, regions,
   regions settled by the warp-per-region kernels, best-first iterations; with several devices also [6..9): sequences cut
   into pieces to balance the devices, pieces made, sequences evaluated again in one piece because a cut could
   not be verified. */
int  vce_submit_stats(double* out, int32_t n);

/*
 * Summary of the calling thread's last vce_submit, per sequence: out[k * 8 + i] = baseline TP, baseline FN (incl. FN_CA),
 * call TP, call FP (incl. FP_CA), skipped variants (both sides), mis-phasings, correct phasings, unphaseable -- the classes
 * of the split writers (SplitEvalSynchronizer.java:112-156: a matched call of weight 0 and OUTSIDE_EVAL variants count
 * nowhere).  Counted on the device together with the per-variant outputs when the device-resident result path ran.
 */
int vce_submit_summary(int64_t* out, int32_t n_seq);

/* RocContainer accumulate + sort + cumulative scan on the device. */
int vce_roc(const vce_roc_in* in, vce_roc_out* out);
/* Same, also reporting host->device and kernel milliseconds (CUDA events) and the number of kernel launches. */
int vce_roc_timed(const vce_roc_in* in, vce_roc_out* out, double* h2d_ms, double* kernel_ms, int32_t* launches);

/*
 * SURVEY.md 8(f) rank 3 -- variant loading on the device: the text of ONE sequence's VCF records (what
 * VcfRecordTabixCallable.call reads for a sequence, VcfRecordTabixCallable.java:91-141; header lines and records of other
 * sequences are passed over) becomes the vce_variants arrays vce_submit takes.  Restated: VcfSortRefiner.java:91-108
 * (records sharing a start leave a java.util.PriorityQueue ordered by end), the loader filters (:103-118: longest allele
 * over --Xmax-length, record beyond the template, factory says no / skips), the sample factory
 * (VariantFactory.java:127-196: FILTER / FT unless --all-records, GT rules, haploid GT doubled, ploidy mismatch skipped),
 * VcfUtils.getAlleleStrings / hasRedundantFirstNucleotide (:785-828), VariantType.getType (:107-190) and Allele.getAlleles
 * (Allele.java:145-192) and, with ref_overlap, Variant.trimAlleles (Variant.java:93-190: REF nulled, ALTs stripped of what they
 * share with REF, the side decided by the overlapping neighbours; clusters of overlapping variants one thread each).
 * With eval_regions, the OUTSIDE_EVAL status bit of VcfRecordTabixCallable.java:125-127 (MergedIntervals.overlapped :172-179);
 * with sample -1 the all-alts factory of `--sample ALT` (VariantFactory.java:195-239); with a region, only the records
 * TabixLineReader hands on for it.  vce_vcf_roc_fields adds the ROC inputs of the same variants.
 * A record the reference would throw VcfFormatException for ends the call with VCE_ERR_BAD_ARGUMENT (vce_last_error names
 * the line).  The result owns host copies of the arrays until vce_vcf_free.
 */
typedef struct vce_vcf_in {
  const char* text;           /* uncompressed VCF text                            */
  int64_t     text_len;       /* < 2 GB                                            */
  const char* name;           /* CHROM to accept; NULL = every record              */
  int32_t     template_len;   /* records ending beyond it are skipped; -1 = unknown */
  int32_t     sample;         /* sample column, 0-based; -1 = `--sample ALT` (VariantFactory.AllAlts :195-239: every
                                 ALT that is a sequence, no genotypes: gt_ploidy 0) */
  int32_t     ploidy;         /* --sample-ploidy (2)                               */
  int32_t     pass_only;      /* 0 with --all-records                              */
  int32_t     max_length;     /* --Xmax-length (1000), -1 = no limit               */
  int32_t     ref_overlap;    /* --ref-overlap: Variant.trimAlleles (Variant.java:93-190) on the loaded variants */
  int32_t     n_eval_regions; /* --evaluation-regions of this sequence; ignored when eval_regions is NULL */
  int32_t     region_begin;   /* --region within the sequence: only records overlapping [region_begin, region_end) belong to the */
  int32_t     region_end;     /* stream (0-based, end exclusive; what TabixLineReader hands on); region_end <= region_begin (a
                                 zeroed struct) = the whole sequence */
  const int32_t* eval_regions;/* [n_eval_regions * 2] merged intervals [start, end), ascending (ReferenceRegions / MergedIntervals):
                                 variants whose locus overlaps none get VCE_STATUS_OUTSIDE_EVAL in status_in
                                 (VcfRecordTabixCallable.java:125-127); NULL = no regions given */
  const char* score_field;    /* --vcf-score-field for vce_vcf_roc_fields: NULL or "QUAL" = the QUAL column, else a FORMAT key of the
                                 sample (vcfeval's default is "GQ"; RocScoreField.java:50-80,120-140) */
} vce_vcf_in;
typedef struct vce_vcf_stats {
  int64_t lines, records, kept, skipped;   /* text lines, records of the sequence, variants, counted skips (:103-118) */
  double  h2d_ms, kernel_ms, d2h_ms;       /* CUDA events on the loader's stream   */
  int32_t launches;
  int32_t reserved;
} vce_vcf_stats;
int  vce_vcf_parse(const vce_vcf_in* in, void** result);
int  vce_vcf_variants(const void* result, vce_variants* out, vce_vcf_stats* stats);
void vce_vcf_free(void* result);
/*
 * The ROC inputs of the variants of a result, made by the same device passes ("RocFilter classification"):
 * what RocContainer.addRocLine derives from a call's record (RocContainer.java:225-248) -- qual[i] = the QUAL column as
 * Double.parseDouble reads it (RocScoreField.java:50-80; NaN when missing) and filter_mask[i] = the RocFilters that accept the
 * record, one bit per filter in declaration order (RocFilter.java:48-137): 0 ALL, 1 HOM, 2 HET, 3 SNP, 4 NON_SNP, 5 MNP, 6 INDEL,
 * 7 XRX, 8 NON_XRX (VariantType.getType(rec, gt) :130-160 on the padding-stripped alleles, VcfUtils.isHomozygousAlt /
 * isHeterozygous, INFO key XRX).  Bit 31 (VCE_ROC_SCORE_UNPARSED): the QUAL text is none of the plain decimal forms the device
 * parses exactly (more than 15 significant digits, hex floats, "NaN" ...): parse it yourself.  Select the bits of the filters
 * in use and hand (qual, tp, fp, tp_raw, mask) to vce_roc.  The arrays live as long as the result.
 */
#define VCE_ROC_SCORE_UNPARSED 0x80000000u
int vce_vcf_roc_fields(const void* result, const uint32_t** filter_mask, const double** qual);

/*
 * SURVEY.md 8(f) rank 4 -- the on-disk formats either side of the loader, decoded on the device.
 *
 * BGZF (`bgzip`ped VCF and its .tbi): the reference reads it through htsjdk's BlockCompressedInputStream (sam-rtg jar,
 * third party) from TabixLineReader / BlockCompressedLineReader (src/com/rtg/tabix/TabixLineReader.java,
 * BlockCompressedLineReader.java:58-133); restated from SAM spec 4.1, RFC 1952 and RFC 1951.  `data` holds the bytes of
 * the file; a virtual offset is (offset of a BGZF member << 16) | offset within that member's text
 * (src/com/rtg/tabix/VirtualOffsets.java); voffset_end < 0 = to the end of the file.  One BGZF member per warp.  A corrupt
 * member ends the call with VCE_ERR_BAD_ARGUMENT (vce_last_error names its offset and what zlib would have said);
 * check_crc != 0 also verifies every member's CRC32 on the device (htsjdk's default is not to).
 */
typedef struct vce_bgzf_stats {
  int64_t blocks, compressed_bytes, text_bytes;
  double  h2d_ms, kernel_ms, d2h_ms;          /* CUDA events on the loader's stream */
  int32_t launches;
  int32_t reserved;
} vce_bgzf_stats;
/* Host only: bytes of text between the two virtual offsets (sum of the members' ISIZE fields). */
int vce_bgzf_text_size(const uint8_t* data, int64_t len, int64_t voffset_begin, int64_t voffset_end, int64_t* text_len);
/* The text between the two virtual offsets into `text` (VCE_ERR_CAPACITY and *text_len = bytes needed when it is too small). */
int vce_bgzf_inflate(const uint8_t* data, int64_t len, int64_t voffset_begin, int64_t voffset_end, int32_t check_crc,
                     uint8_t* text, int64_t text_cap, int64_t* text_len, vce_bgzf_stats* stats);
/* vce_vcf_parse on block-compressed input: in->text / in->text_len are the bytes of the .vcf.gz; the members are inflated and
   the records parsed without the text leaving the device.  Results as for vce_vcf_parse; stats->kernel_ms covers both. */
int vce_vcf_parse_bgzf(const vce_vcf_in* in, int64_t voffset_begin, int64_t voffset_end, int32_t check_crc, void** result,
                       vce_bgzf_stats* stats);

/*
 * Tabix index (TabixIndexReader.java:63-146): the .tbi bytes are inflated on the device, the small tables stay with the host.
 * vce_tabix_info: out = { sequences, format, seq column, begin column, end column (0-based, -1 = none), meta, skip }.
 * vce_tabix_query = AbstractIndexReader.getFilePointers for one interval [begin0, end0) of `name` (0-based, end exclusive;
 * -1 = open), AbstractIndexReader.java:102-205: the virtual offsets that bracket the region's records, or -1 / -1 when the
 * index holds nothing for it (the reference returns null).
 */
int  vce_tabix_open(const uint8_t* tbi, int64_t len, void** index);
int  vce_tabix_info(const void* index, int32_t out[7]);
const char* vce_tabix_name(const void* index, int32_t i);
int  vce_tabix_query(const void* index, const char* name, int32_t begin0, int32_t end0, int64_t* voffset_begin, int64_t* voffset_end);
void vce_tabix_close(void* index);

/*
 * SDF sequence data straight into a resident template (replaces CompressedMemorySequencesReader's load of one sequence +
 * vce_template_register): `seqdata` = the bytes of the SDF's seqdataN file written by BitwiseByteArray.dumpCompressedValues
 * (src/com/rtg/util/bytecompression/BitwiseByteArray.java:426-439; read back by FileBitwiseInputStream.java:151-169): per
 * 64 residues three big-endian 64-bit bit planes.  first_residue = the sequence's start within the file (seqpointerN),
 * len = its length.  bases_out[len] receives the byte-per-base array (0 = N, 1..4 = A C G T) that vce_sequence.template_bases
 * points at; the 2-bit copy stays on every device of the pool, registered under that pointer.  ms3 (optional): H2D, kernel,
 * D2H milliseconds of the last device.
 */
int vce_template_register_sdf(const uint8_t* seqdata, int64_t seqdata_len, int64_t first_residue, int32_t len, uint8_t* bases_out,
                              double* ms3);

/* Formats the reference's warning text (PathFinder.java:163). Returns bytes written (excluding NUL). */
int vce_format_warning(const vce_warning* w, const char* seq_name, char* buf, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* VCE_H_ */

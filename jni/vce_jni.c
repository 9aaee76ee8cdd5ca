/*
 * vce_jni.c -- JNI binding of include/vce.h for com.rtg.vcf.eval.NativeVce (jni/NativeVce.java).
 *
 * Replaces nothing in the reference by itself: it is what a maintainer links into libvce_jni.so so that
 * SequenceEvaluator.run() (src/com/rtg/vcf/eval/SequenceEvaluator.java:87-155) and RocContainer.writeRocs
 * (src/com/rtg/vcf/eval/RocContainer.java:267-318) can call the native engine.  Build: jni/build.sh (needs a JDK for jni.h;
 * none exists in the build image, so this file is compiled in CI only where javac is found -- tests/c/test_submit_threads.c
 * drives vce_submit exactly as this stub does, from 8 threads, without a JVM).
 *
 * All arrays are pinned with GetPrimitiveArrayCritical for the duration of the call: the engine copies what it needs into
 * page-locked staging memory itself, so no Java-side copy is made.  No JNI call is made between Get...Critical and the
 * matching Release (the restriction of the Critical functions).
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/vce.h"

#define MAX_KEEP 40

typedef struct {
  jarray arr[MAX_KEEP];
  void* ptr[MAX_KEEP];
  jint mode[MAX_KEEP];
  int n;
} keep_t;

static jfieldID fid(JNIEnv* e, jobject o, const char* name, const char* sig) {
  return (*e)->GetFieldID(e, (*e)->GetObjectClass(e, o), name, sig);
}
static jarray field_array(JNIEnv* e, jobject o, const char* name, const char* sig) {
  return (jarray) (*e)->GetObjectField(e, o, fid(e, o, name, sig));
}
static void* pin(JNIEnv* e, keep_t* k, jarray a, jint releaseMode) {
  if (a == NULL || k->n >= MAX_KEEP) return NULL;
  void* p = (*e)->GetPrimitiveArrayCritical(e, a, NULL);
  k->arr[k->n] = a;
  k->ptr[k->n] = p;
  k->mode[k->n] = releaseMode;
  ++k->n;
  return p;
}
static void unpin_all(JNIEnv* e, keep_t* k) {
  for (int i = k->n - 1; i >= 0; --i) (*e)->ReleasePrimitiveArrayCritical(e, k->arr[i], k->ptr[i], k->mode[i]);
  k->n = 0;
}

/* jarray handles of one SideArrays / SideResults object, fetched BEFORE anything is pinned */
typedef struct {
  jint n, gtPloidy;
  jarray id, phased, statusIn, gt, alleleOff, alleleStart, alleleEnd, alleleNull, alleleNtOff, nt;
} side_handles;
typedef struct { jarray status, alleleIds, original, weight, syncId; } result_handles;

static void side_handles_of(JNIEnv* e, jobject s, side_handles* h) {
  h->n = (*e)->GetIntField(e, s, fid(e, s, "mN", "I"));
  h->gtPloidy = (*e)->GetIntField(e, s, fid(e, s, "mGtPloidy", "I"));
  h->id = field_array(e, s, "mId", "[I");
  h->phased = field_array(e, s, "mPhased", "[B");
  h->statusIn = field_array(e, s, "mStatusIn", "[B");
  h->gt = field_array(e, s, "mGt", "[B");
  h->alleleOff = field_array(e, s, "mAlleleOff", "[I");
  h->alleleStart = field_array(e, s, "mAlleleStart", "[I");
  h->alleleEnd = field_array(e, s, "mAlleleEnd", "[I");
  h->alleleNull = field_array(e, s, "mAlleleNull", "[B");
  h->alleleNtOff = field_array(e, s, "mAlleleNtOff", "[I");
  h->nt = field_array(e, s, "mNt", "[B");
}
static void result_handles_of(JNIEnv* e, jobject r, result_handles* h) {
  h->status = field_array(e, r, "mStatus", "[B");
  h->alleleIds = field_array(e, r, "mAlleleIds", "[B");
  h->original = field_array(e, r, "mOriginal", "[B");
  h->weight = field_array(e, r, "mWeight", "[D");
  h->syncId = field_array(e, r, "mSyncId", "[I");
}
static void side_pin(JNIEnv* e, keep_t* k, const side_handles* h, vce_variants* v) {
  memset(v, 0, sizeof(*v));
  v->n = h->n;
  v->gt_ploidy = h->gtPloidy;
  v->id = (const int32_t*) pin(e, k, h->id, JNI_ABORT);
  v->phased = (const uint8_t*) pin(e, k, h->phased, JNI_ABORT);
  v->status_in = (const uint8_t*) pin(e, k, h->statusIn, JNI_ABORT);
  v->gt = (const int8_t*) pin(e, k, h->gt, JNI_ABORT);
  v->allele_off = (const int32_t*) pin(e, k, h->alleleOff, JNI_ABORT);
  v->allele_start = (const int32_t*) pin(e, k, h->alleleStart, JNI_ABORT);
  v->allele_end = (const int32_t*) pin(e, k, h->alleleEnd, JNI_ABORT);
  v->allele_null = (const uint8_t*) pin(e, k, h->alleleNull, JNI_ABORT);
  v->allele_nt_off = (const int32_t*) pin(e, k, h->alleleNtOff, JNI_ABORT);
  v->nt = (const uint8_t*) pin(e, k, h->nt, JNI_ABORT);
  v->n_slots = v->allele_off ? v->allele_off[v->n] : 0;
}
static void result_pin(JNIEnv* e, keep_t* k, const result_handles* h, vce_side_result* r) {
  r->status = (uint8_t*) pin(e, k, h->status, 0);
  r->allele_ids = (int8_t*) pin(e, k, h->alleleIds, 0);
  r->original = (uint8_t*) pin(e, k, h->original, 0);
  r->weight = (double*) pin(e, k, h->weight, 0);
  r->sync_id = (int32_t*) pin(e, k, h->syncId, 0);
}

JNIEXPORT jint JNICALL Java_com_rtg_vcf_eval_NativeVce_init(JNIEnv* e, jclass c, jintArray jdevices) {
  (void) c;
  const jsize n = (*e)->GetArrayLength(e, jdevices);
  jint* d = (*e)->GetIntArrayElements(e, jdevices, NULL);
  const int rc = vce_init_devices((int) n, (const int*) d);
  (*e)->ReleaseIntArrayElements(e, jdevices, d, JNI_ABORT);
  return rc;
}

JNIEXPORT void JNICALL Java_com_rtg_vcf_eval_NativeVce_shutdown(JNIEnv* e, jclass c) {
  (void) e; (void) c;
  vce_shutdown();
}

JNIEXPORT jstring JNICALL Java_com_rtg_vcf_eval_NativeVce_lastError(JNIEnv* e, jclass c) {
  (void) c;
  return (*e)->NewStringUTF(e, vce_last_error());
}

/* The registered bytes must not move: the template is copied once into memory this stub owns (the reference holds the
   sequence as a byte[] the collector may relocate).  Returned handle = registration key for submit(). */
typedef struct tmpl_copy { struct tmpl_copy* next; jsize len; uint8_t* bytes; } tmpl_copy;
static tmpl_copy* g_templates = NULL;

JNIEXPORT jint JNICALL Java_com_rtg_vcf_eval_NativeVce_registerTemplate(JNIEnv* e, jclass c, jbyteArray jt) {
  (void) c;
  const jsize len = (*e)->GetArrayLength(e, jt);
  tmpl_copy* t = (tmpl_copy*) malloc(sizeof(tmpl_copy));
  if (!t) return VCE_ERR_BAD_ARGUMENT;
  t->bytes = (uint8_t*) malloc((size_t) len + 1);
  if (!t->bytes) { free(t); return VCE_ERR_BAD_ARGUMENT; }
  (*e)->GetByteArrayRegion(e, jt, 0, len, (jbyte*) t->bytes);
  t->len = len;
  const int rc = vce_template_register(t->bytes, (int32_t) len);
  if (rc != 0) { free(t->bytes); free(t); return rc; }
  t->next = g_templates;      /* (registration happens in the single-threaded load phase) */
  g_templates = t;
  return 0;
}

/* A registered copy with the same length and content as the pinned template, or the pinned bytes themselves. */
static const uint8_t* resident_or(const uint8_t* pinned, jsize len) {
  for (const tmpl_copy* t = g_templates; t; t = t->next) {
    if (t->len == len && (len == 0 || (t->bytes[0] == pinned[0] && t->bytes[len - 1] == pinned[len - 1] &&
                                       memcmp(t->bytes, pinned, (size_t) (len < 4096 ? len : 4096)) == 0))) {
      return t->bytes;
    }
  }
  return pinned;
}

JNIEXPORT jint JNICALL Java_com_rtg_vcf_eval_NativeVce_submit(JNIEnv* e, jclass c, jintArray jcfg, jstring jname, jbyteArray jtemplate,
    jobject jb, jobject jc, jobject jbo, jobject jco, jobjectArray jsync, jintArray jnsync, jintArray jphasing, jintArray jwarn,
    jintArray jnwarn) {
  (void) c;
  vce_config cfg;
  jint ci[12];
  (*e)->GetIntArrayRegion(e, jcfg, 0, 12, ci);
  cfg.n_passes = ci[0];
  cfg.baseline_orientor[0] = ci[1]; cfg.baseline_orientor[1] = ci[2];
  cfg.calls_orientor[0] = ci[3]; cfg.calls_orientor[1] = ci[4];
  cfg.ploidy[0] = ci[5]; cfg.ploidy[1] = ci[6];
  cfg.max_paths = ci[7]; cfg.max_iterations = ci[8]; cfg.prune_no_ops = ci[9]; cfg.flag_alternates = ci[10]; cfg.preference = ci[11];

  /* every handle first: no JNI call is allowed while arrays are pinned */
  side_handles hb, hc;
  result_handles rb, rc_;
  side_handles_of(e, jb, &hb);
  side_handles_of(e, jc, &hc);
  result_handles_of(e, jbo, &rb);
  result_handles_of(e, jco, &rc_);
  jintArray sync0 = (jintArray) (*e)->GetObjectArrayElement(e, jsync, 0);
  jintArray sync1 = (jintArray) (*e)->GetObjectArrayElement(e, jsync, 1);
  const jsize cap0 = (*e)->GetArrayLength(e, sync0), cap1 = (*e)->GetArrayLength(e, sync1);
  const jsize warnCap = (*e)->GetArrayLength(e, jwarn) / 5;
  const jsize tlen = (*e)->GetArrayLength(e, jtemplate);
  const char* name = (*e)->GetStringUTFChars(e, jname, NULL);
  vce_warning* warnings = (vce_warning*) calloc((size_t) (warnCap > 0 ? warnCap : 1), sizeof(vce_warning));

  keep_t k;
  k.n = 0;
  vce_sequence seq;
  vce_sequence_result res;
  memset(&seq, 0, sizeof(seq));
  memset(&res, 0, sizeof(res));
  seq.name = name;
  seq.template_len = (int32_t) tlen;
  const uint8_t* pinnedTemplate = (const uint8_t*) pin(e, &k, jtemplate, JNI_ABORT);
  seq.template_bases = resident_or(pinnedTemplate, tlen);
  side_pin(e, &k, &hb, &seq.baseline);
  side_pin(e, &k, &hc, &seq.calls);
  result_pin(e, &k, &rb, &res.baseline);
  result_pin(e, &k, &rc_, &res.calls);
  res.sync_points[0] = (int32_t*) pin(e, &k, sync0, 0);
  res.sync_points[1] = (int32_t*) pin(e, &k, sync1, 0);
  res.sync_cap[0] = (int32_t) cap0;
  res.sync_cap[1] = (int32_t) cap1;
  res.warnings = warnings;
  res.warn_cap = (int32_t) warnCap;

  const int rc = vce_submit(&cfg, 1, &seq, &res);

  unpin_all(e, &k);
  (*e)->ReleaseStringUTFChars(e, jname, name);
  jint tmp[3];
  tmp[0] = res.n_sync[0]; tmp[1] = res.n_sync[1];
  (*e)->SetIntArrayRegion(e, jnsync, 0, 2, tmp);
  tmp[0] = res.phasing[0]; tmp[1] = res.phasing[1]; tmp[2] = res.phasing[2];
  (*e)->SetIntArrayRegion(e, jphasing, 0, 3, tmp);
  const jint nw = res.n_warnings < warnCap ? res.n_warnings : warnCap;
  for (jint i = 0; i < nw; ++i) {
    const jint w[5] = {warnings[i].pass, warnings[i].paths, warnings[i].iterations, warnings[i].start1, warnings[i].end1};
    (*e)->SetIntArrayRegion(e, jwarn, i * 5, 5, w);
  }
  tmp[0] = res.n_warnings;
  (*e)->SetIntArrayRegion(e, jnwarn, 0, 1, tmp);
  free(warnings);
  return rc != 0 ? rc : res.error;
}

JNIEXPORT jint JNICALL Java_com_rtg_vcf_eval_NativeVce_roc(JNIEnv* e, jclass c, jdoubleArray jscore, jdoubleArray jtp, jdoubleArray jfp,
    jdoubleArray jraw, jintArray jmask, jint nFilters, jboolean ascending, jlongArray jnRows, jdoubleArray jthr, jdoubleArray jctp,
    jdoubleArray jcfp, jdoubleArray jcraw, jlongArray jnoScore) {
  (void) c;
  vce_roc_in in;
  vce_roc_out out;
  memset(&in, 0, sizeof(in));
  memset(&out, 0, sizeof(out));
  in.n = (int64_t) (*e)->GetArrayLength(e, jscore);
  in.n_filters = nFilters;
  in.ascending = ascending ? 1 : 0;
  out.cap = nFilters > 0 ? (int64_t) (*e)->GetArrayLength(e, jthr) / nFilters : 0;
  keep_t k;
  k.n = 0;
  in.score = (const double*) pin(e, &k, jscore, JNI_ABORT);
  in.tp = (const double*) pin(e, &k, jtp, JNI_ABORT);
  in.fp = (const double*) pin(e, &k, jfp, JNI_ABORT);
  in.tp_raw = (const double*) pin(e, &k, jraw, JNI_ABORT);
  in.filter_mask = (const uint32_t*) pin(e, &k, jmask, JNI_ABORT);
  out.n_rows = (int64_t*) pin(e, &k, jnRows, 0);
  out.threshold = (double*) pin(e, &k, jthr, 0);
  out.cum_tp = (double*) pin(e, &k, jctp, 0);
  out.cum_fp = (double*) pin(e, &k, jcfp, 0);
  out.cum_tp_raw = (double*) pin(e, &k, jcraw, 0);
  out.no_score = (int64_t*) pin(e, &k, jnoScore, 0);
  const int rc = vce_roc(&in, &out);
  unpin_all(e, &k);
  return rc;
}

/* ---- loaders: tabix index, .vcf.gz -> SideArrays, SDF seqdata -> resident template ---- */
JNIEXPORT jlong JNICALL Java_com_rtg_vcf_eval_NativeVce_tabixOpen(JNIEnv* e, jclass c, jbyteArray jtbi) {
  (void) c;
  const jsize len = (*e)->GetArrayLength(e, jtbi);
  void* p = (*e)->GetPrimitiveArrayCritical(e, jtbi, NULL);
  void* index = NULL;
  const int rc = vce_tabix_open((const uint8_t*) p, (int64_t) len, &index);
  (*e)->ReleasePrimitiveArrayCritical(e, jtbi, p, JNI_ABORT);
  return rc == 0 ? (jlong) (intptr_t) index : 0;
}

JNIEXPORT jlongArray JNICALL Java_com_rtg_vcf_eval_NativeVce_tabixQuery(JNIEnv* e, jclass c, jlong index, jstring jname, jint begin0, jint end0) {
  (void) c;
  const char* name = (*e)->GetStringUTFChars(e, jname, NULL);
  int64_t v[2] = {-1, -1};
  const int rc = vce_tabix_query((const void*) (intptr_t) index, name, begin0, end0, &v[0], &v[1]);
  (*e)->ReleaseStringUTFChars(e, jname, name);
  if (rc != 0 || v[0] < 0) return NULL;
  jlongArray out = (*e)->NewLongArray(e, 2);
  const jlong jv[2] = {(jlong) v[0], (jlong) v[1]};
  (*e)->SetLongArrayRegion(e, out, 0, 2, jv);
  return out;
}

JNIEXPORT void JNICALL Java_com_rtg_vcf_eval_NativeVce_tabixClose(JNIEnv* e, jclass c, jlong index) {
  (void) e; (void) c;
  vce_tabix_close((void*) (intptr_t) index);
}

static void set_int_field_array(JNIEnv* e, jobject o, const char* name, const int32_t* src, jsize n) {
  jintArray a = (*e)->NewIntArray(e, n);
  if (n > 0) (*e)->SetIntArrayRegion(e, a, 0, n, (const jint*) src);
  (*e)->SetObjectField(e, o, fid(e, o, name, "[I"), a);
}
static void set_byte_field_array(JNIEnv* e, jobject o, const char* name, const void* src, jsize n) {
  jbyteArray a = (*e)->NewByteArray(e, n);
  if (n > 0) (*e)->SetByteArrayRegion(e, a, 0, n, (const jbyte*) src);
  (*e)->SetObjectField(e, o, fid(e, o, name, "[B"), a);
}

JNIEXPORT jint JNICALL Java_com_rtg_vcf_eval_NativeVce_loadVcfGz(JNIEnv* e, jclass c, jbyteArray jfile, jlong vbeg, jlong vend, jstring jname,
    jint templateLength, jint sample, jint ploidy, jboolean passOnly, jint maxLength, jboolean refOverlap, jintArray jregions, jstring jscore,
    jobject jout, jlongArray jcounters) {
  (void) c;
  const char* name = (*e)->GetStringUTFChars(e, jname, NULL);
  vce_vcf_in in;
  memset(&in, 0, sizeof(in));
  in.text_len = (int64_t) (*e)->GetArrayLength(e, jfile);
  in.name = name;
  in.template_len = templateLength;
  in.sample = sample;
  in.ploidy = ploidy;
  in.pass_only = passOnly ? 1 : 0;
  in.max_length = maxLength;
  in.ref_overlap = refOverlap ? 1 : 0;
  in.region_begin = 0;
  in.region_end = -1;   /* the whole sequence (a finer --region: set both from the RegionRestriction) */
  void* result = NULL;
  /* the file bytes are pinned for the call only: the engine uploads them and keeps nothing */
  const char* score = jscore ? (*e)->GetStringUTFChars(e, jscore, NULL) : NULL;
  in.score_field = score;
  const jsize nreg = jregions ? (*e)->GetArrayLength(e, jregions) / 2 : 0;
  void* p = (*e)->GetPrimitiveArrayCritical(e, jfile, NULL);
  void* reg = jregions ? (*e)->GetPrimitiveArrayCritical(e, jregions, NULL) : NULL;
  in.text = (const char*) p;
  in.eval_regions = (const int32_t*) reg;
  in.n_eval_regions = (int32_t) nreg;
  int rc = vce_vcf_parse_bgzf(&in, (int64_t) vbeg, (int64_t) vend, 0, &result, NULL);
  if (reg) (*e)->ReleasePrimitiveArrayCritical(e, jregions, reg, JNI_ABORT);
  (*e)->ReleasePrimitiveArrayCritical(e, jfile, p, JNI_ABORT);
  (*e)->ReleaseStringUTFChars(e, jname, name);
  if (score) (*e)->ReleaseStringUTFChars(e, jscore, score);
  if (rc != 0) return rc;
  vce_variants v;
  vce_vcf_stats st;
  rc = vce_vcf_variants(result, &v, &st);
  if (rc == 0) {
    const jsize n = (jsize) v.n, slots = (jsize) v.n_slots;
    (*e)->SetIntField(e, jout, fid(e, jout, "mN", "I"), n);
    (*e)->SetIntField(e, jout, fid(e, jout, "mGtPloidy", "I"), v.gt_ploidy);
    set_int_field_array(e, jout, "mId", v.id, n);
    set_byte_field_array(e, jout, "mPhased", v.phased, n);
    set_byte_field_array(e, jout, "mStatusIn", v.status_in, n);
    set_byte_field_array(e, jout, "mGt", v.gt, n * v.gt_ploidy);
    set_int_field_array(e, jout, "mAlleleOff", v.allele_off, n + 1);
    set_int_field_array(e, jout, "mAlleleStart", v.allele_start, slots);
    set_int_field_array(e, jout, "mAlleleEnd", v.allele_end, slots);
    set_byte_field_array(e, jout, "mAlleleNull", v.allele_null, slots);
    set_int_field_array(e, jout, "mAlleleNtOff", v.allele_nt_off, slots + 1);
    set_byte_field_array(e, jout, "mNt", v.nt, slots > 0 ? (jsize) v.allele_nt_off[slots] : 0);
    const uint32_t* mask = NULL;
    const double* qual = NULL;
    if (vce_vcf_roc_fields(result, &mask, &qual) == 0) {
      set_int_field_array(e, jout, "mRocMask", (const int32_t*) mask, n);
      jdoubleArray sc = (*e)->NewDoubleArray(e, n);
      if (n > 0) (*e)->SetDoubleArrayRegion(e, sc, 0, n, (const jdouble*) qual);
      (*e)->SetObjectField(e, jout, fid(e, jout, "mScore", "[D"), sc);
    }
    const jlong cnt[4] = {(jlong) st.lines, (jlong) st.records, (jlong) st.kept, (jlong) st.skipped};
    (*e)->SetLongArrayRegion(e, jcounters, 0, 4, cnt);
  }
  vce_vcf_free(result);
  return rc;
}

JNIEXPORT jbyteArray JNICALL Java_com_rtg_vcf_eval_NativeVce_registerTemplateSdf(JNIEnv* e, jclass c, jbyteArray jseqdata, jlong first, jint len) {
  (void) c;
  if (len <= 0) return NULL;
  /* the registered bytes must not move: this stub owns them, as for registerTemplate */
  tmpl_copy* t = (tmpl_copy*) malloc(sizeof(tmpl_copy));
  if (!t) return NULL;
  t->bytes = (uint8_t*) malloc((size_t) len + 1);
  if (!t->bytes) { free(t); return NULL; }
  t->len = len;
  const jsize flen = (*e)->GetArrayLength(e, jseqdata);
  void* p = (*e)->GetPrimitiveArrayCritical(e, jseqdata, NULL);
  const int rc = vce_template_register_sdf((const uint8_t*) p, (int64_t) flen, (int64_t) first, (int32_t) len, t->bytes, NULL);
  (*e)->ReleasePrimitiveArrayCritical(e, jseqdata, p, JNI_ABORT);
  if (rc != 0) { free(t->bytes); free(t); return NULL; }
  t->next = g_templates;
  g_templates = t;
  jbyteArray out = (*e)->NewByteArray(e, len);
  (*e)->SetByteArrayRegion(e, out, 0, len, (const jbyte*) t->bytes);
  return out;
}

JNIEXPORT jstring JNICALL Java_com_rtg_vcf_eval_NativeVce_formatWarning(JNIEnv* e, jclass c, jint pass, jint paths, jint iterations,
    jint start1, jint end1, jstring jname) {
  (void) c;
  const char* name = (*e)->GetStringUTFChars(e, jname, NULL);
  const vce_warning w = {pass, paths, iterations, start1, end1};
  char buf[512];
  vce_format_warning(&w, name, buf, (int32_t) sizeof(buf));
  (*e)->ReleaseStringUTFChars(e, jname, name);
  return (*e)->NewStringUTF(e, buf);
}

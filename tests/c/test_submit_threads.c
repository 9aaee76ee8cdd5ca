/*
 * test_submit_threads.c -- drives vce_submit exactly as the JNI stub does (jni/vce_jni.c): plain C against include/vce.h,
 * one sequence per call, from 8 threads at once (the reference's SimpleThreadPool workers, VcfEvalTask.java:131-137).
 * Every thread builds its own small sequence (SNPs every 97 bases; calls = baseline with every 5th record dropped and every
 * 7th one replaced by a different ALT) and checks the per-variant outcome it must get.  Prints the mean latency of a call.
 * Built and run by tests/test_gpu_parity.py::test_c_abi_from_plain_c_threads (needs a GPU).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/vce.h"

#define NTHREADS 8
#define CALLS_PER_THREAD 6

typedef struct {
  int32_t n;
  int32_t* id; uint8_t* phased; uint8_t* status_in; int8_t* gt; int32_t* allele_off; int32_t* a_start; int32_t* a_end;
  uint8_t* a_null; int32_t* nt_off; uint8_t* nt;
} side_t;

static void side_alloc(side_t* s, int n) {
  s->n = n;
  s->id = calloc((size_t) n + 1, 4); s->phased = calloc((size_t) n + 1, 1); s->status_in = calloc((size_t) n + 1, 1);
  s->gt = calloc((size_t) 2 * n + 2, 1); s->allele_off = calloc((size_t) n + 2, 4);
  s->a_start = calloc((size_t) 3 * n + 3, 4); s->a_end = calloc((size_t) 3 * n + 3, 4); s->a_null = calloc((size_t) 3 * n + 3, 1);
  s->nt_off = calloc((size_t) 3 * n + 4, 4); s->nt = calloc((size_t) 2 * n + 4, 1);
}
static void side_put(side_t* s, int i, int id, int pos, uint8_t ref, uint8_t alt, int hom) {
  s->id[i] = id;
  s->gt[2 * i] = hom ? 1 : 0; s->gt[2 * i + 1] = 1;
  s->allele_off[i] = 3 * i; s->allele_off[i + 1] = 3 * i + 3;
  s->a_null[3 * i] = 1;                                   /* slot 0: the missing allele */
  for (int k = 1; k <= 2; ++k) { s->a_start[3 * i + k] = pos; s->a_end[3 * i + k] = pos + 1; }
  s->nt_off[3 * i] = 2 * i; s->nt_off[3 * i + 1] = 2 * i; s->nt_off[3 * i + 2] = 2 * i + 1; s->nt_off[3 * i + 3] = 2 * i + 2;
  s->nt[2 * i] = ref; s->nt[2 * i + 1] = alt;
}
static void side_view(const side_t* s, vce_variants* v) {
  memset(v, 0, sizeof(*v));
  v->n = s->n; v->gt_ploidy = 2; v->id = s->id; v->phased = s->phased; v->status_in = s->status_in; v->gt = s->gt;
  v->allele_off = s->allele_off; v->n_slots = 3 * s->n; v->allele_start = s->a_start; v->allele_end = s->a_end;
  v->allele_null = s->a_null; v->allele_nt_off = s->nt_off; v->nt = s->nt;
}

typedef struct { int tid; int failures; double ms; } job_t;

static void* worker(void* arg) {
  job_t* job = (job_t*) arg;
  const int nb = 400 + 37 * job->tid;
  const int tl = nb * 97 + 500;
  uint8_t* tmpl = malloc((size_t) tl);
  uint32_t x = 12345u + (uint32_t) job->tid * 7919u;
  for (int i = 0; i < tl; ++i) { x = x * 1664525u + 1013904223u; tmpl[i] = (uint8_t) (1 + ((x >> 24) & 3)); }
  side_t b, c;
  side_alloc(&b, nb);
  side_alloc(&c, nb);
  int nc = 0;
  uint8_t* expectB = calloc((size_t) nb, 1);
  uint8_t* expectC = calloc((size_t) nb, 1);
  for (int i = 0; i < nb; ++i) {
    const int pos = 100 + 97 * i;
    const uint8_t ref = tmpl[pos], alt = (uint8_t) (ref % 4 + 1);
    side_put(&b, i, i + 1, pos, ref, alt, i % 3 == 0);
    if (i % 5 == 4) { expectB[i] = VCE_STATUS_NO_MATCH; continue; }          /* dropped from the calls: FN */
    if (i % 7 == 6) {                                                        /* a different ALT: FN + FP */
      side_put(&c, nc, nc + 1, pos, ref, (uint8_t) ((alt % 4) + 1 == ref ? (ref + 1) % 4 + 1 : (alt % 4) + 1), i % 3 == 0);
      expectB[i] = VCE_STATUS_NO_MATCH;
      expectC[nc++] = VCE_STATUS_NO_MATCH;
      continue;
    }
    side_put(&c, nc, nc + 1, pos, ref, alt, i % 3 == 0);
    expectB[i] = VCE_STATUS_GT_MATCH;
    expectC[nc++] = VCE_STATUS_GT_MATCH;
  }
  c.n = nc;
  vce_config cfg;
  vce_default_config(&cfg);
  vce_sequence seq;
  memset(&seq, 0, sizeof(seq));
  char name[32];
  snprintf(name, sizeof(name), "thread%d", job->tid);
  seq.name = name; seq.template_bases = tmpl; seq.template_len = tl;
  side_view(&b, &seq.baseline);
  side_view(&c, &seq.calls);
  vce_sequence_result res;
  memset(&res, 0, sizeof(res));
  res.baseline.status = calloc((size_t) nb, 1); res.baseline.allele_ids = calloc((size_t) nb * 4, 1); res.baseline.original = calloc((size_t) nb, 1);
  res.baseline.weight = calloc((size_t) nb, 8); res.baseline.sync_id = calloc((size_t) nb, 4);
  res.calls.status = calloc((size_t) nb, 1); res.calls.allele_ids = calloc((size_t) nb * 4, 1); res.calls.original = calloc((size_t) nb, 1);
  res.calls.weight = calloc((size_t) nb, 8); res.calls.sync_id = calloc((size_t) nb, 4);
  for (int p = 0; p < 2; ++p) { res.sync_points[p] = calloc((size_t) 2 * nb + 4, 4); res.sync_cap[p] = 2 * nb + 2; }
  vce_warning warn[8];
  res.warnings = warn; res.warn_cap = 8;
  struct timespec t0, t1;
  for (int rep = 0; rep < CALLS_PER_THREAD; ++rep) {
    if (rep == 1) clock_gettime(CLOCK_MONOTONIC, &t0);   /* the first call builds the thread's context */
    const int rc = vce_submit(&cfg, 1, &seq, &res);
    if (rc != VCE_OK) { fprintf(stderr, "thread %d: vce_submit -> %d (%s)\n", job->tid, rc, vce_last_error()); ++job->failures; break; }
    for (int i = 0; i < nb; ++i) if ((res.baseline.status[i] & 0x0f) != expectB[i]) ++job->failures;
    for (int i = 0; i < nc; ++i) {
      if ((res.calls.status[i] & 0x0f) != expectC[i]) ++job->failures;
      if (expectC[i] == VCE_STATUS_GT_MATCH && res.calls.weight[i] != 1.0) ++job->failures;
    }
    if (res.n_sync[0] < nb / 2 || res.n_warnings != 0) ++job->failures;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  job->ms = ((double) (t1.tv_sec - t0.tv_sec) * 1e3 + (double) (t1.tv_nsec - t0.tv_nsec) / 1e6) / (CALLS_PER_THREAD - 1);
  return NULL;
}

int main(void) {
  if (vce_init(0) != VCE_OK) { fprintf(stderr, "vce_init: %s\n", vce_last_error()); return 2; }
  pthread_t th[NTHREADS];
  job_t jobs[NTHREADS];
  for (int t = 0; t < NTHREADS; ++t) { jobs[t].tid = t; jobs[t].failures = 0; jobs[t].ms = 0; pthread_create(&th[t], NULL, worker, &jobs[t]); }
  int failures = 0;
  double ms = 0;
  for (int t = 0; t < NTHREADS; ++t) { pthread_join(th[t], NULL); failures += jobs[t].failures; ms += jobs[t].ms; }
  vce_shutdown();
  printf("%s: %d threads x %d calls, %d mismatches, mean %.3f ms per single-sequence vce_submit\n", failures ? "FAILED" : "OK", NTHREADS,
         CALLS_PER_THREAD, failures, ms / NTHREADS);
  return failures ? 1 : 0;
}

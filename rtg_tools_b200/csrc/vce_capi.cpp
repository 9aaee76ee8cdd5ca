// vce_capi.cpp -- the C-ABI of include/vce.h on top of the pipeline + CUDA backend.
//
// There is NO CPU fallback: without a CUDA device every compute entry point fails with VCE_ERR_NO_DEVICE.
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>

#include "../../include/vce.h"
#include "vce_cuda.h"
#include "vce_pipeline.h"

namespace {
std::mutex g_mu;
int g_device = -1;
thread_local std::string t_err;

int fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}
}  // namespace

struct vce_job {
  std::unique_ptr<vce::Backend> backend;
  std::unique_ptr<vce::Pipeline> pipeline;
  vce_config cfg;
  int32_t nSeq = 0;
  const vce_sequence* seqs = nullptr;
  bool prepared = false;
  double executeMs = 0;
};

extern "C" {

const char* vce_version(void) { return "rtg-tools_b200 vce 0.1.0 (sm_100a)"; }

const char* vce_last_error(void) { return t_err.c_str(); }

void vce_default_config(vce_config* cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->n_passes = 1;
  cfg->baseline_orientor[0] = cfg->calls_orientor[0] = VCE_ORIENT_UNPHASED;
  cfg->baseline_orientor[1] = cfg->calls_orientor[1] = VCE_ORIENT_SQUASH_GT;
  cfg->ploidy[0] = 2;
  cfg->ploidy[1] = 1;
  cfg->max_paths = 50000;        // ToolsGlobalFlags.java:139
  cfg->max_iterations = 10000000;  // ToolsGlobalFlags.java:140
  cfg->prune_no_ops = 1;
  cfg->flag_alternates = 0;
  cfg->preference = VCE_PREFER_SUM;
}

int vce_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int n = vce::cudaDeviceCount();
  if (n <= 0) return fail(VCE_ERR_NO_DEVICE, "no CUDA device is visible; this engine has no CPU fallback");
  if (device < 0 || device >= n) return fail(VCE_ERR_BAD_ARGUMENT, "device index out of range");
  g_device = device;
  return VCE_OK;
}

void vce_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_device = -1;
}

int vce_job_create(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_job** job) {
  if (!cfg || !job || n_seq < 0 || (n_seq > 0 && !seqs)) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  int dev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = g_device;
  }
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  std::string err;
  std::unique_ptr<vce::Backend> be(vce::createCudaBackend(dev, err));
  if (!be) return fail(VCE_ERR_CUDA, err);
  std::unique_ptr<vce_job> j(new vce_job());
  j->backend = std::move(be);
  j->pipeline.reset(new vce::Pipeline(j->backend.get(), vce::PipelineOptions()));
  j->cfg = *cfg;
  j->nSeq = n_seq;
  j->seqs = seqs;
  const int rc = j->pipeline->prepare(*cfg, n_seq, seqs);
  if (rc) return fail(rc, j->pipeline->error());
  j->prepared = true;
  *job = j.release();
  return VCE_OK;
}

int vce_job_run(vce_job* job) {
  if (!job || !job->prepared) return fail(VCE_ERR_BAD_ARGUMENT, "job not prepared");
  // inputs are resident on the device since vce_job_create; only the device pipeline runs here
  const int64_t h2d = job->backend->stats.h2dBytes;
  job->backend->stats = vce::SearchStats();
  job->backend->stats.h2dBytes = h2d;
  job->backend->timerStart();
  const int rc = job->pipeline->execute();
  job->executeMs = job->backend->timerStopMs();
  if (rc) return fail(rc, job->pipeline->error());
  return VCE_OK;
}

int vce_job_fetch(vce_job* job, vce_sequence_result* results) {
  if (!job || !results) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const int rc = job->pipeline->finish(results);
  if (rc) return fail(rc, "one or more sequences reported an error (see vce_sequence_result.error)");
  return VCE_OK;
}

void vce_job_destroy(vce_job* job) { delete job; }

int vce_job_stats(const vce_job* job, int64_t* n_launches, double* device_ms, double* search_kernel_ms, int64_t* n_regions,
                  int64_t* n_escalated) {
  if (!job) return fail(VCE_ERR_BAD_ARGUMENT, "null job");
  const vce::SearchStats& s = job->backend->stats;
  if (n_launches) *n_launches = s.launches;
  if (device_ms) *device_ms = s.searchMs;
  if (n_regions) *n_regions = s.regions;
  if (n_escalated) *n_escalated = s.escalated;
  if (search_kernel_ms) {
    double ms = 0;
    long long regs = 0;
    vce::cudaBackendFastStats(job->backend.get(), &ms, &regs, nullptr);
    *search_kernel_ms = ms;
  }
  return VCE_OK;
}

int vce_job_stats_ex(const vce_job* job, double* out, int32_t n) {
  if (!job || !out) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const vce::SearchStats& s = job->backend->stats;
  double fastMs = 0;
  long long fastRegions = 0;
  long long fastVariants = 0;
  vce::cudaBackendFastStats(job->backend.get(), &fastMs, &fastRegions, &fastVariants);
  const double v[13] = {(double) s.launches, s.searchMs, fastMs, (double) fastRegions, (double) s.regions, (double) s.escalated,
                        (double) s.spilled, (double) s.prefixReruns, (double) s.h2dBytes, (double) s.d2hBytes,
                        (double) s.iterations, job->executeMs, (double) fastVariants};
  for (int i = 0; i < n && i < 13; ++i) out[i] = v[i];
  return VCE_OK;
}

int vce_submit(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results) {
  if (!cfg || !results || n_seq < 0 || (n_seq > 0 && !seqs)) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  int dev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = g_device;
  }
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  std::string err;
  std::unique_ptr<vce::Backend> be(vce::createCudaBackend(dev, err));
  if (!be) return fail(VCE_ERR_CUDA, err);
  vce::Pipeline pl(be.get(), vce::PipelineOptions());
  const int rc = pl.run(*cfg, n_seq, seqs, results);
  if (rc) return fail(rc, pl.error().empty() ? std::string("one or more sequences reported an error") : pl.error());
  return VCE_OK;
}

int vce_roc(const vce_roc_in* in, vce_roc_out* out) {
  if (!in || !out) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  int dev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = g_device;
  }
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  std::string err;
  const int rc = vce::rocRun(dev, in, out, err, nullptr);
  if (rc) return fail(rc, err);
  return VCE_OK;
}

// Timing variant used by bench.py: same computation, reports host->device and kernel milliseconds.
int vce_roc_timed(const vce_roc_in* in, vce_roc_out* out, double* h2d_ms, double* kernel_ms, int32_t* launches) {
  if (!in || !out) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  int dev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = g_device;
  }
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  std::string err;
  vce::RocTiming t;
  const int rc = vce::rocRun(dev, in, out, err, &t);
  if (rc) return fail(rc, err);
  if (h2d_ms) *h2d_ms = t.h2dMs;
  if (kernel_ms) *kernel_ms = t.kernelMs;
  if (launches) *launches = t.launches;
  return VCE_OK;
}

int vce_format_warning(const vce_warning* w, const char* seq_name, char* buf, int32_t cap) {
  // PathFinder.java:163
  return snprintf(buf, (size_t) cap,
                  "Evaluation too complex (%d unresolved paths, %d iterations) at reference region %s:%d-%d. "
                  "Variants in this region will not be included in results.",
                  w->paths, w->iterations, seq_name ? seq_name : "", w->start1, w->end1);
}

}  // extern "C"

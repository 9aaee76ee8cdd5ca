// vce_capi.cpp -- the C-ABI of include/vce.h on top of the engine + CUDA backend.
//
// There is NO CPU fallback: without a CUDA device every compute entry point fails with VCE_ERR_NO_DEVICE.
//
// Contexts (CUDA backend: streams, device arenas, page-locked staging + engine) are pooled and reused across calls,
// so that a steady-state vce_submit() pays no allocation; concurrent callers each get their own context and share
// the host worker pool.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vce.h"
#include "vce_cuda.h"
#include "vce_engine.h"

namespace {
std::mutex g_mu;
int g_device = -1;
thread_local std::string t_err;

struct Context {
  std::unique_ptr<vce::Backend> backend;
  std::unique_ptr<vce::Engine> engine;
  int device = -1;
};
std::vector<std::unique_ptr<Context>> g_free;
std::unique_ptr<vce::WorkerPool> g_pool;

int fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}

Context* acquire(int& rc) {
  int dev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = g_device;
    if (dev < 0) { rc = fail(VCE_ERR_NOT_INITIALISED, "call vce_init first"); return nullptr; }
    if (!g_pool) g_pool.reset(new vce::WorkerPool(vce::WorkerPool::defaultThreads()));
    for (size_t i = 0; i < g_free.size(); ++i) {
      if (g_free[i]->device == dev) {
        Context* c = g_free[i].release();
        g_free.erase(g_free.begin() + (long) i);
        return c;
      }
    }
  }
  std::string err;
  std::unique_ptr<vce::Backend> be(vce::createCudaBackend(dev, err));
  if (!be) { rc = fail(VCE_ERR_CUDA, err); return nullptr; }
  Context* c = new Context();
  c->device = dev;
  c->backend = std::move(be);
  vce::EngineOptions opt;
  if (const char* e = std::getenv("VCE_DISABLE_MICRO")) opt.disableMicro = std::atoi(e) != 0;   // diagnostics
  if (const char* e = std::getenv("VCE_CHUNK_REGIONS")) opt.chunkRegions = std::atoi(e);
  if (const char* e = std::getenv("VCE_SORT_PACKETS")) opt.sortPackets = std::atoi(e) != 0;
  if (const char* e = std::getenv("VCE_SPLIT_GAP")) opt.gap = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("VCE_DEVICE_BUILD")) opt.deviceBuild = std::atoi(e) != 0;
  if (const char* e = std::getenv("VCE_RESIDENT")) opt.residentInputs = std::atoi(e) != 0;
  if (const char* e = std::getenv("VCE_DEVICE_RESULTS")) opt.deviceResults = std::atoi(e) != 0;
  c->engine.reset(new vce::Engine(c->backend.get(), g_pool.get(), opt));
  return c;
}

void release(Context* c) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device < 0 || g_free.size() >= 4) { delete c; return; }
  g_free.emplace_back(c);
}

// A context whose call failed may still have work in flight on its streams, or a sticky CUDA error: it does not go back
// into the pool (the backend's destructor waits for the device before it frees anything).
void discard(Context* c) { delete c; }
}  // namespace

struct vce_job {
  Context* ctx = nullptr;
  double executeMs = 0;
  bool failed = false;
};

extern "C" {

const char* vce_version(void) { return "rtg-tools_b200 vce 0.2.0 (sm_100a)"; }

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
  g_free.clear();
  vce::templateRegistryRemove(-1, nullptr);
  g_device = -1;
}

int vce_template_register(const uint8_t* bases, int32_t len) {
  if (!bases || len <= 0) return fail(VCE_ERR_BAD_ARGUMENT, "null template");
  int dev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = g_device;
    if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
    if (!g_pool) g_pool.reset(new vce::WorkerPool(vce::WorkerPool::defaultThreads()));
  }
  std::vector<uint32_t> packed, nmask;
  vce::packTemplate(bases, len, g_pool.get(), packed, nmask);
  std::string err;
  const int rc = vce::templateRegistryAdd(dev, bases, len, packed.data(), packed.size(), nmask.data(), nmask.size(), err);
  if (rc) return fail(rc, err);
  return VCE_OK;
}

int vce_template_unregister(const uint8_t* bases) {
  vce::templateRegistryRemove(-1, bases);
  return VCE_OK;
}

int vce_job_create(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_job** job) {
  if (!cfg || !job || n_seq < 0 || (n_seq > 0 && !seqs)) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  int rc = VCE_OK;
  Context* c = acquire(rc);
  if (!c) return rc;
  c->backend->stats = vce::SearchStats();
  rc = c->engine->prepare(*cfg, n_seq, seqs, true);
  if (rc) {
    const std::string msg = c->engine->error();
    discard(c);
    return fail(rc, msg);
  }
  vce_job* j = new vce_job();
  j->ctx = c;
  *job = j;
  return VCE_OK;
}

int vce_job_run(vce_job* job) {
  if (!job || !job->ctx) return fail(VCE_ERR_BAD_ARGUMENT, "job not prepared");
  // inputs are resident on the device since vce_job_create; only the device pipeline runs here
  vce::Backend* be = job->ctx->backend.get();
  const int64_t h2d = be->stats.h2dBytes;
  be->stats = vce::SearchStats();
  be->stats.h2dBytes = h2d;
  be->timerStart();
  const int rc = job->ctx->engine->execute();
  job->executeMs = be->timerStopMs();
  if (rc) { job->failed = true; return fail(rc, job->ctx->engine->error()); }
  return VCE_OK;
}

int vce_job_fetch(vce_job* job, vce_sequence_result* results) {
  if (!job || !job->ctx || !results) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const int rc = job->ctx->engine->finish(results);
  if (rc) return fail(rc, "one or more sequences reported an error (see vce_sequence_result.error)");
  return VCE_OK;
}

void vce_job_destroy(vce_job* job) {
  if (!job) return;
  if (job->ctx) { if (job->failed) discard(job->ctx); else release(job->ctx); }
  delete job;
}

int vce_job_stats(const vce_job* job, int64_t* n_launches, double* device_ms, double* search_kernel_ms, int64_t* n_regions,
                  int64_t* n_escalated) {
  if (!job || !job->ctx) return fail(VCE_ERR_BAD_ARGUMENT, "null job");
  const vce::SearchStats& s = job->ctx->backend->stats;
  if (n_launches) *n_launches = s.launches;
  if (device_ms) *device_ms = s.searchMs;
  if (n_regions) *n_regions = s.regions;
  if (n_escalated) *n_escalated = s.escalated;
  if (search_kernel_ms) {
    double ms = 0;
    long long regs = 0;
    vce::cudaBackendMicroStats(job->ctx->backend.get(), &ms, &regs);
    *search_kernel_ms = ms;
  }
  return VCE_OK;
}

int vce_job_stats_ex(const vce_job* job, double* out, int32_t n) {
  if (!job || !job->ctx || !out) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const vce::SearchStats& s = job->ctx->backend->stats;
  double microMs = 0;
  long long microRegions = 0;
  vce::cudaBackendMicroStats(job->ctx->backend.get(), &microMs, &microRegions);
  const double v[13] = {(double) s.launches, s.searchMs, microMs, (double) microRegions, (double) s.regions, (double) s.escalated,
                        (double) s.spilled, (double) s.prefixReruns, (double) s.h2dBytes, (double) s.d2hBytes,
                        (double) s.iterations, job->executeMs, (double) job->ctx->engine->microVariants};
  for (int i = 0; i < n && i < 13; ++i) out[i] = v[i];
  return VCE_OK;
}

static thread_local double t_submitStats[6] = {0, 0, 0, 0, 0, 0};
static thread_local std::vector<int64_t> t_summary;

int vce_submit_summary(int64_t* out, int32_t n_seq) {
  if (!out || n_seq < 0) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  if ((size_t) n_seq * 8 > t_summary.size()) return fail(VCE_ERR_BAD_ARGUMENT, "no vce_submit of that many sequences on this thread");
  std::memcpy(out, t_summary.data(), (size_t) n_seq * 8 * sizeof(int64_t));
  return VCE_OK;
}

int vce_submit_stats(double* out, int32_t n) {
  if (!out) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  for (int i = 0; i < n && i < 6; ++i) out[i] = t_submitStats[i];
  return VCE_OK;
}

int vce_submit(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results) {
  if (!cfg || !results || n_seq < 0 || (n_seq > 0 && !seqs)) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  int rc = VCE_OK;
  Context* c = acquire(rc);
  if (!c) return rc;
  c->backend->stats = vce::SearchStats();
  if (const char* e = std::getenv("VCE_DEVICE_BUILD")) c->engine->options().deviceBuild = std::atoi(e) != 0;   // contexts are pooled
  rc = c->engine->run(*cfg, n_seq, seqs, results);
  const std::string msg = rc ? c->engine->error() : std::string();
  t_summary = c->engine->summary();
  {
    const vce::SearchStats& s = c->backend->stats;
    t_submitStats[0] = (double) s.launches; t_submitStats[1] = (double) s.h2dBytes; t_submitStats[2] = (double) s.d2hBytes;
    t_submitStats[3] = (double) s.regions; t_submitStats[4] = (double) s.escalated; t_submitStats[5] = (double) s.iterations;
  }
  // per-sequence errors (results[k].error) are ordinary outcomes; anything else leaves the context in an unknown state
  bool perSequence = false;
  if (rc) for (int32_t k = 0; k < n_seq; ++k) if (results[k].error == rc) perSequence = true;
  if (rc && !perSequence) discard(c); else release(c);
  if (rc) return fail(rc, msg.empty() ? std::string("one or more sequences reported an error") : msg);
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

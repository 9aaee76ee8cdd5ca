// vce_capi.cpp -- the C-ABI of include/vce.h on top of the engine + CUDA backend.
//
// There is NO CPU fallback: without a CUDA device every compute entry point fails with VCE_ERR_NO_DEVICE.
//
// Contexts (CUDA backend: streams, device arenas, page-locked staging + engine) are pooled and reused across calls,
// so that a steady-state vce_submit() pays no allocation; concurrent callers each get their own context and share
// the host worker pool.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <array>
#include <thread>
#include <vector>

#include "../../include/vce.h"
#include "vce_cuda.h"
#include "vce_bgzf.h"
#include "vce_split.h"
#include "vce_vcf.h"
#include "vce_engine.h"

namespace {
std::mutex g_mu;
int g_device = -1;                 // the first device of g_devices (single-device entry points: jobs, ROC)
std::vector<int> g_devices;        // vce_init_devices: a batch's sequences are spread over these
thread_local std::string t_err;

struct Context {
  std::unique_ptr<vce::Backend> backend;
  std::unique_ptr<vce::Engine> engine;
  int device = -1;
};
std::vector<std::unique_ptr<Context>> g_free;
std::unique_ptr<vce::WorkerPool> g_pool;
// with several devices every device's engine gets its own share of the host cores (a WorkerPool runs one parallelFor at a time)
std::vector<std::unique_ptr<vce::WorkerPool>> g_devPools;
std::vector<int64_t> g_devLoad;   // variants queued per device by the one-sequence calls in flight (under g_mu)

int fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}

Context* acquire(int& rc, int devIndex = -1) {
  int dev;
  vce::WorkerPool* pool = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = devIndex >= 0 && devIndex < (int) g_devices.size() ? g_devices[(size_t) devIndex] : g_device;
    if (dev < 0) { rc = fail(VCE_ERR_NOT_INITIALISED, "call vce_init first"); return nullptr; }
    if (!g_pool) g_pool.reset(new vce::WorkerPool(vce::WorkerPool::defaultThreads()));
    pool = g_pool.get();
    if (devIndex >= 0 && g_devices.size() > 1) {
      if (g_devPools.size() != g_devices.size()) {
        g_devPools.clear();
        const int per = std::max(2, vce::WorkerPool::defaultThreads() / (int) g_devices.size());
        for (size_t i = 0; i < g_devices.size(); ++i) g_devPools.emplace_back(new vce::WorkerPool(per));
      }
      pool = g_devPools[(size_t) devIndex].get();
    }
    for (size_t i = 0; i < g_free.size(); ++i) {
      if (g_free[i]->device == dev && (g_devices.size() <= 1 || g_free[i]->engine->pool() == pool)) {
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
  c->engine.reset(new vce::Engine(c->backend.get(), pool, opt));
  return c;
}

void release(Context* c) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device < 0 || g_free.size() >= 4 + 2 * g_devices.size()) { delete c; return; }
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

int vce_init_devices(int device_count, const int* devices) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int n = vce::cudaDeviceCount();
  if (n <= 0) return fail(VCE_ERR_NO_DEVICE, "no CUDA device is visible; this engine has no CPU fallback");
  if (device_count < 1 || !devices) return fail(VCE_ERR_BAD_ARGUMENT, "no device given");
  for (int i = 0; i < device_count; ++i) {
    if (devices[i] < 0 || devices[i] >= n) return fail(VCE_ERR_BAD_ARGUMENT, "device index out of range");
    for (int q = 0; q < i; ++q) if (devices[q] == devices[i]) return fail(VCE_ERR_BAD_ARGUMENT, "device listed twice");
  }
  g_devices.assign(devices, devices + device_count);
  g_device = devices[0];
  g_free.clear();      // (contexts are per device and per worker pool)
  g_devPools.clear();
  return VCE_OK;
}

int vce_init(int device) { return vce_init_devices(1, &device); }

void vce_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_free.clear();
  g_devPools.clear();
  g_devices.clear();
  vce::rocRelease();
  vce::vcfShutdown();
  vce::templateRegistryRemove(-1, nullptr);
  vce::pinnedRegistryRemove(nullptr);
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
  std::vector<int> devs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    devs = g_devices;
  }
  (void) dev;
  std::sort(devs.begin(), devs.end());
  devs.erase(std::unique(devs.begin(), devs.end()), devs.end());
  for (int d : devs) {   // every device of the pool: the batch's sequences may land on any of them
    std::string err;
    const int rc = vce::templateRegistryAdd(d, bases, len, packed.data(), packed.size(), nmask.data(), nmask.size(), err);
    if (rc) return fail(rc, err);
  }
  return VCE_OK;
}

int vce_host_register(const void* p, uint64_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  }
  std::string err;
  const int rc = vce::pinnedRegistryAdd(p, (size_t) bytes, err);
  if (rc) return fail(rc, err);
  return VCE_OK;
}

int vce_host_unregister(const void* p) {
  vce::pinnedRegistryRemove(p);
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
static thread_local double t_splitStats[3] = {0, 0, 0};   // sequences cut, pieces, sequences evaluated again in one piece
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
  for (int i = 6; i < n && i < 9; ++i) out[i] = t_splitStats[i - 6];
  return VCE_OK;
}

// One device's share of a batch: the sequences `idx` through that device's own context and worker pool.
static int submitOn(int devIndex, const vce_config* cfg, const std::vector<int32_t>& idx, const vce_sequence* seqs,
                    vce_sequence_result* results, double stats[6], std::vector<int64_t>& summary, std::string& msg) {
  int rc = VCE_OK;
  Context* c = acquire(rc, devIndex);
  if (!c) { msg = t_err; return rc; }
  std::vector<vce_sequence> sub(idx.size());
  std::vector<vce_sequence_result> res(idx.size());
  for (size_t i = 0; i < idx.size(); ++i) { sub[i] = seqs[idx[i]]; res[i] = results[idx[i]]; }
  c->backend->stats = vce::SearchStats();
  rc = c->engine->run(*cfg, (int32_t) idx.size(), sub.data(), res.data());
  if (rc) msg = c->engine->error();
  const vce::SearchStats& st = c->backend->stats;
  stats[0] = (double) st.launches; stats[1] = (double) st.h2dBytes; stats[2] = (double) st.d2hBytes;
  stats[3] = (double) st.regions; stats[4] = (double) st.escalated; stats[5] = (double) st.iterations;
  const std::vector<int64_t>& sm = c->engine->summary();
  bool perSequence = false;
  for (size_t i = 0; i < idx.size(); ++i) {
    results[idx[i]] = res[i];
    if (rc && res[i].error == rc) perSequence = true;
    if (sm.size() >= (i + 1) * 8) for (int q = 0; q < 8; ++q) summary[(size_t) idx[i] * 8 + q] = sm[i * 8 + q];
  }
  if (rc && !perSequence) discard(c); else release(c);
  return rc;
}

static vce::WorkerPool* rocPool();

int vce_submit(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results) {
  if (!cfg || !results || n_seq < 0 || (n_seq > 0 && !seqs)) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  size_t nDev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    nDev = g_devices.size();
  }
  if (nDev > 1 && n_seq == 1) {
    // one sequence per call (the reference's pool workers, VcfEvalTask.java:131-137): the device with the least work queued
    const int64_t w = (int64_t) seqs[0].baseline.n + (int64_t) seqs[0].calls.n + 1;
    size_t best = 0;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if (g_devLoad.size() != nDev) g_devLoad.assign(nDev, 0);
      for (size_t d = 1; d < nDev; ++d) if (g_devLoad[d] < g_devLoad[best]) best = d;
      g_devLoad[best] += w;
    }
    std::vector<int32_t> idx(1, 0);
    std::array<double, 6> st;
    for (double& v : st) v = 0;
    std::vector<int64_t> summary(8, 0);
    std::string msg;
    const int rc = submitOn((int) best, cfg, idx, seqs, results, st.data(), summary, msg);
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if (g_devLoad.size() == nDev) g_devLoad[best] -= w;
    }
    for (int i = 0; i < 6; ++i) t_submitStats[i] = st[(size_t) i];
    t_summary = summary;
    if (rc) return fail(rc, msg.empty() ? std::string("one or more sequences reported an error") : msg);
    return VCE_OK;
  }
  if (nDev > 1 && n_seq > 1) {
    // Sequences are independent (VcfEvalTask.java:131-135): longest-processing-time greedy by variant count over the devices,
    // one host thread per device, each with its own context, streams and share of the host cores; no data-path exchange.
    // When whole sequences do not balance (max / mean > 1.1) the long ones are cut into pieces at wide gaps,
    // evaluated as sequences of their own and stitched back after verification (vce_split.h).
    vce::SplitOptions so;
    if (const char* e = std::getenv("VCE_SPLIT_MIN")) so.minVariants = std::atoll(e);
    if (const char* e = std::getenv("VCE_SPLIT_OVERLAP")) so.overlap = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("VCE_SPLIT_IMBALANCE")) so.imbalance = std::atof(e);
    if (const char* e = std::getenv("VCE_SPLIT_PIECES")) so.piecesPerDevice = std::max(1, std::atoi(e));
    vce::SplitPlan plan;
    const auto tCall = std::chrono::steady_clock::now();
    // planning and stitching are per-sequence tasks of the shared worker pool (no engine is running at those moments)
    vce::WorkerPool* wp = rocPool();
    auto par = [wp](int n, const std::function<void(int)>& fn) { if (n > 0) wp->parallelFor(n, [&](int t, int) { fn(t); }); };
    const char* se = std::getenv("VCE_SPLIT");
    so.force = se && std::atoi(se) == 1;   // 1: SURVEY 8e's rule alone (max / mean > 1.1), without the cost estimate
    if (se && std::atoi(se) == 0) { for (int32_t k = 0; k < n_seq; ++k) plan.whole.push_back(k); }
    else vce::splitPlan(*cfg, n_seq, seqs, results, (int) nDev, so, plan, par);
    const bool timing = std::getenv("VCE_TIMING") != nullptr;
    const auto tPlan = std::chrono::steady_clock::now();
    if (timing && !plan.split.empty()) std::fprintf(stderr, "[vce] split: planning %.2f ms\n", std::chrono::duration<double, std::milli>(tPlan - tCall).count());
    std::vector<int64_t> summary((size_t) n_seq * 8, 0);
    for (int i = 0; i < 6; ++i) t_submitStats[i] = 0;
    int hardRc = VCE_OK;
    std::string hardMsg;
    // one round: items (whole sequences and pieces) over the devices
    auto round = [&](const std::vector<vce_sequence>& xs, std::vector<vce_sequence_result>& xr, std::vector<int64_t>& xsum) {
      std::vector<int64_t> w(xs.size());
      for (size_t i = 0; i < xs.size(); ++i) w[i] = (int64_t) xs[i].baseline.n + (int64_t) xs[i].calls.n;
      std::vector<int> dev;
      vce::splitAssign(w, (int) nDev, dev);
      std::vector<std::vector<int32_t>> share(nDev);
      for (size_t i = 0; i < xs.size(); ++i) share[(size_t) dev[i]].push_back((int32_t) i);
      std::vector<int> rcs(nDev, VCE_OK);
      std::vector<std::string> msgs(nDev);
      std::vector<std::array<double, 6>> stats(nDev);
      xsum.assign(xs.size() * 8, 0);
      std::vector<std::thread> threads;
      for (size_t d = 0; d < nDev; ++d) {
        for (double& v : stats[d]) v = 0;
        if (share[d].empty()) continue;
        threads.emplace_back([&, d] { rcs[d] = submitOn((int) d, cfg, share[d], xs.data(), xr.data(), stats[d].data(), xsum, msgs[d]); });
      }
      for (std::thread& t : threads) t.join();
      for (int i = 0; i < 6; ++i) for (size_t d = 0; d < nDev; ++d) t_submitStats[i] += stats[d][(size_t) i];
      for (size_t d = 0; d < nDev; ++d) {
        if (!rcs[d]) continue;
        bool perSequence = false;
        for (int32_t i : share[d]) if (xr[(size_t) i].error == rcs[d]) perSequence = true;
        if (!perSequence && !hardRc) { hardRc = rcs[d]; hardMsg = msgs[d]; }
      }
    };
    std::vector<vce_sequence> xs;
    std::vector<vce_sequence_result> xr;
    for (int32_t k : plan.whole) { xs.push_back(seqs[k]); xr.push_back(results[k]); }
    for (auto& p : plan.pieces) { xs.push_back(p->view); xr.push_back(p->res); }
    std::vector<int64_t> xsum;
    round(xs, xr, xsum);
    if (hardRc) return fail(hardRc, hardMsg.empty() ? std::string("device failure") : hardMsg);
    for (size_t i = 0; i < plan.whole.size(); ++i) {
      const int32_t k = plan.whole[i];
      results[k] = xr[i];
      for (int q = 0; q < 8; ++q) summary[(size_t) k * 8 + q] = xsum[i * 8 + (size_t) q];
    }
    for (size_t i = 0; i < plan.pieces.size(); ++i) plan.pieces[i]->res = xr[plan.whole.size() + i];
    std::vector<int32_t> redo;
    const auto tRun = std::chrono::steady_clock::now();
    vce::splitMergeAll(plan, seqs, results, summary.data(), redo, par);
    if (timing && !plan.split.empty()) {
      const auto tMerge = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[vce] split: %zu sequences cut into %zu pieces; devices %.2f ms, stitching %.2f ms, %zu to evaluate again\n",
                   plan.split.size(), plan.pieces.size(), std::chrono::duration<double, std::milli>(tRun - tPlan).count(),
                   std::chrono::duration<double, std::milli>(tMerge - tRun).count(), redo.size());
    }
    t_splitStats[0] = (double) plan.split.size(); t_splitStats[1] = (double) plan.pieces.size(); t_splitStats[2] = (double) redo.size();
    if (!redo.empty()) {   // a cut that could not be verified (or a piece with an error): those sequences in one piece
      std::vector<vce_sequence> ys;
      std::vector<vce_sequence_result> yr;
      for (int32_t k : redo) { ys.push_back(seqs[k]); yr.push_back(results[k]); }
      std::vector<int64_t> ysum;
      round(ys, yr, ysum);
      if (hardRc) return fail(hardRc, hardMsg.empty() ? std::string("device failure") : hardMsg);
      for (size_t i = 0; i < redo.size(); ++i) {
        results[redo[i]] = yr[i];
        for (int q = 0; q < 8; ++q) summary[(size_t) redo[i] * 8 + q] = ysum[i * 8 + (size_t) q];
      }
    }
    t_summary = summary;
    for (int32_t k = 0; k < n_seq; ++k) {
      if (results[k].error) return fail(results[k].error, "one or more sequences reported an error");
    }
    return VCE_OK;
  }
  int rc = VCE_OK;
  Context* c = acquire(rc);
  if (!c) return rc;
  c->backend->stats = vce::SearchStats();
  // (contexts are pooled: the switch is applied on every call, so that a context does not keep the mode of an earlier call)
  { const char* e = std::getenv("VCE_DEVICE_BUILD"); c->engine->options().deviceBuild = e ? std::atoi(e) != 0 : true; }
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

static vce::WorkerPool* rocPool() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_pool) g_pool.reset(new vce::WorkerPool(vce::WorkerPool::defaultThreads()));
  return g_pool.get();
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
  const int rc = vce::rocRun(dev, in, out, err, nullptr, rocPool());
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
  const int rc = vce::rocRun(dev, in, out, err, &t, rocPool());
  if (rc) return fail(rc, err);
  if (h2d_ms) *h2d_ms = t.h2dMs;
  if (kernel_ms) *kernel_ms = t.kernelMs;
  if (launches) *launches = t.launches;
  return VCE_OK;
}

// ------------------------------------------------------------------------------------------- VCF loader (vce_vcf.h)
namespace {
// shared by the CUDA library and nothing else: fills the C view of a result
void vcfView(const vce::VcfResult& R, vce_variants* out) {
  out->n = R.n;
  out->gt_ploidy = R.ploidy;
  out->id = R.id;
  out->phased = R.phased;
  out->status_in = R.statusIn;
  out->gt = R.gt;
  out->allele_off = R.alleleOff;
  out->n_slots = R.nSlots;
  out->allele_start = R.alleleStart;
  out->allele_end = R.alleleEnd;
  out->allele_null = R.alleleNull;
  out->allele_nt_off = R.alleleNtOff;
  out->nt = R.nt;
}
}  // namespace

int vce_vcf_parse(const vce_vcf_in* in, void** result) {
  if (!in || !result || (!in->text && in->text_len > 0)) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  *result = nullptr;
  int dev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    dev = g_device;
  }
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  vce::VcfConfig cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.sample = in->sample; cfg.ploidy = in->ploidy; cfg.passOnly = in->pass_only; cfg.maxLength = in->max_length;
  cfg.templateLen = in->template_len;
  cfg.refOverlap = in->ref_overlap;
  cfg.evalRegions = in->eval_regions;
  cfg.nEvalRegions = in->eval_regions ? in->n_eval_regions : -1;
  cfg.regionBeg = in->region_begin;
  cfg.regionEnd = in->region_end;
  if (in->score_field && std::strcmp(in->score_field, "QUAL") != 0) {
    const size_t sl = std::strlen(in->score_field);
    if (sl == 0 || sl >= sizeof(cfg.scoreField)) return fail(VCE_ERR_BAD_ARGUMENT, "score field name too long");
    std::memcpy(cfg.scoreField, in->score_field, sl);
    cfg.scoreLen = (int32_t) sl;
  }
  if (in->name) {
    const size_t nl = std::strlen(in->name);
    if (nl == 0 || nl >= sizeof(cfg.name)) return fail(VCE_ERR_BAD_ARGUMENT, "sequence name empty or too long");
    std::memcpy(cfg.name, in->name, nl);
    cfg.nameLen = (int32_t) nl;
  }
  std::unique_ptr<vce::VcfResult> R(new vce::VcfResult());
  std::string err;
  const int rc = vce::vcfParseOnDevice(dev, reinterpret_cast<const uint8_t*>(in->text), in->text_len, cfg, *R, err);
  if (rc) return fail(rc, err);
  *result = R.release();
  return VCE_OK;
}

int vce_vcf_variants(const void* result, vce_variants* out, vce_vcf_stats* stats) {
  if (!result || !out) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const vce::VcfResult& R = *static_cast<const vce::VcfResult*>(result);
  vcfView(R, out);
  if (stats) {
    stats->lines = R.stats.lines; stats->records = R.stats.records; stats->kept = R.stats.kept; stats->skipped = R.stats.skipped;
    stats->h2d_ms = R.stats.h2dMs; stats->kernel_ms = R.stats.kernelMs; stats->d2h_ms = R.stats.d2hMs;
    stats->launches = R.stats.launches; stats->reserved = 0;
  }
  return VCE_OK;
}

void vce_vcf_free(void* result) { delete static_cast<vce::VcfResult*>(result); }

int vce_vcf_roc_fields(const void* result, const uint32_t** filter_mask, const double** qual) {
  if (!result || !filter_mask || !qual) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const vce::VcfResult& R = *static_cast<const vce::VcfResult*>(result);
  *filter_mask = R.rocMask;
  *qual = R.qual;
  return VCE_OK;
}

// ------------------------------------------------------------------------------- on-disk formats (vce_bgzf.h)
namespace {
int currentDevice() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_device;
}
void bgzfStatsOut(const vce::BgzfStats& st, vce_bgzf_stats* out) {
  if (!out) return;
  out->blocks = st.blocks; out->compressed_bytes = st.compressed; out->text_bytes = st.text;
  out->h2d_ms = st.h2dMs; out->kernel_ms = st.kernelMs; out->d2h_ms = st.d2hMs;
  out->launches = st.launches; out->reserved = 0;
}
int vcfConfigFrom(const vce_vcf_in* in, vce::VcfConfig& cfg) {
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.sample = in->sample; cfg.ploidy = in->ploidy; cfg.passOnly = in->pass_only; cfg.maxLength = in->max_length;
  cfg.templateLen = in->template_len;
  cfg.refOverlap = in->ref_overlap;
  cfg.evalRegions = in->eval_regions;
  cfg.nEvalRegions = in->eval_regions ? in->n_eval_regions : -1;
  cfg.regionBeg = in->region_begin;
  cfg.regionEnd = in->region_end;
  if (in->score_field && std::strcmp(in->score_field, "QUAL") != 0) {
    const size_t sl = std::strlen(in->score_field);
    if (sl == 0 || sl >= sizeof(cfg.scoreField)) return fail(VCE_ERR_BAD_ARGUMENT, "score field name too long");
    std::memcpy(cfg.scoreField, in->score_field, sl);
    cfg.scoreLen = (int32_t) sl;
  }
  if (in->name) {
    const size_t nl = std::strlen(in->name);
    if (nl == 0 || nl >= sizeof(cfg.name)) return fail(VCE_ERR_BAD_ARGUMENT, "sequence name empty or too long");
    std::memcpy(cfg.name, in->name, nl);
    cfg.nameLen = (int32_t) nl;
  }
  return VCE_OK;
}
}  // namespace

int vce_bgzf_text_size(const uint8_t* data, int64_t len, int64_t voffset_begin, int64_t voffset_end, int64_t* text_len) {
  if (!data || !text_len || len < 0 || voffset_begin < 0) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const std::string w = vce::bgzfTextSize(data, len, voffset_begin, voffset_end, text_len);
  if (!w.empty()) return fail(VCE_ERR_BAD_ARGUMENT, w);
  return VCE_OK;
}

int vce_bgzf_inflate(const uint8_t* data, int64_t len, int64_t voffset_begin, int64_t voffset_end, int32_t check_crc, uint8_t* text,
                     int64_t text_cap, int64_t* text_len, vce_bgzf_stats* stats) {
  if (!data || !text_len || (!text && text_cap > 0)) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const int dev = currentDevice();
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  std::vector<uint8_t> out;
  vce::BgzfStats st;
  std::string err;
  const int rc = vce::bgzfInflateOnDevice(dev, data, len, voffset_begin, voffset_end, check_crc != 0, out, st, err);
  if (rc) return fail(rc, err);
  *text_len = (int64_t) out.size();
  bgzfStatsOut(st, stats);
  if ((int64_t) out.size() > text_cap) return fail(VCE_ERR_CAPACITY, "text buffer too small (vce_bgzf_text_size tells the size)");
  if (!out.empty()) std::memcpy(text, out.data(), out.size());
  return VCE_OK;
}

int vce_vcf_parse_bgzf(const vce_vcf_in* in, int64_t voffset_begin, int64_t voffset_end, int32_t check_crc, void** result,
                       vce_bgzf_stats* stats) {
  if (!in || !result || !in->text) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  *result = nullptr;
  const int dev = currentDevice();
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  vce::VcfConfig cfg;
  if (int rc = vcfConfigFrom(in, cfg)) return rc;
  std::unique_ptr<vce::VcfResult> R(new vce::VcfResult());
  vce::BgzfStats st;
  std::string err;
  const int rc = vce::vcfParseBgzfOnDevice(dev, reinterpret_cast<const uint8_t*>(in->text), in->text_len, voffset_begin, voffset_end,
                                           check_crc != 0, cfg, *R, st, err);
  if (rc) return fail(rc, err);
  bgzfStatsOut(st, stats);
  *result = R.release();
  return VCE_OK;
}

int vce_tabix_open(const uint8_t* tbi, int64_t len, void** index) {
  if (!tbi || !index || len < 0) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  *index = nullptr;
  const int dev = currentDevice();
  if (dev < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
  // TabixIndexReader.java:65-67
  if (len < 18 || tbi[0] != 0x1f || tbi[1] != 0x8b) return fail(VCE_ERR_BAD_ARGUMENT, "not a valid TABIX index. (index file is not block compressed)");
  std::unique_ptr<vce::TabixIndex> t(new vce::TabixIndex());
  vce::BgzfStats st;
  std::string err;
  const int rc = vce::bgzfInflateOnDevice(dev, tbi, len, 0, -1, true, t->raw, st, err);
  if (rc) return fail(rc, err.find("not a block compressed") != std::string::npos ? "not a valid TABIX index. (index file is not block compressed)" : err);
  const std::string w = vce::tabixParse(*t);
  if (!w.empty()) return fail(VCE_ERR_BAD_ARGUMENT, w);
  *index = t.release();
  return VCE_OK;
}

int vce_tabix_info(const void* index, int32_t out[7]) {
  if (!index || !out) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  const vce::TabixIndex& t = *static_cast<const vce::TabixIndex*>(index);
  out[0] = (int32_t) t.names.size(); out[1] = t.format; out[2] = t.seqCol; out[3] = t.begCol; out[4] = t.endCol; out[5] = t.meta; out[6] = t.skip;
  return VCE_OK;
}

const char* vce_tabix_name(const void* index, int32_t i) {
  if (!index) return nullptr;
  const vce::TabixIndex& t = *static_cast<const vce::TabixIndex*>(index);
  return i >= 0 && i < (int32_t) t.names.size() ? t.names[(size_t) i].c_str() : nullptr;
}

int vce_tabix_query(const void* index, const char* name, int32_t begin0, int32_t end0, int64_t* voffset_begin, int64_t* voffset_end) {
  if (!index || !name || !voffset_begin || !voffset_end) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  std::string err;
  *voffset_begin = *voffset_end = -1;
  const int r = vce::tabixQuery(*static_cast<const vce::TabixIndex*>(index), name, begin0, end0, voffset_begin, voffset_end, err);
  if (r < 0) return fail(VCE_ERR_BAD_ARGUMENT, err);
  if (r == 0) *voffset_begin = *voffset_end = -1;
  return VCE_OK;
}

void vce_tabix_close(void* index) { delete static_cast<vce::TabixIndex*>(index); }

int vce_template_register_sdf(const uint8_t* seqdata, int64_t seqdata_len, int64_t first_residue, int32_t len, uint8_t* bases_out,
                              double* ms3) {
  if (!seqdata || !bases_out || len <= 0 || first_residue < 0) return fail(VCE_ERR_BAD_ARGUMENT, "null argument");
  std::vector<int> devs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device < 0) return fail(VCE_ERR_NOT_INITIALISED, "call vce_init first");
    devs = g_devices;
  }
  std::sort(devs.begin(), devs.end());
  devs.erase(std::unique(devs.begin(), devs.end()), devs.end());
  // only the longs that hold the sequence travel: groups of 64 residues, three big-endian longs each
  const int64_t g0 = first_residue >> 6;
  const int64_t rel = first_residue - (g0 << 6);
  const int64_t need = ((rel + len + 63) >> 6) * 24;
  if (g0 * 24 + need > seqdata_len) return fail(VCE_ERR_BAD_ARGUMENT, "SDF: the data does not cover the sequence");
  for (int d : devs) {   // every device of the pool, like vce_template_register
    std::string err;
    const int rc = vce::sdfTemplateOnDevice(d, seqdata + g0 * 24, need, rel, len, bases_out, ms3, err);
    if (rc) return fail(rc, err);
  }
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

// vce_pool.h -- a small persistent worker pool for the host side of the engine (digest of the caller's arrays into
// region packets, decoding of result records, result assembly).  The reference runs one SequenceEvaluator per
// sequence on its SimpleThreadPool (VcfEvalTask.java:131-137); here the unit is a chunk of regions / variants so that
// one large chromosome does not bound the parallelism.
#ifndef VCE_POOL_H_
#define VCE_POOL_H_

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vce {

class WorkerPool {
 public:
  explicit WorkerPool(int threads) {
    if (threads < 1) threads = 1;
    nThreads_ = threads;
    for (int i = 1; i < threads; ++i) workers_.emplace_back([this, i] { loop(i); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
  }
  int threads() const { return nThreads_; }

  // Runs fn(task, worker) for task in [0, nTasks), tasks handed out dynamically in ascending order; the caller
  // participates as worker 0.  Returns when every task has finished.  One parallelFor at a time (callers serialise).
  void parallelFor(int nTasks, const std::function<void(int, int)>& fn) {
    if (nTasks <= 0) return;
    if (nThreads_ == 1 || nTasks == 1) {
      for (int t = 0; t < nTasks; ++t) fn(t, 0);
      return;
    }
    std::unique_lock<std::mutex> run(runMu_);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      nTasks_ = nTasks;
      next_.store(0);
      active_ = (int) workers_.size();
      ++epoch_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> lk(mu_);
    doneCv_.wait(lk, [this] { return active_ == 0; });
    fn_ = nullptr;
  }

  static int defaultThreads() {
    if (const char* e = std::getenv("VCE_THREADS")) {
      const int v = std::atoi(e);
      if (v > 0) return v;
    }
    unsigned h = std::thread::hardware_concurrency();
    if (h == 0) h = 4;
    return (int) (h > 64 ? 64 : h);
  }

 private:
  void work(int worker) {
    const std::function<void(int, int)>& f = *fn_;
    while (true) {
      const int t = next_.fetch_add(1);
      if (t >= nTasks_) break;
      f(t, worker);
    }
  }
  void loop(int worker) {
    long seen = 0;
    while (true) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
      }
      work(worker);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--active_ == 0) doneCv_.notify_all();
      }
    }
  }

  int nThreads_ = 1;
  std::vector<std::thread> workers_;
  std::mutex mu_, runMu_;
  std::condition_variable cv_, doneCv_;
  const std::function<void(int, int)>* fn_ = nullptr;
  int nTasks_ = 0;
  std::atomic<int> next_{0};
  int active_ = 0;
  long epoch_ = 0;
  bool stop_ = false;
};

}  // namespace vce
#endif

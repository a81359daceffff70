// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// Shared-memory ParallelFor standing in for internal/ceres/parallel_for.h:69-183
// + thread_pool.cc: contiguous index ranges handed to a persistent pool of
// worker threads.  f is called as f(thread_id, index).
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace orc {

class Pool {
 public:
  static Pool& Get() {
    static Pool p;
    return p;
  }
  int max_threads() const { return static_cast<int>(workers_.size()) + 1; }

  // Runs job(t) for t in [0, n) on n threads (caller is thread 0).
  void Run(int n, const std::function<void(int)>& job) {
    if (n <= 1) {
      job(0);
      return;
    }
    EnsureWorkers(n - 1);
    {
      std::unique_lock<std::mutex> lk(mu_);
      job_ = &job;
      active_ = n - 1;
      pending_ = n - 1;
      ++generation_;
    }
    cv_.notify_all();
    job(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }

  ~Pool() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }

 private:
  void EnsureWorkers(int n) {
    while (static_cast<int>(workers_.size()) < n) {
      const int id = static_cast<int>(workers_.size());
      std::unique_lock<std::mutex> lk(mu_);
      const long gen = generation_;
      lk.unlock();
      workers_.emplace_back([this, id, gen] { Loop(id, gen); });
    }
  }
  void Loop(int id, long seen) {
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        if (id < active_) job = job_;
      }
      if (job != nullptr) {
        (*job)(id + 1);
        std::unique_lock<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_one();
      }
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* job_ = nullptr;
  long generation_ = 0;
  int active_ = 0, pending_ = 0;
  bool stop_ = false;
};

// f(thread_id, i) for i in [start, end); static contiguous partition.
template <typename F>
inline void ParallelFor(int start, int end, int num_threads, F&& f) {
  const int n = end - start;
  if (n <= 0) return;
  if (num_threads <= 1 || n == 1) {
    for (int i = start; i < end; ++i) f(0, i);
    return;
  }
  const int nt = num_threads < n ? num_threads : n;
  std::function<void(int)> job = [&](int t) {
    const long lo = start + static_cast<long>(n) * t / nt;
    const long hi = start + static_cast<long>(n) * (t + 1) / nt;
    for (long i = lo; i < hi; ++i) f(t, static_cast<int>(i));
  };
  Pool::Get().Run(nt, job);
}

// f(thread_id, i) over [start,end) with partition boundaries chosen so every
// thread gets about the same cumulative weight (parallel_for.h:154-183 uses
// cumulative nnz the same way).  cum[i] = total weight of items [start, i].
template <typename F>
inline void ParallelForWeighted(int start, int end, int num_threads,
                                const long* cum /* indexed by i - start */, F&& f) {
  const int n = end - start;
  if (n <= 0) return;
  if (num_threads <= 1 || n == 1) {
    for (int i = start; i < end; ++i) f(0, i);
    return;
  }
  const int nt = num_threads < n ? num_threads : n;
  std::vector<int> bounds(nt + 1, end);
  bounds[0] = start;
  const long total = cum[n - 1];
  int i = 0;
  for (int t = 1; t < nt; ++t) {
    const long target = total * t / nt;
    while (i < n && cum[i] < target) ++i;
    bounds[t] = start + i;
  }
  std::function<void(int)> job = [&](int t) {
    for (int k = bounds[t]; k < bounds[t + 1]; ++k) f(t, k);
  };
  Pool::Get().Run(nt, job);
}

}  // namespace orc

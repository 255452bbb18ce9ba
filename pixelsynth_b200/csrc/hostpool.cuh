// A small persistent team of host worker threads for the native host code (order / masks in glue.cu, dependency levels in
// lmconv_tc.cu).  That code sits on the step's critical path -- the GPU has only the VQ-VAE encoder to do meanwhile -- and
// its parallel regions are a few hundred microseconds each: creating and joining 16 std::threads three or four times per
// step cost more than the work.  The team is created on first use, parked on a condition variable between regions,
// recreated after a fork (the child has no threads), and a region that cannot get the team (creation failed) runs inline.
#pragma once
#include <unistd.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace ps {

class HostPool {
 public:
  static HostPool& get() {
    static HostPool pool;
    return pool;
  }
  // fn(c) for c in [0, T): worker c - 1 runs fn(c) for c >= 1, the caller runs fn(0).  One region at a time.
  template <class F>
  void run(int T, F&& fn) {
    if (T <= 1) {
      fn(0);
      return;
    }
    std::lock_guard<std::mutex> region(region_mu_);
    ensure(T - 1);
    const int helpers = std::min<int>(T - 1, (int)workers_.size());
    // worker w takes c = w + 1, w + 1 + (helpers + 1), ... so that all T shares run even with fewer helpers than asked
    const int stride = helpers + 1;
    std::function<void(int)> body = [&](int first) {
      for (int c = first; c < T; c += stride) fn(c);
    };
    {
      std::lock_guard<std::mutex> lk(mu_);
      body_ = &body;
      active_ = helpers;
      pending_ = helpers;
      ++generation_;
    }
    cv_work_.notify_all();
    body(0);
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
    body_ = nullptr;
  }

  ~HostPool() { shutdown(); }

 private:
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_work_.notify_all();
    if (pid_ == getpid())
      for (auto& t : workers_)
        if (t.joinable()) t.join();
    else
      for (auto& t : workers_) t.detach();  // after a fork the threads do not exist in this process
    workers_.clear();
    stop_ = false;
  }
  void ensure(int want) {
    if (pid_ != getpid()) {  // forked: the parent's workers are not here
      for (auto& t : workers_) t.detach();
      workers_.clear();
      pid_ = getpid();
    }
    while ((int)workers_.size() < want && (int)workers_.size() < 31) {
      const int w = (int)workers_.size();
      unsigned long long seen0;
      {
        std::lock_guard<std::mutex> lk(mu_);
        seen0 = generation_;  // the region that is about to start must not be missed by a worker that starts late
      }
      try {
        workers_.emplace_back([this, w, seen0] { loop(w, seen0); });
      } catch (...) {
        break;  // fewer helpers: the shares are redistributed in run()
      }
    }
  }
  void loop(int w, unsigned long long seen) {
    for (;;) {
      std::function<void(int)>* body = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        if (w < active_) body = body_;
      }
      if (body) {
        (*body)(w + 1);
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) cv_done_.notify_one();
      }
    }
  }

  std::mutex region_mu_, mu_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  std::function<void(int)>* body_ = nullptr;
  int active_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
  bool stop_ = false;
  pid_t pid_ = getpid();
};

}  // namespace ps

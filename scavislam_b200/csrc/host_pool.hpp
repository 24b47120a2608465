// Host-side helpers of svs_ba_set_problem (no CUDA in here).
//
// SpinPool: a few worker threads that are parked on a condition variable between calls and SPIN for work while a call
// is in flight (begin() ... end()).  A back-end tick arrives every few hundred milliseconds, so the threads of an OpenMP
// runtime are asleep at every `parallel for` of the call and each region pays a futex wake-up; here the workers are
// woken once per call, and a loop never waits for a sleeper: chunks are handed out by an atomic counter and the calling
// thread takes chunks itself.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace svs {

class SpinPool {
 public:
  explicit SpinPool(int nthreads = 1) : want_(nthreads > 1 ? nthreads - 1 : 0) {}
  ~SpinPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
      active_.store(0);
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void set_threads(int nthreads) { if (th_.empty()) want_ = nthreads > 1 ? nthreads - 1 : 0; }
  int threads() const { return want_ + 1; }
  // wake the workers (they spin until end())
  void begin() {
    if (want_ == 0) return;
    {
      std::lock_guard<std::mutex> lk(m_);
      if (th_.empty())
        for (int i = 0; i < want_; ++i) th_.emplace_back([this] { worker(); });
      active_.store(1, std::memory_order_release);
    }
    cv_.notify_all();
  }
  void end() { active_.store(0, std::memory_order_release); }
  // fn(i) for i in [0, n), chunks handed out dynamically; returns when all have run
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    if (want_ == 0 || n == 1 || !active_.load(std::memory_order_acquire)) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    // the job fields are written only while the job is closed and no worker is inside work()
    fn_ = &fn;
    n_ = n;
    done_.store(0, std::memory_order_relaxed);
    next_.store(0, std::memory_order_relaxed);
    open_.store(1, std::memory_order_release);
    gen_.fetch_add(1, std::memory_order_release);
    work();
    while (done_.load(std::memory_order_acquire) < n) relax();
    open_.store(0, std::memory_order_release);
    while (inside_.load(std::memory_order_acquire) != 0) relax();
  }

 private:
  static void relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  void work() {
    for (;;) {
      const int i = next_.fetch_add(1, std::memory_order_acq_rel);
      if (i >= n_) return;
      (*fn_)(i);
      done_.fetch_add(1, std::memory_order_release);
    }
  }
  void worker() {
    unsigned seen = gen_.load(std::memory_order_acquire);
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return quit_ || active_.load(std::memory_order_acquire); });
        if (quit_) return;
      }
      while (active_.load(std::memory_order_acquire)) {
        const unsigned g = gen_.load(std::memory_order_acquire);
        if (g != seen) {
          seen = g;
          inside_.fetch_add(1, std::memory_order_acq_rel);
          if (open_.load(std::memory_order_acquire)) work();
          inside_.fetch_sub(1, std::memory_order_release);
        } else {
          relax();
        }
      }
    }
  }
  int want_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  bool quit_ = false;
  std::atomic<int> active_{0};
  std::atomic<unsigned> gen_{0};
  const std::function<void(int)>* fn_ = nullptr;
  int n_ = 0;
  std::atomic<int> next_{0}, done_{0}, open_{0}, inside_{0};
};

}  // namespace svs

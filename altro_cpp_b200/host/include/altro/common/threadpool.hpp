// altro/common/threadpool.hpp (B200 host mirror) — a small task pool with the reference's interface
// (altro/common/threadpool.hpp:25 there: LaunchThreads / AddTask / Wait / StopThreads).  The solver
// does not use it: the reference's pool spreads UpdateExpansions over knot ranges (ilqr.hpp:354-365
// there), which on the device is the (instance, knot) grid of k_update_expansions.  It is kept for
// programs that use the pool directly (perf/benchmark_threadpool.cpp).
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <future>
#include <mutex>
#include <thread>
#include <vector>

#include "altro/utils/assert.hpp"

namespace altro {

class ThreadPool {
 public:
  ThreadPool() = default;
  ThreadPool(const ThreadPool&) = delete;
  ThreadPool& operator=(const ThreadPool&) = delete;
  // a stopped pool can be moved: its pending tasks (and what Wait() will wait for) move with it
  ThreadPool(ThreadPool&& other) noexcept { TakeFrom(other); }
  ThreadPool& operator=(ThreadPool&& other) noexcept {
    if (this != &other) {
      if (IsRunning()) StopThreads();
      TakeFrom(other);
    }
    return *this;
  }
  ~ThreadPool() {
    if (IsRunning()) StopThreads();
  }

  template <class Task>
  void AddTask(const Task& task) {
    std::packaged_task<void()> ptask(task);
    futures_.emplace_back(ptask.get_future());
    {
      std::lock_guard<std::mutex> lock(mutex_);
      queue_.emplace_back(std::move(ptask));
    }
    cv_.notify_one();
  }
  size_t NumTasks() const {
    std::lock_guard<std::mutex> lock(mutex_);
    return queue_.size();
  }
  size_t NumThreads() const { return threads_.size(); }
  bool IsRunning() const { return running_; }

  // blocks until every task added so far has run
  void Wait() {
    for (std::future<void>& f : futures_) f.get();
    futures_.clear();
  }
  void LaunchThreads(int nthreads) {
    ALTRO_ASSERT(!IsRunning(), "Thread pool is already running.");
    running_ = true;
    for (int i = 0; i < nthreads; ++i) threads_.emplace_back([this]() { Work(); });
  }
  void StopThreads() {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      running_ = false;
    }
    cv_.notify_all();
    for (std::thread& t : threads_)
      if (t.joinable()) t.join();
    threads_.clear();
  }
  template <class Rep, class Period>
  void SetTimeoutPerTask(std::chrono::duration<Rep, Period>) {}

 private:
  void TakeFrom(ThreadPool& other) {
    ALTRO_ASSERT(!other.IsRunning(), "Stop the threads of a pool before moving it.");
    std::lock_guard<std::mutex> lock(other.mutex_);
    queue_ = std::move(other.queue_);
    futures_ = std::move(other.futures_);
    other.queue_.clear();
    other.futures_.clear();
  }
  void Work() {
    for (;;) {
      std::packaged_task<void()> task;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        cv_.wait(lock, [this]() { return !running_ || !queue_.empty(); });
        if (queue_.empty()) return;  // stopped and drained
        task = std::move(queue_.front());
        queue_.pop_front();
      }
      task();
    }
  }
  std::atomic_bool running_{false};
  std::vector<std::thread> threads_;
  std::vector<std::future<void>> futures_;
  std::deque<std::packaged_task<void()>> queue_;
  mutable std::mutex mutex_;
  std::condition_variable cv_;
};

}  // namespace altro

// altro/common/threadsafe_queue.hpp (B200 host mirror) — a FIFO several threads may push to and pop from,
// with the interface of the reference's class of this name (Push / TryPop / Size / IsEmpty / Clear,
// movable).  One mutex around a std::deque: the reference's pool spends microseconds per task, the device
// path does not use the queue at all (the batch axis replaces the thread pool), so nothing finer-grained
// is warranted here.
#pragma once

#include <cstddef>
#include <deque>
#include <memory>
#include <mutex>
#include <utility>

namespace altro {

template <class T>
class ThreadSafeQueue {
 public:
  ThreadSafeQueue() : guard_(new std::mutex) {}
  ThreadSafeQueue(const ThreadSafeQueue&) = delete;
  ThreadSafeQueue& operator=(const ThreadSafeQueue&) = delete;
  // moving is for queues nobody else is using at that moment (like the reference's)
  ThreadSafeQueue(ThreadSafeQueue&& other) noexcept : guard_(new std::mutex) {
    std::lock_guard<std::mutex> lock(*other.guard_);
    items_ = std::move(other.items_);
    other.items_.clear();
  }
  ThreadSafeQueue& operator=(ThreadSafeQueue&& other) noexcept {
    if (this != &other) {
      std::lock(*guard_, *other.guard_);
      std::lock_guard<std::mutex> mine(*guard_, std::adopt_lock);
      std::lock_guard<std::mutex> theirs(*other.guard_, std::adopt_lock);
      items_ = std::move(other.items_);
      other.items_.clear();
    }
    return *this;
  }

  void Push(T value) {
    std::lock_guard<std::mutex> lock(*guard_);
    items_.emplace_back(std::move(value));
  }
  // false (and `value` untouched) when there is nothing to pop
  bool TryPop(T& value) {
    std::lock_guard<std::mutex> lock(*guard_);
    if (items_.empty()) return false;
    value = std::move(items_.front());
    items_.pop_front();
    return true;
  }
  void Clear() {
    std::lock_guard<std::mutex> lock(*guard_);
    items_.clear();
  }
  bool IsEmpty() const {
    std::lock_guard<std::mutex> lock(*guard_);
    return items_.empty();
  }
  std::size_t Size() const {
    std::lock_guard<std::mutex> lock(*guard_);
    return items_.size();
  }

 private:
  std::unique_ptr<std::mutex> guard_;
  std::deque<T> items_;
};

}  // namespace altro

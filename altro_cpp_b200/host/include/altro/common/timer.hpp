// altro/common/timer.hpp (B200 host mirror) — the scoped wall-clock profiler of the reference
// (altro/common/timer.hpp:52,107 there): Timer::Start(name) returns a Stopwatch that adds its
// lifetime to "parent/name" when it goes out of scope; PrintSummary() lists the totals.  Host-side
// observability; device time is measured with CUDA events (bench.py).
#pragma once

#include <chrono>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace altro {

class Stopwatch;

class Timer : public std::enable_shared_from_this<Timer> {
  using microseconds = std::chrono::microseconds;

 public:
  ~Timer() {
    if (!printed_summary_ && active_ && !times_.empty()) PrintSummary();
    if (using_file_ && io_) std::fclose(io_);
  }
  static std::shared_ptr<Timer> MakeShared() { return std::shared_ptr<Timer>(new Timer()); }
  static std::shared_ptr<Timer> MakeUnique() { return std::shared_ptr<Timer>(new Timer()); }
  inline Stopwatch Start(const std::string& name);
  void PrintSummary() { PrintSummary(&times_); }
  void PrintSummary(std::map<std::string, microseconds>* times) {
    microseconds total(0);
    for (const auto& kv : *times)
      if (kv.first.find('/') == std::string::npos) total += kv.second;
    std::fprintf(io_, "%-40s %12s %8s\n", "Description", "Time (us)", "%Total");
    for (const auto& kv : *times)
      std::fprintf(io_, "%-40s %12lld %8.1f\n", kv.first.c_str(), static_cast<long long>(kv.second.count()),
                   total.count() ? 100.0 * static_cast<double>(kv.second.count()) / static_cast<double>(total.count()) : 0.0);
    std::fflush(io_);
    printed_summary_ = true;
  }
  void Activate() { active_ = true; }
  void Deactivate() { active_ = false; }
  bool IsActive() const { return active_; }
  void SetOutput(FILE* io) {  // the caller keeps ownership
    if (using_file_ && io_) std::fclose(io_);
    using_file_ = false;
    io_ = io;
  }
  void SetOutput(const std::string& filename) {  // the timer owns the file
    FILE* io = std::fopen(filename.c_str(), "w");
    if (io == nullptr) throw std::runtime_error("Error opening profiler file \"" + filename + "\"");
    SetOutput(io);
    using_file_ = true;
  }

 private:
  friend class Stopwatch;
  Timer() = default;
  std::vector<std::string> stack_;
  std::map<std::string, microseconds> times_;
  bool active_ = false;
  bool printed_summary_ = false;
  bool using_file_ = false;
  FILE* io_ = stdout;
};

using TimerPtr = std::shared_ptr<Timer>;

class Stopwatch {
 public:
  Stopwatch() = default;
  Stopwatch(Stopwatch&& o) noexcept : name_(std::move(o.name_)), start_(o.start_), parent_(std::move(o.parent_)) {}
  Stopwatch(const Stopwatch&) = delete;
  ~Stopwatch() {
    if (!parent_) return;
    const auto dt = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - start_);
    std::string key;
    for (const std::string& s : parent_->stack_) key += (key.empty() ? "" : "/") + s;
    parent_->times_[key] += dt;
    parent_->stack_.pop_back();
  }

 private:
  friend class Timer;
  Stopwatch(std::string name, std::shared_ptr<Timer> timer)
      : name_(std::move(name)), start_(std::chrono::high_resolution_clock::now()), parent_(std::move(timer)) {
    parent_->stack_.push_back(name_);
  }
  std::string name_;
  std::chrono::time_point<std::chrono::high_resolution_clock> start_;
  std::shared_ptr<Timer> parent_;
};

inline Stopwatch Timer::Start(const std::string& name) {
  if (!active_) return Stopwatch();
  return Stopwatch(name, shared_from_this());
}

}  // namespace altro

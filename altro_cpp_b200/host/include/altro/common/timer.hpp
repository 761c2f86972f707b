// altro/common/timer.hpp (B200 host mirror) — the scoped wall-clock profiler of the reference
// (altro/common/timer.hpp:52,107 there): Timer::Start(name) returns a Stopwatch that adds its
// lifetime to "parent/name" when it goes out of scope; PrintSummary() lists the totals.  Host-side
// observability; device time is measured with CUDA events (bench.py).
#pragma once

#include <algorithm>
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
  // One line per name stack, children indented under their parent, with the share of the whole run and of the
  // parent (altro/common/profile_entry.hpp).  The map's order puts "a/b" right after "a".
  inline void PrintSummary(std::map<std::string, microseconds>* times);
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

}  // namespace altro

#include "altro/common/profile_entry.hpp"

namespace altro {

inline void Timer::PrintSummary(std::map<std::string, microseconds>* times) {
  std::vector<ProfileEntry::Ptr> entries, open_levels;  // open_levels[l] = the last entry seen with l name parts
  entries.emplace_back(std::make_shared<ProfileEntry>("top", microseconds(0)));
  open_levels.emplace_back(entries.front());
  int width = 11;  // "Description"
  for (const auto& kv : *times) {
    ProfileEntry::Ptr entry = std::make_shared<ProfileEntry>(kv.first, kv.second);
    const std::size_t level = entry->NumLevels();
    if (level >= open_levels.size()) open_levels.resize(level + 1);
    open_levels[level] = entry;
    entry->parent = open_levels[level - 1] ? open_levels[level - 1] : entries.front();
    if (level == 1) entries.front()->time += entry->time;
    width = std::max(width, static_cast<int>(2 * (level - 1) + entry->name.back().size()));
    entries.emplace_back(std::move(entry));
  }
  width += 2;
  std::fprintf(io_, "%-*s  %9s  %7s  %7s\n", width, "Description", "Time (us)", "%Total", "%Parent");
  std::fprintf(io_, "%s\n", std::string(static_cast<std::size_t>(width) + 31, '-').c_str());
  for (std::size_t i = 1; i < entries.size(); ++i) {
    entries[i]->CalcStats();
    entries[i]->Print(io_, width);
  }
  std::fflush(io_);
  printed_summary_ = true;
}

inline Stopwatch Timer::Start(const std::string& name) {
  if (!active_) return Stopwatch();
  return Stopwatch(name, shared_from_this());
}

}  // namespace altro

// altro/common/solver_logger.hpp (B200 host mirror) — the console table of the reference's solvers
// (altro/common/solver_logger.hpp there): named columns (LogEntry, altro/common/log_entry.hpp), each with a
// verbosity level, a format and optional bounds; Log(title, value) fills the current row, Print() emits it, with
// the header repeated every `frequency` rows.  Host-side observability: a device solve does not print per iteration
// (the per-iteration numbers are in SolverStats afterwards).  Colours are accepted and ignored; formats are
// the "{:>.3e}"-style specs the reference uses, interpreted here without the fmt library.
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "altro/common/log_entry.hpp"

namespace altro {

class SolverLogger {
 public:
  explicit SolverLogger(LogLevel level = LogLevel::kSilent) : level_(level) {}

  LogLevel GetLevel() const { return level_; }
  void SetLevel(LogLevel level) { level_ = level; }
  void Disable() { level_ = LogLevel::kSilent; }
  void SetFrequency(int freq) { frequency_ = std::max(1, freq); }
  template <class Color>
  void SetHeaderColor(Color) {}
  int NumEntries() const { return static_cast<int>(entries_.size()); }

  // column < 0 counts from the end (-1 = last), like the reference
  LogEntry& AddEntry(int column, const std::string& title, const std::string& format = "{}",
                     LogEntry::EntryType type = LogEntry::kFloat) {
    const int n = NumEntries();
    int at = column < 0 ? n + 1 + column : column;
    at = std::min(std::max(at, 0), n);
    entries_.insert(entries_.begin() + at, LogEntry(title, format, type));
    return entries_[static_cast<std::size_t>(at)];
  }
  LogEntry* Find(const std::string& title) {
    for (LogEntry& e : entries_)
      if (e.GetTitle() == title || e.GetName() == title) return &e;
    return nullptr;
  }
  // the column of that title; asking for one that does not exist appends an empty column of that title (the
  // reference keeps its columns in a map and operator[] does the same)
  LogEntry& GetEntry(const std::string& title) {
    LogEntry* e = Find(title);
    return e != nullptr ? *e : AddEntry(-1, title);
  }
  using iterator = std::vector<LogEntry>::iterator;
  using const_iterator = std::vector<LogEntry>::const_iterator;
  iterator begin() { return entries_.begin(); }
  iterator end() { return entries_.end(); }
  const_iterator begin() const { return entries_.begin(); }
  const_iterator end() const { return entries_.end(); }
  // data logged to a column that is not shown at the current level (or does not exist) is discarded
  template <class T>
  void Log(const std::string& title, T value) {
    LogEntry* e = Find(title);
    if (e != nullptr && level_ != LogLevel::kSilent && e->IsActive(level_)) e->Log(value);
  }
  void PrintHeader() {
    if (level_ == LogLevel::kSilent) return;
    std::string line;
    for (const LogEntry& e : entries_)
      if (e.IsActive(level_)) line += e.HeaderCell() + " ";
    std::fprintf(out_, "%s\n%s\n", line.c_str(), std::string(line.size(), '-').c_str());
    rows_since_header_ = 0;
  }
  // the current row without the header logic
  void PrintData() {
    if (level_ == LogLevel::kSilent) return;
    std::string line;
    for (const LogEntry& e : entries_)
      if (e.IsActive(level_)) line += e.Cell() + " ";
    std::fprintf(out_, "%s\n", line.c_str());
    ++rows_since_header_;
  }
  void Print() {
    if (level_ == LogLevel::kSilent) return;
    if (rows_since_header_ < 0 || rows_since_header_ >= frequency_) PrintHeader();
    PrintData();
  }
  void Clear() {
    for (LogEntry& e : entries_) e.Clear();
    rows_since_header_ = -1;
  }
  void SetOutput(FILE* out) { out_ = out; }

 private:
  LogLevel level_;
  int frequency_ = 10;
  int rows_since_header_ = -1;
  std::vector<LogEntry> entries_;
  FILE* out_ = stdout;
};

}  // namespace altro

// altro/common/solver_logger.hpp (B200 host mirror) — the console table of the reference's solvers
// (altro/common/solver_logger.hpp and log_entry.hpp there): named columns, each with a verbosity level, a
// format and optional bounds; Log(title, value) fills the current row, Print() emits it, with the header
// repeated every `frequency` rows.  Host-side observability: a device solve does not print per iteration
// (the per-iteration numbers are in SolverStats afterwards).  Colours are accepted and ignored; formats are
// the "{:>.3e}"-style specs the reference uses, interpreted here without the fmt library.
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace altro {

enum class LogLevel { kSilent = 0, kOuter = 1, kOuterDebug = 2, kInner = 3, kInnerDebug = 4, kDebug = 5 };

class LogEntry {
 public:
  enum EntryType { kInt, kFloat, kString };
  LogEntry() = default;
  LogEntry(std::string title, std::string format, EntryType type = kFloat)
      : title_(std::move(title)), format_(std::move(format)), type_(type) {}

  LogEntry& SetWidth(int width) { width_ = width; return *this; }
  LogEntry& SetLevel(LogLevel level) { level_ = level; return *this; }
  LogEntry& SetType(EntryType type) { type_ = type; return *this; }
  template <class Color>
  LogEntry& SetLowerBound(double bound, Color) { lower_ = bound; has_lower_ = true; return *this; }
  LogEntry& SetLowerBound(double bound) { lower_ = bound; has_lower_ = true; return *this; }
  template <class Color>
  LogEntry& SetUpperBound(double bound, Color) { upper_ = bound; has_upper_ = true; return *this; }
  LogEntry& SetUpperBound(double bound) { upper_ = bound; has_upper_ = true; return *this; }

  const std::string& GetTitle() const { return title_; }
  const std::string& GetFormat() const { return format_; }
  int GetWidth() const { return width_; }
  LogLevel GetLevel() const { return level_; }
  EntryType GetType() const { return type_; }
  bool IsActive(LogLevel level) const { return level >= level_; }
  bool OutOfBounds() const { return has_value_ && ((has_lower_ && value_ < lower_) || (has_upper_ && value_ > upper_)); }

  template <class T>
  void Log(T value) {
    value_ = static_cast<double>(value);
    has_value_ = true;
    text_ = Render(value_);
  }
  void Log(const std::string& value) {
    has_value_ = false;
    text_ = value;
  }
  void Clear() { text_.clear(); has_value_ = false; }
  // the cell, right-aligned in the column
  std::string Cell() const { return Pad(text_); }
  std::string HeaderCell() const { return Pad(title_); }

 private:
  std::string Pad(const std::string& s) const {
    const int w = std::max(width_, static_cast<int>(s.size()));
    return std::string(static_cast<std::size_t>(w) - s.size(), ' ') + s;
  }
  // "{:>8.3e}" and relatives: [[fill]align][width][.precision][type]; anything unparsable prints with %g
  std::string Render(double v) const {
    int precision = -1;
    char kind = type_ == kInt ? 'd' : 'g';
    const std::size_t colon = format_.find(':'), close = format_.rfind('}');
    if (colon != std::string::npos && close != std::string::npos && close > colon) {
      const std::string spec = format_.substr(colon + 1, close - colon - 1);
      const std::size_t dot = spec.find('.');
      if (dot != std::string::npos) precision = std::atoi(spec.c_str() + dot + 1);
      if (!spec.empty() && std::string("defgxEG").find(spec.back()) != std::string::npos) kind = spec.back();
    }
    char buf[64];
    if (kind == 'd' || kind == 'x') {
      std::snprintf(buf, sizeof(buf), kind == 'd' ? "%lld" : "%llx", static_cast<long long>(v));
    } else {
      const char fmt[5] = {'%', '.', '*', kind, '\0'};
      std::snprintf(buf, sizeof(buf), fmt, precision < 0 ? 6 : precision, v);
    }
    return buf;
  }
  std::string title_, format_ = "{}", text_;
  EntryType type_ = kFloat;
  LogLevel level_ = LogLevel::kInner;
  int width_ = 10;
  double value_ = 0.0, lower_ = 0.0, upper_ = 0.0;
  bool has_value_ = false, has_lower_ = false, has_upper_ = false;
};

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
      if (e.GetTitle() == title) return &e;
    return nullptr;
  }
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
  void Print() {
    if (level_ == LogLevel::kSilent) return;
    if (rows_since_header_ < 0 || rows_since_header_ >= frequency_) PrintHeader();
    std::string line;
    for (const LogEntry& e : entries_)
      if (e.IsActive(level_)) line += e.Cell() + " ";
    std::fprintf(out_, "%s\n", line.c_str());
    ++rows_since_header_;
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

// altro/common/log_entry.hpp (B200 host mirror) — LogLevel and LogEntry: one column of the solver's console table
// (altro/common/log_entry.hpp there): title, verbosity level, format, width and optional bounds, plus the text of
// the current row's cell.  Formats are the "{:>.3e}"-style specs the reference uses, interpreted here without the
// fmt library; colours are accepted and ignored.
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <utility>

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
  LogEntry& SetName(const std::string& name) { name_ = name; return *this; }  // key in the logger; the title is what prints
  const std::string& GetName() const { return name_.empty() ? title_ : name_; }
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
  // one cell / one header cell on stdout, when the column is shown at `level`
  void Print(LogLevel level = LogLevel::kSilent) const {
    if (IsActive(level)) std::fputs(Cell().c_str(), stdout);
  }
  template <class Color>
  void PrintHeader(LogLevel level, Color) const { PrintHeader(level); }
  void PrintHeader(LogLevel level = LogLevel::kSilent) const {
    if (IsActive(level)) std::fputs(HeaderCell().c_str(), stdout);
  }
  template <class T, class = typename std::enable_if<!std::is_same<T, LogLevel>::value>::type>
  void Print(T value) {
    Log(value);
    Print();
  }
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
  std::string title_, name_, format_ = "{}", text_;
  EntryType type_ = kFloat;
  LogLevel level_ = LogLevel::kInner;
  int width_ = 10;
  double value_ = 0.0, lower_ = 0.0, upper_ = 0.0;
  bool has_value_ = false, has_lower_ = false, has_upper_ = false;
};

}  // namespace altro

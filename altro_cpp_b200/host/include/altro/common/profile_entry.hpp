// altro/common/profile_entry.hpp (B200 host mirror) — one line of the timer's summary: a '/'-separated name
// stack ("al/ilqr/cost"), the time spent under it, and its share of the whole run and of its parent.  Same
// public members as the reference's struct of this name; host-side observability only.
#pragma once

#include <chrono>
#include <cstddef>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

namespace altro {

struct ProfileEntry : public std::enable_shared_from_this<ProfileEntry> {
  using time_t = std::chrono::microseconds;
  using Ptr = std::shared_ptr<ProfileEntry>;

  ProfileEntry(const std::string& fullname, time_t time_spent) : time(time_spent), percent_total(0), percent_parent(0) {
    std::size_t begin = 0;
    while (begin <= fullname.size()) {
      const std::size_t slash = fullname.find('/', begin);
      const std::size_t end = slash == std::string::npos ? fullname.size() : slash;
      name.emplace_back(fullname.substr(begin, end - begin));
      if (slash == std::string::npos) break;
      begin = slash + 1;
    }
  }

  std::vector<std::string> name;  // "al/ilqr/cost" -> {"al", "ilqr", "cost"}
  time_t time;                    // total time under this name
  int percent_total;              // of the root's time (whole percent, rounded down)
  int percent_parent;             // of the parent's time
  Ptr parent = nullptr;

  std::size_t NumLevels() const { return name.size(); }
  // the ancestor without a parent: it holds the total recorded time
  Ptr GetRoot() {
    Ptr at = shared_from_this();
    while (at->parent) at = at->parent;
    return at;
  }
  void CalcStats() {
    const long long mine = time.count();
    const long long whole = GetRoot()->time.count();
    const long long above = parent ? parent->time.count() : whole;
    percent_total = whole > 0 ? static_cast<int>(100 * mine / whole) : 0;
    percent_parent = above > 0 ? static_cast<int>(100 * mine / above) : 0;
  }
  // "<indent><last name>  <time us>  <%total>  <%parent>", the description padded to `width`
  void Print(FILE* io, int width) {
    const std::string label = std::string(2 * (NumLevels() > 0 ? NumLevels() - 1 : 0), ' ') + (name.empty() ? std::string() : name.back());
    std::fprintf(io, "%-*s  %9lld  %7d  %7d\n", width, label.c_str(), static_cast<long long>(time.count()), percent_total, percent_parent);
  }
  void Print(int width) { Print(stdout, width); }
};

}  // namespace altro

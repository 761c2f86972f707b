// altro/utils/benchmarking.hpp (B200 host mirror) — Benchmark<Duration>(f, samples): run a callable repeatedly on
// the host and summarise its wall-clock times (altro/utils/benchmarking.hpp:21-113 there).  Host-side convenience
// for user functors; device timings come from CUDA events (bench.py, altro_b200_solver_last_solve_ms).
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <ratio>
#include <type_traits>
#include <vector>

// programs written against the reference get fmt through this header (it prints with fmt there)
#if defined(__has_include)
#if __has_include(<fmt/format.h>)
#include <fmt/format.h>
#endif
#endif

namespace altro {
namespace utils {

namespace detail {
template <class Period>
inline const char* PeriodSuffix() {
  if (std::is_same<Period, std::nano>::value) return "ns";
  if (std::is_same<Period, std::micro>::value) return "us";
  if (std::is_same<Period, std::milli>::value) return "ms";
  if (std::is_same<Period, std::ratio<1>>::value) return "s";
  return "ticks";
}
}  // namespace detail

template <class Duration>
struct BenchmarkResults {
  // statistics are kept in floating point, in the unit of `Duration`
  using time_t = std::chrono::duration<double, typename Duration::period>;
  time_t mean;
  time_t median;
  time_t std;  // population standard deviation
  time_t max;
  time_t min;
  int samples;

  // sorts `times` in place
  static BenchmarkResults Calculate(std::vector<time_t>& times) {
    BenchmarkResults res{};
    res.samples = static_cast<int>(times.size());
    if (times.empty()) return res;
    std::sort(times.begin(), times.end());
    const std::size_t count = times.size(), mid = count / 2;
    res.min = times.front();
    res.max = times.back();
    res.median = (count % 2) ? times[mid] : (times[mid - 1] + times[mid]) / 2.0;
    double sum = 0.0;
    for (const time_t& t : times) sum += t.count();
    const double avg = sum / static_cast<double>(count);
    double spread = 0.0;
    for (const time_t& t : times) spread += (t.count() - avg) * (t.count() - avg);
    res.mean = time_t(avg);
    res.std = time_t(std::sqrt(spread / static_cast<double>(count)));
    return res;
  }

  void Print() {
    const char* unit = detail::PeriodSuffix<typename Duration::period>();
    std::printf("Mean:    %g%s\n", mean.count(), unit);
    std::printf("Median:  %g%s\n", median.count(), unit);
    std::printf("Std:     %g%s\n", std.count(), unit);
    std::printf("Max:     %g%s\n", max.count(), unit);
    std::printf("Min:     %g%s\n", min.count(), unit);
    std::printf("Samples: %d\n", samples);
  }
};

static constexpr int kDefaultSamples = 100;

// calls f() Nsamples times, timing each call
template <class Duration, class Function>
BenchmarkResults<Duration> Benchmark(Function f, int Nsamples = kDefaultSamples) {
  using clock = std::chrono::high_resolution_clock;
  using time_t = typename BenchmarkResults<Duration>::time_t;
  std::vector<time_t> times(static_cast<std::size_t>(std::max(Nsamples, 0)));
  for (time_t& slot : times) {
    const clock::time_point before = clock::now();
    f();
    slot = std::chrono::duration_cast<time_t>(clock::now() - before);
  }
  return BenchmarkResults<Duration>::Calculate(times);
}

}  // namespace utils
}  // namespace altro

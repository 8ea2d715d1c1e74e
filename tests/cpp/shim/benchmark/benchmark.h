// SPDX-License-Identifier: Apache-2.0
// Test infrastructure: the handful of Google Benchmark names the reference's tests/bench-ntt.cpp uses, so that the
// UNMODIFIED reference driver (compiled from /root/reference where it lies) builds against the drop-in headers without
// the library.  Every registered benchmark runs its loop a fixed number of times and prints one line.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace benchmark {

class State {
 public:
  explicit State(std::uint64_t iterations) : left_(iterations), total_(iterations) {}
  struct Iterator {
    State* s;
    bool operator!=(const Iterator&) const { return s->KeepRunning(); }
    Iterator& operator++() { return *this; }
    int operator*() const { return 0; }
  };
  Iterator begin() { return {this}; }
  Iterator end() { return {this}; }
  bool KeepRunning() {
    if (!started_) {
      started_ = true;
      t0_ = std::chrono::steady_clock::now();
    }
    if (left_ == 0) {
      t1_ = std::chrono::steady_clock::now();
      return false;
    }
    --left_;
    return true;
  }
  std::uint64_t iterations() const { return total_; }
  void SetItemsProcessed(std::uint64_t n) { items_ = n; }
  std::map<std::string, double> counters;
  double seconds() const { return std::chrono::duration<double>(t1_ - t0_).count(); }

 private:
  std::uint64_t left_, total_, items_ = 0;
  bool started_ = false;
  std::chrono::steady_clock::time_point t0_, t1_;
};

template <class T>
inline void DoNotOptimize(T& v) {
  asm volatile("" : "+m"(v) : : "memory");
}
template <class T>
inline void DoNotOptimize(const T& v) {
  asm volatile("" : : "m"(v) : "memory");
}
inline void ClobberMemory() { asm volatile("" : : : "memory"); }

namespace detail {
inline std::vector<std::pair<std::string, std::function<void(State&)>>>& registry() {
  static std::vector<std::pair<std::string, std::function<void(State&)>>> r;
  return r;
}
inline std::uint64_t& iterations() {
  static std::uint64_t n = 3;
  return n;
}
}  // namespace detail

template <class F>
inline void RegisterBenchmark(const std::string& name, F&& fn) {
  detail::registry().emplace_back(name, std::function<void(State&)>(std::forward<F>(fn)));
}
inline void Initialize(int* argc, char** argv) {
  // --iterations=N is the only option; everything recognised is removed from argv like the library does
  int out = 1;
  for (int i = 1; i < *argc; ++i) {
    const std::string a = argv[i];
    if (a.rfind("--iterations=", 0) == 0)
      detail::iterations() = std::stoull(a.substr(13));
    else
      argv[out++] = argv[i];
  }
  *argc = out;
}
inline bool ReportUnrecognizedArguments(int argc, char** argv) {
  for (int i = 1; i < argc; ++i) std::fprintf(stderr, "unrecognised argument: %s\n", argv[i]);
  return argc > 1;
}
inline std::size_t RunSpecifiedBenchmarks() {
  for (auto& [name, fn] : detail::registry()) {
    State st(detail::iterations());
    fn(st);  // throws on a mismatch (bench-ntt.cpp:60-64)
    std::printf("%s: %llu iterations, %.3f ms each, m = %.0f: ok\n", name.c_str(), (unsigned long long)st.iterations(),
                st.seconds() * 1e3 / (double)st.iterations(), st.counters.count("m") ? st.counters["m"] : 0.0);
  }
  return detail::registry().size();
}
inline void Shutdown() {}

}  // namespace benchmark

// SPDX-License-Identifier: Apache-2.0
// Test infrastructure: the few GoogleTest names the reference's tests/test-modulus.cpp uses (TEST, EXPECT_EQ, ASSERT_GT
// with streamed messages), so that the UNMODIFIED reference test builds against the drop-in headers without the library.
// main() is part of this header (the reference links gtest_main).
#pragma once
#include <cstdio>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace testing {
namespace shim {
struct Case {
  std::string name;
  std::function<void()> body;
};
inline std::vector<Case>& cases() {
  static std::vector<Case> c;
  return c;
}
inline int& failures() {
  static int f = 0;
  return f;
}
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<void()> body) {
    cases().push_back({std::string(suite) + "." + name, std::move(body)});
  }
};
// collects the streamed message of a failed check and reports it when the full expression ends
class Failure {
 public:
  Failure(const char* file, int line, const std::string& what) { os_ << file << ":" << line << ": Failure\n" << what << "\n"; }
  Failure(const Failure& o) { os_ << o.os_.str(); }
  ~Failure() {
    ++failures();
    std::cerr << os_.str() << std::endl;
  }
  template <class T>
  Failure& operator<<(const T& v) {
    os_ << v;
    return *this;
  }

 private:
  std::ostringstream os_;
};
struct Void {
  void operator=(const Failure&) const {}
};
template <class A, class B>
std::string describe(const char* op, const char* ea, const char* eb, const A& a, const B& b) {
  std::ostringstream os;
  os << "Expected: (" << ea << ") " << op << " (" << eb << "), actual: " << a << " vs " << b;
  return os.str();
}
}  // namespace shim
inline void InitGoogleTest(int*, char**) {}
}  // namespace testing

#define GTEST_SHIM_CHECK_(a, b, op, opname, on_fail)                                                         \
  if (const auto& gtest_a_ = (a); true)                                                                     \
    if (const auto& gtest_b_ = (b); gtest_a_ op gtest_b_)                                                    \
      ;                                                                                                     \
    else                                                                                                    \
      on_fail ::testing::shim::Void{} = ::testing::shim::Failure(__FILE__, __LINE__,                          \
                                                               ::testing::shim::describe(opname, #a, #b, gtest_a_, gtest_b_))

#define EXPECT_EQ(a, b) GTEST_SHIM_CHECK_(a, b, ==, "==", )
#define EXPECT_NE(a, b) GTEST_SHIM_CHECK_(a, b, !=, "!=", )
#define EXPECT_GT(a, b) GTEST_SHIM_CHECK_(a, b, >, ">", )
#define EXPECT_LT(a, b) GTEST_SHIM_CHECK_(a, b, <, "<", )
#define ASSERT_EQ(a, b) GTEST_SHIM_CHECK_(a, b, ==, "==", return)
#define ASSERT_NE(a, b) GTEST_SHIM_CHECK_(a, b, !=, "!=", return)
#define ASSERT_GT(a, b) GTEST_SHIM_CHECK_(a, b, >, ">", return)
#define ASSERT_LT(a, b) GTEST_SHIM_CHECK_(a, b, <, "<", return)

#define TEST(suite, name)                                                                                   \
  static void gtest_shim_##suite##_##name();                                                                \
  static ::testing::shim::Registrar gtest_shim_reg_##suite##_##name(#suite, #name, gtest_shim_##suite##_##name); \
  static void gtest_shim_##suite##_##name()

inline int RUN_ALL_TESTS() {
  for (auto& c : ::testing::shim::cases()) {
    const int before = ::testing::shim::failures();
    std::printf("[ RUN      ] %s\n", c.name.c_str());
    c.body();
    std::printf("[ %s ] %s\n", ::testing::shim::failures() == before ? "      OK" : " FAILED ", c.name.c_str());
  }
  std::printf("%zu test(s), %d failure(s)\n", ::testing::shim::cases().size(), ::testing::shim::failures());
  return ::testing::shim::failures() ? 1 : 0;
}

#ifndef GTEST_SHIM_NO_MAIN
int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
#endif

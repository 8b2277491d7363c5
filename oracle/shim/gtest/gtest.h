// Minimal stand-in for googletest 1.7 (TEST, EXPECT_*, RUN_ALL_TESTS) so that the reference's own
// test helpers (test/uct/uct_test_class.h) compile and run without gtest. TEST INFRASTRUCTURE ONLY.
#ifndef MAMCTS_ORACLE_SHIM_GTEST_H_
#define MAMCTS_ORACLE_SHIM_GTEST_H_
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
struct Registry {
  struct Case { std::string name; std::function<void()> fn; };
  std::vector<Case> cases;
  long checks = 0, failures = 0;
  static Registry& get() { static Registry r; return r; }
};
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<void()> fn) {
    Registry::get().cases.push_back({std::string(suite) + "." + name, fn});
  }
};
// Collects the streamed message; prints it when the expectation failed.
struct Expectation {
  bool ok;
  std::ostringstream msg;
  Expectation(bool o, const char* expr, const char* file, int line) : ok(o) {
    Registry::get().checks++;
    if (!ok) { Registry::get().failures++; msg << file << ":" << line << ": expectation failed: " << expr << " "; }
  }
  template <class T> Expectation& operator<<(const T& v) { if (!ok) msg << v; return *this; }
  Expectation& operator<<(std::ostream& (*f)(std::ostream&)) { if (!ok) msg << f; return *this; }
  ~Expectation() { if (!ok) std::cerr << msg.str() << std::endl; }
};
inline void InitGoogleTest(int*, char**) {}
}  // namespace testing

#define TEST(suite, name)                                                              \
  static void suite##_##name##_body();                                                 \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body); \
  static void suite##_##name##_body()
#define EXPECT_TRUE(c) ::testing::Expectation(static_cast<bool>(c), #c, __FILE__, __LINE__)
#define EXPECT_FALSE(c) ::testing::Expectation(!static_cast<bool>(c), "!(" #c ")", __FILE__, __LINE__)
#define EXPECT_EQ(a, b) ::testing::Expectation((a) == (b), #a " == " #b, __FILE__, __LINE__)
#define EXPECT_NE(a, b) ::testing::Expectation((a) != (b), #a " != " #b, __FILE__, __LINE__)
#define EXPECT_NEAR(a, b, tol) ::testing::Expectation(std::fabs((a) - (b)) <= (tol), #a " ~ " #b, __FILE__, __LINE__)
#define ASSERT_TRUE(c) EXPECT_TRUE(c)
#define ASSERT_EQ(a, b) EXPECT_EQ(a, b)

inline int RUN_ALL_TESTS() {
  auto& r = ::testing::Registry::get();
  for (auto& c : r.cases) {
    const long before = r.failures;
    c.fn();
    std::printf("%s %s\n", r.failures == before ? "PASS" : "FAIL", c.name.c_str());
  }
  std::printf("%ld checks, %ld failures\n", r.checks, r.failures);
  return r.failures ? 1 : 0;
}
#endif

/* gtest/gtest.h of the ROS stand-in: enough of googletest to run the reference's own unit test
 * (tests/testUtils.cpp) unmodified -- TEST, ASSERT_EQ / ASSERT_LE / ..., InitGoogleTest, RUN_ALL_TESTS.
 * TEST INFRASTRUCTURE. */
#ifndef ROS_STUB_GTEST_H
#define ROS_STUB_GTEST_H
#include <cstdio>
#include <functional>
#include <string>
#include <vector>
namespace testing {
struct Registry {
  struct Case {
    std::string name;
    std::function<void(bool &)> run;
  };
  std::vector<Case> cases;
  static Registry &get() {
    static Registry r;
    return r;
  }
};
struct Registrar {
  Registrar(const char *suite, const char *name, std::function<void(bool &)> fn) {
    Registry::get().cases.push_back({std::string(suite) + "." + name, std::move(fn)});
  }
};
inline void InitGoogleTest(int *, char **) {}
inline int RunAll() {
  int failed = 0;
  for (auto &c : Registry::get().cases) {
    bool ok = true;
    std::printf("[ RUN      ] %s\n", c.name.c_str());
    c.run(ok);
    std::printf("%s %s\n", ok ? "[       OK ]" : "[  FAILED  ]", c.name.c_str());
    failed += !ok;
  }
  std::printf("[==========] %zu tests ran, %d failed.\n", Registry::get().cases.size(), failed);
  return failed ? 1 : 0;
}
}  // namespace testing
#define RUN_ALL_TESTS() ::testing::RunAll()
#define TEST(suite, name)                                                                      \
  static void suite##_##name##_body(bool &gtest_ok_);                                          \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);      \
  static void suite##_##name##_body(bool &gtest_ok_)
#define GTEST_STUB_ASSERT_(expr, text)                                              \
  do {                                                                              \
    if (!(expr)) {                                                                  \
      std::printf("%s:%d: Failure\n  %s\n", __FILE__, __LINE__, text);              \
      gtest_ok_ = false;                                                            \
      return;                                                                       \
    }                                                                               \
  } while (0)
#define ASSERT_TRUE(a) GTEST_STUB_ASSERT_((a), #a)
#define ASSERT_FALSE(a) GTEST_STUB_ASSERT_(!(a), "!(" #a ")")
#define ASSERT_EQ(a, b) GTEST_STUB_ASSERT_((a) == (b), #a " == " #b)
#define ASSERT_NE(a, b) GTEST_STUB_ASSERT_((a) != (b), #a " != " #b)
#define ASSERT_LE(a, b) GTEST_STUB_ASSERT_((a) <= (b), #a " <= " #b)
#define ASSERT_LT(a, b) GTEST_STUB_ASSERT_((a) < (b), #a " < " #b)
#define ASSERT_GE(a, b) GTEST_STUB_ASSERT_((a) >= (b), #a " >= " #b)
#define ASSERT_GT(a, b) GTEST_STUB_ASSERT_((a) > (b), #a " > " #b)
#define ASSERT_NEAR(a, b, tol) GTEST_STUB_ASSERT_(((a) - (b)) <= (tol) && ((b) - (a)) <= (tol), #a " ~ " #b)
#define EXPECT_TRUE ASSERT_TRUE
#define EXPECT_EQ ASSERT_EQ
#define EXPECT_LE ASSERT_LE
#endif

// Stand-in for googletest's gtest_prod.h (FRIEND_TEST) so that the reference's own tests compile without
// gtest. TEST INFRASTRUCTURE ONLY. The friend declaration names the function the TEST macro of
// shim/gtest/gtest.h defines.
#ifndef MAMCTS_ORACLE_SHIM_GTEST_PROD_H_
#define MAMCTS_ORACLE_SHIM_GTEST_PROD_H_
#define FRIEND_TEST(suite, name) friend void suite##_##name##_body()
#endif

// TEST INFRASTRUCTURE ONLY.  Stand-in for openfst's glog-style "fst/log.h", which the reference's runtime headers pull in
// through utils/log.h (runtime/core/utils/log.h:20) and which is not installed here: just the CHECK macros that
// frontend/fbank.h uses, so that the reference's own front-end sources compile unmodified into oracle/_ref/fbank_ref.
#ifndef ORACLE_STUB_FST_LOG_H_
#define ORACLE_STUB_FST_LOG_H_
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#define ORACLE_CHECK_(cond)                                                          \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      std::abort();                                                                  \
    }                                                                                \
  } while (0)
#define CHECK(cond) ORACLE_CHECK_(cond)
#define CHECK_GE(a, b) ORACLE_CHECK_((a) >= (b))
#define CHECK_GT(a, b) ORACLE_CHECK_((a) > (b))
#define CHECK_LE(a, b) ORACLE_CHECK_((a) <= (b))
#define CHECK_LT(a, b) ORACLE_CHECK_((a) < (b))
#define CHECK_EQ(a, b) ORACLE_CHECK_((a) == (b))
#endif  // ORACLE_STUB_FST_LOG_H_

// Minimal serial stand-in for <oneapi/tbb/blocked_range.h>.
// TEST INFRASTRUCTURE ONLY: oneTBB is not installed in this image, and the only
// thing the reference's CPU ExSUM needs from it is a range type with
// begin()/end()/grainsize() and a `split` tag (reference use:
// src/cpu/blas/blas1/ExSUM.hpp:46-56, ExSUM.cpp:131-140).  Written from scratch.
#ifndef EXBLAS_B200_ORACLE_TBB_BLOCKED_RANGE_H
#define EXBLAS_B200_ORACLE_TBB_BLOCKED_RANGE_H
#include <cstddef>

namespace oneapi { namespace tbb {

struct split {};

template <typename Value>
class blocked_range {
    Value lo_, hi_;
    std::size_t grain_;
public:
    blocked_range(Value lo, Value hi, std::size_t grain = 1) : lo_(lo), hi_(hi), grain_(grain) {}
    Value begin() const { return lo_; }
    Value end() const { return hi_; }
    std::size_t grainsize() const { return grain_; }
};

} }  // namespace oneapi::tbb

namespace tbb = oneapi::tbb;
#endif

// Minimal serial stand-in for <oneapi/tbb/parallel_reduce.h> (+ task_arena).
// TEST INFRASTRUCTURE ONLY.  parallel_reduce applies the body to the whole
// range on the calling thread; that is a legal TBB schedule (no split happened)
// and the result of the reference's superaccumulator-only path (fpe < 2) does
// not depend on the schedule.  The FPE paths (fpe >= 2) use OpenMP, not TBB.
#ifndef EXBLAS_B200_ORACLE_TBB_PARALLEL_REDUCE_H
#define EXBLAS_B200_ORACLE_TBB_PARALLEL_REDUCE_H
#include "blocked_range.h"

namespace oneapi { namespace tbb {

template <typename Range, typename Body>
void parallel_reduce(const Range& range, Body& body) { body(range); }

template <typename Range, typename Body>
void parallel_deterministic_reduce(const Range& range, Body& body) { body(range); }

class task_arena {
public:
    template <typename F>
    auto execute(F&& f) -> decltype(f()) { return f(); }
};

} }  // namespace oneapi::tbb
#endif

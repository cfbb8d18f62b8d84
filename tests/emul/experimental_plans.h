// Test-only plan variants for the kernel emulator (tests/emul/emul.cpp): candidates for the next GPU tuning round that are
// not (yet) in kissfft_b200/csrc/kf_plan_list.h.  Same make_plan arguments as there.
#pragma once

namespace kf {
#if defined(FIXED_POINT)
// fixed point keeps the reference's radix order 4,4,4,4,4,2 (bit-exactness); paired last group of 8 = {4,2}... is not
// available in that order, so the lane permutation is exercised on the 16-point paired group
KF_PLAN(kX2048a, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 1, 4, 2, 1, 2, 0, 0, 1, 1);
// tools/bank_model.py: 2048 with logpad 5 and 1155 without skew have fewer shared-memory wavefronts in fixed point too
KF_PLAN(kX2048b, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 1, 5, 2, 1);
KF_PLAN(kX1155a, 1155, {3, 5, 7, 11},      {1, 1, 2}, 105, 2, 31, 1, 0);
#define KF_EXPERIMENTAL_LIST(X) X(kX2048a, REAL) X(kX2048b, C2C_REAL) X(kX1155a, C2C)
#else
// R2C: 8-point paired last group, lanes 0-15 take the even pairs (conflict-free mirrored loads), with and without the
// input stage doubling as exchange buffer; C2R with the same lane mapping
KF_PLAN(kX2048a, 2048, {4, 2, 4, 4, 4, 4}, {2, 2, 2}, 128, 1, 4, 4, 1, 1, 1, 0, 1, 1);
KF_PLAN(kX2048b, 2048, {2, 4, 4, 4, 4, 4}, {2, 2, 2}, 128, 2, 4, 2, 1, 2, 1, 0, 1, 1);
KF_PLAN(kX2048c, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 2, 4, 2, 1, 2, 1, 0, 1, 1);
#define KF_EXPERIMENTAL_LIST(X) X(kX2048a, R2C) X(kX2048b, R2C) X(kX2048c, C2R)
#endif
}   // namespace kf

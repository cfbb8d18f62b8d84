// Test-only plan variants for the kernel emulator (tests/emul/emul.cpp): candidates for GPU tuning that are not (yet) in
// kissfft_b200/csrc/kf_plan_list.h.  Same make_plan arguments as there (last one: PlanDesc::hoist).
#pragma once

namespace kf {
#if defined(FIXED_POINT)
// fixed point keeps the reference's radix order 4,4,4,4,4,2 (bit-exactness); hoisted tables are exact copies of the table
// entries, so the results stay bit-identical
KF_PLAN(kX2048a, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 1, 4, 2, 1, 2, 0, 0, 1, 7);
// tools/bank_model.py: 2048 with logpad 5 and 1155 without skew have fewer shared-memory wavefronts in fixed point too
KF_PLAN(kX2048b, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 1, 5, 2, 1, 2, 0, 0, 0, 2);
KF_PLAN(kX1155a, 1155, {3, 5, 7, 11},      {1, 1, 2}, 105, 2, 31, 1, 0);
#define KF_EXPERIMENTAL_LIST(X) X(kX2048a, REAL) X(kX2048b, C2C_REAL) X(kX1155a, C2C)
#else
// R2C with the 8-point paired last group and C2R with the paired first group, loop-invariant tables in registers:
// everything (7), split twiddles only (1), stage twiddles only (6); two transforms per CTA; ping-pong buffers
KF_PLAN(kX2048a, 2048, {4, 2, 4, 4, 4, 4}, {2, 2, 2}, 128, 1, 4, 4, 1, 1, 1, 0, 1, 7);
KF_PLAN(kX2048b, 2048, {2, 4, 4, 4, 4, 4}, {2, 2, 2}, 128, 2, 4, 2, 1, 2, 1, 0, 1, 1);
KF_PLAN(kX2048c, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 2, 4, 2, 1, 2, 1, 0, 1, 7);
KF_PLAN(kX2048d, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 1, 4, 4, 1, 1, 1, 0, 1, 6);
KF_PLAN(kX2048e, 2048, {4, 2, 4, 4, 4, 4}, {2, 2, 2}, 128, 1, 4, 4, 1, 1, 0, 0, 1, 7);
#define KF_EXPERIMENTAL_LIST(X) X(kX2048a, R2C) X(kX2048b, R2C) X(kX2048c, C2R) X(kX2048d, C2R) X(kX2048e, C2C_REAL)
#endif
}   // namespace kf

// kf_plan_list.h -- the compile-time plans of this build, one tag type per registered length.
//
// make_plan(N, {radices: outermost..innermost, exactly kf_factor's order (kiss_fft.c:306-328)},
//           {stages per register group, FIRST-EXECUTED (innermost) group first},
//           team (threads per transform), tpc (transforms per CTA), logpad (smem skew), min CTAs/SM)
//
// Host-compilable: tests/emul instantiates the same tags on the CPU.  The datatype is selected by the usual
// macros (FIXED_POINT / kiss_fft_scalar) because tile shapes depend on sizeof(kiss_fft_cpx).
#pragma once
#include <initializer_list>

#include "kf_plan.h"

namespace kf {

constexpr PlanDesc make_plan(int N, std::initializer_list<int> radices, std::initializer_list<int> groups, int team, int tpc,
                             int logpad, int minblocks, int nstage = 0)
{
    PlanDesc d{};
    d.N = N;
    d.L = 0;
    for (int r : radices) d.p[d.L++] = r;
    d.G = 0;
    for (int g : groups) d.glen[d.G++] = g;
    d.team = team;
    d.tpc = tpc;
    d.logpad = logpad;
    d.minblocks = minblocks;
    d.nstage = nstage;
    return d;
}

#if defined(FIXED_POINT) && (FIXED_POINT == 16)
// ---- Q15: 4-byte complex ------------------------------------------------------------------------------
struct kP16 { static constexpr PlanDesc D = make_plan(16,   {4, 4},             {2},       1,   128, 31, 1); };
struct kP64 { static constexpr PlanDesc D = make_plan(64,   {4, 4, 4},          {1, 2},    16,  16,  4,  1); };
struct kP256 { static constexpr PlanDesc D = make_plan(256,  {4, 4, 4, 4},       {2, 2},    16,  16,  4,  1, 2); };
struct kP1024 { static constexpr PlanDesc D = make_plan(1024, {4, 4, 4, 4, 4},    {2, 2, 1}, 64,  4,   4,  1, 3); };
struct kP2048 { static constexpr PlanDesc D = make_plan(2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 2,   4,  1, 2); };
struct kP1000 { static constexpr PlanDesc D = make_plan(1000, {4, 2, 5, 5, 5},    {2, 2, 1}, 50,  5,   4,  1, 2); };
struct kP1155 { static constexpr PlanDesc D = make_plan(1155, {3, 5, 7, 11},      {1, 1, 2}, 105, 2,   4,  1, 2); };
#elif defined(FIXED_POINT)
// ---- Q31: 8-byte complex ------------------------------------------------------------------------------
struct kP16 { static constexpr PlanDesc D = make_plan(16,   {4, 4},             {2},       1,   128, 31, 1); };
struct kP64 { static constexpr PlanDesc D = make_plan(64,   {4, 4, 4},          {1, 2},    16,  16,  4,  1); };
struct kP256 { static constexpr PlanDesc D = make_plan(256,  {4, 4, 4, 4},       {2, 2},    16,  16,  4,  1, 2); };
struct kP1024 { static constexpr PlanDesc D = make_plan(1024, {4, 4, 4, 4, 4},    {2, 2, 1}, 64,  4,   4,  1, 3); };
struct kP2048 { static constexpr PlanDesc D = make_plan(2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 2,   4,  1, 2); };
struct kP1000 { static constexpr PlanDesc D = make_plan(1000, {4, 2, 5, 5, 5},    {2, 2, 1}, 50,  5,   4,  1, 2); };
struct kP1155 { static constexpr PlanDesc D = make_plan(1155, {3, 5, 7, 11},      {1, 1, 2}, 105, 2,   4,  1, 2); };
#else
// ---- float (8-byte complex) and double (16-byte complex) ----------------------------------------------
static constexpr int kTs = (KF_SCALAR_BYTES == 8) ? 2 : 1;   // double: halve the transforms per CTA
struct kP16 { static constexpr PlanDesc D = make_plan(16,   {4, 4},             {2},       1,   128, 31, 1); };
struct kP64 { static constexpr PlanDesc D = make_plan(64,   {4, 4, 4},          {1, 2},    16,  16 / kTs, 4, 1); };
struct kP256 { static constexpr PlanDesc D = make_plan(256,  {4, 4, 4, 4},       {2, 2},    16,  16 / kTs, 4, 1, 2); };
struct kP1024 { static constexpr PlanDesc D = make_plan(1024, {4, 4, 4, 4, 4},    {2, 2, 1}, 64,  4 / kTs,  4, 1, 3); };
struct kP2048 { static constexpr PlanDesc D = make_plan(2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 2 / kTs,  4, 1, 2); };
struct kP1000 { static constexpr PlanDesc D = make_plan(1000, {4, 2, 5, 5, 5},    {2, 2, 1}, 50,  5,        4, 1, 2); };
struct kP1155 { static constexpr PlanDesc D = make_plan(1155, {3, 5, 7, 11},      {1, 1, 2}, 105, 2,        4, 1, 2); };
#endif


// X(tag, modes) with modes in {ALL, C2C, C2C_REAL}: which kernel modes are instantiated for the length
#define KF_PLAN_LIST(X) \
    X(kP16, C2C)        \
    X(kP64, ALL)        \
    X(kP256, ALL)       \
    X(kP1024, ALL)      \
    X(kP2048, ALL)      \
    X(kP1000, C2C_REAL) \
    X(kP1155, C2C)

}   // namespace kf

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
                             int logpad, int minblocks, int nstage = 0, int nbuf = 2, int twmode = 0, int shfl_post = 0)
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
    d.nbuf = nbuf;
    d.twmode = twmode;
    d.shfl_post = shfl_post;
    return d;
}

// Tuned on B200 with tools/tune_gen.py (logs under profiles/).  Fixed-point plans keep kf_factor's radix order
// (bit-exactness); float/double plans may reorder / regroup radices (parity there is relative RMS).
// Tuned on B200 with tools/tune_gen.py (logs under profiles/).  Fixed-point plans keep kf_factor's radix order
// (bit-exactness); float/double plans may reorder / regroup radices (parity there is relative RMS).
#define KF_PLAN(tag, ...) struct tag { static constexpr PlanDesc D = make_plan(__VA_ARGS__); }

// sizeof(kiss_fft_cpx) of this build, usable by the preprocessor
#if defined(FIXED_POINT) && (FIXED_POINT == 16)
#define KF_E 4
#elif defined(KF_IS_DOUBLE)
#define KF_E 16
#else
#define KF_E 8
#endif

// Power-of-two lengths 16..4096 in kf_factor's radix order, shared by all datatypes (KF_E = sizeof(kiss_fft_cpx)):
// "col" plans serve the kiss_fftnd axis pass and want tpc*KF_E >= 64..128-byte row segments.
#define KF_TPC(bytes) ((bytes) / KF_E > 0 ? (bytes) / KF_E : 1)
KF_PLAN(kP16,      16,   {4, 4},             {2},       1,   128,        31, 1, 0);
KF_PLAN(kP32,      32,   {4, 4, 2},          {2, 1},    8,   16,         4,  1, 0);
KF_PLAN(kP64,      64,   {4, 4, 4},          {1, 2},    16,  8,          4,  1, 0);
KF_PLAN(kP128,     128,  {4, 4, 4, 2},       {2, 2},    16,  8,          4,  1, 2);
KF_PLAN(kP256,     256,  {4, 4, 4, 4},       {2, 2},    16,  8,          4,  1, 2);
KF_PLAN(kP128col,  128,  {4, 4, 4, 2},       {2, 2},    16,  KF_TPC(128), 4, 1, 0, 1);
KF_PLAN(kP256col,  256,  {4, 4, 4, 4},       {2, 2},    16,  KF_TPC(128), 4, 1, 0, 1);
#if KF_E <= 8
KF_PLAN(kP512,     512,  {4, 4, 4, 4, 2},    {3, 2},    32,  4,          4,  2, 2);
KF_PLAN(kP512col,  512,  {4, 4, 4, 4, 2},    {3, 2},    32,  KF_TPC(128), 5, 1, 0, 1);
#else
KF_PLAN(kP512,     512,  {4, 4, 4, 4, 2},    {2, 2, 1}, 64,  2,          4,  2, 2);
KF_PLAN(kP512col,  512,  {4, 4, 4, 4, 2},    {2, 2, 1}, 64,  4,          4,  1, 0);
#endif
KF_PLAN(kP4096,    4096, {4, 4, 4, 4, 4, 4}, {2, 2, 2}, 256, 1,          4,  1, (KF_E <= 8 ? 2 : 0));
#define KF_POW2_LIST(X) X(kP16, C2C) X(kP32, C2C_REAL) X(kP64, ALL) X(kP128, C2C_REAL) X(kP128col, COL) X(kP256, C2C_REAL) \
    X(kP256col, COL) X(kP512, C2C_REAL) X(kP512col, COL) X(kP4096, C2C_REAL)

#if defined(FIXED_POINT) && (FIXED_POINT == 16)
// ---- Q15: 4-byte complex, integer-issue bound -------------------------------------------------------------
KF_PLAN(kP1024,    1024, {4, 4, 4, 4, 4},    {2, 2, 1}, 64,  2, 4, 4, 2, 2, 0, 1);
KF_PLAN(kP2048,    2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 1, 4, 2, 1);
KF_PLAN(kP1000,    1000, {4, 2, 5, 5, 5},    {2, 2, 1}, 50,  4, 4, 1, 2);
KF_PLAN(kP1155,    1155, {3, 5, 7, 11},      {1, 1, 2}, 105, 2, 4, 1, 0);
KF_PLAN(kP1024col, 1024, {4, 4, 4, 4, 4},    {2, 2, 1}, 64,  16, 4, 1, 0);
KF_PLAN(kP2048col, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 8, 4, 1, 0);
#define KF_PLAN_LIST(X) \
    KF_POW2_LIST(X) X(kP1024, C2C_REAL) X(kP1024col, COL) X(kP2048, C2C_REAL) X(kP2048col, COL) X(kP1000, C2C_REAL) X(kP1155, C2C)
#elif defined(FIXED_POINT)
// ---- Q31: 8-byte complex, 64-bit products -------------------------------------------------------------------
KF_PLAN(kP1024,    1024, {4, 4, 4, 4, 4},    {2, 2, 1}, 64,  2, 4, 4, 0);
KF_PLAN(kP2048,    2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 1, 4, 2, 0);
KF_PLAN(kP1000,    1000, {4, 2, 5, 5, 5},    {2, 2, 1}, 50,  4, 4, 1, 2);
KF_PLAN(kP1155,    1155, {3, 5, 7, 11},      {1, 1, 2}, 105, 2, 4, 1, 0);
KF_PLAN(kP1024col, 1024, {4, 4, 4, 4, 4},    {2, 2, 1}, 64,  8, 4, 1, 0);
KF_PLAN(kP2048col, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2}, 128, 4, 4, 1, 0);
#define KF_PLAN_LIST(X) \
    KF_POW2_LIST(X) X(kP1024, C2C_REAL) X(kP1024col, COL) X(kP2048, C2C_REAL) X(kP2048col, COL) X(kP1000, C2C_REAL) X(kP1155, C2C)
#elif defined(KF_IS_DOUBLE)   /* the double build passes -DKF_IS_DOUBLE next to -Dkiss_fft_scalar=double */
// ---- double: 16-byte complex --------------------------------------------------------------------------------
KF_PLAN(kP1024,    1024, {4, 4, 4, 4, 4},    {2, 2, 1},    64,  2, 4, 2, 2);
KF_PLAN(kP2048,    2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2},    128, 1, 4, 2, 2);
KF_PLAN(kP1000,    1000, {4, 2, 5, 5, 5},    {1, 1, 1, 2}, 200, 1, 3, 3, 1, 2, 1);
KF_PLAN(kP1155,    1155, {11, 7, 5, 3},      {2, 1, 1},    77,  1, 5, 4, 1);
KF_PLAN(kP1024col, 1024, {4, 4, 4, 4, 4},    {2, 2, 1},    64,  4, 4, 1, 0);
KF_PLAN(kP2048col, 2048, {4, 4, 4, 4, 4, 2}, {2, 2, 2},    128, 2, 4, 1, 0);
#define KF_PLAN_LIST(X) \
    KF_POW2_LIST(X) X(kP1024, C2C_REAL) X(kP1024col, COL) X(kP2048, C2C_REAL) X(kP2048col, COL) X(kP1000, C2C_REAL) X(kP1155, C2C)
#else
// ---- float: 8-byte complex ------------------------------------------------------------------------------------
// 1024 = (4*2*4) * (4*2*4): two 32-point register groups, one warp per transform, ONE shared-memory exchange
KF_PLAN(kP1024,    1024, {2, 4, 4, 4, 4, 2},    {3, 3},    32,  4, 5, 3, 1, 2, 1);
KF_PLAN(kP2048,    2048, {4, 2, 4, 4, 4, 4},    {2, 2, 2}, 128, 1, 4, 3, 1, 2, 1);
// kiss_fftr / kiss_fftri nfft = 4096 (packed complex length 2048): separately tuned per direction
KF_PLAN(kP2048r2c, 2048, {2, 2, 4, 4, 2, 4, 4}, {2, 2, 3}, 128, 1, 4, 3, 1, 2, 1);
KF_PLAN(kP2048c2r, 2048, {4, 2, 4, 4, 4, 4},    {2, 2, 2}, 128, 2, 4, 3, 1, 2, 1);
KF_PLAN(kP1000,    1000, {5, 5, 5, 4, 2},       {3, 2},    40,  2, 5, 4, 1);
KF_PLAN(kP1155,    1155, {3, 11, 5, 7},         {2, 2},    35,  2, 5, 5, 1);
// kiss_fftnd axis pass: 16 adjacent columns per CTA (128-byte row segments), single exchange buffer
KF_PLAN(kP1024col, 1024, {4, 2, 4, 4, 2, 4},    {3, 3},    32,  16, 5, 1, 0, 1);
KF_PLAN(kP2048col, 2048, {4, 4, 4, 4, 4, 2},    {2, 2, 2}, 128, 4, 4, 1, 0);
#define KF_PLAN_LIST(X) \
    KF_POW2_LIST(X) X(kP1024, C2C_REAL) X(kP1024col, COL) X(kP2048r2c, R2C) X(kP2048c2r, C2R) X(kP2048, C2C) X(kP2048col, COL) \
    X(kP1000, C2C_REAL) X(kP1155, C2C)
#endif

}   // namespace kf

// lv32_packed.cuh -- the headline kernels: Lotka-Volterra UDE with the 2 -> 32 -> 32 -> 2 tanh chain
// (BASELINE config 2), two trajectories per thread packed into FFMA2 lanes.
//
// Blackwell's FFMA2 (PTX fma.rn.f32x2) performs two fp32 FMAs per issue slot and accepts a SCALAR
// operand that is broadcast to both halves -- from a uniform register (weights read from the
// constant bank by LDCU.128, four weights per load) or from a normal register.  A thread therefore
// carries a PAIR of trajectories (a, b) in the two halves of every 64-bit register pair:
//   - chain forward / reverse sweeps:  acc(a,b) = fma2(w_scalar, h(a,b), acc(a,b))
//   - ensemble-summed parameter gradient: acc(a,b)[j][i] = fma2(q(a,b)[j], h(a,b)[i], acc(a,b)[j][i]),
//     halves added together once at the end of the kernel.
// This halves the issue slots of the FP32 work (the scalar version is issue-bound: 1 LDCU per 4 FFMA)
// and amortises every weight load over 8 FMAs.
//
// Reference semantics replaced: see ude_common.cuh / ude_adjoint.cuh headers.
#pragma once
#include "ude_common.cuh"

namespace b200ude {
namespace lv32 {

constexpr int H = 32;
constexpr int OFF_W1 = 0;              // W1[j, m] at m*32 + j      (32 x 2, column-major)
constexpr int OFF_B1 = 64;
constexpr int OFF_W2 = 96;             // W2[j, i] at i*32 + j      (32 x 32, column-major)
constexpr int OFF_B2 = OFF_W2 + 1024;  // 1120
constexpr int OFF_W3 = OFF_B2 + 32;    // 1152, W3[m, j] at j*2 + m (2 x 32, column-major)
constexpr int OFF_B3 = OFF_W3 + 64;    // 1216
constexpr int P = OFF_B3 + 2;          // 1218

// ---- packed helpers -------------------------------------------------------------------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }  // scalar broadcast operand
// s * x + y with scalar s (broadcast form of FFMA2)
__device__ __forceinline__ float2 fmas(float s, float2 x, float2 y) { return fma2(bc(s), x, y); }

template <int TM>
__device__ __forceinline__ float2 tanh2(float2 x)
{
    if constexpr (TM == 1) {
        return make_float2(tanh_dev<1>(x.x), tanh_dev<1>(x.y));
    } else {
        // tanh(x) = 1 - 2 / (2^(2 log2(e) x) + 1)
        const float2 z = mul2(x, bc(2.885390081777927f));
        float2 e, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(z.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(z.y));
        e = add2(e, bc(1.0f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(e.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(e.y));
        return fma2(bc(-2.0f), r, bc(1.0f));
    }
}

constexpr float TANH_ZMAX = 30.0f;   // clamp of 2 log2(e) x: tanh is 1.0f beyond x = 9.02, and (2^30 + 1)^4 < FLT_MAX
// tanh of four values with one reciprocal (TM = 0); tanh.approx (TM = 1)
template <int TM>
__device__ __forceinline__ void tanh_quad(float2 &a, float2 &b)
{
    if constexpr (TM == 1) {
        a = make_float2(tanh_dev<1>(a.x), tanh_dev<1>(a.y));
        b = make_float2(tanh_dev<1>(b.x), tanh_dev<1>(b.y));
    } else {
        float2 za = mul2(a, bc(2.885390081777927f)), zb = mul2(b, bc(2.885390081777927f));
        za.x = fminf(za.x, TANH_ZMAX); za.y = fminf(za.y, TANH_ZMAX);
        zb.x = fminf(zb.x, TANH_ZMAX); zb.y = fminf(zb.y, TANH_ZMAX);
        float2 ea, eb;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea.x) : "f"(za.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea.y) : "f"(za.y));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb.x) : "f"(zb.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb.y) : "f"(zb.y));
        const float2 da = add2(ea, bc(1.0f)), db = add2(eb, bc(1.0f));   // 1 + e^{2x} in [1, 2^30 + 1]
        const float2 pp = mul2(da, db);
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pp.x * pp.y));
        const float2 ip = make_float2(r * pp.y, r * pp.x);   // 1 / (da.x db.x), 1 / (da.y db.y)
        const float2 ia = mul2(ip, db), ib = mul2(ip, da);   // 1 / da, 1 / db
        a = fma2(bc(-2.0f), ia, bc(1.0f));
        b = fma2(bc(-2.0f), ib, bc(1.0f));
    }
}

// the same for four scalars
template <int TM>
__device__ __forceinline__ void tanh_quad_s(float &a, float &b, float &c, float &d)
{
    float2 p = make_float2(a, b), q = make_float2(c, d);
    tanh_quad<TM>(p, q);
    a = p.x; b = p.y; c = q.x; d = q.y;
}

struct Pair2 {
    float2 c0, c1;  // the two state components, each holding trajectories (a, b)
};

// ---- UDE right-hand side for a pair of trajectories (NOT inlined: keeps the weight loads next to
// their FFMA2s instead of hoisted out of the time loop and spilled; see ude_common.cuh) -------------
template <int TM, class W>
__device__ __noinline__ Pair2 rhs_pair(W w, Pair2 u)
{
    float2 h1[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float2 a = bc(w(OFF_B1 + j));
        a = fmas(w(OFF_W1 + j), u.c0, a);
        a = fmas(w(OFF_W1 + H + j), u.c1, a);
        h1[j] = tanh2<TM>(a);
    }
    float2 y0 = bc(w(OFF_B3 + 0)), y1 = bc(w(OFF_B3 + 1));
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
        float2 acc[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[jj] = bc(w(OFF_B2 + jb * 4 + jj));
#pragma unroll
        for (int i = 0; i < H; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[jj] = fmas(w(OFF_W2 + i * H + jb * 4 + jj), h1[i], acc[jj]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float2 v = tanh2<TM>(acc[jj]);
            y0 = fmas(w(OFF_W3 + (jb * 4 + jj) * 2 + 0), v, y0);
            y1 = fmas(w(OFF_W3 + (jb * 4 + jj) * 2 + 1), v, y1);
        }
    }
    Pair2 du;
    du.c0 = fmas(c_consts[0], u.c0, y0);    // du1 =  p1*u1 + NN1(u)   scenario_1.jl:71
    du.c1 = fmas(-c_consts[1], u.c1, y1);   // du2 = -p4*u2 + NN2(u)   scenario_1.jl:72
    return du;
}

// ---- forward kernel ---------------------------------------------------------------------------------
// thread tid integrates trajectories a = tid and b = tid + half (half = ceil(N/2)); all global
// accesses are lane-contiguous (128 B per warp per component).
template <int TM, class W, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) forward_kernel(FwdParams p)
{
    const W w = WeightStage<W>::load(p.theta, p.P);
    const int half = (p.N + 1) >> 1;
    const int tid = blockIdx.x * BLOCK + threadIdx.x;
    if (tid >= half) return;
    const size_t N = (size_t)p.N;
    const size_t na = (size_t)tid;
    const bool live_b = tid + half < p.N;
    const size_t nb = live_b ? (size_t)(tid + half) : na;  // odd N: the last thread's b duplicates a; stores masked
    const float dt = p.dt;
    Pair2 u;
    u.c0 = make_float2(__ldg(p.u0 + na), __ldg(p.u0 + nb));
    u.c1 = make_float2(__ldg(p.u0 + N + na), __ldg(p.u0 + N + nb));
    auto store2 = [&](float *base, int row, const Pair2 &v) {
        float *r0 = base + (size_t)row * 2 * N;
        r0[na] = v.c0.x;
        r0[N + na] = v.c1.x;
        if (live_b) {
            r0[nb] = v.c0.y;
            r0[N + nb] = v.c1.y;
        }
    };
    store2(p.out, 0, u);
    store2(p.ustep, 0, u);
    Pair2 k[7];
    k[0] = rhs_pair<TM>(w, u);
    store2(p.dense, 0, k[0]);
    int isave = 1;
#pragma unroll 1
    for (int s = 0; s < p.n_steps; ++s) {
#pragma unroll
        for (int i = 1; i < 7; ++i) {
            float2 a0 = bc(0.0f), a1 = bc(0.0f);
#pragma unroll
            for (int j = 0; j < i; ++j)
                if (Tsit5::a(i, j) != 0.0) {
                    a0 = fmas((float)Tsit5::a(i, j), k[j].c0, a0);
                    a1 = fmas((float)Tsit5::a(i, j), k[j].c1, a1);
                }
            Pair2 g;
            g.c0 = fmas(dt, a0, u.c0);
            g.c1 = fmas(dt, a1, u.c1);
            if (i == 6) u = g;  // stage 7's argument is u_{n+1} (row 7 = b, FSAL)
            k[i] = rhs_pair<TM>(w, g);
            store2(p.dense, s * 6 + i, k[i]);  // row (s+1)*6 for i = 6: k_7(s) = k_1(s+1)
        }
        store2(p.ustep, s + 1, u);
        if ((s + 1) % p.save_every == 0) {
            store2(p.out, isave, u);
            ++isave;
        }
        k[0] = k[6];
    }
    if (p.status) {
        const bool oka = (fabsf(u.c0.x) <= 3.0e38f) && (fabsf(u.c1.x) <= 3.0e38f);  // false for NaN/Inf
        const bool okb = (fabsf(u.c0.y) <= 3.0e38f) && (fabsf(u.c1.y) <= 3.0e38f);
        p.status[na] = oka ? 0 : 1;
        if (live_b) p.status[nb] = okb ? 0 : 1;
    }
}

// ---- adjoint ---------------------------------------------------------------------------------------
// Per-warp shared-memory staging.  Row r belongs to lane r and holds, for every hidden unit j, the
// PAIR (value for trajectory a, value for trajectory b) -- exactly the thread's register pairs, so the
// 128-bit row stores need no shuffling.  Element (j, a|b) of a row sits at 2j + (a|b) + (j >= 16 ? 4 : 0):
// the 4-float gap in the middle of the row makes the four 64-byte j-tiles (and the eight 32-byte
// i-tiles) that the 32 lanes read in ONE instruction fall on distinct banks; LD = 68 keeps the
// 128-bit row stores conflict-free (8 consecutive lanes -> 8 distinct bank groups).
constexpr int LD = 68;
__device__ __forceinline__ constexpr int rowoff(int j) { return 2 * j + (j >= 16 ? 4 : 0); }
struct __align__(16) WarpStage3 {
    float B1[32 * LD];   // h2 rows                   [lane][j][a|b]
    float B2[32 * LD];   // h1 rows, later q1 rows    [lane][j][a|b]
    float B3[32 * LD];   // q2 rows                   [lane][j][a|b]
    float SG[32 * 4];    // scaled output cotangent   [lane][m][a|b]
    float U[32 * 4];     // chain input               [lane][m][a|b]
};

// Loop-variant zero: adding c_zero[i] (unknown to ptxas) to every constant-bank weight address keeps
// the weight loads INSIDE the stage loop (otherwise ptxas hoists the loop-invariant loads out of the
// time loops and spills them, see ude_common.cuh), at the cost of one LDCU per stage.
static __constant__ int c_zero[8];
// four consecutive weights with one 128-bit constant-bank load (idx must be a multiple of 4)
__device__ __forceinline__ float4 ldw4(int idx) { return *reinterpret_cast<const float4 *>(&c_theta[idx]); }

// u(t_{s+1} - c_I dt) for both trajectories from the stored dense output of forward step s
template <int I>
__device__ __forceinline__ Pair2 interp_pair(const AdjParams &p, int s, size_t na, size_t nb, size_t N, float dt)
{
    float2 a0 = bc(0.0f), a1 = bc(0.0f);
#pragma unroll
    for (int j = 0; j < 7; ++j)
        if (Tsit5::bw(I, j) != 0.0) {
            const float *row = p.dense + (size_t)(s * 6 + j) * 2 * N;
            a0 = fmas((float)Tsit5::bw(I, j), make_float2(__ldg(row + na), __ldg(row + nb)), a0);
            a1 = fmas((float)Tsit5::bw(I, j), make_float2(__ldg(row + N + na), __ldg(row + N + nb)), a1);
        }
    const float *ur = p.ustep + (size_t)s * 2 * N;
    Pair2 x;
    x.c0 = fmas(dt, a0, make_float2(__ldg(ur + na), __ldg(ur + nb)));
    x.c1 = fmas(dt, a1, make_float2(__ldg(ur + N + na), __ldg(ur + N + nb)));
    return x;
}

__device__ __forceinline__ void jump_pair(const AdjParams &p, int isave, size_t na, size_t nb, size_t N, Pair2 &lam, float2 &loss)
{
    const float *cr = p.cot + (size_t)isave * 2 * N;
    const float2 c0 = make_float2(__ldg(cr + na), __ldg(cr + nb)), c1 = make_float2(__ldg(cr + N + na), __ldg(cr + N + nb));
    if (p.fused_l2) {
        const float *ur = p.ustep + (size_t)(isave * p.save_every) * 2 * N;
        const float2 r0 = add2(make_float2(__ldg(ur + na), __ldg(ur + nb)), mul2(c0, bc(-1.0f)));
        const float2 r1 = add2(make_float2(__ldg(ur + N + na), __ldg(ur + N + nb)), mul2(c1, bc(-1.0f)));
        const float w0 = c_lossw[0], w1 = c_lossw[1];
        loss = fma2(mul2(bc(w0), r0), r0, loss);
        loss = fma2(mul2(bc(w1), r1), r1, loss);
        lam.c0 = fmas(2.0f * w0, r0, lam.c0);
        lam.c1 = fmas(2.0f * w1, r1, lam.c1);
    } else {
        lam.c0 = add2(lam.c0, c0);
        lam.c1 = add2(lam.c1, c1);
    }
}


// The whole InterpolatingAdjoint backward solve for 64 trajectories per warp.  One warp per CTA.
// Stage body (executed 6 x n_steps times, as ONE rolled loop so that it fits the instruction cache):
//   1. chain forward: h1 -> registers + B2; h2 -> B1; q2 = W3^T sg * (1 - h2^2) -> B3
//   2. GEMM/thin pass over the 32 staged rows (software-pipelined 128-bit shared loads):
//        dW2[j][i] += q2[t][j] h1[t][i]   (lane (jt, it) owns an 8 x 4 tile, accumulators in registers)
//        db2, dW3, db3 as lane-owned columns
//   3. q1 = (W2^T q2) * (1 - h1^2) -> B2 (in place of h1), (df/du)^T g
//   4. thin pass: dW1, db1 as lane-owned columns
template <int TM, class W, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) adjoint_kernel(AdjParams p)
{
    static_assert(BLOCK == 32, "one warp (64 trajectories) per CTA");
    static_assert(!WeightStage<W>::kSmem, "the adjoint reads its weights from the constant bank");
    const int lane = threadIdx.x & 31;
    const int jt = lane >> 3, it = lane & 7;
    const int half = (p.N + 1) >> 1;
    const int tid = blockIdx.x * BLOCK + threadIdx.x;
    const size_t N = (size_t)p.N;
    const bool live_a = tid < half;
    const size_t na = live_a ? (size_t)tid : 0;
    const bool live_b = live_a && (tid + half < p.N);
    const size_t nb = live_b ? (size_t)(tid + half) : na;
    const float dt = p.dt;
    // padding trajectories carry a zero mask on their cotangents: they contribute nothing to the sums
    const float2 lv = make_float2(live_a ? 1.0f : 0.0f, live_b ? 1.0f : 0.0f);

    extern __shared__ __align__(16) unsigned char s_raw[];
    WarpStage3 *st = reinterpret_cast<WarpStage3 *>(s_raw);
    float *const rowB1 = st->B1 + lane * LD, *const rowB2 = st->B2 + lane * LD, *const rowB3 = st->B3 + lane * LD;
    const int goff = jt * 16 + (jt >= 2 ? 4 : 0);   // this lane's j-tile inside a row
    const int hoff = it * 8 + (it >= 4 ? 4 : 0);    // this lane's i-tile inside a row
    const int coff = rowoff(lane);                  // this lane's column inside a row

    // ---- gradient accumulators (registers, live across the whole backward solve) ----
    float2 acc[32];       // dW2 tile: acc[jj*4 + ii] for j = jt*8 + jj, i = it*4 + ii; halves a|b added at the end
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = bc(0.0f);
    float2 g_w30 = bc(0.f), g_w31 = bc(0.f), g_b30 = bc(0.f), g_b31 = bc(0.f), g_b2 = bc(0.f);
    float2 g_b1 = bc(0.f), g_w10 = bc(0.f), g_w11 = bc(0.f);

    Pair2 lam;
    lam.c0 = bc(0.0f);
    lam.c1 = bc(0.0f);
    float2 loss = bc(0.0f);
    const int n_save = p.n_steps / p.save_every + 1;
    jump_pair(p, n_save - 1, na, nb, N, lam, loss);
    const float p1 = c_consts[0], p4 = c_consts[1];
    const float inv_dt = 1.0f / dt;

#pragma unroll 1
    for (int s = p.n_steps - 1; s >= 0; --s) {
        Pair2 kl[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) { kl[j].c0 = bc(0.0f); kl[j].c1 = bc(0.0f); }
#pragma unroll 1
        for (int stage = 0; stage < 6; ++stage) {
            Pair2 x, g;
            float sc, isc;
            // stage-specific part: interpolation weights, Runge-Kutta row, quadrature weight (all static per case)
#define B200UDE_STAGE_PRE(I)                                                          \
    case I: {                                                                         \
        x = interp_pair<I>(p, s, na, nb, N, dt);                                      \
        float2 a0 = bc(0.0f), a1 = bc(0.0f);                                          \
        _Pragma("unroll") for (int j = 0; j < I; ++j) if (Tsit5::a(I, j) != 0.0) {    \
            a0 = fmas((float)Tsit5::a(I, j), kl[j].c0, a0);                           \
            a1 = fmas((float)Tsit5::a(I, j), kl[j].c1, a1);                           \
        }                                                                             \
        g.c0 = fmas(dt, a0, lam.c0);                                                  \
        g.c1 = fmas(dt, a1, lam.c1);                                                  \
        sc = dt * (float)Tsit5::b(I);                                                 \
        isc = inv_dt * (float)(1.0 / Tsit5::b(I));                                    \
    } break;
            switch (stage) {
                B200UDE_STAGE_PRE(0)
                B200UDE_STAGE_PRE(1)
                B200UDE_STAGE_PRE(2)
                B200UDE_STAGE_PRE(3)
                B200UDE_STAGE_PRE(4)
            default:
                B200UDE_STAGE_PRE(5)
            }
#undef B200UDE_STAGE_PRE
            const int zb = c_zero[stage] << 2;   // == 0 (and provably a multiple of 4), opaque to ptxas: pins the weight loads inside this loop
            // cotangent of the chain output, carrying the quadrature weight and the padding mask
            const float2 sg0 = mul2(mul2(lv, bc(sc)), g.c0), sg1 = mul2(mul2(lv, bc(sc)), g.c1);
            *reinterpret_cast<float4 *>(&st->SG[lane * 4]) = make_float4(sg0.x, sg0.y, sg1.x, sg1.y);
            *reinterpret_cast<float4 *>(&st->U[lane * 4]) = make_float4(x.c0.x, x.c0.y, x.c1.x, x.c1.y);

            // ---- 1. chain forward ----
            float2 h1[H];
#pragma unroll
            for (int j4 = 0; j4 < H; j4 += 4) {
                const float4 wb1 = ldw4(zb + OFF_B1 + j4), w10 = ldw4(zb + OFF_W1 + j4), w11 = ldw4(zb + OFF_W1 + H + j4);
                const float b_[4] = {wb1.x, wb1.y, wb1.z, wb1.w}, w0_[4] = {w10.x, w10.y, w10.z, w10.w}, w1_[4] = {w11.x, w11.y, w11.z, w11.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) h1[j4 + k] = tanh2<TM>(fmas(w1_[k], x.c1, fmas(w0_[k], x.c0, bc(b_[k]))));
            }
#pragma unroll
            for (int c = 0; c < 16; ++c)
                *reinterpret_cast<float4 *>(rowB2 + 4 * c + (c >= 8 ? 4 : 0)) = make_float4(h1[2 * c].x, h1[2 * c].y, h1[2 * c + 1].x, h1[2 * c + 1].y);
#pragma unroll 1
            for (int jb = 0; jb < 4; ++jb) {   // 8 outputs per iteration
                const int wb = zb + jb * 8;
                float2 a8[8];
                {
                    const float4 bA = ldw4(wb + OFF_B2), bB = ldw4(wb + OFF_B2 + 4);
                    a8[0] = bc(bA.x); a8[1] = bc(bA.y); a8[2] = bc(bA.z); a8[3] = bc(bA.w);
                    a8[4] = bc(bB.x); a8[5] = bc(bB.y); a8[6] = bc(bB.z); a8[7] = bc(bB.w);
                }
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float4 wA = ldw4(wb + OFF_W2 + i * H), wB = ldw4(wb + OFF_W2 + i * H + 4);
                    a8[0] = fmas(wA.x, h1[i], a8[0]); a8[1] = fmas(wA.y, h1[i], a8[1]);
                    a8[2] = fmas(wA.z, h1[i], a8[2]); a8[3] = fmas(wA.w, h1[i], a8[3]);
                    a8[4] = fmas(wB.x, h1[i], a8[4]); a8[5] = fmas(wB.y, h1[i], a8[5]);
                    a8[6] = fmas(wB.z, h1[i], a8[6]); a8[7] = fmas(wB.w, h1[i], a8[7]);
                }
                float2 h2[8], q2[8];
#pragma unroll
                for (int c = 0; c < 4; ++c) {   // W3[m][j] at j*2 + m: (w0j, w1j, w0j+1, w1j+1)
                    const float4 w3 = ldw4(2 * wb - zb + OFF_W3 + 4 * c);
                    const float w3_[4] = {w3.x, w3.y, w3.z, w3.w};
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int jj = 2 * c + k;
                        h2[jj] = tanh2<TM>(a8[jj]);
                        const float2 t = fmas(w3_[2 * k + 1], sg1, mul2(bc(w3_[2 * k]), sg0));
                        q2[jj] = mul2(t, fma2(mul2(h2[jj], bc(-1.0f)), h2[jj], bc(1.0f)));   // * (1 - h2^2)
                    }
                }
                const int ro = jb * 16 + (jb >= 2 ? 4 : 0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    *reinterpret_cast<float4 *>(rowB1 + ro + 4 * c) = make_float4(h2[2 * c].x, h2[2 * c].y, h2[2 * c + 1].x, h2[2 * c + 1].y);
                    *reinterpret_cast<float4 *>(rowB3 + ro + 4 * c) = make_float4(q2[2 * c].x, q2[2 * c].y, q2[2 * c + 1].x, q2[2 * c + 1].y);
                }
            }
            __syncwarp();

            // ---- 2. GEMM + thin pass over the staged rows, operands double-buffered in registers ----
            {
                float4 G[2][4], Hh[2][2], S[2];
                float2 hc[2], qc[2];
                auto load_row = [&](int t, int b) {
                    const float *r3 = st->B3 + t * LD, *r2 = st->B2 + t * LD, *r1 = st->B1 + t * LD;
#pragma unroll
                    for (int c = 0; c < 4; ++c) G[b][c] = *reinterpret_cast<const float4 *>(r3 + goff + 4 * c);
#pragma unroll
                    for (int c = 0; c < 2; ++c) Hh[b][c] = *reinterpret_cast<const float4 *>(r2 + hoff + 4 * c);
                    hc[b] = *reinterpret_cast<const float2 *>(r1 + coff);
                    qc[b] = *reinterpret_cast<const float2 *>(r3 + coff);
                    S[b] = *reinterpret_cast<const float4 *>(&st->SG[t * 4]);
                };
                auto use_row = [&](int b) {
                    const float2 gj[8] = {make_float2(G[b][0].x, G[b][0].y), make_float2(G[b][0].z, G[b][0].w),
                                          make_float2(G[b][1].x, G[b][1].y), make_float2(G[b][1].z, G[b][1].w),
                                          make_float2(G[b][2].x, G[b][2].y), make_float2(G[b][2].z, G[b][2].w),
                                          make_float2(G[b][3].x, G[b][3].y), make_float2(G[b][3].z, G[b][3].w)};
                    const float2 hi[4] = {make_float2(Hh[b][0].x, Hh[b][0].y), make_float2(Hh[b][0].z, Hh[b][0].w),
                                          make_float2(Hh[b][1].x, Hh[b][1].y), make_float2(Hh[b][1].z, Hh[b][1].w)};
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) acc[jj * 4 + ii] = fma2(gj[jj], hi[ii], acc[jj * 4 + ii]);
                    const float2 s0 = make_float2(S[b].x, S[b].y), s1 = make_float2(S[b].z, S[b].w);
                    g_w30 = fma2(s0, hc[b], g_w30);
                    g_w31 = fma2(s1, hc[b], g_w31);
                    g_b30 = add2(g_b30, s0);
                    g_b31 = add2(g_b31, s1);
                    g_b2 = add2(g_b2, qc[b]);
                };
                load_row(0, 0);
#pragma unroll 1
                for (int t = 0; t < 32; t += 2) {
                    load_row(t + 1, 1);
                    use_row(0);
                    if (t + 2 < 32) load_row(t + 2, 0);
                    use_row(1);
                }
            }
            __syncwarp();

            // ---- 3. q1 = (W2^T q2) * (1 - h1^2), written over h1 in B2; (df/du)^T g ----
            float2 dx0 = bc(0.0f), dx1 = bc(0.0f);
            {
                float2 q2[H];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float4 v = *reinterpret_cast<const float4 *>(rowB3 + 4 * c + (c >= 8 ? 4 : 0));
                    q2[2 * c] = make_float2(v.x, v.y);
                    q2[2 * c + 1] = make_float2(v.z, v.w);
                }
#pragma unroll 1
                for (int ib = 0; ib < 8; ++ib) {   // 4 inputs i per iteration, two partial sums each
                    const int wb = zb + ib * 4 * H;
                    float2 a0[4], a1[4];
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) { a0[ii] = bc(0.0f); a1[ii] = bc(0.0f); }
#pragma unroll
                    for (int j = 0; j < H; j += 4)
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            const float4 w4 = ldw4(wb + OFF_W2 + ii * H + j);
                            a0[ii] = fmas(w4.x, q2[j], a0[ii]);
                            a1[ii] = fmas(w4.y, q2[j + 1], a1[ii]);
                            a0[ii] = fmas(w4.z, q2[j + 2], a0[ii]);
                            a1[ii] = fmas(w4.w, q2[j + 3], a1[ii]);
                        }
                    const int ro = ib * 8 + (ib >= 4 ? 4 : 0);
                    const float4 hA = *reinterpret_cast<const float4 *>(rowB2 + ro), hB = *reinterpret_cast<const float4 *>(rowB2 + ro + 4);
                    const float2 hb[4] = {make_float2(hA.x, hA.y), make_float2(hA.z, hA.w), make_float2(hB.x, hB.y), make_float2(hB.z, hB.w)};
                    float2 q1[4];
                    const float4 wx0 = ldw4(zb + OFF_W1 + ib * 4), wx1 = ldw4(zb + OFF_W1 + H + ib * 4);
                    const float wx0_[4] = {wx0.x, wx0.y, wx0.z, wx0.w}, wx1_[4] = {wx1.x, wx1.y, wx1.z, wx1.w};
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        q1[ii] = mul2(add2(a0[ii], a1[ii]), fma2(mul2(hb[ii], bc(-1.0f)), hb[ii], bc(1.0f)));
                        dx0 = fmas(wx0_[ii], q1[ii], dx0);
                        dx1 = fmas(wx1_[ii], q1[ii], dx1);
                    }
                    *reinterpret_cast<float4 *>(rowB2 + ro) = make_float4(q1[0].x, q1[0].y, q1[1].x, q1[1].y);
                    *reinterpret_cast<float4 *>(rowB2 + ro + 4) = make_float4(q1[2].x, q1[2].y, q1[3].x, q1[3].y);
                }
            }
            // (df/du)^T g with the LV physics diag(p1, -p4); dx carries sc (and the mask of padding lanes)
            Pair2 kn;
            kn.c0 = fmas(p1, g.c0, mul2(dx0, bc(isc)));
            kn.c1 = fmas(-p4, g.c1, mul2(dx1, bc(isc)));
            switch (stage) {
            case 0: kl[0] = kn; break;
            case 1: kl[1] = kn; break;
            case 2: kl[2] = kn; break;
            case 3: kl[3] = kn; break;
            case 4: kl[4] = kn; break;
            default: kl[5] = kn; break;
            }
            __syncwarp();

            // ---- 4. thin pass: lane i owns input-layer row i: dW1[i][m] += sum_t q1[t][i] x[t][m]; db1 ----
#pragma unroll 8
            for (int t = 0; t < 32; ++t) {
                const float2 qq = *reinterpret_cast<const float2 *>(st->B2 + t * LD + coff);
                const float4 u4 = *reinterpret_cast<const float4 *>(&st->U[t * 4]);
                g_b1 = add2(g_b1, qq);
                g_w10 = fma2(qq, make_float2(u4.x, u4.y), g_w10);
                g_w11 = fma2(qq, make_float2(u4.z, u4.w), g_w11);
            }
            __syncwarp();
        }
        float2 a0 = bc(0.0f), a1 = bc(0.0f);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            a0 = fmas((float)Tsit5::b(j), kl[j].c0, a0);
            a1 = fmas((float)Tsit5::b(j), kl[j].c1, a1);
        }
        lam.c0 = fmas(dt, a0, lam.c0);
        lam.c1 = fmas(dt, a1, lam.c1);
        if (s % p.save_every == 0) jump_pair(p, s / p.save_every, na, nb, N, lam, loss);
    }
    if (p.grad_u0) {
        if (live_a) { p.grad_u0[na] = lam.c0.x; p.grad_u0[N + na] = lam.c1.x; }
        if (live_b) { p.grad_u0[nb] = lam.c0.y; p.grad_u0[N + nb] = lam.c1.y; }
    }
    float lsum = loss.x * lv.x + loss.y * lv.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);

    // ---- this warp's partial gradient (halves a|b added here) ----
    float *dst = p.partial + (size_t)blockIdx.x * (P + 1);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) dst[OFF_W2 + (it * 4 + ii) * H + (jt * 8 + jj)] = acc[jj * 4 + ii].x + acc[jj * 4 + ii].y;
    dst[OFF_W3 + lane * 2 + 0] = g_w30.x + g_w30.y;
    dst[OFF_W3 + lane * 2 + 1] = g_w31.x + g_w31.y;
    dst[OFF_B2 + lane] = g_b2.x + g_b2.y;
    dst[OFF_B1 + lane] = g_b1.x + g_b1.y;
    dst[OFF_W1 + lane] = g_w10.x + g_w10.y;
    dst[OFF_W1 + H + lane] = g_w11.x + g_w11.y;
    if (lane == 0) {
        dst[OFF_B3 + 0] = g_b30.x + g_b30.y;   // identical in every lane (sums over the staged rows)
        dst[OFF_B3 + 1] = g_b31.x + g_b31.y;
        dst[P] = lsum;
    }
}

}  // namespace lv32
}  // namespace b200ude

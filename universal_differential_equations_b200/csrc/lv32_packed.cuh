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
// 128-bit row stores need no shuffling.  LD = 68 floats makes both the row stores (8 consecutive lanes
// -> 8 distinct bank groups) and the broadcast tile loads conflict-free.
constexpr int LD = 68;
struct __align__(16) WarpStage2 {
    float B1[32 * LD];   // h2, then q2, then q1     [lane][j][a|b]
    float B2[32 * LD];   // h1                        [lane][j][a|b]
    float SG[32 * 4];    // scaled output cotangent   [lane][m][a|b]
    float U[32 * 4];     // chain input               [lane][m][a|b]
    float ACC[64 * 32];  // dW2 tile accumulators: lane (jt, it) owns [jj 0..7][ii 0..3][a|b]; stored [(c*32 + lane)*4 + k]
    float THIN[16 * 32]; // lane-owned rows of the thin layers: 8 float2 per lane, stored [(c*32 + lane)*4 + k], c < 4
};

struct AdjStageIn {
    Pair2 x;    // u(t) at this stage (chain input)
    Pair2 g;    // backward stage argument lambda_i
    float sc;   // quadrature weight dt*b_i
    float isc;  // 1/sc
    float lva, lvb;  // 1 for live trajectories, 0 for padding (contribute zero to the gradient)
};

// One backward stage: chain forward, reverse sweep, (df/du)^T g, and this stage's contribution to the
// ensemble-summed parameter gradient (per-warp outer-product GEMM over the warp's 64 trajectories).
template <int TM, class W>
__device__ __noinline__ Pair2 adj_stage(W w, WarpStage2 *st, AdjStageIn in)
{
    const int lane = threadIdx.x & 31;
    const int jt = lane >> 3, it = lane & 7;
    const float2 lv = make_float2(in.lva, in.lvb);
    const float2 sg0 = mul2(bc(in.sc), in.g.c0), sg1 = mul2(bc(in.sc), in.g.c1);
    {
        const float2 m0 = mul2(lv, sg0), m1 = mul2(lv, sg1);
        *reinterpret_cast<float4 *>(&st->SG[lane * 4]) = make_float4(m0.x, m0.y, m1.x, m1.y);
        *reinterpret_cast<float4 *>(&st->U[lane * 4]) = make_float4(in.x.c0.x, in.x.c0.y, in.x.c1.x, in.x.c1.y);
    }
    // ---- layer 1 ----
    float2 h1[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float2 a = bc(w(OFF_B1 + j));
        a = fmas(w(OFF_W1 + j), in.x.c0, a);
        a = fmas(w(OFF_W1 + H + j), in.x.c1, a);
        h1[j] = tanh2<TM>(a);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c)
        *reinterpret_cast<float4 *>(&st->B2[lane * LD + 4 * c]) = make_float4(h1[2 * c].x, h1[2 * c].y, h1[2 * c + 1].x, h1[2 * c + 1].y);
    // ---- layer 2 in blocks of 4 outputs; h2 goes straight to B1, q2 stays in registers ----
    float2 q2[H];
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
        float2 acc[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[jj] = bc(w(OFF_B2 + jb * 4 + jj));
#pragma unroll
        for (int i = 0; i < H; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[jj] = fmas(w(OFF_W2 + i * H + jb * 4 + jj), h1[i], acc[jj]);
        float2 h2[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            h2[jj] = tanh2<TM>(acc[jj]);
            const int j = jb * 4 + jj;
            const float2 t = fmas(w(OFF_W3 + j * 2 + 1), sg1, mul2(bc(w(OFF_W3 + j * 2)), sg0));
            q2[j] = mul2(t, fma2(mul2(h2[jj], bc(-1.0f)), h2[jj], bc(1.0f)));  // * (1 - h2^2)
        }
        *reinterpret_cast<float4 *>(&st->B1[lane * LD + 8 * jb]) = make_float4(h2[0].x, h2[0].y, h2[1].x, h2[1].y);
        *reinterpret_cast<float4 *>(&st->B1[lane * LD + 8 * jb + 4]) = make_float4(h2[2].x, h2[2].y, h2[3].x, h2[3].y);
    }
    __syncwarp();
    // ---- thin pass A: lane j owns output-layer column j: dW3[m][j] += sum_t SG[t][m] * h2[t][j]; db3 ----
    {
        float4 ta = *reinterpret_cast<const float4 *>(&st->THIN[(0 * 32 + lane) * 4]);  // (w30.a, w30.b, w31.a, w31.b)
        float4 tb = *reinterpret_cast<const float4 *>(&st->THIN[(1 * 32 + lane) * 4]);  // (b30.a, b30.b, b31.a, b31.b)
        float2 w30 = make_float2(ta.x, ta.y), w31 = make_float2(ta.z, ta.w);
        float2 b30 = make_float2(tb.x, tb.y), b31 = make_float2(tb.z, tb.w);
#pragma unroll 8
        for (int t = 0; t < 32; ++t) {
            const float2 hh = *reinterpret_cast<const float2 *>(&st->B1[t * LD + 2 * lane]);
            const float4 s4 = *reinterpret_cast<const float4 *>(&st->SG[t * 4]);
            const float2 s0 = make_float2(s4.x, s4.y), s1 = make_float2(s4.z, s4.w);
            w30 = fma2(s0, hh, w30);
            w31 = fma2(s1, hh, w31);
            b30 = add2(b30, s0);
            b31 = add2(b31, s1);
        }
        *reinterpret_cast<float4 *>(&st->THIN[(0 * 32 + lane) * 4]) = make_float4(w30.x, w30.y, w31.x, w31.y);
        *reinterpret_cast<float4 *>(&st->THIN[(1 * 32 + lane) * 4]) = make_float4(b30.x, b30.y, b31.x, b31.y);
    }
    __syncwarp();
    // ---- B1 <- q2 (masked); q1 = (W2^T q2) * (1 - h1^2) overwrites h1 ----
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float2 m0 = mul2(lv, q2[2 * c]), m1 = mul2(lv, q2[2 * c + 1]);
        *reinterpret_cast<float4 *>(&st->B1[lane * LD + 4 * c]) = make_float4(m0.x, m0.y, m1.x, m1.y);
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        float2 a0 = bc(0.0f), a1 = bc(0.0f);
#pragma unroll
        for (int j = 0; j < H; j += 2) {
            a0 = fmas(w(OFF_W2 + i * H + j), q2[j], a0);
            a1 = fmas(w(OFF_W2 + i * H + j + 1), q2[j + 1], a1);
        }
        h1[i] = mul2(add2(a0, a1), fma2(mul2(h1[i], bc(-1.0f)), h1[i], bc(1.0f)));  // h1[] now holds q1[]
    }
    __syncwarp();
    // ---- GEMM pass: ACC[j][i](a|b) += q2[t][j](a|b) * h1[t][i](a|b) over the 32 staged rows ----
    {
        float2 acc[32];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 v = *reinterpret_cast<const float4 *>(&st->ACC[(c * 32 + lane) * 4]);
            acc[2 * c] = make_float2(v.x, v.y);
            acc[2 * c + 1] = make_float2(v.z, v.w);
        }
        float4 tc = *reinterpret_cast<const float4 *>(&st->THIN[(2 * 32 + lane) * 4]);  // (b2.a, b2.b, b1.a, b1.b)
        float2 gb2 = make_float2(tc.x, tc.y);
#pragma unroll 2
        for (int t = 0; t < 32; ++t) {
            float2 gj[8], hi[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 v = *reinterpret_cast<const float4 *>(&st->B1[t * LD + jt * 16 + 4 * c]);
                gj[2 * c] = make_float2(v.x, v.y);
                gj[2 * c + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 v = *reinterpret_cast<const float4 *>(&st->B2[t * LD + it * 8 + 4 * c]);
                hi[2 * c] = make_float2(v.x, v.y);
                hi[2 * c + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) acc[jj * 4 + ii] = fma2(gj[jj], hi[ii], acc[jj * 4 + ii]);
            gb2 = add2(gb2, *reinterpret_cast<const float2 *>(&st->B1[t * LD + 2 * lane]));  // db2[lane]
        }
#pragma unroll
        for (int c = 0; c < 16; ++c)
            *reinterpret_cast<float4 *>(&st->ACC[(c * 32 + lane) * 4]) = make_float4(acc[2 * c].x, acc[2 * c].y, acc[2 * c + 1].x, acc[2 * c + 1].y);
        *reinterpret_cast<float2 *>(&st->THIN[(2 * 32 + lane) * 4]) = gb2;
    }
    // ---- input cotangent and (df/du)^T g ----
    float2 dx0 = bc(0.0f), dx1 = bc(0.0f);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        dx0 = fmas(w(OFF_W1 + j), h1[j], dx0);
        dx1 = fmas(w(OFF_W1 + H + j), h1[j], dx1);
    }
    Pair2 kl;
    kl.c0 = fmas(c_consts[0], in.g.c0, mul2(dx0, bc(in.isc)));    // LV physics: diag(p1, -p4)
    kl.c1 = fmas(-c_consts[1], in.g.c1, mul2(dx1, bc(in.isc)));
    __syncwarp();
    // ---- B1 <- q1 (masked); thin pass B: lane i owns input-layer row i: dW1[i][m] += sum_t q1[t][i] x[t][m]; db1 ----
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float2 m0 = mul2(lv, h1[2 * c]), m1 = mul2(lv, h1[2 * c + 1]);
        *reinterpret_cast<float4 *>(&st->B1[lane * LD + 4 * c]) = make_float4(m0.x, m0.y, m1.x, m1.y);
    }
    __syncwarp();
    {
        float2 gb1 = *reinterpret_cast<const float2 *>(&st->THIN[(2 * 32 + lane) * 4 + 2]);
        float4 td = *reinterpret_cast<const float4 *>(&st->THIN[(3 * 32 + lane) * 4]);  // (w10.a, w10.b, w11.a, w11.b)
        float2 w10 = make_float2(td.x, td.y), w11 = make_float2(td.z, td.w);
#pragma unroll 8
        for (int t = 0; t < 32; ++t) {
            const float2 qq = *reinterpret_cast<const float2 *>(&st->B1[t * LD + 2 * lane]);
            const float4 u4 = *reinterpret_cast<const float4 *>(&st->U[t * 4]);
            gb1 = add2(gb1, qq);
            w10 = fma2(qq, make_float2(u4.x, u4.y), w10);
            w11 = fma2(qq, make_float2(u4.z, u4.w), w11);
        }
        *reinterpret_cast<float2 *>(&st->THIN[(2 * 32 + lane) * 4 + 2]) = gb1;
        *reinterpret_cast<float4 *>(&st->THIN[(3 * 32 + lane) * 4]) = make_float4(w10.x, w10.y, w11.x, w11.y);
    }
    __syncwarp();
    return kl;
}

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

template <int TM, class W, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) adjoint_kernel(AdjParams p)
{
    constexpr int NWARP = BLOCK / 32;
    const W w = WeightStage<W>::load(p.theta, p.P);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int jt = lane >> 3, it = lane & 7;
    const int half = (p.N + 1) >> 1;
    const int tid = blockIdx.x * BLOCK + threadIdx.x;
    const size_t N = (size_t)p.N;
    const bool live_a = tid < half;
    const size_t na = live_a ? (size_t)tid : 0;
    const bool live_b = live_a && (tid + half < p.N);
    const size_t nb = live_b ? (size_t)(tid + half) : na;
    const float dt = p.dt;

    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int THETA_SMEM_FLOATS = WeightStage<W>::kSmem ? ((P + 3) / 4) * 4 : 0;
    WarpStage2 *st = reinterpret_cast<WarpStage2 *>(s_raw + sizeof(float) * THETA_SMEM_FLOATS) + warp;
#pragma unroll
    for (int c = 0; c < 16; ++c) *reinterpret_cast<float4 *>(&st->ACC[(c * 32 + lane) * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<float4 *>(&st->THIN[(c * 32 + lane) * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    Pair2 lam;
    lam.c0 = bc(0.0f);
    lam.c1 = bc(0.0f);
    float2 loss = bc(0.0f);
    const int n_save = p.n_steps / p.save_every + 1;
    jump_pair(p, n_save - 1, na, nb, N, lam, loss);

    AdjStageIn in;
    in.lva = live_a ? 1.0f : 0.0f;
    in.lvb = live_b ? 1.0f : 0.0f;

#pragma unroll 1
    for (int s = p.n_steps - 1; s >= 0; --s) {
        Pair2 kl[6];
#define B200UDE_ADJ_STAGE(I)                                                          \
    {                                                                                 \
        in.x = interp_pair<I>(p, s, na, nb, N, dt);                                   \
        float2 a0 = bc(0.0f), a1 = bc(0.0f);                                          \
        _Pragma("unroll") for (int j = 0; j < I; ++j) if (Tsit5::a(I, j) != 0.0) {    \
            a0 = fmas((float)Tsit5::a(I, j), kl[j].c0, a0);                           \
            a1 = fmas((float)Tsit5::a(I, j), kl[j].c1, a1);                           \
        }                                                                             \
        in.g.c0 = fmas(dt, a0, lam.c0);                                               \
        in.g.c1 = fmas(dt, a1, lam.c1);                                               \
        in.sc = dt * (float)Tsit5::b(I);                                              \
        in.isc = 1.0f / in.sc;                                                        \
        kl[I] = adj_stage<TM>(w, st, in);                                             \
    }
        // k_7 of the backward step only feeds FSAL / error estimation: not needed
        B200UDE_ADJ_STAGE(0)
        B200UDE_ADJ_STAGE(1)
        B200UDE_ADJ_STAGE(2)
        B200UDE_ADJ_STAGE(3)
        B200UDE_ADJ_STAGE(4)
        B200UDE_ADJ_STAGE(5)
#undef B200UDE_ADJ_STAGE
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
    float lsum = loss.x * in.lva + loss.y * in.lvb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);

    // ---- CTA reduction, fixed warp order -> deterministic ----
    float acc[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>(&st->ACC[(c * 32 + lane) * 4]);
        acc[2 * c] = v.x + v.y;       // halves a|b added here
        acc[2 * c + 1] = v.z + v.w;
    }
    const float4 t0 = *reinterpret_cast<const float4 *>(&st->THIN[(0 * 32 + lane) * 4]);
    const float4 t1 = *reinterpret_cast<const float4 *>(&st->THIN[(1 * 32 + lane) * 4]);
    const float4 t2 = *reinterpret_cast<const float4 *>(&st->THIN[(2 * 32 + lane) * 4]);
    const float4 t3 = *reinterpret_cast<const float4 *>(&st->THIN[(3 * 32 + lane) * 4]);
    __syncthreads();
    float *red = reinterpret_cast<float *>(s_raw + sizeof(float) * THETA_SMEM_FLOATS);  // [P+1], aliases the stages
    for (int q = threadIdx.x; q < P + 1; q += BLOCK) red[q] = 0.0f;
    __syncthreads();
    for (int wv = 0; wv < NWARP; ++wv) {
        if (warp == wv) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) red[OFF_W2 + (it * 4 + ii) * H + (jt * 8 + jj)] += acc[jj * 4 + ii];
            red[OFF_W3 + lane * 2 + 0] += t0.x + t0.y;
            red[OFF_W3 + lane * 2 + 1] += t0.z + t0.w;
            red[OFF_B2 + lane] += t2.x + t2.y;
            red[OFF_B1 + lane] += t2.z + t2.w;
            red[OFF_W1 + lane] += t3.x + t3.y;
            red[OFF_W1 + H + lane] += t3.z + t3.w;
            if (lane == 0) {
                red[OFF_B3 + 0] += t1.x + t1.y;   // identical in every lane (sum over the staged rows)
                red[OFF_B3 + 1] += t1.z + t1.w;
                red[P] += lsum;
            }
        }
        __syncthreads();
    }
    float *dst = p.partial + (size_t)blockIdx.x * (P + 1);
    for (int q = threadIdx.x; q < P + 1; q += BLOCK) dst[q] = red[q];
}

}  // namespace lv32
}  // namespace b200ude

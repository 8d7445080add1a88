// k_f64.cu -- double-precision kernels (dtype = B200UDE_F64): launchers of ude_f64.cuh + the fp64 reduction / L2 helpers.
#include "ude_f64.cuh"

#include "params.h"

namespace b200ude {

constexpr int F64_BLOCK = 64;   // two warps per CTA: the per-warp gradient vector (P + 1 doubles) lives in shared memory

int adj_rows_f64(int N) { return ((N + F64_BLOCK - 1) / F64_BLOCK) * (F64_BLOCK / 32); }

static cudaError_t upload_shape(const Shape64Host &h, cudaStream_t st)
{
    // theta's address and the constants are per handle; the struct is small (0.4 KB): re-sent when the handle changes
    static SerialCache cache;
    if (cache.hit(h.serial)) return cudaSuccess;
    f64::Shape64 d;
    d.serial = h.serial; d.model = h.model; d.D = h.D; d.din = h.din; d.dout = h.dout; d.n_layers = h.n_layers; d.n_prefix = h.n_prefix; d.P = h.P;
    int off = h.n_prefix;
    for (int l = 0; l < 6; ++l) d.widths[l] = l <= h.n_layers ? h.widths[l] : 0;
    for (int l = 0; l < 5; ++l) {
        d.acts[l] = l < h.n_layers ? h.acts[l] : 0;
        d.woff[l] = off;
        if (l < h.n_layers) off += h.widths[l] * h.widths[l + 1] + h.widths[l + 1];
    }
    d.theta = h.theta;
    for (int i = 0; i < 16; ++i) { d.consts[i] = h.consts[i]; d.lossw[i] = h.lossw[i]; }
    cudaError_t e = cudaMemcpyToSymbolAsync(f64::c_gen, &d, sizeof(d), 0, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        e = cudaStreamSynchronize(st);   // `d` is a stack object
        cache.set(h.serial);
    }
    return e;
}

static f64::Fwd64 fwd(const Fwd64Host &p)
{
    f64::Fwd64 q;
    q.u0 = p.u0; q.out = p.out; q.ustep = p.ustep; q.dense = p.dense; q.status = p.status; q.N = p.N; q.n_steps = p.n_steps; q.save_every = p.save_every; q.dt = p.dt;
    return q;
}
static f64::Adj64 adj(const Adj64Host &p)
{
    f64::Adj64 q;
    q.ustep = p.ustep; q.dense = p.dense; q.cot = p.cot; q.grad_u0 = p.grad_u0; q.partial = p.partial; q.N = p.N; q.n_steps = p.n_steps;
    q.save_every = p.save_every; q.fused_l2 = p.fused_l2; q.dt = p.dt;
    return q;
}
static f64::Adapt64 adapt(const Adapt64Host &a, bool record)
{
    return f64::Adapt64{a.t0, a.save_dt, a.abstol, a.reltol, a.n_save, a.max_steps, record ? a.tgrid : nullptr, record ? a.nacc : nullptr};
}

cudaError_t launch_fwd_f64(const Shape64Host &s, const Fwd64Host &p, int solver, const Adapt64Host *ad, cudaStream_t st)
{
    cudaError_t e = upload_shape(s, st);
    if (e != cudaSuccess) return e;
    const int grid = (p.N + F64_BLOCK - 1) / F64_BLOCK;
    if (solver == 1) {   // Vern7
        if (ad) f64::vern7_adaptive_forward_kernel<F64_BLOCK><<<grid, F64_BLOCK, 0, st>>>(fwd(p), adapt(*ad, false));
        else f64::vern7_forward_kernel<F64_BLOCK><<<grid, F64_BLOCK, 0, st>>>(fwd(p));
    } else {
        if (ad) f64::adaptive_forward_kernel<F64_BLOCK><<<grid, F64_BLOCK, 0, st>>>(fwd(p), adapt(*ad, true));
        else f64::forward_kernel<F64_BLOCK><<<grid, F64_BLOCK, 0, st>>>(fwd(p));
    }
    return cudaGetLastError();
}

cudaError_t launch_adj_f64(const Shape64Host &s, const Adj64Host &p, const Adapt64Host *ad, cudaStream_t st, int *rows_out)
{
    cudaError_t e = upload_shape(s, st);
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(double) * (size_t)(F64_BLOCK / 32) * (s.P + 1);
    const int grid = (p.N + F64_BLOCK - 1) / F64_BLOCK;
    *rows_out = grid * (F64_BLOCK / 32);
    if (ad) {
        auto kern = f64::adaptive_adjoint_kernel<F64_BLOCK>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, F64_BLOCK, smem, st>>>(adj(p), adapt(*ad, true));
    } else {
        auto kern = f64::adjoint_kernel<F64_BLOCK>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, F64_BLOCK, smem, st>>>(adj(p));
    }
    return cudaGetLastError();
}

// fixed-order sum of the per-warp partial rows (one warp per entry; lane l sums rows l, l + 32, ... then a fixed butterfly)
static __global__ void reduce64_kernel(const double *__restrict__ partial, int nrows, int P1, double *__restrict__ grad, double *__restrict__ loss)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= P1) return;
    double acc = 0.0;
    for (int b = lane; b < nrows; b += 32) acc += partial[(size_t)b * P1 + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        if (q == P1 - 1) {
            if (loss) *loss = acc;
        } else {
            grad[q] = acc;
        }
    }
}
cudaError_t launch_reduce_f64(const double *partial, int nrows, int P1, double *grad, double *loss, cudaStream_t st)
{
    reduce64_kernel<<<(P1 + 7) / 8, 256, 0, st>>>(partial, nrows, P1, grad, loss);
    return cudaGetLastError();
}

// L2 cotangent 2 w (out - data) and loss for the adaptive path (generic cotangent); one CTA, fixed order
struct LossW64 { double w[16]; };
static __global__ void l2_cot64_kernel(const double *__restrict__ out, const double *__restrict__ data, double *__restrict__ cot,
                                       double *__restrict__ loss, LossW64 lw, int D, size_t N, size_t total)
{
    __shared__ double red[1024];
    double acc = 0.0;
    for (size_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int c = (int)((e / N) % (size_t)D);
        const double wgt = D <= 16 ? lw.w[c] : 1.0;
        const double r = out[e] - data[e];
        cot[e] = 2.0 * wgt * r;
        acc = fma(wgt * r, r, acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0 && loss) *loss = red[0];
}
cudaError_t launch_l2_cot_f64(const Shape64Host &s, const double *out, const double *data, double *cot, double *loss, size_t N, int n_save, cudaStream_t st)
{
    LossW64 lw;
    for (int i = 0; i < 16; ++i) lw.w[i] = s.lossw[i];
    l2_cot64_kernel<<<1, 1024, 0, st>>>(out, data, cot, loss, lw, s.D, N, (size_t)n_save * s.D * N);
    return cudaGetLastError();
}

}  // namespace b200ude

// k_generic.cu -- shape-generic UDE kernels (runtime chain widths / activations; LV, SEIR, NODE forms).
#include "ude_generic.cuh"

namespace b200ude {

constexpr int GEN_BLOCK = 128;

int adj_rows_generic(int N) { return ((N + GEN_BLOCK - 1) / GEN_BLOCK) * (GEN_BLOCK / 32); }

static cudaError_t upload_gen(const GenericShape &g, cudaStream_t st)
{
    static SerialCache cache;
    if (cache.hit(g.serial)) return cudaSuccess;
    generic::GenDesc d;
    d.model = g.model; d.D = g.D; d.din = g.din; d.dout = g.dout; d.n_layers = g.n_layers; d.n_prefix = g.n_prefix; d.P = g.P;
    int off = g.n_prefix;
    for (int l = 0; l <= generic::MAXL; ++l) d.widths[l] = l <= g.n_layers ? g.widths[l] : 0;
    for (int l = 0; l < generic::MAXL; ++l) {
        d.acts[l] = l < g.n_layers ? g.acts[l] : 0;
        d.woff[l] = off;
        if (l < g.n_layers) off += g.widths[l] * g.widths[l + 1] + g.widths[l + 1];
    }
    cudaError_t e = cudaMemcpyToSymbolAsync(generic::c_gen, &d, sizeof(d), 0, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) cache.set(g.serial);
    return e;
}

cudaError_t launch_fwd_generic(const GenericShape &g, const ConstTables &t, const FwdParams &p, cudaStream_t st)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(g, st);
    if (e != cudaSuccess) return e;
    generic::forward_kernel<GEN_BLOCK><<<(p.N + GEN_BLOCK - 1) / GEN_BLOCK, GEN_BLOCK, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_adj_generic(const GenericShape &g, const ConstTables &t, const AdjParams &p, cudaStream_t st, int *rows_out)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(g, st);
    if (e != cudaSuccess) return e;
    auto kern = generic::adjoint_kernel<GEN_BLOCK>;
    const size_t smem = sizeof(float) * (size_t)(GEN_BLOCK / 32) * (g.P + 1);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (p.N + GEN_BLOCK - 1) / GEN_BLOCK;
    *rows_out = grid * (GEN_BLOCK / 32);
    kern<<<grid, GEN_BLOCK, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_vern7(const GenericShape &g, const ConstTables &t, const FwdParams &p, cudaStream_t st)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(g, st);
    if (e != cudaSuccess) return e;
    generic::vern7_forward_kernel<GEN_BLOCK><<<(p.N + GEN_BLOCK - 1) / GEN_BLOCK, GEN_BLOCK, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_vern7_adaptive(const GenericShape &g, const ConstTables &t, const FwdParams &p, const AdaptiveGrid &ag, cudaStream_t st)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(g, st);
    if (e != cudaSuccess) return e;
    generic::AdaptParams ap{ag.t0, ag.save_dt, ag.abstol, ag.reltol, ag.n_save, ag.max_steps, nullptr, nullptr};   // no step record
    generic::vern7_adaptive_forward_kernel<GEN_BLOCK><<<(p.N + GEN_BLOCK - 1) / GEN_BLOCK, GEN_BLOCK, 0, st>>>(p, ap);
    return cudaGetLastError();
}

// ---- adaptive stepping ----
cudaError_t launch_fwd_adaptive(const GenericShape &g, const ConstTables &t, const FwdParams &p, const AdaptiveGrid &ag, cudaStream_t st)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(g, st);
    if (e != cudaSuccess) return e;
    generic::AdaptParams ap{ag.t0, ag.save_dt, ag.abstol, ag.reltol, ag.n_save, ag.max_steps, ag.tgrid, ag.nacc};
    generic::adaptive_forward_kernel<GEN_BLOCK><<<(p.N + GEN_BLOCK - 1) / GEN_BLOCK, GEN_BLOCK, 0, st>>>(p, ap);
    return cudaGetLastError();
}

cudaError_t launch_adj_adaptive(const GenericShape &g, const ConstTables &t, const AdjParams &p, const AdaptiveGrid &ag, cudaStream_t st, int *rows_out)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(g, st);
    if (e != cudaSuccess) return e;
    generic::AdaptParams ap{ag.t0, ag.save_dt, ag.abstol, ag.reltol, ag.n_save, ag.max_steps, ag.tgrid, ag.nacc};
    auto kern = generic::adaptive_adjoint_kernel<GEN_BLOCK>;
    const size_t smem = sizeof(float) * (size_t)(GEN_BLOCK / 32) * (g.P + 1);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (p.N + GEN_BLOCK - 1) / GEN_BLOCK;
    *rows_out = grid * (GEN_BLOCK / 32);
    kern<<<grid, GEN_BLOCK, smem, st>>>(p, ap);
    return cudaGetLastError();
}

// L2 cotangent 2 w (out - data) and loss for paths that take a generic cotangent (adaptive solves)
struct LossW { float w[16]; };
static __global__ void l2_cot_kernel(const float *__restrict__ out, const float *__restrict__ data, float *__restrict__ cot,
                                     float *__restrict__ block_loss, LossW lw, int D, size_t N, size_t total)
{
    __shared__ float red[256];
    float acc = 0.0f;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const int c = (int)((e / N) % (size_t)D);
        const float wgt = D <= 16 ? lw.w[c] : 1.0f;
        const float r = out[e] - data[e];
        cot[e] = 2.0f * wgt * r;
        acc = fmaf(wgt * r, r, acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_loss[blockIdx.x] = red[0];
}
static __global__ void l2_finish_kernel(const float *__restrict__ block_loss, int nb, float *__restrict__ loss)
{
    float acc = 0.0f;
    for (int b = 0; b < nb; ++b) acc += block_loss[b];   // fixed order
    *loss = acc;
}
cudaError_t launch_l2_cot(const ConstTables &t, const float *out, const float *data, float *cot, float *block_loss, int D, size_t N, int n_save, cudaStream_t st)
{
    LossW lw;
    for (int i = 0; i < 16; ++i) lw.w[i] = t.lossw[i];
    l2_cot_kernel<<<L2_BLOCKS, 256, 0, st>>>(out, data, cot, block_loss, lw, D, N, (size_t)n_save * D * N);
    return cudaGetLastError();
}
cudaError_t launch_l2_finish(const float *block_loss, float *loss, cudaStream_t st)
{
    l2_finish_kernel<<<1, 1, 0, st>>>(block_loss, L2_BLOCKS, loss);
    return cudaGetLastError();
}

// ---- Fisher-KPP ----
static void fkpp_geom(int Nx, generic::FkppGeom *g, int *threads)
{
    g->Nx = Nx;
    g->tpc = Nx >= 128 ? 1 : 128 / Nx;
    *threads = ((g->tpc * Nx + 31) / 32) * 32;
}

int adj_rows_fkpp(int N, int Nx)
{
    generic::FkppGeom g; int th;
    fkpp_geom(Nx, &g, &th);
    return ((N + g.tpc - 1) / g.tpc) * (th / 32);
}

cudaError_t launch_fwd_fkpp(const GenericShape &gs, const ConstTables &t, const FwdParams &p, cudaStream_t st)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(gs, st);
    if (e != cudaSuccess) return e;
    generic::FkppGeom g; int th;
    fkpp_geom(gs.D, &g, &th);
    const int grid = (p.N + g.tpc - 1) / g.tpc;
    generic::fkpp_forward_kernel<0><<<grid, th, sizeof(float) * g.tpc * g.Nx, st>>>(p, g);
    return cudaGetLastError();
}

cudaError_t launch_adj_fkpp(const GenericShape &gs, const ConstTables &t, const AdjParams &p, cudaStream_t st, int *rows_out)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    e = upload_gen(gs, st);
    if (e != cudaSuccess) return e;
    generic::FkppGeom g; int th;
    fkpp_geom(gs.D, &g, &th);
    const int grid = (p.N + g.tpc - 1) / g.tpc;
    const size_t smem = sizeof(float) * (((g.tpc * g.Nx + 3) / 4) * 4 + (size_t)(th / 32) * (gs.P + 1));
    auto kern = generic::fkpp_adjoint_kernel<0>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    *rows_out = grid * (th / 32);
    kern<<<grid, th, smem, st>>>(p, g);
    return cudaGetLastError();
}

}  // namespace b200ude

// k_lv32_fwd.cu -- forward-solve kernels of the LV 2 -> 32 -> 32 -> 2 tanh UDE (BASELINE config 2).
#include "lv32_packed.cuh"
#include "lv32_tc.cuh"

namespace b200ude {

template <int TM, class W>
static cudaError_t launch_one(const FwdParams &p, cudaStream_t st)
{
    auto kern = lv32::forward_kernel<TM, W, FWD_BLOCK_LV32, 1>;
    const int half = (p.N + 1) / 2;
    const int grid = (half + FWD_BLOCK_LV32 - 1) / FWD_BLOCK_LV32;
    const size_t smem = WeightStage<W>::kSmem ? sizeof(float) * ((lv32::P + 3) / 4) * 4 : 0;
    kern<<<grid, FWD_BLOCK_LV32, smem, st>>>(p);
    return cudaGetLastError();
}

template <int TM>
static cudaError_t launch_tc(const FwdParams &p, cudaStream_t st)
{
    auto kern = lv32::tc::forward_kernel<TM, 128, 4>;
    const int grid = (p.N + 127) / 128;
    kern<<<grid, 128, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_lv32(const Variant &v, const ConstTables &t, const FwdParams &p, cudaStream_t st)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    if (v.fwd_tc) return v.approx_tanh ? launch_tc<1>(p, st) : launch_tc<0>(p, st);
    if (v.approx_tanh) return launch_one<1, WConst>(p, st);
    if (v.fwd_smem) return launch_one<0, WSmem>(p, st);
    return launch_one<0, WConst>(p, st);
}

cudaError_t launch_fwd_lv32_adaptive(const Variant &v, const ConstTables &t, const FwdParams &p, const AdaptiveGrid &ag, cudaStream_t st)
{
    cudaError_t e = upload_tables(t, st);
    if (e != cudaSuccess) return e;
    const int grid = (p.N + 127) / 128;
    if (v.approx_tanh) lv32::tc::adaptive_forward_kernel<1, 128, 4><<<grid, 128, 0, st>>>(p, ag);
    else lv32::tc::adaptive_forward_kernel<0, 128, 4><<<grid, 128, 0, st>>>(p, ag);
    return cudaGetLastError();
}

}  // namespace b200ude

// k_lv32_wm.cu -- warp-collective (mma.sync, 16 trajectories per warp) kernels of the LV 2 -> 32 -> 32 -> 2 tanh UDE:
// the low-latency family for small and medium ensembles (lv32_wm.cuh).
#include "lv32_wm.cuh"

namespace b200ude {

int adj_rows_lv32_wm(int N) { return (N + 15) / 16; }

static lv32::wm::Consts make_consts(const ConstTables &t)
{
    lv32::wm::Consts c;
    c.p1 = t.consts[0]; c.p4 = t.consts[1]; c.lw0 = t.lossw[0]; c.lw1 = t.lossw[1];
    return c;
}

template <int TM, int G>
static cudaError_t launch_fwd(const ConstTables &t, const FwdParams &p, cudaStream_t st)
{
    constexpr int WPC = 1;
    const int warps = (p.N + 16 * G - 1) / (16 * G);
    lv32::wm::forward_kernel<TM, G, WPC><<<(warps + WPC - 1) / WPC, 32 * WPC, 0, st>>>(p, make_consts(t));
    return cudaGetLastError();
}

cudaError_t launch_fwd_lv32_wm(const Variant &v, const ConstTables &t, const FwdParams &p, cudaStream_t st)
{
    const int G = v.wm_groups;
    if (v.approx_tanh) return G >= 4 ? launch_fwd<1, 4>(t, p, st) : G == 2 ? launch_fwd<1, 2>(t, p, st) : launch_fwd<1, 1>(t, p, st);
    return G >= 4 ? launch_fwd<0, 4>(t, p, st) : G == 2 ? launch_fwd<0, 2>(t, p, st) : launch_fwd<0, 1>(t, p, st);
}

template <int TM>
static cudaError_t launch_adj(const ConstTables &t, const AdjParams &p, int *rows_out, cudaStream_t st)
{
    constexpr int WPC = 1;
    auto kern = lv32::wm::adjoint_kernel<TM, WPC>;
    constexpr size_t smem = sizeof(lv32::wm::WarpTiles) * WPC;
    static PerDeviceOnce once;
    bool &attr_set = once.flag();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int warps = adj_rows_lv32_wm(p.N);
    *rows_out = warps;
    kern<<<(warps + WPC - 1) / WPC, 32 * WPC, smem, st>>>(p, make_consts(t));
    return cudaGetLastError();
}

cudaError_t launch_adj_lv32_wm(const Variant &v, const ConstTables &t, const AdjParams &p, cudaStream_t st, int *rows_out)
{
    return v.approx_tanh ? launch_adj<1>(t, p, rows_out, st) : launch_adj<0>(t, p, rows_out, st);
}

}  // namespace b200ude

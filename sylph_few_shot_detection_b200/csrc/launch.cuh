// Kernel launch helper: every kernel of the engine is launched with programmatic dependent launch (PDL) so that the
// next kernel's CTAs are scheduled, run their prologue (barrier init, TMEM allocation, tensor-map prefetch) and park in
// `griddepcontrol.wait` while the previous kernel drains.  Contract: every kernel launched through `launch_k` executes
// `ptx::griddep_launch()` early and `ptx::griddep_wait()` BEFORE its first global-memory access that can depend on (or
// be depended on by) the previous kernel in the stream.  SYLPH_PDL=0 turns the attribute off (plain stream order).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace sylph {

namespace ptx {
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace ptx

inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SYLPH_PDL");
        v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
    }
    return v == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace sylph

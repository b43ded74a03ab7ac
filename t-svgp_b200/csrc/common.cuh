// Shared device helpers for the t-SVGP sm_100a kernels: FP64 tensor-core MMA (DMMA), cp.async staging.
// FP64 has no tcgen05/TMEM path on Blackwell; the FP64 tensor instruction is mma.sync m8n8k4 (SASS DMMA.8x8x4),
// measured at 37.1 TFLOP/s on B200 (profiles/fp64_peak_r01.txt), equal to and sharing a pipe with DFMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TSVGP_TILE 128  // all matrix dimensions inside the library are padded to a multiple of this

namespace tsvgp {

// kernels launched by this host thread (reported by tsvgp_get_timings; the bench's gpu_launches)
extern thread_local long g_launches;
extern int g_debug_sync;   // TSVGP_DEBUG_SYNC=1 : synchronise after every launch so a faulting kernel is reported at its call site
inline int count_launch_at(const char* where) {
    ++g_launches;
    if (g_debug_sync) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            fprintf(stderr, "[tsvgp] launch #%ld in %s: %s\n", g_launches, where, cudaGetErrorString(e));
            return (int)e;
        }
    }
    return (int)cudaGetLastError();
}
#define count_launch() count_launch_at(__func__)

// D(8x8) += A(8x4,row) * B(4x8,col).  lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c0,c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier (shared::cta) -----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// arrive on `bar` once every cp.async this thread issued so far has landed (does not change the expected count)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity)
        : "memory");
}

// ---- TMA bulk copy (global -> shared, 1-D), completion counted in bytes on an mbarrier ---------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

// ---- programmatic dependent launch (sm_90+) --------------------------------------------------------------------------------
// The M x M phases are chains of ~100 dependent small kernels.  A kernel launched with the programmatic-serialisation attribute may
// be dispatched (CTAs resident, parameters loaded) as soon as every CTA of its predecessor has started, and blocks in
// griddepcontrol.wait until the predecessor has completed and its memory is visible: the launch latency leaves the critical path.
// Every kernel launched through launch_k(pdl = true, ...) MUST run PDL_PROLOGUE() before it touches global memory.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#define PDL_PROLOGUE() pdl_prologue()

extern int g_pdl;   // 1 = use programmatic dependent launch (default), 0 = plain stream order (env TSVGP_PDL=0; A/B timing)
// set by a host thread while it enqueues two concurrent chains of LARGE kernels (prepare phase at M >= 2048): there the CTAs of a
// pre-dispatched, still waiting kernel of one chain take SM slots from the running GEMM of the other (measured: +0.16 ms at M = 2048)
extern thread_local int g_pdl_suspended;

template <typename... KArgs, typename... Args>
inline void launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl && g_pdl && !g_pdl_suspended) ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through count_launch() / cudaGetLastError
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace tsvgp

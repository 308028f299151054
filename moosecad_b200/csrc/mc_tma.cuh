// TMA (cp.async.bulk.tensor) staging of lattice boxes into shared memory, sm_100a.
//
// The per-point arrays are stored x fastest with a row pitch of L.PX points (a multiple of 16, so rows
// are 128-byte aligned), z slowest: a 3-D tensor {NX, NY, planes}.  A tile kernel brings the H x H x HZ
// box of doubles around its tile into shared memory with ONE bulk tensor copy issued by one thread —
// no per-point address arithmetic, out-of-range coordinates (lattice border, negative indices) are
// zero-filled by the hardware — and waits for it on an mbarrier.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace mc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses to a shared buffer are ordered before the async-proxy (TMA) writes that follow
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one box of the tensor -> shared memory; completion (byte count) is signalled on `bar`
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"((unsigned long long)tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

}  // namespace mc

// tma.cuh — the few PTX wrappers the source-tile staging needs on sm_100a: one mbarrier per CTA and
// cp.async.bulk.tensor.2d (SASS: UTMALDG) global -> shared loads that complete on it.
//
// The tensor maps are encoded on the host (cuTensorMapEncodeTiled, see hgwarp.cu: encode_tmaps) over the
// source image as a 2-D tensor of 32-bit pixels [H][W]; elements outside the image are zero-filled by the
// TMA unit, which is exactly what an out-of-image nearest-neighbour read must produce (transparent black).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace hg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// make the barrier initialisation visible to the async proxy (the TMA unit) before a copy signals it
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, unsigned parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// blocks until the phase with the given parity has completed; the waiting thread then sees the bytes the
// async proxy wrote.  A wait that lasts seconds is a protocol bug, never load: it traps (the launch then fails with
// an error the host reports) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity, int site = 0)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
            printf("hgwarp: mbarrier wait timed out (site %d, block %d, thread %d, parity %u)\n", site, (int)blockIdx.x,
                   (int)threadIdx.x, parity);
            __trap();
        }
    }
}

// one box of the tensor map `tm` whose first element is (x, y) -> shared memory at `dst`; signals `bar`
// with the box's byte count (elements outside the tensor arrive as zeros and are counted too)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tm), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}

}  // namespace hg

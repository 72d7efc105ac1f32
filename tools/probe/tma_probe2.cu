// probe 2: the staged kernel's box stack (32-row boxes then 8-row boxes on one mbarrier), descriptors by value
// usage: tma_probe2 <W> <H> <bx0> <by0> <sel> <nstrips>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../homography.js_b200/csrc/tma.cuh"
using namespace hg;
static __host__ __device__ int box_w(int i) { return i == 0 ? 72 : i == 1 ? 80 : i == 2 ? 96 : i == 3 ? 112 : i == 4 ? 128 : i == 5 ? 160 : i == 6 ? 192 : 256; }
struct Params { int bx0, by0, sel, nstrips, pad[12]; CUtensorMap tm[16]; };
__global__ void k(const __grid_constant__ Params P, uint32_t *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const int pitch = box_w(P.sel);
    if (threadIdx.x == 0) {
        const uint32_t b = smem_u32(&bar);
        mbar_init(b, 1);
        fence_barrier_init();
        const unsigned strip_bytes = 8u * pitch * 4u;
        mbar_arrive_expect_tx(b, strip_bytes * P.nstrips);
        uint32_t dst = smem_u32(sm);
        int y = P.by0, left = P.nstrips;
        for (; left >= 4; left -= 4) { tma_load_2d(dst, P.tm + 8 + P.sel, P.bx0, y, b); dst += 4 * strip_bytes; y += 32; }
        for (; left > 0; --left) { tma_load_2d(dst, P.tm + P.sel, P.bx0, y, b); dst += strip_bytes; y += 8; }
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar), 0, 9);
    const uint32_t *s = (const uint32_t *)sm;
    for (int i = threadIdx.x; i < pitch * 8 * P.nstrips; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    const int W = atoi(argv[1]), H = atoi(argv[2]);
    Params P{}; P.bx0 = atoi(argv[3]); P.by0 = atoi(argv[4]); P.sel = atoi(argv[5]); P.nstrips = atoi(argv[6]);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    std::vector<uint32_t> h((size_t)W * H);
    for (int i = 0; i < W * H; ++i) h[i] = 0x10000u * (i / W) + (i % W) + 1;
    uint32_t *d, *o; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int pitch = box_w(P.sel), n = pitch * 8 * P.nstrips;
    cudaMalloc(&o, n * 4);
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, str[1] = {(cuuint64_t)W * 4}; cuuint32_t es[2] = {1, 1};
    for (int i = 0; i < 16; ++i) {
        cuuint32_t box[2] = {(cuuint32_t)box_w(i % 8), (cuuint32_t)(i < 8 ? 8 : 32)};
        CUresult r = ((enc_fn)fn)(&P.tm[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode %d failed %d\n", i, (int)r); return 0; }
    }
    k<<<1, 128, n * 4>>>(P, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("W=%d H=%d box (%d,%d) sel %d strips %d: ", W, H, P.bx0, P.by0, P.sel, P.nstrips);
    if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 0; }
    std::vector<uint32_t> res((size_t)n); cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; ++i) { int x = P.bx0 + i % pitch, y = P.by0 + i / pitch; uint32_t want = (x >= 0 && x < W && y >= 0 && y < H) ? 0x10000u * y + x + 1 : 0; bad += res[i] != want; }
    printf("ok, mismatches=%d\n", bad);
    return 0;
}

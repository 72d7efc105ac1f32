// standalone probe: which tensor-map / box / descriptor-location variants does UTMALDG accept on this part?
// usage: tma_probe <boxw> <boxh> <where: 0 param, 1 global> <x0> <y0> <W>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../homography.js_b200/csrc/tma.cuh"
using namespace hg;
struct Params { int x0, y0, boxw, boxh, n; const CUtensorMap *g; CUtensorMap tm; };
__global__ void k(const __grid_constant__ Params P, uint32_t *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        const CUtensorMap *tm = P.g ? P.g : &P.tm;
        const uint32_t b = smem_u32(&bar);
        mbar_init(b, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(b, (unsigned)(P.boxw * P.boxh * 4 * P.n));
        for (int i = 0; i < P.n; ++i) tma_load_2d(smem_u32(sm) + i * P.boxw * P.boxh * 4, tm, P.x0, P.y0 + i * P.boxh, b);
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar), 0);
    const uint32_t *s = (const uint32_t *)sm;
    for (int i = threadIdx.x; i < P.boxw * P.boxh * P.n; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    const int boxw = atoi(argv[1]), boxh = atoi(argv[2]), where = atoi(argv[3]), x0 = atoi(argv[4]), y0 = atoi(argv[5]), W = atoi(argv[6]), H = 64, n = 2;
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    std::vector<uint32_t> h((size_t)W * H);
    for (int i = 0; i < W * H; ++i) h[i] = 0x10000u * (i / W) + (i % W) + 1;
    uint32_t *d, *o; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, boxw * boxh * n * 4);
    Params P{}; P.x0 = x0; P.y0 = y0; P.boxw = boxw; P.boxh = boxh; P.n = n;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, str[1] = {(cuuint64_t)W * 4}; cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh}, es[2] = {1, 1};
    CUresult r = ((enc_fn)fn)(&P.tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box %dx%d W=%d: encode failed %d\n", boxw, boxh, W, (int)r); return 0; }
    if (where == 1) { CUtensorMap *g; cudaMalloc(&g, sizeof(CUtensorMap)); cudaMemcpy(g, &P.tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice); P.g = g; }
    k<<<1, 128, boxw * boxh * n * 4>>>(P, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("box %dx%d where=%d at (%d,%d) W=%d: %s\n", boxw, boxh, where, x0, y0, W, cudaGetErrorString(e)); return 0; }
    std::vector<uint32_t> res((size_t)boxw * boxh * n); cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < boxw * boxh * n; ++i) { int x = x0 + i % boxw, y = y0 + i / boxw; uint32_t want = (x >= 0 && x < W && y >= 0 && y < H) ? 0x10000u * y + x + 1 : 0; bad += res[i] != want; }
    printf("box %dx%d where=%d at (%d,%d) W=%d: ok, mismatches=%d\n", boxw, boxh, where, x0, y0, W, bad);
    return 0;
}

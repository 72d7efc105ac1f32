// diag.cuh — device self-checks of the numerical building blocks (called from tests/test_gpu_numerics.py).
#pragma once
#include "warp_geo.cuh"

namespace hg {

// Exhaustive over the 2^20 high-mantissa patterns that MUFU.RCP64H reads, each with several low words:
// residual |1 - d * rcp_newton1(d)| evaluated with one FMA (exact up to 2^-53).  The maximum is the
// relative error bound the projective fast path relies on (needs < 2^-36, see warp_geo.cuh).
__global__ void rcp_error_kernel(int biased_exp, int negative, unsigned long long *max_bits)
{
    const unsigned hi20 = blockIdx.x * blockDim.x + threadIdx.x;
    if (hi20 >= (1u << 20)) return;
    const unsigned lows[6] = {0u, 0xFFFFFFFFu, 0x80000000u, 0x00000001u, hi20 * 2654435761u, ~(hi20 * 40503u)};
    double worst = 0.0;
    for (int i = 0; i < 6; ++i) {
        const unsigned hi = (negative ? 0x80000000u : 0u) | ((unsigned)biased_exp << 20) | hi20;
        const double d = __hiloint2double((int)hi, (int)lows[i]);
        const double rc = rcp_newton1(d);
        const double res = fabs(__fma_rn(-d, rc, 1.0));
        if (!(res <= worst)) worst = res;  // NaN propagates as "worst"
    }
    atomicMax(max_bits, (unsigned long long)__double_as_longlong(worst));
}

// quotient_at_least / exact_doubled_floor (warp_geo.cuh) on caller-supplied operands: out[i] = 1 iff the device decides
// RN(N[i] / D[i]) >= b[i]; tests/test_gpu_numerics.py compares with exact rational arithmetic.
__global__ void quotient_decision_kernel(const double *N, const double *D, const double *b, int n, int *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = quotient_at_least(N[i], D[i], b[i]) ? 1 : 0;
}

}  // namespace hg

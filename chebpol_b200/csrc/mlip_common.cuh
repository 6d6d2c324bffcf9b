// mlip_common.cuh -- pieces shared by the multilinear kernels (mlip_pair, mlip_cell, mlip_quad).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mlip {

// blendfun, src/chebpol.h:24-45
__device__ __forceinline__ double blend_w(double w, int blend)
{
    switch (blend) {
    case 1: return w < 0.5 ? 0.5 * exp(2 - 1 / w) : 1 - 0.5 * exp(2 - 1 / (1 - w));
    case 2: return w < 0.5 ? 0.5 * exp(4 - 1 / (w * w)) : 1 - 0.5 * exp(4 - 1 / ((1 - w) * (1 - w)));
    case 3: return (-2 * w + 3) * w * w;
    case 4: return (w < 0.5) ? 0.0 : 1.0;
    case 5: return 0.5;
    default: return w;
    }
}

// Cell search: upper knot index in [1, n-1], identical to the reference's bisection
// (src/chebpol.c:543-556: smallest index with g[idx] >= v, clamped) for every sorted grid.
// The bucket table gives a starting index; the two fix-up scans use the real knots, so the
// result does not depend on the table's rounding.  NaN terminates (the reference's loop does
// not, :549-553).  Also returns the cell's two knots.
__device__ __forceinline__ int cell_search(const double *g, int n, const int *lut, int nb, double scale, double v,
                                           double &g0, double &g1)
{
    double t = (v - g[0]) * scale;
    t = t > 0.0 ? t : 0.0;  // NaN -> 0
    const double tmax = (double)(nb - 1);
    t = t < tmax ? t : tmax;
    int k = lut[(int)t];
    double hi = g[k], lo = g[k - 1];
    while (k < n - 1 && hi < v) {
        ++k;
        lo = hi;
        hi = g[k];
    }
    while (k > 1 && lo >= v) {
        --k;
        hi = lo;
        lo = g[k - 1];
    }
    g0 = lo;
    g1 = hi;
    return k;
}

}  // namespace mlip

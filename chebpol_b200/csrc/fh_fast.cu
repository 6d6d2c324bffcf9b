// fh_fast.cu -- fast Floater-Hormann evaluation (placeholder until the
// barycentric-basis contraction kernel lands; the strict kernel serves FH).
#include "common.cuh"

bool fh_fast_supported(int rank, const int *dims)
{
    (void)rank; (void)dims;
    return false;
}

cudaError_t launch_fh_fast(const FhArgs &a, int variant, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    (void)a; (void)variant; (void)num_sms; (void)s; (void)li;
    return cudaErrorNotSupported;
}

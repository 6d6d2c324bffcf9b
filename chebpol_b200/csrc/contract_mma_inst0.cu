// contract_mma_inst0.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_3_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_3_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 2, 8>(a, num_sms, s, li);
}

// contract_mma_inst4.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_3_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 3, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_20_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<20, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 1, 8>(a, num_sms, s, li);
}

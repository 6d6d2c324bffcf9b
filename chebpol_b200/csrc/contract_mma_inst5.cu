// contract_mma_inst5.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_4_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_20_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<20, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_20_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<20, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_1_16(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 1, 16>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_1_16(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 1, 16>(a, num_sms, s, li);
}
cudaError_t mma_launch_48_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<48, 1, 4>(a, num_sms, s, li);
}

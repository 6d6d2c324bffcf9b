// contract_mma_inst7.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 2, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 2, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 1, 4, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 2, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_4_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 4, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 1, 4, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_3_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 3, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_20_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<20, 2, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_1_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 1, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_1_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 1, 8, 4>(a, num_sms, s, li);
}

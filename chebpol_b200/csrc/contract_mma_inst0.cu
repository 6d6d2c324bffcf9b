// contract_mma_inst0.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 2, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_2_4_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 4, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_1_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 1, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_4_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 4, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 1, 4, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_4_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 4, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 2, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 4, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_1_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 1, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 2, 8, 2>(a, num_sms, s, li);
}

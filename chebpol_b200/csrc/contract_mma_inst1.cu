// contract_mma_inst1.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_3_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_3_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 4, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 2, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_4_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 4, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 2, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_20_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<20, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_1_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 1, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 1, 4, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 2, 8, 4>(a, num_sms, s, li);
}

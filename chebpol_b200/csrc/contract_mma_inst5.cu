// contract_mma_inst5.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_2_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 4, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 2, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 4, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_4_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 4, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_4_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 4, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 2, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 1, 4, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_40_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<40, 1, 4, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_48_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<48, 1, 4, 4>(a, num_sms, s, li);
}

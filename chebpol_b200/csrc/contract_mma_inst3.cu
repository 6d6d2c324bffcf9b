// contract_mma_inst3.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_3_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_3_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 2, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_2_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 2, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_20_1_8_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<20, 1, 8, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_1_4_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 1, 4, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_1_4_2(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 1, 4, 2>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_2_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 2, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_40_1_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<40, 1, 8, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_48_1_8_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<48, 1, 8, 4>(a, num_sms, s, li);
}

// contract_mma_inst2.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_3_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_2_16(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 2, 16>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_36_1_16(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<36, 1, 16>(a, num_sms, s, li);
}
cudaError_t mma_launch_40_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<40, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_48_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<48, 1, 8>(a, num_sms, s, li);
}

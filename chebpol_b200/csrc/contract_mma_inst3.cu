// contract_mma_inst3.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_3_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_6_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<6, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_48_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<48, 1, 8>(a, num_sms, s, li);
}

// contract_mma_inst2.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_3_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<3, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_4_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 4, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_10_2_16(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<10, 2, 16>(a, num_sms, s, li);
}
cudaError_t mma_launch_16_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<16, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_25_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<25, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_40_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<40, 1, 4>(a, num_sms, s, li);
}

// contract_mma_inst1.cu -- explicit instantiations of the DMMA contraction kernel (generated list;
// see contract_mma.cuh for the kernel and contract_mma.cu for the dispatcher).
#include "contract_mma.cuh"

cudaError_t mma_launch_2_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<2, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_4_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<4, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_8_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<8, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 1, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_12_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<12, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_1_4(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 1, 4>(a, num_sms, s, li);
}
cudaError_t mma_launch_30_2_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<30, 2, 8>(a, num_sms, s, li);
}
cudaError_t mma_launch_40_1_8(const MmaArgs &a, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    return mma_detail::launch_one<40, 1, 8>(a, num_sms, s, li);
}

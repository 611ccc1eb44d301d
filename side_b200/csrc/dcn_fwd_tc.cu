// dcn_fwd_tc.cu -- tcgen05 / TMEM path of the DCNv2 forward (placeholder until the kernel lands).
#include "dcn_common.cuh"

namespace side {
struct DcnFwdArgs;
size_t dcn_fwd_tc_ws_bytes(int Cin, int Cout, int KK, int flags)
{
    (void)flags;
    return sizeof(float) * 2 * (size_t)Cin * Cout * KK;
}
int dcn_fwd_tc(const DcnFwdArgs &, const float *, void *, size_t, cudaStream_t)
{
    set_error("side_dcn_fwd: tcgen05 path not built in this version");
    return SIDE_ERR_UNSUPPORTED;
}
}  // namespace side

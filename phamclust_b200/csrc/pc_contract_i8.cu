// pc_contract_i8.cu -- int8 tcgen05 engine for the shared-count contraction.
// (placeholder until the TMA/TMEM kernel lands; the popc engine is the default)
#include "pc_common.cuh"

int pc_i8_contract(pc_ctx *ctx, int metric, const EpilogueArgs &ep)
{
    (void)metric; (void)ep;
    return pc_fail(ctx, PC_EINVAL, "the int8 tcgen05 engine is not available in this build");
}

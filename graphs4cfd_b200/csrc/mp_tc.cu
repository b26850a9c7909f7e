// mp_tc.cu — tcgen05 tensor-core message-passing kernel (placeholder until the kernel lands).
#include "common.cuh"
namespace g4c {
int mp_tc_dispatch(const G4cMpDesc& d, cudaStream_t) {
    set_error("g4c_mp_fwd: precision=%d not built yet", d.precision);
    return G4C_EUNSUPPORTED;
}
}  // namespace g4c

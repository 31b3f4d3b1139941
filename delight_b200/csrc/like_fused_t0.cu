// Instantiations of the fused likelihood kernel for 1..8 bands per target
// (split over four translation units so that they compile in parallel).
#include "like_fused_kernel.cuh"

namespace dl {

int launch_fused_tu0(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s,
                     bool timed) {
    switch (NF) {
        case 1: return launch_fused_nf<1>(zdep, het, mode, p, s, timed);
        case 2: return launch_fused_nf<2>(zdep, het, mode, p, s, timed);
        case 3: return launch_fused_nf<3>(zdep, het, mode, p, s, timed);
        case 4: return launch_fused_nf<4>(zdep, het, mode, p, s, timed);
        case 5: return launch_fused_nf<5>(zdep, het, mode, p, s, timed);
        case 6: return launch_fused_nf<6>(zdep, het, mode, p, s, timed);
        case 7: return launch_fused_nf<7>(zdep, het, mode, p, s, timed);
        case 8: return launch_fused_nf<8>(zdep, het, mode, p, s, timed);
    }
    dl_set_error("launch_fused_tu0: %d bands are not in this unit", NF);
    return DL_ERR_UNSUPPORTED;
}

}  // namespace dl

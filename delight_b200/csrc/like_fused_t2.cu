// Instantiations of the fused likelihood kernel for 17..24 bands per target
// (split over four translation units so that they compile in parallel).
#include "like_fused_kernel.cuh"

namespace dl {

int launch_fused_tu2(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s,
                     bool timed) {
    switch (NF) {
        case 17: return launch_fused_nf<17>(zdep, het, mode, p, s, timed);
        case 18: return launch_fused_nf<18>(zdep, het, mode, p, s, timed);
        case 19: return launch_fused_nf<19>(zdep, het, mode, p, s, timed);
        case 20: return launch_fused_nf<20>(zdep, het, mode, p, s, timed);
        case 21: return launch_fused_nf<21>(zdep, het, mode, p, s, timed);
        case 22: return launch_fused_nf<22>(zdep, het, mode, p, s, timed);
        case 23: return launch_fused_nf<23>(zdep, het, mode, p, s, timed);
        case 24: return launch_fused_nf<24>(zdep, het, mode, p, s, timed);
    }
    dl_set_error("launch_fused_tu2: %d bands are not in this unit", NF);
    return DL_ERR_UNSUPPORTED;
}

}  // namespace dl

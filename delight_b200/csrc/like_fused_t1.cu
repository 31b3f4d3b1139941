// Instantiations of the fused likelihood kernel for 9..16 bands per target
// (split over four translation units so that they compile in parallel).
#include "like_fused_kernel.cuh"

namespace dl {

int launch_fused_tu1(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s,
                     bool timed) {
    switch (NF) {
        case 9: return launch_fused_nf<9>(zdep, het, mode, p, s, timed);
        case 10: return launch_fused_nf<10>(zdep, het, mode, p, s, timed);
        case 11: return launch_fused_nf<11>(zdep, het, mode, p, s, timed);
        case 12: return launch_fused_nf<12>(zdep, het, mode, p, s, timed);
        case 13: return launch_fused_nf<13>(zdep, het, mode, p, s, timed);
        case 14: return launch_fused_nf<14>(zdep, het, mode, p, s, timed);
        case 15: return launch_fused_nf<15>(zdep, het, mode, p, s, timed);
        case 16: return launch_fused_nf<16>(zdep, het, mode, p, s, timed);
    }
    dl_set_error("launch_fused_tu1: %d bands are not in this unit", NF);
    return DL_ERR_UNSUPPORTED;
}

}  // namespace dl

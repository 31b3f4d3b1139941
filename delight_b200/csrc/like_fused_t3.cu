// Instantiations of the fused likelihood kernel for 25..32 bands per target
// (split over four translation units so that they compile in parallel).
#include "like_fused_kernel.cuh"

namespace dl {

int launch_fused_tu3(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s,
                     bool timed) {
    switch (NF) {
        case 25: return launch_fused_nf<25>(zdep, het, mode, p, s, timed);
        case 26: return launch_fused_nf<26>(zdep, het, mode, p, s, timed);
        case 27: return launch_fused_nf<27>(zdep, het, mode, p, s, timed);
        case 28: return launch_fused_nf<28>(zdep, het, mode, p, s, timed);
        case 29: return launch_fused_nf<29>(zdep, het, mode, p, s, timed);
        case 30: return launch_fused_nf<30>(zdep, het, mode, p, s, timed);
        case 31: return launch_fused_nf<31>(zdep, het, mode, p, s, timed);
        case 32: return launch_fused_nf<32>(zdep, het, mode, p, s, timed);
    }
    dl_set_error("launch_fused_tu3: %d bands are not in this unit", NF);
    return DL_ERR_UNSUPPORTED;
}

}  // namespace dl

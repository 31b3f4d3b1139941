// Fused photo-z likelihood (subsystem 1): host side of the C ABI and the small kernels around
// the main one (like_fused_kernel.cuh; instantiated for 1..32 bands in like_fused_t*.cu).
//
//   dl_like_fused_dev = zaux + prep_targets                      per-target constants
//                     + fused<SAMPLE> + anchor_fold              anchor exponent per target
//                     + fused<MAIN>                              PDFs, evidences, max(v)
//                     + [anchor_reset + fused<SAMPLE, all t> + anchor_fold + fused<MAIN>]
//                                                                rescue of flagged targets; these
//                                                                launches exit at once when no
//                                                                target was flagged
//                     + finalize_scale / finalize_pdf / top-K
#include "like_fused_kernel.cuh"
#include <cstdlib>

namespace dl {

int launch_fused_tu0(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s, bool timed);
int launch_fused_tu1(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s, bool timed);
int launch_fused_tu2(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s, bool timed);
int launch_fused_tu3(int NF, bool zdep, bool het, int mode, const FusedParams& p, cudaStream_t s, bool timed);

static int launch_fused(int NF, bool zdep, bool het, int mode, const FusedParams& p,
                        cudaStream_t s, bool timed) {
    if (NF < 1 || NF > MAXNF) {
        dl_set_error("fused likelihood supports 1..%d bands per target, got %d", MAXNF, NF);
        return DL_ERR_UNSUPPORTED;
    }
    switch ((NF - 1) / 8) {
        case 0: return launch_fused_tu0(NF, zdep, het, mode, p, s, timed);
        case 1: return launch_fused_tu1(NF, zdep, het, mode, p, s, timed);
        case 2: return launch_fused_tu2(NF, zdep, het, mode, p, s, timed);
        default: return launch_fused_tu3(NF, zdep, het, mode, p, s, timed);
    }
}

// ------------------------------------------------------------------------------------------
// per-target constants (layout: like_fused_kernel.cuh), variances scaled by the power of two
// c = 2^round(mean log2 var) (exact) so that products over bands stay in range.
// in:  obs/var [nO][NF] (restricted to the target's NF bands), ell_hat/ell_var scalars,
//      rows [nO][NF] int32 (per-target band lists: index of each band in the staged union) or NULL
__global__ void prep_targets_kernel(int nO, int NF, int TCS, int fzc,
                                    const double* __restrict__ obs, const double* __restrict__ var,
                                    const int* __restrict__ rows, double ell_hat, double ell_var,
                                    double cov_scale, double* __restrict__ tconst) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nO) return;
    double* tc = tconst + (size_t)o * TCS;
    double s = 0.0;
    for (int f = 0; f < NF; ++f) s += log2(var[(size_t)o * NF + f]);
    int k = (int)rint(s / NF);
    k = k < -1000 ? -1000 : (k > 1000 ? 1000 : k);
#if DL_HCFOLD
    k |= 1;                                    // odd: h = sqrt(1 / (2 c)) is a power of two
    const double h = ldexp(1.0, -(k + 1) / 2);
#endif
    const double c = ldexp(1.0, k), ic = ldexp(1.0, -k);
    const double sic = sqrt(ic * cov_scale);   // ell^2 * (cov_scale * model variance) / c
    for (int f = 0; f < NF; ++f) {
        const double ob = obs[(size_t)o * NF + f];
        const double vs = var[(size_t)o * NF + f] * ic;
        const double w = 1.0 / vs;
        tc[2 * f] = ob * w * sic;       // pass-0 weights carry sqrt(ic cov_scale): e2 = (FOT'/FTT)^2
        tc[2 * f + 1] = w;
        tc[2 * NF + 2 * f] = vs;
#if DL_HCFOLD
        tc[2 * NF + 2 * f + 1] = ob * h;
        tc[4 * NF + f] = -(ob * h) * (ob * h);
#else
        tc[2 * NF + 2 * f + 1] = ob;
        tc[4 * NF + f] = ob * ob;
#endif
    }
    double* tk = tc + 5 * NF;
    tk[TC_HC] = -0.5 * ic;
    tk[TC_POT0] = c * (ell_hat / ell_var) * sic;      // pass-0 start value of FOT'
    tk[TC_POT] = c * (ell_hat / ell_var);
    tk[TC_PTT] = c / ell_var;
    tk[TC_POO] = tk[TC_POO0] = c * (ell_hat * ell_hat / ell_var);
    tk[TC_C] = c;
    tk[TC_CS] = c * sic;                              // z-dependent ell priors: pOT0 = cs * zot
    tk[TC_XOFF] = 0.0;
    tk[TC_SPARE] = 0.0;
#if DL_HCFOLD
    tk[TC_POT] *= h;
    tk[TC_SPARE] = c * h;                             // z-dependent ell priors: pOT = (c h) * zot
    tk[TC_POO] = tk[TC_POO0] = -0.5 * (ell_hat * ell_hat / ell_var);     // (c h_c = -1/2)
#endif
    if (rows != nullptr) {
        int* off = reinterpret_cast<int*>(tk + TC_NCONST);
        for (int f = 0; f < NF; ++f) off[f] = rows[(size_t)o * NF + f] * fzc;
        if (NF & 1) off[NF] = 0;
    }
    for (int i = 5 * NF + TC_NCONST + (rows ? (NF + 1) / 2 : 0); i < TCS; ++i) tc[i] = 0.0;
}

__device__ __forceinline__ double key_to_double(unsigned long long k) {
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    return __longlong_as_double((long long)b);
}

// Anchor of every target from the sampled max (first call), or, in the rescue passes, the exact
// max for the flagged targets (whose accumulators are reset) and a "contributes nothing" mark
// (x = -inf) for the others.  One block per target.
__global__ void anchor_fold_kernel(int nO, int NF, int TCS, double* __restrict__ tconst,
                                   const unsigned long long* __restrict__ Akey,
                                   const int* __restrict__ flag, const int* __restrict__ nflag,
                                   int rescue, double* __restrict__ pdf_acc, int nzp,
                                   unsigned long long* __restrict__ Vbits) {
    const int o = blockIdx.x;
    if (rescue && *nflag == 0) return;
    const bool mine = rescue ? flag[o] != 0 : true;
    double* tk = tconst + (size_t)o * TCS + 5 * NF;
    if (threadIdx.x == 0) {
        if (mine) {
            double A = Akey[o] ? key_to_double(Akey[o]) : 0.0;
            if (!(fabs(A) <= 1.79769313486231570815e308)) A = 0.0;
#if DL_HCFOLD
            tk[TC_POO] = tk[TC_POO0] - A;
#else
            tk[TC_POO] = tk[TC_POO0] - A / tk[TC_HC];
#endif
            tk[TC_XOFF] = -A;
            if (rescue) Vbits[o] = 0ull;
        } else {
            tk[TC_POO] = DL_HCFOLD ? -INFINITY : INFINITY;
            tk[TC_XOFF] = -INFINITY;
        }
    }
    if (rescue && mine)
        for (int i = threadIdx.x; i < nzp; i += blockDim.x) pdf_acc[(size_t)o * nzp + i] = 0.0;
}

// rescue: forget the sampled max of the flagged targets before the exact one is taken
__global__ void anchor_reset_kernel(int nO, unsigned long long* __restrict__ Akey,
                                    const int* __restrict__ flag, const int* __restrict__ nflag) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nO || *nflag == 0) return;
    if (flag[o]) Akey[o] = 0ull;
}

// 1 / max(v) per target: results = accumulators / max(v) = exp(L - max L)  (utils_cy.pyx:120-128)
__global__ void finalize_scale_kernel(int nO, const unsigned long long* __restrict__ Vbits,
                                      double* __restrict__ invV) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nO) return;
    invV[o] = 1.0 / __longlong_as_double((long long)Vbits[o]);
}

// pdf[o][z] = pdf_acc[o][z] / max(v)     (out: [nO][ld_out])
__global__ void finalize_pdf_kernel(int nO, int nz, int nzp, const double* __restrict__ pdf_acc,
                                    const double* __restrict__ invV, double* __restrict__ pdf,
                                    long long ld_out) {
    const int z = blockIdx.y * blockDim.x + threadIdx.x;
    const int o = blockIdx.x;
    if (z >= nz) return;
    pdf[(size_t)o * ld_out + z] = pdf_acc[(size_t)o * nzp + z] * invV[o];
}

// evid[o][t] *= 1 / max(v)   (in place: the kernel wrote raw evidences into the output array)
__global__ void finalize_evid_kernel(int nO, int nT, const double* __restrict__ invV,
                                     double* __restrict__ evid) {
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    const int o = blockIdx.x;
    if (t >= nT) return;
    evid[(size_t)o * nT + t] *= invV[o];
}

// p(z, t) = p_t z / p_z_t^2 exp(-z^2 / (2 p_z_t^2))   (scripts/templateFitting.py:79-83), padded
__global__ void template_prior_kernel(int nt, int nz, int nzp, const double* __restrict__ zgrid,
                                      const double* __restrict__ p_t,
                                      const double* __restrict__ p_z_t,
                                      double* __restrict__ prior) {
    const int z = blockIdx.y * blockDim.x + threadIdx.x;
    const int t = blockIdx.x;
    if (z >= nzp) return;
    double v = 0.0;
    if (z < nz) {
        const double b2 = p_z_t[t] * p_z_t[t];
        const double zz = zgrid[z];
        v = p_t[t] * zz / b2 * exp(-0.5 * zz * zz / b2);
    }
    prior[(size_t)t * nzp + z] = v;
}

// ------------------------------------------------------------------------------------------
// top-K training galaxies per target by evidence (scripts/delight-apply.py:150-153).
// Order: descending evidence, equal evidences -> larger index first
// (= np.argsort(kind='stable')[::-1]).  The rows hold the raw evidences of
// the main kernel; the positive per-target factor 1/max(v) does not change the order and is
// applied to the K winners only.  One CTA per target:
// 8 x 8-bit radix-select passes over the order-preserving 64-bit keys, gather, bitonic sort.
constexpr int TK_THREADS = 256;
constexpr int TK_MAXK = 1024;

__device__ __forceinline__ unsigned long long order_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(TK_THREADS) topk_kernel(int nT, int K,
                                                          const double* __restrict__ evid,
                                                          const double* __restrict__ invV,
                                                          long long* __restrict__ top_idx,
                                                          double* __restrict__ top_ev) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_count, s_eq_skip, s_run;
    __shared__ unsigned int warp_tot[TK_THREADS / 32];
    __shared__ unsigned long long skey[TK_MAXK];
    __shared__ int sidx[TK_MAXK];

    const int o = blockIdx.x;
    const double* row = evid + (size_t)o * nT;
    const int tid = threadIdx.x;

    if (tid == 0) { s_prefix = 0ull; s_need = K; }
    __syncthreads();
    // radix select of the K-th largest key
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        hist[tid] = 0;   // TK_THREADS == 256
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int t = tid; t < nT; t += TK_THREADS) {
            const unsigned long long k = order_key(row[t]);
            if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int need = s_need;
            int b = 255;
            for (; b > 0; --b) {
                if ((int)hist[b] >= need) break;
                need -= (int)hist[b];
            }
            s_need = need;
            s_prefix = prefix | ((unsigned long long)b << shift);
        }
        __syncthreads();
    }
    const unsigned long long kth = s_prefix;
    const int need_eq = s_need;      // how many elements equal to kth belong to the top K
    // count ties
    if (tid == 0) { s_count = 0; s_run = 0; }
    __syncthreads();
    int my_eq = 0;
    for (int t = tid; t < nT; t += TK_THREADS) my_eq += (order_key(row[t]) == kth);
    my_eq = __reduce_add_sync(0xffffffffu, my_eq);
    if ((tid & 31) == 0) atomicAdd(&s_run, my_eq);
    __syncthreads();
    if (tid == 0) { s_eq_skip = s_run - need_eq; s_run = 0; }   // ties to skip (lowest indices)
    __syncthreads();
    const int eq_skip = s_eq_skip;
    // gather: everything above kth, and the ties with the largest indices
    for (int base = 0; base < nT; base += TK_THREADS) {
        const int t = base + tid;
        unsigned long long k = 0ull;
        bool gt = false, eq = false;
        if (t < nT) { k = order_key(row[t]); gt = k > kth; eq = k == kth; }
        bool take = gt;
        if (eq_skip == 0) {
            take = gt || eq;
        } else {
            // rank of this tie in index order
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            const int wrank = __popc(bal & ((1u << (tid & 31)) - 1u));
            if ((tid & 31) == 0) warp_tot[tid >> 5] = __popc(bal);
            __syncthreads();
            int off = s_run;
            for (int w = 0; w < (tid >> 5); ++w) off += warp_tot[w];
            if (eq && off + wrank >= eq_skip) take = true;
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int w = 0; w < TK_THREADS / 32; ++w) tot += warp_tot[w];
                s_run += tot;
            }
            __syncthreads();
        }
        if (take) {
            const int pos = atomicAdd(&s_count, 1);
            if (pos < TK_MAXK) { skey[pos] = k; sidx[pos] = t; }
        }
    }
    __syncthreads();
    // pad to a power of two and bitonic-sort descending by (key, index)
    int n2 = 1;
    while (n2 < K) n2 <<= 1;
    for (int i = K + tid; i < n2; i += TK_THREADS) { skey[i] = 0ull; sidx[i] = -1; }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < n2; i += TK_THREADS) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const unsigned long long ki = skey[i], kj = skey[j];
                    const int ii = sidx[i], ij = sidx[j];
                    const bool i_before_j = (ki > kj) || (ki == kj && ii > ij);
                    if (i_before_j != desc) {
                        skey[i] = kj; skey[j] = ki; sidx[i] = ij; sidx[j] = ii;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < K; i += TK_THREADS) {
        const int t = sidx[i];
        top_idx[(size_t)o * K + i] = t;
        top_ev[(size_t)o * K + i] = t >= 0 ? row[t] * invV[o] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// First seven columns of computeMetrics (delight/utils.py:213-227) from the final PDF; one warp
// per target.  out [nO][8]: ztrue zmean std(zmean) zMAP std(zMAP) pdf(ztrue) cumpdf(ztrue) iMAP
__global__ void metrics_kernel(int nO, int nz, const double* __restrict__ pdf, long long ld,
                               const double* __restrict__ zgrid,
                               const double* __restrict__ ztrue, double* __restrict__ out) {
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (o >= nO) return;
    const double* p = pdf + (size_t)o * ld;
    double sw = 0, swz = 0, best = -INFINITY;
    int ibest = 0x7fffffff;
    for (int i = lane; i < nz; i += 32) {
        const double w = p[i];
        sw += w; swz += w * zgrid[i];
        if (w > best) { best = w; ibest = i; }
    }
    sw = warp_sum(sw); swz = warp_sum(swz);
    // argmax, first maximum (np.argmax)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, ibest, s);
        if (ob > best || (ob == best && oi < ibest)) { best = ob; ibest = oi; }
    }
    const double zmean = swz / sw;
    const double zmap = zgrid[ibest < nz ? ibest : 0];
    double v1 = 0, v2 = 0;
    for (int i = lane; i < nz; i += 32) {
        const double w = p[i], d1 = zgrid[i] - zmean, d2 = zgrid[i] - zmap;
        v1 += w * d1 * d1; v2 += w * d2 * d2;
    }
    v1 = warp_sum(v1); v2 = warp_sum(v2);
    // np.interp(ztrue, grid, pdf) and np.interp(ztrue, grid, cumsum(pdf))
    const double zt = ztrue ? ztrue[o] : -1.0;
    double pz, cz;
    int lo = 0;
    if (zt <= zgrid[0]) lo = -1;
    else if (zt >= zgrid[nz - 1]) lo = nz;
    else {
        int a = 0, b = nz - 1;               // grid[a] <= zt < grid[b]
        while (b - a > 1) { const int mid = (a + b) >> 1; if (zgrid[mid] <= zt) a = mid; else b = mid; }
        lo = a;
    }
    const int upto = lo < 0 ? 0 : (lo >= nz ? nz - 1 : lo);
    double cs = 0;
    for (int i = lane; i <= upto; i += 32) cs += p[i];
    cs = warp_sum(cs);
    if (lo < 0) { pz = p[0]; cz = p[0]; }
    else if (lo >= nz) { pz = p[nz - 1]; cz = cs; }
    else {
        const double fr = (zt - zgrid[lo]) / (zgrid[lo + 1] - zgrid[lo]);
        pz = (p[lo + 1] - p[lo]) * fr + p[lo];
        cz = p[lo + 1] * fr + cs;
    }
    if (lane == 0) {
        double* m = out + (size_t)o * 8;
        m[0] = zt; m[1] = zmean; m[2] = sqrt(v1 / sw); m[3] = zmap; m[4] = sqrt(v2 / sw);
        m[5] = pz; m[6] = cz; m[7] = (double)ibest;
    }
}

// trapezoid weights of np.trapz(y, x=zgrid) written as a dot product (delight-apply.py:148):
// w[0] = (z1-z0)/2, w[i] = (z[i+1]-z[i-1])/2, w[nz-1] = (z[nz-1]-z[nz-2])/2, 0 in the padding;
// plus the z-dependent ell-prior terms of utils_cy.pyx:102-104,116 when ell_hat/ell_var vary.
__global__ void zaux_kernel(int nz, int nzp, const double* __restrict__ zgrid,
                            const double* __restrict__ eh, const double* __restrict__ ev,
                            double* __restrict__ wz, double* __restrict__ zd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nzp) return;
    double w = 0.0;
    if (i < nz) {
        const double lo = zgrid[i > 0 ? i - 1 : 0];
        const double hi = zgrid[i + 1 < nz ? i + 1 : nz - 1];
        w = 0.5 * (hi - lo);
    }
    wz[i] = w;
    if (eh != nullptr) {
        const double h = i < nz ? eh[i] : 1.0, v = i < nz ? ev[i] : 1.0;
        zd[i] = h / v; zd[nzp + i] = 1.0 / v; zd[2 * nzp + i] = h * h / v; zd[3 * nzp + i] = v;
    }
}

// CIlevel (delight/utils.py:196-205) for up to 8 fractions at once; one warp per target.
//   evidence = trapz(PDF, grid)
//   for level in linspace(0, PDF.max(), 200):  first level with
//       trapz(PDF[PDF <= level], grid[PDF <= level]) >= fraction * evidence
// The trapezoid over the compressed subset bridges the gaps between selected points, so each
// selected point needs its previous selected point: found with a ballot inside a 32-point
// chunk and carried between chunks.  Levels are scanned once for all fractions.
constexpr int CI_MAXF = 8;
__global__ void ci_levels_kernel(int nO, int nz, const double* __restrict__ pdf, long long ld,
                                 const double* __restrict__ zgrid, int nfrac,
                                 const double* __restrict__ fractions, int numlevels,
                                 double* __restrict__ out) {
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (o >= nO) return;
    const double* p = pdf + (size_t)o * ld;
    // evidence and max
    double ev = 0.0, mx = -INFINITY;
    for (int i = lane; i < nz; i += 32) {
        mx = fmax(mx, p[i]);
        if (i + 1 < nz) ev += (zgrid[i + 1] - zgrid[i]) * (p[i + 1] + p[i]) * 0.5;
    }
    ev = warp_sum(ev);
    mx = warp_max(mx);
    const double step = (mx - 0.0) / (double)(numlevels - 1);      // np.linspace
    unsigned done = 0;
    double res[CI_MAXF];
#pragma unroll
    for (int k = 0; k < CI_MAXF; ++k) res[k] = nan("");
    const unsigned allmask = (1u << nfrac) - 1u;
    for (int lv = 0; lv < numlevels && done != allmask; ++lv) {
        const double level = (lv == numlevels - 1) ? mx : lv * step + 0.0;
        double acc = 0.0, cz = 0.0, cp = 0.0;      // carried previous selected point
        bool have = false;
        for (int base = 0; base < nz; base += 32) {
            const int i = base + lane;
            const double y = i < nz ? p[i] : 0.0;
            const double x = i < nz ? zgrid[i] : 0.0;
            const bool sel = i < nz && y <= level;
            const unsigned bal = __ballot_sync(0xffffffffu, sel);
            const unsigned below = bal & ((1u << lane) - 1u);
            const int src = below ? 31 - __clz(below) : 0;
            const double px = __shfl_sync(0xffffffffu, x, src);
            const double py = __shfl_sync(0xffffffffu, y, src);
            if (sel) {
                if (below) acc += (x - px) * (y + py) * 0.5;
                else if (have) acc += (x - cz) * (y + cp) * 0.5;
            }
            if (bal) {
                const int last = 31 - __clz(bal);
                cz = __shfl_sync(0xffffffffu, x, last);
                cp = __shfl_sync(0xffffffffu, y, last);
                have = true;
            }
        }
        acc = warp_sum(acc);
#pragma unroll
        for (int k = 0; k < CI_MAXF; ++k) {
            if (k < nfrac && !(done & (1u << k)) && acc >= fractions[k] * ev) {
                res[k] = level;
                done |= 1u << k;
            }
        }
    }
    if (lane == 0)
        for (int k = 0; k < nfrac; ++k) out[(size_t)o * nfrac + k] = res[k];
}

// The same scan with the LEVELS spread over the lanes (lane l owns levels l, l + 32, ...; up to
// 8 per lane, i.e. numlevels <= 256): every lane walks the whole PDF once, keeping per level the
// previous selected point and the running trapezoid sum -- no ballots, no shuffles, no pass over
// the PDF per level.  6-10x faster than ci_levels_kernel (which scans level after level and
// stays as the general form for numlevels > 256): 56 -> 8 ms for the 100k targets of config 2.
constexpr int CI_LPL = 8;
template <int LPL>
__global__ void __launch_bounds__(128) ci_levels_lanes_kernel(
    int nO, int nz, const double* __restrict__ pdf, long long ld, const double* __restrict__ zgrid,
    int nfrac, const double* __restrict__ fractions, int numlevels, double* __restrict__ out) {
    extern __shared__ double ci_sm[];                 // per warp: p[nz], then (once) zgrid[nz]
    const int warp = threadIdx.x >> 5;
    const int o = blockIdx.x * (blockDim.x >> 5) + warp;
    const int lane = threadIdx.x & 31;
    double* zs = ci_sm;                               // [nz] shared by the CTA's warps
    double* ps = ci_sm + nz + (size_t)warp * nz;      // [nz] of this warp's target
    for (int i = threadIdx.x; i < nz; i += blockDim.x) zs[i] = zgrid[i];
    __syncthreads();
    if (o >= nO) return;
    const double* p = pdf + (size_t)o * ld;
    double ev = 0.0, mx = -INFINITY;
    for (int i = lane; i < nz; i += 32) {
        const double y = p[i];
        ps[i] = y;
        mx = fmax(mx, y);
        if (i + 1 < nz) ev += (zs[i + 1] - zs[i]) * (p[i + 1] + y) * 0.5;
    }
    __syncwarp();
    ev = warp_sum(ev);
    mx = warp_max(mx);
    const double step = (mx - 0.0) / (double)(numlevels - 1);      // np.linspace
    // acc2 = twice the trapezoid sum (the 1/2 is applied to the threshold side, exactly)
    double lev[LPL], acc2[LPL], px[LPL], py[LPL];
#pragma unroll
    for (int k = 0; k < LPL; ++k) {
        const int lv = lane + 32 * k;
        lev[k] = lv >= numlevels ? -INFINITY : (lv == numlevels - 1 ? mx : lv * step + 0.0);
        acc2[k] = 0.0; px[k] = 0.0; py[k] = 0.0;
    }
    unsigned have = 0;
#pragma unroll 4
    for (int i = 0; i < nz; ++i) {
        const double y = ps[i], x = zs[i];                         // broadcast LDS
#pragma unroll
        for (int k = 0; k < LPL; ++k) {
            const bool sel = y <= lev[k];
            const double term = (x - px[k]) * (y + py[k]);
            if (sel && (have >> k & 1u)) acc2[k] += term;
            if (sel) { px[k] = x; py[k] = y; have |= 1u << k; }
        }
    }
    // first level (ascending) whose integral reaches fraction * evidence
    for (int f = 0; f < nfrac; ++f) {
        const double want = fractions[f] * ev;
        int first = 0x7fffffff;
#pragma unroll
        for (int k = LPL - 1; k >= 0; --k)
            if (lane + 32 * k < numlevels && acc2[k] * 0.5 >= want) first = lane + 32 * k;
        first = __reduce_min_sync(0xffffffffu, first);
        if (lane == 0)
            out[(size_t)o * nfrac + f] = first == 0x7fffffff ? nan("")
                : (first == numlevels - 1 ? mx : first * step + 0.0);
    }
}

// prior[t][z] = exp(-0.5 ((z - z_t) / sigma)^2)   (scripts/delight-apply.py:104), padded rows
__global__ void zprior_kernel(int nT, int nz, int nzp, const double* __restrict__ zgrid,
                              const double* __restrict__ ztrain, double sigma,
                              double* __restrict__ prior) {
    const int z = blockIdx.y * blockDim.x + threadIdx.x;
    const int t = blockIdx.x;
    if (z >= nzp) return;
    double v = 0.0;
    if (z < nz) {
        const double d = (zgrid[z] - ztrain[t]) / sigma;
        v = exp(-0.5 * d * d);
    }
    prior[(size_t)t * nzp + z] = v;
}

// legacy [nz][nT][nB] (scripts/delight-apply.py:84-85) -> device layout [nT][2nB][nzp]
__global__ void relayout_grid_kernel(int nz, int nT, int nB, int nzp,
                                     const double* __restrict__ mean,
                                     const double* __restrict__ covar,
                                     double* __restrict__ grid) {
    // block = (32 z) x (all bands of one t), transposed through smem
    extern __shared__ double tile[];     // [2][32][nB+1]
    const int t = blockIdx.x;
    const int zb = blockIdx.y * 32;
    const int ld = nB + 1;
    for (int i = threadIdx.x; i < 32 * nB; i += blockDim.x) {
        const int zz = i / nB, b = i % nB;
        const int z = zb + zz;
        double a = 0.0, c = 0.0;            // padding: mean 0, model variance 0
        if (z < nz) {
            a = mean[((size_t)z * nT + t) * nB + b];
            c = covar[((size_t)z * nT + t) * nB + b];
        }
        tile[zz * ld + b] = a;
        tile[32 * ld + zz * ld + b] = c;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * nB; i += blockDim.x) {
        const int b = i / 32, zz = i % 32;
        const int z = zb + zz;
        if (z < nzp) {
            grid[((size_t)t * 2 * nB + b) * nzp + z] = tile[zz * ld + b];
            grid[((size_t)t * 2 * nB + nB + b) * nzp + z] = tile[32 * ld + zz * ld + b];
        }
    }
}

// explicit prior [nz][nT] (e.g. templateFitting.py:79-82) -> [nT][nzp]
__global__ void relayout_prior_kernel(int nz, int nT, int nzp, const double* __restrict__ src,
                                      double* __restrict__ dst) {
    __shared__ double tile[32][33];
    const int zb = blockIdx.x * 32, tb = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int z = zb + r, t = tb + tx;
        tile[r][tx] = (z < nz && t < nT) ? src[(size_t)z * nT + t] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int t = tb + r, z = zb + tx;
        if (t < nT && z < nzp) dst[(size_t)t * nzp + z] = tile[tx][r];
    }
}

}  // namespace dl

using namespace dl;

// ============================================================================ C ABI (device)
namespace {

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

struct LaunchPlan {
    int cfg, stages, TT;
};

// ring depth and training-block length that fit the configuration's shared-memory budget:
// config 0 shares an SM between two CTAs (228 KB per SM, 1 KB reserved per CTA), config 1 owns it
int plan_launch(int nf, bool het, bool zdep, int nU, long long nTlist, LaunchPlan& pl) {
    pl.cfg = cfg_of(nf, het);
    const size_t budget = pl.cfg == 0 ? (228 * 1024 / 2 - 1024) : 227 * 1024;
    int TT = env_int("DL_FUSED_TT", cfg_tt(pl.cfg));
    if (TT > nTlist) TT = (int)nTlist;
    if (TT < 1) TT = 1;
    const int smax = env_int("DL_FUSED_STAGES", DL_STAGES);
    for (;; TT = (TT + 1) / 2) {
        for (int s = smax; s >= 2; --s)
            if (fused_smem_bytes(pl.cfg, nf, het, zdep, nU, s, TT) <= budget) {
                pl.stages = s; pl.TT = TT;
                return DL_OK;
            }
        if (TT <= 8) break;
    }
    dl_set_error("fused likelihood: %d bands (%d staged rows) do not fit shared memory", nf,
                 2 * nU + 2);
    return DL_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" {

int dl_fused_chunk() { return 128; }      // nz is padded to this (both configurations divide it)
int dl_fused_targets_per_cta(int nf, int het) {
    return cfg_warps(cfg_of(nf, het != 0)) * FTO;
}
// CTAs of the fused kernel that share an SM: 2 in configuration 0, 1 in configuration 1
int dl_fused_ctas_per_sm(int nf, int het) { return cfg_of(nf, het != 0) == 0 ? 2 : 1; }

long long dl_fused_padded_nz(long long nz) { return (nz + 127) / 128 * 128; }

// bytes of scratch dl_like_fused_dev needs for nO targets against nT training galaxies
// (evid_is_output_or_unused: the caller passes an evidence array, or wants no evidences at all)
long long dl_fused_workspace_bytes(long long nO, long long nT, long long nz,
                                   int evid_is_output_or_unused) {
    const long long nzp = dl_fused_padded_nz(nz);
    auto al = [](long long x) { return (x + 255) & ~255ll; };
    long long b = 0;
    b += al((long long)tc_stride(MAXNF, true) * nO * 8);     // tconst
    b += al(nO * nzp * 8);                                   // pdf_acc
    b += 3 * al(nO * 8);                                     // Vbits, Akey, invV
    b += al((nO + 1) * 4);                                   // flag, nflag
    b += al(nzp * 8) + al(4 * nzp * 8);                      // wz, zdep
    if (!evid_is_output_or_unused) b += al(nO * nT * 8);     // raw evidences (internal)
    return b;
}

int dl_relayout_grid_dev(long long nz, long long nT, long long nB, const double* model_mean,
                         const double* model_covar, double* grid, void* stream) {
    DL_REQUIRE(nz > 0 && nT > 0 && nB > 0 && nB <= 64, "bad shape");
    const int nzp = (int)dl_fused_padded_nz(nz);
    dim3 g((unsigned)nT, (nzp + 31) / 32);
    const size_t sm = sizeof(double) * 2 * 32 * (nB + 1);
    relayout_grid_kernel<<<g, 256, sm, (cudaStream_t)stream>>>((int)nz, (int)nT, (int)nB, nzp,
                                                               model_mean, model_covar, grid);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

int dl_relayout_prior_dev(long long nz, long long nT, const double* prior_zt, double* prior_tz,
                          void* stream) {
    DL_REQUIRE(nz > 0 && nT > 0, "bad shape");
    const int nzp = (int)dl_fused_padded_nz(nz);
    dim3 g((nzp + 31) / 32, (unsigned)((nT + 31) / 32));
    relayout_prior_kernel<<<g, 256, 0, (cudaStream_t)stream>>>((int)nz, (int)nT, nzp, prior_zt,
                                                               prior_tz);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

int dl_zprior_dev(long long nT, long long nz, const double* zgrid, const double* ztrain,
                  double sigma, double* prior_tz, void* stream) {
    DL_REQUIRE(nz > 0 && nT > 0 && sigma > 0, "bad shape");
    const int nzp = (int)dl_fused_padded_nz(nz);
    dim3 g((unsigned)nT, (nzp + 127) / 128);
    zprior_kernel<<<g, 128, 0, (cudaStream_t)stream>>>((int)nT, (int)nz, nzp, zgrid, ztrain,
                                                       sigma, prior_tz);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

// Template prior of scripts/templateFitting.py:79-83 in the fused layout [nt][nzp]:
// p(z, t) = p_t[t] * z / p_z_t[t]^2 * exp(-z^2 / (2 p_z_t[t]^2))
int dl_template_prior_dev(long long nt, long long nz, const double* zgrid, const double* p_t,
                          const double* p_z_t, double* prior_tz, void* stream) {
    DL_REQUIRE(nz > 0 && nt > 0 && zgrid && p_t && p_z_t && prior_tz, "bad arguments");
    const int nzp = (int)dl_fused_padded_nz(nz);
    dim3 g((unsigned)nt, (nzp + 127) / 128);
    template_prior_kernel<<<g, 128, 0, (cudaStream_t)stream>>>((int)nt, (int)nz, nzp, zgrid, p_t,
                                                               p_z_t, prior_tz);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

// Fused likelihood + apply epilogue for targets observed in nf bands each.
//   obs, var     [nO][nf]   target fluxes / variances in the target's own bands
//   n_union, union_bands    HOST array: the band rows of the grid staged per step (0 <= b < nB)
//   target_rows  NULL: every target uses exactly union_bands (n_union == nf); else DEVICE int32
//                [nO][nf]: for target o, band f is union_bands[target_rows[o][f]]
//   grid         [nT][2nB][nzp] device, prior [nT][nzp] device (nzp = dl_fused_padded_nz(nz))
//   zgrid        [nz] device
//   ell_hat/ell_var: scalars, or (ell_hat_z, ell_var_z != NULL) device vectors [nz]
//   cov_scale    factor on the model variance (1 in delight-apply; V_C in delight-optimize.py:122)
//   outputs (device): pdf [nO][ld_pdf]; evid [nO][nT] or NULL; top_idx [nO][K] int64,
//                     top_ev [nO][K] (K may be 0)
int dl_like_fused_dev(long long nO, long long nT, long long nz, long long nB, int nf,
                      int n_union, const int* union_bands, const int* target_rows,
                      const double* obs, const double* var,
                      const double* grid, const double* prior, const double* zgrid,
                      double ell_hat, double ell_var, const double* ell_hat_z,
                      const double* ell_var_z, double cov_scale, int K, double* pdf,
                      long long ld_pdf,
                      double* evid, long long* top_idx, double* top_ev, void* workspace,
                      long long workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DL_REQUIRE(nO > 0 && nT > 0 && nz > 1 && nB > 0, "bad shape");
    DL_REQUIRE(nO < (1ll << 31) / 64 && nT < (1ll << 31) / 2, "too many targets / training galaxies for one call");
    DL_REQUIRE(nf >= 1 && nf <= MAXNF, "1..32 bands per target");
    DL_REQUIRE(union_bands && n_union >= nf && n_union <= MAXNF, "bad band union");
    const bool het = target_rows != nullptr;
    DL_REQUIRE(het || n_union == nf, "n_union != nf without per-target band lists");
    DL_REQUIRE(K >= 0 && K <= TK_MAXK && K <= nT, "0 <= K <= min(nT, 1024)");
    DL_REQUIRE(ld_pdf >= nz, "ld_pdf < nz");
    DL_REQUIRE(cov_scale >= 0.0, "cov_scale < 0");
    const bool zdep = ell_hat_z != nullptr && ell_var_z != nullptr;
    const long long need = dl_fused_workspace_bytes(nO, nT, nz, evid != nullptr || K == 0);
    DL_REQUIRE(workspace && workspace_bytes >= need, "workspace too small");
    const int nzp = (int)dl_fused_padded_nz(nz);

    char* w = (char*)workspace;
    auto take = [&](size_t bytes) { char* r = w; w += (bytes + 255) & ~(size_t)255; return r; };
    double* tconst = (double*)take((size_t)tc_stride(MAXNF, true) * nO * 8);
    double* pdf_acc = (double*)take((size_t)nO * nzp * 8);
    unsigned long long* Vbits = (unsigned long long*)take((size_t)nO * 8);
    unsigned long long* Akey = (unsigned long long*)take((size_t)nO * 8);
    double* invV = (double*)take((size_t)nO * 8);
    int* flag = (int*)take((size_t)(nO + 1) * 4);
    int* nflag = flag + nO;
    double* wz = (double*)take((size_t)nzp * 8);
    double* zd = (double*)take((size_t)4 * nzp * 8);
    double* E = evid ? evid : (K > 0 ? (double*)take((size_t)nO * nT * 8) : nullptr);
    DL_REQUIRE((size_t)(w - (char*)workspace) <= (size_t)workspace_bytes, "workspace accounting");

    // Vbits, Akey, invV, flag are contiguous: one memset; the PDF accumulators another
    DL_CUDA_CHECK(cudaMemsetAsync(Vbits, 0, (char*)wz - (char*)Vbits, stream));
    DL_CUDA_CHECK(cudaMemsetAsync(pdf_acc, 0, (size_t)nO * nzp * 8, stream));

    zaux_kernel<<<(unsigned)((nzp + 127) / 128), 128, 0, stream>>>(
        (int)nz, nzp, zgrid, zdep ? ell_hat_z : nullptr, zdep ? ell_var_z : nullptr, wz, zd);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();

    LaunchPlan pl;
    {
        int rc = plan_launch(nf, het, zdep, n_union, nT, pl);
        if (rc != DL_OK) return rc;
    }
    const int TCS = tc_stride(nf, het);
    prep_targets_kernel<<<(unsigned)((nO + 127) / 128), 128, 0, stream>>>(
        (int)nO, nf, TCS, cfg_fzc(pl.cfg), obs, var, target_rows, zdep ? 1.0 : ell_hat,
        zdep ? 1.0 : ell_var, cov_scale, tconst);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();

    int sm_count = 148, l2_bytes = 126 << 20;
    {
        int dev = 0;
        DL_CUDA_CHECK(cudaGetDevice(&dev));
        DL_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        DL_CUDA_CHECK(cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, dev));
    }
    const int ftpc = cfg_warps(pl.cfg) * FTO;
    const int slots = sm_count * (pl.cfg == 0 ? 2 : 1);

    FusedParams p;
    p.grid = grid; p.prior = prior; p.tconst = tconst; p.wz = wz; p.zdep = zd;
    p.pdf_acc = pdf_acc; p.evid = E; p.Vbits = Vbits; p.Akey = Akey; p.flag = flag; p.nflag = nflag;
    for (int f = 0; f < MAXNF; ++f) {
        p.bands[f] = f < n_union ? union_bands[f] : 0;
        DL_REQUIRE(p.bands[f] >= 0 && p.bands[f] < nB, "band index out of range");
    }
    p.nU = n_union;
    p.nB = (int)nB; p.nz = (int)nz; p.nzp = nzp; p.nT = (int)nT; p.nO = (int)nO;
    p.nTiles = (int)((nO + ftpc - 1) / ftpc);
    // tiles per L2 group: their PDF rows (tile x nzp x 8 B each, updated with red.add at every
    // chunk end) must stay L2-resident next to the training blocks in flight.  Measured at the
    // bench shape (ncu dram bytes of one main pass, profiles/): a third of the L2 for the PDF
    // rows (G = 111 of 32 targets) -> 5.0 GB read + 0.95 GB written; G = 148: 5.0 + 6.9 GB
    // (rows evicted between flushes); G = 74: 8.4 + 1.2 GB (model grid re-read per group).
    {
        const long long tile_bytes = (long long)ftpc * nzp * 8;
        long long g = (long long)l2_bytes / 3 / tile_bytes;
        g = g < 4 ? 4 : (g > slots ? slots : g);
        p.G = env_int("DL_FUSED_G", (int)g);
    }
    p.stages = pl.stages;

    auto set_blocks = [&](long long nTlist, int stride, int TT) {
        // small problems: shorter training blocks until the machine is (twice) full
        while (TT > 8 && (long long)p.nTiles * ((nTlist + TT - 1) / TT) < 2ll * slots) TT = (TT + 1) / 2;
        p.nTlist = (int)nTlist; p.t_stride = stride; p.TT = TT;
        p.nTB = (int)((nTlist + TT - 1) / TT);
    };

    // 1. anchor exponent per target: max of -chi2/2 over a strided sample of training galaxies
    //    (every stride-th galaxy, spread over the whole catalogue: it may be sorted)
    long long nS_want = env_int("DL_FUSED_NSAMPLE", 64);
    if (nS_want < 1) nS_want = 1;
    const long long stride = (nT + nS_want - 1) / nS_want;
    const long long nS = (nT + stride - 1) / stride;
    p.rescue = 0;
    set_blocks(nS, (int)stride, pl.TT < 16 ? pl.TT : 16);
    int rc = launch_fused(nf, zdep, het, MODE_SAMPLE, p, stream, false);
    if (rc != DL_OK) return rc;
    anchor_fold_kernel<<<(unsigned)nO, 128, 0, stream>>>((int)nO, nf, TCS, tconst, Akey, flag,
                                                         nflag, 0, pdf_acc, nzp, Vbits);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();

    // 2. the main pass
    set_blocks(nT, 1, pl.TT);
    rc = launch_fused(nf, zdep, het, MODE_MAIN, p, stream, true);
    if (rc != DL_OK) return rc;

    // 3. rescue: targets with a cell > e^500 above their anchor are redone with the exact max
    //    (every CTA of these launches returns at once when nothing was flagged)
    if (!env_int("DL_FUSED_NORESCUE", 0)) {
        anchor_reset_kernel<<<(unsigned)((nO + 127) / 128), 128, 0, stream>>>((int)nO, Akey, flag,
                                                                              nflag);
        DL_CUDA_CHECK(cudaGetLastError());
        p.rescue = 1;
        rc = launch_fused(nf, zdep, het, MODE_SAMPLE, p, stream, false);
        if (rc != DL_OK) return rc;
        anchor_fold_kernel<<<(unsigned)nO, 128, 0, stream>>>((int)nO, nf, TCS, tconst, Akey, flag,
                                                             nflag, 1, pdf_acc, nzp, Vbits);
        DL_CUDA_CHECK(cudaGetLastError());
        rc = launch_fused(nf, zdep, het, MODE_MAIN, p, stream, false);
        if (rc != DL_OK) return rc;
        dl_count_launch(2);
    }

    // 4. normalise: everything is divided by the target's max(v)
    finalize_scale_kernel<<<(unsigned)((nO + 127) / 128), 128, 0, stream>>>((int)nO, Vbits, invV);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    {
        dim3 g((unsigned)nO, (unsigned)((nz + 127) / 128));
        finalize_pdf_kernel<<<g, 128, 0, stream>>>((int)nO, (int)nz, nzp, pdf_acc, invV, pdf,
                                                   ld_pdf);
        DL_CUDA_CHECK(cudaGetLastError());
        dl_count_launch();
    }
    if (K > 0) {
        DL_REQUIRE(top_idx && top_ev, "top-K outputs missing");
        topk_kernel<<<(unsigned)nO, TK_THREADS, 0, stream>>>((int)nT, K, E, invV, top_idx, top_ev);
        DL_CUDA_CHECK(cudaGetLastError());
        dl_count_launch();
    }
    if (evid) {
        dim3 g((unsigned)nO, (unsigned)((nT + 127) / 128));
        finalize_evid_kernel<<<g, 128, 0, stream>>>((int)nO, (int)nT, invV, evid);
        DL_CUDA_CHECK(cudaGetLastError());
        dl_count_launch();
    }
    return DL_OK;
}

int dl_metrics_dev(long long nO, long long nz, const double* pdf, long long ld_pdf,
                   const double* zgrid, const double* ztrue, double* out, void* stream) {
    DL_REQUIRE(nO > 0 && nz > 1 && ld_pdf >= nz, "bad shape");
    metrics_kernel<<<(unsigned)((nO + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        (int)nO, (int)nz, pdf, ld_pdf, zgrid, ztrue, out);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

// CIlevel for nfrac <= 8 fractions (device array); out [nO][nfrac].  NaN where the reference
// would return None (never for fractions <= 1).
int dl_ci_levels_dev(long long nO, long long nz, const double* pdf, long long ld_pdf,
                     const double* zgrid, int nfrac, const double* fractions, int numlevels,
                     double* out, void* stream) {
    DL_REQUIRE(nO > 0 && nz > 1 && ld_pdf >= nz, "bad shape");
    DL_REQUIRE(nfrac >= 1 && nfrac <= CI_MAXF && numlevels >= 2, "1..8 fractions");
    const size_t sm = sizeof(double) * (size_t)nz * 5;         // zgrid + 4 warps' PDFs
    if (numlevels <= 32 * CI_LPL && sm <= 200 * 1024) {
        auto kern = numlevels <= 32 * 7 ? ci_levels_lanes_kernel<7> : ci_levels_lanes_kernel<CI_LPL>;
        DL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sm));
        kern<<<(unsigned)((nO + 3) / 4), 128, sm, (cudaStream_t)stream>>>(
            (int)nO, (int)nz, pdf, ld_pdf, zgrid, nfrac, fractions, numlevels, out);
    } else {
        ci_levels_kernel<<<(unsigned)((nO + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
            (int)nO, (int)nz, pdf, ld_pdf, zgrid, nfrac, fractions, numlevels, out);
    }
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

}  // extern "C"

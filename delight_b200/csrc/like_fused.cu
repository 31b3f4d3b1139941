// Fused photo-z likelihood for a batch of targets (subsystem 1 of the north star).
//
// Replaces the per-target host loop of scripts/delight-apply.py:114-162:
//     approx_flux_likelihood_cy (delight/utils_cy.pyx:83-128)  ->  like[nz][nT]
//     like *= prior; pdf = sum_t like; evid = trapz_z like; top-K by evidence
// with one launch over (z-chunk x target-tile x training-range).  The like[nz][nT] array of a
// target never exists in HBM.
//
// Work decomposition (B200: 148 SMs, FP64 pipe = 16 lanes / SMSP, 227 KB smem / SM)
//   CTA      = 8 warps (warp 0 also drives TMA); 2 CTAs / SM (97 KB smem, <=128 regs)
//   CTA tile = one z-chunk of 128 redshifts  x  32 targets (4 per compute warp)
//              x  a contiguous range of training galaxies (all of them unless nsplit > 1)
//   lane     = 4 redshifts of the chunk (z = z0 + 32*j + lane) -> conflict-free smem rows
//   staging  = per training galaxy one pipeline stage: 2*NF+1 rows (NF band means, NF band
//              variances, the prior row) of 128 doubles each, fetched with 1-D TMA bulk copies
//              (cp.async.bulk ... mbarrier::complete_tx) into a full/empty mbarrier ring
//   compute  = per (target, training galaxy, z) cell, all in FP64 on the vector pipe:
//              pass 0 with the observed variances only (3 FMA-class ops / band, per-target
//              constants broadcast from smem), one reciprocal for the luminosity scale, pass 1
//              with the model variance folded in (one reciprocal per PAIR of bands + 7 FMA-class
//              ops / band), then chi2, (prod var * FTT)^-1/2 by one rsqrt, one table-based exp.
//              No log, no division; 96 FP64-pipe instructions per cell at 6 bands.
//
// Normalisation.  The reference subtracts the GLOBAL max log-likelihood of a target before
// exp (utils_cy.pyx:120-128).  Here every warp keeps, per target, a lazily updated reference
// exponent R (running max of -chi2/2 over what the warp has seen) and the running max V of the
// un-priored value exp(-chi2/2 - R) / sqrt(det); accumulators are rescaled on the rare update.
// A finalize pass combines the (R, V) of all warps of a target and rescales pdf / evidences, so
// results equal exp(L - max L) up to rounding; per-target constant factors cancel and are never
// computed.  All variances are pre-scaled by a per-target power of two c (exact) so products
// over bands stay in range.
#include "common.cuh"
#ifdef DL_DBGWAIT
#include <cstdio>
#include <vector>
#endif

namespace dl {

#ifndef DL_FW
#define DL_FW 8
#endif
#ifndef DL_FTO
#define DL_FTO 4
#endif
#ifndef DL_ZJ
#define DL_ZJ 2
#endif
#ifndef DL_STAGES
#define DL_STAGES 4
#endif
#ifndef DL_MINB
#define DL_MINB 2
#endif
constexpr int FW = DL_FW;         // compute warps per CTA
constexpr int FTO = DL_FTO;       // targets per compute warp
constexpr int FZT = 4;            // redshifts per lane
constexpr int ZJ_SMALL = DL_ZJ;   // redshifts a lane processes jointly (independent chains),
constexpr int ZJ_LARGE = 1;       //   NF <= 8 / NF > 8 (register budget: 8 NF ZJ registers of model values)
constexpr int FZC = 32 * FZT;     // redshifts per CTA chunk
constexpr int FSTAGES = DL_STAGES;  // TMA ring depth
#ifndef DL_PREF
#define DL_PREF (DL_STAGES - 2)
#endif
constexpr int FPREF = DL_PREF;      // prefetch distance in steps (< FSTAGES)
// Warp 0 also issues the TMA copies of every stage, which makes it the straggler the other
// seven warps wait for.  Tuning switches around that (all measured at config 3):
//   DL_ELECT=1    (default) the copies are issued by the elect.sync leader of warp 0: the compiler
//                 then keeps the addresses in uniform registers and emits the UBLKCPs back to back
//                 instead of a 16-instruction "for each active thread" loop per copy
//                 (1.384 -> 1.346 s).
//   DL_DISTPROD=1 every warp issues its share (rows warp, warp+FW, ...) of each stage and arrives
//                 on `full` with its own expect_tx, so no warp is special (1.350 s with DL_ELECT).
//   DL_ASYM=1     warp DL_LIGHTWARP gets one target less (1.384 s without DL_ELECT: the 3 % more
//                 CTAs eat the gain).
// Round-1 note: before the barrier polls were made whole-warp (see issue_stage) every one of
// these variants measured 1.8-2.8x SLOWER, because the polling lane left its warp split in two
// fragments for the whole compute phase.
#ifndef DL_ASYM
#define DL_ASYM 0
#endif
#ifndef DL_DISTPROD
#define DL_DISTPROD 0
#endif
#ifndef DL_PRODWARP
#define DL_PRODWARP 0
#endif
#ifndef DL_ELECT
#define DL_ELECT 1
#endif
// exp through a table of 2^(j/128) in shared memory + degree-3 polynomial (7 FP64-pipe
// instructions) instead of the degree-9 polynomial (14): 1.311 -> 1.267 s at config 3.
// DL_PAIRRCP = 2 / 3: one reciprocal per pair / triple of bands in pass 1 (0: one per band).
#ifndef DL_EXPTAB
#define DL_EXPTAB 1
#endif
#ifndef DL_PAIRRCP
#define DL_PAIRRCP 2
#endif
#if DL_EXPTAB
#define DL_EXP(x) exp_nonpos_tab((x), etab)
#else
#define DL_EXP(x) exp_nonpos(x)
#endif


constexpr int FTPC = FW * FTO - DL_ASYM;    // targets per CTA
constexpr int FTHREADS = FW * 32;
constexpr int MAXNF = 32;   // bands per target group
// Band counts above DL_REGNF keep the model values in the stage (shared memory) instead of in
// registers, which leaves room for redshift PAIRS there too.  Measured (2048 x 1500 x 1500 cells,
// ms, REGNF 16 / ZJ 1 -> REGNF 8 / ZJ 2): nf 9: 64.7 -> 61.6, 13: 86.6 -> 81.4, 14: 128 -> 100,
// 16: 136 -> 118, 28: 206 -> 195.
#ifndef DL_GRPUNROLL
#define DL_GRPUNROLL 2      // unroll factor of the z-group loop (2 = full for redshift pairs)
#endif
#ifndef DL_REGNF
#define DL_REGNF 8
#endif
#ifndef DL_SMODEL_ZJ
#define DL_SMODEL_ZJ 2
#endif
constexpr int GRPUNROLL = DL_GRPUNROLL;   // (macros are not expanded inside #pragma unroll)
constexpr int REGNF = DL_REGNF;   // up to here the model values of a z-group are cached in registers

// per-target constants: [f][2] = (a' s, w') | [f][4] = (var', obs, obs^2, 0 (f = 0: c s)) |
// hc, pOT s, pOT, pTT, pOO, c   with s = sqrt(cov_scale / c)
__host__ __device__ constexpr int tc_stride(int nf) { return 6 * nf + 6; }

// Shared memory of one CTA with a ring of `stages`, and the ring depth used for NF bands: the
// deepest ring up to DL_STAGES (at least 2) that still lets two CTAs share an SM
// (228 KB per SM, 1 KB reserved per CTA); beyond 13 bands only one CTA fits anyway.
__host__ __device__ constexpr size_t fused_smem_bytes(int nf, bool zdep, int stages) {
    return sizeof(double) * ((size_t)stages * (2 * nf + 1) * FZC + FTPC * FZC + FTPC * tc_stride(nf) +
                             FTPC * 2 + FZC + (zdep ? 4 * FZC : 0) + EXP_TAB) +
           2 * stages * sizeof(uint64_t);
}
__host__ __device__ constexpr int fused_stages(int nf, bool zdep) {
    int s = FSTAGES;
    while (s > 2 && 2 * (fused_smem_bytes(nf, zdep, s) + 1024) > 228 * 1024) --s;
    if (2 * (fused_smem_bytes(nf, zdep, s) + 1024) > 228 * 1024) {
        s = FSTAGES;                                                            // one CTA per SM
        while (s > 2 && fused_smem_bytes(nf, zdep, s) > 227 * 1024) --s;
    }
    return s;
}

struct FusedParams {
    const double* grid;    // [nT][2*nB][nzp]  rows 0..nB-1 mean, nB..2nB-1 variance
    const double* prior;   // [nT][nzp]
    const double* tconst;  // [nO][tc_stride(NF)]
    const double* wz;      // [nzp] trapezoid weights (0 in the padding)
    const double* zdep;    // [4][nzp]: ell_hat/ell_var, 1/ell_var, ell_hat^2/ell_var, ell_var
    double* pdf_part;      // [nsplit][nO][nzp]
    double2* evid_part;    // [nchunk][nO][nT]  (E, R)
    double* Rw;            // [nsplit][nchunk][nO]
    double* Vw;            // [nsplit][nchunk][nO]
    int bands[MAXNF];
    int nB, nz, nzp, nT, nO, nchunk, nsplit;
    long long* dbg;        // DL_DBGWAIT builds: [grid][FW][4] cycle counters
};

// x = -chi2/2 (in the target's units, + the padding mask) and y = (det)^-1/2 of ZJ cells of one
// target against one training galaxy: utils_cy.pyx:98-118 restructured (see the header comment).
// tc: the target's constants, m / cv: band means / variances of the cells.
// Where the band means / variances of the ZJ cells come from: registers (loaded once per z-group
// and reused by the warp's targets; NF <= 16) or the stage in shared memory (NF > 16: 2 NF ZJ
// doubles do not fit the register budget, every use is an LDS).
template <int NF, int ZJ>
struct RegModel {
    const double (&mm)[ZJ][NF];
    const double (&cc)[ZJ][NF];
    __device__ __forceinline__ double m(int c, int f) const { return mm[c][f]; }
    __device__ __forceinline__ double cv(int c, int f) const { return cc[c][f]; }
};
template <int NF, int ZJ>
struct SmemModel {
    const double* stz;      // stage + the lane's first redshift
    __device__ __forceinline__ double m(int c, int f) const { return stz[f * FZC + 32 * c]; }
    __device__ __forceinline__ double cv(int c, int f) const { return stz[(NF + f) * FZC + 32 * c]; }
};

template <int NF, int ZJ, class Model>
__device__ __forceinline__ void cells_xy(const double* tc, const Model md, const double (&pOT0)[ZJ],
                                         const double (&pOT)[ZJ],
                                         const double (&pTT)[ZJ], const double (&pOO)[ZJ],
                                         const double (&zev)[ZJ], const double (&xmask)[ZJ],
                                         double (&x)[ZJ], double (&y)[ZJ]) {
    const double hc = tc[6 * NF + 0];
    // pass 0: observed variances only (utils_cy.pyx:101-113 with ellML = 0)
    double fot[ZJ], ftt[ZJ], foo[ZJ], P[ZJ], P2[ZJ], e2[ZJ];   // det = P * P2 (P2 only for NF > REGNF)
#pragma unroll
    for (int c = 0; c < ZJ; ++c) { fot[c] = pOT0[c]; ftt[c] = pTT[c]; }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const double a = tc[2 * f], w = tc[2 * f + 1];
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            const double mf = md.m(c, f);
            const double tw = mf * w;
            fot[c] = fma(mf, a, fot[c]);
            ftt[c] = fma(tw, mf, ftt[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < ZJ; ++c) {
        const double ell = fot[c] * rcp_fast(ftt[c]);     // ellML * sqrt(cov_scale / c)
        e2[c] = ell * ell;
        fot[c] = pOT[c]; ftt[c] = pTT[c]; foo[c] = pOO[c]; P[c] = zev[c]; P2[c] = 1.0;
    }
    // pass 1: variance inflated by ell^2 * model variance (utils_cy.pyx:107-112)
#if DL_PAIRRCP >= 2
    // Bands in groups of DL_PAIRRCP (2 or 3): ONE reciprocal of the product of the group's
    // variances gives every 1/var (r_a = var_b / (var_a var_b), ...) and the group's factor of
    // the determinant.  Pairs: 9 FP64 instructions + 1 MUFU instead of 10 + 2; triples: 13 + 1
    // instead of 15 + 3.
    auto acc = [&](int c, int f, double r) {
        const double mf = md.m(c, f);
        const double tm = mf * r;
        fot[c] = fma(tm, tc[2 * NF + 4 * f + 1], fot[c]);
        ftt[c] = fma(tm, mf, ftt[c]);
        foo[c] = fma(tc[2 * NF + 4 * f + 2], r, foo[c]);
    };
    constexpr int G = DL_PAIRRCP;
    constexpr int NG = (NF / G) * G;          // bands covered by full groups
#pragma unroll
    for (int f = 0; f < NG; f += G) {
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            const double va = fma(e2[c], md.cv(c, f), tc[2 * NF + 4 * f]);
            const double vb = fma(e2[c], md.cv(c, f + 1), tc[2 * NF + 4 * (f + 1)]);
            const double pab = va * vb;
            if (G == 2) {
                const double ri = rcp_fast(pab);
                acc(c, f, vb * ri);
                acc(c, f + 1, va * ri);
                if (NF > REGNF && f >= NF / 2) P2[c] *= pab; else P[c] *= pab;
            } else {
                const double vc = fma(e2[c], md.cv(c, f + 2), tc[2 * NF + 4 * (f + 2)]);
                const double pabc = pab * vc;
                const double ri = rcp_fast(pabc);
                const double rab = ri * vc;       // 1 / (va vb)
                acc(c, f, vb * rab);
                acc(c, f + 1, va * rab);
                acc(c, f + 2, ri * pab);
                if (NF > REGNF && f >= NF / 2) P2[c] *= pabc; else P[c] *= pabc;
            }
        }
    }
    if (NF - NG == 2) {
        constexpr int f = NG < NF - 1 ? NG : 0;
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            const double va = fma(e2[c], md.cv(c, f), tc[2 * NF + 4 * f]);
            const double vb = fma(e2[c], md.cv(c, f + 1), tc[2 * NF + 4 * (f + 1)]);
            const double pab = va * vb;
            const double ri = rcp_fast(pab);
            acc(c, f, vb * ri);
            acc(c, f + 1, va * ri);
            if (NF > REGNF) P2[c] *= pab; else P[c] *= pab;
        }
    } else if (NF - NG == 1) {
        constexpr int f = NF - 1;
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            const double var = fma(e2[c], md.cv(c, f), tc[2 * NF + 4 * f]);
            acc(c, f, rcp_fast(var));
            if (NF > REGNF) P2[c] *= var; else P[c] *= var;
        }
    }
#else
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const double ov = tc[2 * NF + 4 * f], ob = tc[2 * NF + 4 * f + 1],
                     o2 = tc[2 * NF + 4 * f + 2];
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            const double var = fma(e2[c], md.cv(c, f), ov);
            const double r = rcp_fast(var);
            const double mf = md.m(c, f);
            const double tm = mf * r;
            fot[c] = fma(tm, ob, fot[c]);
            ftt[c] = fma(tm, mf, ftt[c]);
            foo[c] = fma(o2, r, foo[c]);
            P[c] *= var;
        }
    }
#endif
    // chi2 and (det)^-1/2 (utils_cy.pyx:111-118)
#pragma unroll
    for (int c = 0; c < ZJ; ++c) {
        double rF;
        if (NF > REGNF) {
            // many bands: the determinant is kept as two factors so that it cannot leave the
            // double range (model variance >> observed variance in every band), and 1/FTT has
            // its own reciprocal
            y[c] = rsqrt_fast(P[c] * ftt[c]) * rsqrt_fast(P2[c]);
            rF = rcp_fast(ftt[c]);
        } else {
            y[c] = rsqrt_fast(P[c] * ftt[c]);
            rF = (y[c] * y[c]) * P[c];               // 1 / FTT
        }
        const double chi2 = fma(-(fot[c] * rF), fot[c], foo[c]);
        x[c] = fma(hc, chi2, xmask[c]);              // -chi2/2 in the target's units
    }
}

template <int NF, bool ZDEP, int ZJ>
__global__ void __launch_bounds__(FTHREADS, DL_MINB) fused_like_kernel(const FusedParams p) {
    constexpr int ROWS = 2 * NF + 1;
    constexpr int TCS = tc_stride(NF);
    constexpr int STG = fused_stages(NF, ZDEP);                 // ring depth
    constexpr int PREF = (FPREF < STG - 2 ? FPREF : STG - 2) > 1 ? (FPREF < STG - 2 ? FPREF : STG - 2) : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);        // [STG][ROWS][FZC]
    double* pdf_sm = stage + STG * ROWS * FZC;              // [FTPC][FZC]
    double* tc_sm = pdf_sm + FTPC * FZC;                        // [FTPC][TCS]
    double* rv_sm = tc_sm + FTPC * TCS;                         // [FTPC][2]  (R, V) per target
    double* wz_sm = rv_sm + FTPC * 2;                           // [FZC]
    double* zd_sm = wz_sm + FZC;                                // [4][FZC] (ZDEP only)
    uint64_t* full = reinterpret_cast<uint64_t*>(zd_sm + (ZDEP ? 4 * FZC : 0));
    uint64_t* empty = full + STG;
    const double* etab = reinterpret_cast<const double*>(empty + STG);   // [EXP_TAB] 2^(j/32)

    // Target tile fastest, z-chunk slowest: the ~300 CTAs resident at any time then work on at
    // most two z-chunks, i.e. they stream the SAME 13-row x 128-redshift slices of the model
    // grid within a few hundred training galaxies of each other, so the slices are read from
    // HBM once per wave and served from L2 to everybody else (chunk-fastest order: 48 % L2 hit
    // rate and 245 GB of DRAM reads per launch at config 3, see profiles/).
    const int notile = gridDim.x / p.nchunk;
    const int chunk = blockIdx.x / notile;
    const int otile = blockIdx.x % notile;
    const int split = blockIdx.y;
    const int z0 = chunk * FZC;
    const int t_begin = (int)(((long long)split * p.nT) / p.nsplit);
    const int t_end = (int)(((long long)(split + 1) * p.nT) / p.nsplit);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STG; ++s) {
            mbar_init(&full[s], DL_DISTPROD ? FW : 1);
            mbar_init(&empty[s], FW);
        }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < FTPC * TCS; i += FTHREADS) {
        int o = otile * FTPC + i / TCS;
        o = o < p.nO ? o : p.nO - 1;
        tc_sm[i] = p.tconst[(size_t)o * TCS + (i % TCS)];
    }
    exp_tab_init(const_cast<double*>(etab), threadIdx.x, FTHREADS);
    for (int i = threadIdx.x; i < FTPC * FZC; i += FTHREADS) pdf_sm[i] = 0.0;
    for (int i = threadIdx.x; i < FTPC; i += FTHREADS) { rv_sm[2 * i] = -1e300; rv_sm[2 * i + 1] = 0.0; }
    for (int i = threadIdx.x; i < FZC; i += FTHREADS) {
        wz_sm[i] = p.wz[z0 + i];
        if (ZDEP) {
#pragma unroll
            for (int k = 0; k < 4; ++k) zd_sm[k * FZC + i] = p.zdep[(size_t)k * p.nzp + z0 + i];
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------------- TMA producer
    // No dedicated producer warp (it would cost 32 registers/thread of occupancy): warp 0
    // refills, before it computes step `it`, the stage every warp released PREF steps ago.
    // (Measured alternatives: deeper rings 3-6 % slower; prefetch distance 1 or 3 no change.)
    const int nsteps = t_end - t_begin;
#ifdef DL_DBGWAIT
    long long dbg_empty = 0, dbg_issue = 0, dbg_wait = 0, dbg_div = 0;
    const long long dbg_t0 = clock64();
#endif
    auto issue_stage = [&](int step) {
        const int s = step % STG;
        const uint32_t ph = (step / STG) & 1;
        const size_t row = (size_t)p.nzp;
#ifdef DL_DBGWAIT
        const long long ta = clock64();
#endif
        // The WHOLE warp polls the barrier and only the straight-line copy issue below is
        // lane-0 code: a single lane spinning in a wait loop while the other 31 are parked at the
        // reconvergence point leaves the warp split for good (measured with __activemask():
        // lane 0 and lanes 1-31 then run the entire compute phase as two fragments, i.e. the
        // warp issues every instruction twice -- this, not the producer's work, was the
        // "2-2.8x slower" of all round-1 producer variants, see tools/ubench_wait.cu).
        mbar_wait(&empty[s], ph ^ 1);
#ifdef DL_DBGWAIT
        const long long tb = clock64();
        dbg_empty += tb - ta;
#endif
        if (DL_ELECT ? elect_one() : lane == 0) {
#if DL_DISTPROD
        const int nrows = warp < ROWS ? (ROWS - warp + FW - 1) / FW : 0;
        if (nrows > 0) mbar_arrive_expect_tx(&full[s], nrows * FZC * sizeof(double));
        else mbar_arrive(&full[s]);
        const double* gt = p.grid + (size_t)(t_begin + step) * 2 * p.nB * row + z0;
        double* dst = stage + s * ROWS * FZC;
#pragma unroll
        for (int k = 0; k < (ROWS + FW - 1) / FW; ++k) {
            const int r = warp + k * FW;
            if (r < ROWS) {
                const double* src = r < NF ? gt + (size_t)p.bands[r] * row
                                  : r < 2 * NF ? gt + (size_t)(p.nB + p.bands[r - NF]) * row
                                  : p.prior + (size_t)(t_begin + step) * row + z0;
                tma_bulk_g2s(dst + r * FZC, src, FZC * sizeof(double), &full[s]);
            }
        }
#else
        mbar_arrive_expect_tx(&full[s], ROWS * FZC * sizeof(double));
        const double* gt = p.grid + (size_t)(t_begin + step) * 2 * p.nB * row + z0;
        double* dst = stage + s * ROWS * FZC;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            tma_bulk_g2s(dst + f * FZC, gt + (size_t)p.bands[f] * row, FZC * sizeof(double),
                         &full[s]);
            tma_bulk_g2s(dst + (NF + f) * FZC, gt + (size_t)(p.nB + p.bands[f]) * row,
                         FZC * sizeof(double), &full[s]);
        }
        tma_bulk_g2s(dst + 2 * NF * FZC, p.prior + (size_t)(t_begin + step) * row + z0,
                     FZC * sizeof(double), &full[s]);
#endif
        }
        __syncwarp();
#ifdef DL_DBGWAIT
        dbg_issue += clock64() - tb;
#endif
    };
    if (DL_DISTPROD || warp == DL_PRODWARP) {
        for (int step = 0; step < PREF && step < nsteps; ++step) issue_stage(step);
    }

    // ---------------------------------------------------------------------- compute
    // Each lane works on PAIRS of redshifts (z, z+32) of the chunk for one target at a time:
    // the target's constants are fetched once (broadcast LDS.128) for two cells and every
    // statement below carries two independent dependency chains for the FP64 pipe.
    double E[FTO];
#pragma unroll
    for (int o = 0; o < FTO; ++o) E[o] = 0.0;
#ifdef DL_JUNK
    unsigned junk = threadIdx.x;
#endif
    // targets of this warp: warp 0 owns FTO - DL_ASYM, the others FTO (contiguous in the tile)
#ifndef DL_LIGHTWARP
#define DL_LIGHTWARP 0
#endif
    const int nto = warp == DL_LIGHTWARP ? FTO - DL_ASYM : FTO;
    const int tfirst = warp <= DL_LIGHTWARP ? warp * FTO : warp * FTO - DL_ASYM;
    const double* tcw = tc_sm + tfirst * TCS;
    double* pdfw = pdf_sm + tfirst * FZC;
    double* rvw = rv_sm + tfirst * 2;
    const int og0 = otile * FTPC + tfirst;

    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
        const int s = it % STG;
        const uint32_t ph = (it / STG) & 1;
        if ((DL_DISTPROD || warp == DL_PRODWARP) && it + PREF < nsteps)    // warp-uniform
            issue_stage(it + PREF);
#ifdef DL_DBGWAIT
        const long long tw0 = clock64();
        mbar_wait(&full[s], ph);
        dbg_wait += clock64() - tw0;
#else
        mbar_wait(&full[s], ph);
#endif
        const double* st = stage + s * ROWS * FZC;
        // fully unrolled for the register-cached variants; rolled for NF > REGNF (code size and
        // compile time: 4 x 4 x NF bands otherwise)
#pragma unroll (NF > REGNF ? 1 : GRPUNROLL)
        for (int grp = 0; grp < FZT / ZJ; ++grp) {
            const int zl = grp * 32 * ZJ + lane;      // cells: zl + 32 c, c < ZJ
            constexpr bool SMODEL = NF > REGNF;           // model values stay in the stage
            constexpr int RN = SMODEL ? 1 : NF;
            double m[ZJ][RN], cv[ZJ][RN];
            if (!SMODEL) {
#pragma unroll
                for (int f = 0; f < RN; ++f) {
#pragma unroll
                    for (int c = 0; c < ZJ; ++c) {
                        m[c][f] = st[f * FZC + zl + 32 * c];
                        cv[c][f] = st[(NF + f) * FZC + zl + 32 * c];
                    }
                }
            }
            double pr[ZJ], wzv[ZJ], xmask[ZJ], zot[ZJ], ztt[ZJ], zoo[ZJ], zev[ZJ];
#pragma unroll
            for (int c = 0; c < ZJ; ++c) {
                pr[c] = st[2 * NF * FZC + zl + 32 * c];
                wzv[c] = wz_sm[zl + 32 * c];
                // padded redshifts get x = -inf: they never move R or V and contribute exactly 0
                xmask[c] = ((z0 + zl + 32 * c) < p.nz) ? 0.0 : -INFINITY;
                zot[c] = ztt[c] = zoo[c] = 0.0; zev[c] = 1.0;
                if (ZDEP) {
                    zot[c] = zd_sm[zl + 32 * c]; ztt[c] = zd_sm[FZC + zl + 32 * c];
                    zoo[c] = zd_sm[2 * FZC + zl + 32 * c]; zev[c] = zd_sm[3 * FZC + zl + 32 * c];
                }
            }
#pragma unroll (NF > REGNF ? 1 : FTO)
            for (int o = 0; o < FTO; ++o) {
                // E[] is rotated once per target so the loop can stay rolled with the
                // accumulators in registers (FTO rotations restore the order)
                double Eo = E[0];
                if (!(DL_ASYM && o == FTO - 1 && nto < FTO)) {     // warp-uniform
                const double* tc = tcw + o * TCS;
                double pOT0[ZJ], pOT[ZJ], pTT[ZJ], pOO[ZJ];
#pragma unroll
                for (int c = 0; c < ZJ; ++c) {
                    if (ZDEP) {
                        const double cc = tc[6 * NF + 5];
                        pOT0[c] = tc[2 * NF + 3] * zot[c];
                        pOT[c] = cc * zot[c]; pTT[c] = cc * ztt[c]; pOO[c] = cc * zoo[c];
                    } else {
                        pOT0[c] = tc[6 * NF + 1];
                        pOT[c] = tc[6 * NF + 2]; pTT[c] = tc[6 * NF + 3]; pOO[c] = tc[6 * NF + 4];
                    }
                }
                double x[ZJ], y[ZJ], sarg[ZJ];
                const double R = rvw[2 * o];
                if constexpr (SMODEL)
                    cells_xy<NF, ZJ>(tc, SmemModel<NF, ZJ>{st + zl}, pOT0, pOT, pTT, pOO, zev,
                                     xmask, x, y);
                else
                    cells_xy<NF, ZJ>(tc, RegModel<NF, ZJ>{m, cv}, pOT0, pOT, pTT, pOO, zev, xmask,
                                     x, y);
#pragma unroll
                for (int c = 0; c < ZJ; ++c) sarg[c] = x[c] - R;
#ifdef DL_DBGWAIT
                {
                    bool dead = true;
#pragma unroll
                    for (int c = 0; c < ZJ; ++c) dead &= sarg[c] < -708.0;
                    dbg_div += __all_sync(0xffffffffu, dead);
                }
#endif
                // Speculate that neither the reference exponent R nor the running max V moves
                // (they do O(log) times per target): evaluate exp first, then ONE vote covers
                // both conditions, so the tail of the cell has a single warp-wide
                // serialisation point.
                double v[ZJ];
#pragma unroll
                for (int c = 0; c < ZJ; ++c) v[c] = DL_EXP(sarg[c]) * y[c];   // :118,128
                const double V = rvw[2 * o + 1];
                bool up = false;
#pragma unroll
                for (int c = 0; c < ZJ; ++c)       // high words only: a superset of v > V
                    up |= (__double2hiint(sarg[c]) > 0) |
                          (__double2hiint(v[c]) >= __double2hiint(V));
                if (__any_sync(0xffffffffu, up)) {
                    double xl = x[0];
#pragma unroll
                    for (int c = 1; c < ZJ; ++c) xl = fmax(xl, x[c]);
                    const double xm = warp_max(xl);
                    double Vc = V;
                    if (xm > R) {
                        // the warp met a better cell than its reference exponent: rescale
                        const double fs = exp(R - xm);
                        Eo *= fs;
                        Vc *= fs;
#pragma unroll
                        for (int zj = 0; zj < FZT; ++zj) pdfw[o * FZC + zj * 32 + lane] *= fs;
                        __syncwarp();
                        if (lane == 0) { rvw[2 * o] = xm; rvw[2 * o + 1] = Vc; }
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < ZJ; ++c) v[c] = DL_EXP(x[c] - xm) * y[c];
                    }
                    double vl = v[0];
#pragma unroll
                    for (int c = 1; c < ZJ; ++c) vl = fmax(vl, v[c]);
                    const double vm = warp_max(vl);
                    __syncwarp();
                    if (lane == 0 && vm > Vc) rvw[2 * o + 1] = vm;
                    __syncwarp();
                }
#ifdef DL_JUNK
                // tuning probe: DL_JUNK extra integer instructions per cell group
#pragma unroll
                for (int j = 0; j < DL_JUNK; ++j)
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(junk) : "r"(j + lane));
#endif
                // prior, PDF and evidence accumulation  (delight-apply.py:145-148)
#pragma unroll
                for (int c = 0; c < ZJ; ++c) {
                    const double vp = v[c] * pr[c];
                    pdfw[o * FZC + zl + 32 * c] += vp;
                    Eo = fma(vp, wzv[c], Eo);
                }
                }
#pragma unroll
                for (int k = 0; k + 1 < FTO; ++k) E[k] = E[k + 1];
                E[FTO - 1] = Eo;
            }
        }
        // evidence of (target, training galaxy t) over this z-chunk
#if DL_FTO == 4 && !DL_ASYM
        {
            // four warp sums folded into one butterfly: after the xor-16 step a lane keeps two of
            // the four partial sums, after xor-8 one; lane 8 k ends up with the sum of target k
            // (6 shuffle + add steps instead of 20)
            const bool hi16 = lane & 16, hi8 = lane & 8;
            const double s0 = hi16 ? E[0] : E[2], s1 = hi16 ? E[1] : E[3];      // what I give away
            double k0 = hi16 ? E[2] : E[0], k1 = hi16 ? E[3] : E[1];            // what I keep
            k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            double ev = (hi8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
            ev += __shfl_xor_sync(0xffffffffu, ev, 4);
            ev += __shfl_xor_sync(0xffffffffu, ev, 2);
            ev += __shfl_xor_sync(0xffffffffu, ev, 1);
            const int o = lane >> 3;                   // (bit 4, bit 3) = (target >> 1, target & 1)
            const int og = og0 + o;
            if ((lane & 7) == 0 && og < p.nO)
                p.evid_part[((size_t)chunk * p.nO + og) * p.nT + t] =
                    make_double2(ev, rvw[2 * o]);
        }
#else
#pragma unroll
        for (int o = 0; o < FTO; ++o) E[o] = warp_sum(E[o]);
        {
            double ev = E[0];
#pragma unroll
            for (int o = 1; o < FTO; ++o) if (lane == o) ev = E[o];
            const int og = og0 + lane;
            if (lane < nto && og < p.nO)
                p.evid_part[((size_t)chunk * p.nO + og) * p.nT + t] =
                    make_double2(ev, rvw[2 * lane]);
        }
#endif
#pragma unroll
        for (int o = 0; o < FTO; ++o) E[o] = 0.0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

#ifdef DL_JUNK
    if (junk == 0xdeadbeefu && p.nO < 0) p.Rw[0] = 0.0;
#endif
#ifdef DL_DBGWAIT
    if (p.dbg && lane == 0) {
        long long* d = p.dbg + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * FW + warp) * 4;
        d[0] = dbg_wait; d[1] = clock64() - dbg_t0; d[2] = dbg_empty + dbg_issue; d[3] = dbg_div;
    }
#endif
#pragma unroll
    for (int o = 0; o < FTO; ++o) {
        const int og = og0 + o;
        if (o < nto && og < p.nO) {
            double* dst = p.pdf_part + ((size_t)split * p.nO + og) * p.nzp + z0;
#pragma unroll
            for (int zj = 0; zj < FZT; ++zj) dst[zj * 32 + lane] = pdfw[o * FZC + zj * 32 + lane];
            const double vfin = rvw[2 * o + 1];
            if (lane == 0) {
                const size_t k = ((size_t)split * p.nchunk + chunk) * p.nO + og;
                p.Rw[k] = rvw[2 * o];
                p.Vw[k] = vfin;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// per-target constants (see header comment): scaled by c = 2^round(mean log2 var)
// in:  obs/var [nO][NF] (already restricted to the group's NF bands), ell_hat/ell_var scalars
// out: tconst [nO][tc_stride(NF)], layout documented at tc_stride()
__global__ void prep_targets_kernel(int nO, int NF, const double* __restrict__ obs,
                                    const double* __restrict__ var, double ell_hat,
                                    double ell_var, double cov_scale,
                                    double* __restrict__ tconst) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nO) return;
    const int TCS = tc_stride(NF);
    double* tc = tconst + (size_t)o * TCS;
    double s = 0.0;
    for (int f = 0; f < NF; ++f) s += log2(var[(size_t)o * NF + f]);
    int k = (int)rint(s / NF);
    k = k < -1000 ? -1000 : (k > 1000 ? 1000 : k);
    const double c = ldexp(1.0, k), ic = ldexp(1.0, -k);
    const double sic = sqrt(ic * cov_scale);   // ell^2 * (cov_scale * model variance) / c
    for (int f = 0; f < NF; ++f) {
        const double ob = obs[(size_t)o * NF + f];
        const double vs = var[(size_t)o * NF + f] * ic;
        const double w = 1.0 / vs;
        tc[2 * f] = ob * w * sic;       // pass-0 weights carry sqrt(ic cov_scale): e2 = (FOT'/FTT)^2
        tc[2 * f + 1] = w;
        tc[2 * NF + 4 * f] = vs;
        tc[2 * NF + 4 * f + 1] = ob;
        tc[2 * NF + 4 * f + 2] = ob * ob;
        tc[2 * NF + 4 * f + 3] = 0.0;
    }
    tc[6 * NF + 0] = -0.5 * ic;
    tc[6 * NF + 1] = c * (ell_hat / ell_var) * sic;   // pass-0 start value of FOT'
    tc[2 * NF + 3] = c * sic;                          // same for z-dependent ell priors
    tc[6 * NF + 2] = c * (ell_hat / ell_var);
    tc[6 * NF + 3] = c / ell_var;
    tc[6 * NF + 4] = c * (ell_hat * ell_hat / ell_var);
    tc[6 * NF + 5] = c;
}

// combine the per-warp (R, V) of each target: Rfin = max R, Vfin = max V * exp(R - Rfin)
__global__ void finalize_scales_kernel(int nO, int nparts, const double* __restrict__ Rw,
                                       const double* __restrict__ Vw, double* __restrict__ Rfin,
                                       double* __restrict__ invV) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nO) return;
    double rmax = -INFINITY;
    for (int k = 0; k < nparts; ++k) rmax = fmax(rmax, Rw[(size_t)k * nO + o]);
    double vmax = 0.0;
    for (int k = 0; k < nparts; ++k)
        vmax = fmax(vmax, Vw[(size_t)k * nO + o] * exp(Rw[(size_t)k * nO + o] - rmax));
    Rfin[o] = rmax;
    invV[o] = 1.0 / vmax;
}

// pdf[o][z] = sum_split pdf_part * exp(R_part - Rfin) / Vfin     (out: [nO][ld_out])
__global__ void finalize_pdf_kernel(int nO, int nz, int nzp, int nchunk, int nsplit,
                                    const double* __restrict__ pdf_part,
                                    const double* __restrict__ Rw,
                                    const double* __restrict__ Rfin,
                                    const double* __restrict__ invV, double* __restrict__ pdf,
                                    long long ld_out) {
    const int z = blockIdx.y * blockDim.x + threadIdx.x;
    const int o = blockIdx.x;
    if (z >= nz) return;
    const int chunk = z / FZC;
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s) {
        const double r = Rw[((size_t)s * nchunk + chunk) * nO + o];
        acc += pdf_part[((size_t)s * nO + o) * nzp + z] * exp(r - Rfin[o]);
    }
    pdf[(size_t)o * ld_out + z] = acc * invV[o];
}

// evid[o][t] = sum_chunk E * exp(R - Rfin) / Vfin
__global__ void finalize_evid_kernel(int nO, int nT, int nchunk,
                                     const double2* __restrict__ evid_part,
                                     const double* __restrict__ Rfin,
                                     const double* __restrict__ invV,
                                     double* __restrict__ evid) {
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    const int o = blockIdx.x;
    if (t >= nT) return;
    const double rf = Rfin[o];
    double acc = 0.0;
    for (int c = 0; c < nchunk; ++c) {
        const double2 e = evid_part[((size_t)c * nO + o) * nT + t];
        acc += e.x * exp(e.y - rf);
    }
    evid[(size_t)o * nT + t] = acc * invV[o];
}

// ------------------------------------------------------------------------------------------
// top-K training galaxies per target by evidence (scripts/delight-apply.py:150-153).
// Order: descending evidence, equal evidences -> larger index first
// (= np.argsort(kind='stable')[::-1]).  One CTA per target:
// 8 x 8-bit radix-select passes over the order-preserving 64-bit keys, gather, bitonic sort.
constexpr int TK_THREADS = 256;
constexpr int TK_MAXK = 1024;

__device__ __forceinline__ unsigned long long order_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(TK_THREADS) topk_kernel(int nT, int K,
                                                          const double* __restrict__ evid,
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
        top_ev[(size_t)o * K + i] = t >= 0 ? row[t] : 0.0;
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

template <int NF, bool ZDEP>
static int launch_fused_t(const FusedParams& p, cudaStream_t stream) {
    constexpr int ZJ = NF > REGNF ? DL_SMODEL_ZJ : (NF <= 8 ? ZJ_SMALL : ZJ_LARGE);
    const size_t smem = fused_smem_bytes(NF, ZDEP, fused_stages(NF, ZDEP));
    auto kern = fused_like_kernel<NF, ZDEP, ZJ>;
    DL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    const int notile = (p.nO + FTPC - 1) / FTPC;
    dim3 grid(p.nchunk * notile, p.nsplit);
    if (dl_timing_enabled()) dl_timing_mark(stream, 0);
    kern<<<grid, FTHREADS, smem, stream>>>(p);
    if (dl_timing_enabled()) dl_timing_mark(stream, 1);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

template <bool ZDEP>
static int launch_fused_nf(int NF, const FusedParams& p, cudaStream_t stream) {
    switch (NF) {
        case 1: return launch_fused_t<1, ZDEP>(p, stream);
        case 2: return launch_fused_t<2, ZDEP>(p, stream);
        case 3: return launch_fused_t<3, ZDEP>(p, stream);
        case 4: return launch_fused_t<4, ZDEP>(p, stream);
        case 5: return launch_fused_t<5, ZDEP>(p, stream);
        case 6: return launch_fused_t<6, ZDEP>(p, stream);
        case 7: return launch_fused_t<7, ZDEP>(p, stream);
        case 8: return launch_fused_t<8, ZDEP>(p, stream);
        case 9: return launch_fused_t<9, ZDEP>(p, stream);
        case 10: return launch_fused_t<10, ZDEP>(p, stream);
        case 11: return launch_fused_t<11, ZDEP>(p, stream);
        case 12: return launch_fused_t<12, ZDEP>(p, stream);
        case 13: return launch_fused_t<13, ZDEP>(p, stream);
        case 14: return launch_fused_t<14, ZDEP>(p, stream);
        case 15: return launch_fused_t<15, ZDEP>(p, stream);
        case 16: return launch_fused_t<16, ZDEP>(p, stream);
        case 17: return launch_fused_t<17, ZDEP>(p, stream);
        case 18: return launch_fused_t<18, ZDEP>(p, stream);
        case 19: return launch_fused_t<19, ZDEP>(p, stream);
        case 20: return launch_fused_t<20, ZDEP>(p, stream);
        case 21: return launch_fused_t<21, ZDEP>(p, stream);
        case 22: return launch_fused_t<22, ZDEP>(p, stream);
        case 23: return launch_fused_t<23, ZDEP>(p, stream);
        case 24: return launch_fused_t<24, ZDEP>(p, stream);
        case 25: return launch_fused_t<25, ZDEP>(p, stream);
        case 26: return launch_fused_t<26, ZDEP>(p, stream);
        case 27: return launch_fused_t<27, ZDEP>(p, stream);
        case 28: return launch_fused_t<28, ZDEP>(p, stream);
        case 29: return launch_fused_t<29, ZDEP>(p, stream);
        case 30: return launch_fused_t<30, ZDEP>(p, stream);
        case 31: return launch_fused_t<31, ZDEP>(p, stream);
        case 32: return launch_fused_t<32, ZDEP>(p, stream);
    }
    dl_set_error("fused likelihood supports 1..%d bands per target group, got %d", MAXNF, NF);
    return DL_ERR_UNSUPPORTED;
}

}  // namespace dl

using namespace dl;

// ============================================================================ C ABI (device)
extern "C" {

int dl_fused_chunk() { return FZC; }
int dl_fused_targets_per_cta() { return FTPC; }
// CTAs of the fused kernel that share an SM for nf bands (2 up to 13 bands, else 1)
int dl_fused_ctas_per_sm(int nf, int zdep) {
    if (nf < 1 || nf > MAXNF) return 0;
    const int s = fused_stages(nf, zdep != 0);
    return 2 * (fused_smem_bytes(nf, zdep != 0, s) + 1024) <= 228 * 1024 ? 2 : 1;
}

long long dl_fused_padded_nz(long long nz) { return (nz + FZC - 1) / FZC * FZC; }

// bytes of scratch dl_like_fused_dev needs for (nO targets, nT training galaxies, nz, nsplit)
long long dl_fused_workspace_bytes(long long nO, long long nT, long long nz, int nsplit,
                                   int want_evid) {
    const long long nzp = dl_fused_padded_nz(nz);
    const long long nchunk = nzp / FZC;
    auto al = [](long long x) { return (x + 255) & ~255ll; };
    long long b = 0;
    b += al((long long)tc_stride(MAXNF) * nO * 8);      // tconst
    b += al((long long)nsplit * nO * nzp * 8);          // pdf_part
    b += al(nchunk * nO * nT * 16);                     // evid_part
    b += 2 * al((long long)nsplit * nchunk * nO * 8);   // Rw, Vw
    b += 2 * al(nO * 8);                                // Rfin, invV
    b += al(nzp * 8) + al(4 * nzp * 8);                 // wz, zdep
    if (!want_evid) b += al(nO * nT * 8);               // evid (internal)
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

// Fused likelihood + apply epilogue for one group of targets sharing the same band list.
//   obs, var     [nO][nf]   target fluxes / variances restricted to `bands`
//   bands        [nf]       HOST array of band indices into the grid (0 <= b < nB)
//   grid         [nT][2nB][nzp] device, prior [nT][nzp] device (nzp = dl_fused_padded_nz(nz))
//   zgrid        [nz] device
//   ell_hat/ell_var: scalars, or (ell_hat_z, ell_var_z != NULL) device vectors [nz]
//   cov_scale    factor on the model variance (1 in delight-apply; V_C in delight-optimize.py:122)
//   outputs (device): pdf [nO][ld_pdf]; evid [nO][nT] or NULL; top_idx [nO][K] int64,
//                     top_ev [nO][K] (K may be 0)
int dl_like_fused_dev(long long nO, long long nT, long long nz, long long nB, int nf,
                      const int* bands, const double* obs, const double* var,
                      const double* grid, const double* prior, const double* zgrid,
                      double ell_hat, double ell_var, const double* ell_hat_z,
                      const double* ell_var_z, double cov_scale, int nsplit, int K, double* pdf,
                      long long ld_pdf,
                      double* evid, long long* top_idx, double* top_ev, void* workspace,
                      long long workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DL_REQUIRE(nO > 0 && nT > 0 && nz > 1 && nB > 0, "bad shape");
    DL_REQUIRE(nf >= 1 && nf <= MAXNF, "1..32 bands per target group");
    DL_REQUIRE(nsplit >= 1 && nsplit <= nT, "bad nsplit");
    DL_REQUIRE(K >= 0 && K <= TK_MAXK && K <= nT, "0 <= K <= min(nT, 1024)");
    DL_REQUIRE(ld_pdf >= nz, "ld_pdf < nz");
    DL_REQUIRE(cov_scale >= 0.0, "cov_scale < 0");
    const bool zdep = ell_hat_z != nullptr && ell_var_z != nullptr;
    const long long need = dl_fused_workspace_bytes(nO, nT, nz, nsplit, evid != nullptr || K == 0);
    DL_REQUIRE(workspace && workspace_bytes >= need, "workspace too small");
    const int nzp = (int)dl_fused_padded_nz(nz);
    const int nchunk = nzp / FZC;

    char* w = (char*)workspace;
    auto take = [&](size_t bytes) { char* r = w; w += (bytes + 255) & ~(size_t)255; return r; };
    double* tconst = (double*)take((size_t)tc_stride(MAXNF) * nO * 8);
    double* pdf_part = (double*)take((size_t)nsplit * nO * nzp * 8);
    double2* evid_part = (double2*)take((size_t)nchunk * nO * nT * 16);
    double* Rw = (double*)take((size_t)nsplit * nchunk * nO * 8);
    double* Vw = (double*)take((size_t)nsplit * nchunk * nO * 8);
    double* Rfin = (double*)take((size_t)nO * 8);
    double* invV = (double*)take((size_t)nO * 8);
    double* wz = (double*)take((size_t)nzp * 8);
    double* zd = (double*)take((size_t)4 * nzp * 8);
    double* evid_int = evid ? evid : (K > 0 ? (double*)take((size_t)nO * nT * 8) : nullptr);
    DL_REQUIRE((size_t)(w - (char*)workspace) <= (size_t)workspace_bytes, "workspace accounting");

    zaux_kernel<<<(unsigned)((nzp + 127) / 128), 128, 0, stream>>>(
        (int)nz, nzp, zgrid, zdep ? ell_hat_z : nullptr, zdep ? ell_var_z : nullptr, wz, zd);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();

    prep_targets_kernel<<<(unsigned)((nO + 127) / 128), 128, 0, stream>>>(
        (int)nO, nf, obs, var, zdep ? 1.0 : ell_hat, zdep ? 1.0 : ell_var, cov_scale, tconst);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();

    FusedParams p;
    p.grid = grid; p.prior = prior; p.tconst = tconst; p.wz = wz; p.zdep = zd;
    p.pdf_part = pdf_part; p.evid_part = evid_part; p.Rw = Rw; p.Vw = Vw;
    for (int f = 0; f < MAXNF; ++f) {
        p.bands[f] = f < nf ? bands[f] : 0;
        DL_REQUIRE(p.bands[f] >= 0 && p.bands[f] < nB, "band index out of range");
    }
    p.nB = (int)nB; p.nz = (int)nz; p.nzp = nzp; p.nT = (int)nT; p.nO = (int)nO;
    p.nchunk = nchunk; p.nsplit = nsplit;
    p.dbg = nullptr;
#ifdef DL_DBGWAIT
    const size_t dbg_nb = (size_t)nchunk * ((nO + FTPC - 1) / FTPC) * nsplit;
    {
        static long long* dbg = nullptr;
        if (dbg) cudaFree(dbg);
        cudaMalloc(&dbg, dbg_nb * FW * 4 * 8);
        cudaMemset(dbg, 0, dbg_nb * FW * 4 * 8);
        p.dbg = dbg;
    }
#endif
    int rc = zdep ? launch_fused_nf<true>(nf, p, stream) : launch_fused_nf<false>(nf, p, stream);
    if (rc != DL_OK) return rc;

    finalize_scales_kernel<<<(unsigned)((nO + 127) / 128), 128, 0, stream>>>(
        (int)nO, nsplit * nchunk, Rw, Vw, Rfin, invV);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    {
        dim3 g((unsigned)nO, (unsigned)((nz + 127) / 128));
        finalize_pdf_kernel<<<g, 128, 0, stream>>>((int)nO, (int)nz, nzp, nchunk, nsplit,
                                                   pdf_part, Rw, Rfin, invV, pdf, ld_pdf);
        DL_CUDA_CHECK(cudaGetLastError());
        dl_count_launch();
    }
    if (evid_int) {
        dim3 g((unsigned)nO, (unsigned)((nT + 127) / 128));
        finalize_evid_kernel<<<g, 128, 0, stream>>>((int)nO, (int)nT, nchunk, evid_part, Rfin,
                                                    invV, evid_int);
        DL_CUDA_CHECK(cudaGetLastError());
        dl_count_launch();
    }
#ifdef DL_DBGWAIT
    {
        std::vector<long long> h(dbg_nb * FW * 4);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), p.dbg, h.size() * 8, cudaMemcpyDeviceToHost);
        double w[FW] = {0}, tt[FW] = {0}, em[FW] = {0}, is[FW] = {0};
        for (size_t b = 0; b < dbg_nb; ++b)
            for (int q = 0; q < FW; ++q) {
                w[q] += h[(b * FW + q) * 4]; tt[q] += h[(b * FW + q) * 4 + 1];
                em[q] += h[(b * FW + q) * 4 + 2]; is[q] += h[(b * FW + q) * 4 + 3];
            }
        fprintf(stderr, "[dbgwait] per warp: full-wait / producer work (fractions of lifetime) / warp-iterations with all 64 cells underflowing, per CTA:");
        for (int q = 0; q < FW; ++q)
            fprintf(stderr, " w%d=%.3f/%.3f/%.0f", q, w[q] / tt[q], em[q] / tt[q], is[q] / dbg_nb);
        fprintf(stderr, " | CTA lifetime %.3g cyc\n", tt[0] / dbg_nb);
    }
#endif
    if (K > 0) {
        DL_REQUIRE(top_idx && top_ev, "top-K outputs missing");
        topk_kernel<<<(unsigned)nO, TK_THREADS, 0, stream>>>((int)nT, K, evid_int, top_idx,
                                                            top_ev);
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
    ci_levels_kernel<<<(unsigned)((nO + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        (int)nO, (int)nz, pdf, ld_pdf, zgrid, nfrac, fractions, numlevels, out);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

}  // extern "C"

// Fused photo-z likelihood for a batch of targets (subsystem 1 of the north star): the kernel.
//
// Replaces the per-target host loop of scripts/delight-apply.py:114-162:
//     approx_flux_likelihood_cy (delight/utils_cy.pyx:83-128)  ->  like[nz][nT]
//     like *= prior; pdf = sum_t like; evid = trapz_z like; top-K by evidence
// The like[nz][nT] array of a target never exists in HBM, and (round 2) neither do per-z-chunk
// evidence partials: one CTA owns ALL redshifts of its cells.
//
// Work decomposition (B200: 148 SMs, FP64 pipe = 16 lanes / SMSP, 227 KB smem / SM)
//   unit     = one tile of targets (32 / 64)  x  one block of TT training galaxies  x  ALL z
//   CTA      = one unit; warps own 4 targets each; the TMA producer role rotates over the warps
//              config 0 (<= 8 bands, homogeneous band set): 8 warps, 128-redshift stages, 2 CTAs/SM
//              config 1 (more bands / per-target band lists): 16 warps, 64-redshift stages, 1 CTA/SM
//   step     = (z-chunk c, training galaxy k): one pipeline stage of 2 nU + 2 rows (nU band means,
//              nU band variances, the prior row, the trapezoid-weight row) of FZC doubles, fetched
//              with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a
//              full/empty mbarrier ring; steps run chunk-major (c outer, k inner)
//   lane     = FZC/32 redshifts of the chunk (z = z0 + 32 j + lane) -> conflict-free smem rows
//   compute  = per (target, training galaxy, z) cell, all in FP64 on the vector pipe: pass 0 with
//              the observed variances only, one reciprocal for the luminosity scale, pass 1 with
//              the model variance folded in (one reciprocal per PAIR of bands), chi2,
//              (prod var * FTT)^-1/2 by one rsqrt, one table-based exp.  No log, no division.
//   reduce   = PDF: shared-memory accumulator per (target, z of the chunk), flushed with
//              red.global.add.f64 at the end of every chunk (the PDF rows of the ~110 tiles in
//              flight stay in L2);  evidence: warp sum per step added to a shared-memory
//              accumulator per (target, k), complete after the last chunk, stored ONCE.
//
// Normalisation.  The reference subtracts the GLOBAL max log-likelihood of a target before exp
// (utils_cy.pyx:120-128).  Here every target gets a fixed ANCHOR exponent A before the main
// pass -- the max of -chi2/2 over a strided sample of training galaxies (the same kernel in
// SAMPLE mode, < 1 % of the work) -- folded into a per-target constant, so that every partial
// sum of every CTA is expressed in the same units and can simply be added.  The kernel tracks
// max(v) per target (the final divisor: results equal exp(L - max L) up to rounding) and whether
// any cell exceeded A by more than e^500; such targets (never seen with photometric S/N < 100)
// are recomputed by a rescue pass with the exact max as anchor (dl_like_fused_dev).
#pragma once
#include "common.cuh"

namespace dl {

constexpr int FTO = 4;        // targets per warp
constexpr int MAXNF = 32;     // bands per target
#ifndef DL_REGNF
#define DL_REGNF 7
#endif
constexpr int REGNF = DL_REGNF;   // up to here the model values of a z-group are cached in registers
constexpr int BIGNF = 8;      // beyond: band loops partially rolled, determinant in two factors
constexpr int ZJ = 2;         // redshifts a lane processes jointly (independent dependency chains)
#ifndef DL_STAGES
#define DL_STAGES 4
#endif
#ifndef DL_PREF
#define DL_PREF 2
#endif
#ifndef DL_ROTPROD
#define DL_ROTPROD 1
#endif
#ifndef DL_CHI2_EARLY
#define DL_CHI2_EARLY 1
#endif
#ifndef DL_BAND_UNROLL0
#define DL_BAND_UNROLL0 4     // pass-0 band loop beyond REGNF bands
#endif
#ifndef DL_BAND_UNROLL1
#define DL_BAND_UNROLL1 2     // pass-1 band-pair loop beyond REGNF bands
#endif
#ifndef DL_HYBRID
#define DL_HYBRID 0
#endif
#ifndef DL_TAILPAIR
#define DL_TAILPAIR 2
#endif
#ifndef DL_HCFOLD
#define DL_HCFOLD 1           // -1/2 and the unit of the variances folded into the target constants
#endif
#ifndef DL_PRW
#define DL_PRW 1              // evidence accumulated as v * (prior * weight)
#endif
#ifndef DL_MONT
#define DL_MONT 1             // band pairs share reciprocals (Montgomery's trick, up to BIGNF bands)
#endif
#ifndef DL_MONT_GROUP
#define DL_MONT_GROUP 2       // at most this many factors share a reciprocal (model in registers)
#endif
#ifndef DL_MONT_GROUP_S
#define DL_MONT_GROUP_S 2     // the same with the model values read from the stage
#endif
#ifndef DL_MONT_TJ2
#define DL_MONT_TJ2 0         // ... also where two targets are evaluated jointly (8 bands)
#endif
#ifndef DL_MONT_ROLL
#define DL_MONT_ROLL 1        // beyond BIGNF bands: two band pairs share a reciprocal
#endif
#ifndef DL_TJ2_NF
#define DL_TJ2_NF 8
#endif
#ifndef DL_CFG1_NF
#define DL_CFG1_NF 9          // homogeneous band sets from here on use config 1
#endif

#ifndef DL_GRP_ROLL_NF
#define DL_GRP_ROLL_NF 6       // from here on the two z-groups of a stage are a rolled loop
#endif
constexpr int GRP_ROLL_NF = DL_GRP_ROLL_NF;
#ifndef DL_TGT_UNROLL
#define DL_TGT_UNROLL 4        // target (pair) loop unrolled in the register-model kernels
#endif
constexpr int TGT_UNROLL = DL_TGT_UNROLL;
constexpr int BAND_UNROLL0 = DL_BAND_UNROLL0;   // (macros are not expanded inside #pragma unroll)
constexpr int BAND_UNROLL1 = DL_BAND_UNROLL1;

enum { MODE_MAIN = 0, MODE_SAMPLE = 1 };

__host__ __device__ constexpr int cfg_of(int nf, bool het) { return (het || nf >= DL_CFG1_NF) ? 1 : 0; }
__host__ __device__ constexpr int cfg_warps(int cfg) { return cfg ? 16 : 8; }
__host__ __device__ constexpr int cfg_fzc(int cfg) { return cfg ? 64 : 128; }
__host__ __device__ constexpr int cfg_tt(int cfg) { return 32; }                // default t-block

// per-target constants (doubles), s = sqrt(cov_scale / c), c = power of two near the variances:
//   [0, 2NF)      (a_f s, w_f)          pass 0:  a = obs / var', w = 1 / var'
//   [2NF, 4NF)    (var'_f, obs_f)       pass 1
//   [4NF, 5NF)    obs_f^2
// DL_HCFOLD: c is an ODD power of two, so that h = sqrt(1 / (2 c)) is a power of two as well, and
//   the pass-1 constants hold (var'_f, h obs_f), -(h obs_f)^2, TC_POT = h x, TC_POO = -x / (2 c)
//   - anchor: the sums come out as FOT' = h FOT and FOO'' = -FOO / (2 c) - A, and
//   x = -chi2 / (2 c) - A = FOT'^2 / FTT + FOO'' is one FMA behind FOT'^2 (no h_c multiplication,
//   no mask addition: padded cells are zeroed in the select of the exponential).
//   5NF + TC_*    scalars below
//   then (per-target band lists only) NF int32 element offsets of the target's band rows in a stage
enum { TC_HC = 0, TC_POT0, TC_POT, TC_PTT, TC_POO, TC_C, TC_CS, TC_XOFF, TC_POO0, TC_SPARE, TC_NCONST };
__host__ __device__ constexpr int tc_stride(int nf, bool het) {
    return (5 * nf + TC_NCONST + (het ? (nf + 1) / 2 : 0) + 1) & ~1;
}

struct FusedParams {
    const double* grid;        // [nT][2*nB][nzp]  rows 0..nB-1 mean, nB..2nB-1 variance
    const double* prior;       // [nT][nzp]
    const double* tconst;      // [nO][tc_stride]
    const double* wz;          // [nzp] trapezoid weights (0 in the padding)
    const double* zdep;        // [4][nzp]: ell_hat/ell_var, 1/ell_var, ell_hat^2/ell_var, ell_var
    double* pdf_acc;           // [nO][nzp]   zero-initialised, accumulated with red.add
    double* evid;              // [nO][nT]    raw evidences (units of the anchor)
    unsigned long long* Vbits; // [nO] bits of max v (non-negative doubles order like integers)
    unsigned long long* Akey;  // [nO] SAMPLE mode: ordered key of max x
    int* flag;                 // [nO] MAIN: 1 = a cell exceeded the anchor by > e^500
    int* nflag;                // [1]  != 0 if any flag is set
    int bands[MAXNF];          // band rows staged per step (the union of the targets' bands)
    int nU;                    // how many
    int nB, nz, nzp, nT, nO;
    int TT, nTB;               // training galaxies per unit, units per target tile
    int nTlist, t_stride;      // the units cover t = i * t_stride, i < nTlist
    int nTiles, G;             // target tiles, tiles per L2 group
    int stages;                // TMA ring depth
    int rescue;                // 1: only tiles with flagged targets run; only those are stored
};

__host__ __device__ inline size_t fused_smem_bytes(int cfg, int nf, bool het, bool zdep, int nU,
                                                   int stages, int TT) {
    const int fzc = cfg_fzc(cfg), ftpc = cfg_warps(cfg) * FTO;
    const int rows = 2 * nU + 2 + (zdep ? 4 : 0);
    return sizeof(double) * ((size_t)stages * rows * fzc + (size_t)ftpc * fzc +
                             (size_t)ftpc * tc_stride(nf, het) + (size_t)ftpc * TT + EXP_TAB + ftpc) +
           2 * stages * sizeof(uint64_t);
}

// x = -chi2/2 - A (in the target's units, + the padding mask) and y = (det)^-1/2 of ZJ cells of
// one target against one training galaxy: utils_cy.pyx:98-118 restructured (header comment).
// Where the band means / variances of the cells come from: registers (loaded once per z-group and
// reused by the warp's targets; homogeneous band sets up to REGNF bands) or the stage in shared
// memory (more bands: 2 NF ZJ doubles do not fit the register budget; per-target band lists: the
// rows differ from target to target).
template <int NF>
struct RegModel {
    static constexpr bool kInRegs = true;
    const double (&mm)[ZJ][NF];
    const double (&cc)[ZJ][NF];
    __device__ __forceinline__ double m(int c, int f) const { return mm[c][f]; }
    __device__ __forceinline__ double cv(int c, int f) const { return cc[c][f]; }
};
// band means in registers, band variances read from the stage: half the register footprint of
// RegModel, which makes room for two targets jointly (DL_HYBRID)
template <int NF, int FZC>
struct HybridModel {
    static constexpr bool kInRegs = true;
    const double (&mm)[ZJ][NF];
    const double* stz;      // stage + the lane's first redshift
    __device__ __forceinline__ double m(int c, int f) const { return mm[c][f]; }
    __device__ __forceinline__ double cv(int c, int f) const { return stz[(NF + f) * FZC + 32 * c]; }
};
template <int NF, int FZC>
struct SmemModel {
    static constexpr bool kInRegs = false;
    const double* stz;      // stage + the lane's first redshift
    __device__ __forceinline__ double m(int c, int f) const { return stz[f * FZC + 32 * c]; }
    __device__ __forceinline__ double cv(int c, int f) const { return stz[(NF + f) * FZC + 32 * c]; }
};
template <int NF, int FZC>
struct SmemModelHet {
    static constexpr bool kInRegs = false;
    const double* stz;      // stage + the lane's first redshift
    const int* off;         // the target's NF row offsets (elements) into a stage
    int var0;               // nU * FZC: first variance row
    __device__ __forceinline__ double m(int c, int f) const { return stz[off[f] + 32 * c]; }
    __device__ __forceinline__ double cv(int c, int f) const { return stz[off[f] + var0 + 32 * c]; }
};

// TJ targets are evaluated jointly against the same ZJ cells: every model value is read once
// for all of them (with the model in shared memory that halves the dominant LDS traffic) and
// every statement carries TJ x ZJ independent dependency chains.
template <int NF, int TJ, bool ADDMASK, class Model>
__device__ __forceinline__ void cells_xy(const double* const (&tcs)[TJ], const Model md,
                                         const double (&pOT0)[TJ][ZJ], const double (&pOT)[TJ][ZJ],
                                         const double (&pTT)[TJ][ZJ], const double (&pOO)[TJ][ZJ],
                                         const double (&zev)[ZJ], const double (&xmask)[TJ][ZJ],
                                         double (&x)[TJ][ZJ], double (&y)[TJ][ZJ]) {
    const double2* tc0[TJ];     // (a s, w)
    const double2* tc1[TJ];     // (var', obs)
    const double* tc2[TJ];      // obs^2
#pragma unroll
    for (int j = 0; j < TJ; ++j) {
        tc0[j] = reinterpret_cast<const double2*>(tcs[j]);
        tc1[j] = reinterpret_cast<const double2*>(tcs[j] + 2 * NF);
        tc2[j] = tcs[j] + 4 * NF;
    }
    // pass 0: observed variances only (utils_cy.pyx:101-113 with ellML = 0)
    double fot[TJ][ZJ], ftt[TJ][ZJ], foo[TJ][ZJ], P[TJ][ZJ], P2[TJ][ZJ], e2[TJ][ZJ];
#pragma unroll
    for (int j = 0; j < TJ; ++j)
#pragma unroll
        for (int c = 0; c < ZJ; ++c) { fot[j][c] = pOT0[j][c]; ftt[j][c] = pTT[j][c]; }
    // (band loops: fully unrolled with the model in registers -- the register arrays need
    // static indices; partially unrolled beyond REGNF bands, where full unrolling lets the
    // compiler hoist the loads of all bands and spill: 21 bands, two targets jointly: 25 LDL +
    // 21 STL per 64 cells and 16 % long-scoreboard stalls, profiles/r2_fused_nf21_ncu_summary.txt)
    constexpr bool ROLL = NF > BIGNF;
#pragma unroll (ROLL ? BAND_UNROLL0 : NF)
    for (int f = 0; f < NF; ++f) {
        double mf[ZJ];
#pragma unroll
        for (int c = 0; c < ZJ; ++c) mf[c] = md.m(c, f);
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const double2 aw = tc0[j][f];
#pragma unroll
            for (int c = 0; c < ZJ; ++c) {
                const double tw = mf[c] * aw.y;
                fot[j][c] = fma(mf[c], aw.x, fot[j][c]);
                ftt[j][c] = fma(tw, mf[c], ftt[j][c]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < TJ; ++j)
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            const double ell = fot[j][c] * rcp_fast(ftt[j][c]);     // ellML * sqrt(cov_scale / c)
            e2[j][c] = ell * ell;
            fot[j][c] = pOT[j][c]; ftt[j][c] = pTT[j][c]; foo[j][c] = pOO[j][c];
            P[j][c] = zev[c]; P2[j][c] = 1.0;
        }
    // pass 1: variance inflated by ell^2 * model variance (utils_cy.pyx:107-112).  Bands in
    // pairs: ONE reciprocal of the product of the pair's variances gives both 1/var
    // (r_a = var_b / (var_a var_b)) and the pair's factor of the determinant: 9 FP64
    // instructions + 1 MUFU per pair instead of 10 + 2.
    auto acc = [&](int j, int c, int f, double mf, double r, double ob) {
        const double tm = mf * r;
        fot[j][c] = fma(tm, ob, fot[j][c]);
        ftt[j][c] = fma(tm, mf, ftt[j][c]);
        foo[j][c] = fma(tc2[j][f], r, foo[j][c]);
    };
    constexpr int NG = (NF / 2) * 2;
    // beyond REGNF bands the determinant is carried as two factors (pairs from F0 on go to P2)
    constexpr int F0 = ROLL ? ((NF / 2 + 1) & ~1) : NG;
    auto pair = [&](int f, double (&Pd)[TJ][ZJ]) {
        double ma[ZJ], mb[ZJ], ca[ZJ], cb[ZJ];
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            ma[c] = md.m(c, f); mb[c] = md.m(c, f + 1);
            ca[c] = md.cv(c, f); cb[c] = md.cv(c, f + 1);
        }
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const double2 voa = tc1[j][f], vob = tc1[j][f + 1];
#pragma unroll
            for (int c = 0; c < ZJ; ++c) {
                const double va = fma(e2[j][c], ca[c], voa.x);
                const double vb = fma(e2[j][c], cb[c], vob.x);
                const double pab = va * vb;
                const double ri = rcp_fast(pab);
                acc(j, c, f, ma[c], vb * ri, voa.y);
                acc(j, c, f + 1, mb[c], va * ri, vob.y);
                Pd[j][c] *= pab;
            }
        }
    };
    if constexpr (!ROLL && DL_MONT && (TJ == 1 || DL_MONT_TJ2)) {
        // One reciprocal per GROUP of band pairs (Montgomery's trick over the pair products q_i,
        // the odd last band being a factor of its own): the prefix products are the determinant
        // anyway and 1/q_i = (prod_{k<i} q_k) (prod_{k>i} q_k) / prod_k q_k.  n factors: 2 (n - 1)
        // multiplications instead of n - 1 reciprocals (3 FP64 + MUFU + 1 move each).
        constexpr int NP = NF / 2, NQ = NP + (NF & 1);
        // factors per reciprocal: with the model in registers two (three keep 12 more doubles
        // alive across the reciprocal and spill: 1.131 s against 1.071 s at 6 bands)
        constexpr int MGW = Model::kInRegs ? DL_MONT_GROUP : DL_MONT_GROUP_S;
        constexpr int MG = MGW < NQ ? MGW : NQ;
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
#pragma unroll
            for (int g0 = 0; g0 < NQ; g0 += MG) {
                const int n = (NQ - g0) < MG ? (NQ - g0) : MG;          // factors g0 .. g0 + n - 1
                double v[ZJ][2 * MG], q[ZJ][MG], pre[ZJ][MG], R[ZJ], inv[ZJ][MG];
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const bool odd = (g0 + i) == NP;                    // the unpaired last band
                    const int f = 2 * (g0 + i);
#pragma unroll
                    for (int c = 0; c < ZJ; ++c) {
                        v[c][2 * i] = fma(e2[j][c], md.cv(c, f), tc1[j][f].x);
                        if (!odd) {
                            v[c][2 * i + 1] = fma(e2[j][c], md.cv(c, f + 1), tc1[j][f + 1].x);
                            q[c][i] = v[c][2 * i] * v[c][2 * i + 1];
                        } else {
                            q[c][i] = v[c][2 * i];
                        }
                        pre[c][i] = i ? pre[c][i - 1] * q[c][i] : q[c][i];
                    }
                }
#pragma unroll
                for (int c = 0; c < ZJ; ++c) {
                    P[j][c] *= pre[c][n - 1];
                    R[c] = rcp_fast(pre[c][n - 1]);
                }
#pragma unroll
                for (int i = n - 1; i >= 1; --i)
#pragma unroll
                    for (int c = 0; c < ZJ; ++c) {
                        inv[c][i] = pre[c][i - 1] * R[c];
                        R[c] *= q[c][i];
                    }
#pragma unroll
                for (int c = 0; c < ZJ; ++c) inv[c][0] = R[c];
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const bool odd = (g0 + i) == NP;
                    const int f = 2 * (g0 + i);
                    if (!odd) {
                        const double oa = tc1[j][f].y, ob = tc1[j][f + 1].y;
#pragma unroll
                        for (int c = 0; c < ZJ; ++c) {
                            acc(j, c, f, md.m(c, f), v[c][2 * i + 1] * inv[c][i], oa);
                            acc(j, c, f + 1, md.m(c, f + 1), v[c][2 * i] * inv[c][i], ob);
                        }
                    } else {
                        const double oa = tc1[j][f].y;
#pragma unroll
                        for (int c = 0; c < ZJ; ++c) acc(j, c, f, md.m(c, f), inv[c][i], oa);
                    }
                }
            }
        }
    } else {
        // four bands (two pairs) on one reciprocal: 1/q_a = q_b / (q_a q_b)
        auto pair2 = [&](int f, double (&Pd)[TJ][ZJ]) {
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                double v[ZJ][4], R[ZJ], ia[ZJ], ib[ZJ];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double vo = tc1[j][f + i].x;
#pragma unroll
                    for (int c = 0; c < ZJ; ++c) v[c][i] = fma(e2[j][c], md.cv(c, f + i), vo);
                }
#pragma unroll
                for (int c = 0; c < ZJ; ++c) {
                    const double qa = v[c][0] * v[c][1], qb = v[c][2] * v[c][3];
                    const double qq = qa * qb;
                    R[c] = rcp_fast(qq);
                    ia[c] = qb * R[c]; ib[c] = qa * R[c];
                    Pd[j][c] *= qq;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double ob = tc1[j][f + i].y;
#pragma unroll
                    for (int c = 0; c < ZJ; ++c)
                        acc(j, c, f + i, md.m(c, f + i), v[c][i ^ 1] * (i < 2 ? ia[c] : ib[c]), ob);
                }
            }
        };
        if constexpr (ROLL && DL_MONT_ROLL) {
            constexpr int A4 = F0 & ~3, B4 = F0 + ((NG - F0) & ~3);
#pragma unroll 1
            for (int f = 0; f < A4; f += 4) pair2(f, P);
            if (A4 < F0) pair(A4, P);
#pragma unroll 1
            for (int f = F0; f < B4; f += 4) pair2(f, P2);
            if (B4 < NG) pair(B4, P2);
        } else {
#pragma unroll (ROLL ? BAND_UNROLL1 : NF)
            for (int f = 0; f < F0; f += 2) pair(f, P);
#pragma unroll (ROLL ? BAND_UNROLL1 : NF)
            for (int f = F0; f < NG; f += 2) pair(f, P2);
        }
        if (NF - NG == 1) {
            constexpr int f = NF - 1;
            double ma[ZJ], ca[ZJ];
#pragma unroll
            for (int c = 0; c < ZJ; ++c) { ma[c] = md.m(c, f); ca[c] = md.cv(c, f); }
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                const double2 vo = tc1[j][f];
#pragma unroll
                for (int c = 0; c < ZJ; ++c) {
                    const double var = fma(e2[j][c], ca[c], vo.x);
                    acc(j, c, f, ma[c], rcp_fast(var), vo.y);
                    if (NF > BIGNF) P2[j][c] *= var; else P[j][c] *= var;
                }
            }
        }
    }
    // chi2 and (det)^-1/2 (utils_cy.pyx:111-118)
#pragma unroll
    for (int j = 0; j < TJ; ++j) {
        const double hc = tcs[j][5 * NF + TC_HC];
#pragma unroll
        for (int c = 0; c < ZJ; ++c) {
            double rF;
            if (NF > BIGNF) {
                // many bands: the determinant is kept as two factors so that it cannot leave
                // the double range (model variance >> observed variance in every band), and
                // 1/FTT has its own reciprocal
                y[j][c] = rsqrt_fast(P[j][c] * ftt[j][c]) * rsqrt_fast(P2[j][c]);
                rF = rcp_fast(ftt[j][c]);
            } else {
                y[j][c] = rsqrt_fast(P[j][c] * ftt[j][c]);
                rF = (y[j][c] * y[j][c]) * P[j][c];               // 1 / FTT
            }
#if DL_HCFOLD
            (void)hc;
            x[j][c] = fma(fot[j][c] * fot[j][c], rF, foo[j][c]);
            if (ADDMASK) x[j][c] += xmask[j][c];
#else
#if DL_CHI2_EARLY
            // FOT^2 does not wait for 1/FTT: one dependent operation less behind the rsqrt
            const double chi2 = fma(-(fot[j][c] * fot[j][c]), rF, foo[j][c]);
#else
            const double chi2 = fma(-(fot[j][c] * rF), fot[j][c], foo[j][c]);
#endif
            x[j][c] = fma(hc, chi2, xmask[j][c]);     // -chi2/2 - A in the target's units
#endif
        }
    }
}

// 500.0 = 0x407F4000'00000000: a cell more than e^500 above the anchor flags its target
constexpr int HI_OVERFLOW = 0x407F4000;

template <int NF, bool ZDEP, bool HET, int MODE>
__global__ void __launch_bounds__(cfg_warps(cfg_of(NF, HET)) * 32, cfg_of(NF, HET) ? 1 : 2)
fused_like_kernel(const FusedParams p) {
    constexpr int CFG = cfg_of(NF, HET);
    constexpr int FW = cfg_warps(CFG), FZC = cfg_fzc(CFG), FZT = FZC / 32;
    constexpr int FTPC = FW * FTO, THREADS = FW * 32;
    constexpr int TCS = tc_stride(NF, HET);
    constexpr bool SMODEL = HET || NF > REGNF;        // model values stay in the stage
    constexpr bool SAMPLE = MODE == MODE_SAMPLE;
    // targets evaluated jointly: two from DL_TJ2_NF bands on, where the model values are read from
    // the stage for every use and sharing them pays (measured, 2048 x 1500 x 1500 cells, one vs
    // two targets: 13 bands 77.3 / 79.9 ms, 16: 92.5 / 98.3, 21: 131.6 / 126.5, 28: 178.5 / 173.1)
    constexpr bool HYBRID = DL_HYBRID && !SMODEL;     // means in registers, variances from the stage
    constexpr int TJ = ((SMODEL && !HET && NF >= DL_TJ2_NF) || HYBRID) ? 2 : 1;
    static_assert(FTO % TJ == 0, "TJ must divide the targets of a warp");
    // targets whose TAILS (exp, max tracking, accumulation) are evaluated together: the tail of
    // a cell is one long dependent chain, so with the model in registers (heads one target at a
    // time) the tails of DL_TAILPAIR targets are interleaved
    constexpr int TP = (!SMODEL && !SAMPLE && TJ < DL_TAILPAIR) ? DL_TAILPAIR : TJ;
    static_assert(FTO % TP == 0 && TP % TJ == 0, "bad tail pairing");
    const int nU = HET ? p.nU : NF;
    const int ROWS = 2 * nU + 2 + (ZDEP ? 4 : 0);
    const int STG = p.stages;
    const int PREF = (DL_PREF < STG - 2 ? DL_PREF : STG - 2) > 1 ? (DL_PREF < STG - 2 ? DL_PREF : STG - 2) : 1;
    const int TTs = p.TT;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);            // [STG][ROWS][FZC]
    double* pdf_sm = stage + (size_t)STG * ROWS * FZC;              // [FTPC][FZC]
    double* tc_sm = pdf_sm + FTPC * FZC;                            // [FTPC][TCS]
    double* E_sm = tc_sm + FTPC * TCS;                              // [FTPC][TT]
    double* etab_w = E_sm + (size_t)FTPC * TTs;                     // [EXP_TAB]
    unsigned long long* V_sm = reinterpret_cast<unsigned long long*>(etab_w + EXP_TAB);   // [FTPC]
    uint64_t* full = reinterpret_cast<uint64_t*>(V_sm + FTPC);
    uint64_t* empty = full + STG;
    const double* etab = etab_w;

    // unit -> (target tile, training block).  Tiles are grouped G at a time and, inside a group,
    // run tile-fastest / block-slowest: the CTAs in flight then share a few training blocks (the
    // model grid is read from HBM about nTiles / G times in all) and the PDF rows of one group
    // (G x tile x nzp x 8 B) stay in L2 between the flushes of successive blocks.
    int tile, tb;
    {
        const int u = blockIdx.x;
        const int gsz = p.G * p.nTB;
        const int grp = u / gsz, r = u - grp * gsz;
        const int left = p.nTiles - grp * p.G;
        const int tig = left < p.G ? left : p.G;
        tb = r / tig;
        tile = grp * p.G + (r - tb * tig);
    }
    const int ts0 = tb * p.TT;
    const int tcount = (p.nTlist - ts0) < p.TT ? (p.nTlist - ts0) : p.TT;
    const int nchunk = p.nzp / FZC;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (p.rescue) {
        // rescue passes: only tiles that hold a flagged target do anything
        int any = 0;
        if (*p.nflag != 0)
            for (int i = threadIdx.x; i < FTPC; i += THREADS) {
                const int o = tile * FTPC + i;
                any |= (o < p.nO) ? p.flag[o] : 0;
            }
        if (!__syncthreads_or(any)) return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < STG; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], FW);
        }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < FTPC * TCS; i += THREADS) {
        int o = tile * FTPC + i / TCS;
        o = o < p.nO ? o : p.nO - 1;
        tc_sm[i] = p.tconst[(size_t)o * TCS + (i % TCS)];
    }
    if (!SAMPLE) {
        exp_tab_init(etab_w, threadIdx.x, THREADS);
        for (int i = threadIdx.x; i < FTPC * FZC; i += THREADS) pdf_sm[i] = 0.0;
        for (int i = threadIdx.x; i < FTPC * TTs; i += THREADS) E_sm[i] = 0.0;
        for (int i = threadIdx.x; i < FTPC; i += THREADS) {
            const int o = tile * FTPC + i;
            // any lower bound of the final max is fine as a start; never below the smallest
            // normal double, so that cells that underflow (v = 0) do not take the slow path
            const unsigned long long v0 = o < p.nO ? p.Vbits[o] : 0ull;
            V_sm[i] = v0 > 0x0010000000000000ull ? v0 : 0x0010000000000000ull;
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------------- TMA producer
    // No dedicated producer warp (it would cost 32 registers/thread of occupancy): warp 0
    // refills, before it computes step `it`, the stage every warp released PREF steps ago.
    // The WHOLE warp polls the barrier and only the straight-line copy issue is single-lane
    // code (elect.sync): a lane spinning alone leaves its warp split for good (tools/ubench_wait.cu).
    // The producer keeps no cursor of its own: the step it issues is always `ahead` steps past
    // the consumer cursor (cs, cph, k, chunk) of the same warp.
    int cs = 0, cph = 0, k = 0, chunk = 0;        // consumer cursor: stage, phase, k, z-chunk
    auto issue_ahead = [&](int ahead) {
        int s = cs + ahead, ph = cph;
        if (s >= STG) { s -= STG; ph ^= 1; }
        int pk = k + ahead, pc = chunk;
        while (pk >= tcount) { pk -= tcount; ++pc; }
        if (pc >= nchunk) return;                 // past the last step (warp-uniform)
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
            const size_t row = (size_t)p.nzp;
            const size_t t = (size_t)(ts0 + pk) * p.t_stride;
            const int z0 = pc * FZC;
            mbar_arrive_expect_tx(&full[s], ROWS * FZC * sizeof(double));
            const double* gt = p.grid + t * 2 * p.nB * row + z0;
            double* dst = stage + (size_t)s * ROWS * FZC;
            if (HET) {
                for (int f = 0; f < nU; ++f) {
                    tma_bulk_g2s(dst + f * FZC, gt + (size_t)p.bands[f] * row,
                                 FZC * sizeof(double), &full[s]);
                    tma_bulk_g2s(dst + (nU + f) * FZC, gt + (size_t)(p.nB + p.bands[f]) * row,
                                 FZC * sizeof(double), &full[s]);
                }
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    tma_bulk_g2s(dst + f * FZC, gt + (size_t)p.bands[f] * row,
                                 FZC * sizeof(double), &full[s]);
                    tma_bulk_g2s(dst + (NF + f) * FZC, gt + (size_t)(p.nB + p.bands[f]) * row,
                                 FZC * sizeof(double), &full[s]);
                }
            }
            tma_bulk_g2s(dst + 2 * nU * FZC, p.prior + t * row + z0, FZC * sizeof(double),
                         &full[s]);
            tma_bulk_g2s(dst + (2 * nU + 1) * FZC, p.wz + z0, FZC * sizeof(double), &full[s]);
            if (ZDEP) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    tma_bulk_g2s(dst + (2 * nU + 2 + q) * FZC, p.zdep + (size_t)q * row + z0,
                                 FZC * sizeof(double), &full[s]);
            }
        }
        __syncwarp();
    };
    if (warp == 0)
        for (int j = 0; j < PREF; ++j) issue_ahead(j);
#if DL_ROTPROD
    int prod = 0;                                 // whose turn it is to issue (step mod FW)
#endif

    // ---------------------------------------------------------------------- compute
    // Each lane works on PAIRS of redshifts (z, z+32) of the chunk for one target at a time:
    // the target's constants are fetched once (broadcast LDS.128) for two cells and every
    // statement below carries two independent dependency chains for the FP64 pipe.
    double E[FTO];
#pragma unroll
    for (int o = 0; o < FTO; ++o) E[o] = SAMPLE ? -INFINITY : 0.0;
    int ovf = (int)0x80000000;                    // running max of the high words of x
    const int tfirst = warp * FTO;
    const double* tcw = tc_sm + tfirst * TCS;
    double* pdfw = pdf_sm + tfirst * FZC;
    const int og0 = tile * FTPC + tfirst;
    const int var0 = nU * FZC;

    while (chunk < nchunk) {
#if DL_ROTPROD
        // the producer role rotates over the warps step by step, so that no warp is the
        // permanent straggler the other seven wait for
        if (warp == prod) issue_ahead(PREF);                              // warp-uniform
        if (++prod == FW) prod = 0;
#else
        if (warp == 0) issue_ahead(PREF);                                 // warp-uniform
#endif
        mbar_wait(&full[cs], cph);
        const double* st = stage + (size_t)cs * ROWS * FZC;
        const int z0 = chunk * FZC;
        // z-group loop: unrolled for the small register-cached variants; rolled from 6 bands on
        // (a fully unrolled step is 53 KB of SASS at 6 bands, beyond the 32 KB instruction cache:
        // 6 % no-instruction stalls, 1.0998 -> 1.0891 s), when the model values stay in the
        // stage (code size and compile time: 4 x NF bands x groups otherwise) and in SAMPLE
        // mode (no branch separates the targets there: unrolled, the scheduler interleaves all
        // eight target visits of a step and spills 1.7 KB per thread)
#pragma unroll (SMODEL || NF >= GRP_ROLL_NF ? 1 : FZT / ZJ)
        for (int grp = 0; grp < FZT / ZJ; ++grp) {
            const int zl = grp * 32 * ZJ + lane;      // cells: zl + 32 c, c < ZJ
            constexpr int RN = SMODEL ? 1 : NF;
            double m[ZJ][RN], cv[ZJ][RN];
            if (!SMODEL) {
#pragma unroll
                for (int f = 0; f < RN; ++f) {
#pragma unroll
                    for (int c = 0; c < ZJ; ++c) {
                        m[c][f] = st[f * FZC + zl + 32 * c];
                        if (!HYBRID) cv[c][f] = st[(NF + f) * FZC + zl + 32 * c];
                    }
                }
            }
            double pr[ZJ], wzv[ZJ], xmask[ZJ], zot[ZJ], ztt[ZJ], zoo[ZJ], zev[ZJ];
            unsigned pad[ZJ];          // exp_nonpos_tab's limit: EXP_ZERO for padded redshifts
#pragma unroll
            for (int c = 0; c < ZJ; ++c) {
                pr[c] = st[2 * var0 + zl + 32 * c];
                wzv[c] = st[2 * var0 + FZC + zl + 32 * c];
#if DL_PRW
                wzv[c] *= pr[c];
#endif
                // padded redshifts get x = -inf: they never count for the max and contribute 0
                pad[c] = (DL_HCFOLD && (z0 + zl + 32 * c) >= p.nz) ? EXP_ZERO : EXP_UNDER;
                xmask[c] = ((z0 + zl + 32 * c) < p.nz) ? 0.0 : -INFINITY;
                zot[c] = ztt[c] = zoo[c] = 0.0; zev[c] = 1.0;
                if (ZDEP) {
                    const double* zd = st + 2 * var0 + 2 * FZC + zl + 32 * c;
                    zot[c] = zd[0]; ztt[c] = zd[FZC]; zoo[c] = zd[2 * FZC]; zev[c] = zd[3 * FZC];
                }
            }
#pragma unroll (SMODEL || SAMPLE ? 1 : TGT_UNROLL)
            for (int o = 0; o < FTO; o += TP) {
                // E[] is rotated once per target group so the loop can stay rolled with the
                // accumulators in registers (FTO / TP rotations restore the order)
                double Eo[TP];
                double x[TP][ZJ], y[TP][ZJ];
#pragma unroll
                for (int j = 0; j < TP; ++j) Eo[j] = E[j];
                // heads: x = -chi2/2 - A and y = det^-1/2, TJ targets at a time
#pragma unroll
                for (int g = 0; g < TP; g += TJ) {
                    const double* tcs[TJ];
                    double pOT0[TJ][ZJ], pOT[TJ][ZJ], pTT[TJ][ZJ], pOO[TJ][ZJ], xm[TJ][ZJ];
                    double xg[TJ][ZJ], yg[TJ][ZJ];
#pragma unroll
                    for (int j = 0; j < TJ; ++j) {
                        tcs[j] = tcw + (o + g + j) * TCS;
                        const double* tk = tcs[j] + 5 * NF;
#pragma unroll
                        for (int c = 0; c < ZJ; ++c) {
                            if (ZDEP) {
                                const double cc = tk[TC_C];
                                pOT0[j][c] = tk[TC_CS] * zot[c];
#if DL_HCFOLD
                                // (TC_SPARE = c h, c h_c = -1/2; the anchor rides on FOO'')
                                pOT[j][c] = tk[TC_SPARE] * zot[c]; pTT[j][c] = cc * ztt[c];
                                pOO[j][c] = SAMPLE ? -0.5 * zoo[c] : fma(-0.5, zoo[c], tk[TC_XOFF]);
                                xm[j][c] = xmask[c];
#else
                                pOT[j][c] = cc * zot[c]; pTT[j][c] = cc * ztt[c];
                                pOO[j][c] = cc * zoo[c];
                                xm[j][c] = SAMPLE ? xmask[c] : xmask[c] + tk[TC_XOFF];   // -A
#endif
                            } else {
                                // the anchor is folded into the constant term of FOO (TC_POO);
                                // the sampling pass measures the un-anchored exponent (TC_POO0)
                                pOT0[j][c] = tk[TC_POT0];
                                pOT[j][c] = tk[TC_POT]; pTT[j][c] = tk[TC_PTT];
                                pOO[j][c] = tk[SAMPLE ? TC_POO0 : TC_POO];
                                xm[j][c] = xmask[c];
                            }
                        }
                    }
                    if constexpr (HET)
                        cells_xy<NF, TJ, SAMPLE || !DL_HCFOLD>(tcs, SmemModelHet<NF, FZC>{st + zl,
                                             reinterpret_cast<const int*>(tcs[0] + 5 * NF + TC_NCONST),
                                             var0},
                                         pOT0, pOT, pTT, pOO, zev, xm, xg, yg);
                    else if constexpr (SMODEL)
                        cells_xy<NF, TJ, SAMPLE || !DL_HCFOLD>(tcs, SmemModel<NF, FZC>{st + zl}, pOT0, pOT, pTT, pOO,
                                         zev, xm, xg, yg);
                    else if constexpr (HYBRID)
                        cells_xy<NF, TJ, SAMPLE || !DL_HCFOLD>(tcs, HybridModel<NF, FZC>{m, st + zl}, pOT0, pOT, pTT, pOO,
                                         zev, xm, xg, yg);
                    else
                        cells_xy<NF, TJ, SAMPLE || !DL_HCFOLD>(tcs, RegModel<NF>{m, cv}, pOT0, pOT, pTT, pOO, zev, xm,
                                         xg, yg);
#pragma unroll
                    for (int j = 0; j < TJ; ++j)
#pragma unroll
                        for (int c = 0; c < ZJ; ++c) { x[g + j][c] = xg[j][c]; y[g + j][c] = yg[j][c]; }
                }
                // tails of the TP targets together: the exp chains are long and dependent, so
                // they are interleaved (TP x ZJ independent chains) before the first branch
                if constexpr (SAMPLE) {
#pragma unroll
                    for (int j = 0; j < TP; ++j) {
                        double xl = x[j][0];
#pragma unroll
                        for (int c = 1; c < ZJ; ++c) xl = fmax(xl, x[j][c]);
                        Eo[j] = fmax(Eo[j], xl);
                    }
                } else {
                    double v[TP][ZJ];
                    int vh[TP];
#pragma unroll
                    for (int j = 0; j < TP; ++j) {
                        vh[j] = 0;
#pragma unroll
                        for (int c = 0; c < ZJ; ++c) {
                            v[j][c] = exp_nonpos_tab(x[j][c], etab, pad[c]) * y[j][c];   // :118,128
                            ovf = max(ovf, __double2hiint(x[j][c]));
                            vh[j] = max(vh[j], __double2hiint(v[j][c]));
                        }
                    }
#pragma unroll
                    for (int j = 0; j < TP; ++j) {
                        // running max of v (the final divisor), kept per target in shared
                        // memory: only the high word is read here; the rare record-breaking
                        // cells take the slow path (O(log) times per lane)
                        unsigned long long* Vt = V_sm + tfirst + o + j;
                        if (vh[j] >= reinterpret_cast<const int*>(Vt)[1]) {
#pragma unroll
                            for (int c = 0; c < ZJ; ++c)
                                if (v[j][c] == v[j][c])
                                    atomicMax(Vt, (unsigned long long)__double_as_longlong(v[j][c]));
                        }
                    }
                    // prior, PDF and evidence accumulation  (delight-apply.py:145-148)
#pragma unroll
                    for (int j = 0; j < TP; ++j)
#pragma unroll
                        for (int c = 0; c < ZJ; ++c) {
#if DL_PRW
                            // (wzv holds prior x trapezoid weight: one multiplication less)
                            double* pd = pdfw + (o + j) * FZC + zl + 32 * c;
                            *pd = fma(v[j][c], pr[c], *pd);
                            Eo[j] = fma(v[j][c], wzv[c], Eo[j]);
#else
                            const double vp = v[j][c] * pr[c];
                            pdfw[(o + j) * FZC + zl + 32 * c] += vp;
                            Eo[j] = fma(vp, wzv[c], Eo[j]);
#endif
                        }
                }
#pragma unroll
                for (int q = 0; q + TP < FTO; ++q) E[q] = E[q + TP];
#pragma unroll
                for (int j = 0; j < TP; ++j) E[FTO - TP + j] = Eo[j];
            }
        }
        if constexpr (!SAMPLE) {
            // evidence of (target, training galaxy k) over this z-chunk: four warp sums folded
            // into one butterfly -- after the xor-16 step a lane keeps two of the four partial
            // sums, after xor-8 one; lane 8 j ends up with the sum of target j
            static_assert(FTO == 4, "butterfly written for 4 targets per warp");
            const bool hi16 = lane & 16, hi8 = lane & 8;
            const double s0 = hi16 ? E[0] : E[2], s1 = hi16 ? E[1] : E[3];      // what I give away
            double k0 = hi16 ? E[2] : E[0], k1 = hi16 ? E[3] : E[1];            // what I keep
            k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            double ev = (hi8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
            ev += __shfl_xor_sync(0xffffffffu, ev, 4);
            ev += __shfl_xor_sync(0xffffffffu, ev, 2);
            ev += __shfl_xor_sync(0xffffffffu, ev, 1);
            if ((lane & 7) == 0) E_sm[(size_t)(tfirst + (lane >> 3)) * TTs + k] += ev;
#pragma unroll
            for (int o = 0; o < FTO; ++o) E[o] = 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[cs]);
        if (++cs == STG) { cs = 0; cph ^= 1; }
        if (++k == tcount) {
            if constexpr (!SAMPLE) {
                // end of a z-chunk: add the warp's PDF slices to the global accumulators
#pragma unroll
                for (int o = 0; o < FTO; ++o) {
                    const int og = og0 + o;
                    if (og < p.nO) {
                        double* dst = p.pdf_acc + (size_t)og * p.nzp + z0;
#pragma unroll
                        for (int zj = 0; zj < FZT; ++zj) {
                            const double a = pdfw[o * FZC + zj * 32 + lane];
                            if (a != 0.0) atomicAdd(dst + zj * 32 + lane, a);
                            pdfw[o * FZC + zj * 32 + lane] = 0.0;
                        }
                    }
                }
            }
            k = 0; ++chunk;
        }
    }

    // ---------------------------------------------------------------------- unit epilogue
    __syncwarp();
    if constexpr (SAMPLE) {
#pragma unroll
        for (int o = 0; o < FTO; ++o) {
            const double mx = warp_max(E[o]);
            const int og = og0 + o;
            if (lane == 0 && og < p.nO && mx == mx) {
                unsigned long long b = (unsigned long long)__double_as_longlong(mx);
                b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
                atomicMax(p.Akey + og, b);
            }
        }
    } else {
        const int wovf = __reduce_max_sync(0xffffffffu, ovf);
        const bool over = wovf >= HI_OVERFLOW;
#pragma unroll
        for (int o = 0; o < FTO; ++o) {
            const int og = og0 + o;
            if (og >= p.nO) continue;
            if (p.rescue && p.flag[og] == 0) continue;         // warp-uniform
            if (p.evid != nullptr) {
                double* dst = p.evid + (size_t)og * p.nT + ts0;
                // streaming store: written once, read by the top-K kernel much later -- keep
                // these lines from displacing the PDF accumulators and the model grid in L2
                for (int q = lane; q < tcount; q += 32)
                    __stcs(dst + q, E_sm[(size_t)(tfirst + o) * TTs + q]);
            }
            if (lane == 0) {
                atomicMax(p.Vbits + og, V_sm[tfirst + o]);
                if (over && !p.rescue) { p.flag[og] = 1; *p.nflag = 1; }
            }
        }
    }
}

template <int NF, bool ZDEP, bool HET, int MODE>
static int launch_fused_t(const FusedParams& p, cudaStream_t stream, bool timed) {
    constexpr int CFG = cfg_of(NF, HET);
    const int nU = HET ? p.nU : NF;
    const size_t smem = fused_smem_bytes(CFG, NF, HET, ZDEP, nU, p.stages, p.TT);
    auto kern = fused_like_kernel<NF, ZDEP, HET, MODE>;
    DL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    dim3 grid((unsigned)(p.nTiles * p.nTB));
    if (timed && dl_timing_enabled()) dl_timing_mark(stream, 0);
    kern<<<grid, cfg_warps(CFG) * 32, smem, stream>>>(p);
    if (timed && dl_timing_enabled()) dl_timing_mark(stream, 1);
    DL_CUDA_CHECK(cudaGetLastError());
    dl_count_launch();
    return DL_OK;
}

template <int NF>
static int launch_fused_nf(bool zdep, bool het, int mode, const FusedParams& p,
                           cudaStream_t stream, bool timed) {
    if (het) {
        if (zdep) {
            dl_set_error("per-target band lists with z-dependent ell priors: group by band set");
            return DL_ERR_UNSUPPORTED;
        }
        return mode == MODE_SAMPLE ? launch_fused_t<NF, false, true, MODE_SAMPLE>(p, stream, timed)
                                   : launch_fused_t<NF, false, true, MODE_MAIN>(p, stream, timed);
    }
    if (zdep)
        return mode == MODE_SAMPLE ? launch_fused_t<NF, true, false, MODE_SAMPLE>(p, stream, timed)
                                   : launch_fused_t<NF, true, false, MODE_MAIN>(p, stream, timed);
    return mode == MODE_SAMPLE ? launch_fused_t<NF, false, false, MODE_SAMPLE>(p, stream, timed)
                               : launch_fused_t<NF, false, false, MODE_MAIN>(p, stream, timed);
}

}  // namespace dl

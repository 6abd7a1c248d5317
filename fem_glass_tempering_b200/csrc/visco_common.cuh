// Constants and per-term factors of the viscoelastic chain shared by visco.cu (hot path A) and mech.cu (the
// equilibrium extension): both must evaluate lambda*(1 - taylor)/xi with the same operation sequence.
// Translation units including this are compiled with -fmad=false.
#pragma once

#include <cuda.h>

#include "sg_common.cuh"

struct VKParams {
    int N;
    double c_HRg;    // H / Rg            (VM:158)
    double inv_Tb;   // 1 / Tb
    double dt;
    double half_dt;  // dt / 2            (VM:171)
    double inv_d;    // 1 / dim           (VM:144)
    double alpha_s;
    double d_alpha;  // alpha_liquid - alpha_solid (VM:130)
    double m[SG_MAX_TERMS], lm[SG_MAX_TERMS];
    double g2[SG_MAX_TERMS], lg[SG_MAX_TERMS];  // g2 = 2.0 * g_n (VM:178)
    double k[SG_MAX_TERMS], lk[SG_MAX_TERMS];
    int mode;        // SG_VISCO_REFERENCE: the reference's expressions as executed; SG_VISCO_CORRECTED: see below
    double chi;      // VM:15
};

// VM:233-242   (1.0 + a) + 0.5*a^2,  a = (-xi)/lambda
__device__ __forceinline__ double taylor3(double xi, double lambda) {
    const double a = (-1.0 * xi) / lambda;
    return (1.0 + a) + 0.5 * (a * a);
}

__device__ __forceinline__ void decay_fac(double xi, double lambda, double &decay, double &fac) {
    const double x = xi / lambda;
    const double em1 = expm1(-x);
    decay = 1.0 + em1;
    fac = (x != 0.0) ? (-em1) / x : 1.0;
}

// Tensor maps (2-D TMA) of the two history arrays viewed as [n_nodes rows][N*d*d doubles]: the chunked fast kernel
// fetches a [32 nodes][chunk of terms] box with ONE cp.async.bulk.tensor per array (SASS UTMALDG / UTMASTG).
struct ViscoTmaps {
    CUtensorMap s, k;
};
typedef void (*visco_fast_fn)(const VKParams, const sg_visco_fields, const long, const ViscoTmaps);

struct sg_visco_plan {
    sg_ctx *ctx;
    sg_visco_params p;
    VKParams k;
    visco_fast_fn fast;   // nullptr when (dim, n_terms) has no compiled fast path
    uint32_t fast_smem;
    int fast_grid;        // resident one-warp CTAs on the whole GPU
    int fast_chunk;       // terms per staged chunk (< n_terms: the kernel needs tensor maps of the history arrays)
    ViscoTmaps tmaps;     // for the arrays below (rebuilt when the caller passes other buffers)
    const void *tm_s, *tm_k;
    int64_t tm_rows;
};


// Hot path (A): fused Tool–Narayanaswamy / Prony viscoelastic update.
//
// One launch replaces the 17 Function.interpolate(Expression) passes and 7 array
// copies that ThermoViscoProblem.solve_timestep makes per step (TVP:370-373, the
// expressions are VM:111-228, the Taylor stand-in for exp is VM:233-242).
//
// COMPILED WITH -fmad=false: the reference's pointwise kernels are FMA-free (FFCx
// default cffi flags, no -march=native), and lambda*(1 - taylor)/xi cancels
// catastrophically, so the IEEE operation sequence is mirrored exactly.  Every
// product/sum below is written in the association order of the Python source.
//
// Data movement per CTA tile of VTILE nodes:
//   * the two fat history tensors s_tilde / sigma_tilde ([node, N, d, d], 432 B per
//     node each at d=3, N=6) are contiguous per tile, so they are staged into shared
//     memory with one 1-D TMA bulk copy each (cp.async.bulk -> UBLKCP) signalled on
//     an mbarrier, updated in place in shared memory and written back with a bulk
//     store.  No register staging, full-line HBM transactions in both directions.
//   * while the bulk loads are in flight the CTA evaluates the per-node scalars
//     (2 exp, the N fictive-temperature relaxations, the 2N Taylor factors and 3N
//     division chains) spread over (node, term) pairs.
//   * thin arrays (T, phi, xi, Tf, Tf_partial, sigma) are read/written with plain
//     coalesced accesses: the flat index of the tile is the thread index.
#include "sg_common.cuh"
#include "visco_common.cuh"

namespace {

constexpr int VTILE = 32;      // nodes per CTA
constexpr int VTHREADS = 256;  // threads per CTA

struct VGather {
    int n_ld;
    const int32_t *dofs;
    const uint8_t *local_point;
    const double *weights;
};

// VM:156-161 / VM:162-167
__device__ __forceinline__ double shift_phi(const VKParams &P, double T) {
    return exp(P.c_HRg * (P.inv_Tb - 1.0 / T));
}

// ---- SG_VISCO_CORRECTED: the scheme the reference's comments cite (Nielsen et al. 2010) without the quirks of the
// executed code:
//   phi      = exp(H/Rg (1/Tb - chi/T_cur  - (1-chi)/Tf_prev))          Eq. 25 as written at VM:100-108 (dead there)
//   Tf_partial, Tf                                                      VM:111-125 unchanged, with this phi
//   d eps_th = alpha_s (T_cur - T_prev) + (alpha_l - alpha_s)(Tf_cur - Tf_prev)     VM:128-133 with the OLD Tf_prev
//   xi       = dt/2 (phi(T_prev, Tf_prev) + phi(T_cur, Tf_cur))         trapezoidal shifted-time increment (> 0)
//   per term x = xi/lambda:  decay = exp(-x),  fac = (1 - exp(-x))/x = -expm1(-x)/x   (no cancellation, no 0/0)
//   s_n      <- s_n decay + 2 g_n dev fac,   sigma_n <- sigma_n decay + k_n tr I fac       the history IS the partial stress
__device__ __forceinline__ double phi_eq25(const VKParams &P, double T, double Tf) {
    return exp(P.c_HRg * (P.inv_Tb - P.chi / T - (1.0 - P.chi) / Tf));
}
__device__ __forceinline__ double gather_eval(const VGather &G, const double *__restrict__ arr, long node) {
    const double *w = G.weights + (int)G.local_point[node] * G.n_ld;
    const int32_t *dj = G.dofs + node * G.n_ld;
    double acc = 0.0;
    for (int j = 0; j < G.n_ld; ++j) {
        const double wj = w[j];
        if (wj != 0.0) acc = acc + wj * arr[dj[j]];
    }
    return acc;
}

__host__ __device__ inline size_t visco_smem_bytes(int N, int DD, bool tensor) {
    size_t doubles = 6 * (size_t)VTILE * N  // tfp, tg, tk, ds_d, ds_o, dk_d
                     + 5 * (size_t)VTILE;   // dT, xi, tf, Tc, phi
    if (tensor) doubles += 2 * (size_t)VTILE * N * DD;
    return 128 + doubles * sizeof(double);
}

template <int D, bool SCALAR, bool TENSOR>
__global__ void __launch_bounds__(VTHREADS)
visco_kernel(const VKParams P, const sg_visco_fields f, const VGather G, const long n_nodes,
             const unsigned phases, const int bulk_ok) {
    constexpr int DD = D * D;
    constexpr bool GATHER = TENSOR && !SCALAR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *sm = reinterpret_cast<double *>(smem_raw + 128);

    const int N = P.N;
    const int tid = threadIdx.x;
    const long tile0 = (long)blockIdx.x * VTILE;
    const int nn = (int)min((long)VTILE, n_nodes - tile0);

    // shared-memory carve-up
    double *buf_s = sm;                                   // [nn, N, DD]   s_tilde tile
    double *buf_k = buf_s + (TENSOR ? VTILE * N * DD : 0);  // [nn, N, DD]   sigma_tilde tile
    double *s_tfp = buf_k + (TENSOR ? VTILE * N * DD : 0);  // [nn*N]
    double *s_tg = s_tfp + VTILE * N;
    double *s_tk = s_tg + VTILE * N;
    double *s_dsd = s_tk + VTILE * N;
    double *s_dso = s_dsd + VTILE * N;
    double *s_dkd = s_dso + VTILE * N;
    double *s_dT = s_dkd + VTILE * N;  // [nn]
    double *s_xi = s_dT + VTILE;
    double *s_tf = s_xi + VTILE;
    double *s_Tc = s_tf + VTILE;
    double *s_phi = s_Tc + VTILE;

    const bool ph_tf = SCALAR && (phases & SG_PHASE_TF);
    const bool ph_shift = SCALAR && (phases & SG_PHASE_SHIFT);
    const bool ph_strain = TENSOR && (phases & SG_PHASE_STRAIN);
    const bool ph_stress = TENSOR && (phases & SG_PHASE_STRESS);

    const int tile_elems = nn * N * DD;
    const uint32_t tile_bytes = (uint32_t)tile_elems * 8u;
    const bool use_bulk = ph_stress && bulk_ok && (tile_bytes % 16u == 0);
    const long gbase = tile0 * N * DD;

    // ---- stage 0: start the TMA bulk loads of the two history tiles -------------------
    if (ph_stress) {
        if (use_bulk) {
            if (tid == 0) {
                sgptx::mbar_init(bar, 1);
                sgptx::fence_mbar_init();
            }
            __syncthreads();
            if (tid == 0) {
                sgptx::mbar_expect_tx(bar, 2u * tile_bytes);
                sgptx::bulk_g2s(buf_s, f.s_tilde + gbase, tile_bytes, bar);
                sgptx::bulk_g2s(buf_k, f.sigma_tilde + gbase, tile_bytes, bar);
            }
        } else {
            for (int e = tid; e < tile_elems; e += VTHREADS) {
                buf_s[e] = f.s_tilde[gbase + e];
                buf_k[e] = f.sigma_tilde[gbase + e];
            }
        }
    }

    if (P.mode == SG_VISCO_CORRECTED) {
        // Only the fused same-space call reaches here (sg_visco_update checks): all four phases at once.
        if constexpr (SCALAR && TENSOR) {
            // node scalars that do not need the new Tf
            if (tid < nn) {
                const long node = tile0 + tid;
                const double Tc = f.T_cur[node], Tp = f.T_prev[node], Tfo = f.Tf[node];
                const double phi = phi_eq25(P, Tc, Tfo);
                f.phi[node] = phi;
                if (f.T_next) f.T_next[node] = Tc + (Tc - Tp);
                s_Tc[tid] = Tc;
                s_dT[tid] = Tc - Tp;
                s_phi[tid] = phi;
                s_tf[tid] = Tfo;                                 // old Tf
                s_xi[tid] = phi_eq25(P, Tp, Tfo);                // phi at t_n, completed to xi below
            }
            __syncthreads();
            const int n_pairs = nn * N;
            for (int it = tid; it < n_pairs; it += VTHREADS) {
                const int t = it / N, i = it - t * N;
                const double Tc = s_Tc[t], phi = s_phi[t];
                const double v = (P.lm[i] * f.Tf_partial[tile0 * N + it] + (Tc * P.dt) * phi) / (P.lm[i] + P.dt * phi);
                f.Tf_partial[tile0 * N + it] = v;
                s_tfp[it] = v;
            }
            __syncthreads();
            if (tid < nn) {
                const long node = tile0 + tid;
                double tf = P.m[0] * s_tfp[tid * N];
                for (int j = 1; j < N; ++j) tf = tf + P.m[j] * s_tfp[tid * N + j];
                f.Tf[node] = tf;
                const double Tfo = s_tf[tid];
                const double phin = phi_eq25(P, s_Tc[tid], tf);
                const double xi = P.half_dt * (s_xi[tid] + phin);
                f.xi[node] = xi;
                if (f.phi_next) f.phi_next[node] = phin;
                s_xi[tid] = xi;
                s_tf[tid] = tf - Tfo;                            // structural part of the strain increment
            }
            __syncthreads();
            for (int it = tid; it < n_pairs; it += VTHREADS) {
                const int t = it / N, i = it - t * N;
                const long node = tile0 + t;
                const double eth = P.alpha_s * s_dT[t] + P.d_alpha * s_tf[t];
                const double tot_d = -1.0 * eth, tot_o = -1.0 * 0.0;
                double tr = tot_d;
#pragma unroll
                for (int a = 1; a < D; ++a) tr = tr + tot_d;
                const double dev_d = tot_d - P.inv_d * tr, dev_o = tot_o;
                if (i == 0) {
#pragma unroll
                    for (int c = 0; c < DD; ++c) {
                        const bool diag = (c % (D + 1)) == 0;
                        if (f.thermal_strain) f.thermal_strain[node * DD + c] = diag ? eth : 0.0;
                        if (f.total_strain) f.total_strain[node * DD + c] = diag ? tot_d : tot_o;
                        if (f.deviatoric_strain) f.deviatoric_strain[node * DD + c] = diag ? dev_d : dev_o;
                    }
                }
                double dg, fg, dk, fk;
                decay_fac(s_xi[t], P.lg[i], dg, fg);
                decay_fac(s_xi[t], P.lk[i], dk, fk);
                s_tg[it] = dg;
                s_tk[it] = dk;
                s_dsd[it] = (P.g2[i] * dev_d) * fg;
                s_dso[it] = (P.g2[i] * dev_o) * fg;
                s_dkd[it] = (P.k[i] * tr) * fk;
            }
        }
    } else {
    // ---- stage 1a: per-node scalars (thread per node) ---------------------------------
        if (tid < nn) {
            const long node = tile0 + tid;
            double Tc, Tp, xi, phi = 0.0;
            if constexpr (GATHER) {
                Tc = gather_eval(G, f.T_cur, node);
                Tp = gather_eval(G, f.T_prev, node);
                xi = gather_eval(G, f.xi, node);
                s_tf[tid] = gather_eval(G, f.Tf, node);
            } else {
                Tc = f.T_cur[node];
                Tp = f.T_prev[node];
                phi = shift_phi(P, Tc);                     // VM:156
                const double Tn = Tc + (Tc - Tp);           // VM:151
                const double phin = shift_phi(P, Tn);       // VM:162
                xi = P.half_dt * (phin - phi);              // VM:171
                if (ph_tf || ph_shift) f.phi[node] = phi;   // TVP:456, TVP:531
                if (ph_shift) {
                    f.xi[node] = xi;                        // TVP:541
                    if (f.T_next) f.T_next[node] = Tn;      // TVP:524
                    if (f.phi_next) f.phi_next[node] = phin;  // TVP:533
                }
                if (!ph_tf) s_tf[tid] = f.Tf[node];
            }
            s_Tc[tid] = Tc;
            s_dT[tid] = Tc - Tp;
            s_xi[tid] = xi;
            s_phi[tid] = phi;
        }
        __syncthreads();
    
        // ---- stage 1b: (node, term) pairs: fictive-temperature relaxation + Taylor factors -
        const int n_pairs = nn * N;
        for (int it = tid; it < n_pairs; it += VTHREADS) {
            const int t = it / N, i = it - t * N;
            if (ph_tf) {
                // VM:111-119  (lambda_m*Tfp_prev + T*dt*phi) / (lambda_m + dt*phi)
                const double Tc = s_Tc[t], phi = s_phi[t];
                const double num = P.lm[i] * f.Tf_partial[tile0 * N + it] + (Tc * P.dt) * phi;
                const double den = P.lm[i] + P.dt * phi;
                const double v = num / den;
                f.Tf_partial[tile0 * N + it] = v;  // TVP:466 + copy TVP:469
                s_tfp[it] = v;
            }
            if (ph_stress) {
                const double xi = s_xi[t];
                s_tg[it] = taylor3(xi, P.lg[i]);
                s_tk[it] = taylor3(xi, P.lk[i]);
            }
        }
        __syncthreads();
    
        // ---- stage 1c: Tf, strains, Prony increment coefficients ---------------------------
        for (int it = tid; it < n_pairs; it += VTHREADS) {
            const int t = it / N, i = it - t * N;
            const long node = tile0 + t;
            double tf;
            if (ph_tf) {
                // VM:122-125  inner(m, Tf_partial), left-to-right
                tf = P.m[0] * s_tfp[t * N];
                for (int j = 1; j < N; ++j) tf = tf + P.m[j] * s_tfp[t * N + j];
                if (i == 0) f.Tf[node] = tf;  // TVP:480 + copy TVP:481
            } else {
                tf = s_tf[t];
            }
            if constexpr (TENSOR) {
                // VM:128-133; Tf_cur == Tf_prev bitwise after TVP:481, so the structural term is 0 (or NaN)
                const double eth = P.alpha_s * s_dT[t] + P.d_alpha * (tf - tf);
                const double tot_d = -1.0 * eth;   // VM:137
                const double tot_o = -1.0 * 0.0;
                double tr = tot_d;
    #pragma unroll
                for (int a = 1; a < D; ++a) tr = tr + tot_d;
                const double dev_d = tot_d - P.inv_d * tr;  // VM:144
                const double dev_o = tot_o;
                if (ph_strain && i == 0) {
    #pragma unroll
                    for (int c = 0; c < DD; ++c) {
                        const bool diag = (c % (D + 1)) == 0;
                        if (f.thermal_strain) f.thermal_strain[node * DD + c] = diag ? eth : 0.0;  // TVP:492
                        if (f.total_strain) f.total_strain[node * DD + c] = diag ? tot_d : tot_o;  // TVP:504
                        if (f.deviatoric_strain) f.deviatoric_strain[node * DD + c] = diag ? dev_d : dev_o;  // TVP:516
                    }
                }
                if (ph_stress) {
                    const double xi = s_xi[t];
                    const double one_g = 1.0 - s_tg[it], one_k = 1.0 - s_tk[it];
                    // VM:176-182   2.0*g_n*dev/xi*lambda_g_n*(1.0 - taylor)
                    s_dsd[it] = ((P.g2[i] * dev_d) / xi) * P.lg[i] * one_g;
                    s_dso[it] = ((P.g2[i] * dev_o) / xi) * P.lg[i] * one_g;
                    // VM:185-191   k_n*(tr*I)/xi*lambda_k_n*(1.0 - taylor)
                    s_dkd[it] = ((P.k[i] * tr) / xi) * P.lk[i] * one_k;
                }
            }
        }
}
    if constexpr (TENSOR) {
        if (!ph_stress) return;
        if (use_bulk) sgptx::mbar_wait(bar, 0);
        __syncthreads();

        // ---- stage 2: (node, component) pairs sweep the N terms in shared memory -------
        const int n_items = nn * DD;
        for (int it = tid; it < n_items; it += VTHREADS) {
            const int t = it / DD, c = it - t * DD;
            const bool diag = (c % (D + 1)) == 0;
            double acc = 0.0;
            for (int n = 0; n < N; ++n) {
                const int e = (t * N + n) * DD + c;
                const int q = t * N + n;
                const double st = buf_s[e] * s_tg[q];    // VM:194-200
                const double sg = buf_k[e] * s_tk[q];    // VM:203-209
                const double ds = diag ? s_dsd[q] : s_dso[q];
                const double dk = diag ? s_dkd[q] : 0.0;
                const double sp = ds + st;               // VM:212-215
                const double kp = dk + sg;               // VM:218-221
                const bool corrected = P.mode == SG_VISCO_CORRECTED;
                buf_s[e] = corrected ? sp : st;          // TVP:552 + copy TVP:559 (corrected: the history is the partial stress)
                buf_k[e] = corrected ? kp : sg;          // TVP:571 + copy TVP:578
                const double pn = sp + kp;               // VM:224-228
                acc = (n == 0) ? pn : acc + pn;
                if (f.ds_partial) f.ds_partial[gbase + e] = ds;          // TVP:549
                if (f.dsigma_partial) f.dsigma_partial[gbase + e] = dk;  // TVP:568
                if (f.s_partial) f.s_partial[gbase + e] = sp;            // TVP:555 (+561)
                if (f.sigma_partial) f.sigma_partial[gbase + e] = kp;    // TVP:574 (+582)
            }
            f.sigma[tile0 * DD + it] = acc;  // TVP:591
        }

        // ---- stage 3: write the updated history tiles back --------------------------------
        if (use_bulk) {
            sgptx::fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                sgptx::bulk_s2g(f.s_tilde + gbase, buf_s, tile_bytes);
                sgptx::bulk_s2g(f.sigma_tilde + gbase, buf_k, tile_bytes);
                sgptx::bulk_commit();
                sgptx::bulk_wait_read0();
            }
        } else {
            __syncthreads();
            for (int e = tid; e < tile_elems; e += VTHREADS) {
                f.s_tilde[gbase + e] = buf_s[e];
                f.sigma_tilde[gbase + e] = buf_k[e];
            }
        }
    }
}


// =====================================================================================
// Fast path: persistent, barrier-free kernel for the steady-state call
// (phases == ALL, no optional outputs, same-space, N known at compile time).
//
// One WARP owns a tile of 32 nodes; lane == node.  The warp's s_tilde / sigma_tilde /
// Tf_partial rows are contiguous in HBM, so lane 0 fetches them with three 1-D TMA bulk
// copies onto a warp-private mbarrier while all lanes evaluate the exp()-dependent
// scalars.  Each lane then runs the whole chain for its node out of registers, sweeping
// its own row in shared memory with 128-bit accesses (row stride N*d*d*8 B is
// bank-conflict-free for LDS.128), and the warp writes the four result tiles back with
// bulk stores.  No __syncthreads, no cross-lane traffic, CTAs of a single warp so the
// scheduler packs as many tiles per SM as shared memory allows.
// Arithmetic (order of every operation) is identical to visco_kernel above.
// =====================================================================================
constexpr int WT = 32;  // nodes per warp tile

#ifndef SG_VISCO_CHUNK
#define SG_VISCO_CHUNK 6
#endif
// Terms per shared-memory chunk.  While a node's two history rows fit in 32 KB per warp tile (N d^2 <= 64 doubles) they are
// staged whole (three 1-D bulk copies per tile); beyond that — 8, 10, 12 Prony terms in 3-D, BASELINE config 5 — a row of
// 12 x 9 doubles would cost 60 KB of shared memory per warp and leave 3 warps per SM, so the terms are staged
// SG_VISCO_CHUNK = 6 at a time as a [32 nodes][6 terms] BOX of a 2-D tensor map of the history array (one
// cp.async.bulk.tensor per array and chunk, SASS UTMALDG/UTMASTG; a box that ends past the last term is zero-filled on
// load and clipped on store).  Measured on B200, d = 3, fraction of the measured HBM peak for N = 8 / 10 / 12:
//   whole rows 0.81 / 0.78 / 0.58;  per-lane 1-D copies of 4-term segments (round 1) 0.76 / 0.76 / 0.81;
//   tensor boxes of 2 terms 0.71 / 0.65 / 0.66, of 4 terms 0.82 / 0.77 / 0.76, of 6 terms **0.90 / 0.87 / 0.85**:
// every chunk is a serial load -> compute -> store round trip of the warp, so few large chunks win as long as seven warps
// still fit on an SM (the 6-term footprint).
__host__ __device__ constexpr int fast_chunk_terms(int D, int N) {
    const int DD = D * D;
    if (N * DD <= 64) return N;
    // a chunk of a row must be a multiple of 16 bytes (TMA box width): an even number of doubles
    return ((SG_VISCO_CHUNK * DD) % 2 == 0 && SG_VISCO_CHUNK < N) ? SG_VISCO_CHUNK : N;
}

template <int D, int N>
struct FastCfg {
    static constexpr int DD = D * D;
    static constexpr int ROW = N * DD;             // doubles per node in a history tensor
    static constexpr int CH = fast_chunk_terms(D, N);   // terms staged at a time
    static constexpr int CROW = CH * DD;           // doubles per node in a staged chunk
    static constexpr bool VEC = (ROW % 2) == 0 && (CROW % 2) == 0;    // rows 16-B aligned -> LDS.128 / STS.128
    static constexpr int G = VEC ? ((DD % 2) == 0 ? 1 : 2) : 1;  // terms per register group
    static constexpr uint32_t S_BYTES = WT * CROW * 8;
    static constexpr uint32_t TFP_BYTES = WT * N * 8;
    static constexpr uint32_t SIG_BYTES = WT * DD * 8;
    static constexpr uint32_t SMEM = 2 * S_BYTES + TFP_BYTES + SIG_BYTES + 16;
};

template <int D, int N, bool CORR = false>
__global__ void __launch_bounds__(32) visco_fast_kernel(const VKParams P, const sg_visco_fields f, const long n_tiles,
                                                        const __grid_constant__ ViscoTmaps tm) {
    using C = FastCfg<D, N>;
    constexpr int DD = C::DD, ROW = C::ROW, G = C::G, CH = C::CH, CROW = C::CROW;
    constexpr bool CHUNKED = CH < N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *buf_s = reinterpret_cast<double *>(smem_raw);
    double *buf_k = buf_s + WT * CROW;
    double *buf_t = buf_k + WT * CROW;
    double *buf_o = buf_t + WT * N;
    uint64_t *bar = reinterpret_cast<uint64_t *>(buf_o + WT * DD);
    const int lane = threadIdx.x;

    if (lane == 0) {
        sgptx::mbar_init(bar, 1);
        sgptx::fence_mbar_init();
    }
    __syncwarp();
    uint32_t parity = 0;

    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long node0 = tile * WT;
        const long node = node0 + lane;
        if constexpr (!CHUNKED) {
            if (lane == 0) {
                sgptx::mbar_expect_tx(bar, 2u * C::S_BYTES + C::TFP_BYTES);
                sgptx::bulk_g2s(buf_s, f.s_tilde + node0 * ROW, C::S_BYTES, bar);
                sgptx::bulk_g2s(buf_k, f.sigma_tilde + node0 * ROW, C::S_BYTES, bar);
                sgptx::bulk_g2s(buf_t, f.Tf_partial + node0 * N, C::TFP_BYTES, bar);
            }
        } else {   // first chunk of terms: one [32 nodes][CH terms] box per history array
            if (lane == 0) {
                sgptx::mbar_expect_tx(bar, 2u * C::S_BYTES + C::TFP_BYTES);
                sgptx::bulk_g2s(buf_t, f.Tf_partial + node0 * N, C::TFP_BYTES, bar);
                sgptx::tensor_g2s_2d(buf_s, &tm.s, 0, (int)node0, bar);
                sgptx::tensor_g2s_2d(buf_k, &tm.k, 0, (int)node0, bar);
            }
        }
        const double Tc = f.T_cur[node], Tp = f.T_prev[node];
        double phi, xi, Tf_old = 0.0, phi_old = 0.0;
        if constexpr (CORR) {
            Tf_old = f.Tf[node];
            phi = phi_eq25(P, Tc, Tf_old);                 // VM:100-108
            phi_old = phi_eq25(P, Tp, Tf_old);
            xi = 0.0;                                      // needs the new Tf, see below
        } else {
            phi = shift_phi(P, Tc);                        // VM:156
            const double Tn = Tc + (Tc - Tp);              // VM:151
            const double phin = shift_phi(P, Tn);          // VM:162
            xi = P.half_dt * (phin - phi);                 // VM:171
            f.xi[node] = xi;
        }
        f.phi[node] = phi;
        const double Tdt_phi = (Tc * P.dt) * phi, dt_phi = P.dt * phi;

        sgptx::mbar_wait(bar, parity);
        parity ^= 1u;

        // ---- fictive temperatures (VM:111-125) ----
        double tf = 0.0;
        {
            double *row = buf_t + lane * N;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double v = (P.lm[i] * row[i] + Tdt_phi) / (P.lm[i] + dt_phi);
                row[i] = v;
                tf = (i == 0) ? P.m[0] * v : tf + P.m[i] * v;
            }
        }
        f.Tf[node] = tf;
        if constexpr (CORR) {
            xi = P.half_dt * (phi_old + phi_eq25(P, Tc, tf));
            f.xi[node] = xi;
        }
        // ---- strains (VM:128-146) ----
        const double eth = P.alpha_s * (Tc - Tp) + P.d_alpha * (CORR ? tf - Tf_old : tf - tf);
        const double tot_d = -1.0 * eth, tot_o = -1.0 * 0.0;
        double tr = tot_d;
#pragma unroll
        for (int a = 1; a < D; ++a) tr = tr + tot_d;
        const double dev_d = tot_d - P.inv_d * tr, dev_o = tot_o;

        // ---- Prony recursions (VM:176-228), G terms at a time out of registers, CH terms per staged chunk ----
        double acc[DD];
        double *rs = buf_s + lane * CROW, *rk = buf_k + lane * CROW;
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += CH) {
            constexpr int dummy = 0;
            (void)dummy;
            const int len = (N - c0 < CH) ? N - c0 : CH;      // compile-time after unrolling
            if (CHUNKED && c0 > 0) {                          // stage the next chunk (the previous one has been written back)
                if (lane == 0) {                              // a box that ends past the last term is zero-filled there; the
                    sgptx::mbar_expect_tx(bar, 2u * C::S_BYTES);   // transaction still counts the whole box
                    sgptx::tensor_g2s_2d(buf_s, &tm.s, c0 * DD, (int)node0, bar);
                    sgptx::tensor_g2s_2d(buf_k, &tm.k, c0 * DD, (int)node0, bar);
                }
                sgptx::mbar_wait(bar, parity);
                parity ^= 1u;
            }
#pragma unroll
            for (int n0 = c0; n0 < c0 + len; n0 += G) {
                double tg[G], tk[G], dsd[G], dso[G], dkd[G];
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const int n = n0 + u;
                    if constexpr (CORR) {
                        double fg, fk;
                        decay_fac(xi, P.lg[n], tg[u], fg);
                        decay_fac(xi, P.lk[n], tk[u], fk);
                        dsd[u] = (P.g2[n] * dev_d) * fg;
                        dso[u] = (P.g2[n] * dev_o) * fg;
                        dkd[u] = (P.k[n] * tr) * fk;
                    } else {
                        tg[u] = taylor3(xi, P.lg[n]);
                        tk[u] = taylor3(xi, P.lk[n]);
                        const double one_g = 1.0 - tg[u], one_k = 1.0 - tk[u];
                        dsd[u] = ((P.g2[n] * dev_d) / xi) * P.lg[n] * one_g;
                        dso[u] = ((P.g2[n] * dev_o) / xi) * P.lg[n] * one_g;
                        dkd[u] = ((P.k[n] * tr) / xi) * P.lk[n] * one_k;
                    }
                }
                auto item = [&](double &s, double &k, const int e) {  // e: element within the group
                    const int u = e / DD, c = e % DD;
                    const bool diag = (c % (D + 1)) == 0;
                    s = s * tg[u];
                    k = k * tk[u];
                    const double sp = (diag ? dsd[u] : dso[u]) + s, kp = (diag ? dkd[u] : 0.0) + k;
                    const double pn = sp + kp;
                    acc[c] = (n0 + u == 0) ? pn : acc[c] + pn;
                    if constexpr (CORR) {   // the history is the partial stress itself
                        s = sp;
                        k = kp;
                    }
                };
                const int off = (n0 - c0) * DD;               // position inside the staged chunk
                if constexpr (C::VEC) {
                    double2 *vs = reinterpret_cast<double2 *>(rs + off);
                    double2 *vk = reinterpret_cast<double2 *>(rk + off);
#pragma unroll
                    for (int j = 0; j < G * DD / 2; ++j) {
                        double2 s = vs[j], k = vk[j];
                        item(s.x, k.x, 2 * j);
                        item(s.y, k.y, 2 * j + 1);
                        vs[j] = s;
                        vk[j] = k;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < G * DD; ++e) {
                        double s = rs[off + e], k = rk[off + e];
                        item(s, k, e);
                        rs[off + e] = s;
                        rk[off + e] = k;
                    }
                }
            }
            if constexpr (CHUNKED) {                          // write this chunk back before its buffer is reused
                sgptx::fence_async_smem();
                __syncwarp();
                if (lane == 0) {                              // columns past the last term are not stored
                    sgptx::tensor_s2g_2d(&tm.s, c0 * DD, (int)node0, buf_s);
                    sgptx::tensor_s2g_2d(&tm.k, c0 * DD, (int)node0, buf_k);
                    sgptx::bulk_commit();
                    sgptx::bulk_wait_read0();
                }
                __syncwarp();
            }
        }
#pragma unroll
        for (int c = 0; c < DD; ++c) buf_o[lane * DD + c] = acc[c];

        // ---- write the four tiles back ----
        sgptx::fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if constexpr (!CHUNKED) {
                sgptx::bulk_s2g(f.s_tilde + node0 * ROW, buf_s, C::S_BYTES);
                sgptx::bulk_s2g(f.sigma_tilde + node0 * ROW, buf_k, C::S_BYTES);
            }
            sgptx::bulk_s2g(f.Tf_partial + node0 * N, buf_t, C::TFP_BYTES);
            sgptx::bulk_s2g(f.sigma + node0 * DD, buf_o, C::SIG_BYTES);
            sgptx::bulk_commit();
            sgptx::bulk_wait_read0();
        }
        __syncwarp();
    }
}

}  // namespace

namespace {

template <int D, bool SCALAR, bool TENSOR>
int launch_visco_d(const sg_visco_plan *plan, int64_t n, const sg_visco_fields &f, const VGather &G,
                   uint32_t phases, cudaStream_t st) {
    const size_t smem = visco_smem_bytes(plan->k.N, D * D, TENSOR);
    auto kern = visco_kernel<D, SCALAR, TENSOR>;
    SG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (n + VTILE - 1) / VTILE;
    SG_REQUIRE(grid < (int64_t)2147483647, "sg_visco_update: too many nodes for one launch");
    int bulk_ok = 0;
    if (TENSOR) {
        bulk_ok = ((reinterpret_cast<uintptr_t>(f.s_tilde) | reinterpret_cast<uintptr_t>(f.sigma_tilde)) % 16 == 0) &&
                  ((VTILE * plan->k.N * D * D) % 2 == 0);
    }
    kern<<<(unsigned)grid, VTHREADS, smem, st>>>(plan->k, f, G, (long)n, phases, bulk_ok);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

template <bool SCALAR, bool TENSOR>
int launch_visco(const sg_visco_plan *plan, int64_t n, const sg_visco_fields &f, const VGather &G,
                 uint32_t phases, cudaStream_t st) {
    if (n == 0) return SG_OK;
    switch (plan->p.dim) {
        case 1: return launch_visco_d<1, SCALAR, TENSOR>(plan, n, f, G, phases, st);
        case 2: return launch_visco_d<2, SCALAR, TENSOR>(plan, n, f, G, phases, st);
        case 3: return launch_visco_d<3, SCALAR, TENSOR>(plan, n, f, G, phases, st);
    }
    sg_set_error("sg_visco: dim must be 1, 2 or 3 (got %d)", plan->p.dim);
    return SG_E_INVALID;
}

template <int D, int N>
int setup_fast(sg_visco_plan *pl) {
    auto kern = pl->k.mode == SG_VISCO_CORRECTED ? visco_fast_kernel<D, N, true> : visco_fast_kernel<D, N, false>;
    const uint32_t smem = FastCfg<D, N>::SMEM;
    if (smem > 227u * 1024u) return SG_OK;
    SG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    SG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) return SG_OK;
    pl->fast = kern;
    pl->fast_smem = smem;
    pl->fast_grid = per_sm * pl->ctx->sm_count;
    pl->fast_chunk = FastCfg<D, N>::CH;
    return SG_OK;
}

template <int D>
int setup_fast_d(sg_visco_plan *pl) {
    switch (pl->p.n_terms) {
        case 3: return setup_fast<D, 3>(pl);
        case 4: return setup_fast<D, 4>(pl);
        case 6: return setup_fast<D, 6>(pl);
        case 8: return setup_fast<D, 8>(pl);
        case 10: return setup_fast<D, 10>(pl);
        case 12: return setup_fast<D, 12>(pl);
    }
    return SG_OK;
}

// Tensor maps of s_tilde / sigma_tilde as [rows][N*d*d] float64 with a [32][chunk*d*d] box (see ViscoTmaps); cached in the
// plan for the buffers of the last call (the problem passes the same arrays every step).
int ensure_tmaps(sg_visco_plan *pl, const sg_visco_fields *f, int64_t rows) {
    if (pl->tm_s == f->s_tilde && pl->tm_k == f->sigma_tilde && pl->tm_rows == rows) return SG_OK;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        SG_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "sg_visco_update: cuTensorMapEncodeTiled is not available in this driver");
        encode = (encode_fn)fn;
    }
    const int dd = pl->p.dim * pl->p.dim;
    const cuuint64_t row = (cuuint64_t)pl->k.N * dd;
    const cuuint64_t gdim[2] = {row, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {row * 8};
    const cuuint32_t box[2] = {(cuuint32_t)(pl->fast_chunk * dd), (cuuint32_t)WT};
    const cuuint32_t estr[2] = {1, 1};
    double *bases[2] = {f->s_tilde, f->sigma_tilde};
    CUtensorMap *maps[2] = {&pl->tmaps.s, &pl->tmaps.k};
    for (int i = 0; i < 2; ++i) {
        const CUresult r = encode(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, bases[i], gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG_REQUIRE(r == CUDA_SUCCESS, "sg_visco_update: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    pl->tm_s = f->s_tilde;
    pl->tm_k = f->sigma_tilde;
    pl->tm_rows = rows;
    return SG_OK;
}

int check_fields(const sg_visco_fields *f, uint32_t phases, bool scalar, bool tensor, int64_t n) {
    SG_REQUIRE(f != nullptr, "sg_visco: fields is NULL");
    SG_REQUIRE(n >= 0, "sg_visco: negative node count");
    SG_REQUIRE((phases & ~SG_PHASE_ALL) == 0 && phases != 0, "sg_visco: bad phase mask 0x%x", phases);
    SG_REQUIRE(f->T_cur && f->T_prev, "sg_visco: T_cur/T_prev are required");
    SG_REQUIRE(f->Tf, "sg_visco: Tf is required");
    if (scalar) {
        SG_REQUIRE(f->phi && f->xi && f->Tf_partial, "sg_visco: phi, xi and Tf_partial are required");
    }
    if (tensor && (phases & SG_PHASE_STRESS)) {
        SG_REQUIRE(f->s_tilde && f->sigma_tilde && f->sigma, "sg_visco: s_tilde, sigma_tilde, sigma are required");
    }
    if (tensor && !scalar) SG_REQUIRE(f->xi, "sg_visco: xi is required by the tensor phase");
    return SG_OK;
}

}  // namespace

extern "C" {

int sg_visco_plan_create(sg_ctx *ctx, const sg_visco_params *p, sg_visco_plan **out) {
    SG_REQUIRE(ctx && p && out, "sg_visco_plan_create: NULL argument");
    SG_REQUIRE(p->dim >= 1 && p->dim <= 3, "sg_visco_plan_create: dim must be 1..3 (got %d)", p->dim);
    SG_REQUIRE(p->n_terms >= 1 && p->n_terms <= SG_MAX_TERMS, "sg_visco_plan_create: n_terms must be 1..%d (got %d)",
               SG_MAX_TERMS, p->n_terms);
    sg_visco_plan *pl = new sg_visco_plan();
    pl->ctx = ctx;
    pl->p = *p;
    VKParams &k = pl->k;
    memset(&k, 0, sizeof(k));
    k.N = p->n_terms;
    k.c_HRg = p->H / p->Rg;
    k.inv_Tb = 1.0 / p->Tb;
    k.dt = p->dt;
    k.half_dt = p->dt / 2;
    k.inv_d = 1.0 / (double)p->dim;
    k.alpha_s = p->alpha_solid;
    k.d_alpha = p->alpha_liquid - p->alpha_solid;
    SG_REQUIRE(p->mode == SG_VISCO_REFERENCE || p->mode == SG_VISCO_CORRECTED, "sg_visco_plan_create: unknown mode %d", p->mode);
    k.mode = p->mode;
    k.chi = p->chi;
    for (int i = 0; i < p->n_terms; ++i) {
        k.m[i] = p->m[i];
        k.lm[i] = p->lambda_m[i];
        k.g2[i] = 2.0 * p->g[i];
        k.lg[i] = p->lambda_g[i];
        k.k[i] = p->k[i];
        k.lk[i] = p->lambda_k[i];
    }
    pl->fast = nullptr;
    pl->fast_smem = 0;
    pl->fast_grid = 0;
    pl->fast_chunk = 0;
    memset(&pl->tmaps, 0, sizeof(pl->tmaps));
    pl->tm_s = pl->tm_k = nullptr;
    pl->tm_rows = 0;
    int rc = SG_OK;
    if (p->dim == 1) rc = setup_fast_d<1>(pl);
    if (p->dim == 2) rc = setup_fast_d<2>(pl);
    if (p->dim == 3) rc = setup_fast_d<3>(pl);
    if (rc != SG_OK) {
        delete pl;
        return rc;
    }
    *out = pl;
    return SG_OK;
}

int sg_visco_plan_destroy(sg_visco_plan *plan) {
    delete plan;
    return SG_OK;
}

int sg_visco_update(sg_visco_plan *plan, int64_t n_nodes, const sg_visco_fields *f, uint32_t phases, void *stream) {
    SG_REQUIRE(plan, "sg_visco_update: plan is NULL");
    SG_REQUIRE(plan->k.mode == SG_VISCO_REFERENCE || phases == SG_PHASE_ALL,
               "sg_visco_update: the corrected scheme updates Tf and reads its old value in one pass; call it with SG_PHASE_ALL");
    int rc = check_fields(f, phases, true, true, n_nodes);
    if (rc) return rc;
    VGather G{0, nullptr, nullptr, nullptr};
    cudaStream_t st = (cudaStream_t)stream;
    // Steady-state call: everything but the tail (< 32 nodes) goes through the persistent fast path.
    const bool no_optional = !f->T_next && !f->phi_next && !f->thermal_strain && !f->total_strain &&
                             !f->deviatoric_strain && !f->ds_partial && !f->dsigma_partial && !f->s_partial &&
                             !f->sigma_partial;
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(f->s_tilde) | reinterpret_cast<uintptr_t>(f->sigma_tilde) |
                               reinterpret_cast<uintptr_t>(f->Tf_partial) | reinterpret_cast<uintptr_t>(f->sigma);
    int64_t done = 0;
    if (plan->fast && phases == SG_PHASE_ALL && no_optional && (align_or % 16) == 0 && n_nodes >= WT) {
        const int64_t n_tiles = n_nodes / WT;
        const int grid = (int)(n_tiles < plan->fast_grid ? n_tiles : plan->fast_grid);
        if (plan->fast_chunk < plan->k.N) {
            rc = ensure_tmaps(plan, f, n_tiles * WT);
            if (rc) return rc;
        }
        plan->fast<<<grid, 32, plan->fast_smem, st>>>(plan->k, *f, (long)n_tiles, plan->tmaps);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
        done = n_tiles * WT;
        if (done == n_nodes) return SG_OK;
    }
    sg_visco_fields t = *f;
    if (done) {
        const int64_t N = plan->k.N, dd = (int64_t)plan->p.dim * plan->p.dim;
        t.T_cur += done;
        t.T_prev += done;
        t.Tf_partial += done * N;
        t.Tf += done;
        t.phi += done;
        t.xi += done;
        t.s_tilde += done * N * dd;
        t.sigma_tilde += done * N * dd;
        t.sigma += done * dd;
    }
    return launch_visco<true, true>(plan, n_nodes - done, t, G, phases, st);
}

int sg_visco_update_scalar(sg_visco_plan *plan, int64_t n, const sg_visco_fields *f, uint32_t phases, void *stream) {
    SG_REQUIRE(plan, "sg_visco_update_scalar: plan is NULL");
    SG_REQUIRE(plan->k.mode == SG_VISCO_REFERENCE, "the corrected scheme needs equal T and sigma spaces (sg_visco_update)");
    int rc = check_fields(f, phases, true, false, n);
    if (rc) return rc;
    VGather G{0, nullptr, nullptr, nullptr};
    return launch_visco<true, false>(plan, n, *f, G, phases, (cudaStream_t)stream);
}

int sg_visco_update_tensor(sg_visco_plan *plan, int64_t n, const sg_visco_fields *f, const sg_visco_gather *g,
                           uint32_t phases, void *stream) {
    SG_REQUIRE(plan, "sg_visco_update_tensor: plan is NULL");
    SG_REQUIRE(plan->k.mode == SG_VISCO_REFERENCE, "the corrected scheme needs equal T and sigma spaces (sg_visco_update)");
    SG_REQUIRE(g && g->dofs && g->local_point && g->weights && g->n_ld > 0, "sg_visco_update_tensor: bad gather map");
    int rc = check_fields(f, phases, false, true, n);
    if (rc) return rc;
    VGather G{g->n_ld, g->dofs, g->local_point, g->weights};
    return launch_visco<false, true>(plan, n, *f, G, phases, (cudaStream_t)stream);
}

int64_t sg_visco_bytes_per_node(const sg_visco_params *p, const sg_visco_fields *f, uint32_t phases) {
    if (!p || !f) return -1;
    const int64_t N = p->n_terms, dd = (int64_t)p->dim * p->dim;
    int64_t w = 2;  // read T_cur, T_prev
    if (p->mode == SG_VISCO_CORRECTED) w += 1;                    // reads the old Tf
    if (phases & SG_PHASE_TF) w += 2 * N + 1;                    // Tf_partial r+w, Tf w
    if (phases & (SG_PHASE_TF | SG_PHASE_SHIFT)) w += 1;          // phi
    if (phases & SG_PHASE_SHIFT) w += 1 + (f->T_next ? 1 : 0) + (f->phi_next ? 1 : 0);
    if (phases & SG_PHASE_STRAIN)
        w += dd * ((f->thermal_strain ? 1 : 0) + (f->total_strain ? 1 : 0) + (f->deviatoric_strain ? 1 : 0));
    if (phases & SG_PHASE_STRESS)
        w += 4 * N * dd + dd +
             N * dd * ((f->ds_partial ? 1 : 0) + (f->dsigma_partial ? 1 : 0) + (f->s_partial ? 1 : 0) +
                       (f->sigma_partial ? 1 : 0));
    return 8 * w;
}

}  // extern "C"

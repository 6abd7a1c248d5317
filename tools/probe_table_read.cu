// Micro-probe for DESIGN.md §9 item 3: what does it cost a class kernel to read its (warp-uniform) local matrices?
//
// dg_cheb_step reads 80 doubles of class matrices per cell as 40 broadcast LDS.128; ncu attributes ~40 % of the kernel's
// L1 wavefronts to them.  This probe times the same access pattern - every lane of a warp reads the SAME 16-double
// matrix, picked by a per-warp class id, and feeds it to DFMAs - through three paths:
//   lds   shared memory, LDS.128 broadcast        (what the kernels do today)
//   ldc   __constant__ memory, indexed LDC         (constant cache instead of the L1/LSU pipe)
//   ldg   global memory via __ldg, uniform address (L1 read-only path)
// at the launch shape of the class kernels (256 threads, 3 blocks/SM).  Output: ns per matrix read per warp and the
// equivalent cycles, so the alternatives can be compared with the ~1 cycle/wavefront of the LSU pipe.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/probe_table_read tools/probe_table_read.cu
//   ./gpurun_out/probe_table_read
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int N_CLASSES = 55;   // config 3: 28 cell + 27 facet matrices
constexpr int MAT = 16;         // 4 x 4 doubles (DG1 tetrahedron)
constexpr int ITERS = 2000;     // matrix reads per thread
constexpr int TB = 256;

__constant__ double c_tab[N_CLASSES * MAT];

#define CHECK(x)                                                                            \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

// one 4x4 matrix-vector product with the matrix coming from `A` (any address space)
__device__ __forceinline__ void matvec(const double *A, const double (&x)[4], double (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double2 a = reinterpret_cast<const double2 *>(A)[2 * i], b = reinterpret_cast<const double2 *>(A)[2 * i + 1];
        y[i] = fma(a.x, x[0], fma(a.y, x[1], fma(b.x, x[2], fma(b.y, x[3], y[i]))));
    }
}

template <int PATH>   // 0 lds, 1 ldc, 2 ldg
__global__ void __launch_bounds__(TB, 3) k_probe(const double *__restrict__ g_tab, const int *__restrict__ cls, double *out) {
    __shared__ __align__(16) double s_tab[N_CLASSES * MAT];
    for (int i = threadIdx.x; i < N_CLASSES * MAT; i += TB) s_tab[i] = g_tab[i];
    __syncthreads();
    const int warp = (blockIdx.x * TB + threadIdx.x) >> 5;
    double x[4] = {1.0 + threadIdx.x, 0.5, 0.25, 0.125}, y[4] = {0.0, 0.0, 0.0, 0.0};
    int c = cls[warp % 4096];
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        const double *A = PATH == 0 ? s_tab + c * MAT : (PATH == 1 ? c_tab + c * MAT : g_tab + c * MAT);
        if (PATH == 2) {
            double t[MAT];
#pragma unroll
            for (int i = 0; i < MAT; ++i) t[i] = __ldg(A + i);
            matvec(t, x, y);
        } else {
            matvec(A, x, y);
        }
        c = (c * 5 + 7 + (int)(y[0] != 12345.678)) % N_CLASSES;   // next class depends on the result: no hoisting
        x[0] = y[1] * 1e-30 + x[0];
    }
    if (y[0] + y[1] + y[2] + y[3] == 0.123456) out[0] = y[0];
}

template <int PATH>
static void run(const char *name, const double *g_tab, const int *cls, double *out, int grid, double clock_ghz) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    k_probe<PATH><<<grid, TB>>>(g_tab, cls, out);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    for (int r = 0; r < 5; ++r) k_probe<PATH><<<grid, TB>>>(g_tab, cls, out);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 5;
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const double warps_per_sm = (double)grid * (TB / 32) / sms;
    const double ns_per_read_per_sm = ms * 1e6 / (warps_per_sm * ITERS);   // one SM issues warps_per_sm * ITERS matrix reads
    printf("{\"path\": \"%s\", \"ms\": %.4f, \"ns_per_matrix_read_per_sm\": %.3f, \"cycles_per_matrix_read_per_sm\": %.2f}\n", name, ms,
           ns_per_read_per_sm, ns_per_read_per_sm * clock_ghz);
}

int main() {
    std::vector<double> tab(N_CLASSES * MAT);
    for (size_t i = 0; i < tab.size(); ++i) tab[i] = 1.0 / (1.0 + (double)i);
    std::vector<int> cls(4096);
    for (size_t i = 0; i < cls.size(); ++i) cls[i] = (int)((i * 7919u) % N_CLASSES);
    double *g_tab, *out;
    int *g_cls;
    CHECK(cudaMalloc(&g_tab, tab.size() * sizeof(double)));
    CHECK(cudaMalloc(&g_cls, cls.size() * sizeof(int)));
    CHECK(cudaMalloc(&out, sizeof(double)));
    CHECK(cudaMemcpy(g_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(g_cls, cls.data(), cls.size() * sizeof(int), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpyToSymbol(c_tab, tab.data(), tab.size() * sizeof(double)));
    int sms = 0, khz = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    const int grid = 3 * sms;
    printf("{\"sms\": %d, \"clock_ghz_nominal\": %.3f, \"grid\": %d, \"block\": %d, \"matrix_reads_per_thread\": %d}\n", sms, ghz, grid, TB, ITERS);
    run<0>("lds_broadcast", g_tab, g_cls, out, grid, ghz);
    run<1>("ldc_indexed", g_tab, g_cls, out, grid, ghz);
    run<2>("ldg_uniform", g_tab, g_cls, out, grid, ghz);
    return 0;
}

// Equivalence classes of 64-bit keys on the GPU (operator set-up only): sort, unique, binary search.
// Used by thermal.cu to find the local-matrix classes of a mesh.
#include <thrust/binary_search.h>
#include <thrust/scan.h>
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include "sg_common.cuh"

int sg_classify_u64(const uint64_t *keys_dev, int64_t n, int32_t **cls_out, int32_t *n_cls, int32_t **rep_out) {
    SG_REQUIRE(keys_dev && cls_out && n_cls && rep_out && n > 0 && n < (int64_t)0x7fffffff, "sg_classify_u64: bad argument");
    *cls_out = nullptr;
    *rep_out = nullptr;
    try {
        thrust::device_ptr<const uint64_t> kp(keys_dev);
        thrust::device_vector<uint64_t> sorted(kp, kp + n);
        thrust::device_vector<int32_t> idx(n);
        thrust::sequence(idx.begin(), idx.end());
        thrust::sort_by_key(sorted.begin(), sorted.end(), idx.begin());
        auto ends = thrust::unique_by_key(sorted.begin(), sorted.end(), idx.begin());
        const int64_t m = ends.first - sorted.begin();
        SG_CHECK_CUDA(cudaMalloc(cls_out, sizeof(int32_t) * (size_t)n));
        SG_CHECK_CUDA(cudaMalloc(rep_out, sizeof(int32_t) * (size_t)m));
        SG_CHECK_CUDA(cudaMemcpy(*rep_out, thrust::raw_pointer_cast(idx.data()), sizeof(int32_t) * (size_t)m,
                                 cudaMemcpyDeviceToDevice));
        thrust::device_ptr<int32_t> cp(*cls_out);
        thrust::lower_bound(sorted.begin(), sorted.begin() + m, kp, kp + n, cp);
        SG_CHECK_CUDA(cudaDeviceSynchronize());
        *n_cls = (int32_t)m;
    } catch (const std::exception &e) {
        if (*cls_out) cudaFree(*cls_out);
        if (*rep_out) cudaFree(*rep_out);
        *cls_out = *rep_out = nullptr;
        sg_set_error("sg_classify_u64: %s", e.what());
        return SG_E_CUDA;
    }
    return SG_OK;
}

// In-place exclusive prefix sum of n int32 flags/counts on the device; *total = their sum (set-up only).
int sg_exclusive_scan_i32(int32_t *data_dev, int64_t n, int64_t *total) {
    SG_REQUIRE(data_dev && total && n > 0, "sg_exclusive_scan_i32: bad argument");
    try {
        thrust::device_ptr<int32_t> p(data_dev);
        int32_t last_in = 0, last_out = 0;
        SG_CHECK_CUDA(cudaMemcpy(&last_in, data_dev + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
        thrust::exclusive_scan(p, p + n, p);
        SG_CHECK_CUDA(cudaMemcpy(&last_out, data_dev + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
        *total = (int64_t)last_in + last_out;
    } catch (const std::exception &e) {
        sg_set_error("sg_exclusive_scan_i32: %s", e.what());
        return SG_E_CUDA;
    }
    return SG_OK;
}

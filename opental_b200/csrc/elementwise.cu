// HBM-bound elementwise / layout kernels: fp32 <-> bf16 hi/lo planes, clip ingest (NCDHW fp32 -> NDHWC planes).
// All are grid-stride, 16-byte vectorised where alignment allows, sized as multiples of the SM count.
#include "common.cuh"

namespace otal {

__global__ void split_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                  long long n) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        __nv_bfloat16 h[4], l[4];
        split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]);
        split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
        reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        if (lo) reinterpret_cast<uint2*>(lo)[i] = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        __nv_bfloat16 h, l;
        split_bf16(x[i], h, l);
        hi[i] = __bfloat16_as_ushort(h);
        if (lo) lo[i] = __bfloat16_as_ushort(l);
    }
}

__global__ void merge_bf16_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                  float* __restrict__ x, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = __bfloat162float(__ushort_as_bfloat16(hi[i]));
        if (lo) v += __bfloat162float(__ushort_as_bfloat16(lo[i]));
        x[i] = v;
    }
}

// One thread per (n, t, h, w) position: reads C strided fp32 values (coalesced across w), writes Cpad bf16.
__global__ void ncdhw_to_ndhwc_split_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi,
                                            uint16_t* __restrict__ lo, int N, int C, long long THW, int Cpad) {
    const long long total = (long long)N * THW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long n = i / THW, s = i - n * THW;
        for (int c = 0; c < Cpad; ++c) {
            __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
            if (c < C) split_bf16(x[(n * C + c) * THW + s], h, l);
            hi[i * Cpad + c] = __bfloat16_as_ushort(h);
            if (lo) lo[i * Cpad + c] = __bfloat16_as_ushort(l);
        }
    }
}

static inline int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = 148LL * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, long long n, void* stream) {
    if (n < 0 || (n > 0 && (!x || !hi))) { set_last_error_msg("split_bf16: bad argument"); return OTAL_ERR_BAD_ARG; }
    if (n == 0) return OTAL_OK;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(hi) & 7) ||
        (reinterpret_cast<uintptr_t>(lo) & 7)) {
        set_last_error_msg("split_bf16: pointers must be 16-byte (x) / 8-byte (hi, lo) aligned");
        return OTAL_ERR_BAD_ARG;
    }
    split_bf16_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, hi, lo, n);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_merge_bf16(const uint16_t* hi, const uint16_t* lo, float* x, long long n, void* stream) {
    if (n < 0 || (n > 0 && (!x || !hi))) { set_last_error_msg("merge_bf16: bad argument"); return OTAL_ERR_BAD_ARG; }
    if (n == 0) return OTAL_OK;
    merge_bf16_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(hi, lo, x, n);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_ncdhw_to_ndhwc_split(const float* x, uint16_t* hi, uint16_t* lo, int N, int C, int T, int H, int W,
                              int Cpad, void* stream) {
    if (N < 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0 || Cpad < C || !x || !hi) {
        set_last_error_msg("ncdhw_to_ndhwc_split: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    if (N == 0) return OTAL_OK;
    const long long THW = (long long)T * H * W;
    ncdhw_to_ndhwc_split_kernel<<<grid_for((long long)N * THW, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, hi, lo, N, C, THW, Cpad);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

// BoundaryMaxPooling forward / backward for sm_100a.
//
// Semantics follow the reference operator (AFSD/prop_pooling/boundary_max_pooling_kernel.cu:18-46 forward,
// :49-82 backward, launchers :84-145):
//   out[n,c,k] = max_{i in [l,r]} in[n,c,i],  (l,r) = trunc(seg[n,k,0:2]) if c < C/2 else trunc(seg[n,k,2:4]),
//   both clamped to [0,T-1]; r < l gives in[n,c,l]; strict '>' so NaNs never win and ties keep the first index.
//   backward adds grad_out[n,c,k] to grad_in[n,c,argmax].
//
// Design (not the reference's one-thread-per-output grid-stride loop): one CTA owns R whole rows (n, c..c+R) —
// the rows are staged once into shared memory with coalesced, vectorised loads, the K windows are scanned out
// of shared memory, and the outputs are written coalesced along k.  The backward is a *gather*: the argmax of
// every window is written to shared memory and each thread t sums, in ascending k, the grad_out entries whose
// argmax is t.  That makes the backward deterministic (the reference's float atomicAdd is not) and
// needs no zero-initialised grad_in (every element of every row is written exactly once).
//
// compat_tscale_bug=1 reproduces the reference backward's use of grad_output.size(2) (=K) as the time extent
// for clamping *and* addressing (boundary_max_pooling_kernel.cu:121): rows are then K floats
// apart in the flat input / grad_in buffers and the tail [B*C*K, B*C*T) of grad_in stays zero.
//
// Half precision (the reference dispatches AT_DISPATCH_FLOATING_TYPES_AND_HALF, boundary_max_pooling_kernel.cu:96,128): same
// kernels on __half.  Segments are half as well (the reference casts them with static_cast<int>: truncation).  The backward sums
// a frame's contributions in fp32 and rounds once — the reference's half atomicAdd rounds after every addition, in an order that
// changes from run to run, so the two agree to a few half ulps of the partial sums, not bit for bit.
#include "common.cuh"
#include <cuda_fp16.h>

namespace otal {

template <typename T> struct BmpAcc { using type = T; };
template <> struct BmpAcc<__half> { using type = float; };
template <typename A, typename T> __device__ __forceinline__ A bmp_to_acc(T v) { return static_cast<A>(v); }
template <> __device__ __forceinline__ float bmp_to_acc<float, __half>(__half v) { return __half2float(v); }
template <typename T, typename A> __device__ __forceinline__ T bmp_from_acc(A v) { return static_cast<T>(v); }
template <> __device__ __forceinline__ __half bmp_from_acc<__half, float>(float v) { return __float2half_rn(v); }

constexpr int kBmpThreads = 256;
// bytes of the staged rows, rounded up so that the int / gradient arrays behind them stay aligned (half rows of odd length)
template <typename T>
__host__ __device__ inline size_t bmp_rows_bytes(int R, int tlen) { return ((size_t)R * tlen * sizeof(T) + 15) & ~(size_t)15; }
constexpr int kBmpSmemBudget = 40 * 1024;  // bytes for staged rows

template <typename T>
__device__ __forceinline__ int seg_to_int(T v);
template <>
__device__ __forceinline__ int seg_to_int<float>(float v) { return __float2int_rz(v); }
template <>
__device__ __forceinline__ int seg_to_int<double>(double v) { return __double2int_rz(v); }
template <>
__device__ __forceinline__ int seg_to_int<__half>(__half v) { return __half2int_rz(v); }

// Stage segments (clamped ints) for sample n: seg_s[k*4 + j]
template <typename T>
__device__ __forceinline__ void stage_segments(const T* __restrict__ seg, int n, int K, int tlen, int* seg_s) {
    for (int i = threadIdx.x; i < K * 4; i += blockDim.x) {
        int v = seg_to_int<T>(seg[(size_t)n * K * 4 + i]);
        seg_s[i] = min(max(0, v), tlen - 1);
    }
}

// rows are `tstride` elements apart and `tlen` long (tlen == tstride always; kept separate for clarity)
template <typename T>
__global__ void __launch_bounds__(kBmpThreads)
bmp_forward_kernel(const T* __restrict__ in, const T* __restrict__ seg, T* __restrict__ out, int C, int tlen,
                   int K, int R, int blocks_per_sample) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* rows = reinterpret_cast<T*>(smem_raw);
    int* seg_s = reinterpret_cast<int*>(smem_raw + bmp_rows_bytes<T>(R, tlen));

    const int n = blockIdx.x / blocks_per_sample;
    const int c0 = (blockIdx.x % blocks_per_sample) * R;
    const int nrows = min(R, C - c0);
    const int half = C / 2;

    stage_segments(seg, n, K, tlen, seg_s);
    const T* src = in + ((size_t)n * C + c0) * tlen;
    const int total = nrows * tlen;
    for (int i = threadIdx.x; i < total; i += blockDim.x) rows[i] = src[i];
    __syncthreads();

    T* dst = out + ((size_t)n * C + c0) * K;
    const int nout = nrows * K;
    for (int o = threadIdx.x; o < nout; o += blockDim.x) {
        const int row = o / K, k = o - row * K;
        const int st = (c0 + row) / half;  // 0 or 1 (C even is validated on the host)
        const int l = seg_s[k * 4 + st * 2], r = seg_s[k * 4 + st * 2 + 1];
        const T* p = rows + row * tlen;
        T m = p[l];
        for (int i = l + 1; i <= r; ++i) {
            T v = p[i];
            if (v > m) m = v;
        }
        dst[o] = m;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBmpThreads)
bmp_backward_kernel(const T* __restrict__ gout, const T* __restrict__ in, const T* __restrict__ seg,
                    T* __restrict__ gin, int C, int tlen, int K, int R, int blocks_per_sample) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* rows = reinterpret_cast<T*>(smem_raw);
    int* seg_s = reinterpret_cast<int*>(smem_raw + bmp_rows_bytes<T>(R, tlen));
    int* amax_s = seg_s + K * 4;                                   // [R*K]
    T* go_s = reinterpret_cast<T*>(amax_s + (size_t)R * K + ((R * K) & 1));  // [R*K], 8-byte aligned

    const int n = blockIdx.x / blocks_per_sample;
    const int c0 = (blockIdx.x % blocks_per_sample) * R;
    const int nrows = min(R, C - c0);
    const int half = C / 2;

    stage_segments(seg, n, K, tlen, seg_s);
    const T* src = in + ((size_t)n * C + c0) * tlen;
    const int total = nrows * tlen;
    for (int i = threadIdx.x; i < total; i += blockDim.x) rows[i] = src[i];
    const T* gsrc = gout + ((size_t)n * C + c0) * K;
    const int nout = nrows * K;
    for (int o = threadIdx.x; o < nout; o += blockDim.x) go_s[o] = gsrc[o];
    __syncthreads();

    for (int o = threadIdx.x; o < nout; o += blockDim.x) {
        const int row = o / K, k = o - row * K;
        const int st = (c0 + row) / half;
        const int l = seg_s[k * 4 + st * 2], r = seg_s[k * 4 + st * 2 + 1];
        const T* p = rows + row * tlen;
        T m = p[l];
        int am = l;
        for (int i = l + 1; i <= r; ++i) {
            T v = p[i];
            if (v > m) { m = v; am = i; }
        }
        amax_s[o] = am;
    }
    __syncthreads();

    T* dst = gin + ((size_t)n * C + c0) * tlen;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int row = i / tlen, t = i - row * tlen;
        const int* am = amax_s + row * K;
        const T* g = go_s + row * K;
        using A = typename BmpAcc<T>::type;
        A acc = A(0);
        for (int k = 0; k < K; ++k)
            if (am[k] == t) acc += bmp_to_acc<A, T>(g[k]);
        dst[i] = bmp_from_acc<T, A>(acc);
    }
}

// Fallback for rows too long to stage (T*sizeof > budget): one thread per output straight from global memory;
// the backward then needs atomics and a zeroed grad_in.
template <typename T>
__global__ void bmp_forward_direct_kernel(const T* __restrict__ in, const T* __restrict__ seg, T* __restrict__ out,
                                          long long total, int C, int tlen, int K) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = idx % K;
        const int c = (idx / K) % C;
        const long long n = idx / K / C;
        const int st = c / (C / 2);
        const T* s = seg + (n * K + k) * 4 + st * 2;
        int l = min(max(0, seg_to_int<T>(s[0])), tlen - 1);
        int r = min(max(0, seg_to_int<T>(s[1])), tlen - 1);
        const T* p = in + (n * C + c) * (long long)tlen;
        T m = p[l];
        for (int i = l + 1; i <= r; ++i) {
            T v = p[i];
            if (v > m) m = v;
        }
        out[idx] = m;
    }
}
template <typename T>
__global__ void bmp_backward_direct_kernel(const T* __restrict__ gout, const T* __restrict__ in,
                                           const T* __restrict__ seg, T* __restrict__ gin, long long total, int C,
                                           int tlen, int K) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = idx % K;
        const int c = (idx / K) % C;
        const long long n = idx / K / C;
        const int st = c / (C / 2);
        const T* s = seg + (n * K + k) * 4 + st * 2;
        int l = min(max(0, seg_to_int<T>(s[0])), tlen - 1);
        int r = min(max(0, seg_to_int<T>(s[1])), tlen - 1);
        const T* p = in + (n * C + c) * (long long)tlen;
        T m = p[l];
        int am = l;
        for (int i = l + 1; i <= r; ++i) {
            T v = p[i];
            if (v > m) { m = v; am = i; }
        }
        atomicAdd(gin + (n * C + c) * (long long)tlen + am, gout[idx]);
    }
}

static int pick_rows(int tlen, int K, size_t elt, bool backward, int C) {
    // rows per CTA: as many as fit the staging budget, at most 32, at least 1; 0 = does not fit at all
    size_t per_row = (size_t)tlen * elt + (backward ? (size_t)K * (4 + elt) : 0);
    size_t fixed = (size_t)K * 16 + 32;
    if (per_row + fixed > (size_t)kBmpSmemBudget) return 0;
    int r = (int)(((size_t)kBmpSmemBudget - fixed) / per_row);
    r = r > 32 ? 32 : r;
    r = r > C ? C : r;
    // keep at least ~2 waves of CTAs on 148 SMs when the problem is big enough
    return r < 1 ? 1 : r;
}

template <typename T>
static int bmp_forward_impl(const T* in, const T* seg, T* out, int B, int C, int tlen, int K, cudaStream_t s) {
    if (B == 0 || C == 0 || K == 0) return OTAL_OK;
    int R = pick_rows(tlen, K, sizeof(T), false, C);
    if (R > 0) {
        while (R > 1 && (long long)B * ((C + R - 1) / R) < 296) R = (R + 1) / 2;
        const int bps = (C + R - 1) / R;
        size_t smem = bmp_rows_bytes<T>(R, tlen) + (size_t)K * 16;
        bmp_forward_kernel<T><<<B * bps, kBmpThreads, smem, s>>>(in, seg, out, C, tlen, K, R, bps);
    } else {
        long long total = (long long)B * C * K;
        int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
        bmp_forward_direct_kernel<T><<<grid, 256, 0, s>>>(in, seg, out, total, C, tlen, K);
    }
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

template <typename T>
static int bmp_backward_impl(const T* gout, const T* in, const T* seg, T* gin, int B, int C, int T_in, int K,
                             int compat, cudaStream_t s) {
    if (B == 0 || C == 0 || T_in == 0) return OTAL_OK;
    int tlen = T_in;
    if (compat && K != T_in) {
        if (K > T_in) {
            set_last_error_msg("bmp_backward: compat_tscale_bug with K > T would read out of bounds");
            return OTAL_ERR_BAD_ARG;
        }
        // reference addressing: rows are K apart in the flat buffers; tail of grad_in is never written
        OTAL_CUDA_TRY(cudaMemsetAsync(gin, 0, (size_t)B * C * T_in * sizeof(T), s));
        tlen = K;
    }
    if (K == 0) {
        OTAL_CUDA_TRY(cudaMemsetAsync(gin, 0, (size_t)B * C * T_in * sizeof(T), s));
        return OTAL_OK;
    }
    int R = pick_rows(tlen, K, sizeof(T), true, C);
    if (R > 0) {
        while (R > 1 && (long long)B * ((C + R - 1) / R) < 296) R = (R + 1) / 2;
        const int bps = (C + R - 1) / R;
        size_t smem = bmp_rows_bytes<T>(R, tlen) + (size_t)K * 16 + (size_t)(R * K + 1) * 4 +
                      (size_t)R * K * sizeof(T) + 8;
        bmp_backward_kernel<T><<<B * bps, kBmpThreads, smem, s>>>(gout, in, seg, gin, C, tlen, K, R, bps);
    } else {
        OTAL_CUDA_TRY(cudaMemsetAsync(gin, 0, (size_t)B * C * T_in * sizeof(T), s));
        long long total = (long long)B * C * K;
        int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
        bmp_backward_direct_kernel<T><<<grid, 256, 0, s>>>(gout, in, seg, gin, total, C, tlen, K);
    }
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

static int check_bmp_args(const void* a, const void* b, const void* c, int B, int C, int T, int K) {
    if (B < 0 || C < 0 || T < 0 || K < 0) { set_last_error_msg("bmp: negative dimension"); return OTAL_ERR_BAD_ARG; }
    if ((long long)B * C * T > 0 && (!a || !b || !c)) { set_last_error_msg("bmp: null pointer"); return OTAL_ERR_BAD_ARG; }
    if (C % 2 != 0) { set_last_error_msg("bmp: channel count must be even (start/end halves)"); return OTAL_ERR_BAD_ARG; }
    if (T < 1 && (long long)B * C * K > 0) { set_last_error_msg("bmp: empty time axis"); return OTAL_ERR_BAD_ARG; }
    return OTAL_OK;
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_bmp_forward_f32(const float* in, const float* seg, float* out, int B, int C, int T, int K, void* stream) {
    int rc = check_bmp_args(in, seg, out, B, C, T, K);
    if (rc) return rc;
    return bmp_forward_impl<float>(in, seg, out, B, C, T, K, static_cast<cudaStream_t>(stream));
}
int otal_bmp_backward_f32(const float* gout, const float* in, const float* seg, float* gin, int B, int C, int T,
                          int K, int compat_tscale_bug, void* stream) {
    int rc = check_bmp_args(in, seg, gin, B, C, T, K);
    if (rc) return rc;
    return bmp_backward_impl<float>(gout, in, seg, gin, B, C, T, K, compat_tscale_bug,
                                    static_cast<cudaStream_t>(stream));
}
int otal_bmp_forward_f64(const double* in, const double* seg, double* out, int B, int C, int T, int K,
                         void* stream) {
    int rc = check_bmp_args(in, seg, out, B, C, T, K);
    if (rc) return rc;
    return bmp_forward_impl<double>(in, seg, out, B, C, T, K, static_cast<cudaStream_t>(stream));
}
int otal_bmp_backward_f64(const double* gout, const double* in, const double* seg, double* gin, int B, int C, int T,
                          int K, int compat_tscale_bug, void* stream) {
    int rc = check_bmp_args(in, seg, gin, B, C, T, K);
    if (rc) return rc;
    return bmp_backward_impl<double>(gout, in, seg, gin, B, C, T, K, compat_tscale_bug,
                                     static_cast<cudaStream_t>(stream));
}

/* half: pointers are the raw 16-bit patterns (IEEE binary16), C has no half type */
int otal_bmp_forward_f16(const uint16_t* in, const uint16_t* seg, uint16_t* out, int B, int C, int T, int K, void* stream) {
    int rc = check_bmp_args(in, seg, out, B, C, T, K);
    if (rc) return rc;
    return bmp_forward_impl<__half>(reinterpret_cast<const __half*>(in), reinterpret_cast<const __half*>(seg),
                                    reinterpret_cast<__half*>(out), B, C, T, K, static_cast<cudaStream_t>(stream));
}
int otal_bmp_backward_f16(const uint16_t* gout, const uint16_t* in, const uint16_t* seg, uint16_t* gin, int B, int C, int T,
                          int K, int compat_tscale_bug, void* stream) {
    int rc = check_bmp_args(in, seg, gin, B, C, T, K);
    if (rc) return rc;
    return bmp_backward_impl<__half>(reinterpret_cast<const __half*>(gout), reinterpret_cast<const __half*>(in),
                                     reinterpret_cast<const __half*>(seg), reinterpret_cast<__half*>(gin), B, C, T, K,
                                     compat_tscale_bug, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

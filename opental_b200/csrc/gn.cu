// GroupNorm(32, C) + ReLU, forward and backward, on the reference's [B, C, T] fp32 layout.
//
// Replaces nn.GroupNorm(32, C) followed by nn.ReLU(inplace=True) after every pyramid / tower / proposal-branch /
// deconv convolution of CoarsePyramid (AFSD/thumos14/BDNet.py:72-73, :79-80, :86-87, :93-94, :139-140, :153-154,
// :166-167, :176-177, :276-283): 21 modules, 81 calls per forward.  torch runs 3 kernels per call
// forward and 5 backward; here it is one launch each way.
//
// One CTA per (sample, group): the group's (C/32) x T values are contiguous in [B,C,T], are read once with
// coalesced loads into shared memory, reduced with warp shuffles (two-pass mean / variance, like torch's
// RowwiseMoments, biased variance, eps inside the sqrt) and written back normalised, scaled, shifted and clamped.
// HBM-bound by construction (one read + one write of the tensor); the tensors are <= 4 MB, so in practice the
// launch latency is the roofline.  Backward recomputes the ReLU mask from the saved statistics, applies the
// closed-form GroupNorm input gradient and emits per-sample partial sums of d gamma / d beta ([B,2,C], reduced over
// the batch by the caller in a fixed order: deterministic).
#include "common.cuh"

namespace otal {

constexpr int kGnThreads = 256;

__device__ __forceinline__ float gn_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ float gn_block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = gn_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = lane < (kGnThreads >> 5) ? red[lane] : 0.f;
    return gn_warp_sum(t);
}

// x, y: [B, C, T]; group g of sample b covers channels [g*cpg, (g+1)*cpg) = n = cpg*T contiguous floats.
// staged != 0: the group fits the dynamic shared memory and is read from HBM once.
__global__ void __launch_bounds__(kGnThreads)
gn_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int C, int T, int G,
                   float eps, int relu, int staged) {
    extern __shared__ float gn_smem[];
    __shared__ float red[32];
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G, n = cpg * T;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const float* xs = x + base;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += kGnThreads) {
        const float v = xs[i];
        if (staged) gn_smem[i] = v;
        s += v;
    }
    const float mean = gn_block_sum(s, red) / (float)n;
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += kGnThreads) {
        const float d = (staged ? gn_smem[i] : xs[i]) - mean;
        q += d * d;
    }
    const float var = gn_block_sum(q, red) / (float)n;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) { mean_out[blockIdx.x] = mean; rstd_out[blockIdx.x] = rstd; }
    float* ys = y + base;
    for (int i = threadIdx.x; i < n; i += kGnThreads) {
        const int c = g * cpg + i / T;
        float v = ((staged ? gn_smem[i] : xs[i]) - mean) * rstd * gamma[c] + beta[c];
        if (relu) v = fmaxf(v, 0.f);
        ys[i] = v;
    }
}

// gx = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)),  dxh = gy * [y > 0] * gamma[c],  xh = (x - mean) * rstd
// dgb[b][0][c] = sum_t gy*[y>0]*xh,  dgb[b][1][c] = sum_t gy*[y>0]
__global__ void __launch_bounds__(kGnThreads)
gn_relu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                   float* __restrict__ gx, float* __restrict__ dgb, int C, int T, int G, int relu, int staged) {
    extern __shared__ float gn_smem[];      // [2][n]: xh, masked gy (when staged)
    __shared__ float red[32];
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G, n = cpg * T;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const float mean = mean_in[blockIdx.x], rstd = rstd_in[blockIdx.x];
    const float* xs = x + base;
    const float* gs = gy + base;
    float* sxh = gn_smem;
    float* sg = gn_smem + n;
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < n; i += kGnThreads) {
        const int c = g * cpg + i / T;
        const float xh = (xs[i] - mean) * rstd;
        float gv = gs[i];
        if (relu && xh * gamma[c] + beta[c] <= 0.f) gv = 0.f;
        if (staged) { sxh[i] = xh; sg[i] = gv; }
        const float dxh = gv * gamma[c];
        s1 += dxh; s2 += dxh * xh;
    }
    const float m1 = gn_block_sum(s1, red) / (float)n;
    const float m2 = gn_block_sum(s2, red) / (float)n;
    float* gxs = gx + base;
    for (int i = threadIdx.x; i < n; i += kGnThreads) {
        const int c = g * cpg + i / T;
        float xh, gv;
        if (staged) { xh = sxh[i]; gv = sg[i]; }
        else {
            xh = (xs[i] - mean) * rstd; gv = gs[i];
            if (relu && xh * gamma[c] + beta[c] <= 0.f) gv = 0.f;
        }
        gxs[i] = rstd * (gv * gamma[c] - m1 - xh * m2);
    }
    // per-channel parameter gradients of this sample: one warp per channel, fixed order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cl = warp; cl < cpg; cl += (kGnThreads >> 5)) {
        const int c = g * cpg + cl;
        float a = 0.f, bsum = 0.f;
        for (int t = lane; t < T; t += 32) {
            const int i = cl * T + t;
            float xh, gv;
            if (staged) { xh = sxh[i]; gv = sg[i]; }
            else {
                xh = (xs[i] - mean) * rstd; gv = gs[i];
                if (relu && xh * gamma[c] + beta[c] <= 0.f) gv = 0.f;
            }
            a += gv * xh; bsum += gv;
        }
        a = gn_warp_sum(a); bsum = gn_warp_sum(bsum);
        if (lane == 0) {
            dgb[((size_t)b * 2 + 0) * C + c] = a;
            dgb[((size_t)b * 2 + 1) * C + c] = bsum;
        }
    }
}

static int gn_check(int B, int C, int T, int G) {
    if (B <= 0 || C <= 0 || T <= 0 || G <= 0 || C % G) { set_last_error_msg("groupnorm: bad dimensions (C must be a multiple of the group count)"); return OTAL_ERR_BAD_ARG; }
    return OTAL_OK;
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_groupnorm_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int B,
                            int C, int T, int groups, float eps, int relu, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = gn_check(B, C, T, groups);
    if (rc) return rc;
    if (!x || !gamma || !beta || !y || !mean || !rstd) { set_last_error_msg("groupnorm: null pointer"); return OTAL_ERR_BAD_ARG; }
    const size_t n = (size_t)(C / groups) * T;
    const int staged = n * 4 <= 96 * 1024;
    static bool configured = false;
    if (!configured) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        configured = true;
    }
    gn_relu_fwd_kernel<<<B * groups, kGnThreads, staged ? n * 4 : 0, stream>>>(x, gamma, beta, y, mean, rstd, C, T, groups, eps,
                                                                              relu, staged);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_groupnorm_relu_bwd(const float* gy, const float* x, const float* gamma, const float* beta, const float* mean,
                            const float* rstd, float* gx, float* dgamma_dbeta, int B, int C, int T, int groups, int relu,
                            void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = gn_check(B, C, T, groups);
    if (rc) return rc;
    if (!gy || !x || !gamma || !beta || !mean || !rstd || !gx || !dgamma_dbeta) { set_last_error_msg("groupnorm: null pointer"); return OTAL_ERR_BAD_ARG; }
    const size_t n = (size_t)(C / groups) * T;
    const int staged = n * 8 <= 192 * 1024;
    static bool configured = false;
    if (!configured) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        configured = true;
    }
    gn_relu_bwd_kernel<<<B * groups, kGnThreads, staged ? n * 8 : 0, stream>>>(gy, x, gamma, beta, mean, rstd, gx, dgamma_dbeta,
                                                                              C, T, groups, relu, staged);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

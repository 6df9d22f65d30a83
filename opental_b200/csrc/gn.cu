// GroupNorm(32, C) + ReLU, forward and backward, on the reference's [B, C, T] fp32 layout — optionally SEGMENTED along T.
//
// Replaces nn.GroupNorm(32, C) followed by nn.ReLU(inplace=True) after every pyramid / tower / proposal-branch /
// deconv convolution of CoarsePyramid (AFSD/thumos14/BDNet.py:72-73, :79-80, :86-87, :93-94, :139-140, :153-154,
// :166-167, :176-177, :276-283): 21 modules, 81 calls per forward.  torch runs 3 kernels per call
// forward and 5 backward; here it is one launch each way.
//
// Segments: the towers, heads and proposal branches share their weights across the 6 pyramid levels
// (BDNet.py:333-412), so the host runs them ONCE on all levels laid side by side along T.  GroupNorm statistics
// must then stay per (sample, level, group): the kernel takes up to 8 (offset, length) column ranges, normalises
// each range on its own, and writes ZERO to every column outside the ranges (the separator columns that give the
// k=3 "same" convolutions their per-level zero padding).  nseg = 1, (0, T) is plain GroupNorm.
//
// One CTA per (sample, group): the group's (C/32) x T values are read once with coalesced loads into shared memory,
// reduced with warp shuffles (two-pass mean / variance, biased variance, eps inside the sqrt, like torch) and written
// back normalised, scaled, shifted and clamped.  HBM-bound by construction (one read + one write of the tensor); the
// tensors are <= 4 MB, so in practice the launch latency is the roofline.  Backward recomputes the ReLU mask from the
// saved statistics, applies the closed-form GroupNorm input gradient and emits per-sample partial sums of
// d gamma / d beta ([B,2,C], reduced over the batch by the caller in a fixed order: deterministic).
#include "common.cuh"

namespace otal {

constexpr int kGnThreads = 256;
constexpr int kGnMaxSeg = 8;

struct GnSegs { int n; int off[kGnMaxSeg]; int len[kGnMaxSeg]; };

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

// x, y: [B, C, T]; group g of sample b covers channels [g*cpg, (g+1)*cpg): cpg rows of T floats, contiguous.
// staged != 0: the rows fit the dynamic shared memory and are read from HBM once.
__global__ void __launch_bounds__(kGnThreads)
gn_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int C, int T, int G,
                   float eps, int relu, int staged, const GnSegs segs) {
    extern __shared__ float gn_smem[];
    __shared__ float red[32];
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G, n = cpg * T;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const float* xs = x + base;
    float* ys = y + base;
    if (staged) {
        for (int i = threadIdx.x; i < n; i += kGnThreads) gn_smem[i] = xs[i];
        __syncthreads();
    }
    const float* src = staged ? gn_smem : xs;
    if (segs.n != 1 || segs.off[0] != 0 || segs.len[0] != T)      // columns outside every segment are separators
        for (int i = threadIdx.x; i < n; i += kGnThreads) ys[i] = 0.f;
    for (int s = 0; s < segs.n; ++s) {
        const int off = segs.off[s], len = segs.len[s], ns = cpg * len;
        float a = 0.f;
        for (int i = threadIdx.x; i < ns; i += kGnThreads) a += src[(i / len) * T + off + i % len];
        const float mean = gn_block_sum(a, red) / (float)ns;
        float q = 0.f;
        for (int i = threadIdx.x; i < ns; i += kGnThreads) {
            const float d = src[(i / len) * T + off + i % len] - mean;
            q += d * d;
        }
        const float var = gn_block_sum(q, red) / (float)ns;
        const float rstd = rsqrtf(var + eps);
        if (threadIdx.x == 0) { mean_out[blockIdx.x * segs.n + s] = mean; rstd_out[blockIdx.x * segs.n + s] = rstd; }
        __syncthreads();                                           // the zero fill above must land before the segment stores
        for (int i = threadIdx.x; i < ns; i += kGnThreads) {
            const int cl = i / len, idx = cl * T + off + i % len;
            const int c = g * cpg + cl;
            float v = (src[idx] - mean) * rstd * gamma[c] + beta[c];
            if (relu) v = fmaxf(v, 0.f);
            ys[idx] = v;
        }
    }
}

// gx = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)),  dxh = gy * [y > 0] * gamma[c],  xh = (x - mean) * rstd
// dgb[b][0][c] = sum_t gy*[y>0]*xh,  dgb[b][1][c] = sum_t gy*[y>0]   (sums over the segment columns only)
__global__ void __launch_bounds__(kGnThreads)
gn_relu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                   float* __restrict__ gx, float* __restrict__ dgb, int C, int T, int G, int relu, const GnSegs segs) {
    extern __shared__ float gn_smem[];      // [2][n]: xh, masked gy
    __shared__ float red[32];
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G, n = cpg * T;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const float* xs = x + base;
    const float* gs = gy + base;
    float* gxs = gx + base;
    float* sxh = gn_smem;
    float* sg = gn_smem + n;
    for (int i = threadIdx.x; i < n; i += kGnThreads) { sxh[i] = 0.f; sg[i] = 0.f; gxs[i] = 0.f; }
    __syncthreads();
    for (int s = 0; s < segs.n; ++s) {
        const int off = segs.off[s], len = segs.len[s], ns = cpg * len;
        const float mean = mean_in[blockIdx.x * segs.n + s], rstd = rstd_in[blockIdx.x * segs.n + s];
        float s1 = 0.f, s2 = 0.f;
        for (int i = threadIdx.x; i < ns; i += kGnThreads) {
            const int cl = i / len, idx = cl * T + off + i % len;
            const int c = g * cpg + cl;
            const float xh = (xs[idx] - mean) * rstd;
            float gv = gs[idx];
            if (relu && xh * gamma[c] + beta[c] <= 0.f) gv = 0.f;
            sxh[idx] = xh; sg[idx] = gv;
            const float dxh = gv * gamma[c];
            s1 += dxh; s2 += dxh * xh;
        }
        const float m1 = gn_block_sum(s1, red) / (float)ns;
        const float m2 = gn_block_sum(s2, red) / (float)ns;
        for (int i = threadIdx.x; i < ns; i += kGnThreads) {
            const int cl = i / len, idx = cl * T + off + i % len;
            const int c = g * cpg + cl;
            gxs[idx] = rstd * (sg[idx] * gamma[c] - m1 - sxh[idx] * m2);
        }
    }
    __syncthreads();
    // per-channel parameter gradients of this sample: one warp per channel, fixed order (separators hold zeros)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cl = warp; cl < cpg; cl += (kGnThreads >> 5)) {
        const int c = g * cpg + cl;
        float a = 0.f, bsum = 0.f;
        for (int t = lane; t < T; t += 32) {
            const int i = cl * T + t;
            a += sg[i] * sxh[i]; bsum += sg[i];
        }
        a = gn_warp_sum(a); bsum = gn_warp_sum(bsum);
        if (lane == 0) {
            dgb[((size_t)b * 2 + 0) * C + c] = a;
            dgb[((size_t)b * 2 + 1) * C + c] = bsum;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Extended forms for the explicit head schedule (opental_b200/head_schedule.py): the same arithmetic, more destinations.
//   forward:  besides (or instead of) y [B,C,T] fp32, the result as channels-last bf16 hi/lo PLANES [B,T,p_cstride] at channel
//             offset p_coff — the operand layout of the tensor-core convolution that consumes it (replaces one
//             ncl_to_nlc_split launch per conv input, and torch.cat when the destination is a slice of a concat buffer)
//   backward: gy may be a channel slice of a wider [B,Ctot,T] tensor (gy_bstride); gx is written as channels-last planes
//             [B,T,C] — the output-gradient operand of the preceding conv's data / weight gradient — and the parameter
//             gradients are ACCUMULATED in place: dgamma[c], dbeta[c], and dbias[c] += sum_t gx (the bias gradient of the
//             preceding conv: its output feeds only this GroupNorm), atomically over the batch.
// ---------------------------------------------------------------------------------------------------------------------
struct GnExFwd {
    const float *x, *gamma, *beta;
    float *y, *mean, *rstd;
    uint16_t *p_hi, *p_lo;
    int C, T, G, relu, p_cstride, p_coff;
    float eps;
    float* yt;            // optional fp32 channels-last copy [B, yt_T, C] of the columns [yt_off, yt_off + yt_T)
    int yt_off, yt_T;
};

__global__ void __launch_bounds__(kGnThreads)
gn_relu_fwd_ex_kernel(const GnExFwd p, const GnSegs segs) {
    extern __shared__ float gn_smem[];      // [cpg][T]: x, overwritten by y
    __shared__ float red[32];
    const int C = p.C, T = p.T, G = p.G;
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G, n = cpg * T;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const float* xs = p.x + base;
    for (int i = threadIdx.x; i < n; i += kGnThreads) gn_smem[i] = xs[i];
    __syncthreads();
    float stat_m[kGnMaxSeg], stat_r[kGnMaxSeg];
    for (int s = 0; s < segs.n; ++s) {
        const int off = segs.off[s], len = segs.len[s], ns = cpg * len;
        float a = 0.f;
        for (int i = threadIdx.x; i < ns; i += kGnThreads) a += gn_smem[(i / len) * T + off + i % len];
        const float mean = gn_block_sum(a, red) / (float)ns;
        float q = 0.f;
        for (int i = threadIdx.x; i < ns; i += kGnThreads) {
            const float d = gn_smem[(i / len) * T + off + i % len] - mean;
            q += d * d;
        }
        const float var = gn_block_sum(q, red) / (float)ns;
        stat_m[s] = mean; stat_r[s] = rsqrtf(var + p.eps);
        if (threadIdx.x == 0) { p.mean[blockIdx.x * segs.n + s] = mean; p.rstd[blockIdx.x * segs.n + s] = stat_r[s]; }
    }
    __syncthreads();
    // normalise in place; columns outside every segment (the separators) become 0
    for (int i = threadIdx.x; i < n; i += kGnThreads) {
        const int cl = i / T, t = i - cl * T;
        float v = 0.f;
#pragma unroll 1
        for (int s = 0; s < segs.n; ++s)
            if (t >= segs.off[s] && t < segs.off[s] + segs.len[s]) {
                const int c = g * cpg + cl;
                v = (gn_smem[i] - stat_m[s]) * stat_r[s] * p.gamma[c] + p.beta[c];
                if (p.relu) v = fmaxf(v, 0.f);
            }
        gn_smem[i] = v;
        if (p.y) p.y[base + i] = v;
    }
    if (!p.p_hi && !p.yt) return;
    __syncthreads();
    if (p.yt)
        for (int i = threadIdx.x; i < p.yt_T * cpg; i += kGnThreads) {
            const int t = i / cpg, cl = i - t * cpg;
            p.yt[((size_t)b * p.yt_T + t) * C + g * cpg + cl] = gn_smem[cl * T + p.yt_off + t];
        }
    if (!p.p_hi) return;
    // channels-last planes: thread = (t, channel pair)
    const int half = cpg >> 1;
    for (int i = threadIdx.x; i < T * half; i += kGnThreads) {
        const int t = i / half, j = i - t * half;
        uint32_t h, l;
        split_bf16x2(gn_smem[(2 * j) * T + t], gn_smem[(2 * j + 1) * T + t], h, l);
        const size_t o = ((size_t)b * T + t) * p.p_cstride + p.p_coff + g * cpg + 2 * j;
        *reinterpret_cast<uint32_t*>(p.p_hi + o) = h;
        if (p.p_lo) *reinterpret_cast<uint32_t*>(p.p_lo + o) = l;
    }
}

struct GnExBwd {
    const float *gy, *x, *gamma, *beta, *mean, *rstd;
    // optional extra gradient in channels-last form for the columns [gy2_off, gy2_off + gy2_T): one [B, gy2_T, C/2] tensor per
    // channel half (the start / end maps that the training script reads off this feature: BDNet.py:331, :392-395)
    const float *gy2a, *gy2b;
    int gy2_off, gy2_T;
    long long gy_bstride;
    float* gx;
    uint16_t *d_hi, *d_lo;
    float *dgamma, *dbeta, *dbias;
    int C, T, G, relu;
};

__global__ void __launch_bounds__(kGnThreads)
gn_relu_bwd_ex_kernel(const GnExBwd p, const GnSegs segs) {
    extern __shared__ float gn_smem[];      // [2][n]: xh -> gx, masked gy
    __shared__ float red[32];
    const int C = p.C, T = p.T, G = p.G;
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G, n = cpg * T;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const float* xs = p.x + base;
    const float* gs = p.gy ? p.gy + (size_t)b * p.gy_bstride + (size_t)g * cpg * T : nullptr;
    float* sxh = gn_smem;
    float* sg = gn_smem + n;
    for (int i = threadIdx.x; i < n; i += kGnThreads) { sxh[i] = 0.f; sg[i] = 0.f; }
    __syncthreads();
    float stat_r[kGnMaxSeg], m1s[kGnMaxSeg], m2s[kGnMaxSeg];
    for (int s = 0; s < segs.n; ++s) {
        const int off = segs.off[s], len = segs.len[s], ns = cpg * len;
        const float mean = p.mean[blockIdx.x * segs.n + s], rstd = p.rstd[blockIdx.x * segs.n + s];
        float s1 = 0.f, s2 = 0.f;
        for (int i = threadIdx.x; i < ns; i += kGnThreads) {
            const int cl = i / len, idx = cl * T + off + i % len;
            const int c = g * cpg + cl;
            const float xh = (xs[idx] - mean) * rstd;
            float gv = p.gy ? gs[idx] : 0.f;
            const int t = off + i % len, t2 = t - p.gy2_off;
            if (t2 >= 0 && t2 < p.gy2_T) {
                const int hc = C >> 1;
                const float* extra = c < hc ? p.gy2a : p.gy2b;
                if (extra) gv += extra[((size_t)b * p.gy2_T + t2) * hc + (c < hc ? c : c - hc)];
            }
            if (p.relu && xh * p.gamma[c] + p.beta[c] <= 0.f) gv = 0.f;
            sxh[idx] = xh; sg[idx] = gv;
            const float dxh = gv * p.gamma[c];
            s1 += dxh; s2 += dxh * xh;
        }
        stat_r[s] = rstd;
        m1s[s] = gn_block_sum(s1, red) / (float)ns;
        m2s[s] = gn_block_sum(s2, red) / (float)ns;
    }
    __syncthreads();
    // per-channel parameter gradients (one warp per channel) BEFORE xh is overwritten by gx
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cl = warp; cl < cpg; cl += (kGnThreads >> 5)) {
        const int c = g * cpg + cl;
        float a = 0.f, bsum = 0.f;
        for (int t = lane; t < T; t += 32) {
            const int i = cl * T + t;
            a += sg[i] * sxh[i]; bsum += sg[i];
        }
        a = gn_warp_sum(a); bsum = gn_warp_sum(bsum);
        if (lane == 0) { atomicAdd(p.dgamma + c, a); atomicAdd(p.dbeta + c, bsum); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kGnThreads) {
        const int cl = i / T, t = i - cl * T;
        float v = 0.f;
#pragma unroll 1
        for (int s = 0; s < segs.n; ++s)
            if (t >= segs.off[s] && t < segs.off[s] + segs.len[s])
                v = stat_r[s] * (sg[i] * p.gamma[g * cpg + cl] - m1s[s] - sxh[i] * m2s[s]);
        sxh[i] = v;
        if (p.gx) p.gx[base + i] = v;
    }
    __syncthreads();
    if (p.dbias)
        for (int cl = warp; cl < cpg; cl += (kGnThreads >> 5)) {
            float a = 0.f;
            for (int t = lane; t < T; t += 32) a += sxh[cl * T + t];
            a = gn_warp_sum(a);
            if (lane == 0) atomicAdd(p.dbias + g * cpg + cl, a);
        }
    if (!p.d_hi) return;
    const int half = cpg >> 1;
    for (int i = threadIdx.x; i < T * half; i += kGnThreads) {
        const int t = i / half, j = i - t * half;
        uint32_t h, l;
        split_bf16x2(sxh[(2 * j) * T + t], sxh[(2 * j + 1) * T + t], h, l);
        const size_t o = ((size_t)b * T + t) * C + g * cpg + 2 * j;
        *reinterpret_cast<uint32_t*>(p.d_hi + o) = h;
        if (p.d_lo) *reinterpret_cast<uint32_t*>(p.d_lo + o) = l;
    }
}

static int gn_setup(int B, int C, int T, int G, int nseg, const int* seg_off, const int* seg_len, GnSegs& segs) {
    if (B <= 0 || C <= 0 || T <= 0 || G <= 0 || C % G) { set_last_error_msg("groupnorm: bad dimensions (C must be a multiple of the group count)"); return OTAL_ERR_BAD_ARG; }
    if (nseg < 0 || nseg > kGnMaxSeg || (nseg > 0 && (!seg_off || !seg_len))) { set_last_error_msg("groupnorm: at most 8 segments"); return OTAL_ERR_BAD_ARG; }
    if (nseg == 0) { segs.n = 1; segs.off[0] = 0; segs.len[0] = T; return OTAL_OK; }
    segs.n = nseg;
    for (int s = 0; s < nseg; ++s) {
        if (seg_off[s] < 0 || seg_len[s] <= 0 || seg_off[s] + seg_len[s] > T) { set_last_error_msg("groupnorm: segment outside [0,T)"); return OTAL_ERR_BAD_ARG; }
        segs.off[s] = seg_off[s]; segs.len[s] = seg_len[s];
    }
    return OTAL_OK;
}

static int gn_configure() {
    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_fwd_ex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        OTAL_CUDA_TRY(cudaFuncSetAttribute(gn_relu_bwd_ex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        once.mark(once_dev);
    }
    return OTAL_OK;
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_groupnorm_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int B,
                            int C, int T, int groups, float eps, int relu, int nseg, const int* seg_off, const int* seg_len,
                            void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    GnSegs segs{};
    int rc = gn_setup(B, C, T, groups, nseg, seg_off, seg_len, segs);
    if (rc) return rc;
    if (!x || !gamma || !beta || !y || !mean || !rstd) { set_last_error_msg("groupnorm: null pointer"); return OTAL_ERR_BAD_ARG; }
    const size_t n = (size_t)(C / groups) * T;
    const int staged = n * 4 <= 96 * 1024;
    if ((rc = gn_configure())) return rc;
    gn_relu_fwd_kernel<<<B * groups, kGnThreads, staged ? n * 4 : 0, stream>>>(x, gamma, beta, y, mean, rstd, C, T, groups, eps,
                                                                              relu, staged, segs);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_groupnorm_relu_bwd(const float* gy, const float* x, const float* gamma, const float* beta, const float* mean,
                            const float* rstd, float* gx, float* dgamma_dbeta, int B, int C, int T, int groups, int relu,
                            int nseg, const int* seg_off, const int* seg_len, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    GnSegs segs{};
    int rc = gn_setup(B, C, T, groups, nseg, seg_off, seg_len, segs);
    if (rc) return rc;
    if (!gy || !x || !gamma || !beta || !mean || !rstd || !gx || !dgamma_dbeta) { set_last_error_msg("groupnorm: null pointer"); return OTAL_ERR_BAD_ARG; }
    const size_t n = (size_t)(C / groups) * T;
    if (n * 8 > 192 * 1024) { set_last_error_msg("groupnorm backward: (C/groups)*T exceeds the shared-memory staging (24576 values)"); return OTAL_ERR_UNSUPPORTED; }
    if ((rc = gn_configure())) return rc;
    gn_relu_bwd_kernel<<<B * groups, kGnThreads, n * 8, stream>>>(gy, x, gamma, beta, mean, rstd, gx, dgamma_dbeta, C, T, groups,
                                                                  relu, segs);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_groupnorm_relu_fwd_ex(const otal_gn_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("groupnorm_ex: null descriptor"); return OTAL_ERR_BAD_ARG; }
    GnSegs segs{};
    int rc = gn_setup(d->B, d->C, d->T, d->groups, d->nseg, d->seg_off, d->seg_len, segs);
    if (rc) return rc;
    if (!d->x || !d->gamma || !d->beta || !d->mean || !d->rstd || (!d->y && !d->p_hi && !d->yt)) { set_last_error_msg("groupnorm_ex: null pointer"); return OTAL_ERR_BAD_ARG; }
    const int cpg = d->C / d->groups;
    if (d->p_hi && ((cpg & 1) || (d->p_coff & 1) || (d->p_cstride & 1) || d->p_coff + d->C > d->p_cstride)) {
        set_last_error_msg("groupnorm_ex: planes need an even group width and an even channel slice inside p_cstride"); return OTAL_ERR_BAD_ARG;
    }
    const size_t n = (size_t)cpg * d->T;
    if (n * 4 > 96 * 1024) { set_last_error_msg("groupnorm_ex: (C/groups)*T exceeds the shared-memory staging (24576 values)"); return OTAL_ERR_UNSUPPORTED; }
    if ((rc = gn_configure())) return rc;
    if (d->yt && (d->yt_off < 0 || d->yt_T <= 0 || d->yt_off + d->yt_T > d->T)) { set_last_error_msg("groupnorm_ex: yt column range outside [0,T)"); return OTAL_ERR_BAD_ARG; }
    GnExFwd p{d->x, d->gamma, d->beta, d->y, d->mean, d->rstd, d->p_hi, d->p_lo, d->C, d->T, d->groups, d->relu, d->p_cstride, d->p_coff, d->eps,
              d->yt, d->yt_off, d->yt_T};
    gn_relu_fwd_ex_kernel<<<d->B * d->groups, kGnThreads, n * 4, stream>>>(p, segs);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_groupnorm_relu_bwd_ex(const otal_gn_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("groupnorm_ex: null descriptor"); return OTAL_ERR_BAD_ARG; }
    GnSegs segs{};
    int rc = gn_setup(d->B, d->C, d->T, d->groups, d->nseg, d->seg_off, d->seg_len, segs);
    if (rc) return rc;
    if ((!d->gy && !d->gy2a && !d->gy2b) || !d->x || !d->gamma || !d->beta || !d->mean || !d->rstd || !d->dgamma || !d->dbeta || (!d->gx && !d->d_hi)) {
        set_last_error_msg("groupnorm_ex backward: null pointer"); return OTAL_ERR_BAD_ARG;
    }
    if ((d->gy2a || d->gy2b) && (d->gy2_off < 0 || d->gy2_T <= 0 || d->gy2_off + d->gy2_T > d->T || (d->C & 1))) {
        set_last_error_msg("groupnorm_ex backward: gy2 column range outside [0,T)"); return OTAL_ERR_BAD_ARG;
    }
    const int cpg = d->C / d->groups;
    if (d->d_hi && (cpg & 1)) { set_last_error_msg("groupnorm_ex backward: planes need an even group width"); return OTAL_ERR_BAD_ARG; }
    const size_t n = (size_t)cpg * d->T;
    if (n * 8 > 192 * 1024) { set_last_error_msg("groupnorm_ex backward: (C/groups)*T exceeds the shared-memory staging (24576 values)"); return OTAL_ERR_UNSUPPORTED; }
    if ((rc = gn_configure())) return rc;
    GnExBwd p{d->gy, d->x, d->gamma, d->beta, d->mean, d->rstd, d->gy2a, d->gy2b, d->gy2_off, (d->gy2a || d->gy2b) ? d->gy2_T : 0,
              d->gy_bstride > 0 ? d->gy_bstride : (long long)d->C * d->T, d->gx,
              d->d_hi, d->d_lo, d->dgamma, d->dbeta, d->dbias, d->C, d->T, d->groups, d->relu};
    gn_relu_bwd_ex_kernel<<<d->B * d->groups, kGnThreads, n * 8, stream>>>(p, segs);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

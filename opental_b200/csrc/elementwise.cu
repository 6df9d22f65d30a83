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


// Clip ingest for Conv3d_1a: NCDHW fp32 [N,C,T,H,W] -> [N,T,H,Wp,4] bf16 planes with Wp = W + 8: image column w lands at
// padded column w + 2 (2 zero columns left, 6 right), 4 channel slots per pixel (zero for c >= C).  The 8-pixel x 4-slot
// window of output column w' of the stride-2 conv is then the 32 contiguous elements starting at padded column 2*w'
// (read by TMA through a strided tensor map: overlapping windows, no expansion in memory).  One thread per padded pixel:
// reads are coalesced along w per channel plane, each thread writes 8 bytes per plane.
__global__ void clip_ingest_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                   int N, int C, int T, int H, int W) {
    const int Wp = W + 8;
    const long long total = (long long)N * T * H * Wp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long THW = (long long)T * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int wp = (int)(i % Wp);
        const long long row = i / Wp;                 // (n*T + t)*H + h
        const int w = wp - 2;
        uint32_t h32[2] = {0, 0}, l32[2] = {0, 0};
        if (w >= 0 && w < W) {
            const long long n = row / ((long long)T * H);
            const long long th = row - n * (long long)T * H;
            const float* src = x + n * C * THW + th * W + w;
            for (int c = 0; c < C && c < 4; ++c) {
                __nv_bfloat16 hb, lb;
                split_bf16(src[c * THW], hb, lb);
                h32[c >> 1] |= (uint32_t)__bfloat16_as_ushort(hb) << ((c & 1) * 16);
                l32[c >> 1] |= (uint32_t)__bfloat16_as_ushort(lb) << ((c & 1) * 16);
            }
        }
        reinterpret_cast<uint2*>(hi)[i] = make_uint2(h32[0], h32[1]);
        if (lo) reinterpret_cast<uint2*>(lo)[i] = make_uint2(l32[0], l32[1]);
    }
}

// Backward of  y = relu(conv * scale + shift)  w.r.t. the conv output, fused with the bf16 hi/lo split that the
// dgrad / wgrad tensor-core kernels consume:  d = g * [y > 0] * scale[c].   (ReLU backward of torch is
// grad * (result > 0); frozen BN contributes only its scale, AFSD/thumos14/BDNet.py:39-49.)
// One thread per (position, 8 channels): 32 B of g + 16 B of y_hi in, 2 x 16 B out.
__global__ void relu_bn_bwd_split_kernel(const float* __restrict__ g, const uint16_t* __restrict__ y_hi,
                                         const float* __restrict__ scale, uint16_t* __restrict__ d_hi,
                                         uint16_t* __restrict__ d_lo, long long npos, int C, int g_cs, int g_co,
                                         int y_cs, int y_co, int d_cs, int d_co, int relu) {
    const int cgs = C >> 3;
    const long long total = npos * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cg = (int)(i % cgs);
        const long long pos = i / cgs;
        const float* gp = g + pos * g_cs + g_co + cg * 8;
        const float4 a = *reinterpret_cast<const float4*>(gp), b = *reinterpret_cast<const float4*>(gp + 4);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t yw[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
        if (relu) {
            const uint4 yv = *reinterpret_cast<const uint4*>(y_hi + pos * y_cs + y_co + cg * 8);
            yw[0] = yv.x; yw[1] = yv.y; yw[2] = yv.z; yw[3] = yv.w;
        }
        uint32_t ho[4], lo_[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x0 = v[2 * j], x1 = v[2 * j + 1];
            // y > 0  <=>  bf16 bits: sign clear and magnitude non-zero
            const uint32_t y0 = yw[j] & 0xffffu, y1 = yw[j] >> 16;
            if (!(y0 != 0 && y0 < 0x8000u)) x0 = 0.f;
            if (!(y1 != 0 && y1 < 0x8000u)) x1 = 0.f;
            if (scale) { x0 *= __ldg(scale + cg * 8 + 2 * j); x1 *= __ldg(scale + cg * 8 + 2 * j + 1); }
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(x0, h0, l0); split_bf16(x1, h1, l1);
            ho[j] = pack_bf16x2(h0, h1); lo_[j] = pack_bf16x2(l0, l1);
        }
        const size_t off = (size_t)pos * d_cs + d_co + cg * 8;
        *reinterpret_cast<uint4*>(d_hi + off) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
        if (d_lo) *reinterpret_cast<uint4*>(d_lo + off) = make_uint4(lo_[0], lo_[1], lo_[2], lo_[3]);
    }
}

// Adam with classic L2 weight decay (grad += wd * p before the moments; torch.optim.Adam as used by
// AFSD/thumos14/train.py:321-323) over one flat fp32 parameter buffer, optional 1/world gradient
// scaling fused in (data-parallel all-reduce epilogue).  step_size = lr / (1 - beta1^t), bc2 = 1 - beta2^t.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                            float wd, float grad_scale, float bc1, float bc2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float pi = p[i];
        float gi = g[i] * grad_scale + wd * pi;
        float mi = beta1 * m[i] + (1.f - beta1) * gi;
        float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

// The same update with the step counter read from device memory: the launch can then be part of a captured CUDA graph (the bias
// corrections of a host-side counter would be frozen into the graph).  The counter is incremented by the caller before the launch.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                float wd, float grad_scale, const int* __restrict__ step) {
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
        const double t = (double)step[0];
        s_bc[0] = (float)(1.0 - pow((double)beta1, t));
        s_bc[1] = (float)(1.0 - pow((double)beta2, t));
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float step_size = lr / s_bc[0];
    const float inv_sqrt_bc2 = rsqrtf(s_bc[1]);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float pi = p[i];
        float gi = g[i] * grad_scale + wd * pi;
        float mi = beta1 * m[i] + (1.f - beta1) * gi;
        float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

// [B,C,T] fp32 -> channels-last planes [B,Ttot,Cpad]; thread = (b, t, 8-channel group).  Reads are strided by T
// (tiny head tensors, L2 resident), writes are 16-byte vectors.
__global__ void ncl_to_nlc_split_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                        int B, int C, int T, int Cpad, int Ttot, int dilate, int offset) {
    const int cgs = Cpad >> 3;
    const long long total = (long long)B * T * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int t = (int)(i % T);
        const long long r = i / T;
        const int cg = (int)(r % cgs);
        const int b = (int)(r / cgs);
        uint32_t h32[4] = {0, 0, 0, 0}, l32[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cg * 8 + j;
            if (c < C) {
                __nv_bfloat16 hb, lb;
                split_bf16(x[((size_t)b * C + c) * T + t], hb, lb);
                h32[j >> 1] |= (uint32_t)__bfloat16_as_ushort(hb) << ((j & 1) * 16);
                l32[j >> 1] |= (uint32_t)__bfloat16_as_ushort(lb) << ((j & 1) * 16);
            }
        }
        const size_t off = ((size_t)b * Ttot + offset + (size_t)t * dilate) * Cpad + cg * 8;
        *reinterpret_cast<uint4*>(hi + off) = make_uint4(h32[0], h32[1], h32[2], h32[3]);
        if (lo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(l32[0], l32[1], l32[2], l32[3]);
    }
}

static inline int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = 148LL * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// Same output as clip_ingest_kernel, straight from the dataset's storage format: uint8 [N,T,Hs,Ws,3] frames
// (AFSD/common/video2npy.py:61-74), cropped to H x W at a per-sample offset, optionally mirrored along W, normalised as
// (x / 255) * 2 - 1 (AFSD/common/thumos_dataset.py:261-263; _rn intrinsics in torch's operation order: bit-identical).
// Replaces the CPU crop / flip / normalise of the data loader (videotransforms.py:30-124) and a 4x larger H2D copy.
// frame_map (optional, [N,T]): output frame t reads source frame frame_map[n*T+t] — the SSL cut-paste augmentation
// (thumos_dataset.py:187-229) is a re-ordering of the clip's own frames, so the augmented clip is never materialised.
__global__ void clip_ingest_u8_kernel(const unsigned char* __restrict__ px, const int* __restrict__ crop,
                                      const int* __restrict__ frame_map, uint16_t* __restrict__ hi,
                                      uint16_t* __restrict__ lo, int N, int T, int Hs, int Ws, int H, int W, int oh_def, int ow_def) {
    const int Wp = W + 8;
    const long long total = (long long)N * T * H * Wp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int wp = (int)(i % Wp);
        long long row = i / Wp;                       // (n*T + t)*H + h
        const int h = (int)(row % H); row /= H;
        const int t = (int)(row % T);
        const int n = (int)(row / T);
        const int w = wp - 2;
        uint32_t h32[2] = {0, 0}, l32[2] = {0, 0};
        if (w >= 0 && w < W) {
            const int oh = crop ? crop[3 * n] : oh_def, ow = crop ? crop[3 * n + 1] : ow_def, flip = crop ? crop[3 * n + 2] : 0;
            const int ws = ow + (flip ? W - 1 - w : w);
            const int ts = frame_map ? min(max(frame_map[n * T + t], 0), T - 1) : t;
            const unsigned char* src = px + ((((size_t)n * T + ts) * Hs + (oh + h)) * Ws + ws) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __fsub_rn(__fmul_rn(__fdiv_rn((float)src[c], 255.f), 2.f), 1.f);
                __nv_bfloat16 hb, lb;
                split_bf16(v, hb, lb);
                h32[c >> 1] |= (uint32_t)__bfloat16_as_ushort(hb) << ((c & 1) * 16);
                l32[c >> 1] |= (uint32_t)__bfloat16_as_ushort(lb) << ((c & 1) * 16);
            }
        }
        reinterpret_cast<uint2*>(hi)[i] = make_uint2(h32[0], h32[1]);
        if (lo) reinterpret_cast<uint2*>(lo)[i] = make_uint2(l32[0], l32[1]);
    }
}

// The raw-pixel form of clip_ingest_u8_kernel: the same crop / mirror / temporal gather, but the plane holds the uint8 pixel
// VALUES (0..255, exact in bf16) instead of the normalised clip; zero outside the image and in channel slot 3.  The
// normalisation (x/255)*2-1 and the reference's zero padding of the normalised clip are folded into the scale / border-class
// shift of otal_conv1a_fwd_u8, so one plane replaces two and the tensor core sees an operand without rounding error.
__global__ void clip_ingest_u8_raw_kernel(const unsigned char* __restrict__ px, const int* __restrict__ crop,
                                          const int* __restrict__ frame_map, uint16_t* __restrict__ out, int N, int T, int Hs, int Ws,
                                          int H, int W, int oh_def, int ow_def) {
    const int Wp = W + 8;
    const long long total = (long long)N * T * H * Wp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int wp = (int)(i % Wp);
        long long row = i / Wp;                       // (n*T + t)*H + h
        const int h = (int)(row % H); row /= H;
        const int t = (int)(row % T);
        const int n = (int)(row / T);
        const int w = wp - 2;
        uint32_t v32[2] = {0, 0};
        if (w >= 0 && w < W) {
            const int oh = crop ? crop[3 * n] : oh_def, ow = crop ? crop[3 * n + 1] : ow_def, flip = crop ? crop[3 * n + 2] : 0;
            const int ws = ow + (flip ? W - 1 - w : w);
            const int ts = frame_map ? min(max(frame_map[n * T + t], 0), T - 1) : t;
            const unsigned char* src = px + ((((size_t)n * T + ts) * Hs + (oh + h)) * Ws + ws) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v32[c >> 1] |= (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)src[c])) << ((c & 1) * 16);
            // channel slot 3 = 1.0 inside the image (0 in the padding): the forward weights of that slot are zero, and in the
            // weight gradient its column is sum_p D[p] * [tap inside the image at p] — the term that turns the raw-pixel
            // gradient into the gradient of the conv on the normalised, zero-padded clip (ops.conv1a_u8_weight_grad), for free
            v32[1] |= 0x3F80u << 16;
        }
        reinterpret_cast<uint2*>(out)[i] = make_uint2(v32[0], v32[1]);
    }
}

// Sums of an NDHWC gradient (hi + lo planes) per border class of the position and per channel: sums[(ct*4+ch)*4+cw][c] +=
// d[n,t,h,w,c], class of an index o along a dim of n outputs = 1 (o == 0), 2 (o == n-2), 3 (o == n-1), else 0 (the classes
// of ConvParams::shift_classes).  From these 64 x C numbers the caller forms, for every tap of the 7x7x7 stride-2 conv, the
// sum of d over the positions where the tap lies inside the image — the term that turns the raw-pixel weight gradient
// (otal_conv1a_wgrad_u8) into the gradient of the conv on the normalised, zero-padded clip.  One pass over d: a block owns a
// contiguous range of positions, interior positions accumulate in registers, border positions in a shared-memory table.
__global__ void border_class_sums_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ sums,
                                         long long npos, int To, int Ho, int Wo, int C, int cstride, int coff, long long chunk) {
    extern __shared__ float bcs_tab[];               // [64][C]
    for (int i = threadIdx.x; i < 64 * C; i += blockDim.x) bcs_tab[i] = 0.f;
    __syncthreads();
    const int tpp = C >> 3;                          // threads per position (8 channels = 16 bytes each)
    const int sub = threadIdx.x % tpp, slot = threadIdx.x / tpp, pslots = blockDim.x / tpp;
    const long long p0 = (long long)blockIdx.x * chunk;
    const long long p1 = p0 + chunk < npos ? p0 + chunk : npos;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    auto cls = [](int o, int n) { return o == 0 ? 1 : (o == n - 2 ? 2 : (o == n - 1 ? 3 : 0)); };
    for (long long pos = p0 + slot; pos < p1; pos += pslots) {
        const int w = (int)(pos % Wo);
        long long r = pos / Wo;
        const int h = (int)(r % Ho); r /= Ho;
        const int t = (int)(r % To);
        const int k = (cls(t, To) * 4 + cls(h, Ho)) * 4 + cls(w, Wo);
        const size_t off = (size_t)pos * cstride + coff + sub * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(hi + off);
        uint4 b = make_uint4(0, 0, 0, 0);
        if (lo) b = *reinterpret_cast<const uint4*>(lo + off);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        float v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[2 * e] = __uint_as_float(aw[e] << 16) + __uint_as_float(bw[e] << 16);
            v[2 * e + 1] = __uint_as_float(aw[e] & 0xffff0000u) + __uint_as_float(bw[e] & 0xffff0000u);
        }
        if (k == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e];
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&bcs_tab[k * C + sub * 8 + e], v[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&bcs_tab[sub * 8 + e], acc[e]);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * C; i += blockDim.x) {
        const float v = bcs_tab[i];
        if (v != 0.f) atomicAdd(&sums[i], v);
    }
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

int otal_clip_ingest(const float* x, uint16_t* hi, uint16_t* lo, int N, int C, int T, int H, int W, void* stream) {
    if (N < 0 || C <= 0 || C > 4 || T <= 0 || H <= 0 || W <= 0 || (W & 1) || !x || !hi) {
        set_last_error_msg("clip_ingest: bad argument (C <= 4, W even)"); return OTAL_ERR_BAD_ARG;
    }
    if (N == 0) return OTAL_OK;
    clip_ingest_kernel<<<grid_for((long long)N * T * H * (W + 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, hi, lo, N, C, T, H, W);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_relu_bn_bwd_split(const float* g, const uint16_t* y_hi, const float* scale, uint16_t* d_hi, uint16_t* d_lo,
                           long long npos, int C, int g_cstride, int g_coff, int y_cstride, int y_coff,
                           int d_cstride, int d_coff, int relu, void* stream) {
    if (npos < 0 || C <= 0 || C % 8 || g_cstride % 4 || g_coff % 4 || y_cstride % 8 || y_coff % 8 || d_cstride % 8 ||
        d_coff % 8 || !g || !d_hi || (relu && !y_hi)) {
        set_last_error_msg("relu_bn_bwd_split: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    if (npos == 0) return OTAL_OK;
    relu_bn_bwd_split_kernel<<<grid_for(npos * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        g, y_hi, scale, d_hi, d_lo, npos, C, g_cstride, g_coff, y_cstride, y_coff, d_cstride, d_coff, relu);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, float grad_scale, int step, void* stream) {
    if (n < 0 || step < 1 || (n > 0 && (!p || !g || !m || !v))) { set_last_error_msg("adam: bad argument"); return OTAL_ERR_BAD_ARG; }
    if (n == 0) return OTAL_OK;
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step)), bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    adam_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                             weight_decay, grad_scale, bc1, bc2);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float grad_scale, const int* step_dev, void* stream) {
    if (n < 0 || !step_dev || (n > 0 && (!p || !g || !m || !v))) { set_last_error_msg("adam_dev: bad argument"); return OTAL_ERR_BAD_ARG; }
    if (n == 0) return OTAL_OK;
    adam_dev_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                                 weight_decay, grad_scale, step_dev);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_ncl_to_nlc_split(const float* x, uint16_t* hi, uint16_t* lo, int B, int C, int T, int Cpad, int Ttot, int dilate,
                          int offset, void* stream) {
    if (B < 0 || C <= 0 || T <= 0 || Cpad < C || Cpad % 8 || dilate < 1 || offset < 0 || offset + (T - 1) * dilate >= Ttot ||
        !x || !hi) {
        set_last_error_msg("ncl_to_nlc_split: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    if (B == 0) return OTAL_OK;
    ncl_to_nlc_split_kernel<<<grid_for((long long)B * T * (Cpad / 8), 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x, hi, lo, B, C, T, Cpad, Ttot, dilate, offset);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

extern "C" int otal_clip_ingest_u8(const unsigned char* px, const int* crop, const int* frame_map, uint16_t* hi, uint16_t* lo,
                                   int N, int T, int Hs, int Ws, int H, int W, void* stream) {
    if (N < 0 || T <= 0 || Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0 || (W & 1) || H > Hs || W > Ws || !px || !hi) {
        otal::set_last_error_msg("clip_ingest_u8: bad argument (W even, crop inside the frame)"); return OTAL_ERR_BAD_ARG;
    }
    if (N == 0) return OTAL_OK;
    otal::clip_ingest_u8_kernel<<<otal::grid_for((long long)N * T * H * (W + 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        px, crop, frame_map, hi, lo, N, T, Hs, Ws, H, W, (Hs - H) / 2, (Ws - W) / 2);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

extern "C" int otal_clip_ingest_u8_raw(const unsigned char* px, const int* crop, const int* frame_map, uint16_t* out, int N, int T,
                                       int Hs, int Ws, int H, int W, void* stream) {
    if (N < 0 || T <= 0 || Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0 || (W & 1) || H > Hs || W > Ws || !px || !out) {
        otal::set_last_error_msg("clip_ingest_u8_raw: bad argument (W even, crop inside the frame)"); return OTAL_ERR_BAD_ARG;
    }
    if (N == 0) return OTAL_OK;
    otal::clip_ingest_u8_raw_kernel<<<otal::grid_for((long long)N * T * H * (W + 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        px, crop, frame_map, out, N, T, Hs, Ws, H, W, (Hs - H) / 2, (Ws - W) / 2);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

extern "C" int otal_border_class_sums(const uint16_t* d_hi, const uint16_t* d_lo, float* sums, int N, int To, int Ho, int Wo, int C,
                                      int d_cstride, int d_coff, void* stream) {
    if (N < 0 || To < 3 || Ho < 3 || Wo < 3 || C < 8 || C > 128 || (C & (C - 1)) || d_cstride % 8 || d_coff % 8 || d_coff + C > d_cstride ||
        !d_hi || !sums) {
        otal::set_last_error_msg("border_class_sums: bad argument (C a power of two in 8..128, extents >= 3, 16-byte aligned slices)");
        return OTAL_ERR_BAD_ARG;
    }
    if (N == 0) return OTAL_OK;
    const long long npos = (long long)N * To * Ho * Wo;
    long long blocks = 4LL * 148;
    const long long pslots = 256 / (C >> 3);
    if (blocks * pslots > npos) blocks = (npos + pslots - 1) / pslots;
    const long long chunk = (npos + blocks - 1) / blocks;
    otal::border_class_sums_kernel<<<(unsigned)blocks, 256, (size_t)64 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        d_hi, d_lo, sums, npos, To, Ho, Wo, C, d_cstride, d_coff, chunk);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

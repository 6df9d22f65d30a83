// Small fixed-function kernels of the detection head.
//
//  * otal_make_segments — proposal window generation for ALL pyramid levels of a batch in one launch
//    (AFSD/thumos14/BDNet.py:355-384; 25 elementwise torch launches per level in the reference).  The arithmetic is
//    written with explicit round-to-nearest intrinsics in the reference's operation order (no FMA contraction), so the
//    windows are bit-identical to torch's: they are rounded to integers and a 1-ulp difference could move a pooling
//    window by one frame (SURVEY "hard part" 2).  torch.round is round-half-to-even = rintf.
//  * otal_dirichlet_uncertainty — DirichletLayer.compute_uncertainty with 'exp' evidence (BDNet.py:544-556):
//    u = K / sum_k (exp(clamp(x_k, -10, 10)) + 1), one warp per prior.
#include "common.cuh"

namespace otal {

__device__ __forceinline__ void window4(float l, float r, float plen, float* out) {
    const float inl = fmaxf(__fmul_rn(plen, 0.25f), 1.f);            // plen / 4.0 is exact as a multiplication
    const float outl = fmaxf(__fdiv_rn(plen, 10.f), 1.f);
    out[0] = rintf(__fsub_rn(l, outl));
    out[1] = rintf(__fadd_rn(l, inl));
    out[2] = rintf(__fsub_rn(r, inl));
    out[3] = rintf(__fadd_rn(r, outl));
}

__global__ void make_segments_kernel(const float* __restrict__ loc, const float* __restrict__ prior, const int* __restrict__ tlen,
                                     const int* __restrict__ coff, float* __restrict__ seg_raw, float* __restrict__ seg_c,
                                     float* __restrict__ fseg, int B, int P, float frame_num, const int* __restrict__ out_row, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * P) return;
    const int p = i % P;
    // where the windows of prior p go: row p of [B,P,4], or row out_row[p] of [B,S,4] (the level-separated layout)
    const size_t o = out_row ? (size_t)(i / P) * S + out_row[p] : (size_t)i;
    const float l0 = loc[2 * (size_t)i], l1 = loc[2 * (size_t)i + 1];
    const float pri = prior[p];
    const int t = tlen[p];
    const float tf = (float)t;
    // level units (BDNet.py:359-371)
    const float s0 = __fmul_rn(__fdiv_rn(l0, frame_num), tf), s1 = __fmul_rn(__fdiv_rn(l1, frame_num), tf);
    const float centre = rintf(__fsub_rn(__fmul_rn(pri, tf), 0.5f));
    float w[4];
    window4(__fsub_rn(centre, s0), __fadd_rn(centre, s1), __fadd_rn(s0, s1), w);
    float* r = seg_raw ? seg_raw + 4 * o : nullptr;
    float* c = seg_c + 4 * o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (r) r[j] = w[j];
        // what BoundaryMaxPooling does with a per-level window (kernel.cu:33-38: truncate, clamp to [0, t-1]), then the
        // offset of the level inside the level-concatenated feature
        int v = __float2int_rz(w[j]);
        v = min(max(v, 0), t - 1);
        c[j] = (float)(v + coff[p]);
    }
    // frame units (BDNet.py:373-384)
    const float pf = __fmul_rn(pri, frame_num);
    const float dl = __fsub_rn(pf, l0), dr = __fadd_rn(pf, l1);
    window4(dl, dr, __fadd_rn(__fsub_rn(dr, dl), 1.f), fseg + 4 * o);
}

__global__ void dirichlet_uncertainty_kernel(const float* __restrict__ logit, float* __restrict__ unct, int M, int K) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= M) return;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += expf(fminf(fmaxf(logit[(size_t)warp * K + k], -10.f), 10.f)) + 1.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) unct[warp] = (float)K / s;
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_make_segments(const float* loc, const float* prior, const int* level_len, const int* level_off, float* seg_level,
                       float* seg_concat, float* frame_seg, int B, int P, float frame_num, void* stream) {
    if (B <= 0 || P <= 0 || !loc || !prior || !level_len || !level_off || !seg_concat || !frame_seg || !(frame_num > 0.f)) {
        set_last_error_msg("make_segments: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    const int total = B * P;
    make_segments_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(loc, prior, level_len, level_off, seg_level,
                                                                                             seg_concat, frame_seg, B, P, frame_num, nullptr, 0);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_make_segments_ex(const float* loc, const float* prior, const int* level_len, const int* level_off, const int* out_row, int S,
                          float* seg_concat, float* frame_seg, int B, int P, float frame_num, void* stream) {
    if (B <= 0 || P <= 0 || S < P || !loc || !prior || !level_len || !level_off || !out_row || !seg_concat || !frame_seg || !(frame_num > 0.f)) {
        set_last_error_msg("make_segments_ex: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    const int total = B * P;
    make_segments_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(loc, prior, level_len, level_off, nullptr,
                                                                                             seg_concat, frame_seg, B, P, frame_num, out_row, S);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_dirichlet_uncertainty(const float* logit, float* unct, long long M, int K, void* stream) {
    if (M < 0 || K <= 0 || (M > 0 && (!logit || !unct))) { set_last_error_msg("dirichlet_uncertainty: bad argument"); return OTAL_ERR_BAD_ARG; }
    if (M == 0) return OTAL_OK;
    const long long threads = M * 32;
    dirichlet_uncertainty_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(logit, unct, (int)M, K);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Boundary BCE of the training script — calc_bce_loss (AFSD/thumos14/train.py:152-161):
//     s[r] = mean_c tanh(x[r, c]);   loss = mean_r BCE(s[r], t[r])          (F.binary_cross_entropy: logs clamped at -100)
// for x = start / end maps [B, T, C] (rows r = (b, t), C contiguous) and t = the start / end score map of the clip.
// One warp per row: coalesced reads, shuffle reduction.  The forward also stores coef[r] = dBCE/ds / (R * C) so that the
// backward is one elementwise pass dx = g * coef[r] * (1 - tanh(x)^2).
// ------------------------------------------------------------------------------------------------------------------
namespace otal {

__global__ void boundary_bce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ target, long long t_bstride, int T,
                                        float* __restrict__ row_loss, float* __restrict__ coef, int R, int C) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= R) return;
    const float* xr = x + (size_t)row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += tanhf(xr[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        s /= (float)C;
        const int b = row / T, t = row - b * T;
        const float y = target[(size_t)b * t_bstride + t];
        const float l1 = fmaxf(logf(s), -100.f), l0 = fmaxf(logf(1.f - s), -100.f);
        row_loss[row] = -(y * l1 + (1.f - y) * l0);
        // torch's binary_cross_entropy backward: (s - y) / max(s (1 - s), 1e-12)
        coef[row] = (s - y) / fmaxf(s * (1.f - s), 1e-12f) / ((float)R * (float)C);
    }
}

__global__ void boundary_bce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ coef, const float* __restrict__ g,
                                        float* __restrict__ gx, long long n, int C) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float gv = g[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float th = tanhf(x[i]);
        gx[i] = gv * coef[i / C] * (1.f - th * th);
    }
}

}  // namespace otal

extern "C" int otal_boundary_bce_fwd(const float* x, const float* target, long long target_batch_stride, float* row_loss,
                                     float* coef, int B, int T, int C, void* stream) {
    if (B <= 0 || T <= 0 || C <= 0 || !x || !target || !row_loss || !coef) { otal::set_last_error_msg("boundary_bce: bad argument"); return OTAL_ERR_BAD_ARG; }
    const long long threads = (long long)B * T * 32;
    otal::boundary_bce_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, target, target_batch_stride, T, row_loss, coef, B * T, C);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

extern "C" int otal_boundary_bce_bwd(const float* x, const float* coef, const float* grad_loss, float* grad_x, int B, int T, int C,
                                     void* stream) {
    if (B <= 0 || T <= 0 || C <= 0 || !x || !coef || !grad_loss || !grad_x) { otal::set_last_error_msg("boundary_bce: bad argument"); return OTAL_ERR_BAD_ARG; }
    const long long n = (long long)B * T * C;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    otal::boundary_bce_bwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, coef, grad_loss, grad_x, n, C);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Glue of the explicit head schedule (opental_b200/head_schedule.py): what CoarsePyramid.forward does with F.interpolate, +,
// torch.cat, index_select and permute between its convolutions (AFSD/thumos14/BDNet.py:311-331, :340-353, :399-412;
// AFSD/anet/BDNet.py:281-311), and the transposes of those steps in the backward — as three small kernels over the [B,C,T]
// fp32 tensors the convolutions write.
//
//  * otal_rows_combine: dst[b,c,j] = sum_k src[table[j][k].src][b,c,table[j][k].col] — nearest-neighbour upsampling, the
//    top-down addition, the level-separated ("sep") layout of the 6 pyramid levels and, with the transposed table, their
//    gradients.  The result is written as fp32 [B,C,Td] and / or as channels-last bf16 planes (the conv operand layout).
//  * otal_head_gather_fwd / _bwd: the head convolutions' raw outputs [B,Cpad,S] (sep layout, channels padded to 8) -> the
//    reference's [B,P,Cout] tensors, with ScaleExp (exp(x * scale_level), x FPN stride for ActivityNet) on the loc head; the
//    backward writes the raw outputs' gradient as zero-padded channels-last planes and accumulates d scale_level.
// ------------------------------------------------------------------------------------------------------------------
namespace otal {

struct RowsParams {
    int B, C, Td, npairs, nsrc;
    const float* src[8];
    int src_T[8];
    const int* table;
    float* dst;
    uint16_t *p_hi, *p_lo;
};

__device__ __forceinline__ float rows_value(const RowsParams& p, int b, int c, int j) {
    float v = 0.f;
    const int* e = p.table + (size_t)j * p.npairs * 2;
    for (int k = 0; k < p.npairs; ++k) {
        const int s = e[2 * k];
        if (s >= 0) v += p.src[s][((size_t)b * p.C + c) * p.src_T[s] + e[2 * k + 1]];
    }
    return v;
}

__global__ void rows_combine_kernel(const RowsParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p.dst) {
        const long long total = (long long)p.B * p.C * p.Td;
        for (long long i = i0; i < total; i += stride) {
            const int j = (int)(i % p.Td);
            const long long r = i / p.Td;
            p.dst[i] = rows_value(p, (int)(r / p.C), (int)(r % p.C), j);
        }
    }
    if (p.p_hi) {
        const int half = p.C >> 1;
        const long long total = (long long)p.B * p.Td * half;
        for (long long i = i0; i < total; i += stride) {
            const int cp = (int)(i % half);
            const long long r = i / half;
            const int j = (int)(r % p.Td), b = (int)(r / p.Td);
            uint32_t h, l;
            split_bf16x2(rows_value(p, b, 2 * cp, j), rows_value(p, b, 2 * cp + 1, j), h, l);
            const size_t o = ((size_t)b * p.Td + j) * p.C + 2 * cp;
            *reinterpret_cast<uint32_t*>(p.p_hi + o) = h;
            if (p.p_lo) *reinterpret_cast<uint32_t*>(p.p_lo + o) = l;
        }
    }
}

struct HeadGatherParams {
    int B, S, P, n;
    const int* sep_idx;       // [P] column of prior p in the sep layout
    const int* level_id;      // [P]
    const float* mult;        // [P] or null (ActivityNet: FPN stride of the prior's level)
    const float* scale[8];    // ScaleExp parameter of each level
    float* dscale[8];         // their gradients (accumulated) — backward only
    const float* raw[4];      // [B,cpad,S]
    const float* bias[4];     // [cout] or null: added to raw (the conv ran without it: padded output channels)
    float* dbias[4];          // backward: accumulated, or null
    int cpad[4], cout[4], mode[4];
    float* out[4];            // forward: [B,P,cout]
    const float* gout[4];     // backward: [B,P,cout] or null
    const float* outv[4];     // backward, mode 1: the forward's output values
    uint16_t *d_hi[4], *d_lo[4];   // backward: channels-last planes [B,S,cpad]
};

__global__ void head_gather_fwd_kernel(const HeadGatherParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int k = 0; k < p.n; ++k) {
        const int co = p.cout[k];
        const long long total = (long long)p.B * p.P * co;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c = (int)(i % co);
            const long long r = i / co;
            const int pr = (int)(r % p.P), b = (int)(r / p.P);
            float v = p.raw[k][((size_t)b * p.cpad[k] + c) * p.S + p.sep_idx[pr]];
            if (p.bias[k]) v += p.bias[k][c];
            if (p.mode[k] == 1) {
                v = expf(v * p.scale[p.level_id[pr]][0]);
                if (p.mult) v *= p.mult[pr];
            }
            p.out[k][i] = v;
        }
    }
}

// thread = (head, b, sep column): zero rows for separators; mode 1: d raw = g * out * scale_level, d scale_level += g * out * raw
__global__ void head_gather_bwd_kernel(const HeadGatherParams p, const int* __restrict__ prior_of_col) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int k = 0; k < p.n; ++k) {
        const int cp = p.cpad[k], co = p.cout[k];
        const long long total = (long long)p.B * p.S;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
            const int s = (int)(i % p.S), b = (int)(i / p.S);
            const int pr = prior_of_col[s];
            uint16_t* hi = p.d_hi[k] + (size_t)i * cp;
            uint16_t* lo = p.d_lo[k] ? p.d_lo[k] + (size_t)i * cp : nullptr;
            float acc_scale = 0.f;
            for (int c = 0; c < cp; ++c) {
                float g = 0.f;
                if (pr >= 0 && c < co && p.gout[k]) {
                    g = p.gout[k][((size_t)b * p.P + pr) * co + c];
                    if (p.mode[k] == 1) {
                        const float o = p.outv[k][((size_t)b * p.P + pr) * co + c];
                        const float x = p.raw[k][((size_t)b * cp + c) * p.S + s] + (p.bias[k] ? p.bias[k][c] : 0.f);
                        acc_scale += g * o * x;
                        g = g * o * p.scale[p.level_id[pr]][0];
                    }
                    if (p.dbias[k] && g != 0.f) atomicAdd(p.dbias[k] + c, g);
                }
                __nv_bfloat16 hb, lb;
                split_bf16(g, hb, lb);
                hi[c] = __bfloat16_as_ushort(hb);
                if (lo) lo[c] = __bfloat16_as_ushort(lb);
            }
            if (p.mode[k] == 1 && pr >= 0 && acc_scale != 0.f) atomicAdd(p.dscale[p.level_id[pr]], acc_scale);
        }
    }
}

// [B,C,T] fp32 (sample stride x_bstride) -> channels-last planes [B,T,cstride] at channel offset coff; thread = (b, t, 8 channels)
__global__ void ncl_to_nlc_split_ex_kernel(const float* __restrict__ x, long long x_bstride, uint16_t* __restrict__ hi,
                                           uint16_t* __restrict__ lo, int B, int C, int T, int cstride, int coff) {
    const int cgs = C >> 3;
    const long long total = (long long)B * T * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int t = (int)(i % T);
        const long long r = i / T;
        const int cg = (int)(r % cgs), b = (int)(r / cgs);
        uint32_t h32[4], l32[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* px = x + (size_t)b * x_bstride + (size_t)(cg * 8 + 2 * j) * T + t;
            split_bf16x2(px[0], px[T], h32[j], l32[j]);
        }
        const size_t off = ((size_t)b * T + t) * cstride + coff + cg * 8;
        *reinterpret_cast<uint4*>(hi + off) = make_uint4(h32[0], h32[1], h32[2], h32[3]);
        if (lo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(l32[0], l32[1], l32[2], l32[3]);
    }
}

__global__ void boundary_bce_fwd_ex_kernel(const float* __restrict__ x, int x_rstride, const float* __restrict__ target,
                                           long long t_bstride, int T, float* __restrict__ row_loss, float* __restrict__ coef, int R,
                                           int C) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= R) return;
    const float* xr = x + (size_t)row * x_rstride;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += tanhf(xr[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        s /= (float)C;
        const int b = row / T, t = row - b * T;
        const float y = target[(size_t)b * t_bstride + t];
        const float l1 = fmaxf(logf(s), -100.f), l0 = fmaxf(logf(1.f - s), -100.f);
        row_loss[row] = -(y * l1 + (1.f - y) * l0);
        coef[row] = (s - y) / fmaxf(s * (1.f - s), 1e-12f) / ((float)R * (float)C);
    }
}

__global__ void boundary_bce_bwd_ex_kernel(const float* __restrict__ x, int x_rstride, const float* __restrict__ coef,
                                           const float* __restrict__ g, float* __restrict__ gx, long long n, int C) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float gv = g[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long row = i / C;
        const float th = tanhf(x[row * x_rstride + (i - row * C)]);
        gx[i] = gv * coef[row] * (1.f - th * th);
    }
}

static inline int glue_grid(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = 148LL * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace otal

extern "C" int otal_rows_combine(const otal_rows_desc* d, void* stream) {
    using namespace otal;
    if (!d || d->B <= 0 || d->C <= 0 || d->Td <= 0 || d->npairs <= 0 || d->npairs > 16 || d->nsrc <= 0 || d->nsrc > 8 || !d->table ||
        (!d->dst && !d->p_hi) || (d->p_hi && (d->C & 1))) {
        set_last_error_msg("rows_combine: bad argument (1..8 sources, 1..16 pairs per column, even C for planes)"); return OTAL_ERR_BAD_ARG;
    }
    RowsParams p{};
    p.B = d->B; p.C = d->C; p.Td = d->Td; p.npairs = d->npairs; p.nsrc = d->nsrc; p.table = d->table; p.dst = d->dst;
    p.p_hi = d->p_hi; p.p_lo = d->p_lo;
    for (int k = 0; k < d->nsrc; ++k) {
        if (!d->src[k] || d->src_T[k] <= 0) { set_last_error_msg("rows_combine: null source"); return OTAL_ERR_BAD_ARG; }
        p.src[k] = d->src[k]; p.src_T[k] = d->src_T[k];
    }
    rows_combine_kernel<<<glue_grid((long long)d->B * d->C * d->Td, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

static int head_gather_params(const otal_headout_desc* d, otal::HeadGatherParams& p, bool bwd) {
    using namespace otal;
    if (!d || d->B <= 0 || d->S <= 0 || d->P <= 0 || d->n <= 0 || d->n > 4 || !d->sep_idx) {
        set_last_error_msg("head_gather: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    p.B = d->B; p.S = d->S; p.P = d->P; p.n = d->n; p.sep_idx = d->sep_idx; p.level_id = d->level_id; p.mult = d->mult;
    for (int l = 0; l < 8; ++l) { p.scale[l] = d->scale[l]; p.dscale[l] = d->dscale[l]; }
    for (int k = 0; k < d->n; ++k) {
        if (!d->raw[k] || d->cout[k] <= 0 || d->cpad[k] < d->cout[k] || (d->mode[k] == 1 && (!d->level_id || !d->scale[0])) ||
            (!bwd && !d->out[k]) || (bwd && (!d->d_hi[k] || (d->mode[k] == 1 && (!d->out[k] || !d->dscale[0]))))) {
            set_last_error_msg("head_gather: incomplete head entry"); return OTAL_ERR_BAD_ARG;
        }
        p.raw[k] = d->raw[k]; p.cpad[k] = d->cpad[k]; p.cout[k] = d->cout[k]; p.mode[k] = d->mode[k];
        p.bias[k] = d->bias[k]; p.dbias[k] = bwd ? d->dbias[k] : nullptr;
        p.out[k] = d->out[k]; p.outv[k] = d->out[k]; p.gout[k] = d->gout[k]; p.d_hi[k] = d->d_hi[k]; p.d_lo[k] = d->d_lo[k];
    }
    return OTAL_OK;
}

extern "C" int otal_head_gather_fwd(const otal_headout_desc* d, void* stream) {
    otal::HeadGatherParams p{};
    int rc = head_gather_params(d, p, false);
    if (rc) return rc;
    otal::head_gather_fwd_kernel<<<otal::glue_grid((long long)d->B * d->P * 16, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

extern "C" int otal_head_gather_bwd(const otal_headout_desc* d, const int* prior_of_col, void* stream) {
    otal::HeadGatherParams p{};
    int rc = head_gather_params(d, p, true);
    if (rc) return rc;
    if (!prior_of_col) { otal::set_last_error_msg("head_gather_bwd: null column table"); return OTAL_ERR_BAD_ARG; }
    otal::head_gather_bwd_kernel<<<otal::glue_grid((long long)d->B * d->S, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p, prior_of_col);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

extern "C" int otal_ncl_to_nlc_split_ex(const float* x, long long x_bstride, uint16_t* hi, uint16_t* lo, int B, int C, int T,
                                        int cstride, int coff, void* stream) {
    if (B <= 0 || C <= 0 || C % 8 || T <= 0 || cstride % 8 || coff % 8 || coff + C > cstride || !x || !hi) {
        otal::set_last_error_msg("ncl_to_nlc_split_ex: bad argument (C, cstride, coff multiples of 8)"); return OTAL_ERR_BAD_ARG;
    }
    otal::ncl_to_nlc_split_ex_kernel<<<otal::glue_grid((long long)B * T * (C / 8), 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x, x_bstride > 0 ? x_bstride : (long long)C * T, hi, lo, B, C, T, cstride, coff);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

extern "C" int otal_boundary_bce_fwd_ex(const float* x, int x_rstride, const float* target, long long target_batch_stride,
                                        float* row_loss, float* coef, int B, int T, int C, void* stream) {
    if (B <= 0 || T <= 0 || C <= 0 || x_rstride < C || !x || !target || !row_loss || !coef) { otal::set_last_error_msg("boundary_bce_ex: bad argument"); return OTAL_ERR_BAD_ARG; }
    const long long threads = (long long)B * T * 32;
    otal::boundary_bce_fwd_ex_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, x_rstride, target, target_batch_stride, T, row_loss, coef, B * T, C);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

extern "C" int otal_boundary_bce_bwd_ex(const float* x, int x_rstride, const float* coef, const float* grad_loss, float* grad_x, int B,
                                        int T, int C, void* stream) {
    if (B <= 0 || T <= 0 || C <= 0 || x_rstride < C || !x || !coef || !grad_loss || !grad_x) { otal::set_last_error_msg("boundary_bce_ex: bad argument"); return OTAL_ERR_BAD_ARG; }
    const long long n = (long long)B * T * C;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    otal::boundary_bce_bwd_ex_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_rstride, coef, grad_loss, grad_x, n, C);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}


// out[s] = mean of x[off[s] .. off[s+1]) for nseg <= 16 segments, one CTA per segment, fixed summation order (deterministic): the
// six boundary-BCE means of one training step in one launch (calc_bce_loss, AFSD/thumos14/train.py:152-161, x 6: :186-200)
namespace otal {
struct SegMeanParams { int n; long long off[17]; };
__global__ void segment_mean_kernel(const float* __restrict__ x, float* __restrict__ out, const SegMeanParams p) {
    __shared__ float red[32];
    const int s = blockIdx.x;
    const long long lo = p.off[s], hi = p.off[s + 1];
    float a = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) a += x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) out[s] = hi > lo ? t / (float)(hi - lo) : 0.f;
    }
}
}  // namespace otal

extern "C" int otal_segment_mean(const float* x, const long long* offsets_host, int nseg, float* out, void* stream) {
    if (!x || !offsets_host || !out || nseg <= 0 || nseg > 16) { otal::set_last_error_msg("segment_mean: bad argument (1..16 segments)"); return OTAL_ERR_BAD_ARG; }
    otal::SegMeanParams p{};
    p.n = nseg;
    for (int i = 0; i <= nseg; ++i) p.off[i] = offsets_host[i];
    otal::segment_mean_kernel<<<nseg, 512, 0, static_cast<cudaStream_t>(stream)>>>(x, out, p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

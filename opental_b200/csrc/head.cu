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
                                     float* __restrict__ fseg, int B, int P, float frame_num) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * P) return;
    const int p = i % P;
    const float l0 = loc[2 * (size_t)i], l1 = loc[2 * (size_t)i + 1];
    const float pri = prior[p];
    const int t = tlen[p];
    const float tf = (float)t;
    // level units (BDNet.py:359-371)
    const float s0 = __fmul_rn(__fdiv_rn(l0, frame_num), tf), s1 = __fmul_rn(__fdiv_rn(l1, frame_num), tf);
    const float centre = rintf(__fsub_rn(__fmul_rn(pri, tf), 0.5f));
    float w[4];
    window4(__fsub_rn(centre, s0), __fadd_rn(centre, s1), __fadd_rn(s0, s1), w);
    float* r = seg_raw ? seg_raw + 4 * (size_t)i : nullptr;
    float* c = seg_c + 4 * (size_t)i;
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
    window4(dl, dr, __fadd_rn(__fsub_rn(dr, dl), 1.f), fseg + 4 * (size_t)i);
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
                                                                                             seg_concat, frame_seg, B, P, frame_num);
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

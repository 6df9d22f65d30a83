// Inference post-processing on the device.
//
//  * otal_decode_scores — decode_predictions of the test scripts (AFSD/thumos14/test.py:112-140) for a whole batch of
//    sliding-window clips: refined segments in seconds, per-class detection scores
//    (Dirichlet mean probability, averaged over the coarse and refined heads) x sigmoid(centre-ness) x actionness,
//    the averaged Dirichlet uncertainty and the averaged actionness.  One warp per prior.
//  * otal_softnms — Gaussian soft-NMS softnms_v2 (AFSD/common/segment_utils.py:128-162), which the reference runs as a
//    Python while-loop on the CPU after a device->host copy: one CTA per class keeps the scores of all candidates in
//    shared memory and repeats {block arg-max over the undone candidates, decay the overlapping ones}.  Semantics kept
//    bit for bit in structure: first index wins ties, candidates below the score threshold drop out, at most top_k are
//    kept, and the loop stops when ONE undone candidate is left (that candidate is never kept — reference behaviour).
#include "common.cuh"

namespace otal {

__device__ __forceinline__ float pp_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void decode_scores_kernel(const float* __restrict__ loc, const float* __restrict__ ploc, const float* __restrict__ conf,
                                     const float* __restrict__ pconf, const float* __restrict__ center, const float* __restrict__ act,
                                     const float* __restrict__ pact, const float* __restrict__ prior, const float* __restrict__ offset,
                                     float* __restrict__ seg, float* __restrict__ scores, float* __restrict__ unct,
                                     float* __restrict__ actn, int B, int P, int K, float clip, float fps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * P) return;
    const int b = warp / P, p = warp - b * P;
    const float* z1 = conf + (size_t)warp * K;
    const float* z2 = pconf + (size_t)warp * K;
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < K; k += 32) {
        s1 += expf(fminf(fmaxf(z1[k], -10.f), 10.f)) + 1.f;
        s2 += expf(fminf(fmaxf(z2[k], -10.f), 10.f)) + 1.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    float a = 1.f;
    if (act) a = (pp_sigmoid(act[warp]) + pp_sigmoid(pact[warp])) * 0.5f;
    const float c = pp_sigmoid(center[warp]);
    for (int k = lane; k < K; k += 32) {
        const float p1 = (expf(fminf(fmaxf(z1[k], -10.f), 10.f)) + 1.f) / s1;
        const float p2 = (expf(fminf(fmaxf(z2[k], -10.f), 10.f)) + 1.f) / s2;
        float v = (p1 + p2) * 0.5f * c;
        if (act) v *= a;
        scores[((size_t)b * K + k) * P + p] = v;
    }
    if (lane == 0) {
        const float l0 = loc[2 * (size_t)warp], l1 = loc[2 * (size_t)warp + 1];
        const float hw = 0.5f * (l0 + l1);
        const float r0 = hw * ploc[2 * (size_t)warp] + l0, r1 = hw * ploc[2 * (size_t)warp + 1] + l1;
        const float pc = prior[p] * clip;
        const float off = offset ? offset[b] : 0.f;
        seg[2 * (size_t)warp] = (fminf(fmaxf(pc - r0, 0.f), clip) + off) / fps;
        seg[2 * (size_t)warp + 1] = (fminf(fmaxf(pc + r1, 0.f), clip) + off) / fps;
        unct[warp] = ((float)K / s1 + (float)K / s2) * 0.5f;
        actn[warp] = a;
    }
}

constexpr int kNmsThreads = 1024;

__global__ void __launch_bounds__(kNmsThreads, 1)
softnms_kernel(const float* __restrict__ seg, long long seg_class_stride, float* __restrict__ scores, unsigned char* __restrict__ keep,
               int* __restrict__ count, int M, float sigma, int top_k, float thr) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    float* sc = reinterpret_cast<float*>(nms_smem);                    // [M]
    unsigned char* state = nms_smem + (size_t)M * 4;                    // [M]: 0 undone, 1 done, 2 dropped
    __shared__ float red_v[32];
    __shared__ int red_i[32], red_n[32];
    __shared__ float top_s, top_e;
    __shared__ int top_idx, n_undone;
    const int cls = blockIdx.x;
    const float* sg = seg + (size_t)cls * seg_class_stride;
    float* gs = scores + (size_t)cls * M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < M; j += kNmsThreads) {
        const float v = gs[j];
        sc[j] = v;
        state[j] = v >= thr ? 0 : 2;
    }
    __syncthreads();
    int done = 0;
    while (true) {
        // arg-max (lowest index among equal scores) and count over the undone candidates
        float bv = -INFINITY; int bi = 0x7fffffff, n = 0;
        for (int j = tid; j < M; j += kNmsThreads)
            if (state[j] == 0) { ++n; if (sc[j] > bv) { bv = sc[j]; bi = j; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; red_n[warp] = n; }
        __syncthreads();
        if (warp == 0) {
            bv = red_v[lane]; bi = red_i[lane]; n = red_n[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                n += __shfl_xor_sync(0xffffffffu, n, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                n_undone = n; top_idx = bi;
                if (n > 1 && done < top_k) { top_s = sg[2 * (size_t)bi]; top_e = sg[2 * (size_t)bi + 1]; state[bi] = 1; }
            }
        }
        __syncthreads();
        if (!(n_undone > 1 && done < top_k)) break;       // `done` is uniform: every thread counts the same iterations
        ++done;
        const float ts = top_s, te = top_e;
        const float width = fmaxf(te - ts, 1e-5f);
        for (int j = tid; j < M; j += kNmsThreads) {
            if (state[j] != 0) continue;
            const float s = sg[2 * (size_t)j], e = sg[2 * (size_t)j + 1];
            const float inter = fmaxf(fminf(e, te) - fmaxf(s, ts), 0.f);
            const float iou = inter / (width + (e - s) - inter);
            const float v = sc[j] * expf(__fdiv_rn(-(iou * iou), sigma));
            sc[j] = v;
            if (v < thr) state[j] = 2;
        }
        __syncthreads();
    }
    for (int j = tid; j < M; j += kNmsThreads) {
        gs[j] = sc[j];
        keep[(size_t)cls * M + j] = state[j] == 1 ? 1 : 0;
    }
    if (tid == 0) count[cls] = done;
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_decode_scores(const float* loc, const float* prop_loc, const float* conf, const float* prop_conf, const float* center,
                       const float* act, const float* prop_act, const float* prior, const float* offset, float* segments,
                       float* scores, float* uncertainty, float* actionness, int B, int P, int K, float clip_length,
                       float sample_fps, void* stream) {
    if (B <= 0 || P <= 0 || K <= 0 || !loc || !prop_loc || !conf || !prop_conf || !center || !prior || !segments || !scores ||
        !uncertainty || !actionness || (act != nullptr) != (prop_act != nullptr) || !(sample_fps > 0.f)) {
        set_last_error_msg("decode_scores: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    const long long threads = (long long)B * P * 32;
    decode_scores_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        loc, prop_loc, conf, prop_conf, center, act, prop_act, prior, offset, segments, scores, uncertainty, actionness, B, P, K,
        clip_length, sample_fps);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_softnms(const float* segments, long long seg_class_stride, float* scores, unsigned char* keep, int* count, int C, int M,
                 float sigma, int top_k, float score_threshold, void* stream) {
    if (C <= 0 || M <= 0 || !segments || !scores || !keep || !count || !(sigma > 0.f) || top_k < 0 || seg_class_stride < 0) {
        set_last_error_msg("softnms: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    const size_t smem = (size_t)M * 5 + 16;
    if (smem > 200 * 1024) { set_last_error_msg("softnms: more than ~40000 candidates per class do not fit shared memory"); return OTAL_ERR_UNSUPPORTED; }
    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(softnms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        once.mark(once_dev);
    }
    softnms_kernel<<<C, kNmsThreads, smem, static_cast<cudaStream_t>(stream)>>>(segments, seg_class_stride, scores, keep, count, M, sigma,
                                                                              top_k, score_threshold);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

// MultiSegmentLoss (THUMOS14 OpenTAL flavour) as ONE single-CTA kernel: prior<->GT matching, GIoU / L1 / IoU-quality
// BCE regression terms, the evidential (Dirichlet) classification loss with IBM re-weighting and IoU-aware calibration,
// and the positive-unlabeled actionness loss — forward values AND the gradient of every term w.r.t. every head output.
//
// Reference semantics (file:line relative to the OpenTAL repository):
//   AFSD/thumos14/multisegment_loss.py:92-259   MultiSegmentLoss.forward (matching :120-153, terms :155-241, norm :243-256)
//   AFSD/thumos14/multisegment_loss.py:20-53    iou_loss ('calc iou', 'giou'), eps = fp32 machine epsilon
//   AFSD/thumos14/cls_loss.py:120-129           EvidenceLoss.iou_calib
//   AFSD/thumos14/cls_loss.py:132-168,212-278   EvidenceLoss.forward / edl_loss (loss_type 'log', evidence 'exp', with_ibm)
//   AFSD/thumos14/cls_loss.py:299-339           ActionnessLoss.forward (top-M lowest-scoring negatives, rank term)
//
// Why one CTA: the whole problem is B x 126 priors (1008 elements at batch 8, 18.6 KB of inputs per clip)
// but every normaliser is batch-global (N, PN, AN, the 50 IBM bins, the top-M rank of every negative).  The reference
// spends ~300 launches and ~100 host syncs on it; here the elements are staged in shared memory once and every
// global quantity is a warp-shuffle + shared-memory block reduction in a fixed order (deterministic).  The kernel is
// latency-bound by construction: its roofline is the launch, not HBM.
//
// Gradients: the kernel writes "unit" gradients (d loss_i / d input for each of the 7 returned losses) into a small
// workspace; otal_msl_backward combines them with the 7 upstream scalars.  Sub-gradient conventions follow torch:
// min/max ties split the gradient in half, clamp passes the gradient on the closed interval, |x| has gradient 0 at 0.
#include "common.cuh"

namespace otal {

constexpr int kMslThreads = 1024;
constexpr int kMslMaxBins = 256;

struct MslParams {
    int B, P, K, G, M;
    float clip, thresh;
    int use_ibm, num_bins;
    float momentum;
    int iou_aware;
    float act_weight, act_margin;
    int prior_stride;
    const float *loc, *conf, *ploc, *pconf, *center, *act, *pact, *priors, *targets;
    const unsigned char* valid;
    float* weight_accum;
    float* losses;   // [16]: 7 losses, then N, PN, AN, PAN, loss_iouc
    float* ws;       // unit gradients, see offsets below
};

// workspace layout (floats): [d_loc_l 2M][d_loc_ct 2M][d_ploc_l 2M][d_ploc_ct 2M][d_center M][d_act M][d_pact M][d_conf KM][d_pconf KM]
__host__ __device__ inline size_t msl_ws_floats(int M, int K) { return (size_t)M * (11 + 2 * (size_t)K); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block, result broadcast to every thread.  Fixed reduction order -> deterministic.
__device__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect `red` from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.f;
    t = warp_sum(t);
    return t;
}

// max with the lowest index among equal maxima; (value, index) broadcast to every thread
__device__ void block_argmax(float v, int idx, float* red, int* redi, float& out_v, int& out_i) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[warp] = v; redi[warp] = idx; }
    __syncthreads();
    float tv = (lane < (int)(blockDim.x >> 5)) ? red[lane] : -INFINITY;
    int ti = (lane < (int)(blockDim.x >> 5)) ? redi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, tv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ti, o);
        if (ov > tv || (ov == tv && oi < ti)) { tv = ov; ti = oi; }
    }
    out_v = tv; out_i = ti;
}

// 1-D IoU of (left, right) offset pairs and its gradient w.r.t. the prediction (multisegment_loss.py:20-36)
struct IouGrad { float iou, d0, d1, I, U, dI0, dI1; };
__device__ __forceinline__ IouGrad iou_with_grad(float p0, float p1, float t0, float t1) {
    const float eps = 1.1920928955078125e-07f;
    IouGrad r;
    r.dI0 = p0 < t0 ? 1.f : (p0 == t0 ? 0.5f : 0.f);
    r.dI1 = p1 < t1 ? 1.f : (p1 == t1 ? 0.5f : 0.f);
    r.I = fminf(p0, t0) + fminf(p1, t1);
    r.U = (t0 + t1) + (p0 + p1) - r.I;
    const float Uc = fmaxf(r.U, eps);
    const float pass = r.U >= eps ? 1.f : 0.f;
    r.iou = r.I / Uc;
    r.d0 = r.dI0 / Uc - r.I * pass * (1.f - r.dI0) / (Uc * Uc);
    r.d1 = r.dI1 / Uc - r.I * pass * (1.f - r.dI1) / (Uc * Uc);
    return r;
}

__device__ __forceinline__ float bce_logits(float x, float t) {
    return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// EDL 'log' loss of one sample (cls_loss.py:212-216 with func = log, evidence = exp(clamp(.,-10,10))):
// per = log S - log alpha_y; also the IBM statistics grad_norm = |1/alpha_y - K/S|, feat_norm = sum |z| (cls_loss.py:257-262)
__device__ __forceinline__ void edl_terms(const float* z, int K, int y, float& per, float& S, float& alpha_y, float& gnorm, float& fnorm) {
    S = 0.f; fnorm = 0.f; alpha_y = 1.f;
    for (int k = 0; k < K; ++k) {
        const float zk = z[k];
        const float a = expf(fminf(fmaxf(zk, -10.f), 10.f)) + 1.f;
        S += a; fnorm += fabsf(zk);
        if (k == y) alpha_y = a;
    }
    per = logf(S) - logf(alpha_y);
    gnorm = fabsf(1.f / alpha_y - (float)K / S);
}

__global__ void __launch_bounds__(kMslThreads, 1)
msl_forward_kernel(const MslParams p) {
    extern __shared__ __align__(16) unsigned char msl_smem[];
    const int M = p.M, K = p.K, nb = p.num_bins;
    float* s_iou = reinterpret_cast<float*>(msl_smem);       // [M] IoU(loc, loc_t) (no_grad)
    int* s_meta = reinterpret_cast<int*>(s_iou + M);         // [M] pos | ppos << 1 | label << 2
    float* s_gh_c = reinterpret_cast<float*>(s_meta + M);    // [M] IBM grad_hat, coarse
    int* s_bin_c = reinterpret_cast<int*>(s_gh_c + M);       // [M]
    float* s_per_c = reinterpret_cast<float*>(s_bin_c + M);  // [M] unweighted EDL loss, coarse
    float* s_gh_p = s_per_c + M;
    int* s_bin_p = reinterpret_cast<int*>(s_gh_p + M);
    float* s_per_p = reinterpret_cast<float*>(s_bin_p + M);
    float* s_x = s_per_p + M;                                // [M] actionness logits
    float* s_xp = s_x + M;                                   // [M] refined actionness logits
    float* s_unc = s_xp + M;                                 // [M] Dirichlet uncertainty of prop_conf
    float* s_acc = s_unc + M;                                // [nb] IBM EMA
    float* s_red = s_acc + kMslMaxBins;                      // [32]
    int* s_redi = reinterpret_cast<int*>(s_red + 32);        // [32]

    float* d_loc_l = p.ws;
    float* d_loc_ct = d_loc_l + 2 * (size_t)M;
    float* d_ploc_l = d_loc_ct + 2 * (size_t)M;
    float* d_ploc_ct = d_ploc_l + 2 * (size_t)M;
    float* d_center = d_ploc_ct + 2 * (size_t)M;
    float* d_act = d_center + M;
    float* d_pact = d_act + M;
    float* d_conf = d_pact + M;
    float* d_pconf = d_conf + (size_t)K * M;

    const int tid = threadIdx.x;
    if (p.use_ibm)
        for (int i = tid; i < nb; i += blockDim.x) s_acc[i] = p.weight_accum[i];

    // ------------------------------------------------------------------ phase A: matching + element-local terms
    float sum_l = 0.f, sum_pl = 0.f, sum_ct = 0.f, cnt_pos = 0.f, cnt_ppos = 0.f;
    for (int j = tid; j < M; j += blockDim.x) {
        const int b = j / p.P, pr = j - b * p.P;
        const float c = p.priors[(size_t)pr * p.prior_stride];
        // prior <-> GT matching (multisegment_loss.py:129-143); padding rows never win
        const float maxn = p.clip * 2.f;
        float best = INFINITY; int bi = 0;
        for (int g = 0; g < p.G; ++g) {
            const float* t = p.targets + ((size_t)b * p.G + g) * 3;
            const float left = (c - t[0]) * p.clip, right = (t[1] - c) * p.clip;
            float area = left + right;
            if (left < 0.f || right < 0.f) area = maxn;
            if (!p.valid[(size_t)b * p.G + g]) area = maxn * 2.f;
            if (area < best) { best = area; bi = g; }
        }
        const float* tb = p.targets + ((size_t)b * p.G + bi) * 3;
        const float t0 = (c - tb[0]) * p.clip, t1 = (tb[1] - c) * p.clip;
        int label = (int)(long long)tb[2];
        if (best >= maxn) label = 0;
        const float l0 = p.loc[2 * (size_t)j], l1 = p.loc[2 * (size_t)j + 1];
        const IouGrad ig = iou_with_grad(l0, l1, t0, t1);
        const int plabel = ig.iou < p.thresh ? 0 : label;          // :145-150
        const bool pos = label > 0, ppos = plabel > 0;
        s_iou[j] = ig.iou;
        s_meta[j] = (pos ? 1 : 0) | (ppos ? 2 : 0) | (label << 2);
        cnt_pos += pos ? 1.f : 0.f; cnt_ppos += ppos ? 1.f : 0.f;

        // --- GIoU loss on positives (:155-163, iou_loss 'giou')
        float g0 = 0.f, g1 = 0.f;
        if (pos) {
            const float eps = 1.1920928955078125e-07f;
            const float H = fmaxf(l0, t0) + fmaxf(l1, t1);
            const float Hc = fmaxf(H, eps), hp = H >= eps ? 1.f : 0.f;
            const float dH0 = l0 > t0 ? 1.f : (l0 == t0 ? 0.5f : 0.f), dH1 = l1 > t1 ? 1.f : (l1 == t1 ? 0.5f : 0.f);
            const float giou = ig.iou - (H - ig.U) / Hc;
            sum_l += 1.f - giou;
            const float dU0 = 1.f - ig.dI0, dU1 = 1.f - ig.dI1;
            g0 = -(ig.d0 - ((dH0 - dU0) / Hc - (H - ig.U) * hp * dH0 / (Hc * Hc)));
            g1 = -(ig.d1 - ((dH1 - dU1) / Hc - (H - ig.U) * hp * dH1 / (Hc * Hc)));
        }
        d_loc_l[2 * (size_t)j] = g0; d_loc_l[2 * (size_t)j + 1] = g1;

        // --- L1 on refined positives (:165-173); target (no_grad) = (loc_t - loc) / (0.5 * (l0 + l1))  (:151-153)
        const float q0 = p.ploc[2 * (size_t)j], q1 = p.ploc[2 * (size_t)j + 1];
        const float w = l0 + l1;
        float e0 = 0.f, e1 = 0.f;
        if (ppos) {
            const float pt0 = (t0 - l0) / (0.5f * w), pt1 = (t1 - l1) / (0.5f * w);
            const float r0 = q0 - pt0, r1 = q1 - pt1;
            sum_pl += fabsf(r0) + fabsf(r1);
            e0 = r0 > 0.f ? 1.f : (r0 < 0.f ? -1.f : 0.f);
            e1 = r1 > 0.f ? 1.f : (r1 < 0.f ? -1.f : 0.f);
        }
        d_ploc_l[2 * (size_t)j] = e0; d_ploc_l[2 * (size_t)j + 1] = e1;

        // --- IoU-quality BCE on positives (:175-189): target = clamp(IoU(refined segment, gt), 0), NOT detached
        float dc = 0.f, dl0 = 0.f, dl1 = 0.f, dq0 = 0.f, dq1 = 0.f;
        if (pos) {
            const float x = p.center[j];
            const float c0 = 0.5f * w * q0 + l0, c1 = 0.5f * w * q1 + l1;
            const IouGrad cg = iou_with_grad(c0, c1, t0, t1);
            const float q = fmaxf(cg.iou, 0.f);
            const float pass = cg.iou >= 0.f ? 1.f : 0.f;
            sum_ct += bce_logits(x, q);
            dc = sigmoidf(x) - q;
            const float a0 = -x * pass * cg.d0, a1 = -x * pass * cg.d1;     // d loss / d cur_i
            dq0 = a0 * 0.5f * w; dq1 = a1 * 0.5f * w;
            const float cross = 0.5f * (a0 * q0 + a1 * q1);                   // through w = l0 + l1
            dl0 = a0 + cross; dl1 = a1 + cross;
        }
        d_center[j] = dc;
        d_loc_ct[2 * (size_t)j] = dl0; d_loc_ct[2 * (size_t)j + 1] = dl1;
        d_ploc_ct[2 * (size_t)j] = dq0; d_ploc_ct[2 * (size_t)j + 1] = dq1;

        // --- EDL statistics, coarse and refined (unweighted; the IBM weight needs the batch-global bins)
        {
            float per, S, ay, gn, fn;
            edl_terms(p.conf + (size_t)j * K, K, label - 1, per, S, ay, gn, fn);
            s_per_c[j] = pos ? per : 0.f;
            s_gh_c[j] = gn * fn;
            s_bin_c[j] = pos ? (int)ceilf(gn * (float)nb) : -1;
            edl_terms(p.pconf + (size_t)j * K, K, plabel - 1, per, S, ay, gn, fn);
            s_per_p[j] = ppos ? per : 0.f;
            s_gh_p[j] = gn * fn;
            s_bin_p[j] = ppos ? (int)ceilf(gn * (float)nb) : -1;
            s_unc[j] = (float)K / S;
        }
        s_x[j] = p.act ? p.act[j] : 0.f;
        s_xp[j] = p.pact ? p.pact[j] : 0.f;
    }
    const float npos = block_sum(cnt_pos, s_red);
    const float nppos = block_sum(cnt_ppos, s_red);
    const float N = fmaxf(npos, 1.f), PN = fmaxf(nppos, 1.f);      // :243-244
    sum_l = block_sum(sum_l, s_red);
    sum_pl = block_sum(sum_pl, s_red);
    sum_ct = block_sum(sum_ct, s_red);
    __syncthreads();

    // ------------------------------------------------------------------ phase B/C: IBM per-bin EMA (cls_loss.py:263-268)
    // coarse call first, then the refined call sees the already updated buffer — the reference's call order.
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    if (p.use_ibm) {
        for (int pass = 0; pass < 2; ++pass) {
            const int* bins = pass ? s_bin_p : s_bin_c;
            const float* gh = pass ? s_gh_p : s_gh_c;
            for (int i = warp; i < nb; i += nwarps) {
                float s = 0.f, n = 0.f;
                for (int j = lane; j < M; j += 32)
                    if (bins[j] == i + 1) { s += gh[j]; n += 1.f; }
                s = warp_sum(s); n = warp_sum(n);
                if (lane == 0 && n > 0.f) s_acc[i] = p.momentum * s_acc[i] + (1.f - p.momentum) * (s / n);
            }
            __syncthreads();
            // weights = weight_accum[bin - 1] with python index wrap (bin 0 -> last bin, cls_loss.py:268)
            float* per = pass ? s_per_p : s_per_c;
            float* ghw = pass ? s_gh_p : s_gh_c;           // grad_hat is dead after the bin pass: reuse as the weight
            for (int j = tid; j < M; j += blockDim.x) {
                float wgt = 0.f;
                if (bins[j] >= 0) {
                    int idx = (bins[j] - 1) % nb; if (idx < 0) idx += nb;
                    wgt = s_acc[idx];
                }
                ghw[j] = wgt;
                per[j] *= wgt;
            }
            __syncthreads();
        }
        for (int i = tid; i < nb; i += blockDim.x) p.weight_accum[i] = s_acc[i];
    } else {
        for (int j = tid; j < M; j += blockDim.x) { s_gh_c[j] = 1.f; s_gh_p[j] = 1.f; }
        __syncthreads();
    }

    // ------------------------------------------------------------------ phase D: actionness (cls_loss.py:299-339)
    float loss_act[2] = {0.f, 0.f}, an_out[2] = {0.f, 0.f};
    for (int pass = 0; pass < 2; ++pass) {
        const float* xs = pass ? s_xp : s_x;
        const float* src = pass ? p.pact : p.act;
        float* dst = pass ? d_pact : d_act;
        if (!src) { for (int j = tid; j < M; j += blockDim.x) dst[j] = 0.f; continue; }
        const int bit = pass ? 2 : 1;
        const float np_ = pass ? nppos : npos;
        const float nn_ = (float)M - np_;
        const int topM = (int)fminf(np_, nn_) - 1;
        float bsum = 0.f, cnt = 0.f;
        float nmax = -INFINITY, pmax = -INFINITY; int nmi = 0x7fffffff, pmi = 0x7fffffff;
        for (int j = tid; j < M; j += blockDim.x) {
            const bool pos = (s_meta[j] & bit) != 0;
            const float x = xs[j];
            bool sel = true;
            if (topM > 0 && !pos) {
                int rank = 0;                      // number of negatives that sort before this one (ascending, index breaks ties)
                for (int i = 0; i < M; ++i) {
                    const bool ineg = (s_meta[i] & bit) == 0;
                    const float xi = xs[i];
                    rank += (ineg && (xi < x || (xi == x && i < j))) ? 1 : 0;
                }
                sel = rank < topM;
            }
            float g = 0.f;
            if (sel) {
                const float t = pos ? 1.f : 0.f;
                bsum += bce_logits(x, t); cnt += 1.f;
                g = sigmoidf(x) - t;
            }
            dst[j] = g;
            if (pos) { if (x > pmax) { pmax = x; pmi = j; } }
            else { if (x > nmax) { nmax = x; nmi = j; } }
        }
        bsum = block_sum(bsum, s_red);
        cnt = block_sum(cnt, s_red);
        float loss = bsum;
        int rank_arg = -1; float rank_g = 0.f;
        if (p.act_weight != 0.f && topM > 0) {     // rank term: max(0, margin - max(neg) + max(pos).detach())
            float nv, pv; int ni, pi;
            block_argmax(nmax, nmi, s_red, s_redi, nv, ni);
            block_argmax(pmax, pmi, s_red, s_redi, pv, pi);
            const float r = p.act_margin - nv + pv;
            if (r > 0.f) { loss += p.act_weight * r; rank_arg = ni; rank_g = -p.act_weight; }
        }
        const float AN = fmaxf(cnt, 1.f);
        loss_act[pass] = loss / AN; an_out[pass] = cnt;
        __syncthreads();
        for (int j = tid; j < M; j += blockDim.x) {
            float g = dst[j];
            if (j == rank_arg) g += rank_g;
            dst[j] = g / AN;
        }
    }

    // ------------------------------------------------------------------ phase E: weighted sums, calibration, gradient scaling
    float sum_c = 0.f, sum_pc = 0.f, sum_cal = 0.f;
    const float invM = 1.f / (float)M;
    for (int j = tid; j < M; j += blockDim.x) {
        const int meta = s_meta[j];
        const bool pos = meta & 1, ppos = (meta & 2) != 0;
        const int label = meta >> 2;
        sum_c += s_per_c[j]; sum_pc += s_per_p[j];
        d_loc_l[2 * (size_t)j] /= N; d_loc_l[2 * (size_t)j + 1] /= N;
        d_loc_ct[2 * (size_t)j] /= N; d_loc_ct[2 * (size_t)j + 1] /= N;
        d_ploc_ct[2 * (size_t)j] /= N; d_ploc_ct[2 * (size_t)j + 1] /= N;
        d_center[j] /= N;
        d_ploc_l[2 * (size_t)j] /= PN; d_ploc_l[2 * (size_t)j + 1] /= PN;
        // coarse EDL gradient: w/N * (1/S - [k == y]/alpha_y) * d alpha_k / d z_k
        {
            const float* z = p.conf + (size_t)j * K;
            float* dz = d_conf + (size_t)j * K;
            if (pos) {
                float S = 0.f;
                for (int k = 0; k < K; ++k) S += expf(fminf(fmaxf(z[k], -10.f), 10.f)) + 1.f;
                const float wgt = s_gh_c[j] / N;
                for (int k = 0; k < K; ++k) {
                    const float zk = z[k];
                    const float ev = expf(fminf(fmaxf(zk, -10.f), 10.f));
                    const float da = (zk >= -10.f && zk <= 10.f) ? ev : 0.f;
                    float g = 1.f / S;
                    if (k == label - 1) g -= 1.f / (ev + 1.f);
                    dz[k] = wgt * g * da;
                }
            } else {
                for (int k = 0; k < K; ++k) dz[k] = 0.f;
            }
        }
        // refined EDL gradient + IoU-aware calibration over ALL priors (cls_loss.py:120-129, mean).  The reference
        // flattens its [P,B] IoU buffer against [B*P] logits (multisegment_loss.py:116,146,236): element j pairs with
        // the IoU of prior j / B of sample j % B.
        {
            const float* z = p.pconf + (size_t)j * K;
            float* dz = d_pconf + (size_t)j * K;
            float cal_g = 0.f;
            const float unc = s_unc[j];
            if (p.iou_aware) {
                float iou = s_iou[(size_t)(j % p.B) * p.P + (j / p.B)];
                if (iou < 0.f) iou = 1e-3f;
                sum_cal += -iou * logf(1.f - unc) - (1.f - iou) * logf(unc);
                cal_g = (iou / (1.f - unc) - (1.f - iou) / unc) * invM;      // d mean(reg) / d unc_j
            }
            const float S = (float)K / unc;
            const float wgt = ppos ? s_gh_p[j] / PN : 0.f;
            const int plabel = ppos ? label : 0;
            for (int k = 0; k < K; ++k) {
                const float zk = z[k];
                const float ev = expf(fminf(fmaxf(zk, -10.f), 10.f));
                const float da = (zk >= -10.f && zk <= 10.f) ? ev : 0.f;
                float g = 0.f;
                if (ppos) {
                    g = 1.f / S;
                    if (k == plabel - 1) g -= 1.f / (ev + 1.f);
                    g *= wgt;
                }
                g += cal_g * (-unc / S);                                      // d unc / d alpha_k = -K / S^2
                dz[k] = g * da;
            }
        }
    }
    sum_c = block_sum(sum_c, s_red);
    sum_pc = block_sum(sum_pc, s_red);
    sum_cal = block_sum(sum_cal, s_red);
    if (tid == 0) {
        const float iouc = p.iou_aware ? sum_cal * invM : 0.f;
        p.losses[0] = sum_l / N;
        p.losses[1] = sum_c / N;
        p.losses[2] = sum_pl / PN;
        p.losses[3] = sum_pc / PN + iouc;
        p.losses[4] = sum_ct / N;
        p.losses[5] = loss_act[0];
        p.losses[6] = loss_act[1];
        p.losses[7] = npos; p.losses[8] = nppos; p.losses[9] = an_out[0]; p.losses[10] = an_out[1]; p.losses[11] = iouc;
    }
}

struct MslBwdParams {
    int M, K;
    const float* ws;
    const float* gl;      // [7] upstream gradients of the 7 losses
    float *g_loc, *g_conf, *g_ploc, *g_pconf, *g_center, *g_act, *g_pact;
};

__global__ void msl_backward_kernel(const MslBwdParams p) {
    const int M = p.M, K = p.K;
    const float* d_loc_l = p.ws;
    const float* d_loc_ct = d_loc_l + 2 * (size_t)M;
    const float* d_ploc_l = d_loc_ct + 2 * (size_t)M;
    const float* d_ploc_ct = d_ploc_l + 2 * (size_t)M;
    const float* d_center = d_ploc_ct + 2 * (size_t)M;
    const float* d_act = d_center + M;
    const float* d_pact = d_act + M;
    const float* d_conf = d_pact + M;
    const float* d_pconf = d_conf + (size_t)K * M;
    const float g0 = p.gl[0], g1 = p.gl[1], g2 = p.gl[2], g3 = p.gl[3], g4 = p.gl[4], g5 = p.gl[5], g6 = p.gl[6];
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * K; i += stride) {
        p.g_conf[i] = g1 * d_conf[i];
        p.g_pconf[i] = g3 * d_pconf[i];
        if (i < 2 * M) {
            p.g_loc[i] = g0 * d_loc_l[i] + g4 * d_loc_ct[i];
            p.g_ploc[i] = g2 * d_ploc_l[i] + g4 * d_ploc_ct[i];
        }
        if (i < M) {
            p.g_center[i] = g4 * d_center[i];
            if (p.g_act) p.g_act[i] = g5 * d_act[i];
            if (p.g_pact) p.g_pact[i] = g6 * d_pact[i];
        }
    }
}

static size_t msl_smem_bytes(int M) { return (size_t)M * 11 * 4 + (kMslMaxBins + 64) * 4; }

}  // namespace otal

using namespace otal;

extern "C" {

long long otal_msl_workspace_floats(int B, int P, int K) { return (long long)msl_ws_floats(B * P, K); }

int otal_msl_forward(const otal_msl_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("msl: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (d->B <= 0 || d->P <= 0 || d->K <= 1 || d->G <= 0) { set_last_error_msg("msl: bad dimension"); return OTAL_ERR_BAD_ARG; }
    if (!d->loc || !d->conf || !d->prop_loc || !d->prop_conf || !d->center || !d->priors || !d->targets || !d->valid ||
        !d->losses || !d->workspace) { set_last_error_msg("msl: null pointer"); return OTAL_ERR_BAD_ARG; }
    if (d->use_ibm && (!d->weight_accum || d->num_bins <= 0 || d->num_bins > kMslMaxBins)) {
        set_last_error_msg("msl: IBM needs weight_accum and 1..256 bins"); return OTAL_ERR_BAD_ARG;
    }
    const long long M = (long long)d->B * d->P;
    const size_t smem = msl_smem_bytes((int)M);
    if (M > (1 << 20) || smem > 227 * 1024) {
        set_last_error_msg("msl: B*P too large for the single-CTA kernel (shared-memory staging)"); return OTAL_ERR_UNSUPPORTED;
    }
    MslParams p{};
    p.B = d->B; p.P = d->P; p.K = d->K; p.G = d->G; p.M = (int)M;
    p.clip = d->clip_length; p.thresh = d->overlap_thresh;
    p.use_ibm = d->use_ibm; p.num_bins = d->use_ibm ? d->num_bins : 1; p.momentum = d->momentum;
    p.iou_aware = d->iou_aware; p.act_weight = d->act_weight; p.act_margin = d->act_margin;
    p.prior_stride = d->prior_stride > 0 ? d->prior_stride : 1;
    p.loc = d->loc; p.conf = d->conf; p.ploc = d->prop_loc; p.pconf = d->prop_conf; p.center = d->center;
    p.act = d->act; p.pact = d->prop_act; p.priors = d->priors; p.targets = d->targets; p.valid = d->valid;
    p.weight_accum = d->weight_accum; p.losses = d->losses; p.ws = d->workspace;
    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(msl_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        once.mark(once_dev);
    }
    msl_forward_kernel<<<1, kMslThreads, smem, stream>>>(p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_msl_backward(int B, int P, int K, const float* workspace, const float* grad_losses, float* g_loc, float* g_conf,
                      float* g_prop_loc, float* g_prop_conf, float* g_center, float* g_act, float* g_prop_act, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B <= 0 || P <= 0 || K <= 1 || !workspace || !grad_losses || !g_loc || !g_conf || !g_prop_loc || !g_prop_conf || !g_center) {
        set_last_error_msg("msl_backward: bad argument"); return OTAL_ERR_BAD_ARG;
    }
    MslBwdParams p{};
    p.M = B * P; p.K = K; p.ws = workspace; p.gl = grad_losses;
    p.g_loc = g_loc; p.g_conf = g_conf; p.g_ploc = g_prop_loc; p.g_pconf = g_prop_conf; p.g_center = g_center;
    p.g_act = g_act; p.g_pact = g_prop_act;
    const int total = p.M * K;
    int grid = (total + 255) / 256; if (grid > 148 * 4) grid = 148 * 4;
    msl_backward_kernel<<<grid, 256, 0, stream>>>(p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

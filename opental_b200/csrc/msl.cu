// MultiSegmentLoss as ONE single-CTA kernel: prior<->GT matching, GIoU / L1 / IoU-quality BCE regression terms, the
// classification loss (evidential / Dirichlet with IBM re-weighting and IoU-aware calibration, or softmax focal) and the
// positive-unlabeled actionness loss — forward values AND the gradient of every term w.r.t. every head output.
// Three flavours of the reference loss share the kernel (otal_msl_desc.flavour):
//   0  THUMOS14 OpenTAL   AFSD/thumos14/multisegment_loss.py:92-259   MultiSegmentLoss.forward (matching :120-153, terms
//                         :155-241, batch-global normalisation :243-256); AFSD/thumos14/cls_loss.py:120-129 iou_calib,
//                         :132-168,:212-278 EvidenceLoss (loss_type 'log', evidence 'exp', with_ibm: 50-bin EMA),
//                         :299-339 ActionnessLoss (top-M lowest-scoring negatives, rank term)
//   1  ActivityNet OpenTAL AFSD/anet/multisegment_loss.py:106-301: level-range gated matching (bounds :69-83, :156-166),
//                         refined positives need IoU >= min(piou, best IoU among the sample's positives) (:178-184),
//                         smooth-L1 refinement loss (:206), every term normalised PER SAMPLE and averaged over the batch
//                         (:268-297); AFSD/anet/cls_loss.py:116-152,:225-232 stateless IBM weight
//                         1 / (||z||_1 exp(c g) + 1e-10) that back-propagates through ||z||_1; ActionnessLoss per sample
//   2  THUMOS14 closed set AFSD/thumos14/multisegment_loss.py:193-195,:217-218 with cls_loss.py:6-78 FocalLoss_Ori on the
//                         softmax scores of ALL priors (background = class 0, alpha 0.25 / 0.75, gamma 2), no actionness
// all with AFSD/thumos14/multisegment_loss.py:20-53 iou_loss ('calc iou', 'giou'), eps = fp32 machine epsilon.
//
// Why one CTA: the whole problem is B x 126 (189) priors (1008 elements at batch 8, 18.6 KB of inputs per clip, SURVEY
// §8d) but every normaliser is global to a GROUP of elements — the batch (flavours 0, 2) or the sample (flavour 1): N, PN,
// AN, the 50 IBM bins, the top-M rank of every negative.  The reference spends ~300 launches and ~100 host syncs on it;
// here the elements are staged in shared memory once and every group quantity is a warp-shuffle + shared-memory reduction
// in a fixed order (deterministic).  The kernel is latency-bound by construction: its roofline is the launch, not HBM.
//
// Gradients: the kernel writes "unit" gradients (d loss_i / d input for each of the 7 returned losses) into a small
// workspace; otal_msl_backward combines them with the 7 upstream scalars.  Sub-gradient conventions follow torch:
// min/max ties split the gradient in half, clamp passes the gradient on the closed interval, |x| has gradient 0 at 0.
#include "common.cuh"

namespace otal {

constexpr int kMslThreads = 1024;
constexpr int kMslWarps = kMslThreads / 32;
constexpr int kMslMaxBins = 256;
constexpr int kMslMaxGroups = 64;
enum { kMslThumos = 0, kMslAnet = 1, kMslFocal = 2 };
enum { kRwNone = 0, kRwIbm = 1, kRwIb = 2, kRwFocal = 3, kRwGhm = 4 };

struct MslParams {
    int B, P, K, G, M;
    int flavour, groups, L;      // normalisation groups: 1 x M (the batch) or B x P (per sample); elements of a group are contiguous
    float clip, thresh;
    int use_ibm, num_bins;
    float momentum;
    int iou_aware;
    float act_weight, act_margin;
    int prior_stride;
    float ibm_coeff, focal_alpha, focal_gamma;
    int reweight;                // flavour 0 only: kRwNone / kRwIbm (binned EMA) / kRwIb / kRwFocal / kRwGhm (cls_loss.py:221-272)
    int cls_all;                 // flavour 0 only: no os_head — every prior is a classification sample, class 0 = background
    float edl_alpha0, edl_gamma; // focal-EDL: weight of class 0 (1 - alpha0 for the others), exponent
    double* ghm_acc;             // GHM: fp64 per-bin EMA state (EvidenceLoss.acc_sum), updated when momentum > 0
    float bounds[16];            // flavour 1: (left, right] range of max(left, right) per pyramid level
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

// NQ sums per group, deterministic: a group's L elements are split over `wpg` warps (lane-strided), the warp partials are
// added in a fixed order.  f(j, v) writes the NQ contributions of element j.  out[q * kMslMaxGroups + g].
template <int NQ, class F>
__device__ void group_sums(int groups, int L, float* out, float* s_part, F f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpg = groups >= kMslWarps ? 1 : kMslWarps / groups;
    const int per_round = kMslWarps / wpg;
    for (int base = 0; base < groups; base += per_round) {
        const int g = base + warp / wpg, sub = warp % wpg;
        float acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
        if (warp / wpg < per_round && g < groups)
            for (int i = sub * 32 + lane; i < L; i += wpg * 32) {
                float v[NQ];
                f(g * L + i, v);
#pragma unroll
                for (int q = 0; q < NQ; ++q) acc[q] += v[q];
            }
        __syncthreads();                 // s_part may still be read by the previous round / call
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float t = warp_sum(acc[q]);
            if (lane == 0) s_part[q * kMslWarps + warp] = t;
        }
        __syncthreads();
        if ((int)threadIdx.x < per_round && base + (int)threadIdx.x < groups)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float t = 0.f;
                for (int k = 0; k < wpg; ++k) t += s_part[q * kMslWarps + threadIdx.x * wpg + k];
                out[q * kMslMaxGroups + base + threadIdx.x] = t;
            }
    }
    __syncthreads();
}

// per group: max of f(j) over the elements with ok(j), lowest index among equal maxima; (-inf, 0x7fffffff) for an empty set
template <class F>
__device__ void group_argmax(int groups, int L, float* out_v, int* out_i, float* s_part, int* s_parti, F f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpg = groups >= kMslWarps ? 1 : kMslWarps / groups;
    const int per_round = kMslWarps / wpg;
    for (int base = 0; base < groups; base += per_round) {
        const int g = base + warp / wpg, sub = warp % wpg;
        float v = -INFINITY; int idx = 0x7fffffff;
        if (warp / wpg < per_round && g < groups)
            for (int i = sub * 32 + lane; i < L; i += wpg * 32) {
                float x; 
                if (f(g * L + i, x) && (x > v || (x == v && g * L + i < idx))) { v = x; idx = g * L + i; }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        __syncthreads();
        if (lane == 0) { s_part[warp] = v; s_parti[warp] = idx; }
        __syncthreads();
        if ((int)threadIdx.x < per_round && base + (int)threadIdx.x < groups) {
            float tv = -INFINITY; int ti = 0x7fffffff;
            for (int k = 0; k < wpg; ++k) {
                const float ov = s_part[threadIdx.x * wpg + k];
                const int oi = s_parti[threadIdx.x * wpg + k];
                if (ov > tv || (ov == tv && oi < ti)) { tv = ov; ti = oi; }
            }
            out_v[base + threadIdx.x] = tv; out_i[base + threadIdx.x] = ti;
        }
    }
    __syncthreads();
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
__device__ __forceinline__ void edl_terms(const float* z, int K, int y, float& per, float& S, float& alpha_y, float& gnorm, float& fnorm,
                                          float* alpha_max = nullptr) {
    S = 0.f; fnorm = 0.f; alpha_y = 1.f;
    float amax = 0.f;
    for (int k = 0; k < K; ++k) {
        const float zk = z[k];
        const float a = expf(fminf(fmaxf(zk, -10.f), 10.f)) + 1.f;
        S += a; fnorm += fabsf(zk);
        amax = fmaxf(amax, a);
        if (k == y) alpha_y = a;
    }
    if (alpha_max) *alpha_max = amax;
    per = logf(S) - logf(alpha_y);
    gnorm = fabsf(1.f / alpha_y - (float)K / S);
}

// FocalLoss_Ori of one row of logits (cls_loss.py:60-78 on F.softmax, multisegment_loss.py:193-195): value and d / d z
__device__ __forceinline__ float focal_row(const float* z, float* dz, int K, int y, float alpha0, float gamma) {
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, z[k]);
    float Z = 0.f;
    for (int k = 0; k < K; ++k) Z += expf(z[k] - m);
    const float py = expf(z[y] - m) / Z;
    const float pt = py + 1e-6f;
    const float a = y == 0 ? alpha0 : 1.f - alpha0;
    const float om = 1.f - pt;
    const float loss = -powf(om, gamma) * (a * logf(pt));
    const float dpt = a * (gamma * powf(om, gamma - 1.f) * logf(pt) - powf(om, gamma) / pt);
    for (int k = 0; k < K; ++k) {
        const float pk = expf(z[k] - m) / Z;
        dz[k] = dpt * py * ((k == y ? 1.f : 0.f) - pk);
    }
    return loss;
}

// prior <-> GT matching of one prior (multisegment_loss.py:129-143; anet :156-166 adds the level range): target offsets
// and label (0 = background); padding rows never win
__device__ __forceinline__ void match_prior(const MslParams& p, int b, int pr, float& t0, float& t1, int& label) {
    const float c = p.priors[(size_t)pr * p.prior_stride];
    const float maxn = p.clip * 2.f;
    float lb = -INFINITY, rb = INFINITY;
    if (p.flavour == kMslAnet) {
        const int lvl = (int)p.priors[(size_t)pr * p.prior_stride + 1];
        lb = p.bounds[2 * lvl]; rb = p.bounds[2 * lvl + 1];
    }
    float best = INFINITY; int bi = 0;
    for (int g = 0; g < p.G; ++g) {
        const float* t = p.targets + ((size_t)b * p.G + g) * 3;
        const float left = (c - t[0]) * p.clip, right = (t[1] - c) * p.clip;
        const float md = fmaxf(left, right);
        float area = left + right;
        if (left < 0.f || right < 0.f || md <= lb || md > rb) area = maxn;
        if (!p.valid[(size_t)b * p.G + g]) area = maxn * 2.f;
        if (area < best) { best = area; bi = g; }
    }
    const float* tb = p.targets + ((size_t)b * p.G + bi) * 3;
    t0 = (c - tb[0]) * p.clip; t1 = (tb[1] - c) * p.clip;
    label = (int)(long long)tb[2];
    if (best >= maxn) label = 0;
}

// GHM bin of a gradient length g in [0, 1]: edges i / nb as fp32 (python: float(i) / nb -> float32 tensor), the last one + 1e-6
// (cls_loss.py:105-108, :236-238)
__device__ __forceinline__ int ghm_bin(float g, int nb) {
    auto edge = [nb](int i) { return i == nb ? (float)(1.0 + 1e-6) : (float)((double)i / (double)nb); };
    int i = min(max((int)(g * (float)nb), 0), nb - 1);
    while (i > 0 && g < edge(i)) --i;
    while (i < nb - 1 && g >= edge(i + 1)) ++i;
    return i;
}

// s_meta: pos | ppos << 1 | label << 2 (10 bits) | (coarse IBM bin + 1) << 12 (9 bits) | (refined IBM bin + 1) << 21 (9 bits)
__device__ __forceinline__ int meta_bin(int meta, int pass) { return ((meta >> (pass ? 21 : 12)) & 0x1ff) - 1; }

__global__ void __launch_bounds__(kMslThreads, 1)
msl_forward_kernel(const MslParams p) {
    extern __shared__ __align__(16) unsigned char msl_smem[];
    const int M = p.M, K = p.K, nb = p.num_bins, NG = p.groups, L = p.L;
    float* s_iou = reinterpret_cast<float*>(msl_smem);       // [M] IoU(loc, loc_t) (no_grad)
    int* s_meta = reinterpret_cast<int*>(s_iou + M);         // [M]
    float* s_gh_c = reinterpret_cast<float*>(s_meta + M);    // [M] IBM grad_hat, coarse -> the sample's weight
    float* s_per_c = s_gh_c + M;                             // [M] unweighted -> weighted classification loss, coarse
    float* s_gh_p = s_per_c + M;
    float* s_per_p = s_gh_p + M;
    float* s_x = s_per_p + M;                                // [M] actionness logits
    float* s_xp = s_x + M;                                   // [M] refined actionness logits
    float* s_unc = s_xp + M;                                 // [M] Dirichlet uncertainty of prop_conf
    float* s_a = s_unc + M;                                  // [M] x 3 per-element loss terms on their way to the group sums
    float* s_b = s_a + M;
    float* s_c = s_b + M;
    float* s_acc = s_c + M;                                  // [nb] IBM EMA / GHM per-bin weight
    double* s_accd = reinterpret_cast<double*>(s_acc + kMslMaxBins);   // [nb] GHM EMA (fp64 like the reference's python floats)
    float* s_part = reinterpret_cast<float*>(s_accd + kMslMaxBins);    // [4 x 32] warp partials
    int* s_parti = reinterpret_cast<int*>(s_part + 4 * kMslWarps);   // [32]
    float* s_grp = reinterpret_cast<float*>(s_parti + kMslWarps);    // [12 x 64] per-group results
    int* s_grpi = reinterpret_cast<int*>(s_grp + 12 * kMslMaxGroups); // [2 x 64]
    float* g_npos = s_grp, *g_nppos = s_grp + kMslMaxGroups, *g_thr = s_grp + 2 * kMslMaxGroups;
    float* g_sum = s_grp + 3 * kMslMaxGroups;                // [4 x 64] scratch of the current reduction
    float* g_nmax = s_grp + 7 * kMslMaxGroups, *g_pmax = s_grp + 8 * kMslMaxGroups;
    float* g_AN = s_grp + 9 * kMslMaxGroups;                 // [2 x 64] actionness normalisers, coarse / refined
    int* g_nmi = s_grpi, *g_pmi = s_grpi + kMslMaxGroups;

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
    const bool anet = p.flavour == kMslAnet, focal = p.flavour == kMslFocal;
    const int rw = (anet || focal) ? kRwNone : p.reweight;
    const bool ghm = rw == kRwGhm;
    const bool binned = rw == kRwIbm || ghm;                 // batch-global bins: the THUMOS14 IBM (50-bin EMA state) or GHM
    const bool cls_all = !anet && !focal && p.cls_all;
    if (rw == kRwIbm)
        for (int i = tid; i < nb; i += blockDim.x) s_acc[i] = p.weight_accum[i];
    if (ghm)
        for (int i = tid; i < nb; i += blockDim.x) s_accd[i] = p.ghm_acc[i];

    // ------------------------------------------------------------------ phase A1: matching, IoU of the coarse prediction
    for (int j = tid; j < M; j += blockDim.x) {
        const int b = j / p.P, pr = j - b * p.P;
        float t0, t1; int label;
        match_prior(p, b, pr, t0, t1, label);
        const IouGrad ig = iou_with_grad(p.loc[2 * (size_t)j], p.loc[2 * (size_t)j + 1], t0, t1);
        s_iou[j] = ig.iou;
        s_meta[j] = (label > 0 ? 1 : 0) | (label << 2);
    }
    __syncthreads();
    // refined-positive threshold per group: piou, or min(piou, best IoU among the sample's positives) (anet :178-184)
    if (anet) {
        group_argmax(NG, L, g_thr, g_nmi, s_part, s_parti, [&](int j, float& x) { x = s_iou[j]; return (s_meta[j] & 1) != 0; });
        if (tid < NG) g_thr[tid] = g_nmi[tid] == 0x7fffffff ? p.thresh : fminf(p.thresh, g_thr[tid]);
    } else if (tid < NG) {
        g_thr[tid] = p.thresh;
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase A2: element-local terms
    for (int j = tid; j < M; j += blockDim.x) {
        const int b = j / p.P, pr = j - b * p.P;
        const int grp = NG == 1 ? 0 : j / L;
        float t0, t1; int label;
        match_prior(p, b, pr, t0, t1, label);
        const float l0 = p.loc[2 * (size_t)j], l1 = p.loc[2 * (size_t)j + 1];
        const IouGrad ig = iou_with_grad(l0, l1, t0, t1);
        const int plabel = ig.iou < g_thr[grp] ? 0 : label;        // :145-150
        const bool pos = label > 0, ppos = plabel > 0;
        int meta = (pos ? 1 : 0) | (ppos ? 2 : 0) | (label << 2);

        // --- GIoU loss on positives (:155-163, iou_loss 'giou')
        float g0 = 0.f, g1 = 0.f, term_l = 0.f;
        if (pos) {
            const float eps = 1.1920928955078125e-07f;
            const float H = fmaxf(l0, t0) + fmaxf(l1, t1);
            const float Hc = fmaxf(H, eps), hp = H >= eps ? 1.f : 0.f;
            const float dH0 = l0 > t0 ? 1.f : (l0 == t0 ? 0.5f : 0.f), dH1 = l1 > t1 ? 1.f : (l1 == t1 ? 0.5f : 0.f);
            const float giou = ig.iou - (H - ig.U) / Hc;
            term_l = 1.f - giou;
            const float dU0 = 1.f - ig.dI0, dU1 = 1.f - ig.dI1;
            g0 = -(ig.d0 - ((dH0 - dU0) / Hc - (H - ig.U) * hp * dH0 / (Hc * Hc)));
            g1 = -(ig.d1 - ((dH1 - dU1) / Hc - (H - ig.U) * hp * dH1 / (Hc * Hc)));
        }
        d_loc_l[2 * (size_t)j] = g0; d_loc_l[2 * (size_t)j + 1] = g1;
        s_a[j] = term_l;

        // --- L1 (smooth-L1, beta 1, in the ActivityNet flavour: anet :206) on refined positives (:165-173); target (no_grad)
        // = (loc_t - loc) / (0.5 * (l0 + l1))  (:151-153)
        const float q0 = p.ploc[2 * (size_t)j], q1 = p.ploc[2 * (size_t)j + 1];
        const float w = l0 + l1;
        float e0 = 0.f, e1 = 0.f, term_pl = 0.f;
        if (ppos) {
            const float pt0 = (t0 - l0) / (0.5f * w), pt1 = (t1 - l1) / (0.5f * w);
            const float r0 = q0 - pt0, r1 = q1 - pt1;
            if (anet) {
                term_pl = (fabsf(r0) < 1.f ? 0.5f * r0 * r0 : fabsf(r0) - 0.5f) + (fabsf(r1) < 1.f ? 0.5f * r1 * r1 : fabsf(r1) - 0.5f);
                e0 = fabsf(r0) < 1.f ? r0 : (r0 > 0.f ? 1.f : -1.f);
                e1 = fabsf(r1) < 1.f ? r1 : (r1 > 0.f ? 1.f : -1.f);
            } else {
                term_pl = fabsf(r0) + fabsf(r1);
                e0 = r0 > 0.f ? 1.f : (r0 < 0.f ? -1.f : 0.f);
                e1 = r1 > 0.f ? 1.f : (r1 < 0.f ? -1.f : 0.f);
            }
        }
        d_ploc_l[2 * (size_t)j] = e0; d_ploc_l[2 * (size_t)j + 1] = e1;
        s_b[j] = term_pl;

        // --- IoU-quality BCE on positives (:175-189): target = clamp(IoU(refined segment, gt), 0), NOT detached
        float dc = 0.f, dl0 = 0.f, dl1 = 0.f, dq0 = 0.f, dq1 = 0.f, term_ct = 0.f;
        if (pos) {
            const float x = p.center[j];
            const float c0 = 0.5f * w * q0 + l0, c1 = 0.5f * w * q1 + l1;
            const IouGrad cg = iou_with_grad(c0, c1, t0, t1);
            const float q = fmaxf(cg.iou, 0.f);
            const float pass = cg.iou >= 0.f ? 1.f : 0.f;
            term_ct = bce_logits(x, q);
            dc = sigmoidf(x) - q;
            const float a0 = -x * pass * cg.d0, a1 = -x * pass * cg.d1;     // d loss / d cur_i
            dq0 = a0 * 0.5f * w; dq1 = a1 * 0.5f * w;
            const float cross = 0.5f * (a0 * q0 + a1 * q1);                   // through w = l0 + l1
            dl0 = a0 + cross; dl1 = a1 + cross;
        }
        d_center[j] = dc;
        d_loc_ct[2 * (size_t)j] = dl0; d_loc_ct[2 * (size_t)j + 1] = dl1;
        d_ploc_ct[2 * (size_t)j] = dq0; d_ploc_ct[2 * (size_t)j + 1] = dq1;
        s_c[j] = term_ct;

        // --- classification statistics, coarse and refined
        if (focal) {
            // every prior is a sample, background = class 0; the unit gradient is final up to the 1/N scale
            s_per_c[j] = focal_row(p.conf + (size_t)j * K, d_conf + (size_t)j * K, K, label, p.focal_alpha, p.focal_gamma);
            s_per_p[j] = focal_row(p.pconf + (size_t)j * K, d_pconf + (size_t)j * K, K, plabel, p.focal_alpha, p.focal_gamma);
            s_gh_c[j] = 1.f; s_gh_p[j] = 1.f; s_unc[j] = 1.f;
        } else {
            // EDL, unweighted; the THUMOS14 IBM weight needs the batch-global bins (phase B), the ActivityNet one is local:
            // w = 1 / (||z||_1 exp(c g) + 1e-10); s_gh_* then holds the weight and s_per_* the weighted loss
            float per, S, ay, gn, fn, amax;
            auto local_weight = [&](int y, bool inc) {
                if (anet && p.use_ibm) return 1.f / (fn * expf(p.ibm_coeff * gn) + 1e-10f);
                if (rw == kRwIb) return inc ? 1.f / (gn * fn) : 0.f;                        // cls_loss.py:250-256
                if (rw == kRwFocal) return (y == 0 ? p.edl_alpha0 : 1.f - p.edl_alpha0) * powf(1.f - amax / S, p.edl_gamma);   // :221-227
                return 1.f;
            };
            const int yc = cls_all ? label : label - 1;
            const bool inc_c = cls_all || pos;
            edl_terms(p.conf + (size_t)j * K, K, yc, per, S, ay, gn, fn, &amax);
            float wgt = local_weight(yc, inc_c);
            s_per_c[j] = inc_c ? per * (binned ? 1.f : wgt) : 0.f;
            s_gh_c[j] = binned ? (ghm ? gn : gn * fn) : wgt;
            if (binned) meta |= ((inc_c ? (ghm ? ghm_bin(gn, nb) + 1 : (int)ceilf(gn * (float)nb)) : -1) + 1) << 12;
            const int yp = cls_all ? plabel : plabel - 1;
            const bool inc_p = cls_all || ppos;
            edl_terms(p.pconf + (size_t)j * K, K, yp, per, S, ay, gn, fn, &amax);
            wgt = local_weight(yp, inc_p);
            s_per_p[j] = inc_p ? per * (binned ? 1.f : wgt) : 0.f;
            s_gh_p[j] = binned ? (ghm ? gn : gn * fn) : wgt;
            if (binned) meta |= ((inc_p ? (ghm ? ghm_bin(gn, nb) + 1 : (int)ceilf(gn * (float)nb)) : -1) + 1) << 21;
            s_unc[j] = (float)K / S;
        }
        s_meta[j] = meta;
        s_x[j] = p.act ? p.act[j] : 0.f;
        s_xp[j] = p.pact ? p.pact[j] : 0.f;
    }
    __syncthreads();
    group_sums<2>(NG, L, g_npos, s_part, [&](int j, float* v) { v[0] = (float)(s_meta[j] & 1); v[1] = (float)((s_meta[j] >> 1) & 1); });
    // g_npos / g_nppos are adjacent rows of s_grp: group_sums<2> filled both.  Normalisers (:243-244; anet :268-297)
    const float invG = 1.f / (float)NG;
    group_sums<3>(NG, L, g_sum, s_part, [&](int j, float* v) { v[0] = s_a[j]; v[1] = s_b[j]; v[2] = s_c[j]; });
    if (tid == 0) {
        float a = 0.f, b2 = 0.f, c = 0.f;
        for (int g = 0; g < NG; ++g) {
            const float N = fmaxf(g_npos[g], 1.f), PN = fmaxf(g_nppos[g], 1.f);
            a += g_sum[g] / N; b2 += g_sum[kMslMaxGroups + g] / PN; c += g_sum[2 * kMslMaxGroups + g] / N;
        }
        p.losses[0] = a * invG; p.losses[2] = b2 * invG; p.losses[4] = c * invG;
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase B/C: THUMOS14 IBM per-bin EMA (cls_loss.py:263-268)
    // coarse call first, then the refined call sees the already updated buffer — the reference's call order.
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    if (rw == kRwIbm) {
        for (int pass = 0; pass < 2; ++pass) {
            const float* gh = pass ? s_gh_p : s_gh_c;
            for (int i = warp; i < nb; i += nwarps) {
                float s = 0.f, n = 0.f;
                for (int j = lane; j < M; j += 32)
                    if (meta_bin(s_meta[j], pass) == i + 1) { s += gh[j]; n += 1.f; }
                s = warp_sum(s); n = warp_sum(n);
                if (lane == 0 && n > 0.f) s_acc[i] = p.momentum * s_acc[i] + (1.f - p.momentum) * (s / n);
            }
            __syncthreads();
            // weights = weight_accum[bin - 1] with python index wrap (bin 0 -> last bin, cls_loss.py:268)
            float* per = pass ? s_per_p : s_per_c;
            float* ghw = pass ? s_gh_p : s_gh_c;           // grad_hat is dead after the bin pass: reuse as the weight
            for (int j = tid; j < M; j += blockDim.x) {
                float wgt = 0.f;
                const int bin = meta_bin(s_meta[j], pass);
                if (bin >= 0) {
                    int idx = (bin - 1) % nb; if (idx < 0) idx += nb;
                    wgt = s_acc[idx];
                }
                ghw[j] = wgt;
                per[j] *= wgt;
            }
            __syncthreads();
        }
        for (int i = tid; i < nb; i += blockDim.x) p.weight_accum[i] = s_acc[i];
    } else if (ghm) {
        // GHM (cls_loss.py:228-249): the [M,K] gradient lengths |1/alpha - K/S| * y are zero off the target class, so every
        // included row puts K - 1 zeros into bin 0 and its target entry into ghm_bin(g); per-bin EMA of the COUNTS in fp64,
        // weight = 1 / acc[bin] / (number of non-empty bins).  Coarse call first, the refined call sees the updated state.
        int& s_nonempty = s_grpi[2 * kMslMaxGroups - 1];       // a spare slot of the per-group index table (GHM runs with one group)
        for (int pass = 0; pass < 2; ++pass) {
            const float n_inc = cls_all ? (float)M : (pass ? g_nppos[0] : g_npos[0]);
            if (tid == 0) s_nonempty = 0;
            __syncthreads();
            for (int i = warp; i < nb; i += nwarps) {
                float n = 0.f;
                for (int j = lane; j < M; j += 32)
                    if (meta_bin(s_meta[j], pass) == i + 1) n += 1.f;      // stored as ghm_bin + 1
                n = warp_sum(n);
                if (lane == 0) {
                    const double cnt = (double)n + (i == 0 ? (double)(K - 1) * (double)n_inc : 0.0);
                    float w = 0.f;
                    if (cnt > 0.0) {
                        const double acc = p.momentum > 0.f ? (double)p.momentum * s_accd[i] + (1.0 - (double)p.momentum) * cnt : cnt;
                        if (p.momentum > 0.f) s_accd[i] = acc;
                        w = (float)(1.0 / fmax(acc, 1e-300));
                        atomicAdd(&s_nonempty, 1);
                    }
                    s_acc[i] = w;
                }
            }
            __syncthreads();
            float* per = pass ? s_per_p : s_per_c;
            float* ghw = pass ? s_gh_p : s_gh_c;
            const int ne = s_nonempty;
            for (int j = tid; j < M; j += blockDim.x) {
                float wgt = 0.f;
                const int bin = meta_bin(s_meta[j], pass);
                if (bin >= 0) {
                    wgt = s_acc[bin - 1];
                    if (ne > 0) wgt = wgt / (float)ne;
                }
                ghw[j] = wgt;
                per[j] *= wgt;
            }
            __syncthreads();
        }
        if (p.momentum > 0.f)
            for (int i = tid; i < nb; i += blockDim.x) p.ghm_acc[i] = s_accd[i];
    }

    // ------------------------------------------------------------------ phase D: actionness (cls_loss.py:299-339), per group
    for (int pass = 0; pass < 2; ++pass) {
        const float* xs = pass ? s_xp : s_x;
        const float* src = pass ? p.pact : p.act;
        float* dst = pass ? d_pact : d_act;
        if (!src) {
            for (int j = tid; j < M; j += blockDim.x) dst[j] = 0.f;
            if (tid == 0) p.losses[5 + pass] = 0.f;
            continue;
        }
        const int bit = pass ? 2 : 1;
        const float* g_np = pass ? g_nppos : g_npos;
        for (int j = tid; j < M; j += blockDim.x) {
            const int grp = NG == 1 ? 0 : j / L;
            const int topM = (int)fminf(g_np[grp], (float)L - g_np[grp]) - 1;
            const bool pos = (s_meta[j] & bit) != 0;
            const float x = xs[j];
            bool sel = true;
            if (topM > 0 && !pos) {
                int rank = 0;                      // number of the group's negatives that sort before this one (ascending, index breaks ties)
                const int lo = grp * L;
                for (int i = lo; i < lo + L; ++i) {
                    const bool ineg = (s_meta[i] & bit) == 0;
                    const float xi = xs[i];
                    rank += (ineg && (xi < x || (xi == x && i < j))) ? 1 : 0;
                }
                sel = rank < topM;
            }
            float g = 0.f, l = 0.f;
            if (sel) {
                const float t = pos ? 1.f : 0.f;
                l = bce_logits(x, t);
                g = sigmoidf(x) - t;
            }
            dst[j] = g;
            s_a[j] = l; s_b[j] = sel ? 1.f : 0.f;
        }
        __syncthreads();
        group_sums<2>(NG, L, g_sum, s_part, [&](int j, float* v) { v[0] = s_a[j]; v[1] = s_b[j]; });
        const bool ranked = p.act_weight != 0.f;
        if (ranked) {     // rank term: max(0, margin - max(neg) + max(pos).detach())
            group_argmax(NG, L, g_nmax, g_nmi, s_part, s_parti, [&](int j, float& x) { x = xs[j]; return (s_meta[j] & bit) == 0; });
            group_argmax(NG, L, g_pmax, g_pmi, s_part, s_parti, [&](int j, float& x) { x = xs[j]; return (s_meta[j] & bit) != 0; });
        }
        if (tid == 0) {
            float tot = 0.f, an = 0.f;
            for (int g = 0; g < NG; ++g) {
                const int topM = (int)fminf(g_np[g], (float)L - g_np[g]) - 1;
                float loss = g_sum[g];
                int arg = -1;
                if (ranked && topM > 0) {
                    const float r = p.act_margin - g_nmax[g] + g_pmax[g];
                    if (r > 0.f) { loss += p.act_weight * r; arg = g_nmi[g]; }
                }
                const float AN = fmaxf(g_sum[kMslMaxGroups + g], 1.f);
                g_AN[pass * kMslMaxGroups + g] = AN;
                g_nmi[g] = arg;
                tot += loss / AN; an += g_sum[kMslMaxGroups + g];
            }
            p.losses[5 + pass] = tot * invG;
            p.losses[9 + pass] = an;
        }
        __syncthreads();
        for (int j = tid; j < M; j += blockDim.x) {
            const int grp = NG == 1 ? 0 : j / L;
            float g = dst[j];
            if (ranked && j == g_nmi[grp]) g -= p.act_weight;
            dst[j] = g / g_AN[pass * kMslMaxGroups + grp] * invG;
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ phase E: weighted sums, calibration, gradient scaling
    const float invM = 1.f / (float)M;
    for (int j = tid; j < M; j += blockDim.x) {
        const int grp = NG == 1 ? 0 : j / L;
        const float N = fmaxf(g_npos[grp], 1.f) * (float)NG, PN = fmaxf(g_nppos[grp], 1.f) * (float)NG;   // incl. the batch mean
        const int meta = s_meta[j];
        const bool pos = meta & 1, ppos = (meta & 2) != 0;
        const int label = (meta >> 2) & 0x3ff;
        d_loc_l[2 * (size_t)j] /= N; d_loc_l[2 * (size_t)j + 1] /= N;
        d_loc_ct[2 * (size_t)j] /= N; d_loc_ct[2 * (size_t)j + 1] /= N;
        d_ploc_ct[2 * (size_t)j] /= N; d_ploc_ct[2 * (size_t)j + 1] /= N;
        d_center[j] /= N;
        d_ploc_l[2 * (size_t)j] /= PN; d_ploc_l[2 * (size_t)j + 1] /= PN;
        float cal = 0.f;
        if (focal) {
            float* dz = d_conf + (size_t)j * K;
            float* dzp = d_pconf + (size_t)j * K;
            for (int k = 0; k < K; ++k) { dz[k] /= N; dzp[k] /= PN; }
        } else {
            // coarse EDL gradient: w/N * (1/S - [k == y]/alpha_y) * d alpha_k / d z_k  (+ the ActivityNet weight's own gradient
            // through ||z||_1: per * dw/dz_k = -per_weighted * w * exp(c g) * sign(z_k))
            // one row: d/dz of  w * (log S - log alpha_y)  scaled by 1/norm, plus a per-row extra term on alpha (calibration)
            auto row_grad = [&](const float* z, float* dz, int y, bool inc, float w, float per_w, float norm, float cal_alpha) {
                float S = 0.f, ay = 1.f, amax = 0.f; int m = 0;
                for (int k = 0; k < K; ++k) {
                    const float a = expf(fminf(fmaxf(z[k], -10.f), 10.f)) + 1.f;
                    S += a; if (k == y) ay = a;
                    if (a > amax) { amax = a; m = k; }
                }
                const float wn = inc ? w / norm : 0.f;
                float through_norm = 0.f, fcoef = 0.f, pred = 0.f;
                if (inc && anet && p.use_ibm) {        // the ActivityNet weight's own gradient through ||z||_1
                    const float gn = fabsf(1.f / ay - (float)K / S);
                    through_norm = -(per_w / norm) * w * expf(p.ibm_coeff * gn);
                }
                if (inc && rw == kRwFocal) {           // the modulating factor (1 - max_k alpha_k / S)^gamma is NOT detached (:221-227)
                    pred = amax / S;
                    const float a_cls = y == 0 ? p.edl_alpha0 : 1.f - p.edl_alpha0;
                    fcoef = -((logf(S) - logf(ay)) / norm) * a_cls * p.edl_gamma * powf(1.f - pred, p.edl_gamma - 1.f) / S;
                }
                for (int k = 0; k < K; ++k) {
                    const float zk = z[k];
                    const float ev = expf(fminf(fmaxf(zk, -10.f), 10.f));
                    const float da = (zk >= -10.f && zk <= 10.f) ? ev : 0.f;
                    float g = 0.f;
                    if (inc) {
                        g = 1.f / S;
                        if (k == y) g -= 1.f / (ev + 1.f);
                        g = g * wn + fcoef * ((k == m ? 1.f : 0.f) - pred);
                    }
                    g += cal_alpha / S;                                       // caller passes cal_g * (-unc): d unc / d alpha_k = -unc / S
                    dz[k] = g * da + through_norm * (zk > 0.f ? 1.f : (zk < 0.f ? -1.f : 0.f));
                }
            };
            // coarse: w/N * (1/S - [k == y]/alpha_y) * d alpha_k / d z_k (+ the weight's own gradient where it has one)
            row_grad(p.conf + (size_t)j * K, d_conf + (size_t)j * K, cls_all ? label : label - 1, cls_all || pos, s_gh_c[j], s_per_c[j], N, 0.f);
            // refined EDL gradient + IoU-aware calibration over ALL priors (cls_loss.py:120-129, mean).  The THUMOS14 reference
            // flattens its [P,B] IoU buffer against [B*P] logits (multisegment_loss.py:116,146,236): element j pairs with the
            // IoU of prior j / B of sample j % B; the ActivityNet loss pairs them per sample (anet :258-260).
            float cal_g = 0.f;
            const float unc = s_unc[j];
            if (p.iou_aware) {
                float iou = anet ? s_iou[j] : s_iou[(size_t)(j % p.B) * p.P + (j / p.B)];
                if (iou < 0.f) iou = 1e-3f;
                cal = -iou * logf(1.f - unc) - (1.f - iou) * logf(unc);
                cal_g = (iou / (1.f - unc) - (1.f - iou) / unc) * invM;      // d mean(reg) / d unc_j
            }
            const int plabel = ppos ? label : 0;
            row_grad(p.pconf + (size_t)j * K, d_pconf + (size_t)j * K, cls_all ? plabel : plabel - 1, cls_all || ppos, s_gh_p[j], s_per_p[j], PN,
                     cal_g * (-unc));
        }
        s_a[j] = cal;
    }
    __syncthreads();
    group_sums<3>(NG, L, g_sum, s_part, [&](int j, float* v) { v[0] = s_per_c[j]; v[1] = s_per_p[j]; v[2] = s_a[j]; });
    if (tid == 0) {
        float c = 0.f, pc = 0.f, calsum = 0.f, npos = 0.f, nppos = 0.f;
        for (int g = 0; g < NG; ++g) {
            c += g_sum[g] / fmaxf(g_npos[g], 1.f);
            pc += g_sum[kMslMaxGroups + g] / fmaxf(g_nppos[g], 1.f);
            calsum += g_sum[2 * kMslMaxGroups + g];
            npos += g_npos[g]; nppos += g_nppos[g];
        }
        const float iouc = p.iou_aware ? calsum * invM : 0.f;
        p.losses[1] = c * invG;
        p.losses[3] = pc * invG + iouc;
        p.losses[7] = npos; p.losses[8] = nppos; p.losses[11] = iouc;
        if (!p.act) p.losses[9] = 0.f;
        if (!p.pact) p.losses[10] = 0.f;
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

static size_t msl_smem_bytes(int M) {
    return (size_t)M * 12 * 4 + (kMslMaxBins + 2 * kMslMaxBins + 4 * kMslWarps + kMslWarps + 12 * kMslMaxGroups + 2 * kMslMaxGroups) * 4 + 16;
}

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
    if (d->flavour < kMslThumos || d->flavour > kMslFocal) { set_last_error_msg("msl: flavour must be 0 (THUMOS14 EDL), 1 (ActivityNet EDL) or 2 (closed-set focal)"); return OTAL_ERR_BAD_ARG; }
    // re-weighting branch of the THUMOS14 EDL loss: use_ibm (the round-1 field) means kRwIbm unless `reweight` names another
    int rw = d->flavour == kMslThumos ? (d->reweight ? d->reweight : (d->use_ibm ? kRwIbm : kRwNone)) : kRwNone;
    if (rw < kRwNone || rw > kRwGhm) { set_last_error_msg("msl: reweight must be 0..4"); return OTAL_ERR_BAD_ARG; }
    const bool binned = rw == kRwIbm || rw == kRwGhm;
    if (rw == kRwIbm && (!d->weight_accum || d->num_bins <= 0 || d->num_bins > kMslMaxBins)) {
        set_last_error_msg("msl: IBM needs weight_accum and 1..256 bins"); return OTAL_ERR_BAD_ARG;
    }
    if (rw == kRwGhm && (!d->ghm_acc_sum || d->num_bins <= 0 || d->num_bins > kMslMaxBins - 2)) {
        set_last_error_msg("msl: GHM needs the fp64 acc_sum buffer and 1..254 bins"); return OTAL_ERR_BAD_ARG;
    }
    if (d->cls_all && (d->flavour != kMslThumos || d->act || d->prop_act)) {
        set_last_error_msg("msl: cls_all (no os_head) is a THUMOS14 EDL option without actionness inputs"); return OTAL_ERR_BAD_ARG;
    }
    if (d->flavour == kMslFocal && (d->act || d->prop_act || d->iou_aware || d->use_ibm)) {
        set_last_error_msg("msl: the closed-set focal flavour has no actionness head, IBM or IoU calibration"); return OTAL_ERR_BAD_ARG;
    }
    if (d->flavour == kMslAnet && d->prior_stride < 2) {
        set_last_error_msg("msl: the ActivityNet flavour reads the pyramid level from priors[p * prior_stride + 1]"); return OTAL_ERR_BAD_ARG;
    }
    if (d->K > 1023) { set_last_error_msg("msl: more than 1023 classes"); return OTAL_ERR_UNSUPPORTED; }
    const long long M = (long long)d->B * d->P;
    const size_t smem = msl_smem_bytes((int)M);
    if (M > (1 << 20) || smem > 227 * 1024 || (d->flavour == kMslAnet && d->B > kMslMaxGroups)) {
        set_last_error_msg("msl: B*P too large for the single-CTA kernel (shared-memory staging)"); return OTAL_ERR_UNSUPPORTED;
    }
    MslParams p{};
    p.B = d->B; p.P = d->P; p.K = d->K; p.G = d->G; p.M = (int)M;
    p.flavour = d->flavour;
    p.groups = d->flavour == kMslAnet ? d->B : 1;
    p.L = d->flavour == kMslAnet ? d->P : (int)M;
    p.clip = d->clip_length; p.thresh = d->overlap_thresh;
    p.use_ibm = d->use_ibm; p.num_bins = binned ? d->num_bins : 1; p.momentum = d->momentum;
    p.iou_aware = d->iou_aware; p.act_weight = d->act_weight; p.act_margin = d->act_margin;
    p.prior_stride = d->prior_stride > 0 ? d->prior_stride : 1;
    p.ibm_coeff = d->ibm_coeff; p.focal_alpha = d->focal_alpha; p.focal_gamma = d->focal_gamma;
    p.reweight = rw; p.cls_all = d->cls_all ? 1 : 0; p.edl_alpha0 = d->edl_focal_alpha; p.edl_gamma = d->edl_focal_gamma;
    p.ghm_acc = d->ghm_acc_sum;
    for (int i = 0; i < 16; ++i) p.bounds[i] = d->level_bounds[i];
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

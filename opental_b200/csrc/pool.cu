// MaxPool3dSamePadding on NDHWC bf16 hi/lo planes (forward) and its gradient routing (backward).
//
// Reference: AFSD/common/layers.py:9-35 — F.pad with constant ZERO ("same" rule, front = pad//2) followed by
// nn.MaxPool3d; used by the four stage pools of I3D and the b3a branch of every inception module
// (AFSD/common/i3d_backbone.py:104-105, :204, :224, :242, :279).  A window that reaches into the padding therefore
// competes against 0, which we reproduce exactly (inputs are post-ReLU, so it never changes the result).
//
// HBM-bound: one thread owns 8 channels (one 16-byte vector per plane) of one output position; a warp covers
// 256 consecutive channels, i.e. fully coalesced 512-byte rows in NDHWC.  Values are compared as hi + lo, which is
// exact in fp32 (8 + 8 significant bits), and the winning (hi, lo) pair is forwarded unchanged, so the pool is
// bit-exact on the represented values.
//
// Backward: the argmax of each window is recomputed (first maximum in (t,h,w) window order, strict '>', like
// ATen's max_pool3d) and the fp32 gradient is added to the fp32 input-gradient buffer with red.global.add.f32.
// Overlapping windows (k3 s1 / k3 s2) make the adds collide; sums have at most 27 terms.
#include "common.cuh"
#include <stdlib.h>

namespace otal {

struct PoolParams {
    int N, T, H, W, C;
    int To, Ho, Wo;
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int in_cstride, in_coff, out_cstride, out_coff;
    int gout_cstride, gout_coff, gin_cstride, gin_coff;
    const uint16_t *x_hi, *x_lo;
    uint16_t *y_hi, *y_lo;
    const float* g_out;
    float* g_in;
};

__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }

__device__ __forceinline__ void unpack8(const uint4& h, const uint4* l, float (&v)[8]) {
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = bf16_bits_to_float(hw[i] & 0xffffu);
        v[2 * i + 1] = bf16_bits_to_float(hw[i] >> 16);
    }
    if (l) {
        const uint32_t lw[4] = {l->x, l->y, l->z, l->w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] += bf16_bits_to_float(lw[i] & 0xffffu);
            v[2 * i + 1] += bf16_bits_to_float(lw[i] >> 16);
        }
    }
}

template <bool kBackward>
__global__ void __launch_bounds__(256)
maxpool_kernel(const PoolParams p) {
    const int cgs = p.C >> 3;
    const long long total = (long long)p.N * p.To * p.Ho * p.Wo * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int cg = (int)(idx % cgs);
        long long pos = idx / cgs;
        const int wo = (int)(pos % p.Wo); pos /= p.Wo;
        const int ho = (int)(pos % p.Ho); pos /= p.Ho;
        const int to = (int)(pos % p.To);
        const int n = (int)(pos / p.To);
        const int t0 = to * p.st - p.pt, h0 = ho * p.sh - p.ph, w0 = wo * p.sw - p.pw;
        const bool touches_pad = t0 < 0 || h0 < 0 || w0 < 0 || t0 + p.kt > p.T || h0 + p.kh > p.H || w0 + p.kw > p.W;

        float best[8];
        uint32_t bh[8], bl[8];
        long long barg[8];
        bool have = false;
        if (touches_pad) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { best[j] = 0.f; bh[j] = 0; bl[j] = 0; barg[j] = -1; }
            have = true;
        }
        for (int dt = 0; dt < p.kt; ++dt) {
            const int t = t0 + dt;
            if (t < 0 || t >= p.T) continue;
            for (int dh = 0; dh < p.kh; ++dh) {
                const int h = h0 + dh;
                if (h < 0 || h >= p.H) continue;
                for (int dw = 0; dw < p.kw; ++dw) {
                    const int w = w0 + dw;
                    if (w < 0 || w >= p.W) continue;
                    const long long ipos = (((long long)n * p.T + t) * p.H + h) * p.W + w;
                    const size_t off = (size_t)ipos * p.in_cstride + p.in_coff + cg * 8;
                    const uint4 hv = *reinterpret_cast<const uint4*>(p.x_hi + off);
                    uint4 lv = make_uint4(0, 0, 0, 0);
                    if (p.x_lo) lv = *reinterpret_cast<const uint4*>(p.x_lo + off);
                    float v[8];
                    unpack8(hv, p.x_lo ? &lv : nullptr, v);
                    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (!have || v[j] > best[j]) {
                            best[j] = v[j];
                            bh[j] = (j & 1) ? (hw[j >> 1] >> 16) : (hw[j >> 1] & 0xffffu);
                            bl[j] = (j & 1) ? (lw[j >> 1] >> 16) : (lw[j >> 1] & 0xffffu);
                            barg[j] = ipos;
                        }
                    }
                    have = true;
                }
            }
        }
        const long long opos = (((long long)n * p.To + to) * p.Ho + ho) * p.Wo + wo;
        if (!kBackward) {
            const size_t off = (size_t)opos * p.out_cstride + p.out_coff + cg * 8;
            *reinterpret_cast<uint4*>(p.y_hi + off) =
                make_uint4(bh[0] | (bh[1] << 16), bh[2] | (bh[3] << 16), bh[4] | (bh[5] << 16), bh[6] | (bh[7] << 16));
            if (p.y_lo)
                *reinterpret_cast<uint4*>(p.y_lo + off) =
                    make_uint4(bl[0] | (bl[1] << 16), bl[2] | (bl[3] << 16), bl[4] | (bl[5] << 16), bl[6] | (bl[7] << 16));
        } else {
            const float* g = p.g_out + (size_t)opos * p.gout_cstride + p.gout_coff + cg * 8;
            const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
            const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (barg[j] >= 0 && gv[j] != 0.f)
                    atomicAdd(p.g_in + (size_t)barg[j] * p.gin_cstride + p.gin_coff + cg * 8 + j, gv[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fast path (the five window shapes of I3D).  Two ideas on top of the generic kernel above:
//  * order-preserving integer keys: a value is the pair (hi, lo) of bf16 with |lo| <= ulp(hi)/2, so the order of the
//    represented values hi + lo is the lexicographic order of (hi, lo).  Mapping each 16-bit pattern b to
//    T(b) = b ^ 0xffff (negative) / b ^ 0x8000 (non-negative) makes that an unsigned 32-bit compare of
//    (T(hi) << 16) | T(lo): one max per candidate instead of unpack + add + compare + three selects, and the winning key
//    converts back to the exact (hi, lo) pair.
//  * WB consecutive outputs along W per thread share the loaded input columns ((WB-1)*SW + KW columns instead of
//    WB*KW), and the forward optionally records the arg-max (window-relative index, one byte per element; 255 = the
//    zero padding won) so that the backward is a pure scatter: read 8 gradients + 8 bytes, issue <= 8 red.add.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t key_fwd2(uint32_t w) {          // T() on both 16-bit halves
    const uint32_t s = (w >> 15) & 0x00010001u;
    return w ^ ((s * 0xffffu) | 0x80008000u);
}
__device__ __forceinline__ uint32_t key_inv2(uint32_t w) {
    const uint32_t m = ((w >> 15) & 0x00010001u) ^ 0x00010001u;
    return w ^ ((m * 0xffffu) | 0x80008000u);
}

template <int KT, int KH, int KW, int ST, int SH, int SW, int WB>
__global__ void __launch_bounds__(128)
maxpool_fwd_fast_kernel(const PoolParams p, unsigned char* __restrict__ argmax) {
    constexpr int NCOL = (WB - 1) * SW + KW;
    constexpr uint32_t KEY_ZERO = 0x80008000u;                       // key of +0.0 (hi = lo = +0)
    const int cgs = p.C >> 3;
    const int wblocks = (p.Wo + WB - 1) / WB;
    const long long total = (long long)p.N * p.To * p.Ho * wblocks * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int cg = (int)(idx % cgs);
        long long r = idx / cgs;
        const int wbk = (int)(r % wblocks); r /= wblocks;
        const int ho = (int)(r % p.Ho); r /= p.Ho;
        const int to = (int)(r % p.To);
        const int n = (int)(r / p.To);
        const int t0 = to * ST - p.pt, h0 = ho * SH - p.ph, wi0 = wbk * WB * SW - p.pw;
        const bool th_pad = t0 < 0 || h0 < 0 || t0 + KT > p.T || h0 + KH > p.H;
        uint32_t best[WB][8];
        uint32_t arg[WB][8];
#pragma unroll
        for (int o = 0; o < WB; ++o) {
            const int w0 = wi0 + o * SW;
            const bool pad = th_pad || w0 < 0 || w0 + KW > p.W;
#pragma unroll
            for (int j = 0; j < 8; ++j) { best[o][j] = pad ? KEY_ZERO : 0u; arg[o][j] = 255u; }
        }
#pragma unroll
        for (int dt = 0; dt < KT; ++dt) {
            const int t = t0 + dt;
            if (t < 0 || t >= p.T) continue;
#pragma unroll
            for (int dh = 0; dh < KH; ++dh) {
                const int h = h0 + dh;
                if (h < 0 || h >= p.H) continue;
                const long long rowpos = (((long long)n * p.T + t) * p.H + h) * p.W;
#pragma unroll
                for (int c = 0; c < NCOL; ++c) {
                    const int w = wi0 + c;
                    if (w < 0 || w >= p.W) continue;
                    const size_t off = (size_t)(rowpos + w) * p.in_cstride + p.in_coff + cg * 8;
                    const uint4 hv = *reinterpret_cast<const uint4*>(p.x_hi + off);
                    uint4 lv = make_uint4(0, 0, 0, 0);
                    if (p.x_lo) lv = *reinterpret_cast<const uint4*>(p.x_lo + off);
                    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
                    uint32_t key[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t th = key_fwd2(hw[i]), tl = key_fwd2(lw[i]);
                        key[2 * i] = __byte_perm(tl, th, 0x5410);
                        key[2 * i + 1] = __byte_perm(tl, th, 0x7632);
                    }
#pragma unroll
                    for (int o = 0; o < WB; ++o) {
                        const int dw = c - o * SW;                   // compile-time after unrolling
                        if (dw < 0 || dw >= KW) continue;
                        const uint32_t code = (uint32_t)((dt * KH + dh) * KW + dw);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (key[j] > best[o][j]) { best[o][j] = key[j]; arg[o][j] = code; }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < WB; ++o) {
            const int wo = wbk * WB + o;
            if (wo >= p.Wo) continue;
            const long long opos = (((long long)n * p.To + to) * p.Ho + ho) * p.Wo + wo;
            uint32_t oh[4], ol[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t ka = best[o][2 * i] ? best[o][2 * i] : KEY_ZERO, kb = best[o][2 * i + 1] ? best[o][2 * i + 1] : KEY_ZERO;
                oh[i] = key_inv2(__byte_perm(ka, kb, 0x7632));
                ol[i] = key_inv2(__byte_perm(ka, kb, 0x5410));
            }
            const size_t off = (size_t)opos * p.out_cstride + p.out_coff + cg * 8;
            *reinterpret_cast<uint4*>(p.y_hi + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
            if (p.y_lo) *reinterpret_cast<uint4*>(p.y_lo + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
            if (argmax) {
                const uint32_t a0 = arg[o][0] | (arg[o][1] << 8) | (arg[o][2] << 16) | (arg[o][3] << 24);
                const uint32_t a1 = arg[o][4] | (arg[o][5] << 8) | (arg[o][6] << 16) | (arg[o][7] << 24);
                *reinterpret_cast<uint2*>(argmax + (size_t)opos * p.C + cg * 8) = make_uint2(a0, a1);
            }
        }
    }
}

// Backward from the recorded arg-max: pure scatter.
__global__ void __launch_bounds__(256)
maxpool_bwd_argmax_kernel(const PoolParams p, const unsigned char* __restrict__ argmax) {
    const int cgs = p.C >> 3;
    const long long total = (long long)p.N * p.To * p.Ho * p.Wo * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int khw = p.kh * p.kw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int cg = (int)(idx % cgs);
        long long pos = idx / cgs;
        const long long opos = pos;
        const int wo = (int)(pos % p.Wo); pos /= p.Wo;
        const int ho = (int)(pos % p.Ho); pos /= p.Ho;
        const int to = (int)(pos % p.To);
        const int n = (int)(pos / p.To);
        const uint2 a = *reinterpret_cast<const uint2*>(argmax + (size_t)opos * p.C + cg * 8);
        const float* g = p.g_out + (size_t)opos * p.gout_cstride + p.gout_coff + cg * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const int t0 = to * p.st - p.pt, h0 = ho * p.sh - p.ph, w0 = wo * p.sw - p.pw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t code = ((j < 4 ? a.x : a.y) >> ((j & 3) * 8)) & 0xffu;
            if (code == 255u || gv[j] == 0.f) continue;
            const int dt = (int)code / khw, rem = (int)code - dt * khw, dh = rem / p.kw, dw = rem - dh * p.kw;
            const long long ipos = (((long long)n * p.T + (t0 + dt)) * p.H + (h0 + dh)) * p.W + (w0 + dw);
            atomicAdd(p.g_in + (size_t)ipos * p.gin_cstride + p.gin_coff + cg * 8 + j, gv[j]);
        }
    }
}

// (3,3,3) / stride 1 (the b3a pool of every inception module): each input is needed by 27 windows, so the W-blocked
// kernel above re-reads it ~13x from L2 (measured ~8 TB/s of L2 traffic, profiles/r01_pool_bench.txt).  Here one CTA
// stages the integer keys of a (tT+2) x (tH+2) x (tW+2) halo tile for 64 channels in shared memory (each input is
// fetched ~2.7x instead) and every thread scans its windows out of shared memory with two conflict-free 16-byte loads
// per tap.  Shared layout per position: 256 bytes = [8 channel groups x 4 keys (channels 0-3)][8 groups x 4 keys (4-7)].
struct PoolTile { int tT, tH, tW, tilesT, tilesH, tilesW, cchunks; };

__global__ void __launch_bounds__(320)
maxpool333_tiled_kernel(const PoolParams p, const PoolTile tl, unsigned char* __restrict__ argmax) {
    extern __shared__ __align__(16) unsigned char pool_smem[];
    uint4* keys = reinterpret_cast<uint4*>(pool_smem);                 // [halo positions][2 halves][8 cgs] x 16 B
    constexpr uint32_t KEY_ZERO = 0x80008000u;
    const int cgs = p.C >> 3;
    int b = blockIdx.x;
    const int cchunk = b % tl.cchunks; b /= tl.cchunks;
    const int iw = b % tl.tilesW; b /= tl.tilesW;
    const int ih = b % tl.tilesH; b /= tl.tilesH;
    const int it = b % tl.tilesT;
    const int n = b / tl.tilesT;
    const int t0 = it * tl.tT, h0 = ih * tl.tH, w0 = iw * tl.tW;      // first output of the tile (= input coordinate)
    const int eT = tl.tT + 2, eH = tl.tH + 2, eW = tl.tW + 2;
    const int halo = eT * eH * eW;
    // ---- stage keys: item = (halo position, channel group).  Four items per thread and round, all loads issued before the
    // first conversion: with 2 CTAs per SM the global-load latency of this phase is otherwise exposed once per item.
    constexpr int kStageUnroll = 4;
    for (int i0 = threadIdx.x; i0 < halo * 8; i0 += blockDim.x * kStageUnroll) {
        uint4 hv[kStageUnroll], lv[kStageUnroll];
        bool ok[kStageUnroll];
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
            const int i = i0 + u * blockDim.x;
            const int cgl = i & 7, hp = i >> 3;
            const int cg = cchunk * 8 + cgl;
            const int xw = hp % eW, xh = (hp / eW) % eH, xt = hp / (eW * eH);
            const int t = t0 - 1 + xt, h = h0 - 1 + xh, w = w0 - 1 + xw;
            ok[u] = i < halo * 8 && cg < cgs && t >= 0 && t < p.T && h >= 0 && h < p.H && w >= 0 && w < p.W;
            hv[u] = make_uint4(0, 0, 0, 0); lv[u] = make_uint4(0, 0, 0, 0);
            if (ok[u]) {
                const size_t off = ((((size_t)n * p.T + t) * p.H + h) * p.W + w) * p.in_cstride + p.in_coff + cg * 8;
                hv[u] = *reinterpret_cast<const uint4*>(p.x_hi + off);
                if (p.x_lo) lv[u] = *reinterpret_cast<const uint4*>(p.x_lo + off);
            }
        }
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i >= halo * 8) break;
            const int cgl = i & 7, hp = i >> 3;
            uint4 k0 = make_uint4(KEY_ZERO, KEY_ZERO, KEY_ZERO, KEY_ZERO), k1 = k0;   // zero padding competes as 0
            if (ok[u]) {
                const uint32_t hw[4] = {hv[u].x, hv[u].y, hv[u].z, hv[u].w}, lw[4] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w};
                uint32_t key[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t th = key_fwd2(hw[j]), tl_ = key_fwd2(lw[j]);
                    key[2 * j] = __byte_perm(tl_, th, 0x5410);
                    key[2 * j + 1] = __byte_perm(tl_, th, 0x7632);
                }
                k0 = make_uint4(key[0], key[1], key[2], key[3]);
                k1 = make_uint4(key[4], key[5], key[6], key[7]);
            }
            keys[hp * 16 + cgl] = k0;
            keys[hp * 16 + 8 + cgl] = k1;
        }
    }
    __syncthreads();
    // ---- scan: item = (output row/column of the tile, channel group); the thread walks along T.  The 3x3 maximum of every
    // halo plane (first maximum in (dh, dw) order) is computed ONCE and merged over a sliding window of three planes, strict
    // '>' in dt order: the same result as scanning the 27 taps in (dt, dh, dw) order — final key = max(init, taps), arg = first
    // tap that attains it when it beats init — for (tT + 2) / tT x 9 taps per output instead of 27.
    const int ncol = tl.tH * tl.tW;
    for (int i = threadIdx.x; i < ncol * 8; i += blockDim.x) {
        const int cgl = i & 7, op = i >> 3;
        const int cg = cchunk * 8 + cgl;
        const int ow = op % tl.tW, oh = op / tl.tW;
        const int ho = h0 + oh, wo = w0 + ow;
        if (cg >= cgs || ho >= p.Ho || wo >= p.Wo) continue;
        const bool pad_hw = ho == 0 || wo == 0 || ho + 1 >= p.H || wo + 1 >= p.W;
        uint32_t pk[3][8];                       // plane maxima of the last three halo planes
        uint32_t pa[3];                          // their in-plane arg (dh * 3 + dw), one nibble per channel
#pragma unroll
        for (int s_ = 0; s_ < 3; ++s_) { pa[s_] = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[s_][j] = 0; }
        for (int xt = 0; xt < eT; ++xt) {
            // shift the window and scan plane xt
#pragma unroll
            for (int j = 0; j < 8; ++j) { pk[0][j] = pk[1][j]; pk[1][j] = pk[2][j]; }
            pa[0] = pa[1]; pa[1] = pa[2];
            uint32_t bk[8], ba[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { bk[j] = 0u; ba[j] = 0u; }
#pragma unroll
            for (int dh = 0; dh < 3; ++dh)
#pragma unroll
                for (int dw = 0; dw < 3; ++dw) {
                    const int hp = (xt * eH + (oh + dh)) * eW + (ow + dw);
                    const uint4 a = keys[hp * 16 + cgl], c = keys[hp * 16 + 8 + cgl];
                    const uint32_t k[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (k[j] > bk[j]) { bk[j] = k[j]; ba[j] = (uint32_t)(dh * 3 + dw); }
                }
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { pk[2][j] = bk[j]; packed |= ba[j] << (4 * j); }
            pa[2] = packed;
            const int ot = xt - 2;
            if (ot < 0) continue;
            const int to = t0 + ot;
            if (to >= p.To) break;
            const bool pad = pad_hw || to == 0 || to + 1 >= p.T;
            uint32_t best[8], arg[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { best[j] = pad ? KEY_ZERO : 0u; arg[j] = 255u; }
#pragma unroll
            for (int dt = 0; dt < 3; ++dt)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (pk[dt][j] > best[j]) { best[j] = pk[dt][j]; arg[j] = (uint32_t)(dt * 9) + ((pa[dt] >> (4 * j)) & 0xfu); }
            const long long opos = (((long long)n * p.To + to) * p.Ho + ho) * p.Wo + wo;
            uint32_t oh_[4], ol_[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ka = best[2 * j] ? best[2 * j] : KEY_ZERO, kb = best[2 * j + 1] ? best[2 * j + 1] : KEY_ZERO;
                oh_[j] = key_inv2(__byte_perm(ka, kb, 0x7632));
                ol_[j] = key_inv2(__byte_perm(ka, kb, 0x5410));
            }
            const size_t off = (size_t)opos * p.out_cstride + p.out_coff + cg * 8;
            *reinterpret_cast<uint4*>(p.y_hi + off) = make_uint4(oh_[0], oh_[1], oh_[2], oh_[3]);
            if (p.y_lo) *reinterpret_cast<uint4*>(p.y_lo + off) = make_uint4(ol_[0], ol_[1], ol_[2], ol_[3]);
            if (argmax) {
                const uint32_t a0 = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
                const uint32_t a1 = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
                *reinterpret_cast<uint2*>(argmax + (size_t)opos * p.C + cg * 8) = make_uint2(a0, a1);
            }
        }
    }
}

static int launch_pool333_tiled(const PoolParams& p, unsigned char* argmax, cudaStream_t s) {
    PoolTile tl;
    static const int tt_env = getenv("OTAL_POOL_TT") ? atoi(getenv("OTAL_POOL_TT")) : 4;      // developer A/B: T extent of a tile
    tl.tH = p.H < 6 ? p.H : 6; tl.tW = p.W < 6 ? p.W : 6; tl.tT = p.T < tt_env ? p.T : tt_env;
    tl.tilesT = (p.T + tl.tT - 1) / tl.tT; tl.tilesH = (p.H + tl.tH - 1) / tl.tH; tl.tilesW = (p.W + tl.tW - 1) / tl.tW;
    tl.cchunks = ((p.C >> 3) + 7) / 8;
    const size_t smem = (size_t)(tl.tT + 2) * (tl.tH + 2) * (tl.tW + 2) * 256;
    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(maxpool333_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        once.mark(once_dev);
    }
    const long long blocks = (long long)p.N * tl.tilesT * tl.tilesH * tl.tilesW * tl.cchunks;
    if (smem > 200 * 1024) { set_last_error_msg("maxpool: tile does not fit shared memory"); return OTAL_ERR_UNSUPPORTED; }
    const int threads = (tl.tH * tl.tW * 8 + 31) / 32 * 32;          // one scan item (output column, channel group) per thread
    maxpool333_tiled_kernel<<<(unsigned)blocks, threads > 320 ? 320 : threads, smem, s>>>(p, tl, argmax);
    return OTAL_OK;
}

// Backward of a pool FUSED with the ReLU / frozen-BN backward of the layer that produced the pool's input, as a gather:
// every input position sums the gradients of the (few) windows whose recorded arg-max points at it, adds an optional
// extra gradient (another consumer of the same tensor), applies d = g * [y > 0] * scale[c] and writes the bf16 (hi, lo)
// planes the tensor-core dgrad / wgrad kernels read.  No atomics, no zero-initialised fp32 gradient buffer, no separate
// relu_bn_bwd_split pass: the fp32 input gradient of the pool never touches HBM.  Used for the stride-2 stage pools
// (<= 8 windows per input); the stride-1 inception pools (27 windows per input) keep the scatter form.
struct PoolFuse {
    const unsigned char* argmax;     // [N,To,Ho,Wo,C]
    const float* g_add;              // optional fp32 [N,T,H,W,add_cstride] gradient to add (NULL = none)
    int add_cstride, add_coff;
    const uint16_t* y_hi;            // forward value of the pool input (ReLU mask), planes layout of x
    const float* scale;              // [C] folded BN scale or NULL
    uint16_t *d_hi, *d_lo;           // output planes [N,T,H,W,d_cstride]
    int d_cstride, d_coff;
};

__global__ void __launch_bounds__(256)
maxpool_bwd_gather_fused_kernel(const PoolParams p, const PoolFuse f) {
    const int cgs = p.C >> 3;
    const long long total = (long long)p.N * p.T * p.H * p.W * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int cg = (int)(idx % cgs);
        long long pos = idx / cgs;
        const long long ipos = pos;
        const int w = (int)(pos % p.W); pos /= p.W;
        const int h = (int)(pos % p.H); pos /= p.H;
        const int t = (int)(pos % p.T);
        const int n = (int)(pos / p.T);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        // windows (to, ho, wo) that contain this input: o*s - pad <= x <= o*s - pad + k - 1
        const int to_lo = max(0, (t + p.pt - p.kt + p.st) / p.st), to_hi = min(p.To - 1, (t + p.pt) / p.st);
        const int ho_lo = max(0, (h + p.ph - p.kh + p.sh) / p.sh), ho_hi = min(p.Ho - 1, (h + p.ph) / p.sh);
        const int wo_lo = max(0, (w + p.pw - p.kw + p.sw) / p.sw), wo_hi = min(p.Wo - 1, (w + p.pw) / p.sw);
        for (int to = to_lo; to <= to_hi; ++to)
            for (int ho = ho_lo; ho <= ho_hi; ++ho)
                for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                    const uint32_t code = (uint32_t)(((t - (to * p.st - p.pt)) * p.kh + (h - (ho * p.sh - p.ph))) * p.kw +
                                                     (w - (wo * p.sw - p.pw)));
                    const long long opos = (((long long)n * p.To + to) * p.Ho + ho) * p.Wo + wo;
                    const uint2 a = *reinterpret_cast<const uint2*>(f.argmax + (size_t)opos * p.C + cg * 8);
                    const float* g = p.g_out + (size_t)opos * p.gout_cstride + p.gout_coff + cg * 8;
                    const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
                    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t cj = ((j < 4 ? a.x : a.y) >> ((j & 3) * 8)) & 0xffu;
                        if (cj == code) acc[j] += gv[j];
                    }
                }
        if (f.g_add) {
            const float* ga = f.g_add + (size_t)ipos * f.add_cstride + f.add_coff + cg * 8;
            const float4 a0 = *reinterpret_cast<const float4*>(ga), a1 = *reinterpret_cast<const float4*>(ga + 4);
            acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
            acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
        }
        const uint4 yv = *reinterpret_cast<const uint4*>(f.y_hi + (size_t)ipos * p.in_cstride + p.in_coff + cg * 8);
        const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
        uint32_t ho_[4], lo_[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x0 = acc[2 * j], x1 = acc[2 * j + 1];
            const uint32_t y0 = yw[j] & 0xffffu, y1 = yw[j] >> 16;
            if (!(y0 != 0 && y0 < 0x8000u)) x0 = 0.f;             // y > 0 <=> sign clear and magnitude non-zero
            if (!(y1 != 0 && y1 < 0x8000u)) x1 = 0.f;
            if (f.scale) { x0 *= __ldg(f.scale + cg * 8 + 2 * j); x1 *= __ldg(f.scale + cg * 8 + 2 * j + 1); }
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(x0, h0, l0); split_bf16(x1, h1, l1);
            ho_[j] = pack_bf16x2(h0, h1); lo_[j] = pack_bf16x2(l0, l1);
        }
        const size_t off = (size_t)ipos * f.d_cstride + f.d_coff + cg * 8;
        *reinterpret_cast<uint4*>(f.d_hi + off) = make_uint4(ho_[0], ho_[1], ho_[2], ho_[3]);
        if (f.d_lo) *reinterpret_cast<uint4*>(f.d_lo + off) = make_uint4(lo_[0], lo_[1], lo_[2], lo_[3]);
    }
}

// The same gather with the window geometry known at compile time (the three stage pools of I3D): at most NT x NH x NW windows
// contain an input position; their arg-max bytes and gradients are ALL loaded before the first compare (the generic kernel's
// data-dependent loop bounds serialise 2-8 dependent L2 round trips per thread: measured 3.2-3.8x off the HBM roofline).
template <int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(128)
maxpool_bwd_gather_fused_t_kernel(const PoolParams p, const PoolFuse f) {
    constexpr int NT = (KT + ST - 1) / ST, NH = (KH + SH - 1) / SH, NW = (KW + SW - 1) / SW;
    constexpr int NWIN = NT * NH * NW;
    const int cgs = p.C >> 3;
    // one CTA per input row (n, t, h): the row decomposition and the T / H window ranges are uniform per CTA and 32-bit — the
    // per-thread 64-bit div / mod chain of the generic kernel costs more issue slots than the memory system needs time
    const int rows = p.N * p.T * p.H, items = p.W * cgs;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
      const int h = row % p.H, nt = row / p.H, t = nt % p.T, n = nt / p.T;
      const int to_lo = max(0, (t + p.pt - KT + ST) / ST), to_hi = min(p.To - 1, (t + p.pt) / ST);
      const int ho_lo = max(0, (h + p.ph - KH + SH) / SH), ho_hi = min(p.Ho - 1, (h + p.ph) / SH);
      for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int w = i / cgs, cg = i - w * cgs;
        const long long ipos = (long long)row * p.W + w;
        // windows (to, ho, wo) that contain this input: o*s - pad <= x <= o*s - pad + k - 1, ascending order
        const int wo_lo = max(0, (w + p.pw - KW + SW) / SW), wo_hi = min(p.Wo - 1, (w + p.pw) / SW);
        uint2 a[NWIN];
        float4 g0[NWIN], g1[NWIN];
        uint32_t code[NWIN];
#pragma unroll
        for (int it = 0; it < NT; ++it)
#pragma unroll
            for (int ih = 0; ih < NH; ++ih)
#pragma unroll
                for (int iw = 0; iw < NW; ++iw) {
                    const int k = (it * NH + ih) * NW + iw;
                    const int to = to_lo + it, ho = ho_lo + ih, wo = wo_lo + iw;
                    const bool ok = to <= to_hi && ho <= ho_hi && wo <= wo_hi;
                    code[k] = 0xffffffffu;                       // matches no arg-max byte
                    a[k] = make_uint2(0, 0);
                    g0[k] = make_float4(0.f, 0.f, 0.f, 0.f); g1[k] = g0[k];
                    if (ok) {
                        code[k] = (uint32_t)(((t - (to * ST - p.pt)) * KH + (h - (ho * SH - p.ph))) * KW + (w - (wo * SW - p.pw)));
                        const long long opos = (((long long)n * p.To + to) * p.Ho + ho) * p.Wo + wo;
                        a[k] = __ldg(reinterpret_cast<const uint2*>(f.argmax + (size_t)opos * p.C + cg * 8));
                        const float* g = p.g_out + (size_t)opos * p.gout_cstride + p.gout_coff + cg * 8;
                        g0[k] = __ldg(reinterpret_cast<const float4*>(g)); g1[k] = __ldg(reinterpret_cast<const float4*>(g + 4));
                    }
                }
        const uint4 yv = __ldg(reinterpret_cast<const uint4*>(f.y_hi + (size_t)ipos * p.in_cstride + p.in_coff + cg * 8));
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (f.g_add) {
            const float* ga = f.g_add + (size_t)ipos * f.add_cstride + f.add_coff + cg * 8;
            a0 = __ldg(reinterpret_cast<const float4*>(ga)); a1 = __ldg(reinterpret_cast<const float4*>(ga + 4));
        }
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < NWIN; ++k) {
            const float gv[8] = {g0[k].x, g0[k].y, g0[k].z, g0[k].w, g1[k].x, g1[k].y, g1[k].z, g1[k].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t cj = ((j < 4 ? a[k].x : a[k].y) >> ((j & 3) * 8)) & 0xffu;
                if (cj == code[k]) acc[j] += gv[j];
            }
        }
        acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
        acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
        const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
        uint32_t ho_[4], lo_[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x0 = acc[2 * j], x1 = acc[2 * j + 1];
            const uint32_t y0 = yw[j] & 0xffffu, y1 = yw[j] >> 16;
            if (!(y0 != 0 && y0 < 0x8000u)) x0 = 0.f;             // y > 0 <=> sign clear and magnitude non-zero
            if (!(y1 != 0 && y1 < 0x8000u)) x1 = 0.f;
            if (f.scale) { x0 *= __ldg(f.scale + cg * 8 + 2 * j); x1 *= __ldg(f.scale + cg * 8 + 2 * j + 1); }
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(x0, h0, l0); split_bf16(x1, h1, l1);
            ho_[j] = pack_bf16x2(h0, h1); lo_[j] = pack_bf16x2(l0, l1);
        }
        const size_t off = (size_t)ipos * f.d_cstride + f.d_coff + cg * 8;
        *reinterpret_cast<uint4*>(f.d_hi + off) = make_uint4(ho_[0], ho_[1], ho_[2], ho_[3]);
        if (f.d_lo) *reinterpret_cast<uint4*>(f.d_lo + off) = make_uint4(lo_[0], lo_[1], lo_[2], lo_[3]);
      }
    }
}

template <int KT, int KH, int KW, int ST, int SH, int SW, int WB>
static void launch_pool_fast(const PoolParams& p, unsigned char* argmax, cudaStream_t s) {
    const int wblocks = (p.Wo + WB - 1) / WB;
    const long long total = (long long)p.N * p.To * p.Ho * wblocks * (p.C >> 3);
    long long b = (total + 127) / 128;
    const long long cap = 148LL * 32;
    maxpool_fwd_fast_kernel<KT, KH, KW, ST, SH, SW, WB><<<(int)(b < 1 ? 1 : (b > cap ? cap : b)), 128, 0, s>>>(p, argmax);
}

static int fill_pool(const otal_pool_desc* d, PoolParams& p, bool backward) {
    if (!d) { set_last_error_msg("pool: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (d->N <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->C % 8 || d->in_cstride % 8 || d->in_coff % 8) {
        set_last_error_msg("pool: bad dimension (channels / strides / offsets must be multiples of 8)"); return OTAL_ERR_BAD_ARG;
    }
    if (d->kt < 1 || d->kh < 1 || d->kw < 1 || d->st < 1 || d->sh < 1 || d->sw < 1) {
        set_last_error_msg("pool: bad window"); return OTAL_ERR_BAD_ARG;
    }
    if (!d->x_hi && !(backward && d->argmax)) { set_last_error_msg("pool: null input"); return OTAL_ERR_BAD_ARG; }
    if (!backward && (!d->y_hi || d->out_cstride % 8 || d->out_coff % 8 || (d->x_lo && !d->y_lo))) {
        set_last_error_msg("pool: bad output"); return OTAL_ERR_BAD_ARG;
    }
    if (backward && (!d->g_out || !d->g_in || d->gout_cstride % 4 || d->gout_coff % 4)) {
        set_last_error_msg("pool: bad gradient buffers"); return OTAL_ERR_BAD_ARG;
    }
    p.N = d->N; p.T = d->T; p.H = d->H; p.W = d->W; p.C = d->C;
    p.kt = d->kt; p.kh = d->kh; p.kw = d->kw; p.st = d->st; p.sh = d->sh; p.sw = d->sw;
    p.pt = d->pt; p.ph = d->ph; p.pw = d->pw;
    p.To = (d->T + d->st - 1) / d->st; p.Ho = (d->H + d->sh - 1) / d->sh; p.Wo = (d->W + d->sw - 1) / d->sw;
    p.in_cstride = d->in_cstride; p.in_coff = d->in_coff; p.out_cstride = d->out_cstride; p.out_coff = d->out_coff;
    p.gout_cstride = d->gout_cstride; p.gout_coff = d->gout_coff; p.gin_cstride = d->gin_cstride; p.gin_coff = d->gin_coff;
    p.x_hi = d->x_hi; p.x_lo = d->x_lo; p.y_hi = d->y_hi; p.y_lo = d->y_lo; p.g_out = d->g_out; p.g_in = d->g_in;
    return OTAL_OK;
}

static int pool_grid(const PoolParams& p) {
    const long long total = (long long)p.N * p.To * p.Ho * p.Wo * (p.C >> 3);
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_maxpool_fwd(const otal_pool_desc* d, void* stream) {
    PoolParams p{};
    int rc = fill_pool(d, p, false);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto is = [&](int kt, int kh, int kw, int st, int sh, int sw) {
        return p.kt == kt && p.kh == kh && p.kw == kw && p.st == st && p.sh == sh && p.sw == sw;
    };
    // WB = 2 measured best at batch 8 (profiles/r01_pool_bench.txt): these launches are bound by L2 re-reads of the
    // 27 window positions, WB = 4 shares more columns but halves the warps in flight
    if (is(3, 3, 3, 1, 1, 1) && p.pt == 1 && p.ph == 1 && p.pw == 1) { if ((rc = launch_pool333_tiled(p, d->argmax, s))) return rc; }
    else if (is(3, 3, 3, 1, 1, 1)) launch_pool_fast<3, 3, 3, 1, 1, 1, 2>(p, d->argmax, s);
    else if (is(1, 3, 3, 1, 2, 2)) launch_pool_fast<1, 3, 3, 1, 2, 2, 2>(p, d->argmax, s);
    else if (is(3, 3, 3, 2, 2, 2)) launch_pool_fast<3, 3, 3, 2, 2, 2, 2>(p, d->argmax, s);
    else if (is(2, 2, 2, 2, 2, 2)) launch_pool_fast<2, 2, 2, 2, 2, 2, 2>(p, d->argmax, s);
    else {
        if (d->argmax) { set_last_error_msg("pool: arg-max recording is implemented for the I3D window shapes only"); return OTAL_ERR_UNSUPPORTED; }
        maxpool_kernel<false><<<pool_grid(p), 256, 0, s>>>(p);
    }
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

int otal_maxpool_bwd(const otal_pool_desc* d, void* stream) {
    PoolParams p{};
    int rc = fill_pool(d, p, true);
    if (rc) return rc;
    if (d->argmax) maxpool_bwd_argmax_kernel<<<pool_grid(p), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, d->argmax);
    else maxpool_kernel<true><<<pool_grid(p), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}


int otal_maxpool_bwd_relu_bn_split(const otal_pool_desc* d, const float* g_add, int add_cstride, int add_coff, const float* scale,
                                   uint16_t* d_hi, uint16_t* d_lo, int d_cstride, int d_coff, void* stream) {
    if (!d || !d->argmax || !d->g_out || !d->x_hi || !d_hi || d_cstride % 8 || d_coff % 8 || (g_add && (add_cstride % 4 || add_coff % 4))) {
        set_last_error_msg("maxpool_bwd_relu_bn_split: needs the recorded arg-max, g_out, the forward input planes and d planes");
        return OTAL_ERR_BAD_ARG;
    }
    otal_pool_desc dd = *d;
    float dummy = 0.f;
    dd.g_in = &dummy;                       // fill_pool insists on a gradient destination; the fused kernel has none
    PoolParams p{};
    int rc = fill_pool(&dd, p, true);
    if (rc) return rc;
    PoolFuse f{};
    f.argmax = d->argmax; f.g_add = g_add; f.add_cstride = add_cstride; f.add_coff = add_coff; f.y_hi = d->x_hi; f.scale = scale;
    f.d_hi = d_hi; f.d_lo = d_lo; f.d_cstride = d_cstride; f.d_coff = d_coff;
    const long long total = (long long)p.N * p.T * p.H * p.W * (p.C >> 3);
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 16;
    const int grid = (int)(b < 1 ? 1 : (b > cap ? cap : b));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto is = [&](int kt, int kh, int kw, int st, int sh, int sw) {
        return p.kt == kt && p.kh == kh && p.kw == kw && p.st == st && p.sh == sh && p.sw == sw;
    };
    static const bool generic = getenv("OTAL_POOL_BWD_GENERIC") != nullptr;       // developer A/B
    const int rows = p.N * p.T * p.H;
    const int rgrid = rows < 148 * 16 ? rows : 148 * 16;                        // 16 CTAs of 128 threads per SM
    if (!generic && is(1, 3, 3, 1, 2, 2)) maxpool_bwd_gather_fused_t_kernel<1, 3, 3, 1, 2, 2><<<rgrid, 128, 0, s>>>(p, f);
    else if (!generic && is(3, 3, 3, 2, 2, 2)) maxpool_bwd_gather_fused_t_kernel<3, 3, 3, 2, 2, 2><<<rgrid, 128, 0, s>>>(p, f);
    else if (!generic && is(2, 2, 2, 2, 2, 2)) maxpool_bwd_gather_fused_t_kernel<2, 2, 2, 2, 2, 2><<<rgrid, 128, 0, s>>>(p, f);
    else maxpool_bwd_gather_fused_kernel<<<grid, 256, 0, s>>>(p, f);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

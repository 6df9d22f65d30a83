// Weight gradient of the implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces torch's convolution_backward (weight part) for Unit3D / Unit1D
// (AFSD/common/i3d_backbone.py:82, AFSD/common/layers.py:211):
//   dW[tap][co][ci] = sum_{n, p}  D[n, p, co] * X[n, s*p + tap - pad, ci]
// with D = gradient w.r.t. the conv output (already multiplied by the ReLU mask and the folded-BN scale,
// see relu_bn_bwd_split) and X the saved conv input, both NDHWC bf16 hi/lo planes.
//
// GEMM view per tap:  M = Cout (tile 128), N = Cin (tile <= 256), K = output positions.  In NDHWC both operands
// have the *channel* (M resp. N) dimension contiguous, i.e. they are "MN-major" for the tensor core: the TMA boxes
// [64 positions x 64 channels] land in shared memory as 128-byte-swizzled rows indexed by K, which is exactly the
// canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units — no transposition
// anywhere.  Padding is the TMA zero fill of the shifted X box; stride-2 convs read parity-split views of X.
//
// One work item = (tap, 128-row block of Cout, N block of Cin, K split).  The K split exists because K (positions,
// up to 2.4 M at batch 8) is the long dimension while M x N x taps can be as small as 64 x 64 x 49; partial sums
// are added to dW with fp32 reductions (red.global.add.f32), dW must be zeroed (or hold the running gradient) on
// entry.  bf16x3: D_hi*X_hi + D_lo*X_hi + D_hi*X_lo with fp32 accumulation in TMEM.
//
// Warp roles (256 threads, persistent over work items): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 epilogue (tcgen05.ld -> red.add).
#include "common.cuh"
#include "tensormap.h"

namespace otal {

constexpr int kWgThreads = 256;
constexpr int kWgKP = 64;                 // positions per pipeline stage
constexpr int kWgBox = kWgKP * 128;       // bytes of one [64 positions x 64 channels] bf16 box = 8 KB
constexpr int kWgMaxStages = 6;

struct WgParams {
    int N, To, Ho, Wo;        // output (D) extent
    int Cin, Cout;            // Cin = weight row width actually accumulated (64 for the folded conv1a)
    int kt, kh, kw, pt, ph, pw, st, sh, sw;
    int tT, tH, tW, tilesT, tilesH, tilesW;
    int BN, n_blocks, m_blocks, ksplit, ktiles;   // ktiles = N * tilesT*tilesH*tilesW
    int BNs;                  // swap mode: MMA N = Cout rounded up to 16
    int x32;                  // X operand rows are 32 elements = 64 bytes (folded Conv3d_1a), SWIZZLE_64B boxes of 4 KB
    int swap;                 // x32 only: the 4 taps' X tiles are the M operand (4 x 32 rows), D is the N operand (Cout <= 128
                              // columns): one MMA covers 4 taps, so the D tile is read from shared memory once per 4 taps
                              // instead of once per tap (the per-tap form is bound by those reads: SMEM 128 B/clk)
    int G, ngroups, cstride;  // taps per work item (they share the D tile in shared memory), number of tap groups, TMEM
                              // column stride between the accumulators of consecutive taps (BN rounded up to 32)
    int nsplit, nstages;
    int total_items;
    int x_single;             // XS instantiation (swap mode, Cout == 64): X is ONE exact bf16 plane (raw uint8 pixels) and D keeps
                              // hi/lo; the D_lo box sits right behind the D_hi box, so ONE MMA of N = 128 computes
                              // X * [D_hi | D_lo] into accumulator columns [0,64) and [64,128), which the epilogue adds
    float* dw;
};

struct alignas(64) WgMaps {
    CUtensorMap X_hi[8], X_lo[8];     // per parity class, as in the forward kernel
    CUtensorMap D_hi, D_lo;
};

__device__ __forceinline__ void wg_split_parity(int d, int s, int& q, int& par) {
    if (s == 1) { q = d; par = 0; }
    else { par = d & 1; q = (d - par) >> 1; }
}

struct WgSmem { uint32_t a_bytes, b_bytes, tap_bytes, stage_bytes, bar_off, total; };
__host__ __device__ inline WgSmem wg_smem_layout(int BN, int nsplit, int nstages, int G, int x32 = 0, int x_single = 0) {
    WgSmem s;
    const uint32_t planes = nsplit == 3 ? 2u : 1u;
    const uint32_t xbox = x32 ? kWgBox / 2 : kWgBox;          // [64 positions x 32 | 64 channels]
    s.a_bytes = 2u * kWgBox * planes;                         // 128 rows of Cout = 2 boxes
    s.tap_bytes = (uint32_t)((BN + 63) / 64) * xbox * (x_single ? 1u : planes);   // X boxes of ONE tap (hi boxes of all taps come first,
                                                                  // then the lo boxes: every plane is one run of boxes)
    s.b_bytes = s.tap_bytes * (uint32_t)G;
    s.stage_bytes = s.a_bytes + s.b_bytes;
    s.bar_off = s.stage_bytes * (uint32_t)nstages;
    s.total = s.bar_off + 256;
    return s;
}

// XS: single-plane X operand (WgParams::x_single), swap mode only — a separate instantiation, the default one is unchanged.
template <bool XS>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const WgSmem L = wg_smem_layout(p.BN, p.nsplit, p.nstages, p.G, p.x32, XS ? 1 : 0);
    const uint32_t xbox = p.x32 ? kWgBox / 2 : kWgBox;        // bytes of one X box
    const uint32_t xstep = p.x32 ? 1024u : 2048u;             // bytes of 16 K rows of X
    const uint32_t lo_off = (uint32_t)(p.G * ((p.BN + 63) / 64)) * xbox;   // lo boxes follow the hi boxes of all G taps
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* empty_bar = full_bar + kWgMaxStages;
    uint64_t* tmem_full = empty_bar + kWgMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler (see conv_igemm.cu)
    const int lane = threadIdx.x & 31;
    const bool split = p.nsplit == 3;
    const uint32_t planes = split ? 2u : 1u;
    const int nboxes_b = (p.BN + 63) / 64;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.X_hi[0]);
        tma_prefetch_desc(&maps.D_hi);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.nstages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // item -> (tap group, nb, mb, ks), tap group fastest: the CTAs that run concurrently work on the SAME K chunk (range
    // of output positions) for different taps / channel blocks, so the chunk of D and X is fetched from HBM once and
    // re-read by the other taps from L2.  (With ks fastest every tap streamed the whole D and X tensors from HBM again:
    // taps x (|D| + |X|) of DRAM traffic, e.g. 16 GB for Conv3d_2c at batch 8.)
    // A work item covers G consecutive taps: the D tile of a stage is loaded once and multiplied against the G shifted X
    // tiles into G accumulators (TMEM columns [j*cstride, j*cstride + BN)), which cuts the L2->SMEM operand traffic per tap from
    // |D tile| + |X tile| to |D tile|/G + |X tile| — these kernels are bound by that traffic (~43 B/clk/SM), not by the MMA.
    const int ntaps_ = p.kt * p.kh * p.kw;
    auto decode = [&](int item, int& tap0, int& gsz, int& mb, int& nb, int& k_begin, int& k_end) {
        const int tg = item % p.ngroups; item /= p.ngroups;
        nb = item % p.n_blocks; item /= p.n_blocks;
        mb = item % p.m_blocks; item /= p.m_blocks;
        const int ks = item;
        tap0 = tg * p.G;
        gsz = min(p.G, ntaps_ - tap0);
        const long long kt_ = p.ktiles;
        k_begin = (int)(kt_ * ks / p.ksplit);
        k_end = (int)(kt_ * (ks + 1) / p.ksplit);
    };

    if (warp == 0) {
        // ---- TMA producer: the whole warp runs the loop (uniform registers), one elected lane issues
        int stage = 0; uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int tap0, gsz, mb, nb, k0, k1;
            decode(item, tap0, gsz, mb, nb, k0, k1);
            int qt[4], qh[4], qw[4], mi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tap = min(tap0 + j, ntaps_ - 1);
                const int dw = tap % p.kw, dh = (tap / p.kw) % p.kh, dt = tap / (p.kw * p.kh);
                int rt, rh, rw;
                wg_split_parity(dt - p.pt, p.st, qt[j], rt);
                wg_split_parity(dh - p.ph, p.sh, qh[j], rh);
                wg_split_parity(dw - p.pw, p.sw, qw[j], rw);
                mi[j] = rt * 4 + rh * 2 + rw;
            }
            const int m_valid = min(128, p.Cout - mb * 128);
            const int a_boxes = m_valid > 64 ? 2 : 1;     // rows 64..127 of a short block are never read back
            const uint32_t tx = XS ? (uint32_t)a_boxes * kWgBox * planes + (uint32_t)(gsz * nboxes_b) * xbox
                                   : ((uint32_t)a_boxes * kWgBox + (uint32_t)(gsz * nboxes_b) * xbox) * planes;
            // K tile k -> (n, t, h, w) tile indices, advanced with running counters (no integer division per stage)
            int iw, ih, it, n;
            {
                int m = k0;
                iw = m % p.tilesW; m /= p.tilesW;
                ih = m % p.tilesH; m /= p.tilesH;
                it = m % p.tilesT; m /= p.tilesT;
                n = m;
            }
            for (int k = k0; k < k1; ++k) {
                const int w0 = iw * p.tW, h0 = ih * p.tH, t0 = it * p.tT;
                const int n_cur = n;
                if (++iw == p.tilesW) { iw = 0; if (++ih == p.tilesH) { ih = 0; if (++it == p.tilesT) { it = 0; ++n; } } }
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    unsigned char* sA = smem + (size_t)stage * L.stage_bytes;
                    mbar_expect_tx(&full_bar[stage], tx);
                    for (int j = 0; j < a_boxes; ++j) {
                        tma_load_5d(&maps.D_hi, &full_bar[stage], sA + j * kWgBox, mb * 128 + j * 64, w0, h0, t0, n_cur);
                        if (split)   // XS: D_lo is the second 64-wide N block of the [D_hi | D_lo] operand (a_boxes == 1)
                            tma_load_5d(&maps.D_lo, &full_bar[stage], sA + (XS ? 1 : 2) * kWgBox + j * kWgBox, mb * 128 + j * 64, w0, h0,
                                        t0, n_cur);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g < gsz) {
                            unsigned char* sB = sA + L.a_bytes + (size_t)(g * nboxes_b) * xbox;      // hi boxes of tap g
                            for (int j = 0; j < nboxes_b; ++j) {
                                tma_load_5d(&maps.X_hi[mi[g]], &full_bar[stage], sB + j * xbox, nb * p.BN + j * 64, w0 + qw[g],
                                            h0 + qh[g], t0 + qt[g], n_cur);
                                if (split && !XS)
                                    tma_load_5d(&maps.X_lo[mi[g]], &full_bar[stage], sB + lo_off + j * xbox, nb * p.BN + j * 64,
                                                w0 + qw[g], h0 + qh[g], t0 + qt[g], n_cur);
                            }
                        }
                    }
                }
                __syncwarp();
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: whole warp runs the loop, one elected lane issues
        const uint32_t idesc = umma_idesc_bf16(128, p.BN, 1, 1);   // both operands MN-major
        // descriptor template: 16 K rows = 2048 bytes per K step; LBO = distance between 64-channel boxes, SBO = 8 K rows
        const uint64_t tmpl = umma_smem_desc_sw128(0, kWgBox, 1024);
        // X with 64-byte rows: MN-major SWIZZLE_64B, 32 contiguous elements per K row, 8-row groups 512 bytes apart
        const uint64_t tmpl_x = p.x32 ? umma_smem_desc(0, kWgBox / 2, 512, 4) : tmpl;
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int tap0, gsz, mb, nb, k0, k1;
            decode(item, tap0, gsz, mb, nb, k0, k1);
            if (k1 <= k0) continue;                       // empty split: nothing to add (producer/epilogue skip too)
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256;
            for (int k = k0; k < k1; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + (size_t)stage * L.stage_bytes);
                const uint64_t a_hi0 = tmpl + (uint64_t)(sA >> 4);
                const uint64_t a_lo0 = tmpl + (uint64_t)((sA + 2 * kWgBox) >> 4);
                if (p.swap) {
                    if (elect_one()) {
                        // A = X tiles of the 4 taps (MN-major SW64, 32-row blocks LBO = tap_bytes apart), B = D tile
                        const uint32_t sB = sA + L.a_bytes;
                        const uint64_t tx_ = umma_smem_desc(0, xbox, 512, 4);
                        const uint32_t idesc_s = umma_idesc_bf16(128, p.BNs, 1, 1);
#pragma unroll
                        for (int ks = 0; ks < kWgKP / 16; ++ks) {
                            const uint64_t xa_hi = tx_ + (uint64_t)((sB + ks * 1024) >> 4);
                            const uint64_t xa_lo = tx_ + (uint64_t)((sB + lo_off + ks * 1024) >> 4);
                            const uint64_t d_hi = a_hi0 + (uint64_t)(ks * (2048 >> 4));
                            const uint64_t d_lo = a_lo0 + (uint64_t)(ks * (2048 >> 4));
                            if constexpr (XS) {      // X * [D_hi | D_lo]: N = 2 * 64
                                umma_f16(d_tmem, xa_hi, d_hi, umma_idesc_bf16(128, 128, 1, 1), (k != k0 || ks != 0));
                                continue;
                            }
                            umma_f16(d_tmem, xa_hi, d_hi, idesc_s, (k != k0 || ks != 0));
                            if (split) {
                                umma_f16(d_tmem, xa_lo, d_hi, idesc_s, 1);
                                umma_f16(d_tmem, xa_hi, d_lo, idesc_s, 1);
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                    }
                } else if (elect_one()) {
                    // ONE MMA per term covers all taps of the group: their X boxes sit side by side in shared memory, i.e.
                    // they are consecutive 64-wide N blocks (LBO = one box) of an N = gsz * nboxes * 64 operand, and the
                    // accumulator of tap g is the column range [g*cstride, g*cstride + BN).  The D tile is then read from
                    // shared memory once per term instead of once per tap and term (these MMAs are SMEM-bandwidth-bound).
                    const uint32_t sB = sA + L.a_bytes;
                    const uint32_t idesc_g = umma_idesc_bf16(128, gsz * nboxes_b * 64, 1, 1);
#pragma unroll
                    for (int ks = 0; ks < kWgKP / 16; ++ks) {
                        const uint64_t a_hi = a_hi0 + (uint64_t)(ks * (2048 >> 4));
                        const uint64_t a_lo = a_lo0 + (uint64_t)(ks * (2048 >> 4));
                        const uint64_t b_hi = tmpl_x + (uint64_t)((sB + ks * xstep) >> 4);
                        umma_f16(d_tmem, a_hi, b_hi, idesc_g, (k != k0 || ks != 0));
                        if (split) {
                            const uint64_t b_lo = tmpl_x + (uint64_t)((sB + lo_off + ks * xstep) >> 4);
                            umma_f16(d_tmem, a_lo, b_hi, idesc_g, 1);
                            umma_f16(d_tmem, a_hi, b_lo, idesc_g, 1);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tmem_full[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int tap0, gsz, mb, nb, k0, k1;
            decode(item, tap0, gsz, mb, nb, k0, k1);
            if (k1 <= k0) continue;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)acc * 256 + ((uint32_t)(q * 32) << 16);
            if (p.swap) {
                // accumulator row = tap-in-group * 32 + folded input index j, column = output channel: for a fixed
                // column the 32 lanes of a warp add to 32 consecutive floats of dW[tap][co][:]
                if (q < gsz) {
                    float* dst = p.dw + (size_t)(tap0 + q) * p.Cout * 32 + lane;
                    for (int col0 = 0; col0 < p.Cout; col0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(t_acc + col0, v);
                        if constexpr (XS) {                  // second accumulator half: the X * D_lo products
                            uint32_t v2[32];
                            tmem_ld32(t_acc + 64 + col0, v2);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                        } else {
                            tmem_ld_wait();
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.Cout) atomicAdd(dst + (size_t)(col0 + j) * 32, __uint_as_float(v[j]));
                    }
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            const int co = mb * 128 + row;
            for (int g = 0; g < gsz; ++g) {
                float* dst = p.dw + ((size_t)(tap0 + g) * p.Cout + co) * p.Cin + nb * p.BN;
                for (int col0 = 0; col0 < p.BN; col0 += 32) {
                    uint32_t v[32];
                    if (p.BN - col0 >= 32) tmem_ld32(t_acc + g * p.cstride + col0, v);
                    else {
                        uint32_t v16[16];
                        tmem_ld16(t_acc + g * p.cstride + col0, v16);
#pragma unroll
                        for (int j = 0; j < 16; ++j) { v[j] = v16[j]; v[16 + j] = 0; }
                    }
                    tmem_ld_wait();
                    if (co < p.Cout) {
                        // 16-byte vector reductions (red.global.add.v4.f32): Cin, BN and col0 are multiples of 4 and dW
                        // blocks are 32-byte aligned, so groups of 4 columns are all-in or all-out and aligned
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int ci = nb * p.BN + col0 + j;
                            if (col0 + j < p.BN && ci < p.Cin)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + col0 + j),
                                             "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                             "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                             : "memory");
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int wg_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static int wg_finish_and_launch(WgParams& p, WgMaps& maps, const uint16_t* d_hi, const uint16_t* d_lo, int d_cstride,
                                int d_coff, cudaStream_t stream) {
    const bool split = p.nsplit == 3;
    p.tilesT = (p.To + p.tT - 1) / p.tT; p.tilesH = (p.Ho + p.tH - 1) / p.tH; p.tilesW = (p.Wo + p.tW - 1) / p.tW;
    p.ktiles = p.N * p.tilesT * p.tilesH * p.tilesW;
    p.m_blocks = (p.Cout + 127) / 128;
    // N block: whole Cin when it fits 256 accumulator columns, else the multiple of 64 that wastes least
    if (p.Cin <= 256) { p.BN = (p.Cin + 15) / 16 * 16; p.n_blocks = 1; }
    else {
        int best = 256, best_pad = 1 << 30;
        for (int bn = 256; bn >= 64; bn -= 64) {
            int pad = (p.Cin + bn - 1) / bn * bn;
            if (pad < best_pad) { best_pad = pad; best = bn; }
        }
        p.BN = best; p.n_blocks = (p.Cin + best - 1) / best;
    }
    const int ntaps = p.kt * p.kh * p.kw;
    // short reductions (the 1-D head: K = 16..512 positions) cannot be split along K: narrow the N block instead so
    // that enough CTAs share the epilogue's reductions into dW
    while (p.BN > 64 && p.BN % 64 == 0 && (long long)ntaps * p.m_blocks * p.n_blocks * p.ktiles < wg_num_sms()) {
        p.BN -= 64;
        p.n_blocks = (p.Cin + p.BN - 1) / p.BN;
    }
    p.swap = (p.x32 && p.Cout <= 128 && p.Cout % 32 == 0 && p.Cin == 32) ? 1 : 0;
    p.BNs = (p.Cout + 15) / 16 * 16;
    if (p.x32 && !p.swap) { set_last_error_msg("conv1a_wgrad: Cout must be a multiple of 32 and <= 128"); return OTAL_ERR_UNSUPPORTED; }
    if (p.x_single && !(p.swap && split && p.Cout == 64)) {
        set_last_error_msg("conv1a_wgrad_u8: needs nsplit 3 and Cout == 64"); return OTAL_ERR_UNSUPPORTED;
    }
    // taps per work item: as many accumulators as fit 256 TMEM columns (double buffered) and two pipeline stages of
    // shared memory (D tile + G X tiles, 16 KB per box pair in bf16x3)
    {
        const int nbx = (p.BN + 63) / 64;
        p.cstride = nbx * 64;                                        // one 64-wide N block per X box
        int G = 256 / p.cstride;
        if (G > 4 / nbx) G = 4 / nbx;
        if (G > ntaps) G = ntaps;
        if (G < 1) G = 1;
        if (p.swap) G = 4;                                           // 4 x 32 rows = one M = 128 operand
        p.G = G;
        p.ngroups = (ntaps + G - 1) / G;
    }
    const int base_items = p.ngroups * p.m_blocks * p.n_blocks;
    int ks = (2 * wg_num_sms() + base_items - 1) / base_items;       // at least two waves of work items
    {
        // K chunk sized so that the chunks in flight (concurrent CTAs / items per chunk, + 1 for the transition) stay
        // L2-resident: bytes per 64-position K tile = 64 x (Cout + Cin) channels x 2 B x planes
        const double tile_bytes = 64.0 * (p.Cout + p.Cin) * 2.0 * (split ? 2.0 : 1.0);
        const int in_flight = (wg_num_sms() + base_items - 1) / base_items + 1;
        const double budget = 40.0 * 1024 * 1024 / in_flight;
        long long chunk_tiles = (long long)(budget / tile_bytes);
        if (chunk_tiles < 16) chunk_tiles = 16;                      // keep the accumulation long enough to amortise the epilogue
        const long long ks_l2 = (p.ktiles + chunk_tiles - 1) / chunk_tiles;
        if (ks_l2 > ks) ks = (int)(ks_l2 > 4096 ? 4096 : ks_l2);
    }
    if (ks > p.ktiles) ks = p.ktiles;
    if (ks < 1) ks = 1;
    p.ksplit = ks;
    p.total_items = base_items * ks;

    const uint32_t smem_cap = 227 * 1024 - 1024;
    int nst = 0;
    for (int s = kWgMaxStages; s >= 2 && !nst; --s)
        if (wg_smem_layout(p.BN, p.nsplit, s, p.G, p.x32, p.x_single).total <= smem_cap) nst = s;
    if (!nst) { set_last_error_msg("wgrad: tile does not fit shared memory"); return OTAL_ERR_UNSUPPORTED; }
    p.nstages = nst;
    const WgSmem SL = wg_smem_layout(p.BN, p.nsplit, nst, p.G, p.x32, p.x_single);

    int rc;
    const uint32_t box[5] = {64, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tT, 1};
    const uint64_t ddims[5] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.To, (uint64_t)p.N};
    const uint64_t cs = (uint64_t)d_cstride * 2;
    const uint64_t dst_[4] = {cs, cs * p.Wo, cs * p.Wo * p.Ho, cs * p.Wo * p.Ho * p.To};
    if ((rc = make_tensor_map_bf16(&maps.D_hi, d_hi + d_coff, 5, ddims, dst_, box, 1))) return rc;
    if (split && (rc = make_tensor_map_bf16(&maps.D_lo, d_lo + d_coff, 5, ddims, dst_, box, 1))) return rc;

    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        once.mark(once_dev);
    }
    if (p.x_single) {       // the staged instantiation is configured on its first use only
        static OncePerDevice once_staged;
        if (once_staged.need(&once_dev)) {
            OTAL_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            once_staged.mark(once_dev);
        }
    }
    const int grid = p.total_items < wg_num_sms() ? p.total_items : wg_num_sms();
    if (p.x_single) conv_wgrad_kernel<true><<<grid, kWgThreads, SL.total + 1024, stream>>>(maps, p);
    else conv_wgrad_kernel<false><<<grid, kWgThreads, SL.total + 1024, stream>>>(maps, p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // namespace otal

using namespace otal;

extern "C" {

int otal_conv_wgrad(const otal_wgrad_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("wgrad: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (d->N <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) {
        set_last_error_msg("wgrad: non-positive dimension"); return OTAL_ERR_BAD_ARG;
    }
    if (d->Cin % 8 || d->Cout % 8 || d->x_cstride % 8 || d->x_coff % 8 || d->d_cstride % 8 || d->d_coff % 8) {
        set_last_error_msg("wgrad: channel counts / strides / offsets must be multiples of 8"); return OTAL_ERR_BAD_ARG;
    }
    if (d->tT * d->tH * d->tW != kWgKP) { set_last_error_msg("wgrad: K tile box must hold 64 positions"); return OTAL_ERR_BAD_ARG; }
    if (d->nsplit != 1 && d->nsplit != 3) { set_last_error_msg("wgrad: nsplit must be 1 or 3"); return OTAL_ERR_BAD_ARG; }
    const int st = d->sT ? d->sT : 1, sh = d->sH ? d->sH : 1, sw = d->sW ? d->sW : 1;
    if ((st != 1 && st != 2) || (sh != 1 && sh != 2) || (sw != 1 && sw != 2)) {
        set_last_error_msg("wgrad: stride must be 1 or 2"); return OTAL_ERR_BAD_ARG;
    }
    const bool split = d->nsplit == 3;
    if (!d->x_hi || !d->d_hi || !d->dw || (split && (!d->x_lo || !d->d_lo))) { set_last_error_msg("wgrad: null pointer"); return OTAL_ERR_BAD_ARG; }

    WgParams p{};
    p.N = d->N; p.Cin = d->Cin; p.Cout = d->Cout;
    p.To = (d->T + st - 1) / st; p.Ho = (d->H + sh - 1) / sh; p.Wo = (d->W + sw - 1) / sw;
    p.kt = d->kt; p.kh = d->kh; p.kw = d->kw; p.pt = d->pt; p.ph = d->ph; p.pw = d->pw;
    p.st = st; p.sh = sh; p.sw = sw;
    p.tT = d->tT; p.tH = d->tH; p.tW = d->tW;
    p.nsplit = d->nsplit; p.dw = d->dw;

    WgMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    const uint64_t cs = (uint64_t)d->x_cstride * 2;
    const uint64_t sW_ = cs, sH_ = cs * d->W, sT_ = cs * d->W * d->H, sN_ = cs * d->W * d->H * d->T;
    const uint32_t box[5] = {64, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tT, 1};
    for (int rt = 0; rt < st; ++rt) for (int rh = 0; rh < sh; ++rh) for (int rw = 0; rw < sw; ++rw) {
        const int eT = (d->T - rt + st - 1) / st, eH = (d->H - rh + sh - 1) / sh, eW = (d->W - rw + sw - 1) / sw;
        if (eT <= 0 || eH <= 0 || eW <= 0) continue;
        const uint64_t xdims[5] = {(uint64_t)d->Cin, (uint64_t)eW, (uint64_t)eH, (uint64_t)eT, (uint64_t)d->N};
        const uint64_t xst[4] = {sW_ * sw, sH_ * sh, sT_ * st, sN_};
        const size_t off = (size_t)d->x_coff + ((size_t)rt * d->H * d->W + (size_t)rh * d->W + rw) * d->x_cstride;
        const int mi = rt * 4 + rh * 2 + rw;
        if ((rc = make_tensor_map_bf16(&maps.X_hi[mi], d->x_hi + off, 5, xdims, xst, box, 1))) return rc;
        if (split && (rc = make_tensor_map_bf16(&maps.X_lo[mi], d->x_lo + off, 5, xdims, xst, box, 1))) return rc;
    }
    return wg_finish_and_launch(p, maps, d->d_hi, d->d_lo, d->d_cstride, d->d_coff, stream);
}

// Weight gradient of Conv3d_1a_7x7 in the folded layout of otal_conv1a_fwd: dw is [49][Cout][32] fp32.
static int conv1a_wgrad_impl(const otal_conv1a_wgrad_desc* d, void* stream_, bool u8) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("conv1a_wgrad: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (u8 && d->nsplit != 3) { set_last_error_msg("conv1a_wgrad_u8: needs nsplit 3 (hi/lo output-gradient planes)"); return OTAL_ERR_BAD_ARG; }
    if (d->N <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->W % 2 || d->Cout <= 0 || d->Cout % 8 ||
        d->d_cstride % 8 || d->d_coff % 8) {
        set_last_error_msg("conv1a_wgrad: bad dimension"); return OTAL_ERR_BAD_ARG;
    }
    if (d->tT * d->tH * d->tW != kWgKP) { set_last_error_msg("conv1a_wgrad: K tile box must hold 64 positions"); return OTAL_ERR_BAD_ARG; }
    if (d->nsplit != 1 && d->nsplit != 3) { set_last_error_msg("conv1a_wgrad: nsplit must be 1 or 3"); return OTAL_ERR_BAD_ARG; }
    const bool split = d->nsplit == 3;
    if (!d->x_hi || !d->d_hi || !d->dw || (split && ((!u8 && !d->x_lo) || !d->d_lo))) { set_last_error_msg("conv1a_wgrad: null pointer"); return OTAL_ERR_BAD_ARG; }

    WgParams p{};
    p.N = d->N; p.Cin = 32; p.Cout = d->Cout;
    p.x32 = 1;
    p.x_single = u8 ? 1 : 0;
    p.To = (d->T + 1) / 2; p.Ho = (d->H + 1) / 2; p.Wo = d->W / 2;
    p.kt = 7; p.kh = 7; p.kw = 1;
    p.pt = (d->T % 2 == 0) ? 2 : 3; p.ph = (d->H % 2 == 0) ? 2 : 3; p.pw = 0;
    p.st = 2; p.sh = 2; p.sw = 1;
    p.tT = d->tT; p.tH = d->tH; p.tW = d->tW;
    p.nsplit = d->nsplit; p.dw = d->dw;

    WgMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    const uint64_t px = 4 * 2;                                  // bytes per padded pixel; rows are W + 8 pixels (otal_clip_ingest)
    const uint64_t Wp = (uint64_t)d->W + 8;
    const uint64_t sH_ = px * Wp, sT_ = sH_ * d->H, sN_ = sT_ * d->T;
    const uint32_t box[5] = {32, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tT, 1};
    for (int rt = 0; rt < 2; ++rt) for (int rh = 0; rh < 2; ++rh) {
        const int eT = (d->T - rt + 1) / 2, eH = (d->H - rh + 1) / 2;
        if (eT <= 0 || eH <= 0) continue;
        const uint64_t xdims[5] = {32, (uint64_t)p.Wo, (uint64_t)eH, (uint64_t)eT, (uint64_t)d->N};
        const uint64_t xst[4] = {2 * px, sH_ * 2, sT_ * 2, sN_};
        const size_t off = ((size_t)rt * d->H + rh) * Wp * 4;
        const int mi = rt * 4 + rh * 2;
        if ((rc = make_tensor_map_bf16(&maps.X_hi[mi], d->x_hi + off, 5, xdims, xst, box, 2))) return rc;
        if (split && !u8 && (rc = make_tensor_map_bf16(&maps.X_lo[mi], d->x_lo + off, 5, xdims, xst, box, 2))) return rc;
    }
    return wg_finish_and_launch(p, maps, d->d_hi, d->d_lo, d->d_cstride, d->d_coff, stream);
}

int otal_conv1a_wgrad(const otal_conv1a_wgrad_desc* d, void* stream) { return conv1a_wgrad_impl(d, stream, false); }

// Weight gradient of Conv3d_1a against the RAW uint8 clip (one exact bf16 plane, x_lo ignored): dw += sum_p D[p] * u[p + tap]
// in ONE tensor-core pass (u * [D_hi | D_lo]) instead of three.  The caller turns it into the gradient of the reference's
// conv on the normalised, zero-padded clip: dW = (2/255) * dw - (sum of D over the positions where the tap is inside the
// image) — see otal_conv1a_fwd_u8 and otal_border_class_sums.
int otal_conv1a_wgrad_u8(const otal_conv1a_wgrad_desc* d, void* stream) { return conv1a_wgrad_impl(d, stream, true); }

}  // extern "C"

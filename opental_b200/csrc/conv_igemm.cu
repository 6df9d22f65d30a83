// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): the I3D Unit3D / AFSD Unit1D hot op.
//
// Replaces, for one conv layer, the reference chain  F.pad -> nn.Conv3d(bias=False) -> BatchNorm3d(eval) -> ReLU
// (AFSD/common/i3d_backbone.py:51-87) and  F.pad -> nn.Conv1d(bias) (AFSD/common/layers.py:204-214; a 1-D conv
// is the H=W=1 case).  Stride-1 "same"/explicit padding only; the stride-2 convs have their own entry points.
//
// GEMM view:  D[p, co] = sum_{tap, ci} X[p + tap - pad, ci] * Wt[tap, co, ci]
//   M = output positions, tiled as boxes tT x tH x tW = 128 positions of one sample
//   N = output channels (BN <= 256 per tile), K = taps x Cin (64-channel chunks)
//
// Data layout in HBM: activations are NDHWC ("channels-last"), each tensor stored as two bf16 planes
// (hi, lo) with x ~= hi + lo; weights are [tap][Cout][Cin] bf16 hi/lo planes.  In `nsplit == 3` mode the
// kernel computes  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  with fp32 accumulation in TMEM ("bf16x3", ~3e-5 relative
// error end to end); `nsplit == 1` is plain bf16.
//
// Pipeline (one CTA per SM, persistent over tiles, 256 threads):
//   warp 0   : TMA producer.  For every (tap, 64-channel chunk) one 5-D box load per activation plane — the
//              box origin is shifted by (tap - pad), out-of-bounds elements are zero-filled by the TMA unit,
//              which *is* the padding (no F.pad copy) — plus one 3-D box load per weight plane.  Boxes land
//              in shared memory in the 128-byte-swizzled K-major layout tcgen05 consumes directly.
//   warp 1   : MMA issuer (one elected thread): 4 (K=16) x nsplit tcgen05.mma per stage into a TMEM
//              accumulator (128 lanes x BN fp32 columns, double buffered), tcgen05.commit frees the stage.
//   warp 2   : TMEM allocator.
//   warps 4-7: epilogue.  tcgen05.ld -> per-channel scale/shift (folded frozen BN, or bias) -> ReLU ->
//              bf16 hi/lo split -> swizzled staging in shared memory -> TMA store into the channel slice of
//              the destination (so inception branch outputs land directly inside the concat buffer); TMA
//              clips partial tiles.  Optionally also writes fp32 with plain stores.
#include "common.cuh"
#include "tensormap.h"
#include <stdlib.h>

namespace otal {

constexpr int kConvThreads = 256;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // bf16 elements per K chunk = 128 bytes = one swizzle row
constexpr int kATileBytes = kTileM * 128;         // 16 KB
constexpr int kMaxStages = 8;
constexpr int kAccStride = 256;                   // TMEM columns per accumulator
constexpr int kTmemCols = 512;

struct ConvParams {
    int N, T, H, W;
    int Cin, Cout;
    int kt, kh, kw, pt, ph, pw;
    int st, sh, sw;  // conv stride per dim (1 or 2); stride-2 dims read parity-split views of the input
    int tT, tH, tW;
    int tilesT, tilesH, tilesW;
    int BN, n_blocks, kchunks;
    int nsplit;      // 1 (bf16) or 3 (bf16x3)
    int relu;
    int store_bf16;  // TMA store of hi (and lo when nsplit==3) planes
    int nstages;
    int nbuf;        // staging buffers for the TMA store (1 or 2; 0 when store_bf16 == 0)
    int total_tiles;
    int out_cstride; // fp32 output: channel stride (elements) of one position and channel offset of the slice
    int out_coff;
    int accumulate;  // fp32 destination: y += result (dgrad into a shared gradient buffer)
    int out_ncdhw;   // fp32 destination is channel-major [N, out_cstride, T, H, W] instead of channels-last
    int b_mn;        // dgrad mode: weights are read as [tap][K][N] (N contiguous, "MN-major" B) and taps are flipped
    int kchunks2;    // second K segment (1x1 convs only): extra 64-channel chunks read through the A2 / B2 maps, i.e.
                     // D = [x | x2] . [w ; w2] — the data gradients of several 1x1 convs that share their input, in ONE pass
    int ncat;        // bf16x3 with BN <= 128: B_lo sits right behind B_hi in shared memory, so ONE MMA of N = 2*BN computes
                     // a_hi*[b_hi | b_lo] into columns [0,BN) and [BN,2BN) and a second one adds a_lo*b_hi to [0,BN): the A
                     // tile is read from shared memory twice instead of three times per K step (these shapes are bound by the
                     // 128 B/clk shared-memory operand bandwidth, not by the tensor pipe); the epilogue adds the two halves
    int k32;         // K chunk of 32 elements = 64-byte rows, SWIZZLE_64B operands (the folded Conv3d_1a: 8 W taps x 4 channels)
    int tail_ksteps;   // K = 16 steps the LAST chunk of a tap really holds (Cin = 144: 2 full chunks + 1 step instead of 4): the padded
    int tail_ksteps2;  // part of a chunk is TMA zero fill in both operands, so its MMAs add exact zeros — not issued.  2: second K segment
    const float* scale;  // [Cout] or nullptr (=1)
    const float* shift;  // [Cout] or nullptr (=0)
    float* out_f32;      // optional NDHWC fp32 destination (nullptr = skip)
    int a_single;    // U8 instantiation: the A operand is ONE exact bf16 plane (raw uint8 pixel values 0..255) while B and the
                     // output keep hi/lo planes: a*[b_hi | b_lo] is ONE N-concatenated MMA per K step (needs ncat)
    int ksplit;      // KS instantiation: a tile's K iterations (taps x chunks) are split over `ksplit` CTAs, each adds its partial
                     // sums to the fp32 destination with atomics (zeroed by the caller unless it accumulates); small problems
                     // only — the 1-D head, where one CTA per tile streams 24..72 stages alone while half the SMs idle
    unsigned long long* timeline;   // developer timeline buffer (OTAL_TIMELINE builds only), else nullptr
    int shift_classes;  // U8: `shift` is a table [4*4*4 border classes][Cout]; class of an output index o along a dim of n
                     // outputs = 1 (o == 0), 2 (o == n-2), 3 (o == n-1), else 0 — which taps of a 7-tap stride-2 window fall
                     // outside the image, where the reference pads the NORMALISED clip with 0 and the raw clip holds 0 = -1
};

// Developer timeline (compiled in only with -DOTAL_TIMELINE, see tools/conv_timeline.py): per CTA and tile, clock64() of the
// pipeline events of each role.  g_timeline is set by otal_debug_set_timeline(); slots: 0 producer starts the tile, 1 first
// stage issued, 2 last stage issued, 3 MMA got the first stage, 4 MMA committed the tile, 5 epilogue waits for the
// accumulator, 6 got it, 7 first chunk stored, 8 accumulator released, 9 MMA waits for a free accumulator, 10 got it.
#ifdef OTAL_TIMELINE
#define OTAL_TL(tileno, slot) do { if (p.timeline && (tileno) < 16 && lane == 0) \
    p.timeline[((size_t)blockIdx.x * 16 + (tileno)) * 16 + (slot)] = (unsigned long long)clock64(); } while (0)
static unsigned long long* g_timeline = nullptr;
#else
#define OTAL_TL(tileno, slot) do { } while (0)
#endif

struct ConvSmem {
    // dynamic shared memory carve-up, all offsets relative to a 1024-byte aligned base
    uint32_t stage_bytes, a_bytes, b_bytes;
    uint32_t staging_off, bar_off, total;
};

__host__ __device__ inline uint32_t conv_b_rows(int BN, int b_mn) { return b_mn ? (uint32_t)((BN + 63) / 64) * 64u : (uint32_t)BN; }

__host__ __device__ inline ConvSmem conv_smem_layout(int BN, int nsplit, int nstages, int nbuf, int b_mn, int k32 = 0,
                                                      int a_single = 0) {
    ConvSmem s;
    const uint32_t planes = nsplit == 3 ? 2u : 1u;
    const uint32_t rowb = k32 ? 64u : 128u;                        // operand row pitch in shared memory
    s.a_bytes = kTileM * rowb * (a_single ? 1u : planes);
    s.b_bytes = ((conv_b_rows(BN, b_mn) * rowb + 1023u) & ~1023u) * planes;
    s.stage_bytes = s.a_bytes + s.b_bytes;
    s.staging_off = s.stage_bytes * (uint32_t)nstages;
    uint32_t staging = (uint32_t)nbuf * planes * kATileBytes;      // nbuf (0, 1 or 2) buffers x planes x 16 KB
    s.bar_off = s.staging_off + staging;
    s.total = s.bar_off + 256 + 4096;      // barriers, then the epilogue's per-tile (scale, shift) table [2][256] float2
    return s;
}

// All TMA descriptors of one launch.  A has one map per input parity class (pt*4 + ph*2 + pw) so that a
// stride-2 conv reads the (even|odd) sub-lattice of the input as a dense tensor; stride-1 convs use A[0] only.
struct alignas(64) ConvMaps {
    CUtensorMap A_hi[8], A_lo[8];
    CUtensorMap B_hi, B_lo, O_hi, O_lo;
    CUtensorMap A2_hi, A2_lo, B2_hi, B2_lo;      // second K segment (kchunks2 > 0)
};

// d = tap offset - front pad along one dim with stride s: input index = s*o + d = s*(o + q) + par
__device__ __forceinline__ void split_parity(int d, int s, int& q, int& par) {
    if (s == 1) { q = d; par = 0; }
    else { par = d & 1; q = (d - par) >> 1; }
}

// U8 (implies NCAT): single-plane A operand + border-class shift table, see ConvParams::a_single / shift_classes.  A separate
// instantiation, so the code of the other two is unchanged by it.
// KS: split-K over CTAs with an atomic fp32 epilogue (ConvParams::ksplit) — also a separate instantiation.
template <bool NCAT, bool U8 = false, bool KS = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p) {
    static_assert(NCAT || !U8, "the single-plane A operand needs the N-concatenated weight tile");
    static_assert(!(U8 && KS), "split-K is for the small fp32-output problems");
    const CUtensorMap& mapB_hi = maps.B_hi; const CUtensorMap& mapB_lo = maps.B_lo;
    const CUtensorMap& mapO_hi = maps.O_hi; const CUtensorMap& mapO_lo = maps.O_lo;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment is required by the 128B swizzle pattern (pattern repeats every 8 rows x 128 B)
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const ConvSmem L = conv_smem_layout(p.BN, p.nsplit, p.nstages, p.nbuf, p.b_mn, p.k32, U8 ? 1 : 0);
    const uint32_t planes_ = p.nsplit == 3 ? 2u : 1u;
    const uint32_t a_plane = U8 ? L.a_bytes : L.a_bytes / planes_;   // bytes of one A / B plane per stage
    const uint32_t b_plane = L.b_bytes / planes_;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops (addresses, UMMA /
    // TMA descriptors, barrier addresses) in uniform registers; only the issue instructions sit behind elect_one().
    // (With the loops inside `if (lane == 0)` every tcgen05.mma paid a ~15-instruction R2UR waterfall: issue-bound at N <= 64.)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const bool split = p.nsplit == 3;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.A_hi[0]);
        tma_prefetch_desc(&mapB_hi);
        if (split) { if (!U8) tma_prefetch_desc(&maps.A_lo[0]); tma_prefetch_desc(&mapB_lo); }
        if (p.store_bf16) { tma_prefetch_desc(&mapO_hi); if (split) tma_prefetch_desc(&mapO_lo); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.nstages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int ntaps = p.kt * p.kh * p.kw;
    const int kiters = ntaps * p.kchunks + p.kchunks2;
    const int kchunk_elems = p.k32 ? 32 : kChunkK;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (whole warp runs the loop)
        int stage = 0; uint32_t phase = 0;
        int tl_n = 0, tl_first = 1;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl_n) {
            OTAL_TL(tl_n, 0); tl_first = 1;
            int tdec = tile, it_lo = 0, it_hi = 0x7fffffff, it_cur = 0;
            if constexpr (KS) {                       // K range of this CTA's share of the tile
                const int ks = tdec % p.ksplit; tdec /= p.ksplit;
                it_lo = kiters * ks / p.ksplit; it_hi = kiters * (ks + 1) / p.ksplit;
            }
            const int nb = tdec % p.n_blocks;
            int m = tdec / p.n_blocks;
            const int w0 = (m % p.tilesW) * p.tW; m /= p.tilesW;
            const int h0 = (m % p.tilesH) * p.tH; m /= p.tilesH;
            const int t0 = (m % p.tilesT) * p.tT; m /= p.tilesT;
            const int n = m;
            // taps in (dt, dh, dw) order with running counters: no integer division in the per-stage path (the producer
            // warp's address arithmetic, not the TMA unit, paces short stages such as the folded Conv3d_1a)
            int tap = 0;
            for (int dt = 0; dt < p.kt; ++dt) {
                int qt, rt;
                split_parity(dt - p.pt, p.st, qt, rt);
                for (int dh = 0; dh < p.kh; ++dh) {
                    int qh, rh;
                    split_parity(dh - p.ph, p.sh, qh, rh);
                    for (int dw = 0; dw < p.kw; ++dw, ++tap) {
                        int qw, rw;
                        split_parity(dw - p.pw, p.sw, qw, rw);
                        const int mi = rt * 4 + rh * 2 + rw;
                        const CUtensorMap* mapA_hi = &maps.A_hi[mi];
                        const CUtensorMap* mapA_lo = &maps.A_lo[mi];
                        const int cw = w0 + qw, ch = h0 + qh, ct = t0 + qt;
                        const int btap = p.b_mn ? ntaps - 1 - tap : tap;
                        for (int kc = 0; kc < p.kchunks; ++kc) {
                            if constexpr (KS) {
                                const int it = it_cur++;
                                if (it < it_lo || it >= it_hi) continue;
                            }
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            if (elect_one()) {
                                unsigned char* sA = smem + (size_t)stage * L.stage_bytes;
                                unsigned char* sB = sA + L.a_bytes;
                                mbar_expect_tx(&full_bar[stage], L.stage_bytes);
                                const int c0 = kc * kchunk_elems;
                                tma_load_5d(mapA_hi, &full_bar[stage], sA, c0, cw, ch, ct, n);
                                if (!p.b_mn) {
                                    tma_load_3d(&mapB_hi, &full_bar[stage], sB, c0, nb * p.BN, btap);
                                    if (split) tma_load_3d(&mapB_lo, &full_bar[stage], sB + b_plane, c0, nb * p.BN, btap);
                                } else {
                                    // [64 K rows x 64 N] boxes of the flipped tap: N contiguous = MN-major B
                                    const int nbx = (p.BN + 63) / 64;
                                    for (int j = 0; j < nbx; ++j) {
                                        tma_load_3d(&mapB_hi, &full_bar[stage], sB + j * 8192, nb * p.BN + j * 64, c0, btap);
                                        if (split)
                                            tma_load_3d(&mapB_lo, &full_bar[stage], sB + b_plane + j * 8192, nb * p.BN + j * 64, c0, btap);
                                    }
                                }
                                if (split && !U8) tma_load_5d(mapA_lo, &full_bar[stage], sA + a_plane, c0, cw, ch, ct, n);
                            }
                            __syncwarp();
                            if (tl_first) { OTAL_TL(tl_n, 1); tl_first = 0; }
                            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            if (p.kchunks2 == 0) OTAL_TL(tl_n, 2);
            for (int kc = 0; kc < p.kchunks2; ++kc) {          // second K segment (1x1, stride 1: no tap offsets)
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    unsigned char* sA = smem + (size_t)stage * L.stage_bytes;
                    unsigned char* sB = sA + L.a_bytes;
                    mbar_expect_tx(&full_bar[stage], L.stage_bytes);
                    const int c0 = kc * kChunkK;
                    tma_load_5d(&maps.A2_hi, &full_bar[stage], sA, c0, w0, h0, t0, n);
                    if (!p.b_mn) {
                        tma_load_3d(&maps.B2_hi, &full_bar[stage], sB, c0, nb * p.BN, 0);
                        if (split) tma_load_3d(&maps.B2_lo, &full_bar[stage], sB + b_plane, c0, nb * p.BN, 0);
                    } else {
                        const int nbx = (p.BN + 63) / 64;
                        for (int j = 0; j < nbx; ++j) {
                            tma_load_3d(&maps.B2_hi, &full_bar[stage], sB + j * 8192, nb * p.BN + j * 64, c0, 0);
                            if (split) tma_load_3d(&maps.B2_lo, &full_bar[stage], sB + b_plane + j * 8192, nb * p.BN + j * 64, c0, 0);
                        }
                    }
                    if (split) tma_load_5d(&maps.A2_lo, &full_bar[stage], sA + a_plane, c0, w0, h0, t0, n);
                }
                __syncwarp();
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
            if (p.kchunks2 != 0) OTAL_TL(tl_n, 2);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (whole warp runs the loop)
        const uint32_t idesc = umma_idesc_bf16(kTileM, p.BN, 0, p.b_mn ? 1 : 0);
        const uint32_t idesc_cat = umma_idesc_bf16(kTileM, 2 * p.BN, 0, p.b_mn ? 1 : 0);
        // descriptor templates: everything but the start address.  K-major: 32 bytes per K=16 step inside the swizzled
        // row.  MN-major B (dgrad): 16 K rows = 2048 bytes per step, LBO = 8 KB between 64-wide N boxes, SBO = 1 KB.
        // k32: 64-byte rows, SWIZZLE_64B, 8-row groups 512 bytes apart, two K=16 steps per stage.
        const uint64_t tmpl_k = p.k32 ? umma_smem_desc(0, 16, 512, 4) : umma_smem_desc_sw128(0, 16, 1024);
        const uint64_t tmpl_b = p.b_mn ? umma_smem_desc_sw128(0, 8192, 1024) : tmpl_k;
        const uint32_t b_step = p.b_mn ? (2048u >> 4) : (32u >> 4);
        const int ksteps = p.k32 ? 2 : 4;
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        int tl_n = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl_n) {
            OTAL_TL(tl_n, 9);
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            OTAL_TL(tl_n, 10);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
            int n_it = kiters;
            if constexpr (KS) {
                const int ks = tile % p.ksplit;
                n_it = kiters * (ks + 1) / p.ksplit - kiters * ks / p.ksplit;
            }
            const int it_seg1 = kiters - p.kchunks2;       // iterations of the first K segment: taps x chunks, chunk fastest
            int kc = 0;
            for (int it = 0; it < n_it; ++it) {
                int ks_here = ksteps;                      // the last chunk of a tap / of the second segment may hold fewer K steps
                if (it < it_seg1) {
                    if (++kc == p.kchunks) { kc = 0; ks_here = p.tail_ksteps; }
                } else if (it == n_it - 1) {
                    ks_here = p.tail_ksteps2;
                }
                mbar_wait(&full_bar[stage], phase);
                if (it == 0) OTAL_TL(tl_n, 3);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + (size_t)stage * L.stage_bytes);
                const uint32_t sB = sA + L.a_bytes;
                const uint64_t a_hi0 = tmpl_k + (uint64_t)(sA >> 4);
                const uint64_t a_lo0 = tmpl_k + (uint64_t)((sA + a_plane) >> 4);
                const uint64_t b_hi0 = tmpl_b + (uint64_t)(sB >> 4);
                const uint64_t b_lo0 = tmpl_b + (uint64_t)((sB + b_plane) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k >= ks_here) break;
                        const uint64_t a_hi = a_hi0 + (uint64_t)(k * 2);
                        const uint64_t b_hi = b_hi0 + (uint64_t)(k * b_step);
                        if constexpr (NCAT) {
                            umma_f16(d_tmem, a_hi, b_hi, idesc_cat, (it | k) != 0);          // [a_hi*b_hi | a_hi*b_lo]
                            if constexpr (!U8) umma_f16(d_tmem, a_lo0 + (uint64_t)(k * 2), b_hi, idesc, 1);   // += a_lo*b_hi
                            continue;
                        }
                        umma_f16(d_tmem, a_hi, b_hi, idesc, (it | k) != 0);
                        if (split) {
                            umma_f16(d_tmem, a_lo0 + (uint64_t)(k * 2), b_hi, idesc, 1);
                            umma_f16(d_tmem, a_hi, b_lo0 + (uint64_t)(k * b_step), idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tmem_full[acc]);
            __syncwarp();
            OTAL_TL(tl_n, 4);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (128 threads)
        // One warp per scheduler does all of this, so the loop has to be short and free of long-latency dependent loads:
        // the per-channel scale / shift of the tile's N block are staged in shared memory once per tile (ss), every launch
        // parameter the loop reads lives in a register, conversions use the packed bf16x2 form.  (Round-2 timeline: the
        // previous form — per-element __ldg of scale / shift behind branches, scalar conversions — took ~10 000 cycles per
        // 64-column chunk and was THE bound of every 1x1 convolution.)
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;           // accumulator row = position inside the tile box
        const int et = threadIdx.x - 128;        // 0..127
        unsigned char* stg = smem + L.staging_off;
        float2* ss = reinterpret_cast<float2*>(smem + L.bar_off + 256);       // [2][256] (scale, shift), by tile parity
        const uint32_t planes = split ? 2u : 1u;
        const int BN = p.BN, Cout = p.Cout, n_blocks = p.n_blocks;
        const int tW = p.tW, tH = p.tH, tT = p.tT, tilesW = p.tilesW, tilesH = p.tilesH, tilesT = p.tilesT;
        const bool store_bf16 = p.store_bf16 != 0, ncdhw = p.out_ncdhw != 0, accumulate = p.accumulate != 0, two_buf = p.nbuf == 2;
        const float* const scale_g = p.scale; const float* const shift_g = p.shift;
        float* const out_f32 = p.out_f32;
        const bool has_ss = (scale_g != nullptr) || (shift_g != nullptr && !U8);
        const float relu_floor = p.relu ? 0.f : -3.402823466e38f;
        const int nchunks = (BN + 63) / 64;
        int acc = 0; uint32_t acc_phase = 0;
        int sbuf = 0;
        int tl_n = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl_n) {
            int tdec = tile;
            bool first_split = true;                  // the bias / shift is added by ONE of the K shares
            if constexpr (KS) { first_split = (tdec % p.ksplit) == 0; tdec /= p.ksplit; }
            const int nb = tdec % n_blocks;
            int m = tdec / n_blocks;
            const int w0 = (m % tilesW) * tW; m /= tilesW;
            const int h0 = (m % tilesH) * tH; m /= tilesH;
            const int t0 = (m % tilesT) * tT; m /= tilesT;
            const int n = m;
            float2* sst = ss + (tl_n & 1) * 256;
            if (has_ss) {
                // columns of this N block: (scale, shift), identity beyond Cout; the K shares other than the first add no shift
                for (int c = et; c < BN; c += 128) {
                    const int cg = nb * BN + c;
                    float sc = 1.f, sh = 0.f;
                    if (cg < Cout) {
                        if (scale_g) sc = __ldg(scale_g + cg);
                        if (shift_g && !U8 && (!KS || first_split)) sh = __ldg(shift_g + cg);
                    }
                    sst[c] = make_float2(sc, sh);
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
            }
            float* orow = nullptr;
            int shift_off = 0;                       // U8: row of the border-class shift table for this position
            if (out_f32 != nullptr || U8) {
                // position of this thread's row (fp32 store / border class)
                const int rw = row % tW, rh = (row / tW) % tH, rt = row / (tW * tH);
                const bool row_ok = (w0 + rw < p.W) && (h0 + rh < p.H) && (t0 + rt < p.T);
                if constexpr (U8) {
                    auto cls = [](int o, int n_) { return o == 0 ? 1 : (o == n_ - 2 ? 2 : (o == n_ - 1 ? 3 : 0)); };
                    shift_off = ((cls(t0 + rt, p.T) * 4 + cls(h0 + rh, p.H)) * 4 + cls(w0 + rw, p.W)) * Cout;
                }
                if (out_f32 && row_ok) {
                    const size_t pos = ((size_t)(t0 + rt) * p.H + (h0 + rh)) * p.W + (w0 + rw);
                    if (ncdhw)   // element (n, c, pos): channel stride = T*H*W, consecutive rows -> consecutive addresses
                        orow = out_f32 + ((size_t)n * p.out_cstride + p.out_coff) * ((size_t)p.T * p.H * p.W) + pos;
                    else
                        orow = out_f32 + ((size_t)n * p.T * p.H * p.W + pos) * p.out_cstride + p.out_coff;
                }
            }

            if (warp == 4) OTAL_TL(tl_n, 5);
            mbar_wait(&tmem_full[acc], acc_phase);
            if (warp == 4) OTAL_TL(tl_n, 6);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)acc * kAccStride + ((uint32_t)(q * 32) << 16);
            for (int ch = 0; ch < nchunks; ++ch) {
                unsigned char* buf_hi = stg + (size_t)sbuf * planes * kATileBytes;
                unsigned char* buf_lo = buf_hi + kATileBytes;
                if (store_bf16) {
                    // the store that last read this staging buffer was committed nbuf chunks ago
                    if (et == 0) { if (two_buf) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int col0 = ch * 64 + half * 32;
                    if (col0 >= BN) break;
                    uint32_t v[32];
                    tmem_ld32(t_acc + col0, v);
                    float f[32];
                    if constexpr (NCAT) {                 // second half of the accumulator: the a_hi*b_lo products
                        uint32_t v2[32];
                        tmem_ld32(t_acc + BN + col0, v2);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    }
                    const int cbase = nb * BN + col0;     // channel inside the slice
                    if (has_ss) {
                        const float4* s4 = reinterpret_cast<const float4*>(sst + col0);      // 2 columns per 16-byte broadcast load
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float4 s2 = s4[j >> 1];
                            f[j] = fmaf(f[j], s2.x, s2.y);
                            f[j + 1] = fmaf(f[j + 1], s2.z, s2.w);
                        }
                    }
                    if constexpr (U8) {
                        // border-class shift table: rows differ between threads, channels beyond Cout do not exist (Cout % 8 == 0)
                        const float4* tb = reinterpret_cast<const float4*>(shift_g + shift_off + cbase);
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (cbase + j < Cout) {
                                const float4 t4 = __ldg(tb + (j >> 2));
                                f[j] += t4.x; f[j + 1] += t4.y; f[j + 2] += t4.z; f[j + 3] += t4.w;
                            }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], relu_floor);
                    if (orow && !ncdhw) {
                        // Cout, out_coff and out_cstride are multiples of 8: groups of 4 are all-in or all-out
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (cbase + j < Cout) {
                                float4* dst = reinterpret_cast<float4*>(orow + cbase + j);
                                float4 v4 = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                                if constexpr (KS) {           // partial sums of the K shares meet in the destination
                                    float* d1 = reinterpret_cast<float*>(dst);
                                    atomicAdd(d1, v4.x); atomicAdd(d1 + 1, v4.y); atomicAdd(d1 + 2, v4.z); atomicAdd(d1 + 3, v4.w);
                                    continue;
                                }
                                if (accumulate) { const float4 o = *dst; v4.x += o.x; v4.y += o.y; v4.z += o.z; v4.w += o.w; }
                                *dst = v4;
                            }
                    } else if (orow) {
                        // channel-major: for a fixed channel the 32 rows of a warp are 32 consecutive floats
                        const size_t cs = (size_t)p.T * p.H * p.W;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (cbase + j < Cout) {
                                float* dst = orow + (size_t)(cbase + j) * cs;
                                if constexpr (KS) { atomicAdd(dst, f[j]); continue; }
                                *dst = accumulate ? *dst + f[j] : f[j];
                            }
                    }
                    if (store_bf16) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {          // 4 x 16-byte groups (8 channels each)
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) split_bf16x2(f[g * 8 + e * 2], f[g * 8 + e * 2 + 1], hi[e], lo[e]);
                            const int j16 = half * 4 + g;                       // 16-byte chunk inside the 128-byte row
                            const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j16 ^ (row & 7)) << 4);
                            *reinterpret_cast<uint4*>(buf_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            if (split) *reinterpret_cast<uint4*>(buf_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
                if (store_bf16) {
                    fence_proxy_async_smem();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et == 0) {
                        const int cc = nb * BN + ch * 64;
                        tma_store_5d(&mapO_hi, buf_hi, cc, w0, h0, t0, n);
                        if (split) tma_store_5d(&mapO_lo, buf_lo, cc, w0, h0, t0, n);
                        tma_store_commit();
                    }
                    if (two_buf) sbuf ^= 1;
                }
                if (ch == 0 && warp == 4) OTAL_TL(tl_n, 7);
            }
            // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (warp == 4) OTAL_TL(tl_n, 8);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (store_bf16 && et == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// Everything the launcher needs besides the A maps (which differ between the generic and the folded conv1a path).
struct ConvLaunch {
    ConvParams p;
    const uint16_t *w_hi, *w_lo;      // [taps][Cout][KW] bf16 planes, KW = weight row width (Cin or 64 for conv1a);
                                      // dgrad (b_mn): [taps][K = w_k][N = Cout]
    int w_k;                          // reduction width in elements
    const uint16_t *w2_hi = nullptr, *w2_lo = nullptr;   // second K segment weights (w2_k reduction elements), or null
    int w2_k = 0;
    uint16_t *y_hi, *y_lo;
    int To, Ho, Wo;                   // output extent
};

static int finish_and_launch(ConvLaunch& L, ConvMaps& maps, cudaStream_t stream) {
    ConvParams& p = L.p;
    const bool split = p.nsplit == 3;
    p.T = L.To; p.H = L.Ho; p.W = L.Wo;      // the kernel tiles OUTPUT positions
    p.tilesT = (p.T + p.tT - 1) / p.tT; p.tilesH = (p.H + p.tH - 1) / p.tH; p.tilesW = (p.W + p.tW - 1) / p.tW;
    if (p.Cout <= 256) { p.n_blocks = 1; p.BN = (p.Cout + 15) / 16 * 16; }
    else {
        int best = 256, best_pad = 1 << 30;
        for (int bn = 256; bn >= 64; bn -= 64) {
            int pad = (p.Cout + bn - 1) / bn * bn;
            if (pad < best_pad) { best_pad = pad; best = bn; }
        }
        p.BN = best; p.n_blocks = (p.Cout + best - 1) / best;
    }
    // Small problems (the 1-D head: 16..512 positions, 512 channels) are bound by ONE SM streaming its weight block
    // from L2 (~42 B/clk per SM): narrow the N block so that more CTAs share the weight traffic.
    {
        const int m_tiles = p.N * p.tilesT * p.tilesH * p.tilesW;
        while (p.BN > 64 && p.BN % 64 == 0 && m_tiles * p.n_blocks < num_sms() / 2) {
            p.BN -= 64;
            p.n_blocks = (p.Cout + p.BN - 1) / p.BN;
        }
        if (p.BN > 64 && p.BN % 64 != 0 && p.Cout > 64 && m_tiles * p.n_blocks < num_sms() / 2) {
            p.BN = 64; p.n_blocks = (p.Cout + 63) / 64;
        }
    }
    {
        // The widest tiles (BN > 128 in bf16x3: 96 KB per stage) leave no room for the epilogue's (scale, shift) table next to
        // two stages and a staging buffer: such a launch runs as 128-wide N blocks instead (which also get the N-concatenated form).
        const uint32_t cap = 227 * 1024 - 1024;
        const int nb_try = L.y_hi != nullptr ? 1 : 0;
        if (p.BN > 128 && conv_smem_layout(p.BN, p.nsplit, 2, nb_try, p.b_mn, p.k32, p.a_single).total > cap) {
            p.BN = 128; p.n_blocks = (p.Cout + 127) / 128;
        }
    }
    {
        // N-concatenated hi|lo weights (see ConvParams::ncat): needs 2*BN accumulator columns and B_lo contiguous after B_hi
        static const bool off = getenv("OTAL_NO_NCAT") != nullptr;
        const uint32_t rowb = p.k32 ? 64u : 128u;
        p.ncat = (!off && split && 2 * p.BN <= kAccStride && (conv_b_rows(p.BN, p.b_mn) * rowb) % 1024u == 0 &&
                  (!p.b_mn || p.BN % 64 == 0)) ? 1 : 0;
        if (p.a_single) {
            // the single-plane A operand exists only in the N-concatenated form (not switchable by OTAL_NO_NCAT)
            p.ncat = (split && 2 * p.BN <= kAccStride && (conv_b_rows(p.BN, p.b_mn) * rowb) % 1024u == 0 && !p.b_mn) ? 1 : 0;
            if (!p.ncat) { set_last_error_msg("conv (u8): needs bf16x3 weights and Cout <= 128"); return OTAL_ERR_UNSUPPORTED; }
        }
    }
    const int chunk = p.k32 ? 32 : kChunkK;
    p.kchunks = (L.w_k + chunk - 1) / chunk;
    static const bool no_ktail = getenv("OTAL_NO_KTAIL") != nullptr;          // developer A/B: issue the zero K steps as before
    p.tail_ksteps = no_ktail ? chunk / 16 : (L.w_k - (p.kchunks - 1) * chunk + 15) / 16;
    p.tail_ksteps2 = (no_ktail || L.w2_k <= 0) ? kChunkK / 16 : (L.w2_k - (p.kchunks2 - 1) * kChunkK + 15) / 16;
    p.store_bf16 = L.y_hi != nullptr;
    p.total_tiles = p.N * p.tilesT * p.tilesH * p.tilesW * p.n_blocks;
    if (p.ksplit > 1) {
        // split-K over CTAs was staged in round 1 (template parameter KS) and withdrawn in round 2: -0.15 ms per step at best, and
        // never validated on the explicit head schedule.  The instantiation is not built; the descriptor field must be 0 or 1.
        set_last_error_msg("conv: ksplit is not supported (withdrawn experiment)");
        return OTAL_ERR_UNSUPPORTED;
    }
    p.ksplit = 1;

    const uint32_t smem_cap = 227 * 1024 - 1024;  // minus alignment slack
    // prefer (>= 3 stages, double-buffered staging), then (3 stages, 1 buffer), then (2 stages, 2 buffers), then (2 stages, 1 buffer)
    int nst = 0, nbuf = p.store_bf16 ? 2 : 0;
    for (int s = kMaxStages; s >= 3 && !nst; --s)
        if (conv_smem_layout(p.BN, p.nsplit, s, nbuf, p.b_mn, p.k32, p.a_single).total <= smem_cap) nst = s;
    {
        // a third pipeline stage with ONE staging buffer before two stages with two (fits BN <= 112 in bf16x3): more operand bytes
        // in flight for the HBM-bound 1x1 convs; 17.42 / 17.36 / 17.46 / 17.35 ms per step off / on / off / on
        // (profiles/r02_prefer_stages_ab.txt).  OTAL_CONV_NO_PREFER_STAGES=1: the old order (developer A/B)
        static const bool prefer_stages = getenv("OTAL_CONV_NO_PREFER_STAGES") == nullptr;
        if (!nst && prefer_stages && p.store_bf16 &&
            conv_smem_layout(p.BN, p.nsplit, 3, 1, p.b_mn, p.k32, p.a_single).total <= smem_cap) { nst = 3; nbuf = 1; }
    }
    if (!nst && conv_smem_layout(p.BN, p.nsplit, 2, nbuf, p.b_mn, p.k32, p.a_single).total <= smem_cap) nst = 2;
    if (!nst && p.store_bf16 && conv_smem_layout(p.BN, p.nsplit, 2, 1, p.b_mn, p.k32, p.a_single).total <= smem_cap) { nst = 2; nbuf = 1; }
    if (!nst) { set_last_error_msg("conv: tile does not fit shared memory"); return OTAL_ERR_UNSUPPORTED; }
    p.nstages = nst; p.nbuf = nbuf;
    const ConvSmem SL = conv_smem_layout(p.BN, p.nsplit, nst, nbuf, p.b_mn, p.k32, p.a_single);

    int rc;
    const int ntaps = p.kt * p.kh * p.kw;
    // forward: rows of w_k reduction elements per output channel; dgrad: rows of Cout (N) elements per reduction index
    const uint64_t brow = p.b_mn ? (uint64_t)p.Cout : (uint64_t)L.w_k;
    const uint64_t bcol = p.b_mn ? (uint64_t)L.w_k : (uint64_t)p.Cout;
    const uint64_t bdims[3] = {brow, bcol, (uint64_t)ntaps};
    const uint64_t bst[2] = {brow * 2, brow * 2 * bcol};
    const uint32_t bbox[3] = {(uint32_t)(p.k32 ? 32 : 64), p.b_mn ? 64u : (uint32_t)p.BN, 1};
    const int bswz = p.k32 ? 2 : 1;                                   // 2 = SWIZZLE_64B (make_tensor_map_bf16)
    if ((rc = make_tensor_map_bf16(&maps.B_hi, L.w_hi, 3, bdims, bst, bbox, bswz))) return rc;
    if (split && (rc = make_tensor_map_bf16(&maps.B_lo, L.w_lo, 3, bdims, bst, bbox, bswz))) return rc;
    if (L.w2_k > 0) {
        const uint64_t brow2 = p.b_mn ? (uint64_t)p.Cout : (uint64_t)L.w2_k;
        const uint64_t bcol2 = p.b_mn ? (uint64_t)L.w2_k : (uint64_t)p.Cout;
        const uint64_t bdims2[3] = {brow2, bcol2, 1};
        const uint64_t bst2[2] = {brow2 * 2, brow2 * 2 * bcol2};
        if ((rc = make_tensor_map_bf16(&maps.B2_hi, L.w2_hi, 3, bdims2, bst2, bbox, 1))) return rc;
        if (split && (rc = make_tensor_map_bf16(&maps.B2_lo, L.w2_lo, 3, bdims2, bst2, bbox, 1))) return rc;
    }
    if (p.store_bf16) {
        const uint32_t obox[5] = {64, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tT, 1};
        const uint64_t odims[5] = {(uint64_t)p.Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.T, (uint64_t)p.N};
        const uint64_t cs = (uint64_t)p.out_cstride * 2;
        const uint64_t ost[4] = {cs, cs * p.W, cs * p.W * p.H, cs * p.W * p.H * p.T};
        if ((rc = make_tensor_map_bf16(&maps.O_hi, L.y_hi + p.out_coff, 5, odims, ost, obox, 1))) return rc;
        if (split && (rc = make_tensor_map_bf16(&maps.O_lo, L.y_lo + p.out_coff, 5, odims, ost, obox, 1))) return rc;
    }

#ifdef OTAL_TIMELINE
    p.timeline = g_timeline;
    fprintf(stderr, "otal timeline: BN %d n_blocks %d stages %d nbuf %d ncat %d kchunks %d taps %d tiles %d b_mn %d f32 %d\n", p.BN, p.n_blocks,
            p.nstages, p.nbuf, p.ncat, p.kchunks, p.kt * p.kh * p.kw, p.total_tiles, p.b_mn, p.out_f32 != nullptr);
#endif
    const size_t smem_bytes = SL.total + 1024;
    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        OTAL_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        once.mark(once_dev);
    }
    if (p.a_single) {
        static OncePerDevice once_u8;
        if (once_u8.need(&once_dev)) {
            auto* kernel_u8 = conv_igemm_kernel<true, true>;
            OTAL_CUDA_TRY(cudaFuncSetAttribute(kernel_u8, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            once_u8.mark(once_dev);
        }
    }
    int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    if (p.a_single) conv_igemm_kernel<true, true><<<grid, kConvThreads, smem_bytes, stream>>>(maps, p);
    else if (p.ncat) conv_igemm_kernel<true><<<grid, kConvThreads, smem_bytes, stream>>>(maps, p);
    else conv_igemm_kernel<false><<<grid, kConvThreads, smem_bytes, stream>>>(maps, p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // namespace otal

using namespace otal;

extern "C" {

#ifdef OTAL_TIMELINE
// developer builds only: [grid][16 tiles][16 slots] of clock64() values, or nullptr to switch the recording off
__attribute__((visibility("default"))) void otal_debug_set_timeline(unsigned long long* buf) { otal::g_timeline = buf; }
#endif

// See include/opental_b200.h for the contract.
int otal_conv_igemm_fwd(const otal_conv_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("conv: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (d->N <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) {
        set_last_error_msg("conv: non-positive dimension"); return OTAL_ERR_BAD_ARG;
    }
    const bool ncdhw = d->y_f32_ncdhw != 0;   // channel-major fp32 output: no vector stores, no alignment constraint
    if (ncdhw && d->y_hi) { set_last_error_msg("conv: NCDHW output is fp32 only"); return OTAL_ERR_BAD_ARG; }
    if (d->Cin % 8 || d->in_cstride % 8 || d->in_coff % 8 ||
        (!ncdhw && (d->out_cstride % 8 || d->out_coff % 8 || d->Cout % 8))) {
        set_last_error_msg("conv: channel counts / strides / offsets must be multiples of 8 (16-byte TMA rows)");
        return OTAL_ERR_BAD_ARG;
    }
    if (d->tT * d->tH * d->tW != kTileM) { set_last_error_msg("conv: tile box must hold 128 positions"); return OTAL_ERR_BAD_ARG; }
    if (d->nsplit != 1 && d->nsplit != 3) { set_last_error_msg("conv: nsplit must be 1 or 3"); return OTAL_ERR_BAD_ARG; }
    const int st = d->sT ? d->sT : 1, sh = d->sH ? d->sH : 1, sw = d->sW ? d->sW : 1;
    if ((st != 1 && st != 2) || (sh != 1 && sh != 2) || (sw != 1 && sw != 2)) {
        set_last_error_msg("conv: stride must be 1 or 2"); return OTAL_ERR_BAD_ARG;
    }
    const bool split = d->nsplit == 3;
    if (!d->x_hi || !d->w_hi || (split && (!d->x_lo || !d->w_lo))) { set_last_error_msg("conv: null operand plane"); return OTAL_ERR_BAD_ARG; }
    if (d->y_hi && split && !d->y_lo) { set_last_error_msg("conv: y_lo missing"); return OTAL_ERR_BAD_ARG; }
    if (!d->y_hi && !d->y_f32) { set_last_error_msg("conv: no destination"); return OTAL_ERR_BAD_ARG; }
    if (d->dgrad && (st != 1 || sh != 1 || sw != 1)) { set_last_error_msg("conv: dgrad mode is stride-1 only"); return OTAL_ERR_BAD_ARG; }
    if (d->accumulate && !d->y_f32) { set_last_error_msg("conv: accumulate needs the fp32 destination"); return OTAL_ERR_BAD_ARG; }

    ConvLaunch L{};
    ConvParams& p = L.p;
    p.N = d->N; p.Cin = d->Cin; p.Cout = d->Cout;
    p.kt = d->kt; p.kh = d->kh; p.kw = d->kw; p.pt = d->pt; p.ph = d->ph; p.pw = d->pw;
    p.st = st; p.sh = sh; p.sw = sw;
    p.tT = d->tT; p.tH = d->tH; p.tW = d->tW;
    p.nsplit = d->nsplit; p.relu = d->relu; p.accumulate = d->accumulate; p.b_mn = d->dgrad ? 1 : 0;
    p.out_ncdhw = d->y_f32_ncdhw ? 1 : 0;
    p.scale = d->scale; p.shift = d->shift; p.out_f32 = d->y_f32;
    p.out_cstride = d->out_cstride; p.out_coff = d->out_coff;
    p.ksplit = d->ksplit > 1 ? d->ksplit : 1;
    L.w_hi = d->w_hi; L.w_lo = d->w_lo; L.w_k = d->Cin; L.y_hi = d->y_hi; L.y_lo = d->y_lo;
    L.To = (d->T + st - 1) / st; L.Ho = (d->H + sh - 1) / sh; L.Wo = (d->W + sw - 1) / sw;

    ConvMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    // activations: dims (C, W, H, T, N), channel slice [in_coff, in_coff+Cin) of rows of in_cstride channels; one
    // dense view per parity class of the strided dims
    const uint64_t cs = (uint64_t)d->in_cstride * 2;
    const uint64_t sW_ = cs, sH_ = cs * d->W, sT_ = cs * d->W * d->H, sN_ = cs * d->W * d->H * d->T;
    const uint32_t abox[5] = {64, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tT, 1};
    for (int rt = 0; rt < st; ++rt) for (int rh = 0; rh < sh; ++rh) for (int rw = 0; rw < sw; ++rw) {
        const int eT = (d->T - rt + st - 1) / st, eH = (d->H - rh + sh - 1) / sh, eW = (d->W - rw + sw - 1) / sw;
        if (eT <= 0 || eH <= 0 || eW <= 0) continue;
        const uint64_t adims[5] = {(uint64_t)p.Cin, (uint64_t)eW, (uint64_t)eH, (uint64_t)eT, (uint64_t)p.N};
        const uint64_t ast[4] = {sW_ * sw, sH_ * sh, sT_ * st, sN_};
        const size_t off = (size_t)d->in_coff + ((size_t)rt * d->H * d->W + (size_t)rh * d->W + rw) * d->in_cstride;
        const int mi = rt * 4 + rh * 2 + rw;
        if ((rc = make_tensor_map_bf16(&maps.A_hi[mi], d->x_hi + off, 5, adims, ast, abox, 1))) return rc;
        if (split && (rc = make_tensor_map_bf16(&maps.A_lo[mi], d->x_lo + off, 5, adims, ast, abox, 1))) return rc;
    }
    if (d->Cin2 > 0) {
        // second K segment: x2 [N,T,H,W,in2_cstride] planes (channels [in2_coff, in2_coff + Cin2)), w2 [1][Cout][Cin2]
        // (dgrad: [1][Cin2][Cout]); 1x1 stride-1 only
        if (d->kt * d->kh * d->kw != 1 || st != 1 || sh != 1 || sw != 1 || d->Cin2 % 8 || d->in2_cstride % 8 || d->in2_coff % 8 ||
            !d->x2_hi || !d->w2_hi || (split && (!d->x2_lo || !d->w2_lo))) {
            set_last_error_msg("conv: the second K segment needs a 1x1 stride-1 conv and complete x2 / w2 planes"); return OTAL_ERR_BAD_ARG;
        }
        p.kchunks2 = (d->Cin2 + kChunkK - 1) / kChunkK;
        const uint64_t cs2 = (uint64_t)d->in2_cstride * 2;
        const uint64_t adims[5] = {(uint64_t)d->Cin2, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->T, (uint64_t)p.N};
        const uint64_t ast[4] = {cs2, cs2 * d->W, cs2 * d->W * d->H, cs2 * d->W * d->H * d->T};
        if ((rc = make_tensor_map_bf16(&maps.A2_hi, d->x2_hi + d->in2_coff, 5, adims, ast, abox, 1))) return rc;
        if (split && (rc = make_tensor_map_bf16(&maps.A2_lo, d->x2_lo + d->in2_coff, 5, adims, ast, abox, 1))) return rc;
        L.w2_hi = d->w2_hi; L.w2_lo = d->w2_lo; L.w2_k = d->Cin2;
    }
    return finish_and_launch(L, maps, stream);
}

// Conv3d_1a_7x7 (AFSD/common/i3d_backbone.py:196-199): 7x7x7, stride 2, 3 input channels.
// The clip is stored W-padded with 4 channel slots per pixel, [N,T,H,W+8,4] (otal_clip_ingest): the 7 W-taps x 4 slots of
// one (dt,dh) tap of output column w' are 28 contiguous values inside the 32-element (64-byte) window that starts at
// padded column 2*w' (8th W-tap and channel 3 carry zero weights): K = 49 x 32 instead of 343 x 3 = 1029 useful (1.5x
// padding).  Operand rows are 64 bytes -> SWIZZLE_64B TMA boxes and UMMA descriptors.  T and H use the stride-2 parity
// views with TMA zero fill as padding; W padding is physical.  (A window-expanded copy of the clip — dense TMA rows —
// was measured: same kernel time, 4x the clip traffic; the strided map is kept.)
static int conv1a_fwd_impl(const otal_conv1a_desc* d, void* stream_, bool u8) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("conv1a: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (u8) {
        // raw-pixel form: one exact A plane, hi/lo weights, per-border-class shift table (see otal_conv1a_fwd_u8)
        if (d->nsplit != 3 || !d->shift || d->T % 2 || d->H % 2 || d->T < 6 || d->H < 6 || d->W < 6) {
            set_last_error_msg("conv1a_u8: needs nsplit 3, a shift table, even T / H and extents >= 6"); return OTAL_ERR_BAD_ARG;
        }
    }
    if (d->N <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->Cout <= 0 || d->Cout % 8 || d->out_cstride % 8 || d->out_coff % 8) {
        set_last_error_msg("conv1a: bad dimension"); return OTAL_ERR_BAD_ARG;
    }
    if (d->W % 2) { set_last_error_msg("conv1a: W must be even"); return OTAL_ERR_BAD_ARG; }
    if (d->tT * d->tH * d->tW != kTileM) { set_last_error_msg("conv1a: tile box must hold 128 positions"); return OTAL_ERR_BAD_ARG; }
    if (d->nsplit != 1 && d->nsplit != 3) { set_last_error_msg("conv1a: nsplit must be 1 or 3"); return OTAL_ERR_BAD_ARG; }
    const bool split = d->nsplit == 3;
    if (!d->x_hi || !d->w_hi || (split && ((!u8 && !d->x_lo) || !d->w_lo)) || !d->y_hi || (split && !d->y_lo)) {
        set_last_error_msg("conv1a: null plane"); return OTAL_ERR_BAD_ARG;
    }
    ConvLaunch L{};
    ConvParams& p = L.p;
    p.N = d->N; p.Cin = 32; p.Cout = d->Cout;
    p.k32 = 1;                                                  // 8 W taps x 4 channels = 32-element K chunks, SWIZZLE_64B
    p.a_single = u8 ? 1 : 0; p.shift_classes = u8 ? 1 : 0;
    p.kt = 7; p.kh = 7; p.kw = 1;
    // "same" padding of k=7, s=2 on an even extent: total 5, front 2 (i3d_backbone.py:45-69)
    p.pt = (d->T % 2 == 0) ? 2 : 3; p.ph = (d->H % 2 == 0) ? 2 : 3; p.pw = 0;
    p.st = 2; p.sh = 2; p.sw = 1;
    p.tT = d->tT; p.tH = d->tH; p.tW = d->tW;
    p.nsplit = d->nsplit; p.relu = d->relu; p.accumulate = 0;
    p.scale = d->scale; p.shift = d->shift; p.out_f32 = nullptr;
    p.out_cstride = d->out_cstride; p.out_coff = d->out_coff;
    L.w_hi = d->w_hi; L.w_lo = d->w_lo; L.w_k = 32; L.y_hi = d->y_hi; L.y_lo = d->y_lo;
    L.To = (d->T + 1) / 2; L.Ho = (d->H + 1) / 2; L.Wo = d->W / 2;

    ConvMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    const uint64_t px = 4 * 2;                                  // bytes per padded pixel (4 channel slots)
    const uint64_t Wp = (uint64_t)d->W + 8;                     // padded row: 2 zero pixels left, 6 right (otal_clip_ingest)
    const uint64_t sH_ = px * Wp, sT_ = sH_ * d->H, sN_ = sT_ * d->T;
    const uint32_t abox[5] = {32, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tT, 1};
    for (int rt = 0; rt < 2; ++rt) for (int rh = 0; rh < 2; ++rh) {
        const int eT = (d->T - rt + 1) / 2, eH = (d->H - rh + 1) / 2;
        if (eT <= 0 || eH <= 0) continue;
        // dim0 = the 32-element window, dim1 = output column: the window origin advances 2 pixels = 16 bytes, i.e. the
        // windows of neighbouring columns overlap in memory (strided tensor map, nothing is duplicated in HBM)
        const uint64_t adims[5] = {32, (uint64_t)L.Wo, (uint64_t)eH, (uint64_t)eT, (uint64_t)p.N};
        const uint64_t ast[4] = {2 * px, sH_ * 2, sT_ * 2, sN_};
        const size_t off = ((size_t)rt * d->H + rh) * Wp * 4;
        const int mi = rt * 4 + rh * 2;
        if ((rc = make_tensor_map_bf16(&maps.A_hi[mi], d->x_hi + off, 5, adims, ast, abox, 2))) return rc;
        if (split && !u8 && (rc = make_tensor_map_bf16(&maps.A_lo[mi], d->x_lo + off, 5, adims, ast, abox, 2))) return rc;
    }
    return finish_and_launch(L, maps, stream);
}

int otal_conv1a_fwd(const otal_conv1a_desc* d, void* stream) { return conv1a_fwd_impl(d, stream, false); }

// Conv3d_1a on the RAW uint8 clip (otal_clip_ingest_u8_raw: pixel values 0..255, exact in ONE bf16 plane; x_lo is ignored).
// With x = (2/255) u - 1 inside the image and the reference's zero padding of x outside it,
//   conv(x, W)[p, co] = (2/255) * sum_taps W * u_zero-padded  -  sum_{taps inside the image at p} W,
// so the caller passes scale = bn_scale * 2/255 and `shift` = a table [4][4][4][Cout] over the border classes of the output
// position (ConvParams::shift_classes; entry = bn_shift - bn_scale * sum of the in-bounds weights).  One tensor-core pass
// (u * [w_hi | w_lo], N-concatenated) instead of the two of the bf16x3 form, and half the activation fill; at least as
// accurate (the activation operand has no rounding error at all).
int otal_conv1a_fwd_u8(const otal_conv1a_desc* d, void* stream) { return conv1a_fwd_impl(d, stream, true); }

}  // extern "C"

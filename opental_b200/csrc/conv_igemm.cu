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
    const float* scale;  // [Cout] or nullptr (=1)
    const float* shift;  // [Cout] or nullptr (=0)
    float* out_f32;      // optional NDHWC fp32 destination (nullptr = skip)
};

struct ConvSmem {
    // dynamic shared memory carve-up, all offsets relative to a 1024-byte aligned base
    uint32_t stage_bytes, a_bytes, b_bytes;
    uint32_t staging_off, bar_off, total;
};

__host__ __device__ inline ConvSmem conv_smem_layout(int BN, int nsplit, int nstages, int nbuf) {
    ConvSmem s;
    const uint32_t planes = nsplit == 3 ? 2u : 1u;
    s.a_bytes = kATileBytes * planes;
    s.b_bytes = (uint32_t)BN * 128u * planes;
    s.stage_bytes = s.a_bytes + s.b_bytes;
    s.staging_off = s.stage_bytes * (uint32_t)nstages;
    uint32_t staging = (uint32_t)nbuf * planes * kATileBytes;      // nbuf (0, 1 or 2) buffers x planes x 16 KB
    s.bar_off = s.staging_off + staging;
    s.total = s.bar_off + 256;
    return s;
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                  const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                  const __grid_constant__ CUtensorMap mapO_hi, const __grid_constant__ CUtensorMap mapO_lo,
                  const ConvParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment is required by the 128B swizzle pattern (pattern repeats every 8 rows x 128 B)
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const ConvSmem L = conv_smem_layout(p.BN, p.nsplit, p.nstages, p.nbuf);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool split = p.nsplit == 3;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA_hi);
        tma_prefetch_desc(&mapB_hi);
        if (split) { tma_prefetch_desc(&mapA_lo); tma_prefetch_desc(&mapB_lo); }
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
    const int kiters = ntaps * p.kchunks;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nb = tile % p.n_blocks;
                int m = tile / p.n_blocks;
                const int w0 = (m % p.tilesW) * p.tW; m /= p.tilesW;
                const int h0 = (m % p.tilesH) * p.tH; m /= p.tilesH;
                const int t0 = (m % p.tilesT) * p.tT; m /= p.tilesT;
                const int n = m;
                for (int tap = 0; tap < ntaps; ++tap) {
                    const int dw = tap % p.kw, dh = (tap / p.kw) % p.kh, dt = tap / (p.kw * p.kh);
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        unsigned char* sA = smem + (size_t)stage * L.stage_bytes;
                        unsigned char* sB = sA + L.a_bytes;
                        mbar_expect_tx(&full_bar[stage], L.stage_bytes);
                        const int c0 = kc * kChunkK;
                        tma_load_5d(&mapA_hi, &full_bar[stage], sA, c0, w0 + dw - p.pw, h0 + dh - p.ph,
                                    t0 + dt - p.pt, n);
                        tma_load_3d(&mapB_hi, &full_bar[stage], sB, c0, nb * p.BN, tap);
                        if (split) {
                            tma_load_5d(&mapA_lo, &full_bar[stage], sA + kATileBytes, c0, w0 + dw - p.pw,
                                        h0 + dh - p.ph, t0 + dt - p.pt, n);
                            tma_load_3d(&mapB_lo, &full_bar[stage], sB + (size_t)p.BN * 128, c0, nb * p.BN, tap);
                        }
                        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kTileM, p.BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + (size_t)stage * L.stage_bytes);
                    const uint32_t sB = sA + L.a_bytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t a_hi = umma_smem_desc_sw128(sA + k * 32, 16, 1024);
                        const uint64_t b_hi = umma_smem_desc_sw128(sB + k * 32, 16, 1024);
                        umma_f16(d_tmem, a_hi, b_hi, idesc, (it | k) != 0);
                        if (split) {
                            const uint64_t a_lo = umma_smem_desc_sw128(sA + kATileBytes + k * 32, 16, 1024);
                            const uint64_t b_lo = umma_smem_desc_sw128(sB + p.BN * 128 + k * 32, 16, 1024);
                            umma_f16(d_tmem, a_lo, b_hi, idesc, 1);
                            umma_f16(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == p.nstages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (128 threads)
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;           // accumulator row = position inside the tile box
        const int et = threadIdx.x - 128;        // 0..127
        unsigned char* stg = smem + L.staging_off;
        const uint32_t planes = split ? 2u : 1u;
        int acc = 0; uint32_t acc_phase = 0;
        int sbuf = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int nb = tile % p.n_blocks;
            int m = tile / p.n_blocks;
            const int w0 = (m % p.tilesW) * p.tW; m /= p.tilesW;
            const int h0 = (m % p.tilesH) * p.tH; m /= p.tilesH;
            const int t0 = (m % p.tilesT) * p.tT; m /= p.tilesT;
            const int n = m;
            // position of this thread's row (for the optional fp32 store)
            const int rw = row % p.tW, rh = (row / p.tW) % p.tH, rt = row / (p.tW * p.tH);
            const bool row_ok = (w0 + rw < p.W) && (h0 + rh < p.H) && (t0 + rt < p.T);
            float* orow = nullptr;
            if (p.out_f32 && row_ok) {
                size_t pos = (((size_t)n * p.T + (t0 + rt)) * p.H + (h0 + rh)) * p.W + (w0 + rw);
                orow = p.out_f32 + pos * p.out_cstride + p.out_coff;
            }

            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)acc * kAccStride + ((uint32_t)(q * 32) << 16);
            const int nchunks = (p.BN + 63) / 64;
            for (int ch = 0; ch < nchunks; ++ch) {
                unsigned char* buf_hi = stg + (size_t)sbuf * planes * kATileBytes;
                unsigned char* buf_lo = buf_hi + kATileBytes;
                if (p.store_bf16) {
                    // the store that last read this staging buffer was committed nbuf chunks ago
                    if (et == 0) { if (p.nbuf == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int col0 = ch * 64 + half * 32;
                    if (col0 >= p.BN) break;
                    uint32_t v[32];
                    tmem_ld32(t_acc + col0, v);
                    tmem_ld_wait();
                    const int cbase = nb * p.BN + col0;   // channel inside the slice
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int c = cbase + j;
                        float sc = 1.f, sh = 0.f;
                        if (c < p.Cout) {
                            if (p.scale) sc = __ldg(p.scale + c);
                            if (p.shift) sh = __ldg(p.shift + c);
                        }
                        float x = fmaf(__uint_as_float(v[j]), sc, sh);
                        if (p.relu) x = fmaxf(x, 0.f);
                        f[j] = x;
                    }
                    if (orow) {
                        // Cout, out_coff and out_cstride are multiples of 8: groups of 4 are all-in or all-out
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (cbase + j < p.Cout)
                                *reinterpret_cast<float4*>(orow + cbase + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    }
                    if (p.store_bf16) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {          // 4 x 16-byte groups (8 channels each)
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                __nv_bfloat16 h0b, l0b, h1b, l1b;
                                split_bf16(f[g * 8 + e * 2], h0b, l0b);
                                split_bf16(f[g * 8 + e * 2 + 1], h1b, l1b);
                                hi[e] = pack_bf16x2(h0b, h1b);
                                lo[e] = pack_bf16x2(l0b, l1b);
                            }
                            const int j16 = half * 4 + g;                       // 16-byte chunk inside the 128-byte row
                            const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j16 ^ (row & 7)) << 4);
                            *reinterpret_cast<uint4*>(buf_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            if (split) *reinterpret_cast<uint4*>(buf_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
                if (p.store_bf16) {
                    fence_proxy_async_smem();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et == 0) {
                        const int cc = nb * p.BN + ch * 64;
                        tma_store_5d(&mapO_hi, buf_hi, cc, w0, h0, t0, n);
                        if (split) tma_store_5d(&mapO_lo, buf_lo, cc, w0, h0, t0, n);
                        tma_store_commit();
                    }
                    if (p.nbuf == 2) sbuf ^= 1;
                }
            }
            // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.store_bf16 && et == 0) tma_store_wait_all<0>();
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

}  // namespace otal

using namespace otal;

extern "C" {

// See include/opental_b200.h for the contract.
int otal_conv_igemm_fwd(const otal_conv_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("conv: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (d->N <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) {
        set_last_error_msg("conv: non-positive dimension"); return OTAL_ERR_BAD_ARG;
    }
    if (d->Cin % 8 || d->in_cstride % 8 || d->in_coff % 8 || d->out_cstride % 8 || d->out_coff % 8 || d->Cout % 8) {
        set_last_error_msg("conv: channel counts / strides / offsets must be multiples of 8 (16-byte TMA rows)");
        return OTAL_ERR_BAD_ARG;
    }
    if (d->tT * d->tH * d->tW != kTileM) { set_last_error_msg("conv: tile box must hold 128 positions"); return OTAL_ERR_BAD_ARG; }
    if (d->nsplit != 1 && d->nsplit != 3) { set_last_error_msg("conv: nsplit must be 1 or 3"); return OTAL_ERR_BAD_ARG; }
    const bool split = d->nsplit == 3;
    if (!d->x_hi || !d->w_hi || (split && (!d->x_lo || !d->w_lo))) { set_last_error_msg("conv: null operand plane"); return OTAL_ERR_BAD_ARG; }
    const int store_bf16 = d->y_hi != nullptr;
    if (store_bf16 && split && !d->y_lo) { set_last_error_msg("conv: y_lo missing"); return OTAL_ERR_BAD_ARG; }
    if (!store_bf16 && !d->y_f32) { set_last_error_msg("conv: no destination"); return OTAL_ERR_BAD_ARG; }

    ConvParams p{};
    p.N = d->N; p.T = d->T; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
    p.kt = d->kt; p.kh = d->kh; p.kw = d->kw; p.pt = d->pt; p.ph = d->ph; p.pw = d->pw;
    p.tT = d->tT; p.tH = d->tH; p.tW = d->tW;
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
    p.kchunks = (p.Cin + kChunkK - 1) / kChunkK;
    p.nsplit = d->nsplit; p.relu = d->relu; p.store_bf16 = store_bf16;
    p.total_tiles = p.N * p.tilesT * p.tilesH * p.tilesW * p.n_blocks;
    p.scale = d->scale; p.shift = d->shift; p.out_f32 = d->y_f32;
    p.out_cstride = d->out_cstride; p.out_coff = d->out_coff;

    const uint32_t smem_cap = 227 * 1024 - 1024;  // minus alignment slack
    // prefer (>= 3 stages, double-buffered staging), then (2 stages, 2 buffers), then (2 stages, 1 buffer)
    int nst = 0, nbuf = store_bf16 ? 2 : 0;
    for (int s = kMaxStages; s >= 3 && !nst; --s)
        if (conv_smem_layout(p.BN, p.nsplit, s, nbuf).total <= smem_cap) nst = s;
    if (!nst && conv_smem_layout(p.BN, p.nsplit, 2, nbuf).total <= smem_cap) nst = 2;
    if (!nst && store_bf16 && conv_smem_layout(p.BN, p.nsplit, 2, 1).total <= smem_cap) { nst = 2; nbuf = 1; }
    if (!nst) { set_last_error_msg("conv: tile does not fit shared memory"); return OTAL_ERR_UNSUPPORTED; }
    p.nstages = nst; p.nbuf = nbuf;
    const ConvSmem L = conv_smem_layout(p.BN, p.nsplit, nst, nbuf);

    CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mO_hi, mO_lo;
    memset(&mA_lo, 0, sizeof(mA_lo)); memset(&mB_lo, 0, sizeof(mB_lo));
    memset(&mO_hi, 0, sizeof(mO_hi)); memset(&mO_lo, 0, sizeof(mO_lo));
    int rc;
    // activations: dims (C, W, H, T, N), channel slice [in_coff, in_coff+Cin) of rows of in_cstride channels
    const uint64_t adims[5] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.T, (uint64_t)p.N};
    const uint64_t ast[4] = {(uint64_t)d->in_cstride * 2, (uint64_t)d->in_cstride * 2 * p.W,
                             (uint64_t)d->in_cstride * 2 * p.W * p.H, (uint64_t)d->in_cstride * 2 * p.W * p.H * p.T};
    const uint32_t abox[5] = {64, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tT, 1};
    if ((rc = make_tensor_map_bf16(&mA_hi, d->x_hi + d->in_coff, 5, adims, ast, abox, /*swizzle128=*/1))) return rc;
    if (split && (rc = make_tensor_map_bf16(&mA_lo, d->x_lo + d->in_coff, 5, adims, ast, abox, 1))) return rc;
    // weights: dims (Cin, Cout, taps)
    const int ntaps = p.kt * p.kh * p.kw;
    const uint64_t bdims[3] = {(uint64_t)p.Cin, (uint64_t)p.Cout, (uint64_t)ntaps};
    const uint64_t bst[2] = {(uint64_t)p.Cin * 2, (uint64_t)p.Cin * 2 * p.Cout};
    const uint32_t bbox[3] = {64, (uint32_t)p.BN, 1};
    if ((rc = make_tensor_map_bf16(&mB_hi, d->w_hi, 3, bdims, bst, bbox, 1))) return rc;
    if (split && (rc = make_tensor_map_bf16(&mB_lo, d->w_lo, 3, bdims, bst, bbox, 1))) return rc;
    if (store_bf16) {
        const uint64_t odims[5] = {(uint64_t)p.Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.T, (uint64_t)p.N};
        const uint64_t ost[4] = {(uint64_t)d->out_cstride * 2, (uint64_t)d->out_cstride * 2 * p.W,
                                 (uint64_t)d->out_cstride * 2 * p.W * p.H,
                                 (uint64_t)d->out_cstride * 2 * p.W * p.H * p.T};
        if ((rc = make_tensor_map_bf16(&mO_hi, d->y_hi + d->out_coff, 5, odims, ost, abox, 1))) return rc;
        if (split && (rc = make_tensor_map_bf16(&mO_lo, d->y_lo + d->out_coff, 5, odims, ost, abox, 1))) return rc;
    }

    const size_t smem_bytes = L.total + 1024;
    static size_t configured = 0;
    if (smem_bytes > configured) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024));
        configured = 227 * 1024;
    }
    int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    conv_igemm_kernel<<<grid, kConvThreads, smem_bytes, stream>>>(mA_hi, mA_lo, mB_hi, mB_lo, mO_hi, mO_lo, p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

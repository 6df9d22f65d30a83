// Conv3d_1a_7x7 WEIGHT GRADIENT on the raw uint8 clip with a RESIDENT INPUT HALO (sm_100a, tcgen05 / TMA).
//
// Same operator as otal_conv1a_wgrad_u8 (weight part of the backward of AFSD/common/i3d_backbone.py:196-199 in the folded layout,
// see include/opental_b200.h):   dw[(dt,dh)][co][j] += sum_{n,t',h',w'} D[n,t',h',w',co] * u[n, 2t'+dt-2, 2h'+dh-2, window(w')[j]]
// with j = the 32-element (8 pixels x 4 slots) window of output column w' — a different data flow.  The generic weight-gradient
// kernel loads, for every (dt, dh) tap of a 64-position K chunk, its own 4 KB X tile (and the 16 KB D tile once per 4 taps):
// 808 KB of L2 -> shared-memory fill per 128 positions, which is what bounds it (1.73 ms at batch 8 against a tensor-pipe floor
// of 0.75 ms).  Here, as in conv1a_halo.cu:
//   * a work unit is ONE tile of 16 x 8 output positions of one output frame = the K = 128 of the MMAs;
//   * for one dt the unit's input rows are loaded ONCE: a box of 38 input rows x 8 column windows x 64 bytes (19 KB).  With
//     K = positions the X operand is MN-major: an 8-row swizzle atom is ONE input row of the box (its 8 windows = 8 K indices, 64
//     bytes = 32 M indices each), the atom of tap dh for K-group hh lives at (2 hh + dh) x 512 bytes.  Consecutive dh taps are
//     consecutive M blocks (leading offset 512 B), consecutive hh are 1 KB apart (stride offset): the A operand of ONE MMA with
//     M = 128 is the window of taps dh = 0..3 (or 4..7; tap 7 does not exist and its rows are dropped) — the seven dh taps cost
//     no further fill and no per-tap MMA;
//   * the [D_hi | D_lo] tile (16 x 8 positions x 64 channels x 2 planes = 32 KB, MN-major SWIZZLE_128B, N = 128) arrives with
//     the same stage and serves the TWO dt of the CTA;
//   * a CTA owns a PAIR of dt (dt = 6 alone): four accumulators of 128 lanes x 128 columns = all 512 TMEM columns, accumulated
//     over kWhFlush units of the CTA's share and then flushed (coalesced red.add into dw).  The flush interval bounds the length
//     of the fp32 accumulation chain inside the tensor core: one chain over a whole share (3500 accumulate steps) measured
//     1.5e-4 relative on the final gradient (the accumulator's rounding is not unbiased), 512 steps stay at the 2e-5 of the
//     generic kernel.  Fill per 128 positions: 7 x 19 + 4 x 32 = 261 KB.
// Warp roles and barriers follow conv_wgrad.cu.
#include "common.cuh"
#include "tensormap.h"

namespace otal {

constexpr int kWhThreads = 256;
constexpr int kWhRows = 38;                          // 2 * 16 + 6 input rows: tap dh = 7 of the second M block reads row 37
constexpr int kWhXBytes = kWhRows * 512;             // [38 rows][8 windows][64 B]
constexpr int kWhXSlot = 20 * 1024;                  // rounded up: 1 KB aligned slots
constexpr int kWhDBytes = 16384;                     // [16 x 8 positions][64 channels] of one plane
constexpr int kWhStage = 2 * kWhXSlot + 2 * kWhDBytes;      // 72 KB
constexpr int kWhStages = 3;
constexpr int kWhFlush = 64;                         // units (x 8 K steps) per accumulation chain
constexpr int kWhBarOff = kWhStages * kWhStage;
constexpr int kWhSmem = kWhBarOff + 256;

struct WhParams {
    int N, To, Ho, Wo;
    int hblocks, wblocks, total_units;
    int ctas_pair, ctas_single;        // CTAs per dt pair (three pairs) and for dt = 6
    float* dw;                         // [49][64][32]
};

struct alignas(64) WhMaps { CUtensorMap X, D_hi, D_lo; };

__global__ void __launch_bounds__(kWhThreads, 1)
conv1a_wgrad_halo_kernel(const __grid_constant__ WhMaps maps, const WhParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWhBarOff);
    uint64_t* empty_bar = full_bar + kWhStages;
    uint64_t* tmem_full = empty_bar + kWhStages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    // CTA -> (dt group, share of the units).  Groups 0..2 run two dt per unit, group 3 one: it gets half as many CTAs.
    int group, share, nshare;
    {
        const int b = blockIdx.x;
        if (b < 3 * p.ctas_pair) { group = b / p.ctas_pair; share = b % p.ctas_pair; nshare = p.ctas_pair; }
        else { group = 3; share = b - 3 * p.ctas_pair; nshare = p.ctas_single; }
    }
    const int ndt = group == 3 ? 1 : 2;
    const int dt0 = 2 * group;
    // units are dealt round-robin: at any time the CTAs of ALL groups work inside one moving window of the clip (a pair CTA
    // takes twice as long per unit but owns half as many), so a D tile / input frame fetched from HBM by one group is an L2 hit
    // for the others (contiguous shares re-read D from HBM for the odd group: 1.51 GB of DRAM reads against 0.77 GB of operands)
    const int u_begin = share, u_step = nshare;
    const int u_end = p.total_units;
    const int n_units = share < p.total_units ? (p.total_units - share + nshare - 1) / nshare : 0;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&maps.X); tma_prefetch_desc(&maps.D_hi); tma_prefetch_desc(&maps.D_lo); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kWhStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tmem_full, 1); mbar_init(tmem_empty, 128);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // unit -> (n, t', h block, w block); w block fastest so that neighbouring units share input rows in L2
    auto decode = [&](int unit, int& n, int& to, int& h0, int& w0) {
        const int wb = unit % p.wblocks; unit /= p.wblocks;
        const int hb = unit % p.hblocks; unit /= p.hblocks;
        to = unit % p.To; n = unit / p.To;
        h0 = hb * 16; w0 = wb * 8;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0; uint32_t phase = 0;
        const uint32_t tx = (uint32_t)ndt * kWhXBytes + 2u * kWhDBytes;
        for (int unit = u_begin; unit < u_end; unit += u_step) {
            int n, to, h0, w0;
            decode(unit, n, to, h0, w0);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
                unsigned char* s = smem + (size_t)stage * kWhStage;
                mbar_expect_tx(&full_bar[stage], tx);
                // input rows 2 h0 - 2 .. + 37 of input frame 2 t' + dt - 2; rows / frames outside the clip are zero-filled
                for (int i = 0; i < ndt; ++i)
                    tma_load_5d(&maps.X, &full_bar[stage], s + i * kWhXSlot, 0, w0, 2 * h0 - 2, 2 * to + dt0 + i - 2, n);
                tma_load_5d(&maps.D_hi, &full_bar[stage], s + 2 * kWhXSlot, 0, w0, h0, to, n);
                tma_load_5d(&maps.D_lo, &full_bar[stage], s + 2 * kWhXSlot + kWhDBytes, 0, w0, h0, to, n);
            }
            __syncwarp();
            if (++stage == kWhStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_bf16(128, 128, 1, 1);                 // both operands MN-major: X^T-like views x [D_hi | D_lo]
        // X: MN-major SWIZZLE_64B; M blocks (32 window elements of tap dh) 512 B apart, 8-position K groups (hh) 1 KB apart
        const uint64_t tmpl_x = umma_smem_desc(0, 512, 1024, 4);
        // D: MN-major SWIZZLE_128B; the lo plane is the second 64-wide N block (16 KB further), K groups (hh) 1 KB apart
        const uint64_t tmpl_d = umma_smem_desc_sw128(0, kWhDBytes, 1024);
        int stage = 0; uint32_t phase = 0;
        int in_chain = 0; uint32_t flush_phase = 0;
        for (int unit = u_begin; unit < u_end; unit += u_step) {
            if (in_chain == 0 && unit != u_begin) {            // the epilogue must have read the previous chain out of TMEM
                mbar_wait(tmem_empty, flush_phase ^ 1);
                tc_fence_after();
            }
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t s = smem_u32(smem + (size_t)stage * kWhStage);
            const uint32_t sD = s + 2 * kWhXSlot;
            if (elect_one()) {
                for (int i = 0; i < ndt; ++i) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(i * 2 + half) * 128;
                        const uint32_t sX = s + i * kWhXSlot + half * 4 * 512;          // taps dh = 4 half .. 4 half + 3
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {                                 // 16 positions = 2 hh = 2 KB of X rows, 2 KB of D
                            const uint64_t a = tmpl_x + (uint64_t)((sX + ks * 2048) >> 4);
                            const uint64_t b = tmpl_d + (uint64_t)((sD + ks * 2048) >> 4);
                            umma_f16(d_tmem, a, b, idesc, !(in_chain == 0 && ks == 0));
                        }
                    }
                }
                umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == kWhStages) { stage = 0; phase ^= 1; }
            if (++in_chain == kWhFlush || unit + u_step >= u_end) {
                if (elect_one()) umma_commit(tmem_full);
                __syncwarp();
                in_chain = 0; flush_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue: one flush per accumulation chain
        const int q = warp & 3;                         // tap inside the M block: lanes = its 32 window elements
        const int nflush = (n_units + kWhFlush - 1) / kWhFlush;
        uint32_t fphase = 0;
        for (int f = 0; f < nflush; ++f) {
            mbar_wait(tmem_full, fphase);
            tc_fence_after();
            for (int i = 0; i < ndt; ++i)
                for (int half = 0; half < 2; ++half) {
                    const int dh = half * 4 + q;
                    if (dh >= 7) continue;                  // tap 7 does not exist (warp-uniform)
                    const uint32_t t_acc = tmem_base + (uint32_t)(i * 2 + half) * 128 + ((uint32_t)(q * 32) << 16);
                    float* dst = p.dw + (size_t)((dt0 + i) * 7 + dh) * 64 * 32 + lane;
#pragma unroll
                    for (int col0 = 0; col0 < 64; col0 += 32) {
                        uint32_t v[32], v2[32];
                        tmem_ld32(t_acc + col0, v);
                        tmem_ld32(t_acc + 64 + col0, v2);    // the u * D_lo products
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            atomicAdd(dst + (size_t)(col0 + j) * 32, __uint_as_float(v[j]) + __uint_as_float(v2[j]));
                    }
                }
            tc_fence_before();
            mbar_arrive(tmem_empty);
            fphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace otal

using namespace otal;

extern "C" {

// See include/opental_b200.h: same contract as otal_conv1a_wgrad_u8 (tT / tH / tW of the descriptor are ignored).
int otal_conv1a_wgrad_u8_halo(const otal_conv1a_wgrad_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("conv1a_wgrad_halo: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (d->N <= 0 || d->T < 6 || d->H < 6 || d->W < 6 || d->T % 2 || d->H % 2 || d->W % 2) {
        set_last_error_msg("conv1a_wgrad_halo: needs even extents >= 6"); return OTAL_ERR_BAD_ARG;
    }
    if (d->Cout != 64 || d->nsplit != 3 || d->d_cstride % 8 || d->d_coff % 8) {
        set_last_error_msg("conv1a_wgrad_halo: Cout must be 64 (bf16x3), gradient slice 16-byte aligned"); return OTAL_ERR_UNSUPPORTED;
    }
    if (!d->x_hi || !d->d_hi || !d->d_lo || !d->dw) { set_last_error_msg("conv1a_wgrad_halo: null pointer"); return OTAL_ERR_BAD_ARG; }
    WhParams p{};
    p.N = d->N; p.To = d->T / 2; p.Ho = d->H / 2; p.Wo = d->W / 2;
    p.hblocks = (p.Ho + 15) / 16; p.wblocks = (p.Wo + 7) / 8;
    p.total_units = p.N * p.To * p.hblocks * p.wblocks;
    p.dw = d->dw;
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    // 3 S + S / 2 CTAs: a pair CTA does twice the work of a single-dt CTA per unit
    int S = (2 * sms) / 7;
    if (S > p.total_units) S = p.total_units;
    if (S < 1) S = 1;
    p.ctas_pair = S;
    p.ctas_single = sms - 3 * S > 0 ? sms - 3 * S : 1;
    if (p.ctas_single > p.total_units) p.ctas_single = p.total_units;

    WhMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    {
        // the clip [N,T,H,W+8,4] through the overlapping-window view of conv1a_halo.cu, 8 windows per box row
        const uint64_t px = 4 * 2, Wp = (uint64_t)d->W + 8;
        const uint64_t sH = px * Wp, sT = sH * d->H, sN = sT * d->T;
        const uint64_t adims[5] = {32, (uint64_t)p.Wo, (uint64_t)d->H, (uint64_t)d->T, (uint64_t)d->N};
        const uint64_t ast[4] = {2 * px, sH, sT, sN};
        const uint32_t abox[5] = {32, 8, (uint32_t)kWhRows, 1, 1};
        if ((rc = make_tensor_map_bf16(&maps.X, d->x_hi, 5, adims, ast, abox, 2))) return rc;
    }
    {
        const uint32_t dbox[5] = {64, 8, 16, 1, 1};
        const uint64_t ddims[5] = {64, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.To, (uint64_t)p.N};
        const uint64_t cs = (uint64_t)d->d_cstride * 2;
        const uint64_t dst_[4] = {cs, cs * p.Wo, cs * p.Wo * p.Ho, cs * p.Wo * p.Ho * p.To};
        if ((rc = make_tensor_map_bf16(&maps.D_hi, d->d_hi + d->d_coff, 5, ddims, dst_, dbox, 1))) return rc;
        if ((rc = make_tensor_map_bf16(&maps.D_lo, d->d_lo + d->d_coff, 5, ddims, dst_, dbox, 1))) return rc;
    }
    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(conv1a_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        once.mark(once_dev);
    }
    conv1a_wgrad_halo_kernel<<<3 * p.ctas_pair + p.ctas_single, kWhThreads, kWhSmem + 1024, stream>>>(maps, p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

// Conv3d_1a_7x7 forward on the RAW uint8 clip with a RESIDENT INPUT HALO (sm_100a, tcgen05 / TMA).
//
// Same operator as otal_conv1a_fwd_u8 (AFSD/common/i3d_backbone.py:196-199 + the loader's normalisation folded into the
// epilogue, see include/opental_b200.h) — a different data flow.  The generic implicit-GEMM kernel loads, for every one of the
// 49 (dt, dh) taps of a 128-position tile, an 8 KB activation box and an 8 KB weight tile: 784 KB of L2 -> shared-memory fill
// per tile, which is what bounds it (tensor pipe 23 % active in ncu, round 2).  Here
//   * a work unit is 256 output positions of one output frame: two tiles of 16 rows x 8 columns side by side in W;
//   * for one dt the unit's input rows are loaded ONCE: a box of 37 input rows x 16 column windows x 64 bytes (the 8-pixel x
//     4-slot window of an output column, expanded by the overlapping-window tensor map) = 37 KB.  The A operand of tap
//     (dt, dh) is a VIEW of that box — row m = 8 hh + ww of a tile lives at (2 hh + dh) KB + tile x 512 + 64 ww bytes, i.e. the
//     8-row swizzle atoms of consecutive hh are a constant 2 KB apart — so the seven dh taps cost no further activation fill;
//   * the seven [w_hi | w_lo] weight tiles of the dt (56 KB) arrive with the same stage and serve both tiles.
// Fill per 256 positions: 7 x (37 + 56) KB = 651 KB instead of 2 x 784 KB; the MMAs are unchanged (one N-concatenated
// 128 x 128 x 16 per K step and tile, the single-plane u8 form).  Warp roles, barriers and the epilogue follow conv_igemm.cu.
#include "common.cuh"
#include "tensormap.h"

namespace otal {

constexpr int kHaloThreads = 256;
constexpr int kHaloRows = 37;                        // 2 * 16 + 5 input rows of a 16-row output tile
constexpr int kHaloABytes = kHaloRows * 1024;        // [37 rows][16 windows][64 B]
constexpr int kHaloAStage = 38 * 1024;               // rounded up: the weight tiles start 1 KB aligned
constexpr int kHaloBBytes = 7 * 8192;                // 7 dh taps x [128 rows (hi | lo)][64 B]
constexpr int kHaloStage = kHaloAStage + kHaloBBytes;
constexpr int kHaloStages = 2;
constexpr int kHaloStaging = 2 * 16384;              // one 64-channel chunk of one tile: hi + lo planes
constexpr int kHaloBarOff = kHaloStages * kHaloStage + kHaloStaging;
constexpr int kHaloSmem = kHaloBarOff + 256 + 512;   // barriers, then 64 x (scale) floats + padding

struct HaloParams {
    int N, To, Ho, Wo, T, H, W;        // output extent and input extent (W = image width, before the 8-pixel padding)
    int Cout;                          // 64
    int hblocks, wpairs;               // ceil(Ho / 16), ceil(Wo / 16)
    int total_units;
    int relu;
    const float* scale;                // [Cout]
    const float* shift;                // [4][4][4][Cout] border-class table
};

struct alignas(64) HaloMaps { CUtensorMap A, B, O_hi, O_lo; };

__global__ void __launch_bounds__(kHaloThreads, 1)
conv1a_halo_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kHaloBarOff);
    uint64_t* empty_bar = full_bar + kHaloStages;
    uint64_t* tmem_full = empty_bar + kHaloStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* sscale = reinterpret_cast<float*>(smem + kHaloBarOff + 256);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.A); tma_prefetch_desc(&maps.B); tma_prefetch_desc(&maps.O_hi); tma_prefetch_desc(&maps.O_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kHaloStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + 64) sscale[threadIdx.x - 128] = p.scale ? __ldg(p.scale + threadIdx.x - 128) : 1.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // unit -> (n, t', h block, w pair); w pair fastest so that neighbouring CTAs share input rows in L2
    auto decode = [&](int unit, int& n, int& to, int& h0, int& w0) {
        const int wp = unit % p.wpairs; unit /= p.wpairs;
        const int hb = unit % p.hblocks; unit /= p.hblocks;
        to = unit % p.To; n = unit / p.To;
        h0 = hb * 16; w0 = wp * 16;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0; uint32_t phase = 0;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            int n, to, h0, w0;
            decode(unit, n, to, h0, w0);
            for (int dt = 0; dt < 7; ++dt) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    unsigned char* sA = smem + (size_t)stage * kHaloStage;
                    mbar_expect_tx(&full_bar[stage], kHaloABytes + kHaloBBytes);
                    // input rows 2 h0 - 2 .. + 36 of input frame 2 t' + dt - 2 ("same" padding of k = 7, s = 2 on even extents:
                    // front 2); rows / frames outside the clip are zero-filled by the TMA unit = the raw clip's zero padding
                    tma_load_5d(&maps.A, &full_bar[stage], sA, 0, w0, 2 * h0 - 2, 2 * to + dt - 2, n);
                    tma_load_3d(&maps.B, &full_bar[stage], sA + kHaloAStage, 0, 0, dt * 7);
                }
                __syncwarp();
                if (++stage == kHaloStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);              // u * [w_hi | w_lo]
        const uint64_t tmpl_a = umma_smem_desc(0, 16, 2048, 4);              // SWIZZLE_64B, 8-row atoms 2 KB apart (two input rows)
        const uint64_t tmpl_b = umma_smem_desc(0, 16, 512, 4);               // weight rows: atoms back to back
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256;
            for (int dt = 0; dt < 7; ++dt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + (size_t)stage * kHaloStage);
                const uint32_t sB = sA + kHaloAStage;
                if (elect_one()) {
#pragma unroll
                    for (int dh = 0; dh < 7; ++dh) {
                        const uint64_t b0 = tmpl_b + (uint64_t)((sB + dh * 8192) >> 4);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t a0 = tmpl_a + (uint64_t)((sA + dh * 1024 + j * 512) >> 4);
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(d_tmem + (uint32_t)j * 128, a0 + (uint64_t)(k * 2), b0 + (uint64_t)(k * 2), idesc,
                                         (dt | dh | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == kHaloStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tmem_full[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (128 threads)
        const int q = warp & 3;
        const int row = q * 32 + lane;                  // position inside a tile: hh = row / 8, ww = row % 8
        const int et = threadIdx.x - 128;
        unsigned char* buf_hi = smem + kHaloStages * kHaloStage;
        unsigned char* buf_lo = buf_hi + 16384;
        const int Cout = p.Cout;
        const float relu_floor = p.relu ? 0.f : -3.402823466e38f;
        auto cls = [](int o, int n_) { return o == 0 ? 1 : (o == n_ - 2 ? 2 : (o == n_ - 1 ? 3 : 0)); };
        int acc = 0; uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            int n, to, h0, w0;
            decode(unit, n, to, h0, w0);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            for (int j = 0; j < 2; ++j) {
                const int ho = h0 + (row >> 3), wo = w0 + j * 8 + (row & 7);
                // border class of the output position (ConvParams::shift_classes in conv_igemm.cu); positions beyond the
                // extent are clipped by the TMA store, their class only has to stay inside the table
                const int ct = cls(to, p.To), chh = ho < p.Ho ? cls(ho, p.Ho) : 0, cw = wo < p.Wo ? cls(wo, p.Wo) : 0;
                const float4* tb = reinterpret_cast<const float4*>(p.shift + ((ct * 4 + chh) * 4 + cw) * Cout);
                const uint32_t t_acc = tmem_base + (uint32_t)acc * 256 + (uint32_t)j * 128 + ((uint32_t)(q * 32) << 16);
                // the previous store must have read the staging buffer before it is overwritten
                if (et == 0) tma_store_wait_read<0>();
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int col0 = half * 32;
                    uint32_t v[32], v2[32];
                    tmem_ld32(t_acc + col0, v);
                    tmem_ld32(t_acc + 64 + col0, v2);          // the u * w_lo products
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 s4 = *reinterpret_cast<const float4*>(sscale + col0 + c);
                        const float4 t4 = __ldg(tb + ((col0 + c) >> 2));
                        f[c] = fmaxf(fmaf(__uint_as_float(v[c]) + __uint_as_float(v2[c]), s4.x, t4.x), relu_floor);
                        f[c + 1] = fmaxf(fmaf(__uint_as_float(v[c + 1]) + __uint_as_float(v2[c + 1]), s4.y, t4.y), relu_floor);
                        f[c + 2] = fmaxf(fmaf(__uint_as_float(v[c + 2]) + __uint_as_float(v2[c + 2]), s4.z, t4.z), relu_floor);
                        f[c + 3] = fmaxf(fmaf(__uint_as_float(v[c + 3]) + __uint_as_float(v2[c + 3]), s4.w, t4.w), relu_floor);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) split_bf16x2(f[g * 8 + e * 2], f[g * 8 + e * 2 + 1], hi[e], lo[e]);
                        const int j16 = half * 4 + g;
                        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j16 ^ (row & 7)) << 4);
                        *reinterpret_cast<uint4*>(buf_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(buf_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) {
                    tma_store_5d(&maps.O_hi, buf_hi, 0, w0 + j * 8, h0, to, n);
                    tma_store_5d(&maps.O_lo, buf_lo, 0, w0 + j * 8, h0, to, n);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (et == 0) tma_store_wait_all<0>();
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

// See include/opental_b200.h.  d->w_hi = the PACKED weights [49][2 * Cout][32] (per tap: the Cout w_hi rows, then the Cout w_lo
// rows); d->w_lo and d->x_lo are ignored.
int otal_conv1a_fwd_u8_halo(const otal_conv1a_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d) { set_last_error_msg("conv1a_halo: null descriptor"); return OTAL_ERR_BAD_ARG; }
    if (d->N <= 0 || d->T < 6 || d->H < 6 || d->W < 6 || d->T % 2 || d->H % 2 || d->W % 2) {
        set_last_error_msg("conv1a_halo: needs even extents >= 6"); return OTAL_ERR_BAD_ARG;
    }
    if (d->Cout != 64 || d->out_cstride % 8 || d->out_coff % 8 || d->nsplit != 3) {
        set_last_error_msg("conv1a_halo: Cout must be 64 (bf16x3), output slice 16-byte aligned"); return OTAL_ERR_UNSUPPORTED;
    }
    if (!d->x_hi || !d->w_hi || !d->y_hi || !d->y_lo || !d->shift) { set_last_error_msg("conv1a_halo: null plane"); return OTAL_ERR_BAD_ARG; }
    HaloParams p{};
    p.N = d->N; p.T = d->T; p.H = d->H; p.W = d->W;
    p.To = d->T / 2; p.Ho = d->H / 2; p.Wo = d->W / 2;
    p.Cout = d->Cout; p.relu = d->relu; p.scale = d->scale; p.shift = d->shift;
    p.hblocks = (p.Ho + 15) / 16; p.wpairs = (p.Wo + 15) / 16;
    p.total_units = p.N * p.To * p.hblocks * p.wpairs;

    HaloMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    {
        // the clip [N,T,H,W+8,4]: dim0 = the 32-element window of an output column, dim1 = output column (window origin advances
        // 2 pixels = 16 bytes: overlapping windows, nothing is duplicated in HBM), dim2 = input row, dim3 = frame, dim4 = sample
        const uint64_t px = 4 * 2, Wp = (uint64_t)d->W + 8;
        const uint64_t sH = px * Wp, sT = sH * d->H, sN = sT * d->T;
        const uint64_t adims[5] = {32, (uint64_t)p.Wo, (uint64_t)d->H, (uint64_t)d->T, (uint64_t)d->N};
        const uint64_t ast[4] = {2 * px, sH, sT, sN};
        const uint32_t abox[5] = {32, 16, (uint32_t)kHaloRows, 1, 1};
        if ((rc = make_tensor_map_bf16(&maps.A, d->x_hi, 5, adims, ast, abox, 2))) return rc;
    }
    {
        const uint64_t bdims[3] = {32, 128, 49};
        const uint64_t bst[2] = {64, 64 * 128};
        const uint32_t bbox[3] = {32, 128, 7};
        if ((rc = make_tensor_map_bf16(&maps.B, d->w_hi, 3, bdims, bst, bbox, 2))) return rc;
    }
    {
        const uint32_t obox[5] = {64, 8, 16, 1, 1};
        const uint64_t odims[5] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.To, (uint64_t)p.N};
        const uint64_t cs = (uint64_t)d->out_cstride * 2;
        const uint64_t ost[4] = {cs, cs * p.Wo, cs * p.Wo * p.Ho, cs * p.Wo * p.Ho * p.To};
        if ((rc = make_tensor_map_bf16(&maps.O_hi, d->y_hi + d->out_coff, 5, odims, ost, obox, 1))) return rc;
        if ((rc = make_tensor_map_bf16(&maps.O_lo, d->y_lo + d->out_coff, 5, odims, ost, obox, 1))) return rc;
    }
    static OncePerDevice once;
    int once_dev = 0;
    if (once.need(&once_dev)) {
        OTAL_CUDA_TRY(cudaFuncSetAttribute(conv1a_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        once.mark(once_dev);
    }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const int grid = p.total_units < sms ? p.total_units : sms;
    conv1a_halo_kernel<<<grid, kHaloThreads, kHaloSmem + 1024, stream>>>(maps, p);
    OTAL_CUDA_TRY(cudaGetLastError());
    return OTAL_OK;
}

}  // extern "C"

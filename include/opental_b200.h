/*
 * opental_b200 — C ABI of the B200 (sm_100a) OpenTAL hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 *   - takes DEVICE pointers (unless stated otherwise) and a `void* stream` that is a cudaStream_t,
 *   - launches asynchronously on that stream, never allocates device memory, never synchronises,
 *   - returns 0 on success or a negative OTAL_ERR_* code; otal_last_error() then describes the failure
 *     (thread-local, valid until the next failing call on the same thread),
 *   - is re-entrant (no global mutable state besides one-time attribute caches).
 *
 * The reference interface each function replaces is cited as file:line relative to the OpenTAL repository.
 */
#ifndef OPENTAL_B200_H
#define OPENTAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTAL_ABI_VERSION 1

#if defined(__GNUC__)
#define OTAL_API __attribute__((visibility("default")))
#else
#define OTAL_API
#endif

#define OTAL_OK 0
#define OTAL_ERR_BAD_ARG (-1)     /* invalid shape / null pointer / unsupported combination of arguments */
#define OTAL_ERR_CUDA (-2)        /* a CUDA runtime call or kernel launch failed */
#define OTAL_ERR_DRIVER (-3)      /* TMA tensor-map encoding failed */
#define OTAL_ERR_UNSUPPORTED (-4) /* configuration outside what the kernels implement */

OTAL_API const char* otal_last_error(void);
OTAL_API int otal_abi_version(void);
/* sizeof of a descriptor struct of this header by name ("otal_conv_desc", ...), 0 if unknown: a binding compares it with
 * the size of its own mirror (opental_b200/_lib.py does at load time) so that a stale mirror fails loudly, not silently. */
OTAL_API int otal_abi_sizeof(const char* name);

/* ------------------------------------------------------------------------------------------------------------
 * BoundaryMaxPooling — replaces boundary_max_pooling_cuda.forward / .backward
 *   AFSD/prop_pooling/boundary_max_pooling_cuda.cpp:21-34 (forward), :36-50 (backward), pybind :52-55
 *   AFSD/prop_pooling/boundary_max_pooling_kernel.cu:18-46, :49-82 (kernels), :84-145 (launchers)
 *
 * in  [B,C,T] contiguous, seg [B,K,4] contiguous *same dtype as in* (float->int truncation, then clamp to
 * [0,T-1]), out [B,C,K].  Channels c < C/2 use seg[...,0:2], the others seg[...,2:4]; C must be even.
 * Unlike the reference, `out` / `grad_in` need not be zero-initialised: every element is written.
 *
 * backward: compat_tscale_bug != 0 reproduces the reference's use of K (= grad_output.size(2)) instead of T as
 * the time extent for clamping and addressing (boundary_max_pooling_kernel.cu:121); exact when T == K.
 * compat_tscale_bug == 0 is the mathematically correct gradient.  The backward is deterministic.
 * ---------------------------------------------------------------------------------------------------------- */
OTAL_API int otal_bmp_forward_f32(const float* in, const float* seg, float* out, int B, int C, int T, int K, void* stream);
OTAL_API int otal_bmp_backward_f32(const float* grad_out, const float* in, const float* seg, float* grad_in, int B, int C,
                          int T, int K, int compat_tscale_bug, void* stream);
OTAL_API int otal_bmp_forward_f64(const double* in, const double* seg, double* out, int B, int C, int T, int K,
                         void* stream);
OTAL_API int otal_bmp_backward_f64(const double* grad_out, const double* in, const double* seg, double* grad_in, int B,
                          int C, int T, int K, int compat_tscale_bug, void* stream);
/* half precision (AT_DISPATCH_FLOATING_TYPES_AND_HALF, boundary_max_pooling_kernel.cu:96,128): raw IEEE binary16 bit patterns,
 * segments half as well (truncated like the reference's static_cast<int>).  The backward sums a frame's contributions in fp32
 * and rounds once (the reference's half atomicAdd rounds after every addition, in a run-dependent order). */
OTAL_API int otal_bmp_forward_f16(const uint16_t* in, const uint16_t* seg, uint16_t* out, int B, int C, int T, int K,
                         void* stream);
OTAL_API int otal_bmp_backward_f16(const uint16_t* grad_out, const uint16_t* in, const uint16_t* seg, uint16_t* grad_in,
                          int B, int C, int T, int K, int compat_tscale_bug, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution (tcgen05 / TMA), stride 1 — replaces the per-layer chain
 *   F.pad -> nn.Conv3d(bias=False) -> BatchNorm3d(eval, frozen) -> ReLU   AFSD/common/i3d_backbone.py:51-87
 *   F.pad -> nn.Conv1d(bias)                                               AFSD/common/layers.py:204-214
 *   F.pad -> nn.Conv3d(bias) with full-extent spatial kernel               AFSD/common/layers.py:143-175
 *
 * Activations are NDHWC, stored as two bf16 planes (hi, lo), x ~= hi + lo (nsplit == 3, "bf16x3") or one
 * plane (nsplit == 1).  A tensor may be a channel slice [coff, coff+C) of rows that are `cstride` channels
 * wide, which is how inception branches write straight into the concat buffer.
 * Weights: [kt*kh*kw][Cout][Cin] bf16 planes (tap-major, K-major rows).
 * y = relu?( conv(x, w) * scale[co] + shift[co] ), written as bf16 planes (y_hi/y_lo) and/or fp32 (y_f32).
 * Output extent = ceil(input extent / stride) (pad front = pt/ph/pw, the rest of the "same" padding is implicit
 * zero: TMA out-of-bounds fill).  With accumulate = 1 the fp32 destination is read-modify-written, which is how the
 * data-gradient of several consumers of one tensor is summed (dgrad = this same entry point run on the
 * output-gradient with tap-flipped, channel-transposed weights).
 * (tT,tH,tW) is the 128-position tile box, tT*tH*tW == 128.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct otal_conv_desc {
    int N, T, H, W;          /* batch and INPUT spatial extent */
    int Cin, Cout;           /* channels read / written by this launch */
    int kt, kh, kw;          /* filter taps */
    int pt, ph, pw;          /* front padding */
    int tT, tH, tW;          /* tile box */
    int sT, sH, sW;          /* conv stride per dim: 1 or 2 (0 = 1).  Output extent = ceil(input extent / stride) */
    int nsplit;              /* 1 = bf16, 3 = bf16x3 */
    int relu;
    int accumulate;          /* y_f32 += result instead of y_f32 = result (dgrad into a shared gradient buffer) */
    int dgrad;               /* data-gradient mode: x = output gradient planes [.., Cin = forward Cout], w = the
                                FORWARD weights [taps][Cin][Cout] read transposed with flipped taps, pt/ph/pw =
                                k-1-forward pad, Cout = forward Cin.  Stride 1 only. */
    int y_f32_ncdhw;         /* fp32 destination layout: 0 = NDHWC (out_cstride channels per position), 1 = NCDHW
                                ([N, out_cstride, To, Ho, Wo], any Cout) — the reference's layout, used by the 1-D head */
    int in_cstride, in_coff;   /* input row width and slice offset, in channels */
    int out_cstride, out_coff; /* output row width and slice offset, in channels (bf16 planes and fp32 alike) */
    const uint16_t* x_hi; const uint16_t* x_lo;   /* bf16 bit patterns */
    const uint16_t* w_hi; const uint16_t* w_lo;
    const float* scale; const float* shift;       /* [Cout] or NULL */
    uint16_t* y_hi; uint16_t* y_lo;               /* NULL = do not store bf16 planes */
    float* y_f32;                                 /* NULL = do not store fp32 */
    /* Optional second K segment, 1x1 stride-1 convs only (Cin2 = 0: unused): y = [x | x2] . [w ; w2].  Used to compute the
     * data gradient of several 1x1 convs that share their input (the b0 / b1a / b2a branches of an inception module,
     * AFSD/common/i3d_backbone.py:116-121) in one pass instead of one read-modify-write pass per branch.
     * x2: planes [N,T,H,W,in2_cstride], channels [in2_coff, in2_coff + Cin2); w2: [1][Cout][Cin2] (dgrad: [1][Cin2][Cout]). */
    int Cin2, in2_cstride, in2_coff;
    const uint16_t* x2_hi; const uint16_t* x2_lo;
    const uint16_t* w2_hi; const uint16_t* w2_lo;
    /* must be 0 or 1 (a split-K experiment of round 1, withdrawn: the field stays so that the struct layout is unchanged) */
    int ksplit;
} otal_conv_desc;

OTAL_API int otal_conv_igemm_fwd(const otal_conv_desc* desc, void* stream);

/* Conv3d_1a_7x7: 7x7x7, stride 2, 3 input channels -> Cout, + folded BN + ReLU
 *   AFSD/common/i3d_backbone.py:196-199 (end point), :51-87 (Unit3D.forward incl. the (2,3) "same" padding)
 * x: the clip as written by otal_clip_ingest, [N,T,H,W+8,4] bf16 planes (image column w at padded column w + 2, zero
 * elsewhere, channel slot 3 zero): output column w' reads the 8-pixel x 4-slot window starting at padded column 2*w'.
 * w: [49 (dt,dh)][Cout][32] planes, element dw*4 + c of a row = W[co,c,dt,dh,dw] (zero for dw == 7 or c == 3).
 * y: [N,ceil(T/2),ceil(H/2),W/2,out_cstride] planes at channel offset out_coff. */
typedef struct otal_conv1a_desc {
    int N, T, H, W;
    int Cout;
    int tT, tH, tW;
    int nsplit, relu;
    int out_cstride, out_coff;
    const uint16_t* x_hi; const uint16_t* x_lo;
    const uint16_t* w_hi; const uint16_t* w_lo;
    const float* scale; const float* shift;
    uint16_t* y_hi; uint16_t* y_lo;
} otal_conv1a_desc;
OTAL_API int otal_conv1a_fwd(const otal_conv1a_desc* desc, void* stream);

/* Conv3d_1a on the RAW uint8 clip — same reference lines as otal_conv1a_fwd plus the loader's normalisation
 * (AFSD/common/thumos_dataset.py:261-263), which is folded into the epilogue instead of being applied to the input:
 * x_hi = the plane of otal_clip_ingest_u8_raw (pixel values 0..255, exact in bf16; x_lo is ignored), w as above.
 * With x = (2/255) u - 1 inside the image and zero padding of x outside it (i3d_backbone.py:59-79),
 *   conv(x, W)[p, co] = (2/255) * conv(u zero-padded, W)[p, co] - sum over the taps that lie inside the image at p of W[co,.],
 * so `scale` = bn_scale * 2/255 and `shift` is a TABLE [4][4][4][Cout] indexed by the border class of the output position
 * along (T, H, W) — class of index o among n outputs: 1 if o == 0, 2 if o == n-2, 3 if o == n-1, else 0 — holding
 * bn_shift - bn_scale * (sum of the in-bounds weights of that class).  Needs nsplit = 3, even T and H, extents >= 6.
 * One tensor-core pass u * [w_hi | w_lo] instead of two, half the activation traffic, no activation rounding error.
 * Default for uint8 input since round 2 (OTAL_U8_CONV1A=0 in opental_b200/backbone.py switches back). */
OTAL_API int otal_conv1a_fwd_u8(const otal_conv1a_desc* desc, void* stream);

/* The same operator with a RESIDENT INPUT HALO (csrc/conv1a_halo.cu): a work unit = 256 positions of one output frame (two
 * 16 x 8 tiles side by side); per dt ONE 37-row input box is loaded and the seven dh taps read it as shifted views, the seven
 * [w_hi | w_lo] weight tiles of the dt arrive with it: 651 KB of L2 -> shared-memory fill per 256 positions instead of 1568 KB.
 * Same arguments except: w_hi = the PACKED weights [49][2*Cout][32] (per tap the Cout w_hi rows, then the Cout w_lo rows),
 * w_lo and x_lo ignored; Cout = 64, nsplit = 3, even T / H / W >= 6; tT/tH/tW are ignored (the tile is 1 x 16 x 8). */
OTAL_API int otal_conv1a_fwd_u8_halo(const otal_conv1a_desc* desc, void* stream);

/* Weight gradient — replaces the weight part of torch's convolution_backward for Unit3D / Unit1D
 *   AFSD/common/i3d_backbone.py:82 (conv3d), AFSD/common/layers.py:211 (conv1d)
 * dw[tap][co][ci] += sum_{n,p} d[n,p,co] * x[n, s*p + tap - pad, ci]; x = saved conv input planes [N,T,H,W,x_cstride],
 * d = output-gradient planes [N,ceil(T/s),..,d_cstride] (see otal_relu_bn_bwd_split).  dw is fp32, same layout as
 * the forward weights, and is accumulated with float reductions: zero it (or keep the running gradient in it).
 * (tT,tH,tW) is the 64-position K tile box. */
typedef struct otal_wgrad_desc {
    int N, T, H, W;          /* batch and INPUT (x) extent */
    int Cin, Cout;
    int kt, kh, kw;
    int pt, ph, pw;
    int sT, sH, sW;
    int tT, tH, tW;
    int nsplit;
    int x_cstride, x_coff, d_cstride, d_coff;
    const uint16_t* x_hi; const uint16_t* x_lo;
    const uint16_t* d_hi; const uint16_t* d_lo;
    float* dw;
} otal_wgrad_desc;
OTAL_API int otal_conv_wgrad(const otal_wgrad_desc* desc, void* stream);

/* Weight gradient of Conv3d_1a_7x7 in the folded layout of otal_conv1a_fwd: dw is [49][Cout][32] fp32. */
typedef struct otal_conv1a_wgrad_desc {
    int N, T, H, W;
    int Cout;
    int tT, tH, tW;
    int nsplit;
    int d_cstride, d_coff;
    const uint16_t* x_hi; const uint16_t* x_lo;
    const uint16_t* d_hi; const uint16_t* d_lo;
    float* dw;
} otal_conv1a_wgrad_desc;
OTAL_API int otal_conv1a_wgrad(const otal_conv1a_wgrad_desc* desc, void* stream);

/* Weight gradient of Conv3d_1a against the raw uint8 clip (x_hi = plane of otal_clip_ingest_u8_raw, x_lo ignored, nsplit = 3,
 * Cout = 64): dw[49][Cout][32] += sum_p D[p, co] * u[p + tap] in one tensor-core pass u * [D_hi | D_lo].  The gradient of the
 * reference's conv is (2/255) * dw - R with R[co, tap] = sum of D over the positions where the tap lies inside the image,
 * which the caller forms from otal_border_class_sums.  STAGED like otal_conv1a_fwd_u8. */
OTAL_API int otal_conv1a_wgrad_u8(const otal_conv1a_wgrad_desc* desc, void* stream);
/* The same operator with a resident input halo (csrc/conv1a_wgrad_halo.cu): one 38-row input box per dt serves the seven dh taps as
 * shifted MN-major operand views, a CTA accumulates a pair of dt over its share of the positions in TMEM and flushes once.  Same
 * arguments and result (summation order differs); tT / tH / tW are ignored.  Even extents >= 6, Cout = 64. */
OTAL_API int otal_conv1a_wgrad_u8_halo(const otal_conv1a_wgrad_desc* desc, void* stream);

/* sums[(ct*4+ch)*4+cw][c] += sum over the positions of border class (ct,ch,cw) of (d_hi + d_lo)[n,t,h,w,c] — d: NDHWC bf16
 * planes [N,To,Ho,Wo,d_cstride], channels [d_coff, d_coff + C), C a power of two in 8..128; sums: fp32 [64][C], zeroed by the
 * caller.  Classes as in otal_conv1a_fwd_u8.  d_lo may be NULL. */
OTAL_API int otal_border_class_sums(const uint16_t* d_hi, const uint16_t* d_lo, float* sums, int N, int To, int Ho, int Wo, int C,
                                    int d_cstride, int d_coff, void* stream);

/* MaxPool3dSamePadding on NDHWC planes — replaces AFSD/common/layers.py:9-35 (zero pad + nn.MaxPool3d).
 * Output extent = ceil(input / stride); (pt,ph,pw) = front padding of the "same" rule; padding competes as 0.
 * forward: x planes -> y planes.  backward: g_in[argmax] += g_out (fp32, float reductions; zero g_in first unless it
 * already holds the other consumers' gradient); x planes are the saved forward input, or `argmax` as recorded by the
 * forward pass.  Inputs are compared through order-preserving integer keys of the (hi, lo) pairs: exact for any sign. */
typedef struct otal_pool_desc {
    int N, T, H, W, C;
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int in_cstride, in_coff, out_cstride, out_coff;
    int gout_cstride, gout_coff, gin_cstride, gin_coff;
    const uint16_t* x_hi; const uint16_t* x_lo;
    uint16_t* y_hi; uint16_t* y_lo;
    const float* g_out; float* g_in;
    unsigned char* argmax;   /* optional [N,To,Ho,Wo,C] bytes: forward records the window-relative arg-max (255 = the zero
                                padding won), backward then scatters without re-reading x (x_hi may be NULL in that case) */
} otal_pool_desc;
OTAL_API int otal_maxpool_fwd(const otal_pool_desc* desc, void* stream);
OTAL_API int otal_maxpool_bwd(const otal_pool_desc* desc, void* stream);
/* Pool backward fused with the ReLU / frozen-BN backward of the layer that produced the pool's input (gather form, for
 * the stride-2 stage pools): d = (pool_bwd(g_out) [+ g_add]) * [x > 0] * scale[c], written as bf16 planes
 * [N,T,H,W,d_cstride] at d_coff.  Needs desc->argmax (recorded by the forward), desc->g_out and desc->x_hi (the pool
 * input = the producer's post-ReLU output); g_add: optional fp32 [N,T,H,W,add_cstride] gradient of another consumer. */
OTAL_API int otal_maxpool_bwd_relu_bn_split(const otal_pool_desc* desc, const float* g_add, int add_cstride, int add_coff,
                                            const float* scale, uint16_t* d_hi, uint16_t* d_lo, int d_cstride, int d_coff,
                                            void* stream);

/* Clip ingest for Conv3d_1a — replaces `clips.cuda()` + the first F.pad (AFSD/thumos14/train.py:165,
 * AFSD/common/i3d_backbone.py:59-79): NCDHW fp32 [N,C<=4,T,H,W] (W even) -> [N,T,H,W+8,4] bf16 planes: image column w at
 * padded column w + 2 (the conv's "same" front padding of 2 plus room for the 8-pixel windows at the right edge), 4
 * channel slots per pixel, zero outside the image / for channels >= C.  lo may be NULL. */
OTAL_API int otal_clip_ingest(const float* x, uint16_t* hi, uint16_t* lo, int N, int C, int T, int H, int W, void* stream);

/* The same planes straight from the dataset's storage format — replaces the data loader's crop / flip / normalise
 * (AFSD/common/thumos_dataset.py:239-275 `__getitem__`: videotransforms RandomCrop / RandomHorizontalFlip :30-124, then
 * `(x / 255.0) * 2.0 - 1.0` :261-263) and `clips.cuda()` (AFSD/thumos14/train.py:165) of a 4x larger fp32 tensor.
 * px: uint8 [N,T,Hs,Ws,3] frames (AFSD/common/video2npy.py:61-74); crop: device int32 [N,3] = (row offset, column offset,
 * mirror flag) per sample, or NULL for the centre crop without mirroring.  Normalisation is bit-identical to torch's.
 * frame_map: device int32 [N,T] or NULL (identity): output frame t is source frame frame_map[n*T+t] (clamped to [0,T)).
 * This is the cut-paste augmentation of the SSL pass (`THUMOS_Dataset.augment_`, thumos_dataset.py:187-229: two temporal
 * slice copies on a clone of the fp32 clip = a re-ordering of the clip's own frames), applied while the frames are read. */
OTAL_API int otal_clip_ingest_u8(const unsigned char* px, const int* crop, const int* frame_map, uint16_t* hi, uint16_t* lo, int N,
                                 int T, int Hs, int Ws, int H, int W, void* stream);

/* The same crop / mirror / temporal gather, but the ONE output plane holds the pixel values 0..255 (exact in bf16) instead
 * of the normalised clip: the input of otal_conv1a_fwd_u8 / otal_conv1a_wgrad_u8. */
OTAL_API int otal_clip_ingest_u8_raw(const unsigned char* px, const int* crop, const int* frame_map, uint16_t* out, int N, int T,
                                     int Hs, int Ws, int H, int W, void* stream);

/* Backward of relu(conv*scale+shift) w.r.t. the conv output, fused with the hi/lo split the tensor-core kernels read:
 * d = g * [y > 0] * scale[c]   (torch relu backward + frozen BatchNorm3d, AFSD/thumos14/BDNet.py:39-49).
 * g fp32 [npos, g_cstride] at g_coff, y_hi bf16 plane of the forward output, d planes.  relu = 0 skips the mask,
 * scale may be NULL (=1), d_lo may be NULL. */
OTAL_API int otal_relu_bn_bwd_split(const float* g, const uint16_t* y_hi, const float* scale, uint16_t* d_hi, uint16_t* d_lo,
                                    long long npos, int C, int g_cstride, int g_coff, int y_cstride, int y_coff,
                                    int d_cstride, int d_coff, int relu, void* stream);

/* Adam with L2-in-gradient weight decay over a flat fp32 buffer — replaces torch.optim.Adam.step as configured by
 * AFSD/thumos14/train.py:321-323.  g is multiplied by grad_scale first (1/world_size after a summing all-reduce). */
OTAL_API int otal_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, float grad_scale, int step, void* stream);
/* The same update with the step counter t >= 1 read from DEVICE memory (step_dev[0], incremented by the caller beforehand): the
 * launch can be captured into a CUDA graph together with the gradient all-reduce (bias corrections are computed on the device). */
OTAL_API int otal_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                                float eps, float weight_decay, float grad_scale, const int* step_dev, void* stream);

/* [B,C,T] fp32 (the reference's conv1d layout) -> channels-last bf16 planes [B,Ttot,Cpad]: element (b,c,t) lands at
 * position offset + t*dilate, channel c.  With dilate > 1 or Cpad > C the caller zero-fills the planes first (this is
 * the zero-upsampled output gradient a stride-2 conv's dgrad reads).  lo may be NULL. */
OTAL_API int otal_ncl_to_nlc_split(const float* x, uint16_t* hi, uint16_t* lo, int B, int C, int T, int Cpad, int Ttot,
                                   int dilate, int offset, void* stream);

/* fp32 -> (hi, lo) bf16 planes, elementwise over n values (layout preserving). lo may be NULL. */
OTAL_API int otal_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, long long n, void* stream);
/* (hi, lo) -> fp32.  lo may be NULL. */
OTAL_API int otal_merge_bf16(const uint16_t* hi, const uint16_t* lo, float* x, long long n, void* stream);
/* NCDHW fp32 -> NDHWC (hi, lo) planes with the channel dimension zero-padded to Cpad (clip ingest).
 * Replaces the implicit layout of `clips.cuda()` + first F.pad (AFSD/thumos14/train.py:165). */
OTAL_API int otal_ncdhw_to_ndhwc_split(const float* x, uint16_t* hi, uint16_t* lo, int N, int C, int T, int H, int W,
                              int Cpad, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * MultiSegmentLoss, three flavours (desc.flavour):
 *  0  THUMOS14 OpenTAL (cls_loss_type 'edl' with loss_type 'log', evidence 'exp', os_head) — replaces
 *            MultiSegmentLoss.forward                 AFSD/thumos14/multisegment_loss.py:92-259 (iou_loss :20-53)
 *            EvidenceLoss.forward/edl_loss/iou_calib  AFSD/thumos14/cls_loss.py:120-168, :212-278
 *            ActionnessLoss.forward                   AFSD/thumos14/cls_loss.py:299-339
 *  1  ActivityNet OpenTAL — replaces MultiSegmentLoss.forward AFSD/anet/multisegment_loss.py:106-301 (level_bounds = its
 *            `bounds` :69-83 as (left, right] per pyramid level, the level read from priors[p*prior_stride + 1]; per-sample
 *            normalisation, smooth-L1, min(piou, max IoU) refinement threshold) and AFSD/anet/cls_loss.py:116-152, :225-232
 *            (stateless IBM weight 1 / (||z||_1 exp(ibm_coeff g) + 1e-10); weight_accum unused)
 *  2  THUMOS14 closed set (configs/thumos14.yaml: cls_loss_type 'focal', no os_head) — multisegment_loss.py:92-259 with
 *            FocalLoss_Ori AFSD/thumos14/cls_loss.py:6-78 on the softmax of all priors (K counts the background class 0;
 *            focal_alpha = weight of class 0, 1 - focal_alpha of the others; act / prop_act NULL; losses[5], [6] = 0)
 * and their autograd backward.  One single-CTA launch computes the 7 losses and the gradient of each loss w.r.t. each
 * head output ("unit gradients", stored in `workspace`); otal_msl_backward scales them by the 7 upstream gradients.
 *
 * Inputs are the reference's tensors, contiguous fp32: loc/prop_loc [B,P,2], conf/prop_conf [B,P,K], center/act/
 * prop_act [B,P] (act / prop_act may be NULL: no os_head), priors element p at priors[p*prior_stride], targets
 * [B,G,3] = (start, end, label 1..K) normalised to the clip, zero padded, valid [B,G] bytes (1 = real row).
 * weight_accum [num_bins] is the IBM EMA state (cls_loss.py:114), updated in place when use_ibm != 0 (the caller sets
 * use_ibm = epoch >= ibm_start).  losses [16]: loss_l, loss_c, loss_prop_l, loss_prop_c, loss_ct, loss_act,
 * loss_prop_act, then N, PN, AN, PAN, loss_iouc.  workspace: otal_msl_workspace_floats(B,P,K) floats.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct otal_msl_desc {
    int B, P, K, G;
    float clip_length, overlap_thresh;
    int use_ibm, num_bins;
    float momentum;
    int iou_aware;
    float act_weight, act_margin;
    int prior_stride;
    const float* loc; const float* conf; const float* prop_loc; const float* prop_conf;
    const float* center; const float* act; const float* prop_act;
    const float* priors; const float* targets; const unsigned char* valid;
    float* weight_accum;
    float* losses;
    float* workspace;
    int flavour;                 /* 0 THUMOS14 EDL, 1 ActivityNet EDL, 2 THUMOS14 closed-set focal */
    float ibm_coeff;             /* flavour 1 */
    float focal_alpha, focal_gamma;   /* flavour 2 */
    float level_bounds[16];      /* flavour 1: (left, right) per pyramid level, up to 8 levels */
    /* flavour 0, the re-weighting branches of EvidenceLoss.edl_loss (AFSD/thumos14/cls_loss.py:221-272; one at a time, in the
     * reference's precedence focal > GHM > IB > IBM): 0 = use_ibm decides (IBM or none), 1 IBM, 2 IB (1 / (grad_norm feat_norm)),
     * 3 focal-EDL (edl_focal_alpha = weight of class 0, edl_focal_gamma; the modulating factor is differentiated), 4 GHM
     * (num_bins bins, momentum; ghm_acc_sum = the fp64 acc_sum state [num_bins], updated when momentum > 0).  The caller selects
     * a branch only from its start epoch on (ghm_start, ib_start, ibm_start). */
    int reweight;
    int cls_all;                 /* flavour 0: 1 = no os_head (configs/ablations/thumos14_opental_noACT.yaml): every prior is a
                                  * classification sample with class 0 = background (K counts it); act / prop_act NULL */
    float edl_focal_alpha, edl_focal_gamma;
    double* ghm_acc_sum;
} otal_msl_desc;
OTAL_API long long otal_msl_workspace_floats(int B, int P, int K);
OTAL_API int otal_msl_forward(const otal_msl_desc* desc, void* stream);
/* grad_losses [7] (device): upstream gradients of the 7 losses.  g_act / g_prop_act may be NULL. */
OTAL_API int otal_msl_backward(int B, int P, int K, const float* workspace, const float* grad_losses, float* g_loc,
                               float* g_conf, float* g_prop_loc, float* g_prop_conf, float* g_center, float* g_act,
                               float* g_prop_act, void* stream);

/* GroupNorm(groups, C) + ReLU on [B,C,T] fp32 — replaces nn.GroupNorm(32, C) + nn.ReLU(inplace=True) after every
 * pyramid / tower / proposal-branch / deconv conv (AFSD/thumos14/BDNet.py:72-73, :139-140, :166-167, :176-177, :276-283).
 * Segments: nseg (0..8) column ranges [seg_off[s], seg_off[s] + seg_len[s]) along T (HOST int arrays) are normalised
 * independently — the 6 pyramid levels laid side by side, which share these modules' weights (BDNet.py:333-412) — and
 * every column outside the ranges is written as 0 (forward) / receives gradient 0 (backward).  nseg = 0: one segment
 * [0,T), i.e. plain GroupNorm.
 * forward also writes the per-(sample, group, segment) mean and 1/sqrt(var + eps) ([B*groups*max(nseg,1)] each) that
 * backward consumes.  backward: gx [B,C,T]; dgamma_dbeta [B,2,C] = per-sample partial sums of (d gamma, d beta) — sum
 * over B for the parameter gradients.  relu = 0 gives plain GroupNorm. */
OTAL_API int otal_groupnorm_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                     int B, int C, int T, int groups, float eps, int relu, int nseg, const int* seg_off,
                                     const int* seg_len, void* stream);
OTAL_API int otal_groupnorm_relu_bwd(const float* gy, const float* x, const float* gamma, const float* beta, const float* mean,
                                     const float* rstd, float* gx, float* dgamma_dbeta, int B, int C, int T, int groups,
                                     int relu, int nseg, const int* seg_off, const int* seg_len, void* stream);

/* Extended GroupNorm + ReLU for the explicit head schedule: same arithmetic and segments as above, more destinations.
 * forward (reads x, gamma, beta; writes mean, rstd): y [B,C,T] fp32 (may be NULL) and / or the result as channels-last bf16
 *   planes p_hi / p_lo [B,T,p_cstride] at channel offset p_coff (p_lo may be NULL) — the operand layout of the tensor-core conv
 *   that consumes it, or a slice of the concat buffer of ProposalBranch.forward (AFSD/thumos14/BDNet.py:111).
 * backward (reads gy, x, gamma, beta, mean, rstd): gy element (b,c,t) at gy[b*gy_bstride + c*T + t] (gy_bstride = 0: C*T), i.e.
 *   a channel slice of a wider gradient tensor; gx [B,C,T] fp32 (may be NULL) and / or channels-last planes d_hi / d_lo [B,T,C];
 *   dgamma[c], dbeta[c] and (if not NULL) dbias[c] += sum over (b,t) of gx are ACCUMULATED atomically. */
typedef struct otal_gn_desc {
    int B, C, T, groups;
    float eps;
    int relu, nseg;
    int seg_off[8], seg_len[8];
    const float* x; const float* gamma; const float* beta;
    float* mean; float* rstd;
    float* y; uint16_t* p_hi; uint16_t* p_lo; int p_cstride, p_coff;          /* forward */
    float* yt; int yt_off, yt_T;     /* forward: optional fp32 channels-last copy [B,yt_T,C] of the columns [yt_off, yt_off+yt_T) */
    const float* gy; long long gy_bstride;                                      /* backward (gy may be NULL if gy2a / gy2b is given) */
    const float* gy2a; const float* gy2b; int gy2_off, gy2_T;   /* backward: optional extra gradient, channels-last [B,gy2_T,C/2] per
                                                                 * channel half, for the columns [gy2_off, gy2_off+gy2_T) */
    float* gx; uint16_t* d_hi; uint16_t* d_lo; float* dgamma; float* dbeta; float* dbias;
} otal_gn_desc;
OTAL_API int otal_groupnorm_relu_fwd_ex(const otal_gn_desc* desc, void* stream);
OTAL_API int otal_groupnorm_relu_bwd_ex(const otal_gn_desc* desc, void* stream);

/* Glue of the explicit head schedule: what CoarsePyramid.forward does with F.interpolate, +, torch.cat, index_select and permute
 * between its convolutions (AFSD/thumos14/BDNet.py:311-331, :340-353, :399-412; AFSD/anet/BDNet.py:281-311) and the transposes of
 * those steps in the backward.
 * otal_rows_combine: dst[b,c,j] = sum over k < npairs of src[table[j][k][0]][b,c,table[j][k][1]] (source index -1 = no term);
 *   sources [B,C,src_T[i]] fp32, `table` a DEVICE int32 array [Td][npairs][2]; the result goes to dst [B,C,Td] fp32 and / or
 *   channels-last bf16 planes p_hi / p_lo [B,Td,C] (either destination may be NULL). */
typedef struct otal_rows_desc {
    int B, C, Td, npairs, nsrc;
    const float* src[8];
    int src_T[8];
    const int* table;
    float* dst; uint16_t* p_hi; uint16_t* p_lo;
} otal_rows_desc;
OTAL_API int otal_rows_combine(const otal_rows_desc* desc, void* stream);
/* otal_head_gather_fwd: up to 4 head convolutions' raw outputs raw[k] [B,cpad[k],S] (S columns in the level-separated layout,
 *   channels padded to cpad) -> out[k] [B,P,cout[k]], prior p read from column sep_idx[p] (DEVICE int32 [P]).  mode[k] = 1 applies
 *   ScaleExp (BDNet.py:55-61, :341-346): exp(x * *scale[level_id[p]]) (* mult[p] if mult != NULL: anet/BDNet.py:307-311).
 * otal_head_gather_bwd: gout[k] [B,P,cout[k]] (NULL = zero) -> the raw outputs' gradient as channels-last planes d_hi / d_lo
 *   [B,S,cpad[k]] (zero in separator columns and padded channels; prior_of_col DEVICE int32 [S], -1 = separator); mode 1 reads the
 *   forward's out[k] and raw[k] and accumulates *dscale[level] atomically. */
typedef struct otal_headout_desc {
    int B, S, P, n;
    const int* sep_idx; const int* level_id; const float* mult;
    const float* scale[8]; float* dscale[8];
    const float* raw[4]; int cpad[4], cout[4], mode[4];
    float* out[4]; const float* gout[4]; uint16_t* d_hi[4]; uint16_t* d_lo[4];
    const float* bias[4];        /* [cout[k]] or NULL: added to raw before everything else (the conv's bias, when its padded output
                                  * channels keep it out of the conv epilogue) */
    float* dbias[4];             /* backward: its gradient, accumulated; or NULL */
} otal_headout_desc;
OTAL_API int otal_head_gather_fwd(const otal_headout_desc* desc, void* stream);
OTAL_API int otal_head_gather_bwd(const otal_headout_desc* desc, const int* prior_of_col, void* stream);
/* [B,C,T] fp32 (sample stride x_bstride elements, 0 = C*T) -> channels-last planes [B,T,cstride] at channel offset coff (a slice of
 * the concat buffer of ProposalBranch.forward, BDNet.py:111).  C, cstride, coff multiples of 8.  lo may be NULL. */
OTAL_API int otal_ncl_to_nlc_split_ex(const float* x, long long x_bstride, uint16_t* hi, uint16_t* lo, int B, int C, int T,
                                      int cstride, int coff, void* stream);
/* otal_boundary_bce_fwd / _bwd (below) on rows that are a channel slice of wider channels-last rows: row r starts at
 * x + r * x_rstride (x_rstride >= C elements); grad_x is dense [B,T,C]. */
OTAL_API int otal_boundary_bce_fwd_ex(const float* x, int x_rstride, const float* target, long long target_batch_stride,
                                      float* row_loss, float* coef, int B, int T, int C, void* stream);
OTAL_API int otal_boundary_bce_bwd_ex(const float* x, int x_rstride, const float* coef, const float* grad_loss, float* grad_x,
                                      int B, int T, int C, void* stream);

/* out[s] = mean of x[offsets[s] .. offsets[s+1]) for s < nseg <= 16 (offsets: HOST array of nseg + 1 element offsets), fixed
 * summation order: the six boundary-BCE means of a training step (AFSD/thumos14/train.py:152-161, :186-200) in one launch. */
OTAL_API int otal_segment_mean(const float* x, const long long* offsets_host, int nseg, float* out, void* stream);

/* Proposal window generation for all pyramid levels at once — replaces the no_grad block AFSD/thumos14/BDNet.py:355-384.
 * loc [B,P,2] (frames) over the P = sum of level lengths priors; per-prior tables prior [P] ((c+0.5)/t), level_len [P]
 * (t of the prior's level), level_off [P] (first column of that level in the level-concatenated feature).
 *   seg_level  [B,P,4] (optional, may be NULL): the reference's `segments` in level units, bit-identical to torch
 *   seg_concat [B,P,4]: the same windows truncated + clamped to [0, t-1] (what boundary_max_pooling_kernel.cu:33-38 does
 *                       with them) + level_off: windows into the level-concatenated feature
 *   frame_seg  [B,P,4]: the reference's `frame_segments` (frame units), bit-identical to torch */
OTAL_API int otal_make_segments(const float* loc, const float* prior, const int* level_len, const int* level_off,
                                float* seg_level, float* seg_concat, float* frame_seg, int B, int P, float frame_num, void* stream);
/* The same windows written to rows out_row[p] of [B,S,4] arrays (DEVICE int32 [P]; the other rows are left untouched): the
 * level-separated layout of the explicit head schedule, where level_off holds the levels' first columns in that layout. */
OTAL_API int otal_make_segments_ex(const float* loc, const float* prior, const int* level_len, const int* level_off, const int* out_row,
                                   int S, float* seg_concat, float* frame_seg, int B, int P, float frame_num, void* stream);

/* DirichletLayer.compute_uncertainty, 'exp' evidence (AFSD/thumos14/BDNet.py:544-556): unct[m] = K / sum_k(exp(clamp(
 * logit[m,k], -10, 10)) + 1).  logit [M,K] contiguous. */
OTAL_API int otal_dirichlet_uncertainty(const float* logit, float* unct, long long M, int K, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Inference post-processing — replaces decode_predictions (AFSD/thumos14/test.py:112-140) and softnms_v2
 * (AFSD/common/segment_utils.py:128-162, a Python loop on the CPU in the reference).
 *
 * otal_decode_scores: for B sliding-window clips, loc/prop_loc [B,P,2], conf/prop_conf [B,P,K] logits, center/act/
 * prop_act [B,P] (act, prop_act NULL = closed-set head), prior [P] centres, offset [B] = first frame of each clip (or
 * NULL).  segments [B,P,2] = refined (start, end) in seconds, scores [B,K,P] = mean Dirichlet probability of the two
 * heads x sigmoid(center) x actionness, uncertainty [B,P] = mean of K / sum(alpha), actionness [B,P].
 *
 * otal_softnms: Gaussian soft-NMS per class.  segments [.., M, 2] (class c reads segments + c*seg_class_stride floats;
 * stride 0 = all classes share one candidate list), scores [C,M] decayed IN PLACE (entries below score_threshold never
 * take part: set filtered-out candidates to 0), keep [C,M] bytes = 1 for the kept candidates, count [C].
 * ---------------------------------------------------------------------------------------------------------- */
OTAL_API int otal_decode_scores(const float* loc, const float* prop_loc, const float* conf, const float* prop_conf,
                                const float* center, const float* act, const float* prop_act, const float* prior,
                                const float* offset, float* segments, float* scores, float* uncertainty, float* actionness,
                                int B, int P, int K, float clip_length, float sample_fps, void* stream);
OTAL_API int otal_softnms(const float* segments, long long seg_class_stride, float* scores, unsigned char* keep, int* count, int C,
                          int M, float sigma, int top_k, float score_threshold, void* stream);

/* Boundary BCE of the training script — replaces calc_bce_loss (AFSD/thumos14/train.py:152-161; anet/train.py:134-143):
 * loss = mean over rows r = (b, t) of BCE(mean_c tanh(x[b,t,c]), target[b*target_batch_stride + t]).  x [B,T,C] contiguous.
 * forward writes the per-row losses row_loss [B*T] (their mean is the loss) and coef [B*T] for the backward;
 * backward: grad_x = grad_loss[0] * coef[r] * (1 - tanh(x)^2). */
OTAL_API int otal_boundary_bce_fwd(const float* x, const float* target, long long target_batch_stride, float* row_loss, float* coef,
                                   int B, int T, int C, void* stream);
OTAL_API int otal_boundary_bce_bwd(const float* x, const float* coef, const float* grad_loss, float* grad_x, int B, int T, int C,
                                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPENTAL_B200_H */

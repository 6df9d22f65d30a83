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
 * Output extent equals input extent (pad front = pt/ph/pw, the rest of the "same" padding is implicit zero).
 * (tT,tH,tW) is the 128-position tile box, tT*tH*tW == 128.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct otal_conv_desc {
    int N, T, H, W;          /* batch and spatial extent (output == input extent) */
    int Cin, Cout;           /* channels read / written by this launch */
    int kt, kh, kw;          /* filter taps */
    int pt, ph, pw;          /* front padding */
    int tT, tH, tW;          /* tile box */
    int nsplit;              /* 1 = bf16, 3 = bf16x3 */
    int relu;
    int in_cstride, in_coff;   /* input row width and slice offset, in channels */
    int out_cstride, out_coff; /* output row width and slice offset, in channels (bf16 planes and fp32 alike) */
    const uint16_t* x_hi; const uint16_t* x_lo;   /* bf16 bit patterns */
    const uint16_t* w_hi; const uint16_t* w_lo;
    const float* scale; const float* shift;       /* [Cout] or NULL */
    uint16_t* y_hi; uint16_t* y_lo;               /* NULL = do not store bf16 planes */
    float* y_f32;                                 /* NULL = do not store fp32 */
} otal_conv_desc;

OTAL_API int otal_conv_igemm_fwd(const otal_conv_desc* desc, void* stream);

/* fp32 -> (hi, lo) bf16 planes, elementwise over n values (layout preserving). lo may be NULL. */
OTAL_API int otal_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, long long n, void* stream);
/* (hi, lo) -> fp32.  lo may be NULL. */
OTAL_API int otal_merge_bf16(const uint16_t* hi, const uint16_t* lo, float* x, long long n, void* stream);
/* NCDHW fp32 -> NDHWC (hi, lo) planes with the channel dimension zero-padded to Cpad (clip ingest).
 * Replaces the implicit layout of `clips.cuda()` + first F.pad (AFSD/thumos14/train.py:165). */
OTAL_API int otal_ncdhw_to_ndhwc_split(const float* x, uint16_t* hi, uint16_t* lo, int N, int C, int T, int H, int W,
                              int Cpad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPENTAL_B200_H */

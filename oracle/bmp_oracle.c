/* TEST INFRASTRUCTURE — plain-C restatement of BoundaryMaxPooling, independent of the torch formulation in
 * opental_oracle.py (two restatements + the reference's own CUDA kernel in oracle/_ref check the product operator).
 * Follows AFSD/prop_pooling/boundary_max_pooling_kernel.cu:
 *   window of output (n, c, k): columns 0:2 of segments[n, k] for the first half of the channels, 2:4 for the second
 *   (:30-31); float -> int by C truncation, then clamp to [0, extent-1] (:33-36); the scan starts AT l, so r < l
 *   degenerates to the single element l; strict '>' keeps the first maximum (:37-43, :67-78).
 *   backward (:49-82, :114-145): the gradient of output (n, c, k) goes to the arg-max position.  The reference launcher
 *   passes grad_output.size(2) = K as the time extent (:121) for both clamping and addressing — `compat` reproduces that
 *   (the flat buffers are re-read as rows of K values), compat = 0 is the correct gradient.
 * Built by oracle/build_c.py (gcc -O2 -shared); only tests/ load it. */
#include <stddef.h>

static int window(const float* seg4, int second_half, int extent, int* right) {
    int l = (int)seg4[second_half ? 2 : 0];
    int r = (int)seg4[second_half ? 3 : 1];
    if (l < 0) l = 0;
    if (l > extent - 1) l = extent - 1;
    if (r < 0) r = 0;
    if (r > extent - 1) r = extent - 1;
    *right = r;
    return l;
}

static int first_max(const float* row, int l, int r) {
    int best = l;
    for (int i = l + 1; i <= r; ++i)
        if (row[i] > row[best]) best = i;
    return best;
}

void bmp_oracle_forward(const float* in, const float* seg, float* out, int B, int C, int T, int K) {
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c) {
            const float* row = in + ((size_t)n * C + c) * T;
            for (int k = 0; k < K; ++k) {
                int r, l = window(seg + ((size_t)n * K + k) * 4, c >= C / 2, T, &r);
                out[((size_t)n * C + c) * K + k] = row[first_max(row, l, r)];
            }
        }
}

/* grad_in must hold B*C*T floats; it is zero-filled here. */
void bmp_oracle_backward(const float* gout, const float* in, const float* seg, float* gin, int B, int C, int T, int K, int compat) {
    const int ext = compat ? K : T;            /* the extent the reference kernel is launched with */
    for (size_t i = 0; i < (size_t)B * C * T; ++i) gin[i] = 0.f;
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c) {
            const size_t base = ((size_t)n * C + c) * ext;     /* row start in the (possibly mis-strided) flat view */
            for (int k = 0; k < K; ++k) {
                int r, l = window(seg + ((size_t)n * K + k) * 4, c >= C / 2, ext, &r);
                gin[base + first_max(in + base, l, r)] += gout[((size_t)n * C + c) * K + k];
            }
        }
}

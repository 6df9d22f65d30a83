// Developer probe (not part of the library): how fast can ONE CTA per SM stream [128 rows x 64 bf16] operand boxes from HBM / L2
// into shared memory, as a function of the pipeline depth and of the mechanism (TMA tiled boxes vs cp.async 16-byte copies)?
// Decides how the memory-bound convolutions (1x1, Conv3d_1a) should be fed.  Build: see tools/probe/build.sh.
#include "../../opental_b200/csrc/common.cuh"
#include "../../opental_b200/csrc/tensormap.h"
#include <vector>
#include <stdlib.h>

using namespace otal;

struct ProbeParams {
    int rows;            // rows of the [rows, C] bf16 matrix (positions)
    int C;               // channels (row pitch = 2*C bytes)
    int nstages;         // pipeline depth
    int boxes;           // boxes (16 KB each) per stage: consecutive 64-channel chunks of the tile, wrapping to the next tile
    int planes;          // 1 or 2 source tensors (hi / lo)
    int wboxes;          // extra boxes per stage from a small L2-resident tensor (the weights)
    int consume_cycles;  // the consumer holds a stage this long before releasing it (models the MMA)
    int mode;            // 0 = TMA 2-D boxes, 1 = cp.async by 4 warps, 2 = TMA 5-D boxes (4x4x8 positions)
    const uint16_t* src0; const uint16_t* src1; const uint16_t* wsrc;
    int tiles;
};
struct alignas(64) ProbeMaps { CUtensorMap A[2]; CUtensorMap W; CUtensorMap A5[2]; };

__global__ void __launch_bounds__(256, 1) probe_kernel(const __grid_constant__ ProbeMaps maps, const ProbeParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = (uint32_t)(p.boxes * p.planes + p.wboxes) * 16384u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stage_bytes * p.nstages);
    uint64_t* empty = full + 8;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nstages; ++i) { mbar_init(&full[i], p.mode == 1 ? 128 : 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int kchunks = p.C / 64;
    const int total_boxes = p.tiles * kchunks;           // A boxes (per plane) of the whole problem
    const int stages_total = (total_boxes + p.boxes - 1) / p.boxes;
    if (p.mode != 1 && warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int s = blockIdx.x; s < stages_total; s += gridDim.x) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
                unsigned char* dst = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx(&full[stage], stage_bytes);
                for (int b = 0; b < p.boxes; ++b) {
                    int box = s * p.boxes + b; if (box >= total_boxes) box = total_boxes - 1;
                    const int tile = box / kchunks, kc = box % kchunks;
                    for (int pl = 0; pl < p.planes; ++pl) {
                        if (p.mode == 0) tma_load_2d(&maps.A[pl], &full[stage], dst, kc * 64, tile * 128);
                        else {
                            // 5-D view [N, T, 12, 12, C] with 4x4x8 position boxes: tile -> (n, t0, h0, w0)
                            int m = tile; const int w0 = (m % 3) * 4; m /= 3; const int h0 = (m % 3) * 4; m /= 3;
                            const int t0 = (m % 16) * 8; m /= 16;
                            tma_load_5d(&maps.A5[pl], &full[stage], dst, kc * 64, w0, h0, t0, m);
                        }
                        dst += 16384;
                    }
                }
                for (int b = 0; b < p.wboxes; ++b) { tma_load_2d(&maps.W, &full[stage], dst, 0, ((s + b) & 7) * 128); dst += 16384; }
            }
            __syncwarp();
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
    } else if (p.mode == 1 && warp >= 4) {
        // cp.async producers: 128 threads, a box = 128 rows x 8 x 16-byte chunks; thread t copies chunk (t & 7) of rows (t >> 3) + 16 j
        const int t = threadIdx.x - 128;
        int stage = 0; uint32_t phase = 0;
        for (int s = blockIdx.x; s < stages_total; s += gridDim.x) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* dst = smem + (size_t)stage * stage_bytes;
            for (int b = 0; b < p.boxes; ++b) {
                int box = s * p.boxes + b; if (box >= total_boxes) box = total_boxes - 1;
                const int tile = box / kchunks, kc = box % kchunks;
                for (int pl = 0; pl < p.planes; ++pl) {
                    const uint16_t* src = (pl ? p.src1 : p.src0) + ((size_t)tile * 128) * p.C + kc * 64;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int row = (t >> 3) + 16 * j, ch = t & 7;
                        const uint32_t d = smem_u32(dst) + (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
                        const void* g = src + (size_t)row * p.C + ch * 8;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
                    }
                    dst += 16384;
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        for (int s = blockIdx.x; s < stages_total; s += gridDim.x) {
            mbar_wait(&full[stage], phase);
            if (p.consume_cycles > 0) { const long long t0 = clock64(); while (clock64() - t0 < p.consume_cycles) {} }
            if (lane == 0) mbar_arrive(&empty[stage]);
            __syncwarp();
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
    }
}

static float run(const ProbeMaps& maps, ProbeParams p, int reps = 5) {
    const size_t smem = (size_t)(p.boxes * p.planes + p.wboxes) * 16384 * p.nstages + 256 + 1024;
    if (smem > 227 * 1024) return -1.f;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe_kernel<<<148, 256, smem>>>(maps, p);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) probe_kernel<<<148, 256, smem>>>(maps, p);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int kNT = 8 * 128, HW = 144;                    // 8 clips x 128 frames x 12 x 12 positions = 147456 rows (Mixed_3b/3c)
    const int rows = kNT * HW;
    printf("# rows %d; GB/s = A bytes (+ W bytes) moved / kernel time; B200, 148 CTAs x 1 per SM\n", rows);
    for (int C : {64, 256, 512}) {
        const size_t n = (size_t)rows * C;
        uint16_t *a0, *a1, *w;
        cudaMalloc(&a0, n * 2); cudaMalloc(&a1, n * 2); cudaMalloc(&w, 1024 * 64 * 2);
        cudaMemset(a0, 0, n * 2); cudaMemset(a1, 0, n * 2); cudaMemset(w, 0, 1024 * 64 * 2);
        ProbeMaps maps; memset(&maps, 0, sizeof(maps));
        const uint64_t d2[2] = {(uint64_t)C, (uint64_t)rows}; const uint64_t s2[1] = {(uint64_t)C * 2}; const uint32_t b2[2] = {64, 128};
        make_tensor_map_bf16(&maps.A[0], a0, 2, d2, s2, b2, 1); make_tensor_map_bf16(&maps.A[1], a1, 2, d2, s2, b2, 1);
        const uint64_t dw[2] = {64, 1024}; const uint64_t sw[1] = {128};
        make_tensor_map_bf16(&maps.W, w, 2, dw, sw, b2, 1);
        const uint64_t d5[5] = {(uint64_t)C, 12, 12, 128, 8}; const uint64_t cs = (uint64_t)C * 2;
        const uint64_t s5[4] = {cs, cs * 12, cs * 144, cs * 144 * 128}; const uint32_t b5[5] = {64, 4, 4, 8, 1};
        make_tensor_map_bf16(&maps.A5[0], a0, 5, d5, s5, b5, 1); make_tensor_map_bf16(&maps.A5[1], a1, 5, d5, s5, b5, 1);
        for (int mode : {0, 2, 1}) for (int planes : {1, 2}) for (int wboxes : {0, 2}) for (int boxes : {1, 2}) for (int nst : {2, 3, 4, 6}) {
            if (mode == 2 && C % 64) continue;
            if (mode == 1 && wboxes) continue;
            ProbeParams p{};
            p.rows = rows; p.C = C; p.nstages = nst; p.boxes = boxes; p.planes = planes; p.wboxes = wboxes; p.consume_cycles = 0;
            p.mode = mode; p.src0 = a0; p.src1 = a1; p.wsrc = w; p.tiles = rows / 128;
            const float ms = run(maps, p);
            if (ms < 0) continue;
            const double abytes = (double)n * 2 * planes, wbytes = (double)(p.tiles * (C / 64) / boxes) * wboxes * 16384.0;
            printf("C %3d mode %s planes %d wboxes %d boxes/stage %d stages %d in-flight %3d KB : %7.3f ms  A %6.0f GB/s  A+W %6.0f GB/s\n", C,
                   mode == 0 ? "tma2d" : mode == 2 ? "tma5d" : "cpasy", planes, wboxes, boxes, nst,
                   (boxes * planes + wboxes) * 16 * nst, ms, abytes / ms * 1e-6, (abytes + wbytes) / ms * 1e-6);
        }
        cudaFree(a0); cudaFree(a1); cudaFree(w);
    }
    return 0;
}

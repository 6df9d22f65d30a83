// C-ABI plumbing shared by all kernels: last-error string, version, TMA tensor-map encoding.
#include "common.cuh"
#include "tensormap.h"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>

namespace otal {

static thread_local char g_err[512] = "";

void set_last_error(const char* what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
void set_last_error_msg(const char* what) { snprintf(g_err, sizeof(g_err), "%s", what); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
            set_last_error_msg("cuTensorMapEncodeTiled entry point not available");
            return nullptr;
        }
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swz) {
    auto fn = get_encode();
    if (!fn) return OTAL_ERR_DRIVER;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof(g_err),
                 "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] "
                 "stride0 %llu base %p",
                 (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                 (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                 (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                 rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                 base);
        return OTAL_ERR_DRIVER;
    }
    return OTAL_OK;
}

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
    return make_tensor_map(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, nullptr,
                           swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                           : swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace otal

extern "C" {
const char* otal_last_error(void) { return otal::g_err; }
int otal_abi_version(void) { return OTAL_ABI_VERSION; }

// sizeof of a descriptor struct by name (0 = unknown): lets a binding check its mirror of the struct against this build.
int otal_abi_sizeof(const char* name) {
    if (!name) return 0;
    if (!strcmp(name, "otal_conv_desc")) return (int)sizeof(otal_conv_desc);
    if (!strcmp(name, "otal_conv1a_desc")) return (int)sizeof(otal_conv1a_desc);
    if (!strcmp(name, "otal_wgrad_desc")) return (int)sizeof(otal_wgrad_desc);
    if (!strcmp(name, "otal_conv1a_wgrad_desc")) return (int)sizeof(otal_conv1a_wgrad_desc);
    if (!strcmp(name, "otal_pool_desc")) return (int)sizeof(otal_pool_desc);
    if (!strcmp(name, "otal_msl_desc")) return (int)sizeof(otal_msl_desc);
    if (!strcmp(name, "otal_gn_desc")) return (int)sizeof(otal_gn_desc);
    if (!strcmp(name, "otal_rows_desc")) return (int)sizeof(otal_rows_desc);
    if (!strcmp(name, "otal_headout_desc")) return (int)sizeof(otal_headout_desc);
    return 0;
}
}

// minimal TMA sanity probe (scratch): one box load, print values
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap* gtm, const float* feat, int c0, int c1, int c2, int bytes, float* out, int n, int variant) {
    // (variant >= 10: same as variant - 10 with the destination rounded up to 1024 bytes)
    extern __shared__ __align__(1024) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar);
    unsigned dst = (unsigned)__cvta_generic_to_shared(dyn);
    if (threadIdx.x == 0) printf("dyn smem address 0x%x (mod 1024 = %u, mod 128 = %u)\n", dst, dst & 1023u, dst & 127u);
    if (variant >= 10) { dst = (dst + 1023u) & ~1023u; variant -= 10; }      // force a 1024-byte aligned destination
    const unsigned char* src_s = dyn + (dst - (unsigned)__cvta_generic_to_shared(dyn));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (variant == 6) {
        if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" :: "l"(&tm) : "memory");
        __syncthreads();
        if (threadIdx.x == 0) out[0] = 42.f;
        return;
    }
    if (variant == 4) {
        if (threadIdx.x < 32) {
            unsigned pred = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) {
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar_s), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             :: "r"(dst), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar_s) : "memory");
            }
        }
    } else if (threadIdx.x == 0) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar_s), "r"(bytes) : "memory");
        if (variant == 0)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(dst), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar_s) : "memory");
        else if (variant == 1)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(gtm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar_s) : "memory");
        else if (variant == 3)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(c0), "r"(c1), "r"(bar_s) : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(feat), "r"(bytes), "r"(bar_s) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(bar_s), "r"(0) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<const float*>(src_s)[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    // args: order bx by bz variant
    const int order = atoi(argv[1]), bx = atoi(argv[2]), b1 = atoi(argv[3]), b2 = atoi(argv[4]), variant = atoi(argv[5]), cfgbits = argc > 6 ? atoi(argv[6]) : 0;
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    printf("entry %p q=%d\n", (void*)enc, (int)q);
    const int NC = 64, H = 24, W = 168;
    std::vector<float> h((size_t)NC * H * W);
    for (int c = 0; c < NC; c++) for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) h[((size_t)c * H + y) * W + x] = c * 10000 + y * 100 + x * 0.1f;
    float* feat; CK(cudaMalloc(&feat, h.size() * 4)); CK(cudaMemcpy(feat, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t dims[3], strides[2]; cuuint32_t box[3], es[3] = {1, 1, 1};
    if (order == 0) { dims[0] = W; dims[1] = NC; dims[2] = H; strides[0] = (cuuint64_t)H * W * 4; strides[1] = (cuuint64_t)W * 4; }
    else { dims[0] = W; dims[1] = H; dims[2] = NC; strides[0] = (cuuint64_t)W * 4; strides[1] = (cuuint64_t)H * W * 4; }
    box[0] = bx; box[1] = b1; box[2] = b2;
    if (variant == 3) { dims[0] = W; dims[1] = (cuuint64_t)NC * H; strides[0] = (cuuint64_t)W * 4; box[0] = bx; box[1] = b1 * b2; }
    CUresult rc = enc(&tm, (cfgbits & 4) ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, variant == 3 ? 2 : 3, feat, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      (cfgbits & 2) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      (cfgbits & 1) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int i = 0; i < 16; i++) printf("%016llx%c", (unsigned long long)tm.opaque[i], i % 4 == 3 ? '\n' : ' ');
    printf("order=%d box=(%d,%d,%d) variant=%d encode rc=%d\n", order, bx, b1, b2, variant, (int)rc);
    if (rc) return 0;
    const int n = bx * b1 * b2;
    float* out; CK(cudaMalloc(&out, n * 4));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4 + 2048));
    CUtensorMap* gtm; CK(cudaMalloc(&gtm, sizeof(tm))); CK(cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    const int cx = argc > 7 ? atoi(argv[7]) : 5;
    printf("x coordinate %d (start offset %d bytes)\n", cx, cx * 4);
    k<<<1, 128, n * 4 + 2048>>>(tm, gtm, feat, cx, 2, 3, n * 4, out, n, variant);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(n); CK(cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost));
    printf("first: %.1f %.1f %.1f | row1: %.1f | plane1: %.1f\n", o[0], o[1], o[2], o[bx], o[bx * b1]);
    return 0;
}

// Probe: one warp loads a {bx u32, rows, 1} box through a 3-D tensor map and writes it back (debug aid for k1_strips.cu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, uint32_t *out, int cx, int cy, int cz, int bytes, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 8192);
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (mode >= 1 && lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         smem_u32(smem)),
                     "l"(&tmap), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(bar))
                     : "memory");
    }
    if (mode >= 2) {
        asm volatile(
            "{\n.reg .pred p;\nPW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra PD;\nbra PW;\nPD:\n}\n" ::"r"(smem_u32(bar)),
            "r"(0)
            : "memory");
        for (int i = lane; i < bytes / 4; i += 32) out[i] = reinterpret_cast<uint32_t *>(smem)[i];
    }
    if (mode == 3) {
        uint32_t v = __dp2a_lo(out[lane], 0x0101u, 0u);
        v = __dp4a(v, 0x00d2f04eu, v);
        v = __funnelshift_l(v, v, 1);
        out[lane] = v;
    }
}
int main(int argc, char **argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 2, bx = argc > 2 ? atoi(argv[2]) : 192, rows = argc > 3 ? atoi(argv[3]) : 2;
    int cx = argc > 4 ? atoi(argv[4]) : -6, cy = argc > 5 ? atoi(argv[5]) : -1;
    const int w = 1920, h = 64, n = 2, pitch = w * 3;
    std::vector<uint8_t> host((size_t)n * h * pitch);
    for (size_t i = 0; i < host.size(); i++) host[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d; uint32_t *out;
    cudaMalloc(&d, host.size()); cudaMalloc(&out, 65536);
    cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fp=%p\n", cudaGetErrorString(e), (int)q, fp);
    CUtensorMap map;
    cuuint64_t gdim[3] = {(cuuint64_t)w * 3 / 4, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t gstride[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
    CUresult cr = ((EncodeTiledFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)cr);
    int bytes = bx * 4 * rows;
    probe<<<1, 32, 16384>>>(map, out, cx, cy, 1, bytes, mode);
    e = cudaDeviceSynchronize();
    printf("mode %d box %dx%d at (%d,%d,1): %s\n", mode, bx, rows, cx, cy, cudaGetErrorString(e));
    if (e == cudaSuccess && mode == 2) {
        std::vector<uint32_t> res(bytes / 4);
        cudaMemcpy(res.data(), out, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < rows; r++)
            for (int i = 0; i < bx; i++) {
                int gx = cx + i, gy = cy + r;
                uint32_t want = 0;
                if (gx >= 0 && gx < w * 3 / 4 && gy >= 0 && gy < h) memcpy(&want, &host[(size_t)1 * h * pitch + (size_t)gy * pitch + (size_t)gx * 4], 4);
                if (res[r * bx + i] != want) bad++;
            }
        printf("mismatches: %d of %d\n", bad, bytes / 4);
    }
    return 0;
}

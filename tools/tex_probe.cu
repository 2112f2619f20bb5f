// Probe for K2's sampling: does tex2Dgather work on a pitch-linear u8 texture on this device, in which order do the four
// texels come back, and how does one gather per bilinear sample compare with four byte loads on K2's access pattern
// (49 x 49 samples along a rotated square of side S per warp)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o aruco3_b200/csrc/build/tex_probe tools/tex_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__global__ void order_kernel(cudaTextureObject_t tex, int l, int t, uint32_t *out) {
    const uchar4 g = tex2Dgather<uchar4>(tex, (float)l + 1.0f, (float)t + 1.0f, 0);
    out[0] = g.x; out[1] = g.y; out[2] = g.z; out[3] = g.w;
}

template <int MODE>  // 0: four byte loads, 1: one gather
__global__ void __launch_bounds__(256) sample_kernel(const uint8_t *grey, const cudaTextureObject_t *tex, uint32_t w, uint32_t h, uint32_t nquads, float side,
                                                     uint32_t *sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    for (uint32_t q = blockIdx.x * 8 + warp; q < nquads; q += gridDim.x * 8) {
        const uint32_t frame = q % 256;
        const float ang = 0.1f + 0.37f * (float)(q % 17), cx = 200.0f + (float)((q * 97) % (w - 400)), cy = 200.0f + (float)((q * 57) % (h - 400));
        const float ca = cosf(ang) * side / 49.0f, sa = sinf(ang) * side / 49.0f;
        const uint8_t *g = grey + (size_t)frame * w * h;
        const cudaTextureObject_t t = MODE ? tex[frame] : 0;
        for (uint32_t i = lane; i < 2401; i += 32) {
            const float ox = (float)(i % 49) - 24.5f, oy = (float)(i / 49) - 24.5f;
            const float px = cx + ca * ox - sa * oy, py = cy + sa * ox + ca * oy;
            const uint32_t l = (uint32_t)px, tp = (uint32_t)py;
            if (MODE) {
                const uchar4 v = tex2Dgather<uchar4>(t, (float)l + 1.0f, (float)tp + 1.0f, 0);
                acc += v.x + 2u * v.y + 3u * v.z + 5u * v.w;
            } else {
                acc += g[(size_t)(tp + 1) * w + l] + 2u * g[(size_t)(tp + 1) * w + l + 1] + 3u * g[(size_t)tp * w + l + 1] + 5u * g[(size_t)tp * w + l];
            }
        }
    }
    if (acc == 0xdeadbeefu) sink[0] = acc;
    atomicAdd(&sink[1], acc);
}

int main() {
    const uint32_t w = 1920, h = 1080, n = 256;
    uint8_t *grey;
    CK(cudaMalloc(&grey, (size_t)n * w * h));
    std::vector<uint8_t> host((size_t)w * h);
    for (size_t i = 0; i < host.size(); i++) host[i] = (uint8_t)((i * 2654435761u) >> 13);
    for (uint32_t f = 0; f < n; f++) CK(cudaMemcpy(grey + (size_t)f * w * h, host.data(), host.size(), cudaMemcpyHostToDevice));
    std::vector<cudaTextureObject_t> tex(n);
    for (uint32_t f = 0; f < n; f++) {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = grey + (size_t)f * w * h;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<unsigned char>();
        rd.res.pitch2D.width = w; rd.res.pitch2D.height = h; rd.res.pitch2D.pitchInBytes = w;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        CK(cudaCreateTextureObject(&tex[f], &rd, &td, nullptr));
    }
    cudaTextureObject_t *dtex;
    CK(cudaMalloc(&dtex, n * sizeof(cudaTextureObject_t)));
    CK(cudaMemcpy(dtex, tex.data(), n * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice));
    uint32_t *out;
    CK(cudaMalloc(&out, 64));
    // order of the four texels
    const int l = 100, t = 50;
    order_kernel<<<1, 1>>>(tex[0], l, t, out);
    CK(cudaDeviceSynchronize());
    uint32_t ho[4];
    CK(cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost));
    const uint32_t tl = host[(size_t)t * w + l], tr = host[(size_t)t * w + l + 1], bl = host[(size_t)(t + 1) * w + l], br = host[(size_t)(t + 1) * w + l + 1];
    printf("texels: tl %u tr %u bl %u br %u ; gather x %u y %u z %u w %u\n", tl, tr, bl, br, ho[0], ho[1], ho[2], ho[3]);
    // timing
    for (float side : {48.0f, 140.0f}) {
        for (int mode = 0; mode < 2; mode++) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            CK(cudaMemset(out, 0, 64));
            for (int rep = 0; rep < 3; rep++) {
                if (rep == 1) cudaEventRecord(a);
                if (mode) sample_kernel<1><<<148 * 3, 256>>>(grey, dtex, w, h, 5600, side, out);
                else sample_kernel<0><<<148 * 3, 256>>>(grey, dtex, w, h, 5600, side, out);
            }
            cudaEventRecord(b);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            uint32_t hs[2];
            CK(cudaMemcpy(hs, out, 8, cudaMemcpyDeviceToHost));
            printf("side %.0f mode %s: %.4f ms per 5600 quads (checksum %u)\n", side, mode ? "gather" : "4 loads", ms / 2, hs[1]);
        }
    }
    return 0;
}

// Micro-benchmark: peak bilinear fetch rate of a 256x64 RGBA8 (unorm8 -> float4) texture on this GPU.
// Build/run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/tex_rate.cu -o /tmp/tex_rate && /tmp/tex_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <int COMPONENTS>
__global__ void k_tex(cudaTextureObject_t t, float *out, int iters)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    float u = (gid & 1023) * (1.f / 1024.f), v = (gid >> 10 & 63) * (1.f / 64.f);
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 c = tex2D<float4>(t, u + k * 0.0123f, v + k * 0.0071f);
            if (COMPONENTS == 4) acc += c.x + c.y + c.z + c.w;
            else if (COMPONENTS == 3) acc += c.x + c.y + c.z;
            else acc += c.x;
        }
        u += 0.00037f;
        v += 0.00011f;
    }
    if (acc == 123456.f) out[gid] = acc;
}

int main()
{
    std::vector<unsigned char> lut(256 * 64 * 4);
    for (size_t i = 0; i < lut.size(); ++i) lut[i] = (unsigned char)(i * 37);
    cudaArray_t arr;
    cudaChannelFormatDesc ch = cudaCreateChannelDesc<uchar4>();
    cudaMallocArray(&arr, &ch, 256, 64);
    cudaMemcpy2DToArray(arr, 0, 0, lut.data(), 1024, 1024, 64, cudaMemcpyHostToDevice);
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    cudaTextureObject_t tex;
    cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2000, blocks = 148 * 8, threads = 256;
    for (int comp : {4, 3, 1}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (comp == 4) k_tex<4><<<blocks, threads>>>(tex, out, iters);
            else if (comp == 3) k_tex<3><<<blocks, threads>>>(tex, out, iters);
            else k_tex<1><<<blocks, threads>>>(tex, out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fetches = (double)blocks * threads * iters * 8;
        printf("components used %d: %.1f Gfetch/s (%.2f fetches/clk/SM at 1.965 GHz)\n", comp, fetches / ms / 1e6,
               fetches / ms / 1e6 / 148 / 1.965);
    }
    return 0;
}

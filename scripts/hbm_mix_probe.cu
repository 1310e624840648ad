// What does HBM sustain for the READ : WRITE mix of each step kernel?  The roofline denominator (MEASURED_PEAKS.json) is a
// 1 : 1 copy; the fused f32 step reads 4 B and writes 8 B per cell, the fused f64 step reads 8 B and writes 12 B.  Plain
// streaming kernels with exactly those mixes (16-byte accesses, grid-stride, arrays far larger than L2), best launch shape of {2..32 blocks per SM} x {256, 512, 1024 threads} x {1, 2, 4 indices per thread}.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/hbm_mix_probe.cu -o scripts/hbm_mix_probe && scripts/hbm_mix_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int NR, int NW>
__global__ void k_mix(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n)
{
    // per index: NR float4 loads from NR planes, NW float4 stores to NW planes
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = in[i];
#pragma unroll
        for (int r = 1; r < NR; ++r) { const float4 b = in[i + r * n]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
#pragma unroll
        for (int w = 0; w < NW; ++w) out[i + w * n] = make_float4(a.x + w, a.y, a.z, a.w);
    }
}

// the same with U independent indices per thread and trip (more loads in flight per thread)
template <int NR, int NW, int U>
__global__ void k_mix_u(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            if (i < n) {
                a[u] = in[i];
#pragma unroll
                for (int r = 1; r < NR; ++r) { const float4 b = in[i + r * n]; a[u].x += b.x; a[u].y += b.y; a[u].z += b.z; a[u].w += b.w; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            if (i < n) {
#pragma unroll
                for (int w = 0; w < NW; ++w) out[i + w * n] = make_float4(a[u].x + w, a[u].y, a[u].z, a[u].w);
            }
        }
    }
}

template <int NR, int NW>
static void run(const char *name, float4 *in, float4 *out, size_t n)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    int best_cfg[3] = {0, 0, 0};
    const int bps[] = {2, 4, 8, 16, 32}, thr[] = {256, 512, 1024};
    for (int bi = 0; bi < 5; ++bi)
        for (int ti = 0; ti < 3; ++ti)
            for (int u = 1; u <= 4; u *= 2) {
                if (bps[bi] * thr[ti] > 2048 * 16) continue;
                for (int rep = 0; rep < 6; ++rep) {
                    cudaEventRecord(e0);
                    if (u == 1) k_mix<NR, NW><<<148 * bps[bi], thr[ti]>>>(in, out, n);
                    else if (u == 2) k_mix_u<NR, NW, 2><<<148 * bps[bi], thr[ti]>>>(in, out, n);
                    else k_mix_u<NR, NW, 4><<<148 * bps[bi], thr[ti]>>>(in, out, n);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep >= 1 && ms < best) { best = ms; best_cfg[0] = bps[bi]; best_cfg[1] = thr[ti]; best_cfg[2] = u; }
                }
            }
    const double bytes = (double)n * 16.0 * (NR + NW);
    printf("{\"mix\": \"%s\", \"read_planes\": %d, \"write_planes\": %d, \"GBps\": %.1f, \"ms\": %.4f, \"best\": \"%d blocks/SM x %d threads, unroll %d\"}\n",
           name, NR, NW, bytes / best / 1e6, best, best_cfg[0], best_cfg[1], best_cfg[2]);
}

int main()
{
    const size_t n = (size_t)32 << 20;  // 32 Mi float4 = 512 MiB per plane
    float4 *in, *out;
    cudaMalloc(&in, n * 16 * 2);
    cudaMalloc(&out, n * 16 * 3);
    cudaMemset(in, 0, n * 16 * 2);
    run<1, 1>("copy 1:1 (drainage only: read h, write h')", in, out, n);
    run<1, 2>("1:2 (fused f32 step: read h, write h' + RGBA8)", in, out, n);
    run<2, 3>("2:3 (fused f64 step: read 8 B, write 8 + 4 B)", in, out, n);
    run<2, 2>("2:2 (WGSL semantics: two fields in, two out)", in, out, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

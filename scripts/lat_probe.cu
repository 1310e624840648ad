// lat_probe.cu -- dependent-chain latencies of the instructions the resident kernel's pole reduction is built from.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double *out, long long *cyc, double seed, const float *gsrc)
{
    __shared__ float sm[1024];
    __shared__ unsigned long long bar;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = gsrc[i];
    __syncthreads();
    const int n = 256;
    long long t0, t1;
    double v = seed;
    // DADD chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) v += seed;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / n;
    // DMUL+DADD (DFMA) chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) v = fma(v, 1.0000001, seed);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / n;
    // double shuffle + DADD chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) v += __shfl_xor_sync(0xffffffffu, v, 1);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / n;
    // F2F f32->f64->f32 chain
    float f = (float)seed;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { double d = (double)f; d += seed; f = (float)d; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0) / n;
    // LDS dependent chain
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) idx = (int)sm[idx] & 1023;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0) / n;
    // FADD chain
    float g = (float)seed;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) g += f;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0) / n;
    // __syncthreads chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0) / n;
    // 32-bit shuffle + FADD
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) g += __shfl_xor_sync(0xffffffffu, g, 1);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0) / n;
    // mbarrier try_wait on a completed phase
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar))); }
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    __syncthreads();
    unsigned ok = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        unsigned p;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(p) : "r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
        ok += p;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = (t1 - t0) / n;
    out[threadIdx.x] = v + f + g + idx + ok;
}
int main()
{
    double *out; long long *cyc; float *src;
    cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&src, 4096);
    cudaMemset(src, 0, 4096);
    const char *names[] = {"DADD", "DFMA", "2xSHFL+DADD", "F2F.64.32+DADD+F2F.32.64", "LDS(+cvt+and)", "FADD", "BAR.SYNC", "SHFL+FADD", "TRYWAIT(done)"};
    for (int threads : {32, 512}) {
        probe<<<1, threads>>>(out, cyc, 1e-30, src);
        cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("threads=%d:", threads);
        for (int i = 0; i < 9; ++i) printf("  %s %lld", names[i], h[i]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

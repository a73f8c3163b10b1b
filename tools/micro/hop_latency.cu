// Microbenchmark: latency of a dependent shared-memory pointer chase, (a) through a generic pointer into dynamic
// shared memory (the compiler re-derives the shared window base with S2R SR_CgaCtaId inside the loop) and (b) with an
// explicit 32-bit shared address and ld.shared.  nvcc -arch=sm_100a -O3 -o hop_latency hop_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_generic(unsigned long long* out, int hops, int stride)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* E2 = reinterpret_cast<uint16_t*>(smem + 50160 + 4224 + 2 * 8448);
    uint16_t* seg = reinterpret_cast<uint16_t*>(smem + 80000);
    for (int i = threadIdx.x; i < 4224; i += blockDim.x) E2[i] = (uint16_t)((i + stride) % 4224);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int b = 0, nseg = 0;
        const long long t0 = clock64();
        for (int h = 0; h < hops; ++h) {
            if (lane == 0) seg[nseg] = (uint16_t)b;
            ++nseg;
            const unsigned e = E2[b];
            if (e == 0xFFFF) break;
            b = (int)e;
        }
        const long long t1 = clock64();
        if (lane == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = b + nseg; }
    }
}

__global__ void k_explicit(unsigned long long* out, int hops, int stride)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* E2 = reinterpret_cast<uint16_t*>(smem + 50160 + 4224 + 2 * 8448);
    uint16_t* seg = reinterpret_cast<uint16_t*>(smem + 80000);
    for (int i = threadIdx.x; i < 4224; i += blockDim.x) E2[i] = (uint16_t)((i + stride) % 4224);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const unsigned aE2 = (unsigned)__cvta_generic_to_shared(E2), aSeg = (unsigned)__cvta_generic_to_shared(seg);
        int b = 0, nseg = 0;
        const long long t0 = clock64();
        for (int h = 0; h < hops; ++h) {
            if (lane == 0) asm volatile("st.shared.u16 [%0], %1;" ::"r"(aSeg + 2 * nseg), "h"((unsigned short)b) : "memory");
            ++nseg;
            unsigned short e;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(e) : "r"(aE2 + 2 * b) : "memory");
            if (e == 0xFFFF) break;
            b = (int)e;
        }
        const long long t1 = clock64();
        if (lane == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = b + nseg; }
    }
}

__global__ void k_s2r(unsigned long long* out, int hops, int stride)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* E2 = reinterpret_cast<uint16_t*>(smem + 50160 + 4224 + 2 * 8448);
    for (int i = threadIdx.x; i < 4224; i += blockDim.x) E2[i] = (uint16_t)((i + stride) % 4224);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const unsigned aE2 = (unsigned)__cvta_generic_to_shared(E2);
        int b = 0, nseg = 0;
        const long long t0 = clock64();
        for (int h = 0; h < hops; ++h) {
            unsigned r;
            asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));      // an S2R on the dependent chain
            ++nseg;
            unsigned short e;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(e) : "r"(aE2 + 2 * b + (r << 24)) : "memory");
            if (e == 0xFFFF) break;
            b = (int)e;
        }
        const long long t1 = clock64();
        if (lane == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = b + nseg; }
    }
}

int main()
{
    unsigned long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 83328);
    cudaFuncSetAttribute(k_explicit, cudaFuncAttributeMaxDynamicSharedMemorySize, 83328);
    cudaFuncSetAttribute(k_s2r, cudaFuncAttributeMaxDynamicSharedMemorySize, 83328);
    for (int rep = 0; rep < 2; ++rep)
        for (int threads : {32, 512}) {
            unsigned long long h[2];
            k_generic<<<1, threads, 83328>>>(d, 1000, 257); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("generic  threads %4d: %.1f cycles/hop\n", threads, h[0] / 1000.0);
            k_explicit<<<1, threads, 83328>>>(d, 1000, 257); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("explicit threads %4d: %.1f cycles/hop\n", threads, h[0] / 1000.0);
            k_s2r<<<1, threads, 83328>>>(d, 1000, 257); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("with S2R threads %4d: %.1f cycles/hop\n", threads, h[0] / 1000.0);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

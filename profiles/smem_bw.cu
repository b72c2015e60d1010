// Shared-memory bandwidth microbenchmark (SURVEY.md 8d asks for the measured peak the shared-memory roofline is
// quoted against). Every thread streams 128-bit conflict-free loads out of a 64 KB shared buffer; a dependent XOR
// keeps the loads live. Reports aggregate GB/s for 32- and 128-bit accesses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/smem_bw profiles/smem_bw.cu && profiles/smem_bw
#include <cstdio>
#include <cuda_runtime.h>

#define SMEM_BYTES 65536
template <class T>
__global__ void __launch_bounds__(1024) k_smem(T* out, int iters) {
    extern __shared__ __align__(16) unsigned char raw[];
    T* buf = reinterpret_cast<T*>(raw);
    const int n = SMEM_BYTES / sizeof(T);
    const int U = n / 1024; // distinct addresses per thread and iteration: nothing for the compiler to merge
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T v = T();
        reinterpret_cast<unsigned*>(&v)[0] = i * 2654435761u;
        buf[i] = v;
    }
    __syncthreads();
    T acc = T();
    int idx = threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            T v = buf[(idx + u * 1024) & (n - 1)];
            if constexpr (sizeof(T) == 16) {
                acc.x ^= v.x;
                acc.y ^= v.y;
                acc.z ^= v.z;
                acc.w ^= v.w;
            }
            else {
                acc ^= v;
            }
        }
        idx = (idx + 32) & (n - 1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class T>
double run(int blocks_per_sm, int sms, int iters) {
    T* out;
    int blocks = blocks_per_sm * sms;
    cudaMalloc(&out, sizeof(T) * 1024 * blocks);
    cudaFuncSetAttribute(k_smem<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_smem<T><<<blocks, 1024, SMEM_BYTES>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_smem<T><<<blocks, 1024, SMEM_BYTES>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaFree(out);
    double bytes = (double)blocks * 1024 * iters * (SMEM_BYTES / sizeof(T) / 1024) * sizeof(T);
    return bytes / (ms * 1e-3) / 1e9;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double best32 = 0, best128 = 0;
    for (int bps = 1; bps <= 2; bps++) {
        double g32 = run<unsigned>(bps, sms, 4000), g128 = run<uint4>(bps, sms, 4000);
        if (g32 > best32) best32 = g32;
        if (g128 > best128) best128 = g128;
    }
    printf("{\"device\": \"%s\", \"sms\": %d, \"smem_gbs_32bit\": %.1f, \"smem_gbs\": %.1f, \"note\": \"conflict-free LDS.128 streaming, 1024 threads per block, 1-2 blocks per SM\"}\n",
           p.name, sms, best32, best128);
    return 0;
}

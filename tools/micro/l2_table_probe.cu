// Microbenchmark for a one-level partition + direct-address tables in L2:
// n items stream through; item i touches table slot (window(i) * W + hash(i) % W), where the window
// advances with i (as the first-level buckets would), W slots of 8 bytes = 2 MiB.  Three access kinds:
// gather (probe), atomicExch (build with duplicate detection), plain store (build without).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/l2probe tools/micro/l2_table_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int KIND>
__global__ void k(const u64 *__restrict__ in, u64 n, u64 *table, uint32_t wbits, u64 per_window, u64 *acc) {
    u64 cnt = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 8 * stride) {
        u64 x[8];
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = (i0 + u * stride < n) ? __ldcs(in + i0 + u * stride) : 0;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const u64 i = i0 + u * stride;
            if (i < n) {
                const u64 slot = ((i / per_window) << wbits) | (mix((uint32_t)x[u]) & ((1u << wbits) - 1));
                if (KIND == 0) cnt += table[slot] != 0;
                if (KIND == 1) cnt += atomicExch(&table[slot], ~x[u]) != 0;
                if (KIND == 2) table[slot] = ~x[u];
            }
        }
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(acc, cnt);
}
int main() {
    const u64 n = 128000000ull;
    const uint32_t wbits = 18;  // 2 MiB windows
    const u64 windows = 512, per_window = n / windows;
    u64 *in, *table, *acc;
    cudaMalloc(&in, n * 8); cudaMalloc(&table, (windows << wbits) * 8); cudaMalloc(&acc, 8);
    cudaMemset(in, 0x5a, n * 8);
    // fill `in` with distinct-ish words
    {
        u64 *h = (u64 *)malloc(1 << 20);
        for (int i = 0; i < (1 << 17); i++) h[i] = (u64)i * 0x9E3779B97F4A7C15ull;
        for (u64 off = 0; off < n * 8; off += 1 << 20) cudaMemcpy((char *)in + off, h, (n * 8 - off) < (1 << 20) ? (n * 8 - off) : (1 << 20), cudaMemcpyHostToDevice);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[3] = {"gather (probe)", "atomicExch (build)", "plain store (build)"};
    for (int grid : {148 * 8, 148 * 16}) {
        for (int kind = 0; kind < 3; kind++) {
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaMemset(table, kind == 0 ? 0xff : 0, (windows << wbits) * 8);
                cudaMemset(acc, 0, 8);
                cudaEventRecord(e0);
                if (kind == 0) k<0><<<grid, 256>>>(in, n, table, wbits, per_window, acc);
                if (kind == 1) k<1><<<grid, 256>>>(in, n, table, wbits, per_window, acc);
                if (kind == 2) k<2><<<grid, 256>>>(in, n, table, wbits, per_window, acc);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("grid %5d  %-22s %.3f ms per 128 M items (%.0f G items/s)\n", grid, names[kind], best, n / best / 1e6);
        }
    }
    // memset of the 1 GiB of tables
    float ms; cudaEventRecord(e0); cudaMemsetAsync(table, 0, (windows << wbits) * 8); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("memset of the tables (1 GiB): %.3f ms\n", ms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

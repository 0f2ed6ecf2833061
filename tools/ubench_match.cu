// Micro-benchmark: issue cost of MATCH.ANY (__match_any_sync) against the eight-ballot form and the shared-memory atomicOr
// form of "which lanes of my warp hold the same 8-bit digit" (the ranking step of csrc/dense_sort.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_match tools/ubench_match.cu && /tmp/ubench_match
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 2) k(const uint32_t *digits, int iters, unsigned long long *out, long long *cyc)
{
    __shared__ uint32_t mm[16][256];
    const unsigned lane = threadIdx.x & 31u, wp = threadIdx.x >> 5;
    for (int d = lane; d < 256; d += 32) mm[wp][d] = 0u;
    __syncwarp();
    uint32_t my[8];
    for (int j = 0; j < 8; ++j) my[j] = digits[(blockIdx.x * 8 + j) * 512 + threadIdx.x] & 255u;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t d = (my[j] + it) & 255u;
            uint32_t peers;
            if (MODE == 0) {
                peers = __match_any_sync(0xFFFFFFFFu, d);
            } else if (MODE == 1) {
                peers = 0xFFFFFFFFu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = (d >> b) & 1u;
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, bit);
                    peers &= bit ? m : ~m;
                }
            } else {
                atomicOr(&mm[wp][d], 1u << lane);
                __syncwarp();
                peers = mm[wp][d];
                __syncwarp();
                if ((peers & ((1u << lane) - 1u)) == 0u) mm[wp][d] = 0u;
                __syncwarp();
            }
            acc += __popc(peers);
        }
    }
    long long t1 = clock64();
    atomicAdd(out, (unsigned long long)acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    const int grid = 296, iters = 500;
    uint32_t *h = (uint32_t *)malloc(grid * 8 * 512 * 4), *d;
    srand(3);
    for (int i = 0; i < grid * 8 * 512; ++i) h[i] = rand();
    unsigned long long *out; long long *cyc, hc[296];
    cudaMalloc(&d, grid * 8 * 512 * 4); cudaMalloc(&out, 8); cudaMalloc(&cyc, grid * 8);
    cudaMemcpy(d, h, grid * 8 * 512 * 4, cudaMemcpyHostToDevice);
    const char *names[] = {"__match_any_sync", "8 ballots", "atomicOr + read back + clear"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<grid, 512>>>(d, iters, out, cyc);
            if (mode == 1) k<1><<<grid, 512>>>(d, iters, out, cyc);
            if (mode == 2) k<2><<<grid, 512>>>(d, iters, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += hc[i]; avg /= grid;
        // 32 warps per SM = 8 per scheduler; each warp executes iters*8 items
        printf("%-32s %8.1f cycles per warp-item per scheduler (32 warps/SM, random 8-bit digits)  err=%s\n", names[mode], avg / (iters * 8.0 * 8.0), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

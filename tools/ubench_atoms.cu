// Micro-benchmark: cost of shared-memory atomics per warp instruction under different address patterns.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_atoms tools/ubench_atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: atomicAdd(.,1) 1: atomicAdd(., v!=1) 2: red.shared.add 3: non-atomic ld/add/st 4: atomicAdd u64? 
__global__ void __launch_bounds__(1024, 1) k(const uint32_t *idx, int iters, int nbins, unsigned long long *out, long long *cyc)
{
    extern __shared__ uint32_t h[];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t my[16];
    for (int j = 0; j < 16; ++j) my[j] = idx[(blockIdx.x * 16 + j) * 1024 + threadIdx.x];
    const uint32_t hb = smem_u32(h);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t b = my[j];
            if (MODE == 0) atomicAdd(&h[b], 1u);
            else if (MODE == 1) atomicAdd(&h[b], (uint32_t)(it + 2));
            else if (MODE == 2) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hb + 4 * b), "r"(it + 2) : "memory");
            else if (MODE == 3) { uint32_t v; asm volatile("ld.shared.u32 %0,[%1];" : "=r"(v) : "r"(hb + 4 * b) : "memory"); asm volatile("st.shared.u32 [%0], %1;" ::"r"(hb + 4 * b), "r"(v + 1) : "memory"); }
            else if (MODE == 4) { uint32_t v; asm volatile("ld.shared.u32 %0,[%1];" : "=r"(v) : "r"(hb + 4 * b) : "memory"); my[j] = (my[j] + (v & 0)) ; }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s += h[i];
    atomicAdd(out, s + my[3]);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    const int grid = 148, nb = 8192, iters = 200;
    uint32_t *hidx = (uint32_t *)malloc(grid * 16 * 1024 * 4), *d;
    unsigned long long *out; long long *cyc, hc[148];
    cudaMalloc(&d, grid * 16 * 1024 * 4); cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
    const char *names[] = {"lane-distinct banks", "random uniform", "gaussian sd=600", "all lanes same addr", "2 addrs per warp", "stride-32 (same bank, 32 addrs)"};
    for (int pat = 0; pat < 6; ++pat) {
        srand(1);
        for (int i = 0; i < grid * 16 * 1024; ++i) {
            int lane = i & 31; uint32_t v = 0;
            switch (pat) {
            case 0: v = ((rand() % (nb / 32)) * 32 + lane); break;
            case 1: v = rand() % nb; break;
            case 2: { double s = 0; for (int q = 0; q < 12; ++q) s += rand() / (double)RAND_MAX; v = (uint32_t)(nb / 2 + (s - 6.0) * 600); if (v >= (uint32_t)nb) v = nb - 1; } break;
            case 3: v = (i >> 5) % nb; break;
            case 4: v = ((i >> 5) * 2 + (lane & 1)) % nb; break;
            case 5: v = (lane * 32 + ((i >> 5) & 31)) % nb; break;
            }
            hidx[i] = v;
        }
        cudaMemcpy(d, hidx, grid * 16 * 1024 * 4, cudaMemcpyHostToDevice);
        printf("%-34s", names[pat]);
        for (int mode = 0; mode < 5; ++mode) {
            cudaMemset(out, 0, 8);
            cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, nb * 4);
            if (mode == 0) k<0><<<grid, 1024, nb * 4>>>(d, iters, nb, out, cyc);
            if (mode == 1) k<1><<<grid, 1024, nb * 4>>>(d, iters, nb, out, cyc);
            if (mode == 2) k<2><<<grid, 1024, nb * 4>>>(d, iters, nb, out, cyc);
            if (mode == 3) k<3><<<grid, 1024, nb * 4>>>(d, iters, nb, out, cyc);
            if (mode == 4) k<4><<<grid, 1024, nb * 4>>>(d, iters, nb, out, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hc, cyc, 148 * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < grid; ++i) avg += hc[i]; avg /= grid;
            // cycles per warp-instruction per SM: 32 warps * iters * 16 instr
            printf("  m%d %6.2f", mode, avg / (32.0 * iters * 16));
        }
        printf("   (SM cycles per warp-instr; m0 atomicAdd 1, m1 atomicAdd v, m2 red, m3 ld+st, m4 ld)\n");
    }
    return 0;
}

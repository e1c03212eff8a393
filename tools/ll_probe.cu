// Micro-benchmark: cost of one gather round of the tagged-partials protocol (512 threads, 10 x 16-B loads
// per thread over a 50 KB L2-resident array) for different load flavours.
#include <cstdio>
#include <cuda_runtime.h>
template <int FL>
__device__ __forceinline__ uint4 ld(const uint4 *p) {
    uint4 q;
    if (FL == 0) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    if (FL == 1) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    if (FL == 2) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    if (FL == 3) asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    return q;
}
template <int FL>
__global__ void round_cost(const uint4 *part, int nb, long long *cyc, unsigned *sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    long long t0 = clock64();
    unsigned acc = 0;
    for (int rep = 0; rep < 4; ++rep) {
        uint4 q[10];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            int b = r * 32 + lane; if (b >= nb) b = nb - 1;
            q[2 * r] = ld<FL>(&part[(size_t)warp * nb + b]);
            q[2 * r + 1] = ld<FL>(&part[(size_t)((warp + 16) % 21) * nb + b]);
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) acc += q[i].x + q[i].y + q[i].z + q[i].w;
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 4;
    sink[threadIdx.x] = acc;
}
int main() {
    const int nb = 148; uint4 *d; long long *c; unsigned *s;
    cudaMalloc(&d, 21 * nb * 16); cudaMemset(d, 1, 21 * nb * 16); cudaMalloc(&c, 8); cudaMalloc(&s, 4096);
    const char *names[] = {"ld.relaxed.gpu.v4", "ld.volatile.v4", "ld.cg.v4", "ld.ca.v4"};
    long long hc;
    for (int rep = 0; rep < 2; ++rep) {
        round_cost<0><<<1, 512>>>(d, nb, c, s); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("%-20s %lld cycles/round\n", names[0], hc);
        round_cost<1><<<1, 512>>>(d, nb, c, s); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("%-20s %lld cycles/round\n", names[1], hc);
        round_cost<2><<<1, 512>>>(d, nb, c, s); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("%-20s %lld cycles/round\n", names[2], hc);
        round_cost<3><<<1, 512>>>(d, nb, c, s); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("%-20s %lld cycles/round\n", names[3], hc);
    }
    return 0;
}

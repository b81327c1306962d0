// Dev micro-benchmark: what does ONE level of a barrier-separated shared-memory chain cost on this SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o level_floor level_floor.cu && ./level_floor
// Variants: consumer warps (1/4/8), barrier flavour (named bar.sync / __syncthreads / __syncwarp), with or without the
// extra per-level work of the real chain kernel (16-byte record read from shared memory, 16-byte global store).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>  // 0: named barrier (register count), 1: __syncthreads, 2: __syncwarp (1 warp), 3: named barrier, immediate count 128
__global__ void chain(int* out, uint4* gout, int iters, int nt, int extra, unsigned seed) {
    __shared__ int ctr[4096];
    __shared__ uint4 ring[1024];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) ctr[i] = i & 7;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) ring[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    if ((int)threadIdx.x >= nt) return;
    unsigned s = seed + threadIdx.x * 2654435761u;
    int a0 = (s >> 3) & 4095, a1 = (s >> 9) & 4095, a2 = (s >> 15) & 4095;
    int acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        volatile int* c = ctr;
        const int x0 = c[a0], x1 = c[a1], x2 = c[a2];
        uint4 r = make_uint4(0, 0, 0, 0);
        if (extra & 1) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(&ring[(it * 128 + threadIdx.x) & 1023]);
            asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(sa));
        }
        const bool L10 = x1 < x0, L20 = x2 < x0, L21 = x2 < x1;
        const bool is2 = L10 ? L21 : L20, is1 = L10 && !L21;
        const int oA = is2 ? a2 : (is1 ? a1 : a0), vA = is2 ? x2 : (is1 ? x1 : x0);
        c[oA] = vA + 1;
        if (extra & 2) gout[(size_t)it * nt + threadIdx.x] = make_uint4(oA, r.x, r.y, r.z);
        if (extra & 4) ring[(it * 128 + threadIdx.x + 512) & 1023] = make_uint4(oA, r.x, r.y, r.z);
        acc += oA;
        // next level's addresses (data-dependent so nothing is hoisted)
        a0 = (a0 * 5 + 1 + (r.w & 1)) & 4095; a1 = (a1 * 5 + 3) & 4095; a2 = (a2 * 5 + 7) & 4095;
        if (MODE == 0) asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory");
        else if (MODE == 1) __syncthreads();
        else if (MODE == 3) asm volatile("bar.sync 1, 128;" ::: "memory");
        else __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = acc; out[1] = (int)((t1 - t0) / iters); }
}

int main() {
    int* d; uint4* g;
    cudaMalloc(&d, 64); cudaMalloc(&g, (size_t)20000 * 256 * 16);
    const int iters = 20000;
    int h[2];
    for (int extra : {0, 1, 2, 3, 4, 5}) {   // bit 0: ring read, bit 1: global 16 B store, bit 2: shared 16 B store instead
        for (int nt : {32, 128, 256}) {
            chain<0><<<1, nt>>>(d, g, iters, nt, extra, 1u); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            printf("extra=%d nt=%3d named-bar(reg) %4d", extra, nt, h[1]);
            chain<1><<<1, nt>>>(d, g, iters, nt, extra, 1u); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            printf("  syncthreads %4d", h[1]);
            if (nt == 128) { chain<3><<<1, nt>>>(d, g, iters, nt, extra, 1u); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); printf("  named-bar(imm) %4d", h[1]); }
            if (nt == 32) { chain<2><<<1, 32>>>(d, g, iters, 32, extra, 1u); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); printf("  syncwarp %4d", h[1]); }
            printf("  cycles/level\n");
        }
    }
    // a fifth warp that never joins the barrier (like the TMA producer warp) next to 4 consumer warps
    chain<0><<<1, 160>>>(d, g, iters, 128, 3, 1u); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("extra=3 nt=128 (+1 idle warp launched) named-bar(reg) %4d cycles/level\n", h[1]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// ubench_pingpong.cu -- one-way latency of a {value, tag} slot hand-off between two CTAs through L2 (the primitive of the
// persistent kernel's grid-wide exchange), for np concurrent pairs (CTA i <-> CTA i + np), and the same between two CTAs of
// one cluster through distributed shared memory.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void st_vol(uint2* p, unsigned x, unsigned y) { asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint2 ld_vol(const uint2* p) { uint2 v; asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rlx(uint2* p, unsigned x, unsigned y) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint2 ld_rlx(const uint2* p) { uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }

// mode 0: volatile, mode 1: relaxed.gpu; nlanes threads of warp 0 each run their own slot (nlanes slots per pair, 128 B apart
// when spread != 0, adjacent 8-byte slots otherwise)
__global__ void k_pp(uint2* slots, int np, int iters, int mode, int nlanes, int spread, long long* out, unsigned* smid)
{
    const int pair = blockIdx.x % np, side = blockIdx.x / np;
    if (threadIdx.x == 0) { unsigned id; asm("mov.u32 %0, %smid;" : "=r"(id)); smid[blockIdx.x] = id; }
    if (threadIdx.x >= nlanes) return;
    const int stride = spread ? 16 : 1;
    uint2* mine = slots + ((size_t)(2 * pair + side) * 32 + threadIdx.x) * stride;       // I write here
    uint2* theirs = slots + ((size_t)(2 * pair + (side ^ 1)) * 32 + threadIdx.x) * stride; // I poll here
    long long t0 = clock64();
    for (int k = 1; k <= iters; k++) {
        if (side == 0) {
            if (mode == 0) st_vol(mine, k, k); else st_rlx(mine, k, k);
            uint2 v;
            do { v = mode == 0 ? ld_vol(theirs) : ld_rlx(theirs); } while (v.y != (unsigned)k);
        } else {
            uint2 v;
            do { v = mode == 0 ? ld_vol(theirs) : ld_rlx(theirs); } while (v.y != (unsigned)k);
            if (mode == 0) st_vol(mine, k, k); else st_rlx(mine, k, k);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// the same hand-off between the two CTAs of a cluster through DSMEM
__global__ void __cluster_dims__(2, 1, 1) k_pp_dsmem(int iters, long long* out)
{
    __shared__ unsigned long long box;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    if (threadIdx.x == 0) box = 0;
    cl.sync();
    if (threadIdx.x == 0) {
        unsigned long long* peer = cl.map_shared_rank(&box, rank ^ 1);
        volatile unsigned long long* me = &box;
        long long t0 = clock64();
        for (int k = 1; k <= iters; k++) {
            if (rank == 0) { *(volatile unsigned long long*)peer = k; while (*me != (unsigned long long)k) {} }
            else { while (*me != (unsigned long long)k) {} *(volatile unsigned long long*)peer = k; }
        }
        out[blockIdx.x] = clock64() - t0;
    }
    cl.sync();
}

int main()
{
    const int iters = 2000;
    uint2* slots; long long* out; unsigned* smid;
    CK(cudaMalloc(&slots, 1 << 22)); CK(cudaMalloc(&out, 4096 * 8)); CK(cudaMalloc(&smid, 4096 * 4));
    for (int mode = 0; mode < 2; mode++)
        for (int np : {1, 8, 74})
            for (int nl : {1, 32})
                for (int spread : {0, 1}) {
                    if (nl == 1 && spread) continue;
                    CK(cudaMemset(slots, 0, 1 << 22));
                    void* args[] = {&slots, (void*)&np, (void*)&iters, &mode, &nl, &spread, &out, &smid};
                    CK(cudaLaunchCooperativeKernel((void*)k_pp, dim3(2 * np), dim3(32), args, 0, 0));
                    CK(cudaDeviceSynchronize());
                    long long h[256]; unsigned sm[256];
                    CK(cudaMemcpy(h, out, 2 * np * 8, cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(sm, smid, 2 * np * 4, cudaMemcpyDeviceToHost));
                    double mn = 1e30, mx = 0, av = 0;
                    for (int i = 0; i < np; i++) { double v = (double)h[i] / iters / 2; av += v / np; if (v < mn) mn = v; if (v > mx) mx = v; }
                    printf("%s pairs %3d lanes %2d %s: one-way hand-off cycles avg %.0f min %.0f max %.0f (pair 0: SM %u <-> SM %u)\n", mode ? "relaxed.gpu" : "volatile   ", np, nl,
                           spread ? "128B-spread" : "adjacent   ", av, mn, mx, sm[0], sm[np]);
                }
    k_pp_dsmem<<<2, 32>>>(iters, out);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
    printf("DSMEM (cluster of 2): one-way hand-off cycles %.0f\n", (double)h[0] / iters / 2);
    return 0;
}

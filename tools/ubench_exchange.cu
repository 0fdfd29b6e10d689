// ubench_exchange.cu -- the grid-wide sum of the persistent kernel (eh_epoch_kernel.cuh) in isolation:
// G CTAs, each publishes an NPART-float partial per step, every CTA needs all NPART totals.  Measures cycles per
// step of the exchange alone (a calibrated busy loop stands in for the compute phase), for several variants:
//   mode 0  reduce-scatter + all-gather through L2, self-validating {value, tag} slots, ld/st.volatile (the kernel's scheme)
//   mode 1  same, partial published with coalesced 16-byte stores (identity element order)
//   mode 2  same as 1 with .relaxed.gpu accesses instead of .volatile
//   mode 3  same as 1, owners = the FIRST NSL CTAs' service warps but totals pushed into per-CTA... (unused)
//   mode 4  counters: partial stored plainly, __threadfence + atomicAdd on a counter; owners poll the counter, totals likewise
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_exchange ubench_exchange.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void st_vol_v2(uint2* p, unsigned x, unsigned y) { asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void st_vol_v4(uint2* p, unsigned a, unsigned b, unsigned c, unsigned d) { asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory"); }
__device__ __forceinline__ uint4 ld_vol_v4(const uint2* p) { uint4 v; asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rlx_v4(uint2* p, unsigned a, unsigned b, unsigned c, unsigned d) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory"); }
__device__ __forceinline__ uint4 ld_rlx_v4(const uint2* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_cg_v4(uint2* p, unsigned a, unsigned b, unsigned c, unsigned d) { asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory"); }
__device__ __forceinline__ uint4 ld_cg_v4(const uint2* p) { uint4 v; asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

struct Args {
    uint2* part;      // [2][G][NP]
    uint2* tot;       // [2][NP]
    unsigned* cnt;    // [4] counters (mode 4)
    float* plain;     // [2][G][NP] plain floats (mode 4)
    float* ptot;      // [2][NP]
    int NP, nsteps, mode, compute_cycles, jitter, maxown;
    long long* out;   // [G][8] accumulated phase cycles
    unsigned long long* gt;  // [nsteps][G][4] global-timer stamps (optional)
    float* sink;
};

template <int MODE>
__global__ void __launch_bounds__(480, 1) k_x(const Args a)
{
    extern __shared__ float red[];
    const int G = gridDim.x, bid = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool service = warp == nwarps - 1;
    const int NP = a.NP, NSL = NP / 4;
    const int PG = 16, EPP = 2;
    const int nown = (NSL - bid + G - 1) / G;
    const int npasses = (2 * nown + EPP - 1) / EPP;
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    float fake = threadIdx.x;
    for (int s = 0; s < a.nsteps; s++) {
        const int par = s & 1;
        const unsigned tag = (unsigned)s + 1u;
        uint2* part = a.part + (size_t)par * (MODE == 5 ? (size_t)G * a.maxown * G * 4 : (size_t)G * NP);
        uint2* tot = a.tot + (size_t)par * (MODE == 5 ? (size_t)G * NP : (size_t)NP);
        long long t0 = clock64();
        // "compute": busy loop (compute warps only), a little longer on some CTAs
        if (!service) {
            const long long until = t0 + a.compute_cycles + ((bid * 37 + s * 11) % 16) * a.jitter / 16;
            while (clock64() < until) fake = fake * 1.0001f + 0.5f;
        }
        long long t1 = clock64();
        __syncthreads();
        long long t2 = clock64();
        if (a.gt && threadIdx.x == 0) a.gt[((size_t)s * G + bid) * 4 + 0] = gtime();
        // A
        if (MODE == 0) {
            for (int q = threadIdx.x; q < NP; q += blockDim.x) {
                const int NB = NP / 16;
                const int p = q < NB * 16 ? (q % NB) * 16 + q / NB : q;
                st_vol_v2(part + (size_t)bid * NP + p, __float_as_uint(fake + q), tag);
            }
        } else if (MODE == 1) {
            for (int q = threadIdx.x; q < NP / 2; q += blockDim.x)
                st_vol_v4(part + (size_t)bid * NP + 2 * q, __float_as_uint(fake + q), tag, __float_as_uint(fake - q), tag);
        } else if (MODE == 2) {
            for (int q = threadIdx.x; q < NP / 2; q += blockDim.x)
                st_rlx_v4(part + (size_t)bid * NP + 2 * q, __float_as_uint(fake + q), tag, __float_as_uint(fake - q), tag);
        } else if (MODE == 6 || MODE == 7) {
            for (int q = threadIdx.x; q < NP / 2; q += blockDim.x)
                st_cg_v4(part + (size_t)bid * NP + 2 * q, __float_as_uint(fake + q), tag, __float_as_uint(fake - q), tag);
        } else if (MODE == 5) {
            // push by owner: slice j = q / 2 goes to owner j % G, row j / G of its inbox: [row][peer][4 slots]
            for (int q = threadIdx.x; q < NP / 2; q += blockDim.x) {
                const int j = q >> 1, half = q & 1, owner = j % G, row = j / G;
                uint2* dst = part + ((size_t)(owner * a.maxown + row) * G + bid) * 4 + half * 2;
                st_vol_v4(dst, __float_as_uint(fake + q), tag, __float_as_uint(fake - q), tag);
            }
        } else if (MODE == 4) {
            float* pp = a.plain + ((size_t)par * G + bid) * NP;
            for (int q = threadIdx.x; q < NP; q += blockDim.x) pp[q] = fake + q;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(a.cnt + par, 1u);
        }
        long long t3 = clock64();
        if (a.gt && threadIdx.x == 0) a.gt[((size_t)s * G + bid) * 4 + 1] = gtime();
        // B
        if (MODE == 7) {
            __shared__ float2 sb[512];
            const int t = threadIdx.x;
            float2 v = make_float2(0.f, 0.f);
            if (nown > 0 && t < 2 * G) {
                const uint2* src = part + (size_t)(t >> 1) * NP + bid * 4 + (t & 1) * 2;
                uint4 x = ld_cg_v4(src);
                while (x.y != tag || x.w != tag) x = ld_cg_v4(src);
                v = make_float2(__uint_as_float(x.x), __uint_as_float(x.z));
            }
            // lanes with equal parity hold the same element pair: butterfly over xor 2..16, then warps through smem
            for (int o = 16; o > 1; o >>= 1) { v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o); }
            if (lane < 2) sb[warp * 2 + lane] = v;
            __syncthreads();
            if (nown > 0 && t < 2) {
                float2 sum = make_float2(0.f, 0.f);
                for (int w = 0; w < nwarps; w++) { sum.x += sb[w * 2 + t].x; sum.y += sb[w * 2 + t].y; }
                st_cg_v4(tot + bid * 4 + t * 2, __float_as_uint(sum.x), tag, __float_as_uint(sum.y), tag);
            }
        } else if (MODE == 5) {
            if (service) {
                for (int row = 0; row < nown; row++) {
                    const int j = bid + row * G;
                    const uint2* src = part + ((size_t)(bid * a.maxown + row) * G) * 4;   // [G][4] slots, contiguous
                    const int half = lane & 1;
                    float acc0 = 0.f, acc1 = 0.f;
                    constexpr int U = 10;
                    for (int c0 = lane >> 1; c0 < G; c0 += 16 * U) {
                        uint4 t[U];
#pragma unroll
                        for (int u = 0; u < U; u++)
                            if (c0 + 16 * u < G) t[u] = ld_vol_v4(src + (size_t)(c0 + 16 * u) * 4 + half * 2);
#pragma unroll
                        for (int u = 0; u < U; u++)
                            if (c0 + 16 * u < G) {
                                while (t[u].y != tag || t[u].w != tag) t[u] = ld_vol_v4(src + (size_t)(c0 + 16 * u) * 4 + half * 2);
                                acc0 += __uint_as_float(t[u].x);
                                acc1 += __uint_as_float(t[u].z);
                            }
                    }
                    for (int o = 16; o > 1; o >>= 1) {
                        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
                        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
                    }
                    // push the totals of this slice into every CTA's private inbox
                    for (int c = lane >> 1; c < G; c += 16)
                        st_vol_v4(tot + (size_t)c * NP + j * 4 + half * 2, __float_as_uint(acc0), tag, __float_as_uint(acc1), tag);
                }
            }
        } else if (MODE != 4) {
            for (int pass = nwarps - 1 - warp; pass < npasses; pass += nwarps) {
                const int sub = lane & (PG - 1);
                const int op = pass * EPP + (lane >> 4);
                const int slice = bid + (op >> 1) * G;
                const int ge = slice * 4 + (op & 1) * 2;
                const bool act = slice < NSL;
                float acc0 = 0.f, acc1 = 0.f;
                if (act) {
                    constexpr int U = 10;
                    const uint2* src = part + ge;
                    for (int cb = sub; cb < G; cb += PG * U) {
                        uint4 t[U];
#pragma unroll
                        for (int u = 0; u < U; u++)
                            if (cb + u * PG < G) t[u] = (MODE == 2 ? ld_rlx_v4(src + (size_t)(cb + u * PG) * NP) : MODE == 6 ? ld_cg_v4(src + (size_t)(cb + u * PG) * NP) : ld_vol_v4(src + (size_t)(cb + u * PG) * NP));
#pragma unroll
                        for (int u = 0; u < U; u++)
                            if (cb + u * PG < G) {
                                while (t[u].y != tag || t[u].w != tag)
                                    t[u] = (MODE == 2 ? ld_rlx_v4(src + (size_t)(cb + u * PG) * NP) : MODE == 6 ? ld_cg_v4(src + (size_t)(cb + u * PG) * NP) : ld_vol_v4(src + (size_t)(cb + u * PG) * NP));
                                acc0 += __uint_as_float(t[u].x);
                                acc1 += __uint_as_float(t[u].z);
                            }
                    }
                }
                for (int o = PG >> 1; o > 0; o >>= 1) {
                    acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
                    acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
                }
                if (act && sub == 0) {
                    if (MODE == 2) st_rlx_v4(tot + ge, __float_as_uint(acc0), tag, __float_as_uint(acc1), tag);
                    else if (MODE == 6) st_cg_v4(tot + ge, __float_as_uint(acc0), tag, __float_as_uint(acc1), tag);
                    else st_vol_v4(tot + ge, __float_as_uint(acc0), tag, __float_as_uint(acc1), tag);
                }
            }
        } else {
            // counters: wait until all G partials are complete, then plain loads
            if (service) {
                const unsigned want = (unsigned)G * (unsigned)(s / 2 + 1);
                if (lane == 0) { while (*(volatile unsigned*)(a.cnt + par) < want) {} }
                __syncwarp();
                __threadfence();
                for (int pass = 0; pass < npasses; pass++) {
                    const int sub = lane & (PG - 1);
                    const int op = pass * EPP + (lane >> 4);
                    const int slice = bid + (op >> 1) * G;
                    const int ge = slice * 4 + (op & 1) * 2;
                    const bool act = slice < NSL;
                    float acc0 = 0.f, acc1 = 0.f;
                    if (act)
                        for (int c = sub; c < G; c += PG) {
                            const float2 v = __ldcg(reinterpret_cast<const float2*>(a.plain + ((size_t)par * G + c) * NP + ge));
                            acc0 += v.x; acc1 += v.y;
                        }
                    for (int o = PG >> 1; o > 0; o >>= 1) {
                        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
                        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
                    }
                    if (act && sub == 0) *reinterpret_cast<float2*>(a.ptot + (size_t)par * NP + ge) = make_float2(acc0, acc1);
                }
                __threadfence();
                __syncwarp();
                if (lane == 0 && nown > 0) atomicAdd(a.cnt + 2 + par, 1u);
            }
        }
        long long t4 = clock64();
        if (a.gt && service && lane == 0) a.gt[((size_t)s * G + bid) * 4 + 2] = gtime();
        // C
        if (service) {
            if (MODE == 5) {
                constexpr int U = 8;
                const uint2* mine = tot + (size_t)bid * NP;
                for (int k0 = lane; k0 < NP / 2; k0 += 32 * U) {
                    uint4 t[U];
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (k0 + 32 * u < NP / 2) t[u] = ld_vol_v4(mine + 2 * (k0 + 32 * u));
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (k0 + 32 * u < NP / 2) {
                            while (t[u].y != tag || t[u].w != tag) t[u] = ld_vol_v4(mine + 2 * (k0 + 32 * u));
                            red[2 * (k0 + 32 * u)] = __uint_as_float(t[u].x);
                            red[2 * (k0 + 32 * u) + 1] = __uint_as_float(t[u].z);
                        }
                }
            } else if (MODE != 4) {
                constexpr int U = 8;
                for (int k0 = lane; k0 < NP / 2; k0 += 32 * U) {
                    uint4 t[U];
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (k0 + 32 * u < NP / 2) t[u] = (MODE == 2 ? ld_rlx_v4(tot + 2 * (k0 + 32 * u)) : (MODE == 6 || MODE == 7) ? ld_cg_v4(tot + 2 * (k0 + 32 * u)) : ld_vol_v4(tot + 2 * (k0 + 32 * u)));
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (k0 + 32 * u < NP / 2) {
                            while (t[u].y != tag || t[u].w != tag) t[u] = (MODE == 2 ? ld_rlx_v4(tot + 2 * (k0 + 32 * u)) : (MODE == 6 || MODE == 7) ? ld_cg_v4(tot + 2 * (k0 + 32 * u)) : ld_vol_v4(tot + 2 * (k0 + 32 * u)));
                            red[2 * (k0 + 32 * u)] = __uint_as_float(t[u].x);
                            red[2 * (k0 + 32 * u) + 1] = __uint_as_float(t[u].z);
                        }
                }
            } else {
                const int nowners = NSL < G ? NSL : G;
                const unsigned want = (unsigned)nowners * (unsigned)(s / 2 + 1);
                if (lane == 0) { while (*(volatile unsigned*)(a.cnt + 2 + par) < want) {} }
                __syncwarp();
                __threadfence();
                for (int k = lane; k < NP; k += 32) red[k] = __ldcg(a.ptot + (size_t)par * NP + k);
            }
        }
        if (a.gt && service && lane == 0) a.gt[((size_t)s * G + bid) * 4 + 3] = gtime();
        __syncthreads();
        long long t5 = clock64();
        fake += red[threadIdx.x % NP];
        if (threadIdx.x == 0) {
            acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; acc[5] += t5 - t0;
        }
    }
    if (threadIdx.x == 0)
        for (int i = 0; i < 6; i++) a.out[bid * 8 + i] = acc[i];
    if (fake == 12345.678f) a.sink[0] = fake;
}

int main(int argc, char** argv)
{
    int G = argc > 1 ? atoi(argv[1]) : 148, NP = 476, nsteps = 200;
    int compute = argc > 2 ? atoi(argv[2]) : 10000, jitter = argc > 3 ? atoi(argv[3]) : 1000;
    Args a{};
    const int maxown = (NP / 4 + G - 1) / G;
    a.maxown = maxown;
    const size_t part_bytes = (size_t)2 * G * (size_t)std::max(NP, maxown * G * 4) * 8, tot_bytes = (size_t)2 * G * NP * 8;
    CK(cudaMalloc(&a.part, part_bytes));
    CK(cudaMalloc(&a.tot, tot_bytes));
    CK(cudaMalloc(&a.cnt, 16));
    CK(cudaMalloc(&a.plain, (size_t)2 * G * NP * 4));
    CK(cudaMalloc(&a.ptot, (size_t)2 * NP * 4));
    CK(cudaMalloc(&a.out, (size_t)G * 8 * 8));
    CK(cudaMalloc(&a.sink, 4));
    unsigned long long* gtbuf; CK(cudaMalloc(&gtbuf, (size_t)nsteps * G * 4 * 8));
    const bool use_gt = argc > 4 && atoi(argv[4]);
    a.NP = NP; a.nsteps = nsteps; a.compute_cycles = compute; a.jitter = jitter;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode : {1, 6, 7}) {
        CK(cudaMemset(a.part, 0, part_bytes));
        CK(cudaMemset(a.tot, 0, tot_bytes));
        CK(cudaMemset(a.cnt, 0, 16));
        CK(cudaMemset(gtbuf, 0, (size_t)nsteps * G * 4 * 8));
        a.gt = use_gt ? gtbuf : nullptr;
        a.mode = mode;
        void* args[] = {(void*)&a};
        const void* fn = mode == 0 ? (const void*)k_x<0> : mode == 1 ? (const void*)k_x<1> : mode == 2 ? (const void*)k_x<2> : mode == 5 ? (const void*)k_x<5> : mode == 6 ? (const void*)k_x<6> : mode == 7 ? (const void*)k_x<7> : (const void*)k_x<4>;
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(480), args, NP * 4 + 64, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h((size_t)G * 8);
        CK(cudaMemcpy(h.data(), a.out, h.size() * 8, cudaMemcpyDeviceToHost));
        double m[6] = {0, 0, 0, 0, 0, 0};
        for (int b = 0; b < G; b++)
            for (int i = 0; i < 6; i++) m[i] += (double)h[(size_t)b * 8 + i] / G / nsteps;
        std::vector<unsigned long long> gt((size_t)nsteps * G * 4);
        CK(cudaMemcpy(gt.data(), gtbuf, gt.size() * 8, cudaMemcpyDeviceToHost));
        // global-timer view of a late step: spread of publish times, time from the LAST publish to the last B-done / C-done
        double lastA = 0, lastB = 0, lastC = 0, firstA = 0;
        int cntS = 0;
        for (int s = nsteps / 2; s < nsteps; s++) {
            unsigned long long a0 = ~0ull, a1 = 0, b1 = 0, c1 = 0;
            for (int b = 0; b < G; b++) {
                const unsigned long long* g = &gt[((size_t)s * G + b) * 4];
                if (g[1] < a0) a0 = g[1];
                if (g[1] > a1) a1 = g[1];
                if (g[2] > b1) b1 = g[2];
                if (g[3] > c1) c1 = g[3];
            }
            firstA += 0; lastA += (double)(a1 - a0); lastB += (double)(b1 - a1); lastC += (double)(c1 - a1);
            cntS++;
        }
        printf("mode %d  G %d: %.2f us/step (busy loop %d + jitter %d cycles)  cycles: compute %.0f  barrier %.0f  A %.0f  B %.0f  C+barrier %.0f  total %.0f | globaltimer ns: publish spread %.0f, last publish -> last B done %.0f, -> last C done %.0f\n",
               mode, G, 1e3 * ms / nsteps, compute, jitter, m[0], m[1], m[2], m[3], m[4], m[5], lastA / cntS, lastB / cntS, lastC / cntS);
    }
    return 0;
}

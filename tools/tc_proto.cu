// tc_proto.cu -- stand-alone check of the tcgen05 building blocks of the small-MLP tensor engine (eh_engine_tc.cuh):
// kind::tf32 MMAs with M = 128, N = 16, K = 8 on UNSWIZZLED shared-memory operands that the threads write themselves
// (plane layout [feature / 4][row][4 floats]), 3xTF32 splitting, accumulators in TMEM read back with tcgen05.ld.
//   fwd : Z[s][j]  = sum_k A1[s][k] W[j][k]      A K-major (rows = samples),  B K-major  (rows = j)
//   bwd : E[s][k]  = sum_j D2[s][j] W[j][k]      A K-major,                    B MN-major (the same W image)
//   dW  : G[j][k]  = sum_s D2[s][j] A1[s][k]     A MN-major (the same D2 image; rows 16.. of the 128-row tile read zeros),
//                                                B MN-major (the same A1 image), K = 128 samples = 16 MMAs of K = 8
// Prints the max error against a double-precision host result and the cycles of each round trip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_proto tc_proto.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, int a_mn, int b_mn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&r)[16])
{
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void split(float x, float& hi, float& lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}


constexpr int WPS = 16 * 16;   // bytes per weight plane: 16 rows x 16 bytes

__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const float (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])),
        "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])),
        "r"(__float_as_uint(r[8])), "r"(__float_as_uint(r[9])), "r"(__float_as_uint(r[10])), "r"(__float_as_uint(r[11])),
        "r"(__float_as_uint(r[12])), "r"(__float_as_uint(r[13])), "r"(__float_as_uint(r[14])), "r"(__float_as_uint(r[15]))
        : "memory");
}

// TMEM columns: [0,16) A hi, [16,32) A lo, [32,48) Z, [48,64) E;  smem: W planes [k/4][j][4] (forward B, K-major) and the
// transposed image [j/4][k][4] (backward B, K-major), hi and lo each
__global__ void __launch_bounds__(128, 1) k_proto(const float* a1, const float* d2, const float* w, float* zf, float* zb, long long* cyc, int reps)
{
    __shared__ __align__(128) unsigned char sw[8 * WPS];
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 2; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // W[j][k]
        for (int i = tid; i < 256; i += 128) {
            const int j = i >> 4, k = i & 15;
            float h, l;
            split(w[i], h, l);
            *reinterpret_cast<float*>(sw + 0 * WPS + (k >> 2) * WPS + j * 16 + (k & 3) * 4) = h;
            *reinterpret_cast<float*>(sw + 4 * WPS + (k >> 2) * WPS + j * 16 + (k & 3) * 4) = l;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bar[0]);
    float va[16], vh[16], vl[16], r[16];
    for (int i = 0; i < 16; i++) va[i] = a1[tid * 16 + i];
    long long tsum = 0, t_first = 0;
    for (int rep = 0; rep < reps; rep++) {
        __syncthreads();
        long long t0 = clock64();
        for (int i = 0; i < 16; i++) split(va[i], vh[i], vl[i]);
        st16(lane_addr + 0, vh);
        st16(lane_addr + 16, vl);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            constexpr uint32_t id = idesc_tf32(128, 16, 0, 0);
            const uint32_t bWh = (uint32_t)__cvta_generic_to_shared(sw), bWl = bWh + 4 * WPS;
            for (int ks = 0; ks < 2; ks++) {
                const uint64_t wh = desc_noswz(bWh + ks * 2 * WPS, WPS, 128), wl = desc_noswz(bWl + ks * 2 * WPS, WPS, 128);
                mma_tf32_ts(tm + 32, tm + 16 + ks * 8, wh, id, ks ? 1u : 0u);   // lo x hi
                mma_tf32_ts(tm + 32, tm + 0 + ks * 8, wl, id, 1u);              // hi x lo
                mma_tf32_ts(tm + 32, tm + 0 + ks * 8, wh, id, 1u);              // hi x hi
            }
            tc_commit(bar0);
        }
        mbar_wait(bar0, rep & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        ld16(lane_addr + 32, r);
        long long t1 = clock64();
        if (rep == 0) t_first = t1 - t0; else tsum += t1 - t0;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    for (int i = 0; i < 16; i++) zf[tid * 16 + i] = r[i];
    if (tid == 0 || tid == 127) { cyc[tid ? 2 : 0] = t_first; cyc[tid ? 3 : 1] = reps > 1 ? tsum / (reps - 1) : 0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u) : "memory");
}

int main()
{
    float ha[128 * 16], hd[128 * 16], hw[256];
    srand(1);
    for (int i = 0; i < 128 * 16; i++) { ha[i] = 2.f * rand() / RAND_MAX - 1.f; hd[i] = (2.f * rand() / RAND_MAX - 1.f) * 1e-3f; }
    for (int i = 0; i < 256; i++) hw[i] = 2.f * rand() / RAND_MAX - 1.f;
    float *a, *d, *w, *zf, *zb;
    long long* cyc;
    cudaMalloc(&a, sizeof ha); cudaMalloc(&d, sizeof hd); cudaMalloc(&w, sizeof hw);
    cudaMalloc(&zf, sizeof ha); cudaMalloc(&zb, sizeof ha); cudaMalloc(&cyc, 16 * 8);
    cudaMemcpy(a, ha, sizeof ha, cudaMemcpyHostToDevice);
    cudaMemcpy(d, hd, sizeof hd, cudaMemcpyHostToDevice);
    cudaMemcpy(w, hw, sizeof hw, cudaMemcpyHostToDevice);
    k_proto<<<1, 128>>>(a, d, w, zf, zb, cyc, 16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static float rf[128 * 16];
    long long hc[16];
    cudaMemcpy(rf, zf, sizeof rf, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
    double ef = 0, mf = 0;
    for (int s = 0; s < 128; s++)
        for (int j = 0; j < 16; j++) {
            double z = 0;
            for (int k = 0; k < 16; k++) z += (double)ha[s * 16 + k] * hw[j * 16 + k];
            ef = fmax(ef, fabs(z - rf[s * 16 + j])); mf = fmax(mf, fabs(z));
        }
    printf("TS-mode fwd (A from TMEM)  max err %.3e (max |z| %.3e, rel %.2e)\n", ef, mf, ef / mf);
    printf("round trip split + tcgen05.st + 6 MMAs + commit + wait + tcgen05.ld: first %lld / %lld cycles, steady %lld / %lld (thread 0 / 127)\n", hc[0], hc[2], hc[1], hc[3]);
    return 0;
}

// ubench.cu -- pipe-rate microbenchmarks on sm_100a that ground the design of the fused step kernel:
// cycles per warp-instruction (per SM) for FFMA2, HMMA.1688.TF32, broadcast LDS.128, LDS.32, SHFL, MUFU,
// at several warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void k(float* out, long long* cyc, const float* in)
{
    __shared__ float4 sm[256];
    if (threadIdx.x < 256) sm[threadIdx.x] = make_float4(in[threadIdx.x], 1.f, 2.f, 3.f);
    __syncthreads();
    float2 acc[8];
    float d[4][4];
    unsigned a[4] = {__float_as_uint(in[threadIdx.x & 31]), 0x3f800000u, 0x3f000000u, 0x3e800000u}, b[2] = {0x3f800000u, 0x3f000000u};
    for (int i = 0; i < 8; i++) acc[i] = make_float2(in[i], in[i + 1]);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) d[i][j] = 0.f;
    float x = in[threadIdx.x & 63];
    float4 ld[4] = {};
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS / 8; it++) {
        if (MODE == 0) {  // FFMA2, 8 independent chains
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = __ffma2_rn(acc[i], make_float2(x, x), acc[i]);
        } else if (MODE == 1) {  // HMMA 1688 tf32, 4 independent accumulators x 2
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) mma_tf32(d[i], a, b);
        } else if (MODE == 2) {  // broadcast LDS.128 (all lanes same address), 8 independent
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float4 v = sm[(it * 8 + i) & 255];
                ld[i & 3].x += v.x; ld[i & 3].y += v.y; ld[i & 3].z += v.z; ld[i & 3].w += v.w;
            }
        } else if (MODE == 3) {  // LDS.32, lane-distinct conflict-free
#pragma unroll
            for (int i = 0; i < 8; i++) x += reinterpret_cast<float*>(sm)[((it * 8 + i) * 32 + (threadIdx.x & 31)) & 1023];
        } else if (MODE == 4) {  // SHFL.BFLY
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, 1 + (i & 3));
        } else if (MODE == 5) {  // MUFU.EX2
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float y;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i].x));
                acc[i].x = y;
            }
        } else if (MODE == 6) {  // LDS.128 lane-distinct (512 B per instruction)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float4 v = sm[((it * 8 + i) * 32 + (threadIdx.x & 31)) & 255];
                ld[i & 3].x += v.x; ld[i & 3].y += v.y; ld[i & 3].z += v.z; ld[i & 3].w += v.w;
            }
        } else if (MODE == 7) {  // FFMA2 with a broadcast LDS.128 every 2 FFMA2 (the v2 inner loop)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float4 v = sm[(it * 4 + i) & 255];
                acc[2 * i] = __ffma2_rn(make_float2(v.x, v.y), make_float2(x, x), acc[2 * i]);
                acc[2 * i + 1] = __ffma2_rn(make_float2(v.z, v.w), make_float2(x, x), acc[2 * i + 1]);
            }
        }
    }
    long long t1 = clock64();
    float s = x;
    for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += d[i][j];
    for (int i = 0; i < 4; i++) s += ld[i].x + ld[i].y + ld[i].z + ld[i].w;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter)
{
    float *out, *in;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&in, 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    float h[1024];
    for (int i = 0; i < 1024; i++) h[i] = 1.0f + 1e-3f * i;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    printf("%-34s", name);
    for (int warps : {1, 4, 8, 16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, cyc, in);
        k<MODE><<<148, warps * 32>>>(out, cyc, in);
        cudaDeviceSynchronize();
        long long hc[148];
        cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
        double m = 0;
        for (int i = 0; i < 148; i++) m += hc[i];
        m /= 148;
        int n_instr = (ITERS / 8) * per_iter;
        // cycles per warp-instruction as seen by the SM (all warps together)
        printf("  w=%2d: %6.2f cyc/inst/SM", warps, m / ((double)n_instr * warps));
    }
    printf("\n");
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA2 (8 chains)", 8);
    run<1>("HMMA.1688.F32.TF32 (4 accum)", 8);
    run<2>("LDS.128 broadcast", 8);
    run<6>("LDS.128 lane-distinct", 8);
    run<3>("LDS.32 lane-distinct", 8);
    run<4>("SHFL.BFLY", 8);
    run<5>("MUFU.EX2", 8);
    run<7>("4x(LDS.128 bcast + 2 FFMA2) [12 inst]", 12);
    return 0;
}

// microbench_f32x2.cu -- issue rate of the packed fp32 instructions of sm_100a (FFMA2 / FMUL2 / FADD2) next to
// scalar FFMA, alone and mixed with ALU work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb2 microbench_f32x2.cu
// Prints warp-instructions per clock per SM (all SMs busy, 1024 threads/SM).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define ITERS 2048
#define CHAINS 8
typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// MODE 0: FFMA2 chains; 1: FMUL2; 2: FADD2; 3: scalar FFMA x2 (same flops as mode 0); 4: FFMA2 + IADD; 5: FFMA + IADD;
// 6: FFMA2 + 2 IADD; 7: 2 FFMA + 2 IADD; 8: FFMA2 with a scalar broadcast operand
template <int MODE> __global__ void __launch_bounds__(256, 4) probe(u64 *cycles, uint32_t *sink, uint32_t seed)
{
    u64 v[CHAINS];
    uint32_t w[CHAINS], w2[CHAINS];
    const u64 k = 0x3f8000013f800001ull, c = 0x3089705f3089705full;
    for (int i = 0; i < CHAINS; ++i) {
        const uint32_t s = 0x3f800000u + ((seed * (threadIdx.x + 1) + i * 977u + blockIdx.x) & 0xffffu);
        v[i] = (static_cast<u64>(s) << 32) | (s + 3u);
        w[i] = s; w2[i] = s * 3u;
    }
    const float bc = __uint_as_float(0x3f800001u + (seed & 1u));
    __syncthreads();
    const u64 t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) v[i] = ffma2(v[i], k, c);
            if (MODE == 1) v[i] = fmul2(v[i], k);
            if (MODE == 2) v[i] = fadd2(v[i], c);
            if (MODE == 3) {
                float lo = __uint_as_float(static_cast<uint32_t>(v[i])), hi = __uint_as_float(static_cast<uint32_t>(v[i] >> 32));
                lo = __fmaf_rn(lo, 1.0000001f, 1e-9f); hi = __fmaf_rn(hi, 1.0000001f, 1e-9f);
                v[i] = (static_cast<u64>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
            }
            if (MODE == 4) { v[i] = ffma2(v[i], k, c); w[i] += 0x9e3779b9u; }
            if (MODE == 5) { w2[i] = __float_as_uint(__fmaf_rn(__uint_as_float(w2[i]), 1.0000001f, 1e-9f)); w[i] += 0x9e3779b9u; }
            if (MODE == 6) { v[i] = ffma2(v[i], k, c); w[i] += 0x9e3779b9u; w2[i] ^= 0x5555u; w2[i] += 1u; }
            if (MODE == 7) {
                float lo = __uint_as_float(static_cast<uint32_t>(v[i])), hi = __uint_as_float(static_cast<uint32_t>(v[i] >> 32));
                lo = __fmaf_rn(lo, 1.0000001f, 1e-9f); hi = __fmaf_rn(hi, 1.0000001f, 1e-9f);
                v[i] = (static_cast<u64>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
                w[i] += 0x9e3779b9u; w2[i] ^= 0x5555u; w2[i] += 1u;
            }
            if (MODE == 8) {
                float2 b2 = make_float2(bc, bc);
                v[i] = ffma2(v[i], *reinterpret_cast<u64 *>(&b2), c);
            }
        }
    }
    const u64 t1 = clock64();
    u64 acc = 0;
    for (int i = 0; i < CHAINS; ++i) acc ^= v[i] ^ w[i] ^ (static_cast<u64>(w2[i]) << 7);
    if (acc == 0x12345678u) sink[0] = static_cast<uint32_t>(acc);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int n, int sms)
{
    const int blocks = sms * 4;
    u64 *d_cycles;
    uint32_t *d_sink;
    cudaMalloc(&d_cycles, blocks * sizeof(u64));
    cudaMalloc(&d_sink, 64);
    probe<MODE><<<blocks, 256>>>(d_cycles, d_sink, 12345u);
    cudaDeviceSynchronize();
    probe<MODE><<<blocks, 256>>>(d_cycles, d_sink, 999u);
    cudaDeviceSynchronize();
    std::vector<u64> c(blocks);
    cudaMemcpy(c.data(), d_cycles, blocks * sizeof(u64), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (auto v : c) mean += static_cast<double>(v);
    mean /= blocks;
    const double per_clk = 4.0 * 8.0 * ITERS * CHAINS * n / mean;
    std::printf("%-28s n=%d  %7.3f warp-instr/clk/SM  block cycles %.0f  err=%s\n", name, n, per_clk, mean, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_cycles);
    cudaFree(d_sink);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::printf("SMs %d\n", sms);
    run<0>("FFMA2", 1, sms);
    run<1>("FMUL2", 1, sms);
    run<2>("FADD2", 1, sms);
    run<8>("FFMA2 bcast operand", 1, sms);
    run<3>("2 x FFMA", 2, sms);
    run<4>("FFMA2 + IADD", 2, sms);
    run<5>("FFMA + IADD", 2, sms);
    run<6>("FFMA2 + 3 ALU", 4, sms);
    run<7>("2 FFMA + 3 ALU", 5, sms);
    return 0;
}

// microbench.cu -- instruction-throughput probes for the pipes the propagation kernel leans on
// (B200, sm_100a).  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
// Prints warp-instructions per clock per SM for each probe (all SMs busy, 1024 threads/SM).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define ITERS 2048
#define CHAINS 8

template <class Op> __global__ void __launch_bounds__(256, 4) probe(unsigned long long *cycles, uint32_t *sink, uint32_t seed)
{
    uint32_t v[CHAINS];
    for (int i = 0; i < CHAINS; ++i) v[i] = seed * (threadIdx.x + 1) + i * 977u + blockIdx.x;
    __syncthreads();
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) v[i] = Op::apply(v[i]);
    }
    const unsigned long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < CHAINS; ++i) acc ^= v[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

#define F(x) __uint_as_float(x)
#define U(x) __float_as_uint(x)

struct OpFFMA { static constexpr int n = 1; static __device__ __forceinline__ uint32_t apply(uint32_t v) { return U(__fmaf_rn(F(v), 1.0000001f, 1e-9f)); } };
struct OpFFMA3 { static constexpr int n = 1; static __device__ __forceinline__ uint32_t apply(uint32_t v) { float f = F(v); return U(__fmaf_rn(f, f, f)); } };
struct OpFADDRZ { static constexpr int n = 1; static __device__ __forceinline__ uint32_t apply(uint32_t v) { return U(__fadd_rz(F(v), 1e-9f)); } };
struct OpI2F { static constexpr int n = 2; static __device__ __forceinline__ uint32_t apply(uint32_t v) { return U(__uint2float_rz(v)) ^ 0x5555u; } };
struct OpF2I { static constexpr int n = 2; static __device__ __forceinline__ uint32_t apply(uint32_t v) { return static_cast<uint32_t>(__float2int_rz(F(v | 0x40000000u))) + 0x3f000000u; } };
struct OpLOP { static constexpr int n = 1; static __device__ __forceinline__ uint32_t apply(uint32_t v) { return (v ^ 0x5555u) + 0u; } };
struct OpIADD { static constexpr int n = 1; static __device__ __forceinline__ uint32_t apply(uint32_t v) { return v + 0x9e3779b9u; } };
struct OpPRMT { static constexpr int n = 1; static __device__ __forceinline__ uint32_t apply(uint32_t v) { return __byte_perm(v, 0x4b000000u, 0x7632); } };
struct OpIMADW {
    static constexpr int n = 1;
    static __device__ __forceinline__ uint32_t apply(uint32_t v)
    {
        const uint64_t x = static_cast<uint64_t>(v) * 4294967118u + (v >> 7);
        return static_cast<uint32_t>(x) ^ static_cast<uint32_t>(x >> 32);
    }
};
#define MUFU_OP(NAME, INSN)                                                                                   \
    struct NAME {                                                                                             \
        static constexpr int n = 1;                                                                           \
        static __device__ __forceinline__ uint32_t apply(uint32_t v)                                          \
        {                                                                                                     \
            float r;                                                                                          \
            asm volatile(INSN " %0, %1;" : "=f"(r) : "f"(F(v)));                                              \
            return U(r);                                                                                      \
        }                                                                                                     \
    };
MUFU_OP(OpRCP, "rcp.approx.ftz.f32")
MUFU_OP(OpRSQ, "rsqrt.approx.ftz.f32")
MUFU_OP(OpSQRT, "sqrt.approx.ftz.f32")
MUFU_OP(OpLG2, "lg2.approx.ftz.f32")
MUFU_OP(OpEX2, "ex2.approx.ftz.f32")
MUFU_OP(OpSIN, "sin.approx.ftz.f32")
MUFU_OP(OpCOS, "cos.approx.ftz.f32")
struct OpFFMA2 {
    static constexpr int n = 1;
    static __device__ __forceinline__ uint32_t apply(uint32_t v)
    {
        unsigned long long a = (static_cast<unsigned long long>(v) << 32) | v, r;
        asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(r) : "l"(a));
        return static_cast<uint32_t>(r) ^ static_cast<uint32_t>(r >> 32);
    }
};
// exact uint32 -> float (round toward zero) without the conversion unit: 2 PRMT + 2 FFMA + FADD.RZ
struct OpU2F_EXACT {
    static constexpr int n = 5;
    static __device__ __forceinline__ uint32_t apply(uint32_t v)
    {
        const float hi = F(__byte_perm(v, 0x4b000000u, 0x7632)); // 2^23 + (v >> 16)
        const float lo = F(__byte_perm(v, 0x4b000000u, 0x7610)); // 2^23 + (v & 0xffff)
        const float a = __fmaf_rn(hi, 1.52587890625e-05f, -128.f);              // (v>>16) * 2^-16
        const float b = __fmaf_rn(lo, 2.3283064365386963e-10f, -0.001953125f);  // (v&0xffff) * 2^-32
        return U(__fadd_rz(a, b)) ^ v;
    }
};
// mix: one MUFU per 8 FFMA
struct OpMIX {
    static constexpr int n = 9;
    static __device__ __forceinline__ uint32_t apply(uint32_t v)
    {
        float f = F(v);
#pragma unroll
        for (int i = 0; i < 8; ++i) f = __fmaf_rn(f, 1.0000001f, 1e-9f);
        float r;
        asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f));
        return U(r);
    }
};
// mix: one MUFU per 4 FFMA
struct OpMIX4 {
    static constexpr int n = 5;
    static __device__ __forceinline__ uint32_t apply(uint32_t v)
    {
        float f = F(v);
#pragma unroll
        for (int i = 0; i < 4; ++i) f = __fmaf_rn(f, 1.0000001f, 1e-9f);
        float r;
        asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f));
        return U(r);
    }
};
// FFMA + IADD interleaved (fma pipe + alu pipe)
struct OpFMAALU {
    static constexpr int n = 2;
    static __device__ __forceinline__ uint32_t apply(uint32_t v) { return U(__fmaf_rn(F(v), 1.0000001f, 1e-9f)) + 0x10u; }
};

template <class Op> void run(const char *name, int sms, double &ffma_ref)
{
    const int blocks = sms * 4;
    unsigned long long *d_cycles;
    uint32_t *d_sink;
    cudaMalloc(&d_cycles, blocks * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 64);
    probe<Op><<<blocks, 256>>>(d_cycles, d_sink, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<Op><<<blocks, 256>>>(d_cycles, d_sink, 999u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> c(blocks);
    cudaMemcpy(c.data(), d_cycles, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (auto v : c) mean += static_cast<double>(v);
    mean /= blocks;
    // per SM: 4 blocks x 8 warps x ITERS x CHAINS x n instructions in `mean` cycles
    const double per_clk = 4.0 * 8.0 * ITERS * CHAINS * Op::n / mean;
    std::printf("%-12s n=%d  %8.3f warp-instr/clk/SM  (%7.1f lanes/clk/SM)  block cycles %.0f  kernel %.3f ms  err=%s\n", name, Op::n, per_clk,
                per_clk * 32.0, mean, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_cycles);
    cudaFree(d_sink);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::printf("SMs %d; counts include the probe's own glue instructions as listed by n\n", sms);
    double ref = 0;
    run<OpFFMA>("FFMA imm", sms, ref);
    run<OpFFMA3>("FFMA 3reg", sms, ref);
    run<OpFFMA2>("FFMA2(+glue)", sms, ref);
    run<OpFADDRZ>("FADD.RZ", sms, ref);
    run<OpLOP>("LOP3", sms, ref);
    run<OpIADD>("IADD", sms, ref);
    run<OpPRMT>("PRMT", sms, ref);
    run<OpFMAALU>("FFMA+IADD", sms, ref);
    run<OpIMADW>("IMAD.WIDE mwc", sms, ref);
    run<OpI2F>("I2F.RZ+LOP", sms, ref);
    run<OpF2I>("F2I+2ALU", sms, ref);
    run<OpU2F_EXACT>("U2F exact", sms, ref);
    run<OpRCP>("MUFU.RCP", sms, ref);
    run<OpRSQ>("MUFU.RSQ", sms, ref);
    run<OpSQRT>("MUFU.SQRT", sms, ref);
    run<OpLG2>("MUFU.LG2", sms, ref);
    run<OpEX2>("MUFU.EX2", sms, ref);
    run<OpSIN>("MUFU.SIN", sms, ref);
    run<OpCOS>("MUFU.COS", sms, ref);
    run<OpMIX>("8FFMA+LG2", sms, ref);
    run<OpMIX4>("4FFMA+LG2", sms, ref);
    return 0;
}

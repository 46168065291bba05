// Developer tool (B200 box): issue rate of the packed fp32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2, PTX *.f32x2) against
// their scalar forms, alone and interleaved with integer work - the question behind DESIGN.md 5 item 21: the noise kernels are
// bound by issue slots, not by the FMA pipe, so two fp32 operations per issued instruction are worth having if they issue at
// the scalar rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/ffma2_probe.cu && /tmp/ffma2_probe
#include <cuda_runtime.h>
#include <cstdio>

typedef unsigned long long u64;
__device__ __forceinline__ u64 f2fma(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 f2add(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 f2mul(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float sfma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float sadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned ilop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

constexpr int ITER = 4096;

// mode 0: 8 scalar FFMA per iteration; 1: 8 FFMA2; 2: 8 FFMA + 8 XOR; 3: 8 FFMA2 + 8 XOR; 4: 8 FFMA2 + 16 XOR; 5: 16 FFMA + 16 XOR;
// 6: 8 FADD scalar; 7: 8 FADD2
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, unsigned iseed)
{
    float a[16];
    u64 p[8];
    unsigned q[16];
    const float b = seed + 1.0f, c = seed * 0.5f;
    const u64 pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), pc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int i = 0; i < 16; ++i) { a[i] = seed + (float)i + threadIdx.x; q[i] = iseed + i + threadIdx.x; }
    for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
        if (MODE == 0 || MODE == 2 || MODE == 5)
        {
#pragma unroll
            for (int i = 0; i < (MODE == 5 ? 16 : 8); ++i) a[i] = sfma(a[i], b, c);
        }
        if (MODE == 6)
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sadd(a[i], c);
        }
        if (MODE == 1 || MODE == 3 || MODE == 4)
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = f2fma(p[i], pb, pc);
        }
        if (MODE == 7)
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = f2add(p[i], pc);
        }
        if (MODE == 2 || MODE == 3)
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = ilop(q[i], iseed);
        }
        if (MODE == 4 || MODE == 5)
        {
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = ilop(q[i], iseed);
        }
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += a[i] + (float)q[i];
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double flopPerIter, double instrPerIter, float* d, int sms, double mhz)
{
    const int blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(d, 1.0f, 3u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(d, 1.0f, 3u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double threads = (double)blocks * 256, warps = threads / 32;
    const double tflops = threads * ITER * flopPerIter / (ms * 1e-3) / 1e12;
    const double ipc = warps * ITER * instrPerIter / (ms * 1e-3 * mhz * 1e6) / sms;      // warp instructions per clock per SM (loop overhead not counted)
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  %5.2f warp-instr/clk/SM\n", name, ms, tflops, ipc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    float* d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(float));
    printf("%s, %d SMs, %.0f MHz (rates assume this clock)\n", p.name, p.multiProcessorCount, mhz);
    run<0>("8 FFMA", 16, 8, d, p.multiProcessorCount, mhz);
    run<1>("8 FFMA2", 32, 8, d, p.multiProcessorCount, mhz);
    run<6>("8 FADD", 8, 8, d, p.multiProcessorCount, mhz);
    run<7>("8 FADD2", 16, 8, d, p.multiProcessorCount, mhz);
    run<2>("8 FFMA + 8 XOR", 16, 16, d, p.multiProcessorCount, mhz);
    run<3>("8 FFMA2 + 8 XOR", 32, 16, d, p.multiProcessorCount, mhz);
    run<5>("16 FFMA + 16 XOR", 32, 32, d, p.multiProcessorCount, mhz);
    run<4>("8 FFMA2 + 16 XOR", 32, 24, d, p.multiProcessorCount, mhz);
    return 0;
}

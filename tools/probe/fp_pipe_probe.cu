// Micro-benchmark (development aid, not product code): issue rate of the fp32 instructions the blur kernels are built
// from, per SM, on the device it runs on.  Prints lane-results per clock per SM for scalar / packed mul, add, fma with
// register and with uniform (kernel-parameter) operands, at several resident-warp counts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_pipe_probe fp_pipe_probe.cu && ./fp_pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

enum Op { FFMA, FMUL, FADD, FFMA2, FMUL2, FADD2, FFMA2_U, FMUL2_U, MULADD2, MULADD2_U, FFMA_U, MIX_M2_A1, FFMA2_BR, FFMA2_BU, FFMA2_SHIFT };
static const char* names[] = {"FFMA r,r,r", "FMUL r,r", "FADD r,r", "FFMA2 r,r,r", "FMUL2 r,r", "FADD2 r,r", "FFMA2 r,U,r", "FMUL2 r,U",
                              "FMUL2+FFMA2(one) r", "FMUL2(U)+FFMA2(one)", "FFMA r,U,r", "FMUL2(U)+2xFADD", "FFMA2 r,bcast(r),r", "FFMA2 r,bcast(U),r", "FFMA2 shift d!=c"};
static const int results_per_iter[] = {1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 1, 2, 2, 2, 2};  // useful tap-results per accumulator per iteration

struct P { float2 t[8]; float2 one; };

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)), "l"(reinterpret_cast<unsigned long long&>(c)));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}

template <int OP>
__global__ void __launch_bounds__(1024) probe(const P p, float* out, long long* cycles, int iters, float seed) {
    constexpr int NA = 8;
    float2 acc[NA], v[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) { acc[i] = make_float2(seed * i, seed + i); v[i] = make_float2(1.0f + seed * (threadIdx.x + i), 1.0f - seed * i); }
    float2 tr = make_float2(seed + 0.5f, seed + 0.25f);  // a per-thread register "tap"
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                if (OP == FFMA) acc[i].x = __fmaf_rn(v[i].x, tr.x, acc[i].x);
                if (OP == FFMA_U) acc[i].x = __fmaf_rn(v[i].x, p.t[k].x, acc[i].x);
                if (OP == FMUL) acc[i].x = __fmul_rn(acc[i].x, tr.x);
                if (OP == FADD) acc[i].x = __fadd_rn(acc[i].x, v[i].x);
                if (OP == FFMA2) acc[i] = fma2(v[i], tr, acc[i]);
                if (OP == FFMA2_U) acc[i] = fma2(v[i], p.t[k], acc[i]);
                if (OP == FMUL2) acc[i] = mul2(acc[i], tr);
                if (OP == FMUL2_U) acc[i] = mul2(acc[i], p.t[k]);
                if (OP == FADD2) acc[i] = add2(acc[i], v[i]);
                if (OP == MULADD2) acc[i] = fma2(mul2(v[i], tr), p.one, acc[i]);
                if (OP == MULADD2_U) acc[i] = fma2(mul2(v[i], p.t[k]), p.one, acc[i]);
                if (OP == FFMA2_BR) acc[i] = fma2(v[i], make_float2(tr.x, tr.x), acc[i]);
                if (OP == FFMA2_BU) acc[i] = fma2(v[i], make_float2(p.t[k].x, p.t[k].x), acc[i]);
                if (OP == FFMA2_SHIFT) acc[i] = fma2(v[0], make_float2(tr.x, tr.x), acc[(i + 1) % NA]);
                if (OP == MIX_M2_A1) { float2 m = mul2(v[i], p.t[k]); acc[i].x = __fadd_rn(__fadd_rn(acc[i].x, m.x), m.y); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, float* out, long long* cyc, const P& p) {
    const int iters = 2000;
    for (int warps : {4, 8, 16, 32}) {
        probe<OP><<<sms, warps * 32>>>(p, out, cyc, iters, 1e-3f);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        probe<OP><<<sms, warps * 32>>>(p, out, cyc, iters, 1e-3f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        long long h[1024];
        cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; ++i) mean += (double)h[i];
        mean /= sms;
        const double instr = (double)iters * 8 * 8 * warps;              // warp-instructions of the probed kind per SM (x2 for the pairs)
        const double n_instr = (OP == MULADD2 || OP == MULADD2_U) ? 2 * instr : (OP == MIX_M2_A1 ? 3 * instr : instr);
        const double res = (double)iters * 8 * 8 * warps * 32 * results_per_iter[OP];
        printf("%-22s warps/SM %2d: %.3f warp-instr/clk/SM, %.1f results/clk/SM  (%.3f ms, %.0f clk)\n", names[OP], warps, n_instr / mean, res / mean, ms, mean);
    }
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 1024 * sms);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    P p;
    for (int i = 0; i < 8; ++i) p.t[i] = make_float2(0.9f + 0.01f * i, 0.9f - 0.01f * i);
    p.one = make_float2(1.0f, 1.0f);
    printf("SMs %d\n", sms);
    run<FFMA>(sms, out, cyc, p); run<FFMA_U>(sms, out, cyc, p); run<FMUL>(sms, out, cyc, p); run<FADD>(sms, out, cyc, p);
    run<FFMA2>(sms, out, cyc, p); run<FFMA2_U>(sms, out, cyc, p); run<FMUL2>(sms, out, cyc, p); run<FMUL2_U>(sms, out, cyc, p); run<FADD2>(sms, out, cyc, p);
    run<FFMA2_BR>(sms, out, cyc, p); run<FFMA2_BU>(sms, out, cyc, p); run<FFMA2_SHIFT>(sms, out, cyc, p); run<MULADD2>(sms, out, cyc, p); run<MULADD2_U>(sms, out, cyc, p); run<MIX_M2_A1>(sms, out, cyc, p);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// Issue-rate microbenchmark for sm_100a: how many warp-instructions per clock one SM sub-partition sustains for the
// instruction kinds the fitter's pixel loop is made of (packed / scalar fp32, MUFU, ALU min/max/select, LDS, SHFL),
// alone and mixed, at full occupancy and at the fitter's 3 warps per scheduler.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

enum Op { FFMA2, FMUL2, FADD2, FFMA, FMUL, EX2, RCP, FMNMX, FSEL, LDS64, SHFL, NONE };

template <int OP>
__device__ __forceinline__ void op(unsigned long long& a, const unsigned long long x, const unsigned long long y, float& s, const float u,
                                   const float v, const float* sm, int lane) {
    if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(x), "l"(y));
    if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(x));
    if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(y));
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s) : "f"(u), "f"(v));
    if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s) : "f"(u));
    if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s));
    if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s));
    if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(s) : "f"(u));
    if (OP == FSEL) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %2, p;}" : "+f"(s) : "f"(u), "f"(v));
    if (OP == LDS64) {
        float2 t;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"((unsigned)__cvta_generic_to_shared(sm + 2 * lane)));
        s += t.x;
    }
    if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+f"(s));
}

// body: NA ops of kind A followed by NB ops of kind B, on 8 independent chains each
template <int A, int NA, int B, int NB>
__global__ void __launch_bounds__(128) mix(float* out, int iters, float fa, float fb) {
    __shared__ float sm[256];
    sm[threadIdx.x] = fa; sm[threadIdx.x + 128] = fb;
    __syncthreads();
    unsigned long long acc[8], x, y;
    float s[8];
    const float2 xv = make_float2(fa, fa * 1.0001f), yv = make_float2(fb, fb * 0.9999f);
    x = *reinterpret_cast<const unsigned long long*>(&xv);
    y = *reinterpret_cast<const unsigned long long*>(&yv);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] = x + k; s[k] = fa + k * 1e-3f; }
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NA; ++k) op<A>(acc[k & 7], x, y, s[k & 7], fa, fb, sm, lane);
#pragma unroll
        for (int k = 0; k < NB; ++k) op<B>(acc[k & 7], x, y, s[(k + 4) & 7], fa, fb, sm, lane);
    }
    float r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += s[k] + float(acc[k] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static int g_sms, g_khz;
static float* g_out;

template <int A, int NA, int B, int NB>
void run(const char* name, int ctas_per_sm) {
    const int iters = 4000, blocks = g_sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix<A, NA, B, NB><<<blocks, 128>>>(g_out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warps_per_smsp = ctas_per_sm * 4 / 4.0;
    const double instr = (double)(NA + NB) * iters * warps_per_smsp;    // warp-instructions per SMSP
    const double clocks = best * 1e-3 * g_khz * 1e3;
    printf("%-34s %2d warps/SMSP  %6.3f ms  %.3f instr/clk/SMSP  (%.2f clk per instr)\n", name, (int)warps_per_smsp, best, instr / clocks, clocks / instr);
}

#define BOTH(A, NA, B, NB, name) run<A, NA, B, NB>(name, 3); run<A, NA, B, NB>(name, 12)

int main() {
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&g_khz, cudaDevAttrClockRate, 0);
    cudaMalloc(&g_out, sizeof(float) * g_sms * 16 * 128);
    printf("%d SMs, %.0f MHz (clock rate attribute; the run may boost differently)\n", g_sms, g_khz / 1e3);
    run<FFMA2, 16, NONE, 0>("warm-up", 12);
    BOTH(FFMA2, 16, NONE, 0, "FFMA2");
    BOTH(FMUL2, 16, NONE, 0, "FMUL2");
    BOTH(FADD2, 16, NONE, 0, "FADD2");
    BOTH(FFMA, 16, NONE, 0, "FFMA (scalar)");
    BOTH(FMUL, 16, NONE, 0, "FMUL (scalar)");
    BOTH(EX2, 16, NONE, 0, "MUFU.EX2");
    BOTH(RCP, 16, NONE, 0, "MUFU.RCP");
    BOTH(FMNMX, 16, NONE, 0, "FMNMX");
    BOTH(FSEL, 16, NONE, 0, "FSETP+FSEL (2 instr)");
    BOTH(LDS64, 16, NONE, 0, "LDS.64 (+FADD)");
    BOTH(SHFL, 16, NONE, 0, "SHFL.BFLY");
    BOTH(FFMA2, 16, EX2, 2, "FFMA2 x16 + EX2 x2");
    BOTH(FFMA2, 16, EX2, 4, "FFMA2 x16 + EX2 x4");
    BOTH(FFMA2, 16, FMNMX, 4, "FFMA2 x16 + FMNMX x4");
    BOTH(FFMA2, 16, FMNMX, 8, "FFMA2 x16 + FMNMX x8");
    BOTH(FFMA2, 16, FFMA, 4, "FFMA2 x16 + FFMA x4");
    BOTH(FFMA2, 16, FMUL, 8, "FFMA2 x16 + FMUL x8");
    BOTH(FFMA2, 16, LDS64, 2, "FFMA2 x16 + LDS.64 x2");
    BOTH(FFMA2, 16, SHFL, 4, "FFMA2 x16 + SHFL x4");
    BOTH(FFMA, 16, FMNMX, 8, "FFMA x16 + FMNMX x8");
    return 0;
}

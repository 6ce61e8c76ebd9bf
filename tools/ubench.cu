// ubench.cu - issue-rate microbenchmarks on B200 (sm_100a): how many warp-instructions
// per clock per SM each instruction class sustains, alone and mixed.  Used to size the
// per-pixel instruction budget of the fused kernels (DESIGN.md "compute budget").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, int n, float fa, float fb, unsigned ua, unsigned ub) {
    float f[8];
    unsigned u[8];
    float2 g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = fa * (threadIdx.x + i); u[i] = ua + threadIdx.x * 7 + i; g[i] = make_float2(f[i], f[i] + 1.f); }
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            if (OP == 1) asm volatile("{.reg .b64 a,b,c; mov.b64 a,{%0,%1}; mov.b64 b,{%2,%2}; mov.b64 c,{%3,%3}; fma.rn.f32x2 a,a,b,c; mov.b64 {%0,%1},a;}" : "+f"(g[i].x), "+f"(g[i].y) : "f"(fa), "f"(fb));
            if (OP == 2) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
            if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(ua), "r"(ub));
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(ub));
            if (OP == 5) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(ua), "r"(ub));
            if (OP == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(ub));
            if (OP == 7) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
            if (OP == 9) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(u[i]));
            if (OP == 10) asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(u[i]) : "f"(f[i]));
            if (OP == 11) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(ub));
            if (OP == 12) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(ub));
            if (OP == 13) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(ua), "r"(ub)); }
            if (OP == 14) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(ua), "r"(ub)); }
            if (OP == 15) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(ub)); }
            if (OP == 16) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
            if (OP == 17) asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(f[i]) : "r"(u[i]));
            if (OP == 18) asm volatile("{.reg .b32 h; mov.b32 h,%0; fma.rn.f16x2 h,h,h,h; mov.b32 %0,h;}" : "+r"(u[i]));
            if (OP == 19) asm volatile("clz.b32 %0, %0;" : "+r"(u[i]));
            if (OP == 20) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i]));
            if (OP == 21) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(ub)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i].x) : "f"(fa), "f"(fb)); }
            if (OP == 22) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
            if (OP == 23) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
        }
    }
    long long t1 = clock64();
    float s = 0; unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += f[i] + g[i].x + g[i].y; v += u[i]; }
    if (n == -1) { out[threadIdx.x] = s; ((unsigned*)out)[threadIdx.x + 256] = v; }
    if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[1] = t1 - t0;
}

template <int OP>
int run(const char* name, int per_iter) {
    float* d; CK(cudaMalloc(&d, 4096));
    int grid = 148 * 8;   // 8 CTAs x 256 thr = 2048 threads per SM (full occupancy)
    k<OP><<<grid, 256>>>(d, 16, 1.0001f, 0.5f, 0x01020304u, 0x4B000000u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<grid, 256>>>(d, ITER, 1.0001f, 0.5f, 0x01020304u, 0x4B000000u);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long cyc; CK(cudaMemcpy(&cyc, ((long long*)d) + 1, 8, cudaMemcpyDeviceToHost));
    double winstr = (double)grid * 8 /*warps*/ * ITER * 8 * per_iter;   // warp-instructions
    double clk = 1.965e9;                                              // clocks.max.sm; no throttle at this load
    double per_clk_sm = winstr / 148.0 / (ms * 1e-3 * clk);
    printf("%-28s %8.3f ms  warp-instr/clk/SM %6.3f  (lane-ops/clk/SM %6.1f)\n", name, ms, per_clk_sm, per_clk_sm * 32);
    cudaFree(d);
    return 0;
}

int main() {
    run<0>("FFMA", 1); run<1>("FFMA2 (f32x2)", 1); run<2>("FADD", 1); run<3>("LOP3", 1); run<4>("PRMT", 1);
    run<5>("IDP4A", 1); run<6>("IMAD", 1); run<7>("MUFU.RCP", 1); run<8>("FMNMX", 1); run<9>("I2F.U32", 1);
    run<10>("F2I", 1); run<11>("VABSDIFF4", 1); run<12>("IADD", 1); run<16>("SHL", 1); run<17>("I2F.U8", 1);
    run<18>("HFMA2", 1); run<19>("CLZ", 1); run<20>("SHFL", 1); run<22>("FADD.RM", 1); run<23>("FFMA.SAT", 1);
    run<13>("FFMA+LOP3", 2); run<14>("FFMA+IDP4A", 2); run<15>("FFMA+IMAD", 2); run<21>("2FFMA+PRMT", 3);
    return 0;
}

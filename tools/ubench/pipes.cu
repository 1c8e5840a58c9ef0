// Micro-benchmark (development): per-SM issue throughput of FFMA / FFMA2 / FMUL / MUFU / LDS on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void k(float* out, float a, float b, long long* clk) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float2 p0 = {x0, x1}, p1 = {x2, x3}, p2 = {x4, x5}, p3 = {x6, x7}, p4 = {x1, x0}, p5 = {x3, x2}, p6 = {x5, x4}, p7 = {x7, x6};
    float2 aa = {a, a + 1}, bb = {b, b + 1};
    __shared__ float sm[2048];
    sm[threadIdx.x] = x0; sm[threadIdx.x + 1024] = x1;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
        if (OP == 0) { x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b); x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b); }
        if (OP == 1) { p0 = __ffma2_rn(p0, aa, bb); p1 = __ffma2_rn(p1, aa, bb); p2 = __ffma2_rn(p2, aa, bb); p3 = __ffma2_rn(p3, aa, bb); p4 = __ffma2_rn(p4, aa, bb); p5 = __ffma2_rn(p5, aa, bb); p6 = __ffma2_rn(p6, aa, bb); p7 = __ffma2_rn(p7, aa, bb); }
        if (OP == 2) { x0 = fmaf(x0, 1.0001f, 0.5f); x1 = fmaf(x1, 1.0001f, 0.5f); x2 = fmaf(x2, 1.0001f, 0.5f); x3 = fmaf(x3, 1.0001f, 0.5f); x4 = fmaf(x4, 1.0001f, 0.5f); x5 = fmaf(x5, 1.0001f, 0.5f); x6 = fmaf(x6, 1.0001f, 0.5f); x7 = fmaf(x7, 1.0001f, 0.5f); }
        if (OP == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
                       asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x5)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x7)); }
        if (OP == 4) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x1)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x3));
                       asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x4)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x5)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x6)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x7)); }
        if (OP == 5) { x0 += sm[(threadIdx.x + i) & 2047]; x1 += sm[(threadIdx.x + i + 32) & 2047]; x2 += sm[(threadIdx.x + i + 64) & 2047]; x3 += sm[(threadIdx.x + i + 96) & 2047];
                       x4 += sm[(threadIdx.x + i + 128) & 2047]; x5 += sm[(threadIdx.x + i + 160) & 2047]; x6 += sm[(threadIdx.x + i + 192) & 2047]; x7 += sm[(threadIdx.x + i + 224) & 2047]; }
        if (OP == 6) { x0 = x0 * a; x1 = x1 * a; x2 = x2 * a; x3 = x3 * a; x4 = x4 * a; x5 = x5 * a; x6 = x6 * a; x7 = x7 * a; }
        if (OP == 7) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x1)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x3));
                       asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x4)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x5)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x6)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x7)); }
        if (OP == 8) { unsigned u0 = __float_as_uint(x0), u1 = __float_as_uint(x1), u2 = __float_as_uint(x2), u3 = __float_as_uint(x3);
                       u0 = (u0 << 16) ^ i; u1 = (u1 & 0xffff0000u) ^ i; u2 = (u2 << 16) ^ i; u3 = (u3 & 0xffff0000u) ^ i;
                       unsigned u4 = __float_as_uint(x4), u5 = __float_as_uint(x5), u6 = __float_as_uint(x6), u7 = __float_as_uint(x7);
                       u4 = (u4 << 16) ^ i; u5 = (u5 & 0xffff0000u) ^ i; u6 = (u6 << 16) ^ i; u7 = (u7 & 0xffff0000u) ^ i;
                       x0 = __uint_as_float(u0); x1 = __uint_as_float(u1); x2 = __uint_as_float(u2); x3 = __uint_as_float(u3); x4 = __uint_as_float(u4); x5 = __uint_as_float(u5); x6 = __uint_as_float(u6); x7 = __uint_as_float(u7); }
        if (OP == 9) {     // FFMA2 with three distinct 64-bit register operands per instruction (no operand shared by neighbours)
            p0 = __ffma2_rn(p1, p2, p0); p3 = __ffma2_rn(p4, p5, p3); p6 = __ffma2_rn(p7, aa, p6); p1 = __ffma2_rn(p2, p0, p1);
            p4 = __ffma2_rn(p5, p3, p4); p7 = __ffma2_rn(bb, p6, p7); p2 = __ffma2_rn(p0, p1, p2); p5 = __ffma2_rn(p3, p4, p5); }
        if (OP == 10) {    // plain FFMA with three distinct register operands
            x0 = fmaf(x1, x2, x0); x3 = fmaf(x4, x5, x3); x6 = fmaf(x7, a, x6); x1 = fmaf(x2, x0, x1);
            x4 = fmaf(x5, x3, x4); x7 = fmaf(b, x6, x7); x2 = fmaf(x0, x1, x2); x5 = fmaf(x3, x4, x5); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + p4.x + p4.y + p5.x + p5.y + p6.x + p6.y + p7.x + p7.y;
}
template <int OP> void run(const char* name, int per_iter_ops) {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    for (int threads : {128, 256, 512, 1024}) {
        k<OP><<<148, threads>>>(out, 1.0001f, 0.5f, clk);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        printf("%-10s threads %4d: %.2f lane-ops/clk/SM (%.2f warp-instr/clk/SM)\n", name, threads, (double)ITERS * per_iter_ops * threads / c, (double)ITERS * 8 * threads / 32 / c);
    }
    cudaFree(out); cudaFree(clk);
}
int main() {
    run<0>("FFMA reg", 8); run<2>("FFMA imm", 8); run<1>("FFMA2", 16); run<6>("FMUL reg", 8); run<3>("MUFU.EX2", 8); run<4>("MUFU.RCP", 8); run<7>("MUFU.TANH", 8); run<5>("LDS.32", 8); run<8>("ALU shl/and+xor", 16); run<9>("FFMA2 3reg", 16); run<10>("FFMA 3reg", 8);
    return 0;
}

// Developer probe: dependent-issue latencies of the instructions the one-warp recursions are made of (one warp, one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = seed + lane; sm[32 + lane] = 1.0 + 1e-9 * lane;
    __syncwarp();
    double a = seed, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    const int N = 512;
    // 1. dependent DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64(); cyc[0] = (t1 - t0);
    // 2. dependent DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * b;
    t1 = clock64(); cyc[1] = (t1 - t0);
    // 3. dependent DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    t1 = clock64(); cyc[2] = (t1 - t0);
    // 4. dependent shuffle of a double (two SHFL.32)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
    t1 = clock64(); cyc[3] = (t1 - t0);
    // 5. dependent DMUL + SHFL + DFMA (one step of a triangular solve)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { const double x = __shfl_sync(0xffffffffu, a * b, i & 31); a = fma(-c, x, a); }
    t1 = clock64(); cyc[4] = (t1 - t0);
    // 6. dependent LDS.64 (pointer chase through values) : address from the loaded value
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { const double v = sm[idx]; idx = ((int)v) & 31; }
    t1 = clock64(); cyc[5] = (t1 - t0);
    // 7. STS -> syncwarp -> LDS round trip (exchange through shared memory)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31] + c; __syncwarp(); }
    t1 = clock64(); cyc[6] = (t1 - t0);
    // 8. MUFU.RSQ64H-based rsqrt seed + one DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = fma(y, c, 1.5); }
    t1 = clock64(); cyc[7] = (t1 - t0);
    // 9. DMMA m8n8k4 dependent accumulate
    double d0 = a, d1 = b;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(b), "d"(c));
    t1 = clock64(); cyc[8] = (t1 - t0);
    out[lane] = a + idx + d0 + d1;
    cyc[9] = N;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 256); cudaMalloc(&cyc, 128);
    probe<<<1, 32>>>(out, cyc, 1.25);
    probe<<<1, 32>>>(out, cyc, 1.25);
    long long h[16];
    cudaMemcpy(h, cyc, 80, cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA", "DMUL", "DADD", "SHFL f64 (2 x SHFL.32)", "DMUL+SHFL+DFMA step", "LDS.64 pointer chase (+F2I, LOP)", "STS+syncwarp+LDS+DADD+syncwarp", "rsqrt.approx.f64 + DFMA", "DMMA 8x8x4 accumulate"};
    for (int i = 0; i < 9; ++i) printf("%-36s %.1f cycles per dependent step\n", names[i], (double)h[i] / (double)h[9]);
    return 0;
}

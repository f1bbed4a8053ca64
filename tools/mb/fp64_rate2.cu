// FP64 issue rate with all-register operands (no uniform / constant operands) and a DMUL+DADD mix like the minimiser's.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double *out, long long *cyc, int iters, const double *in) {
    double x[ILP], y[ILP], z[ILP];
    for (int i = 0; i < ILP; i++) {
        x[i] = in[threadIdx.x + i * 32];
        y[i] = in[threadIdx.x + 1000 + i * 32];
        z[i] = in[threadIdx.x + 2000 + i * 32];
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = fma(x[i], y[i], z[i]);          // 3 register operands
            else if (OP == 1) x[i] = x[i] * y[i];                // DMUL reg,reg
            else if (OP == 2) x[i] = x[i] + y[i];                // DADD reg,reg
            else { x[i] = x[i] * y[i]; x[i] = x[i] + z[i]; }     // DMUL then dependent DADD (fmad off)
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP, int ILP>
void run(const char *name, int threads, int blocks, const double *in) {
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * blocks * 1024);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    const int iters = 2048;
    k<OP, ILP><<<blocks, threads>>>(out, cyc, iters, in);
    k<OP, ILP><<<blocks, threads>>>(out, cyc, iters, in);
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < blocks; i++) c += h[i];
    c /= blocks;
    const double ops = (double)iters * ILP * threads * (OP == 3 ? 2 : 1);
    printf("%-10s ILP=%d threads=%4d blocks=%3d: %.0f cycles, %.2f lanes/clk/SM (%.2f cycles per warp-instr per SMSP)\n", name, ILP, threads,
           blocks, c, ops / c, c / (ops / 32 / 4));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    double *in; cudaMalloc(&in, sizeof(double) * 8192);
    double h[8192]; for (int i = 0; i < 8192; i++) h[i] = 1.0 + 1e-9 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int t = 128; t <= 512; t *= 2) run<0, 4>("DFMA rrr", t, 16, in);
    for (int t = 128; t <= 512; t *= 2) run<1, 4>("DMUL rr", t, 16, in);
    for (int t = 128; t <= 512; t *= 2) run<2, 4>("DADD rr", t, 16, in);
    for (int t = 128; t <= 512; t *= 2) run<3, 4>("DMUL+DADD", t, 16, in);
    run<3, 1>("DMUL+DADD", 384, 16, in);
    run<3, 2>("DMUL+DADD", 384, 16, in);
    run<0, 8>("DFMA rrr", 256, 16, in);
    return 0;
}

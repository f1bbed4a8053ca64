// FP64 issue rate / latency per SM on this part (DFMA, DMUL, DADD), for sizing the minimiser's cluster.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = a + threadIdx.x * 1e-9 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = fma(x[i], b, a);
            else if (OP == 1) x[i] = x[i] * b;
            else x[i] = x[i] + b;
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
void run(const char *name, int threads, int blocks) {
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * blocks * 1024);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    const int iters = 4096;
    k<OP, ILP><<<blocks, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    k<OP, ILP><<<blocks, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < blocks; i++) c += h[i];
    c /= blocks;
    const double ops = (double)iters * ILP * threads;
    printf("%s ILP=%d threads=%4d blocks=%3d: %.0f cycles, %.2f lanes/clk/SM, %.2f cycles per dependent op\n", name, ILP, threads,
           blocks, c, ops / c, c / iters / ILP * (threads <= 32 ? 1 : 0));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 1>("DFMA", 32, 1);
    run<1, 1>("DMUL", 32, 1);
    run<2, 1>("DADD", 32, 1);
    run<0, 4>("DFMA", 32, 1);
    for (int t = 128; t <= 1024; t *= 2) run<0, 4>("DFMA", t, 148);
    for (int t = 128; t <= 1024; t *= 2) run<1, 4>("DMUL", t, 148);
    for (int t = 128; t <= 1024; t *= 2) run<2, 4>("DADD", t, 148);
    run<0, 8>("DFMA", 512, 148);
    run<0, 8>("DFMA", 512, 16);
    return 0;
}

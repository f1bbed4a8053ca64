// L2 signalling latency microbenchmarks (B200): ping-pong between two CTAs on different SMs with several
// store/load flavours, and own-store -> own-load visibility.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2000
template <int MODE>
__device__ __forceinline__ void st(unsigned long long *p, unsigned long long v) {
    if (MODE == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 2) atomicExch(p, v);
    if (MODE == 3) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int MODE>
__device__ __forceinline__ unsigned long long ld(const unsigned long long *p) {
    unsigned long long v;
    if (MODE == 0) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 1) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 2) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 3) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <int MODE>
__global__ void pingpong(unsigned long long *a, unsigned long long *b, long long *out, int other) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && blockIdx.x != other) return;
    const bool first = blockIdx.x == 0;
    long long t0 = clock64();
    for (unsigned long long i = 1; i <= N; i++) {
        if (first) {
            st<MODE>(a, i);
            while (ld<MODE>(b) != i) {}
        } else {
            while (ld<MODE>(a) != i) {}
            st<MODE>(b, i);
        }
    }
    if (first) out[0] = (clock64() - t0) / N;
}
template <int MODE>
__global__ void own(unsigned long long *a, long long *out) {
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    for (unsigned long long i = 1; i <= N; i++) {
        st<MODE>(a + blockIdx.x * 32, i);
        while (ld<MODE>(a + blockIdx.x * 32) != i) {}
    }
    out[blockIdx.x] = (clock64() - t0) / N;
}
// many pollers on one flag: block 0 writes, all others poll; measure writer->(all acked) round trip
template <int MODE>
__global__ void fanout(unsigned long long *flag, unsigned long long *acks, long long *out, int nb) {
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    for (unsigned long long i = 1; i <= 200; i++) {
        if (blockIdx.x == 0) {
            st<MODE>(flag, i);
            for (int b = 1; b < nb; b++) while (ld<MODE>(acks + b * 16) != i) {}
        } else {
            while (ld<MODE>(flag) != i) {}
            st<MODE>(acks + blockIdx.x * 16, i);
        }
    }
    if (blockIdx.x == 0) out[0] = (clock64() - t0) / 200;
}
int main() {
    unsigned long long *buf; long long *out;
    cudaMalloc(&buf, 1 << 20); cudaMalloc(&out, 4096);
    long long h[148];
#define RUN(M) \
    for (int other : {1, 2, 37, 74, 147}) { cudaMemset(buf, 0, 1 << 20); pingpong<M><<<148, 32>>>(buf, buf + 1024, out, other); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); printf("mode %d pingpong blk0<->blk%d: %lld cycles per round trip (2 hops)\n", M, other, h[0]); } \
    cudaMemset(buf, 0, 1 << 20); own<M><<<148, 32>>>(buf, out); cudaMemcpy(h, out, 8 * 148, cudaMemcpyDeviceToHost); { long long mn = 1 << 30, mx = 0; for (int i = 0; i < 148; i++) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; } printf("mode %d own store->load: min %lld max %lld cycles\n", M, mn, mx); } \
    for (int nb : {8, 65, 148}) { cudaMemset(buf, 0, 1 << 20); fanout<M><<<nb, 32>>>(buf, buf + 4096, out, nb); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); printf("mode %d fanout %d blocks: %lld cycles per flag->all acks\n", M, nb, h[0]); }
    RUN(0) RUN(1) RUN(2) RUN(3)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

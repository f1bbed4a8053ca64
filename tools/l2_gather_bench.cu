// l2_gather_bench.cu -- the communication pattern of k_minimizer_persist in isolation, for trying hand-over variants
// without the rest of the pipeline (round-1 finding: ~2/3 of an evaluation is hand-over latency; request hop ~500 ns,
// gather of 65 blocks x 59 slots ~2.2-3.5 us; see profiles/r1_summary.md section 4).
//
// NB blocks; per round every block publishes W 8-byte {data32, seq32} slots, block 0 gathers them with variant V and
// publishes a 34-word "request" that the others poll before the next round.  Reports cycles per round.
//   V=0  thread b polls all W slots of block b (what the product does)
//   V=1  G threads per block share its slots (thread g*NB + b takes slots [g*W/G, (g+1)*W/G))
//   V=2  two levels: every 8th block first gathers its group of 8 and republishes the group's (summed) slots; block 0
//        gathers the leaders only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_bench tools/l2_gather_bench.cu
// Run  : ./l2_gather_bench [NB=65] [rounds=2000] [work_cycles=4000]   (work = simulated per-round keyline work)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define W 59
#define T 256
#define REQ 34

__device__ __forceinline__ unsigned long long ldv(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stv(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void spin_work(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {
    }
}

// slots[w * T + b], lead[w * T + g], req[k]
template <int V, int G>
__global__ void __launch_bounds__(T) k_rounds(unsigned long long *slots, unsigned long long *lead, unsigned long long *req,
                                              int nb, int rounds, int work, long long *out, unsigned int *sink) {
    const int tid = threadIdx.x, b = blockIdx.x;
    __shared__ unsigned int acc_s[T];
    unsigned int acc = 0;
    acc_s[tid] = 0;
    __syncthreads();
    const long long t_begin = clock64();
    for (int r = 1; r <= rounds; r++) {
        const unsigned int seq = (unsigned int)r;
        // ---- wait for the request of this round (block 0 owns it) ----
        if (b != 0 && r > 1 && tid < REQ) {
            while ((unsigned int)(ldv(req + tid) >> 32) != seq - 1) {
            }
        }
        __syncthreads();
        if (tid == 0) spin_work(work);   // "keylines + block sums"
        __syncthreads();
        if (tid < W) stv(slots + (size_t)tid * T + b, ((unsigned long long)seq << 32) | (unsigned int)(b * 131 + tid));
        if (V == 2 && (b % 8) == 0 && b != 0) {   // group leader: gather own group, republish
            if (tid < 8 * 8) {
                // 8 threads per member, ~8 slots each
                const int mb = b + (tid & 7), part = tid >> 3;
                if (mb < nb) {
                    for (int w = part * 8; w < W && w < part * 8 + 8; w++) {
                        unsigned long long x;
                        while ((unsigned int)((x = ldv(slots + (size_t)w * T + mb)) >> 32) != seq) {
                        }
                        atomicAdd(&acc_s[w], (unsigned int)x);
                    }
                }
            }
            __syncthreads();
            if (tid < W) {
                stv(lead + (size_t)tid * T + b / 8, ((unsigned long long)seq << 32) | acc_s[tid]);
                acc_s[tid] = 0;
            }
        }
        if (b != 0) continue;
        // ---- block 0: gather ----
        if (V == 0) {
            if (tid < nb) {
                bool ok;
                do {
                    ok = true;
                    unsigned int a = 0;
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        const unsigned long long x = ldv(slots + (size_t)w * T + tid);
                        ok = ok && (unsigned int)(x >> 32) == seq;
                        a += (unsigned int)x;
                    }
                    if (ok) acc += a;
                } while (!ok);
            }
        } else if (V == 1) {
            const int g = tid / nb, bb = tid - g * nb;
            if (g < G) {
                const int w0 = g * ((W + G - 1) / G), w1 = min(W, w0 + (W + G - 1) / G);
                bool ok;
                do {
                    ok = true;
                    unsigned int a = 0;
                    for (int w = w0; w < w1; w++) {
                        const unsigned long long x = ldv(slots + (size_t)w * T + bb);
                        ok = ok && (unsigned int)(x >> 32) == seq;
                        a += (unsigned int)x;
                    }
                    if (ok) acc += a;
                } while (!ok);
            }
        } else {
            const int ng = (nb + 7) / 8;   // leaders 8, 16, ... publish in lead[.][1..]; group 0 = blocks 0..7 read directly
            if (tid < 8 * 8) {
                const int mb = (tid & 7), part = tid >> 3;
                if (mb < nb)
                    for (int w = part * 8; w < W && w < part * 8 + 8; w++) {
                        unsigned long long x;
                        while ((unsigned int)((x = ldv(slots + (size_t)w * T + mb)) >> 32) != seq) {
                        }
                        acc += (unsigned int)x;
                    }
            } else if (tid - 64 < (ng - 1) * 3) {
                const int g = 1 + (tid - 64) / 3, part = (tid - 64) % 3;
                for (int w = part * 20; w < W && w < part * 20 + 20; w++) {
                    unsigned long long x;
                    while ((unsigned int)((x = ldv(lead + (size_t)w * T + g)) >> 32) != seq) {
                    }
                    acc += (unsigned int)x;
                }
            }
        }
        __syncthreads();
        if (tid == 0) spin_work(2500);   // "LM step + pose"
        __syncthreads();
        if (tid < REQ) stv(req + tid, ((unsigned long long)seq << 32) | (unsigned int)tid);
    }
    if (b == 0) {
        acc_s[tid] = acc;
        __syncthreads();
        if (tid == 0) {
            unsigned int s = 0;
            for (int i = 0; i < T; i++) s += acc_s[i];
            *sink = s;
            out[0] = (clock64() - t_begin) / rounds;
        }
    }
}

template <int V, int G>
static void run(const char *name, int nb, int rounds, int work) {
    unsigned long long *slots, *lead, *req;
    long long *out, h = 0;
    unsigned int *sink;
    cudaMalloc(&slots, sizeof(unsigned long long) * 64 * T);
    cudaMalloc(&lead, sizeof(unsigned long long) * 64 * T);
    cudaMalloc(&req, sizeof(unsigned long long) * 64);
    cudaMalloc(&out, 64);
    cudaMalloc(&sink, 64);
    cudaMemset(slots, 0, sizeof(unsigned long long) * 64 * T);
    cudaMemset(lead, 0, sizeof(unsigned long long) * 64 * T);
    cudaMemset(req, 0, sizeof(unsigned long long) * 64);
    k_rounds<V, G><<<nb, T>>>(slots, lead, req, nb, rounds, work, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %6lld cycles/round (of which %d simulated work)  %s\n", name, h, work + 2500, cudaGetErrorString(e));
    cudaFree(slots);
    cudaFree(lead);
    cudaFree(req);
    cudaFree(out);
    cudaFree(sink);
}

int main(int argc, char **argv) {
    const int nb = argc > 1 ? atoi(argv[1]) : 65, rounds = argc > 2 ? atoi(argv[2]) : 2000, work = argc > 3 ? atoi(argv[3]) : 4000;
    if (nb < 1 || nb > 148) {
        printf("NB must be 1..148 (all blocks must be co-resident)\n");
        return 1;
    }
    printf("NB=%d rounds=%d\n", nb, rounds);
    run<0, 1>("V0 thread-per-block, 59 slots each", nb, rounds, work);
    if (nb * 2 <= T) run<1, 2>("V1 two threads per block", nb, rounds, work);
    if (nb * 3 <= T) run<1, 3>("V1 three threads per block", nb, rounds, work);
    run<2, 1>("V2 two-level (groups of 8)", nb, rounds, work);
    return 0;
}

// lsu_bench.cu -- per-key costs of the primitives a shared-memory bucket sort is built from, on this GPU:
// shared atomics with / without return, scattered shared stores, L2 gathers, scattered global stores to
// append streams, warp-ballot ranking.  Prints cycles per key per SM (148 SMs, nominal 1.965 GHz).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o lsu_bench lsu_bench.cu && ./lsu_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

constexpr int T = 512;
constexpr int ITEMS = 16;

// mode 0: ATOMS no return, 1: ATOMS return, 2: plain LDS+STS read-modify-write (racy, cost only),
// 3: scattered STS, 4: scattered LDS
template <int MODE, int BINS>
__global__ void __launch_bounds__(T, 2) smem_kernel(const uint32_t *__restrict__ keys, uint32_t *out, int reps)
{
    extern __shared__ uint32_t s[];
    for (int i = threadIdx.x; i < BINS; i += T) s[i] = 0;
    __syncthreads();
    uint32_t k[ITEMS];
    const int64_t base = (int64_t)blockIdx.x * T * ITEMS;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) k[j] = keys[base + j * T + threadIdx.x] & (BINS - 1);
    uint32_t acc = 0;
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            if (MODE == 0) atomicAdd(&s[k[j]], 1u);
            if (MODE == 1) acc += atomicAdd(&s[k[j]], 1u);
            if (MODE == 2) { uint32_t v = s[k[j]]; s[k[j]] = v + 1; acc += v; }
            if (MODE == 3) s[k[j]] = acc + j;
            if (MODE == 4) acc += s[k[j]];
        }
        if (MODE == 4 || MODE == 2) { __syncwarp(); }
    }
    __syncthreads();
    if (acc == 0x12345678u || MODE == 0 || MODE == 3) out[blockIdx.x * T + threadIdx.x] = acc + s[threadIdx.x & (BINS - 1)];
}

// L2 gather: 2 loads (w, w+1) from a table of `words` words at a random position per key
__global__ void __launch_bounds__(T, 2) gather_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ table,
                                                      uint32_t words, uint32_t *out, int two)
{
    const int64_t base = (int64_t)blockIdx.x * T * ITEMS;
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        uint32_t p = keys[base + j * T + threadIdx.x] % words;
        uint32_t a = __ldg(table + p);
        uint32_t b = two ? __ldg(table + p + 1) : 0u;
        acc += __funnelshift_l(b, a, p & 31);
    }
    out[blockIdx.x * T + threadIdx.x] = acc;
}

// scattered global store to 8192 append streams (cursor in shared memory), returning atomics
template <int RET>
__global__ void __launch_bounds__(T, 2) scatter_kernel(const uint32_t *__restrict__ keys, uint32_t *__restrict__ dst,
                                                       uint32_t stream_len, int tiles)
{
    __shared__ uint32_t s_c[8192];
    for (int i = threadIdx.x; i < 8192; i += T) s_c[i] = i * stream_len + blockIdx.x * (stream_len / tiles);
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * T * ITEMS;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        uint32_t k = keys[base + j * T + threadIdx.x] & 8191u;
        if (RET) {
            uint32_t slot = atomicAdd(&s_c[k], 1u);
            dst[slot] = (uint32_t)(base + j);
        } else {
            dst[k * stream_len + (blockIdx.x * ITEMS + j) % stream_len] = (uint32_t)(base + j);   // no atomics: store cost only
        }
    }
}

// ballot ranking on 8 bits (as rsort::onesweep does), warp-private counters updated by the leaders without atomics
__global__ void __launch_bounds__(T, 2) ballot_kernel(const uint32_t *__restrict__ keys, uint32_t *out, int nbits, int atomic)
{
    __shared__ uint32_t s_h[T / 32][256];
    for (int i = threadIdx.x; i < (T / 32) * 256; i += T) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * T * ITEMS;
    uint32_t *mine = s_h[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        uint32_t d = keys[base + j * T + threadIdx.x] & ((1u << nbits) - 1u);
        unsigned pm = 0xffffffffu;
        for (int b = 0; b < nbits; b++) {
            unsigned bm = __ballot_sync(0xffffffffu, (d >> b) & 1u);
            pm &= ((d >> b) & 1u) ? bm : ~bm;
        }
        int leader = __ffs(pm) - 1;
        uint32_t before = 0;
        if (atomic) {
            if ((int)lane == leader) before = atomicAdd(&mine[d], (uint32_t)__popc(pm));
        } else {
            if ((int)lane == leader) { before = mine[d]; mine[d] = before + __popc(pm); }
            __syncwarp();
        }
        before = __shfl_sync(0xffffffffu, before, leader);
        acc += before + __popc(pm & lt);
    }
    out[blockIdx.x * T + threadIdx.x] = acc;
}

int main()
{
    const int64_t n = 46'710'784 / (T * ITEMS) * (T * ITEMS);
    const int blocks = (int)(n / (T * ITEMS));
    uint32_t *keys, *out, *table, *dst;
    CK(cudaMalloc(&keys, n * 4));
    CK(cudaMalloc(&out, (size_t)blocks * T * 4 + 4096));
    const uint32_t words = 3'000'000;   // 12 MB, as a chr21-sized packed text
    CK(cudaMalloc(&table, (words + 8) * 4));
    CK(cudaMalloc(&dst, (n + (1 << 20)) * 4));
    uint32_t *h = (uint32_t *)malloc(n * 4);
    uint32_t x = 12345;
    for (int64_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; h[i] = (x >> 7) ^ (x << 11); }
    CK(cudaMemcpy(keys, h, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(table, 1, (words + 8) * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char *name, float ms, double keys_done) {
        double cyc = ms * 1e-3 * 1.965e9 * 148 / keys_done;
        printf("%-44s %8.3f ms  %6.2f Gkeys/s  %6.3f cyc/key/SM\n", name, ms, keys_done / ms / 1e6, cyc);
    };
#define RUN(name, keys_done, ...) do { __VA_ARGS__; CK(cudaDeviceSynchronize()); cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1); report(name, ms, keys_done); } while (0)
    const int reps = 4;
    CK(cudaFuncSetAttribute(smem_kernel<0, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(smem_kernel<1, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    RUN("ATOMS no return, 4096 bins", (double)n * reps, (smem_kernel<0, 4096><<<blocks, T, 4096 * 4>>>(keys, out, reps)));
    RUN("ATOMS return, 4096 bins", (double)n * reps, (smem_kernel<1, 4096><<<blocks, T, 4096 * 4>>>(keys, out, reps)));
    RUN("ATOMS no return, 256 bins", (double)n * reps, (smem_kernel<0, 256><<<blocks, T, 256 * 4>>>(keys, out, reps)));
    RUN("ATOMS return, 256 bins", (double)n * reps, (smem_kernel<1, 256><<<blocks, T, 256 * 4>>>(keys, out, reps)));
    RUN("ATOMS return, 16384 bins", (double)n * reps, (smem_kernel<1, 16384><<<blocks, T, 16384 * 4>>>(keys, out, reps)));
    RUN("LDS+STS rmw, 4096 bins", (double)n * reps, (smem_kernel<2, 4096><<<blocks, T, 4096 * 4>>>(keys, out, reps)));
    RUN("scattered STS, 4096 words", (double)n * reps, (smem_kernel<3, 4096><<<blocks, T, 4096 * 4>>>(keys, out, reps)));
    RUN("scattered LDS, 4096 words", (double)n * reps, (smem_kernel<4, 4096><<<blocks, T, 4096 * 4>>>(keys, out, reps)));
    RUN("keys read only (1 gather word=0)", (double)n, (gather_kernel<<<blocks, T>>>(keys, table, 1, out, 0)));
    RUN("L2 gather 1 word / key, 12 MB table", (double)n, (gather_kernel<<<blocks, T>>>(keys, table, words, out, 0)));
    RUN("L2 gather 2 words / key, 12 MB table", (double)n, (gather_kernel<<<blocks, T>>>(keys, table, words, out, 1)));
    const uint32_t stream_len = (uint32_t)(n / 8192) + 64;
    RUN("scatter STG, 8192 streams, ATOMS cursor", (double)n, (scatter_kernel<1><<<blocks, T>>>(keys, dst, stream_len, blocks)));
    RUN("scatter STG, 8192 streams, no atomics", (double)n, (scatter_kernel<0><<<blocks, T>>>(keys, dst, stream_len, blocks)));
    RUN("ballot rank 8 bits, leader ATOMS", (double)n, (ballot_kernel<<<blocks, T>>>(keys, out, 8, 1)));
    RUN("ballot rank 8 bits, leader LDS/STS", (double)n, (ballot_kernel<<<blocks, T>>>(keys, out, 8, 0)));
    RUN("ballot rank 4 bits, leader LDS/STS", (double)n, (ballot_kernel<<<blocks, T>>>(keys, out, 4, 0)));
    return 0;
}

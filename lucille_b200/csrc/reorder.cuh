// K6 of the survey's kernel list: ray reordering -- built, measured, and left OFF by default.
//
// On the L2-resident scene ray order buys nothing (scripts/exp_sort.py on C3: batch / octant-major / Morton / random = 948 / 953 /
// 947 / 947 Mrays/s): a lane reads its own 64-byte record at every step whatever its neighbours do, and every record is an L2 hit
// anyway.  On the 10 M-triangle scene (1.2 GB of records, L2 hit 92 %) the order decides what the other 8 % cost: with the batch
// permuted on the HOST for free, octant-major order over the WHOLE batch (all rays of one direction octant together, neighbouring
// shading points adjacent) gives 981 -> 1060 Mrays/s, while grouping by octant inside small neighbourhoods (8 / 32 points) or by
// Morton cell does nothing and a random order costs 3 %.  That +8 % is the upper bound for a device sort, which has to be paid for:
// a STABLE counting sort on the 3-bit octant -- per-tile histograms, one scan over (octant, tile), scatter by tile base + rank inside
// the tile -- into a scratch batch, plus the permutation through which the traverser writes its results back in input order.
// (Stability matters: a first version with warp-aggregated atomic cursors, i.e. positions in the order the CTAs happened to run,
// interleaved points thousands apart inside an octant and LOST 6 %.)  Measured inside the pipelines on the 10 M-triangle scene
// (scripts/gpu_r3t.sh): point entry, 16 Mi rays, generation + sort + traversal 18.19 -> 17.95 ms (+1.3 %); a 2048^2 AO frame 126.9 ->
// 129.3 ms (-2 %): the three passes move 1.1 GB per 16 Mi rays and the per-sample counters lose their locality, which eats the gain.
// So the kernels stay in the tree behind B200_K6=1 (parity-tested: tests/test_gpu_parity.py::test_k6_reordered_batches_...), and the
// default paths trace batches in the order their generators write them -- point-major, neighbouring pixels adjacent, which is already
// within 8 % of the best order found.  Generating the rays in octant-major order directly (two generator passes instead of a sort)
// is what would keep most of the 8 %.
#pragma once

namespace b200 {

__device__ __forceinline__ uint32_t ray_octant(const float4 d) { return (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u); }

constexpr uint32_t kK6Tile = 4096;                   // rays per CTA: 256 threads x 16 rounds

// pass 1: per-tile histogram, hist[o * ntiles + tile]
__global__ void __launch_bounds__(256)
k6_hist_kernel(const float4 *__restrict__ rays, const uint32_t n, const uint32_t ntiles, unsigned int *__restrict__ hist)
{
    __shared__ unsigned int s_h[8];
    if (threadIdx.x < 8) s_h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t t0 = blockIdx.x * kK6Tile;
    unsigned int mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t r = 0; r < 16u; ++r) {
        const uint32_t i = t0 + r * 256u + threadIdx.x;
        const uint32_t oct = i < n ? ray_octant(__ldg(rays + 2 * (size_t)i + 1)) : 8u;
#pragma unroll
        for (uint32_t o = 0; o < 8u; ++o) mine[o] += (unsigned)__popc(__ballot_sync(0xffffffffu, oct == o));
    }
    if ((threadIdx.x & 31u) == 0u) {
#pragma unroll
        for (uint32_t o = 0; o < 8u; ++o) if (mine[o]) atomicAdd(&s_h[o], mine[o]);
    }
    __syncthreads();
    if (threadIdx.x < 8) hist[threadIdx.x * ntiles + blockIdx.x] = s_h[threadIdx.x];
}

// pass 2: exclusive scan of the 8 * ntiles counters in octant-major order, in place (one CTA; 8 * ntiles <= 2^20 for a 2^31-ray batch / 4096)
__global__ void __launch_bounds__(1024)
k6_scan_kernel(unsigned int *__restrict__ hist, const uint32_t m)
{
    __shared__ unsigned int s_part[1024];
    const uint32_t per = (m + 1023u) / 1024u, lo = threadIdx.x * per, hi = lo + per < m ? lo + per : m;
    unsigned int sum = 0;
    for (uint32_t k = lo; k < hi; ++k) sum += hist[k];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) {
        const unsigned int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = s_part[threadIdx.x] - sum;
    for (uint32_t k = lo; k < hi; ++k) { const unsigned int c = hist[k]; hist[k] = run; run += c; }
}

// pass 3: stable scatter -- a ray's place = its tile's base for its octant + the number of same-octant rays before it in the tile
__global__ void __launch_bounds__(256)
k6_scatter_kernel(const float4 *__restrict__ rays, const uint32_t n, const uint32_t ntiles, const unsigned int *__restrict__ base,
                  float4 *__restrict__ sorted, uint32_t *__restrict__ perm)
{
    __shared__ unsigned int s_run[8], s_warp[8][8];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x < 8) s_run[threadIdx.x] = base[threadIdx.x * ntiles + blockIdx.x];
    __syncthreads();
    const uint32_t t0 = blockIdx.x * kK6Tile;
    for (uint32_t r = 0; r < 16u; ++r) {
        const uint32_t i = t0 + r * 256u + threadIdx.x;
        float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d = a;
        uint32_t oct = 8u;
        if (i < n) { a = __ldg(rays + 2 * (size_t)i); d = __ldg(rays + 2 * (size_t)i + 1); oct = ray_octant(d); }
        unsigned my_mask = 0;
#pragma unroll
        for (uint32_t o = 0; o < 8u; ++o) {
            const unsigned m = __ballot_sync(0xffffffffu, oct == o);
            if (lane == o) s_warp[warp][o] = (unsigned)__popc(m);
            if (oct == o) my_mask = m;
        }
        __syncthreads();
        if (oct < 8u) {
            unsigned before = 0;
            for (unsigned w = 0; w < warp; ++w) before += s_warp[w][oct];
            const uint32_t pos = s_run[oct] + before + (uint32_t)__popc(my_mask & ((1u << lane) - 1u));
            sorted[2 * (size_t)pos] = a; sorted[2 * (size_t)pos + 1] = d;
            perm[pos] = i;
        }
        __syncthreads();
        if (threadIdx.x < 8) { unsigned tot = 0; for (unsigned w = 0; w < 8; ++w) tot += s_warp[w][threadIdx.x]; s_run[threadIdx.x] += tot; }
        __syncthreads();
    }
}

}  // namespace b200

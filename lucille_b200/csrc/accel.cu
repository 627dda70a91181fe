// libb200accel.so -- CUDA kernels (sm_100a) and the C ABI declared in include/lucille_b200.h.
//
// Compile with:  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3
// (-fmad=false is part of the parity contract: see trace.cuh).
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/lucille_b200.h"
#include "bvh_build.h"
#include "trace.cuh"

using namespace b200;

// ------------------------------------------------------------------------------------------------
// error channel + launch accounting
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// ------------------------------------------------------------------------------------------------
// the accelerator object
// ------------------------------------------------------------------------------------------------
struct ri_b200_accel {
    HostTree   tree;
    FlatTree   flat;          // host copy of the flat records is dropped after upload
    int        device = 0;
    uint32_t   precisions = 0;
    int        sm_count = 148;
    cudaStream_t stream = nullptr, copy_stream[2] = {nullptr, nullptr};
    cudaEvent_t  ev[8] = {};
    Node32 *d_nodes32 = nullptr; Tri32 *d_tris32 = nullptr;
    Node64 *d_nodes64 = nullptr; Tri64 *d_tris64 = nullptr;
    Tri32 *d_tris32t = nullptr; Tri64 *d_tris64t = nullptr;   // leaf-transposed copies (pool.cuh)
    uint32_t *d_slot_of_prim = nullptr;
    double *d_nrm64 = nullptr; float *d_nrm32 = nullptr;     // optional vertex normals, [prim][9]
    float *d_tex = nullptr; int tex_w = 0, tex_h = 0;          // optional material texture, [h][w][4] floats
    double *d_col = nullptr, *d_st = nullptr; uint8_t *d_attr_flags = nullptr;   // optional vertex colours [prim][9], st [prim][6], flags (1 colour, 2 st, 4 inside)
    uint64_t device_bytes = 0;
    double   upload_seconds = 0.0;
    // staging for host-buffer batches (double buffered)
    void    *d_in[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
    uint64_t stage_in_bytes = 0, stage_out_bytes = 0;
    void    *d_whole_in = nullptr, *d_whole_out = nullptr;     // whole-batch staging of the streamed occlusion path
    uint64_t whole_in_bytes = 0, whole_out_bytes = 0;
    unsigned long long *d_counters = nullptr;
    // MT19937 jump-ahead: polynomial table and cached window states at segment starts (frame.cuh)
    uint32_t *d_mt_polys = nullptr, *d_mt_states = nullptr;
    uint32_t  mt_states_cap = 0, mt_states_seed = 0;
    ri_b200_hit_exchange_fn hit_exchange = nullptr;      // rng_mode 0 over several ranks (frame.cuh)
    void     *hit_exchange_user = nullptr;
    float4 *d_k6_rays = nullptr; uint32_t *d_k6_perm = nullptr; unsigned int *d_k6_ctr = nullptr; uint64_t k6_cap = 0;   // reorder.cuh scratch (lock held)
    uint32_t *pixc_pix = nullptr; uint64_t pixc_cap = 0, pixc_n = 0; int pixc_w = 0, pixc_h = 0, pixc_bucket = 0, pixc_rank = 0, pixc_world = 0;   // frame.cuh: cached pixel list
    bool verts_f32 = false;               // every vertex coordinate is an fp32 number (hybrid.cuh: no absolute error in the fp32 slots)
    Node32 *d_nodesH = nullptr; Tri32 *d_trisH = nullptr;      // hybrid.cuh's own fp32 records, relative to hyb_c (NULL: it reads the shared ones)
    double hyb_c[3] = {0.0, 0.0, 0.0};
    bool streamed_faulted = false;        // host_batch: the streamed upload timed out once -> launch-per-piece from then on
    unsigned int *d_work = nullptr;      // ring of work counters for the persistent kernels
    std::atomic<unsigned> work_slot{0};   // the _dev entry points may be called from several host threads / streams
    // single-ray path
    void *h_pin = nullptr, *d_one = nullptr;
    std::mutex mu;
    // frame scratch (grown on demand)
    void *d_frame[12] = {};
    uint64_t frame_bytes[12] = {};
};

template <typename Real> static SceneView<Real> make_view(const ri_b200_accel *a);

template <> SceneView<float> make_view<float>(const ri_b200_accel *a)
{
    SceneView<float> v;
    v.nodes = a->d_nodes32; v.tris = a->d_tris32; v.slot_of_prim = a->d_slot_of_prim; v.normals = a->d_nrm32;
    for (int k = 0; k < 3; ++k) { v.smin[k] = a->flat.smin32[k]; v.smax[k] = a->flat.smax32[k]; }
    v.root_word = a->flat.root_word; v.top_count = a->flat.top_count;
    return v;
}
template <> SceneView<double> make_view<double>(const ri_b200_accel *a)
{
    SceneView<double> v;
    v.nodes = a->d_nodes64; v.tris = a->d_tris64; v.slot_of_prim = a->d_slot_of_prim; v.normals = a->d_nrm64;
    for (int k = 0; k < 3; ++k) { v.smin[k] = a->tree.bmin[k]; v.smax[k] = a->tree.bmax[k]; }
    v.root_word = a->flat.root_word; v.top_count = a->flat.top_count;
    return v;
}

// ------------------------------------------------------------------------------------------------
// ray-batch kernels
// ------------------------------------------------------------------------------------------------
constexpr int kBlock = 256;
constexpr unsigned kWorkRing = 1024;     // work counters of the persistent kernels: one per launch, reused round-robin

template <typename Real> struct RayIO;
template <> struct RayIO<float> {
    using Hit = ri_b200_hit_f32;
    static constexpr int kRayStride = 8;
    static __device__ __forceinline__ void load(const float *rays, uint64_t i, float org[3], float dir[3])
    {
        const float4 *p = reinterpret_cast<const float4 *>(rays) + 2 * i;
        const float4 a = __ldg(p), b = __ldg(p + 1);
        org[0] = a.x; org[1] = a.y; org[2] = a.z;
        dir[0] = b.x; dir[1] = b.y; dir[2] = b.z;
    }
    // a buffer the copy engine is still writing while the kernel runs (streamed host batches): ld.global.cg, never the
    // non-coherent path -- ld.global.nc requires the data to be read-only for the kernel's lifetime
    static __device__ __forceinline__ void load_coherent(const float *rays, uint64_t i, float org[3], float dir[3])
    {
        const float4 *p = reinterpret_cast<const float4 *>(rays) + 2 * i;
        const float4 a = __ldcg(p), b = __ldcg(p + 1);
        org[0] = a.x; org[1] = a.y; org[2] = a.z;
        dir[0] = b.x; dir[1] = b.y; dir[2] = b.z;
    }
    static __device__ __forceinline__ void store(Hit *out, uint64_t i, bool hit, float t, float u, float v, uint32_t prim)
    {
        float4 r;
        r.x = hit ? t : 1.0e38f; r.y = hit ? u : 0.0f; r.z = hit ? v : 0.0f;
        r.w = __uint_as_float(hit ? prim : 0xffffffffu);
        reinterpret_cast<float4 *>(out)[i] = r;
    }
};
template <> struct RayIO<double> {
    using Hit = ri_b200_hit_f64;
    static constexpr int kRayStride = 6;
    static __device__ __forceinline__ void load(const double *rays, uint64_t i, double org[3], double dir[3])
    {
        const double2 *p = reinterpret_cast<const double2 *>(rays) + 3 * i;
        const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        org[0] = a.x; org[1] = a.y; org[2] = b.x;
        dir[0] = b.y; dir[1] = c.x; dir[2] = c.y;
    }
    static __device__ __forceinline__ void load_coherent(const double *rays, uint64_t i, double org[3], double dir[3])
    {
        const double2 *p = reinterpret_cast<const double2 *>(rays) + 3 * i;
        const double2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
        org[0] = a.x; org[1] = a.y; org[2] = b.x;
        dir[0] = b.y; dir[1] = c.x; dir[2] = c.y;
    }
    static __device__ __forceinline__ void store(Hit *out, uint64_t i, bool hit, double t, double u, double v, uint32_t prim)
    {
        Hit h;
        h.t = hit ? t : 1.0e38; h.u = hit ? u : 0.0; h.v = hit ? v : 0.0;
        h.prim = hit ? prim : 0xffffffffu; h.hit = hit ? 1u : 0u;
        out[i] = h;
    }
};

template <typename Real, bool ANYHIT, bool COUNT>
__global__ void __launch_bounds__(kBlock)
trace_batch_kernel(const SceneView<Real> S, const Real *__restrict__ rays, const uint64_t n,
                   typename RayIO<Real>::Hit *__restrict__ hits, uint8_t *__restrict__ occ,
                   unsigned long long *__restrict__ counters)
{
    extern __shared__ uint32_t s_stack[];
    uint32_t *stk = s_stack + threadIdx.x;
    const uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    LaneCounters lc = {0, 0, 0, 0};
    bool active = i < n;
    if (active) {
        Real org[3], dir[3], t, u, v;
        uint32_t prim;
        RayIO<Real>::load(rays, i, org, dir);
        const bool hit = trace_ray<Real, ANYHIT, COUNT>(S, org, dir, stk, kBlock, t, u, v, prim, &lc);
        if (ANYHIT) occ[i] = hit ? 1 : 0;
        else RayIO<Real>::store(hits, i, hit, t, u, v, prim);
    }
    if (COUNT) {
        // the reference bumps nrays only for rays that reach ri_bvh_intersect on a non-empty tree (bvh.c:446,460)
        unsigned long long c[5] = {(active && S.root_word != kDoneWord) ? 1ull : 0ull, lc.ninner, lc.nleaf, lc.ntris, lc.nhit};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            unsigned long long x = c[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) == 0 && x) atomicAdd(&counters[k], x);
        }
    }
}

#include "persistent.cuh"
#include "packed.cuh"
#include "pool.cuh"
#include "pool32.cuh"
#include "reorder.cuh"
#include "pool_closest.cuh"
#include "hybrid.cuh"

static const char *pool_tris(const ri_b200_accel *a, float)  { return reinterpret_cast<const char *>(a->d_tris32t); }
static const char *pool_tris(const ri_b200_accel *a, double) { return reinterpret_cast<const char *>(a->d_tris64t); }

static int stack_capacity(const ri_b200_accel *a)
{
    int cap = a->tree.max_depth;
    if (cap < 1) cap = 1;
    return cap;
}

template <int kCap>
static void launch_closest32_cap(ri_b200_accel *a, const float *d_rays, uint32_t m, uint32_t chunk, ri_b200_hit_f32 *d_hits, unsigned int *ctr, cudaStream_t st);
static bool launch_closest32(ri_b200_accel *a, const float *d_rays, uint32_t m, uint32_t chunk, ri_b200_hit_f32 *d_hits, unsigned int *ctr, cudaStream_t st);
static bool launch_closest32(ri_b200_accel *, const double *, uint32_t, uint32_t, ri_b200_hit_f64 *, unsigned int *, cudaStream_t) { return false; }
static bool launch_closest_hybrid(ri_b200_accel *a, const double *d_rays, uint32_t m, uint32_t chunk, ri_b200_hit_f64 *d_hits, unsigned int *ctr, cudaStream_t st);
static bool launch_closest_hybrid(ri_b200_accel *, const float *, uint32_t, uint32_t, ri_b200_hit_f32 *, unsigned int *, cudaStream_t);

template <typename Real>
static void launch_closest_pool(ri_b200_accel *a, const Real *d_rays, uint32_t m, uint32_t chunk, typename RayIO<Real>::Hit *d_hits,
                                unsigned int *ctr, uint32_t refill_at, cudaStream_t st)
{
    if (launch_closest32(a, d_rays, m, chunk, d_hits, ctr, st)) return;       // fp32 on a tree that fits a static stack: pool32.cuh
    if (launch_closest_hybrid(a, d_rays, m, chunk, d_hits, ctr, st)) return;  // double rays, both record sets resident: hybrid.cuh
    const int cap = stack_capacity(a);
    const size_t smem = pool_closest_smem_bytes<Real>(cap);
    auto kern = closest_pool_kernel<Real>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPcThreads, smem);
    if (per_sm < 1) per_sm = 1;
    const uint64_t capb = (uint64_t)per_sm * (uint64_t)a->sm_count;
    uint64_t want = ((uint64_t)m + chunk - 1) / chunk;
    want = (want + (kPcThreads / 32) - 1) / (kPcThreads / 32);
    const unsigned blocks = (unsigned)(want < capb ? want : capb);
    kern<<<blocks, kPcThreads, smem, st>>>(make_view<Real>(a), pool_tris(a, Real(0)), d_rays, m, chunk, d_hits, ctr, refill_at, (uint32_t)cap, make_pack_k());
}

// pool32.cuh: fp32 occlusion with static shared memory.  Returns false when it does not apply (fp64 records, a tree deeper than
// the largest instantiated stack, B200_POOL32=0 or one of pool.cuh's experiment knobs) and the generic pooled kernel should run.
// Child order of the occlusion traversers (pool32.cuh has the measurements): longer-path-first while the fp32 records fit in L2 with
// room to spare, nearer-box-first beyond; B200_ANYHIT_ORDER=0|1|2 overrides (0 = the reference's order).
static int anyhit_order(const ri_b200_accel *a)
{
    if (const char *e = getenv("B200_ANYHIT_ORDER")) return atoi(e);
    const uint64_t bytes = (uint64_t)a->flat.ninner * sizeof(Node32) + a->flat.nslots * sizeof(Tri32);
    return bytes <= (96ull << 20) ? 1 : 2;
}

template <int kCap>
static void launch_pool32_cap(ri_b200_accel *a, const float *d_rays, uint32_t m, uint32_t chunk, uint8_t *d_occ, uint32_t *d_counts,
                              uint32_t rays_per_count, unsigned int *ctr, const unsigned int *d_ready, unsigned int *d_fault, unsigned blocks, cudaStream_t st,
                              const uint32_t *d_perm)
{
    (void)blocks;                       // the caller's count assumes its own CTA size: this kernel has its own
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, occluded_pool32_kernel<kCap, false, 1>, kOccThreads, 0);
    if (per_sm < 1) per_sm = 1;
    const uint64_t capb = (uint64_t)per_sm * (uint64_t)a->sm_count;
    uint64_t want = ((uint64_t)m + chunk - 1) / chunk;
    want = (want + (kOccThreads / 32) - 1) / (kOccThreads / 32);
    const unsigned nb = (unsigned)(want < capb ? want : capb);
    const int order = anyhit_order(a);
#define B200_P32_LAUNCH(C, O) occluded_pool32_kernel<kCap, C, O><<<nb, kOccThreads, 0, st>>>(make_view<float>(a), pool_tris(a, 0.0f), d_rays, m, chunk, \
        (C) ? nullptr : d_occ, (C) ? d_counts : nullptr, rays_per_count, ctr, d_ready, d_fault, make_pack_k(), d_perm)
    if (d_counts) { if (order == 1) B200_P32_LAUNCH(true, 1); else if (order == 2) B200_P32_LAUNCH(true, 2); else if (order == 3) B200_P32_LAUNCH(true, 3); else B200_P32_LAUNCH(true, 0); }
    else          { if (order == 1) B200_P32_LAUNCH(false, 1); else if (order == 2) B200_P32_LAUNCH(false, 2); else if (order == 3) B200_P32_LAUNCH(false, 3); else B200_P32_LAUNCH(false, 0); }
#undef B200_P32_LAUNCH
}
static bool pool32_applies(const ri_b200_accel *a)
{
    static const bool off = (getenv("B200_POOL32") && atoi(getenv("B200_POOL32")) == 0) || getenv("B200_POOL_TOPSMEM") || getenv("B200_REFILL") ||
                            getenv("B200_LEAF_AT") || getenv("B200_QUADFETCH");
    static const bool pool_off = getenv("B200_POOL") && atoi(getenv("B200_POOL")) == 0;
    return !off && !pool_off && stack_capacity(a) <= 36;
}
static bool launch_pool32(ri_b200_accel *a, const float *d_rays, uint32_t m, uint32_t chunk, uint8_t *d_occ, uint32_t *d_counts,
                          uint32_t rays_per_count, unsigned int *ctr, const unsigned int *d_ready, unsigned int *d_fault, unsigned blocks, cudaStream_t st,
                          const uint32_t *d_perm)
{
    const int cap = stack_capacity(a);
    if (!pool32_applies(a)) return false;
    if (cap <= 20) launch_pool32_cap<20>(a, d_rays, m, chunk, d_occ, d_counts, rays_per_count, ctr, d_ready, d_fault, blocks, st, d_perm);
    else if (cap <= 28) launch_pool32_cap<28>(a, d_rays, m, chunk, d_occ, d_counts, rays_per_count, ctr, d_ready, d_fault, blocks, st, d_perm);
    else launch_pool32_cap<36>(a, d_rays, m, chunk, d_occ, d_counts, rays_per_count, ctr, d_ready, d_fault, blocks, st, d_perm);
    return true;
}
static bool launch_pool32(ri_b200_accel *, const double *, uint32_t, uint32_t, uint8_t *, uint32_t *, uint32_t, unsigned int *,
                          const unsigned int *, unsigned int *, unsigned, cudaStream_t, const uint32_t *) { return false; }

// hybrid.cuh: double-exact occlusion through the fp32 records with certified decisions, the double records only where fp32 cannot
// decide.  Applies to double rays when BOTH record sets are resident and the tree fits a static stack; B200_HYBRID=0 turns it off.
static int anyhit_order(const ri_b200_accel *a);

static bool hybrid_applies(const ri_b200_accel *a)
{
    const char *env = getenv("B200_HYBRID");                      // read per launch: tests and A/B scripts switch it inside one process
    const int mode = env ? atoi(env) : 1;                         // 0: never (the double kernels run)
    return mode != 0 && stack_capacity(a) <= 28 && a->d_nodes64 && a->d_tris64 &&
           ((a->d_nodesH && a->d_trisH) || (a->d_nodes32 && a->d_tris32t));      // its own fp32 records, or the shared ones
}

// is the scene close enough to the world origin for the SHARED fp32 records (absolute coordinates) to give good bounds?
static bool hybrid_near_origin(const ri_b200_accel *a)
{
    double far = 0.0, ext = 0.0;
    for (int k = 0; k < 3; ++k) {
        far = std::fmax(far, std::fmax(std::fabs(a->tree.bmin[k]), std::fabs(a->tree.bmax[k])));
        ext = std::fmax(ext, a->tree.bmax[k] - a->tree.bmin[k]);
    }
    return far <= 2.0 * ext;
}

static HybK hybrid_consts(const ri_b200_accel *a)
{
    HybK H;
    float bm = 0.0f;
    const bool own = a->d_nodesH != nullptr;
    for (int k = 0; k < 3; ++k) {
        H.c[k] = own ? a->hyb_c[k] : 0.0;
        const double lo = own ? a->tree.bmin[k] - H.c[k] : (double)a->flat.smin32[k], hi = own ? a->tree.bmax[k] - H.c[k] : (double)a->flat.smax32[k];
        H.bmax[k] = std::nextafterf((float)std::fmax(std::fabs(lo), std::fabs(hi)), INFINITY);
        bm = std::fmax(bm, H.bmax[k]);
    }
    H.eta0 = (!own && a->verts_f32) ? 0.0f : 5.9604645e-8f * bm;   // own records: v0 - c is rounded once to fp32
    H.de = own ? 0.0f : 2.0f * H.eta0;                             // own records: edges are the double edges rounded once (relative error only)
    H.order = anyhit_order(a);
    return H;
}
static SceneView<float> hybrid_view(const ri_b200_accel *a)
{
    SceneView<float> S = make_view<float>(a);
    if (a->d_nodesH) S.nodes = a->d_nodesH;
    return S;
}
static const char *hybrid_tris(const ri_b200_accel *a)
{ return a->d_trisH ? reinterpret_cast<const char *>(a->d_trisH) : pool_tris(a, 0.0f); }
static bool hybrid_exact(const ri_b200_accel *a) { return a->verts_f32 && !a->d_nodesH; }

// The filter's own records (hybrid.cuh, end): built when the shared fp32 records would bound poorly -- vertices that are not fp32
// numbers (their slots carry an absolute error, and so do the edges subtracted in fp32) or a scene far from the world origin.
// Measured on the C3 scene (4 Mi double AO rays): non-fp32 vertices 797 -> see DESIGN.md; moved out to |coordinates| ~ 1000 the
// shared records lose to the double kernel (386 against 504 Mrays/s), the translated ones do not care where the scene sits.
static int hybrid_build_records(ri_b200_accel *a)
{
    if (a->tree.empty || !a->d_nodes64 || !a->d_tris64 || stack_capacity(a) > 28) return 0;
    const char *env = getenv("B200_HYBRID_OWN");                   // A/B: 0 = never build them, 1 = always
    const int force = env ? atoi(env) : -1;
    const bool shared_ok = a->d_nodes32 && a->d_tris32t && a->verts_f32 && hybrid_near_origin(a);     // the shared records serve as well
    if (force == 0 || (force != 1 && shared_ok)) return 0;
    // (an accelerator built with double records ONLY gets them too: its batched double queries then run at the hybrid rate instead of
    // the double kernels' -- + 48 bytes per triangle slot and 64 per node next to the 96 / 128 of the double records)
    for (int k = 0; k < 3; ++k) a->hyb_c[k] = (double)(float)(0.5 * (a->tree.bmin[k] + a->tree.bmax[k]));
    const uint32_t ninner = a->flat.ninner;
    const uint64_t nslots = a->flat.nslots;
    CUDA_OK(cudaMalloc((void **)&a->d_trisH, nslots * sizeof(Tri32)));
    CUDA_OK(cudaMemsetAsync(a->d_trisH, 0, nslots * sizeof(Tri32), a->stream));
    CUDA_OK(cudaMalloc((void **)&a->d_nodesH, (size_t)(ninner ? ninner : 1) * sizeof(Node32)));
    a->device_bytes += nslots * sizeof(Tri32) + (uint64_t)ninner * sizeof(Node32);
    if (ninner) {
        hyb_nodes_kernel<<<(ninner + 255) / 256, 256, 0, a->stream>>>(a->d_nodes64, ninner, a->hyb_c[0], a->hyb_c[1], a->hyb_c[2], a->d_nodesH);
        LAUNCHED();
        hyb_tris_kernel<<<(2 * ninner + 255) / 256, 256, 0, a->stream>>>(a->d_nodes64, 2 * ninner, 0u, a->d_tris64, a->hyb_c[0], a->hyb_c[1], a->hyb_c[2],
                                                                            reinterpret_cast<char *>(a->d_trisH));
        LAUNCHED();
    } else if (a->flat.root_word & kLeafFlag) {
        hyb_tris_kernel<<<1, 256, 0, a->stream>>>(a->d_nodes64, 1u, a->flat.root_word, a->d_tris64, a->hyb_c[0], a->hyb_c[1], a->hyb_c[2],
                                                  reinterpret_cast<char *>(a->d_trisH));
        LAUNCHED();
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(a->stream));
    return 0;
}

// double-exact closest hit through the fp32 records (hybrid.cuh, second kernel)
template <int kCap>
static void launch_closest_hybrid_cap(ri_b200_accel *a, const double *d_rays, uint32_t m, uint32_t chunk, ri_b200_hit_f64 *d_hits, unsigned int *ctr, cudaStream_t st)
{
    const size_t smem = sizeof(HybCSmem<kCap>);
    const HybK H = hybrid_consts(a);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kHcThreads, smem);
        if (per_sm < 1) per_sm = 1;
        const uint64_t capb = (uint64_t)per_sm * (uint64_t)a->sm_count;
        uint64_t want = ((uint64_t)m + chunk - 1) / chunk;
        want = (want + (kHcThreads / 32) - 1) / (kHcThreads / 32);
        const unsigned blocks = (unsigned)(want < capb ? want : capb);
        kern<<<blocks, kHcThreads, smem, st>>>(hybrid_view(a), make_view<double>(a), hybrid_tris(a), d_rays, m, chunk, d_hits, ctr, make_pack_k(), H);
    };
    if (hybrid_exact(a)) go(closest_hybrid_kernel<kCap, true>); else go(closest_hybrid_kernel<kCap, false>);
}
static bool launch_closest_hybrid(ri_b200_accel *a, const double *d_rays, uint32_t m, uint32_t chunk, ri_b200_hit_f64 *d_hits, unsigned int *ctr, cudaStream_t st)
{
    const char *env = getenv("B200_HYBRID_CLOSEST");
    if ((env && atoi(env) == 0) || !hybrid_applies(a)) return false;
    if (stack_capacity(a) <= 20) launch_closest_hybrid_cap<20>(a, d_rays, m, chunk, d_hits, ctr, st);
    else launch_closest_hybrid_cap<28>(a, d_rays, m, chunk, d_hits, ctr, st);
    return true;
}
static bool launch_closest_hybrid(ri_b200_accel *, const float *, uint32_t, uint32_t, ri_b200_hit_f32 *, unsigned int *, cudaStream_t) { return false; }

template <int kCap>
static void launch_hybrid_cap(ri_b200_accel *a, const double *d_rays, uint32_t m, uint32_t chunk, uint8_t *d_occ, uint32_t *d_counts,
                              uint32_t rays_per_count, unsigned int *ctr, const unsigned int *d_ready, unsigned int *d_fault, unsigned blocks, cudaStream_t st)
{
    const HybK H = hybrid_consts(a);
    const SceneView<float> S = hybrid_view(a);
    const SceneView<double> S64 = make_view<double>(a);
    const char *tt = hybrid_tris(a);
    const bool exact = hybrid_exact(a);
    const PackK K = make_pack_k();
    const size_t smem = sizeof(HybSmem<kCap>);
#define B200_HYB_LAUNCH(C, X) do { auto kern = occluded_hybrid_kernel<kCap, C, X>;                                               \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                      \
        kern<<<blocks, kHybThreads, smem, st>>>(S, S64, tt, d_rays, m, chunk, d_occ, d_counts, rays_per_count, ctr, d_ready, d_fault, K, H); } while (0)
    if (d_counts) { if (exact) B200_HYB_LAUNCH(true, true); else B200_HYB_LAUNCH(true, false); }
    else          { if (exact) B200_HYB_LAUNCH(false, true); else B200_HYB_LAUNCH(false, false); }
#undef B200_HYB_LAUNCH
}
static bool launch_hybrid(ri_b200_accel *a, const double *d_rays, uint32_t m, uint32_t chunk, uint8_t *d_occ, uint32_t *d_counts,
                          uint32_t rays_per_count, unsigned int *ctr, const unsigned int *d_ready, unsigned int *d_fault, unsigned, cudaStream_t st)
{
    if (!hybrid_applies(a)) return false;
    const int cap = stack_capacity(a);

    auto blocks_for = [&](const void *kern, size_t smem) -> unsigned {
        int per_sm = 0;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kHybThreads, smem);
        if (per_sm < 1) per_sm = 1;
        const uint64_t capb = (uint64_t)per_sm * (uint64_t)a->sm_count;
        uint64_t want = ((uint64_t)m + chunk - 1) / chunk;
        want = (want + (kHybThreads / 32) - 1) / (kHybThreads / 32);
        return (unsigned)(want < capb ? want : capb);
    };
    if (cap <= 20) launch_hybrid_cap<20>(a, d_rays, m, chunk, d_occ, d_counts, rays_per_count, ctr, d_ready, d_fault,
                                         blocks_for((const void *)occluded_hybrid_kernel<20, false, true>, sizeof(HybSmem<20>)), st);
    else launch_hybrid_cap<28>(a, d_rays, m, chunk, d_occ, d_counts, rays_per_count, ctr, d_ready, d_fault,
                               blocks_for((const void *)occluded_hybrid_kernel<28, false, true>, sizeof(HybSmem<28>)), st);
    return true;
}
static bool launch_hybrid(ri_b200_accel *, const float *, uint32_t, uint32_t, uint8_t *, uint32_t *, uint32_t, unsigned int *,
                          const unsigned int *, unsigned int *, unsigned, cudaStream_t) { return false; }

template <int kCap>
static void launch_closest32_cap(ri_b200_accel *a, const float *d_rays, uint32_t m, uint32_t chunk, ri_b200_hit_f32 *d_hits, unsigned int *ctr, cudaStream_t st)
{
    auto kern = closest_pool32_kernel<kCap>;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCloseThreads, 0);
    if (per_sm < 1) per_sm = 1;
    const uint64_t capb = (uint64_t)per_sm * (uint64_t)a->sm_count;
    uint64_t want = ((uint64_t)m + chunk - 1) / chunk;
    want = (want + (kCloseThreads / 32) - 1) / (kCloseThreads / 32);
    const unsigned blocks = (unsigned)(want < capb ? want : capb);
    kern<<<blocks, kCloseThreads, 0, st>>>(make_view<float>(a), pool_tris(a, 0.0f), d_rays, m, chunk, d_hits, ctr, make_pack_k());
}
static bool launch_closest32(ri_b200_accel *a, const float *d_rays, uint32_t m, uint32_t chunk, ri_b200_hit_f32 *d_hits, unsigned int *ctr, cudaStream_t st)
{
    static const bool off = (getenv("B200_POOL32") && atoi(getenv("B200_POOL32")) == 0) || getenv("B200_REFILL");
    const int cap = stack_capacity(a);
    if (off || cap > 24) return false;
    if (cap <= 20) launch_closest32_cap<20>(a, d_rays, m, chunk, d_hits, ctr, st);
    else launch_closest32_cap<24>(a, d_rays, m, chunk, d_hits, ctr, st);
    return true;
}

template <typename Real, bool ANYHIT, bool COUNT>
static int launch_trace(ri_b200_accel *a, const Real *d_rays, uint64_t n, typename RayIO<Real>::Hit *d_hits, uint8_t *d_occ,
                        unsigned long long *d_counters, cudaStream_t st, uint32_t *d_counts = nullptr, uint32_t rays_per_count = 1,
                        const unsigned int *d_ready = nullptr, unsigned int *d_fault = nullptr, const uint32_t *d_perm = nullptr)
{
    if (n == 0) return 0;
    const int cap = stack_capacity(a);
    const size_t smem = (size_t)cap * kBlock * sizeof(uint32_t);
    if (smem > 200 * 1024) return fail("BVH depth %d exceeds the shared-memory traversal stack", cap);
    if (!COUNT) {
        // production path: persistent warps with ray replacement; batches above 2^31 rays are split
        auto pk = trace_persistent_kernel<Real, ANYHIT>;
        // 4 CTAs x 256 threads per SM at 64 registers (fp32): measured 979 Mrays/s on C3 against 878 with 3 CTAs at 72 registers
        static const bool top_smem = getenv("B200_POOL_TOPSMEM") && atoi(getenv("B200_POOL_TOPSMEM")) != 0;      // experiment, fp32 only (pool.cuh)
        const bool use_top = top_smem && sizeof(Real) == 4;
        auto pool = use_top ? occluded_pool_kernel<Real, (sizeof(Real) == 4 ? 4 : 3), (sizeof(Real) == 4)> : occluded_pool_kernel<Real, (sizeof(Real) == 4 ? 4 : 3)>;
        static const bool use_pool = !(getenv("B200_POOL") && atoi(getenv("B200_POOL")) == 0);   // A/B knob: 0 = vote-scheduled kernel
        const bool pooled = ANYHIT && use_pool;
        static const bool use_pool_closest = !(getenv("B200_POOL_CLOSEST") && atoi(getenv("B200_POOL_CLOSEST")) == 0);
        static const bool use_pool_closest64 = !(getenv("B200_POOL_CLOSEST64") && atoi(getenv("B200_POOL_CLOSEST64")) == 0);
        const bool pooled_closest = !ANYHIT && use_pool_closest && (sizeof(Real) == 4 || use_pool_closest64) && d_hits != nullptr;
        static const uint32_t refill_at = getenv("B200_REFILL") ? (uint32_t)atoi(getenv("B200_REFILL")) : 4u;   // measured best of 1,4,8,16,24 on C3
        static const uint32_t leaf_at = getenv("B200_LEAF_AT") ? (uint32_t)atoi(getenv("B200_LEAF_AT")) : 32u;  // items waiting before a leaf round runs
        static const uint32_t quad_fetch = (getenv("B200_QUADFETCH") && atoi(getenv("B200_QUADFETCH")) != 0) ? 1u : 0u;   // pool.cuh: refill whole lane quads
        if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const size_t pool_smem = pool_smem_bytes<Real>(cap) + (use_top ? kTopNodes * sizeof(Node32) + 16 : 0);
        if (pool_smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(pool, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pool_smem));
        int per_sm = 0;
        if (pooled) CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pool, kBlock, pool_smem));
        else CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk, kBlock, smem));
        if (per_sm < 1) return fail("persistent traversal kernel does not fit on an SM");
        const uint64_t kMax = 1ull << 31;
        for (uint64_t done = 0; done < n; done += kMax) {
            const uint32_t m = (uint32_t)((n - done) < kMax ? (n - done) : kMax);
            const uint64_t cap = (uint64_t)per_sm * (uint64_t)a->sm_count;
            const uint64_t warps = cap * (kBlock / 32);
            // rays per atomic fetch: one warp-load.  Measured on the C3 batch (scripts/gpu_r3f.sh): 32 / 64 / 128 / 256 / 512 rays ->
            // 1113 / 1112 / 1106 / 1088 / 1053 Mrays/s -- the fetch is one atomic per 32 rays either way, a larger chunk only
            // lengthens the tail of the launch
            uint32_t chunk = 32u;
            (void)warps;
            if (const char *e = getenv("B200_RAYCHUNK")) { const int v = atoi(e); if (v >= 32) chunk = (uint32_t)v & ~31u; }   // A/B knob
            uint64_t want = ((uint64_t)m + chunk - 1) / chunk;             // one chunk per warp at least
            want = (want + (kBlock / 32) - 1) / (kBlock / 32);
            const unsigned blocks = (unsigned)(want < cap ? want : cap);
            unsigned int *ctr = a->d_work + (a->work_slot.fetch_add(1) & (kWorkRing - 1u));      // documented limit: kWorkRing launches in flight (lucille_b200.h)
            CUDA_OK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st));
            if (pooled_closest)
                launch_closest_pool<Real>(a, d_rays + done * RayIO<Real>::kRayStride, m, chunk, d_hits + done, ctr, refill_at, st);
            else if (pooled && launch_hybrid(a, d_rays + done * RayIO<Real>::kRayStride, m, chunk, d_occ ? d_occ + done : nullptr,
                                             d_counts ? d_counts + done / rays_per_count : nullptr, rays_per_count, ctr, d_ready, d_fault, blocks, st))
                ;                                // double rays, both record sets resident: certified fp32 decisions, doubles where needed (hybrid.cuh)
            else if (pooled && launch_pool32(a, d_rays + done * RayIO<Real>::kRayStride, m, chunk, d_occ ? d_occ + done : nullptr,
                                             d_counts ? d_counts + done / rays_per_count : nullptr, rays_per_count, ctr, d_ready, d_fault, blocks, st, d_perm))
                ;                                // fp32 occlusion on a tree that fits a static stack: the specialised kernel (pool32.cuh)
            else if (pooled)
                pool<<<blocks, kBlock, pool_smem, st>>>(make_view<Real>(a), pool_tris(a, Real(0)), d_rays + done * RayIO<Real>::kRayStride, m, chunk,
                                                      d_occ ? d_occ + done : nullptr, d_counts ? d_counts + done / rays_per_count : nullptr,
                                                      rays_per_count, ctr, refill_at | (leaf_at << 8) | (quad_fetch << 16), (uint32_t)stack_capacity(a), d_ready, d_fault, make_pack_k());
            else
                pk<<<blocks, kBlock, smem, st>>>(make_view<Real>(a), d_rays + done * RayIO<Real>::kRayStride, m, chunk,
                                               d_hits ? d_hits + done : nullptr, d_occ ? d_occ + done : nullptr,
                                               d_counts ? d_counts + done / rays_per_count : nullptr, rays_per_count, ctr, refill_at);
            LAUNCHED();
            CUDA_OK(cudaGetLastError());
        }
        return 0;
    }
    auto kern = trace_batch_kernel<Real, ANYHIT, COUNT>;
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t blocks = (n + kBlock - 1) / kBlock;
    if (blocks > 0x7fffffffull) return fail("batch too large");
    kern<<<(unsigned)blocks, kBlock, smem, st>>>(make_view<Real>(a), d_rays, n, d_hits, d_occ, d_counters);
    LAUNCHED();
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Occlusion batches of the entry points that hold the accelerator's lock (frames, point entries): with B200_K6=1 the batch is first put
// into octant-major order (reorder.cuh, K6) in the accelerator's scratch and traced through the permutation.
template <typename Real>
static int trace_occlusion_locked(ri_b200_accel *a, const Real *d_rays, uint64_t n, uint8_t *d_occ, uint32_t *d_counts, uint32_t rays_per_count,
                                  cudaStream_t st)
{ return launch_trace<Real, true, false>(a, d_rays, n, nullptr, d_occ, nullptr, st, d_counts, rays_per_count); }

template <>
int trace_occlusion_locked<float>(ri_b200_accel *a, const float *d_rays, uint64_t n, uint8_t *d_occ, uint32_t *d_counts, uint32_t rays_per_count,
                                  cudaStream_t st)
{
    const char *env = getenv("B200_K6");
    const int mode = env ? atoi(env) : -1;
    const bool sort = mode == 1;          // off by default: see reorder.cuh for the measurements that decided it
    if (!sort || !pool32_applies(a) || n >= (1ull << 31) || a->tree.empty)
        return launch_trace<float, true, false>(a, d_rays, n, nullptr, d_occ, nullptr, st, d_counts, rays_per_count);
    if (a->k6_cap < n) {
        cudaFree(a->d_k6_rays); cudaFree(a->d_k6_perm); a->d_k6_rays = nullptr; a->d_k6_perm = nullptr; a->k6_cap = 0;
        cudaFree(a->d_k6_ctr); a->d_k6_ctr = nullptr;
        const uint64_t nt = (n + kK6Tile - 1) / kK6Tile;
        if (cudaMalloc((void **)&a->d_k6_ctr, 8 * nt * sizeof(unsigned int)) != cudaSuccess ||
            cudaMalloc((void **)&a->d_k6_rays, n * 32) != cudaSuccess || cudaMalloc((void **)&a->d_k6_perm, n * 4) != cudaSuccess) {
            cudaGetLastError();                      // no room for the scratch batch: trace in input order
            cudaFree(a->d_k6_rays); a->d_k6_rays = nullptr;
            return launch_trace<float, true, false>(a, d_rays, n, nullptr, d_occ, nullptr, st, d_counts, rays_per_count);
        }
        a->k6_cap = n;
    }
    const uint32_t ntiles = (uint32_t)((n + kK6Tile - 1) / kK6Tile);
    k6_hist_kernel<<<ntiles, 256, 0, st>>>(reinterpret_cast<const float4 *>(d_rays), (uint32_t)n, ntiles, a->d_k6_ctr);
    LAUNCHED();
    k6_scan_kernel<<<1, 1024, 0, st>>>(a->d_k6_ctr, 8u * ntiles);
    LAUNCHED();
    k6_scatter_kernel<<<ntiles, 256, 0, st>>>(reinterpret_cast<const float4 *>(d_rays), (uint32_t)n, ntiles, a->d_k6_ctr, a->d_k6_rays, a->d_k6_perm);
    LAUNCHED();
    CUDA_OK(cudaGetLastError());
    return launch_trace<float, true, false>(a, reinterpret_cast<const float *>(a->d_k6_rays), n, nullptr, d_occ, nullptr, st, d_counts, rays_per_count,
                                            nullptr, nullptr, a->d_k6_perm);
}

// ------------------------------------------------------------------------------------------------
// build / free / info
// ------------------------------------------------------------------------------------------------
extern "C" const char *ri_b200_last_error(void) { return g_err; }

extern "C" int ri_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" uint64_t ri_b200_launch_count(void) { return g_launches.load(); }

template <typename T> static int upload(T **dst, const std::vector<T> &src, uint64_t &bytes)
{
    *dst = nullptr;
    if (src.empty()) return 0;
    CUDA_OK(cudaMalloc((void **)dst, src.size() * sizeof(T)));
    CUDA_OK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    bytes += src.size() * sizeof(T);
    return 0;
}

namespace b200 { struct DeviceBuild; }
static int build_on_device(ri_b200_accel *a, const double *tri_xyz, uint64_t ntris);      // bvh_build_gpu.cuh, end of this file

extern "C" ri_b200_accel_t *ri_b200_build(const double *tri_xyz, uint64_t ntris, uint32_t precisions, int device)
{
    if (ntris > kMaxTris) { fail("too many triangles (%llu > %llu)", (unsigned long long)ntris, (unsigned long long)kMaxTris); return nullptr; }
    if (!(precisions & (RI_B200_PREC_F32 | RI_B200_PREC_F64))) { fail("no precision requested"); return nullptr; }
    if (ntris && !tri_xyz) { fail("tri_xyz is NULL"); return nullptr; }
    const bool host_only = (precisions & RI_B200_HOST_ONLY) != 0;
    int ndev = 0;
    if (host_only) {
        ri_b200_accel *h = new (std::nothrow) ri_b200_accel();
        if (!h) { fail("out of memory"); return nullptr; }
        h->device = -1;
        h->precisions = precisions;
        build_tree(tri_xyz, ntris, h->tree);
        flatten_tree(h->tree, 1024, (precisions & RI_B200_PREC_F32) != 0, (precisions & RI_B200_PREC_F64) != 0, h->flat);
        if (h->flat.overflow) { fail("too many triangle slots for the 27-bit leaf word"); delete h; return nullptr; }
        return h;
    }
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail("no CUDA device: libb200accel has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= ndev) { fail("bad device %d", device); return nullptr; }

    ri_b200_accel *a = new (std::nothrow) ri_b200_accel();
    if (!a) { fail("out of memory"); return nullptr; }
    a->device = device;
    a->precisions = precisions;
    if (precisions & RI_B200_PREC_F64) {                                             // hybrid.cuh: are all vertex coordinates fp32 numbers?
        bool exact = true;
        for (uint64_t i = 0; i < 9 * ntris && exact; ++i) exact = (double)(float)tri_xyz[i] == tri_xyz[i];
        a->verts_f32 = exact;
    }

    // Where the tree is built: RI_B200_BUILD_DEVICE / RI_B200_BUILD_HOST say so; otherwise scenes of at least 32 Ki triangles are built
    // on the device (1 M triangles: 30 ms against 170 ms with 16 host threads, same tree) and small ones on the host (a device
    // build costs a dozen launches per level however small the scene).  B200_BUILD=host|device overrides the automatic choice.
    bool on_device = (precisions & RI_B200_BUILD_DEVICE) != 0;
    if (!on_device && !(precisions & RI_B200_BUILD_HOST)) {
        const char *env = getenv("B200_BUILD");
        on_device = env ? (strcmp(env, "device") == 0) : (ntris >= (1u << 15));
    }
    if (on_device) {
        // tree, flat node records AND the triangle-slot buffers are produced by the device path; body() below adds the rest
        if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice(%d) failed", device); delete a; return nullptr; }
        if (build_on_device(a, tri_xyz, ntris) != 0) {
            if (precisions & RI_B200_BUILD_DEVICE) { ri_b200_free(a); return nullptr; }       // asked for explicitly: report it
            cudaGetLastError();                                                                 // automatic choice: the host builder makes the same tree --
            fprintf(stderr, "[b200] device BVH build failed (%s); building the same tree on the host\n", g_err);   // said out loud, never silent
            cudaFree(a->d_tris32); cudaFree(a->d_tris64); cudaFree(a->d_tris32t); cudaFree(a->d_tris64t); cudaFree(a->d_slot_of_prim);
            a->d_tris32 = a->d_tris32t = nullptr; a->d_tris64 = a->d_tris64t = nullptr; a->d_slot_of_prim = nullptr; a->device_bytes = 0;
            on_device = false;
        }
    }
    if (!on_device) {
        build_tree(tri_xyz, ntris, a->tree);
        flatten_tree(a->tree, 1024, (precisions & RI_B200_PREC_F32) != 0, (precisions & RI_B200_PREC_F64) != 0, a->flat);
    }
    if (a->flat.overflow) { fail("too many triangle slots for the 27-bit leaf word"); ri_b200_free(a); return nullptr; }

    auto t0 = std::chrono::steady_clock::now();
    auto body = [&]() -> int {
        CUDA_OK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        a->sm_count = prop.multiProcessorCount;
        CUDA_OK(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&a->copy_stream[0], cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&a->copy_stream[1], cudaStreamNonBlocking));
        for (auto &e : a->ev) CUDA_OK(cudaEventCreate(&e));
        if (upload(&a->d_nodes32, a->flat.nodes32, a->device_bytes)) return -1;
        if (upload(&a->d_nodes64, a->flat.nodes64, a->device_bytes)) return -1;
        if (!on_device) {                       // the device builder has written these buffers itself
            if (upload(&a->d_tris32, a->flat.tris32, a->device_bytes)) return -1;
            if (upload(&a->d_tris64, a->flat.tris64, a->device_bytes)) return -1;
            if (upload(&a->d_tris32t, a->flat.tris32t, a->device_bytes)) return -1;
            if (upload(&a->d_tris64t, a->flat.tris64t, a->device_bytes)) return -1;
            if (upload(&a->d_slot_of_prim, a->flat.slot_of_prim, a->device_bytes)) return -1;
        }
        CUDA_OK(cudaMalloc((void **)&a->d_counters, 8 * sizeof(unsigned long long)));
        CUDA_OK(cudaMalloc((void **)&a->d_work, (kWorkRing + 2) * sizeof(unsigned int)));      // work counters, then the upload cursor and fault flag of the streamed host-buffer path
        CUDA_OK(cudaMallocHost(&a->h_pin, 4096));
        CUDA_OK(cudaMalloc(&a->d_one, 4096));
        return 0;
    };
    if (body() != 0) { ri_b200_free(a); return nullptr; }
    if ((precisions & RI_B200_PREC_F64) && hybrid_build_records(a) != 0) {
        // not fatal: without them the filter reads the shared fp32 records with the wider bounds that go with them (same answers, slower)
        fprintf(stderr, "[b200] the hybrid kernels' own fp32 records could not be built (%s); using the shared records\n", g_err);
        cudaGetLastError();
        cudaFree(a->d_nodesH); cudaFree(a->d_trisH);
        a->d_nodesH = nullptr; a->d_trisH = nullptr;
    }
    a->upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // flat host copies are no longer needed (keep the small header fields)
    std::vector<Node32>().swap(a->flat.nodes32); std::vector<Tri32>().swap(a->flat.tris32);
    std::vector<Node64>().swap(a->flat.nodes64); std::vector<Tri64>().swap(a->flat.tris64);
    std::vector<Tri32>().swap(a->flat.tris32t); std::vector<Tri64>().swap(a->flat.tris64t);
    std::vector<uint32_t>().swap(a->flat.slot_of_prim);
    return a;
}

extern "C" void ri_b200_free(ri_b200_accel_t *a)
{
    if (!a) return;
    if (a->device < 0) { delete a; return; }
    cudaSetDevice(a->device);
    if (a->stream) cudaStreamSynchronize(a->stream);
    cudaFree(a->pixc_pix);
    cudaFree(a->d_nodesH); cudaFree(a->d_trisH); cudaFree(a->d_k6_rays); cudaFree(a->d_k6_perm); cudaFree(a->d_k6_ctr);
    cudaFree(a->d_nodes32); cudaFree(a->d_tris32); cudaFree(a->d_nodes64); cudaFree(a->d_tris64); cudaFree(a->d_tris32t); cudaFree(a->d_tris64t); cudaFree(a->d_slot_of_prim); cudaFree(a->d_nrm64); cudaFree(a->d_nrm32); cudaFree(a->d_col); cudaFree(a->d_st); cudaFree(a->d_attr_flags); cudaFree(a->d_tex);
    for (int i = 0; i < 2; ++i) { cudaFree(a->d_in[i]); cudaFree(a->d_out[i]); }
    cudaFree(a->d_whole_in); cudaFree(a->d_whole_out);
    for (auto p : a->d_frame) cudaFree(p);
    cudaFree(a->d_counters); cudaFree(a->d_one); cudaFree(a->d_work); cudaFree(a->d_mt_polys); cudaFree(a->d_mt_states);
    if (a->h_pin) cudaFreeHost(a->h_pin);
    for (auto &e : a->ev) if (e) cudaEventDestroy(e);
    if (a->stream) cudaStreamDestroy(a->stream);
    for (auto &s : a->copy_stream) if (s) cudaStreamDestroy(s);
    cudaGetLastError();
    delete a;
}

extern "C" int ri_b200_info(const ri_b200_accel_t *a, ri_b200_info_t *out)
{
    if (!a || !out) return fail("null argument");
    std::memset(out, 0, sizeof(*out));
    out->ntris = a->tree.ntris; out->ninner = a->tree.ninner; out->nleaf = a->tree.nleaf;
    out->max_depth = a->tree.max_depth; out->empty = a->tree.empty ? 1 : 0;
    out->precisions = a->precisions; out->device = a->device;
    for (int k = 0; k < 3; ++k) { out->bmin[k] = a->tree.bmin[k]; out->bmax[k] = a->tree.bmax[k]; }
    out->build_seconds = a->tree.build_seconds; out->upload_seconds = a->upload_seconds;
    out->device_bytes = a->device_bytes;
    return 0;
}

extern "C" int64_t ri_b200_export_nodes(const ri_b200_accel_t *a, ri_b200_node_t *out, int64_t capacity)
{
    static_assert(sizeof(ri_b200_node_t) == sizeof(CanonNode), "layout");
    if (!a) return fail("null argument");
    const int64_t n = (int64_t)a->tree.nodes.size();
    if (!out) return n;
    if (capacity < n) return fail("capacity %lld < %lld nodes", (long long)capacity, (long long)n);
    if (n) std::memcpy(out, a->tree.nodes.data(), (size_t)n * sizeof(CanonNode));
    return n;
}

extern "C" int ri_b200_triorder(const ri_b200_accel_t *a, uint32_t *orig_out)
{
    if (!a || !orig_out) return fail("null argument");
    if (!a->tree.orig.empty()) std::memcpy(orig_out, a->tree.orig.data(), a->tree.orig.size() * sizeof(uint32_t));
    return 0;
}

extern "C" int64_t ri_b200_export_flat(const ri_b200_accel_t *a, void *nodes32, void *nodes64, void *tris32, void *tris64,
                                       uint32_t *header_out)
{
    if (!a) return fail("null argument");
    if (a->device >= 0) return fail("flat records are only kept on RI_B200_HOST_ONLY accelerators");
    const FlatTree &f = a->flat;
    if (nodes32 && !f.nodes32.empty()) std::memcpy(nodes32, f.nodes32.data(), f.nodes32.size() * sizeof(Node32));
    if (nodes64 && !f.nodes64.empty()) std::memcpy(nodes64, f.nodes64.data(), f.nodes64.size() * sizeof(Node64));
    if (tris32 && !f.tris32.empty()) std::memcpy(tris32, f.tris32.data(), f.tris32.size() * sizeof(Tri32));
    if (tris64 && !f.tris64.empty()) std::memcpy(tris64, f.tris64.data(), f.tris64.size() * sizeof(Tri64));
    if (header_out) { header_out[0] = f.root_word; header_out[1] = f.ninner; header_out[2] = f.top_count; header_out[3] = (uint32_t)f.nslots; }
    return (int64_t)f.ninner;
}

extern "C" int ri_b200_export_flat_transposed(const ri_b200_accel_t *a, void *tris32t, void *tris64t)
{
    if (!a) return fail("null argument");
    if (a->device >= 0) return fail("flat records are only kept on RI_B200_HOST_ONLY accelerators");
    const FlatTree &f = a->flat;
    if (tris32t && !f.tris32t.empty()) std::memcpy(tris32t, f.tris32t.data(), f.tris32t.size() * sizeof(Tri32));
    if (tris64t && !f.tris64t.empty()) std::memcpy(tris64t, f.tris64t.data(), f.tris64t.size() * sizeof(Tri64));
    return 0;
}

extern "C" int ri_b200_set_normals(ri_b200_accel_t *a, const double *tri_normals)
{
    if (!a) return fail("null argument");
    if (a->device < 0) return fail("host-only accelerator: no device records, no CPU fallback");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaFree(a->d_nrm64); cudaFree(a->d_nrm32); a->d_nrm64 = nullptr; a->d_nrm32 = nullptr;
    if (!tri_normals || a->tree.empty) return 0;
    const size_t n = (size_t)a->tree.ntris;
    std::vector<double> h64(9 * n);
    std::vector<float> h32(9 * n);
    for (size_t p = 0; p < n; ++p)
        for (int k = 0; k < 9; ++k) {
            const double v = tri_normals[9 * (size_t)a->tree.orig[p] + k];
            h64[9 * p + k] = v; h32[9 * p + k] = (float)v;
        }
    if (a->precisions & RI_B200_PREC_F64) {
        CUDA_OK(cudaMalloc((void **)&a->d_nrm64, h64.size() * sizeof(double)));
        CUDA_OK(cudaMemcpy(a->d_nrm64, h64.data(), h64.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (a->precisions & RI_B200_PREC_F32) {
        CUDA_OK(cudaMalloc((void **)&a->d_nrm32, h32.size() * sizeof(float)));
        CUDA_OK(cudaMemcpy(a->d_nrm32, h32.data(), h32.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" void *ri_b200_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); fail("cudaMallocHost(%llu) failed", (unsigned long long)bytes); return nullptr; }
    return p;
}
extern "C" void ri_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------------
// device-buffer entry points
// ------------------------------------------------------------------------------------------------
static int need(const ri_b200_accel *a, uint32_t prec)
{
    if (!a) return fail("null accelerator");
    if (a->device < 0) return fail("host-only accelerator: no device records, no CPU fallback");
    if (!(a->precisions & prec)) return fail("accelerator was built without %s records", prec == RI_B200_PREC_F32 ? "fp32" : "fp64");
    return 0;
}

extern "C" int ri_b200_intersect_dev_f32(ri_b200_accel_t *a, const float *d_rays, uint64_t n, ri_b200_hit_f32 *d_out, void *stream)
{
    if (need(a, RI_B200_PREC_F32)) return -1;
    CUDA_OK(cudaSetDevice(a->device));
    return launch_trace<float, false, false>(a, d_rays, n, d_out, nullptr, nullptr, stream ? (cudaStream_t)stream : a->stream);
}
extern "C" int ri_b200_occluded_dev_f32(ri_b200_accel_t *a, const float *d_rays, uint64_t n, uint8_t *d_out, void *stream)
{
    if (need(a, RI_B200_PREC_F32)) return -1;
    CUDA_OK(cudaSetDevice(a->device));
    return launch_trace<float, true, false>(a, d_rays, n, nullptr, d_out, nullptr, stream ? (cudaStream_t)stream : a->stream);
}
extern "C" int ri_b200_intersect_dev_f64(ri_b200_accel_t *a, const double *d_rays, uint64_t n, ri_b200_hit_f64 *d_out, void *stream)
{
    if (need(a, RI_B200_PREC_F64)) return -1;
    CUDA_OK(cudaSetDevice(a->device));
    return launch_trace<double, false, false>(a, d_rays, n, d_out, nullptr, nullptr, stream ? (cudaStream_t)stream : a->stream);
}
extern "C" int ri_b200_occluded_dev_f64(ri_b200_accel_t *a, const double *d_rays, uint64_t n, uint8_t *d_out, void *stream)
{
    if (need(a, RI_B200_PREC_F64)) return -1;
    CUDA_OK(cudaSetDevice(a->device));
    return launch_trace<double, true, false>(a, d_rays, n, nullptr, d_out, nullptr, stream ? (cudaStream_t)stream : a->stream);
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points: chunked, double-buffered H2D -> kernel -> D2H
// ------------------------------------------------------------------------------------------------
static uint64_t chunk_rays()
{
    static const uint64_t v = getenv("B200_CHUNK") ? (uint64_t)atoll(getenv("B200_CHUNK")) : (1ull << 21);   // measured: 0.5M 450, 1M 541, 2M 587, 4M 592, 8M 555 Mrays/s e2e on C3
    return v < 1024 ? 1024 : v;
}

static uint64_t chunk_ramp()
{
    static const uint64_t v = getenv("B200_CHUNK0") ? (uint64_t)atoll(getenv("B200_CHUNK0")) : (1ull << 18);
    return v < 1024 ? 1024 : v;
}

static int ensure_stage(ri_b200_accel *a, uint64_t in_bytes, uint64_t out_bytes)
{
    if (a->stage_in_bytes < in_bytes) {
        for (int i = 0; i < 2; ++i) { cudaFree(a->d_in[i]); a->d_in[i] = nullptr; CUDA_OK(cudaMalloc(&a->d_in[i], in_bytes)); }
        a->stage_in_bytes = in_bytes;
    }
    if (a->stage_out_bytes < out_bytes) {
        for (int i = 0; i < 2; ++i) { cudaFree(a->d_out[i]); a->d_out[i] = nullptr; CUDA_OK(cudaMalloc(&a->d_out[i], out_bytes)); }
        a->stage_out_bytes = out_bytes;
    }
    return 0;
}

// Occlusion batches from HOST buffers, streamed: ONE persistent launch consumes the batch while the copy engine is still
// uploading it.  The upload goes piece by piece on a copy stream, each piece followed (in stream order) by a 4-byte write of
// the number of rays that have landed; a warp that claims rays beyond that cursor waits for it (pool.cuh).  Against one launch
// per piece this removes the ramp and the tail of every launch but one (2 Mi-ray pieces: 2.35 ms each instead of 2.14).
template <typename Real>
static int host_occluded_streamed(ri_b200_accel *a, const Real *rays, uint64_t n, uint8_t *out)
{
    const uint64_t ray_bytes = RayIO<Real>::kRayStride * sizeof(Real);
    // whole-batch staging; a batch that does not fit (or more than 8 GiB of rays) takes the launch-per-piece path instead
    if (n * ray_bytes > (8ull << 30)) return 1;
    if (a->whole_in_bytes < n * ray_bytes) {
        cudaFree(a->d_whole_in); a->d_whole_in = nullptr; a->whole_in_bytes = 0;
        if (cudaMalloc(&a->d_whole_in, n * ray_bytes) != cudaSuccess) { cudaGetLastError(); a->d_whole_in = nullptr; return 1; }
        a->whole_in_bytes = n * ray_bytes;
    }
    if (a->whole_out_bytes < n) {
        cudaFree(a->d_whole_out); a->d_whole_out = nullptr; a->whole_out_bytes = 0;
        if (cudaMalloc(&a->d_whole_out, n) != cudaSuccess) { cudaGetLastError(); a->d_whole_out = nullptr; return 1; }
        a->whole_out_bytes = n;
    }
    unsigned int *d_ready = a->d_work + kWorkRing, *d_fault = a->d_work + kWorkRing + 1;
    uint32_t *cursor = (uint32_t *)a->h_pin;                       // pinned: one value per piece, <= 1024 pieces
    cudaStream_t ks = a->stream, cs = a->copy_stream[0];
    // piece size: measured on C3 (16 Mi rays, kernel alone 17.1 ms): 2 Mi 18.6 ms, 1 Mi 18.2, 512 Ki 18.0, 256 Ki 17.9, 128 Ki 17.95
    static const uint64_t piece0 = getenv("B200_PIECE") ? (uint64_t)atoll(getenv("B200_PIECE")) : (1ull << 18);
    uint64_t piece = piece0 < 1024 ? 1024 : piece0;
    while ((n + piece - 1) / piece > 1000) piece *= 2;

    CUDA_OK(cudaMemsetAsync(d_ready, 0, 2 * sizeof(unsigned int), ks));
    CUDA_OK(cudaEventRecord(a->ev[6], ks));
    CUDA_OK(cudaStreamWaitEvent(cs, a->ev[6], 0));
    if (launch_trace<Real, true, false>(a, (const Real *)a->d_whole_in, n, nullptr, (uint8_t *)a->d_whole_out, nullptr, ks, nullptr, 1,
                                        d_ready, d_fault)) return -1;
    uint64_t done = 0;
    int k = 0;
    while (done < n) {
        const uint64_t m = (n - done) < piece ? (n - done) : piece;
        CUDA_OK(cudaMemcpyAsync((char *)a->d_whole_in + done * ray_bytes, (const char *)rays + done * ray_bytes, m * ray_bytes,
                                cudaMemcpyHostToDevice, cs));
        done += m;
        cursor[k] = (uint32_t)done;
        CUDA_OK(cudaMemcpyAsync(d_ready, &cursor[k], sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
        ++k;
    }
    CUDA_OK(cudaMemcpyAsync(out, a->d_whole_out, n, cudaMemcpyDeviceToHost, ks));
    unsigned int *h_fault = (unsigned int *)((char *)a->h_pin + 4092);
    CUDA_OK(cudaMemcpyAsync(h_fault, d_fault, sizeof(unsigned int), cudaMemcpyDeviceToHost, ks));
    CUDA_OK(cudaStreamSynchronize(cs));
    CUDA_OK(cudaStreamSynchronize(ks));
    return *h_fault ? 1 : 0;          // 1: the upload could not run beside the kernel -- the caller takes the launch-per-piece path
}

template <typename Real, bool ANYHIT>
static int host_batch(ri_b200_accel *a, const Real *rays, uint64_t n, void *out)
{
    using Hit = typename RayIO<Real>::Hit;
    if (need(a, sizeof(Real) == 4 ? RI_B200_PREC_F32 : RI_B200_PREC_F64)) return -1;
    if (n == 0) return 0;
    if (!rays || !out) return fail("null buffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    {
        static const bool streamed = !(getenv("B200_STREAMED") && atoi(getenv("B200_STREAMED")) == 0) &&
                                     !(getenv("B200_POOL") && atoi(getenv("B200_POOL")) == 0);
        if (ANYHIT && streamed && !a->streamed_faulted && n < (1ull << 31)) {
            const int rc = host_occluded_streamed<Real>(a, rays, n, (uint8_t *)out);
            if (rc <= 0) return rc;
            // the upload could not run beside the kernel (a profiler or CUDA_LAUNCH_BLOCKING serialises the two streams): LATCH the
            // launch-per-piece path for this accelerator, so that only this one call pays the spin cap (~0.1 s), and say so once
            a->streamed_faulted = true;
            fprintf(stderr, "[b200] streamed host-buffer upload cannot overlap the kernel here; using one launch per piece from now on\n");
        }
    }
    const uint64_t ray_bytes = RayIO<Real>::kRayStride * sizeof(Real);
    const uint64_t out_bytes = ANYHIT ? 1 : sizeof(Hit);
    const uint64_t chunk = n < chunk_rays() ? n : chunk_rays();
    if (ensure_stage(a, chunk * ray_bytes, chunk * out_bytes)) return -1;

    // chunk sizes ramp up (256 Ki, 512 Ki, ... up to `chunk`): the first upload, which nothing can hide, stays short
    uint64_t done = 0, m_next = chunk_ramp() < chunk ? chunk_ramp() : chunk;
    int slot = 0;
    const bool trace = getenv("B200_E2E_TRACE") != nullptr;     // diagnostics: per-chunk device timeline on stderr
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); tev.push_back(e); } };
    while (done < n) {
        const uint64_t m = (n - done) < m_next ? (n - done) : m_next;
        m_next = (m_next * 2 < chunk) ? m_next * 2 : chunk;
        cudaStream_t st = a->copy_stream[slot];
        const int dbg_skip = getenv("B200_E2E_SKIP") ? atoi(getenv("B200_E2E_SKIP")) : 0;    // diagnostics only: 1 = no upload, 2 = no kernel
        mark(st);
        if (dbg_skip != 1) CUDA_OK(cudaMemcpyAsync(a->d_in[slot], (const char *)rays + done * ray_bytes, m * ray_bytes, cudaMemcpyHostToDevice, st));
        mark(st);
        if (dbg_skip != 2 && launch_trace<Real, ANYHIT, false>(a, (const Real *)a->d_in[slot], m, ANYHIT ? nullptr : (Hit *)a->d_out[slot],
                                              ANYHIT ? (uint8_t *)a->d_out[slot] : nullptr, nullptr, st)) return -1;
        mark(st);
        CUDA_OK(cudaMemcpyAsync((char *)out + done * out_bytes, a->d_out[slot], m * out_bytes, cudaMemcpyDeviceToHost, st));
        mark(st);
        done += m;
        slot ^= 1;
        // the slot we are about to reuse must have drained (its stream runs copy->kernel->copy in order)
        if (done < n) CUDA_OK(cudaStreamSynchronize(a->copy_stream[slot]));
    }
    CUDA_OK(cudaStreamSynchronize(a->copy_stream[0]));
    CUDA_OK(cudaStreamSynchronize(a->copy_stream[1]));
    for (size_t i = 0; i + 3 < tev.size(); i += 4) {
        float t[4];
        for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], tev[0], tev[i + k]);
        fprintf(stderr, "[e2e] chunk %zu: upload %.3f-%.3f kernel -%.3f download -%.3f ms\n", i / 4, t[0], t[1], t[2], t[3]);
    }
    for (auto e : tev) cudaEventDestroy(e);
    return 0;
}

extern "C" int ri_b200_intersect_batch_f32(ri_b200_accel_t *a, const float *rays, uint64_t n, ri_b200_hit_f32 *out)
{ return host_batch<float, false>(a, rays, n, out); }
extern "C" int ri_b200_occluded_batch_f32(ri_b200_accel_t *a, const float *rays, uint64_t n, uint8_t *out)
{ return host_batch<float, true>(a, rays, n, out); }
extern "C" int ri_b200_intersect_batch_f64(ri_b200_accel_t *a, const double *rays, uint64_t n, ri_b200_hit_f64 *out)
{ return host_batch<double, false>(a, rays, n, out); }
extern "C" int ri_b200_occluded_batch_f64(ri_b200_accel_t *a, const double *rays, uint64_t n, uint8_t *out)
{ return host_batch<double, true>(a, rays, n, out); }

extern "C" int ri_b200_count_batch(ri_b200_accel_t *a, const void *rays, uint64_t n, uint32_t precision, int anyhit,
                                   ri_b200_counters_t *out)
{
    if (need(a, precision)) return -1;
    if (!out) return fail("null argument");
    std::memset(out, 0, sizeof(*out));
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const bool f32 = precision == RI_B200_PREC_F32;
    const uint64_t ray_bytes = f32 ? 32 : 48;
    void *d_rays = nullptr, *d_res = nullptr;
    CUDA_OK(cudaMalloc(&d_rays, n * ray_bytes));
    CUDA_OK(cudaMalloc(&d_res, n * (anyhit ? 1 : (f32 ? sizeof(ri_b200_hit_f32) : sizeof(ri_b200_hit_f64)))));
    int rc = 0;
    auto body = [&]() -> int {
        CUDA_OK(cudaMemcpyAsync(d_rays, rays, n * ray_bytes, cudaMemcpyHostToDevice, a->stream));
        CUDA_OK(cudaMemsetAsync(a->d_counters, 0, 8 * sizeof(unsigned long long), a->stream));
        int r;
        if (f32) r = anyhit ? launch_trace<float, true, true>(a, (const float *)d_rays, n, nullptr, (uint8_t *)d_res, a->d_counters, a->stream)
                            : launch_trace<float, false, true>(a, (const float *)d_rays, n, (ri_b200_hit_f32 *)d_res, nullptr, a->d_counters, a->stream);
        else     r = anyhit ? launch_trace<double, true, true>(a, (const double *)d_rays, n, nullptr, (uint8_t *)d_res, a->d_counters, a->stream)
                            : launch_trace<double, false, true>(a, (const double *)d_rays, n, (ri_b200_hit_f64 *)d_res, nullptr, a->d_counters, a->stream);
        if (r) return r;
        unsigned long long h[5];
        CUDA_OK(cudaMemcpyAsync(h, a->d_counters, sizeof(h), cudaMemcpyDeviceToHost, a->stream));
        CUDA_OK(cudaStreamSynchronize(a->stream));
        out->nrays = h[0]; out->ninner = h[1]; out->nleaf = h[2]; out->ntris = h[3]; out->nhit_tris = h[4];
        return 0;
    };
    rc = body();
    cudaFree(d_rays); cudaFree(d_res);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// post-hit state (intersection_state.c:99-248 for geometry carrying only "P")
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void normalize3(double d[3])          // vector.h:75-87
{
    const double n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    if (n2 > (double)1.0e-17f) {
        const double r = 1.0 / sqrt(n2);
        d[0] *= r; d[1] *= r; d[2] *= r;
    }
}
__device__ __forceinline__ void cross3(double d[3], const double a[3], const double b[3])
{
    d[0] = a[1] * b[2] - a[2] * b[1];
    d[1] = a[2] * b[0] - a[0] * b[2];
    d[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void ortho_basis(double b0[3], double b1[3], const double n[3])   // reflection.c:312-333
{
    int i;
    for (i = 0; i < 3; ++i) if (n[i] < 0.6 && n[i] > -0.6) break;
    if (i >= 3) i = 0;
    double e[3] = {0.0, 0.0, 0.0};
    e[i] = 1.0;
    cross3(b0, e, n);
    normalize3(b0);
    cross3(b1, n, b0);
    normalize3(b1);
}

// normals: optional [prim][9]; a triangle with nine zero components has none (its geom's `normals` is NULL) -> Ns = Ng
__device__ __forceinline__ void state_from_hit(const Tri64 *tris, const uint32_t *slot_of_prim, const double org[3], const double dir[3],
                                               double t, uint32_t prim, ri_b200_state_f64 &s,
                                               const double *normals = nullptr, double bu = 0.0, double bv = 0.0)
{
    TriRegs<double> tr;
    load_tri(tris + slot_of_prim[prim], tr);
    for (int k = 0; k < 3; ++k) s.P[k] = org[k] + dir[k] * t;
    cross3(s.Ng, tr.e1, tr.e2);                                   // (v1-v0) x (v2-v0), geometric.c:20-33
    normalize3(s.Ng);
    bool has_n = false;
    double n[9];
    if (normals) {
        for (int k = 0; k < 9; ++k) { n[k] = normals[9 * (size_t)prim + k]; has_n = has_n || (n[k] != 0.0); }
    }
    if (has_n) {                                                  // ri_lerp_vector, geometric.c:40-62 (not normalised)
        const double w0 = 1.0 - bu - bv;
        for (int k = 0; k < 3; ++k) {
            const double a = n[k] * w0, b = n[3 + k] * bu, c = n[6 + k] * bv;
            s.Ns[k] = (a + b) + c;
        }
    } else {
        for (int k = 0; k < 3; ++k) s.Ns[k] = s.Ng[k];
    }
    ortho_basis(s.tangent, s.binormal, s.Ng);
}

__global__ void state_kernel(const Tri64 *__restrict__ tris, const uint32_t *__restrict__ slot_of_prim, const double *__restrict__ normals,
                             const double *__restrict__ rays,
                             const ri_b200_hit_f64 *__restrict__ hits, uint64_t n, ri_b200_state_f64 *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ri_b200_state_f64 s;
    const ri_b200_hit_f64 h = hits[i];
    if (h.hit) {
        double org[3], dir[3];
        RayIO<double>::load(rays, i, org, dir);
        state_from_hit(tris, slot_of_prim, org, dir, h.t, h.prim, s, normals, h.u, h.v);
    } else {
        memset(&s, 0, sizeof(s));
    }
    out[i] = s;
}

extern "C" int ri_b200_state_batch_f64(ri_b200_accel_t *a, const double *rays, const ri_b200_hit_f64 *hits, uint64_t n,
                                       ri_b200_state_f64 *out)
{
    if (need(a, RI_B200_PREC_F64)) return -1;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    double *d_rays = nullptr; ri_b200_hit_f64 *d_hits = nullptr; ri_b200_state_f64 *d_out = nullptr;
    auto body = [&]() -> int {
        CUDA_OK(cudaMalloc((void **)&d_rays, n * 48));
        CUDA_OK(cudaMalloc((void **)&d_hits, n * sizeof(ri_b200_hit_f64)));
        CUDA_OK(cudaMalloc((void **)&d_out, n * sizeof(ri_b200_state_f64)));
        CUDA_OK(cudaMemcpyAsync(d_rays, rays, n * 48, cudaMemcpyHostToDevice, a->stream));
        CUDA_OK(cudaMemcpyAsync(d_hits, hits, n * sizeof(ri_b200_hit_f64), cudaMemcpyHostToDevice, a->stream));
        state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, a->stream>>>(a->d_tris64, a->d_slot_of_prim, a->d_nrm64, d_rays, d_hits, n, d_out);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(out, d_out, n * sizeof(ri_b200_state_f64), cudaMemcpyDeviceToHost, a->stream));
        CUDA_OK(cudaStreamSynchronize(a->stream));
        return 0;
    };
    const int rc = body();
    cudaFree(d_rays); cudaFree(d_hits); cudaFree(d_out);
    return rc;
}

// ---- the rest of ri_intersection_state_build: E, I, colour, st, inside (intersection_state.c:123-133, 192-246) ----------------
extern "C" int ri_b200_set_attributes(ri_b200_accel_t *a, const double *tri_colors, const uint8_t *has_color, const double *tri_st,
                                      const uint8_t *has_st, const uint8_t *tri_inside)
{
    if (!a) return fail("null argument");
    if (a->device < 0) return fail("host-only accelerator: no device records, no CPU fallback");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const size_t n = (size_t)a->tree.ntris;
    std::vector<uint8_t> fl(n, 0);
    if (a->d_attr_flags && n) {                              // keep the "carries the material texture" marks of ri_b200_set_texture
        CUDA_OK(cudaMemcpy(fl.data(), a->d_attr_flags, n, cudaMemcpyDeviceToHost));
        for (auto &f : fl) f &= 8u;
    }
    cudaFree(a->d_col); cudaFree(a->d_st); cudaFree(a->d_attr_flags);
    a->d_col = a->d_st = nullptr; a->d_attr_flags = nullptr;
    if (a->tree.empty) return 0;
    std::vector<double> col(9 * n, 0.0), st(6 * n, 0.0);
    for (size_t p = 0; p < n; ++p) {                         // post-build order, like the normals
        const size_t o = (size_t)a->tree.orig[p];
        if (tri_colors && has_color && has_color[o]) { std::memcpy(&col[9 * p], tri_colors + 9 * o, 9 * sizeof(double)); fl[p] |= 1; }
        if (tri_st && has_st && has_st[o]) { std::memcpy(&st[6 * p], tri_st + 6 * o, 6 * sizeof(double)); fl[p] |= 2; }
        if (tri_inside && tri_inside[o]) fl[p] |= 4;
    }
    CUDA_OK(cudaMalloc((void **)&a->d_col, col.size() * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&a->d_st, st.size() * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&a->d_attr_flags, n));
    CUDA_OK(cudaMemcpy(a->d_col, col.data(), col.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(a->d_st, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(a->d_attr_flags, fl.data(), n, cudaMemcpyHostToDevice));
    return 0;
}

// Material texture of the AO transport (ambientocclusion.c:393-401; what Surface "..." "texture" [...] loads, ri/attribute.c:309-327):
// rgba [h][w][4] floats = ri_texture_t.data after ri_texture_scale; tri_textured[ntris] marks the triangles whose geom carries it
// (NULL: all).  The st come from ri_b200_set_attributes (call it first); a textured triangle without st fetches at (0,0).
__global__ void mark_textured_kernel(uint8_t *flags, const uint8_t *textured, uint32_t n)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) flags[p] = (uint8_t)((flags[p] & ~8u) | (textured[p] ? 8u : 0u));
}

extern "C" int ri_b200_set_texture(ri_b200_accel_t *a, const float *rgba, int width, int height, const uint8_t *tri_textured)
{
    if (!a) return fail("null argument");
    if (a->device < 0) return fail("host-only accelerator: no device records, no CPU fallback");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaFree(a->d_tex); a->d_tex = nullptr; a->tex_w = a->tex_h = 0;
    if (!rgba || a->tree.empty) return 0;
    if (width < 1 || height < 1) return fail("bad texture size");
    const size_t n = (size_t)a->tree.ntris;
    std::vector<uint8_t> mark(n, 1);
    if (tri_textured) for (size_t p = 0; p < n; ++p) mark[p] = tri_textured[(size_t)a->tree.orig[p]] ? 1 : 0;
    if (!a->d_attr_flags) {
        CUDA_OK(cudaMalloc((void **)&a->d_attr_flags, n));
        CUDA_OK(cudaMemset(a->d_attr_flags, 0, n));
    }
    uint8_t *d_mark = nullptr;
    CUDA_OK(cudaMalloc((void **)&d_mark, n));
    cudaError_t e = cudaMemcpy(d_mark, mark.data(), n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { mark_textured_kernel<<<(unsigned)((n + 255) / 256), 256>>>(a->d_attr_flags, d_mark, (uint32_t)n); LAUNCHED(); e = cudaDeviceSynchronize(); }
    cudaFree(d_mark);
    if (e != cudaSuccess) return fail("texture mask upload failed: %s", cudaGetErrorString(e));
    CUDA_OK(cudaMalloc((void **)&a->d_tex, sizeof(float) * 4 * (size_t)width * height));
    CUDA_OK(cudaMemcpy(a->d_tex, rgba, sizeof(float) * 4 * (size_t)width * height, cudaMemcpyHostToDevice));
    a->tex_w = width; a->tex_h = height;
    return 0;
}

__global__ void state_ext_kernel(const double *__restrict__ col, const double *__restrict__ st, const uint8_t *__restrict__ flags,
                                 const double *__restrict__ rays, const ri_b200_hit_f64 *__restrict__ hits, uint64_t n,
                                 ri_b200_state_ext_f64 *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ri_b200_state_ext_f64 s;
    memset(&s, 0, sizeof(s));
    const ri_b200_hit_f64 h = hits[i];
    s.hit = (int32_t)h.hit;
    if (h.hit) {
        double org[3], dir[3];
        RayIO<double>::load(rays, i, org, dir);
        const uint8_t fl = flags ? flags[h.prim] : 0;
        for (int k = 0; k < 3; ++k) s.E[k] = org[k];                       // intersection_state.c:133
        normalize3(dir);                                                   // :130-131
        for (int k = 0; k < 3; ++k) s.I[k] = dir[k];
        if (fl & 1) {                                                      // ri_lerp_vector, geometric.c:40-62
            const double *c = col + 9 * (size_t)h.prim;
            const double w0 = 1.0 - h.u - h.v;
            for (int k = 0; k < 3; ++k) { const double x = c[k] * w0, y = c[3 + k] * h.u, z = c[6 + k] * h.v; s.color[k] = (x + y) + z; }
        } else {
            for (int k = 0; k < 3; ++k) s.color[k] = 1.0;                  // :204-207
        }
        if (fl & 2) {                                                      // lerp_uv, :266-280
            const double *t = st + 6 * (size_t)h.prim;
            s.st[0] = (1 - h.u - h.v) * t[0] + h.u * t[2] + h.v * t[4];
            s.st[1] = (1 - h.u - h.v) * t[1] + h.u * t[3] + h.v * t[5];
        }
        s.t = h.t;
        s.inside = (fl & 4) ? 1 : 0;                                       // :233-246
    }
    out[i] = s;
}

extern "C" int ri_b200_state_ext_batch_f64(ri_b200_accel_t *a, const double *rays, const ri_b200_hit_f64 *hits, uint64_t n,
                                           ri_b200_state_ext_f64 *out)
{
    if (need(a, RI_B200_PREC_F64)) return -1;
    if (n == 0) return 0;
    if (!rays || !hits || !out) return fail("null buffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    double *d_rays = nullptr; ri_b200_hit_f64 *d_hits = nullptr; ri_b200_state_ext_f64 *d_out = nullptr;
    auto body = [&]() -> int {
        CUDA_OK(cudaMalloc((void **)&d_rays, n * 48));
        CUDA_OK(cudaMalloc((void **)&d_hits, n * sizeof(ri_b200_hit_f64)));
        CUDA_OK(cudaMalloc((void **)&d_out, n * sizeof(ri_b200_state_ext_f64)));
        CUDA_OK(cudaMemcpyAsync(d_rays, rays, n * 48, cudaMemcpyHostToDevice, a->stream));
        CUDA_OK(cudaMemcpyAsync(d_hits, hits, n * sizeof(ri_b200_hit_f64), cudaMemcpyHostToDevice, a->stream));
        state_ext_kernel<<<(unsigned)((n + 255) / 256), 256, 0, a->stream>>>(a->d_col, a->d_st, a->d_attr_flags, d_rays, d_hits, n, d_out);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(out, d_out, n * sizeof(ri_b200_state_ext_f64), cudaMemcpyDeviceToHost, a->stream));
        CUDA_OK(cudaStreamSynchronize(a->stream));
        return 0;
    };
    const int rc = body();
    cudaFree(d_rays); cudaFree(d_hits); cudaFree(d_out);
    return rc;
}

// single ray: the vtable-compatible path (accel.h:30-34).  One launch of the f64 kernel + state.
extern "C" int ri_b200_intersect1(ri_b200_accel_t *a, const double org[3], const double dir[3],
                                  ri_b200_hit_f64 *hit, ri_b200_state_f64 *state)
{
    if (need(a, RI_B200_PREC_F64)) return -1;
    if (!org || !dir || !hit) return fail("null argument");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    double *h = (double *)a->h_pin;
    char *d = (char *)a->d_one;
    for (int k = 0; k < 3; ++k) { h[k] = org[k]; h[3 + k] = dir[k]; }
    double *d_ray = (double *)d;
    ri_b200_hit_f64 *d_hit = (ri_b200_hit_f64 *)(d + 64);
    ri_b200_state_f64 *d_state = (ri_b200_state_f64 *)(d + 128);
    CUDA_OK(cudaMemcpyAsync(d_ray, h, 48, cudaMemcpyHostToDevice, a->stream));
    if (launch_trace<double, false, false>(a, d_ray, 1, d_hit, nullptr, nullptr, a->stream)) return -1;
    if (state) {
        state_kernel<<<1, 32, 0, a->stream>>>(a->d_tris64, a->d_slot_of_prim, a->d_nrm64, d_ray, d_hit, 1, d_state);
        LAUNCHED();
    }
    CUDA_OK(cudaMemcpyAsync((char *)a->h_pin + 64, d + 64, 64 + sizeof(ri_b200_state_f64), cudaMemcpyDeviceToHost, a->stream));
    CUDA_OK(cudaStreamSynchronize(a->stream));
    std::memcpy(hit, (char *)a->h_pin + 64, sizeof(*hit));
    if (state) std::memcpy(state, (char *)a->h_pin + 128, sizeof(*state));
    return hit->hit ? 1 : 0;
}

// rays per wave of the wavefront entry points (generator -> traverser -> accumulator): 2^24, or B200_WAVE_RAYS (tests: a small value
// drives the chunk loops of the point / gather / shade entry points with small inputs)
static uint64_t wave_rays()
{
    if (const char *e = getenv("B200_WAVE_RAYS")) { const long long v = atoll(e); if (v >= 64) return (uint64_t)v; }
    return 1ull << 24;
}

#include "frame.cuh"
#include "pathtrace.cuh"
#include "points.cuh"
#include "beam.cuh"
#include "whitted.cuh"
#include "gather.cuh"
#include "shade.cuh"
#include "bvh_build_gpu.cuh"
#include "hdr.cuh"

// Device build path of ri_b200_build: tree on the device, topology numbering + node records on the host (a few hundred
// thousand records), triangle slots on the device again.  The triangles cross PCIe once and are never gathered on the host.
static int build_on_device(ri_b200_accel *a, const double *tri_xyz, uint64_t ntris)
{
    const bool want32 = (a->precisions & RI_B200_PREC_F32) != 0, want64 = (a->precisions & RI_B200_PREC_F64) != 0;
    DeviceBuild db;
    if (build_tree_device(tri_xyz, ntris, a->tree, nullptr, db) != 0) return -1;
    flatten_tree(a->tree, 1024, want32, want64, a->flat, false);
    if (a->flat.overflow || a->tree.empty) return 0;
    const uint64_t nslots = a->flat.nslots;
    std::vector<uint32_t> leaf_slot(db.total, 0u);                       // breadth-first node -> first slot of its leaf
    for (uint32_t i = 0; i < db.total; ++i) leaf_slot[i] = a->flat.leaf_slot[db.canon_of[i]];
    uint32_t *d_leaf_slot = db.s.d_hist;                                 // scratch of the builder, free again (>= 2n/17 * 384 words)
    if ((uint64_t)db.total > ((uint64_t)2 * (db.n / 17u + 2u) + 8u) * 384u) return fail("device builder: scratch too small for the slot table");
    CUDA_OK(cudaMemcpy(d_leaf_slot, leaf_slot.data(), (size_t)db.total * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (want32) {
        CUDA_OK(cudaMalloc((void **)&a->d_tris32, nslots * sizeof(Tri32)));  CUDA_OK(cudaMemset(a->d_tris32, 0, nslots * sizeof(Tri32)));
        CUDA_OK(cudaMalloc((void **)&a->d_tris32t, nslots * sizeof(Tri32))); CUDA_OK(cudaMemset(a->d_tris32t, 0, nslots * sizeof(Tri32)));
        a->device_bytes += 2 * nslots * sizeof(Tri32);
    }
    if (want64) {
        CUDA_OK(cudaMalloc((void **)&a->d_tris64, nslots * sizeof(Tri64)));  CUDA_OK(cudaMemset(a->d_tris64, 0, nslots * sizeof(Tri64)));
        CUDA_OK(cudaMalloc((void **)&a->d_tris64t, nslots * sizeof(Tri64))); CUDA_OK(cudaMemset(a->d_tris64t, 0, nslots * sizeof(Tri64)));
        a->device_bytes += 2 * nslots * sizeof(Tri64);
    }
    CUDA_OK(cudaMalloc((void **)&a->d_slot_of_prim, (size_t)db.n * sizeof(uint32_t)));
    a->device_bytes += (uint64_t)db.n * sizeof(uint32_t);
    gb_fill_slots<<<(db.n + 255) / 256, 256>>>(db.s.d_tri, db.s.d_box[db.cur], db.n, db.s.d_nodes, d_leaf_slot, a->d_tris32, a->d_tris64,
                                              reinterpret_cast<char *>(a->d_tris32t), reinterpret_cast<char *>(a->d_tris64t), a->d_slot_of_prim);
    LAUNCHED();
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

// The fp32 occlusion traverser the AO rays run through: pool.cuh's pooled scheme (same answers, same argument for exactness),
// specialised and put on an instruction diet.
//
// Why: profiles/r02_* show the pooled kernel limited by warp-instruction ISSUE (an A/B with ~10 dead bookkeeping instructions per
// loop iteration added cost 3 % of the rays/s; ray order, L1 hints and the L1 wavefront count did not move it), so every instruction
// that is not arithmetic of the reference is overhead to remove:
//   * STATIC shared memory, templated on the stack capacity: every shared-memory address is lane offset + compile-time constant.
//     With the dynamic array of pool.cuh the compiler re-derived the shared window base (S2UR SR_CgaCtaId, UMOV, ULEA, LDCU) at
//     each use because it cannot spare registers for the pointers at 64 registers / 4 CTAs per SM.
//   * the per-lane stack pointer is kept as a BYTE OFFSET that already contains the lane's column (push / pop = one add);
//   * the ray's three sign bits live in one mask: near child = (mask >> axis) & 1 instead of two selects and a compare;
//   * leaf-round / refill thresholds and the result kind (occlusion bytes or per-point counters) are template constants;
//   * the acceptance window of triangle_isect is a chain of predicated SETPs (one instruction per comparison);
//   * when both children of a node are kept, the one more likely to end the query is visited first (kOrder, at the node step): an
//     occlusion query may visit leaves in any order, the near-child-first rule of the reference only matters for closest hits.
// Arithmetic: packed.cuh (FFMA2 on (A, B) triangle pairs and (lo, hi) box planes, each result rounded like the scalar operation).
// Trees deeper than the largest instantiated stack, fp64 records and the experiments stay with pool.cuh.
#pragma once

namespace b200 {

// acceptance of both triangles of an item in leaf order from t_leaf = 1e38 (bvh.c:833-848; NEGATED comparisons like tri_test_bf):
// returns tl < 1e38 after the pair.
__device__ __forceinline__ bool pair_accept_occluded(const PairMT &r, const bool valid_b)
{
    float aA, aB, uA, uB, vA, vB, tA, tB, wA, wB;
    upk2(r.a, aA, aB); upk2(r.u, uA, uB); upk2(r.v, vA, vB); upk2(r.t, tA, tB); upk2(r.uv, wA, wB);
    uint32_t res;
    asm("{\n\t.reg .pred p, q;\n\t.reg .f32 x, tl;\n\t"
        "abs.f32 x, %1;\n\t"
        "setp.gt.f32 p, x, 0f283424DC;\n\t"             // |a| > 1e-14f
        "setp.geu.and.f32 p, %2, 0f00000000, p;\n\t"    // !(u < 0)
        "setp.leu.and.f32 p, %2, 0f3F800000, p;\n\t"    // !(u > 1)
        "setp.geu.and.f32 p, %3, 0f00000000, p;\n\t"    // !(v < 0)
        "setp.leu.and.f32 p, %4, 0f3F800000, p;\n\t"    // !(u + v > 1)
        "setp.geu.and.f32 p, %5, 0f00000000, p;\n\t"    // !(t < 0)
        "setp.leu.and.f32 p, %5, 0f7E967699, p;\n\t"    // !(t > 1e38)
        "selp.f32 tl, %5, 0f7E967699, p;\n\t"           // t_leaf after A
        "abs.f32 x, %6;\n\t"
        "setp.ne.u32 q, %11, 0;\n\t"                    // B is a real triangle
        "setp.gt.and.f32 q, x, 0f283424DC, q;\n\t"
        "setp.geu.and.f32 q, %7, 0f00000000, q;\n\t"
        "setp.leu.and.f32 q, %7, 0f3F800000, q;\n\t"
        "setp.geu.and.f32 q, %8, 0f00000000, q;\n\t"
        "setp.leu.and.f32 q, %9, 0f3F800000, q;\n\t"
        "setp.geu.and.f32 q, %10, 0f00000000, q;\n\t"
        "setp.leu.and.f32 q, %10, tl, q;\n\t"           // !(t > t_leaf)
        "selp.f32 tl, %10, tl, q;\n\t"
        "setp.lt.f32 p, tl, 0f7E967699;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(res)
        : "f"(aA), "f"(uA), "f"(vA), "f"(wA), "f"(tA), "f"(aB), "f"(uB), "f"(vB), "f"(wB), "f"(tB), "r"((uint32_t)valid_b));
    return res != 0u;
}

// shared-memory accesses by 32-bit shared-window address: [lane offset + link-time constant] in SASS.  Through generic pointers the
// compiler re-derives the shared window base (S2UR SR_CgaCtaId, UMOV, ULEA) at every use when it has no register to keep it in.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{ asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

struct Pool32Warp {
    float4 rays[64];                     // a lane's (org, dir): 32 bytes, read by the lanes that test its leaf items
    uint2  desc[32];                     // leaf-round descriptors of this warp (one address register serves both areas)
};
#ifndef B200_OCC_THREADS
#define B200_OCC_THREADS 64       // C3 / C5 occlusion, Mrays/s (scripts/gpu_r3i.sh): 64 x 16: 1123 / 983; 256 x 4: 1113 / 968; 128 x 8: 1114 / 970;
#endif
#ifndef B200_OCC_CTAS
#define B200_OCC_CTAS 16          // 192 x 5: 1093 / 937; 128 x 7 and 128 x 6 at 71 registers: 1072 / 927, 1079 / 938
#endif
constexpr int kOccThreads = B200_OCC_THREADS;       // CTA size / CTAs per SM of the occlusion kernel (A/B knobs)
template <int kCap> struct Pool32Smem {
    uint32_t   stack[kCap * kOccThreads];     // [depth][thread]: conflict-free columns
    Pool32Warp warp[kOccThreads / 32];
};

template <int kCap, bool kCounts, int kOrder>
__global__ void __launch_bounds__(kOccThreads, B200_OCC_CTAS)
occluded_pool32_kernel(const SceneView<float> S, const char *__restrict__ trisT, const float *__restrict__ rays, const uint32_t n,
                       const uint32_t chunk, uint8_t *__restrict__ occ, uint32_t *__restrict__ counts, const uint32_t rays_per_count,
                       unsigned int *__restrict__ work_counter, const unsigned int *__restrict__ ready, unsigned int *__restrict__ fault,
                       const PackK K, const uint32_t *__restrict__ perm)
{
    constexpr unsigned FULL = 0xffffffffu;
#ifndef B200_OCC_REFILL
#define B200_OCC_REFILL 8        // 2 / 4 / 6 / 8 / 12 -> 1040 / 1079 / 1099 / 1105 / 1099 Mrays/s on C3 (scripts/gpu_r3e.sh)
#endif
#ifndef B200_OCC_LEAFAT
#define B200_OCC_LEAFAT 32
#endif
    constexpr uint32_t kRefillAt = B200_OCC_REFILL, kLeafAt = B200_OCC_LEAFAT;
    constexpr uint32_t kRow = kOccThreads * 4u;                        // bytes between two stack levels of a lane
    __shared__ __align__(16) Pool32Smem<kCap> sm;
    const unsigned lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = (2u << lane) - 1u;
    // shared-window address of the block's record, taken ONCE through an opaque asm so that it is kept (the compiler's own
    // conversion is rematerialised at every use); everything else is a compile-time offset from it
    uint32_t sm_a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sm_a) : "l"(&sm));
    const uint32_t stack_a = sm_a + (uint32_t)offsetof(Pool32Smem<kCap>, stack);
    const uint32_t rays_a = sm_a + (uint32_t)offsetof(Pool32Smem<kCap>, warp) + (wbase >> 5) * (uint32_t)sizeof(Pool32Warp);   // this warp's 32 ray slots
    const uint32_t desc_a = rays_a + (uint32_t)offsetof(Pool32Warp, desc);                                                        // and its descriptor run

    uint32_t chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool exhausted = false;                      // warp-uniform

    // per-lane ray state.  cur: kIdle | inner-node index | leaf word; prog: items of that leaf already tested;
    // spa: byte offset of the lane's next free stack entry (level * kRow + 4 * thread); sgn: bit k = dir[k] < 0
    uint32_t cur = kIdle, prog = 0, idx = 0, spa = threadIdx.x * 4u, sgn = 0;
    float org[3] = {0.0f, 0.0f, 0.0f}, inv[3] = {0.0f, 0.0f, 0.0f};

    auto retire = [&](const bool hit) {
        if (kCounts) { if (hit) atomicAdd(&counts[idx / rays_per_count], 1u); }
        else occ[idx] = hit ? 1 : 0;
    };

    for (;;) {
        // ------------------------------------------------------------------ fetch (pool.cuh)
        unsigned idle = __ballot_sync(FULL, cur == kIdle);
        while (idle && !exhausted) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = (n - base < chunk) ? n : base + chunk;
                if (ready) {                     // streamed upload: wait until the copy engine has delivered this chunk
                    if (lane == 0) {
                        unsigned spins = 0, have;
                        for (;;) {
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(ready) : "memory");
                            if (have >= chunk_end) break;
                            __nanosleep(200);
                            if (++spins > kPoolSpinCap) { atomicExch(fault, 1u); break; }
                        }
                    }
                    __syncwarp();
                    if (*(volatile unsigned int *)fault) { exhausted = true; break; }
                }
            }
            const unsigned avail = chunk_end - chunk_next;
            const unsigned n_idle = __popc(idle);
            const unsigned take = n_idle < avail ? n_idle : avail;
            const unsigned rank = __popc(idle & lt_mask);
            if (cur == kIdle && rank < take) {
                idx = chunk_next + rank;
                float dir[3];
                if (ready) RayIO<float>::load_coherent(rays, idx, org, dir);      // the copy engine is still writing this buffer: no ld.global.nc
                else RayIO<float>::load(rays, idx, org, dir);
                if (perm) idx = __ldg(perm + idx);          // a reordered batch (reorder.cuh): results go back to the ray's place in the input
                sts128(rays_a + lane * 32u, make_float4(org[0], org[1], org[2], 0.0f));
                sts128(rays_a + lane * 32u + 16u, make_float4(dir[0], dir[1], dir[2], 0.0f));
                const bool sx = dir[0] < 0.0f, sy = dir[1] < 0.0f, sz = dir[2] < 0.0f;
                sgn = (sx ? 1u : 0u) | (sy ? 2u : 0u) | (sz ? 4u : 0u);
#pragma unroll
                for (int k = 0; k < 3; ++k)      // bvh.c:473-497
                    inv[k] = (fabsf(dir[k]) > 1.0e-14f) ? 1.0f / dir[k] : ((dir[k] < 0.0f) ? -FLT_MAX : FLT_MAX);
                float tmin;
                const bool in_scene = (S.root_word != kDoneWord) &&
                    slab<float>(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], org, inv, sx, sy, sz, tmin);
                spa = threadIdx.x * 4u; prog = 0;
                if (in_scene) cur = S.root_word;
                else retire(false);              // bvh.c:446 / 522-526: miss without traversal
            }
            chunk_next += take;
            idle = __ballot_sync(FULL, cur == kIdle);
        }
        if (idle == FULL) break;                 // nothing in flight and nothing left to fetch

        // ------------------------------------------------------------------ traverse
        for (;;) {
            const bool in_leaf = (int32_t)cur < 0;
            const uint32_t nitems = (((cur >> kLeafShift) & 15u) + 2u) >> 1;      // (ntris + 1) / 2 pairs
            const uint32_t cnt = in_leaf ? nitems - prog : 0u;                    // >= 1 for a lane standing in a leaf
            const uint32_t total = __reduce_add_sync(FULL, cnt);
            const unsigned owners = __ballot_sync(FULL, in_leaf);
            const unsigned in_node = __ballot_sync(FULL, cur < kIdle);            // inner-node indices are < kIdle, leaf words above
            const unsigned n_node = __popc(in_node);
            if ((in_node | owners) == 0u) break;
            if (!exhausted && (uint32_t)__popc(~(in_node | owners)) >= kRefillAt) break;

            if (total >= kLeafAt || total > n_node) {
                // ---- leaf round: items 0..31 of the pool, one per lane (pool.cuh)
                uint32_t excl = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    excl += (uint32_t)__popc(ballot_bit(cnt, 1u << b) & lt_mask) << b;
                const bool owner = in_leaf && excl < 32u;
                const unsigned starts = __reduce_or_sync(FULL, owner ? (1u << excl) : 0u);
                if (owner) sts64(desc_a + (uint32_t)__popc(owners & lt_mask) * 8u, make_uint2(cur, lane | ((prog - excl + 64u) << 8)));
                __syncwarp();
                bool hit = false;
                if (lane < total) {
                    const uint2 d = lds64(desc_a + (uint32_t)__popc(starts & le_mask) * 8u - 8u);
                    const unsigned own = d.y & 31u;
                    const uint32_t item = lane + (d.y >> 8) - 64u;                // item number inside the owner's leaf
                    const uint32_t ntris = ((d.x >> kLeafShift) & 15u) + 1u, slot0 = d.x & kSlotMask;
                    const uint32_t m = ((ntris + 3u) >> 2) << 1;                  // row length in pairs
                    const uint32_t o0 = slot0 * 3u + item * 2u, o1 = o0 + 2u * m, o2 = slot0 * 3u + 4u * m + item;    // 16-byte units
                    const P4 c0 = ldg256p(trisT + (size_t)o0 * 16u), c1 = ldg256p(trisT + (size_t)o1 * 16u);
                    const P2 c2 = ldg128p(trisT + (size_t)o2 * 16u);
                    const float4 ro = lds128(rays_a + own * 32u), rd = lds128(rays_a + own * 32u + 16u);
                    const float oorg[3] = {ro.x, ro.y, ro.z}, odir[3] = {rd.x, rd.y, rd.z};
                    hit = pair_accept_occluded(pair_mt(K, c0, c1, c2, oorg, odir), 2u * item + 1u < ntris);
                }
                const unsigned hits = __ballot_sync(FULL, hit);
                if (owner) {
                    const uint32_t room = 32u - excl, took = cnt < room ? cnt : room;
                    const unsigned mine = (FULL >> (32u - took)) << excl;         // took >= 1
                    if (hits & mine) { retire(true); cur = kIdle; }               // occluded: bvh.c:850 commits, the query is decided
                    else {
                        prog += took;
                        if (prog == nitems) {                                     // leaf exhausted without a hit: pop, or the ray escapes
                            prog = 0;
                            if (spa < kRow) { retire(false); cur = kIdle; }
                            else { spa -= kRow; cur = lds32(stack_a + spa); }
                        }
                    }
                }
            }
            if (cur < kIdle) {
                // ---- node step: bvh.c:1153-1179 with best_t == 1e38 (no hit yet)
                const Node32 *p = S.nodes + cur;
                const P4 a = ldg256p(p), b = ldg256p(reinterpret_cast<const char *>(p) + 32);
                const pk_t ox = pkb(org[0]), oy = pkb(org[1]), oz = pkb(org[2]);
                const pk_t ix = pkb(inv[0]), iy = pkb(inv[1]), iz = pkb(inv[2]);
                const bool sx = (sgn & 1u) != 0u, sy = (sgn & 2u) != 0u, sz = (sgn & 4u) != 0u;
                // An occlusion query may visit the children in ANY order (pool.cuh): the answer is the OR over the same leaves.  Which
                // child first finds a hit soonest?  kOrder 1: the one the ray runs through for longer (hit probability ~ path length
                // x density) -- best when the records are L2-resident (C3: 1123 -> 1195 Mrays/s), but it sends the rays of one point to
                // distant subtrees, which costs DRAM traffic on a scene beyond L2 (10 M triangles: 975 -> 914); kOrder 2: the nearer
                // box by its entry distance (1137 / 989): the host picks by the size of the records.  kOrder 0: the reference's
                // near child = child[sign[axis0]], kept for A/B.
                float tn0, tf0, tn1, tf1;
                slab_pk_t(K, a.v[0], a.v[2], b.v[0], ox, oy, oz, ix, iy, iz, sx, sy, sz, tn0, tf0);
                slab_pk_t(K, a.v[1], a.v[3], b.v[1], ox, oy, oz, ix, iy, iz, sx, sy, sz, tn1, tf1);
                const bool h0 = (tf0 > 0.0f) && (tn0 <= tf0) && (tn0 < 1.0e38f), h1 = (tf1 > 0.0f) && (tn1 <= tf1) && (tn1 < 1.0e38f);
                const uint32_t c0 = (uint32_t)b.v[2], c1 = (uint32_t)(b.v[2] >> 32), axis = (uint32_t)b.v[3];
                const bool longer = (tf1 - fmaxf(tn1, 0.0f)) > (tf0 - fmaxf(tn0, 0.0f)), nearer = tn1 < tn0;
                const bool order = kOrder == 1 ? longer
                                 : kOrder == 2 ? nearer
                                 : kOrder == 3 ? (cur < S.top_count ? nearer : longer)       // experiment: locality at the top, probability below
                                               : ((sgn >> axis) & 1u) != 0u;
                const uint32_t near = order ? c1 : c0, far = order ? c0 : c1;
                uint32_t next;
                const bool both = h0 && h1, none = !h0 && !h1;                   // predicated: measured 1080 vs 1072 Mrays/s for the branchy form
                const bool pop = none && (spa >= kRow);
                if (both) sts32(stack_a + spa, far);
                const uint32_t popped = pop ? lds32(stack_a + spa - kRow) : kIdle;
                spa = spa + (both ? kRow : 0u) - (pop ? kRow : 0u);
                next = both ? near : (none ? popped : (h0 ? c0 : c1));
                if (next == kIdle) retire(false);                                 // stack ran dry
#if B200_PF_LEAF
                // experiment (X1: bulk / TMA staging where the records really come from HBM): the lane will stand in this leaf until a
                // leaf round takes it -- ask the TMA unit to pull the leaf's whole slab (<= 768 B) into L2 meanwhile, one instruction
                if ((int32_t)next < 0 && next != kIdle) {
                    const uint32_t nt = ((next >> kLeafShift) & 15u) + 1u;
                    const char *slab = trisT + (size_t)(next & kSlotMask) * 48u;
                    const uint32_t bytes = ((nt + 3u) >> 2) * 192u;               // round_up(ntris, 4) slots of 48 bytes
#if B200_PF_LEAF == 2
                    for (uint32_t o = 0; o < bytes; o += 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(slab + o));
#else
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(slab), "r"(bytes) : "memory");
#endif
                }
#endif
#if B200_PF_FAR
                // ... and the record of the far child the lane has just pushed (an inner node: 64 bytes)
                if (both && far < kIdle) asm volatile("prefetch.global.L2 [%0];" :: "l"(S.nodes + far));
#endif
                prog = 0;
                cur = next;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// closest hit, fp32 records: pool_closest.cuh's scheme (a lane waits in its leaf until the leaf is committed; the leaf's winner is one
// shared-memory atomicMin on (bits of |t|) << 32 | (31 - lane); exactness argument there) with the same diet as the kernel above.
struct Close32Warp {
    float4 rays[64];                     // ray slots
    uint2  desc[32];                     // leaf-round descriptors
    unsigned long long key[32];          // per owner: smallest (|t| bits, 31 - lane) offered this round
    float4 res[32];                      // per item lane: (t, u, v, prim bits) of its accepted triangle
    float4 leaf[32];                     // per lane: the leaf-local winner so far (t, u, v, prim bits); its t is also in a register
    float4 best[32];                     // per lane: the committed closest hit so far; its t is also in a register
};
#ifndef B200_CLOSE_THREADS
#define B200_CLOSE_THREADS 128
#endif
constexpr int kCloseThreads = B200_CLOSE_THREADS;      // CTA size of the closest-hit kernel (A/B knob with B200_CLOSE_CTAS)
template <int kCap> struct Close32Smem {
    uint32_t    stack[kCap * kCloseThreads];
    Close32Warp warp[kCloseThreads / 32];
};
__device__ __forceinline__ void atom_min_shared_u64(uint32_t a, unsigned long long v)
{ asm volatile("red.shared.min.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a)
{ unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }

// Occupancy, measured on the C3 batch (scripts/gpu_r2u.sh): CTAs of 128 threads, 7 per SM at 72 registers (28 warps, no spills):
// 944 Mrays/s; 256 x 3 at 77 registers (24 warps): 892; 128 x 6 and 192 x 4 (24 warps): 891 / 892; 256 x 4 and 128 x 8 at 64
// registers (32 warps, spills): 838 / 837.  The kernel waits on memory with 61 % of its issue slots used: warps in flight are what
// it is short of, until the register allocation starts to spill.
#ifndef B200_CLOSE_CTAS
#define B200_CLOSE_CTAS 7
#endif
template <int kCap>
__global__ void __launch_bounds__(kCloseThreads, B200_CLOSE_CTAS)
closest_pool32_kernel(const SceneView<float> S, const char *__restrict__ trisT, const float *__restrict__ rays, const uint32_t n,
                      const uint32_t chunk, ri_b200_hit_f32 *__restrict__ hits_out, unsigned int *__restrict__ work_counter, const PackK K)
{
    constexpr unsigned FULL = 0xffffffffu;
#ifndef B200_CLOSE_REFILL
#define B200_CLOSE_REFILL 8      // idle lanes that trigger a refill: 1 / 2 / 4 / 8 / 12 -> 876 / 915 / 945 / 958 / 948 Mrays/s on C3
#endif
    constexpr uint32_t kRefillAt = B200_CLOSE_REFILL;
    constexpr uint32_t kRow = kCloseThreads * 4u;
    __shared__ __align__(16) Close32Smem<kCap> sm;
    const unsigned lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = (2u << lane) - 1u;
    uint32_t sm_a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sm_a) : "l"(&sm));
    const uint32_t stack_a = sm_a + (uint32_t)offsetof(Close32Smem<kCap>, stack);
    const uint32_t rays_a = sm_a + (uint32_t)offsetof(Close32Smem<kCap>, warp) + (wbase >> 5) * (uint32_t)sizeof(Close32Warp);
    const uint32_t desc_a = rays_a + (uint32_t)offsetof(Close32Warp, desc);
    const uint32_t key_a = rays_a + (uint32_t)offsetof(Close32Warp, key);
    const uint32_t res_a = rays_a + (uint32_t)offsetof(Close32Warp, res);
    const uint32_t leaf_a = rays_a + (uint32_t)offsetof(Close32Warp, leaf) + lane * 16u;      // (u, v, prim) of the two records live in
    const uint32_t best_a = rays_a + (uint32_t)offsetof(Close32Warp, best) + lane * 16u;      // shared memory, their t in registers

    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;

    uint32_t cur = kIdle, prog = 0, idx = 0, spa = threadIdx.x * 4u, sgn = 0;
    float org[3] = {0.0f, 0.0f, 0.0f}, inv[3] = {0.0f, 0.0f, 0.0f};
    float best_t = 1.0e38f, tl = 1.0e38f;        // tl == 1e38 <=> the leaf has no accepted triangle yet (a t of exactly 1e38 never commits: bvh.c:850)

    auto retire = [&]() {                        // bvh.c:1187
        const float4 b = lds128(best_a);
        RayIO<float>::store(hits_out, idx, best_t < 1.0e38f, best_t, b.y, b.z, __float_as_uint(b.w));
    };
    auto enter = [&](const uint32_t word) {      // step onto `word`; a leaf starts with a fresh leaf-local record (bvh.c:833-836)
        cur = word; prog = 0;
        tl = 1.0e38f;
    };

    for (;;) {
        unsigned idle = __ballot_sync(FULL, cur == kIdle);
        while (idle && !exhausted) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = (n - base < chunk) ? n : base + chunk;
            }
            const unsigned avail = chunk_end - chunk_next;
            const unsigned n_idle = __popc(idle);
            const unsigned take = n_idle < avail ? n_idle : avail;
            const unsigned rank = __popc(idle & lt_mask);
            if (cur == kIdle && rank < take) {
                idx = chunk_next + rank;
                float dir[3];
                RayIO<float>::load(rays, idx, org, dir);
                sts128(rays_a + lane * 32u, make_float4(org[0], org[1], org[2], 0.0f));
                sts128(rays_a + lane * 32u + 16u, make_float4(dir[0], dir[1], dir[2], 0.0f));
                best_t = 1.0e38f;
                const bool sx = dir[0] < 0.0f, sy = dir[1] < 0.0f, sz = dir[2] < 0.0f;
                sgn = (sx ? 1u : 0u) | (sy ? 2u : 0u) | (sz ? 4u : 0u);
#pragma unroll
                for (int k = 0; k < 3; ++k)      // bvh.c:473-497
                    inv[k] = (fabsf(dir[k]) > 1.0e-14f) ? 1.0f / dir[k] : ((dir[k] < 0.0f) ? -FLT_MAX : FLT_MAX);
                float tmin;
                const bool in_scene = (S.root_word != kDoneWord) &&
                    slab<float>(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], org, inv, sx, sy, sz, tmin);
                spa = threadIdx.x * 4u;
                if (in_scene) enter(S.root_word);
                else retire();                   // bvh.c:446 / 522-526: miss without traversal
            }
            chunk_next += take;
            idle = __ballot_sync(FULL, cur == kIdle);
        }
        if (idle == FULL) break;

        for (;;) {
            const bool in_leaf = (int32_t)cur < 0;
            const uint32_t nitems = (((cur >> kLeafShift) & 15u) + 2u) >> 1;
            const uint32_t cnt = in_leaf ? nitems - prog : 0u;
            const uint32_t total = __reduce_add_sync(FULL, cnt);
            const unsigned owners = __ballot_sync(FULL, in_leaf);
            const unsigned in_node = __ballot_sync(FULL, cur < kIdle);
            const unsigned n_node = __popc(in_node);
            if ((in_node | owners) == 0u) break;
            if (!exhausted && (uint32_t)__popc(~(in_node | owners)) >= kRefillAt) break;

            if (total >= 32u || total > n_node) {
                // ---- leaf round: items 0..31 of the pool, one per lane
                uint32_t excl = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    excl += (uint32_t)__popc(ballot_bit(cnt, 1u << b) & lt_mask) << b;
                const bool owner = in_leaf && excl < 32u;
                const unsigned starts = __reduce_or_sync(FULL, owner ? (1u << excl) : 0u);
                if (owner) {
                    sts64(desc_a + (uint32_t)__popc(owners & lt_mask) * 8u, make_uint2(cur, lane | ((prog - excl + 64u) << 8)));
                    sts_u64(key_a + lane * 8u, ~0ull);
                }
                __syncwarp();
                if (lane < total) {
                    const uint2 d = lds64(desc_a + (uint32_t)__popc(starts & le_mask) * 8u - 8u);
                    const unsigned own = d.y & 31u;
                    const uint32_t item = lane + (d.y >> 8) - 64u;
                    const uint32_t ntris = ((d.x >> kLeafShift) & 15u) + 1u, slot0 = d.x & kSlotMask;
                    const uint32_t m = ((ntris + 3u) >> 2) << 1;
                    const uint32_t o0 = slot0 * 3u + item * 2u, o1 = o0 + 2u * m, o2 = slot0 * 3u + 4u * m + item;
                    const float4 ro = lds128(rays_a + own * 32u), rd = lds128(rays_a + own * 32u + 16u);
                    const float oorg[3] = {ro.x, ro.y, ro.z}, odir[3] = {rd.x, rd.y, rd.z};
                    float t, u, v;
                    uint32_t prim;
                    pair_closest(K, trisT + (size_t)o0 * 16u, trisT + (size_t)o1 * 16u, trisT + (size_t)o2 * 16u, 2u * item + 1u < ntris, oorg, odir,
                                 t, u, v, prim);
                    if (prim != 0xffffffffu) {   // the item accepted a triangle: offer it to the owner
                        sts128(res_a + lane * 16u, make_float4(t, u, v, __uint_as_float(prim)));
                        atom_min_shared_u64(key_a + own * 8u, ((unsigned long long)(__float_as_uint(t) & 0x7fffffffu) << 32) | (unsigned long long)(31u - lane));
                    }
                }
                __syncwarp();
                if (owner) {
                    const uint32_t room = 32u - excl, took = cnt < room ? cnt : room;
                    const unsigned long long key = lds_u64(key_a + lane * 8u);
                    if (key != ~0ull) {          // winner of this round's items of my leaf; accepted unless t > t_leaf (bvh.c:780)
                        const float4 r = lds128(res_a + (31u - (unsigned)(key & 31ull)) * 16u);
                        if (!(r.x > tl)) { tl = r.x; sts128(leaf_a, r); }
                    }
                    prog += took;
                    if (prog == nitems) {        // leaf finished: commit (bvh.c:850), then pop or retire
                        if (tl < best_t) { best_t = tl; sts128(best_a, lds128(leaf_a)); }      // tl < best_t <= 1e38: the leaf accepted a triangle
                        if (spa < kRow) { retire(); cur = kIdle; }
                        else { spa -= kRow; enter(lds32(stack_a + spa)); }
                    }
                }
                __syncwarp();                    // key / res are rewritten by the next round
            }
            if (cur < kIdle) {
                // ---- node step: bvh.c:1153-1179
                const Node32 *p = S.nodes + cur;
                const P4 a = ldg256p(p), b = ldg256p(reinterpret_cast<const char *>(p) + 32);
                const pk_t ox = pkb(org[0]), oy = pkb(org[1]), oz = pkb(org[2]);
                const pk_t ix = pkb(inv[0]), iy = pkb(inv[1]), iz = pkb(inv[2]);
                const bool sx = (sgn & 1u) != 0u, sy = (sgn & 2u) != 0u, sz = (sgn & 4u) != 0u;
                const bool h0 = slab_pk(K, a.v[0], a.v[2], b.v[0], ox, oy, oz, ix, iy, iz, sx, sy, sz, best_t);
                const bool h1 = slab_pk(K, a.v[1], a.v[3], b.v[1], ox, oy, oz, ix, iy, iz, sx, sy, sz, best_t);
                const uint32_t c0 = (uint32_t)b.v[2], c1 = (uint32_t)(b.v[2] >> 32), axis = (uint32_t)b.v[3];
                const bool order = ((sgn >> axis) & 1u) != 0u;
                const uint32_t near = order ? c1 : c0, far = order ? c0 : c1;
                const bool both = h0 && h1, none = !h0 && !h1;
                const bool pop = none && (spa >= kRow);
                if (both) sts32(stack_a + spa, far);
                const uint32_t popped = pop ? lds32(stack_a + spa - kRow) : kIdle;
                spa = spa + (both ? kRow : 0u) - (pop ? kRow : 0u);
                const uint32_t next = both ? near : (none ? popped : (h0 ? c0 : c1));
                if (next == kIdle) { retire(); cur = kIdle; }
                else enter(next);
            }
        }
    }
}

}  // namespace b200

// Pooled occlusion (any-hit) traverser: the production kernel for ri_b200_occluded_* and the AO transport.
//
// The persistent kernel of persistent.cuh gives every lane one ray and lets the warp vote between a node step and a leaf
// step; with the reference's tree (leaves of up to 16 triangles) a ray spends as long inside leaves as between them, so
// each step runs with about half of the lanes.  Here a lane still OWNS one ray and walks the inner nodes for it, but the
// triangle tests of all the leaves the warp is standing in are POOLED: every leaf contributes its remaining test items
// (fp32: a pair of triangle slots, fp64: one slot), an exclusive prefix sum built from bit-sliced ballots numbers them, the
// owners publish (leaf word, first item, lane) under their ordinal in shared memory and one REDUX.OR gives the bitmap of
// first items; a leaf round then hands items 0..31 to the 32 lanes -- lane g finds its owner with one popc over that
// bitmap, reads the owner's ray from its shared-memory slot, tests its item, and the owners read the verdicts back from
// one __ballot_sync.  A leaf round therefore runs with 32 lanes whenever 32 items are waiting, and node steps run with
// every lane that is not waiting for a leaf.  (The first version found owners by a shuffle binary search and fetched rays
// with shuffles: 906 Mrays/s on the C3 batch against 1032 for this one, on the first 4 Mi rays.)
//
// Why this is still the reference's answer, bit for bit: an occlusion query returns `bvh_traverse(...) != 0`
// (bvh.c:1187), i.e. whether ANY visited leaf holds a triangle that triangle_isect accepts with t < 1e38.  Before the
// first acceptance the closest-so-far values are their initial 1e38 (bvh.c:1114, 833), so every box test and every
// triangle test of the query is evaluated against the same constants in whatever order it happens: the set of leaves
// reached, each per-triangle verdict and hence the OR are order-independent.  Each lane performs its ray's node steps in
// the reference's order with the reference's arithmetic (slab_mm / tri_test_bf of trace.cuh); only the triangle tests of
// one leaf are spread over several lanes.  (A NaN t, the one value that makes `t > t_leaf` order-dependent inside a leaf,
// can only turn a later rejection `t > 1e38` into an acceptance with t >= 1e38, which does not commit: bvh.c:850.)
//
// Memory: node records as in persistent.cuh (2 x LDG.256 per node).  Triangles come from the leaf-TRANSPOSED copy
// (FlatTree::tris32t / tris64t): chunk k of item j of a leaf whose rows hold m items sits at slot0*sizeof(slot) + (k*m + j)*32
// (rows start on 64-byte boundaries for fp32, 128-byte ones for fp64), so
// the lanes that test consecutive items of one leaf read consecutive 32-byte chunks -- one L1 wavefront per 128-byte
// line instead of one per lane.
#pragma once

namespace b200 {

template <typename Real> struct PoolLeaf;

template <> struct PoolLeaf<float> {                 // item = two triangle slots (96 B = 3 chunks)
    static constexpr int kCntBits = 4;               // items per leaf <= 8
    static __device__ __forceinline__ uint32_t items(uint32_t ntris) { return (ntris + 1u) >> 1; }
    static __device__ __forceinline__ bool test(const PackK &K, const char *trisT, uint32_t slot0, uint32_t ntris, uint32_t j,
                                                const float org[3], const float dir[3])
    {
        const uint32_t m = ((ntris + 3u) >> 2) << 1;                                      // row length in pairs: the leaf owns round_up(ntris, 4) slots
        const uint32_t o0 = slot0 * 3u + j * 2u, o1 = o0 + 2u * m, o2 = slot0 * 3u + 4u * m + j;     // 16-byte units: < 2^29; row 2 holds 16 B per item
        return pair_occluded(K, trisT + (size_t)o0 * 16u, trisT + (size_t)o1 * 16u, trisT + (size_t)o2 * 16u, 2u * j + 1u < ntris, org, dir);
    }
};

template <> struct PoolLeaf<double> {                // item = one triangle slot (96 B = 3 chunks)
    static constexpr int kCntBits = 5;               // items per leaf <= 16
    static __device__ __forceinline__ uint32_t items(uint32_t ntris) { return ntris; }
    static __device__ __forceinline__ bool test(const PackK &, const char *trisT, uint32_t slot0, uint32_t ntris, uint32_t j,
                                                const double org[3], const double dir[3])
    {
        const uint32_t m = (ntris + 3u) & ~3u;       // slots owned by the leaf = row length
        const uint32_t o0 = slot0 * 3u + j, o1 = o0 + m, o2 = o1 + m;                       // 32-byte units: < 2^29
        const D4 q0 = ldg256d(trisT + (size_t)o0 * 32u), q1 = ldg256d(trisT + (size_t)o1 * 32u), q2 = ldg256d(trisT + (size_t)o2 * 32u);
        TriRegs<double> a;
        a.v0[0] = q0.v[0]; a.v0[1] = q0.v[1]; a.v0[2] = q0.v[2]; a.prim = 0;
        a.e1[0] = q1.v[0]; a.e1[1] = q1.v[1]; a.e1[2] = q1.v[2];
        a.e2[0] = q2.v[0]; a.e2[1] = q2.v[1]; a.e2[2] = q2.v[2];
        double tl = Prec<double>::inf(), ul = 0.0, vl = 0.0;
        uint32_t tprim = 0xffffffffu;
        tri_test_bf<double>(a, org, dir, true, tl, ul, vl, tprim);
        return tl < Prec<double>::inf();
    }
};

// a lane's ray (org, dir) in shared memory, where the lanes that test its leaf items pick it up
template <typename Real> struct RaySlot;
template <> struct RaySlot<float> {
    static constexpr uint32_t kBytes = 32;
    static __device__ __forceinline__ void store(char *base, unsigned slot, const float org[3], const float dir[3])
    {
        float4 *p = reinterpret_cast<float4 *>(base + slot * kBytes);
        p[0] = make_float4(org[0], org[1], org[2], 0.0f); p[1] = make_float4(dir[0], dir[1], dir[2], 0.0f);
    }
    static __device__ __forceinline__ void load(const char *base, unsigned slot, float org[3], float dir[3])
    {
        const float4 *p = reinterpret_cast<const float4 *>(base + slot * kBytes);
        const float4 a = p[0], b = p[1];
        org[0] = a.x; org[1] = a.y; org[2] = a.z; dir[0] = b.x; dir[1] = b.y; dir[2] = b.z;
    }
};
template <> struct RaySlot<double> {
    static constexpr uint32_t kBytes = 48;
    static __device__ __forceinline__ void store(char *base, unsigned slot, const double org[3], const double dir[3])
    {
        double2 *p = reinterpret_cast<double2 *>(base + slot * kBytes);
        p[0] = make_double2(org[0], org[1]); p[1] = make_double2(org[2], dir[0]); p[2] = make_double2(dir[1], dir[2]);
    }
    static __device__ __forceinline__ void load(const char *base, unsigned slot, double org[3], double dir[3])
    {
        const double2 *p = reinterpret_cast<const double2 *>(base + slot * kBytes);
        const double2 a = p[0], b = p[1], c = p[2];
        org[0] = a.x; org[1] = a.y; org[2] = b.x; dir[0] = b.y; dir[1] = c.x; dir[2] = c.y;
    }
};
// __ballot_sync(full, (x & bit) != 0) pinned to LOP3-with-predicate + VOTE (the compiler's own form is shift, and, compare, vote)
__device__ __forceinline__ unsigned ballot_bit(uint32_t x, uint32_t bit)
{
    unsigned r;
    asm volatile("{ .reg .pred p; .reg .b32 t; and.b32 t, %1, %2; setp.ne.u32 p, t, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }"
                 : "=r"(r) : "r"(x), "r"(bit));
    return r;
}
template <typename Real> constexpr size_t pool_smem_bytes(int stack_cap)
{ return (size_t)stack_cap * kBlock * sizeof(uint32_t) + (size_t)kBlock * (RaySlot<Real>::kBytes + sizeof(uint2)); }

// x 200 ns ~ 0.1 s and more: a piece of the upload lands every 0.15 ms.  The cap is what ends the launch when the copies cannot
// run beside the kernel at all (a profiler that serialises the two streams): the host then falls back to one launch per piece.
constexpr unsigned kPoolSpinCap = 500u * 1000u;

// EXPERIMENT (B200_POOL_TOPSMEM=1, fp32): the first kTopNodes inner nodes (breadth-first: the top levels of the tree, which every ray
// walks) staged once per CTA into shared memory by ONE bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA engine), node
// steps on them served by LDS.128 instead of LDG.256.  Measured slower than the plain kernel -- see DESIGN.md -- and kept as evidence.
constexpr uint32_t kTopNodes = 256;

__device__ __forceinline__ void load_node_shared(const char *p, NodeRegs<float> &r)
{
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 16),
                 c = *reinterpret_cast<const float4 *>(p + 32);
    const uint4 d = *reinterpret_cast<const uint4 *>(p + 48);
    r.x[0] = a.x; r.x[1] = a.y; r.x[2] = a.z; r.x[3] = a.w;
    r.y[0] = b.x; r.y[1] = b.y; r.y[2] = b.z; r.y[3] = b.w;
    r.z[0] = c.x; r.z[1] = c.y; r.z[2] = c.z; r.z[3] = c.w;
    r.c0 = d.x; r.c1 = d.y; r.axis = d.z;
}
__device__ __forceinline__ void load_node_shared(const char *, NodeRegs<double> &) {}

// one node step on packed pairs: two LDG.256 bring (lo, hi) of both child boxes per axis as aligned register pairs
__device__ __forceinline__ void node_step_pk(const PackK &K, const Node32 *p, const float org[3], const float inv[3], bool sx, bool sy, bool sz,
                                             float best_t, bool &h0, bool &h1, uint32_t &c0, uint32_t &c1, uint32_t &axis)
{
    const P4 a = ldg256p_keep(p), b = ldg256p_keep(reinterpret_cast<const char *>(p) + 32);
    const pk_t ox = pkb(org[0]), oy = pkb(org[1]), oz = pkb(org[2]);
    const pk_t ix = pkb(inv[0]), iy = pkb(inv[1]), iz = pkb(inv[2]);
    h0 = slab_pk(K, a.v[0], a.v[2], b.v[0], ox, oy, oz, ix, iy, iz, sx, sy, sz, best_t);
    h1 = slab_pk(K, a.v[1], a.v[3], b.v[1], ox, oy, oz, ix, iy, iz, sx, sy, sz, best_t);
    c0 = (uint32_t)b.v[2]; c1 = (uint32_t)(b.v[2] >> 32); axis = (uint32_t)b.v[3];
}
__device__ __forceinline__ void node_step_pk(const PackK &, const Node64 *, const double *, const double *, bool, bool, bool, double,
                                             bool &, bool &, uint32_t &, uint32_t &, uint32_t &) {}

template <typename Real, int kMinBlocks, bool kTopSmem = false>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
occluded_pool_kernel(const SceneView<Real> S, const char *__restrict__ trisT, const Real *__restrict__ rays, const uint32_t n,
                     const uint32_t chunk, uint8_t *__restrict__ occ, uint32_t *__restrict__ counts, const uint32_t rays_per_count,
                     unsigned int *__restrict__ work_counter, const uint32_t refill_at, const uint32_t stack_cap,
                     const unsigned int *__restrict__ ready, unsigned int *__restrict__ fault, const PackK K)
{
    using P = Prec<Real>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint32_t s_stack[];          // [stack_cap][kBlock] words, then ray slots, then descriptors
    uint32_t *stk = s_stack + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = (2u << lane) - 1u;
    char *s_tail = reinterpret_cast<char *>(s_stack + (size_t)stack_cap * kBlock);
    char *s_rays = s_tail + (size_t)(threadIdx.x & ~31u) * RaySlot<Real>::kBytes;              // this warp's 32 ray slots
    uint2 *s_desc = reinterpret_cast<uint2 *>(s_tail + (size_t)kBlock * RaySlot<Real>::kBytes) + (threadIdx.x & ~31u);
    char *s_top = s_tail + (size_t)kBlock * (RaySlot<Real>::kBytes + sizeof(uint2));            // kTopSmem: [kTopNodes] node records, then the mbarrier
    const uint32_t ntop = kTopSmem ? (S.top_count < kTopNodes ? S.top_count : kTopNodes) : 0u;
    if (kTopSmem && ntop) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_top + kTopNodes * sizeof(Node32));
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_top);
        const uint32_t bytes = ntop * (uint32_t)sizeof(Node32);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(S.nodes), "r"(bytes), "r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    }

    uint32_t chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool exhausted = false;                      // warp-uniform

    // per-lane ray state.  cur: kIdle | inner-node index | leaf word; prog: items of that leaf already tested
    uint32_t cur = kIdle, prog = 0, idx = 0, sp = 0;
    Real org[3], dir[3], inv[3];
    bool sx = false, sy = false, sz = false;
    org[0] = org[1] = org[2] = dir[0] = dir[1] = dir[2] = inv[0] = inv[1] = inv[2] = Real(0);

    auto retire = [&](const bool hit) {
        if (counts) { if (hit) atomicAdd(&counts[idx / rays_per_count], 1u); }
        else occ[idx] = hit ? 1 : 0;
    };

    for (;;) {
        // ------------------------------------------------------------------ fetch (as in persistent.cuh)
        // quad mode (refill_at bit 16): new rays go only to ALIGNED LANE QUADS that are idle as a whole, four consecutive rays each.
        // An LDG.256 is served four lanes per L1 pass, and four rays of one AO point that start at the root together walk the
        // same nodes until their paths part -- one wavefront per node record instead of four while they do.
#ifdef B200_NO_QUAD
        const bool quad_mode = false;
#else
        const bool quad_mode = ((refill_at >> 16) & 1u) != 0u;
#endif
        auto fetchable = [&](unsigned m) -> unsigned {
            if (!quad_mode) return m;
            const unsigned q = m & (m >> 1) & (m >> 2) & (m >> 3) & 0x11111111u;
            return q * 15u;
        };
        unsigned idle_all = __ballot_sync(FULL, cur == kIdle);
        unsigned idle = fetchable(idle_all);
        while (idle && !exhausted) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = (n - base < chunk) ? n : base + chunk;
                if (ready) {
                    // streamed upload (host-buffer entry points): the copy engine is still writing the batch while this kernel
                    // runs; `*ready` = number of leading rays that have landed, bumped in stream order after every piece
                    if (lane == 0) {
                        unsigned spins = 0, have;
                        for (;;) {
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(ready) : "memory");
                            if (have >= chunk_end) break;
                            __nanosleep(200);
                            if (++spins > kPoolSpinCap) { atomicExch(fault, 1u); break; }     // upload never arrived: give up loudly
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
            if (((idle >> lane) & 1u) && rank < take) {
                idx = chunk_next + rank;
                if (ready) RayIO<Real>::load_coherent(rays, idx, org, dir);      // the copy engine is still writing this buffer: no ld.global.nc
                else RayIO<Real>::load(rays, idx, org, dir);
                RaySlot<Real>::store(s_rays, lane, org, dir);
                sx = dir[0] < Real(0); sy = dir[1] < Real(0); sz = dir[2] < Real(0);
#pragma unroll
                for (int k = 0; k < 3; ++k)      // bvh.c:473-497
                    inv[k] = (P::rabs(dir[k]) > P::eps()) ? Real(1) / dir[k] : ((dir[k] < Real(0)) ? -P::vmax() : P::vmax());
                Real tmin;
                const bool in_scene = (S.root_word != kDoneWord) &&
                    slab<Real>(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], org, inv, sx, sy, sz, tmin);
                sp = 0; prog = 0;
                if (in_scene) cur = S.root_word;
                else retire(false);              // bvh.c:446 / 522-526: miss without traversal
            }
            chunk_next += take;
            idle_all = __ballot_sync(FULL, cur == kIdle);
            idle = fetchable(idle_all);
        }
        if (idle_all == FULL) break;             // nothing in flight and nothing left to fetch

        // ------------------------------------------------------------------ traverse
        for (;;) {
            const bool in_leaf = (cur & kLeafFlag) != 0u;
            const uint32_t ntris = ((cur >> kLeafShift) & 15u) + 1u;
            const uint32_t nitems = PoolLeaf<Real>::items(ntris);
            const uint32_t cnt = in_leaf ? nitems - prog : 0u;        // >= 1 for a lane standing in a leaf
            const uint32_t total = __reduce_add_sync(FULL, cnt);
            const unsigned owners = __ballot_sync(FULL, in_leaf);
            const unsigned in_node = __ballot_sync(FULL, cur < kIdle);            // inner-node indices are < kIdle, leaf words above
            const unsigned n_node = __popc(in_node);
            if (n_node == 0u && total == 0u) break;
            if (!exhausted && (uint32_t)__popc(fetchable(~(in_node | owners))) >= (refill_at & 255u)) break;

            if (total >= ((refill_at >> 8) & 255u) || total > n_node) {           // leaf-round threshold rides in the upper bits (B200_LEAF_AT, default 32)
                // ---- leaf round: items 0..31 of the pool, one per lane.
                // exclusive prefix sum of cnt (<= 16) from bit-sliced ballots: no dependent shuffle chain
                uint32_t excl = 0;
#pragma unroll
                for (int b = 0; b < PoolLeaf<Real>::kCntBits; ++b)
                    excl += (uint32_t)__popc(ballot_bit(cnt, 1u << b) & lt_mask) << b;
                const bool owner = in_leaf && excl < 32u;             // my leaf has items in this round
                // bit e of `starts` = an owner's first item is item e; owners publish a descriptor under their ordinal
                const unsigned starts = __reduce_or_sync(FULL, owner ? (1u << excl) : 0u);
                if (owner) s_desc[__popc(owners & lt_mask)] = make_uint2(cur, lane | ((prog - excl + 64u) << 8));
                __syncwarp();
                const bool have = lane < total;
                bool hit = false;
                if (have) {
                    const uint2 d = s_desc[__popc(starts & le_mask) - 1u];
                    const unsigned own = d.y & 31u;
                    const uint32_t item = lane + (d.y >> 8) - 64u;    // item number inside the owner's leaf
                    Real oorg[3], odir[3];
                    RaySlot<Real>::load(s_rays, own, oorg, odir);
                    hit = PoolLeaf<Real>::test(K, trisT, d.x & kSlotMask, ((d.x >> kLeafShift) & 15u) + 1u, item, oorg, odir);
                }
                const unsigned hits = __ballot_sync(FULL, hit);
                if (owner) {
                    const uint32_t took = (cnt < 32u - excl) ? cnt : 32u - excl;
                    const unsigned mine = ((took >= 32u) ? FULL : ((1u << took) - 1u)) << excl;
                    if (hits & mine) { retire(true); cur = kIdle; }   // occluded: bvh.c:850 commits, the query is decided
                    else {
                        prog += took;
                        if (prog == nitems) {                         // leaf exhausted without a hit: pop, or the ray escapes
                            prog = 0;
                            if (sp == 0u) { retire(false); cur = kIdle; }
                            else { --sp; cur = stk[sp * kBlock]; }
                        }
                    }
                }
            }
            if (cur < kIdle) {
                // ---- node step: bvh.c:1153-1179 with best_t == 1e38 (no hit yet)
                bool h0, h1;
                uint32_t c0, c1, axis;
                if (sizeof(Real) == 4 && !kTopSmem) {                 // packed fp32: (lo, hi) of each child box as a register pair (packed.cuh)
                    node_step_pk(K, S.nodes + cur, org, inv, sx, sy, sz, P::inf(), h0, h1, c0, c1, axis);
                } else {
                    NodeRegs<Real> nd;
                    if (kTopSmem && cur < ntop) load_node_shared(s_top + (size_t)cur * sizeof(Node32), nd);
                    else load_node_wide(S.nodes + cur, nd);
                    h0 = slab_mm<Real>(nd.x[0], nd.x[1], nd.y[0], nd.y[1], nd.z[0], nd.z[1], org, inv, sx, sy, sz, P::inf());
                    h1 = slab_mm<Real>(nd.x[2], nd.x[3], nd.y[2], nd.y[3], nd.z[2], nd.z[3], org, inv, sx, sy, sz, P::inf());
                    c0 = nd.c0; c1 = nd.c1; axis = nd.axis;
                }
                const bool order = (axis == 0) ? sx : ((axis == 1) ? sy : sz);           // near child = child[sign[axis0]]
                const bool both = h0 && h1, none = !h0 && !h1;
                const bool pop = none && (sp != 0u);
                if (both) stk[sp * kBlock] = order ? c0 : c1;
                const uint32_t popped = pop ? stk[(sp - 1u) * kBlock] : kIdle;
                sp = sp + (both ? 1u : 0u) - (pop ? 1u : 0u);
                const uint32_t one = h0 ? c0 : c1;
                const uint32_t next = both ? (order ? c1 : c0) : (none ? popped : one);
                if (next == kIdle) retire(false);                     // stack ran dry
                prog = 0;
                cur = next;
            }
        }
    }
}

}  // namespace b200

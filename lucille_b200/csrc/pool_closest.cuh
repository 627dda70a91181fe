// Pooled closest-hit traverser: pool.cuh's scheme for ri_b200_intersect_* (fp32 records: items are slot pairs, one 64-bit key;
// fp64 records: items are single slots and the winner is found in two steps -- atomicMin on the bits of |t|, then atomicMax on the
// item number among the lanes that hold that t).
//
// A closest-hit query is order-DEPENDENT: the best t so far culls boxes (bvh.c:1038-1044) and leaves commit in visiting order
// (bvh.c:850).  That order is kept: a lane walks its ray's nodes in the reference's order, and when it reaches a leaf it WAITS
// there until every triangle of the leaf has been tested and the leaf's result committed, exactly where bvh_intersect_leaf_node
// would have returned.  Only the tests of one leaf are spread over lanes (and possibly two rounds).  Inside a leaf the reference
// keeps a leaf-local closest t starting at 1e38 and accepts a triangle unless `t > t_leaf`, so the leaf's answer is the smallest
// accepted t and, among equal ones, the LAST triangle in leaf order; a pair's own winner (tri_test_bf on its two triangles, from
// 1e38) combined in leaf order by the same rule gives the same triangle (the case analysis is in DESIGN.md).  Across lanes the
// combination is one 64-bit shared-memory atomicMin on (bits of |t|) << 32 | (31 - item number): smallest t first, latest item
// on ties; -0.0 and +0.0 compare equal as they do in the reference's float comparisons, and the winner's (t, u, v, prim) are
// then read back unchanged.  (A NaN t needs coordinates whose products overflow fp32; such scenes are outside this kernel's
// contract -- B200_POOL_CLOSEST=0 selects the one-lane-per-ray kernel, which is exact for them too.)
#pragma once

namespace b200 {

// pair-local winner of item j of a leaf, tested from t_leaf = 1e38 (bvh.c:833-848 restricted to the pair)
__device__ __forceinline__ void pool_pair_closest(const PackK &K, const char *trisT, uint32_t slot0, uint32_t ntris, uint32_t j, const float org[3],
                                                  const float dir[3], float &tl, float &ul, float &vl, uint32_t &tprim)
{
    const uint32_t m = ((ntris + 3u) >> 2) << 1;
    const uint32_t o0 = slot0 * 3u + j * 2u, o1 = o0 + 2u * m, o2 = slot0 * 3u + 4u * m + j;     // row 2 holds 16 B per item
    pair_closest(K, trisT + (size_t)o0 * 16u, trisT + (size_t)o1 * 16u, trisT + (size_t)o2 * 16u, 2u * j + 1u < ntris, org, dir, tl, ul, vl, tprim);
}

__device__ __forceinline__ void pool_item_closest(const PackK &K, const char *trisT, uint32_t slot0, uint32_t ntris, uint32_t j, const float org[3],
                                                  const float dir[3], float &tl, float &ul, float &vl, uint32_t &tprim)
{ pool_pair_closest(K, trisT, slot0, ntris, j, org, dir, tl, ul, vl, tprim); }

// fp64: one triangle slot per item, tested from t_leaf = 1e38
__device__ __forceinline__ void pool_item_closest(const PackK &, const char *trisT, uint32_t slot0, uint32_t ntris, uint32_t j, const double org[3],
                                                  const double dir[3], double &tl, double &ul, double &vl, uint32_t &tprim)
{
    const uint32_t m = (ntris + 3u) & ~3u;
    const uint32_t o0 = slot0 * 3u + j, o1 = o0 + m, o2 = o1 + m;
    const D4 q0 = ldg256d(trisT + (size_t)o0 * 32u), q1 = ldg256d(trisT + (size_t)o1 * 32u), q2 = ldg256d(trisT + (size_t)o2 * 32u);
    TriRegs<double> a;
    a.v0[0] = q0.v[0]; a.v0[1] = q0.v[1]; a.v0[2] = q0.v[2]; a.prim = (uint32_t)__double_as_longlong(q0.v[3]);
    a.e1[0] = q1.v[0]; a.e1[1] = q1.v[1]; a.e1[2] = q1.v[2];
    a.e2[0] = q2.v[0]; a.e2[1] = q2.v[1]; a.e2[2] = q2.v[2];
    tl = Prec<double>::inf(); ul = 0.0; vl = 0.0; tprim = 0xffffffffu;
    tri_test_bf<double>(a, org, dir, true, tl, ul, vl, tprim);
}

template <typename Real> struct PoolRes;                      // an item's offer to its owner: (t, u, v, prim)
template <> struct PoolRes<float>  { float  t, u, v; uint32_t prim; };
template <> struct PoolRes<double> { double t, u, v; uint32_t prim, pad; };
__device__ __forceinline__ unsigned long long abs_bits(float t)  { return (unsigned long long)(__float_as_uint(t) & 0x7fffffffu); }
__device__ __forceinline__ unsigned long long abs_bits(double t) { return (unsigned long long)__double_as_longlong(t) & 0x7fffffffffffffffull; }

// CTA size / CTAs per SM of this kernel (A/B knobs; the fp32 default path is pool32.cuh's specialised kernel).  Double records, C3
// batch (scripts/gpu_r2v.sh): 128 threads x 5 CTAs at 96 registers (20 warps): 425 Mrays/s; 256 x 2 at 110 registers (16 warps):
// 382; 128 x 4 at 110: 384; 128 x 6 and 256 x 3 at 80 registers (24 warps, spills): 364.
#ifndef B200_PC_THREADS
#define B200_PC_THREADS 128
#endif
#ifndef B200_PC64_CTAS
#define B200_PC64_CTAS 5
#endif
constexpr int kPcThreads = B200_PC_THREADS;
template <typename Real> constexpr size_t pool_closest_smem_bytes(int stack_cap)
{ return (size_t)stack_cap * kPcThreads * sizeof(uint32_t) + (size_t)kPcThreads * (RaySlot<Real>::kBytes + sizeof(uint2) + sizeof(unsigned long long) + sizeof(PoolRes<Real>) + sizeof(uint32_t)); }

// fp32: 3 CTAs x 256 threads per SM at 79 registers: measured 909 Mrays/s on the C3 batch against 877 with 4 CTAs at 64 (spills)
template <typename Real>
__global__ void __launch_bounds__(kPcThreads, sizeof(Real) == 4 ? (768 / kPcThreads) : B200_PC64_CTAS)
closest_pool_kernel(const SceneView<Real> S, const char *__restrict__ trisT, const Real *__restrict__ rays, const uint32_t n,
                    const uint32_t chunk, typename RayIO<Real>::Hit *__restrict__ hits_out, unsigned int *__restrict__ work_counter,
                    const uint32_t refill_at, const uint32_t stack_cap, const PackK K)
{
    using P = Prec<Real>;
    using L = PoolLeaf<Real>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint32_t s_stack[];  // [stack_cap][kPcThreads] words, ray slots, descriptors, keys, results
    uint32_t *stk = s_stack + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = (2u << lane) - 1u;
    char *s_tail = reinterpret_cast<char *>(s_stack + (size_t)stack_cap * kPcThreads);
    char *s_rays = s_tail + (size_t)wbase * RaySlot<Real>::kBytes;
    uint2 *s_desc = reinterpret_cast<uint2 *>(s_tail + (size_t)kPcThreads * RaySlot<Real>::kBytes) + wbase;
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_tail + (size_t)kPcThreads * (RaySlot<Real>::kBytes + sizeof(uint2))) + wbase;
    PoolRes<Real> *s_res = reinterpret_cast<PoolRes<Real> *>(s_tail + (size_t)kPcThreads * (RaySlot<Real>::kBytes + sizeof(uint2) + sizeof(unsigned long long))) + wbase;
    uint32_t *s_win = reinterpret_cast<uint32_t *>(s_tail + (size_t)kPcThreads * (RaySlot<Real>::kBytes + sizeof(uint2) + sizeof(unsigned long long) + sizeof(PoolRes<Real>))) + wbase;

    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;

    uint32_t cur = kIdle, prog = 0, idx = 0, sp = 0, best_prim = 0xffffffffu, tprim = 0xffffffffu;
    Real org[3], dir[3], inv[3], best_t = P::inf(), best_u = Real(0), best_v = Real(0), tl = P::inf(), ul = Real(0), vl = Real(0);
    bool sx = false, sy = false, sz = false;
    org[0] = org[1] = org[2] = dir[0] = dir[1] = dir[2] = inv[0] = inv[1] = inv[2] = Real(0);

    auto retire = [&]() { RayIO<Real>::store(hits_out, idx, best_t < P::inf(), best_t, best_u, best_v, best_prim); };   // bvh.c:1187
    auto enter = [&](const uint32_t word) {      // step onto `word`; a leaf starts with a fresh leaf-local record (bvh.c:833-836)
        cur = word; prog = 0;
        tl = P::inf(); ul = Real(0); vl = Real(0); tprim = 0xffffffffu;
    };

    for (;;) {
        // ------------------------------------------------------------------ fetch
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
                RayIO<Real>::load(rays, idx, org, dir);
                RaySlot<Real>::store(s_rays, lane, org, dir);
                best_t = P::inf(); best_u = Real(0); best_v = Real(0); best_prim = 0xffffffffu;
                sx = dir[0] < Real(0); sy = dir[1] < Real(0); sz = dir[2] < Real(0);
#pragma unroll
                for (int k = 0; k < 3; ++k)      // bvh.c:473-497
                    inv[k] = (P::rabs(dir[k]) > P::eps()) ? Real(1) / dir[k] : ((dir[k] < Real(0)) ? -P::vmax() : P::vmax());
                Real tmin;
                const bool in_scene = (S.root_word != kDoneWord) &&
                    slab<Real>(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], org, inv, sx, sy, sz, tmin);
                sp = 0;
                if (in_scene) enter(S.root_word);
                else retire();                   // bvh.c:446 / 522-526: miss without traversal
            }
            chunk_next += take;
            idle = __ballot_sync(FULL, cur == kIdle);
        }
        if (idle == FULL) break;

        // ------------------------------------------------------------------ traverse
        for (;;) {
            const bool in_leaf = (cur & kLeafFlag) != 0u;
            const uint32_t ntris = ((cur >> kLeafShift) & 15u) + 1u;
            const uint32_t nitems = L::items(ntris);
            const uint32_t cnt = in_leaf ? nitems - prog : 0u;
            const uint32_t total = __reduce_add_sync(FULL, cnt);
            const unsigned owners = __ballot_sync(FULL, in_leaf);
            const unsigned n_node = __popc(__ballot_sync(FULL, cur < kIdle));
            if (n_node == 0u && total == 0u) break;
            if (!exhausted && 32u - n_node - (uint32_t)__popc(owners) >= refill_at) break;

            if (total >= 32u || total > n_node) {
                // ---- leaf round (pool.cuh): items 0..31 of the pool, one per lane
                uint32_t excl = 0;
#pragma unroll
                for (int b = 0; b < L::kCntBits; ++b)
                    excl += (uint32_t)__popc(ballot_bit(cnt, 1u << b) & lt_mask) << b;
                const bool owner = in_leaf && excl < 32u;
                const unsigned starts = __reduce_or_sync(FULL, owner ? (1u << excl) : 0u);
                if (owner) {
                    s_desc[__popc(owners & lt_mask)] = make_uint2(cur, lane | ((prog - excl + 64u) << 8));
                    s_key[lane] = ~0ull;
                    if (sizeof(Real) == 8) s_win[lane] = 0u;
                }
                __syncwarp();
                unsigned own = 0;
                unsigned long long my_bits = ~0ull;
                if (lane < total) {
                    const uint2 d = s_desc[__popc(starts & le_mask) - 1u];
                    own = d.y & 31u;
                    const uint32_t item = lane + (d.y >> 8) - 64u;
                    Real oorg[3], odir[3], t, u, v;
                    uint32_t prim;
                    RaySlot<Real>::load(s_rays, own, oorg, odir);
                    pool_item_closest(K, trisT, d.x & kSlotMask, ((d.x >> kLeafShift) & 15u) + 1u, item, oorg, odir, t, u, v, prim);
                    if (prim != 0xffffffffu) {   // the item accepted a triangle: offer it to the owner
                        PoolRes<Real> r;
                        r.t = t; r.u = u; r.v = v; r.prim = prim;
                        s_res[lane] = r;
                        my_bits = abs_bits(t);
                        if (sizeof(Real) == 4) atomicMin(&s_key[own], (my_bits << 32) | (unsigned long long)(31u - lane));
                        else atomicMin(&s_key[own], my_bits);
                    }
                }
                __syncwarp();
                if (sizeof(Real) == 8) {         // fp64: among the items that hold the smallest t, the latest one
                    if (my_bits != ~0ull && s_key[own] == my_bits) atomicMax(&s_win[own], lane + 1u);
                    __syncwarp();
                }
                if (owner) {
                    const uint32_t took = (cnt < 32u - excl) ? cnt : 32u - excl;
                    const unsigned long long key = s_key[lane];
                    if (key != ~0ull) {          // winner of this round's items of my leaf; accepted unless t > t_leaf (bvh.c:780)
                        const unsigned w = (sizeof(Real) == 4) ? 31u - (unsigned)(key & 31ull) : s_win[lane] - 1u;
                        const PoolRes<Real> r = s_res[w];
                        if (!(r.t > tl)) { tl = r.t; ul = r.u; vl = r.v; tprim = r.prim; }
                    }
                    prog += took;
                    if (prog == nitems) {        // leaf finished: commit (bvh.c:850), then pop or retire
                        const bool commit = (tprim != 0xffffffffu) && (tl < best_t);
                        best_t = commit ? tl : best_t; best_u = commit ? ul : best_u; best_v = commit ? vl : best_v;
                        best_prim = commit ? tprim : best_prim;
                        if (sp == 0u) { retire(); cur = kIdle; }
                        else { --sp; enter(stk[sp * kPcThreads]); }
                    }
                }
                __syncwarp();                    // s_key / s_res are rewritten by the next round
            }
            if (cur < kIdle) {
                // ---- node step: bvh.c:1153-1179
                bool h0, h1;
                uint32_t c0, c1, axis;
                if (sizeof(Real) == 4) {
                    node_step_pk(K, S.nodes + cur, org, inv, sx, sy, sz, best_t, h0, h1, c0, c1, axis);
                } else {
                    NodeRegs<Real> nd;
                    load_node_wide(S.nodes + cur, nd);
                    h0 = slab_mm<Real>(nd.x[0], nd.x[1], nd.y[0], nd.y[1], nd.z[0], nd.z[1], org, inv, sx, sy, sz, best_t);
                    h1 = slab_mm<Real>(nd.x[2], nd.x[3], nd.y[2], nd.y[3], nd.z[2], nd.z[3], org, inv, sx, sy, sz, best_t);
                    c0 = nd.c0; c1 = nd.c1; axis = nd.axis;
                }
                const bool order = (axis == 0) ? sx : ((axis == 1) ? sy : sz);
                const bool both = h0 && h1, none = !h0 && !h1;
                const bool pop = none && (sp != 0u);
                if (both) stk[sp * kPcThreads] = order ? c0 : c1;
                const uint32_t popped = pop ? stk[(sp - 1u) * kPcThreads] : kIdle;
                sp = sp + (both ? 1u : 0u) - (pop ? 1u : 0u);
                const uint32_t one = h0 ? c0 : c1;
                const uint32_t next = both ? (order ? c1 : c0) : (none ? popped : one);
                if (next == kIdle) { retire(); cur = kIdle; }
                else enter(next);
            }
        }
    }
}

}  // namespace b200

// Persistent-threads wavefront traverser (the production ray-batch kernel).
//
// The grid is sized to the machine (resident CTAs per SM x 148 SMs) and every warp loops over three kinds of step:
//   fetch     idle lanes are refilled from the ray batch: the warp takes chunks of consecutive rays with ONE global
//             atomicAdd per chunk and hands them to idle lanes by ballot rank, so a lane whose ray terminates (any-hit
//             found, or the stack ran dry) is replaced at once instead of idling until the slowest ray of its warp ends
//   node step every lane that holds an inner-node word tests the two child boxes of ITS node (2 x LDG.256 per fp32 node)
//             and descends / pushes / pops -- written branch-free (selects + predicated stack accesses)
//   leaf step every lane that holds a leaf word tests the next PAIR of triangles of ITS leaf (3 x LDG.256 per fp32 pair),
//             branch-free as well; progress through the leaf lives in the leaf word itself (slot += 2, count -= 2)
// Which of node/leaf step runs next is decided by a warp vote (__ballot_sync + __popc): the kind more lanes are waiting
// for.  A lane's own sequence of box tests, triangle tests, pushes and pops -- and therefore its arithmetic, visiting
// order and tie rules -- is exactly that of trace_ray() (trace.cuh) / bvh_traverse (bvh.c:1092-1188); only the
// interleaving between lanes changes, so results are bit-identical to the one-ray-per-thread kernel.
// Occlusion results go out as one byte per ray, or -- for the AO transport -- are accumulated with atomicAdd into one
// counter per `rays_per_count` consecutive rays (the per-sample "occlusion += 1.0" of ambientocclusion.c:125-129).
#pragma once

namespace b200 {

constexpr uint32_t kIdle = kDoneWord;                 // lane holds no ray (bit 31 clear, never a valid node index)
constexpr uint32_t kSlotMask = (1u << kLeafShift) - 1u;

template <typename Real> struct LeafStep;

template <> struct LeafStep<float> {                 // two slots per step
    static constexpr uint32_t kPerStep = 2;
    static __device__ __forceinline__ void run(const Tri32 *tris, uint32_t slot, uint32_t left, const float org[3], const float dir[3],
                                               float &tl, float &ul, float &vl, uint32_t &tprim)
    {
        TriRegs<float> a, b;
        load_tri_pair(tris + slot, a, b);
        tri_test_bf<float>(a, org, dir, true, tl, ul, vl, tprim);
        tri_test_bf<float>(b, org, dir, left > 1u, tl, ul, vl, tprim);
    }
};
template <> struct LeafStep<double> {                // one slot per step
    static constexpr uint32_t kPerStep = 1;
    static __device__ __forceinline__ void run(const Tri64 *tris, uint32_t slot, uint32_t left, const double org[3], const double dir[3],
                                               double &tl, double &ul, double &vl, uint32_t &tprim)
    {
        TriRegs<double> a;
        load_tri_wide(tris + slot, a);
        tri_test_bf<double>(a, org, dir, true, tl, ul, vl, tprim);
    }
};

template <typename Real, bool ANYHIT>
__global__ void __launch_bounds__(kBlock)
trace_persistent_kernel(const SceneView<Real> S, const Real *__restrict__ rays, const uint32_t n, const uint32_t chunk,
                        typename RayIO<Real>::Hit *__restrict__ hits, uint8_t *__restrict__ occ,
                        uint32_t *__restrict__ counts, const uint32_t rays_per_count,
                        unsigned int *__restrict__ work_counter, const uint32_t refill_at)
{
    using P = Prec<Real>;
    extern __shared__ uint32_t s_stack[];
    uint32_t *stk = s_stack + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;

    uint32_t chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool exhausted = false;                      // warp-uniform

    // per-lane ray state.  cur: kIdle | inner-node index | leaf word (flag | (triangles left - 1) << 27 | next slot)
    uint32_t cur = kIdle, idx = 0, sp = 0, best_prim = 0xffffffffu, tprim = 0xffffffffu;
    Real org[3], dir[3], inv[3], best_t = P::inf(), best_u = Real(0), best_v = Real(0);
    Real tl = P::inf(), ul = Real(0), vl = Real(0);
    bool sx = false, sy = false, sz = false;
    org[0] = org[1] = org[2] = dir[0] = dir[1] = dir[2] = inv[0] = inv[1] = inv[2] = Real(0);

    auto retire = [&]() {                         // write the result of this lane's ray
        const bool hit = best_t < P::inf();       // bvh.c:1187
        if (ANYHIT && counts) { if (hit) atomicAdd(&counts[idx / rays_per_count], 1u); }
        else if (ANYHIT) occ[idx] = hit ? 1 : 0;
        else RayIO<Real>::store(hits, idx, hit, best_t, best_u, best_v, best_prim);
    };

    for (;;) {
        // ------------------------------------------------------------------ fetch
        unsigned idle = __ballot_sync(0xffffffffu, cur == kIdle);
        while (idle && !exhausted) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
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
                best_t = P::inf(); best_u = Real(0); best_v = Real(0); best_prim = 0xffffffffu;
                tl = P::inf(); ul = Real(0); vl = Real(0); tprim = 0xffffffffu;
                sx = dir[0] < Real(0); sy = dir[1] < Real(0); sz = dir[2] < Real(0);
#pragma unroll
                for (int k = 0; k < 3; ++k)      // bvh.c:473-497
                    inv[k] = (P::rabs(dir[k]) > P::eps()) ? Real(1) / dir[k] : ((dir[k] < Real(0)) ? -P::vmax() : P::vmax());
                Real tmin;
                const bool in_scene = (S.root_word != kDoneWord) &&
                    slab<Real>(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], org, inv, sx, sy, sz, tmin);
                sp = 0;
                if (in_scene) cur = S.root_word;
                else retire();                   // bvh.c:446 / 522-526: miss without traversal
            }
            chunk_next += take;
            idle = __ballot_sync(0xffffffffu, cur == kIdle);
        }
        if (idle == 0xffffffffu) break;          // nothing in flight and nothing left to fetch

        // ------------------------------------------------------------------ traverse: vote, step, repeat
        for (;;) {
            const bool in_leaf = (cur & kLeafFlag) != 0u;
            const bool in_node = !in_leaf && (cur != kIdle);
            const unsigned want_node = __ballot_sync(0xffffffffu, in_node);
            const unsigned want_leaf = __ballot_sync(0xffffffffu, in_leaf);
            if ((want_node | want_leaf) == 0u) break;
            // refill once `refill_at` lanes idle and rays remain (rays fetched together start in lock step: their first
            // node and leaf loads coalesce)
            if (!exhausted && (uint32_t)__popc(~(want_node | want_leaf)) >= refill_at) break;

            if (__popc(want_node) >= __popc(want_leaf)) {
                if (in_node) {                   // ---- node step: bvh.c:1153-1179
                    NodeRegs<Real> nd;
                    load_node_wide(S.nodes + cur, nd);
                    const bool h0 = slab_mm<Real>(nd.x[0], nd.x[1], nd.y[0], nd.y[1], nd.z[0], nd.z[1], org, inv, sx, sy, sz, best_t);
                    const bool h1 = slab_mm<Real>(nd.x[2], nd.x[3], nd.y[2], nd.y[3], nd.z[2], nd.z[3], org, inv, sx, sy, sz, best_t);
                    const bool order = (nd.axis == 0) ? sx : ((nd.axis == 1) ? sy : sz);     // near child = child[sign[axis0]]
                    const bool both = h0 && h1, none = !h0 && !h1;
                    const bool pop = none && (sp != 0u);
                    if (both) stk[sp * kBlock] = order ? nd.c0 : nd.c1;
                    const uint32_t popped = pop ? stk[(sp - 1u) * kBlock] : kIdle;
                    sp = sp + (both ? 1u : 0u) - (pop ? 1u : 0u);
                    const uint32_t one = h0 ? nd.c0 : nd.c1;
                    const uint32_t next = both ? (order ? nd.c1 : nd.c0) : (none ? popped : one);
                    if (next == kIdle) retire();                                  // stack ran dry
                    if (next & kLeafFlag) { tl = P::inf(); ul = Real(0); vl = Real(0); tprim = 0xffffffffu; }   // bvh.c:833-836
                    cur = next;
                }
            } else {
                if (in_leaf) {                   // ---- leaf step: bvh.c:838-861
                    const uint32_t left = ((cur >> kLeafShift) & 15u) + 1u;
                    LeafStep<Real>::run(S.tris, cur & kSlotMask, left, org, dir, tl, ul, vl, tprim);
                    // occlusion query: the first accepted triangle already decides the answer -- the rest of the leaf
                    // could only lower the leaf-local t, and the commit test of bvh.c:850 is tl < 1e38 here
                    const bool decided = ANYHIT && (tl < P::inf());
                    if (left > LeafStep<Real>::kPerStep && !decided) {
                        cur += LeafStep<Real>::kPerStep - (LeafStep<Real>::kPerStep << kLeafShift);
                    } else {                     // leaf finished: commit (bvh.c:850), then pop or retire
                        const bool commit = (tprim != 0xffffffffu) && (tl < best_t);
                        best_t = commit ? tl : best_t; best_u = commit ? ul : best_u; best_v = commit ? vl : best_v;
                        best_prim = commit ? tprim : best_prim;
                        const bool done = (ANYHIT && commit) || (sp == 0u);
                        const uint32_t popped = done ? kIdle : stk[(sp - 1u) * kBlock];
                        sp -= done ? 0u : 1u;
                        if (done) retire();
                        tl = P::inf(); ul = Real(0); vl = Real(0); tprim = 0xffffffffu;
                        cur = popped;
                    }
                }
            }
        }
    }
}

}  // namespace b200

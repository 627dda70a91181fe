// Persistent-threads wavefront traverser (the production ray-batch kernel).
//
// The grid is sized to the machine (resident CTAs per SM x 148 SMs) and every warp loops:
//   fetch    idle lanes are refilled from the ray batch: the warp takes chunks of kChunk consecutive rays
//            with ONE global atomicAdd per chunk and hands them to idle lanes by ballot rank, so a lane
//            whose ray terminates (any-hit found, or the stack ran dry) is replaced at once instead of
//            idling until the slowest ray of its warp finishes
//   traverse synchronised while-while: all lanes descend inner nodes until each holds a leaf word,
//            then all lanes run their leaf's Moeller-Trumbore loop, pop, and the warp re-ballots
// Per-ray arithmetic, visiting order and tie rules are exactly those of trace_ray() (trace.cuh); only
// the assignment of rays to lanes changes, so results are bit-identical to the one-ray-per-thread kernel.
#pragma once

namespace b200 {

constexpr unsigned kChunk = 128;        // rays per atomic fetch (one warp)

template <typename Real, bool ANYHIT>
__global__ void __launch_bounds__(kBlock)
trace_persistent_kernel(const SceneView<Real> S, const Real *__restrict__ rays, const uint32_t n,
                        typename RayIO<Real>::Hit *__restrict__ hits, uint8_t *__restrict__ occ,
                        unsigned int *__restrict__ work_counter)
{
    using P = Prec<Real>;
    extern __shared__ uint32_t s_stack[];
    uint32_t *stk = s_stack + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;

    uint32_t chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool exhausted = (S.root_word == kDoneWord) && false;

    // per-lane ray state
    bool busy = false;
    uint32_t idx = 0, cur = kDoneWord, sp = 0, best_prim = 0xffffffffu;
    Real org[3], dir[3], inv[3], best_t = P::inf(), best_u = Real(0), best_v = Real(0);
    bool sx = false, sy = false, sz = false;
    org[0] = org[1] = org[2] = dir[0] = dir[1] = dir[2] = inv[0] = inv[1] = inv[2] = Real(0);

    for (;;) {
        // ------------------------------------------------------------------ fetch
        unsigned idle = __ballot_sync(0xffffffffu, !busy);
        while (idle && !exhausted) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = (n - base < kChunk) ? n : base + kChunk;
            }
            const unsigned avail = chunk_end - chunk_next;
            const unsigned n_idle = __popc(idle);
            const unsigned take = n_idle < avail ? n_idle : avail;
            const unsigned rank = __popc(idle & lt_mask);
            if (!busy && rank < take) {
                idx = chunk_next + rank;
                RayIO<Real>::load(rays, idx, org, dir);
                best_t = P::inf(); best_u = Real(0); best_v = Real(0); best_prim = 0xffffffffu;
                sx = dir[0] < Real(0); sy = dir[1] < Real(0); sz = dir[2] < Real(0);
#pragma unroll
                for (int k = 0; k < 3; ++k)      // bvh.c:473-497
                    inv[k] = (P::rabs(dir[k]) > P::eps()) ? Real(1) / dir[k] : ((dir[k] < Real(0)) ? -P::vmax() : P::vmax());
                Real tmin;
                const bool in_scene = (S.root_word != kDoneWord) &&
                    slab<Real>(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], org, inv, sx, sy, sz, tmin);
                if (in_scene) {
                    busy = true; cur = S.root_word; sp = 0;
                } else {                           // bvh.c:446 / 522-526: miss without traversal
                    if (ANYHIT) occ[idx] = 0;
                    else RayIO<Real>::store(hits, idx, false, best_t, best_u, best_v, best_prim);
                }
            }
            chunk_next += take;
            idle = __ballot_sync(0xffffffffu, !busy);
        }
        if (idle == 0xffffffffu) break;          // nothing in flight and nothing left to fetch

        // ------------------------------------------------------------------ traverse
        for (;;) {
            // inner phase: every busy lane descends until it holds a leaf word (or finishes)
            while (busy && !(cur & kLeafFlag)) {
                NodeRegs<Real> nd;
                load_node(S.nodes + cur, nd);
                Real tmin0, tmin1;
                const bool h0 = slab<Real>(nd.x[0], nd.x[1], nd.y[0], nd.y[1], nd.z[0], nd.z[1], org, inv, sx, sy, sz, tmin0) && (tmin0 < best_t);
                const bool h1 = slab<Real>(nd.x[2], nd.x[3], nd.y[2], nd.y[3], nd.z[2], nd.z[3], org, inv, sx, sy, sz, tmin1) && (tmin1 < best_t);
                if (h0 && h1) {
                    const bool order = (nd.axis == 0) ? sx : ((nd.axis == 1) ? sy : sz);
                    stk[sp * kBlock] = order ? nd.c0 : nd.c1;
                    ++sp;
                    cur = order ? nd.c1 : nd.c0;
                } else if (h0) {
                    cur = nd.c0;
                } else if (h1) {
                    cur = nd.c1;
                } else if (sp == 0) {
                    busy = false;                 // stack ran dry: this ray is finished
                } else {
                    --sp;
                    cur = stk[sp * kBlock];
                }
            }
            // leaf phase
            if (busy) {
                const uint32_t start = cur & ((1u << kLeafShift) - 1u);
                const uint32_t count = ((cur >> kLeafShift) & 15u) + 1u;
                Real tl = P::inf(), ul = Real(0), vl = Real(0);
                uint32_t tid = 0;
                bool any = false;
                for (uint32_t i = 0; i < count; ++i) {
                    TriRegs<Real> tr;
                    load_tri(S.tris + start + i, tr);
                    if (tri_test<Real>(tr, org, dir, tl, ul, vl)) { tid = i; any = true; }
                }
                if (any && (tl < best_t)) {       // bvh.c:850
                    best_t = tl; best_u = ul; best_v = vl; best_prim = start + tid;
                    if (ANYHIT) busy = false;
                }
                if (busy) {
                    if (sp == 0) busy = false;
                    else { --sp; cur = stk[sp * kBlock]; }
                }
            }
            // retire finished rays
            const bool finished = !busy && (cur != kDoneWord);
            if (finished) {
                const bool hit = best_t < P::inf();
                if (ANYHIT) occ[idx] = hit ? 1 : 0;
                else RayIO<Real>::store(hits, idx, hit, best_t, best_u, best_v, best_prim);
                cur = kDoneWord;
            }
            const unsigned live = __ballot_sync(0xffffffffu, busy);
            if (live == 0u) break;
            if (!exhausted && live != 0xffffffffu) break;     // some lane idles and rays remain: refill
        }
    }
}

}  // namespace b200

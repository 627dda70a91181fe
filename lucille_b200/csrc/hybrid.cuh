// fp64-EXACT occlusion at (nearly) fp32 cost: the pooled fp32 traverser of pool32.cuh used as a FILTER whose every decision carries
// a rigorous error bound, with the reference's double arithmetic run on the spot for the few decisions the bound cannot settle.
//
// Why: the double records are what real scenes need (RIB-scale coordinates put lucille's 1e-6 origin offset below an fp32 ulp, so
// plain fp32 records self-occlude at random: SURVEY 7 hard part 1), and the double kernels run at half the fp32 rate because
// every record is twice as many bytes, every operation half rate, and occupancy is lower.  But almost every decision of a
// traversal -- does the ray's slab interval of this box exist, is this (u, v, t) inside the triangle's window -- is far from its
// threshold; only those within a few fp32 ulps of it need doubles.
//
// What is computed.  The query is ri_b200_occluded_*_f64's: does ANY leaf the reference reaches hold a triangle triangle_isect accepts
// with t < 1e38 (pool.cuh has the argument that this is order-independent).  Rays come in double.  Per ray: inverse direction and the
// scene-box test in double exactly as bvh.c:473-526; then origin = hi + lo floats, direction and inverse rounded to float.
//   * child box (bvh.c:869-936): t-values from the OUTWARD-rounded fp32 boxes.  With u = 2^-24, |t32 - t64| <= E for every plane, where
//         E = 8 u max_k |inv_k| (Bmax_k + |org_k|)          (Bmax = the scene box's largest |coordinate|: every node box is inside)
//     [derivation: box rounding 2u|b|, origin rounding u|O|, subtraction u|n|, inverse u, product u  =>  u |inv| (5|b| + 4|O|)].
//     max / min are 1-Lipschitz, so tmin64 in [tmin32 - E, tmin32 + E] and the same for tmax.  The reference passes the child iff
//     tmax > 0 && tmin <= tmax && tmin < 1e38:   certainly so iff  tmax32 - E > 0 && tmax32 - tmin32 >= 2E && tmin32 < 1e37;
//     certainly not iff  tmax32 + E <= 0 || tmin32 - tmax32 > 2E;  otherwise the double record of that child is tested (hyb_box64).
//   * triangle (bvh.c:730-791), division-free: a = e1.(d x e2), U = s.(d x e2), V = (s x e1).d, T = e2.(s x e1) in fp32 with
//     s = (org_hi - v0) + org_lo.  With the max-norms Md, Me1, Me2, Ms of d, e1, e2, s and eta = the absolute error of s:
//         |a32 - a| <= 48 u Md Me1 Me2          |U32 - U| <= Md Me2 (6 eta + 42 u Ms)     (V: Md Me1 ..., T: Me1 Me2 ...)
//     (three-term dot products of two-term cross products: input roundings, product and sum roundings, all bounded through the
//     max-norms).  The kernel uses 96 u and G = 8 eta0 + 72 u Ms, a third more than derived, which also covers the reference's own
//     2^-53 roundings and the arithmetic that evaluates the bounds.  eta0 = the part of s's error that is not relative to s: 0 when
//     every vertex coordinate of the scene is an fp32 number (RIB files hold floats; so does the synthetic soup), else u Bmax -- and then
//     (or when the scene lies far from the world origin) the filter reads its OWN fp32 records, relative to the scene centre and
//     rounded once from the double records (end of this file), so that Bmax is the scene's half extent wherever the scene sits.
//     With sg = sign(a): u >= 0 <=> sg U >= 0, u + v <= 1 <=> sg (U + V) <= |a|, t >= 0 <=> sg T >= 0, |a| > 1e-14.  A triangle
//     is certainly accepted / certainly rejected when every / some comparison holds with its bound to spare; otherwise its double
//     slot is tested with the reference's expression tree (hyb_tri64).  NaNs and infinities make every comparison false = undecided.
// Every decision the reference takes for the ray is therefore either reproduced with certainty or recomputed in double, so the set
// of leaves reached and the OR of the acceptances are the reference's: the result equals ri_b200_occluded_dev_f64's bit for bit
// (tests/test_gpu_parity.py::test_hybrid_occlusion_is_fp64_exact, test_gpu_fullsize.py).  The bounds themselves are checked as
// mathematics on the CPU (tests/test_hybrid_bounds.py: this classification restated in numpy float32 against the double expression
// trees on millions of adversarial cases; with the constants cut to 1/8 the same inputs find contradictions, with 1/4 they do not).
#pragma once

namespace b200 {

struct HybK {
    float  bmax[3];         // largest |coordinate| of the (translated) scene box per axis
    float  eta0;            // absolute error of a vertex coordinate in the fp32 slots: 0 (all vertices are fp32 numbers) or u * max_k bmax_k
    float  de;              // absolute error of an fp32 edge component beyond its relative rounding: 0 (edges rounded from doubles) or 2 * eta0
    double c[3];            // origin of the fp32 records the kernel reads (end of this file): the filter works on coordinates minus c
    int    order;           // occlusion kernel: child order when both children are kept (pool32.cuh): 1 longer path first, 2 nearer first, 0 reference
};

struct HybWarp {
    float4 r32[96];         // per lane three float4: (org_hi.xyz, Md) (dir.xyz, G0) (org_lo.xyz, -)
    double r64[32 * 9];     // per lane: org.xyz dir.xyz inv.xyz in double, for the on-the-spot double tests
    uint2  desc[32];        // leaf-round descriptors
};
#ifndef B200_HYB_THREADS
#define B200_HYB_THREADS 256
#endif
#ifndef B200_HYB_CTAS
#define B200_HYB_CTAS 4
#endif
constexpr int kHybThreads = B200_HYB_THREADS;
template <int kCap> struct HybSmem {
    uint32_t stack[kCap * kHybThreads];
    HybWarp  warp[kHybThreads / 32];
};

__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }

// the reference's child-box test in double on the double node record (bvh.c:869-936, 1038-1044 with best_t = 1e38)
__device__ __noinline__ bool hyb_box64(const double *__restrict__ bx, const uint32_t r64_a, const uint32_t sgn)
{
    const double lox = __ldg(bx), hix = __ldg(bx + 1), loy = __ldg(bx + 4), hiy = __ldg(bx + 5), loz = __ldg(bx + 8), hiz = __ldg(bx + 9);
    const double org[3] = {lds_f64(r64_a), lds_f64(r64_a + 8u), lds_f64(r64_a + 16u)};
    const double inv[3] = {lds_f64(r64_a + 48u), lds_f64(r64_a + 56u), lds_f64(r64_a + 64u)};
    double tmin;
    const bool pass = slab<double>(lox, hix, loy, hiy, loz, hiz, org, inv, (sgn & 1u) != 0u, (sgn & 2u) != 0u, (sgn & 4u) != 0u, tmin);
    return pass && (tmin < 1.0e38);
}

// the reference's triangle test in double on the double slot (bvh.c:730-791 from t_leaf = 1e38; commits iff t < 1e38, bvh.c:850)
__device__ __noinline__ bool hyb_tri64(const Tri64 *__restrict__ tp, const uint32_t r64_a)
{
    TriRegs<double> tr;
    load_tri(tp, tr);
    const double org[3] = {lds_f64(r64_a), lds_f64(r64_a + 8u), lds_f64(r64_a + 16u)};
    const double dir[3] = {lds_f64(r64_a + 24u), lds_f64(r64_a + 32u), lds_f64(r64_a + 40u)};
    double t = 1.0e38, u = 0.0, v = 0.0;
    const bool ok = tri_test<double>(tr, org, dir, t, u, v);
    return ok && (t < 1.0e38);
}

// 1 = the reference certainly accepts, 0 = certainly rejects, 2 = undecided at fp32
template <bool kExact>
__device__ __forceinline__ uint32_t hyb_tri_class(const float a, const float U, const float V, const float T, const float Me1, const float Me2,
                                                  const float Ms, const float Md, const float G0, const float de)
{
    constexpr float u72 = 72.0f * 5.9604645e-8f, u96 = 96.0f * 5.9604645e-8f;
    const float G = __fmaf_rn(u72, Ms, G0);
    const float P12 = Me1 * Me2, MdE1 = Md * Me1, MdE2 = Md * Me2;
    float ea = u96 * (Md * P12), eU = MdE2 * G, eV = MdE1 * G, eT = P12 * G;
    if (!kExact) {                                   // edges carry an absolute error de on top of the relative one
        const float Ms1 = Ms + G, Es = (Me1 + Me2) + de, d8 = 8.0f * de;
        ea = __fmaf_rn(d8 * Md, Es, ea);
        eU = __fmaf_rn(d8 * Md, Ms1, eU);
        eV = __fmaf_rn(d8 * Md, Ms1, eV);
        eT = __fmaf_rn(d8 * Ms1, Es, eT);
    }
    const uint32_t sbit = __float_as_uint(a) & 0x80000000u;
    const float A = fabsf(a);
    const float Us = __uint_as_float(__float_as_uint(U) ^ sbit), Vs = __uint_as_float(__float_as_uint(V) ^ sbit),
                Ts = __uint_as_float(__float_as_uint(T) ^ sbit);
    const float Alo = A - ea, Ahi = A + ea;
    const float Ulo = Us - eU, Uhi = Us + eU, Vlo = Vs - eV, Vhi = Vs + eV, Tlo = Ts - eT, Thi = Ts + eT;
    const bool acc = (Alo > 1.0001e-14f) && (Ulo >= 0.0f) && (Vlo >= 0.0f) && (Tlo >= 0.0f) && ((Alo - Uhi) - Vhi >= 0.0f) && (Thi <= 1.0e30f * Alo);
    const bool rej = (Ahi <= 0.9999e-14f) || (Uhi < 0.0f) || (Vhi < 0.0f) || (Thi < 0.0f) || (Ulo > Ahi) || (Ulo + Vlo > Ahi);
    return acc ? 1u : (rej ? 0u : 2u);
}

// both triangles of a leaf item: bit 0 = some triangle is certainly accepted; bits 1, 2 = triangle A / B is undecided
template <bool kExact>
__device__ __forceinline__ uint32_t hyb_pair(const PackK &K, const P4 &c0, const P4 &c1, const P2 &c2, const float4 rA, const float4 rB, const float4 rC,
                                             const bool valid_b, const float de)
{
    const pk_t v0x = c0.v[0], v0y = c0.v[1], v0z = c0.v[2], e1x = c0.v[3];
    const pk_t e1y = c1.v[0], e1z = c1.v[1], e2x = c1.v[2], e2y = c1.v[3];
    const pk_t e2z = c2.v[0];
    const pk_t dx = pkb(rB.x), dy = pkb(rB.y), dz = pkb(rB.z);
    const pk_t px = psub(K, pmul(K, dy, e2z), pmul(K, dz, e2y));
    const pk_t py = psub(K, pmul(K, dz, e2x), pmul(K, dx, e2z));
    const pk_t pz = psub(K, pmul(K, dx, e2y), pmul(K, dy, e2x));
    const pk_t a = padd(K, padd(K, pmul(K, e1x, px), pmul(K, e1y, py)), pmul(K, e1z, pz));
    const pk_t sx = padd(K, psub(K, pkb(rA.x), v0x), pkb(rC.x)), sy = padd(K, psub(K, pkb(rA.y), v0y), pkb(rC.y)),
               sz = padd(K, psub(K, pkb(rA.z), v0z), pkb(rC.z));
    const pk_t qx = psub(K, pmul(K, sy, e1z), pmul(K, sz, e1y));
    const pk_t qy = psub(K, pmul(K, sz, e1x), pmul(K, sx, e1z));
    const pk_t qz = psub(K, pmul(K, sx, e1y), pmul(K, sy, e1x));
    const pk_t U = padd(K, padd(K, pmul(K, sx, px), pmul(K, sy, py)), pmul(K, sz, pz));
    const pk_t V = padd(K, padd(K, pmul(K, qx, dx), pmul(K, qy, dy)), pmul(K, qz, dz));
    const pk_t T = padd(K, padd(K, pmul(K, e2x, qx), pmul(K, e2y, qy)), pmul(K, e2z, qz));
    float aA, aB, UA, UB, VA, VB, TA, TB, x0, x1, y0, y1, z0, z1;
    upk2(a, aA, aB); upk2(U, UA, UB); upk2(V, VA, VB); upk2(T, TA, TB);
    upk2(e1x, x0, x1); upk2(e1y, y0, y1); upk2(e1z, z0, z1);
    const float Me1A = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fabsf(z0)), Me1B = fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fabsf(z1));
    upk2(e2x, x0, x1); upk2(e2y, y0, y1); upk2(e2z, z0, z1);
    const float Me2A = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fabsf(z0)), Me2B = fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fabsf(z1));
    upk2(sx, x0, x1); upk2(sy, y0, y1); upk2(sz, z0, z1);
    const float MsA = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fabsf(z0)), MsB = fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fabsf(z1));
    const uint32_t kA = hyb_tri_class<kExact>(aA, UA, VA, TA, Me1A, Me2A, MsA, rA.w, rB.w, de);
    const uint32_t kB = valid_b ? hyb_tri_class<kExact>(aB, UB, VB, TB, Me1B, Me2B, MsB, rA.w, rB.w, de) : 0u;
    return ((kA | kB) & 1u) | (kA & 2u) | ((kB & 2u) << 1);
}

template <int kCap, bool kCounts, bool kExact>
__global__ void __launch_bounds__(kHybThreads, B200_HYB_CTAS)
occluded_hybrid_kernel(const SceneView<float> S, const SceneView<double> S64, const char *__restrict__ trisT, const double *__restrict__ rays,
                       const uint32_t n, const uint32_t chunk, uint8_t *__restrict__ occ, uint32_t *__restrict__ counts,
                       const uint32_t rays_per_count, unsigned int *__restrict__ work_counter, const unsigned int *__restrict__ ready,
                       unsigned int *__restrict__ fault, const PackK K, const HybK H)
{
    constexpr unsigned FULL = 0xffffffffu;
#ifndef B200_HYB_REFILL
#define B200_HYB_REFILL 8
#endif
    constexpr uint32_t kRefillAt = B200_HYB_REFILL, kLeafAt = 32u;
    constexpr uint32_t kRow = kHybThreads * 4u;
    constexpr float kU = 5.9604645e-8f;                                // 2^-24
    extern __shared__ __align__(16) unsigned char hyb_smem[];          // one HybSmem<kCap> (more than the 48 KB a static array may have)
    const unsigned lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = (2u << lane) - 1u;
    uint32_t sm_a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sm_a) : "l"(hyb_smem));
    const uint32_t stack_a = sm_a + (uint32_t)offsetof(HybSmem<kCap>, stack);
    const uint32_t r32_a = sm_a + (uint32_t)offsetof(HybSmem<kCap>, warp) + (wbase >> 5) * (uint32_t)sizeof(HybWarp);
    const uint32_t r64_a = r32_a + (uint32_t)offsetof(HybWarp, r64);
    const uint32_t desc_a = r32_a + (uint32_t)offsetof(HybWarp, desc);

    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;

    uint32_t cur = kIdle, prog = 0, idx = 0, spa = threadIdx.x * 4u, sgn = 0;
    float org[3] = {0.0f, 0.0f, 0.0f}, inv[3] = {0.0f, 0.0f, 0.0f}, E = 0.0f;

    auto retire = [&](const bool hit) {
        if (kCounts) { if (hit) atomicAdd(&counts[idx / rays_per_count], 1u); }
        else occ[idx] = hit ? 1 : 0;
    };

    for (;;) {
        // ------------------------------------------------------------------ fetch (pool32.cuh), ray set-up in double
        unsigned idle = __ballot_sync(FULL, cur == kIdle);
        while (idle && !exhausted) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = (n - base < chunk) ? n : base + chunk;
                if (ready) {
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
                double O[3], D[3], I[3];
                if (ready) RayIO<double>::load_coherent(rays, idx, O, D);
                else RayIO<double>::load(rays, idx, O, D);
                const bool sx = D[0] < 0.0, sy = D[1] < 0.0, sz = D[2] < 0.0;
                sgn = (sx ? 1u : 0u) | (sy ? 2u : 0u) | (sz ? 4u : 0u);
#pragma unroll
                for (int k = 0; k < 3; ++k)      // bvh.c:473-497
                    I[k] = (fabs(D[k]) > 1.0e-14) ? 1.0 / D[k] : ((D[k] < 0.0) ? -DBL_MAX : DBL_MAX);
                double tmin64;
                const bool in_scene = (S.root_word != kDoneWord) &&        // bvh.c:522-526, in double: exact
                    slab<double>(S64.smin[0], S64.smax[0], S64.smin[1], S64.smax[1], S64.smin[2], S64.smax[2], O, I, sx, sy, sz, tmin64);
                const uint32_t q = r64_a + lane * 72u;
#pragma unroll
                for (int k = 0; k < 3; ++k) { sts_f64(q + 8u * k, O[k]); sts_f64(q + 24u + 8u * k, D[k]); sts_f64(q + 48u + 8u * k, I[k]); }
                float ol[3], d[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double oc = O[k] - H.c[k];             // the filter's frame (one rounding of 2^-53 |O|: inside G0's 4 u^2 term)
                    org[k] = (float)oc;
                    ol[k] = (float)(oc - (double)org[k]);
                    d[k] = (float)D[k];
                    inv[k] = (float)I[k];
                }
                const float Md = fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fabsf(d[2]));
                const float Mo = fmaxf(fmaxf(fabsf(org[0]), fabsf(org[1])), fabsf(org[2])) +
                                 (float)fmax(fmax(fabs(H.c[0]), fabs(H.c[1])), fabs(H.c[2]));
                E = (8.0f * kU) * fmaxf(fmaxf(fabsf(inv[0]) * (H.bmax[0] + fabsf(org[0])), fabsf(inv[1]) * (H.bmax[1] + fabsf(org[1]))),
                                        fabsf(inv[2]) * (H.bmax[2] + fabsf(org[2])));
                if (!(E < 1.0e30f)) E = __int_as_float(0x7f800000);       // a component of 1/dir beyond fp32 (or NaN): nothing is certain
                const float G0 = 8.0f * (H.eta0 + (4.0f * kU * kU) * Mo) + 1.0e-30f;
                sts128(r32_a + lane * 48u, make_float4(org[0], org[1], org[2], Md));
                sts128(r32_a + lane * 48u + 16u, make_float4(d[0], d[1], d[2], G0));
                sts128(r32_a + lane * 48u + 32u, make_float4(ol[0], ol[1], ol[2], 0.0f));
                spa = threadIdx.x * 4u; prog = 0;
                if (in_scene) cur = S.root_word;
                else retire(false);
            }
            chunk_next += take;
            idle = __ballot_sync(FULL, cur == kIdle);
        }
        if (idle == FULL) break;

        // ------------------------------------------------------------------ traverse
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

            if (total >= kLeafAt || total > n_node) {
                // ---- leaf round (pool32.cuh); an item's verdict is certain at fp32 or settled in double on the spot
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
                    const uint32_t item = lane + (d.y >> 8) - 64u;
                    const uint32_t ntris = ((d.x >> kLeafShift) & 15u) + 1u, slot0 = d.x & kSlotMask;
                    const uint32_t m = ((ntris + 3u) >> 2) << 1;
                    const uint32_t o0 = slot0 * 3u + item * 2u, o1 = o0 + 2u * m, o2 = slot0 * 3u + 4u * m + item;
                    const P4 c0 = ldg256p(trisT + (size_t)o0 * 16u), c1 = ldg256p(trisT + (size_t)o1 * 16u);
                    const P2 c2 = ldg128p(trisT + (size_t)o2 * 16u);
                    const float4 rA = lds128(r32_a + own * 48u), rB = lds128(r32_a + own * 48u + 16u), rC = lds128(r32_a + own * 48u + 32u);
                    const uint32_t k = hyb_pair<kExact>(K, c0, c1, c2, rA, rB, rC, 2u * item + 1u < ntris, H.de);
                    hit = (k & 1u) != 0u;
                    if (!hit && (k & 6u)) {                                      // undecided triangles: the reference's test on the double slot
                        const Tri64 *tp = S64.tris + slot0 + 2u * item;
                        const uint32_t q = r64_a + own * 72u;
                        hit = hyb_tri64((k & 2u) ? tp : tp + 1, q);                 // ONE call site serves either triangle: the warp runs the
                        if (!hit && (k & 6u) == 6u) hit = hyb_tri64(tp + 1, q);     // double routine once per round, twice only when a lane needs both
                    }
                }
                const unsigned hits = __ballot_sync(FULL, hit);
                if (owner) {
                    const uint32_t room = 32u - excl, took = cnt < room ? cnt : room;
                    const unsigned mine = (FULL >> (32u - took)) << excl;
                    if (hits & mine) { retire(true); cur = kIdle; }
                    else {
                        prog += took;
                        if (prog == nitems) {
                            prog = 0;
                            if (spa < kRow) { retire(false); cur = kIdle; }
                            else { spa -= kRow; cur = lds32(stack_a + spa); }
                        }
                    }
                }
            }
            if (cur < kIdle) {
                // ---- node step: bvh.c:1153-1179 with best_t == 1e38
                const Node32 *p = S.nodes + cur;
                const P4 a = ldg256p(p), b = ldg256p(reinterpret_cast<const char *>(p) + 32);
                const pk_t ox = pkb(org[0]), oy = pkb(org[1]), oz = pkb(org[2]);
                const pk_t ix = pkb(inv[0]), iy = pkb(inv[1]), iz = pkb(inv[2]);
                const bool sx = (sgn & 1u) != 0u, sy = (sgn & 2u) != 0u, sz = (sgn & 4u) != 0u;
                float tn0, tf0, tn1, tf1;
                slab_pk_t(K, a.v[0], a.v[2], b.v[0], ox, oy, oz, ix, iy, iz, sx, sy, sz, tn0, tf0);
                slab_pk_t(K, a.v[1], a.v[3], b.v[1], ox, oy, oz, ix, iy, iz, sx, sy, sz, tn1, tf1);
                const float E2 = E + E;
                bool h0 = (tf0 - E > 0.0f) && (tf0 - tn0 >= E2) && (tn0 < 1.0e37f);
                bool h1 = (tf1 - E > 0.0f) && (tf1 - tn1 >= E2) && (tn1 < 1.0e37f);
                const bool u0 = !h0 && !((tf0 + E <= 0.0f) || (tn0 - tf0 > E2));
                const bool u1 = !h1 && !((tf1 + E <= 0.0f) || (tn1 - tf1 > E2));
                if (u0 || u1) {                                                  // undecided: the double record of this node
                    const double *bx = S64.nodes[cur].x;
                    const uint32_t q = r64_a + lane * 72u;
                    const bool r = hyb_box64(u0 ? bx : bx + 2, q, sgn);                 // one call site for either child
                    if (u0) h0 = r; else h1 = r;
                    if (u0 && u1) h1 = hyb_box64(bx + 2, q, sgn);
                }
                const uint32_t c0 = (uint32_t)b.v[2], c1 = (uint32_t)(b.v[2] >> 32), axis = (uint32_t)b.v[3];
                const bool order = H.order == 1 ? (tf1 - fmaxf(tn1, 0.0f)) > (tf0 - fmaxf(tn0, 0.0f))      // any order gives the same answer
                                 : H.order == 2 ? tn1 < tn0 : ((sgn >> axis) & 1u) != 0u;
                const uint32_t near = order ? c1 : c0, far = order ? c0 : c1;
                const bool both = h0 && h1, none = !h0 && !h1;
                const bool pop = none && (spa >= kRow);
                if (both) sts32(stack_a + spa, far);
                const uint32_t popped = pop ? lds32(stack_a + spa - kRow) : kIdle;
                spa = spa + (both ? kRow : 0u) - (pop ? kRow : 0u);
                const uint32_t next = both ? near : (none ? popped : (h0 ? c0 : c1));
                if (next == kIdle) retire(false);
                prog = 0;
                cur = next;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------------------------
// Double-exact CLOSEST hit through the fp32 records: the pooled closest-hit scheme of pool_closest.cuh (a lane walks its ray's nodes in
// the reference's order and waits in a leaf until the leaf is committed) with the same certified fp32 decisions as the occlusion kernel
// above, and the reference's double arithmetic for everything that can reach the answer:
//   * best_t, the leaf-local t and the (t, u, v) of every offer are DOUBLES computed by the reference's own triangle_isect on the
//     double slot (hyb_tri64_hit); fp32 only decides which triangles need not be looked at: those whose window test certainly fails,
//     and those whose t is certainly greater than the best t so far (T_lo > best_hi * |a|_hi) -- a leaf whose smallest t is not below
//     best_t does not commit (bvh.c:850), so such a triangle can never change the result, nor can it change which of the OTHER
//     triangles of its leaf wins (it could only raise nothing: the leaf-local t only ever decreases);
//   * a child box is kept iff the reference keeps it, `pass && tmin < best_t` (bvh.c:1038-1044): certainly so iff the fp32 interval
//     passes with its bound to spare and tmin32 + E < round_down(best_t), certainly not iff it certainly fails or tmin32 - E >=
//     round_up(best_t); otherwise the double record decides with the exact best_t (hyb_box64_best).
// Visiting order = tree order + the ray's sign bits, which are taken from the double direction: the same as the reference's; the tie
// rules (later triangle of a leaf wins equal t, earlier leaf wins across leaves) act on the exact doubles in that order.  Result:
// (hit, t, u, v, prim) bit-identical to ri_b200_intersect_*_f64's double kernel and to the reference.
__device__ __noinline__ bool hyb_box64_best(const double *__restrict__ bx, const uint32_t r64_a, const uint32_t sgn, const double best_t)
{
    const double lox = __ldg(bx), hix = __ldg(bx + 1), loy = __ldg(bx + 4), hiy = __ldg(bx + 5), loz = __ldg(bx + 8), hiz = __ldg(bx + 9);
    const double org[3] = {lds_f64(r64_a), lds_f64(r64_a + 8u), lds_f64(r64_a + 16u)};
    const double inv[3] = {lds_f64(r64_a + 48u), lds_f64(r64_a + 56u), lds_f64(r64_a + 64u)};
    double tmin;
    const bool pass = slab<double>(lox, hix, loy, hiy, loz, hiz, org, inv, (sgn & 1u) != 0u, (sgn & 2u) != 0u, (sgn & 4u) != 0u, tmin);
    return pass && (tmin < best_t);
}

// triangle_isect on the double slot against t_io (the pair-local t so far): updates (t, u, v, prim) on acceptance
__device__ __noinline__ void hyb_tri64_hit(const Tri64 *__restrict__ tp, const uint32_t r64_a, double &t_io, double &u_io, double &v_io, uint32_t &prim_io)
{
    TriRegs<double> tr;
    load_tri(tp, tr);
    const double org[3] = {lds_f64(r64_a), lds_f64(r64_a + 8u), lds_f64(r64_a + 16u)};
    const double dir[3] = {lds_f64(r64_a + 24u), lds_f64(r64_a + 32u), lds_f64(r64_a + 40u)};
    if (tri_test<double>(tr, org, dir, t_io, u_io, v_io)) prim_io = tr.prim;
}

// could the reference accept this triangle with a t that is not certainly above best_hi?  (false = it cannot matter)
template <bool kExact>
__device__ __forceinline__ bool hyb_tri_maybe(const float a, const float U, const float V, const float T, const float Me1, const float Me2,
                                              const float Ms, const float Md, const float G0, const float de, const float best_hi)
{
    constexpr float u72 = 72.0f * 5.9604645e-8f, u96 = 96.0f * 5.9604645e-8f;
    const float G = __fmaf_rn(u72, Ms, G0);
    const float P12 = Me1 * Me2, MdE1 = Md * Me1, MdE2 = Md * Me2;
    float ea = u96 * (Md * P12), eU = MdE2 * G, eV = MdE1 * G, eT = P12 * G;
    if (!kExact) {
        const float Ms1 = Ms + G, Es = (Me1 + Me2) + de, d8 = 8.0f * de;
        ea = __fmaf_rn(d8 * Md, Es, ea);
        eU = __fmaf_rn(d8 * Md, Ms1, eU);
        eV = __fmaf_rn(d8 * Md, Ms1, eV);
        eT = __fmaf_rn(d8 * Ms1, Es, eT);
    }
    const uint32_t sbit = __float_as_uint(a) & 0x80000000u;
    const float A = fabsf(a);
    const float Us = __uint_as_float(__float_as_uint(U) ^ sbit), Vs = __uint_as_float(__float_as_uint(V) ^ sbit),
                Ts = __uint_as_float(__float_as_uint(T) ^ sbit);
    const float Ahi = A + ea;
    const float Ulo = Us - eU, Uhi = Us + eU, Vlo = Vs - eV, Vhi = Vs + eV, Tlo = Ts - eT, Thi = Ts + eT;
    // certainly rejected by the window (as in hyb_tri_class), or certainly beyond the best t so far: t = T / |a| >= Tlo / Ahi > best_hi
    const bool rej = (Ahi <= 0.9999e-14f) || (Uhi < 0.0f) || (Vhi < 0.0f) || (Thi < 0.0f) || (Ulo > Ahi) || (Ulo + Vlo > Ahi) ||
                     (Tlo > best_hi * Ahi * 1.000001f);
    return !rej;
}

struct HybCWarp {
    float4 r32[96];                    // per lane: (org_hi.xyz, Md) (dir.xyz, G0) (org_lo.xyz, best_hi)
    double r64[32 * 9];                // per lane: org, dir, inv in double
    uint2  desc[32];
    unsigned long long key[32];        // per owner: bits of the smallest t offered this round
    PoolRes<double> res[32];           // per item lane: its offer
    double leaf_uv[64], best_uv[64];   // per lane: (u, v) of the leaf-local / committed record; their t live in registers
    uint32_t win[32], leaf_prim[32], best_prim[32];
};
#ifndef B200_HC_THREADS
#define B200_HC_THREADS 128
#endif
constexpr int kHcThreads = B200_HC_THREADS;
template <int kCap> struct HybCSmem {
    uint32_t stack[kCap * kHcThreads];
    HybCWarp warp[kHcThreads / 32];
};
#ifndef B200_HC_CTAS
#define B200_HC_CTAS 5
#endif

template <int kCap, bool kExact>
__global__ void __launch_bounds__(kHcThreads, B200_HC_CTAS)
closest_hybrid_kernel(const SceneView<float> S, const SceneView<double> S64, const char *__restrict__ trisT, const double *__restrict__ rays,
                      const uint32_t n, const uint32_t chunk, ri_b200_hit_f64 *__restrict__ hits_out, unsigned int *__restrict__ work_counter,
                      const PackK K, const HybK H)
{
    constexpr unsigned FULL = 0xffffffffu;
#ifndef B200_HC_REFILL
#define B200_HC_REFILL 8
#endif
    constexpr uint32_t kRefillAt = B200_HC_REFILL;
    constexpr uint32_t kRow = kHcThreads * 4u;
    constexpr float kU = 5.9604645e-8f;
    extern __shared__ __align__(16) unsigned char hyb_smem[];
    const unsigned lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = (2u << lane) - 1u;
    HybCWarp *W = reinterpret_cast<HybCSmem<kCap> *>(hyb_smem)->warp + (wbase >> 5);
    uint32_t sm_a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sm_a) : "l"(hyb_smem));
    const uint32_t stack_a = sm_a + (uint32_t)offsetof(HybCSmem<kCap>, stack);
    const uint32_t r32_a = sm_a + (uint32_t)offsetof(HybCSmem<kCap>, warp) + (wbase >> 5) * (uint32_t)sizeof(HybCWarp);
    const uint32_t r64_a = r32_a + (uint32_t)offsetof(HybCWarp, r64);
    const uint32_t desc_a = r32_a + (uint32_t)offsetof(HybCWarp, desc);

    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;

    uint32_t cur = kIdle, prog = 0, idx = 0, spa = threadIdx.x * 4u, sgn = 0;
    float org[3] = {0.0f, 0.0f, 0.0f}, inv[3] = {0.0f, 0.0f, 0.0f}, E = 0.0f, best_lo = 1.0e38f, best_hi = 1.0e38f;
    double best_t = 1.0e38, tl = 1.0e38;

    auto set_best = [&](const double t) {            // best_t and its fp32 brackets; the upper one is what item lanes read
        best_t = t;
        best_lo = __double2float_rd(t); best_hi = __double2float_ru(t);
        asm volatile("st.shared.f32 [%0], %1;" :: "r"(r32_a + lane * 48u + 44u), "f"(best_hi) : "memory");
    };
    auto retire = [&]() {                            // bvh.c:1187
        const bool hit = best_t < 1.0e38;
        RayIO<double>::store(hits_out, idx, hit, best_t, W->best_uv[2 * lane], W->best_uv[2 * lane + 1], W->best_prim[lane]);
    };
    auto enter = [&](const uint32_t word) {          // a leaf starts with a fresh leaf-local record (bvh.c:833-836)
        cur = word; prog = 0;
        tl = 1.0e38;
        W->leaf_prim[lane] = 0xffffffffu;
    };

    for (;;) {
        // ------------------------------------------------------------------ fetch, ray set-up in double (as in occluded_hybrid_kernel)
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
                double O[3], D[3], I[3];
                RayIO<double>::load(rays, idx, O, D);
                const bool sx = D[0] < 0.0, sy = D[1] < 0.0, sz = D[2] < 0.0;
                sgn = (sx ? 1u : 0u) | (sy ? 2u : 0u) | (sz ? 4u : 0u);
#pragma unroll
                for (int k = 0; k < 3; ++k)          // bvh.c:473-497
                    I[k] = (fabs(D[k]) > 1.0e-14) ? 1.0 / D[k] : ((D[k] < 0.0) ? -DBL_MAX : DBL_MAX);
                double tmin64;
                const bool in_scene = (S.root_word != kDoneWord) &&
                    slab<double>(S64.smin[0], S64.smax[0], S64.smin[1], S64.smax[1], S64.smin[2], S64.smax[2], O, I, sx, sy, sz, tmin64);
                const uint32_t q = r64_a + lane * 72u;
#pragma unroll
                for (int k = 0; k < 3; ++k) { sts_f64(q + 8u * k, O[k]); sts_f64(q + 24u + 8u * k, D[k]); sts_f64(q + 48u + 8u * k, I[k]); }
                float ol[3], d[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double oc = O[k] - H.c[k];             // the filter's frame (one rounding of 2^-53 |O|: inside G0's 4 u^2 term)
                    org[k] = (float)oc;
                    ol[k] = (float)(oc - (double)org[k]);
                    d[k] = (float)D[k];
                    inv[k] = (float)I[k];
                }
                const float Md = fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fabsf(d[2]));
                const float Mo = fmaxf(fmaxf(fabsf(org[0]), fabsf(org[1])), fabsf(org[2])) +
                                 (float)fmax(fmax(fabs(H.c[0]), fabs(H.c[1])), fabs(H.c[2]));
                E = (8.0f * kU) * fmaxf(fmaxf(fabsf(inv[0]) * (H.bmax[0] + fabsf(org[0])), fabsf(inv[1]) * (H.bmax[1] + fabsf(org[1]))),
                                        fabsf(inv[2]) * (H.bmax[2] + fabsf(org[2])));
                if (!(E < 1.0e30f)) E = __int_as_float(0x7f800000);
                const float G0 = 8.0f * (H.eta0 + (4.0f * kU * kU) * Mo) + 1.0e-30f;
                sts128(r32_a + lane * 48u, make_float4(org[0], org[1], org[2], Md));
                sts128(r32_a + lane * 48u + 16u, make_float4(d[0], d[1], d[2], G0));
                sts128(r32_a + lane * 48u + 32u, make_float4(ol[0], ol[1], ol[2], 0.0f));
                set_best(1.0e38);
                W->best_uv[2 * lane] = 0.0; W->best_uv[2 * lane + 1] = 0.0; W->best_prim[lane] = 0xffffffffu;
                spa = threadIdx.x * 4u;
                if (in_scene) enter(S.root_word);
                else retire();
            }
            chunk_next += take;
            idle = __ballot_sync(FULL, cur == kIdle);
        }
        if (idle == FULL) break;

        // ------------------------------------------------------------------ traverse
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
                    W->key[lane] = ~0ull;
                    W->win[lane] = 0u;
                }
                __syncwarp();
                unsigned own = 0;
                unsigned long long my_bits = ~0ull;
                if (lane < total) {
                    const uint2 d = lds64(desc_a + (uint32_t)__popc(starts & le_mask) * 8u - 8u);
                    own = d.y & 31u;
                    const uint32_t item = lane + (d.y >> 8) - 64u;
                    const uint32_t ntris = ((d.x >> kLeafShift) & 15u) + 1u, slot0 = d.x & kSlotMask;
                    const uint32_t m = ((ntris + 3u) >> 2) << 1;
                    const uint32_t o0 = slot0 * 3u + item * 2u, o1 = o0 + 2u * m, o2 = slot0 * 3u + 4u * m + item;
                    const P4 c0 = ldg256p(trisT + (size_t)o0 * 16u), c1 = ldg256p(trisT + (size_t)o1 * 16u);
                    const P2 c2 = ldg128p(trisT + (size_t)o2 * 16u);
                    const float4 rA = lds128(r32_a + own * 48u), rB = lds128(r32_a + own * 48u + 16u), rC = lds128(r32_a + own * 48u + 32u);
                    // the fp32 expression trees of hyb_pair, then: which of the two triangles can still matter?
                    const pk_t v0x = c0.v[0], v0y = c0.v[1], v0z = c0.v[2], e1x = c0.v[3];
                    const pk_t e1y = c1.v[0], e1z = c1.v[1], e2x = c1.v[2], e2y = c1.v[3];
                    const pk_t e2z = c2.v[0];
                    const pk_t dx = pkb(rB.x), dy = pkb(rB.y), dz = pkb(rB.z);
                    const pk_t px = psub(K, pmul(K, dy, e2z), pmul(K, dz, e2y));
                    const pk_t py = psub(K, pmul(K, dz, e2x), pmul(K, dx, e2z));
                    const pk_t pz = psub(K, pmul(K, dx, e2y), pmul(K, dy, e2x));
                    const pk_t a = padd(K, padd(K, pmul(K, e1x, px), pmul(K, e1y, py)), pmul(K, e1z, pz));
                    const pk_t sx = padd(K, psub(K, pkb(rA.x), v0x), pkb(rC.x)), sy = padd(K, psub(K, pkb(rA.y), v0y), pkb(rC.y)),
                               sz = padd(K, psub(K, pkb(rA.z), v0z), pkb(rC.z));
                    const pk_t qx = psub(K, pmul(K, sy, e1z), pmul(K, sz, e1y));
                    const pk_t qy = psub(K, pmul(K, sz, e1x), pmul(K, sx, e1z));
                    const pk_t qz = psub(K, pmul(K, sx, e1y), pmul(K, sy, e1x));
                    const pk_t U = padd(K, padd(K, pmul(K, sx, px), pmul(K, sy, py)), pmul(K, sz, pz));
                    const pk_t V = padd(K, padd(K, pmul(K, qx, dx), pmul(K, qy, dy)), pmul(K, qz, dz));
                    const pk_t T = padd(K, padd(K, pmul(K, e2x, qx), pmul(K, e2y, qy)), pmul(K, e2z, qz));
                    float aA, aB, UA, UB, VA, VB, TA, TB, x0, x1, y0, y1, z0, z1;
                    upk2(a, aA, aB); upk2(U, UA, UB); upk2(V, VA, VB); upk2(T, TA, TB);
                    upk2(e1x, x0, x1); upk2(e1y, y0, y1); upk2(e1z, z0, z1);
                    const float Me1A = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fabsf(z0)), Me1B = fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fabsf(z1));
                    upk2(e2x, x0, x1); upk2(e2y, y0, y1); upk2(e2z, z0, z1);
                    const float Me2A = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fabsf(z0)), Me2B = fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fabsf(z1));
                    upk2(sx, x0, x1); upk2(sy, y0, y1); upk2(sz, z0, z1);
                    const float MsA = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fabsf(z0)), MsB = fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fabsf(z1));
                    const bool mA = hyb_tri_maybe<kExact>(aA, UA, VA, TA, Me1A, Me2A, MsA, rA.w, rB.w, H.de, rC.w);
                    const bool mB = (2u * item + 1u < ntris) && hyb_tri_maybe<kExact>(aB, UB, VB, TB, Me1B, Me2B, MsB, rA.w, rB.w, H.de, rC.w);
                    if (mA || mB) {                  // the reference's test on the double slots, in leaf order from t = 1e38 (bvh.c:833-848)
                        double t = 1.0e38, u = 0.0, v = 0.0;
                        uint32_t prim = 0xffffffffu;
                        const Tri64 *tp = S64.tris + slot0 + 2u * item;
                        const uint32_t q = r64_a + own * 72u;
                        hyb_tri64_hit(mA ? tp : tp + 1, q, t, u, v, prim);           // one call site for either triangle (see the occlusion kernel)
                        if (mA && mB) hyb_tri64_hit(tp + 1, q, t, u, v, prim);
                        if (prim != 0xffffffffu) {
                            PoolRes<double> r;
                            r.t = t; r.u = u; r.v = v; r.prim = prim; r.pad = 0u;
                            W->res[lane] = r;
                            my_bits = abs_bits(t);
                            atomicMin(&W->key[own], my_bits);
                        }
                    }
                }
                __syncwarp();
                if (my_bits != ~0ull && W->key[own] == my_bits) atomicMax(&W->win[own], lane + 1u);   // among equal t, the latest item
                __syncwarp();
                if (owner) {
                    const uint32_t room = 32u - excl, took = cnt < room ? cnt : room;
                    if (W->key[lane] != ~0ull) {     // this round's winner among my leaf's items; accepted unless t > t_leaf (bvh.c:780)
                        const PoolRes<double> r = W->res[W->win[lane] - 1u];
                        if (!(r.t > tl)) { tl = r.t; W->leaf_uv[2 * lane] = r.u; W->leaf_uv[2 * lane + 1] = r.v; W->leaf_prim[lane] = r.prim; }
                    }
                    prog += took;
                    if (prog == nitems) {            // leaf finished: commit (bvh.c:850), then pop or retire
                        if ((W->leaf_prim[lane] != 0xffffffffu) && (tl < best_t)) {
                            set_best(tl);
                            W->best_uv[2 * lane] = W->leaf_uv[2 * lane]; W->best_uv[2 * lane + 1] = W->leaf_uv[2 * lane + 1];
                            W->best_prim[lane] = W->leaf_prim[lane];
                        }
                        if (spa < kRow) { retire(); cur = kIdle; }
                        else { spa -= kRow; enter(lds32(stack_a + spa)); }
                    }
                }
                __syncwarp();                        // key / res / win are rewritten by the next round
            }
            if (cur < kIdle) {
                // ---- node step: bvh.c:1153-1179
                const Node32 *p = S.nodes + cur;
                const P4 a = ldg256p(p), b = ldg256p(reinterpret_cast<const char *>(p) + 32);
                const pk_t ox = pkb(org[0]), oy = pkb(org[1]), oz = pkb(org[2]);
                const pk_t ix = pkb(inv[0]), iy = pkb(inv[1]), iz = pkb(inv[2]);
                const bool sx = (sgn & 1u) != 0u, sy = (sgn & 2u) != 0u, sz = (sgn & 4u) != 0u;
                float tn0, tf0, tn1, tf1;
                slab_pk_t(K, a.v[0], a.v[2], b.v[0], ox, oy, oz, ix, iy, iz, sx, sy, sz, tn0, tf0);
                slab_pk_t(K, a.v[1], a.v[3], b.v[1], ox, oy, oz, ix, iy, iz, sx, sy, sz, tn1, tf1);
                const float E2 = E + E;
                bool h0 = (tf0 - E > 0.0f) && (tf0 - tn0 >= E2) && (tn0 + E < best_lo);
                bool h1 = (tf1 - E > 0.0f) && (tf1 - tn1 >= E2) && (tn1 + E < best_lo);
                const bool u0 = !h0 && !((tf0 + E <= 0.0f) || (tn0 - tf0 > E2) || (tn0 - E >= best_hi));
                const bool u1 = !h1 && !((tf1 + E <= 0.0f) || (tn1 - tf1 > E2) || (tn1 - E >= best_hi));
                if (u0 || u1) {
                    const double *bx = S64.nodes[cur].x;
                    const uint32_t q = r64_a + lane * 72u;
                    const bool r = hyb_box64_best(u0 ? bx : bx + 2, q, sgn, best_t);    // one call site for either child
                    if (u0) h0 = r; else h1 = r;
                    if (u0 && u1) h1 = hyb_box64_best(bx + 2, q, sgn, best_t);
                }
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

// ------------------------------------------------------------------------------------------------------------------------------------
// The filter's own fp32 records, derived on the device from the DOUBLE records (whichever builder made them) when the shared fp32
// records would give poor bounds: coordinates relative to an origin c (the scene box's centre) so that the bounds scale with the
// scene's size instead of its distance from the world origin, boxes rounded outward with directed roundings end to end, vertices
// fl32(v0 - c), edges fl32 of the double edges (relative error only: no absolute term `de`).
__global__ void __launch_bounds__(256)
hyb_nodes_kernel(const Node64 *__restrict__ n64, const uint32_t ninner, const double cx, const double cy, const double cz, Node32 *__restrict__ out)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= ninner) return;
    const Node64 s = n64[i];
    Node32 d;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j & 1) {
            d.x[j] = __double2float_ru(__dsub_ru(s.x[j], cx)); d.y[j] = __double2float_ru(__dsub_ru(s.y[j], cy)); d.z[j] = __double2float_ru(__dsub_ru(s.z[j], cz));
        } else {
            d.x[j] = __double2float_rd(__dsub_rd(s.x[j], cx)); d.y[j] = __double2float_rd(__dsub_rd(s.y[j], cy)); d.z[j] = __double2float_rd(__dsub_rd(s.z[j], cz));
        }
    }
    d.c0 = s.c0; d.c1 = s.c1; d.axis = s.axis; d.pad = 0u;
    out[i] = d;
}

// one thread per child word (2 per inner node; `root_leaf` != 0: the single leaf of a tree without inner nodes): the leaf's pairs in
// the interleaved, leaf-transposed layout of FlatTree::tris32t (bvh_build.cpp)
__global__ void __launch_bounds__(256)
hyb_tris_kernel(const Node64 *__restrict__ n64, const uint32_t nwords, const uint32_t root_leaf, const Tri64 *__restrict__ t64,
                const double cx, const double cy, const double cz, char *__restrict__ out)
{
    const uint32_t w = blockIdx.x * 256u + threadIdx.x;
    if (w >= nwords) return;
    const uint32_t word = root_leaf ? root_leaf : ((w & 1u) ? n64[w >> 1].c1 : n64[w >> 1].c0);
    if (!(word & kLeafFlag)) return;
    const uint32_t ntris = ((word >> kLeafShift) & 15u) + 1u, slot0 = word & kSlotMask;
    const uint32_t m = ((ntris + 3u) >> 2) << 1, used = (ntris + 1u) >> 1;
    const double c[3] = {cx, cy, cz};
    char *dst = out + (size_t)slot0 * 48u;
    for (uint32_t j = 0; j < m; ++j) {
        uint32_t wd[20];
#pragma unroll
        for (int q = 0; q < 20; ++q) wd[q] = 0u;
        wd[18] = wd[19] = 0xffffffffu;
        for (uint32_t h = 0; h < 2u; ++h) {
            const uint32_t k = 2u * j + h;
            if (k < ntris) {
                TriRegs<double> tr;
                load_tri(t64 + slot0 + k, tr);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    wd[2 * q + h] = __float_as_uint((float)(tr.v0[q] - c[q]));
                    wd[2 * (3 + q) + h] = __float_as_uint((float)tr.e1[q]);
                    wd[2 * (6 + q) + h] = __float_as_uint((float)tr.e2[q]);
                }
                wd[18 + h] = tr.prim;
            } else if (j + 1u == used) {                     // filler half of the last pair: unit edges (masked by its validity bit)
                wd[2 * 3 + h] = __float_as_uint(1.0f);
                wd[2 * 7 + h] = __float_as_uint(1.0f);
            }
        }
        uint32_t *p0 = reinterpret_cast<uint32_t *>(dst + (size_t)j * 32u), *p1 = reinterpret_cast<uint32_t *>(dst + (size_t)(m + j) * 32u),
                 *p2 = reinterpret_cast<uint32_t *>(dst + (size_t)2u * m * 32u + (size_t)j * 16u);
#pragma unroll
        for (int q = 0; q < 8; ++q) { p0[q] = wd[q]; p1[q] = wd[8 + q]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) p2[q] = wd[16 + q];
    }
}

}  // namespace b200

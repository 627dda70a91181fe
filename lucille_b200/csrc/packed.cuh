// Packed fp32 arithmetic for the pooled traversers: two IEEE-754 single operations per issued instruction (sm_100 FFMA2).
//
// The fp32 kernels are limited by warp-instruction issue, and a third of what they issue are the separately rounded multiplies
// and adds of the reference's expression trees (trace.cuh; -fmad=false is part of the parity contract).  Blackwell executes
// `fma.rn.f32x2` on an aligned register pair as ONE instruction, so the two triangles of a leaf item (A, B) -- and the
// (lo, hi) planes of a child box -- are carried as pairs and every multiply / add / subtract of the reference becomes one FFMA2:
//
//      a * b   ==  fma(a, b, -0.0)        exact product, one rounding; -0.0 keeps the sign of a zero product
//      a + b   ==  fma(a, 1.0, b)         a * 1 is exact, one rounding of the sum
//      a - b   ==  fma(b, -1.0, a)        b * -1 is exact, one rounding of the difference
//
// so each result is bit-identical to the scalar FMUL / FADD it replaces (IEEE-754 fusedMultiplyAdd rounds once; the products
// by +-1 and the sums with -0.0 are exact, signs of zero included: (+0) + (-0) = +0 and (-0) + (-0) = -0 in round-to-nearest).
//
// Why not mul.rn.f32x2 / add.rn.f32x2: ptxas 12.9 contracts a packed multiply feeding a packed add into FFMA2 even under
// --fmad=false and in spite of the .rn qualifiers (checked in SASS), which would round once where the reference rounds twice.
// A chain of fma's cannot be contracted further; the constants come in as a kernel parameter (uniform registers in SASS: no
// per-thread register cost), so ptxas cannot see that they are +-1 / -0 and fold them back into packed adds and multiplies.
#pragma once

#include <cstdint>

namespace b200 {

typedef unsigned long long pk_t;                     // (lo, hi) = two floats in one aligned 64-bit register pair

struct PackK { pk_t one, mone, nzero; };             // (1,1), (-1,-1), (-0,-0): filled by make_pack_k() on the host

static inline PackK make_pack_k()
{
    PackK k;
    k.one = 0x3f8000003f800000ull; k.mone = 0xbf800000bf800000ull; k.nzero = 0x8000000080000000ull;
    return k;
}

#ifdef __CUDACC__
__device__ __forceinline__ pk_t pk2(float lo, float hi)
{
    pk_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
// broadcast pair (x, x).  volatile: keeps the mov next to its use, where ptxas folds it into the FFMA2 operand (Rx.F32);
// hoisted out of the traversal loop it would pin a second register per ray component for the whole kernel.
__device__ __forceinline__ pk_t pkb(float x)
{
    pk_t r;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void upk2(pk_t x, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x)); }
__device__ __forceinline__ pk_t pfma(pk_t a, pk_t b, pk_t c)
{
    pk_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ pk_t pmul(const PackK &K, pk_t a, pk_t b) { return pfma(a, b, K.nzero); }
__device__ __forceinline__ pk_t padd(const PackK &K, pk_t a, pk_t b) { return pfma(a, K.one, b); }
__device__ __forceinline__ pk_t psub(const PackK &K, pk_t a, pk_t b) { return pfma(b, K.mone, a); }

struct __align__(32) P4 { pk_t v[4]; };
__device__ __forceinline__ P4 ldg256p(const void *p)          // LDG.E.256 into four aligned register pairs
{
    P4 r;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
}
// the same load with an L1 eviction hint: node records are re-used by every ray (keep), triangle chunks stream through (do not allocate)
__device__ __forceinline__ P4 ldg256p_keep(const void *p)
{
#ifdef B200_L1_HINTS
    P4 r;
    asm volatile("ld.global.nc.L1::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
#else
    return ldg256p(p);
#endif
}
__device__ __forceinline__ P4 ldg256p_stream(const void *p)
{
#ifdef B200_L1_HINTS
    P4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
#else
    return ldg256p(p);
#endif
}

struct __align__(16) P2 { pk_t v[2]; };
__device__ __forceinline__ P2 ldg128p(const void *p)          // LDG.E.128 into two aligned register pairs
{
    P2 r;
    asm volatile("ld.global.nc.v2.b64 {%0,%1}, [%2];" : "=l"(r.v[0]), "=l"(r.v[1]) : "l"(p));
    return r;
}

// triangle_isect (bvh.c:730-791) on BOTH triangles of a leaf item at once.  The item's chunks hold the pair interleaved
// (bvh_build.cpp): c0 = v0.x v0.y v0.z e1.x, c1 = e1.y e1.z e2.x e2.y (32 B each), c2 = e2.z prim (16 B), each word pair = (A, B).
// Returns det, u, v, t and u+v as (A, B) pairs -- the same expression trees as tri_test_bf (trace.cuh), every operation rounded
// separately.
struct PairMT { pk_t a, u, v, t, uv; };
__device__ __forceinline__ PairMT pair_mt(const PackK &K, const P4 &c0, const P4 &c1, const P2 &c2, const float org[3], const float dir[3])
{
    const pk_t v0x = c0.v[0], v0y = c0.v[1], v0z = c0.v[2], e1x = c0.v[3];
    const pk_t e1y = c1.v[0], e1z = c1.v[1], e2x = c1.v[2], e2y = c1.v[3];
    const pk_t e2z = c2.v[0];
    const pk_t dx = pkb(dir[0]), dy = pkb(dir[1]), dz = pkb(dir[2]);
    const pk_t ox = pkb(org[0]), oy = pkb(org[1]), oz = pkb(org[2]);
    const pk_t px = psub(K, pmul(K, dy, e2z), pmul(K, dz, e2y));
    const pk_t py = psub(K, pmul(K, dz, e2x), pmul(K, dx, e2z));
    const pk_t pz = psub(K, pmul(K, dx, e2y), pmul(K, dy, e2x));
    PairMT r;
    r.a = padd(K, padd(K, pmul(K, e1x, px), pmul(K, e1y, py)), pmul(K, e1z, pz));
    float a0, a1;
    upk2(r.a, a0, a1);
    const pk_t inva = pk2(1.0f / a0, 1.0f / a1);
    const pk_t sx = psub(K, ox, v0x), sy = psub(K, oy, v0y), sz = psub(K, oz, v0z);
    const pk_t qx = psub(K, pmul(K, sy, e1z), pmul(K, sz, e1y));
    const pk_t qy = psub(K, pmul(K, sz, e1x), pmul(K, sx, e1z));
    const pk_t qz = psub(K, pmul(K, sx, e1y), pmul(K, sy, e1x));
    r.u = pmul(K, padd(K, padd(K, pmul(K, sx, px), pmul(K, sy, py)), pmul(K, sz, pz)), inva);
    r.v = pmul(K, padd(K, padd(K, pmul(K, qx, dx), pmul(K, qy, dy)), pmul(K, qz, dz)), inva);
    r.t = pmul(K, padd(K, padd(K, pmul(K, e2x, qx), pmul(K, e2y, qy)), pmul(K, e2z, qz)), inva);
    r.uv = padd(K, r.u, r.v);
    return r;
}

// acceptance window of triangle_isect against the leaf-local t (NEGATED comparisons as in tri_test_bf: NaNs fall the same way)
__device__ __forceinline__ bool mt_accept(float a, float u, float v, float uv, float t, float t_leaf)
{
    return (fabsf(a) > 1.0e-14f) && !(u < 0.0f) && !(u > 1.0f) && !(v < 0.0f) && !(uv > 1.0f) && !(t < 0.0f) && !(t > t_leaf);
}

// leaf-local winner of the pair, tested in leaf order from t_leaf = 1e38 (bvh.c:833-848 restricted to the pair)
__device__ __forceinline__ void pair_closest(const PackK &K, const char *p0, const char *p1, const char *p2, const bool valid_b,
                                             const float org[3], const float dir[3], float &tl, float &ul, float &vl, uint32_t &tprim)
{
    const P4 c0 = ldg256p_stream(p0), c1 = ldg256p_stream(p1);
    const P2 c2 = ldg128p(p2);
    const PairMT r = pair_mt(K, c0, c1, c2, org, dir);
    float aA, aB, uA, uB, vA, vB, tA, tB, wA, wB, primA, primB;
    upk2(r.a, aA, aB); upk2(r.u, uA, uB); upk2(r.v, vA, vB); upk2(r.t, tA, tB); upk2(r.uv, wA, wB); upk2(c2.v[1], primA, primB);
    tl = 1.0e38f; ul = 0.0f; vl = 0.0f; tprim = 0xffffffffu;
    const bool okA = mt_accept(aA, uA, vA, wA, tA, tl);
    tl = okA ? tA : tl; ul = okA ? uA : ul; vl = okA ? vA : vl; tprim = okA ? __float_as_uint(primA) : tprim;
    const bool okB = valid_b && mt_accept(aB, uB, vB, wB, tB, tl);
    tl = okB ? tB : tl; ul = okB ? uB : ul; vl = okB ? vB : vl; tprim = okB ? __float_as_uint(primB) : tprim;
}

// occlusion form: does the pair hold a triangle the reference accepts with t < 1e38?
__device__ __forceinline__ bool pair_occluded(const PackK &K, const char *p0, const char *p1, const char *p2, const bool valid_b,
                                              const float org[3], const float dir[3])
{
    const P4 c0 = ldg256p_stream(p0), c1 = ldg256p_stream(p1);
    const P2 c2 = ldg128p(p2);
    const PairMT r = pair_mt(K, c0, c1, c2, org, dir);
    float aA, aB, uA, uB, vA, vB, tA, tB, wA, wB;
    upk2(r.a, aA, aB); upk2(r.u, uA, uB); upk2(r.v, vA, vB); upk2(r.t, tA, tB); upk2(r.uv, wA, wB);
    float tl = 1.0e38f;
    const bool okA = mt_accept(aA, uA, vA, wA, tA, tl);
    tl = okA ? tA : tl;
    const bool okB = valid_b && mt_accept(aB, uB, vB, wB, tB, tl);
    tl = okB ? tB : tl;
    return tl < 1.0e38f;
}

// Two child boxes of a node record (bvh.c:869-936, 1030-1044).  The record keeps (lo, hi) of a child adjacent per axis, so
// (plane - org) * inv runs on the pair; near / far are then picked by the ray's sign bit exactly as the reference does.
__device__ __forceinline__ bool slab_pk(const PackK &K, pk_t bx, pk_t by, pk_t bz, pk_t ox, pk_t oy, pk_t oz, pk_t ix, pk_t iy, pk_t iz,
                                        bool sx, bool sy, bool sz, float best_t)
{
    float lx, hx, ly, hy, lz, hz;
    upk2(pmul(K, psub(K, bx, ox), ix), lx, hx);
    upk2(pmul(K, psub(K, by, oy), iy), ly, hy);
    upk2(pmul(K, psub(K, bz, oz), iz), lz, hz);
    const float tnx = sx ? hx : lx, tfx = sx ? lx : hx;
    const float tny = sy ? hy : ly, tfy = sy ? ly : hy;
    const float tnz = sz ? hz : lz, tfz = sz ? lz : hz;
    const float tmin = fmaxf(fmaxf(tnx, tny), tnz);
    const float tmax = fminf(fminf(tfx, tfy), tfz);
    return (tmax > 0.0f) && (tmin <= tmax) && (tmin < best_t);
}
// packed slab arithmetic of slab_pk (packed.cuh), returning the interval instead of the verdict
__device__ __forceinline__ void slab_pk_t(const PackK &K, pk_t bx, pk_t by, pk_t bz, pk_t ox, pk_t oy, pk_t oz, pk_t ix, pk_t iy, pk_t iz,
                                          bool sx, bool sy, bool sz, float &tmin, float &tmax)
{
    float lx, hx, ly, hy, lz, hz;
    upk2(pmul(K, psub(K, bx, ox), ix), lx, hx);
    upk2(pmul(K, psub(K, by, oy), iy), ly, hy);
    upk2(pmul(K, psub(K, bz, oz), iz), lz, hz);
    const float tnx = sx ? hx : lx, tfx = sx ? lx : hx;
    const float tny = sy ? hy : ly, tfy = sy ? ly : hy;
    const float tnz = sz ? hz : lz, tfz = sz ? lz : hz;
    tmin = fmaxf(fmaxf(tnx, tny), tnz);
    tmax = fminf(fminf(tfx, tfy), tfz);
}

#endif

}  // namespace b200

// Device-side ray traversal for the B200 accelerator (sm_100a).
//
// One ray per lane, ordered depth-first traversal of the reference-identical binary BVH with a
// per-lane short stack kept in shared memory ([depth][lane] layout: conflict-free), 128-bit loads
// of node records and triangle slots (this reference kernel: LDG.128; the persistent kernel: LDG.256), leaf
// Moeller-Trumbore with the reference's acceptance window and tie rules.
//
// Semantics follow lucille's src/render/bvh.c exactly (430-542 ray setup + scene-box rejection,
// 1092-1188 traversal order, 938-1083/869-936 slab tests, 793-864/730-791 leaf test); the file
// must be compiled with -fmad=false so that every multiply and add rounds separately, like the
// reference (double) and the fp32 restatement in oracle/.
#pragma once

#include <cfloat>
#include <cstdint>

#include "bvh_build.h"

namespace b200 {

template <typename Real> struct Prec;

template <> struct Prec<float> {
    using Node = Node32;
    using Tri  = Tri32;
    static __device__ __forceinline__ float inf()  { return 1.0e38f; }
    static __device__ __forceinline__ float eps()  { return 1.0e-14f; }
    static __device__ __forceinline__ float vmax() { return FLT_MAX; }
    static __device__ __forceinline__ float rabs(float x) { return fabsf(x); }
};
template <> struct Prec<double> {
    using Node = Node64;
    using Tri  = Tri64;
    static __device__ __forceinline__ double inf()  { return 1.0e38; }
    static __device__ __forceinline__ double eps()  { return 1.0e-14; }
    static __device__ __forceinline__ double vmax() { return DBL_MAX; }
    static __device__ __forceinline__ double rabs(double x) { return fabs(x); }
};

template <typename Real> struct SceneView {
    const typename Prec<Real>::Node *nodes;
    const typename Prec<Real>::Tri  *tris;
    const uint32_t *slot_of_prim;      // hit id (post-build triangle position) -> triangle slot
    const Real     *normals;           // optional [prim][9] vertex normals n0 n1 n2 (NULL: Ns = Ng)
    Real     smin[3], smax[3];
    uint32_t root_word;
    uint32_t top_count;
};

struct LaneCounters { uint32_t ninner, nleaf, ntris, nhit; };

// node record -> registers -------------------------------------------------------------------
template <typename Real> struct NodeRegs { Real x[4], y[4], z[4]; uint32_t c0, c1, axis; };

__device__ __forceinline__ void load_node(const Node32 *p, NodeRegs<float> &r)
{
    const float4 *q = reinterpret_cast<const float4 *>(p);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const uint4  d = __ldg(reinterpret_cast<const uint4 *>(q + 3));
    r.x[0] = a.x; r.x[1] = a.y; r.x[2] = a.z; r.x[3] = a.w;
    r.y[0] = b.x; r.y[1] = b.y; r.y[2] = b.z; r.y[3] = b.w;
    r.z[0] = c.x; r.z[1] = c.y; r.z[2] = c.z; r.z[3] = c.w;
    r.c0 = d.x; r.c1 = d.y; r.axis = d.z;
}

__device__ __forceinline__ void load_node(const Node64 *p, NodeRegs<double> &r)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a0 = __ldg(q), a1 = __ldg(q + 1), b0 = __ldg(q + 2), b1 = __ldg(q + 3), c0 = __ldg(q + 4), c1 = __ldg(q + 5);
    const uint4   d  = __ldg(reinterpret_cast<const uint4 *>(q + 6));
    r.x[0] = a0.x; r.x[1] = a0.y; r.x[2] = a1.x; r.x[3] = a1.y;
    r.y[0] = b0.x; r.y[1] = b0.y; r.y[2] = b1.x; r.y[3] = b1.y;
    r.z[0] = c0.x; r.z[1] = c0.y; r.z[2] = c1.x; r.z[3] = c1.y;
    r.c0 = d.x; r.c1 = d.y; r.axis = d.z;
}

template <typename Real> struct TriRegs { Real v0[3], e1[3], e2[3]; uint32_t prim; };

__device__ __forceinline__ void load_tri(const Tri32 *p, TriRegs<float> &r)
{
    const float4 *q = reinterpret_cast<const float4 *>(p);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    r.v0[0] = a.x; r.v0[1] = a.y; r.v0[2] = a.z; r.prim = __float_as_uint(a.w);
    r.e1[0] = b.x; r.e1[1] = b.y; r.e1[2] = b.z;
    r.e2[0] = c.x; r.e2[1] = c.y; r.e2[2] = c.z;
}

__device__ __forceinline__ void load_tri(const Tri64 *p, TriRegs<double> &r)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4), f = __ldg(q + 5);
    r.v0[0] = a.x; r.v0[1] = a.y; r.v0[2] = b.x; r.prim = (uint32_t)__double_as_longlong(b.y);
    r.e1[0] = c.x; r.e1[1] = c.y; r.e1[2] = d.x;
    r.e2[0] = e.x; r.e2[1] = e.y; r.e2[2] = f.x;
}

// 256-bit global loads (sm_100: LDG.E.256): one L1 wavefront moves 32 bytes of a lane's record instead of 16 ---------
struct __align__(32) F8 { float v[8]; };
struct __align__(32) D4 { double v[4]; };

__device__ __forceinline__ F8 ldg256(const void *p)
{
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ D4 ldg256d(const void *p)
{
    D4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}

__device__ __forceinline__ void load_node_wide(const Node32 *p, NodeRegs<float> &r)
{
    const F8 a = ldg256(p), b = ldg256(reinterpret_cast<const char *>(p) + 32);
    r.x[0] = a.v[0]; r.x[1] = a.v[1]; r.x[2] = a.v[2]; r.x[3] = a.v[3];
    r.y[0] = a.v[4]; r.y[1] = a.v[5]; r.y[2] = a.v[6]; r.y[3] = a.v[7];
    r.z[0] = b.v[0]; r.z[1] = b.v[1]; r.z[2] = b.v[2]; r.z[3] = b.v[3];
    r.c0 = __float_as_uint(b.v[4]); r.c1 = __float_as_uint(b.v[5]); r.axis = __float_as_uint(b.v[6]);
}
__device__ __forceinline__ void load_node_wide(const Node64 *p, NodeRegs<double> &r)
{
    const char *c = reinterpret_cast<const char *>(p);
    const D4 a = ldg256d(c), b = ldg256d(c + 32), d = ldg256d(c + 64);
    const uint4 w = __ldg(reinterpret_cast<const uint4 *>(c + 96));
    r.x[0] = a.v[0]; r.x[1] = a.v[1]; r.x[2] = a.v[2]; r.x[3] = a.v[3];
    r.y[0] = b.v[0]; r.y[1] = b.v[1]; r.y[2] = b.v[2]; r.y[3] = b.v[3];
    r.z[0] = d.v[0]; r.z[1] = d.v[1]; r.z[2] = d.v[2]; r.z[3] = d.v[3];
    r.c0 = w.x; r.c1 = w.y; r.axis = w.z;
}

// a PAIR of fp32 triangle slots (even slot first; leaves start at multiples of four): 96 contiguous, 32-byte-aligned bytes = 3 x LDG.256
__device__ __forceinline__ void load_tri_pair(const Tri32 *p, TriRegs<float> &a, TriRegs<float> &b)
{
    const char *c = reinterpret_cast<const char *>(p);
    const F8 q0 = ldg256(c), q1 = ldg256(c + 32), q2 = ldg256(c + 64);
    a.v0[0] = q0.v[0]; a.v0[1] = q0.v[1]; a.v0[2] = q0.v[2]; a.prim = __float_as_uint(q0.v[3]);
    a.e1[0] = q0.v[4]; a.e1[1] = q0.v[5]; a.e1[2] = q0.v[6];
    a.e2[0] = q1.v[0]; a.e2[1] = q1.v[1]; a.e2[2] = q1.v[2];
    b.v0[0] = q1.v[4]; b.v0[1] = q1.v[5]; b.v0[2] = q1.v[6]; b.prim = __float_as_uint(q1.v[7]);
    b.e1[0] = q2.v[0]; b.e1[1] = q2.v[1]; b.e1[2] = q2.v[2];
    b.e2[0] = q2.v[4]; b.e2[1] = q2.v[5]; b.e2[2] = q2.v[6];
}
// one fp64 triangle slot: 96 bytes, 32-byte aligned = 3 x LDG.256
__device__ __forceinline__ void load_tri_wide(const Tri64 *p, TriRegs<double> &a)
{
    const char *c = reinterpret_cast<const char *>(p);
    const D4 q0 = ldg256d(c), q1 = ldg256d(c + 32), q2 = ldg256d(c + 64);
    a.v0[0] = q0.v[0]; a.v0[1] = q0.v[1]; a.v0[2] = q0.v[2]; a.prim = (uint32_t)__double_as_longlong(q0.v[3]);
    a.e1[0] = q1.v[0]; a.e1[1] = q1.v[1]; a.e1[2] = q1.v[2];
    a.e2[0] = q2.v[0]; a.e2[1] = q2.v[1]; a.e2[2] = q2.v[2];
}

// bvh.c:869-936: one child's slab test; lo/hi picked by the ray's sign bits ---------------------
template <typename Real>
__device__ __forceinline__ bool slab(Real lox, Real hix, Real loy, Real hiy, Real loz, Real hiz,
                                     const Real org[3], const Real inv[3], bool sx, bool sy, bool sz, Real &tmin_out)
{
    const Real nx = sx ? hix : lox, fx = sx ? lox : hix;
    const Real ny = sy ? hiy : loy, fy = sy ? loy : hiy;
    const Real nz = sz ? hiz : loz, fz = sz ? loz : hiz;
    const Real tnx = (nx - org[0]) * inv[0], tfx = (fx - org[0]) * inv[0];
    const Real tny = (ny - org[1]) * inv[1], tfy = (fy - org[1]) * inv[1];
    Real tmin = (tnx > tny) ? tnx : tny;
    Real tmax = (tfx < tfy) ? tfx : tfy;
    const Real tnz = (nz - org[2]) * inv[2], tfz = (fz - org[2]) * inv[2];
    tmin = (tmin > tnz) ? tmin : tnz;
    tmax = (tmax < tfz) ? tmax : tfz;
    tmin_out = tmin;
    return (tmax > Real(0)) && (tmin <= tmax);
}

// Same test with hardware min/max (FMNMX).  `(a > b) ? a : b` and fmax(a, b) return the same VALUE for every non-NaN
// pair (they can differ only in the sign of a zero, which the comparisons that consume tmin/tmax do not see), and no NaN
// can arise here: 1/dir is finite (|dir| > 1e-14 or +-REAL_MAX), so (plane - org) * inv is finite or +-inf, never 0 * inf.
__device__ __forceinline__ float  rmax(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ float  rmin(float a, float b)   { return fminf(a, b); }
__device__ __forceinline__ double rmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double rmin(double a, double b) { return fmin(a, b); }

template <typename Real>
__device__ __forceinline__ bool slab_mm(Real lox, Real hix, Real loy, Real hiy, Real loz, Real hiz,
                                        const Real org[3], const Real inv[3], bool sx, bool sy, bool sz, Real best_t)
{
    const Real nx = sx ? hix : lox, fx = sx ? lox : hix;
    const Real ny = sy ? hiy : loy, fy = sy ? loy : hiy;
    const Real nz = sz ? hiz : loz, fz = sz ? loz : hiz;
    const Real tnx = (nx - org[0]) * inv[0], tfx = (fx - org[0]) * inv[0];
    const Real tny = (ny - org[1]) * inv[1], tfy = (fy - org[1]) * inv[1];
    const Real tnz = (nz - org[2]) * inv[2], tfz = (fz - org[2]) * inv[2];
    const Real tmin = rmax(rmax(tnx, tny), tnz);
    const Real tmax = rmin(rmin(tfx, tfy), tfz);
    return (tmax > Real(0)) && (tmin <= tmax) && (tmin < best_t);          // bvh.c:925 and 1038-1044
}

// bvh.c:730-791 triangle_isect against the leaf-local closest t ------------------------------
template <typename Real>
__device__ __forceinline__ bool tri_test(const TriRegs<Real> &tr, const Real org[3], const Real dir[3],
                                         Real &t_io, Real &u_io, Real &v_io)
{
    const Real px = dir[1] * tr.e2[2] - dir[2] * tr.e2[1];
    const Real py = dir[2] * tr.e2[0] - dir[0] * tr.e2[2];
    const Real pz = dir[0] * tr.e2[1] - dir[1] * tr.e2[0];
    const Real a  = tr.e1[0] * px + tr.e1[1] * py + tr.e1[2] * pz;
    if (!(Prec<Real>::rabs(a) > Prec<Real>::eps())) return false;
    const Real inva = Real(1) / a;
    const Real sx = org[0] - tr.v0[0], sy = org[1] - tr.v0[1], sz = org[2] - tr.v0[2];
    const Real qx = sy * tr.e1[2] - sz * tr.e1[1];
    const Real qy = sz * tr.e1[0] - sx * tr.e1[2];
    const Real qz = sx * tr.e1[1] - sy * tr.e1[0];
    const Real u = (sx * px + sy * py + sz * pz) * inva;
    const Real v = (qx * dir[0] + qy * dir[1] + qz * dir[2]) * inva;
    const Real t = (tr.e2[0] * qx + tr.e2[1] * qy + tr.e2[2] * qz) * inva;
    if ((u < Real(0)) || (u > Real(1))) return false;
    if ((v < Real(0)) || ((u + v) > Real(1))) return false;
    if ((t < Real(0)) || (t > t_io)) return false;
    t_io = t; u_io = u; v_io = v;
    return true;
}

// Branch-free form of the same test for the persistent kernel: every lane computes the whole expression tree (in SIMT the
// warp pays for it as soon as one lane needs it) and the acceptance window becomes one predicate.  The NEGATED comparisons
// are the reference's own (`if ((u < 0) || (u > 1)) return 0;` ...), so NaNs fall the same way.  `valid` masks the filler
// slot of an odd leaf.  Returns the updated leaf-local (t, u, v, prim) through the references.
template <typename Real>
__device__ __forceinline__ void tri_test_bf(const TriRegs<Real> &tr, const Real org[3], const Real dir[3], const bool valid,
                                            Real &t_io, Real &u_io, Real &v_io, uint32_t &prim_io)
{
    const Real px = dir[1] * tr.e2[2] - dir[2] * tr.e2[1];
    const Real py = dir[2] * tr.e2[0] - dir[0] * tr.e2[2];
    const Real pz = dir[0] * tr.e2[1] - dir[1] * tr.e2[0];
    const Real a  = tr.e1[0] * px + tr.e1[1] * py + tr.e1[2] * pz;
    const Real inva = Real(1) / a;
    const Real sx = org[0] - tr.v0[0], sy = org[1] - tr.v0[1], sz = org[2] - tr.v0[2];
    const Real qx = sy * tr.e1[2] - sz * tr.e1[1];
    const Real qy = sz * tr.e1[0] - sx * tr.e1[2];
    const Real qz = sx * tr.e1[1] - sy * tr.e1[0];
    const Real u = (sx * px + sy * py + sz * pz) * inva;
    const Real v = (qx * dir[0] + qy * dir[1] + qz * dir[2]) * inva;
    const Real t = (tr.e2[0] * qx + tr.e2[1] * qy + tr.e2[2] * qz) * inva;
    const bool ok = valid && (Prec<Real>::rabs(a) > Prec<Real>::eps())
                    && !(u < Real(0)) && !(u > Real(1)) && !(v < Real(0)) && !((u + v) > Real(1))
                    && !(t < Real(0)) && !(t > t_io);
    t_io = ok ? t : t_io; u_io = ok ? u : u_io; v_io = ok ? v : v_io; prim_io = ok ? tr.prim : prim_io;
}

// One ray through the tree.  `stk` points at this lane's column of the shared stack, entries are
// `stride` words apart.  Returns hit flag; on hit fills t,u,v,prim (post-build triangle position).
template <typename Real, bool ANYHIT, bool COUNT>
__device__ __forceinline__ bool trace_ray(const SceneView<Real> &S, const Real org[3], const Real dir[3],
                                          uint32_t *stk, const uint32_t stride,
                                          Real &best_t, Real &best_u, Real &best_v, uint32_t &best_prim,
                                          LaneCounters *cnt)
{
    using P = Prec<Real>;
    best_t = P::inf(); best_u = Real(0); best_v = Real(0); best_prim = 0xffffffffu;

    uint32_t cur = S.root_word;
    if (cur == kDoneWord) return false;                       // empty scene, bvh.c:446

    // bvh.c:473-497 (intended rule for all three axes)
    Real inv[3];
    const bool sx = dir[0] < Real(0), sy = dir[1] < Real(0), sz = dir[2] < Real(0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        inv[k] = (P::rabs(dir[k]) > P::eps()) ? Real(1) / dir[k] : ((dir[k] < Real(0)) ? -P::vmax() : P::vmax());
    }
    {                                                         // bvh.c:522-526
        Real tmin;
        if (!slab<Real>(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], org, inv, sx, sy, sz, tmin))
            return false;
    }

    uint32_t sp = 0;
    for (;;) {
        // ---- inner nodes: descend until a leaf word or the stack runs dry
        while (!(cur & kLeafFlag)) {
            NodeRegs<Real> n;
            load_node(S.nodes + cur, n);
            if (COUNT) cnt->ninner++;
            Real tmin0, tmin1;
            const bool h0 = slab<Real>(n.x[0], n.x[1], n.y[0], n.y[1], n.z[0], n.z[1], org, inv, sx, sy, sz, tmin0) && (tmin0 < best_t);
            const bool h1 = slab<Real>(n.x[2], n.x[3], n.y[2], n.y[3], n.z[2], n.z[3], org, inv, sx, sy, sz, tmin1) && (tmin1 < best_t);
            if (h0 && h1) {                                   // bvh.c:1171-1178: near child = child[dir_sign[axis0]]
                const bool order = (n.axis == 0) ? sx : ((n.axis == 1) ? sy : sz);
                stk[sp * stride] = order ? n.c0 : n.c1;
                ++sp;
                cur = order ? n.c1 : n.c0;
            } else if (h0) {
                cur = n.c0;
            } else if (h1) {
                cur = n.c1;
            } else {
                if (sp == 0) { cur = kDoneWord; break; }
                --sp;
                cur = stk[sp * stride];
            }
        }
        if (cur == kDoneWord) break;

        // ---- leaf: bvh.c:793-864
        {
            const uint32_t start = cur & ((1u << kLeafShift) - 1u);
            const uint32_t count = ((cur >> kLeafShift) & 15u) + 1u;
            Real tl = P::inf(), ul = Real(0), vl = Real(0);
            uint32_t tprim = 0;
            bool any = false;
            if (COUNT) { cnt->nleaf++; cnt->ntris += count; }
            for (uint32_t i = 0; i < count; ++i) {
                TriRegs<Real> tr;
                load_tri(S.tris + start + i, tr);
                if (tri_test<Real>(tr, org, dir, tl, ul, vl)) { tprim = tr.prim; any = true; if (COUNT) cnt->nhit++; }
            }
            if (any && (tl < best_t)) {                       // bvh.c:850
                best_t = tl; best_u = ul; best_v = vl; best_prim = tprim;
                if (ANYHIT) break;
            }
        }
        if (sp == 0) break;
        --sp;
        cur = stk[sp * stride];
    }
    return best_t < P::inf();                                 // bvh.c:1187
}

}  // namespace b200

// Host-side BVH construction -- see bvh_build.h.  Reference behaviour: src/render/bvh.c
// (ri_bvh_build 276-379, bvh_construct 1328-1564, bin_triangle_edge 1571-1692,
//  find_cut_from_bin 1230-1326, SAH 1210-1228, bbox_add_margin 1697-1731).
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <deque>
#include <memory>
#include <thread>

namespace b200 {
namespace {

constexpr int    kLeafTris = 16;       // BVH_NTRIS_LEAF, bvh.c:81
constexpr int    kBins     = 64;       // BVH_BIN_SIZE,  bvh.c:82
constexpr double kEps      = 1.0e-14;  // RI_EPS, base/common.h:27
constexpr double kInf      = 1.0e38;   // RI_INFINITY, include/ri.h:47

struct Box {
    double   lo[3], hi[3];
    uint64_t idx;
};

struct TmpNode {
    int32_t  is_leaf = 0, axis = 0;
    TmpNode *c[2] = {nullptr, nullptr};
    uint64_t left = 0, n = 0;          // leaf range in the final order
    double   lbox[6], rbox[6];
};

struct Task {
    TmpNode **slot;
    int       buf;
    uint64_t  left, right;
    double    bmin[3], bmax[3];
    int       depth;
};

struct Arena {
    std::deque<TmpNode> nodes;
    int max_depth = 0;
    TmpNode *alloc() { nodes.emplace_back(); return &nodes.back(); }
};

struct Ctx {
    Box         *buf[2];
    const double *tri_in;
    double      *tri_out;
    uint32_t    *orig_out;
    uint64_t     grain;        // ranges <= grain become parallel tasks (0: never)
    std::vector<Task> *tasks;
    int          nthreads;     // for the big top-level ranges
};

inline void add_margin(double lo[3], double hi[3])
{
    double m[3];
    for (int k = 0; k < 3; ++k) {
        const double extent = hi[k] - lo[k];
        m[k] = (extent < kEps) ? kEps : kEps * extent;
    }
    for (int k = 0; k < 3; ++k) { lo[k] -= m[k]; hi[k] += m[k]; }
}

inline double half_area2(const double lo[3], const double hi[3])
{
    double sa = (hi[0] - lo[0]) * (hi[1] - lo[1]) + (hi[1] - lo[1]) * (hi[2] - lo[2]) + (hi[2] - lo[2]) * (hi[0] - lo[0]);
    sa *= 2.0;
    return sa;
}

// The reference evaluates the cost in double and stores it in a float before comparing (bvh.c:1218-1227).
inline double sah(uint64_t nl, double al, uint64_t nr, double ar, double total)
{
    const float t_aabb = 0.2f, t_tri = 0.8f;
    float cost = 2.0f * t_aabb + (al / total) * (double)(int)nl * t_tri + (ar / total) * (double)(int)nr * t_tri;
    return cost;
}

struct Split { int axis; double pos; };

// Ranges of at least kBigRange boxes (the top few levels of a large scene) are binned and partitioned by all host threads;
// integer bin counts, min/max box unions and the stable-left / reversed-right placement are order-independent, so the
// result is the same tree.
constexpr uint64_t kBigRange = 1ull << 19;

template <typename F> void parallel_chunks(int nthreads, uint64_t n, F fn)
{
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) {
        const uint64_t lo = n * (uint64_t)w / (uint64_t)nthreads, hi = n * (uint64_t)(w + 1) / (uint64_t)nthreads;
        pool.emplace_back([=]() { fn(w, lo, hi); });
    }
    for (auto &t : pool) t.join();
}

struct Bins { uint32_t c[2][3][kBins]; };

void count_bins(const Box *b, uint64_t lo_i, uint64_t hi_i, const double lo[3], const double inv[3], Bins &bins)
{
    std::memset(&bins, 0, sizeof(bins));
    for (uint64_t i = lo_i; i < hi_i; ++i) {
        for (int k = 0; k < 3; ++k) {
            uint32_t a = (uint32_t)((b[i].lo[k] - lo[k]) * inv[k]);
            uint32_t c = (uint32_t)((b[i].hi[k] - lo[k]) * inv[k]);
            if (a >= (uint32_t)kBins) a = kBins - 1;
            if (c >= (uint32_t)kBins) c = kBins - 1;
            bins.c[0][k][a]++;
            bins.c[1][k][c]++;
        }
    }
}

Split choose_split(const Box *b, uint64_t n, const double lo[3], const double hi[3], int nthreads)
{
    double inv[3];
    for (int k = 0; k < 3; ++k) {
        const double extent = hi[k] - lo[k];
        inv[k] = (extent > kEps) ? (double)kBins / extent : 0.0;
    }
    Bins tot;
    if (nthreads > 1 && n >= kBigRange) {
        std::vector<Bins> part((size_t)nthreads);
        parallel_chunks(nthreads, n, [&](int w, uint64_t a, uint64_t c) { count_bins(b, a, c, lo, inv, part[(size_t)w]); });
        std::memset(&tot, 0, sizeof(tot));
        for (const Bins &p : part)
            for (int s = 0; s < 2; ++s) for (int k = 0; k < 3; ++k) for (int i = 0; i < kBins; ++i) tot.c[s][k][i] += p.c[s][k][i];
    } else {
        count_bins(b, 0, n, lo, inv, tot);
    }
    uint32_t (*bins)[3][kBins] = tot.c;

    Split best{0, 0.0};
    double best_cost = kInf;
    const double total = half_area2(lo, hi);
    for (int j = 0; j < 3; ++j) {
        const double step = (hi[j] - lo[j]) / (double)kBins;
        uint64_t nl = 0, nr = n;
        double llo[3] = {lo[0], lo[1], lo[2]}, lhi[3] = {hi[0], hi[1], hi[2]};
        double rlo[3] = {lo[0], lo[1], lo[2]}, rhi[3] = {hi[0], hi[1], hi[2]};
        for (int i = 0; i < kBins - 1; ++i) {
            nl += bins[0][j][i];
            nr -= bins[1][j][i];
            const double pos = lo[j] + (i + 1) * step;
            lhi[j] = pos;
            rlo[j] = pos;
            const double cost = sah(nl, half_area2(llo, lhi), nr, half_area2(rlo, rhi), total);
            if (cost < best_cost) { best_cost = cost; best.axis = j; best.pos = pos; }
        }
    }
    return best;
}

inline void grow(double lo[3], double hi[3], const Box &b)
{
    for (int k = 0; k < 3; ++k) {
        lo[k] = (lo[k] < b.lo[k]) ? lo[k] : b.lo[k];
        hi[k] = (hi[k] > b.hi[k]) ? hi[k] : b.hi[k];
    }
}

void range_box(const Box *b, uint64_t n, double lo[3], double hi[3])
{
    for (int k = 0; k < 3; ++k) { lo[k] = b[0].lo[k]; hi[k] = b[0].hi[k]; }
    for (uint64_t i = 1; i < n; ++i) grow(lo, hi, b[i]);
}

void build_range(Ctx &cx, Arena &ar, TmpNode **slot, int cur, uint64_t left, uint64_t right,
                 const double lo[3], const double hi[3], int depth)
{
    const uint64_t n = right - left;
    if (cx.tasks && depth > 0 && n <= cx.grain && n > (uint64_t)kLeafTris) {   // defer to the pool
        Task t;
        t.slot = slot; t.buf = cur; t.left = left; t.right = right; t.depth = depth;
        for (int k = 0; k < 3; ++k) { t.bmin[k] = lo[k]; t.bmax[k] = hi[k]; }
        cx.tasks->push_back(t);
        return;
    }

    TmpNode *node = ar.alloc();
    *slot = node;
    if (depth > ar.max_depth) ar.max_depth = depth;

    if (n <= (uint64_t)kLeafTris) {                       // bvh.c:1352-1403
        const Box *b = cx.buf[cur] + left;
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t src = b[i].idx;
            std::memcpy(cx.tri_out + 9 * (left + i), cx.tri_in + 9 * src, 9 * sizeof(double));
            cx.orig_out[left + i] = (uint32_t)src;
        }
        node->is_leaf = 1; node->left = left; node->n = n;
        return;
    }

    const Box *src = cx.buf[cur] + left;
    Box *dst = cx.buf[cur ^ 1] + left;
    const Split sp = choose_split(src, n, lo, hi, cx.nthreads);

    // partition: "bmax[axis] < cut" goes left in input order, the rest fills the right part from the
    // back (so it comes out reversed) -- bvh.c:1437-1468.  Child boxes are accumulated in the same sweep.
    uint64_t nl = 0, nr = n - 1;
    double llo[3], lhi[3], rlo[3], rhi[3];
    bool lseen = false, rseen = false;
    if (cx.nthreads > 1 && n >= kBigRange) {
        // two passes: per-chunk left counts, then every chunk writes its elements where the sequential sweep would have
        const int T = cx.nthreads;
        std::vector<uint64_t> cnt((size_t)T, 0);
        parallel_chunks(T, n, [&](int w, uint64_t a, uint64_t c) {
            uint64_t k = 0;
            for (uint64_t i = a; i < c; ++i) k += (src[i].hi[sp.axis] < sp.pos) ? 1 : 0;
            cnt[(size_t)w] = k;
        });
        std::vector<uint64_t> lbase((size_t)T, 0), rbase((size_t)T, 0);      // elements placed before this chunk
        uint64_t lsum = 0, rsum = 0;
        for (int w = 0; w < T; ++w) {
            const uint64_t a = n * (uint64_t)w / (uint64_t)T, c = n * (uint64_t)(w + 1) / (uint64_t)T;
            lbase[(size_t)w] = lsum; rbase[(size_t)w] = rsum;
            lsum += cnt[(size_t)w]; rsum += (c - a) - cnt[(size_t)w];
        }
        struct Acc { double llo[3], lhi[3], rlo[3], rhi[3]; bool ls, rs; };
        std::vector<Acc> acc((size_t)T);
        parallel_chunks(T, n, [&](int w, uint64_t a, uint64_t c) {
            Acc &A = acc[(size_t)w];
            A.ls = A.rs = false;
            uint64_t l = lbase[(size_t)w], r = n - 1 - rbase[(size_t)w];
            for (uint64_t i = a; i < c; ++i) {
                const Box &b = src[i];
                if (b.hi[sp.axis] < sp.pos) {
                    dst[l++] = b;
                    if (!A.ls) { for (int k = 0; k < 3; ++k) { A.llo[k] = b.lo[k]; A.lhi[k] = b.hi[k]; } A.ls = true; }
                    else grow(A.llo, A.lhi, b);
                } else {
                    dst[r--] = b;
                    if (!A.rs) { for (int k = 0; k < 3; ++k) { A.rlo[k] = b.lo[k]; A.rhi[k] = b.hi[k]; } A.rs = true; }
                    else grow(A.rlo, A.rhi, b);
                }
            }
        });
        nl = lsum; nr = n - 1 - rsum;
        for (const Acc &A : acc) {
            if (A.ls) {
                if (!lseen) { for (int k = 0; k < 3; ++k) { llo[k] = A.llo[k]; lhi[k] = A.lhi[k]; } lseen = true; }
                else for (int k = 0; k < 3; ++k) { llo[k] = (llo[k] < A.llo[k]) ? llo[k] : A.llo[k]; lhi[k] = (lhi[k] > A.lhi[k]) ? lhi[k] : A.lhi[k]; }
            }
            if (A.rs) {
                if (!rseen) { for (int k = 0; k < 3; ++k) { rlo[k] = A.rlo[k]; rhi[k] = A.rhi[k]; } rseen = true; }
                else for (int k = 0; k < 3; ++k) { rlo[k] = (rlo[k] < A.rlo[k]) ? rlo[k] : A.rlo[k]; rhi[k] = (rhi[k] > A.rhi[k]) ? rhi[k] : A.rhi[k]; }
            }
        }
    } else
    for (uint64_t i = 0; i < n; ++i) {
        const Box &b = src[i];
        if (b.hi[sp.axis] < sp.pos) {
            dst[nl++] = b;
            if (!lseen) { for (int k = 0; k < 3; ++k) { llo[k] = b.lo[k]; lhi[k] = b.hi[k]; } lseen = true; }
            else grow(llo, lhi, b);
        } else {
            dst[nr--] = b;
            if (!rseen) { for (int k = 0; k < 3; ++k) { rlo[k] = b.lo[k]; rhi[k] = b.hi[k]; } rseen = true; }
            else grow(rlo, rhi, b);
        }
    }
    if (nl == 0 || nl == n) {                             // object-median fallback, bvh.c:1471-1478
        nl = n / 2;
        range_box(dst, nl, llo, lhi);
        range_box(dst + nl, n - nl, rlo, rhi);
    }
    add_margin(llo, lhi);
    add_margin(rlo, rhi);

    node->axis = sp.axis;
    for (int k = 0; k < 3; ++k) {
        node->lbox[k] = llo[k]; node->lbox[3 + k] = lhi[k];
        node->rbox[k] = rlo[k]; node->rbox[3 + k] = rhi[k];
    }
    build_range(cx, ar, &node->c[0], cur ^ 1, left, left + nl, llo, lhi, depth + 1);
    build_range(cx, ar, &node->c[1], cur ^ 1, left + nl, right, rlo, rhi, depth + 1);
}

int64_t emit(const TmpNode *n, std::vector<CanonNode> &out, int depth, HostTree &t)
{
    const int64_t me = (int64_t)out.size();
    out.emplace_back();
    if (depth > t.max_depth) t.max_depth = depth;
    {
        CanonNode &c = out[me];
        std::memset(&c, 0, sizeof(c));
        c.is_leaf = n->is_leaf;
        if (n->is_leaf) {
            c.child0 = c.child1 = -1;
            c.tri_start = (int64_t)n->left; c.ntris = (int64_t)n->n;
            t.nleaf++;
            return me;
        }
        c.axis = n->axis;
        std::memcpy(c.lbox, n->lbox, sizeof(c.lbox));
        std::memcpy(c.rbox, n->rbox, sizeof(c.rbox));
        t.ninner++;
    }
    const int64_t c0 = emit(n->c[0], out, depth + 1, t);
    out[me].child0 = c0;
    const int64_t c1 = emit(n->c[1], out, depth + 1, t);
    out[me].child1 = c1;
    return me;
}

}  // namespace

void build_tree(const double *tri_xyz, uint64_t ntris, HostTree &out, int nthreads)
{
    const auto t0 = std::chrono::steady_clock::now();
    out = HostTree();
    out.ntris = ntris;
    if (ntris == 0) { out.empty = true; return; }         // bvh.c:311-315
    out.empty = false;

    std::unique_ptr<Box[]> a(new Box[ntris]), b(new Box[ntris]);
    for (uint64_t i = 0; i < ntris; ++i) {                // get_bbox_of_triangle, bvh.c:1852-1868
        const double *v = tri_xyz + 9 * i;
        Box &bx = a[i];
        for (int k = 0; k < 3; ++k) {
            double mn = v[k], mx = v[k];
            for (int c = 1; c < 3; ++c) {
                mn = (mn < v[3 * c + k]) ? mn : v[3 * c + k];
                mx = (mx > v[3 * c + k]) ? mx : v[3 * c + k];
            }
            bx.lo[k] = mn; bx.hi[k] = mx;
        }
        bx.idx = i;
    }
    range_box(a.get(), ntris, out.bmin, out.bmax);        // calc_scene_bbox, bvh.c:1829-1849
    add_margin(out.bmin, out.bmax);                       // bvh.c:330

    out.tri.resize(9 * ntris);
    out.orig.resize(ntris);

    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;

    std::vector<Task> tasks;
    Ctx cx;
    cx.buf[0] = a.get(); cx.buf[1] = b.get();
    cx.tri_in = tri_xyz; cx.tri_out = out.tri.data(); cx.orig_out = out.orig.data();
    const bool parallel = nthreads > 1 && ntris >= (1u << 16);
    cx.grain = parallel ? std::max<uint64_t>(ntris / (uint64_t)(8 * nthreads), 4096) : 0;
    cx.tasks = parallel ? &tasks : nullptr;
    cx.nthreads = parallel ? nthreads : 1;

    Arena top;
    TmpNode *root = nullptr;
    build_range(cx, top, &root, 0, 0, ntris, out.bmin, out.bmax, 0);

    std::vector<Arena> arenas((size_t)nthreads);
    if (!tasks.empty()) {
        std::sort(tasks.begin(), tasks.end(),
                  [](const Task &x, const Task &y) { return (x.right - x.left) > (y.right - y.left); });
        Ctx sub = cx;
        sub.tasks = nullptr; sub.grain = 0; sub.nthreads = 1;
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        for (int w = 0; w < nthreads; ++w) {
            pool.emplace_back([&, w]() {
                Ctx local = sub;
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= tasks.size()) break;
                    const Task &t = tasks[i];
                    build_range(local, arenas[(size_t)w], t.slot, t.buf, t.left, t.right, t.bmin, t.bmax, t.depth);
                }
            });
        }
        for (auto &th : pool) th.join();
    }

    out.nodes.reserve((size_t)(ntris / 4 + 16));
    emit(root, out.nodes, 0, out);
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------------

static float f32_down(double x) { float f = (float)x; if ((double)f > x) f = std::nextafterf(f, -INFINITY); return f; }
static float f32_up(double x)   { float f = (float)x; if ((double)f < x) f = std::nextafterf(f,  INFINITY); return f; }

void flatten_tree(const HostTree &t, uint32_t top_nodes, bool want32, bool want64, FlatTree &out, bool fill_tris)
{
    out = FlatTree();
    for (int k = 0; k < 3; ++k) { out.smin32[k] = f32_down(t.bmin[k]); out.smax32[k] = f32_up(t.bmax[k]); }
    if (t.empty) { out.root_word = kDoneWord; return; }

    const size_t nn = t.nodes.size();
    std::vector<uint32_t> dev(nn, 0xffffffffu);           // canonical index -> device inner index
    std::vector<int64_t>  order;                          // device inner index -> canonical index
    order.reserve((size_t)t.ninner);

    // BFS top cluster
    std::vector<int64_t> frontier;
    if (!t.nodes[0].is_leaf) {
        std::deque<int64_t> q;
        q.push_back(0);
        while (!q.empty() && order.size() < top_nodes) {
            const int64_t c = q.front(); q.pop_front();
            dev[(size_t)c] = (uint32_t)order.size();
            order.push_back(c);
            const CanonNode &n = t.nodes[(size_t)c];
            if (!t.nodes[(size_t)n.child0].is_leaf) q.push_back(n.child0);
            if (!t.nodes[(size_t)n.child1].is_leaf) q.push_back(n.child1);
        }
        out.top_count = (uint32_t)order.size();
        frontier.assign(q.begin(), q.end());
    }
    // DFS-preorder remainder, one frontier subtree after another
    std::vector<int64_t> stack;
    for (int64_t f : frontier) {
        stack.push_back(f);
        while (!stack.empty()) {
            const int64_t c = stack.back(); stack.pop_back();
            dev[(size_t)c] = (uint32_t)order.size();
            order.push_back(c);
            const CanonNode &n = t.nodes[(size_t)c];
            if (!t.nodes[(size_t)n.child1].is_leaf) stack.push_back(n.child1);
            if (!t.nodes[(size_t)n.child0].is_leaf) stack.push_back(n.child0);
        }
    }
    out.ninner = (uint32_t)order.size();

    // triangle slots: each leaf starts at a slot that is a multiple of four and owns round_up(ntris, 4) slots (DFS order of
    // leaves == post-build triangle order): fp32 pairs stay 32-byte aligned, and the rows of the leaf-transposed copies below
    // start on 64-byte (fp32) / 128-byte (fp64) boundaries, which is what lets neighbouring lanes share L1 wavefronts
    std::vector<uint32_t> slot_of(nn, 0);
    uint64_t nslots = 0;
    for (size_t c = 0; c < nn; ++c) {
        const CanonNode &n = t.nodes[c];
        if (!n.is_leaf) continue;
        slot_of[c] = (uint32_t)nslots;
        nslots += (uint64_t)((n.ntris + 3) & ~3ll);
    }
    out.nslots = nslots;
    if (nslots >= (1ull << kLeafShift)) { out = FlatTree(); out.overflow = true; return; }     // the leaf word holds a 27-bit slot
    out.leaf_slot = slot_of;
    if (fill_tris) out.slot_of_prim.assign((size_t)t.ntris, 0);

    auto word = [&](int64_t c) -> uint32_t {
        const CanonNode &n = t.nodes[(size_t)c];
        if (n.is_leaf) return kLeafFlag | ((uint32_t)(n.ntris - 1) << kLeafShift) | slot_of[(size_t)c];
        return dev[(size_t)c];
    };
    out.root_word = word(0);

    if (want32) out.nodes32.resize(order.size());
    if (want64) out.nodes64.resize(order.size());
    for (size_t i = 0; i < order.size(); ++i) {
        const CanonNode &n = t.nodes[(size_t)order[i]];
        const uint32_t c0 = word(n.child0), c1 = word(n.child1);
        if (want32) {
            Node32 &d = out.nodes32[i];
            float *ax[3] = {d.x, d.y, d.z};
            for (int k = 0; k < 3; ++k) {
                ax[k][0] = f32_down(n.lbox[k]); ax[k][1] = f32_up(n.lbox[3 + k]);
                ax[k][2] = f32_down(n.rbox[k]); ax[k][3] = f32_up(n.rbox[3 + k]);
            }
            d.c0 = c0; d.c1 = c1; d.axis = (uint32_t)n.axis; d.pad = 0;
        }
        if (want64) {
            Node64 &d = out.nodes64[i];
            double *ax[3] = {d.x, d.y, d.z};
            for (int k = 0; k < 3; ++k) {
                ax[k][0] = n.lbox[k]; ax[k][1] = n.lbox[3 + k];
                ax[k][2] = n.rbox[k]; ax[k][3] = n.rbox[3 + k];
            }
            d.c0 = c0; d.c1 = c1; d.axis = (uint32_t)n.axis; d.pad = 0;
            d.pad2[0] = d.pad2[1] = d.pad2[2] = d.pad2[3] = 0;
        }
    }

    if (!fill_tris) return;
    if (want32) { out.tris32.resize((size_t)nslots); std::memset(out.tris32.data(), 0, out.tris32.size() * sizeof(Tri32)); }
    if (want64) { out.tris64.resize((size_t)nslots); std::memset(out.tris64.data(), 0, out.tris64.size() * sizeof(Tri64)); }
    for (size_t c = 0; c < nn; ++c) {
        const CanonNode &n = t.nodes[c];
        if (!n.is_leaf) continue;
        for (int64_t i = 0; i < n.ntris; ++i) {
            const uint64_t p = (uint64_t)(n.tri_start + i), slot = (uint64_t)slot_of[c] + (uint64_t)i;
            const double *v = t.tri.data() + 9 * p;
            out.slot_of_prim[(size_t)p] = (uint32_t)slot;
            if (want32) {
                Tri32 &d = out.tris32[(size_t)slot];
                for (int k = 0; k < 3; ++k) {
                    const float v0 = (float)v[k], v1 = (float)v[3 + k], v2 = (float)v[6 + k];
                    d.v0[k] = v0; d.e1[k] = v1 - v0; d.e2[k] = v2 - v0;      // bvh.c:747-752, in fp32
                }
                d.prim = (uint32_t)p;
            }
            if (want64) {
                Tri64 &d = out.tris64[(size_t)slot];
                for (int k = 0; k < 3; ++k) { d.v0[k] = v[k]; d.e1[k] = v[3 + k] - v[k]; d.e2[k] = v[6 + k] - v[k]; }
                d.prim = p;
            }
        }
        for (int64_t i = n.ntris; i < ((n.ntris + 3) & ~3ll); ++i) {      // filler slots: zero-area triangles, prim = MISS
            const uint64_t slot = (uint64_t)slot_of[c] + (uint64_t)i;
            if (want32) out.tris32[(size_t)slot].prim = 0xffffffffu;
            if (want64) out.tris64[(size_t)slot].prim = 0xffffffffull;
        }
    }

    // leaf-transposed copies for the pooled occlusion kernel
    if (want32) out.tris32t.resize((size_t)nslots);
    if (want64) out.tris64t.resize((size_t)nslots);
    for (size_t c = 0; c < nn; ++c) {
        const CanonNode &n = t.nodes[c];
        if (!n.is_leaf) continue;
        const uint64_t slot0 = slot_of[c], ns = (uint64_t)((n.ntris + 3) & ~3ll);
        if (want32) {                                        // item = pair of slots = 3 chunks of 32 B; row = ns/2 chunks
            // Inside an item the two triangles A (even slot) and B (odd slot) are INTERLEAVED word by word so that every field
            // arrives as an aligned (A, B) register pair for the packed fp32 arithmetic of the pooled kernels (FFMA2, packed.cuh):
            // words (2f, 2f+1) = field f of (A, B), f = v0.xyz e1.xyz e2.xyz; words 18, 19 = prim of (A, B).  Rows 0 and 1 hold words
            // 0..7 / 8..15 of every item (32 B per item), row 2 holds words 16..19 (16 B per item, at byte 64 m of the leaf's block):
            // a leaf's third row fits one 128-byte line and is read with LDG.128 (8 lanes per L1 pass instead of 4).
            const uint64_t m = ns / 2, used = (uint64_t)((n.ntris + 1) / 2);
            const Tri32 *src = out.tris32.data() + slot0;
            char *dst = reinterpret_cast<char *>(out.tris32t.data() + slot0);
            for (uint64_t j = 0; j < m; ++j) {
                uint32_t w[24];
                std::memset(w, 0, sizeof(w));
                for (int h = 0; h < 2; ++h) {
                    const Tri32 &s = src[2 * j + (uint64_t)h];
                    const float f[9] = {s.v0[0], s.v0[1], s.v0[2], s.e1[0], s.e1[1], s.e1[2], s.e2[0], s.e2[1], s.e2[2]};
                    for (int q = 0; q < 9; ++q) std::memcpy(&w[2 * q + h], &f[q], 4);
                    w[18 + h] = s.prim;
                }
                if ((n.ntris & 1) && j == used - 1) {
                    // the filler half of the last pair is masked by its validity bit, never by its determinant: give it unit edges
                    // so that 1/det stays on the fast path of the reciprocal (a zero determinant takes the slow one)
                    const float one = 1.0f;
                    std::memcpy(&w[2 * 3 + 1], &one, 4);     // B.e1 = (1,0,0)
                    std::memcpy(&w[2 * 7 + 1], &one, 4);     // B.e2 = (0,1,0)
                }
                std::memcpy(dst + j * 32, w, 32);
                std::memcpy(dst + (m + j) * 32, w + 8, 32);
                std::memcpy(dst + 2 * m * 32 + j * 16, w + 16, 16);
            }
        }
        if (want64) {                                        // item = one slot = 3 chunks of 32 B
            const char *src = reinterpret_cast<const char *>(out.tris64.data() + slot0);
            char *dst = reinterpret_cast<char *>(out.tris64t.data() + slot0);
            for (uint64_t j = 0; j < ns; ++j)
                for (uint64_t k = 0; k < 3; ++k) std::memcpy(dst + (k * ns + j) * 32, src + j * 96 + k * 32, 32);
        }
    }
}

}  // namespace b200

// BVH construction ON THE DEVICE: the same tree as ri_bvh_build() / bvh_build.cpp, built level by level.
//
// Reference: src/render/bvh.c -- bvh_construct 1328-1564, bin_triangle_edge 1571-1692, find_cut_from_bin 1230-1326, SAH 1210-1228,
// bbox_add_margin 1697-1731, get_bbox_of_triangle 1852-1868, calc_scene_bbox 1829-1849.  The recursion becomes a loop over tree
// levels; every node of a level is processed at once:
//   gb_bin          one lane per box: bins its min/max edges into the 64 x 3 x 2 counters of ITS node (integer atomics)
//   gb_split        one lane per node: the 63 x 3 candidate planes in the reference's loop order, cost in double rounded to float
//   gb_flags+scan   "bbox.bmax[axis] < cut_pos" as a 0/1 flag per box and its exclusive prefix sum over the whole array
//   gb_scatter      left part in input order from the front of the node's range, right part from the BACK (so it comes out
//                   reversed, bvh.c:1460-1464): position = f(prefix sums), no ordering decisions left to chance
//   gb_children     left count (object-median fallback n/2 when one side is empty, bvh.c:1471-1478), two child records
//   gb_child_boxes  union of the boxes of each child range: atomicMin/Max on order-preserving 64-bit keys of the doubles
//   gb_finalize     margin 1e-14 (relative / absolute), leaf test n <= 16
// Integer counts, min/max unions and prefix-sum placement are exact and order-independent, the cost arithmetic is the host's
// expression compiled with -fmad=false: the result is the reference's tree bit for bit (nodes, boxes, leaf order) -- tests compare
// it with the host builder and the oracle.  (One caveat shared with the threaded host builder: a union of -0.0 and +0.0 keeps
// whichever the reference met last; the keys order -0.0 below +0.0.)
// Nodes come out in breadth-first order; the host renumbers them depth-first (the canonical form everything else consumes).
#pragma once

namespace b200 {

struct GBox  { double lo[3], hi[3]; uint32_t idx, node; };
struct GNode {
    double   lo[3], hi[3];        // bmin/bmax handed to bvh_construct for this node (margin included)
    double   cut;
    uint32_t left, n, nl;         // range in the box array; boxes going to child 0
    uint32_t child0, child1;      // indices in this array (breadth-first)
    int32_t  axis, is_leaf;
    uint32_t pad;
};

constexpr int      kGbLeaf = 16, kGbBins = 64;
constexpr double   kGbEps = 1.0e-14, kGbInf = 1.0e38;

__device__ __forceinline__ unsigned long long gb_key(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double gb_unkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// accumulate a box into the six keys at `acc` (min xyz, max xyz); warp-aggregated when the whole warp targets `acc`
__device__ __forceinline__ void gb_accumulate(unsigned long long *acc, const double lo[3], const double hi[3], bool active)
{
    const unsigned all = __ballot_sync(0xffffffffu, true);
    const unsigned long long me = active ? (unsigned long long)(uintptr_t)acc : 0ull;
    const unsigned long long first = __shfl_sync(0xffffffffu, me, 0);
    const bool uniform = (all == 0xffffffffu) && __all_sync(0xffffffffu, me == first) && first != 0ull;
    if (uniform) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            unsigned long long a = gb_key(lo[k]), b = gb_key(hi[k]);
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
                a = a2 < a ? a2 : a; b = b2 > b ? b2 : b;
            }
            if ((threadIdx.x & 31) == 0) { atomicMin(acc + k, a); atomicMax(acc + 3 + k, b); }
        }
    } else if (active) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(acc + k, gb_key(lo[k])); atomicMax(acc + 3 + k, gb_key(hi[k])); }
    }
}

__device__ __forceinline__ void gb_add_margin(double lo[3], double hi[3])      // bvh.c:1697-1731
{
    double m[3];
    for (int k = 0; k < 3; ++k) {
        const double extent = hi[k] - lo[k];
        m[k] = (extent < kGbEps) ? kGbEps : kGbEps * extent;
    }
    for (int k = 0; k < 3; ++k) { lo[k] -= m[k]; hi[k] += m[k]; }
}

__device__ __forceinline__ double gb_half_area2(const double lo[3], const double hi[3])    // bvh.c:1190-1208
{
    double sa = (hi[0] - lo[0]) * (hi[1] - lo[1]) + (hi[1] - lo[1]) * (hi[2] - lo[2]) + (hi[2] - lo[2]) * (hi[0] - lo[0]);
    sa *= 2.0;
    return sa;
}

__device__ __forceinline__ double gb_sah(uint32_t nl, double al, uint32_t nr, double ar, double total)   // bvh.c:1210-1228
{
    const float t_aabb = 0.2f, t_tri = 0.8f;
    const float cost = (float)((double)(2.0f * t_aabb) + (al / total) * (double)(int)nl * (double)t_tri + (ar / total) * (double)(int)nr * (double)t_tri);
    return (double)cost;
}

// ---- triangle boxes + scene box -------------------------------------------------------------------------------------
__global__ void gb_init_boxes(const double *__restrict__ tri, uint32_t n, GBox *__restrict__ boxes, unsigned long long *__restrict__ root_acc)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    if (active) {
        const double *v = tri + 9 * (size_t)i;
        GBox b;
        for (int k = 0; k < 3; ++k) {                                   // get_bbox_of_triangle, bvh.c:1852-1868
            double mn = v[k], mx = v[k];
            for (int c = 1; c < 3; ++c) {
                mn = (mn < v[3 * c + k]) ? mn : v[3 * c + k];
                mx = (mx > v[3 * c + k]) ? mx : v[3 * c + k];
            }
            b.lo[k] = lo[k] = mn; b.hi[k] = hi[k] = mx;
        }
        b.idx = i; b.node = 0;
        boxes[i] = b;
    }
    gb_accumulate(root_acc, lo, hi, active);
}

__global__ void gb_root(GNode *nodes, const unsigned long long *root_acc, uint32_t n)
{
    GNode r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = gb_unkey(root_acc[k]); r.hi[k] = gb_unkey(root_acc[3 + k]); }
    gb_add_margin(r.lo, r.hi);                                          // bvh.c:330
    r.cut = 0.0; r.left = 0; r.n = n; r.nl = 0; r.child0 = r.child1 = 0; r.axis = 0; r.is_leaf = (n <= (uint32_t)kGbLeaf) ? 1 : 0; r.pad = 0;
    nodes[0] = r;
}

// ---- one level ----------------------------------------------------------------------------------------------------------
__global__ void gb_bin(const GBox *__restrict__ boxes, uint32_t n, const GNode *__restrict__ nodes, uint32_t first, uint32_t *__restrict__ hist)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const GBox b = boxes[p];
    const GNode &nd = nodes[b.node];
    if (nd.is_leaf || b.node < first) return;
    uint32_t *h = hist + (size_t)(b.node - first) * (2 * 3 * kGbBins);
    for (int k = 0; k < 3; ++k) {                                       // bin_triangle_edge, bvh.c:1571-1692
        const double extent = nd.hi[k] - nd.lo[k];
        const double inv = (extent > kGbEps) ? (double)kGbBins / extent : 0.0;
        uint32_t a = (uint32_t)((b.lo[k] - nd.lo[k]) * inv);
        uint32_t c = (uint32_t)((b.hi[k] - nd.lo[k]) * inv);
        if (a >= (uint32_t)kGbBins) a = kGbBins - 1;
        if (c >= (uint32_t)kGbBins) c = kGbBins - 1;
        atomicAdd(h + k * kGbBins + a, 1u);
        atomicAdd(h + 3 * kGbBins + k * kGbBins + c, 1u);
    }
}

__global__ void gb_split(GNode *__restrict__ nodes, uint32_t first, uint32_t count, const uint32_t *__restrict__ hist)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    GNode &nd = nodes[first + j];
    if (nd.is_leaf) return;
    const uint32_t *h = hist + (size_t)j * (2 * 3 * kGbBins);
    const double lo[3] = {nd.lo[0], nd.lo[1], nd.lo[2]}, hi[3] = {nd.hi[0], nd.hi[1], nd.hi[2]};
    int best_axis = 0;
    double best_pos = 0.0, best_cost = kGbInf;
    const double total = gb_half_area2(lo, hi);
    for (int a = 0; a < 3; ++a) {                                       // find_cut_from_bin, bvh.c:1230-1326
        const double step = (hi[a] - lo[a]) / (double)kGbBins;
        uint32_t nl = 0, nr = nd.n;
        double llo[3] = {lo[0], lo[1], lo[2]}, lhi[3] = {hi[0], hi[1], hi[2]};
        double rlo[3] = {lo[0], lo[1], lo[2]}, rhi[3] = {hi[0], hi[1], hi[2]};
        for (int i = 0; i < kGbBins - 1; ++i) {
            nl += h[a * kGbBins + i];
            nr -= h[3 * kGbBins + a * kGbBins + i];
            const double pos = lo[a] + (double)(i + 1) * step;
            lhi[a] = pos;
            rlo[a] = pos;
            const double cost = gb_sah(nl, gb_half_area2(llo, lhi), nr, gb_half_area2(rlo, rhi), total);
            if (cost < best_cost) { best_cost = cost; best_axis = a; best_pos = pos; }
        }
    }
    nd.axis = best_axis; nd.cut = best_pos;
}

// flag per box (1 = goes left) and per-tile sums for the scan
__global__ void __launch_bounds__(kScanBlock)
gb_flags(const GBox *__restrict__ boxes, uint32_t n, const GNode *__restrict__ nodes, uint32_t first, uint8_t *__restrict__ flags, uint32_t *__restrict__ tile_sums)
{
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t c = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const uint32_t p = base + k;
        uint8_t f = 0;
        if (p < n) {
            const GBox &b = boxes[p];
            const GNode &nd = nodes[b.node];
            if (!nd.is_leaf && b.node >= first) f = (b.hi[nd.axis] < nd.cut) ? 1 : 0;      // bvh.c:1442
            flags[p] = f;
        }
        c += f;
    }
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// exclusive prefix sums S[0..n] of the flags from the scanned tile sums
__global__ void __launch_bounds__(kScanBlock)
gb_prefix(const uint8_t *__restrict__ flags, uint32_t n, const uint32_t *__restrict__ tile_offsets, uint32_t *__restrict__ S)
{
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t c = 0;
    for (int k = 0; k < kScanItems; ++k) { const uint32_t p = base + k; if (p < n) c += flags[p]; }
    uint32_t run = tile_offsets[blockIdx.x] + block_exclusive_scan(c, nullptr);
    for (int k = 0; k < kScanItems; ++k) {
        const uint32_t p = base + k;
        if (p > n) break;
        S[p] = run;                                                     // p == n writes the grand total
        if (p < n) run += flags[p];
    }
}

__global__ void gb_scatter(const GBox *__restrict__ src, GBox *__restrict__ dst, uint32_t n, const GNode *__restrict__ nodes, uint32_t first,
                           const uint8_t *__restrict__ flags, const uint32_t *__restrict__ S)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const GBox b = src[p];
    const GNode &nd = nodes[b.node];
    if (nd.is_leaf || b.node < first) { dst[p] = b; return; }           // finished ranges stay where they are
    const uint32_t l = nd.left, lrank = S[p] - S[l];
    const uint32_t q = flags[p] ? l + lrank : l + nd.n - 1u - ((p - l) - lrank);      // bvh.c:1437-1468
    dst[q] = b;
}

// acc: six keys per node of the NEXT level, indexed by (child index - next_first)
__global__ void gb_children(GNode *__restrict__ nodes, uint32_t first, uint32_t count, const uint32_t *__restrict__ S,
                            uint32_t *__restrict__ counters /* [0] next node index */, unsigned long long *__restrict__ acc, uint32_t next_first)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    GNode &nd = nodes[first + j];
    if (nd.is_leaf) return;
    uint32_t nl = S[nd.left + nd.n] - S[nd.left];
    if (nl == 0u || nl == nd.n) nl = nd.n / 2u;                         // object-median fallback, bvh.c:1471-1478
    nd.nl = nl;
    const uint32_t c0 = atomicAdd(&counters[0], 2u);
    nd.child0 = c0; nd.child1 = c0 + 1u;
    for (int c = 0; c < 2; ++c) {
        GNode ch;
        for (int k = 0; k < 3; ++k) { ch.lo[k] = 0.0; ch.hi[k] = 0.0; }
        ch.cut = 0.0;
        ch.left = c ? nd.left + nl : nd.left;
        ch.n = c ? nd.n - nl : nl;
        ch.nl = 0; ch.child0 = ch.child1 = 0; ch.axis = 0; ch.is_leaf = 0; ch.pad = 0;
        nodes[c0 + c] = ch;
        unsigned long long *a = acc + (size_t)(c0 + c - next_first) * 6;
        for (int k = 0; k < 3; ++k) { a[k] = ~0ull; a[3 + k] = 0ull; }
    }
}

__global__ void gb_child_boxes(GBox *__restrict__ dst, uint32_t n, const GNode *__restrict__ nodes, uint32_t first,
                               unsigned long long *__restrict__ acc, uint32_t next_first)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = false;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    unsigned long long *a = nullptr;
    if (q < n) {
        GBox &b = dst[q];
        const GNode &nd = nodes[b.node];
        if (!nd.is_leaf && b.node >= first) {
            const uint32_t child = (q < nd.left + nd.nl) ? nd.child0 : nd.child1;
            for (int k = 0; k < 3; ++k) { lo[k] = b.lo[k]; hi[k] = b.hi[k]; }
            b.node = child;
            a = acc + (size_t)(child - next_first) * 6;
            active = true;
        }
    }
    gb_accumulate(a, lo, hi, active);
}

__global__ void gb_finalize(GNode *__restrict__ nodes, uint32_t next_first, uint32_t next_count, const unsigned long long *__restrict__ acc,
                            uint32_t *__restrict__ counters /* [1] inner nodes of the next level */)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= next_count) return;
    GNode &ch = nodes[next_first + j];
    const unsigned long long *a = acc + (size_t)j * 6;
    for (int k = 0; k < 3; ++k) { ch.lo[k] = gb_unkey(a[k]); ch.hi[k] = gb_unkey(a[3 + k]); }
    gb_add_margin(ch.lo, ch.hi);                                        // bvh.c:1486-1561
    ch.is_leaf = (ch.n <= (uint32_t)kGbLeaf) ? 1 : 0;                   // bvh.c:1352
    if (!ch.is_leaf) atomicAdd(&counters[1], 1u);
}

// ---- host driver ------------------------------------------------------------------------------------------------------------
struct GpuBuildScratch {
    double *d_tri = nullptr; GBox *d_box[2] = {nullptr, nullptr}; GNode *d_nodes = nullptr;
    uint32_t *d_hist = nullptr, *d_S = nullptr, *d_tiles = nullptr, *d_counters = nullptr; uint8_t *d_flags = nullptr;
    unsigned long long *d_acc = nullptr;
    ~GpuBuildScratch()
    {
        cudaFree(d_tri); cudaFree(d_box[0]); cudaFree(d_box[1]); cudaFree(d_nodes); cudaFree(d_hist); cudaFree(d_S); cudaFree(d_tiles);
        cudaFree(d_counters); cudaFree(d_flags); cudaFree(d_acc);
    }
};

static int64_t gb_emit(const std::vector<GNode> &g, uint32_t i, std::vector<CanonNode> &out, int depth, HostTree &t, std::vector<uint32_t> &canon_of)
{
    const int64_t me = (int64_t)out.size();
    out.emplace_back();
    canon_of[i] = (uint32_t)me;
    if (depth > t.max_depth) t.max_depth = depth;
    const GNode &n = g[i];
    {
        CanonNode &c = out[(size_t)me];
        std::memset(&c, 0, sizeof(c));
        c.is_leaf = n.is_leaf;
        if (n.is_leaf) {
            c.child0 = c.child1 = -1;
            c.tri_start = (int64_t)n.left; c.ntris = (int64_t)n.n;
            t.nleaf++;
            return me;
        }
        c.axis = n.axis;
        const GNode &l = g[n.child0], &r = g[n.child1];
        for (int k = 0; k < 3; ++k) { c.lbox[k] = l.lo[k]; c.lbox[3 + k] = l.hi[k]; c.rbox[k] = r.lo[k]; c.rbox[3 + k] = r.hi[k]; }
        t.ninner++;
    }
    const int64_t c0 = gb_emit(g, n.child0, out, depth + 1, t, canon_of);
    out[(size_t)me].child0 = c0;
    const int64_t c1 = gb_emit(g, n.child1, out, depth + 1, t, canon_of);
    out[(size_t)me].child1 = c1;
    return me;
}

struct DeviceBuild {                                   // what the slot-filling pass needs after the tree is final
    GpuBuildScratch s;
    int      cur = 0;
    uint32_t n = 0, total = 0;
    std::vector<uint32_t> canon_of;                    // breadth-first node -> canonical (depth-first) node
};

__global__ void gb_extract_idx(const GBox *__restrict__ boxes, uint32_t n, uint32_t *__restrict__ idx)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) idx[p] = boxes[p].idx;
}

// Triangle slots straight from the device copy of the triangles: what flatten_tree()'s per-triangle loops write on the host
// (bvh_build.cpp) -- the slot records of both precisions, their leaf-transposed copies, slot_of_prim, the filler slots.
// One lane per triangle in its final position p (= prim id); buffers are zeroed beforehand.
__global__ void gb_fill_slots(const double *__restrict__ tri, const GBox *__restrict__ boxes, uint32_t n, const GNode *__restrict__ nodes,
                              const uint32_t *__restrict__ leaf_slot, Tri32 *__restrict__ t32, Tri64 *__restrict__ t64,
                              char *__restrict__ t32t, char *__restrict__ t64t, uint32_t *__restrict__ slot_of_prim)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const GBox &b = boxes[p];
    const GNode &nd = nodes[b.node];
    const uint32_t i = p - nd.left, ntris = nd.n, slot0 = leaf_slot[b.node], slot = slot0 + i;
    const uint32_t ns = (ntris + 3u) & ~3u, m = ns / 2u;             // slots / fp32 pairs per row
    const double *v = tri + 9 * (size_t)b.idx;
    slot_of_prim[p] = slot;
    if (t64) {
        Tri64 d;
        for (int k = 0; k < 3; ++k) { d.v0[k] = v[k]; d.e1[k] = v[3 + k] - v[k]; d.e2[k] = v[6 + k] - v[k]; }
        d.prim = p; d.pad1 = 0; d.pad2 = 0;
        t64[slot] = d;
        const uint4 *src = reinterpret_cast<const uint4 *>(&d);
        for (uint32_t k = 0; k < 3; ++k) {
            uint4 *dst = reinterpret_cast<uint4 *>(t64t + (size_t)slot0 * sizeof(Tri64) + ((size_t)k * ns + i) * 32u);
            dst[0] = src[2 * k]; dst[1] = src[2 * k + 1];
        }
    }
    if (t32) {
        Tri32 d;
        for (int k = 0; k < 3; ++k) {
            const float v0 = (float)v[k], v1 = (float)v[3 + k], v2 = (float)v[6 + k];
            d.v0[k] = v0; d.e1[k] = v1 - v0; d.e2[k] = v2 - v0;      // bvh.c:747-752, in fp32
        }
        d.prim = p; d.pad1 = 0; d.pad2 = 0;
        t32[slot] = d;
        // leaf-transposed copy: the pair's two triangles interleaved word by word (bvh_build.cpp): word 2f+h = field f of half h
        const uint32_t j = i >> 1, h = i & 1u;
        const float f[9] = {d.v0[0], d.v0[1], d.v0[2], d.e1[0], d.e1[1], d.e1[2], d.e2[0], d.e2[1], d.e2[2]};
        // rows 0 / 1: words 0..7 / 8..15 of the item (32 B per item); row 2: words 16..19 (16 B per item) at byte 64 m of the block
        char *blk = t32t + (size_t)slot0 * sizeof(Tri32);
        for (uint32_t q = 0; q < 8; ++q) {
            const uint32_t wi = 2u * q + h;
            *reinterpret_cast<float *>(blk + ((size_t)(wi >> 3) * m + j) * 32u + (wi & 7u) * 4u) = f[q];
        }
        *reinterpret_cast<float *>(blk + (size_t)m * 64u + (size_t)j * 16u + h * 4u) = f[8];       // word 16 + h
        *reinterpret_cast<uint32_t *>(blk + (size_t)m * 64u + (size_t)j * 16u + (2u + h) * 4u) = p;   // word 18 + h
    }
    if (i == ntris - 1u) {                                               // filler slots: zero-area triangles, prim = MISS
        for (uint32_t f = ntris; f < ns; ++f) {
            if (t64) {
                t64[slot0 + f].prim = 0xffffffffull;
                *reinterpret_cast<unsigned long long *>(t64t + (size_t)slot0 * sizeof(Tri64) + (size_t)f * 32u + 24u) = 0xffffffffull;
            }
            if (t32) {
                t32[slot0 + f].prim = 0xffffffffu;
                *reinterpret_cast<uint32_t *>(t32t + (size_t)slot0 * sizeof(Tri32) + (size_t)m * 64u + (size_t)(f >> 1) * 16u + (2u + (f & 1u)) * 4u) = 0xffffffffu;
            }
        }
        if (t32 && (ntris & 1u)) {                                       // unit edges for the masked half of the last pair (bvh_build.cpp)
            char *item = t32t + (size_t)slot0 * sizeof(Tri32) + (size_t)((ntris + 1u) / 2u - 1u) * 32u;
            *reinterpret_cast<float *>(item + 7u * 4u) = 1.0f;                                   // word 7  = B.e1.x (chunk 0)
            *reinterpret_cast<float *>(item + (size_t)m * 32u + 7u * 4u) = 1.0f;                 // word 15 = B.e2.y (chunk 1)
        }
    }
}

// tri_xyz: HOST [ntris][9].  Fills `out` like build_tree() except for out.tri (the triangles stay on the device: db);
// build_seconds = wall time incl. the triangle upload and the node download.
static int build_tree_device(const double *tri_xyz, uint64_t ntris64, HostTree &out, cudaStream_t st, DeviceBuild &db)
{
    GpuBuildScratch &s = db.s;
    const auto t0 = std::chrono::steady_clock::now();
    out = HostTree();
    out.ntris = ntris64;
    if (ntris64 == 0) { out.empty = true; return 0; }
    if (ntris64 >= (1ull << 31)) return fail("device builder: too many triangles");
    out.empty = false;
    const uint32_t n = (uint32_t)ntris64;
    const uint32_t max_nodes = 2u * n + 2u, max_level_inner = n / (uint32_t)(kGbLeaf + 1) + 2u;
    const uint32_t ntiles = (n + 1 + kScanTile - 1) / kScanTile;
    CUDA_OK(cudaMalloc((void **)&s.d_tri, (size_t)n * 9 * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&s.d_box[0], (size_t)n * sizeof(GBox)));
    CUDA_OK(cudaMalloc((void **)&s.d_box[1], (size_t)n * sizeof(GBox)));
    CUDA_OK(cudaMalloc((void **)&s.d_nodes, (size_t)max_nodes * sizeof(GNode)));
    CUDA_OK(cudaMalloc((void **)&s.d_hist, ((size_t)2 * max_level_inner + 8) * 2 * 3 * kGbBins * sizeof(uint32_t)));   // one histogram per node of a level, leaves included
    CUDA_OK(cudaMalloc((void **)&s.d_S, ((size_t)n + 1) * sizeof(uint32_t)));
    CUDA_OK(cudaMalloc((void **)&s.d_tiles, ((size_t)ntiles + 4) * sizeof(uint32_t)));
    CUDA_OK(cudaMalloc((void **)&s.d_counters, 4 * sizeof(uint32_t)));
    CUDA_OK(cudaMalloc((void **)&s.d_flags, (size_t)n));
    CUDA_OK(cudaMalloc((void **)&s.d_acc, ((size_t)2 * max_level_inner + 8) * 6 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemcpyAsync(s.d_tri, tri_xyz, (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice, st));

    const unsigned nb = (n + 255) / 256;
    {
        unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
        CUDA_OK(cudaMemcpyAsync(s.d_acc, init, sizeof(init), cudaMemcpyHostToDevice, st));
        gb_init_boxes<<<nb, 256, 0, st>>>(s.d_tri, n, s.d_box[0], s.d_acc);
        LAUNCHED();
        gb_root<<<1, 1, 0, st>>>(s.d_nodes, s.d_acc, n);
        LAUNCHED();
    }
    uint32_t first = 0, count = 1, inner = (n > (uint32_t)kGbLeaf) ? 1u : 0u;
    int cur = 0, levels = 0;
    const bool trace = getenv("B200_BUILD_TRACE") != nullptr;
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(st);
        fprintf(stderr, "[device build] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    lap("alloc+upload+boxes");
    while (inner > 0) {
        ++levels;
        if (count > max_level_inner * 2u + 8u) return fail("device builder: level overflow");
        const uint32_t next_first = first + count;
        uint32_t h_counters[2] = {next_first, 0u};
        CUDA_OK(cudaMemcpyAsync(s.d_counters, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemsetAsync(s.d_hist, 0, (size_t)count * 2 * 3 * kGbBins * sizeof(uint32_t), st));
        gb_bin<<<nb, 256, 0, st>>>(s.d_box[cur], n, s.d_nodes, first, s.d_hist);
        LAUNCHED();
        gb_split<<<(count + 127) / 128, 128, 0, st>>>(s.d_nodes, first, count, s.d_hist);
        LAUNCHED();
        gb_flags<<<ntiles, kScanBlock, 0, st>>>(s.d_box[cur], n, s.d_nodes, first, s.d_flags, s.d_tiles);
        LAUNCHED();
        scan_tile_offsets<<<1, kScanBlock, 0, st>>>(s.d_tiles, ntiles, s.d_tiles + ntiles);
        LAUNCHED();
        gb_prefix<<<ntiles, kScanBlock, 0, st>>>(s.d_flags, n, s.d_tiles, s.d_S);
        LAUNCHED();
        gb_scatter<<<nb, 256, 0, st>>>(s.d_box[cur], s.d_box[cur ^ 1], n, s.d_nodes, first, s.d_flags, s.d_S);
        LAUNCHED();
        gb_children<<<(count + 127) / 128, 128, 0, st>>>(s.d_nodes, first, count, s.d_S, s.d_counters, s.d_acc, next_first);
        LAUNCHED();
        gb_child_boxes<<<nb, 256, 0, st>>>(s.d_box[cur ^ 1], n, s.d_nodes, first, s.d_acc, next_first);
        LAUNCHED();
        const uint32_t next_count = 2u * inner;
        gb_finalize<<<(next_count + 127) / 128, 128, 0, st>>>(s.d_nodes, next_first, next_count, s.d_acc, s.d_counters);
        LAUNCHED();
        CUDA_OK(cudaMemcpyAsync(h_counters, s.d_counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        if (h_counters[0] != next_first + next_count) return fail("device builder: child count mismatch");
        if ((uint64_t)next_first + next_count > max_nodes) return fail("device builder: node overflow");
        first = next_first; count = next_count; inner = h_counters[1];
        cur ^= 1;
    }
    CUDA_OK(cudaGetLastError());
    lap("levels");
    const uint32_t total = first + count;
    std::vector<GNode> g(total);
    out.orig.resize(n);
    gb_extract_idx<<<nb, 256, 0, st>>>(s.d_box[cur], n, s.d_S);
    LAUNCHED();
    CUDA_OK(cudaMemcpyAsync(g.data(), s.d_nodes, (size_t)total * sizeof(GNode), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(out.orig.data(), s.d_S, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    lap("download");
    for (int k = 0; k < 3; ++k) { out.bmin[k] = g[0].lo[k]; out.bmax[k] = g[0].hi[k]; }
    out.nodes.reserve(total);
    db.canon_of.assign(total, 0u);
    gb_emit(g, 0, out.nodes, 0, out, db.canon_of);
    lap("host renumber");
    if (trace) fprintf(stderr, "[device build] %d levels, %u nodes\n", levels, total);
    db.cur = cur; db.n = n; db.total = total;
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

}  // namespace b200

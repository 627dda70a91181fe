// Host-side BVH construction for the B200 accelerator.
//
// Produces the SAME tree as lucille's ri_bvh_build() (src/render/bvh.c:276-379): same binned-SAH
// split decisions (64 bins, float-rounded cost), same partition order (left stable, right reversed),
// same child boxes (union + 1e-14 margin), same leaf contents and order -- because triangle order
// inside leaves and the visiting order are part of the tie-breaking contract.
// The data flow is different: one flat box array partitioned between two ping-pong buffers
// (no per-level memcpy + three extra sweeps), subtrees built by a pool of host threads, and the
// result emitted directly as flat records for the GPU.
#pragma once

#include <cstdint>
#include <vector>

namespace b200 {

struct CanonNode {              // canonical DFS-preorder form (== ri_b200_node_t)
    int32_t is_leaf;
    int32_t axis;
    int64_t child0, child1;
    int64_t tri_start, ntris;
    double  lbox[6];
    double  rbox[6];
};

struct HostTree {
    bool     empty = true;
    uint64_t ntris = 0;
    double   bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    int      max_depth = 0;
    int64_t  ninner = 0, nleaf = 0;
    std::vector<CanonNode> nodes;   // DFS preorder
    std::vector<double>    tri;     // post-build order, [ntris][9] = v0 v1 v2
    std::vector<uint32_t>  orig;    // post-build position -> input triangle
    double   build_seconds = 0.0;
};

// tri_xyz: [ntris][3][3] doubles. nthreads <= 0: use the hardware concurrency.
void build_tree(const double *tri_xyz, uint64_t ntris, HostTree &out, int nthreads = 0);

// ---- flat device records --------------------------------------------------------------------

constexpr uint32_t kLeafFlag  = 0x80000000u;
constexpr uint32_t kDoneWord  = 0x7fffffffu;
constexpr uint32_t kLeafShift = 27;                  // word = flag | (ntris-1)<<27 | first triangle SLOT (always even)
constexpr uint64_t kMaxTris   = (1ull << 27) - 4;              // necessary only: flatten_tree checks the real slot count (sum of round_up(ntris, 4))

struct Node32 {                 // 64 B, read as 4 x LDG.128
    float    x[4];              // lo0.x hi0.x lo1.x hi1.x   (slot 0 = left child, slot 1 = right child)
    float    y[4];
    float    z[4];
    uint32_t c0, c1, axis, pad; // child words; axis = ri_qbvh_node_t.axis0
};
struct Node64 {                 // 128 B
    double   x[4];
    double   y[4];
    double   z[4];
    uint32_t c0, c1, axis, pad;
    uint32_t pad2[4];
};
// Triangle SLOTS.  Every leaf starts at a multiple-of-four slot and owns round_up(ntris, 4) slots, so a pair of fp32 slots is
// one 32-byte-aligned 96-byte run = 3 x LDG.256; the leftover slots of a leaf hold zero-area triangles (never hit:
// |det| <= 1e-14 rejects them, bvh.c:759-763).  `prim` = position in the post-build triangle order (the hit id).
struct Tri32 { float v0[3]; uint32_t prim; float e1[3]; uint32_t pad1; float e2[3]; uint32_t pad2; };          // 48 B
struct Tri64 { double v0[3]; uint64_t prim; double e1[3]; uint64_t pad1; double e2[3]; uint64_t pad2; };       // 96 B = 3 x LDG.256

struct FlatTree {
    bool     overflow = false;       // more triangle slots than the 27-bit field of a leaf word can address
    uint32_t root_word = kDoneWord;  // inner index, leaf word, or kDoneWord for an empty scene
    uint32_t ninner = 0;
    uint32_t top_count = 0;          // the first top_count inner nodes are in BFS order (SMEM-resident cluster)
    std::vector<Node32> nodes32;
    std::vector<Node64> nodes64;
    std::vector<Tri32>  tris32;      // nslots entries
    std::vector<Tri64>  tris64;
    // the same slots, 32-byte chunks TRANSPOSED inside each leaf (pooled occlusion kernel, pool.cuh): a leaf at slot0
    // whose rows hold m items (fp32: m = slots/2 pairs, fp64: m = slots; slots = round_up(ntris, 4)) keeps chunk k of item j at byte
    // slot0 * sizeof(slot) + (k * m + j) * 32, so lanes testing consecutive items read consecutive 32-byte chunks
    std::vector<Tri32>  tris32t;
    std::vector<Tri64>  tris64t;
    uint64_t nslots = 0;
    std::vector<uint32_t> slot_of_prim;   // post-build triangle position -> slot
    std::vector<uint32_t> leaf_slot;      // canonical node index -> first slot (leaves only)
    float  smin32[3], smax32[3];
};

// Lays the inner nodes out as [BFS top cluster | DFS-preorder remainder] and folds leaves into
// their parent's child words.  fp32 boxes are rounded OUTWARD (min down, max up).
// fill_tris = false: node records, leaf words and leaf_slot only -- the device builder writes the triangle slots itself.
void flatten_tree(const HostTree &t, uint32_t top_nodes, bool want32, bool want64, FlatTree &out, bool fill_tris = true);

}  // namespace b200

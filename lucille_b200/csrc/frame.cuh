// On-device ambient-occlusion frame: the wavefront re-expression of lucille's pixel loop
// (render.c:715-823 subsample, 1107-1146 render_bucket, 919-979 bucket_write) and of
// ri_transport_ambientocclusion / calculate_occlusion (ambientocclusion.c:332-415, 42-151).
//
//   K3  primary_kernel   one lane per pixel sub-sample, in the reference's consumption order
//                        (spiral buckets -> row-major pixels -> ys -> xs); camera ray in double,
//                        closest-hit traversal
//   scan_*               exclusive prefix sum of the hit flags: the reference draws 2*ntheta*nphi MT
//                        numbers per *hit* sample and none per miss, so a sample's offset in the
//                        stream is 2*N*(hits before it)
//   compact_kernel       hit samples -> dense work list with the shading frame (P + 1e-6*Ns, basis)
//   mt_kernel            MT19937 stream of randomMT2() (random.c:98-112, 211-247), one CTA
//   K4  ao_kernel        one lane per occlusion ray, any-hit traversal, warp-ballot accumulation
//   resolve_kernel       box average of the sub-samples, float RGB at row H-1-y
//
// Included at the end of accel.cu (same translation unit, -fmad=false).
#pragma once

namespace b200 {

struct FrameDev {
    double c2w[16];
    double flength_signed;      // sign * flength, camera.c:269,275
    double w, h;
    int    width, height, xsamples, ysamples, ntheta, nphi, spp, nao;
    int    rng_mode;
    uint32_t seed;
    double ao_eps;              // shading-point offset along Ns: 1e-6 (ambientocclusion.c:56), 1e-5 for the sun-sky gather (:222)
    // rng_mode 0 on more than one rank: the frame's hit samples are numbered over ALL ranks' buckets in spiral order, so that every
    // gather ray draws the same two words of the single MT19937 stream as in the reference.  delta = (hit samples of the whole
    // frame before bucket b) - (hit samples of this rank before bucket b); both null when world == 1.
    const uint32_t  *pix_bucket = nullptr;       // [pixel of this rank] -> index among this rank's buckets
    const long long *bucket_delta = nullptr;     // [bucket of this rank]
};

// camera.c:248-352 (perspective) + render.c:770-781 normalise -------------------------------------
__device__ __forceinline__ void camera_ray(const FrameDev &F, int px, int py, double jx, double jy, double org[3], double dir[3])
{
    const double x = (double)(px + jx), y = (double)(py + jy);
    double v[4];
    v[0] = (2.0 * x - F.w) / F.w;
    v[1] = (2.0 * y - F.h) / F.h;
    v[2] = F.flength_signed;
    v[3] = 1.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double dp = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) dp += v[i] * F.c2w[4 * i + j];    // ri_vector_transform, vector.h:182-210
        const double pos = F.c2w[12 + j];
        org[j] = pos;
        dir[j] = dp - pos;
    }
    normalize3(dir);
}

template <typename Real>
__global__ void __launch_bounds__(kBlock)
primary_kernel(const SceneView<Real> S, const FrameDev F, const uint32_t *__restrict__ pixels, const double *__restrict__ jitter,
               const uint64_t nsamples, Real *__restrict__ hit_t, uint32_t *__restrict__ hit_prim, Real *__restrict__ hit_uv)
{
    extern __shared__ uint32_t s_stack[];
    const uint64_t s = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (s >= nsamples) return;
    const uint64_t p = s / (uint64_t)F.spp;
    const int sub = (int)(s - p * (uint64_t)F.spp);
    const uint32_t pix = pixels[p];
    double org[3], dir[3];
    camera_ray(F, (int)(pix & 0xffffu), (int)(pix >> 16), jitter[2 * sub], jitter[2 * sub + 1], org, dir);
    Real o[3] = {(Real)org[0], (Real)org[1], (Real)org[2]}, d[3] = {(Real)dir[0], (Real)dir[1], (Real)dir[2]};
    Real t, u, v;
    uint32_t prim;
    const bool hit = trace_ray<Real, false, false>(S, o, d, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr);
    hit_t[s] = hit ? t : Prec<Real>::inf();
    hit_prim[s] = hit ? prim : 0xffffffffu;
    if (hit_uv) { hit_uv[2 * s] = u; hit_uv[2 * s + 1] = v; }    // only needed when the scene carries vertex normals
}

// ---- exclusive scan of hit flags (three small kernels; 2048 flags per block) -------------------
constexpr int kScanBlock = 256, kScanItems = 8, kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t *total)
{
    __shared__ uint32_t warp_sums[kScanBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = x;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < kScanBlock / 32) ? warp_sums[lane] : 0;
        for (int o = 1; o < kScanBlock / 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        if (lane < kScanBlock / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t base = warp ? warp_sums[warp - 1] : 0;
    if (total) *total = warp_sums[kScanBlock / 32 - 1];
    __syncthreads();
    return base + incl - x;
}

__global__ void __launch_bounds__(kScanBlock)
scan_tile_sums(const uint32_t *__restrict__ hit_prim, uint64_t n, uint32_t *__restrict__ tile_sums)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t c = 0;
    for (int k = 0; k < kScanItems; ++k) { const uint64_t i = base + k; if (i < n && hit_prim[i] != 0xffffffffu) ++c; }
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock)
scan_tile_offsets(uint32_t *__restrict__ tile_sums, uint32_t ntiles, uint32_t *__restrict__ grand_total)
{
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += kScanBlock) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = (i < ntiles) ? tile_sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(x, &total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

template <typename Real> struct StateMath;      // hit state in the accelerator's precision

template <> struct StateMath<double> {
    static __device__ __forceinline__ void frame(const SceneView<double> &S, const double org[3], const double dir[3], double t,
                                                 double bu, double bv, uint32_t prim, const double eps, double rec[12])
    {
        ri_b200_state_f64 s;
        state_from_hit(S.tris, S.slot_of_prim, org, dir, t, prim, s, S.normals, bu, bv);                                                 // eps: ambientocclusion.c:56,73-75
        double b0[3], b1[3];
        ortho_basis(b0, b1, s.Ns);                                      // ambientocclusion.c:65
        for (int k = 0; k < 3; ++k) {
            double o = s.P[k];
            o += s.Ns[k] * eps;
            rec[k] = o; rec[3 + k] = b0[k]; rec[6 + k] = b1[k]; rec[9 + k] = s.Ns[k];
        }
    }
};

template <> struct StateMath<float> {
    static __device__ __forceinline__ void nrm(float d[3])
    {
        const float n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        if (n2 > 1.0e-17f) { const float r = 1.0f / sqrtf(n2); d[0] *= r; d[1] *= r; d[2] *= r; }
    }
    static __device__ __forceinline__ void crs(float d[3], const float a[3], const float b[3])
    {
        d[0] = a[1] * b[2] - a[2] * b[1]; d[1] = a[2] * b[0] - a[0] * b[2]; d[2] = a[0] * b[1] - a[1] * b[0];
    }
    static __device__ __forceinline__ void frame(const SceneView<float> &S, const float org[3], const float dir[3], float t,
                                                 float bu, float bv, uint32_t prim, const double eps, float rec[12])
    {
        TriRegs<float> tr;
        load_tri(S.tris + S.slot_of_prim[prim], tr);
        float n[3], b0[3], b1[3], e[3] = {0.f, 0.f, 0.f};
        crs(n, tr.e1, tr.e2);
        nrm(n);
        if (S.normals) {                                                // interpolated shading normal (geometric.c:40-62)
            float vn[9];
            bool has_n = false;
            for (int k = 0; k < 9; ++k) { vn[k] = S.normals[9 * (size_t)prim + k]; has_n = has_n || (vn[k] != 0.0f); }
            if (has_n) {
                const float w0 = 1.0f - bu - bv;
                for (int k = 0; k < 3; ++k) n[k] = (vn[k] * w0 + vn[3 + k] * bu) + vn[6 + k] * bv;
            }
        }
        int i;
        for (i = 0; i < 3; ++i) if (n[i] < 0.6f && n[i] > -0.6f) break;
        if (i >= 3) i = 0;
        e[i] = 1.0f;
        crs(b0, e, n); nrm(b0);
        crs(b1, n, b0); nrm(b1);
        for (int k = 0; k < 3; ++k) {
            float o = org[k] + dir[k] * t;
            o += n[k] * (float)eps;
            rec[k] = o; rec[3 + k] = b0[k]; rec[6 + k] = b1[k]; rec[9 + k] = n[k];
        }
    }
};

// ---- material texture of the AO transport (ambientocclusion.c:393-401): radiance *= ri_texture_fetch(texture, st) per channel ----
struct TexDev {
    const float   *data;      // [h][w][4] floats (ri_texture_t.data, USE_ZORDER 0), NULL: no texture
    int            width, height;
    const double  *st;        // [prim][6] corner texture coordinates (may be NULL)
    const uint8_t *flags;     // [prim]: 2 = has st, 8 = the triangle's geom carries the texture (may be NULL: no triangle does)
    double        *texcol;    // out: [rank][3]; (1,1,1) where the hit triangle is not textured
};

// render/texture.c:86-236: wrap by floor, clamp, bilinear over the 2x2 texels at (u (w-1), v (h-1)), ZERO texels beyond the last
// row / column
__device__ __forceinline__ void texture_fetch_dev(const TexDev &T, double u, double v, double out[3])
{
    const double sx = floor(u), sy = floor(v);
    u = u - sx; v = v - sy;
    if (u < 0.0) u = 0.0;
    if (u >= 1.0) u = 1.0;
    if (v < 0.0) v = 0.0;
    if (v >= 1.0) v = 1.0;
    const double px = u * (double)(T.width - 1), py = v * (double)(T.height - 1);
    const int x = (int)px, y = (int)py;
    const double dx = px - (double)x, dy = py - (double)y;
    const double w0 = (1.0 - dx) * (1.0 - dy), w1 = (1.0 - dx) * dy, w2 = dx * (1.0 - dy), w3 = dx * dy;
    const bool my = y < T.height - 1, mx = x < T.width - 1;
    const float *t0 = T.data + 4 * ((size_t)y * T.width + x);
    const float *t1 = t0 + 4 * (size_t)T.width, *t2 = t0 + 4, *t3 = t1 + 4;
    for (int i = 0; i < 3; ++i) {
        const double a = (double)t0[i];
        const double b = my ? (double)t1[i] : 0.0;                      // texel (x, y+1)
        const double c = mx ? (double)t2[i] : 0.0;                      // texel (x+1, y)
        const double d = (my && mx) ? (double)t3[i] : 0.0;              // texel (x+1, y+1)
        out[i] = w0 * a + w1 * b + w2 * c + w3 * d;
    }
}

template <typename Real>
__global__ void __launch_bounds__(kScanBlock)
compact_kernel(const SceneView<Real> S, const FrameDev F, const uint32_t *__restrict__ pixels, const double *__restrict__ jitter,
               const uint64_t nsamples, const Real *__restrict__ hit_t, const uint32_t *__restrict__ hit_prim,
               const Real *__restrict__ hit_uv, const uint32_t *__restrict__ tile_offsets, uint32_t *__restrict__ sample_rank, uint32_t *__restrict__ rank_sample,
               Real *__restrict__ records, const TexDev tex)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t flags = 0, c = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const uint64_t i = base + k;
        if (i < nsamples && hit_prim[i] != 0xffffffffu) { flags |= 1u << k; ++c; }
    }
    uint32_t rank = tile_offsets[blockIdx.x] + block_exclusive_scan(c, nullptr);
    for (int k = 0; k < kScanItems; ++k) {
        const uint64_t s = base + k;
        if (s >= nsamples) break;
        if (!(flags & (1u << k))) { sample_rank[s] = 0xffffffffu; continue; }
        sample_rank[s] = rank;
        rank_sample[rank] = (uint32_t)s;
        // regenerate the eye ray (deterministic) and build the shading frame
        const uint64_t p = s / (uint64_t)F.spp;
        const int sub = (int)(s - p * (uint64_t)F.spp);
        const uint32_t pix = pixels[p];
        double org[3], dir[3];
        camera_ray(F, (int)(pix & 0xffffu), (int)(pix >> 16), jitter[2 * sub], jitter[2 * sub + 1], org, dir);
        Real o[3] = {(Real)org[0], (Real)org[1], (Real)org[2]}, d[3] = {(Real)dir[0], (Real)dir[1], (Real)dir[2]};
        Real rec[12];
        StateMath<Real>::frame(S, o, d, hit_t[s], hit_uv ? hit_uv[2 * s] : Real(0), hit_uv ? hit_uv[2 * s + 1] : Real(0), hit_prim[s], F.ao_eps, rec);
        Real *dst = records + 12 * (uint64_t)rank;
#pragma unroll
        for (int q = 0; q < 12; ++q) dst[q] = rec[q];
        if (tex.data) {                                                  // ambientocclusion.c:393-401
            double col[3] = {1.0, 1.0, 1.0};
            const uint32_t prim = hit_prim[s];
            const uint8_t fl = tex.flags ? tex.flags[prim] : 0;
            if (fl & 8) {
                double s0 = 0.0, s1 = 0.0;
                if ((fl & 2) && tex.st) {                                // lerp_uv, intersection_state.c:266-280
                    const double u = (double)hit_uv[2 * s], v = (double)hit_uv[2 * s + 1];
                    const double *c = tex.st + 6 * (size_t)prim;
                    s0 = (1 - u - v) * c[0] + u * c[2] + v * c[4];
                    s1 = (1 - u - v) * c[1] + u * c[3] + v * c[5];
                }
                texture_fetch_dev(tex, s0, s1, col);
            }
            double *o = tex.texcol + 3 * (uint64_t)rank;
            o[0] = col[0]; o[1] = col[1]; o[2] = col[2];
        }
        ++rank;
    }
}

// ---- MT19937, the generator behind randomMT2() (random.c:98-112 seeding, 211-247 regeneration + tempering) ----
// One CTA owns the 624-word state in shared memory (double buffered).  new[k] depends on old[k], old[k+1] and on
// the word 227 positions back in the NEW state, so a block regenerates in three dependent phases of <=227 lanes.
constexpr int kMtN = 624, kMtM = 397;
}  // namespace b200
#include "mt_jump_table.h"
namespace b200 {

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t far)
{
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// Segment generator: CTA `blockIdx.x` starts from the window states[blockIdx.x] (624 words: x_m .. x_{m+623}, only the top
// bit of word 0 is significant) and writes its blocks of the stream.  states == NULL: one CTA, seeded here (seedMT2).
// Segment k covers stream words [k*seg_blocks*624, ...); the last segment may be shorter (total_blocks).
__global__ void __launch_bounds__(256)
mt_kernel(uint32_t seed, const uint32_t *__restrict__ states, uint64_t seg_blocks, uint64_t total_blocks, uint32_t *__restrict__ out)
{
    __shared__ uint32_t st[2][kMtN];
    const int k = threadIdx.x;
    if (!states) {
        if (k == 0) {
            uint32_t x = seed;
            st[0][0] = x;
            for (int i = 1; i < kMtN; ++i) { x = 69069u * x; st[0][i] = x; }      // seedMT2, random.c:98-112
        }
    } else {
        for (int i = k; i < kMtN; i += 256) st[0][i] = states[(uint64_t)blockIdx.x * kMtN + i];
    }
    __syncthreads();
    const uint64_t first_block = (uint64_t)blockIdx.x * seg_blocks;
    if (first_block >= total_blocks) return;
    const uint64_t nblocks = (total_blocks - first_block) < seg_blocks ? (total_blocks - first_block) : seg_blocks;
    out += first_block * kMtN;
    int cur = 0;
    for (uint64_t b = 0; b < nblocks; ++b) {
        uint32_t *o = st[cur], *n = st[cur ^ 1];
        uint32_t *dst = out + b * kMtN;
        if (k < kMtN - kMtM) {                                   // phase 1: k in [0,227)
            const uint32_t v = mt_twist(o[k], o[k + 1], o[k + kMtM]);
            n[k] = v; dst[k] = mt_temper(v);
        }
        __syncthreads();
        if (k < kMtN - kMtM) {                                   // phase 2: k in [227,454)
            const int i = k + (kMtN - kMtM);
            const uint32_t v = mt_twist(o[i], o[i + 1], n[i - (kMtN - kMtM)]);
            n[i] = v; dst[i] = mt_temper(v);
        }
        __syncthreads();
        {                                                        // phase 3: k in [454,624)
            const int i = k + 2 * (kMtN - kMtM);
            if (i < kMtN - 1) {
                const uint32_t v = mt_twist(o[i], o[i + 1], n[i - (kMtN - kMtM)]);
                n[i] = v; dst[i] = mt_temper(v);
            } else if (i == kMtN - 1) {
                const uint32_t v = mt_twist(o[kMtN - 1], n[0], n[kMtM - 1]);
                n[i] = v; dst[i] = mt_temper(v);
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

// Jump ahead: dst = g(T) src, g = poly (x^J mod phi, tools/gen_mt_jump.py), by Horner's rule 32 coefficients at a time:
//     r <- T^32 r  xor  sum_{j<32} g[32w+j] * T^j src            for w = 623 .. 0
// T^j src is the source window slid by j words, so the sum is a XOR of shifted reads of the source extended by 31 words;
// T^32 r regenerates 32 words at once (they depend only on old words).  r lives in shared memory as a ring with head h.
// CTA b computes states[b + stride] from states[b]  (b + stride < nstates): one level of the doubling tree.
__global__ void __launch_bounds__(256)
mt_jump_kernel(const uint32_t *__restrict__ poly, uint32_t *__restrict__ states, uint32_t stride, uint32_t nstates)
{
    __shared__ uint32_t r[kMtN], ext[kMtN + 32], g[kMtN];
    const uint32_t src = blockIdx.x, dst = blockIdx.x + stride;
    if (dst >= nstates) return;
    const int t = threadIdx.x;
    for (int i = t; i < kMtN; i += 256) { ext[i] = states[(uint64_t)src * kMtN + i]; g[i] = poly[i]; r[i] = 0u; }
    __syncthreads();
    if (t < 31) ext[kMtN + t] = mt_twist(ext[t], ext[t + 1], ext[t + kMtM]);       // x_{624+t}: inputs are all old words
    __syncthreads();
    uint32_t h = 0;
    for (int w = kMtN - 1; w >= 0; --w) {
        if (t < 32) {                                   // r <- T^32 r
            const uint32_t a = r[(h + t) % kMtN], b = r[(h + t + 1) % kMtN], far = r[(h + t + kMtM) % kMtN];
            __syncwarp();
            r[(h + t) % kMtN] = mt_twist(a, b, far);
        }
        h = (h + 32) % kMtN;
        __syncthreads();
        const uint32_t bits = g[w];
        if (bits) {
            for (int p = t; p < kMtN; p += 256) {
                uint32_t acc = 0, m = bits;
                while (m) { const int j = __ffs(m) - 1; m &= m - 1; acc ^= ext[p + j]; }
                r[(h + p) % kMtN] ^= acc;
            }
        }
        __syncthreads();
    }
    for (int i = t; i < kMtN; i += 256) states[(uint64_t)dst * kMtN + i] = r[(h + i) % kMtN];
}

__device__ __forceinline__ uint64_t splitmix64_dev(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// ---- occlusion-ray direction k = j*ntheta + i of hit sample `rank`: ambientocclusion.c:83-117 (stratified cosine sampling about
// the shading frame in `rec`), uniforms from the MT19937 stream in the reference's consumption order or from the counter RNG
template <typename Real>
__device__ __forceinline__ void ao_direction(const FrameDev &F, const uint32_t rank, const uint32_t k, const Real *__restrict__ rec,
                                             const uint32_t *__restrict__ rank_sample, const uint32_t *__restrict__ pixels,
                                             const uint32_t *__restrict__ mt_stream, Real dir[3])
{
    const uint32_t N = (uint32_t)F.nao;
    const uint32_t j = k / (uint32_t)F.ntheta, i = k - j * (uint32_t)F.ntheta;
    double r0, r1;
    if (F.rng_mode == 0) {
        uint64_t grank = rank;
        if (F.bucket_delta) grank = (uint64_t)((long long)rank + F.bucket_delta[F.pix_bucket[rank_sample[rank] / (uint32_t)F.spp]]);
        const uint64_t base = (uint64_t)2 * N * grank + 2 * k;              // draw order z0 then z1, ambientocclusion.c:91-92
        r0 = (double)mt_stream[base] * 2.3283064365386963e-10;               // random.c:244
        r1 = (double)mt_stream[base + 1] * 2.3283064365386963e-10;
    } else {
        const uint32_t s = rank_sample[rank];
        const uint64_t p = s / (uint32_t)F.spp;
        const uint32_t sub = s - (uint32_t)p * (uint32_t)F.spp;
        const uint32_t pix = pixels[p];
        const uint64_t sid = ((uint64_t)(pix >> 16) * (uint64_t)F.width + (pix & 0xffffu)) * (uint64_t)F.spp + sub;
        const uint64_t idx = (sid * N + k) * 2;
        r0 = (double)(splitmix64_dev((uint64_t)F.seed + idx * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
        r1 = (double)(splitmix64_dev((uint64_t)F.seed + (idx + 1) * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
    }
    if (sizeof(Real) == 8) {                                                 // ambientocclusion.c:91-117, double
        const double z0 = ((double)i + r0) / (double)F.ntheta;
        const double z1 = ((double)j + r1) / (double)F.nphi;
        const double ct = sqrt(z0);
        const double phi = 2.0 * 3.14159265358979323846 * z1;
        const double lx = cos(phi) * ct, ly = sin(phi) * ct, lz = sqrt(1.0 - ct * ct);
#pragma unroll
        for (int q = 0; q < 3; ++q)
            dir[q] = (Real)(lx * (double)rec[3 + q] + ly * (double)rec[6 + q] + lz * (double)rec[9 + q]);
    } else {
        const float z0 = ((float)i + (float)r0) / (float)F.ntheta;
        const float z1 = ((float)j + (float)r1) / (float)F.nphi;
        const float ct = sqrtf(z0);
        const float phi = 6.28318530717958647692f * z1;
        float sp, cp;
        sincosf(phi, &sp, &cp);
        const float lx = cp * ct, ly = sp * ct, lz = sqrtf(1.0f - ct * ct);
#pragma unroll
        for (int q = 0; q < 3; ++q)
            dir[q] = (Real)(lx * (float)rec[3 + q] + ly * (float)rec[6 + q] + lz * (float)rec[9 + q]);
    }
}

// ---- K4: occlusion rays ----------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(kBlock)
ao_kernel(const SceneView<Real> S, const FrameDev F, const uint64_t nrays, const uint32_t rank0,
          const Real *__restrict__ records, const uint32_t *__restrict__ rank_sample, const uint32_t *__restrict__ pixels,
          const uint32_t *__restrict__ mt_stream, uint32_t *__restrict__ occ,
          Real *__restrict__ dump_rays, const uint64_t dump_count)
{
    extern __shared__ uint32_t s_stack[];
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool active = gid < nrays;
    bool hit = false;
    uint32_t rank = 0;
    if (active) {
        const uint32_t N = (uint32_t)F.nao;
        rank = rank0 + (uint32_t)(gid / N);
        const uint32_t k = (uint32_t)(gid - (uint64_t)(rank - rank0) * N);
        const Real *rec = records + 12 * (uint64_t)rank;
        Real org[3] = {rec[0], rec[1], rec[2]}, dir[3];
        ao_direction<Real>(F, rank, k, rec, rank_sample, pixels, mt_stream, dir);
        if (dump_rays && gid < dump_count) {
            const int st = sizeof(Real) == 8 ? 6 : 8, off = sizeof(Real) == 8 ? 3 : 4;
            Real *d = dump_rays + gid * st;
            d[0] = org[0]; d[1] = org[1]; d[2] = org[2];
            d[off] = dir[0]; d[off + 1] = dir[1]; d[off + 2] = dir[2];
            if (sizeof(Real) == 4) { d[3] = (Real)0; d[7] = (Real)1.0e38f; }
        }
        Real t, u, v;
        uint32_t prim;
        hit = trace_ray<Real, true, false>(S, org, dir, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr);
    }
    if ((F.nao & 31) == 0) {                                            // a warp never straddles two samples
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&occ[rank], (uint32_t)__popc(m));
    } else if (hit) {
        atomicAdd(&occ[rank], 1u);
    }
}

// ---- K4 (wavefront form): write the occlusion rays of hit samples [rank0, rank0 + nrays/N) as a ray batch; the batch is
// then traced by the persistent any-hit kernel, which accumulates straight into the per-sample counters.
template <typename Real>
__global__ void __launch_bounds__(kBlock)
ao_gen_kernel(const FrameDev F, const uint64_t nrays, const uint32_t rank0, const Real *__restrict__ records,
              const uint32_t *__restrict__ rank_sample, const uint32_t *__restrict__ pixels,
              const uint32_t *__restrict__ mt_stream, Real *__restrict__ rays_out)
{
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (gid >= nrays) return;
    const uint32_t N = (uint32_t)F.nao;
    const uint32_t rank = rank0 + (uint32_t)(gid / N);
    const uint32_t k = (uint32_t)(gid - (uint64_t)(rank - rank0) * N);
    const Real *rec = records + 12 * (uint64_t)rank;
    Real dir[3];
    ao_direction<Real>(F, rank, k, rec, rank_sample, pixels, mt_stream, dir);
    if (sizeof(Real) == 8) {
        double2 *o = reinterpret_cast<double2 *>(rays_out) + 3 * gid;
        o[0] = make_double2((double)rec[0], (double)rec[1]);
        o[1] = make_double2((double)rec[2], (double)dir[0]);
        o[2] = make_double2((double)dir[1], (double)dir[2]);
    } else {
        float4 *o = reinterpret_cast<float4 *>(rays_out) + 2 * gid;
        o[0] = make_float4((float)rec[0], (float)rec[1], (float)rec[2], 0.0f);
        o[1] = make_float4((float)dir[0], (float)dir[1], (float)dir[2], 1.0e38f);
    }
}

// ---- resolve with a material texture: radiance[k] = Lo * texcol[k] (ambientocclusion.c:398-400), three channels
__global__ void resolve_tex_kernel(const FrameDev F, const uint32_t *__restrict__ pixels, uint64_t npixels,
                                   const uint32_t *__restrict__ sample_rank, const uint32_t *__restrict__ occ, const double *__restrict__ texcol,
                                   float *__restrict__ rgb, const int packed)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npixels) return;
    const uint32_t pix = pixels[p];
    const int x = (int)(pix & 0xffffu), y = (int)(pix >> 16);
    const double ns = (double)F.nao;
    double accum[3] = {0.0, 0.0, 0.0};
    for (int sub = 0; sub < F.spp; ++sub) {
        const uint32_t r = sample_rank[p * (uint64_t)F.spp + sub];
        for (int q = 0; q < 3; ++q) {
            double rad = 0.0;
            if (r != 0xffffffffu) { rad = 1.0 * (ns - (double)occ[r]) / ns; rad *= texcol[3 * (uint64_t)r + q]; }
            accum[q] = accum[q] + rad;
        }
    }
    float *dst = packed ? rgb + 3 * p : rgb + 3 * ((uint64_t)(F.height - y - 1) * F.width + x);
    for (int q = 0; q < 3; ++q) dst[q] = (float)(accum[q] * (1.0 / (double)(F.xsamples * F.ysamples)));
}

// ---- resolve: Lo = (N - occluded)/N per hit sample, mean over sub-samples, float at row H-1-y ---------
__global__ void resolve_kernel(const FrameDev F, const uint32_t *__restrict__ pixels, uint64_t npixels,
                               const uint32_t *__restrict__ sample_rank, const uint32_t *__restrict__ occ, float *__restrict__ rgb,
                               const int packed)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npixels) return;
    const uint32_t pix = pixels[p];
    const int x = (int)(pix & 0xffffu), y = (int)(pix >> 16);
    const double ns = (double)F.nao;
    double accum = 0.0;
    for (int sub = 0; sub < F.spp; ++sub) {
        const uint32_t r = sample_rank[p * (uint64_t)F.spp + sub];
        double rad = 0.0;
        if (r != 0xffffffffu) rad = 1.0 * (ns - (double)occ[r]) / ns;    // ambientocclusion.c:143-147
        accum = accum + rad;                                             // render.c:805
    }
    const double px = accum * (1.0 / (double)(F.xsamples * F.ysamples)); // render.c:820
    // full framebuffer: row H-1-y (render.c:962-964); packed: this rank's pixels in visiting order (the NCCL send buffer)
    float *dst = packed ? rgb + 3 * p : rgb + 3 * ((uint64_t)(F.height - y - 1) * F.width + x);
    const float f = (float)px;
    dst[0] = f; dst[1] = f; dst[2] = f;
}

// hit samples per bucket of this rank (one CTA per bucket; samples of a bucket are contiguous in visiting order)
__global__ void __launch_bounds__(256)
bucket_hits_kernel(const uint32_t *__restrict__ hit_prim, const uint32_t *__restrict__ bucket_first_pixel, uint32_t spp,
                   uint32_t *__restrict__ bucket_hits)
{
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    const uint64_t lo = (uint64_t)bucket_first_pixel[blockIdx.x] * spp, hi = (uint64_t)bucket_first_pixel[blockIdx.x + 1] * spp;
    uint32_t c = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += 256) c += hit_prim[i] != 0xffffffffu;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
    __syncthreads();
    if (threadIdx.x == 0) bucket_hits[blockIdx.x] = total;
}

// ---- host helpers ---------------------------------------------------------------------------------
// spiral.c:97-140 NthBucketSpiral
}  // namespace b200
#include "sunsky.cuh"
namespace b200 {

static void nth_bucket_spiral(int n, int nxb, int nyb, int *bx, int *by)
{
    const int minnb = nxb < nyb ? nxb : nyb;
    const int center = (minnb - 1) / 2;
    int nx = nxb, ny = nyb;
    while (n < nx * ny) { --nx; --ny; }
    const int nxny = nx * ny, m = nx < ny ? nx : ny;
    int x, y;
    if (m % 2 == 1) {
        if (n <= nxny + ny) { x = nx - m / 2; y = -m / 2 + n - nxny; }
        else { x = nx - m / 2 - (n - (nxny + ny)); y = ny - m / 2; }
    } else {
        if (n <= nxny + ny) { x = -m / 2; y = ny - m / 2 - (n - nxny); }
        else { x = -m / 2 + (n - (nxny + ny)); y = -m / 2; }
    }
    *bx = x + center; *by = y + center;
}

// pixel visiting order of the reference for this rank's buckets: render.c:582-710 + 1131-1146
static void pixel_order(const ri_b200_frame_t &f, std::vector<uint32_t> &pix, std::vector<uint32_t> *bucket_first = nullptr)
{
    const int bs = f.bucket_size > 0 ? f.bucket_size : 32;
    const int nxb = (f.width + bs - 1) / bs, nyb = (f.height + bs - 1) / bs;
    const int world = f.world > 0 ? f.world : 1;
    pix.clear();
    if (bucket_first) bucket_first->clear();
    for (int n = 0; n < nxb * nyb; ++n) {
        if (n % world != f.rank) continue;
        if (bucket_first) bucket_first->push_back((uint32_t)pix.size());
        int bx, by;
        nth_bucket_spiral(n, nxb, nyb, &bx, &by);
        const int x0 = bx * bs, y0 = by * bs;
        const int w = (x0 + bs <= f.width) ? bs : f.width - x0;
        const int h = (y0 + bs <= f.height) ? bs : f.height - y0;
        for (int v = y0; v < y0 + h; ++v)
            for (int u = x0; u < x0 + w; ++u) pix.push_back((uint32_t)u | ((uint32_t)v << 16));
    }
    if (bucket_first) bucket_first->push_back((uint32_t)pix.size());
}

// render.c:830-917 sample_subpixel / init_sigma (periodx masks both indices -- reproduced)
static void jitter_table(int xsamples, int ysamples, std::vector<double> &jit)
{
    auto sigma = [](unsigned period, std::vector<unsigned> &s) {
        s.resize(period);
        for (unsigned i = 0; i < period; ++i) {
            unsigned digit = period, inverse = 0;
            for (unsigned bits = i; bits; bits >>= 1) { digit >>= 1; if (bits & 1) inverse += digit; }
            s[i] = inverse;
        }
    };
    std::vector<unsigned> sx, sy;
    sigma((unsigned)xsamples, sx);
    sigma((unsigned)ysamples, sy);
    jit.resize((size_t)2 * xsamples * ysamples);
    for (int ys = 0; ys < ysamples; ++ys)
        for (int xs = 0; xs < xsamples; ++xs) {
            const unsigned j = (unsigned)xs & ((unsigned)xsamples - 1), k = (unsigned)ys & ((unsigned)xsamples - 1);
            double a = (double)xs + (double)sx[k % sx.size()] / (double)xsamples;
            double b = (double)ys + (double)sy[j % sy.size()] / (double)ysamples;
            a /= (double)xsamples;
            b /= (double)ysamples;
            a += 0.5 / (xsamples * xsamples);
            b += 0.5 / (ysamples * ysamples);
            jit[2 * (ys * xsamples + xs)] = a;
            jit[2 * (ys * xsamples + xs) + 1] = b;
        }
}

static int frame_buf(ri_b200_accel *a, int slot, uint64_t bytes, void **out)
{
    if (a->frame_bytes[slot] < bytes) {
        cudaFree(a->d_frame[slot]); a->d_frame[slot] = nullptr; a->frame_bytes[slot] = 0;
        CUDA_OK(cudaMalloc(&a->d_frame[slot], bytes));
        a->frame_bytes[slot] = bytes;
    }
    *out = a->d_frame[slot];
    return 0;
}

// window states at the first `segments` segment starts of the stream seeded `seed`, cached in the accelerator (rebuilt only for
// another seed or a longer stream): log2(capacity) dependent launches, each one Horner pass over a 19937-term jump polynomial
static int mt_states_ensure(ri_b200_accel *a, uint32_t seed, uint32_t segments, cudaStream_t st)
{
    if (segments > (1u << kMtJumpPolys)) return fail("MT19937 stream too long for the jump table (%u segments)", segments);
    if (!a->d_mt_polys) {
        CUDA_OK(cudaMalloc((void **)&a->d_mt_polys, sizeof(kMtJumpPoly)));
        CUDA_OK(cudaMemcpyAsync(a->d_mt_polys, kMtJumpPoly, sizeof(kMtJumpPoly), cudaMemcpyHostToDevice, st));
    }
    if (a->mt_states_seed != seed || a->mt_states_cap < segments) {          // (re)build the table of window states
        uint32_t cap = 1;
        while (cap < segments) cap <<= 1;
        if (a->mt_states_cap < cap) {
            cudaFree(a->d_mt_states); a->d_mt_states = nullptr; a->mt_states_cap = 0;
            CUDA_OK(cudaMalloc((void **)&a->d_mt_states, (size_t)cap * kMtN * sizeof(uint32_t)));
            a->mt_states_cap = cap;
        }
        uint32_t w0[kMtN];
        w0[0] = seed;
        for (int i = 1; i < kMtN; ++i) w0[i] = 69069u * w0[i - 1];                 // seedMT2, random.c:98-112
        CUDA_OK(cudaMemcpyAsync(a->d_mt_states, w0, sizeof(w0), cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));                                        // w0 is on the stack
        int m = 0;
        for (uint32_t have = 1; have < a->mt_states_cap; have <<= 1, ++m) {
            mt_jump_kernel<<<have, 256, 0, st>>>(a->d_mt_polys + (size_t)m * kMtN, a->d_mt_states, have, a->mt_states_cap);
            LAUNCHED();
        }
        a->mt_states_seed = seed;
    }
    return 0;
}

// The whole stream in parallel: window states at every segment start come from the doubling tree of jumps
// (log2(segments) launches, cached per seed in the accelerator), then one CTA per segment regenerates its blocks.
static int mt_stream_launch(ri_b200_accel *a, uint32_t seed, uint32_t segments, uint64_t total_blocks, uint32_t *d_out, cudaStream_t st)
{
    if (segments <= 1) {
        mt_kernel<<<1, 256, 0, st>>>(seed, nullptr, total_blocks, total_blocks, d_out);
        LAUNCHED();
        return 0;
    }
    if (mt_states_ensure(a, seed, segments, st)) return -1;
    mt_kernel<<<segments, 256, 0, st>>>(seed, a->d_mt_states, (uint64_t)kMtSegBlocks, total_blocks, d_out);
    LAUNCHED();
    CUDA_OK(cudaGetLastError());
    return 0;
}

// One-time cost made explicit: a caller that knows it will draw up to `max_words` words of the stream seeded `seed` (every frame and
// gather entry point with rng_mode 0 / the MT19937 gathers) pays for the window-state table here, asynchronously on the
// accelerator's stream, instead of inside its first big call (about 5 ms per doubling of the stream length beyond 638 976 words).
extern "C" int ri_b200_mt_prepare(ri_b200_accel_t *a, uint32_t seed, uint64_t max_words)
{
    if (!a) return fail("null argument");
    if (a->device < 0) return fail("host-only accelerator: no device records, no CPU fallback");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const uint64_t blocks = (max_words + kMtN - 1) / kMtN;
    const uint64_t segments = (blocks + kMtSegBlocks - 1) / kMtSegBlocks;
    if (segments <= 1) return 0;
    if (segments > (1ull << kMtJumpPolys)) return fail("MT19937 stream too long for the jump table (%llu segments)", (unsigned long long)segments);
    return mt_states_ensure(a, seed, (uint32_t)segments, a->stream);
}

template <typename Real>
static int render_ao_impl(ri_b200_accel *a, const ri_b200_frame_t &f, float *d_rgb, cudaStream_t st, ri_b200_frame_stats_t *stats,
                          Real *d_dump, uint64_t dump_count, int packed = 0, const ri_b200_sunsky_t *sky = nullptr, const bool dirtmap = false)
{
    std::vector<uint32_t> pix;
    std::vector<double> jit;
    // one MT19937 stream over several ranks: the per-bucket hit counts are exchanged through the host's callback (below)
    const bool shared_stream = f.rng_mode == 0 && f.world > 1;
    if (shared_stream && !a->hit_exchange) return fail("rng_mode 0 on several ranks needs ri_b200_set_hit_exchange");
    // The exchange callback is a collective of the host's (every rank of the frame calls it once).  A rank that fails BEFORE its
    // call must not leave the others waiting in theirs: until the real call below this guard makes the call on any early return,
    // with a NULL count array = "this rank has failed", so that the host's exchange fails the frame on every rank together.
    struct ExchangeGuard {
        ri_b200_accel *a; bool armed;
        ~ExchangeGuard() { if (armed) a->hit_exchange(a->hit_exchange_user, nullptr, 0, nullptr, nullptr); }
    } exchange_guard{a, shared_stream};
    std::vector<uint32_t> bfirst;
    // the pixel list of a rank depends only on the frame's geometry: frame after frame of the same size it is kept on the device
    // (4096^2: 16.7 M pixels = 30 ms of host work and a 67 MB upload per frame otherwise)
    const bool pix_cached = !shared_stream && a->pixc_pix && a->pixc_w == f.width && a->pixc_h == f.height && a->pixc_bucket == f.bucket_size &&
                            a->pixc_rank == f.rank && a->pixc_world == f.world;
    if (!pix_cached) pixel_order(f, pix, shared_stream ? &bfirst : nullptr);
    jitter_table(f.xsamples, f.ysamples, jit);
    const uint64_t npix = pix_cached ? a->pixc_n : pix.size();
    const uint32_t nbuckets = shared_stream ? (uint32_t)bfirst.size() - 1 : 0;
    // the sun-sky transport gathers with a fixed 8 x 8 pattern whatever Option "gather" says (ambientocclusion.c:371-374)
    // ... and the dirt-map transport with a fixed 4 x 4 one (dirtmap.c:261-265)
    const int ntheta = sky ? 8 : (dirtmap ? 4 : f.ntheta), nphi = sky ? 8 : (dirtmap ? 4 : f.nphi);
    const int spp = f.xsamples * f.ysamples, N = ntheta * nphi;
    const uint64_t nsamples = npix * (uint64_t)spp;
    if (nsamples >= 0xfffffff0ull) return fail("too many samples per rank for 32-bit sample ids");

    FrameDev F;
    for (int i = 0; i < 16; ++i) F.c2w[i] = f.c2w[i];
    F.flength_signed = (double)(float)(f.is_rh ? -1.0 : 1.0) * f.flength;
    F.w = (double)f.width; F.h = (double)f.height;
    F.width = f.width; F.height = f.height; F.xsamples = f.xsamples; F.ysamples = f.ysamples;
    F.ntheta = ntheta; F.nphi = nphi; F.spp = spp; F.nao = N;
    F.rng_mode = f.rng_mode; F.seed = f.seed;
    F.ao_eps = (sky || dirtmap) ? 1.0e-5 : 1.0e-6;                       // dirtmap.c:96

    const int cap = stack_capacity(a);
    const size_t smem = (size_t)cap * kBlock * sizeof(uint32_t);
    if (smem > 200 * 1024) return fail("BVH depth %d exceeds the shared-memory traversal stack", cap);
    if (smem > 48 * 1024) {
        CUDA_OK(cudaFuncSetAttribute(primary_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_OK(cudaFuncSetAttribute(ao_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const size_t sky_smem = smem + 3 * kBlock * sizeof(float);
    if (sky && sky_smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(sunsky_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sky_smem));
    const size_t dirt_smem = smem + kBlock * sizeof(double);
    if (dirtmap && dirt_smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(dirtmap_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dirt_smem));

    const uint32_t ntiles = (uint32_t)((nsamples + kScanTile - 1) / kScanTile);
    void *p = nullptr;
    uint32_t *d_pix, *d_prim, *d_tiles, *d_srank, *d_ranks;
    double *d_jit;
    Real *d_t;
    if (frame_buf(a, 0, (npix + 1) * 4 + jit.size() * 8 + 64 + (shared_stream ? npix * 4 + ((uint64_t)nbuckets + 1) * 16 : 0), &p)) return -1;
    d_jit = (double *)p;
    long long *d_bdelta = (long long *)(d_jit + jit.size());                 // shared_stream only: [nbuckets]
    d_pix = (uint32_t *)(d_bdelta + nbuckets);
    uint32_t *d_pixb = d_pix + npix + 1, *d_bfirst = d_pixb + (shared_stream ? npix : 0), *d_bhits = d_bfirst + nbuckets + 1;
    F.pix_bucket = nullptr; F.bucket_delta = nullptr;
    const bool textured = a->d_tex != nullptr && !sky;                   // the sun-sky branch does not look at the texture (ambientocclusion.c:369-376)
    const bool with_normals = make_view<Real>(a).normals != nullptr, with_uv = with_normals || textured;
    if (frame_buf(a, 1, (nsamples + 1) * sizeof(Real) * (with_uv ? 3 : 1), &p)) return -1;
    d_t = (Real *)p;
    Real *d_uv = with_uv ? d_t + nsamples : nullptr;
    if (frame_buf(a, 2, (nsamples + 1) * 4 * 3 + ((uint64_t)ntiles + 4) * 4, &p)) return -1;
    d_prim = (uint32_t *)p; d_srank = d_prim + nsamples; d_ranks = d_srank + nsamples; d_tiles = d_ranks + nsamples;
    uint32_t *d_total = d_tiles + ntiles;

    SceneView<Real> S = make_view<Real>(a);
    CUDA_OK(cudaEventRecord(a->ev[0], st));
    if (npix) {
        CUDA_OK(cudaMemcpyAsync(d_jit, jit.data(), jit.size() * 8, cudaMemcpyHostToDevice, st));
        if (pix_cached) {
            d_pix = a->pixc_pix;
        } else if (!shared_stream) {                 // (re)fill the cache and read the list from there
            if (a->pixc_cap < npix + 1) {
                cudaFree(a->pixc_pix); a->pixc_pix = nullptr; a->pixc_cap = 0;
                if (cudaMalloc((void **)&a->pixc_pix, (npix + 1) * 4) == cudaSuccess) a->pixc_cap = npix + 1; else cudaGetLastError();
            }
            if (a->pixc_pix) {
                d_pix = a->pixc_pix;
                a->pixc_w = f.width; a->pixc_h = f.height; a->pixc_bucket = f.bucket_size; a->pixc_rank = f.rank; a->pixc_world = f.world; a->pixc_n = npix;
            }
            CUDA_OK(cudaMemcpyAsync(d_pix, pix.data(), npix * 4, cudaMemcpyHostToDevice, st));
            CUDA_OK(cudaStreamSynchronize(st));      // pix is a local; the cached copy must be complete before the key is trusted
        } else {
            CUDA_OK(cudaMemcpyAsync(d_pix, pix.data(), npix * 4, cudaMemcpyHostToDevice, st));
        }
    }
    std::vector<uint32_t> pixb;
    if (shared_stream && npix) {
        pixb.resize(npix);
        for (uint32_t b = 0; b < nbuckets; ++b)
            for (uint32_t q = bfirst[b]; q < bfirst[b + 1]; ++q) pixb[q] = b;
        CUDA_OK(cudaMemcpyAsync(d_pixb, pixb.data(), npix * 4, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(d_bfirst, bfirst.data(), ((size_t)nbuckets + 1) * 4, cudaMemcpyHostToDevice, st));
    }
    // packed: 0 = full framebuffer, cleared first; 1 = this rank's pixels packed in visiting order; 2 = full framebuffer addressing
    // WITHOUT clearing -- the buffer is shared by the ranks of a node (peer memory), each storing its own tiles into it
    if (packed == 0) CUDA_OK(cudaMemsetAsync(d_rgb, 0, (size_t)f.width * f.height * 3 * sizeof(float), st));
    uint32_t nhits = 0;
    if (nsamples) {
        primary_kernel<Real><<<(unsigned)((nsamples + kBlock - 1) / kBlock), kBlock, smem, st>>>(S, F, d_pix, d_jit, nsamples, d_t, d_prim, d_uv);
        LAUNCHED();
        scan_tile_sums<<<ntiles, kScanBlock, 0, st>>>(d_prim, nsamples, d_tiles);
        LAUNCHED();
        scan_tile_offsets<<<1, kScanBlock, 0, st>>>(d_tiles, ntiles, d_total);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(a->h_pin, d_total, 4, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaEventRecord(a->ev[1], st));
        CUDA_OK(cudaStreamSynchronize(st));
        nhits = *(uint32_t *)a->h_pin;
    } else {
        CUDA_OK(cudaEventRecord(a->ev[1], st));
    }

    uint64_t frame_hits = nhits;                 // hit samples of the whole frame (== nhits on one rank)
    if (shared_stream) {
        // ri_b200_set_hit_exchange: this rank's hit samples per bucket go out, the frame-wide count before each of them comes back
        std::vector<uint32_t> bhits(nbuckets, 0u);
        std::vector<uint64_t> bbase(nbuckets, 0ull);
        if (nbuckets && nsamples) {
            bucket_hits_kernel<<<nbuckets, 256, 0, st>>>(d_prim, d_bfirst, (uint32_t)spp, d_bhits);
            LAUNCHED();
            CUDA_OK(cudaMemcpyAsync(bhits.data(), d_bhits, (size_t)nbuckets * 4, cudaMemcpyDeviceToHost, st));
            CUDA_OK(cudaStreamSynchronize(st));
        }
        exchange_guard.armed = false;
        if (a->hit_exchange(a->hit_exchange_user, bhits.data(), nbuckets, bbase.data(), &frame_hits) != 0)
            return fail("the hit-count exchange failed (on this rank or on another one)");
        std::vector<long long> delta(nbuckets);
        uint64_t before = 0;
        for (uint32_t b = 0; b < nbuckets; ++b) {
            if (bbase[b] + bhits[b] > frame_hits) return fail("hit-count exchange: bucket %u ends beyond the frame's hit samples", b);
            delta[b] = (long long)bbase[b] - (long long)before;
            before += bhits[b];
        }
        if (before != nhits) return fail("hit-count exchange: bucket sums disagree with the scan");
        if (nbuckets) {
            CUDA_OK(cudaMemcpyAsync(d_bdelta, delta.data(), (size_t)nbuckets * 8, cudaMemcpyHostToDevice, st));
            CUDA_OK(cudaStreamSynchronize(st));                                    // delta is a local
            F.pix_bucket = d_pixb; F.bucket_delta = d_bdelta;
        }
    }

    const uint64_t nao_rays = (uint64_t)nhits * (uint64_t)N;
    double *d_texcol = nullptr;
    Real *d_rec = nullptr;
    uint32_t *d_occ = nullptr, *d_mt = nullptr;
    if (frame_buf(a, 3, ((uint64_t)nhits + 1) * 12 * sizeof(Real), &p)) return -1;
    d_rec = (Real *)p;
    if (frame_buf(a, 4, ((uint64_t)nhits + 1) * 4, &p)) return -1;
    d_occ = (uint32_t *)p;
    // on several ranks every rank generates the stream of the whole frame (interleaved buckets touch nearly every segment anyway)
    const uint64_t mt_blocks = f.rng_mode == 0 && nhits ? (2 * frame_hits * (uint64_t)N + kMtN - 1) / kMtN : 0;
    if (frame_buf(a, 5, (mt_blocks * kMtN + 4) * 4, &p)) return -1;
    d_mt = (uint32_t *)p;
    const uint32_t mt_segments = (uint32_t)((mt_blocks + kMtSegBlocks - 1) / kMtSegBlocks);

    if (nhits) {
        TexDev tex;
        tex.data = textured ? a->d_tex : nullptr; tex.width = a->tex_w; tex.height = a->tex_h; tex.st = a->d_st; tex.flags = a->d_attr_flags;
        tex.texcol = nullptr;
        if (textured) {
            if (frame_buf(a, 9, ((uint64_t)nhits + 1) * 3 * sizeof(double), &p)) return -1;
            tex.texcol = (double *)p;
        }
        d_texcol = tex.texcol;
        compact_kernel<Real><<<ntiles, kScanBlock, 0, st>>>(S, F, d_pix, d_jit, nsamples, d_t, d_prim, d_uv, d_tiles, d_srank, d_ranks, d_rec, tex);
        LAUNCHED();
        CUDA_OK(cudaMemsetAsync(d_occ, 0, (uint64_t)nhits * 4, st));
    } else if (nsamples) {
        CUDA_OK(cudaMemsetAsync(d_srank, 0xff, nsamples * 4, st));
    }
    CUDA_OK(cudaEventRecord(a->ev[2], st));
    if (mt_blocks) {
        if (mt_stream_launch(a, f.seed, mt_segments, mt_blocks, d_mt, st)) return -1;
    }
    CUDA_OK(cudaEventRecord(a->ev[3], st));
    // tiny scenes (a few hundred triangles, rays that end after a handful of steps): the fused one-ray-per-thread kernel
    // wins (C1: 11.5 ms vs 17.9 ms); everything else goes through the persistent traverser (1M-triangle soup: 167 ms vs 278 ms)
    const char *force = getenv("B200_FUSED_AO_TEST");         // test hook: exercise both paths on the same scene
    const bool fused_ao = force ? atoi(force) != 0 : (a->tree.ntris < 4096);
    double *d_lo = nullptr;
    if (sky) {                                    // sun-sky gather: one lane per ray, sky lookup on misses, per-sample sums in order
        ri_b200_sunsky_t *d_sky = nullptr;
        if (frame_buf(a, 8, sizeof(ri_b200_sunsky_t), &p)) return -1;
        d_sky = (ri_b200_sunsky_t *)p;
        if (frame_buf(a, 9, ((uint64_t)nhits + 1) * 3 * sizeof(double), &p)) return -1;
        d_lo = (double *)p;
        CUDA_OK(cudaMemcpyAsync(d_sky, sky, sizeof(ri_b200_sunsky_t), cudaMemcpyHostToDevice, st));
        if (nao_rays && fused_ao) {
            const uint64_t blocks = (nao_rays + kBlock - 1) / kBlock;
            if (blocks > 0x7fffffffull) return fail("too many gather rays in one frame pass");
            sunsky_kernel<Real><<<(unsigned)blocks, kBlock, sky_smem, st>>>(S, F, d_sky, nao_rays, d_rec, d_ranks, d_pix, d_mt, d_lo, (uint32_t)cap);
            LAUNCHED();
        } else if (nao_rays) {                    // wavefront: gather rays and sun shadow rays through the pooled occlusion traverser
            const uint32_t chunk_samples = (1u << 24) / 64u, nsun = (uint32_t)(sky->nsun > 0 ? sky->nsun : 0);
            const uint64_t ray_words = sizeof(Real) == 8 ? 6 : 8;
            const uint64_t buf_samples = nhits < chunk_samples ? nhits : chunk_samples;
            if (frame_buf(a, 7, buf_samples * 64 * ray_words * sizeof(Real), &p)) return -1;
            Real *d_rays = (Real *)p;
            if (frame_buf(a, 10, buf_samples * 64 + buf_samples * (nsun + 1) + 64, &p)) return -1;
            uint8_t *d_occ8 = (uint8_t *)p, *d_sunocc = d_occ8 + buf_samples * 64;
            if (frame_buf(a, 11, (buf_samples * (nsun + 1) + 1) * ray_words * sizeof(Real), &p)) return -1;
            Real *d_sunrays = (Real *)p;
            for (uint32_t r0 = 0; r0 < nhits; r0 += chunk_samples) {
                const uint32_t ns = (nhits - r0) < chunk_samples ? (nhits - r0) : chunk_samples;
                const uint64_t nr = (uint64_t)ns * 64;
                ao_gen_kernel<Real><<<(unsigned)((nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(F, nr, r0, d_rec, d_ranks, d_pix, d_mt, d_rays);
                LAUNCHED();
                if (trace_occlusion_locked<Real>(a, d_rays, nr, d_occ8, nullptr, 1, st)) return -1;
                if (nsun) {
                    const uint64_t nsr = (uint64_t)ns * nsun;
                    sun_rays_kernel<Real><<<(unsigned)((nsr + 255) / 256), 256, 0, st>>>(d_sky, d_rec, r0, ns, d_sunrays);
                    LAUNCHED();
                    if (launch_trace<Real, true, false>(a, d_sunrays, nsr, nullptr, d_sunocc, nullptr, st)) return -1;
                }
                sky_accum_kernel<Real><<<(ns + kBlock / 32 - 1) / (kBlock / 32), kBlock, 0, st>>>(d_sky, d_rays, d_occ8, d_sunocc, r0, ns, d_lo);
                LAUNCHED();
            }
        }
    } else if (dirtmap) {                         // dirt map: closest hit per gather ray, colour by distance, per-sample sums in order
        if (frame_buf(a, 8, ((uint64_t)nhits + 1) * 3 * sizeof(double), &p)) return -1;
        d_lo = (double *)p;
        if (nao_rays && fused_ao) {               // one lane per ray
            const uint64_t blocks = (nao_rays + kBlock - 1) / kBlock;
            if (blocks > 0x7fffffffull) return fail("too many gather rays in one frame pass");
            dirtmap_kernel<Real><<<(unsigned)blocks, kBlock, dirt_smem, st>>>(S, F, nao_rays, d_rec, d_ranks, d_pix, d_mt, d_texcol, d_lo, (uint32_t)cap);
            LAUNCHED();
        } else if (nao_rays) {                    // wavefront through the pooled closest-hit traverser
            using Hit = typename RayIO<Real>::Hit;
            const uint32_t chunk_samples = (1u << 24) / 16u;
            const uint64_t ray_words = sizeof(Real) == 8 ? 6 : 8;
            const uint64_t buf_samples = nhits < chunk_samples ? nhits : chunk_samples;
            if (frame_buf(a, 7, buf_samples * 16 * ray_words * sizeof(Real), &p)) return -1;
            Real *d_rays = (Real *)p;
            if (frame_buf(a, 10, (buf_samples * 16 + 1) * sizeof(Hit), &p)) return -1;
            Hit *d_hits = (Hit *)p;
            for (uint32_t r0 = 0; r0 < nhits; r0 += chunk_samples) {
                const uint32_t ns = (nhits - r0) < chunk_samples ? (nhits - r0) : chunk_samples;
                const uint64_t nr = (uint64_t)ns * 16;
                ao_gen_kernel<Real><<<(unsigned)((nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(F, nr, r0, d_rec, d_ranks, d_pix, d_mt, d_rays);
                LAUNCHED();
                if (launch_trace<Real, false, false>(a, d_rays, nr, d_hits, nullptr, nullptr, st)) return -1;
                dirt_accum_kernel<Real><<<(ns + 255) / 256, 256, 0, st>>>(d_hits, d_texcol, r0, ns, d_lo);
                LAUNCHED();
            }
        }
    } else if (nao_rays && fused_ao) {                   // one lane per ray, generation fused with a one-ray-per-thread traversal
        const uint64_t blocks = (nao_rays + kBlock - 1) / kBlock;
        if (blocks > 0x7fffffffull) return fail("too many occlusion rays in one frame pass");
        ao_kernel<Real><<<(unsigned)blocks, kBlock, smem, st>>>(S, F, nao_rays, 0u, d_rec, d_ranks, d_pix, d_mt, d_occ, d_dump, dump_count);
        LAUNCHED();
    } else if (nao_rays) {                        // wavefront: ray batches of <= 2^24 rays through the persistent traverser
        const uint32_t chunk_samples = (1u << 24) / (uint32_t)N ? (1u << 24) / (uint32_t)N : 1u;
        const uint64_t ray_words = sizeof(Real) == 8 ? 6 : 8;
        const uint64_t buf_rays = (uint64_t)(nhits < chunk_samples ? nhits : chunk_samples) * (uint64_t)N;
        if (frame_buf(a, 7, buf_rays * ray_words * sizeof(Real), &p)) return -1;
        Real *d_rays = (Real *)p;
        for (uint32_t r0 = 0; r0 < nhits; r0 += chunk_samples) {
            const uint32_t ns = (nhits - r0) < chunk_samples ? (nhits - r0) : chunk_samples;
            const uint64_t nr = (uint64_t)ns * (uint64_t)N;
            ao_gen_kernel<Real><<<(unsigned)((nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(F, nr, r0, d_rec, d_ranks, d_pix, d_mt, d_rays);
            LAUNCHED();
            if (d_dump && r0 == 0 && dump_count)
                CUDA_OK(cudaMemcpyAsync(d_dump, d_rays, (dump_count < nr ? dump_count : nr) * ray_words * sizeof(Real), cudaMemcpyDeviceToDevice, st));
            if (trace_occlusion_locked<Real>(a, d_rays, nr, nullptr, d_occ + r0, (uint32_t)N, st)) return -1;
        }
    }
    CUDA_OK(cudaEventRecord(a->ev[4], st));
    if (npix) {
        if (sky || dirtmap) resolve_rgb_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(F, d_pix, npix, d_srank, d_lo, d_rgb, packed == 1);
        else if (textured && d_texcol) resolve_tex_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(F, d_pix, npix, d_srank, d_occ, d_texcol, d_rgb, packed == 1);
        else resolve_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(F, d_pix, npix, d_srank, d_occ, d_rgb, packed == 1);
        LAUNCHED();
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(a->ev[5], st));
    if (stats) {
        CUDA_OK(cudaEventSynchronize(a->ev[5]));
        float ms;
        std::memset(stats, 0, sizeof(*stats));
        stats->nrays_primary = nsamples; stats->nhits_primary = nhits;
        stats->nrays_ao = nao_rays + (sky ? (uint64_t)nhits * (uint64_t)sky->nsun : 0);     // + one shadow ray per sun light and hit sample
        CUDA_OK(cudaEventElapsedTime(&ms, a->ev[0], a->ev[5])); stats->ms_total = ms;
        CUDA_OK(cudaEventElapsedTime(&ms, a->ev[0], a->ev[1])); stats->ms_primary = ms;
        CUDA_OK(cudaEventElapsedTime(&ms, a->ev[2], a->ev[3])); stats->ms_rng = ms;
        CUDA_OK(cudaEventElapsedTime(&ms, a->ev[3], a->ev[4])); stats->ms_ao = ms;
        CUDA_OK(cudaEventElapsedTime(&ms, a->ev[4], a->ev[5])); stats->ms_resolve = ms;
    }
    return 0;
}

static int check_frame(const ri_b200_accel *a, const ri_b200_frame_t *f)
{
    if (!a || !f) return fail("null argument");
    if (f->width < 1 || f->height < 1 || f->width > 65535 || f->height > 65535) return fail("bad frame size");
    if (f->xsamples < 1 || f->ysamples < 1 || f->ntheta < 1 || f->nphi < 1) return fail("bad sample counts");
    if (f->world < 1 || f->rank < 0 || f->rank >= f->world) return fail("bad rank/world");
    if (f->rng_mode == 0 && f->world != 1 && !a->hit_exchange)
        return fail("rng_mode 0 (single MT19937 stream in reference order) on world > 1 needs ri_b200_set_hit_exchange()");
    if (f->precision != RI_B200_PREC_F32 && f->precision != RI_B200_PREC_F64) return fail("bad precision");
    return need(a, (uint32_t)f->precision);
}

}  // namespace b200

extern "C" int ri_b200_render_ao_dev(ri_b200_accel_t *a, const ri_b200_frame_t *f, float *d_rgb, void *stream,
                                     ri_b200_frame_stats_t *stats)
{
    if (check_frame(a, f)) return -1;
    if (!d_rgb) return fail("null framebuffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : a->stream;
    if (f->precision == RI_B200_PREC_F64) return render_ao_impl<double>(a, *f, d_rgb, st, stats, nullptr, 0);
    return render_ao_impl<float>(a, *f, d_rgb, st, stats, nullptr, 0);
}

extern "C" int ri_b200_render_ao_tiles_dev(ri_b200_accel_t *a, const ri_b200_frame_t *f, float *d_packed, void *stream,
                                           ri_b200_frame_stats_t *stats)
{
    if (check_frame(a, f)) return -1;
    if (!d_packed) return fail("null buffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : a->stream;
    if (f->precision == RI_B200_PREC_F64) return render_ao_impl<double>(a, *f, d_packed, st, stats, nullptr, 0, 1);
    return render_ao_impl<float>(a, *f, d_packed, st, stats, nullptr, 0, 1);
}

// ---- fused multi-GPU resolve: every rank's resolve kernel stores its tiles straight into ONE framebuffer that lives on rank 0's GPU
// and is mapped into the other ranks' address spaces (CUDA IPC over NVLink / NVSwitch peer memory) -- the gather that would follow
// the frame is done by the stores themselves, no packed slab, no NCCL call.
extern "C" int ri_b200_render_ao_peer_dev(ri_b200_accel_t *a, const ri_b200_frame_t *f, float *d_rgb_shared, void *stream,
                                          ri_b200_frame_stats_t *stats)
{
    if (check_frame(a, f)) return -1;
    if (!d_rgb_shared) return fail("null framebuffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : a->stream;
    if (f->precision == RI_B200_PREC_F64) return render_ao_impl<double>(a, *f, d_rgb_shared, st, stats, nullptr, 0, 2);
    return render_ao_impl<float>(a, *f, d_rgb_shared, st, stats, nullptr, 0, 2);
}

extern "C" int ri_b200_set_hit_exchange(ri_b200_accel_t *a, ri_b200_hit_exchange_fn fn, void *user)
{
    if (!a) return fail("null accelerator");
    std::lock_guard<std::mutex> lock(a->mu);
    a->hit_exchange = fn; a->hit_exchange_user = user;
    return 0;
}

extern "C" void *ri_b200_peer_alloc(uint64_t bytes, int device, uint8_t handle_out[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    cudaIpcMemHandle_t h;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); fail("peer_alloc: cudaMalloc(%llu) failed", (unsigned long long)bytes); return nullptr; }
    if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaGetLastError(); cudaFree(p); fail("peer_alloc: cudaIpcGetMemHandle failed"); return nullptr; }
    if (handle_out) std::memcpy(handle_out, &h, 64);
    return p;
}
extern "C" void *ri_b200_peer_open(const uint8_t handle[64], int device)
{
    void *p = nullptr;
    cudaIpcMemHandle_t h;
    if (!handle) { fail("null handle"); return nullptr; }
    std::memcpy(&h, handle, 64);
    if (cudaSetDevice(device) != cudaSuccess || cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        fail("peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
extern "C" int ri_b200_peer_close(void *p, int device)
{
    if (!p) return 0;
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaIpcCloseMemHandle(p));
    return 0;
}
extern "C" int ri_b200_peer_free(void *p, int device)
{
    if (!p) return 0;
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaFree(p));
    return 0;
}
extern "C" int ri_b200_peer_read(const void *p, void *host, uint64_t bytes, int device)
{
    if (!p || !host) return fail("null argument");
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(host, p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int64_t ri_b200_frame_pixels(const ri_b200_frame_t *f, uint32_t *out, int64_t capacity)
{
    if (!f) return fail("null argument");
    if (f->width < 1 || f->height < 1 || f->width > 65535 || f->height > 65535) return fail("bad frame size");
    if (f->world < 1 || f->rank < 0 || f->rank >= f->world) return fail("bad rank/world");
    std::vector<uint32_t> pix;
    pixel_order(*f, pix);
    if (!out) return (int64_t)pix.size();
    if (capacity < (int64_t)pix.size()) return fail("capacity too small");
    if (!pix.empty()) std::memcpy(out, pix.data(), pix.size() * sizeof(uint32_t));
    return (int64_t)pix.size();
}

extern "C" int ri_b200_render_ao(ri_b200_accel_t *a, const ri_b200_frame_t *f, float *rgb_out, ri_b200_frame_stats_t *stats)
{
    if (check_frame(a, f)) return -1;
    if (!rgb_out) return fail("null framebuffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const size_t bytes = (size_t)f->width * f->height * 3 * sizeof(float);
    void *p = nullptr;
    if (frame_buf(a, 6, bytes, &p)) return -1;
    int rc;
    if (f->precision == RI_B200_PREC_F64) rc = render_ao_impl<double>(a, *f, (float *)p, a->stream, stats, nullptr, 0);
    else rc = render_ao_impl<float>(a, *f, (float *)p, a->stream, stats, nullptr, 0);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(rgb_out, p, bytes, cudaMemcpyDeviceToHost, a->stream));
    CUDA_OK(cudaStreamSynchronize(a->stream));
    return 0;
}

extern "C" int ri_b200_render_sunsky(ri_b200_accel_t *a, const ri_b200_frame_t *f, const ri_b200_sunsky_t *sky, float *rgb_out,
                                     ri_b200_frame_stats_t *stats)
{
    if (check_frame(a, f)) return -1;
    if (!rgb_out || !sky) return fail("null argument");
    if (sky->nsun < 0 || sky->nsun > 4) return fail("bad sun light count %d", sky->nsun);
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const size_t bytes = (size_t)f->width * f->height * 3 * sizeof(float);
    void *p = nullptr;
    if (frame_buf(a, 6, bytes, &p)) return -1;
    int rc;
    if (f->precision == RI_B200_PREC_F64) rc = render_ao_impl<double>(a, *f, (float *)p, a->stream, stats, nullptr, 0, 0, sky);
    else rc = render_ao_impl<float>(a, *f, (float *)p, a->stream, stats, nullptr, 0, 0, sky);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(rgb_out, p, bytes, cudaMemcpyDeviceToHost, a->stream));
    CUDA_OK(cudaStreamSynchronize(a->stream));
    return 0;
}

extern "C" int ri_b200_render_sunsky_tiles_dev(ri_b200_accel_t *a, const ri_b200_frame_t *f, const ri_b200_sunsky_t *sky, float *d_packed,
                                               void *stream, ri_b200_frame_stats_t *stats)
{
    if (check_frame(a, f)) return -1;
    if (!d_packed || !sky) return fail("null argument");
    if (sky->nsun < 0 || sky->nsun > 4) return fail("bad sun light count %d", sky->nsun);
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : a->stream;
    if (f->precision == RI_B200_PREC_F64) return render_ao_impl<double>(a, *f, d_packed, st, stats, nullptr, 0, 1, sky);
    return render_ao_impl<float>(a, *f, d_packed, st, stats, nullptr, 0, 1, sky);
}

extern "C" int ri_b200_render_dirtmap(ri_b200_accel_t *a, const ri_b200_frame_t *f, float *rgb_out, ri_b200_frame_stats_t *stats)
{
    if (check_frame(a, f)) return -1;
    if (!rgb_out) return fail("null framebuffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const size_t bytes = (size_t)f->width * f->height * 3 * sizeof(float);
    void *p = nullptr;
    if (frame_buf(a, 6, bytes, &p)) return -1;
    int rc;
    if (f->precision == RI_B200_PREC_F64) rc = render_ao_impl<double>(a, *f, (float *)p, a->stream, stats, nullptr, 0, 0, nullptr, true);
    else rc = render_ao_impl<float>(a, *f, (float *)p, a->stream, stats, nullptr, 0, 0, nullptr, true);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(rgb_out, p, bytes, cudaMemcpyDeviceToHost, a->stream));
    CUDA_OK(cudaStreamSynchronize(a->stream));
    return 0;
}

extern "C" int ri_b200_sunsky_rgb(const ri_b200_sunsky_t *sky, const float *dirs, uint64_t n, float *rgb_out, int device)
{
    if (!sky || (n && (!dirs || !rgb_out))) return fail("null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail("no CUDA device: libb200accel has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail("bad device %d", device);
    if (n == 0) return 0;
    CUDA_OK(cudaSetDevice(device));
    ri_b200_sunsky_t *d_sky = nullptr;
    float *d_dirs = nullptr, *d_rgb = nullptr;
    auto body = [&]() -> int {
        CUDA_OK(cudaMalloc((void **)&d_sky, sizeof(*sky)));
        CUDA_OK(cudaMalloc((void **)&d_dirs, n * 3 * sizeof(float)));
        CUDA_OK(cudaMalloc((void **)&d_rgb, n * 3 * sizeof(float)));
        CUDA_OK(cudaMemcpy(d_sky, sky, sizeof(*sky), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_dirs, dirs, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
        sky_rgb_kernel<<<(unsigned)((n + 127) / 128), 128>>>(d_sky, d_dirs, n, d_rgb);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpy(rgb_out, d_rgb, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        return 0;
    };
    const int rc = body();
    cudaFree(d_sky); cudaFree(d_dirs); cudaFree(d_rgb);
    return rc;
}

extern "C" int ri_b200_mt_stream(ri_b200_accel_t *a, uint32_t seed, uint64_t n, uint32_t *out_u32, int device)
{
    if (!out_u32) return fail("null argument");
    if (n == 0) return 0;
    if (a && a->device < 0) return fail("host-only accelerator: no device records, no CPU fallback");
    CUDA_OK(cudaSetDevice(a ? a->device : device));
    const uint64_t blocks = (n + kMtN - 1) / kMtN;
    uint32_t *d = nullptr;
    CUDA_OK(cudaMalloc((void **)&d, blocks * kMtN * 4));
    int rc = 0;
    if (a) {                                     // parallel: jump-ahead tree + one CTA per segment
        std::lock_guard<std::mutex> lock(a->mu);
        rc = mt_stream_launch(a, seed, (uint32_t)((blocks + kMtSegBlocks - 1) / kMtSegBlocks), blocks, d, a->stream);
        if (!rc && cudaStreamSynchronize(a->stream) != cudaSuccess) rc = fail("mt stream failed");
    } else {                                     // sequential: one CTA walks the whole stream
        mt_kernel<<<1, 256>>>(seed, nullptr, blocks, blocks, d);
        LAUNCHED();
    }
    cudaError_t e = rc ? cudaSuccess : cudaMemcpy(out_u32, d, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (rc) return rc;
    if (e != cudaSuccess) return fail("mt stream copy failed: %s", cudaGetErrorString(e));
    return 0;
}

// Beam (4-corner frustum) visibility query, SURVEY 8a row a10 -- fp64 records, one beam per lane.
// Reference: ri_beam_set (src/render/beam.c:332-466), ri_bvh_intersect_beam_visibility (bvh.c:612-667), test_beam_aabb /
// test_beam_aabb_misses / get_n_point (bvh.c:1997-2089), test_beam_node (2097-2126), bvh_traverse_beam_visibility (2648-2746),
// bvh_intersect_leaf_node_beam_visibility (2435-2542), test_beam_triangle (2139-2281).  Only the testbed calls it in lucille.
#pragma once

namespace b200 {

struct BeamRegs {
    double org[3], dir[4][3], normal[4][3];
    int    dominant_axis, order;          // order = dirsign[dominant_axis]
};

// beam.c:332-466 -- returns false when the four directions do not share a sign on every axis (ri_beam_set returns -1)
__device__ __forceinline__ bool beam_set(BeamRegs &b, const double *p)
{
    double dir[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) dir[j][k] = p[3 + 3 * j + k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int zeros = 0, mask = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (fabs(dir[j][i]) < 1.0e-14) zeros++;
            else mask += (dir[j][i] < 0.0) ? 1 : -1;
        }
        if ((mask != -(4 - zeros)) && (mask != (4 - zeros))) return false;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) b.org[k] = p[k];
    double maxval = fabs(dir[0][0]);
    int ax = 0;
    if (maxval < fabs(dir[0][1])) { maxval = fabs(dir[0][0]); ax = 1; }        // sic, beam.c:389-392
    if (maxval < fabs(dir[0][2])) { maxval = fabs(dir[0][2]); ax = 2; }
    b.dominant_axis = ax;
    const bool sgn = ((ax == 0) ? dir[0][0] : (ax == 1) ? dir[0][1] : dir[0][2]) < 0.0;
    b.order = sgn ? 1 : 0;
    double nrm[3] = {ax == 0 ? 1.0 : 0.0, ax == 1 ? 1.0 : 0.0, ax == 2 ? 1.0 : 0.0};
    if (sgn) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double t = dir[i][0] * nrm[0] + dir[i][1] * nrm[1] + dir[i][2] * nrm[2];
        const double kk = (fabs(t) > 1.0e-14) ? 1024.0 / t : 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) b.dir[i][k] = kk * dir[i][k];
    }
    cross3(b.normal[0], b.dir[1], b.dir[0]);
    cross3(b.normal[1], b.dir[2], b.dir[1]);
    cross3(b.normal[2], b.dir[3], b.dir[2]);
    cross3(b.normal[3], b.dir[0], b.dir[3]);
    return true;
}

// bvh.c:2012-2089: true = the box may be hit
__device__ __forceinline__ bool beam_aabb(double lox, double hix, double loy, double hiy, double loz, double hiz, const BeamRegs &b)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double *n = b.normal[i];
        const double npx = (n[0] > 0.0) ? lox : hix, npy = (n[1] > 0.0) ? loy : hiy, npz = (n[2] > 0.0) ? loz : hiz;
        const double d = (npx - b.org[0]) * n[0] + (npy - b.org[1]) * n[1] + (npz - b.org[2]) * n[2];
        if (d > 0.0) return false;
    }
    return true;
}

// bvh.c:2139-2281
__device__ __forceinline__ int beam_triangle(const TriRegs<double> &tr, const BeamRegs &b)
{
    double u[4], v[4], t[4];
    int mask = 0;
    const double sx = b.org[0] - tr.v0[0], sy = b.org[1] - tr.v0[1], sz = b.org[2] - tr.v0[2];
    double q[3];
    {
        const double s[3] = {sx, sy, sz};
        cross3(q, s, tr.e1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double p[3];
        cross3(p, b.dir[i], tr.e2);
        const double a = tr.e1[0] * p[0] + tr.e1[1] * p[1] + tr.e1[2] * p[2];
        const double inva = (fabs(a) > 1.0e-14) ? 1.0 / a : 0.0;
        u[i] = (sx * p[0] + sy * p[1] + sz * p[2]) * inva;
        v[i] = (q[0] * b.dir[i][0] + q[1] * b.dir[i][1] + q[2] * b.dir[i][2]) * inva;
        t[i] = (tr.e2[0] * q[0] + tr.e2[1] * q[1] + tr.e2[2] * q[2]) * inva;
        const bool out = (u[i] < 0.0) || (u[i] > 1.0) || (v[i] < 0.0) || ((u[i] + v[i]) > 1.0) || (t[i] < 0.0) || (t[i] > 1.0e38);
        if (!out) mask |= 1 << i;
    }
    if (mask == 0xf) return 1;                                   // RI_BEAM_HIT_COMPLETELY
    if (mask != 0) return 2;                                     // RI_BEAM_HIT_PARTIALLY
    int c;
    c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c += (t[i] < 0.0);
    if (c == 4) return 0;
    c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c += (u[i] < 0.0);
    if (c != 0 && c != 4) return 2;
    c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c += (u[i] > 1.0);
    if (c != 0 && c != 4) return 2;
    c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c += (v[i] < 0.0);
    if (c != 0 && c != 4) return 2;
    c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c += ((u[i] + v[i]) >= 1.0);
    if (c != 0 && c != 4) return 2;
    return 0;
}

__global__ void __launch_bounds__(kBlock)
beam_visibility_kernel(const SceneView<double> S, const double *__restrict__ beams, const uint64_t n, int32_t *__restrict__ out)
{
    extern __shared__ uint32_t s_stack[];
    uint32_t *stk = s_stack + threadIdx.x;
    const uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    BeamRegs b;
    if (!beam_set(b, beams + 15 * i)) { out[i] = -1; return; }
    uint32_t cur = S.root_word;
    if (cur == kDoneWord) { out[i] = 0; return; }               // empty scene: bvh.c:625-628
    if (!beam_aabb(S.smin[0], S.smax[0], S.smin[1], S.smax[1], S.smin[2], S.smax[2], b)) { out[i] = 0; return; }
    uint32_t sp = 0;
    int result = 0;
    for (;;) {
        while (!(cur & kLeafFlag)) {
            NodeRegs<double> nd;
            load_node_wide(S.nodes + cur, nd);
            const bool h0 = beam_aabb(nd.x[0], nd.x[1], nd.y[0], nd.y[1], nd.z[0], nd.z[1], b);
            const bool h1 = beam_aabb(nd.x[2], nd.x[3], nd.y[2], nd.y[3], nd.z[2], nd.z[3], b);
            if (h0 && h1) {
                stk[sp * kBlock] = b.order ? nd.c0 : nd.c1;
                ++sp;
                cur = b.order ? nd.c1 : nd.c0;
            } else if (h0) cur = nd.c0;
            else if (h1) cur = nd.c1;
            else {
                if (sp == 0) { cur = kDoneWord; break; }
                --sp;
                cur = stk[sp * kBlock];
            }
        }
        if (cur == kDoneWord) break;
        {
            const uint32_t start = cur & ((1u << kLeafShift) - 1u), count = ((cur >> kLeafShift) & 15u) + 1u;
            for (uint32_t k = 0; k < count && result == 0; ++k) {
                TriRegs<double> tr;
                load_tri_wide(S.tris + start + k, tr);
                result = beam_triangle(tr, b);
            }
            if (result != 0) break;                               // first triangle that is not a complete miss decides
        }
        if (sp == 0) break;
        --sp;
        cur = stk[sp * kBlock];
    }
    out[i] = result;
}

}  // namespace b200

extern "C" int ri_b200_beam_visibility_batch(ri_b200_accel_t *a, const double *beams, uint64_t n, int32_t *out)
{
    if (need(a, RI_B200_PREC_F64)) return -1;
    if (n == 0) return 0;
    if (!beams || !out) return fail("null buffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const int cap = stack_capacity(a);
    const size_t smem = (size_t)cap * kBlock * sizeof(uint32_t);
    if (smem > 200 * 1024) return fail("BVH depth %d exceeds the shared-memory traversal stack", cap);
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(beam_visibility_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double *d_beams = nullptr;
    int32_t *d_out = nullptr;
    auto body = [&]() -> int {
        CUDA_OK(cudaMalloc((void **)&d_beams, n * 15 * sizeof(double)));
        CUDA_OK(cudaMalloc((void **)&d_out, n * sizeof(int32_t)));
        CUDA_OK(cudaMemcpyAsync(d_beams, beams, n * 15 * sizeof(double), cudaMemcpyHostToDevice, a->stream));
        beam_visibility_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, smem, a->stream>>>(make_view<double>(a), d_beams, n, d_out);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(out, d_out, n * sizeof(int32_t), cudaMemcpyDeviceToHost, a->stream));
        CUDA_OK(cudaStreamSynchronize(a->stream));
        return 0;
    };
    const int rc = body();
    cudaFree(d_beams); cudaFree(d_out);
    return rc;
}

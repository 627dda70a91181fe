// Hemisphere gathers at shading points: the per-point loops of three more ri_raytrace callers as ONE batched query
// over n points (P, N) -- the Monte Carlo branches (Option "use_qmc" defaults to 0, option.c:139) and, for the IBL and dome gathers,
// the quasi-Monte Carlo ones (scrambled Halton / Hammersley points over Faure permutations, qmc.c; nsamples rays per point):
//   RI_B200_GATHER_OCCLUSION  occlusion() shadeop        shader.c:680-768   coverage / nsamples
//   RI_B200_GATHER_IBL        ri_ibl_sample_cosweight    ibl.c:53-228       pi * sum(Le/pi) / (ntheta*nphi), Le from the angular map on a miss
//   RI_B200_GATHER_DOME       ri_domelight_sample        ibl.c:231-389      pi * sum(col*intensity/pi) / nsamples on misses
// All three draw a stratified ntheta x 3 ntheta fan (j outer, i inner), two MT19937 words per ray whether it hits or not
// (randomMT / randomMT2: same generator, same seed, random.c:163-247), so point p starts at stream word 2*N*p: the stream comes
// from the jump-ahead generator of frame.cuh and every ray is independent.
//
// Small scenes: one warp per point, lane l takes rays l, l+32, ...; after each group of 32 the contributions are added in ray order (every lane
// runs the same 32-step shuffle sum, adding +0.0 for rays that hit), which is the reference's accumulation order exactly.
// ri_ibl_sample_bruteforce (ibl.c:395-518) is not built: it overwrites the ray origin with the direction before measuring their
// distance (ibl.c:497-505), so every term is multiplied by invdist = 0 and the function returns zero power for any input.
#pragma once

namespace b200 {

struct GatherDev {
    int      kind, nsamples, ntheta, nphi;
    int      nper;                  // rays per point: ntheta * nphi, or nsamples in the quasi-Monte Carlo branches
    int      qmc, qmc_base;         // use_qmc; primes[inray->d] of the dome light's shift
    const int32_t *instance;        // inray->i per point, or null
    int      perm3[3], perm5[5], permd[97];   // Faure permutations of bases 3, 5 and qmc_base (qmc.c:182-260)
    uint64_t stream_offset;         // stream words consumed before point 0
    double   rad[3];                // DOME: col * intensity
    TexDev   env;                   // IBL
};

// qmc.c:329-349 generalized_vdC
__device__ __forceinline__ double generalized_vdc_dev(int i, const int base, const int *perm)
{
    double h = 0.0;
    const double f = 1.0 / (double)base;
    double factor = f;
    while (i > 0) {
        h += (double)perm[i % base] * factor;
        i /= base;
        factor *= f;
    }
    return h;
}

// ray k = j*ntheta + i of point p: stratified cosine fan about N (reflection.c:312-333 basis), normalised, origin per caller
__device__ __forceinline__ void gather_ray(const GatherDev &G, const double *__restrict__ points, const uint32_t *__restrict__ mt_stream,
                                           const uint64_t p, const uint32_t k, double org[3], double dir[3])
{
    const uint32_t N = (uint32_t)G.nper;
    const double P[3] = {points[6 * p], points[6 * p + 1], points[6 * p + 2]};
    const double Nn[3] = {points[6 * p + 3], points[6 * p + 4], points[6 * p + 5]};
    double b0[3], b1[3];
    ortho_basis(b0, b1, Nn);
    double theta, phi;
    if (G.qmc) {                                                            // ibl.c:107-151 + 540-581, 266-320
        const int inst = G.instance ? G.instance[p] : 0;
        double s0, s1;
        if (G.kind == RI_B200_GATHER_IBL) {                                 // scrambled Halton, bases 3 and 5, index i + inray->i
            s0 = 0.0 + generalized_vdc_dev((int)k + inst, 3, G.perm3);
            s1 = 0.0 + generalized_vdc_dev((int)k + inst, 5, G.perm5);
        } else {                                                            // Hammersley (i/n, base 3), both shifted by u
            const double u = generalized_vdc_dev(inst, G.qmc_base, G.permd);
            s0 = u + (double)k / (double)G.nsamples;
            s1 = u + generalized_vdc_dev((int)k, 3, G.perm3);
        }
        s0 = s0 - floor(s0); s1 = s1 - floor(s1);                           // mod_1, qmc.c:523-532
        theta = sqrt(s0);
        phi = 2.0 * 3.14159265358979323846 * s1;
    } else {
        const uint32_t j = k / (uint32_t)G.ntheta, i = k - j * (uint32_t)G.ntheta;
        const uint64_t w = G.stream_offset + 2 * ((uint64_t)N * p + k);
        const double r0 = (double)mt_stream[w] * 2.3283064365386963e-10;    // random.c:196,244
        const double r1 = (double)mt_stream[w + 1] * 2.3283064365386963e-10;
        theta = (G.kind == RI_B200_GATHER_OCCLUSION) ? sqrt((double)i + r0) / (double)G.ntheta     // shader.c:731
                                                     : sqrt(((double)i + r0) / (double)G.ntheta);  // ibl.c:174,337
        phi = 2.0 * 3.14159265358979323846 * ((double)j + r1) / (double)G.nphi;
    }
    const double lx = cos(phi) * theta, ly = sin(phi) * theta, lz = sqrt(1.0 - theta * theta);
#pragma unroll
    for (int q = 0; q < 3; ++q) dir[q] = lx * b0[q] + ly * b1[q] + lz * Nn[q];
    normalize3(dir);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        org[q] = P[q];
        if (G.kind == RI_B200_GATHER_OCCLUSION) org[q] += 0.0001 * dir[q];  // shader.c:750-752
        if (G.kind == RI_B200_GATHER_IBL) org[q] += Nn[q] * 0.0001;          // ibl.c:92-94
    }
}

__global__ void __launch_bounds__(kBlock)
point_gather_kernel(const SceneView<double> S, const GatherDev G, const double *__restrict__ points, const uint64_t npoints,
                    const uint32_t *__restrict__ mt_stream, double *__restrict__ out3, const uint32_t stack_cap)
{
    extern __shared__ uint32_t s_stack[];
    const uint64_t p = ((uint64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (p >= npoints) return;                                               // whole warps leave together
    const uint32_t N = (uint32_t)G.nper;
    double sum[3] = {0.0, 0.0, 0.0};
    uint32_t coverage = 0;
    for (uint32_t base = 0; base < N; base += 32) {
        const uint32_t k = base + lane;
        double c[3] = {0.0, 0.0, 0.0};
        bool hit = false;
        if (k < N) {
            double dir[3], org[3], t, u, v;
            uint32_t prim;
            gather_ray(G, points, mt_stream, p, k, org, dir);
            hit = trace_ray<double, true, false>(S, org, dir, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr);
            if (!hit && G.kind != RI_B200_GATHER_OCCLUSION) {
                double rad[3] = {G.rad[0], G.rad[1], G.rad[2]};
                if (G.kind == RI_B200_GATHER_IBL) ibl_fetch_dev(G.env, dir, rad);
                const double brdf = (G.qmc && G.kind == RI_B200_GATHER_IBL) ? (double)0.31831f : 1.0 / 3.14159265358979323846;   // ibl.c:139
#pragma unroll
                for (int q = 0; q < 3; ++q) c[q] = rad[q] * brdf;
            }
        }
        if (G.kind == RI_B200_GATHER_OCCLUSION) {
            coverage += (uint32_t)__popc(__ballot_sync(0xffffffffu, hit));
        } else {
            for (int l = 0; l < 32; ++l) {
#pragma unroll
                for (int q = 0; q < 3; ++q) sum[q] = sum[q] + __shfl_sync(0xffffffffu, c[q], l);
            }
        }
    }
    if (lane == 0) {
        double *o = out3 + 3 * p;
        if (G.kind == RI_B200_GATHER_OCCLUSION) {
            if (coverage > (uint32_t)G.nsamples) coverage = (uint32_t)G.nsamples;
            o[0] = o[1] = o[2] = (double)(float)((double)coverage / (double)(float)G.nsamples);
        } else if (G.kind == RI_B200_GATHER_IBL) {
            for (int q = 0; q < 3; ++q) o[q] = G.qmc ? sum[q] / (double)G.nsamples : 3.14159265358979323846 * sum[q] / (double)(G.ntheta * G.nphi);
        } else {
            for (int q = 0; q < 3; ++q) o[q] = 3.14159265358979323846 * sum[q] / (double)G.nsamples;
        }
    }
}

// ---- wavefront form (scenes past a few thousand triangles): the same rays written out, traced by the pooled occlusion traverser
// (pool.cuh, fp64 records), and accumulated per point in ray order.
__global__ void __launch_bounds__(kBlock)
gather_gen_kernel(const GatherDev G, const double *__restrict__ points, const uint64_t p0, const uint64_t nrays,
                  const uint32_t *__restrict__ mt_stream, double *__restrict__ rays_out)
{
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (gid >= nrays) return;
    const uint32_t N = (uint32_t)G.nper;
    const uint64_t p = p0 + gid / N;
    const uint32_t k = (uint32_t)(gid % N);
    double org[3], dir[3];
    gather_ray(G, points, mt_stream, p, k, org, dir);
    double2 *o = reinterpret_cast<double2 *>(rays_out) + 3 * gid;
    o[0] = make_double2(org[0], org[1]);
    o[1] = make_double2(org[2], dir[0]);
    o[2] = make_double2(dir[1], dir[2]);
}

__global__ void __launch_bounds__(kBlock)
gather_accum_kernel(const GatherDev G, const double *__restrict__ rays, const uint8_t *__restrict__ occ, const uint64_t p0,
                    const uint64_t npoints, double *__restrict__ out3)
{
    const uint64_t lp = ((uint64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;          // point within the chunk
    const uint32_t lane = threadIdx.x & 31u;
    if (lp >= npoints) return;
    const uint32_t N = (uint32_t)G.nper;
    double sum[3] = {0.0, 0.0, 0.0};
    uint32_t coverage = 0;
    for (uint32_t base = 0; base < N; base += 32) {
        const uint32_t k = base + lane;
        double c[3] = {0.0, 0.0, 0.0};
        bool hit = false;
        if (k < N) {
            const uint64_t r = lp * N + k;
            hit = occ[r] != 0;
            if (!hit && G.kind != RI_B200_GATHER_OCCLUSION) {
                double rad[3] = {G.rad[0], G.rad[1], G.rad[2]};
                if (G.kind == RI_B200_GATHER_IBL) {
                    const double dir[3] = {rays[6 * r + 3], rays[6 * r + 4], rays[6 * r + 5]};
                    ibl_fetch_dev(G.env, dir, rad);
                }
                const double brdf = (G.qmc && G.kind == RI_B200_GATHER_IBL) ? (double)0.31831f : 1.0 / 3.14159265358979323846;   // ibl.c:139
#pragma unroll
                for (int q = 0; q < 3; ++q) c[q] = rad[q] * brdf;
            }
        }
        if (G.kind == RI_B200_GATHER_OCCLUSION) {
            coverage += (uint32_t)__popc(__ballot_sync(0xffffffffu, hit));
        } else {
            for (int l = 0; l < 32; ++l) {
#pragma unroll
                for (int q = 0; q < 3; ++q) sum[q] = sum[q] + __shfl_sync(0xffffffffu, c[q], l);
            }
        }
    }
    if (lane == 0) {
        double *o = out3 + 3 * (p0 + lp);
        if (G.kind == RI_B200_GATHER_OCCLUSION) {
            if (coverage > (uint32_t)G.nsamples) coverage = (uint32_t)G.nsamples;
            o[0] = o[1] = o[2] = (double)(float)((double)coverage / (double)(float)G.nsamples);
        } else if (G.kind == RI_B200_GATHER_IBL) {
            for (int q = 0; q < 3; ++q) o[q] = G.qmc ? sum[q] / (double)G.nsamples : 3.14159265358979323846 * sum[q] / (double)(G.ntheta * G.nphi);
        } else {
            for (int q = 0; q < 3; ++q) o[q] = 3.14159265358979323846 * sum[q] / (double)G.nsamples;
        }
    }
}

}  // namespace b200

// qmc.c:182-260 faure_permutation, one base (host): p_2 = (0,1); odd b: p_{b-1} with the values >= (b-1)/2 raised by one and (b-1)/2
// put in the middle; even b: (2 p_{b/2}, 2 p_{b/2} + 1)
static void faure_perm(int base, int *out)
{
    int tmp[128];
    if (base == 2) { out[0] = 0; out[1] = 1; return; }
    if (base % 2 != 0) {
        const int c = (base - 1) / 2;
        faure_perm(base - 1, tmp);
        for (int j = 0; j < c; ++j) out[j] = (2 * tmp[j] >= base - 1) ? tmp[j] + 1 : tmp[j];
        out[c] = c;
        for (int j = c + 1; j < base; ++j) out[j] = (2 * tmp[j - 1] >= base - 1) ? tmp[j - 1] + 1 : tmp[j - 1];
    } else {
        faure_perm(base / 2, tmp);
        for (int j = 0; j < base / 2; ++j) out[j] = 2 * tmp[j];
        for (int j = base / 2; j < base; ++j) out[j] = out[j - base / 2] + 1;
    }
}

extern "C" int ri_b200_gather_points_f64(ri_b200_accel_t *a, const ri_b200_gather_t *g, const double *points, uint64_t n, double *out3,
                                         uint64_t *nrays_out)
{
    using namespace b200;
    if (!a || !g || (n && (!points || !out3))) return fail("null argument");
    if (g->kind < RI_B200_GATHER_OCCLUSION || g->kind > RI_B200_GATHER_DOME) return fail("bad gather kind");
    if (g->nsamples < 1) return fail("bad sample count");
    if (g->kind == RI_B200_GATHER_IBL && (!g->env_rgba || g->env_width < 1 || g->env_height < 1)) return fail("the IBL gather needs an environment map");
    if (need(a, RI_B200_PREC_F64)) return -1;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = a->stream;

    GatherDev G;
    G.kind = g->kind; G.nsamples = g->nsamples;
    int ntheta = g->kind == RI_B200_GATHER_OCCLUSION ? (int)((float)g->nsamples / 3.0) : (int)(g->nsamples / 3.0);   // shader.c:710, ibl.c:160,326
    ntheta = (int)std::sqrt((double)ntheta);
    if (ntheta < 1) ntheta = 1;
    if (g->kind != RI_B200_GATHER_OCCLUSION && ntheta > 128) ntheta = 128;        // MAX_HEMISAMPLE, ibl.h:20
    G.ntheta = ntheta; G.nphi = 3 * ntheta;
    G.stream_offset = g->stream_offset;
    for (int q = 0; q < 3; ++q) G.rad[q] = g->col[q] * g->intensity;
    G.env.data = nullptr; G.env.width = g->env_width; G.env.height = g->env_height; G.env.st = nullptr; G.env.flags = nullptr; G.env.texcol = nullptr;
    G.qmc = g->use_qmc ? 1 : 0; G.qmc_base = 3; G.instance = nullptr;
    if (G.qmc) {
        if (g->kind == RI_B200_GATHER_OCCLUSION) return fail("the occlusion() shadeop has no quasi-Monte Carlo branch");
        static const int primes[25] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97};   // qmc.c:20-31
        const int dim = g->qmc_dim < 1 ? 1 : g->qmc_dim;
        if (dim > 24) return fail("qmc_dim %d: primes[dim] is beyond lucille's 100-entry permutation table", g->qmc_dim);
        G.qmc_base = primes[dim];
        faure_perm(3, G.perm3); faure_perm(5, G.perm5); faure_perm(G.qmc_base, G.permd);
    }
    G.nper = G.qmc ? g->nsamples : G.ntheta * G.nphi;
    const uint64_t N = (uint64_t)G.nper;
    if (nrays_out) *nrays_out = N * n;
    if (!n) return 0;

    const int cap = stack_capacity(a);
    const size_t smem = (size_t)cap * kBlock * sizeof(uint32_t);
    if (smem > 200 * 1024) return fail("BVH depth %d exceeds the shared-memory traversal stack", cap);
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(point_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    void *p = nullptr;
    if (frame_buf(a, 0, n * 6 * sizeof(double), &p)) return -1;
    double *d_points = (double *)p;
    if (frame_buf(a, 1, n * 3 * sizeof(double), &p)) return -1;
    double *d_out = (double *)p;
    if (G.qmc && g->qmc_instance) {
        if (frame_buf(a, 2, n * sizeof(int32_t), &p)) return -1;
        CUDA_OK(cudaMemcpyAsync(p, g->qmc_instance, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        G.instance = (const int32_t *)p;
    }
    const uint64_t words = G.qmc ? 1 : g->stream_offset + 2 * N * n;
    const uint64_t mt_blocks = (words + kMtN - 1) / kMtN;
    if (frame_buf(a, 5, (mt_blocks * kMtN + 4) * 4, &p)) return -1;
    uint32_t *d_mt = (uint32_t *)p;
    if (g->kind == RI_B200_GATHER_IBL) {
        const size_t bytes = sizeof(float) * 4 * (size_t)g->env_width * g->env_height;
        if (frame_buf(a, 7, bytes, &p)) return -1;
        CUDA_OK(cudaMemcpyAsync(p, g->env_rgba, bytes, cudaMemcpyHostToDevice, st));
        G.env.data = (const float *)p;
    }
    CUDA_OK(cudaMemcpyAsync(d_points, points, n * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!G.qmc && mt_stream_launch(a, g->seed, (uint32_t)((mt_blocks + kMtSegBlocks - 1) / kMtSegBlocks), mt_blocks, d_mt, st)) return -1;
    const char *force = getenv("B200_FUSED_AO_TEST");         // test hook: exercise both paths on the same scene
    const bool fused = force ? atoi(force) != 0 : (a->tree.ntris < 4096);
    if (fused) {                                              // one warp per point, one-ray-per-lane traversal
        const uint64_t blocks = (n * 32 + kBlock - 1) / kBlock;
        if (blocks > 0x7fffffffull) return fail("too many shading points in one gather");
        point_gather_kernel<<<(unsigned)blocks, kBlock, smem, st>>>(make_view<double>(a), G, d_points, n, d_mt, d_out, (uint32_t)cap);
        LAUNCHED();
    } else {                                                  // wavefront: <= 2^24 rays at a time through the pooled traverser
        const uint64_t chunk_points = (wave_rays() / N) ? wave_rays() / N : 1;
        const uint64_t buf_points = n < chunk_points ? n : chunk_points;
        if (frame_buf(a, 10, buf_points * N * 6 * sizeof(double), &p)) return -1;
        double *d_rays = (double *)p;
        if (frame_buf(a, 11, buf_points * N + 64, &p)) return -1;
        uint8_t *d_occ8 = (uint8_t *)p;
        for (uint64_t p0 = 0; p0 < n; p0 += chunk_points) {
            const uint64_t np = (n - p0) < chunk_points ? (n - p0) : chunk_points, nr = np * N;
            gather_gen_kernel<<<(unsigned)((nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(G, d_points, p0, nr, d_mt, d_rays);
            LAUNCHED();
            if (launch_trace<double, true, false>(a, d_rays, nr, nullptr, d_occ8, nullptr, st)) return -1;
            gather_accum_kernel<<<(unsigned)((np * 32 + kBlock - 1) / kBlock), kBlock, 0, st>>>(G, d_rays, d_occ8, p0, np, d_out);
            LAUNCHED();
        }
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(out3, d_out, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

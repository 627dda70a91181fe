// The two shading-language callers of ri_raytrace, as batched queries.  In lucille both run inside a surface
// shader, one call per shading point; what they do between the shader's lines is a ray query plus arithmetic, and that is what a
// batch of shading points hands to the device:
//
//   trace(status, dst, P, R)                  render/shader.c:895-976    ri_b200_shade_trace_f64
//       ray = (P + 0.0001 R, R) with R NOT normalised, closest hit; on a miss dst = the environment along R when the scene's first
//       light is an IBL / sun-sky light, else zero; on a hit the input block of the hit geometry's shader procedure is filled from
//       the hit state (Cs, P, N = Ns, Ng, dPdu, dPdv, I = normalize(P_hit - P), s = u, t = v) and the procedure -- a host function
//       pointer -- is called.  The batch returns that block per point; the procedure call stays with the host.
//   next_lightsource(status, P, N, angle)     render/shader.c:1116-1186 + init_lightsource 1236-1310    ri_b200_light_samples_f64
//       the samples an `illuminance` loop visits at a shading point: m = ntheta * 3 ntheta stratified cosine directions about N
//       (two words of the randomMT stream per sample, all m drawn before the first ray), L = normalize(direction),
//       Cl = environment(L) / m; returned in order are the samples inside the cone (dot(L, N) > 0, acos(dot) < angle) that no
//       triangle occludes along normalize(L) from P + 0.0001 N -- never the LAST sample of the set (shader.c:1170-1177: after the
//       break sample_index >= nsamples holds and the function returns NULL).
//
// Both are wavefronts over the pooled fp64 traversers (pool_closest.cuh / pool.cuh): a set-up kernel writes the ray batch, the
// traverser answers it, a finishing kernel does the per-point arithmetic in the reference's order.
#pragma once

namespace b200 {

__global__ void __launch_bounds__(kBlock)
shade_trace_rays_kernel(const double *__restrict__ pr, const uint64_t n, double *__restrict__ rays_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const double P[3] = {pr[6 * i], pr[6 * i + 1], pr[6 * i + 2]}, R[3] = {pr[6 * i + 3], pr[6 * i + 4], pr[6 * i + 5]};
    double org[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) org[k] = P[k] + 0.0001 * R[k];              // shader.c:913-918 (ri_vector_copy, then += 0.0001 * dir)
    double2 *o = reinterpret_cast<double2 *>(rays_out) + 3 * i;
    o[0] = make_double2(org[0], org[1]);
    o[1] = make_double2(org[2], R[0]);
    o[2] = make_double2(R[1], R[2]);
}

__global__ void __launch_bounds__(kBlock)
shade_trace_finish_kernel(const SceneView<double> S, const double *__restrict__ col, const uint8_t *__restrict__ flags,
                          const double *__restrict__ pr, const double *__restrict__ rays, const ri_b200_hit_f64 *__restrict__ hits,
                          const uint64_t n, const TexDev env, ri_b200_trace_rec_f64 *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    ri_b200_trace_rec_f64 r;
    memset(&r, 0, sizeof(r));
    const ri_b200_hit_f64 h = hits[i];
    double org[3], dir[3];
    RayIO<double>::load(rays, i, org, dir);
    r.hit = (int32_t)h.hit;
    r.prim = h.hit ? h.prim : RI_B200_MISS_PRIM;
    if (h.hit) {
        ri_b200_state_f64 st;
        state_from_hit(S.tris, S.slot_of_prim, org, dir, h.t, h.prim, st, S.normals, h.u, h.v);
        const uint8_t fl = flags ? flags[h.prim] : 0;
        if (fl & 1) {                                                       // ri_lerp_vector, geometric.c:40-62
            const double *c = col + 9 * (size_t)h.prim;
            const double w0 = 1.0 - h.u - h.v;
            for (int k = 0; k < 3; ++k) { const double x = c[k] * w0, y = c[3 + k] * h.u, z = c[6 + k] * h.v; r.Cs[k] = (x + y) + z; }
        } else {
            for (int k = 0; k < 3; ++k) r.Cs[k] = 1.0;                      // intersection_state.c:204-207
        }
        double eye[3];
        for (int k = 0; k < 3; ++k) {
            r.P[k] = st.P[k]; r.N[k] = st.Ns[k]; r.Ng[k] = st.Ng[k]; r.dPdu[k] = st.tangent[k]; r.dPdv[k] = st.binormal[k];
            eye[k] = st.P[k] - pr[6 * i + k];                               // shader.c:960-961: I = state.P - P, normalised
        }
        normalize3(eye);
        for (int k = 0; k < 3; ++k) r.I[k] = eye[k];
        r.t = h.t;
        r.s = (float)h.u; r.tt = (float)h.v;                                // shader.c:967-968 (RtFloat)
    } else if (env.data) {
        double rad[3];
        ibl_fetch_dev(env, dir, rad);                                       // shader.c:927-940
        for (int k = 0; k < 3; ++k) r.Ci[k] = rad[k];
    }
    out[i] = r;
}

struct LightDev {
    int      ntheta, nphi, m;
    double   angle;
    uint64_t stream_offset;
    TexDev   env;
};

// one lane per (point, sample): direction, colour, and -- for the samples inside the cone -- a shadow ray appended to the batch
__global__ void __launch_bounds__(kBlock)
light_gen_kernel(const LightDev G, const double *__restrict__ points, const uint64_t p0, const uint64_t nsamp,
                 const uint32_t *__restrict__ mt_stream, double *__restrict__ L_out, double *__restrict__ Cl_out,
                 uint8_t *__restrict__ visible, double *__restrict__ rays_out, uint32_t *__restrict__ ids_out, unsigned int *__restrict__ nrays)
{
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    bool shoot = false;
    double org[3], dir[3];
    if (gid < nsamp) {
        const uint32_t m = (uint32_t)G.m;
        const uint64_t p = p0 + gid / m;
        const uint32_t k = (uint32_t)(gid % m);
        const uint32_t j = k / (uint32_t)G.ntheta, i = k - j * (uint32_t)G.ntheta;
        const double P[3] = {points[6 * p], points[6 * p + 1], points[6 * p + 2]};
        const double Nn[3] = {points[6 * p + 3], points[6 * p + 4], points[6 * p + 5]};
        double b0[3], b1[3];
        ortho_basis(b0, b1, Nn);
        const uint64_t w = G.stream_offset + 2 * ((uint64_t)m * p + k);
        const double r0 = (double)mt_stream[w] * 2.3283064365386963e-10;    // random.c:196
        const double r1 = (double)mt_stream[w + 1] * 2.3283064365386963e-10;
        const double theta = sqrt(((double)i + r0) / (double)G.ntheta);     // shader.c:1280-1283
        const double phi = 2.0 * 3.14159265358979323846 * ((double)j + r1) / (double)G.nphi;
        const double lx = cos(phi) * theta, ly = sin(phi) * theta, lz = sqrt(1.0 - theta * theta);
        double L[3], texel[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q) L[q] = lx * b0[q] + ly * b1[q] + lz * Nn[q];
        normalize3(L);
        if (G.env.data) ibl_fetch_dev(G.env, L, texel);
        const double scale = (double)(1.0f / (double)m);                    // shader.c:1298-1302
        const uint64_t o = 3 * (p * m + k);
#pragma unroll
        for (int q = 0; q < 3; ++q) { L_out[o + q] = L[q]; Cl_out[o + q] = texel[q] * scale; }
        visible[p * m + k] = 0;
        const double ndotl = L[0] * Nn[0] + L[1] * Nn[1] + L[2] * Nn[2];
        if (ndotl > 0.0 && acos(ndotl) < G.angle) {                         // shader.c:1141-1150
            shoot = true;
#pragma unroll
            for (int q = 0; q < 3; ++q) { org[q] = P[q] + Nn[q] * 0.0001; dir[q] = L[q]; }
            normalize3(dir);                                                // shader.c:1157
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, shoot);
    if (mask) {
        const unsigned lane = threadIdx.x & 31u, leader = (unsigned)__ffs(mask) - 1u;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(nrays, (unsigned)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (shoot) {
            const unsigned r = base + (unsigned)__popc(mask & ((1u << lane) - 1u));
            double2 *o = reinterpret_cast<double2 *>(rays_out) + 3 * (size_t)r;
            o[0] = make_double2(org[0], org[1]);
            o[1] = make_double2(org[2], dir[0]);
            o[2] = make_double2(dir[1], dir[2]);
            ids_out[r] = (uint32_t)gid;
        }
    }
}

__global__ void __launch_bounds__(kBlock)
light_visible_kernel(const uint32_t *__restrict__ ids, const uint8_t *__restrict__ occ, const unsigned int *__restrict__ nrays,
                     const uint64_t p0, const uint32_t m, uint8_t *__restrict__ visible)
{
    const uint64_t r = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (r >= *nrays) return;
    const uint32_t gid = ids[r];
    const uint32_t k = gid % m;
    if (!occ[r] && k + 1u < m) visible[p0 * m + gid] = 1;                   // the last sample of a set is never returned
}

}  // namespace b200

static void env_view(b200::TexDev &T, const float *d_rgba, int w, int h)
{
    T.data = d_rgba; T.width = w; T.height = h; T.st = nullptr; T.flags = nullptr; T.texcol = nullptr;
}

extern "C" int ri_b200_shade_trace_f64(ri_b200_accel_t *a, const double *pr, uint64_t n, const float *env_rgba, int env_width,
                                       int env_height, ri_b200_trace_rec_f64 *out)
{
    using namespace b200;
    if (!a || (n && (!pr || !out))) return fail("null argument");
    if (env_rgba && (env_width < 1 || env_height < 1)) return fail("bad environment map size");
    if (need(a, RI_B200_PREC_F64)) return -1;
    if (!n) return 0;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = a->stream;
    void *p = nullptr;
    TexDev env;
    env_view(env, nullptr, env_width, env_height);
    if (env_rgba) {
        const size_t bytes = sizeof(float) * 4 * (size_t)env_width * env_height;
        if (frame_buf(a, 7, bytes, &p)) return -1;
        CUDA_OK(cudaMemcpyAsync(p, env_rgba, bytes, cudaMemcpyHostToDevice, st));
        env.data = (const float *)p;
    }
    const uint64_t chunk = n < wave_rays() / 4 ? n : wave_rays() / 4;               // 4 Mi points at a time: 0.3 KB of records each
    if (frame_buf(a, 0, chunk * 6 * sizeof(double), &p)) return -1;
    double *d_pr = (double *)p;
    if (frame_buf(a, 1, chunk * 6 * sizeof(double), &p)) return -1;
    double *d_rays = (double *)p;
    if (frame_buf(a, 2, chunk * sizeof(ri_b200_hit_f64), &p)) return -1;
    ri_b200_hit_f64 *d_hits = (ri_b200_hit_f64 *)p;
    if (frame_buf(a, 3, chunk * sizeof(ri_b200_trace_rec_f64), &p)) return -1;
    ri_b200_trace_rec_f64 *d_out = (ri_b200_trace_rec_f64 *)p;
    for (uint64_t i0 = 0; i0 < n; i0 += chunk) {
        const uint64_t m = (n - i0) < chunk ? (n - i0) : chunk;
        const unsigned blocks = (unsigned)((m + kBlock - 1) / kBlock);
        CUDA_OK(cudaMemcpyAsync(d_pr, pr + 6 * i0, m * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
        shade_trace_rays_kernel<<<blocks, kBlock, 0, st>>>(d_pr, m, d_rays);
        LAUNCHED();
        if (launch_trace<double, false, false>(a, d_rays, m, d_hits, nullptr, nullptr, st)) return -1;
        shade_trace_finish_kernel<<<blocks, kBlock, 0, st>>>(make_view<double>(a), a->d_col, a->d_attr_flags, d_pr, d_rays, d_hits, m, env, d_out);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(out + i0, d_out, m * sizeof(ri_b200_trace_rec_f64), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}

static int light_ntheta(int nsamples)
{
    int ntheta = (int)(nsamples / 3.0);                                       // shader.c:1263-1266
    ntheta = (int)std::sqrt((double)ntheta);
    return ntheta < 1 ? 1 : ntheta;
}

extern "C" int ri_b200_light_samples_count(int nsamples)
{
    const int ntheta = light_ntheta(nsamples);
    return ntheta * 3 * ntheta;
}

extern "C" int ri_b200_light_samples_f64(ri_b200_accel_t *a, const ri_b200_light_t *g, const double *points, uint64_t n, double *L_out,
                                         double *Cl_out, uint8_t *visible_out, uint64_t *nrays_out)
{
    using namespace b200;
    if (!a || !g || (n && (!points || !L_out || !Cl_out || !visible_out))) return fail("null argument");
    if (g->nsamples < 1) return fail("bad sample count");
    if (g->env_rgba && (g->env_width < 1 || g->env_height < 1)) return fail("bad environment map size");
    if (need(a, RI_B200_PREC_F64)) return -1;
    LightDev G;
    G.ntheta = light_ntheta(g->nsamples);
    G.nphi = 3 * G.ntheta;
    G.m = G.ntheta * G.nphi;
    G.angle = g->angle;
    G.stream_offset = g->stream_offset;
    env_view(G.env, nullptr, g->env_width, g->env_height);
    if (nrays_out) *nrays_out = 0;
    if (!n) return G.m;
    const uint64_t m = (uint64_t)G.m;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = a->stream;
    void *p = nullptr;
    if (g->env_rgba) {
        const size_t bytes = sizeof(float) * 4 * (size_t)g->env_width * g->env_height;
        if (frame_buf(a, 7, bytes, &p)) return -1;
        CUDA_OK(cudaMemcpyAsync(p, g->env_rgba, bytes, cudaMemcpyHostToDevice, st));
        G.env.data = (const float *)p;
    }
    if (frame_buf(a, 0, n * 6 * sizeof(double), &p)) return -1;
    double *d_points = (double *)p;
    if (frame_buf(a, 1, n * m * 3 * sizeof(double), &p)) return -1;
    double *d_L = (double *)p;
    if (frame_buf(a, 2, n * m * 3 * sizeof(double), &p)) return -1;
    double *d_Cl = (double *)p;
    if (frame_buf(a, 3, n * m + 64, &p)) return -1;
    uint8_t *d_vis = (uint8_t *)p;
    const uint64_t words = g->stream_offset + 2 * m * n;
    const uint64_t mt_blocks = (words + kMtN - 1) / kMtN;
    if (frame_buf(a, 5, (mt_blocks * kMtN + 4) * 4, &p)) return -1;
    uint32_t *d_mt = (uint32_t *)p;
    const uint64_t chunk_points = (wave_rays() / m) ? wave_rays() / m : 1;
    const uint64_t buf_points = n < chunk_points ? n : chunk_points;
    if (frame_buf(a, 10, buf_points * m * 6 * sizeof(double), &p)) return -1;
    double *d_rays = (double *)p;
    if (frame_buf(a, 11, buf_points * m + 64, &p)) return -1;
    uint8_t *d_occ8 = (uint8_t *)p;
    if (frame_buf(a, 8, buf_points * m * 4 + 64, &p)) return -1;
    uint32_t *d_ids = (uint32_t *)p;
    if (frame_buf(a, 9, 64, &p)) return -1;
    unsigned int *d_nrays = (unsigned int *)p;

    CUDA_OK(cudaMemcpyAsync(d_points, points, n * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (mt_stream_launch(a, g->seed, (uint32_t)((mt_blocks + kMtSegBlocks - 1) / kMtSegBlocks), mt_blocks, d_mt, st)) return -1;
    uint64_t total_rays = 0;
    for (uint64_t p0 = 0; p0 < n; p0 += chunk_points) {
        const uint64_t np = (n - p0) < chunk_points ? (n - p0) : chunk_points, ns = np * m;
        CUDA_OK(cudaMemsetAsync(d_nrays, 0, sizeof(unsigned int), st));
        light_gen_kernel<<<(unsigned)((ns + kBlock - 1) / kBlock), kBlock, 0, st>>>(G, d_points, p0, ns, d_mt, d_L, d_Cl, d_vis, d_rays, d_ids, d_nrays);
        LAUNCHED();
        unsigned int nr = 0;
        CUDA_OK(cudaMemcpyAsync(&nr, d_nrays, sizeof(nr), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        total_rays += nr;
        if (!nr) continue;
        if (launch_trace<double, true, false>(a, d_rays, nr, nullptr, d_occ8, nullptr, st)) return -1;
        light_visible_kernel<<<(unsigned)(((uint64_t)nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(d_ids, d_occ8, d_nrays, p0, (uint32_t)m, d_vis);
        LAUNCHED();
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(L_out, d_L, n * m * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(Cl_out, d_Cl, n * m * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(visible_out, d_vis, n * m, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (nrays_out) *nrays_out = total_rays;
    return G.m;
}

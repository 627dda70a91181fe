// Whitted transport on the device: ri_transport_whitted + trace_whitted (transport/whitted.c:31-151), which
// lucille compiles but does not call from its pixel loop (render.c:800-804).  From the eye hit a chain of refracted rays
// (ri_refract, eta 1.33, render/reflection.c:69-127; total internal reflection -> ri_reflect with its float dot product, :25-49),
// each starting at P + 1e-7 Rd, at most MAX_TRACE_DEPTH 8 bounces; the radiance is the angular-map environment looked up along the
// direction that leaves the scene (ri_texture_ibl_fetch, render/texture.c:238-277) and zero when the chain is cut or there is no
// environment.  One lane per pixel sub-sample; fp64 records.  The acos of the environment lookup is the one libm call: against the
// host it can differ in the last place, so frames are compared to a 1e-9 relative tolerance (identical in practice).
#pragma once

namespace b200 {

__device__ __forceinline__ void refract_dev(double out[3], const double in[3], const double n[3], const double eta)
{
    double cos1 = in[0] * n[0] + in[1] * n[1] + in[2] * n[2], N[3], e = 1.0 / eta;
    if (cos1 < 0.0) { cos1 = -cos1; N[0] = n[0]; N[1] = n[1]; N[2] = n[2]; }
    else { e = eta; N[0] = -(n[0]); N[1] = -(n[1]); N[2] = -(n[2]); }
    double coeff = 1.0 - (e * e) * (1.0 - cos1 * cos1);
    if (coeff <= 0.0) {                                               // total internal reflection: ri_reflect, float dot
        const float dot = (float)(in[0] * n[0] + in[1] * n[1] + in[2] * n[2]);
        const double two_dot = (double)(2 * dot);
        for (int k = 0; k < 3; ++k) { const double nd = n[k] * two_dot; out[k] = in[k] - nd; }
        normalize3(out);
        return;
    }
    coeff = e * cos1 - sqrt(coeff);
    for (int k = 0; k < 3; ++k) out[k] = coeff * N[k] + e * in[k];
    normalize3(out);
}

__device__ __forceinline__ void ibl_fetch_dev(const TexDev &env, const double dir[3], double out[3])
{
    const double pi = 3.1415926535;
    double nd[3] = {dir[0], dir[1], dir[2]};
    normalize3(nd);
    double r = (nd[2] >= -1.0 && nd[2] < 1.0) ? (1.0 / pi) * acos(nd[2]) : 0.0;
    const double norm2 = nd[0] * nd[0] + nd[1] * nd[1];
    if (norm2 > 1.0e-6) r /= sqrt(norm2);
    double u = nd[0] * r, v = nd[1] * r;
    u = 0.5 * u + 0.5;
    v = 0.5 - 0.5 * v;
    texture_fetch_dev(env, u, v, out);
}

__global__ void __launch_bounds__(kBlock)
whitted_kernel(const SceneView<double> S, const FrameDev F, const uint32_t *__restrict__ pixels, const double *__restrict__ jitter,
               const uint64_t nsamples, const TexDev env, double *__restrict__ rad_out, unsigned long long *__restrict__ nrays_out,
               const int hitmask_only)
{
    extern __shared__ uint32_t s_stack[];
    const uint64_t s = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    unsigned long long traced = 0;
    if (s < nsamples) {
        const uint64_t p = s / (uint64_t)F.spp;
        const int sub = (int)(s - p * (uint64_t)F.spp);
        const uint32_t pix = pixels[p];
        double org[3], dir[3], rad[3] = {0.0, 0.0, 0.0}, t, u, v;
        uint32_t prim;
        camera_ray(F, (int)(pix & 0xffffu), (int)(pix >> 16), jitter[2 * sub], jitter[2 * sub + 1], org, dir);
        ++traced;
        bool hit = trace_ray<double, false, false>(S, org, dir, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr);
        if (hitmask_only) {                                               // ri_transport_sample, transport.c:135-160: white on a hit
            if (hit) rad[0] = rad[1] = rad[2] = 1.0;
        } else if (!hit) {
            if (env.data) ibl_fetch_dev(env, dir, rad);
        } else {
            for (int depth = 1; depth <= 8 && hit; ++depth) {             // MAX_TRACE_DEPTH, whitted.c:24,48
                ri_b200_state_f64 st;
                state_from_hit(S.tris, S.slot_of_prim, org, dir, t, prim, st, S.normals, u, v);
                double I[3] = {dir[0], dir[1], dir[2]}, Rd[3];
                normalize3(I);                                            // intersection_state.c:130-131
                refract_dev(Rd, I, st.Ns, 1.33);
                for (int k = 0; k < 3; ++k) { org[k] = st.P[k] + 1.0e-7 * Rd[k]; dir[k] = Rd[k]; }
                ++traced;
                hit = trace_ray<double, false, false>(S, org, dir, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr);
                if (!hit && env.data) ibl_fetch_dev(env, Rd, rad);
            }
        }
        rad_out[3 * s] = rad[0]; rad_out[3 * s + 1] = rad[1]; rad_out[3 * s + 2] = rad[2];
    }
    for (int o = 16; o > 0; o >>= 1) traced += __shfl_xor_sync(0xffffffffu, traced, o);
    if ((threadIdx.x & 31) == 0 && traced) atomicAdd(nrays_out, traced);
}

// ---- wavefront form (scenes past a few thousand triangles): the chains advance one refraction at a time, every generation of rays
// goes through the pooled closest-hit traverser (pool_closest.cuh, fp64 records) and the chains that go on are compacted.
__global__ void __launch_bounds__(kBlock)
eye_rays_kernel(const FrameDev F, const uint32_t *__restrict__ pixels, const double *__restrict__ jitter, const uint64_t nsamples,
                double *__restrict__ rays_out, uint32_t *__restrict__ ids_out, double *__restrict__ rad_out)
{
    const uint64_t s = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (s >= nsamples) return;
    const uint64_t p = s / (uint64_t)F.spp;
    const int sub = (int)(s - p * (uint64_t)F.spp);
    const uint32_t pix = pixels[p];
    double org[3], dir[3];
    camera_ray(F, (int)(pix & 0xffffu), (int)(pix >> 16), jitter[2 * sub], jitter[2 * sub + 1], org, dir);
    double2 *o = reinterpret_cast<double2 *>(rays_out) + 3 * s;
    o[0] = make_double2(org[0], org[1]);
    o[1] = make_double2(org[2], dir[0]);
    o[2] = make_double2(dir[1], dir[2]);
    ids_out[s] = (uint32_t)s;
    rad_out[3 * s] = 0.0; rad_out[3 * s + 1] = 0.0; rad_out[3 * s + 2] = 0.0;
}

// generation `depth` has been traced: a chain that left the scene takes the environment's radiance along its ray; a chain that hit
// something refracts (whitted.c:31-83) and joins generation depth + 1, unless it already is the eighth bounce (MAX_TRACE_DEPTH)
__global__ void __launch_bounds__(kBlock)
whitted_step_kernel(const SceneView<double> S, const TexDev env, const int depth, const int hitmask_only, const uint32_t n,
                    const double *__restrict__ rays, const uint32_t *__restrict__ ids, const ri_b200_hit_f64 *__restrict__ hits,
                    double *__restrict__ rays_next, uint32_t *__restrict__ ids_next, unsigned int *__restrict__ n_next,
                    double *__restrict__ rad_out)
{
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    bool go_on = false;
    double org[3], dir[3];
    uint32_t s = 0;
    if (i < n) {
        s = ids[i];
        RayIO<double>::load(rays, i, org, dir);
        const ri_b200_hit_f64 h = hits[i];
        if (!h.hit) {
            if (!hitmask_only && env.data) {
                double rad[3] = {0.0, 0.0, 0.0};
                ibl_fetch_dev(env, dir, rad);
                rad_out[3 * (uint64_t)s] = rad[0]; rad_out[3 * (uint64_t)s + 1] = rad[1]; rad_out[3 * (uint64_t)s + 2] = rad[2];
            }
        } else if (hitmask_only) {
            rad_out[3 * (uint64_t)s] = 1.0; rad_out[3 * (uint64_t)s + 1] = 1.0; rad_out[3 * (uint64_t)s + 2] = 1.0;
        } else if (depth < 8) {
            ri_b200_state_f64 st;
            state_from_hit(S.tris, S.slot_of_prim, org, dir, h.t, h.prim, st, S.normals, h.u, h.v);
            double I[3] = {dir[0], dir[1], dir[2]}, Rd[3];
            normalize3(I);
            refract_dev(Rd, I, st.Ns, 1.33);
            for (int k = 0; k < 3; ++k) { org[k] = st.P[k] + 1.0e-7 * Rd[k]; dir[k] = Rd[k]; }
            go_on = true;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, go_on);
    if (m) {
        const unsigned lane = threadIdx.x & 31u;
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(n_next, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (go_on) {
            const uint64_t pos = base + (unsigned)__popc(m & ((1u << lane) - 1u));
            double2 *o = reinterpret_cast<double2 *>(rays_next) + 3 * pos;
            o[0] = make_double2(org[0], org[1]);
            o[1] = make_double2(org[2], dir[0]);
            o[2] = make_double2(dir[1], dir[2]);
            ids_next[pos] = s;
        }
    }
}

// render.c:805,820 + bucket_write: box average of the sub-sample radiances, float RGB at row H-1-y
__global__ void resolve_samples_kernel(const FrameDev F, const uint32_t *__restrict__ pixels, uint64_t npixels, const double *__restrict__ rad,
                                       float *__restrict__ rgb)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npixels) return;
    const uint32_t pix = pixels[p];
    const int x = (int)(pix & 0xffffu), y = (int)(pix >> 16);
    double accum[3] = {0.0, 0.0, 0.0};
    for (int sub = 0; sub < F.spp; ++sub)
        for (int q = 0; q < 3; ++q) accum[q] = accum[q] + rad[3 * (p * (uint64_t)F.spp + sub) + q];
    float *dst = rgb + 3 * ((uint64_t)(F.height - y - 1) * F.width + x);
    for (int q = 0; q < 3; ++q) dst[q] = (float)(accum[q] * (1.0 / (double)(F.xsamples * F.ysamples)));
}

}  // namespace b200

static int render_eye_transport(ri_b200_accel_t *a, const ri_b200_frame_t *f, const float *env_rgba, int env_width, int env_height,
                                float *rgb_out, ri_b200_frame_stats_t *stats, int hitmask_only);

extern "C" int ri_b200_render_whitted(ri_b200_accel_t *a, const ri_b200_frame_t *f, const float *env_rgba, int env_width, int env_height,
                                      float *rgb_out, ri_b200_frame_stats_t *stats)
{ return render_eye_transport(a, f, env_rgba, env_width, env_height, rgb_out, stats, 0); }

// ri_transport_sample (transport/transport.c:50-173), the integrator Option "renderer" "method" would select if render.c did not
// hard-wire ambient occlusion: white where the eye ray hits, black elsewhere (area-light geometry, which returns the light's colour,
// is not part of the triangle soup this library receives).
extern "C" int ri_b200_render_sample(ri_b200_accel_t *a, const ri_b200_frame_t *f, float *rgb_out, ri_b200_frame_stats_t *stats)
{ return render_eye_transport(a, f, nullptr, 0, 0, rgb_out, stats, 1); }

static int render_eye_transport(ri_b200_accel_t *a, const ri_b200_frame_t *f, const float *env_rgba, int env_width, int env_height,
                                float *rgb_out, ri_b200_frame_stats_t *stats, int hitmask_only)
{
    if (!a || !f || !rgb_out) return fail("null argument");
    if (f->width < 1 || f->height < 1 || f->width > 65535 || f->height > 65535) return fail("bad frame size");
    if (f->xsamples < 1 || f->ysamples < 1) return fail("bad sample counts");
    if (f->world != 1 || f->rank != 0) return fail("ri_b200_render_whitted renders whole frames (world == 1)");
    if (env_rgba && (env_width < 1 || env_height < 1)) return fail("bad environment size");
    if (need(a, RI_B200_PREC_F64)) return -1;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = a->stream;
    std::vector<uint32_t> pix;
    std::vector<double> jit;
    pixel_order(*f, pix);
    jitter_table(f->xsamples, f->ysamples, jit);
    const uint64_t npix = pix.size();
    const int spp = f->xsamples * f->ysamples;
    const uint64_t nsamples = npix * (uint64_t)spp;
    FrameDev F;
    for (int i = 0; i < 16; ++i) F.c2w[i] = f->c2w[i];
    F.flength_signed = (double)(float)(f->is_rh ? -1.0 : 1.0) * f->flength;
    F.w = (double)f->width; F.h = (double)f->height;
    F.width = f->width; F.height = f->height; F.xsamples = f->xsamples; F.ysamples = f->ysamples;
    F.ntheta = F.nphi = 1; F.spp = spp; F.nao = 1; F.rng_mode = 0; F.seed = 0; F.ao_eps = 0.0;
    const int cap = stack_capacity(a);
    const size_t smem = (size_t)cap * kBlock * sizeof(uint32_t);
    if (smem > 200 * 1024) return fail("BVH depth %d exceeds the shared-memory traversal stack", cap);
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(whitted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *p = nullptr;
    if (frame_buf(a, 0, (npix + 1) * 4 + jit.size() * 8 + 64, &p)) return -1;
    double *d_jit = (double *)p;
    uint32_t *d_pix = (uint32_t *)(d_jit + jit.size());
    if (frame_buf(a, 1, (nsamples + 1) * 3 * sizeof(double) + 16, &p)) return -1;
    unsigned long long *d_nrays = (unsigned long long *)p;
    double *d_rad = (double *)((char *)p + 16);
    const size_t fb_bytes = (size_t)f->width * f->height * 3 * sizeof(float);
    if (frame_buf(a, 6, fb_bytes, &p)) return -1;
    float *d_rgb = (float *)p;
    TexDev env;
    env.data = nullptr; env.width = env_width; env.height = env_height; env.st = nullptr; env.flags = nullptr; env.texcol = nullptr;
    if (env_rgba) {
        if (frame_buf(a, 7, sizeof(float) * 4 * (size_t)env_width * env_height, &p)) return -1;
        CUDA_OK(cudaMemcpyAsync(p, env_rgba, sizeof(float) * 4 * (size_t)env_width * env_height, cudaMemcpyHostToDevice, st));
        env.data = (const float *)p;
    }
    CUDA_OK(cudaEventRecord(a->ev[0], st));
    CUDA_OK(cudaMemcpyAsync(d_jit, jit.data(), jit.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(d_pix, pix.data(), npix * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(d_nrays, 0, 16, st));
    CUDA_OK(cudaMemsetAsync(d_rgb, 0, fb_bytes, st));
    const char *force = getenv("B200_FUSED_AO_TEST");         // test hook: exercise both paths on the same scene
    const bool fused = force ? atoi(force) != 0 : (a->tree.ntris < 4096);
    uint64_t wave_rays = 0;
    if (fused || nsamples == 0) {                             // one lane per chain
        whitted_kernel<<<(unsigned)((nsamples + kBlock - 1) / kBlock), kBlock, smem, st>>>(make_view<double>(a), F, d_pix, d_jit, nsamples, env, d_rad, d_nrays, hitmask_only);
        LAUNCHED();
    } else {                                                  // one generation of rays at a time through the pooled traverser
        if (nsamples >= 0xfffffff0ull) return fail("too many samples for 32-bit chain ids");
        if (frame_buf(a, 8, 2 * nsamples * 6 * sizeof(double), &p)) return -1;
        double *d_rays[2] = {(double *)p, (double *)p + nsamples * 6};
        if (frame_buf(a, 9, 2 * nsamples * sizeof(uint32_t) + 64, &p)) return -1;
        uint32_t *d_ids[2] = {(uint32_t *)p, (uint32_t *)p + nsamples};
        unsigned int *d_cnt = (unsigned int *)((uint32_t *)p + 2 * nsamples);            // [16] live chains per generation
        if (frame_buf(a, 10, nsamples * sizeof(ri_b200_hit_f64), &p)) return -1;
        ri_b200_hit_f64 *d_hits = (ri_b200_hit_f64 *)p;
        CUDA_OK(cudaMemsetAsync(d_cnt, 0, 16 * sizeof(unsigned int), st));
        eye_rays_kernel<<<(unsigned)((nsamples + kBlock - 1) / kBlock), kBlock, 0, st>>>(F, d_pix, d_jit, nsamples, d_rays[0], d_ids[0], d_rad);
        LAUNCHED();
        uint32_t live = (uint32_t)nsamples;
        const SceneView<double> S = make_view<double>(a);
        for (int depth = 0, cur = 0; depth <= 8 && live; ++depth, cur ^= 1) {
            wave_rays += live;
            if (launch_trace<double, false, false>(a, d_rays[cur], live, d_hits, nullptr, nullptr, st)) return -1;
            whitted_step_kernel<<<(live + kBlock - 1) / kBlock, kBlock, 0, st>>>(S, env, depth, hitmask_only, live, d_rays[cur], d_ids[cur], d_hits,
                                                                               d_rays[cur ^ 1], d_ids[cur ^ 1], d_cnt + depth + 1, d_rad);
            LAUNCHED();
            CUDA_OK(cudaMemcpyAsync(a->h_pin, d_cnt + depth + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            CUDA_OK(cudaStreamSynchronize(st));
            live = *(unsigned int *)a->h_pin;
        }
        CUDA_OK(cudaMemcpyAsync(d_nrays, &wave_rays, 8, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));                   // wave_rays is a local
    }
    resolve_samples_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(F, d_pix, npix, d_rad, d_rgb);
    LAUNCHED();
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(a->ev[5], st));
    unsigned long long *h_nrays = (unsigned long long *)a->h_pin;
    CUDA_OK(cudaMemcpyAsync(h_nrays, d_nrays, 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(rgb_out, d_rgb, fb_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0.0f;
        std::memset(stats, 0, sizeof(*stats));
        stats->nrays_primary = nsamples; stats->nrays_ao = *h_nrays - nsamples;
        CUDA_OK(cudaEventElapsedTime(&ms, a->ev[0], a->ev[5])); stats->ms_total = ms;
    }
    return 0;
}

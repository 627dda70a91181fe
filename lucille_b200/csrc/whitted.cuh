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
    whitted_kernel<<<(unsigned)((nsamples + kBlock - 1) / kBlock), kBlock, smem, st>>>(make_view<double>(a), F, d_pix, d_jit, nsamples, env, d_rad, d_nrays, hitmask_only);
    LAUNCHED();
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

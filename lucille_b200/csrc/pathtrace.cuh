// Path-trace transport (BASELINE config 4), fp64 records.
//
// lucille's src/transport/pathtrace.c is a sketch that is not in the reference build, so there is no binary
// to match; what is reproduced is its control flow, with builder-stated inputs: Lambert kd (grey), constant environment
// Le, counter-based RNG, deterministic sin/cos.  One lane per path (pixel sample): eye ray, up to max_vertices-2 bounces
// with Russian roulette on kd, cosine sampling about the UN-flipped geometric normal with the sketch's float-rounded local
// vector, next ray starting AT the hit point (no offset, pathtrace.c:289-290), G *= kd/pi per bounce, then one more
// sampled direction and a visibility ray towards the environment (pathtrace.c:189-244, 246-314, 480-537).
// The CPU restatement is oracle/lucille_oracle.c:orc_render_pathtrace; the two are bit-identical.
#pragma once

namespace b200 {

struct PathDev {
    FrameDev cam;
    int      spp, max_vertices;
    uint32_t seed;
    double   kd, Le;
};

// sin/cos of 2*pi*r: same operations, same order as orc_det_sincos2pi (plain IEEE multiplies and adds; -fmad=false)
__device__ __forceinline__ void det_sincos2pi(double r, double &s_out, double &c_out)
{
    const double q = floor(4.0 * r + 0.5);
    const double y = r - 0.25 * q;
    const double th = 6.283185307179586476925286766559 * y;
    const double z = th * th;
    double sp = -7.6471637318198164759e-13;
    sp = sp * z + 1.6059043836821614599e-10;
    sp = sp * z + -2.5052108385441718775e-08;
    sp = sp * z + 2.7557319223985890653e-06;
    sp = sp * z + -1.9841269841269841270e-04;
    sp = sp * z + 8.3333333333333333333e-03;
    sp = sp * z + -1.6666666666666666667e-01;
    const double sn = th + th * (z * sp);
    double cp = 4.7794773323873852974e-14;
    cp = cp * z + -1.1470745597729724714e-11;
    cp = cp * z + 2.0876756987868098979e-09;
    cp = cp * z + -2.7557319223985890653e-07;
    cp = cp * z + 2.4801587301587301587e-05;
    cp = cp * z + -1.3888888888888888889e-03;
    cp = cp * z + 4.1666666666666666667e-02;
    cp = cp * z + -0.5;
    const double cs = 1.0 + z * cp;
    const int qi = ((int)q) & 3;
    if (qi == 0)      { s_out = sn;  c_out = cs; }
    else if (qi == 1) { s_out = cs;  c_out = -sn; }
    else if (qi == 2) { s_out = -sn; c_out = -cs; }
    else              { s_out = -cs; c_out = sn; }
}

__device__ __forceinline__ double path_uniform(uint32_t seed, uint64_t sid, uint32_t k)
{
    const uint64_t idx = sid * 64u + k;
    return (double)(splitmix64_dev((uint64_t)seed + idx * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
}

// pathtrace.c:480-508 sample_cosweight
__device__ __forceinline__ void path_cosweight(double out[3], const double n[3], double r0, double r1)
{
    double b0[3], b1[3], sn, cs;
    const double cost = sqrt(r0), sint = sqrt(1.0 - r0);
    ortho_basis(b0, b1, n);
    det_sincos2pi(r1, sn, cs);
    const double v0 = (double)(float)(cs * sint), v1 = (double)(float)(sn * sint), v2 = (double)(float)cost;
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = v0 * b0[i] + v1 * b1[i] + v2 * n[i];
}

// one lane per path; sample radiances go to `rad[pixel_in_pass * spp + s]`, summed in sample order by path_resolve_kernel
__global__ void __launch_bounds__(kBlock)
pathtrace_kernel(const SceneView<double> S, const PathDev F, const uint32_t *__restrict__ pixels, const uint64_t pixel0,
                 const uint64_t npaths, double *__restrict__ rad, unsigned long long *__restrict__ ray_counter)
{
    extern __shared__ uint32_t s_stack[];
    uint32_t *stk = s_stack + threadIdx.x;
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    unsigned nrays = 0;
    if (gid < npaths) {
        const uint64_t pl = gid / (uint64_t)F.spp;
        const uint32_t s = (uint32_t)(gid - pl * (uint64_t)F.spp);
        const uint32_t pix = pixels[pixel0 + pl];
        const int x = (int)(pix & 0xffffu), y = (int)(pix >> 16);
        const uint64_t sid = ((uint64_t)y * (uint64_t)F.cam.width + (uint64_t)x) * (uint64_t)F.spp + s;
        uint32_t k = 0;
        double org[3], dir[3], t, u, v, G = 1.0, result;
        uint32_t prim;
        const double jx = path_uniform(F.seed, sid, k++), jy = path_uniform(F.seed, sid, k++);
        camera_ray(F.cam, x, y, jx, jy, org, dir);
        ++nrays;
        if (!trace_ray<double, false, false>(S, org, dir, stk, kBlock, t, u, v, prim, nullptr)) {
            result = F.Le;                                          // background (pathtrace.c:205-209)
        } else {
            ri_b200_state_f64 st;
            state_from_hit(S.tris, S.slot_of_prim, org, dir, t, prim, st);
            const double inv_pi = 1.0 / 3.14159265358979323846;
            int depth = 2;
            while (depth < F.max_vertices) {                        // trace_path
                if (path_uniform(F.seed, sid, k++) > F.kd) break;   // russian roulette
                (void)path_uniform(F.seed, sid, k++);               // lobe pick: always 'D'
                const double r0 = path_uniform(F.seed, sid, k++), r1 = path_uniform(F.seed, sid, k++);
                double out[3];
                path_cosweight(out, st.Ng, r0, r1);
                ++nrays;
                if (!trace_ray<double, false, false>(S, st.P, out, stk, kBlock, t, u, v, prim, nullptr)) break;
                G = G * (F.kd * inv_pi);
                ++depth;
                const double o2[3] = {st.P[0], st.P[1], st.P[2]};
                state_from_hit(S.tris, S.slot_of_prim, o2, out, t, prim, st);
            }
            k += 1;                                                 // connection: lobe pick draw
            const double r0 = path_uniform(F.seed, sid, k), r1 = path_uniform(F.seed, sid, k + 1);
            double out[3];
            path_cosweight(out, st.Ng, r0, r1);
            G = G * (F.kd * inv_pi);
            ++nrays;
            const bool blocked = trace_ray<double, false, false>(S, st.P, out, stk, kBlock, t, u, v, prim, nullptr);
            result = (blocked ? 0.0 : F.Le) * G;
        }
        rad[gid] = result;
    }
    for (int o = 16; o > 0; o >>= 1) nrays += __shfl_xor_sync(0xffffffffu, nrays, o);
    if ((threadIdx.x & 31) == 0 && nrays) atomicAdd(ray_counter, (unsigned long long)nrays);
}

__global__ void path_resolve_kernel(const PathDev F, const uint32_t *__restrict__ pixels, const uint64_t pixel0, const uint64_t npix,
                                    const double *__restrict__ rad, float *__restrict__ rgb, const int packed)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    double sum = 0.0;
    for (int s = 0; s < F.spp; ++s) sum = sum + rad[p * (uint64_t)F.spp + s];      // pathtrace.c:160-167, in sample order
    const uint32_t pix = pixels[pixel0 + p];
    const int x = (int)(pix & 0xffffu), y = (int)(pix >> 16);
    float *dst = packed ? rgb + 3 * (pixel0 + p) : rgb + 3 * ((uint64_t)(F.cam.height - y - 1) * F.cam.width + x);
    const float f = (float)(sum / (double)F.spp);
    dst[0] = f; dst[1] = f; dst[2] = f;
}

}  // namespace b200

static int render_pathtrace_impl(ri_b200_accel *a, const ri_b200_path_frame_t &f, float *d_rgb, cudaStream_t st, int packed,
                                 ri_b200_frame_stats_t *stats)
{
    ri_b200_frame_t tile;
    std::memset(&tile, 0, sizeof(tile));
    tile.width = f.width; tile.height = f.height; tile.bucket_size = f.bucket_size; tile.rank = f.rank; tile.world = f.world;
    std::vector<uint32_t> pix;
    pixel_order(tile, pix);
    const uint64_t npix = pix.size();

    PathDev F;
    for (int i = 0; i < 16; ++i) F.cam.c2w[i] = f.c2w[i];
    F.cam.flength_signed = (double)(float)(f.is_rh ? -1.0 : 1.0) * f.flength;
    F.cam.w = (double)f.width; F.cam.h = (double)f.height; F.cam.width = f.width; F.cam.height = f.height;
    F.cam.xsamples = F.cam.ysamples = F.cam.ntheta = F.cam.nphi = F.cam.spp = F.cam.nao = 1; F.cam.rng_mode = 1; F.cam.seed = f.seed;
    F.spp = f.spp; F.max_vertices = f.max_vertices; F.seed = f.seed; F.kd = f.kd; F.Le = f.Le;

    const int cap = stack_capacity(a);
    const size_t smem = (size_t)cap * kBlock * sizeof(uint32_t);
    if (smem > 200 * 1024) return fail("BVH depth %d exceeds the shared-memory traversal stack", cap);
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(pathtrace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    const uint64_t pix_per_pass = ((1ull << 24) / (uint64_t)f.spp) ? ((1ull << 24) / (uint64_t)f.spp) : 1;
    void *p = nullptr;
    if (frame_buf(a, 0, (npix + 1) * 4, &p)) return -1;
    uint32_t *d_pix = (uint32_t *)p;
    if (frame_buf(a, 1, (pix_per_pass < npix ? pix_per_pass : npix) * (uint64_t)f.spp * sizeof(double) + 8, &p)) return -1;
    double *d_rad = (double *)p;

    CUDA_OK(cudaEventRecord(a->ev[0], st));
    if (npix) CUDA_OK(cudaMemcpyAsync(d_pix, pix.data(), npix * 4, cudaMemcpyHostToDevice, st));
    if (!packed) CUDA_OK(cudaMemsetAsync(d_rgb, 0, (size_t)f.width * f.height * 3 * sizeof(float), st));
    CUDA_OK(cudaMemsetAsync(a->d_counters, 0, sizeof(unsigned long long), st));
    SceneView<double> S = make_view<double>(a);
    for (uint64_t p0 = 0; p0 < npix; p0 += pix_per_pass) {
        const uint64_t np = (npix - p0) < pix_per_pass ? (npix - p0) : pix_per_pass;
        const uint64_t npaths = np * (uint64_t)f.spp;
        pathtrace_kernel<<<(unsigned)((npaths + kBlock - 1) / kBlock), kBlock, smem, st>>>(S, F, d_pix, p0, npaths, d_rad, a->d_counters);
        LAUNCHED();
        path_resolve_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(F, d_pix, p0, np, d_rad, d_rgb, packed);
        LAUNCHED();
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(a->ev[1], st));
    if (stats) {
        unsigned long long nr = 0;
        CUDA_OK(cudaMemcpyAsync(&nr, a->d_counters, sizeof(nr), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaEventSynchronize(a->ev[1]));
        CUDA_OK(cudaStreamSynchronize(st));
        float ms;
        std::memset(stats, 0, sizeof(*stats));
        CUDA_OK(cudaEventElapsedTime(&ms, a->ev[0], a->ev[1]));
        stats->ms_total = ms; stats->nrays_primary = npix * (uint64_t)f.spp; stats->nrays_ao = nr - stats->nrays_primary;
    }
    return 0;
}

static int check_path_frame(const ri_b200_accel *a, const ri_b200_path_frame_t *f)
{
    if (!a || !f) return fail("null argument");
    if (f->width < 1 || f->height < 1 || f->width > 65535 || f->height > 65535) return fail("bad frame size");
    if (f->spp < 1 || f->max_vertices < 2) return fail("bad sample counts");
    if (f->world < 1 || f->rank < 0 || f->rank >= f->world) return fail("bad rank/world");
    return need(a, RI_B200_PREC_F64);
}

extern "C" int ri_b200_render_pathtrace(ri_b200_accel_t *a, const ri_b200_path_frame_t *f, float *rgb_out, ri_b200_frame_stats_t *stats)
{
    if (check_path_frame(a, f)) return -1;
    if (!rgb_out) return fail("null framebuffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    const size_t bytes = (size_t)f->width * f->height * 3 * sizeof(float);
    void *p = nullptr;
    if (frame_buf(a, 6, bytes, &p)) return -1;
    if (render_pathtrace_impl(a, *f, (float *)p, a->stream, 0, stats)) return -1;
    CUDA_OK(cudaMemcpyAsync(rgb_out, p, bytes, cudaMemcpyDeviceToHost, a->stream));
    CUDA_OK(cudaStreamSynchronize(a->stream));
    return 0;
}

extern "C" int ri_b200_render_pathtrace_tiles_dev(ri_b200_accel_t *a, const ri_b200_path_frame_t *f, float *d_packed, void *stream,
                                                  ri_b200_frame_stats_t *stats)
{
    if (check_path_frame(a, f)) return -1;
    if (!d_packed) return fail("null buffer");
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    return render_pathtrace_impl(a, *f, d_packed, stream ? (cudaStream_t)stream : a->stream, 1, stats);
}

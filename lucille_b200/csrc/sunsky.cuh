// Sun-sky gather on the device (row a12): gather_sunsky + contribution_from_sunlight (ambientocclusion.c:153-324) and the sky
// lookup they call per MISSED ray, ri_sunsky_get_sky_rgb (render/sunsky.c:24-38, 136-152, 297-408; render/specrend.c:127-172,
// 366-440).  The lookup is float arithmetic around DOUBLE libm calls in the reference (`float x = sin(a) * ...`): every
// expression below keeps the reference's types, so the only difference to the CPU is the last-ulp behaviour of CUDA's double
// sin/cos/acos/atan2/exp against glibc's, visible after rounding to float in a small fraction of the lookups (tolerance in
// tests/test_gpu_parity.py).  The file is compiled with -fmad=false like everything else.
//
// The host side owns the model set-up (ri_sunsky_init, src/render/sunsky.c:176-295) and hands over what it produced:
// ri_b200_sunsky_t (include/lucille_b200.h).
#pragma once

namespace b200 {

__device__ __forceinline__ float sky_angle_between(float v_theta, float v_phi, float s_theta, float s_phi)   // sunsky.c:24-38
{
    const float cospi = (float)(sin((double)v_theta) * sin((double)s_theta) * cos((double)(s_phi - v_phi))
                                + cos((double)v_theta) * cos((double)s_theta));
    if ((double)cospi > 1.0) return 0.0f;
    if ((double)cospi < -1.0) return (float)3.14159265358979323846;
    return (float)acos((double)cospi);
}

__device__ __forceinline__ float sky_perez(const float *lam, float theta, float gamma, float lvz, float sun_theta)   // sunsky.c:136-152
{
    const float l0 = lam[0], l1 = lam[1], l2 = lam[2], l3 = lam[3], l4 = lam[4];
    const float den = (float)((1.0 + (double)l0 * exp((double)l1)) *
                              (1.0 + (double)l2 * exp((double)(l3 * sun_theta)) + (double)l4 * cos((double)sun_theta) * cos((double)sun_theta)));
    const float num = (float)((1.0 + (double)l0 * exp((double)l1 / cos((double)theta))) *
                              (1.0 + (double)l2 * exp((double)(l3 * gamma)) + (double)l4 * cos((double)gamma) * cos((double)gamma)));
    return lvz * num / den;
}

// specrend.c:366-440: 81 five-nanometre bins, each 10 nm sample used for two of them, float accumulation in bin order
__device__ __forceinline__ void sky_spectrum_to_xyz(const float *__restrict__ cie, const float *spec, float &x, float &y, float &z)
{
    float X = 0.0f, Y = 0.0f, Z = 0.0f;
    for (int i = 0; i < 81; ++i) {
        const float m = spec[i >> 1];
        X += m * cie[3 * i]; Y += m * cie[3 * i + 1]; Z += m * cie[3 * i + 2];
    }
    x = X; y = Y; z = Z;
}

__device__ __forceinline__ void sky_xyz_to_rgb(const float *cs, float xc, float yc, float zc, float rgb[3])      // specrend.c:127-172
{
    const float xr = cs[0], yr = cs[1], zr = 1 - (xr + yr);
    const float xg = cs[2], yg = cs[3], zg = 1 - (xg + yg);
    const float xb = cs[4], yb = cs[5], zb = 1 - (xb + yb);
    const float xw = cs[6], yw = cs[7], zw = 1 - (xw + yw);
    float rx = (yg * zb) - (yb * zg), ry = (xb * zg) - (xg * zb), rz = (xg * yb) - (xb * yg);
    float gx = (yb * zr) - (yr * zb), gy = (xr * zb) - (xb * zr), gz = (xb * yr) - (xr * yb);
    float bx = (yr * zg) - (yg * zr), by = (xg * zr) - (xr * zg), bz = (xr * yg) - (xg * yr);
    float rw = 1.0f, gw = 1.0f, bw = 1.0f;
    if (fabs((double)yw) > 1.0e-48) {
        rw = ((rx * xw) + (ry * yw) + (rz * zw)) / yw;
        gw = ((gx * xw) + (gy * yw) + (gz * zw)) / yw;
        bw = ((bx * xw) + (by * yw) + (bz * zw)) / yw;
    }
    rx = rx / rw; ry = ry / rw; rz = rz / rw;
    gx = gx / gw; gy = gy / gw; gz = gz / gw;
    bx = bx / bw; by = by / bw; bz = bz / bw;
    rgb[0] = (rx * xc) + (ry * yc) + (rz * zc);
    rgb[1] = (gx * xc) + (gy * yc) + (gz * zc);
    rgb[2] = (bx * xc) + (by * yc) + (bz * zc);
}

// sunsky.c:322-408 (y and z swapped on entry, :337-339)
__device__ __noinline__ void sky_rgb_dev(const ri_b200_sunsky_t *__restrict__ K, const float v[3], float rgb[3])
{
    float spec[41];
    float t0 = v[0], t1 = v[2], t2 = v[1];
    if ((double)t2 < 0.0) {                                              // under the horizon: zero spectrum
        for (int i = 0; i < 41; ++i) spec[i] = 0.0f;
    } else {
        if ((double)t2 < 0.001) {
            t2 = (float)0.001;
            const float vlen = (float)sqrt((double)(t0 * t0 + t1 * t1 + t2 * t2));
            t0 /= vlen; t1 /= vlen; t2 /= vlen;
        }
        const float theta = (float)acos((double)t2);
        const float phi = (fabs((double)theta) < 1.0e-6) ? 0.0f : (float)atan2((double)t1, (double)t0);
        const float sun_theta = K->sun_theta;
        const float gamma = sky_angle_between(theta, phi, sun_theta, K->sun_phi);
        const float x = sky_perez(K->perez_x, theta, gamma, K->zenith_x, sun_theta);
        const float y = sky_perez(K->perez_y, theta, gamma, K->zenith_y, sun_theta);
        const float Y = sky_perez(K->perez_Y, theta, gamma, K->zenith_Y, sun_theta);
        // chromaticity_to_spectrum, sunsky.c:297-314
        const double den = 0.0241 + 0.2562 * (double)x - 0.7341 * (double)y;
        const float M1 = (float)((-1.3515 - 1.7703 * (double)x + 5.9114 * (double)y) / den);
        const float M2 = (float)((0.03 - 31.4424 * (double)x + 30.0717 * (double)y) / den);
        for (int i = 0; i < 41; ++i) spec[i] = K->S0[i] + M1 * K->S1[i] + M2 * K->S2[i];
        float lx, ly, lz;
        sky_spectrum_to_xyz(&K->cie[0][0], spec, lx, ly, lz);
        if (fabs((double)ly) < 1.0e-48) ly = 1.0f;
        for (int i = 0; i < 41; ++i) spec[i] = Y * spec[i] / ly;
    }
    float X, Yc, Z;
    sky_spectrum_to_xyz(&K->cie[0][0], spec, X, Yc, Z);
    sky_xyz_to_rgb(K->cs, X, Yc, Z, rgb);
}

__global__ void sky_rgb_kernel(const ri_b200_sunsky_t *__restrict__ K, const float *__restrict__ dirs, const uint64_t n, float *__restrict__ rgb)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v[3] = {dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]};
    float c[3];
    sky_rgb_dev(K, v, c);
    rgb[3 * i] = c[0]; rgb[3 * i + 1] = c[1]; rgb[3 * i + 2] = c[2];
}

// One lane per gather ray (8 x 8 per hit sample, four samples per CTA): direction, occlusion query, sky colour on a miss.  The
// first lane of each sample then adds the 64 colours in the reference's loop order (double += float), shoots the shadow ray
// of every sun light from the same offset origin, and writes Lo = (1/pi) * col / 64 (ambientocclusion.c:316-321).
template <typename Real>
__global__ void __launch_bounds__(kBlock)
sunsky_kernel(const SceneView<Real> S, const FrameDev F, const ri_b200_sunsky_t *__restrict__ K, const uint64_t nrays,
              const Real *__restrict__ records, const uint32_t *__restrict__ rank_sample, const uint32_t *__restrict__ pixels,
              const uint32_t *__restrict__ mt_stream, double *__restrict__ lo_out, const uint32_t stack_cap)
{
    extern __shared__ uint32_t s_stack[];
    float *s_col = reinterpret_cast<float *>(s_stack + (size_t)stack_cap * kBlock);       // [3][kBlock]
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const uint32_t N = 64u;
    const uint32_t rank = (uint32_t)(gid / N), k = (uint32_t)(gid % N);
    const bool active = gid < nrays;
    float c[3] = {0.0f, 0.0f, 0.0f};
    Real org[3] = {Real(0), Real(0), Real(0)};
    if (active) {
        const Real *rec = records + 12 * (uint64_t)rank;
        Real dir[3], t, u, v;
        uint32_t prim;
        org[0] = rec[0]; org[1] = rec[1]; org[2] = rec[2];
        ao_direction<Real>(F, rank, k, rec, rank_sample, pixels, mt_stream, dir);
        if (!trace_ray<Real, true, false>(S, org, dir, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr)) {
            const float vf[3] = {(float)dir[0], (float)dir[1], (float)dir[2]};            // ambientocclusion.c:294-296
            sky_rgb_dev(K, vf, c);
        }
    }
    s_col[threadIdx.x] = c[0]; s_col[kBlock + threadIdx.x] = c[1]; s_col[2 * kBlock + threadIdx.x] = c[2];
    __syncthreads();
    if (active && k == 0) {
        double col[3] = {0.0, 0.0, 0.0};
        for (uint32_t q = 0; q < N; ++q) {                               // occluded rays hold +0.0f: x + 0.0 == x
            col[0] += (double)s_col[threadIdx.x + q];
            col[1] += (double)s_col[kBlock + threadIdx.x + q];
            col[2] += (double)s_col[2 * kBlock + threadIdx.x + q];
        }
        for (int l = 0; l < K->nsun; ++l) {                              // contribution_from_sunlight, ambientocclusion.c:153-199
            Real dir[3] = {(Real)K->sun_dir[l][0], (Real)K->sun_dir[l][1], (Real)K->sun_dir[l][2]}, t, u, v;
            uint32_t prim;
            if (!trace_ray<Real, true, false>(S, org, dir, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr)) {
                col[0] += K->sun_col[l][0]; col[1] += K->sun_col[l][1]; col[2] += K->sun_col[l][2];
            }
        }
        const double nsamples = (double)N, m = (1.0 / 3.14159265358979323846);
        lo_out[3 * (uint64_t)rank] = m * col[0] / nsamples;
        lo_out[3 * (uint64_t)rank + 1] = m * col[1] / nsamples;
        lo_out[3 * (uint64_t)rank + 2] = m * col[2] / nsamples;
    }
}

// ---- dirt-map transport (transport/dirtmap.c:84-221 calculate_dirt(4, 4), 223-293 ri_transport_dirtmap) ----
// One lane per gather ray (4 x 4 per hit sample, sixteen samples per CTA): CLOSEST hit, colour by distance -- black within
// near_clip 0.1, white beyond far_clip 0.5 or on a miss, mix_color in between (white * (1-p) - black * p, p = clamp(1 - (t-0.1)/0.4);
// the reference raises it to the power 1.0f / dirt_gain = 1, which is the identity).  The first lane of each sample adds the 16
// values in loop order and writes Lo = sum / 16 on three channels, times the material-texture colour when there is one
// (dirtmap.c:272-281).
template <typename Real>
__global__ void __launch_bounds__(kBlock)
dirtmap_kernel(const SceneView<Real> S, const FrameDev F, const uint64_t nrays, const Real *__restrict__ records,
               const uint32_t *__restrict__ rank_sample, const uint32_t *__restrict__ pixels, const uint32_t *__restrict__ mt_stream,
               const double *__restrict__ texcol, double *__restrict__ lo_out, const uint32_t stack_cap)
{
    extern __shared__ uint32_t s_stack[];
    double *s_col = reinterpret_cast<double *>(s_stack + (size_t)stack_cap * kBlock);     // [kBlock]
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const uint32_t N = 16u;
    const uint32_t rank = (uint32_t)(gid / N), k = (uint32_t)(gid % N);
    const bool active = gid < nrays;
    double c = 0.0;
    if (active) {
        const Real *rec = records + 12 * (uint64_t)rank;
        Real org[3] = {rec[0], rec[1], rec[2]}, dir[3], t, u, v;
        uint32_t prim;
        ao_direction<Real>(F, rank, k, rec, rank_sample, pixels, mt_stream, dir);
        const double near_clip = 0.1, far_clip = 0.5, dirt_color = 0.0, base_color = 1.0;
        if (trace_ray<Real, false, false>(S, org, dir, s_stack + threadIdx.x, kBlock, t, u, v, prim, nullptr)) {
            const double td = (double)t;
            if (td <= near_clip) c = dirt_color;
            else if (td >= far_clip) c = base_color;
            else {
                double p = 1.0 - ((td - near_clip) / (far_clip - near_clip));
                if (p < 0.0) p = 0.0;
                if (p > 1.0) p = 1.0;
                c = (1.0 - p) * base_color - p * dirt_color;
            }
        } else {
            c = base_color;
        }
    }
    s_col[threadIdx.x] = c;
    __syncthreads();
    if (active && k == 0) {
        double sum = 0.0;
        for (uint32_t q = 0; q < N; ++q) sum = sum + s_col[threadIdx.x + q];
        const double lo = sum / (double)N;
        for (int q = 0; q < 3; ++q) {
            double r = lo;
            if (texcol) r *= texcol[3 * (uint64_t)rank + q];
            lo_out[3 * (uint64_t)rank + q] = r;
        }
    }
}

// ---- wavefront forms of the two gathers for scenes past a few thousand triangles: rays written by ao_gen_kernel go through the
// pooled traversers (pool.cuh / pool_closest.cuh) and these kernels do the per-sample arithmetic afterwards, in the same order.

// shadow ray of sun light l for every hit sample of the chunk, from the offset origin in the record (ambientocclusion.c:176-187)
template <typename Real>
__global__ void sun_rays_kernel(const ri_b200_sunsky_t *__restrict__ K, const Real *__restrict__ records, const uint32_t rank0,
                                const uint32_t nsamples, Real *__restrict__ rays_out)
{
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nsun = (uint32_t)K->nsun;
    if (gid >= (uint64_t)nsamples * nsun) return;
    const uint32_t s = (uint32_t)(gid / nsun), l = (uint32_t)(gid - (uint64_t)s * nsun);
    const Real *rec = records + 12 * (uint64_t)(rank0 + s);
    if (sizeof(Real) == 8) {
        double2 *o = reinterpret_cast<double2 *>(rays_out) + 3 * gid;
        o[0] = make_double2((double)rec[0], (double)rec[1]);
        o[1] = make_double2((double)rec[2], (double)K->sun_dir[l][0]);
        o[2] = make_double2((double)K->sun_dir[l][1], (double)K->sun_dir[l][2]);
    } else {
        float4 *o = reinterpret_cast<float4 *>(rays_out) + 2 * gid;
        o[0] = make_float4((float)rec[0], (float)rec[1], (float)rec[2], 0.0f);
        o[1] = make_float4((float)(Real)K->sun_dir[l][0], (float)(Real)K->sun_dir[l][1], (float)(Real)K->sun_dir[l][2], 1.0e38f);
    }
}

// One warp per hit sample: the 64 occlusion flags of its gather rays are read two per lane, the MISSED rays are dealt densely to the
// lanes (the sky lookup is ~1500 instructions; a third of the rays need it), their colours parked in shared memory by ray number,
// then lanes 0..2 add one channel each in ray order (double += float, occluded rays hold +0.0f), add the unshadowed suns and write Lo.
template <typename Real>
__global__ void __launch_bounds__(kBlock)
sky_accum_kernel(const ri_b200_sunsky_t *__restrict__ K, const Real *__restrict__ rays, const uint8_t *__restrict__ occ,
                 const uint8_t *__restrict__ sun_occ, const uint32_t rank0, const uint32_t nsamples, double *__restrict__ lo_out)
{
    __shared__ float s_col[kBlock / 32][3][64];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t s = blockIdx.x * (kBlock / 32) + warp;
    if (s >= nsamples) return;
    const uint64_t r0 = (uint64_t)s * 64;
    const uint32_t m0 = __ballot_sync(0xffffffffu, occ[r0 + lane] == 0), m1 = __ballot_sync(0xffffffffu, occ[r0 + 32 + lane] == 0);
    for (int q = 0; q < 3; ++q) { s_col[warp][q][lane] = 0.0f; s_col[warp][q][32 + lane] = 0.0f; }
    __syncwarp();
    const uint32_t nmiss = (uint32_t)(__popc(m0) + __popc(m1));
    for (uint32_t i = lane; i < nmiss; i += 32) {
        // i-th missed ray of the sample: i-th set bit of (m1:m0)
        uint32_t k;
        const uint32_t c0 = (uint32_t)__popc(m0);
        if (i < c0) k = (uint32_t)__fns(m0, 0, (int)i + 1);
        else k = 32u + (uint32_t)__fns(m1, 0, (int)(i - c0) + 1);
        float v[3], c[3];
        if (sizeof(Real) == 8) {
            const double *r = reinterpret_cast<const double *>(rays) + 6 * (r0 + k);
            v[0] = (float)r[3]; v[1] = (float)r[4]; v[2] = (float)r[5];                        // ambientocclusion.c:294-296
        } else {
            const float *r = reinterpret_cast<const float *>(rays) + 8 * (r0 + k);
            v[0] = r[4]; v[1] = r[5]; v[2] = r[6];
        }
        sky_rgb_dev(K, v, c);
        s_col[warp][0][k] = c[0]; s_col[warp][1][k] = c[1]; s_col[warp][2][k] = c[2];
    }
    __syncwarp();
    if (lane < 3) {
        double col = 0.0;
        for (int q = 0; q < 64; ++q) col += (double)s_col[warp][lane][q];
        for (int l = 0; l < K->nsun; ++l)                                                     // contribution_from_sunlight
            if (!sun_occ[(uint64_t)s * (uint32_t)K->nsun + l]) col += K->sun_col[l][lane];
        lo_out[3 * (uint64_t)(rank0 + s) + lane] = (1.0 / 3.14159265358979323846) * col / 64.0;
    }
}

// dirt map: one lane per hit sample adds the 16 colours of its gather rays (closest-hit records from the pooled traverser) in order
template <typename Real>
__global__ void dirt_accum_kernel(const typename RayIO<Real>::Hit *__restrict__ hits, const double *__restrict__ texcol, const uint32_t rank0,
                                  const uint32_t nsamples, double *__restrict__ lo_out)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsamples) return;
    const double near_clip = 0.1, far_clip = 0.5, dirt_color = 0.0, base_color = 1.0;
    double sum = 0.0;
    for (uint32_t q = 0; q < 16u; ++q) {
        const typename RayIO<Real>::Hit h = hits[(uint64_t)s * 16 + q];
        double c = base_color;
        if (h.prim != 0xffffffffu) {
            const double td = (double)h.t;
            if (td <= near_clip) c = dirt_color;
            else if (td >= far_clip) c = base_color;
            else {
                double p = 1.0 - ((td - near_clip) / (far_clip - near_clip));
                if (p < 0.0) p = 0.0;
                if (p > 1.0) p = 1.0;
                c = (1.0 - p) * base_color - p * dirt_color;
            }
        }
        sum = sum + c;
    }
    const double lo = sum / 16.0;
    const uint64_t rank = (uint64_t)rank0 + s;
    for (int q = 0; q < 3; ++q) lo_out[3 * rank + q] = texcol ? lo * texcol[3 * rank + q] : lo;
}

// render.c:805,820 + bucket_write: three-channel box average of the sub-sample radiances, float RGB at row H-1-y
__global__ void resolve_rgb_kernel(const FrameDev F, const uint32_t *__restrict__ pixels, uint64_t npixels,
                                   const uint32_t *__restrict__ sample_rank, const double *__restrict__ lo, float *__restrict__ rgb, const int packed)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npixels) return;
    const uint32_t pix = pixels[p];
    const int x = (int)(pix & 0xffffu), y = (int)(pix >> 16);
    double accum[3] = {0.0, 0.0, 0.0};
    for (int sub = 0; sub < F.spp; ++sub) {
        const uint32_t r = sample_rank[p * (uint64_t)F.spp + sub];
        for (int q = 0; q < 3; ++q) accum[q] = accum[q] + ((r != 0xffffffffu) ? lo[3 * (uint64_t)r + q] : 0.0);
    }
    float *dst = packed ? rgb + 3 * p : rgb + 3 * ((uint64_t)(F.height - y - 1) * F.width + x);
    for (int q = 0; q < 3; ++q) dst[q] = (float)(accum[q] * (1.0 / (double)(F.xsamples * F.ysamples)));
}

}  // namespace b200

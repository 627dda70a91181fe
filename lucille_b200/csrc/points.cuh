// calculate_occlusion (transport/ambientocclusion.c:42-151) as ONE batched call: shading points (P, Ns) in, number of occluded
// gather rays per point out.  The rays are generated on the device, traced by the pooled occlusion traverser (pool.cuh) and
// counted per point by the traverser itself -- the host moves 48 bytes per point in and 4 bytes per point out instead of
// 32 bytes per ray in and 1 byte per ray out (ri_b200_occluded_batch_f32).
//
// Ray set-up per point, exactly the reference's (ambientocclusion.c:56-117): origin P + eps * Ns, basis = ri_ortho_basis(Ns)
// (reflection.c:312-333), for j < nphi, for i < ntheta: z0 = (i + u0) / ntheta, z1 = (j + u1) / nphi, cos(theta) = sqrt(z0),
// phi = 2 pi z1, local = (cos(phi) cos(theta), sin(phi) cos(theta), sqrt(1 - cos^2(theta))), dir = sum local[k] * basis[k],
// not normalised.  Two substitutions make a batch order-free and reproducible bit for bit on any device:
//   * u0, u1 come from the counter-based generator of scenes.uniform01 keyed by (seed, point, j, i) instead of the next two
//     words of the one sequential randomMT2 stream (the frame path, frame.cuh, keeps that stream);
//   * sin / cos of 2 pi z1 are det_sincos2pi (pathtrace.cuh): plain IEEE multiplies and adds in a fixed order, |error| < 1e-15,
//     instead of libm, whose last place differs between glibc and CUDA.
// Everything is computed in double (like the reference) and rounded once to the fp32 ray record.  The CPU restatement is
// orc_ao_point_rays_f32 (oracle/lucille_oracle.c); tests compare the generated rays and the per-point counts bit for bit.
#pragma once

namespace b200 {

struct AoPointsDev { int ntheta, nphi; uint64_t seed; double eps; };

// Real = float: fp32 ray records (rounded once); Real = double: the rays in the reference's own precision, [n][6]
template <typename Real>
__global__ void __launch_bounds__(kBlock)
ao_points_gen_kernel(const AoPointsDev G, const double *__restrict__ points, const uint64_t point0, const uint64_t nrays, Real *__restrict__ rays_out)
{
    // the shading frame of a point is shared by its N rays: the first thread of the block that touches a point computes
    // ri_ortho_basis once and parks it in shared memory (a block of 256 rays covers at most 256 points)
    __shared__ double s_b0[kBlock][3], s_b1[kBlock][3];
    const uint64_t gid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const uint32_t N = (uint32_t)(G.ntheta * G.nphi);
    const uint64_t pb = point0 + ((uint64_t)blockIdx.x * kBlock) / N;             // first point of this block
    const bool active = gid < nrays;
    const uint64_t p = point0 + (active ? gid : nrays - 1) / N;
    const uint32_t k = (uint32_t)((active ? gid : nrays - 1) % N);
    const double *pt = points + 6 * p;
    const double n[3] = {pt[3], pt[4], pt[5]};
    const uint32_t local = (uint32_t)(p - pb);
    if (active && (k == 0u || threadIdx.x == 0u)) {
        double b0[3], b1[3];
        ortho_basis(b0, b1, n);                                                   // ambientocclusion.c:65
#pragma unroll
        for (int q = 0; q < 3; ++q) { s_b0[local][q] = b0[q]; s_b1[local][q] = b1[q]; }
    }
    __syncthreads();
    if (!active) return;
    const uint32_t j = k / (uint32_t)G.ntheta, i = k - j * (uint32_t)G.ntheta;     // outer loop j (phi), inner loop i (theta)
    const uint64_t idx = (p * N + k) * 2;
    const double u0 = (double)(splitmix64_dev(G.seed + idx * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
    const double u1 = (double)(splitmix64_dev(G.seed + (idx + 1) * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
    const double z0 = ((double)i + u0) / (double)G.ntheta;                        // ambientocclusion.c:91-92
    const double z1 = ((double)j + u1) / (double)G.nphi;
    const double ct = sqrt(z0);
    double sn, cs;
    det_sincos2pi(z1, sn, cs);
    const double lx = cs * ct, ly = sn * ct, lz = sqrt(1.0 - ct * ct);            // ambientocclusion.c:96-100
    double org[3], dir[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        org[q] = pt[q] + n[q] * G.eps;                                            // ambientocclusion.c:73-75
        dir[q] = lx * s_b0[local][q] + ly * s_b1[local][q] + lz * n[q];           // ambientocclusion.c:106-111
    }
    if (sizeof(Real) == 4) {
        float4 *o = reinterpret_cast<float4 *>(rays_out) + 2 * gid;
        o[0] = make_float4((float)org[0], (float)org[1], (float)org[2], 0.0f);
        o[1] = make_float4((float)dir[0], (float)dir[1], (float)dir[2], 1.0e38f);
    } else {
        double2 *o = reinterpret_cast<double2 *>(rays_out) + 3 * gid;
        o[0] = make_double2(org[0], org[1]);
        o[1] = make_double2(org[2], dir[0]);
        o[2] = make_double2(dir[1], dir[2]);
    }
}

}  // namespace b200

static int ao_points_check(const ri_b200_ao_points_t *g)
{
    if (!g) return fail("null argument");
    if (g->ntheta < 1 || g->nphi < 1 || (int64_t)g->ntheta * g->nphi > (1 << 20)) return fail("bad gather sample counts %d x %d", g->ntheta, g->nphi);
    return 0;
}

// device-resident points -> device-resident counts, asynchronous on `st`.  Up to 2^24 rays are in flight at a time.
template <typename Real>
static int ao_points_run(ri_b200_accel *a, const ri_b200_ao_points_t *g, const double *d_points, uint64_t n, uint32_t *d_counts, cudaStream_t st)
{
    using namespace b200;
    constexpr size_t kRay = RayIO<Real>::kRayStride * sizeof(Real);
    const uint64_t N = (uint64_t)g->ntheta * (uint64_t)g->nphi;
    const AoPointsDev G = {g->ntheta, g->nphi, g->seed, g->eps};
    const uint64_t chunk_points = (wave_rays() / N) ? wave_rays() / N : 1;
    const uint64_t buf_points = n < chunk_points ? n : chunk_points;
    void *p = nullptr;
    if (frame_buf(a, 10, buf_points * N * kRay, &p)) return -1;
    Real *d_rays = (Real *)p;
    CUDA_OK(cudaMemsetAsync(d_counts, 0, n * sizeof(uint32_t), st));
    for (uint64_t p0 = 0; p0 < n; p0 += chunk_points) {
        const uint64_t np = (n - p0) < chunk_points ? (n - p0) : chunk_points, nr = np * N;
        ao_points_gen_kernel<Real><<<(unsigned)((nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(G, d_points, p0, nr, d_rays);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        if (trace_occlusion_locked<Real>(a, d_rays, nr, nullptr, d_counts + p0, (uint32_t)N, st)) return -1;
    }
    return 0;
}

extern "C" int ri_b200_occlusion_points_dev_f32(ri_b200_accel_t *a, const ri_b200_ao_points_t *g, const double *d_points, uint64_t n,
                                                uint32_t *d_occluded, void *stream)
{
    if (need(a, RI_B200_PREC_F32) || ao_points_check(g)) return -1;
    if (n && (!d_points || !d_occluded)) return fail("null argument");
    if (!n) return 0;
    std::lock_guard<std::mutex> lock(a->mu);               // the ray scratch buffer belongs to the accelerator
    CUDA_OK(cudaSetDevice(a->device));
    return ao_points_run<float>(a, g, d_points, n, d_occluded, stream ? (cudaStream_t)stream : a->stream);
}

// the same call in the reference's own precision: double rays, the double reference's answer for every ray (through hybrid.cuh when
// the fp32 records are resident too, else the double kernel)
extern "C" int ri_b200_occlusion_points_dev_f64(ri_b200_accel_t *a, const ri_b200_ao_points_t *g, const double *d_points, uint64_t n,
                                                uint32_t *d_occluded, void *stream)
{
    if (need(a, RI_B200_PREC_F64) || ao_points_check(g)) return -1;
    if (n && (!d_points || !d_occluded)) return fail("null argument");
    if (!n) return 0;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    return ao_points_run<double>(a, g, d_points, n, d_occluded, stream ? (cudaStream_t)stream : a->stream);
}

template <typename Real>
static int ao_points_host(ri_b200_accel_t *a, const ri_b200_ao_points_t *g, const double *points, uint64_t n, uint32_t *occluded_out)
{
    using namespace b200;
    constexpr size_t kRay = RayIO<Real>::kRayStride * sizeof(Real);
    if (need(a, sizeof(Real) == 4 ? RI_B200_PREC_F32 : RI_B200_PREC_F64) || ao_points_check(g)) return -1;
    if (n && (!points || !occluded_out)) return fail("null argument");
    if (!n) return 0;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    // Per chunk of <= 2^24 rays: the points are uploaded in four pieces on the copy stream, the ray generation of a piece starts as
    // soon as its points have landed (event), and ONE traversal launch follows -- only the first quarter of the upload is exposed.
    // (Two half-batch traversals on two streams hid the upload and the generation as well, but paid a second launch's ramp and tail:
    // 16.18 ms against 16.12 for the plain sequence on C3.)
    const uint64_t N = (uint64_t)g->ntheta * (uint64_t)g->nphi;
    const AoPointsDev G = {g->ntheta, g->nphi, g->seed, g->eps};
    const uint64_t chunk_points = (wave_rays() / N) ? wave_rays() / N : 1;
    const uint64_t buf_points = n < chunk_points ? n : chunk_points;
    cudaStream_t ks = a->stream, cs = a->copy_stream[0];
    void *p = nullptr;
    if (frame_buf(a, 0, n * 6 * sizeof(double), &p)) return -1;
    double *d_points = (double *)p;
    if (frame_buf(a, 1, n * sizeof(uint32_t), &p)) return -1;
    uint32_t *d_counts = (uint32_t *)p;
    if (frame_buf(a, 10, buf_points * N * kRay, &p)) return -1;
    Real *d_rays = (Real *)p;
    CUDA_OK(cudaEventRecord(a->ev[6], ks));                       // whatever the accelerator's stream was doing comes first
    CUDA_OK(cudaStreamWaitEvent(cs, a->ev[6], 0));
    CUDA_OK(cudaMemsetAsync(d_counts, 0, n * sizeof(uint32_t), ks));
    for (uint64_t p0 = 0; p0 < n; p0 += chunk_points) {
        const uint64_t np = (n - p0) < chunk_points ? (n - p0) : chunk_points;
        const int pieces = np * N >= (1ull << 20) ? 4 : 1;
        for (int k = 0; k < pieces; ++k) {
            const uint64_t q0 = p0 + np * (uint64_t)k / pieces, q1 = p0 + np * (uint64_t)(k + 1) / pieces, nr = (q1 - q0) * N;
            if (q1 == q0) continue;
            CUDA_OK(cudaMemcpyAsync(d_points + 6 * q0, points + 6 * q0, (q1 - q0) * 6 * sizeof(double), cudaMemcpyHostToDevice, cs));
            CUDA_OK(cudaEventRecord(a->ev[k & 3], cs));
            CUDA_OK(cudaStreamWaitEvent(ks, a->ev[k & 3], 0));
            ao_points_gen_kernel<Real><<<(unsigned)((nr + kBlock - 1) / kBlock), kBlock, 0, ks>>>(G, d_points, q0, nr,
                                                                                                   d_rays + (q0 - p0) * N * RayIO<Real>::kRayStride);
            LAUNCHED();
            CUDA_OK(cudaGetLastError());
        }
        if (trace_occlusion_locked<Real>(a, d_rays, np * N, nullptr, d_counts + p0, (uint32_t)N, ks)) return -1;
        if (p0 + chunk_points < n) {                              // the next chunk's generation overwrites the ray buffer: wait for this traversal
            CUDA_OK(cudaEventRecord(a->ev[7], ks));
            CUDA_OK(cudaStreamWaitEvent(cs, a->ev[7], 0));
        }
    }
    CUDA_OK(cudaMemcpyAsync(occluded_out, d_counts, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ks));
    CUDA_OK(cudaStreamSynchronize(ks));
    return 0;
}

extern "C" int ri_b200_occlusion_points_f32(ri_b200_accel_t *a, const ri_b200_ao_points_t *g, const double *points, uint64_t n, uint32_t *occluded_out)
{ return ao_points_host<float>(a, g, points, n, occluded_out); }
extern "C" int ri_b200_occlusion_points_f64(ri_b200_accel_t *a, const ri_b200_ao_points_t *g, const double *points, uint64_t n, uint32_t *occluded_out)
{ return ao_points_host<double>(a, g, points, n, occluded_out); }

// the generated batch itself, [n * ntheta * nphi][8] fp32 ray records on the HOST: what the call above traces (tests; resident-ray benchmarks)
extern "C" int ri_b200_ao_point_rays_f32(ri_b200_accel_t *a, const ri_b200_ao_points_t *g, const double *points, uint64_t n, float *rays_out)
{
    using namespace b200;
    if (need(a, RI_B200_PREC_F32) || ao_points_check(g)) return -1;
    if (n && (!points || !rays_out)) return fail("null argument");
    if (!n) return 0;
    std::lock_guard<std::mutex> lock(a->mu);
    CUDA_OK(cudaSetDevice(a->device));
    cudaStream_t st = a->stream;
    const uint64_t N = (uint64_t)g->ntheta * (uint64_t)g->nphi;
    const AoPointsDev G = {g->ntheta, g->nphi, g->seed, g->eps};
    const uint64_t chunk_points = (wave_rays() / N) ? wave_rays() / N : 1;
    const uint64_t buf_points = n < chunk_points ? n : chunk_points;
    void *p = nullptr;
    if (frame_buf(a, 0, n * 6 * sizeof(double), &p)) return -1;
    double *d_points = (double *)p;
    if (frame_buf(a, 10, buf_points * N * 8 * sizeof(float), &p)) return -1;
    float *d_rays = (float *)p;
    CUDA_OK(cudaMemcpyAsync(d_points, points, n * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    for (uint64_t p0 = 0; p0 < n; p0 += chunk_points) {
        const uint64_t np = (n - p0) < chunk_points ? (n - p0) : chunk_points, nr = np * N;
        ao_points_gen_kernel<float><<<(unsigned)((nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(G, d_points, p0, nr, d_rays);
        LAUNCHED();
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(rays_out + p0 * N * 8, d_rays, nr * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

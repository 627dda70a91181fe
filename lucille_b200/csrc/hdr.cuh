// The output step right after the path: lucille's Radiance .hdr display driver on the device, so that a frame
// rendered into device memory leaves the GPU as the finished file image instead of a float framebuffer.
//   hdr_dd_write   display/hdrdrv.c:62-88    clamp negative components to 0, add onto the zeroed buffer
//   float2rgbe     imageio/rgbe.c:78-96       v = frexp(max) * 256 / max in double rounded to float; bytes by truncation
//   RGBE_WritePixels_RLE / RGBE_WriteBytes_RLE  rgbe.c:296-340, 244-294   per scanline: 2 2 hi lo, then the four channel planes
//                                               run-length coded one after the other; flat pixels for widths < 8 or > 0x7fff
//   RGBE_WriteHeader  rgbe.c:117-139
// hdr_rgbe_kernel: one lane per pixel.  hdr_rle_kernel: one WARP per (scanline, channel) row, a parallel run detector (three warp
// scans per 32 bytes) that emits what the reference's sequential coder emits.  hdr_offsets_kernel: one block scans the row lengths.  hdr_pack_kernel:
// one CTA per row copies it to its place.  Byte-identical to the file the reference writes (tests/test_gpu_parity.py).
#pragma once

namespace b200 {

__device__ __forceinline__ void hdr_float2rgbe(unsigned char rgbe[4], float red, float green, float blue)
{
    float v = red;
    if (green > v) v = green;
    if (blue > v) v = blue;
    if ((double)v < 1e-32) {
        rgbe[0] = rgbe[1] = rgbe[2] = rgbe[3] = 0;
    } else {
        int e;
        v = (float)(frexp((double)v, &e) * 256.0 / (double)v);
        rgbe[0] = (unsigned char)(red * v);
        rgbe[1] = (unsigned char)(green * v);
        rgbe[2] = (unsigned char)(blue * v);
        rgbe[3] = (unsigned char)(e + 128);
    }
}

// planes: [height][4][width] bytes (flat == 0) or the pixel-interleaved rgbe stream itself (flat == 1)
__global__ void hdr_rgbe_kernel(const float *__restrict__ rgb, int width, int height, unsigned char *__restrict__ planes, int flat)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)width * height) return;
    const int y = (int)(i / width), x = (int)(i - (size_t)y * width);
    float acc[3];
    for (int k = 0; k < 3; ++k) {                                    // hdr_dd_write
        float c = rgb[3 * i + k];
        if ((double)c < 0.0) c = 0.0f;
        acc[k] = 0.0f;
        acc[k] += c;
    }
    unsigned char q[4];
    hdr_float2rgbe(q, acc[0], acc[1], acc[2]);
    if (flat) {
        for (int k = 0; k < 4; ++k) planes[4 * i + k] = q[k];
    } else {
        unsigned char *row = planes + (size_t)y * 4 * width;
        for (int k = 0; k < 4; ++k) row[(size_t)k * width + x] = q[k];
    }
}

// The run-length code of one channel row (what RGBE_WriteBytes_RLE, rgbe.c:244-294, produces) as a PARALLEL run detector: one warp
// per row, 32 bytes per step, three warp scans per step and no sequential coder.  The reference's output is a function of the row's
// PIECES -- maximal runs of equal bytes cut every 127 bytes from the run's start -- which are LONG (>= 4 bytes) or short:
//   * a long piece of l bytes becomes [128 + l, value];
//   * a maximal stretch of short pieces between long pieces (or the row's ends) becomes literal packets [count <= 128, bytes...],
//     except that a stretch made of exactly ONE piece of 2 or 3 bytes becomes [128 + l, value] (rgbe.c:268-273).
// Every byte decides what it emits from left context carried by scans (start of its run, end of the last long piece, bytes
// emitted so far) and at most six bytes of lookahead: a long piece is written by its LAST byte (which knows l), a literal packet's
// count by the packet's last byte, at the offset it derives from its own.
__device__ __forceinline__ bool hdr_piece_is_long(const unsigned char *d, int n, int ps)      // the piece starting at ps has >= 4 bytes
{ return ps + 3 < n && d[ps + 1] == d[ps] && d[ps + 2] == d[ps] && d[ps + 3] == d[ps]; }

__global__ void hdr_rle_kernel(const unsigned char *__restrict__ planes, int width, int nrows, unsigned char *__restrict__ tmp, uint32_t row_cap,
                               uint32_t *__restrict__ row_len)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int r = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31u);
    if (r >= nrows) return;                                                      // whole warps leave together
    const unsigned char *d = planes + (size_t)r * width;
    unsigned char *out = tmp + (size_t)r * row_cap;
    const int n = width;
    int carry_s = 0, carry_a = 0;                                                // run start / end of the last long piece, so far
    uint32_t carry_off = 0;                                                      // bytes emitted so far
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const bool in = i < n;
        const unsigned char v = in ? d[i] : 0;
        // (1) start of my maximal run: inclusive max-scan of head positions
        const bool head = in && (i == 0 || d[i - 1] != v);
        int s = head ? i : carry_s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, s, o); if (lane >= o && t > s) s = t; }
        const int k = (i - s) % 127, ps = i - k;                                 // my place in my piece, my piece's start
        const bool last_of_piece = in && (k == 126 || i + 1 == n || d[i + 1] != v);
        const bool lng = in && hdr_piece_is_long(d, n, ps);
        // (2) end of the last long piece before me: exclusive max-scan of (j + 1) over last bytes j of long pieces
        int a_inc = (lng && last_of_piece) ? i + 1 : carry_a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, a_inc, o); if (lane >= o && t > a_inc) a_inc = t; }
        int a = __shfl_up_sync(FULL, a_inc, 1);
        if (lane == 0) a = carry_a;
        // (3) what I emit
        uint32_t c = 0;
        bool single = false;
        int la = 0;
        const int rr = i - a;                                                    // my place in the stretch of short pieces
        if (in && !lng) {
            la = 1;                                                              // length of the (short) piece that opens the stretch
            while (la < 3 && a + la < n && d[a + la] == d[a]) ++la;
            single = la >= 2 && (a + la == n || hdr_piece_is_long(d, n, a + la));
            c = single ? (i == a ? 2u : 0u) : (1u + ((rr & 127) == 0 ? 1u : 0u));
        } else if (lng && last_of_piece) {
            c = 2u;
        }
        uint32_t e = c;                                                          // inclusive prefix sum of c
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, e, o); if (lane >= o) e += t; }
        const uint32_t off = carry_off + e - c;
        if (lng && last_of_piece) {
            out[off] = (unsigned char)(128 + k + 1); out[off + 1] = v;
        } else if (in && !lng) {
            if (single) {
                if (i == a) { out[off] = (unsigned char)(128 + la); out[off + 1] = v; }
            } else {
                const int kk = rr & 127;
                const uint32_t pos = off + (kk == 0 ? 1u : 0u);                  // my data byte
                out[pos] = v;
                const bool stretch_ends = i + 1 == n || (last_of_piece && hdr_piece_is_long(d, n, i + 1));
                if (kk == 127 || stretch_ends) out[pos - (uint32_t)kk - 1u] = (unsigned char)(kk + 1);      // the packet's count byte
            }
        }
        carry_s = __shfl_sync(FULL, s, 31);
        carry_a = __shfl_sync(FULL, a_inc, 31);
        carry_off += __shfl_sync(FULL, e, 31);
    }
    if (lane == 0) row_len[r] = carry_off;
}

// row_off[r] = offset of row r in the body (each scanline = 4 header bytes + its four rows); total body size in *body_bytes
__global__ void __launch_bounds__(kScanBlock)
hdr_offsets_kernel(const uint32_t *__restrict__ row_len, int nrows, unsigned long long *__restrict__ row_off, unsigned long long *__restrict__ body_bytes)
{
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nrows; base += kScanBlock) {
        const int r = base + (int)threadIdx.x;
        const uint32_t x = (r < nrows) ? row_len[r] + (((r & 3) == 0) ? 4u : 0u) : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(x, &total);
        if (r < nrows) row_off[r] = carry + ex + (((r & 3) == 0) ? 4u : 0u);      // the row's bytes start after its scanline's header
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *body_bytes = carry;
}

__global__ void hdr_pack_kernel(const unsigned char *__restrict__ tmp, uint32_t row_cap, const uint32_t *__restrict__ row_len,
                                const unsigned long long *__restrict__ row_off, int width, unsigned char *__restrict__ body)
{
    const int r = blockIdx.x;
    const uint32_t n = row_len[r];
    unsigned char *dst = body + row_off[r];
    const unsigned char *src = tmp + (size_t)r * row_cap;
    if ((r & 3) == 0 && threadIdx.x == 0) {                              // rgbe.c:311-314
        dst[-4] = 2; dst[-3] = 2; dst[-2] = (unsigned char)(width >> 8); dst[-1] = (unsigned char)(width & 0xFF);
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace b200

// rgb: [height][width][3] floats in display order, in DEVICE memory when rgb_on_device != 0, else on the host.
// Returns the file size in bytes (header included) or -1; the file image is written to `out` (HOST) when it fits in cap.
extern "C" int64_t ri_b200_hdr_encode(const float *rgb, int width, int height, uint8_t *out, uint64_t cap, int device, int rgb_on_device)
{
    if (!rgb || width < 1 || height < 1) return fail("bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail("no CUDA device: libb200accel has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail("bad device %d", device);
    CUDA_OK(cudaSetDevice(device));
    char hdr[128];
    const int hlen = snprintf(hdr, sizeof(hdr), "#?%s\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", "RGBE", height, width);   // rgbe.c:117-139
    const size_t npix = (size_t)width * height;
    const int flat = (width < 8) || (width > 0x7fff);                    // rgbe.c:303-305
    const int nrows = 4 * height;
    const uint32_t row_cap = (uint32_t)width + (uint32_t)width / 64u + 8u;
    float *d_rgb = nullptr;
    unsigned char *d_planes = nullptr, *d_tmp = nullptr, *d_body = nullptr;
    uint32_t *d_len = nullptr;
    unsigned long long *d_off = nullptr;
    int64_t result = -1;
    auto body = [&]() -> int {
        const float *src = rgb;
        if (!rgb_on_device) {
            CUDA_OK(cudaMalloc((void **)&d_rgb, npix * 3 * sizeof(float)));
            CUDA_OK(cudaMemcpy(d_rgb, rgb, npix * 3 * sizeof(float), cudaMemcpyHostToDevice));
            src = d_rgb;
        }
        CUDA_OK(cudaMalloc((void **)&d_planes, npix * 4));
        hdr_rgbe_kernel<<<(unsigned)((npix + 255) / 256), 256>>>(src, width, height, d_planes, flat);
        LAUNCHED();
        unsigned long long body_bytes = 0;
        const unsigned char *d_src = d_planes;
        if (flat) {
            body_bytes = npix * 4;
        } else {
            CUDA_OK(cudaMalloc((void **)&d_tmp, (size_t)nrows * row_cap));
            CUDA_OK(cudaMalloc((void **)&d_len, (size_t)nrows * sizeof(uint32_t)));
            CUDA_OK(cudaMalloc((void **)&d_off, ((size_t)nrows + 1) * sizeof(unsigned long long)));
            hdr_rle_kernel<<<(nrows + 3) / 4, 128>>>(d_planes, width, nrows, d_tmp, row_cap, d_len);      // one warp per (scanline, channel) row
            LAUNCHED();
            hdr_offsets_kernel<<<1, kScanBlock>>>(d_len, nrows, d_off, d_off + nrows);
            LAUNCHED();
            CUDA_OK(cudaMemcpy(&body_bytes, d_off + nrows, sizeof(body_bytes), cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMalloc((void **)&d_body, body_bytes ? body_bytes : 1));
            hdr_pack_kernel<<<nrows, 128>>>(d_tmp, row_cap, d_len, d_off, width, d_body);
            LAUNCHED();
            d_src = d_body;
        }
        CUDA_OK(cudaGetLastError());
        result = (int64_t)hlen + (int64_t)body_bytes;
        if (out && cap >= (uint64_t)result) {
            std::memcpy(out, hdr, (size_t)hlen);
            CUDA_OK(cudaMemcpy(out + hlen, d_src, body_bytes, cudaMemcpyDeviceToHost));
        }
        return 0;
    };
    const int rc = body();
    cudaFree(d_rgb); cudaFree(d_planes); cudaFree(d_tmp); cudaFree(d_len); cudaFree(d_off); cudaFree(d_body);
    return rc ? -1 : result;
}

// ---- the other display driver of the reference that takes float pixels: the socket driver's wire format (display/sockdrv.c:118-262,
// sockdrv_defs.h).  [COMMAND_NEW=0][8][width][height], one [COMMAND_PIXEL=2][24*1024] message per MAXPACKETS = 1024
// pixels {int x, int y, float r, g, b, 1.0f} in the order bucket_write hands them to the driver (render.c:919-979), [COMMAND_FINISH=1].
// The pixels left over when the frame is not a multiple of 1024 are never sent by the reference (sock_dd_close does not flush its
// packet array) -- reproduced.  This is the stream only; opening the socket and send() stay with the host.
namespace b200 {
__global__ void sock_pack_kernel(const float *__restrict__ rgb, const uint32_t *__restrict__ pixels, const uint64_t nsent, const int width,
                                 const int height, unsigned char *__restrict__ out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nsent) return;
    const uint32_t pix = pixels[q];
    const int x = (int)(pix & 0xffffu), y = height - (int)(pix >> 16) - 1;
    const uint64_t g = q >> 10, k = q & 1023u;
    unsigned char *msg = out + 16 + g * (8 + 24 * 1024);
    if (k == 0) { int32_t *h = reinterpret_cast<int32_t *>(msg); h[0] = 2; h[1] = 24 * 1024; }
    int32_t *o = reinterpret_cast<int32_t *>(msg + 8 + k * 24);
    const float *px = rgb + 3 * ((uint64_t)y * width + x);
    o[0] = x; o[1] = y;
    o[2] = __float_as_int(px[0]); o[3] = __float_as_int(px[1]); o[4] = __float_as_int(px[2]); o[5] = __float_as_int(1.0f);
}
}  // namespace b200

extern "C" int64_t ri_b200_sockdrv_encode(const float *rgb, const ri_b200_frame_t *f, uint8_t *out, uint64_t cap, int device, int rgb_on_device)
{
    using namespace b200;
    if (!rgb || !f || f->width < 1 || f->height < 1 || f->width > 65535 || f->height > 65535) return fail("bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail("no CUDA device: libb200accel has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail("bad device %d", device);
    CUDA_OK(cudaSetDevice(device));
    ri_b200_frame_t whole = *f;
    whole.rank = 0; whole.world = 1;                             // the display sees the whole frame
    std::vector<uint32_t> pix;
    pixel_order(whole, pix);
    const uint64_t npix = pix.size(), ngroups = npix / 1024, nsent = ngroups * 1024;
    const int64_t total = 16 + (int64_t)ngroups * (8 + 24 * 1024) + 4;
    if (!out || cap < (uint64_t)total) return total;
    float *d_rgb = nullptr;
    uint32_t *d_pix = nullptr;
    unsigned char *d_out = nullptr;
    auto body = [&]() -> int {
        const float *src = rgb;
        if (!rgb_on_device) {
            CUDA_OK(cudaMalloc((void **)&d_rgb, npix * 3 * sizeof(float)));
            CUDA_OK(cudaMemcpy(d_rgb, rgb, npix * 3 * sizeof(float), cudaMemcpyHostToDevice));
            src = d_rgb;
        }
        CUDA_OK(cudaMalloc((void **)&d_pix, (nsent + 1) * sizeof(uint32_t)));
        CUDA_OK(cudaMalloc((void **)&d_out, (size_t)total));
        if (nsent) {
            CUDA_OK(cudaMemcpy(d_pix, pix.data(), nsent * sizeof(uint32_t), cudaMemcpyHostToDevice));
            sock_pack_kernel<<<(unsigned)((nsent + 255) / 256), 256>>>(src, d_pix, nsent, f->width, f->height, d_out);
            LAUNCHED();
            CUDA_OK(cudaGetLastError());
        }
        CUDA_OK(cudaMemcpy(out, d_out, (size_t)total, cudaMemcpyDeviceToHost));
        const int32_t head[4] = {0, 8, f->width, f->height}, fin = 1;
        std::memcpy(out, head, 16);
        std::memcpy(out + total - 4, &fin, 4);
        return 0;
    };
    const int rc = body();
    cudaFree(d_rgb); cudaFree(d_pix); cudaFree(d_out);
    return rc ? -1 : total;
}

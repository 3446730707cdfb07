// Point <-> pixel transforms of the distillation student's fusion path:
//   point -> pixel scatter-mean   core/models/fusion_blocks.py:217-238 (Point2Grid) and the multi-scale loop of
//                                 core/models/nuscenes/spvcnn_swiftnet18_spformer_tsd_full.py:448-478
//   pixel -> point bilinear gather with per-camera masks   fusion_blocks.py:241-278 (Feature_Gather / Feature_Fetch),
//                                 spvcnn_swiftnet18_spformer_tsd_full.py:482-494
// The reference runs, per batch element x camera (x scale), torch.unique(dim=0) (a sort + host sync), scatter_add_,
// sparse_coo_tensor().to_dense() and a permute; here one scatter kernel + one normalise/transpose kernel per (scale) call
// over ALL cameras, and one gather kernel.  It is the 2-D sibling of voxelise / devoxelise: HBM-bound, fp32.
//
// Layouts: feats fp32 [N, C]; coord fp32 [V, N, 2] = (x, y) in [-1, 1] (width, height); mask u8 [V, N]; V = cameras of one
// batch element; grids [V, C, H, W] (what the image branch consumes / produces).
#include "u2_common.cuh"

namespace {

__device__ __forceinline__ bool pix_of(const float *coord, int64_t vp, int H, int W, int &px, int &py) {
    // fusion_blocks.py:225-227: u = (x + 1) / 2 * (w - 1), v = (y + 1) / 2 * (h - 1), floor
    const float u = (coord[2 * vp] + 1.0f) / 2 * ((float)W - 1.0f);
    const float v = (coord[2 * vp + 1] + 1.0f) / 2 * ((float)H - 1.0f);
    px = (int)floorf(u);
    py = (int)floorf(v);
    return px >= 0 && px < W && py >= 0 && py < H;
}

// acc [V, H, W, C] += feats rows, cnt [V, H, W] += 1 for every (camera, masked point); G lanes per point (float4 each)
__global__ void __launch_bounds__(256) p2g_scatter_kernel(const float *__restrict__ feats, const float *__restrict__ coord,
                                                          const uint8_t *__restrict__ mask, int64_t N, int C, int V, int H, int W,
                                                          int G, float *__restrict__ acc, int *__restrict__ cnt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t vp = t / G;
    const int g = (int)(t % G);
    if (vp >= (int64_t)V * N || !mask[vp]) return;
    int px, py;
    if (!pix_of(coord, vp, H, W, px, py)) return;
    const int64_t v = vp / N, p = vp % N;
    const int64_t cell = (v * H + py) * W + px;
    if (g == 0) atomicAdd(cnt + cell, 1);
    const float *src = feats + p * C;
    float *dst = acc + cell * C;
    for (int c = g * 4; c < C; c += G * 4) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(src + c));
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
    }
}

// out [V, C, H, W] = acc [V, H, W, C] / max(cnt, 1): 32 x 32 (pixel, channel) tiles through shared memory
__global__ void __launch_bounds__(256) p2g_finish_kernel(const float *__restrict__ acc, const int *__restrict__ cnt, int C, int64_t HW,
                                                         float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int v = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int64_t p = p0 + r;
        const int c = c0 + tx;
        float x = 0.f;
        if (p < HW && c < C) {
            const int n = cnt[v * HW + p];
            x = n > 0 ? acc[(v * HW + p) * C + c] / (float)n : 0.f;
        }
        tile[r][tx] = x;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const int64_t p = p0 + tx;
        if (p < HW && c < C) out[((int64_t)v * C + c) * HW + p] = tile[tx][r];
    }
}

// backward of the scatter-mean: dfeats[p, :] = sum over cameras v with mask[v, p] of dgrid[v, :, py, px] / cnt[v, py, px]
__global__ void __launch_bounds__(256) p2g_bwd_kernel(const float *__restrict__ dgrid, const float *__restrict__ coord,
                                                      const uint8_t *__restrict__ mask, const int *__restrict__ cnt, int64_t N, int C, int V,
                                                      int H, int W, float *__restrict__ dfeats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t p = t / C;
    const int c = (int)(t % C);
    const int64_t HW = (int64_t)H * W;
    float s = 0.f;
    for (int v = 0; v < V; v++) {
        const int64_t vp = (int64_t)v * N + p;
        if (!mask[vp]) continue;
        int px, py;
        if (!pix_of(coord, vp, H, W, px, py)) continue;
        const int64_t cell = (int64_t)py * W + px;
        s += __ldg(dgrid + ((int64_t)v * C + c) * HW + cell) / (float)cnt[v * HW + cell];
    }
    dfeats[t] = s;
}

// fusion_blocks.py:241-278 + tsd_full.py:489-492: the LAST camera whose mask holds the point wins; bilinear sample with
// align_corners=True and zero padding (torch grid_sample semantics: ix = (x + 1) / 2 * (W - 1))
__device__ __forceinline__ int last_camera(const uint8_t *mask, int64_t N, int V, int64_t p) {
    int vs = -1;
    for (int v = 0; v < V; v++)
        if (mask[(int64_t)v * N + p]) vs = v;
    return vs;
}

struct Bilin {
    int x0, y0;
    float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};
__device__ __forceinline__ Bilin bilin(const float *coord, int64_t vp, int H, int W) {
    const float ix = (coord[2 * vp] + 1.f) / 2.f * (float)(W - 1), iy = (coord[2 * vp + 1] + 1.f) / 2.f * (float)(H - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    Bilin b;
    b.x0 = (int)fx; b.y0 = (int)fy;
    const float ax = ix - fx, ay = iy - fy;
    b.w00 = (1.f - ax) * (1.f - ay); b.w01 = ax * (1.f - ay); b.w10 = (1.f - ax) * ay; b.w11 = ax * ay;
    return b;
}

__global__ void __launch_bounds__(256) pix_gather_fwd_kernel(const float *__restrict__ img, const float *__restrict__ coord,
                                                             const uint8_t *__restrict__ mask, int64_t N, int C, int V, int H, int W,
                                                             float *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    // consecutive threads = consecutive POINTS of one channel: the four taps of neighbouring points share cache lines
    const int64_t p = t % N;
    const int c = (int)(t / N);
    const int v = last_camera(mask, N, V, p);
    float s = 0.f;
    if (v >= 0) {
        const Bilin b = bilin(coord, (int64_t)v * N + p, H, W);
        const float *im = img + ((int64_t)v * C + c) * H * W;
        const bool x0 = b.x0 >= 0 && b.x0 < W, x1 = b.x0 + 1 >= 0 && b.x0 + 1 < W, y0 = b.y0 >= 0 && b.y0 < H, y1 = b.y0 + 1 >= 0 && b.y0 + 1 < H;
        if (y0 && x0) s = fmaf(b.w00, __ldg(im + (int64_t)b.y0 * W + b.x0), s);
        if (y0 && x1) s = fmaf(b.w01, __ldg(im + (int64_t)b.y0 * W + b.x0 + 1), s);
        if (y1 && x0) s = fmaf(b.w10, __ldg(im + (int64_t)(b.y0 + 1) * W + b.x0), s);
        if (y1 && x1) s = fmaf(b.w11, __ldg(im + (int64_t)(b.y0 + 1) * W + b.x0 + 1), s);
    }
    out[p * C + c] = s;
}

__global__ void __launch_bounds__(256) pix_gather_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ coord,
                                                             const uint8_t *__restrict__ mask, int64_t N, int C, int V, int H, int W,
                                                             float *__restrict__ dimg) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t p = t % N;
    const int c = (int)(t / N);
    const int v = last_camera(mask, N, V, p);
    if (v < 0) return;
    const float g = __ldg(dout + p * C + c);
    const Bilin b = bilin(coord, (int64_t)v * N + p, H, W);
    float *im = dimg + ((int64_t)v * C + c) * H * W;
    const bool x0 = b.x0 >= 0 && b.x0 < W, x1 = b.x0 + 1 >= 0 && b.x0 + 1 < W, y0 = b.y0 >= 0 && b.y0 < H, y1 = b.y0 + 1 >= 0 && b.y0 + 1 < H;
    if (y0 && x0) atomicAdd(im + (int64_t)b.y0 * W + b.x0, b.w00 * g);
    if (y0 && x1) atomicAdd(im + (int64_t)b.y0 * W + b.x0 + 1, b.w01 * g);
    if (y1 && x0) atomicAdd(im + (int64_t)(b.y0 + 1) * W + b.x0, b.w10 * g);
    if (y1 && x1) atomicAdd(im + (int64_t)(b.y0 + 1) * W + b.x0 + 1, b.w11 * g);
}

}  // namespace

extern "C" size_t u2_point2grid_scratch_bytes(int32_t C, int32_t V, int32_t H, int32_t W) {
    return (size_t)V * H * W * ((size_t)C * sizeof(float) + sizeof(int));
}

// grid [V, C, H, W] = per-pixel mean of the masked points' features (0 where no point falls); counts int32 [V, H, W] is an
// output too (the backward needs it).  scratch: u2_point2grid_scratch_bytes (accumulator in pixel-major layout).
extern "C" int u2_point2grid_fwd(const float *feats, const float *coord, const uint8_t *mask, int64_t N, int32_t C, int32_t V,
                                 int32_t H, int32_t W, float *grid, int32_t *counts, void *scratch, size_t scratch_bytes,
                                 u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(feats && coord && mask && grid && counts && scratch, "u2_point2grid_fwd: null pointer");
    U2_CHECK_ARG(C > 0 && C % 4 == 0 && V > 0 && H > 0 && W > 0, "u2_point2grid_fwd: bad shape C=%d V=%d H=%d W=%d", C, V, H, W);
    U2_CHECK_ARG(scratch_bytes >= (size_t)V * H * W * C * sizeof(float) && (((uintptr_t)feats | (uintptr_t)scratch) & 15) == 0,
                 "u2_point2grid_fwd: scratch too small or misaligned");
    float *acc = (float *)scratch;
    const int64_t HW = (int64_t)H * W;
    U2_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)V * HW * C * sizeof(float), st));
    U2_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)V * HW * sizeof(int), st));
    if (N > 0) {
        int G = 1;
        while (G < 32 && G * 4 < C) G <<= 1;
        const int64_t threads = (int64_t)V * N * G;
        p2g_scatter_kernel<<<(unsigned)u2_ceil_div(threads, 256), 256, 0, st>>>(feats, coord, mask, N, C, V, H, W, G, acc, counts);
        U2_LAUNCH_OK();
    }
    dim3 grid_dim((unsigned)u2_ceil_div(HW, 32), (unsigned)u2_ceil_div(C, 32), (unsigned)V);
    p2g_finish_kernel<<<grid_dim, 256, 0, st>>>(acc, counts, C, HW, grid);
    U2_LAUNCH_OK();
    return 0;
}

extern "C" int u2_point2grid_bwd(const float *dgrid, const float *coord, const uint8_t *mask, const int32_t *counts, int64_t N,
                                 int32_t C, int32_t V, int32_t H, int32_t W, float *dfeats, u2_stream_t stream) {
    U2_CHECK_ARG(dgrid && coord && mask && counts && dfeats, "u2_point2grid_bwd: null pointer");
    if (N == 0) return 0;
    p2g_bwd_kernel<<<(unsigned)u2_ceil_div(N * C, 256), 256, 0, (cudaStream_t)stream>>>(dgrid, coord, mask, counts, N, C, V, H, W, dfeats);
    U2_LAUNCH_OK();
    return 0;
}

// out [N, C]: bilinear sample (align_corners, zero padding) of img [V, C, H, W] at every point's pixel coordinate in the LAST
// camera that sees it; zeros for points no camera sees
extern "C" int u2_pixel_gather_fwd(const float *img, const float *coord, const uint8_t *mask, int64_t N, int32_t C, int32_t V, int32_t H,
                                   int32_t W, float *out, u2_stream_t stream) {
    U2_CHECK_ARG(img && coord && mask && out, "u2_pixel_gather_fwd: null pointer");
    if (N == 0) return 0;
    pix_gather_fwd_kernel<<<(unsigned)u2_ceil_div(N * C, 256), 256, 0, (cudaStream_t)stream>>>(img, coord, mask, N, C, V, H, W, out);
    U2_LAUNCH_OK();
    return 0;
}

// dimg [V, C, H, W] (zeroed here) += bilinear weights x dout [N, C]
extern "C" int u2_pixel_gather_bwd(const float *dout, const float *coord, const uint8_t *mask, int64_t N, int32_t C, int32_t V, int32_t H,
                                   int32_t W, float *dimg, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(dout && coord && mask && dimg, "u2_pixel_gather_bwd: null pointer");
    U2_CUDA_OK(cudaMemsetAsync(dimg, 0, (size_t)V * C * H * W * sizeof(float), st));
    if (N == 0) return 0;
    pix_gather_bwd_kernel<<<(unsigned)u2_ceil_div(N * C, 256), 256, 0, st>>>(dout, coord, mask, N, C, V, H, W, dimg);
    U2_LAUNCH_OK();
    return 0;
}

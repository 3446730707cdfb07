// Point <-> voxel transforms: scatter-mean voxelise (fwd/bwd), trilinear weights,
// devoxelise (fwd/bwd).  Replaces torchsparse.backend.{voxelize,devoxelize}_*_cuda and the
// elementwise chain of spf.calc_ti_weights (see include/u2mkd.h for the call sites).
//
// HBM-bound fp32 work: every feature row is moved with 16-byte vector accesses by a
// sub-warp group of lanes; scatter-adds are run-length segmented inside the group
// (consecutive points of a scan usually fall into the same coarse voxel) and only the
// run totals go out, as one 16-byte vector reduction each.
#include "u2_common.cuh"

__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

static inline int group_lanes(int V) {  // lanes cooperating on one row: pow2 in [1,32]
    int g = 1;
    while (g < V && g < 32) g <<= 1;
    return g;
}

// ------------------------------------------------------------------ voxelize fwd
// A group of G lanes walks PTS consecutive points; lane l owns float4 columns l, l+G, ...
template <int NV>
__global__ void __launch_bounds__(256) voxelize_fwd_kernel(const float4 *__restrict__ feats, int64_t n_pts, int V, int G,
                                                           int pts_per_group, const int *__restrict__ idx,
                                                           const int *__restrict__ counts, float *__restrict__ out,
                                                           int64_t n_vox) {
    const int lane_in_group = threadIdx.x & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    int64_t p0 = group * pts_per_group;
    if (p0 >= n_pts) return;
    int64_t p1 = p0 + pts_per_group;
    if (p1 > n_pts) p1 = n_pts;

    float4 acc[NV];
    int cur = -1;
    auto flush = [&]() {
        if (cur < 0) return;
        const float inv = 1.0f / (float)__ldg(counts + cur);
#pragma unroll
        for (int j = 0; j < NV; j++) {
            int v = lane_in_group + j * G;
            if (v < V) {
                float4 a = acc[j];
                a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
                red_add_v4(out + ((int64_t)cur * V + v) * 4, a);
            }
        }
    };
    for (int64_t p = p0; p < p1; p++) {
        int v_id = __ldg(idx + p);
        if (v_id < 0 || v_id >= n_vox) continue;
        if (v_id != cur) {
            flush();
            cur = v_id;
#pragma unroll
            for (int j = 0; j < NV; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < NV; j++) {
            int v = lane_in_group + j * G;
            if (v < V) {
                float4 f = __ldg(feats + p * V + v);
                acc[j].x += f.x; acc[j].y += f.y; acc[j].z += f.z; acc[j].w += f.w;
            }
        }
    }
    flush();
}

__global__ void __launch_bounds__(256) voxelize_fwd_scalar_kernel(const float *__restrict__ feats, int64_t n_pts, int C,
                                                                  const int *__restrict__ idx,
                                                                  const int *__restrict__ counts, float *__restrict__ out,
                                                                  int64_t n_vox) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pts * C) return;
    int64_t p = t / C;
    int c = (int)(t - p * C);
    int v = __ldg(idx + p);
    if (v < 0 || v >= n_vox) return;
    atomicAdd(out + (int64_t)v * C + c, feats[t] / (float)__ldg(counts + v));
}

extern "C" int u2_voxelize_fwd(const float *feats, int64_t n_pts, int32_t C, const int32_t *idx, const int32_t *counts,
                               float *out, int64_t n_vox, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(C > 0, "u2_voxelize_fwd: C=%d", C);
    if (n_vox > 0) U2_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)n_vox * C * sizeof(float), st));
    if (n_pts == 0 || n_vox == 0) return 0;
    const bool vec = (C % 4 == 0) && (((uintptr_t)feats | (uintptr_t)out) & 15) == 0 && C <= 1024;
    if (!vec) {
        voxelize_fwd_scalar_kernel<<<(unsigned)u2_ceil_div(n_pts * C, 256), 256, 0, st>>>(feats, n_pts, C, idx, counts,
                                                                                        out, n_vox);
        U2_LAUNCH_OK();
        return 0;
    }
    const int V = C / 4, G = group_lanes(V);
    const int NV = (V + G - 1) / G;
    // enough groups for >= ~8 waves of warps, but runs of >= 8 points so merging can happen
    int pts = 8;
    while (pts < 64 && u2_ceil_div(n_pts, pts) * G > (int64_t)U2_NUM_SMS * 2048 * 4) pts <<= 1;
    const int64_t groups = u2_ceil_div(n_pts, pts);
    const unsigned grid = (unsigned)u2_ceil_div(groups * G, 256);
    const float4 *f4 = (const float4 *)feats;
#define LAUNCH(NVT) voxelize_fwd_kernel<NVT><<<grid, 256, 0, st>>>(f4, n_pts, V, G, pts, idx, counts, out, n_vox)
    if (NV <= 1) LAUNCH(1);
    else if (NV <= 2) LAUNCH(2);
    else if (NV <= 4) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ voxelize bwd (pure gather)
// gfeats[p, :] = gout[idx[p], :] / counts[idx[p]].  A thread owns U 16-byte pieces of one point's row (strided by the group): the
// index / count chain (two dependent loads) is paid once per U pieces and the U row loads are in flight together — with one
// piece per thread the kernel sat at 0.41-0.48 of the HBM peak, bound by that three-deep dependent chain per 16 bytes.
template <int U>
__global__ void __launch_bounds__(256) voxelize_bwd_vec_kernel(const float4 *__restrict__ gout, int64_t n_vox, int per_row,
                                                               const int *__restrict__ idx, const int *__restrict__ counts,
                                                               float4 *__restrict__ gfeats, int64_t n_pts) {
    const int groups = per_row / U;  // thread groups per row
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pts * groups) return;
    const int64_t p = t / groups;
    const int v0 = (int)(t - p * groups);  // pieces v0, v0 + groups, ...: every load / store instruction of a warp stays contiguous
    const int vox = __ldg(idx + p);
    const bool ok = vox >= 0 && vox < n_vox;
    const int cnt = ok ? __ldg(counts + vox) : 0;
    float4 r[U];
#pragma unroll
    for (int u = 0; u < U; u++) r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok && cnt > 0) {
        const float4 *src = gout + (int64_t)vox * per_row + v0;
#pragma unroll
        for (int u = 0; u < U; u++) r[u] = __ldg(src + u * groups);
        const float c = (float)cnt;
#pragma unroll
        for (int u = 0; u < U; u++) { r[u].x /= c; r[u].y /= c; r[u].z /= c; r[u].w /= c; }
    }
    float4 *dst = gfeats + p * per_row + v0;
#pragma unroll
    for (int u = 0; u < U; u++) dst[u * groups] = r[u];
}

__global__ void __launch_bounds__(256) voxelize_bwd_kernel(const float *__restrict__ gout, int64_t n_vox, int C,
                                                           const int *__restrict__ idx, const int *__restrict__ counts,
                                                           float *__restrict__ gfeats, int64_t n_pts) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // scalar path: C not a multiple of 4 / unaligned
    if (t >= n_pts * C) return;
    int64_t p = t / C;
    int v = (int)(t - p * C);
    int vox = __ldg(idx + p);
    const bool ok = vox >= 0 && vox < n_vox;
    int cnt = ok ? __ldg(counts + vox) : 0;
    gfeats[t] = (ok && cnt > 0) ? __ldg(gout + (int64_t)vox * C + v) / (float)cnt : 0.f;
}

extern "C" int u2_voxelize_bwd(const float *gout, int64_t n_vox, int32_t C, const int32_t *idx, const int32_t *counts,
                               float *gfeats, int64_t n_pts, u2_stream_t stream) {
    if (n_pts == 0) return 0;
    U2_CHECK_ARG(C > 0, "u2_voxelize_bwd: C=%d", C);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C % 4 == 0) && (((uintptr_t)gout | (uintptr_t)gfeats) & 15) == 0;
    if (vec) {
        const int per_row = C / 4;
        const float4 *g4 = (const float4 *)gout;
        float4 *o4 = (float4 *)gfeats;
        if (per_row % 4 == 0)
            voxelize_bwd_vec_kernel<4><<<(unsigned)u2_ceil_div(n_pts * (per_row / 4), 256), 256, 0, st>>>(g4, n_vox, per_row, idx, counts, o4, n_pts);
        else if (per_row % 2 == 0)
            voxelize_bwd_vec_kernel<2><<<(unsigned)u2_ceil_div(n_pts * (per_row / 2), 256), 256, 0, st>>>(g4, n_vox, per_row, idx, counts, o4, n_pts);
        else
            voxelize_bwd_vec_kernel<1><<<(unsigned)u2_ceil_div(n_pts * per_row, 256), 256, 0, st>>>(g4, n_vox, per_row, idx, counts, o4, n_pts);
    } else {
        voxelize_bwd_kernel<<<(unsigned)u2_ceil_div(n_pts * C, 256), 256, 0, st>>>(gout, n_vox, C, idx, counts, gfeats, n_pts);
    }
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ trilinear weights
// Follows the reference's fp32 expression order exactly: products left to
// right, divide by scale^3, zero the missing corners, renormalise by (sum + 1e-8).
__global__ void __launch_bounds__(256) ti_weights_kernel(const float4 *__restrict__ coords,
                                                         const int64_t *__restrict__ idx_kn, int64_t n, float scale,
                                                         float *__restrict__ w_kn) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = __ldg(coords + i);
    float xf, yf, zf;
    if (scale != 1.0f) {
        xf = __fmul_rn(floorf(__fdiv_rn(c.x, scale)), scale);
        yf = __fmul_rn(floorf(__fdiv_rn(c.y, scale)), scale);
        zf = __fmul_rn(floorf(__fdiv_rn(c.z, scale)), scale);
    } else {
        xf = floorf(c.x); yf = floorf(c.y); zf = floorf(c.z);
    }
    const float xc = __fadd_rn(xf, scale), yc = __fadd_rn(yf, scale), zc = __fadd_rn(zf, scale);
    const float ax = __fsub_rn(xc, c.x), bx = __fsub_rn(c.x, xf);
    const float ay = __fsub_rn(yc, c.y), by = __fsub_rn(c.y, yf);
    const float az = __fsub_rn(zc, c.z), bz = __fsub_rn(c.z, zf);
    float w[8];
    w[0] = __fmul_rn(__fmul_rn(ax, ay), az);
    w[1] = __fmul_rn(__fmul_rn(ax, ay), bz);
    w[2] = __fmul_rn(__fmul_rn(ax, by), az);
    w[3] = __fmul_rn(__fmul_rn(ax, by), bz);
    w[4] = __fmul_rn(__fmul_rn(bx, ay), az);
    w[5] = __fmul_rn(__fmul_rn(bx, ay), bz);
    w[6] = __fmul_rn(__fmul_rn(bx, by), az);
    w[7] = __fmul_rn(__fmul_rn(bx, by), bz);
    const float s3 = scale * scale * scale;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (scale != 1.0f) w[k] = __fdiv_rn(w[k], s3);
        if (__ldg(idx_kn + (int64_t)k * n + i) == -1) w[k] = 0.f;
        sum = __fadd_rn(sum, w[k]);
    }
    const float den = __fadd_rn(sum, 1e-8f);
#pragma unroll
    for (int k = 0; k < 8; k++) w_kn[(int64_t)k * n + i] = __fdiv_rn(w[k], den);
}

extern "C" int u2_ti_weights(const float *coords, const int64_t *idx_kn, int64_t n_pts, float scale, float *weights_kn,
                             u2_stream_t stream) {
    if (n_pts == 0) return 0;
    U2_CHECK_ARG(((uintptr_t)coords & 15) == 0, "u2_ti_weights: coords must be 16-byte aligned [N,4] fp32");
    ti_weights_kernel<<<(unsigned)u2_ceil_div(n_pts, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)coords, idx_kn, n_pts, scale, weights_kn);
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ devoxelize fwd
// A group of G lanes per point; each lane gathers the <=8 corner rows' float4 columns it
// owns (8 independent 16-byte loads in flight) and blends them. Zero-weight corners are
// not fetched (stride-1 points sit on voxel corners: 7 of 8 weights vanish).
template <int NV>
__global__ void __launch_bounds__(256) devoxelize_fwd_kernel(const float4 *__restrict__ feats, int V, int G,
                                                             const int4 *__restrict__ idx, const float4 *__restrict__ w,
                                                             int64_t n_pts, float4 *__restrict__ out) {
    const int lane_in_group = threadIdx.x & (G - 1);
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (p >= n_pts) return;
    const int4 i0 = __ldg(idx + 2 * p), i1 = __ldg(idx + 2 * p + 1);
    const float4 w0 = __ldg(w + 2 * p), w1 = __ldg(w + 2 * p + 1);
    const int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    const float wt[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int v = lane_in_group + j * G;
        if (v >= V) break;
        float4 f[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
            f[k] = (id[k] >= 0 && wt[k] != 0.f) ? __ldg(feats + (int64_t)id[k] * V + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a.x += wt[k] * f[k].x; a.y += wt[k] * f[k].y; a.z += wt[k] * f[k].z; a.w += wt[k] * f[k].w;
        }
        out[p * V + v] = a;
    }
}

__global__ void __launch_bounds__(256) devoxelize_fwd_scalar_kernel(const float *__restrict__ feats, int C,
                                                                    const int *__restrict__ idx, const float *__restrict__ w,
                                                                    int64_t n_pts, float *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pts * C) return;
    int64_t p = t / C;
    int c = (int)(t - p * C);
    float a = 0.f;
    for (int k = 0; k < 8; k++) {
        int v = __ldg(idx + p * 8 + k);
        if (v >= 0) a += __ldg(w + p * 8 + k) * __ldg(feats + (int64_t)v * C + c);
    }
    out[t] = a;
}

extern "C" int u2_devoxelize_fwd(const float *feats, int64_t n_vox, int32_t C, const int32_t *idx, const float *w,
                                 int64_t n_pts, float *out, u2_stream_t stream) {
    (void)n_vox;
    if (n_pts == 0) return 0;
    U2_CHECK_ARG(C > 0, "u2_devoxelize_fwd: C=%d", C);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C % 4 == 0) && C <= 1024 &&
                     (((uintptr_t)feats | (uintptr_t)out | (uintptr_t)idx | (uintptr_t)w) & 15) == 0;
    if (!vec) {
        devoxelize_fwd_scalar_kernel<<<(unsigned)u2_ceil_div(n_pts * C, 256), 256, 0, st>>>(feats, C, idx, w, n_pts, out);
        U2_LAUNCH_OK();
        return 0;
    }
    const int V = C / 4, G = group_lanes(V), NV = (V + G - 1) / G;
    const unsigned grid = (unsigned)u2_ceil_div(n_pts * G, 256);
#define LAUNCH(NVT) \
    devoxelize_fwd_kernel<NVT><<<grid, 256, 0, st>>>((const float4 *)feats, V, G, (const int4 *)idx, (const float4 *)w, n_pts, (float4 *)out)
    if (NV <= 1) LAUNCH(1);
    else if (NV <= 2) LAUNCH(2);
    else if (NV <= 4) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ devoxelize bwd
// gfeats[idx[i,k]] += w[i,k] * gout[i]. A group of G lanes walks PTS consecutive points and,
// per corner slot k, keeps a running total while the target voxel stays the same
// (neighbouring points share corners at coarse strides); totals leave as 16-byte reductions.
template <int NV>
__global__ void __launch_bounds__(256) devoxelize_bwd_kernel(const float4 *__restrict__ gout, int V, int G,
                                                             int pts_per_group, const int4 *__restrict__ idx,
                                                             const float4 *__restrict__ w, int64_t n_pts,
                                                             float *__restrict__ gfeats) {
    const int lane_in_group = threadIdx.x & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    int64_t p0 = group * pts_per_group;
    if (p0 >= n_pts) return;
    int64_t p1 = p0 + pts_per_group;
    if (p1 > n_pts) p1 = n_pts;
#pragma unroll 1
    for (int j = 0; j < NV; j++) {
        const int v = lane_in_group + j * G;
        if (v >= V) break;
        float4 acc[8];
        int cur[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { cur[k] = -1; acc[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
        // the loads of point p + 1 are issued before point p is processed (the walk is serial: without the prefetch every
        // point exposed a full global-load latency in front of its reductions)
        int4 ni0 = __ldg(idx + 2 * p0), ni1 = __ldg(idx + 2 * p0 + 1);
        float4 nw0 = __ldg(w + 2 * p0), nw1 = __ldg(w + 2 * p0 + 1);
        float4 ng = __ldg(gout + p0 * V + v);
        for (int64_t p = p0; p < p1; p++) {
            const int4 i0 = ni0, i1 = ni1;
            const float4 w0 = nw0, w1 = nw1;
            const float4 g = ng;
            if (p + 1 < p1) {
                ni0 = __ldg(idx + 2 * (p + 1)); ni1 = __ldg(idx + 2 * (p + 1) + 1);
                nw0 = __ldg(w + 2 * (p + 1)); nw1 = __ldg(w + 2 * (p + 1) + 1);
                ng = __ldg(gout + (p + 1) * V + v);
            }
            const int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
            const float wt[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (id[k] < 0 || wt[k] == 0.f) continue;
                if (id[k] != cur[k]) {
                    if (cur[k] >= 0) red_add_v4(gfeats + ((int64_t)cur[k] * V + v) * 4, acc[k]);
                    cur[k] = id[k];
                    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                acc[k].x += wt[k] * g.x; acc[k].y += wt[k] * g.y; acc[k].z += wt[k] * g.z; acc[k].w += wt[k] * g.w;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (cur[k] >= 0) red_add_v4(gfeats + ((int64_t)cur[k] * V + v) * 4, acc[k]);
    }
}

__global__ void __launch_bounds__(256) devoxelize_bwd_scalar_kernel(const float *__restrict__ gout, int C,
                                                                    const int *__restrict__ idx, const float *__restrict__ w,
                                                                    int64_t n_pts, float *__restrict__ gfeats) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pts * C) return;
    int64_t p = t / C;
    int c = (int)(t - p * C);
    const float g = gout[t];
    for (int k = 0; k < 8; k++) {
        int v = __ldg(idx + p * 8 + k);
        if (v >= 0) atomicAdd(gfeats + (int64_t)v * C + c, __ldg(w + p * 8 + k) * g);
    }
}

extern "C" int u2_devoxelize_bwd(const float *gout, int64_t n_pts, int32_t C, const int32_t *idx, const float *w,
                                 float *gfeats, int64_t n_vox, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(C > 0, "u2_devoxelize_bwd: C=%d", C);
    if (n_vox > 0) U2_CUDA_OK(cudaMemsetAsync(gfeats, 0, (size_t)n_vox * C * sizeof(float), st));
    if (n_pts == 0 || n_vox == 0) return 0;
    const bool vec = (C % 4 == 0) && C <= 1024 &&
                     (((uintptr_t)gfeats | (uintptr_t)gout | (uintptr_t)idx | (uintptr_t)w) & 15) == 0;
    if (!vec) {
        devoxelize_bwd_scalar_kernel<<<(unsigned)u2_ceil_div(n_pts * C, 256), 256, 0, st>>>(gout, C, idx, w, n_pts, gfeats);
        U2_LAUNCH_OK();
        return 0;
    }
    const int V = C / 4, G = group_lanes(V), NV = (V + G - 1) / G;
    const int pts = 8;
    const int64_t groups = u2_ceil_div(n_pts, pts);
    const unsigned grid = (unsigned)u2_ceil_div(groups * G, 256);
#define LAUNCH(NVT) \
    devoxelize_bwd_kernel<NVT><<<grid, 256, 0, st>>>((const float4 *)gout, V, G, pts, (const int4 *)idx, (const float4 *)w, n_pts, gfeats)
    if (NV <= 1) LAUNCH(1);
    else if (NV <= 2) LAUNCH(2);
    else if (NV <= 4) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    U2_LAUNCH_OK();
    return 0;
}

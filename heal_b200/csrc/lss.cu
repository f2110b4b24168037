// Lift-Splat-Shoot: frustum geometry -> BEV cell index, and the fused depth-softmax (x) feature outer
// product + BEV pooling.
// Reference: opencood/models/heter_encoders.py:125-147 (get_geometry), :161-217 (voxel_pooling, incl. the
// truncating `.long()` index :173 and the (B,C,Y,X) output layout :206-212),
// opencood/models/sub_modules/lss_submodule.py:132-134 / :227-229 (softmax over D, depth[:,None]*feat[:,:,None]),
// opencood/utils/camera_utils.py:220-246 (QuickCumsum = per-cell sum).
// The (n*cams, C, D, fH, fW) outer-product tensor of the reference (277 MB per agent at 704x256) is never
// materialised: each block stages one image row of depth logits and features in shared memory, computes the
// softmax in place and accumulates prob*feat straight into the channels-last BEV map.
#include "common.cuh"
#include "../../include/heal_b200.h"

namespace {

struct GeoP {
    const float* frustum;      // (D, fH, fW, 3): (u, v, depth)
    const float* post_inv;     // (BN, 3, 3) inverse(post_rots)
    const float* post_trans;   // (BN, 3)
    const float* combine;      // (BN, 3, 3) rots @ inverse(intrins)
    const float* trans;        // (BN, 3)
    float lower[3], dx[3];     // bx - dx/2 and dx, as the reference computes them in fp32
    int nx[3];
    int BN, DHW;
    int* cell;                 // (BN, D, fH, fW): y*nx + x, or -1 when the point falls outside the grid
};

__global__ void k_lss_cell_index(GeoP g) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)g.BN * g.DHW) return;
    int bn = (int)(i / g.DHW), f = (int)(i % g.DHW);
    const float* fr = g.frustum + 3 * (size_t)f;
    const float* pt = g.post_trans + 3 * bn;
    float p0 = fr[0] - pt[0], p1 = fr[1] - pt[1], p2 = fr[2] - pt[2];
    const float* A = g.post_inv + 9 * bn;
    float q0 = A[0] * p0 + A[1] * p1 + A[2] * p2;
    float q1 = A[3] * p0 + A[4] * p1 + A[5] * p2;
    float q2 = A[6] * p0 + A[7] * p1 + A[8] * p2;
    float r0 = q0 * q2, r1 = q1 * q2, r2 = q2;
    const float* C = g.combine + 9 * bn;
    const float* t = g.trans + 3 * bn;
    float e[3];
    e[0] = (C[0] * r0 + C[1] * r1 + C[2] * r2) + t[0];
    e[1] = (C[3] * r0 + C[4] * r1 + C[5] * r2) + t[1];
    e[2] = (C[6] * r0 + C[7] * r1 + C[8] * r2) + t[2];
    int idx[3];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = __fdiv_rn(e[j] - g.lower[j], g.dx[j]);
        // .long() truncates toward zero: values in (-1, 0) land in cell 0 (heter_encoders.py:173,181)
        float tr = truncf(v);
        ok = ok && (tr >= 0.f) && (tr < (float)g.nx[j]);
        idx[j] = ok ? (int)tr : 0;
    }
    g.cell[i] = ok ? ((idx[2] * g.nx[1] + idx[1]) * g.nx[0] + idx[0]) : -1;
}

struct PoolP {
    const float* logits;   // (BN, D, fH, fW)   depth_head output (NCHW as the torch trunk produces it)
    const float* feat;     // (BN, C, fH, fW)   image_head output
    const int* cell;       // (BN, D, fH, fW)
    float* out;            // (B, nz*ny*nx, C) channels-last, pre-zeroed
    int BN, cams, D, C, fH, fW, cells_per_agent;
};

// one block per (bn, h) image row
__global__ void __launch_bounds__(256)
k_lss_pool(PoolP p) {
    extern __shared__ float sm[];
    float* sProb = sm;                          // [fW][D+1]
    float* sFeat = sm + (size_t)p.fW * (p.D + 1);   // [fW][C+1]
    const int bn = blockIdx.x / p.fH, h = blockIdx.x % p.fH;
    const int HW = p.fH * p.fW;
    for (int i = threadIdx.x; i < p.D * p.fW; i += blockDim.x) {
        int d = i / p.fW, w = i % p.fW;
        sProb[w * (p.D + 1) + d] = __ldg(p.logits + ((size_t)bn * p.D + d) * HW + (size_t)h * p.fW + w);
    }
    for (int i = threadIdx.x; i < p.C * p.fW; i += blockDim.x) {
        int c = i / p.fW, w = i % p.fW;
        sFeat[w * (p.C + 1) + c] = __ldg(p.feat + ((size_t)bn * p.C + c) * HW + (size_t)h * p.fW + w);
    }
    __syncthreads();
    // softmax over D per pixel (F.softmax(x, dim=1))
    for (int w = threadIdx.x; w < p.fW; w += blockDim.x) {
        float* pr = sProb + w * (p.D + 1);
        float mx = -INFINITY;
        for (int d = 0; d < p.D; ++d) mx = fmaxf(mx, pr[d]);
        float s = 0.f;
        for (int d = 0; d < p.D; ++d) { float e = expf(pr[d] - mx); pr[d] = e; s += e; }
        for (int d = 0; d < p.D; ++d) pr[d] = pr[d] / s;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int agent = bn / p.cams;
    float* outa = p.out + (size_t)agent * p.cells_per_agent * p.C;
    for (int pt = warp; pt < p.D * p.fW; pt += nwarps) {
        int d = pt / p.fW, w = pt % p.fW;
        int cell = __ldg(p.cell + ((size_t)bn * p.D + d) * HW + (size_t)h * p.fW + w);
        if (cell < 0) continue;
        float pr = sProb[w * (p.D + 1) + d];
        const float* fw = sFeat + w * (p.C + 1);
        float* o = outa + (size_t)cell * p.C;
        for (int c = lane; c < p.C; c += 32) atomicAdd(o + c, pr * fw[c]);
    }
}

}  // namespace

extern "C" int heal_lss_cell_index(const float* frustum, int D, int fH, int fW,
                                   const float* post_rots_inv, const float* post_trans, const float* combine, const float* trans,
                                   int num_images, const float* lower3_host, const float* dx3_host, const int* nx3_host,
                                   int* cell_out, void* stream_) {
    if (!frustum || !post_rots_inv || !post_trans || !combine || !trans || !cell_out) return HEAL_ERR_ARG;
    if (num_images <= 0) return HEAL_OK;
    GeoP g;
    g.frustum = frustum; g.post_inv = post_rots_inv; g.post_trans = post_trans; g.combine = combine; g.trans = trans;
    for (int j = 0; j < 3; ++j) { g.lower[j] = lower3_host[j]; g.dx[j] = dx3_host[j]; g.nx[j] = nx3_host[j]; }
    g.BN = num_images; g.DHW = D * fH * fW; g.cell = cell_out;
    long long total = (long long)num_images * g.DHW;
    k_lss_cell_index<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(g);
    return heal_check_launch();
}

extern "C" int heal_lss_pool(const float* depth_logits, const float* feat, const int* cell, int num_images, int cams_per_agent,
                             int D, int C, int fH, int fW, int cells_per_agent, float* bev_out, void* stream_) {
    if (!depth_logits || !feat || !cell || !bev_out) return HEAL_ERR_ARG;
    if (num_images <= 0) return HEAL_OK;
    if (cams_per_agent < 1 || (num_images % cams_per_agent)) return HEAL_ERR_ARG;
    PoolP p;
    p.logits = depth_logits; p.feat = feat; p.cell = cell; p.out = bev_out;
    p.BN = num_images; p.cams = cams_per_agent; p.D = D; p.C = C; p.fH = fH; p.fW = fW; p.cells_per_agent = cells_per_agent;
    size_t smem = ((size_t)fW * (D + 1) + (size_t)fW * (C + 1)) * sizeof(float);
    if (smem > 200 * 1024) return HEAL_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(k_lss_pool, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return HEAL_ERR_LAUNCH;
    }
    k_lss_pool<<<num_images * fH, 256, smem, (cudaStream_t)stream_>>>(p);
    return heal_check_launch();
}

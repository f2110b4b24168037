// Sparse stem: the first residual block's two stride-2 convolutions evaluated directly from the pillar list, so the dense
// (B, ny, nx, 64) BEV canvas of PointPillarScatter is never materialised (335 MB zero-fill + two full reads at 5 x 512^2).
// Reference ops fused: PointPillarScatter.forward (point_pillar_scatter.py:19-77) -> BasicBlock.conv1 (3x3, stride 2, pad 1) +
// bn1 + ReLU and BasicBlock.downsample (1x1, stride 2) + bn (resblock.py:48-64, 178-187).  Only ~4.6 % of the pillars' cells
// are occupied; an output pixel gathers the <= 9 occupied input cells under its window through an id map (pillar row or -1),
// every other pixel is just ReLU(folded bias).  fp32 FMA math, any output storage format.
#include "common.cuh"
#include "../../include/heal_b200.h"

namespace {

struct StemP {
    const float* feats;     // (M, 64) pillar features
    const int* idmap;       // (B, ny, nx) pillar row or -1
    const float* w1;        // [9][64][64] (tap, cin, cout), BN folded
    const float* b1;        // [64]
    const float* w2;        // [64][64] (cin, cout) 1x1 stride-2 downsample, BN folded
    const float* b2;        // [64]
    ActV out1, out2;        // (B, Ho, Wo, 64)
    int B, ny, nx, Ho, Wo;
};

__device__ __forceinline__ void store2(const ActV& a, size_t pix, int c, float x, float y) {
    size_t idx = pix * (size_t)a.cs + (size_t)(a.co + c);
    if (a.fmt == 0) { *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.p) + idx) = make_float2(x, y); return; }
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(a.p) + idx;
    __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    *reinterpret_cast<__nv_bfloat162*>(b) = h;
    if (a.fmt == 2) {
        float2 hf = __bfloat1622float2(h);
        *reinterpret_cast<__nv_bfloat162*>(b + a.plane) = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    }
}

__global__ void __launch_bounds__(256)
k_sparse_stem(StemP p) {
    const int lane = threadIdx.x & 31;
    const long long npix = (long long)p.B * p.Ho * p.Wo;
    const float2 bias1 = *reinterpret_cast<const float2*>(p.b1 + 2 * lane);
    const float2 bias2 = *reinterpret_cast<const float2*>(p.b2 + 2 * lane);
    for (long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix;
         pix += (long long)gridDim.x * (blockDim.x >> 5)) {
        const int ox = (int)(pix % p.Wo); long long t = pix / p.Wo;
        const int oy = (int)(t % p.Ho); const int b = (int)(t / p.Ho);
        int id = -1;
        if (lane < 9) {
            int iy = 2 * oy - 1 + lane / 3, ix = 2 * ox - 1 + lane % 3;
            if (iy >= 0 && iy < p.ny && ix >= 0 && ix < p.nx) id = __ldg(p.idmap + ((size_t)b * p.ny + iy) * p.nx + ix);
        }
        unsigned hits = __ballot_sync(0xffffffffu, id >= 0);
        float2 a1 = bias1, a2 = bias2;
        while (hits) {
            const int tap = __ffs(hits) - 1;
            hits &= hits - 1;
            const int pid = __shfl_sync(0xffffffffu, id, tap);
            const float2 f = *reinterpret_cast<const float2*>(p.feats + (size_t)pid * 64 + 2 * lane);
            const float* w = p.w1 + (size_t)tap * 64 * 64 + 2 * lane;
#pragma unroll 8
            for (int k = 0; k < 64; ++k) {
                const float fk = __shfl_sync(0xffffffffu, (k & 1) ? f.y : f.x, k >> 1);
                const float2 wv = __ldg(reinterpret_cast<const float2*>(w + (size_t)k * 64));
                a1.x = fmaf(fk, wv.x, a1.x); a1.y = fmaf(fk, wv.y, a1.y);
            }
            if (tap == 4) {                         // centre cell (2oy, 2ox) is the 1x1 stride-2 downsample's input
                const float* w2 = p.w2 + 2 * lane;
#pragma unroll 8
                for (int k = 0; k < 64; ++k) {
                    const float fk = __shfl_sync(0xffffffffu, (k & 1) ? f.y : f.x, k >> 1);
                    const float2 wv = __ldg(reinterpret_cast<const float2*>(w2 + (size_t)k * 64));
                    a2.x = fmaf(fk, wv.x, a2.x); a2.y = fmaf(fk, wv.y, a2.y);
                }
            }
        }
        store2(p.out1, (size_t)pix, 2 * lane, fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f));
        store2(p.out2, (size_t)pix, 2 * lane, a2.x, a2.y);
    }
}

__global__ void k_fill_idmap(const int4* __restrict__ coords, const int* __restrict__ m_dev, int M, int ny, int nx, int* __restrict__ idmap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int Md = m_dev ? min(M, m_dev[0]) : M;
    if (i >= Md) return;
    int4 c = coords[i];   // [b, z, y, x]
    idmap[((size_t)c.x * ny + c.z) * nx + (c.y + c.w)] = i;
}

}  // namespace

extern "C" int heal_pillar_idmap(const int* voxel_coords, const int* num_voxels_dev, int num_voxels, int batch, int ny, int nx,
                                 int* idmap_out, void* stream_) {
    if (!voxel_coords || !idmap_out || batch < 1) return HEAL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    cudaMemsetAsync(idmap_out, 0xFF, (size_t)batch * ny * nx * sizeof(int), st);
    if (num_voxels > 0)
        k_fill_idmap<<<(num_voxels + 255) / 256, 256, 0, st>>>((const int4*)voxel_coords, num_voxels_dev, num_voxels, ny, nx, idmap_out);
    return heal_check_launch();
}

extern "C" int heal_sparse_stem(const float* pillar_features, const int* idmap, int batch, int ny, int nx,
                                const float* w_conv3x3, const float* b_conv3x3, const float* w_down1x1, const float* b_down1x1,
                                int channels, const heal_act_t* out_conv, const heal_act_t* out_down, void* stream_) {
    if (!pillar_features || !idmap || !w_conv3x3 || !b_conv3x3 || !w_down1x1 || !b_down1x1 || !out_conv || !out_down) return HEAL_ERR_ARG;
    if (channels != 64 || (ny & 1) || (nx & 1)) return HEAL_ERR_UNSUPPORTED;
    if ((out_conv->cstride & 1) || (out_conv->coffset & 1) || (out_down->cstride & 1) || (out_down->coffset & 1)) return HEAL_ERR_UNSUPPORTED;
    StemP p;
    p.feats = pillar_features; p.idmap = idmap; p.w1 = w_conv3x3; p.b1 = b_conv3x3; p.w2 = w_down1x1; p.b2 = b_down1x1;
    p.out1.p = out_conv->data; p.out1.fmt = out_conv->fmt; p.out1.cs = out_conv->cstride; p.out1.co = out_conv->coffset; p.out1.plane = out_conv->plane_stride;
    p.out2.p = out_down->data; p.out2.fmt = out_down->fmt; p.out2.cs = out_down->cstride; p.out2.co = out_down->coffset; p.out2.plane = out_down->plane_stride;
    p.B = batch; p.ny = ny; p.nx = nx; p.Ho = ny / 2; p.Wo = nx / 2;
    long long npix = (long long)batch * p.Ho * p.Wo;
    long long blocks = (npix + 7) / 8;
    long long cap = (long long)HEAL_NUM_SMS * 32;
    k_sparse_stem<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream_>>>(p);
    return heal_check_launch();
}

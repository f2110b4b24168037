// Fused PillarVFE (10-d point augmentation -> Linear(10,64) -> BatchNorm1d(eval) -> ReLU -> max over the
// T slots) + PointPillarScatter into a channels-last BEV canvas.
// Reference: opencood/models/sub_modules/pillar_vfe.py:105-155 (PillarVFE.forward), :31-53 (PFNLayer),
//            opencood/models/sub_modules/point_pillar_scatter.py:19-77.
// One warp per pillar. Stage A: lane = point slot (coalesced 16 B loads, warp-reduced mean).
// Stage B: lane = channel pair, point features broadcast from shared memory, running max in
// registers, one coalesced 256 B store of the pillar's 64 channels into canvas[b][y][x][:].
// Parity traps kept: padded slots are zeroed BEFORE the linear layer and therefore contribute
// ReLU(BN(0)) to the max (pillar_vfe.py:149,46); scatter index = z + y*nx + x (scatter.py:58).
#include "common.cuh"
#include "../../include/heal_b200.h"

namespace {

constexpr int PV_WARPS = 8;
constexpr int PV_COUT = 64;
constexpr int PV_CIN = 10;

struct PvCfg {
    float vx, vy, vz;
    float xoff, yoff, zoff;
    int T, nx, ny;
};

__global__ void __launch_bounds__(PV_WARPS * 32)
k_pillar_vfe_scatter(const float4* __restrict__ voxels, const int* __restrict__ num_points,
                     const int4* __restrict__ coords, const int* __restrict__ num_voxels_dev, int M,
                     const float* __restrict__ Wf, const float* __restrict__ bf, PvCfg c,
                     float* __restrict__ pillar_out, __nv_bfloat16* __restrict__ pillar_split_out, ActV canvas) {
    __shared__ float sW[PV_CIN * PV_COUT];
    __shared__ float sB[PV_COUT];
    __shared__ __align__(16) float sF[PV_WARPS][32][12];
    for (int i = threadIdx.x; i < PV_CIN * PV_COUT; i += blockDim.x) sW[i] = Wf[i];
    if (threadIdx.x < PV_COUT) sB[threadIdx.x] = bf[threadIdx.x];
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int Mdev = num_voxels_dev ? min(M, num_voxels_dev[0]) : M;
    // ---- lane = channel pair: this lane's two weight columns stay in registers for every pillar the warp processes ----
    float w0[PV_CIN], w1[PV_CIN];
#pragma unroll
    for (int k = 0; k < PV_CIN; ++k) { w0[k] = sW[k * PV_COUT + 2 * lane]; w1[k] = sW[k * PV_COUT + 2 * lane + 1]; }
    const float b0 = sB[2 * lane], b1 = sB[2 * lane + 1];
    // persistent warps: grid sized for the machine (the live pillar count is on the device, ~5x below the row capacity)
    for (int v = blockIdx.x * PV_WARPS + warp; v < Mdev; v += gridDim.x * PV_WARPS) {

    // ---- stage A: lane = point slot --------------------------------------------------------
    int n = num_points[v];
    int4 cd = coords[v];  // [b, z, y, x]
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    // only the n real slots are read: the padded ones are zero by contract (sp_voxel_preprocessor zero-fills) and would add +0
    // to the three sums, so skipping them is bit-identical and saves (T - n) * 16 B of DRAM reads per pillar (~10x on LiDAR pillars)
    if (lane < min(n, c.T)) p = __ldg(voxels + (size_t)v * c.T + lane);
    float fn = (float)n;
    float mx = __fdiv_rn(warp_sum(p.x), fn), my = __fdiv_rn(warp_sum(p.y), fn), mz = __fdiv_rn(warp_sum(p.z), fn);
    float cx = __fadd_rn(__fmul_rn((float)cd.w, c.vx), c.xoff);
    float cy = __fadd_rn(__fmul_rn((float)cd.z, c.vy), c.yoff);
    float cz = __fadd_rn(__fmul_rn((float)cd.y, c.vz), c.zoff);
    float msk = (lane < n) ? 1.f : 0.f;
    float* f = sF[warp][lane];
    reinterpret_cast<float4*>(f)[0] = make_float4(p.x * msk, p.y * msk, p.z * msk, p.w * msk);
    reinterpret_cast<float4*>(f)[1] = make_float4((p.x - mx) * msk, (p.y - my) * msk, (p.z - mz) * msk, (p.x - cx) * msk);
    reinterpret_cast<float4*>(f)[2] = make_float4((p.y - cy) * msk, (p.z - cz) * msk, 0.f, 0.f);
    __syncwarp();

    // ---- stage B: lane = channel pair ------------------------------------------------------
    float m0 = 0.f, m1 = 0.f;  // ReLU output >= 0, so 0 is the identity of the running max
    // padded slots are all-zero rows: Linear gives +0, so each contributes exactly ReLU(b) to the max -- fold them into one
    // fmaxf instead of T - n dead dot products (LiDAR pillars hold ~3 points of T = 32)
    const int nt = min(n, c.T);
    if (nt < c.T) { m0 = fmaxf(m0, b0); m1 = fmaxf(m1, b1); }
    for (int t = 0; t < nt; ++t) {
        const float4* ft = reinterpret_cast<const float4*>(sF[warp][t]);
        float4 a = ft[0], b = ft[1], d = ft[2];
        float x[PV_CIN] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y};
        float y0 = 0.f, y1 = 0.f;
#pragma unroll
        for (int k = 0; k < PV_CIN; ++k) { y0 = fmaf(x[k], w0[k], y0); y1 = fmaf(x[k], w1[k], y1); }
        m0 = fmaxf(m0, y0 + b0);
        m1 = fmaxf(m1, y1 + b1);
    }
    float2 o = make_float2(m0, m1);
    if (pillar_out) reinterpret_cast<float2*>(pillar_out + (size_t)v * PV_COUT)[lane] = o;
    if (pillar_split_out) {      // split rows [hi 64 | lo 64]: the gather source of the tensor-core sparse stem (heal_spconv_gather_gemm_tc)
        __nv_bfloat16* r = pillar_split_out + (size_t)v * (2 * PV_COUT);
        const __nv_bfloat162 h = __floats2bfloat162_rn(o.x, o.y);
        const float2 hf = __bfloat1622float2(h);
        reinterpret_cast<__nv_bfloat162*>(r)[lane] = h;
        reinterpret_cast<__nv_bfloat162*>(r + PV_COUT)[lane] = __floats2bfloat162_rn(o.x - hf.x, o.y - hf.y);
    }
    if (canvas.p) {
        size_t cell = ((size_t)cd.x * c.ny + (size_t)cd.z) * c.nx + (size_t)(cd.y + cd.w);  // z + y*nx + x, z == 0
        if (canvas.fmt == 0) {
            reinterpret_cast<float2*>(reinterpret_cast<float*>(canvas.p) + cell * canvas.cs + canvas.co)[lane] = o;
        } else {
            __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(canvas.p) + cell * canvas.cs + canvas.co;
            __nv_bfloat162 h = __floats2bfloat162_rn(o.x, o.y);
            reinterpret_cast<__nv_bfloat162*>(b)[lane] = h;
            if (canvas.fmt == 2) {
                float2 hf = __bfloat1622float2(h);
                reinterpret_cast<__nv_bfloat162*>(b + canvas.plane)[lane] = __floats2bfloat162_rn(o.x - hf.x, o.y - hf.y);
            }
        }
    }
    __syncwarp();          // sF[warp] is rewritten by the next pillar
    }
}

// Stand-alone PointPillarScatter (point_pillar_scatter.py:19-77): rows of already computed pillar features -> canvas cells.
// One warp per pillar, lanes over channel pairs; the caller pre-zeroes the canvas.
__global__ void __launch_bounds__(256)
k_pillar_scatter(const float* __restrict__ feats, const int4* __restrict__ coords, const int* __restrict__ num_voxels_dev, int M,
                 int C, int nx, int ny, ActV canvas) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + warp;
    const int Mdev = num_voxels_dev ? min(M, num_voxels_dev[0]) : M;
    if (v >= Mdev) return;
    const int4 cd = coords[v];
    const size_t cell = ((size_t)cd.x * ny + (size_t)cd.z) * nx + (size_t)(cd.y + cd.w);
    for (int c = lane; c < C; c += 32) act_store1(canvas, cell, c, __ldg(feats + (size_t)v * C + c));
}

}  // namespace

extern "C" int heal_pillar_scatter(const float* pillar_features, const int* voxel_coords, const int* num_voxels_dev, int num_voxels,
                                   int channels, int nx, int ny, const heal_act_t* canvas_out, void* stream_) {
    if (!pillar_features || !voxel_coords || !canvas_out || !canvas_out->data || channels < 1) return HEAL_ERR_ARG;
    if (num_voxels <= 0) return HEAL_OK;
    ActV cv;
    cv.p = canvas_out->data; cv.fmt = canvas_out->fmt; cv.cs = canvas_out->cstride; cv.co = canvas_out->coffset; cv.plane = canvas_out->plane_stride;
    k_pillar_scatter<<<(num_voxels + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(pillar_features, (const int4*)voxel_coords, num_voxels_dev,
                                                                            num_voxels, channels, nx, ny, cv);
    return heal_check_launch();
}

extern "C" int heal_pillar_vfe_scatter(const float* voxel_features, const int* voxel_num_points, const int* voxel_coords,
                                       const int* num_voxels_dev, int num_voxels, int max_points_per_voxel,
                                       const float* w_folded, const float* b_folded, int c_in, int c_out,
                                       const float* voxel_size3, const float* offset3, int nx, int ny,
                                       float* pillar_features_out, void* pillar_split_rows_out, const heal_act_t* canvas_out, void* stream_) {
    if (!voxel_features || !voxel_num_points || !voxel_coords || !w_folded || !b_folded) return HEAL_ERR_ARG;
    if (c_in != PV_CIN || c_out != PV_COUT || max_points_per_voxel < 1 || max_points_per_voxel > 32) return HEAL_ERR_UNSUPPORTED;
    if (num_voxels <= 0) return HEAL_OK;
    PvCfg c;
    c.vx = voxel_size3[0]; c.vy = voxel_size3[1]; c.vz = voxel_size3[2];
    c.xoff = offset3[0]; c.yoff = offset3[1]; c.zoff = offset3[2];
    c.T = max_points_per_voxel; c.nx = nx; c.ny = ny;
    ActV cv;
    cv.p = nullptr; cv.fmt = 0; cv.cs = PV_COUT; cv.co = 0; cv.plane = 0;
    if (canvas_out && canvas_out->data) {
        if ((canvas_out->cstride & 1) || (canvas_out->coffset & 1)) return HEAL_ERR_UNSUPPORTED;
        cv.p = canvas_out->data; cv.fmt = canvas_out->fmt; cv.cs = canvas_out->cstride; cv.co = canvas_out->coffset;
        cv.plane = canvas_out->plane_stride;
    }
    int grid = (num_voxels + PV_WARPS - 1) / PV_WARPS;
    if (grid > HEAL_NUM_SMS * 8) grid = HEAL_NUM_SMS * 8;
    k_pillar_vfe_scatter<<<grid, PV_WARPS * 32, 0, (cudaStream_t)stream_>>>(
        (const float4*)voxel_features, voxel_num_points, (const int4*)voxel_coords, num_voxels_dev, num_voxels,
        w_folded, b_folded, c, pillar_features_out, (__nv_bfloat16*)pillar_split_rows_out, cv);
    return heal_check_launch();
}

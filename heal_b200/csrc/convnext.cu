// ConvNeXt aligner front half: depthwise k x k convolution + bias + LayerNorm over the channels of every pixel, one kernel.
// Reference: opencood/models/sub_modules/feature_alignnet_modules.py:299-345 (ConvNeXtBlock: dwconv -> permute -> LayerNorm(eps 1e-6,
// channels_last, :12-25) -> Linear -> GELU -> Linear -> gamma -> + input).  The two Linear layers are 1x1 convolutions and run on
// the tcgen05 conv engine (csrc/conv2d_tc.cu, GELU / residual in its epilogue); this kernel produces their input.
// HBM-bound: reads the map once (the 7x7 neighbourhood of a pixel is shared through L1/L2 by the warps of a block, which cover
// 8 x 4 neighbouring pixels), writes it once.  One warp per pixel, lane = channel pair (C = 64) -> every tap is one coalesced
// 256 B (fp32) row read; mean / variance by warp shuffles, exactly LayerNorm's biased variance.
#include "common.cuh"
#include "../../include/heal_b200.h"

namespace {

struct DwLnP {
    ActV in, out;
    const float* w;      // [k*k][C] depthwise weights (tap-major, channel contiguous)
    const float* b;      // [C] conv bias
    const float* lnw;    // [C] LayerNorm weight
    const float* lnb;    // [C] LayerNorm bias
    int N, H, W, C, k;
    float eps;
};

// block = 32 warps = 8 (x) x 4 (y) pixels; lanes cover C in chunks of 64 (2 channels per lane per chunk)
__global__ void __launch_bounds__(1024)
k_dwconv_ln(DwLnP p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = blockIdx.x * 8 + (warp & 7), y = blockIdx.y * 4 + (warp >> 3), n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int r = p.k / 2;
    const size_t img = (size_t)n * p.H * p.W;
    float acc[8];                                  // up to C = 256: 4 chunks x 2 channels
    const int chunks = p.C / 64;
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[2 * q] = 0.f; acc[2 * q + 1] = 0.f; }
    for (int q = 0; q < chunks; ++q) {
        const int c = q * 64 + 2 * lane;
        float a0 = p.b ? __ldg(p.b + c) : 0.f, a1 = p.b ? __ldg(p.b + c + 1) : 0.f;
        for (int dy = -r; dy <= r; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= p.H) continue;
            for (int dx = -r; dx <= r; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= p.W) continue;
                const float2 wv = __ldg(reinterpret_cast<const float2*>(p.w + (size_t)((dy + r) * p.k + (dx + r)) * p.C + c));
                const size_t pix = img + (size_t)yy * p.W + xx;
                const float v0 = act_load1(p.in, pix, c), v1 = act_load1(p.in, pix, c + 1);
                a0 = fmaf(v0, wv.x, a0);
                a1 = fmaf(v1, wv.y, a1);
            }
        }
        acc[2 * q] = a0; acc[2 * q + 1] = a1;
    }
    float s = 0.f;
    for (int q = 0; q < chunks; ++q) s += acc[2 * q] + acc[2 * q + 1];
    const float mean = warp_sum(s) / (float)p.C;
    float v = 0.f;
    for (int q = 0; q < chunks; ++q) { const float d0 = acc[2 * q] - mean, d1 = acc[2 * q + 1] - mean; v += d0 * d0 + d1 * d1; }
    const float rstd = rsqrtf(warp_sum(v) / (float)p.C + p.eps);
    const size_t opix = img + (size_t)y * p.W + x;
    for (int q = 0; q < chunks; ++q) {
        const int c = q * 64 + 2 * lane;
        act_store1(p.out, opix, c, (acc[2 * q] - mean) * rstd * __ldg(p.lnw + c) + __ldg(p.lnb + c));
        act_store1(p.out, opix, c + 1, (acc[2 * q + 1] - mean) * rstd * __ldg(p.lnw + c + 1) + __ldg(p.lnb + c + 1));
    }
}

// 3x3 / stride-2 / pad-1 max pooling (torchvision ResNet stem), channels-last, thread = (pixel, 4 channels)
__global__ void __launch_bounds__(256)
k_maxpool3s2(ActV in, ActV out, int N, int H, int W, int C, int Ho, int Wo, int d2s) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cq = C / 4;
    if (t >= (long long)N * Ho * Wo * cq) return;
    const int c = (int)(t % cq) * 4;
    long long pix = t / cq;
    const int ox = (int)(pix % Wo); pix /= Wo;
    const int oy = (int)(pix % Ho); const int n = (int)(pix / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dy = 0; dy < 3; ++dy) {
        const int y = 2 * oy - 1 + dy;
        if (y < 0 || y >= H) continue;
        for (int dx = 0; dx < 3; ++dx) {
            const int x = 2 * ox - 1 + dx;
            if (x < 0 || x >= W) continue;
            // d2s: the (H, W, C) input is stored depth-to-space folded as (H/2, W/2, 4C), channel block = (y & 1) * 2 + (x & 1)
            // (the phase-major output of the space-to-depth ResNet stem convolution)
            const float4 v = d2s ? act_load4(in, ((size_t)n * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1), (((y & 1) << 1) | (x & 1)) * C + c)
                                 : act_load4(in, ((size_t)n * H + y) * W + x, c);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    act_store4(out, ((size_t)n * Ho + oy) * Wo + ox, c, m);
}

static inline ActV view_of(const heal_act_t* a) {
    ActV v;
    v.p = a->data; v.fmt = a->fmt; v.cs = a->cstride; v.co = a->coffset; v.plane = a->plane_stride;
    return v;
}

}  // namespace

extern "C" int heal_dwconv_layernorm(const heal_act_t* in, int N, int H, int W, int C, const float* dw_weight, const float* dw_bias,
                                     int ksize, const float* ln_weight, const float* ln_bias, float eps, const heal_act_t* out,
                                     void* stream_) {
    if (!in || !in->data || !out || !out->data || !dw_weight || !ln_weight || !ln_bias) return HEAL_ERR_ARG;
    if ((C % 64) || C > 256 || ksize < 1 || !(ksize & 1) || ksize > 11) return HEAL_ERR_UNSUPPORTED;
    DwLnP p;
    p.in = view_of(in); p.out = view_of(out); p.w = dw_weight; p.b = dw_bias; p.lnw = ln_weight; p.lnb = ln_bias;
    p.N = N; p.H = H; p.W = W; p.C = C; p.k = ksize; p.eps = eps;
    dim3 grid((W + 7) / 8, (H + 3) / 4, N);
    k_dwconv_ln<<<grid, 1024, 0, (cudaStream_t)stream_>>>(p);
    return heal_check_launch();
}

extern "C" int heal_maxpool3x3s2(const heal_act_t* in, int N, int H, int W, int C, int depth_to_space_in, const heal_act_t* out, void* stream_) {
    if (!in || !in->data || !out || !out->data) return HEAL_ERR_ARG;
    if (depth_to_space_in && ((H | W) & 1)) return HEAL_ERR_UNSUPPORTED;
    if ((C & 3) || (in->cstride & 3) || (in->coffset & 3) || (out->cstride & 3) || (out->coffset & 3)) return HEAL_ERR_UNSUPPORTED;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo * (C / 4);
    k_maxpool3s2<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(view_of(in), view_of(out), N, H, W, C, Ho, Wo, depth_to_space_in);
    return heal_check_launch();
}

// fp32 <-> bf16 / split-bf16 activation conversion (boundary plumbing of the tensor-core path).
#include "common.cuh"
#include "../../include/heal_b200.h"

namespace {
__global__ void k_act_convert(ActV src, ActV dst, size_t npix, int chunks) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = npix * (size_t)chunks;
    for (; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t pix = i / chunks;
        int c = (int)(i % chunks) * 4;
        act_store4(dst, pix, c, act_load4(src, pix, c));
    }
}
}  // namespace

extern "C" int heal_act_convert(const heal_act_t* src, const heal_act_t* dst, size_t num_pixels, int channels, void* stream_) {
    if (!src || !dst || !src->data || !dst->data) return HEAL_ERR_ARG;
    if ((channels & 3) || (src->cstride & 3) || (src->coffset & 3) || (dst->cstride & 3) || (dst->coffset & 3)) return HEAL_ERR_UNSUPPORTED;
    if (num_pixels == 0) return HEAL_OK;
    ActV s, d;
    s.p = src->data; s.fmt = src->fmt; s.cs = src->cstride; s.co = src->coffset; s.plane = src->plane_stride;
    d.p = dst->data; d.fmt = dst->fmt; d.cs = dst->cstride; d.co = dst->coffset; d.plane = dst->plane_stride;
    size_t total = num_pixels * (size_t)(channels / 4);
    size_t blocks = (total + 255) / 256;
    size_t cap = (size_t)HEAL_NUM_SMS * 16;
    k_act_convert<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream_>>>(s, d, num_pixels, channels / 4);
    return heal_check_launch();
}

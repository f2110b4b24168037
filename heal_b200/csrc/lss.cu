// Lift-Splat-Shoot: frustum geometry -> BEV cell index, and the fused depth-softmax (x) feature outer
// product + BEV pooling.
// Reference: opencood/models/heter_encoders.py:125-147 (get_geometry), :161-217 (voxel_pooling, incl. the
// truncating `.long()` index :173 and the (B,C,Y,X) output layout :206-212),
// opencood/models/sub_modules/lss_submodule.py:132-134 / :227-229 (softmax over D, depth[:,None]*feat[:,:,None]),
// opencood/utils/camera_utils.py:220-246 (QuickCumsum = per-cell sum).
// The (n*cams, C, D, fH, fW) outer-product tensor of the reference (277 MB per agent at 704x256) is never
// materialised: each block stages one image row of depth logits and features in shared memory, computes the
// softmax in place and accumulates prob*feat straight into the channels-last BEV map.
#include <cub/cub.cuh>
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

// ---- deterministic BEV pooling: cell-sorted interval reduction (no atomics) -----------------------------------------------------
// keys = agent * cells + cell (invalid points get the one-past-the-end key), values = frustum point id; a STABLE radix sort puts
// every BEV cell's points next to each other in ascending point order; one warp per cell then sums prob * feature over its
// interval in that fixed order and writes the cell's C channels once (empty cells are written as zeros: no memset pass).
struct SortedP {
    const float* logits; long long l_img, l_d, l_pix;     // depth logits, element strides (NCHW trunk: HW*D, HW, 1; NHWC heads: HW*S, 1, S)
    const float* feat;   long long f_img, f_c, f_pix;     // image features
    const int* cell;                                        // (BN, D, fH, fW)
    int BN, cams, D, C, HW, cells_per_agent, agents;
    float* prob;                                            // (BN, D, HW) workspace
    ActV out;                                               // (agents, cells_per_agent, C) channels-last, any storage format
};

__global__ void k_lss_keys(const int* __restrict__ cell, long long n, int DHW, int cams, int cells_per_agent, unsigned invalid,
                           unsigned* __restrict__ keys, int* __restrict__ vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell[i];
    const int agent = (int)(i / DHW) / cams;
    keys[i] = (c < 0) ? invalid : (unsigned)agent * (unsigned)cells_per_agent + (unsigned)c;
    vals[i] = (int)i;
}

// softmax over D for every (image, pixel): one warp per pixel, lanes over depth bins (D <= 64); probabilities are stored
// pixel-major [image][pixel][D] so that both this kernel's writes and the pooling kernel's reads are contiguous per pixel
__global__ void __launch_bounds__(256)
k_lss_prob(SortedP p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 8 + warp;
    if (t >= (long long)p.BN * p.HW) return;
    const int bn = (int)(t / p.HW), pix = (int)(t % p.HW);
    const float* lg = p.logits + (long long)bn * p.l_img + (long long)pix * p.l_pix;
    const float v0 = (lane < p.D) ? __ldg(lg + (long long)lane * p.l_d) : -INFINITY;
    const float v1 = (lane + 32 < p.D) ? __ldg(lg + (long long)(lane + 32) * p.l_d) : -INFINITY;
    float mx = fmaxf(v0, v1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e0 = (lane < p.D) ? expf(v0 - mx) : 0.f, e1 = (lane + 32 < p.D) ? expf(v1 - mx) : 0.f;
    const float s = warp_sum(e0 + e1);
    float* pr = p.prob + (size_t)t * p.D;
    if (lane < p.D) pr[lane] = e0 / s;
    if (lane + 32 < p.D) pr[lane + 32] = e1 / s;
}

// Balanced interval pooling.  The sorted point list is cut into chunks of LSS_CH points, one block per chunk, thread = channel.
// (BEV cells next to the cameras collect hundreds of frustum points, cells far away one or two: a cell-per-warp or cells-per-block
// split leaves a few blocks with 10^4 sequential iterations; equal point counts per block do not.)  A thread walks its chunk in
// order; a cell that begins AND ends inside the chunk is written straight to the (pre-zeroed) BEV map; the partial sum of a cell
// that crosses a chunk boundary goes to part[chunk][0 = head (cell began earlier) | 1 = tail (cell continues)].  k_lss_pool_fixup
// then adds, for every cell with a tail partial, the head partials of the following chunks in chunk order.  Summation order is
// fixed (ascending point id, grouped by chunk): bit-identical from run to run.
constexpr int LSS_CH = 128;
constexpr unsigned LSS_NONE = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256)
k_lss_pool_chunks(SortedP p, const unsigned* __restrict__ keys, const int* __restrict__ vals, long long n, unsigned total_cells,
                  float* __restrict__ part, unsigned* __restrict__ pkey) {
    __shared__ unsigned s_key[LSS_CH];
    __shared__ float s_pr[LSS_CH];
    __shared__ long long s_fo[LSS_CH];
    __shared__ unsigned s_edge[2];
    const long long base = (long long)blockIdx.x * LSS_CH;
    const int tid = threadIdx.x, ch = tid;
    const int DHW = p.D * p.HW;
    if (tid < LSS_CH) {
        const long long j = base + tid;
        unsigned k = (j < n) ? keys[j] : LSS_NONE;
        if (k >= total_cells) k = LSS_NONE;
        s_key[tid] = k;
        if (k != LSS_NONE) {
            const int i = vals[j];
            const int bn = i / DHW, rem = i - bn * DHW;
            const int d = rem / p.HW, pix = rem - d * p.HW;
            s_pr[tid] = p.prob[((size_t)bn * p.HW + pix) * p.D + d];
            s_fo[tid] = (long long)bn * p.f_img + (long long)pix * p.f_pix;
        }
    }
    if (tid == 0) {
        s_edge[0] = (base > 0) ? keys[base - 1] : LSS_NONE;
        unsigned nx = (base + LSS_CH < n) ? keys[base + LSS_CH] : LSS_NONE;
        s_edge[1] = (nx >= total_cells) ? LSS_NONE : nx;
    }
    __syncthreads();
    // number of valid points in this chunk (valid keys sort before the invalid tail)
    const int nvalid = __syncthreads_count(tid < LSS_CH && s_key[tid < LSS_CH ? tid : 0] != LSS_NONE);
    if (ch >= p.C) return;
    unsigned cur = s_key[0];
    if (nvalid == 0) return;                           // the whole chunk lies in the invalid tail
    bool started_here = (s_edge[0] != cur);
    float acc = 0.f;
    float* mypart = part + (size_t)blockIdx.x * 2 * p.C;
    const float* fch = p.feat + (long long)ch * p.f_c;
    auto step = [&](int t, float f) {
        const unsigned k = s_key[t];
        if (k != cur) {                                // the cell `cur` ended inside this chunk
            if (started_here) act_store1(p.out, (size_t)cur, ch, acc);
            else { mypart[ch] = acc; if (ch == 0) pkey[2 * blockIdx.x] = cur; }
            acc = 0.f; cur = k; started_here = true;
        }
        acc = fmaf(s_pr[t], f, acc);
    };
    int t = 0;
    for (; t + 4 <= nvalid; t += 4) {                  // four independent feature loads in flight, then the ordered accumulation
        const float f0 = __ldg(fch + s_fo[t]), f1 = __ldg(fch + s_fo[t + 1]), f2 = __ldg(fch + s_fo[t + 2]), f3 = __ldg(fch + s_fo[t + 3]);
        step(t, f0); step(t + 1, f1); step(t + 2, f2); step(t + 3, f3);
    }
    for (; t < nvalid; ++t) step(t, __ldg(fch + s_fo[t]));
    const unsigned nextk = (t < LSS_CH) ? LSS_NONE : s_edge[1];
    const bool ended_here = (nextk != cur);
    if (started_here && ended_here) act_store1(p.out, (size_t)cur, ch, acc);
    else {
        const int slot = started_here ? 1 : 0;         // tail partial (cell continues) | head partial (also: a cell covering the whole chunk)
        mypart[slot * p.C + ch] = acc;
        if (ch == 0) pkey[2 * blockIdx.x + slot] = cur;
    }
}

__global__ void __launch_bounds__(256)
k_lss_pool_fixup(SortedP p, const float* __restrict__ part, const unsigned* __restrict__ pkey, int nchunks) {
    const int b = blockIdx.x, ch = threadIdx.x;
    const unsigned key = pkey[2 * b + 1];
    if (key == LSS_NONE || ch >= p.C) return;
    float acc = part[((size_t)b * 2 + 1) * p.C + ch];
    for (int b2 = b + 1; b2 < nchunks && pkey[2 * b2] == key; ++b2) acc += part[(size_t)b2 * 2 * p.C + ch];
    act_store1(p.out, (size_t)key, ch, acc);
}

struct SortedWs { size_t keys_in, keys_out, vals_in, vals_out, prob, part, pkey, cub, total, cub_bytes; };

SortedWs sorted_layout(long long npts, long long total_cells, int end_bit, int C) {
    SortedWs L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = heal_align_up(o + bytes, 256); return r; };
    L.keys_in = take(npts * 4); L.keys_out = take(npts * 4); L.vals_in = take(npts * 4); L.vals_out = take(npts * 4);
    const long long nchunks = (npts + LSS_CH - 1) / LSS_CH;
    L.prob = take(npts * 4); L.part = take((size_t)nchunks * 2 * C * 4); L.pkey = take((size_t)nchunks * 2 * 4);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr,
                                    (int)npts, 0, end_bit);
    L.cub_bytes = tb;
    L.cub = take(tb);
    L.total = o + 256;
    return L;
}

int key_bits(long long total_cells) { int b = 1; while ((1LL << b) <= total_cells) ++b; return b; }

}  // namespace

extern "C" size_t heal_lss_pool_sorted_workspace(int num_images, int D, int fH, int fW, int agents, int cells_per_agent) {
    const long long npts = (long long)num_images * D * fH * fW, cells = (long long)agents * cells_per_agent;
    if (npts <= 0 || cells <= 0 || npts >= (1LL << 31)) return 0;
    return sorted_layout(npts, cells, key_bits(cells), 256).total;
}

extern "C" int heal_lss_pool_sorted(const float* depth_logits, long long l_img, long long l_d, long long l_pix,
                                    const float* feat, long long f_img, long long f_c, long long f_pix,
                                    const int* cell, int num_images, int cams_per_agent, int D, int C, int fH, int fW,
                                    int cells_per_agent, const heal_act_t* bev_out, void* workspace, size_t workspace_bytes, void* stream_) {
    if (!depth_logits || !feat || !cell || !bev_out || !bev_out->data || !workspace) return HEAL_ERR_ARG;
    if (num_images <= 0) return HEAL_OK;
    if (cams_per_agent < 1 || (num_images % cams_per_agent) || C < 1 || C > 256) return HEAL_ERR_ARG;
    const int agents = num_images / cams_per_agent;
    const long long npts = (long long)num_images * D * fH * fW, cells = (long long)agents * cells_per_agent;
    if (npts >= (1LL << 31) || cells >= (1LL << 31)) return HEAL_ERR_UNSUPPORTED;
    const int bits = key_bits(cells);
    SortedWs L = sorted_layout(npts, cells, bits, 256);
    if (workspace_bytes < L.total) return HEAL_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    cudaStream_t st = (cudaStream_t)stream_;
    SortedP p;
    p.logits = depth_logits; p.l_img = l_img; p.l_d = l_d; p.l_pix = l_pix;
    p.feat = feat; p.f_img = f_img; p.f_c = f_c; p.f_pix = f_pix;
    p.cell = cell; p.BN = num_images; p.cams = cams_per_agent; p.D = D; p.C = C; p.HW = fH * fW;
    p.cells_per_agent = cells_per_agent; p.agents = agents; p.prob = (float*)(ws + L.prob);
    p.out.p = bev_out->data; p.out.fmt = bev_out->fmt; p.out.cs = bev_out->cstride; p.out.co = bev_out->coffset; p.out.plane = bev_out->plane_stride;
    unsigned* k_in = (unsigned*)(ws + L.keys_in); unsigned* k_out = (unsigned*)(ws + L.keys_out);
    int* v_in = (int*)(ws + L.vals_in); int* v_out = (int*)(ws + L.vals_out);
    float* part = (float*)(ws + L.part); unsigned* pkey = (unsigned*)(ws + L.pkey);
    const int nchunks = (int)((npts + LSS_CH - 1) / LSS_CH);
    if (D > 64) return HEAL_ERR_UNSUPPORTED;
    const unsigned gp = (unsigned)((npts + 255) / 256);
    k_lss_keys<<<gp, 256, 0, st>>>(cell, npts, D * fH * fW, cams_per_agent, cells_per_agent, (unsigned)cells, k_in, v_in);
    k_lss_prob<<<(unsigned)(((long long)num_images * p.HW + 7) / 8), 256, 0, st>>>(p);
    size_t tb = L.cub_bytes;
    if (cub::DeviceRadixSort::SortPairs(ws + L.cub, tb, k_in, k_out, v_in, v_out, (int)npts, 0, bits, st) != cudaSuccess) return HEAL_ERR_LAUNCH;
    cudaMemsetAsync(pkey, 0xFF, (size_t)nchunks * 2 * 4, st);
    const int threads = (C + 31) / 32 * 32 < LSS_CH ? LSS_CH : (C + 31) / 32 * 32;
    k_lss_pool_chunks<<<nchunks, threads, 0, st>>>(p, k_out, v_out, npts, (unsigned)cells, part, pkey);
    k_lss_pool_fixup<<<nchunks, threads, 0, st>>>(p, part, pkey, nchunks);
    return heal_check_launch(6);
}

// Per-camera 3x3 algebra of get_geometry (lss_submodule / heter_encoders.py:135,142): post_inv = inverse(post_rots),
// combine = rots @ inverse(intrins).  torch.inverse is not capturable in a CUDA graph (cuSOLVER/cuBLAS pointer-array setup), so the
// two inverses are closed-form adjugates evaluated in fp64 and rounded to fp32 (within 1 ulp of the exact inverse; the reference's
// fp32 LU differs from it by a few ulp, which only matters for frustum points that sit on a cell boundary).
__device__ __forceinline__ void inv3x3_f64(const float* m, double* o) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    const double r = 1.0 / det;
    o[0] = A * r; o[1] = -(b * i - c * h) * r; o[2] = (b * f - c * e) * r;
    o[3] = B * r; o[4] = (a * i - c * g) * r;  o[5] = -(a * f - c * d) * r;
    o[6] = C * r; o[7] = -(a * h - b * g) * r; o[8] = (a * e - b * d) * r;
}

__global__ void k_lss_camera_matrices(const float* __restrict__ rots, const float* __restrict__ intrins, const float* __restrict__ post_rots,
                                      int n, float* __restrict__ post_inv, float* __restrict__ combine) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double pi[9], ii[9];
    inv3x3_f64(post_rots + 9 * c, pi);
    inv3x3_f64(intrins + 9 * c, ii);
    float iif[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { post_inv[9 * c + j] = (float)pi[j]; iif[j] = (float)ii[j]; }
    const float* R = rots + 9 * c;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k)        // fp32 matmul of the fp32 inverse, as rots.matmul(torch.inverse(intrins))
            combine[9 * c + 3 * r + k] = fmaf(R[3 * r + 2], iif[6 + k], fmaf(R[3 * r + 1], iif[3 + k], R[3 * r] * iif[k]));
}

extern "C" int heal_lss_camera_matrices(const float* rots, const float* intrins, const float* post_rots, int num_images,
                                        float* post_rots_inv_out, float* combine_out, void* stream_) {
    if (!rots || !intrins || !post_rots || !post_rots_inv_out || !combine_out) return HEAL_ERR_ARG;
    if (num_images <= 0) return HEAL_OK;
    k_lss_camera_matrices<<<(num_images + 63) / 64, 64, 0, (cudaStream_t)stream_>>>(rots, intrins, post_rots, num_images,
                                                                                   post_rots_inv_out, combine_out);
    return heal_check_launch();
}

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

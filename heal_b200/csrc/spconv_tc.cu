// Sparse 3-D convolution gather-GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) for the
// SECOND middle encoder (VoxelBackBone8x; reference layer list: opencood/models/sub_modules/sparse_backbone_3d.py:48-91,
// post_act_block :11-30 = conv + BatchNorm1d(eps 1e-3) + ReLU; spconv semantics restated in oracle/sparse_conv.py).
//
// Output-stationary rulebook (csrc/spconv.cu): nbr[out_row][k] = input row under kernel offset k, or -1.
// GEMM view per 128-row output tile: D[128 x Cout] = sum over K-blocks  A_kb[128 x 64] . W_kb[64 x Cout]
//   A_kb row r = the gathered input rows of out-row r for the 64/Cin kernel offsets of K-block kb, side by side
//                (Cin = 64: one offset per K-block; 32: two; 16: four), zero where the rulebook says -1;
//   W_kb       = the matching offsets' (Cin x Cout) weight blocks stacked along K (zero rows for padding offsets).
// Features travel between layers as "split rows": (rows, 2*C) bf16 = [hi C | lo C], x ~= hi + lo (16 mantissa bits),
// and a product is a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with fp32 accumulation: fp32-equivalent (same scheme as conv2d_tc.cu).
//
// Roles (288 threads, persistent over tiles):
//   warps 0..3  gather producers: the tile's rulebook rows go to shared memory once; per ACTIVE K-block (a K-block none of
//               the tile's 128 rows hits is skipped altogether) every thread issues sixteen 16-byte cp.async copies (8 lanes
//               per gathered row: full 128 B lines; src-size 0 zero-fills a miss) into the 128B-swizzled K-major A tile, and
//               one lane TMA-loads the K-block's weights.  Copies stay in flight across K-blocks and tiles (cp.async groups,
//               two behind), then fence.proxy.async + mbarrier arrive hand the stage to the tensor core.
//   warp  4     MMA issuer: 4 K-steps x { a_hi x [b_hi | b_lo] (N' = 2 Cout), a_lo x b_hi } per stage, double-buffered TMEM.
//   warps 5..8  epilogue: TMEM -> registers -> main + aux -> + BN shift -> ReLU -> split rows (or fp32 rows) to global.
#include <cuda.h>
#include "common.cuh"
#include "tc_prims.cuh"
#include "../../include/heal_b200.h"

namespace {

constexpr int SPT_THREADS = 288;
constexpr int SPT_M = 128;
constexpr int SPT_ATILE = SPT_M * 128;          // one plane of A: 128 rows x 128 B

struct SptP {
    const __nv_bfloat16* in;     // split rows: hi channels at in + row*in_pitch, lo channels at + in_lo (interleaved rows: pitch 2*CIN, lo CIN;
                                 // a planar split `Act`: pitch = pixel stride, lo = plane stride)
    long long in_pitch, in_lo, out_pitch, out_lo;
    const int* nbr;              // (capacity, K)
    const int* m_dev;            // live output rows on the device (or nullptr)
    int M;                       // capacity
    int K, KB;                   // kernel offsets, K-blocks = ceil(K / (64 / CIN))
    const float* bias;           // (COUT) folded BatchNorm shift
    int relu;
    __nv_bfloat16* out_split;    // (capacity, 2*COUT) or nullptr
    float* out_f32;              // (capacity, COUT) or nullptr
};

__device__ __forceinline__ void prod_bar() { asm volatile("bar.sync 3, 128;" ::: "memory"); }

template <int CIN, int COUT, int STAGES>
__global__ void __launch_bounds__(SPT_THREADS, 1)
k_spconv_tc(const __grid_constant__ CUtensorMap tmB, const SptP p) {
    constexpr int TPK = 64 / CIN;                   // kernel offsets per 64-wide K-block
    constexpr int CPT = CIN / 8;                    // 16-byte chunks per gathered row and plane
    constexpr int B_BYTES = 2 * COUT * 128;         // [plane][COUT rows][128 B]
    constexpr int STAGE_BYTES = 2 * SPT_ATILE + B_BYTES;
    constexpr int ACC_STRIDE = 2 * COUT;            // main | aux accumulator columns
    constexpr int TMEM_COLS = (2 * ACC_STRIDE < 32) ? 32 : 2 * ACC_STRIDE;
    constexpr int LA = (STAGES >= 4) ? 2 : 1;       // cp.async groups a producer keeps in flight behind the one it issues
    constexpr int KPAD = 28;                        // rulebook row pitch in shared memory (K <= 27)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uintptr_t raw_addr = reinterpret_cast<uintptr_t>(smem_raw);
    asm volatile("" : "+l"(raw_addr));
    uint8_t* smem = reinterpret_cast<uint8_t*>(raw_addr);
    if (smem_u32(smem) & 1023u) __trap();
    int* snbr = reinterpret_cast<int*>(smem + (size_t)STAGES * STAGE_BYTES);
    int* sflags = snbr + SPT_M * KPAD;                                  // [STAGES] bit0 = first K-block of its tile, bit1 = last
    unsigned* smask = reinterpret_cast<unsigned*>(sflags + STAGES);     // [2] active-K-block mask of the tile being set up
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(smask + 2) + 7) & ~(uintptr_t)7);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES), bar_tempty = smem_u32(bars + 2 * STAGES + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 128 + 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int Md = p.m_dev ? min(p.M, p.m_dev[0]) : p.M;
    const int ntiles = (Md + SPT_M - 1) / SPT_M;

    if (warp < 4) {
        // ================================ gather producers ================================
        const int tid = threadIdx.x;                 // 0..127
        int stage = 0; uint32_t phase = 0;
        int pend[LA + 1]; int npend = 0;
#pragma unroll
        for (int i = 0; i <= LA; ++i) pend[i] = 0;
        // The tile's rulebook rows (128 x K ints, contiguous) are prefetched into REGISTERS one tile ahead: the loads fly under the
        // current tile's gathers, so a tile starts without an exposed global-memory round trip (the first version re-read them at
        // every tile start: ~2 us of bubble per tile, profiles/ncu_full_r2_summary.json: long_scoreboard-bound at 14 % occupancy)
        int nxt[27];
        auto fetch = [&](int tile_) {
            const long long base = (long long)tile_ * SPT_M * p.K;
            const int lim = min(SPT_M, Md - tile_ * SPT_M) * p.K;
#pragma unroll
            for (int q = 0; q < 27; ++q) {
                const int i = tid + 128 * q;
                nxt[q] = (q < p.K && i < lim) ? __ldg(p.nbr + base + i) : -1;
            }
        };
        if ((int)blockIdx.x < ntiles) fetch(blockIdx.x);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            prod_bar();                               // every producer has issued the previous tile's copies (they read snbr)
            if (tid == 0) smask[0] = 0u;
#pragma unroll
            for (int q = 0; q < 27; ++q) {
                if (q < p.K) {
                    const int i = tid + 128 * q;
                    const int r = i / p.K, k = i - r * p.K;
                    snbr[r * KPAD + k] = nxt[q];
                }
            }
            if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x);
            prod_bar();
            {   // active K-blocks of this tile = OR over its rows
                unsigned m = 0;
                for (int k = 0; k < p.K; ++k) if (snbr[tid * KPAD + k] >= 0) m |= 1u << (k / TPK);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
                if (lane == 0 && m) atomicOr(smask, m);
            }
            prod_bar();
            unsigned mask = smask[0];
            if (mask == 0u) mask = 1u;                // a tile always produces one stage (never happens: the centre / generating offset hits)
            const int kb_first = __ffs(mask) - 1, kb_last = 31 - __clz(mask);
            for (int kb = kb_first; kb <= kb_last; ++kb) {
                if (!((mask >> kb) & 1u)) continue;
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t sa = smem_base + stage * STAGE_BYTES;
                if (warp == 0) {
                    if (elect_one()) {
                        sflags[stage] = (kb == kb_first ? 1 : 0) | (kb == kb_last ? 2 : 0);
                        mbar_expect_tx(bar_full + 8 * stage, (uint32_t)B_BYTES);
                        tma_load_3d(sa + 2 * SPT_ATILE, &tmB, bar_full + 8 * stage, 0, kb * COUT, 0);
                    }
                    __syncwarp();
                }
                // 8 lanes per gathered row: lane & 7 = 16-byte chunk of the 128-byte K-block row, lane >> 3 = row within a group of 4
                const int c = lane & 7;
                const int tl = c / CPT, sc = c - tl * CPT;          // offset within the K-block, chunk within that offset's row
                const int tap = kb * TPK + tl;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = warp * 32 + it * 4 + (lane >> 3);
                    const int src = (tap < p.K) ? snbr[r * KPAD + tap] : -1;
                    const __nv_bfloat16* g = p.in + (src >= 0 ? (long long)src * p.in_pitch + sc * 8 : 0);
                    const uint32_t nb = src >= 0 ? 16u : 0u;
                    const uint32_t d = sa + r * 128 + ((c ^ (r & 7)) << 4);
                    cp_async16(d, g, nb);                            // hi plane
                    cp_async16(d + SPT_ATILE, g + (src >= 0 ? p.in_lo : 0), nb);   // lo plane
                }
                cp_async_commit();
                pend[npend++] = stage;
                if (npend > LA) {
                    cp_async_wait<LA>();
                    fence_async_smem();
                    mbar_arrive(bar_full + 8 * pend[0]);
#pragma unroll
                    for (int i = 0; i < LA; ++i) pend[i] = pend[i + 1];
                    --npend;
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        cp_async_wait<0>();
        fence_async_smem();
        for (int i = 0; i < npend; ++i) mbar_arrive(bar_full + 8 * pend[i]);
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        const uint32_t idesc_cat = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * COUT) >> 3) << 17) | ((uint32_t)(SPT_M >> 4) << 24);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(SPT_M >> 4) << 24);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_STRIDE);
            while (true) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const int fl = sflags[stage];
                const uint32_t sa = smem_base + stage * STAGE_BYTES;
                const uint32_t sb = sa + 2 * SPT_ATILE;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t a_hi = umma_desc_sw128(sa + k * 32), a_lo = umma_desc_sw128(sa + SPT_ATILE + k * 32);
                        const uint64_t b = umma_desc_sw128(sb + k * 32);
                        umma_bf16(tmem_d, a_hi, b, idesc_cat, ((fl & 1) && k == 0) ? 0u : 1u);   // [a_hi*b_hi | a_hi*b_lo]
                        umma_bf16(tmem_d, a_lo, b, idesc, 1u);                                   // += a_lo*b_hi
                    }
                    umma_commit(bar_empty + 8 * stage);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                if (fl & 2) break;
            }
            if (elect_one()) umma_commit(bar_tfull + 8 * acc);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ================================ epilogue ================================
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access (warps 5,6,7,8 -> 1,2,3,0)
        const int row = quarter * 32 + lane;
        constexpr int CH = (COUT >= 32) ? 32 : 16;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int grow = tile * SPT_M + row;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t tb = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += CH) {
                uint32_t x[CH], y[CH];
                if constexpr (CH == 32) { tmem_ld32(tb + (uint32_t)c0, x); tmem_ld32(tb + (uint32_t)(COUT + c0), y); }
                else { tmem_ld16(tb + (uint32_t)c0, x); tmem_ld16(tb + (uint32_t)(COUT + c0), y); }
                tmem_wait_ld();
                if (grow < Md) {
#pragma unroll
                    for (int g8 = 0; g8 < CH / 8; ++g8) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j] = __uint_as_float(x[g8 * 8 + j]) + __uint_as_float(y[g8 * 8 + j]) + (p.bias ? __ldg(p.bias + c0 + g8 * 8 + j) : 0.f);
                            if (p.relu) v[j] = fmaxf(v[j], 0.f);
                        }
                        if (p.out_split) {
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float h0 = __bfloat162float(__float2bfloat16_rn(v[2 * j])), h1 = __bfloat162float(__float2bfloat16_rn(v[2 * j + 1]));
                                hw[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                                lw[j] = pack_bf16(v[2 * j] - h0, v[2 * j + 1] - h1);
                            }
                            __nv_bfloat16* o = p.out_split + (long long)grow * p.out_pitch + c0 + g8 * 8;
                            *reinterpret_cast<uint4*>(o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                            *reinterpret_cast<uint4*>(o + p.out_lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        }
                        if (p.out_f32) {
                            float* o = p.out_f32 + (size_t)grow * COUT + c0 + g8 * 8;
                            stg_f4(o, make_float4(v[0], v[1], v[2], v[3]));
                            stg_f4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int CIN, int COUT>
int launch_spt(const SptP& p, const void* w_packed, cudaStream_t st) {
    constexpr int STAGES = (COUT >= 128) ? 3 : 4;
    constexpr int B_BYTES = 2 * COUT * 128;
    constexpr size_t SMEM = (size_t)STAGES * (2 * SPT_ATILE + B_BYTES) + SPT_M * 28 * 4 + 512;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    PFN_tmEncodeTiled enc = get_encode();
    if (!enc) return HEAL_ERR_DRIVER;
    CUtensorMap tmB;
    {
        cuuint64_t d[3] = {64, (cuuint64_t)p.KB * COUT, 2};
        cuuint64_t s[2] = {128, (cuuint64_t)p.KB * COUT * 128};
        cuuint32_t b[3] = {64u, (cuuint32_t)COUT, 2u};
        cuuint32_t es[3] = {1, 1, 1};
        if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)w_packed, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HEAL_ERR_DRIVER;
    }
    static size_t attr_set[HEAL_MAX_DEVICES] = {};
    if (!heal_ensure_dyn_smem(k_spconv_tc<CIN, COUT, STAGES>, SMEM, attr_set)) return HEAL_ERR_LAUNCH;
    int tiles = (p.M + SPT_M - 1) / SPT_M;
    int grid = tiles < HEAL_NUM_SMS ? tiles : HEAL_NUM_SMS;
    k_spconv_tc<CIN, COUT, STAGES><<<grid, SPT_THREADS, SMEM, st>>>(tmB, p);
    return heal_check_launch();
}

// fp32 rows -> split rows [hi C | lo C]
__global__ void k_rows_to_split(const float* __restrict__ in, const int* __restrict__ m_dev, int M, int C, __nv_bfloat16* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Md = m_dev ? min(M, m_dev[0]) : M;
    if (t >= (long long)Md * C) return;
    const int r = (int)(t / C), c = (int)(t % C);
    const float v = in[t];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[(size_t)r * 2 * C + c] = h;
    out[(size_t)r * 2 * C + C + c] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// PointPillars stem as a sparse 2-D convolution: rulebook of a k x k / stride-2 conv from the pillar list (cell -> pillar row id
// map) to the DENSE output grid.  nbr[(b*Ho + oy)*Wo + ox][r*k + s] = idmap[b][2oy + r - pad][2ox + s - pad] or -1.
__global__ void k_stem_rulebook(const int* __restrict__ idmap, int B, int ny, int nx, int Ho, int Wo, int k, int pad, int* __restrict__ nbr) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int K = k * k;
    if (t >= (long long)B * Ho * Wo * K) return;
    const int tap = (int)(t % K);
    long long pix = t / K;
    const int ox = (int)(pix % Wo); pix /= Wo;
    const int oy = (int)(pix % Ho); const int b = (int)(pix / Ho);
    const int y = 2 * oy + tap / k - pad, x = 2 * ox + tap % k - pad;
    nbr[t] = (y >= 0 && y < ny && x >= 0 && x < nx) ? __ldg(idmap + ((size_t)b * ny + y) * nx + x) : -1;
}

}  // namespace

extern "C" int heal_stem_rulebook(const int* idmap, int batch, int ny, int nx, int ksize, int pad, int* nbr_out, void* stream_) {
    if (!idmap || !nbr_out || batch < 1 || (ny & 1) || (nx & 1) || ksize < 1) return HEAL_ERR_ARG;
    const int Ho = ny / 2, Wo = nx / 2;
    const long long total = (long long)batch * Ho * Wo * ksize * ksize;
    k_stem_rulebook<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(idmap, batch, ny, nx, Ho, Wo, ksize, pad, nbr_out);
    return heal_check_launch();
}

extern "C" int heal_spconv_gather_gemm_tc(const void* in_split_rows, const int* nbr, const int* out_rows_dev, int out_capacity, int kvol,
                                          const void* w_packed, const float* bias, int c_in, int c_out, int relu,
                                          void* out_split_rows, float* out_f32,
                                          long long in_pitch, long long in_lo, long long out_pitch, long long out_lo, void* stream_) {
    if (!in_split_rows || !nbr || !w_packed || (!out_split_rows && !out_f32) || out_capacity < 1) return HEAL_ERR_ARG;
    if (kvol < 1 || kvol > 27) return HEAL_ERR_UNSUPPORTED;
    if (c_in != 16 && c_in != 32 && c_in != 64) return HEAL_ERR_UNSUPPORTED;
    SptP p;
    p.in = (const __nv_bfloat16*)in_split_rows; p.nbr = nbr; p.m_dev = out_rows_dev; p.M = out_capacity;
    p.K = kvol; p.KB = (kvol + (64 / c_in) - 1) / (64 / c_in);
    p.bias = bias; p.relu = relu; p.out_split = (__nv_bfloat16*)out_split_rows; p.out_f32 = out_f32;
    p.in_pitch = in_pitch > 0 ? in_pitch : 2 * c_in; p.in_lo = in_lo > 0 ? in_lo : c_in;
    p.out_pitch = out_pitch > 0 ? out_pitch : 2 * c_out; p.out_lo = out_lo > 0 ? out_lo : c_out;
    if ((p.in_pitch | p.in_lo | p.out_pitch | p.out_lo) & 7) return HEAL_ERR_UNSUPPORTED;      // 16-byte rows
    cudaStream_t st = (cudaStream_t)stream_;
#define SPT(ci, co) if (c_in == ci && c_out == co) return launch_spt<ci, co>(p, w_packed, st)
    SPT(16, 16); SPT(16, 32); SPT(32, 32); SPT(32, 64); SPT(64, 64); SPT(64, 128);
#undef SPT
    return HEAL_ERR_UNSUPPORTED;
}

extern "C" int heal_rows_to_split(const float* rows_f32, const int* rows_dev, int capacity, int channels, void* out_split_rows, void* stream_) {
    if (!rows_f32 || !out_split_rows || capacity < 1 || channels < 1) return HEAL_ERR_ARG;
    long long total = (long long)capacity * channels;
    k_rows_to_split<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(rows_f32, rows_dev, capacity, channels,
                                                                                       (__nv_bfloat16*)out_split_rows);
    return heal_check_launch();
}

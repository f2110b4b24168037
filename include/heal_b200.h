/* heal_b200 — C ABI of the B200-native (sm_100a) per-frame perception hot path of HEAL / OpenCOOD.
 *
 * The reference (yifanlu0227/HEAL) is pure Python/PyTorch on this path and has NO foreign-function
 * interface of its own; each entry point below therefore cites the reference *Python* call site it
 * replaces (path:line relative to the reference root). INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host
 *    or the comment says "host".
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream, never
 *    synchronises, never allocates (callers pass workspaces; heal_*_workspace() gives the size).
 *  - return value: 0 on success, negative HEAL_ERR_* otherwise (no exceptions cross the ABI).
 *  - activations are channels-last: a (N,C,H,W) map is stored as [n][h][w][c]; `cstride` is the
 *    distance in floats between two pixels and `coffset` the first channel used, so that a conv can
 *    read / write a channel slice of a wider concat buffer.
 */
#ifndef HEAL_B200_H_
#define HEAL_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEAL_B200_ABI_VERSION 2

/* Activation tensor view (channels-last).  fmt: 0 = fp32, 1 = bf16 (one plane), 2 = split-bf16
 * (hi = bf16(x) at data, lo = bf16(x-hi) at data + plane_stride elements; x ~ hi+lo, 16 mantissa bits).
 * cstride = elements between consecutive pixels, coffset = first channel used. */
#define HEAL_FMT_F32 0
#define HEAL_FMT_BF16 1
#define HEAL_FMT_SPLIT 2
typedef struct {
    void* data;
    int fmt;
    int cstride;
    int coffset;
    size_t plane_stride;
} heal_act_t;
int heal_abi_version(void);
/* sm_100a build check: returns 0 when a Blackwell (cc 10.x) device is current, negative otherwise. */
int heal_device_check(void);
/* number of kernels this library has launched in this process so far */
long long heal_launch_count(void);

/* ---- voxelization -------------------------------------------------------------------------
 * replaces SpVoxelPreprocessor.preprocess + collate_batch for all agents of a scene
 * (opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:62-85, :145-174 -> spconv CPU
 * VoxelGeneratorV2.generate / Point2VoxelCPU3d.point_to_voxel). Bit-exact: voxel order = order of
 * first appearance, first `max_points_per_voxel` points in input order, cells beyond `max_voxels`
 * dropped, coords = [agent, z, y, x].
 *   points          (P_total,4) f32, agents concatenated
 *   agent_offsets   (A+1) i32 device: points of agent a are [off[a], off[a+1])
 *   range_min3/voxel_size3/grid3   HOST arrays (x,y,z)
 *   capacity        rows available in the outputs (>= number of voxels produced)
 *   voxels_out      (capacity,T,4) f32   coords_out (capacity,4) i32   num_points_out (capacity) i32
 *   num_voxels_out  (1+A) i32: [0] = M total, [1+a] = voxels of agent a */
size_t heal_voxelize_workspace(int num_points_total, int capacity, int num_agents);
int heal_voxelize(const float* points, const int* agent_offsets, int num_agents, int num_points_total,
                  const float* range_min3_host, const float* voxel_size3_host, const int* grid3_host,
                  int max_points_per_voxel, int max_voxels, int capacity,
                  float* voxels_out, int* coords_out, int* num_points_out, int* num_voxels_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* MeanVFE.forward (opencood/models/sub_modules/mean_vfe.py:13-33): (M,T,4) -> (M,4) */
int heal_mean_vfe(const float* voxels, const int* num_points, int num_voxels, int max_points_per_voxel,
                  float* mean_out, void* stream);

/* ---- PillarVFE + PointPillarScatter, fused --------------------------------------------------
 * replaces PillarVFE.forward (opencood/models/sub_modules/pillar_vfe.py:105-155, PFNLayer :31-53)
 * and PointPillarScatter.forward (opencood/models/sub_modules/point_pillar_scatter.py:19-77).
 *   w_folded (10,64) = linear.weight^T * bn_scale, b_folded (64) = bn_shift (folded on the host, fp64)
 *   offset3_host = voxel_size/2 + range_min (x,y,z)
 *   pillar_features_out (M,64) fp32 or NULL; canvas_out: (B,ny,nx,64) channels-last view in any storage
 *   format, pre-zeroed by the caller, or NULL
 *   num_voxels_dev: optional device count (row 0 used) so a graph-captured frame needs no host sync */
int heal_pillar_vfe_scatter(const float* voxel_features, const int* voxel_num_points, const int* voxel_coords,
                            const int* num_voxels_dev, int num_voxels, int max_points_per_voxel,
                            const float* w_folded, const float* b_folded, int c_in, int c_out,
                            const float* voxel_size3_host, const float* offset3_host, int nx, int ny,
                            float* pillar_features_out, const heal_act_t* canvas_out, void* stream);

/* ---- 2-D convolution, fp32 CUDA-core path ---------------------------------------------------
 * replaces nn.Conv2d / nn.ConvTranspose2d(k==stride) + eval BatchNorm2d + ReLU (+ residual add) of
 * resblock.py:48-64,102-122, base_bev_backbone.py:40-86, base_bev_backbone_resnet.py:54-85,
 * downsample_conv.py:16-27, heter_pyramid_collab.py:102-107.  fp32 FMA math whatever the storage
 * format of in / residual / out (heal_act_t).
 *   weight: groups==1 : fp32 [kh][kw][Cin][w_cstride] (w_cstride >= Cout, multiple of 4), BN scale folded
 *           groups>1  : fp32 [kh*kw][Cin/groups][Cout/groups][groups] (3x3 only, Cin==Cout)
 *   output pixel (oh,ow) of the (Ho,Wo) conv grid is stored at (oh*upsample+up_i, ow*upsample+up_j)
 *   of an (Ho*upsample, Wo*upsample) map: a k==stride transposed conv is upsample^2 1x1 launches. */
int heal_conv2d_simt(const heal_act_t* in, int N, int H, int W, int Cin,
                     const float* weight, int w_cstride, const float* bias,
                     int kh, int kw, int stride, int pad, int groups,
                     const heal_act_t* residual, const heal_act_t* out, int Ho, int Wo, int Cout,
                     int upsample, int up_i, int up_j, int relu, void* stream);

/* ---- 2-D convolution, tcgen05 tensor-core path ---------------------------------------------------
 * Same reference ops as heal_conv2d_nhwc_f32, evaluated as an implicit GEMM with tcgen05.mma (TMEM
 * accumulators) fed by TMA.  Activations are "split-bf16": `planes` bf16 channels-last planes
 * (planes=2: hi = bf16(x), lo = bf16(x-hi), fp32-equivalent via 3 MMAs per K step; planes=1: plain bf16).
 *   in_split   bf16 [planes][N][H][W][in_cstride], plane stride in elements; Cin % 64 == 0
 *   w_packed   bf16 [planes][w_rows][Cin]; conv: w_rows = kh*kw*coutp, row = tap*coutp + co;
 *              transposed conv (kh=kw=1, upsample=k): w_rows = k*k*coutp, row = (i*k+j)*coutp + co;
 *              coutp = Cout padded to 16/32/64 or a multiple of 128; BN scale folded; bias fp32 [Cout]
 *   residual   split planes (res_split) or fp32 (res_f32) or neither, at output resolution
 *   outputs    split planes and/or fp32, channels-last, output map (Ho*upsample, Wo*upsample)
 *   stride     1 or 2 (TMA element strides on W/H)
 *   blockdiag  1: grouped 3x3 conv (Cin == Cout == coutp, channels-per-group | 64) evaluated as block-diagonal
 *              64x64 channel blocks; w_packed is then [planes][kh*kw*coutp][64] (row = tap*coutp + co,
 *              column = input channel within co's 64-channel block, zeros outside co's group) */
int heal_conv2d_tc(const void* in_split, size_t in_plane_stride, int N, int H, int W, int Cin, int in_cstride, int in_coffset,
                   const void* w_packed, int w_rows, int coutp, const float* bias,
                   int kh, int kw, int stride, int pad, int blockdiag, int planes,
                   const void* res_split, size_t res_plane_stride, const float* res_f32, int res_cstride, int res_coffset,
                   void* out_split, size_t out_plane_stride, int out_cstride, int out_coffset,
                   float* out_f32, int out32_cstride, int out32_coffset,
                   int Ho, int Wo, int Cout, int upsample, int relu, void* stream);

/* ---- PyramidFusion weighted fuse of one level -------------------------------------------------
 * replaces weighted_fuse (opencood/models/fuse_modules/pyramid_fuse.py:17-63) incl. both
 * warp_affine_simple calls (torch_transformation_utils.py:323-332), score = sigmoid(occ)+1e-4
 * (:145) and the eval-mode camera crop mask (:147-162) for ONE scene.
 *   feat (n,H,W,C) channels-last; occ (n,H,W) logits; theta (n,2,3) f64 = affine_matrix[b,0,:n]
 *   crop_windows (n,4) i32 [h0,h1,w0,w1] (score kept inside, zeroed outside) or NULL
 *   out (H,W,C) */
int heal_pyramid_fuse_level(const heal_act_t* feat, const float* occ, const double* theta,
                            const int* crop_windows, int n_agents, int H, int W, int C, int align_corners,
                            const heal_act_t* out, void* stream);

/* AttFusion.forward for one scene (opencood/models/fuse_modules/fusion_in_one.py:126-151) */
int heal_att_fuse(const heal_act_t* feat, const double* theta, int n_agents, int H, int W, int C,
                  const heal_act_t* out, void* stream);

/* ---- format conversion between fp32 and (split-)bf16 channels-last buffers ------------------- */
int heal_act_convert(const heal_act_t* src, const heal_act_t* dst, size_t num_pixels, int channels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HEAL_B200_H_ */

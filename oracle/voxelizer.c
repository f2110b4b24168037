/* ORACLE — TEST INFRASTRUCTURE ONLY (never linked into or called by the heal_b200 product path).
 *
 * CPU restatement of the point-cloud voxel generator the reference calls at
 *   opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:65 (spconv 1.x VoxelGeneratorV2.generate)
 *   opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:68 (spconv 2.x Point2VoxelCPU3d.point_to_voxel)
 * The algorithm lives in spconv, a third-party dependency that is NOT vendored, NOT version-pinned
 * (README.md:107-116; no spconv line in requirements.txt) and NOT installed in the build container.
 * This file restates the published algorithm of spconv's `points_to_voxel_3d_np`
 * (spconv v1.2.1, include/spconv/point2voxel.h) from its documented behaviour:
 *   for every point, in order: c_j = floor((p_j - range_min_j) / voxel_size_j) in fp32;
 *   skip the point if any c_j is outside [0, grid_j); look the cell up in a dense
 *   coor_to_voxelidx[z][y][x] table; a new cell gets id = voxel_num++ unless voxel_num >= max_voxels
 *   (then the point is skipped); the point is appended to its voxel while the voxel holds fewer than
 *   max_points points.  coors are stored (z, y, x).
 * PARITY UNPINNED at this boundary: no spconv build and no golden vector exists to check it against
 * (SURVEY.md 8c); it is pinned only by hand-derived known-answer cases in tests/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int oracle_points_to_voxel(const float* points, int num_points, int num_features,
                           const float* voxel_size, const float* coors_range, const int* grid_size,
                           int max_points, int max_voxels,
                           float* voxels /* [max_voxels][max_points][num_features], zeroed here */,
                           int* coors /* [max_voxels][3] zyx */, int* num_points_per_voxel /* [max_voxels] */) {
    const int gx = grid_size[0], gy = grid_size[1], gz = grid_size[2];
    size_t cells = (size_t)gx * gy * gz;
    int* coor_to_voxelidx = (int*)malloc(cells * sizeof(int));
    if (!coor_to_voxelidx) return -1;
    memset(coor_to_voxelidx, 0xFF, cells * sizeof(int)); /* -1 */
    memset(voxels, 0, (size_t)max_voxels * max_points * num_features * sizeof(float));
    memset(num_points_per_voxel, 0, (size_t)max_voxels * sizeof(int));
    int voxel_num = 0;
    for (int i = 0; i < num_points; ++i) {
        int coor[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            float q = (points[(size_t)i * num_features + j] - coors_range[j]) / voxel_size[j];
            float f = floorf(q);
            if (!(f >= 0.0f && f < (float)grid_size[j])) { failed = 1; break; }
            coor[2 - j] = (int)f;
        }
        if (failed) continue;
        size_t lin = ((size_t)coor[0] * gy + coor[1]) * gx + coor[2];
        int voxelidx = coor_to_voxelidx[lin];
        if (voxelidx == -1) {
            voxelidx = voxel_num;
            if (voxel_num >= max_voxels) continue;
            voxel_num += 1;
            coor_to_voxelidx[lin] = voxelidx;
            coors[voxelidx * 3 + 0] = coor[0];
            coors[voxelidx * 3 + 1] = coor[1];
            coors[voxelidx * 3 + 2] = coor[2];
        }
        int num = num_points_per_voxel[voxelidx];
        if (num < max_points) {
            memcpy(voxels + ((size_t)voxelidx * max_points + num) * num_features,
                   points + (size_t)i * num_features, num_features * sizeof(float));
            num_points_per_voxel[voxelidx] = num + 1;
        }
    }
    free(coor_to_voxelidx);
    return voxel_num;
}

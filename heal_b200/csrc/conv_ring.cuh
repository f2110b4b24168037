// Interface between conv2d_tc.cu (argument checks, tensor maps, dispatch) and conv3x3_ring.cu (the row-ring kernel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

struct RingP {
    int N, H, W, C;          // images, rows, columns (multiple of 128), channels (multiple of 64; Cin == Cout, grouped)
    int planes, wplanes;     // activation / weight bf16 planes (1 or 2)
    int relu;                // 0 none, 1 ReLU
    int segs, seg_rows;      // row segments per (image, column tile, channel block) strip and rows per segment
    int pdl;                 // launched with programmatic stream serialization (griddepcontrol.wait before the first global read)
    const float* bias;       // [C] or null
};

// tmA: activations {C, W, H, N, plane}, box {64, 130, 1, 1, planes}, 128B swizzle (the halo map of heal_conv2d_tc)
// tmB: packed diagonal weight sub-blocks, box {16, 16, wplanes, 4, 3}, 32B swizzle
// tmO: output {C, W, H, N, plane}, box {64, 128, 1, 1, planes}, 128B swizzle
int heal_conv3x3_ring_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, RingP p, cudaStream_t st);

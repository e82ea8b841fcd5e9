// Device-side interface of the encode kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1b_types.h"

namespace av1b {

// Per-launch parameters of the closed-loop intra encode kernel (one CTA per tile per frame).
struct IntraLaunch {
  Av1bGeom g;
  int32_t bit_depth;
  int32_t base_q_idx;
  int32_t quant_rnd;
  int32_t dc_q, ac_q;
  // frame-batch layout: plane p of frame f starts at base[p] + f * frame_stride[p] (in elements)
  const uint16_t* src[3];
  uint16_t* rec[3];
  int16_t* coef[3];
  size_t plane_elems[3];
  Av1bBlockInfo* blocks;      // [n_frames][h8*w8]
  const uint8_t* part_map;    // [n_frames][h8*w8]   (per-frame partition quadtree)
  size_t map_elems;
};

void upload_tables_once();
cudaError_t launch_partition_fixed(const Av1bGeom& g, int blk_log2, uint8_t* map, int n_frames, cudaStream_t s);
cudaError_t launch_intra_encode(const IntraLaunch& p, int n_frames, cudaStream_t s);

}  // namespace av1b

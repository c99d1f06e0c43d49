// Host-callable launchers of the per-field-count kernels.  kernels_f.cu is compiled once per field
// count F (-DPB_F=1..4, four translation units built in parallel) and exports one FieldOps table;
// engine.cu picks the table of the index's F.  Every entry returns the cudaError_t of the launch.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "union_kernels.cuh"
#include "union_warp_kernel.cuh"

namespace pbk {

struct FieldOps {
  cudaError_t (*score_occupancy)(int scorer, bool gmode, bool narrow, int* per_sm, int threads, size_t smem);
  cudaError_t (*score_launch)(int scorer, bool gmode, bool narrow, const ScoreParams* P, int grid, int threads,
                              size_t smem, cudaStream_t st);
  cudaError_t (*mark_launch)(const ScoreParams* P, int clear, int grid, cudaStream_t st);
  cudaError_t (*fold_launch)(int scorer, const FoldParams* FP, int grid, cudaStream_t st);
  cudaError_t (*binfold_launch)(int scorer, const ScoreParams* P, int grid, cudaStream_t st);
  cudaError_t (*live_df_launch)(const IndexView* ix, unsigned long long* df_live, uint32_t* live_rows, int grid,
                                cudaStream_t st);
  cudaError_t (*union_occupancy)(bool gen, int* per_sm, size_t smem);
  cudaError_t (*union_launch)(bool gen, const UnionParams* P, int grid, size_t smem, cudaStream_t st);
  cudaError_t (*union_warp_occupancy)(bool gen, int* per_sm, size_t smem);
  cudaError_t (*union_warp_launch)(bool gen, const UnionParams* P, int grid, size_t smem, cudaStream_t st);
};

const FieldOps* field_ops_f1();
const FieldOps* field_ops_f2();
const FieldOps* field_ops_f3();
const FieldOps* field_ops_f4();

inline const FieldOps* field_ops(uint32_t F) {
  switch (F) {
    case 1: return field_ops_f1();
    case 2: return field_ops_f2();
    case 3: return field_ops_f3();
    case 4: return field_ops_f4();
    default: return nullptr;
  }
}

}  // namespace pbk

// The per-field-count kernels (kernels.cuh) for ONE field count: compiled four times with
// -DPB_F=1..4 so that the 8 instantiations of the scoring kernel per F build in parallel.
#include <map>
#include <mutex>
#include <utility>

#include "field_ops.hpp"

#ifndef PB_F
#error "compile with -DPB_F=1..4"
#endif

namespace pbk {
namespace {

constexpr int F = PB_F;

// The dynamic shared-memory limit of a kernel is per device and only ever RAISED here: indexes with different table
// sizes (a main image and its delta segment, say) launch the same instantiation, and the callers cache the occupancy.
template <class K>
cudaError_t raise_smem_limit(K kernel, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> limit;      // (kernel, device) -> limit set so far
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  int& cur = limit[{reinterpret_cast<const void*>(kernel), dev}];
  if ((int)smem <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) cur = (int)smem;
  return e;
}

template <int SC, bool G, bool N>
cudaError_t occ_t(int* per_sm, int threads, size_t smem) {
  cudaError_t e = raise_smem_limit(score_kernel<F, SC, G, N>, smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, score_kernel<F, SC, G, N>, threads, smem);
}
template <int SC, bool G, bool N>
cudaError_t launch_t(const ScoreParams* P, int grid, int threads, size_t smem, cudaStream_t st) {
  score_kernel<F, SC, G, N><<<grid, threads, smem, st>>>(*P);
  return cudaGetLastError();
}

// run fn<SC, G, N> for the runtime (scorer, gmode, narrow) triple
#define PB_DISPATCH3(fn, ...)                                                       \
  switch ((scorer ? 4 : 0) | (gmode ? 2 : 0) | (narrow ? 1 : 0)) {                   \
    case 0: return fn<0, false, false>(__VA_ARGS__);                                \
    case 1: return fn<0, false, true>(__VA_ARGS__);                                 \
    case 2: return fn<0, true, false>(__VA_ARGS__);                                 \
    case 3: return fn<0, true, true>(__VA_ARGS__);                                  \
    case 4: return fn<1, false, false>(__VA_ARGS__);                                \
    case 5: return fn<1, false, true>(__VA_ARGS__);                                 \
    case 6: return fn<1, true, false>(__VA_ARGS__);                                 \
    default: return fn<1, true, true>(__VA_ARGS__);                                 \
  }

cudaError_t score_occupancy(int scorer, bool gmode, bool narrow, int* per_sm, int threads, size_t smem) {
  PB_DISPATCH3(occ_t, per_sm, threads, smem)
}
cudaError_t score_launch(int scorer, bool gmode, bool narrow, const ScoreParams* P, int grid, int threads, size_t smem,
                         cudaStream_t st) {
  PB_DISPATCH3(launch_t, P, grid, threads, smem, st)
}
cudaError_t mark_launch(const ScoreParams* P, int clear, int grid, cudaStream_t st) {
  mark_kernel<F><<<grid, CTA_THREADS, 0, st>>>(*P, clear);
  return cudaGetLastError();
}
cudaError_t fold_launch(int scorer, const FoldParams* FP, int grid, cudaStream_t st) {
  if (scorer) fold_kernel<F, 1><<<grid, CTA_THREADS, 0, st>>>(*FP);
  else fold_kernel<F, 0><<<grid, CTA_THREADS, 0, st>>>(*FP);
  return cudaGetLastError();
}
cudaError_t binfold_launch(int scorer, const ScoreParams* P, int grid, cudaStream_t st) {
  if (scorer) binfold_kernel<F, 1><<<grid, CTA_THREADS, 0, st>>>(*P);
  else binfold_kernel<F, 0><<<grid, CTA_THREADS, 0, st>>>(*P);
  return cudaGetLastError();
}
cudaError_t live_df_launch(const IndexView* ix, unsigned long long* df_live, uint32_t* live_rows, int grid, cudaStream_t st) {
  live_df_kernel<F><<<grid, 256, 0, st>>>(*ix, df_live, live_rows);
  return cudaGetLastError();
}

template <bool GEN>
cudaError_t union_occ_t(int* per_sm, size_t smem) {
  cudaError_t e = raise_smem_limit(union_kernel<F, GEN>, smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, union_kernel<F, GEN>, UShape<GEN>::THREADS, smem);
}
cudaError_t union_occupancy(bool gen, int* per_sm, size_t smem) {
  return gen ? union_occ_t<true>(per_sm, smem) : union_occ_t<false>(per_sm, smem);
}
cudaError_t union_launch(bool gen, const UnionParams* P, int grid, size_t smem, cudaStream_t st) {
  if (gen) union_kernel<F, true><<<grid, UShape<true>::THREADS, smem, st>>>(*P);
  else union_kernel<F, false><<<grid, UShape<false>::THREADS, smem, st>>>(*P);
  return cudaGetLastError();
}

template <bool GEN>
cudaError_t union_warp_occ_t(int* per_sm, size_t smem) {
  cudaError_t e = raise_smem_limit(union_warp_kernel<F, GEN>, smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, union_warp_kernel<F, GEN>, U_W_THREADS, smem);
}
cudaError_t union_warp_occupancy(bool gen, int* per_sm, size_t smem) {
  return gen ? union_warp_occ_t<true>(per_sm, smem) : union_warp_occ_t<false>(per_sm, smem);
}
cudaError_t union_warp_launch(bool gen, const UnionParams* P, int grid, size_t smem, cudaStream_t st) {
  if (gen) union_warp_kernel<F, true><<<grid, U_W_THREADS, smem, st>>>(*P);
  else union_warp_kernel<F, false><<<grid, U_W_THREADS, smem, st>>>(*P);
  return cudaGetLastError();
}

const FieldOps OPS = {score_occupancy, score_launch, mark_launch, fold_launch, binfold_launch, live_df_launch,
                      union_occupancy, union_launch, union_warp_occupancy, union_warp_launch};

}  // namespace

#define PB_CAT2(a, b) a##b
#define PB_CAT(a, b) PB_CAT2(a, b)
const FieldOps* PB_CAT(field_ops_f, PB_F)() { return &OPS; }

}  // namespace pbk

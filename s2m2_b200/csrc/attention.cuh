// Shared parameter block of the attention kernels (attention.cu: CUDA-core, attention_mma.cu: tensor-core).
#pragma once
#include "common.cuh"
namespace s2m2 {
struct AttnP {
  const void* qkv;
  void* out;
  void* pesum;
  const float *pe_x, *pe_y;
  long long qkv_ld, out_ld;
  int qh, oh, ph_half;
  int S, N, heads, d, C, kv_shift, pe_h, pe_w;
  float scale;
};
int attention_simt_launch(const AttnP& p, cudaStream_t s);
int attention_mma_supported(const AttnP& p);
int attention_mma_launch(const AttnP& p, cudaStream_t s);
int attention_tc_supported(const AttnP& p);
size_t attention_tc_workspace_bytes(const AttnP& p);
int attention_tc_launch(const AttnP& p, void* workspace, cudaStream_t s);
// tcgen05 2-D attention (N >= 256 keys, head dim 64), optionally with the relative-PE marginals (work = indicator matrix)
int attention2d_tc_supported(const AttnP& p, const void* work, long long work_bytes);
int attention2d_tc_launch(const AttnP& p, const void* indicator, cudaStream_t s);
size_t attention2d_tc_indicator_bytes(int h, int w);
}  // namespace s2m2

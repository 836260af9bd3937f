/* A host that is not Python: plain C over the handle-level C-ABI of libs2m2b200.so (include/s2m2b200.h).
 *
 *   host_forward <weights.bin> <feature_channels> <num_transformer> <B> <H> <W> <pair.bin> <out.bin> [repeat]
 *
 * weights.bin : a state dict written by tools/export_state_dict.py ("S2M2SD01", int32 count, then per entry
 *               int32 name_len, name, int32 ndim, int64 shape[ndim], float32 data) — the reference's keys (s2m2.py:69-78)
 * pair.bin    : left then right image, each (B,3,H,W) float32 in [0,255] (what image_pad hands to the model, model_utils.py:72-76)
 * out.bin     : 3 x (B,1,H,W) float32: disparity | occlusion | confidence (s2m2.py:197)
 *
 * The calls are exactly what a cgo / JNI / Rust-FFI binding of the library would make; cudart appears only for the device
 * buffers and the stream.  tests/test_host_forward_gpu.py runs this program and compares out.bin with the nn.Module bit for bit.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "s2m2b200.h"

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } while (0)
#define CK(call) do { if ((call) != 0) DIE("%s failed: %s", #call, s2m2b200_last_error()); } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) DIE("%s: %s", #call, cudaGetErrorString(e_)); } while (0)

static void read_exact(void* dst, size_t n, FILE* f, const char* what) {
  if (fread(dst, 1, n, f) != n) DIE("short read (%s)", what);
}

int main(int argc, char** argv) {
  if (argc < 9) DIE("usage: %s weights.bin C NTR B H W pair.bin out.bin [repeat]", argv[0]);
  const int C = atoi(argv[2]), NTR = atoi(argv[3]), B = atoi(argv[4]), H = atoi(argv[5]), W = atoi(argv[6]);
  const int repeat = argc > 9 ? atoi(argv[9]) : 1;

  S2m2Model* model = NULL;
  CK(s2m2b200_create(C, NTR, /*use_positivity=*/1, /*output_upsample=*/0, /*refine_iter=*/3, S2M2_F16, &model));

  FILE* f = fopen(argv[1], "rb");
  if (!f) DIE("cannot open %s", argv[1]);
  char magic[8];
  int32_t count;
  read_exact(magic, 8, f, "magic");
  if (memcmp(magic, "S2M2SD01", 8) != 0) DIE("%s is not a state-dict file", argv[1]);
  read_exact(&count, 4, f, "count");
  for (int i = 0; i < count; ++i) {
    int32_t nlen, ndim;
    char name[512];
    int64_t shape[4], numel = 1;
    read_exact(&nlen, 4, f, "name length");
    if (nlen <= 0 || nlen >= (int32_t)sizeof(name)) DIE("bad name length");
    read_exact(name, (size_t)nlen, f, "name");
    name[nlen] = 0;
    read_exact(&ndim, 4, f, "ndim");
    if (ndim < 1 || ndim > 4) DIE("%s: ndim %d", name, ndim);
    read_exact(shape, 8 * (size_t)ndim, f, "shape");
    for (int d = 0; d < ndim; ++d) numel *= shape[d];
    float* data = (float*)malloc((size_t)numel * 4);
    read_exact(data, (size_t)numel * 4, f, name);
    CK(s2m2b200_set_param(model, name, data, shape, ndim));
    free(data);
  }
  fclose(f);
  CK(s2m2b200_finalize_params(model));

  const size_t img = (size_t)B * 3 * H * W, outn = (size_t)3 * B * H * W;
  float* hpair = (float*)malloc(2 * img * 4);
  f = fopen(argv[7], "rb");
  if (!f) DIE("cannot open %s", argv[7]);
  read_exact(hpair, 2 * img * 4, f, "images");
  fclose(f);

  const int64_t wbytes = s2m2b200_workspace_bytes(model, B, H, W, 0, 0, 0);
  if (wbytes < 0) DIE("workspace_bytes: %s", s2m2b200_last_error());
  float *dpair, *dout;
  void* work;
  cudaStream_t stream;
  CU(cudaStreamCreate(&stream));
  CU(cudaMalloc((void**)&dpair, 2 * img * 4));
  CU(cudaMalloc((void**)&dout, outn * 4));
  CU(cudaMalloc(&work, (size_t)wbytes));
  CU(cudaMemcpyAsync(dpair, hpair, 2 * img * 4, cudaMemcpyHostToDevice, stream));

  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CK(s2m2b200_forward(model, dpair, dpair + img, 0, B, H, W, 0, 0, dout, work, wbytes, stream));   /* warm-up */
  CU(cudaEventRecord(e0, stream));
  for (int r = 0; r < repeat; ++r) CK(s2m2b200_forward(model, dpair, dpair + img, 0, B, H, W, 0, 0, dout, work, wbytes, stream));
  CU(cudaEventRecord(e1, stream));
  CU(cudaStreamSynchronize(stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));

  float* hout = (float*)malloc(outn * 4);
  CU(cudaMemcpy(hout, dout, outn * 4, cudaMemcpyDeviceToHost));
  f = fopen(argv[8], "wb");
  if (!f || fwrite(hout, 4, outn, f) != outn) DIE("cannot write %s", argv[8]);
  fclose(f);
  double mean = 0.0;
  for (size_t i = 0; i < (size_t)B * H * W; ++i) mean += hout[i];
  printf("%s | C=%d NTR=%d %dx%dx%d | workspace %.1f MB | %.3f ms / forward over %d | mean disparity %.6f | %lld kernel launches\n",
         s2m2b200_version(), C, NTR, B, H, W, wbytes / 1e6, ms / repeat, repeat, mean / ((double)B * H * W), (long long)s2m2b200_launch_count());

  s2m2b200_destroy(model);
  cudaFree(dpair); cudaFree(dout); cudaFree(work);
  free(hpair); free(hout);
  return 0;
}

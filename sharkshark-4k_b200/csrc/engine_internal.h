// Internals shared by engine.cu (contexts, plans, C ABI) and bsvd.cu (streaming denoiser).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/ss4k.h"
#include "conv_params.h"
#include "elementwise.h"
#include "program.h"

namespace ss4k {

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};
typedef std::map<std::string, HostTensor> WeightMap;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

std::string fmt(const char* f, ...);
int round_up(int a, int b);

// packed weights of one conv on the device + its MMA schedule tables (geometry independent)
struct ConvWeights {
  void* d_w = nullptr;
  float* d_bias = nullptr;
  float* d_slope = nullptr;
  std::vector<KBlock> kb;
  std::vector<Tap> taps;
  std::vector<uint8_t> mask;  // [nkb][kMaxTaps]
  int nkb = 0, ntaps = 0, nsub = 1, max_dr = 2, npad_total = 0;
  bool bf16 = false;
};

// a launchable conv: parameter block + grid
struct ConvExec {
  ConvParams p;
  int grid = 0;
  bool ext_out = false;  // ep.out is the caller's output pointer (patched per run)
};

struct BufView {
  void* ptr = nullptr;
  int64_t lo_off = 0;  // element offset of the low half (0: plain tensor)
};

}  // namespace ss4k

struct ss4k_ctx {
  int device = 0;
  int nsm = 148;
  int desc_mode = 0;
  std::string err;
  std::map<int, ss4k::WeightMap> weights;
  ss4k::EncodeTiledFn encode = nullptr;
  int32_t* err_host = nullptr;  // mapped pinned int[4]: watchdog diagnostics
  int32_t* err_dev = nullptr;
  int64_t launches = 0;
  cudaStream_t stream = nullptr;  // internal stream (ss4k_run_host, graph capture)
};

namespace ss4k {

int fail(ss4k_ctx* ctx, int code, const std::string& msg);
int check_kernel_health(ss4k_ctx* ctx, cudaError_t e, const char* what);

#define SS4K_CK(ctx, call)                                                                                  \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess)                                                                                  \
      return ::ss4k::fail(ctx, SS4K_E_CUDA,                                                                 \
                          ::ss4k::fmt("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
  } while (0)

// look up weight / bias / slope tensors of a conv in a weight map
int find_conv_tensors(ss4k_ctx* ctx, const WeightMap& wmap, const ConvSpec& cs, const HostTensor** W,
                      const HostTensor** B, const HostTensor** S);
int prepare_conv_weights(ss4k_ctx* ctx, const ConvSpec& cs, const HostTensor& W, const HostTensor* B,
                         const HostTensor* S, int act_mode, ConvWeights* cw);
void free_conv_weights(ConvWeights& cw);
int build_conv_params(ss4k_ctx* ctx, const ConvSpec& cs, const ConvWeights& cw,
                      const std::function<BufView(int)>& bufview, ConvExec* ex);
int launch_conv(ss4k_ctx* ctx, const ConvExec& ex, void* ext_out, cudaStream_t st);

// BSVD (bsvd.cu)
struct BsvdEngine;
int bsvd_create(ss4k_ctx* ctx, const ss4k_plan_cfg& cfg, const WeightMap& wmap, BsvdEngine** out);
void bsvd_destroy(BsvdEngine* e);
int bsvd_run_clip(BsvdEngine* e, const void* in_dev, void* out_dev, int nframes, cudaStream_t st);
double bsvd_flops_per_frame(const BsvdEngine* e);
int bsvd_launches_per_frame(const BsvdEngine* e);

}  // namespace ss4k

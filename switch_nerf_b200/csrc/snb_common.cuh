// Shared helpers for the switch_nerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/switch_nerf_b200.h"

namespace snb {

void set_error(const char* fmt, ...);
void count_launch();
struct PhaseEvents { cudaEvent_t e[6]; };   // front [0,1] | route [2,3] | back [4,5]
// returns nullptr when profiling is off (or the event pool is exhausted)
PhaseEvents* profile_next();

#define SNB_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      snb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SNB_ECUDA;                                                             \
    }                                                                               \
  } while (0)

#define SNB_CHECK_LAUNCH(name)                                                      \
  do {                                                                              \
    snb::count_launch();                                                            \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      snb::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SNB_ECUDA;                                                             \
    }                                                                               \
  } while (0)

#define SNB_REQUIRE(cond, ...)                                                      \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      snb::set_error(__VA_ARGS__);                                                  \
      return SNB_EINVAL;                                                            \
    }                                                                               \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over the caller-provided workspace.
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  bool ok;
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), ok(true) {}
  template <typename T>
  T* take(size_t n) {
    size_t bytes = align_up(n * sizeof(T), 256);
    if (off + bytes > cap) { ok = false; return nullptr; }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

// Activation codes for the fp32 linear kernel epilogue.
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_SOFTPLUS_SHIFT = 3 };

struct Ep;   // expert-parallel group (snb_ep.cuh)

// Device-side model storage (library-owned copies, see snb_model_create).
struct Model {
  snb_model_desc d;
  snb_tuning tune;   // per-model knobs (defaults <- SNB_* environment at creation; snb_model_set_tuning)
  int xyz_in;    // 3 + 6*pos_xyz_freqs
  int dir_in;    // 3 + 6*pos_dir_freqs
  int cat_in;    // width + dir_in + appearance_dim
  int x_cols;    // 7 or 10
  // fp32 copies, every GEMM weight as [N(out), K(in)] row-major
  float* f32_blob = nullptr;
  size_t f32_bytes = 0;
  float *xyz_w, *xyz_b, *gate_w[4], *gate_b[4], *ln_w, *ln_b, *wg;
  float *exp_w[16];  // [E][N][K]  (transposed from the reference's [E][K][N])
  float *exp_b[16];  // [E][N]
  float *l1_w, *l1_b, *l2_w, *l2_b, *sigma_w, *sigma_b, *color_w, *color_b, *emb_a;
  // bf16 packed blob for the tcgen05 path (layout documented in snb_tc.cu)
  void* tc_blob = nullptr;
  size_t tc_bytes = 0;
  int sm_count = 148;
  // software pipelining of render_rays: routing kernels run on this stream
  cudaStream_t side_stream = nullptr;
  int* sel_zero = nullptr;               // [4 sets][SEL_ZERO_INTS] scratch of k_select that is zero between uses (self-cleaning)
  cudaStream_t fin_stream = nullptr;     // expert-parallel: waits for the peers' result rows off the main stream
  cudaEvent_t ev_back[4] = {nullptr, nullptr, nullptr, nullptr}, ev_fin[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_front[4] = {nullptr, nullptr, nullptr, nullptr}, ev_route[4] = {nullptr, nullptr, nullptr, nullptr};
  // expert-parallel group attached by snb_model_attach_a2a (not owned); nullptr = every expert is local
  Ep* ep = nullptr;
};

// forward intermediates of one model chunk kept for the backward sweep (snb_fp32.cu fills it, snb_backward.cu reads it)
struct Fp32Saved {
  float *pe, *h, *ga[5], *g, *gates, *gate, *bufx, *act[16], *out_rows, *hr, *sig_pre, *cat, *h2, *rgb;
  int *idx, *loc, *counts, *cap_dev, *ebase, *erows;
  int64_t rows, cap_host;
};
int fp32_forward_saved(Model* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* o,
                       Fp32Saved* sv, Arena& ws, cudaStream_t st);
int ln_forward_launch(const float* g, int64_t S, int M, const float* w, const float* b, float* out, cudaStream_t st);
int fp32_backward(Model* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* o, const float* d_out,
                  const float* d_l_aux, const snb_grads* G, Arena& ws, cudaStream_t st);
size_t fp32_backward_workspace_bytes(const Model* m, int64_t S, double max_cf);
int composite_backward_launch(const float* z, const float* raw, const float* last_delta, int64_t N, int S, const float* d_rgb,
                              float* d_raw, cudaStream_t st);

// ---- entry points implemented in the individual .cu files ----
int fp32_forward(Model* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* o,
                 float* out, int32_t* moe_idx, float* l_aux, float* dbg_gates, int32_t* dbg_loc,
                 Arena& ws, cudaStream_t st);
size_t fp32_workspace_bytes(const Model* m, int64_t S, double max_cf);
int fp32_moe_layer(Model* m, const float* input, const float* gate_input, int64_t S, const snb_route_opts* o, float* y,
                   int32_t* moe_idx, float* l_aux, Arena& ws, cudaStream_t st);
size_t fp32_moe_layer_workspace_bytes(const Model* m, int64_t S, double max_cf);

int tc_pack_weights(Model* m, const snb_weights* w, cudaStream_t st);
int tc_forward(Model* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* o,
               float* out, int32_t* moe_idx, float* l_aux, float* dbg_gates, int32_t* dbg_loc,
               Arena& ws, cudaStream_t st);
size_t tc_workspace_bytes(const Model* m, int64_t S, double max_cf);
// rows of a render pass given as rays + depths instead of the materialised [S,7] tensor (x may then be NULL)
struct RaySource {
  const float* rays;   // [N,8]
  const float* z;      // [N*Sn] ray-major
  const int* img;      // [N] nullable
  int Sn;
};
int tc_forward_chunks(Model* m, const float* x, int64_t B, int64_t chunk, const snb_route_opts* o, float* out,
                      int32_t* moe_idx, float* l_aux, void* ws_base, size_t ws_stride, int nsets, cudaStream_t st,
                      const float* sigma_noise = nullptr, const RaySource* rs = nullptr);
// true when the kernels that will run for this model can take a RaySource (the TS kernels, all experts local)
bool tc_ray_source_ok(const Model* m);
bool tc_supported(const Model* m);
void tuning_from_env(snb_tuning* t);
void tc_release(Model* m);

size_t route_workspace_bytes(int64_t S, int32_t E);
// Routing on device: see snb_route_top1 in the public header.  `ebase/erows` (nullable, [E]):
// first dispatch row and number of computed rows of each expert for (no_batch ? contiguous : padded) layout.
int route_top1(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int32_t* idx, int32_t* loc,
               float* gate, int32_t* counts, int32_t* capacity, float* l_aux, void* ws, size_t ws_bytes,
               cudaStream_t st);

__host__ __device__ inline int capacity_of(int64_t S, int E, double cf) {
  // tutel_fast_dispatch.py:210-211: top_k * int(cf * ceil(S/E)); Python float math is double.
  return (int)(cf * (double)((S + E - 1) / E));
}

}  // namespace snb

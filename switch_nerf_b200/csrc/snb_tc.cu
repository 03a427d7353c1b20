// SNB_PREC_BF16 path: the NeRFMoE forward of one model_chunk as two fused tcgen05 kernels
// (+ the small routing kernels in between):
//
//   k_front  (launch #1)  per 128-sample tile: positional encoding -> xyz Linear -> external gate
//                         MLP -> LayerNorm -> fp32 gate GEMM -> softmax.  Writes h (bf16 [S,M]) and
//                         gates (fp32 [S,E]).  reference: models/nerf_moe.py:330-372,
//                         tutel_moe_layer_nobatch.py:105-126
//   route_top1            snb_route.cu (idx / loc / counts / l_aux on device)
//   k_back   (launch #2)  per (expert, 128-row tile of that expert's bucket): gather rows by the
//                         dispatch index -> 7 expert layers (skip at 3) -> x gate (combine) -> ReLU
//                         -> sigma head -> layer "1" -> [dir PE | appearance] concat -> layer "2" ->
//                         colour head -> sigmoid / softplus -> scatter [rgb, sigma] to out[sample].
//                         Dropped samples form one more bucket that skips the expert stack (h = 0).
//                         reference: tutel_moe_layer_nobatch.py:142-225, 887-924; nerf_moe.py:384-441
//
// Inside a CTA (576 threads = 18 warps, 1 CTA / SM, persistent over tiles):
//   warp 0     weight producer : 1-D bulk async copies (UBLKCP) of pre-packed K-slices (<=64 wide)
//                                into a 3-stage shared-memory ring, mbarrier full/empty
//   warp 1     MMA issuer      : one thread issues tcgen05.mma (M=128, N<=256, K=16 per instr.),
//                                accumulators in TMEM, ping-pong between columns [0,256) / [256,512)
//   warps 2-17 epilogue        : 16 warps = 4 TMEM lane quarters x 4 column sub-slices.  tcgen05.ld ->
//                                bias / ReLU / skip / gate / heads in fp32 registers -> bf16 back into
//                                the shared A tile, one 64-column chunk at a time (each warp owns 16 of
//                                the 64 columns), signalling a per-chunk mbarrier so the next layer's
//                                MMAs start while the rest of the epilogue is still running.  Row-wise
//                                reductions (LayerNorm statistics, sigma / colour dot products) are
//                                finished through a small shared-memory exchange.
// The fp32 gate GEMM (fp32_gate: True) runs on the tensor cores too: LayerNorm is folded into the
// weights, logits = rstd * (g . (gamma*wg)^T - mean * c1) + c0, g is exactly representable in bf16
// (it is the bf16 output of the gate MLP) and gamma*wg is split into bf16 hi + lo parts (N = 2x16),
// which reproduces the fp32 product to ~2^-17 relative.
// Activations never leave the SM between layers; per sample HBM traffic is x (28 B) + h (2x512 B)
// + gates (32 B) + out (16 B).
//
// Operand layout (both A and B, "K-major, no swizzle" UMMA canonical form): 8x8 bf16 core matrices of
// 128 contiguous bytes; LBO = distance between K-adjacent core matrices (128 B), SBO = distance
// between 8-row groups.  Weights are packed into exactly that image on the host side of the C ABI
// (tc_pack_weights) so a K-slice is ONE contiguous bulk copy.
#include "snb_common.cuh"
#include <stdlib.h>
#include <atomic>
#include <mutex>

#include "snb_umma.cuh"
#include "snb_ep.cuh"
#include "snb_select.cuh"

namespace snb {
using namespace ptx;
static_assert(SEL_MAX_E == 16, "k_front_ts keeps one softmax register per possible expert");
extern unsigned long long* g_timeline;

static constexpr int TILE = 128;            // rows per tile (UMMA M)
static constexpr int MW = 256;              // model width handled by this path
static constexpr int KA_MAX = 352;          // widest A operand (layer "2": 256 + 27 + 48 = 331 -> 336), padded
static constexpr uint32_t SBO_A = KA_MAX / 8 * 128;   // 5632 B between 8-row groups of the A tile
static constexpr uint32_t A_BYTES = TILE / 8 * SBO_A; // 90112
static constexpr int NSTAGE = 3;
static constexpr uint32_t STAGE_BYTES = 256 * 64 * 2; // one K-slice of a 256-wide layer
static constexpr int NCHUNK = 6;            // 64-column chunks of the A tile (352/64 rounded up)
static constexpr int EPI_WARPS = 16;
static constexpr int EPI_THREADS = EPI_WARPS * 32;   // 512
static constexpr int THREADS = 64 + EPI_THREADS;      // 576
static constexpr int GATE_N = 32;                     // gate GEMM: 16 "hi" + 16 "lo" output columns
static constexpr int MAX_E = 16;

#ifndef SNB_UMMA_SWAP
#define SNB_UMMA_SWAP 0   // set to 1 if the hardware interprets LBO/SBO the other way round (selftest variant 1)
#endif
__device__ __forceinline__ uint64_t op_desc(uint32_t addr, uint32_t k_stride, uint32_t mn_stride) {
#if SNB_UMMA_SWAP
  return umma_smem_desc(addr, mn_stride, k_stride);
#else
  return umma_smem_desc(addr, k_stride, mn_stride);
#endif
}

// ------------------------------------------------------------------------------------------
// packed weights
// ------------------------------------------------------------------------------------------
struct TcLayer {
  uint32_t w_off_w; // byte offset of the wide image (snb_tc_wide.cuh) inside wblob_w
  uint32_t w_off;   // byte offset of the packed bf16 image inside tc_blob
  uint32_t K16;     // K rounded up to 16
  uint32_t N;       // output features (<= 256)
  uint32_t b_off;   // float offset of the bias inside the fp32 side table
};

struct TcParams {
  const uint8_t* wblob_w;   // packed bf16 weights, wide images (per layer half and 32-wide K-slice); nullptr unless used
  uint32_t expert_w_stride_w;
  int mip;                  // MipNeRFMoE: x = [mean3, cov3, dir3, idx], integrated positional encoding
  const uint8_t* wblob;     // packed bf16 weights
  const float* fblob;       // fp32 side table: biases (bf16-rounded), gate constants, sigma/colour heads
  TcLayer front[5];         // xyz, gate fcs
  int n_front;
  TcLayer gate;             // folded LayerNorm+gate GEMM: rows [0,16) = hi(gamma*wg), rows [16,32) = lo
  TcLayer expert[16];       // layer j of expert 0; expert e adds e*expert_stride(_b)
  int n_expert;
  uint32_t expert_w_stride; // bytes between experts (all layers of one expert are contiguous)
  uint32_t expert_b_stride; // floats between experts
  TcLayer back[2];          // layer "1", layer "2"
  uint32_t o_c0, o_c1, o_wsig, o_bsig, o_wcol, o_bcol;   // float offsets in fblob
  uint32_t o_b3x;           // [E][256]: bias of the skip layer + bias of the xyz layer (h is re-accumulated by the MMA)
  int recompute_h;          // 1: launch #2 recomputes h = xyz Linear(PE) per tile instead of gathering it from HBM
  int E, skip_layer, pos_xyz_freqs, pos_dir_freqs, appearance_dim, appearance_count, hidden2, x_cols;
  const float* emb_a;       // fp32 [count, A]
  const __nv_bfloat16* emb_cat;   // bf16 [count, cat_cols]: the appearance row where launch #2's cat block wants it (columns
                                  // [NDIR, NDIR + A), zeros elsewhere) -- 16-byte chunks copied as they are (TS kernels)
  int32_t cat_cols;
  unsigned long long* tl;   // debug timeline (nullable): [role][TL_N] (tag<<48 | clock) marks of CTA 0
  int ab;                   // debug A/B switches (SNB_FRONT_AB): 1 = no level-0 histogram, 2 = no column sums
  // ray source (render passes on the TS kernels): the [S,7] rows [o + d z, d, image index] of rendering.py:357-362 are
  // never materialised -- a row is rebuilt from its 32-byte ray and its 4-byte depth where it is consumed
  const float* ray_src;     // [N,8] rays (nullptr: rows come from x)
  const float* ray_z;       // [N * ray_sn] depths of the pass, ray-major
  const int* ray_img;       // [N] image indices (nullable)
  int ray_sn;               // samples per ray of the pass
  int64_t ray_s0;           // first sample of this chunk inside the pass
};

// row `s` of the chunk from the ray source: xyz = o + d * z with separate multiply and add (torch: rays_o + rays_d * z)
__device__ __forceinline__ void ray_row_xyz(const TcParams& P, int64_t s, float& x0, float& x1, float& x2) {
  const uint32_t g = (uint32_t)(P.ray_s0 + s);
  const float* ray = P.ray_src + (size_t)(g / (uint32_t)P.ray_sn) * 8;
  const float zv = P.ray_z[g];
  x0 = __fadd_rn(ray[0], __fmul_rn(ray[3], zv));
  x1 = __fadd_rn(ray[1], __fmul_rn(ray[4], zv));
  x2 = __fadd_rn(ray[2], __fmul_rn(ray[5], zv));
}
__device__ __forceinline__ void ray_row_dir(const TcParams& P, int64_t s, float& d0, float& d1, float& d2, int& img) {
  const uint32_t r = (uint32_t)(P.ray_s0 + s) / (uint32_t)P.ray_sn;
  const float* ray = P.ray_src + (size_t)r * 8;
  d0 = ray[3]; d1 = ray[4]; d2 = ray[5];
  img = P.ray_img ? P.ray_img[r] : 0;
}

// canonical image: slices of 64 k; inside a slice (n/8)*(klen*16) + (kk/8)*128 + (n%8)*16 + (kk%8)*2 bytes
__device__ __forceinline__ size_t packed_index(int n, int k, int N, int K16) {
  const int j = k / 64, kk = k % 64;
  const int klen = min(64, K16 - 64 * j);
  return (size_t)N * 64 * j + ((size_t)(n / 8) * (klen * 16) + (size_t)(kk / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2) / 2;
}

// embedding_a rows as launch #2's cat block holds them: [0 x ndir | bf16(row) | 0-pad] (cols columns per row)
__global__ void k_pack_emb_cat(const float* __restrict__ emb, int count, int A, int ndir, int cols, __nv_bfloat16* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)count * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols) - ndir;
  dst[i] = __float2bfloat16_rn((c >= 0 && c < A) ? emb[(int64_t)r * A + c] : 0.f);
}

__global__ void k_pack_layer(const float* __restrict__ w, int N, int K, int K16, __nv_bfloat16* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K16; i += gridDim.x * blockDim.x) {
    const int n = i / K16, k = i % K16;
    const float v = (k < K) ? w[(size_t)n * K + k] : 0.f;
    dst[packed_index(n, k, N, K16)] = __float2bfloat16_rn(v);
  }
}

// gate GEMM operand: W'[e,k] = ln_w[k] * wg[e,k]; row e = bf16 hi part, row 16+e = bf16 lo part
__global__ void k_pack_gate(const float* __restrict__ ln_w, const float* __restrict__ wg, int E, int K,
                            __nv_bfloat16* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < GATE_N * K; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i % K;
    const int e = n % 16;
    float v = 0.f;
    if (e < E) {
      const float wv = ln_w[k] * wg[(size_t)e * K + k];
      const float hi = bf16_round(wv);
      v = (n < 16) ? hi : (wv - hi);
    }
    dst[packed_index(n, k, GATE_N, K)] = __float2bfloat16_rn(v);
  }
}

// c1[e] = sum_k ln_w[k]*wg[e,k] ; c0[e] = sum_k ln_b[k]*wg[e,k]   (one warp per expert)
__global__ void k_gate_consts(const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                              const float* __restrict__ wg, int E, int K, float* __restrict__ c0,
                              float* __restrict__ c1) {
  const int e = blockIdx.x, lane = threadIdx.x;
  if (e >= E) return;
  double a0 = 0.0, a1 = 0.0;
  for (int k = lane; k < K; k += 32) {
    a1 += (double)ln_w[k] * (double)wg[(size_t)e * K + k];
    a0 += (double)ln_b[k] * (double)wg[(size_t)e * K + k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (lane == 0) { c0[e] = (float)a0; c1[e] = (float)a1; }
}

// dst[e][n] = bf16(b_skip[e][n]) + bf16(b_xyz[n])
__global__ void k_sum_bias(const float* __restrict__ b_skip, const float* __restrict__ b_xyz, int E, int N,
                           float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E * N) dst[i] = bf16_round(b_skip[i]) + bf16_round(b_xyz[i % N]);
}

__global__ void k_round_copy(const float* __restrict__ src, int n, int round_bf16, float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = round_bf16 ? bf16_round(src[i]) : src[i];
}

__global__ void k_pack_layer_wide(const float* __restrict__ w, int N, int K, int K16, __nv_bfloat16* __restrict__ dst);
__global__ void k_pack_gate_wide(const float* __restrict__ ln_w, const float* __restrict__ wg, int E, int K,
                                 __nv_bfloat16* __restrict__ dst);

struct TcHost {
  TcParams p;
  float* fblob;
  size_t fblob_floats;
};

bool tc_supported(const Model* m) {
  const snb_model_desc& d = m->d;
  // width 256: TS kernels (hidden activations in tensor memory); width 512 and every mip model: wide kernels
  // (snb_tc_wide.cuh: A operand in shared memory, a layer's accumulators own all of tensor memory)
  return (d.width == 256 || d.width == 512) && d.num_experts <= MAX_E && d.hidden2 <= 256 && d.hidden2 % 64 == 0 &&
         d.gate_layers >= 1 && d.gate_layers <= 4 && m->xyz_in <= 80 && m->cat_in - d.width <= 80 && d.expert_layers <= 16 &&
         d.expert_layers >= 1 && d.appearance_dim % 4 == 0 && d.pos_xyz_freqs == 12 && d.pos_dir_freqs == 4 &&
         (d.skip_layer >= 0 || (d.width == 256 && !d.mip));
}

struct TcOwner { TcHost h; uint8_t* wblob; uint8_t* wblob_w; __nv_bfloat16* emb_cat; };
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
// defaults of a new model object; the SNB_* variables are read here once per model (A/B scripts), never on the call path
void tuning_from_env(snb_tuning* t) {
  const int cg = env_int("SNB_CG", 1);
  t->cta_group_front = env_int("SNB_CG_FRONT", cg) == 2 ? 2 : 1;
  t->cta_group_back = env_int("SNB_CG_BACK", cg) == 2 ? 2 : 1;
  t->ts = env_int("SNB_TS", 1) != 0;
  t->ts_front = t->ts && env_int("SNB_TS_FRONT", 1) != 0;
  t->wide = env_int("SNB_WIDE", 0) != 0;
  t->route_full = getenv("SNB_ROUTE_FULL") != nullptr;
  t->no_overlap = getenv("SNB_NO_OVERLAP") != nullptr;
  t->pipe_depth = env_int("SNB_PIPE_DEPTH", -1);
  t->route_sms = env_int("SNB_ROUTE_SMS", -1);
  t->back_partition = getenv("SNB_BACK_PART") != nullptr;
  t->no_ray_source = getenv("SNB_NO_RAY_SOURCE") != nullptr;
  t->gather_h = getenv("SNB_GATHER_H") != nullptr;
  t->front_ab = env_int("SNB_FRONT_AB", 0);
}
// which kernel family evaluates a model: wide for width 512 and mip models (SNB_WIDE=1 forces it for A/B tests)
static bool tc_use_wide(const Model* m) {
  return m->d.width != 256 || m->d.mip || m->tune.wide;
}

void tc_release(Model* m) {
  TcOwner* own = (TcOwner*)m->tc_blob;
  if (!own) return;
  if (own->wblob) cudaFree(own->wblob);
  if (own->wblob_w) cudaFree(own->wblob_w);
  if (own->h.fblob) cudaFree(own->h.fblob);
  if (own->emb_cat) cudaFree(own->emb_cat);
  delete own;
  m->tc_blob = nullptr;
}

int tc_pack_weights(Model* m, const snb_weights* w, cudaStream_t st) {
  const snb_model_desc& d = m->d;
  const int E = d.num_experts, L = d.expert_layers, H2 = d.hidden2, MW = d.width;
  const bool wide = tc_use_wide(m), narrow = (d.width == 256 && !d.mip);
  TcOwner* own = nullptr;
  if (m->tc_blob == nullptr) {
    own = new TcOwner();
    memset(own, 0, sizeof(TcOwner));
    size_t wbytes = 0, nf = 0;
    TcParams& p = own->h.p;
    auto add_layer = [&](TcLayer& l, int N, int K) {
      l.K16 = (uint32_t)((K + 15) / 16 * 16);
      l.N = (uint32_t)N;
      l.w_off = l.w_off_w = (uint32_t)wbytes;      // both image kinds have the same size: one offset table
      wbytes += (size_t)N * l.K16 * 2;
      wbytes = align_up(wbytes, 128);
      l.b_off = (uint32_t)nf;
      nf += (size_t)align_up((size_t)N, 64);
    };
    p.n_front = 1 + d.gate_layers;
    add_layer(p.front[0], MW, m->xyz_in);
    for (int i = 0; i < d.gate_layers; ++i) add_layer(p.front[1 + i], MW, MW);
    add_layer(p.gate, GATE_N, MW);
    add_layer(p.back[0], MW, MW);
    add_layer(p.back[1], H2, m->cat_in);
    p.n_expert = L;
    size_t e0_w = wbytes, e0_f = nf;
    for (int j = 0; j < L; ++j) add_layer(p.expert[j], MW, MW);
    p.expert_w_stride = p.expert_w_stride_w = (uint32_t)(wbytes - e0_w);
    p.expert_b_stride = (uint32_t)(nf - e0_f);
    wbytes = e0_w + (size_t)p.expert_w_stride * E;
    nf = e0_f + (size_t)p.expert_b_stride * E;
    auto addf = [&](size_t n) { size_t o = nf; nf += align_up(n, 64); return (uint32_t)o; };
    p.o_c0 = addf(MAX_E); p.o_c1 = addf(MAX_E);
    p.o_wsig = addf(MW); p.o_bsig = addf(1); p.o_wcol = addf((size_t)3 * H2); p.o_bcol = addf(3);
    p.o_b3x = addf((size_t)E * MW);
    p.recompute_h = (!m->tune.gather_h && d.skip_layer >= 0) ? 1 : 0;
    p.mip = d.mip;
    p.E = E; p.skip_layer = d.skip_layer; p.pos_xyz_freqs = d.pos_xyz_freqs; p.pos_dir_freqs = d.pos_dir_freqs;
    p.appearance_dim = d.appearance_dim; p.appearance_count = d.appearance_count; p.hidden2 = H2; p.x_cols = m->x_cols;
    if (narrow) SNB_CHECK_CUDA(cudaMalloc((void**)&own->wblob, wbytes));
    if (wide) SNB_CHECK_CUDA(cudaMalloc((void**)&own->wblob_w, wbytes));
    SNB_CHECK_CUDA(cudaMalloc((void**)&own->h.fblob, nf * sizeof(float)));
    own->h.fblob_floats = nf;
    p.wblob = own->wblob;
    p.wblob_w = own->wblob_w;
    p.fblob = own->h.fblob;
    p.emb_a = m->emb_a;
    p.cat_cols = (int32_t)(p.back[1].K16 - (uint32_t)MW);
    if (d.appearance_dim > 0 && p.cat_cols > 0) {
      SNB_CHECK_CUDA(cudaMalloc((void**)&own->emb_cat, (size_t)d.appearance_count * p.cat_cols * sizeof(__nv_bfloat16)));
      p.emb_cat = own->emb_cat;
    }
    m->tc_blob = own;
    m->tc_bytes = wbytes;
  } else {
    own = (TcOwner*)m->tc_blob;
  }
  TcParams& p = own->h.p;
  SNB_CHECK_CUDA(cudaMemsetAsync(own->h.fblob, 0, own->h.fblob_floats * sizeof(float), st));
  auto pack = [&](const TcLayer& l, const float* wsrc, int K, size_t extra_w, const float* bsrc, size_t extra_b) -> int {
    if (own->wblob) {
      k_pack_layer<<<256, 256, 0, st>>>(wsrc, (int)l.N, K, (int)l.K16, (__nv_bfloat16*)(own->wblob + l.w_off + extra_w));
      SNB_CHECK_LAUNCH("k_pack_layer");
    }
    if (own->wblob_w) {
      k_pack_layer_wide<<<256, 256, 0, st>>>(wsrc, (int)l.N, K, (int)l.K16, (__nv_bfloat16*)(own->wblob_w + l.w_off_w + extra_w));
      SNB_CHECK_LAUNCH("k_pack_layer_wide");
    }
    k_round_copy<<<(unsigned)cdiv(l.N, 256), 256, 0, st>>>(bsrc, (int)l.N, 1, own->h.fblob + l.b_off + extra_b);
    SNB_CHECK_LAUNCH("k_round_copy");
    return SNB_OK;
  };
  int rc;
  // sources are the library-owned fp32 copies uploaded by snb_api upload(); experts there are [E][N][K]
  if ((rc = pack(p.front[0], m->xyz_w, m->xyz_in, 0, m->xyz_b, 0))) return rc;
  for (int i = 0; i < d.gate_layers; ++i)
    if ((rc = pack(p.front[1 + i], m->gate_w[i], MW, 0, m->gate_b[i], 0))) return rc;
  if ((rc = pack(p.back[0], m->l1_w, MW, 0, m->l1_b, 0))) return rc;
  if ((rc = pack(p.back[1], m->l2_w, m->cat_in, 0, m->l2_b, 0))) return rc;
  for (int e = 0; e < E; ++e)
    for (int j = 0; j < L; ++j)
      if ((rc = pack(p.expert[j], m->exp_w[j] + (size_t)e * MW * MW, MW, (size_t)e * p.expert_w_stride,
                     m->exp_b[j] + (size_t)e * MW, (size_t)e * p.expert_b_stride)))
        return rc;
  if (own->wblob) {
    k_pack_gate<<<32, 256, 0, st>>>(m->ln_w, m->wg, E, MW, (__nv_bfloat16*)(own->wblob + p.gate.w_off));
    SNB_CHECK_LAUNCH("k_pack_gate");
  }
  if (own->wblob_w) {
    k_pack_gate_wide<<<32, 256, 0, st>>>(m->ln_w, m->wg, E, MW, (__nv_bfloat16*)(own->wblob_w + p.gate.w_off_w));
    SNB_CHECK_LAUNCH("k_pack_gate_wide");
  }
  k_gate_consts<<<E, 32, 0, st>>>(m->ln_w, m->ln_b, m->wg, E, MW, own->h.fblob + p.o_c0, own->h.fblob + p.o_c1);
  SNB_CHECK_LAUNCH("k_gate_consts");
  auto cpf = [&](uint32_t off, const float* src, int n, int round) -> int {
    k_round_copy<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(src, n, round, own->h.fblob + off);
    SNB_CHECK_LAUNCH("k_round_copy");
    return SNB_OK;
  };
  if ((rc = cpf(p.o_wsig, m->sigma_w, MW, 1))) return rc;   // bf16 Linear under autocast
  if ((rc = cpf(p.o_bsig, m->sigma_b, 1, 1))) return rc;
  if ((rc = cpf(p.o_wcol, m->color_w, 3 * d.hidden2, 1))) return rc;
  if ((rc = cpf(p.o_bcol, m->color_b, 3, 1))) return rc;
  if (d.skip_layer >= 0) {
    k_sum_bias<<<(unsigned)cdiv((int64_t)E * MW, 256), 256, 0, st>>>(m->exp_b[d.skip_layer], m->xyz_b, E, MW, own->h.fblob + p.o_b3x);
    SNB_CHECK_LAUNCH("k_sum_bias");
  }
  if (own->emb_cat) {
    k_pack_emb_cat<<<(unsigned)cdiv((int64_t)d.appearance_count * p.cat_cols, 256), 256, 0, st>>>(
        m->emb_a, d.appearance_count, d.appearance_dim, m->dir_in, p.cat_cols, own->emb_cat);
    SNB_CHECK_LAUNCH("k_pack_emb_cat");
  }
  (void)w;
  return SNB_OK;
}

// ------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------
struct __align__(16) SmemCtl {
  uint64_t full[2 * NSTAGE];      // CTA pairs use 2*NSTAGE half-size ring slots
  uint64_t empty[2 * NSTAGE];
  uint64_t acc_full[2];
  uint64_t a_ready[NCHUNK];
  uint64_t peer_ok[2 * NSTAGE];   // CTA pairs: "the peer CTA's A chunk + weight half for this ring slot are ready" (leader only)
  uint32_t tmem_base;
  uint32_t pad;
};
static constexpr size_t SM_A = 0;
static constexpr size_t SM_RING = A_BYTES;
static constexpr size_t SM_BIAS = SM_RING + (size_t)NSTAGE * STAGE_BYTES;   // 2 x 256 floats
static constexpr size_t SM_VEC = SM_BIAS + 2 * 256 * 4;                     // head vectors: wsig[256], wcol[3*256]
static constexpr size_t SM_VEC_FLOATS = 256 + 3 * 256 + 64;
static constexpr size_t SM_RED = SM_VEC + SM_VEC_FLOATS * 4;                // row-reduction scratch [4][4][128] floats
static constexpr size_t SM_RED_FLOATS = 4 * 4 * 128;
static constexpr size_t SM_CTL = SM_RED + SM_RED_FLOATS * 4;
static constexpr size_t SM_TOTAL = SM_CTL + sizeof(SmemCtl) + 1024;         // + alignment slack
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");

static constexpr int TL_N = 2048;
__device__ __forceinline__ void tl_mark(unsigned long long* tl, int role, int& n, int tag) {
  if (tl && blockIdx.x == 0 && n < TL_N) tl[role * TL_N + n++] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
}

struct Pipe {               // running counters of one role
  uint32_t slice = 0;       // weight slices produced / consumed so far
  uint32_t a_use[NCHUNK] = {0, 0, 0, 0, 0, 0};
  uint32_t acc_use[2] = {0, 0};
};

__device__ __forceinline__ uint32_t a_chunk_addr(uint32_t a_base, int row, int col8 /*col/8*/) {
  return a_base + (uint32_t)(row >> 3) * SBO_A + (uint32_t)col8 * 128u + (uint32_t)(row & 7) * 16u;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// 32 lanes x 16 columns: thread t of the warp gets TMEM lane (base_lane + t), columns c..c+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ---- producer: stream the K-slices of one layer through the ring ----
// CG = 2 (CTA pair): each CTA streams its own half of the slice (rows [rank*N/2, +N/2) of the image are the
// contiguous half of the bytes) -- half the L2 traffic and half the shared-memory landing traffic per SM.
template <int CG>
__device__ __forceinline__ void produce_layer(const uint8_t* wsrc, uint32_t N, uint32_t K16, uint8_t* ring,
                                              SmemCtl* ctl, Pipe& pp, uint32_t rank) {
  const uint32_t nsl = (K16 + 63) / 64;
  for (uint32_t j = 0; j < nsl; ++j) {
    const uint32_t klen = min(64u, K16 - 64u * j);
    const uint32_t bytes = N * klen * 2 / CG;
    constexpr uint32_t NS = NSTAGE * CG, SB = STAGE_BYTES / CG;
    const uint32_t stage = pp.slice % NS, phase = (pp.slice / NS) & 1;
    mbar_wait(&ctl->empty[stage], phase ^ 1);
    mbar_arrive_expect_tx(&ctl->full[stage], bytes);
    bulk_g2s(ring + (size_t)stage * SB, wsrc + (size_t)N * 64 * 2 * j + (size_t)rank * bytes, bytes, &ctl->full[stage]);
    ++pp.slice;
  }
}

// ---- MMA issuer: one layer = K16/16 tcgen05.mma instructions into accumulator buffer `buf` ----
// CG = 1: this CTA's 128 x N tile.  CG = 2: the leader CTA (rank 0) issues M = 256 instructions for the pair
// (rows 0-127 = its own A tile / TMEM, rows 128-255 = the peer's); the peer CTA's warp 1 only forwards
// "my A chunk and my weight half are in place" to the leader's peer_ok barrier of the ring slot.
// `chunk0`: first 64-column chunk of the A tile this segment reads (its a_ready barriers are chunk0 + j);
// `cont`: keep accumulating into the buffer (second segment of a layer); `commit`: this segment ends the layer.
template <int CG>
__device__ __forceinline__ void mma_layer(uint32_t N, uint32_t K16, uint32_t a_base, uint32_t ring_base,
                                          uint32_t tmem_base, int buf, SmemCtl* ctl, Pipe& pp, uint32_t rank,
                                          unsigned long long* tl = nullptr, int* tn = nullptr, uint32_t chunk0 = 0,
                                          bool cont = false, bool commit = true) {
  const uint32_t nsl = (K16 + 63) / 64;
  const uint32_t idesc = umma_idesc_bf16(TILE * CG, (int)N);
  const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
  for (uint32_t j = 0; j < nsl; ++j) {
    const uint32_t klen = min(64u, K16 - 64u * j);
    constexpr uint32_t NS = NSTAGE * CG, SB = STAGE_BYTES / CG;
    const uint32_t stage = pp.slice % NS, phase = (pp.slice / NS) & 1;
    if (CG == 2 && rank != 0) {
      // peer: forward "my half of the weight slice landed" (normally far ahead of the MMAs); its A chunks are
      // signalled by its epilogue warps directly on the leader's a_ready.  The slot is recycled through the
      // multicast commit on empty[stage].
      mbar_wait(&ctl->full[stage], phase);
      mbar_arrive_remote(mapa_shared(smem_u32(&ctl->peer_ok[stage]), 0));
      ++pp.slice;
      continue;
    }
    // CG == 2: 16 local + 16 remote (release.cluster) warp arrivals.  The waiting thread itself never reads the
    // peer's data -- the MMA it issues reads each CTA's own shared memory through that CTA's async proxy, which the
    // writers fenced (fence.proxy.async) before arriving -- so the plain (CTA-scope) wait is used: a cluster-scope
    // acquire here costs an L1 invalidation per K-slice.
    mbar_wait(&ctl->a_ready[chunk0 + j], pp.a_use[chunk0 + j] & 1);   // A columns of this slice written + fenced
    ++pp.a_use[chunk0 + j];
    if (tn) tl_mark(tl, 1, *tn, 100 + (int)j);
    mbar_wait(&ctl->full[stage], phase);                // weight slice (or this CTA's half of it) landed
    if (tn) tl_mark(tl, 1, *tn, 110 + (int)j);
    if (CG == 2) mbar_wait(&ctl->peer_ok[stage], phase);
    if (tn) tl_mark(tl, 1, *tn, 130 + (int)j);
    tc_fence_after();
    const uint32_t b_base = ring_base + stage * SB;
    for (uint32_t t = 0; t < klen / 16; ++t) {
      const uint64_t da = op_desc(a_base + (8u * (chunk0 + j) + 2u * t) * 128u, 128u, SBO_A);
      const uint64_t db = op_desc(b_base + (2u * t) * 128u, 128u, klen * 16u);
      const uint32_t acc = (cont || (j | t)) ? 1u : 0u;
      if (CG == 2) umma_bf16_pair(d_tmem, da, db, idesc, acc);
      else umma_bf16(d_tmem, da, db, idesc, acc);
    }
    if (CG == 2) umma_commit_pair(&ctl->empty[stage], 3);   // frees the ring slot in BOTH CTAs
    else umma_commit(&ctl->empty[stage]);
    ++pp.slice;
  }
  if (commit) {
    if (CG == 2) {
      if (rank == 0) umma_commit_pair(&ctl->acc_full[buf], 3);  // accumulators complete -> both epilogues
    } else {
      umma_commit(&ctl->acc_full[buf]);
    }
  }
  if (tn) tl_mark(tl, 1, *tn, 120);
}

// ---- epilogue helpers -------------------------------------------------------------------
struct EpiCtx {
  int lane, q, cs, row;     // TMEM lane quarter, column sub-slice (16 of every 64 columns), tile row
  int et;                   // 0..511
  uint32_t lane_base;       // (q*32) << 16
  uint32_t remote_a_ready;  // CTA pairs, peer CTA only: cluster address of the LEADER's a_ready[0] (0 = arrive locally)
};

__device__ __forceinline__ void epi_wait_acc(SmemCtl* ctl, Pipe& pp, int buf) {
  mbar_wait_backoff(&ctl->acc_full[buf], pp.acc_use[buf] & 1);
  ++pp.acc_use[buf];
  tc_fence_after();
}
// every epilogue warp arrives once per chunk (a_ready count = 16): all lanes fence, lane 0 arrives
// CTA pairs: the MMAs are issued by the leader CTA only, so BOTH CTAs' epilogue warps arrive on the leader's
// a_ready[c] (count 2 x 16): the leader's warps locally, the peer's through the cluster address.
__device__ __forceinline__ void epi_signal_chunk(SmemCtl* ctl, int c, int lane, uint32_t remote_a_ready = 0) {
  fence_proxy_async_smem();    // my st.shared -> visible to the async proxy (tcgen05.mma operand reads)
  tc_fence_before();           // my tcgen05.ld of the old accumulator happen-before the MMA that overwrites it
  __syncwarp();
  if (lane == 0) {
    if (remote_a_ready) mbar_arrive_remote(remote_a_ready + (uint32_t)c * 8u);
    else mbar_arrive(&ctl->a_ready[c]);
  }
}
// stage the (bf16-rounded) bias of a layer into shared memory (double buffered by `slot`)
__device__ __forceinline__ void epi_load_bias(const float* __restrict__ bias, int N, float* sbias, int slot, int et) {
  float* dst = sbias + slot * 256;
  if (et < N) dst[et] = bias[et];
  epi_bar_sync();
}

// pack two fp32 into bf16x2 (lo = first argument) with optional ReLU fused into the conversion
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// y = act(acc + bias (+ skip)) -> bf16 -> A tile; this thread owns columns [64c + 16cs, +16) of every chunk c.
// TMEM loads are issued two chunks at a time (one wait per pair).
// `xrow` (nullable): bf16 row of the expert input for the skip connection (tutel_moe_layer_nobatch.py:911-916).
template <bool RELU>
__device__ __forceinline__ void epi_hidden(uint32_t tmem_acc, const float* sb, int nchunks, uint32_t a_base,
                                           const EpiCtx& ec, const __nv_bfloat16* xrow, SmemCtl* ctl,
                                           unsigned long long* tl = nullptr, int* tn = nullptr) {
  for (int c2 = 0; c2 < nchunks; c2 += 2) {
    uint32_t v[2][16];
    const int colA = c2 * 64 + ec.cs * 16;
    tmem_ld16(tmem_acc + (uint32_t)colA, v[0]);
    if (c2 + 1 < nchunks) tmem_ld16(tmem_acc + (uint32_t)(colA + 64), v[1]);
    uint4 xs[2][2];
    if (xrow) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        xs[h][0] = *reinterpret_cast<const uint4*>(xrow + colA + 64 * h);
        xs[h][1] = *reinterpret_cast<const uint4*>(xrow + colA + 64 * h + 8);
      }
    }
    tmem_ld_wait();
    if (tn) tl_mark(tl, 0, *tn, 70 + c2);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (c2 + h >= nchunks) break;
      const int col0 = colA + 64 * h;
      const float4* b4 = reinterpret_cast<const float4*>(sb + col0);
      uint32_t pk[8];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 b = b4[j4];
        float f0 = __uint_as_float(v[h][4 * j4 + 0]) + b.x;
        float f1 = __uint_as_float(v[h][4 * j4 + 1]) + b.y;
        float f2 = __uint_as_float(v[h][4 * j4 + 2]) + b.z;
        float f3 = __uint_as_float(v[h][4 * j4 + 3]) + b.w;
        if (xrow) {
          // reference: h = bf16(Linear) ; h = bf16(h + x)
          const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xs[h][0]);
          __nv_bfloat162 xa = *reinterpret_cast<const __nv_bfloat162*>(&xw[2 * j4]);
          __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162*>(&xw[2 * j4 + 1]);
          f0 = bf16_round(f0) + __bfloat162float(xa.x);
          f1 = bf16_round(f1) + __bfloat162float(xa.y);
          f2 = bf16_round(f2) + __bfloat162float(xb.x);
          f3 = bf16_round(f3) + __bfloat162float(xb.y);
        }
        pk[2 * j4] = pack2<RELU>(f0, f1);
        pk[2 * j4 + 1] = pack2<RELU>(f2, f3);
      }
      st_shared_v4(a_chunk_addr(a_base, ec.row, col0 / 8), pk[0], pk[1], pk[2], pk[3]);
      st_shared_v4(a_chunk_addr(a_base, ec.row, col0 / 8 + 1), pk[4], pk[5], pk[6], pk[7]);
      if (tn) tl_mark(tl, 0, *tn, 60 + c2 + h);
      epi_signal_chunk(ctl, c2 + h, ec.lane, ec.remote_a_ready);
      if (tn) tl_mark(tl, 0, *tn, 80 + c2 + h);
    }
  }
}

// Store the 256-column bf16 A tile to global memory, one full 512-byte row per warp instruction
// (lane l moves the 16-byte group l of the row): fully coalesced.  Warp (q, cs) moves rows
// q*32 + cs*8 + i, i = 0..7.  `row_ptr(r)` returns the destination row pointer or nullptr.
template <typename RowPtr>
__device__ __forceinline__ void a_tile_to_global(const uint8_t* smem_a, const EpiCtx& ec, RowPtr row_ptr) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ec.q * 32 + ec.cs * 8 + i;
    __nv_bfloat16* dst = row_ptr(r);
    const uint8_t* src = smem_a + (size_t)(r >> 3) * SBO_A + (size_t)ec.lane * 128 + (size_t)(r & 7) * 16;
    if (dst) *reinterpret_cast<uint4*>(dst + ec.lane * 8) = *reinterpret_cast<const uint4*>(src);
  }
}

// ------------------------------------------------------------------------------------------
// positional encoding helpers (models/nerf.py:21-26): [x, sin(2^k x), cos(2^k x)]_k, per-frequency
// layout [sin xyz | cos xyz].  Base angle by sincosf, higher octaves by the double-angle recurrence
// (error ~2^k ulp << bf16 rounding of the operand).
// ------------------------------------------------------------------------------------------
template <int F>
__device__ __forceinline__ void pe_to_bf16(const float (&p)[3], __nv_bfloat16* dst /* 3 + 6F values */) {
  float s[3], c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    dst[a] = __float2bfloat16_rn(p[a]);
    sincosf(p[a], &s[a], &c[a]);
  }
#pragma unroll
  for (int k = 0; k < F; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      dst[3 + 6 * k + a] = __float2bfloat16_rn(s[a]);
      dst[3 + 6 * k + 3 + a] = __float2bfloat16_rn(c[a]);
      const float s2 = 2.f * s[a] * c[a];
      const float c2 = 1.f - 2.f * s[a] * s[a];
      s[a] = s2;
      c[a] = c2;
    }
  }
}

// write `n8` 16-byte groups (8 bf16 each) of a row into A columns starting at col8*8
__device__ __forceinline__ void a_store_row(uint32_t a_base, int row, int col8, const __nv_bfloat16* vals, int n8) {
  const uint4* v = reinterpret_cast<const uint4*>(vals);
  for (int q = 0; q < n8; ++q) {
    uint4 t = v[q];
    st_shared_v4(a_chunk_addr(a_base, row, col8 + q), t.x, t.y, t.z, t.w);
  }
}

// ------------------------------------------------------------------------------------------
// common prologue of both kernels
// ------------------------------------------------------------------------------------------
template <int CG>
__device__ __forceinline__ SmemCtl* cta_setup(uint8_t* smem, int warp) {
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + SM_CTL);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NSTAGE; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], 1);
      mbar_init(&ctl->peer_ok[i], 1);
    }
    mbar_init(&ctl->acc_full[0], 1);
    mbar_init(&ctl->acc_full[1], 1);
    for (int i = 0; i < NCHUNK; ++i) mbar_init(&ctl->a_ready[i], EPI_WARPS * CG);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_pair<512>(&ctl->tmem_base);
    else tmem_alloc<512>(&ctl->tmem_base);
  }
  return ctl;
}
// all threads: make the barrier inits / TMEM address visible (cluster-wide for CTA pairs)
template <int CG>
__device__ __forceinline__ void cta_setup_sync() {
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
}
template <int CG>
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base, int warp) {
  tc_fence_before();
  if (CG == 2) cluster_sync_all();     // the peer may still read this CTA's shared memory / arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}
__device__ __forceinline__ EpiCtx epi_ctx(int warp, int lane, SmemCtl* ctl, int cg, uint32_t rank) {
  EpiCtx ec;
  ec.remote_a_ready = (cg == 2 && rank != 0) ? mapa_shared(smem_u32(&ctl->a_ready[0]), 0) : 0u;
  ec.lane = lane;
  ec.q = warp & 3;                 // hardware rule: a warp reads the TMEM lanes 32*(warp_id % 4) ...
  ec.cs = (warp - 2) >> 2;         // ... so warps {2..5}, {6..9}, {10..13}, {14..17} cover all 4 quarters each
  ec.row = ec.q * 32 + lane;
  ec.et = (int)threadIdx.x - 64;
  ec.lane_base = (uint32_t)(ec.q * 32) << 16;
  return ec;
}

// ------------------------------------------------------------------------------------------
// launch #1: encode + xyz + external gate MLP + folded LayerNorm/gate GEMM + softmax
// ------------------------------------------------------------------------------------------
template <int FX, int CG>
__global__ void __launch_bounds__(THREADS, 1) k_front(TcParams P, const float* __restrict__ x, int64_t S,
                                                      __nv_bfloat16* __restrict__ H, float* __restrict__ gates) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int t_first0 = CG * ((int)blockIdx.x / CG), t_stride = (int)gridDim.x;   // the CG CTAs of a cluster take CG consecutive tiles
  SmemCtl* ctl = cta_setup<CG>(smem, warp);
  float* sbias = reinterpret_cast<float*>(smem + SM_BIAS);
  float* sred = reinterpret_cast<float*>(smem + SM_RED);      // [2][4][128]: sum / sumsq partials per cs
  cta_setup_sync<CG>();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t a_base = smem_u32(smem + SM_A), ring_base = smem_u32(smem + SM_RING);
  const int n_tiles = (int)((S + TILE - 1) / TILE);
  const int NL = P.n_front;
  Pipe pp;

  if (warp == 0) {
    if (lane == 0)
      for (int tb = t_first0; tb < n_tiles; tb += t_stride) {
      const int t = tb + (int)rank;
        for (int l = 0; l < NL; ++l)
          produce_layer<CG>(P.wblob + P.front[l].w_off, P.front[l].N, P.front[l].K16, smem + SM_RING, ctl, pp, rank);
        produce_layer<CG>(P.wblob + P.gate.w_off, GATE_N, MW, smem + SM_RING, ctl, pp, rank);
      }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t li = 0;
      int tn = 0;
      for (int tb = t_first0; tb < n_tiles; tb += t_stride) {
      const int t = tb + (int)rank;
        tl_mark(P.tl, 1, tn, 1);
        for (int l = 0; l < NL; ++l, ++li)
          mma_layer<CG>(P.front[l].N, P.front[l].K16, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn);
        mma_layer<CG>(GATE_N, MW, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn);
        ++li;
      }
    }
  } else {
    const EpiCtx ec = epi_ctx(warp, lane, ctl, CG, rank);
    const int row = ec.row;
    uint32_t li = 0;
    int tn = 0;
    unsigned long long* tl = (warp == 2 && lane == 0) ? P.tl : nullptr;
    float pn[3] = {0.f, 0.f, 0.f};          // xyz of this thread's row in the NEXT tile (prefetched one tile ahead)
    if (ec.cs == 0) {
      const int64_t s0 = (int64_t)(t_first0 + (int)rank) * TILE + row;
      if (t_first0 < n_tiles && s0 < S) { pn[0] = x[s0 * P.x_cols]; pn[1] = x[s0 * P.x_cols + 1]; pn[2] = x[s0 * P.x_cols + 2]; }
    }
    for (int tb = t_first0; tb < n_tiles; tb += t_stride) {
      const int t = tb + (int)rank;
      const int64_t s = (int64_t)t * TILE + row;
      const bool valid = s < S;
      tl_mark(tl, 0, tn, 1);
      // ---- stage PE(xyz) as the A operand of the xyz layer (K16 = 80 -> chunks 0 and 1) ----
      constexpr int NPE = 3 + 6 * FX, NPAD = (NPE + 15) / 16 * 16;
      if (ec.cs == 0) {
        float p[3] = {pn[0], pn[1], pn[2]};
        {
          const int64_t sn = s + (int64_t)t_stride * TILE;
          pn[0] = pn[1] = pn[2] = 0.f;
          if (tb + t_stride < n_tiles && sn < S) { pn[0] = x[sn * P.x_cols]; pn[1] = x[sn * P.x_cols + 1]; pn[2] = x[sn * P.x_cols + 2]; }
        }
        __align__(16) __nv_bfloat16 pe[NPAD];
        pe_to_bf16<FX>(p, pe);
#pragma unroll
        for (int i = NPE; i < NPAD; ++i) pe[i] = __float2bfloat16_rn(0.f);
        a_store_row(a_base, row, 0, pe, NPAD / 8);
      }
      for (int c = 0; c < (NPAD + 63) / 64; ++c) epi_signal_chunk(ctl, c, lane, ec.remote_a_ready);
      tl_mark(tl, 0, tn, 2);
      for (int l = 0; l < NL; ++l, ++li) {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.front[l].b_off, MW, sbias, buf, ec.et);
        if (l == 1 && !P.recompute_h) {
          // h (written into the A tile by the L0 epilogue of ALL warps, complete after the barrier above) -> H:
          // coalesced 512-byte rows, overlapping the L1 MMAs; a second barrier fences the L1 epilogue's A writes
          const int64_t s0 = (int64_t)t * TILE;
          a_tile_to_global(smem + SM_A, ec, [&](int r) { return (s0 + r < S) ? (H + (s0 + r) * MW) : (__nv_bfloat16*)nullptr; });
          epi_bar_sync();
        }
        tl_mark(tl, 0, tn, 10 + l);
        epi_wait_acc(ctl, pp, buf);
        tl_mark(tl, 0, tn, 20 + l);
        const uint32_t tacc = tmem_base + ec.lane_base + (uint32_t)buf * 256u;
        const float* sb = sbias + buf * 256;
        if (l == 0) {
          // h = xyz Linear (act none): bf16 -> A (operand of the gate MLP) and -> H[s] (expert input of launch #2)
          epi_hidden<false>(tacc, sb, 4, a_base, ec, nullptr, ctl, tl, &tn);
          tl_mark(tl, 0, tn, 90);
        } else if (l < NL - 1) {
          epi_hidden<true>(tacc, sb, 4, a_base, ec, nullptr, ctl, tl, &tn);
        } else {
          // last gate-MLP layer: g = bf16(Linear) -> A (operand of the folded gate GEMM) + LayerNorm statistics
          float sum = 0.f, sq = 0.f;
          for (int c = 0; c < 4; ++c) {
            const int col0 = c * 64 + ec.cs * 16;
            uint32_t v[16];
            tmem_ld16(tacc + (uint32_t)col0, v);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float g0 = bf16_round(__uint_as_float(v[j]) + sb[col0 + j]);
              const float g1 = bf16_round(__uint_as_float(v[j + 1]) + sb[col0 + j + 1]);
              sum += g0 + g1;
              sq = fmaf(g0, g0, fmaf(g1, g1, sq));
              pk[j / 2] = pack_bf16x2(g0, g1);
            }
            st_shared_v4(a_chunk_addr(a_base, row, col0 / 8), pk[0], pk[1], pk[2], pk[3]);
            st_shared_v4(a_chunk_addr(a_base, row, col0 / 8 + 1), pk[4], pk[5], pk[6], pk[7]);
            epi_signal_chunk(ctl, c, lane, ec.remote_a_ready);
          }
          sred[(0 * 4 + ec.cs) * 128 + row] = sum;
          sred[(1 * 4 + ec.cs) * 128 + row] = sq;
        }
        tl_mark(tl, 0, tn, 30 + l);
      }
      // ---- folded LayerNorm + gate GEMM epilogue: logits = rstd*(G_hi + G_lo - mean*c1) + c0 ; softmax ----
      {
        const int buf = (int)(li & 1);
        epi_bar_sync();                       // LayerNorm partial sums of all 4 column sub-slices are in sred
        epi_wait_acc(ctl, pp, buf);
        tl_mark(tl, 0, tn, 40);
        if (ec.cs == 0) {
          const float sum = sred[0 * 128 + row] + sred[1 * 128 + row] + sred[2 * 128 + row] + sred[3 * 128 + row];
          const float sq = sred[4 * 128 + row] + sred[5 * 128 + row] + sred[6 * 128 + row] + sred[7 * 128 + row];
          const float mean = sum * (1.f / MW);
          const float var = fmaxf(sq * (1.f / MW) - mean * mean, 0.f);
          const float rstd = rsqrtf(var + 1e-5f);
          const uint32_t tacc = tmem_base + ec.lane_base + (uint32_t)buf * 256u;
          uint32_t hi[16], lo[16];
          tmem_ld16(tacc, hi);
          tmem_ld16(tacc + 16u, lo);
          tmem_ld_wait();
          float lg[MAX_E];
          float mx = -INFINITY;
#pragma unroll
          for (int e = 0; e < MAX_E; ++e) {
            lg[e] = rstd * (__uint_as_float(hi[e]) + __uint_as_float(lo[e]) - mean * P.fblob[P.o_c1 + e]) + P.fblob[P.o_c0 + e];
            if (e < P.E) mx = fmaxf(mx, lg[e]);
          }
          float den = 0.f;
#pragma unroll
          for (int e = 0; e < MAX_E; ++e)
            if (e < P.E) { lg[e] = expf(lg[e] - mx); den += lg[e]; }
          if (valid) {
#pragma unroll
            for (int e = 0; e < MAX_E; ++e)
              if (e < P.E) gates[s * P.E + e] = lg[e] / den;
          }
        }
        tc_fence_before();
        epi_bar_sync();                       // sred is reused by the next tile
        tl_mark(tl, 0, tn, 41);
        ++li;
      }
    }
  }
  cta_teardown<CG>(tmem_base, warp);
}

// ------------------------------------------------------------------------------------------
// tile table for launch #2
// ------------------------------------------------------------------------------------------
static_assert(TILE == EP_TILE, "tile height shared with the expert-parallel plan");

__global__ void __launch_bounds__(256) k_tile_plan(const int* __restrict__ counts, const int* __restrict__ cap_dev, int E,
                                                   int no_batch, int64_t S, int pair, TileTable tt) {
  // pair = 1 (CTA pairs): every bucket gets an even number of tiles (the second CTA of a pair must run the same
  // expert's weights); the padding tile has 0 rows.
  __shared__ int s_row[MAX_E + 2], s_tile[MAX_E + 2], s_kc[MAX_E + 1], s_nt[MAX_E + 1];
  const int cap = *cap_dev;
  if (threadIdx.x == 0) {
    int row = 0, nt = 0, kept_total = 0;
    for (int e = 0; e <= E; ++e) {
      int kc;
      if (e < E) { kc = no_batch ? counts[e] : min(counts[e], cap); kept_total += kc; }
      else kc = (int)S - kept_total;                       // dropped bucket
      int n = (kc + TILE - 1) / TILE;
      if (pair) n = (n + 1) & ~1;
      s_row[e] = row; s_tile[e] = nt; s_kc[e] = kc; s_nt[e] = n;
      tt.seg_start[e] = row;
      row += (kc + TILE - 1) / TILE * TILE;
      nt += n;
    }
    *tt.n_tiles = nt;
    *tt.drop_counter = 0;
  }
  __syncthreads();
  for (int e = 0; e <= E; ++e) {
    const int t0 = s_tile[e], nt = s_nt[e];
    for (int i = threadIdx.x; i < nt; i += blockDim.x) {
      tt.tile_expert[t0 + i] = (e < E) ? e : -1;
      tt.tile_row0[t0 + i] = s_row[e] + i * TILE;
      tt.tile_rows[t0 + i] = max(0, min(TILE, s_kc[e] - i * TILE));
    }
  }
}

__global__ void k_scatter_rows(const int* __restrict__ idx, const int* __restrict__ loc, const int* __restrict__ cap_dev,
                               int E, int no_batch, int64_t S, TileTable tt) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int cap = *cap_dev;
  const int e = idx[s], l = loc[s];
  if (no_batch || l < cap) tt.row2sample[tt.seg_start[e] + l] = (int)s;
  else tt.row2sample[tt.seg_start[E] + atomicAdd(tt.drop_counter, 1)] = (int)s;
}

// ------------------------------------------------------------------------------------------
// launch #2: gather -> experts -> combine -> heads
// ------------------------------------------------------------------------------------------
template <int FD, int CG>
__global__ void __launch_bounds__(THREADS, 1) k_back(TcParams P, TileTable tt, RowIO io,
                                                     const __nv_bfloat16* __restrict__ H) {
  const float* __restrict__ x = io.x;
  const float* __restrict__ gate = io.gate;
  const float* __restrict__ noise = io.noise;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int t_first0 = CG * ((int)blockIdx.x / CG), t_stride = (int)gridDim.x;   // pairs of tiles share the expert (k_tile_plan)
  SmemCtl* ctl = cta_setup<CG>(smem, warp);
  float* sbias = reinterpret_cast<float*>(smem + SM_BIAS);
  float* svec = reinterpret_cast<float*>(smem + SM_VEC);
  float* sred = reinterpret_cast<float*>(smem + SM_RED);      // [4 values][4 cs][128 rows]
  float *s_wsig = svec, *s_wcol = svec + 256;                 // [256], [3][H2]
  const int H2 = P.hidden2;
  for (int i = threadIdx.x; i < MW; i += THREADS) s_wsig[i] = P.fblob[P.o_wsig + i];
  for (int i = threadIdx.x; i < 3 * H2; i += THREADS) s_wcol[i] = P.fblob[P.o_wcol + i];
  cta_setup_sync<CG>();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t a_base = smem_u32(smem + SM_A), ring_base = smem_u32(smem + SM_RING);
  const int n_tiles = *tt.n_tiles;
  const int NE = P.n_expert;
  Pipe pp;

  if (warp == 0) {
    if (lane == 0)
      for (int tb = t_first0; tb < n_tiles; tb += t_stride) {
      const int t = tb + (int)rank;
        const int e = tt.tile_expert[t];
        if (e >= 0) {
          if (P.recompute_h) produce_layer<CG>(P.wblob + P.front[0].w_off, MW, P.front[0].K16, smem + SM_RING, ctl, pp, rank);
          for (int l = 0; l < NE; ++l) {
            produce_layer<CG>(P.wblob + P.expert[l].w_off + (size_t)e * P.expert_w_stride, MW, MW, smem + SM_RING, ctl, pp, rank);
            if (P.recompute_h && l == P.skip_layer)     // the skip term h is re-accumulated by the tensor core
              produce_layer<CG>(P.wblob + P.front[0].w_off, MW, P.front[0].K16, smem + SM_RING, ctl, pp, rank);
          }
        }
        produce_layer<CG>(P.wblob + P.back[0].w_off, P.back[0].N, P.back[0].K16, smem + SM_RING, ctl, pp, rank);
        produce_layer<CG>(P.wblob + P.back[1].w_off, P.back[1].N, P.back[1].K16, smem + SM_RING, ctl, pp, rank);
      }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t li = 0;
      int tn = 0;
      for (int tb = t_first0; tb < n_tiles; tb += t_stride) {
      const int t = tb + (int)rank;
        const int e = tt.tile_expert[t];
        tl_mark(P.tl, 1, tn, 1);
        if (e >= 0) {
          if (P.recompute_h) {      // h = xyz Linear(PE(xyz)); PE lives in A columns [256, 336) = chunks 4, 5
            mma_layer<CG>(MW, P.front[0].K16, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn, 4);
            ++li;
          }
          for (int l = 0; l < NE; ++l, ++li) {
            if (P.recompute_h && l == P.skip_layer) {
              mma_layer<CG>(MW, MW, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn, 0, false, false);
              mma_layer<CG>(MW, P.front[0].K16, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn, 4, true, true);
            } else {
              mma_layer<CG>(MW, MW, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn);
            }
          }
        }
        mma_layer<CG>(P.back[0].N, P.back[0].K16, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn); ++li;
        mma_layer<CG>(P.back[1].N, P.back[1].K16, a_base, ring_base, tmem_base, (int)(li & 1), ctl, pp, rank, P.tl, &tn); ++li;
      }
    }
  } else {
    const EpiCtx ec = epi_ctx(warp, lane, ctl, CG, rank);
    const int row = ec.row;
    const float b_sig = P.fblob[P.o_bsig];
    const float b_col0 = P.fblob[P.o_bcol], b_col1 = P.fblob[P.o_bcol + 1], b_col2 = P.fblob[P.o_bcol + 2];
    uint32_t li = 0;
    int tn = 0;
    unsigned long long* tl = (warp == 2 && lane == 0) ? P.tl : nullptr;
    // per-row inputs of the tile are fetched one tile ahead (dependent-load chain row2sample -> x/gate is
    // hidden behind the previous tile); only the 512-byte h row is gathered at staging time
    struct RowIn { int e, sidx; float g, d0, d1, d2, x0, x1, x2; int ai; };
    auto fetch_row = [&](int t) {
      RowIn r;
      r.e = -1; r.sidx = -1; r.g = 0.f; r.d0 = r.d1 = r.d2 = 0.f; r.x0 = r.x1 = r.x2 = 0.f; r.ai = 0;
      if (t < n_tiles) {
        r.e = tt.tile_expert[t];
        if (row < tt.tile_rows[t]) r.sidx = tt.row2sample[tt.tile_row0[t] + row];
        if (r.sidx >= 0) {
          if (r.e >= 0) r.g = io.wsel ? sel_gate(io.wsel[r.sidx]) : gate[(int64_t)r.sidx * io.g_stride];
          if (ec.cs == 0 && P.recompute_h && r.e >= 0) {
            const float* xr = x + (int64_t)r.sidx * io.x_stride;
            r.x0 = xr[0]; r.x1 = xr[1]; r.x2 = xr[2];
          }
          if (ec.cs == 1) {
            const float* xr = x + (int64_t)r.sidx * io.x_stride;
            r.d0 = xr[P.x_cols - 4]; r.d1 = xr[P.x_cols - 3]; r.d2 = xr[P.x_cols - 2];
            r.ai = min(max((int)xr[P.x_cols - 1], 0), P.appearance_count - 1);
          }
        }
      }
      return r;
    };
    RowIn nxt = fetch_row(t_first0 + (int)rank);
    for (int tb = t_first0; tb < n_tiles; tb += t_stride) {
      const int t = tb + (int)rank;
      const RowIn cur = nxt;
      const int e = cur.e;
      tl_mark(tl, 0, tn, 1);
      const int sidx = cur.sidx;
      const bool valid = sidx >= 0;
      const __nv_bfloat16* hrow = valid ? (H + (int64_t)sidx * MW) : nullptr;
      const float g = cur.g;
      const bool rh = P.recompute_h && e >= 0;     // recompute h = xyz Linear(PE(xyz)) on the tensor core
      // [PE(dir) | appearance | 0-pad] -> A columns [256, K16 of layer "2")  (cs == 1 threads)
      auto write_cat = [&]() {
        if (ec.cs == 1) {
          constexpr int NDIR = 3 + 6 * FD;                   // 27
          constexpr int NCAT_MAX = KA_MAX - MW;              // 96
          __align__(16) __nv_bfloat16 cat[NCAT_MAX];
#pragma unroll
          for (int i = 0; i < NCAT_MAX; ++i) cat[i] = __float2bfloat16_rn(0.f);
          if (valid) {
            float dvec[3] = {cur.d0, cur.d1, cur.d2};
            pe_to_bf16<FD>(dvec, cat);
            const float4* er = reinterpret_cast<const float4*>(P.emb_a + (int64_t)cur.ai * P.appearance_dim);
            for (int i = 0; i < P.appearance_dim / 4; ++i) {
              float4 f = er[i];
              cat[NDIR + 4 * i + 0] = __float2bfloat16_rn(f.x);
              cat[NDIR + 4 * i + 1] = __float2bfloat16_rn(f.y);
              cat[NDIR + 4 * i + 2] = __float2bfloat16_rn(f.z);
              cat[NDIR + 4 * i + 3] = __float2bfloat16_rn(f.w);
            }
          }
          a_store_row(a_base, row, MW / 8, cat, ((int)P.back[1].K16 - MW) / 8);
        }
      };
      if (rh) {
        // ---- stage PE(xyz) into A columns [256, 336): operand of the xyz layer now and of the skip term later ----
        if (ec.cs == 0) {
          constexpr int NPE = 3 + 6 * 12, NPAD = (NPE + 15) / 16 * 16;
          float pxyz[3] = {cur.x0, cur.x1, cur.x2};
          __align__(16) __nv_bfloat16 pe[NPAD];
          pe_to_bf16<12>(pxyz, pe);
#pragma unroll
          for (int i = NPE; i < NPAD; ++i) pe[i] = __float2bfloat16_rn(0.f);
          a_store_row(a_base, row, MW / 8, pe, NPAD / 8);
        }
        epi_signal_chunk(ctl, 4, lane, ec.remote_a_ready);
        epi_signal_chunk(ctl, 5, lane, ec.remote_a_ready);
      } else {
        // ---- stage: expert input rows gathered from H (or zeros for the dropped bucket) + the concat block ----
        // warp (q, cs) gathers rows q*32 + cs*8 + i: one coalesced 512-byte row per instruction
        uint4 hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int sr = __shfl_sync(0xffffffffu, sidx, ec.cs * 8 + i);
          hv[i] = make_uint4(0, 0, 0, 0);
          if (sr >= 0 && e >= 0) hv[i] = *reinterpret_cast<const uint4*>(H + (int64_t)sr * MW + lane * 8);
        }
        write_cat();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          st_shared_v4(a_chunk_addr(a_base, ec.q * 32 + ec.cs * 8 + i, lane), hv[i].x, hv[i].y, hv[i].z, hv[i].w);
        for (int c = 0; c < 4; ++c) epi_signal_chunk(ctl, c, lane, ec.remote_a_ready);
      }
      nxt = fetch_row(t + t_stride);
      tl_mark(tl, 0, tn, 2);
      float sig_acc = 0.f;
      if (e >= 0) {
        if (rh) {
          // xyz layer (act none): h -> A[:, 0:256), the operand of expert layer 0
          const int buf = (int)(li & 1);
          epi_load_bias(P.fblob + P.front[0].b_off, MW, sbias, buf, ec.et);
          epi_wait_acc(ctl, pp, buf);
          epi_hidden<false>(tmem_base + ec.lane_base + (uint32_t)buf * 256u, sbias + buf * 256, 4, a_base, ec, nullptr, ctl);
          if (P.skip_layer == 0) { epi_signal_chunk(ctl, 4, lane, ec.remote_a_ready); epi_signal_chunk(ctl, 5, lane, ec.remote_a_ready); }
          ++li;
        }
        for (int l = 0; l < NE; ++l, ++li) {
          const int buf = (int)(li & 1);
          const bool skip_here = (l == P.skip_layer);
          epi_load_bias((rh && skip_here) ? (P.fblob + P.o_b3x + (size_t)e * MW)
                                          : (P.fblob + P.expert[l].b_off + (size_t)e * P.expert_b_stride), MW, sbias, buf, ec.et);
          tl_mark(tl, 0, tn, 10 + l);
          epi_wait_acc(ctl, pp, buf);
          tl_mark(tl, 0, tn, 20 + l);
          if (rh && skip_here) write_cat();      // the skip layer's MMAs have retired: the PE(xyz) block is free
          const uint32_t tacc = tmem_base + ec.lane_base + (uint32_t)buf * 256u;
          const float* sb = sbias + buf * 256;
          if (l < NE - 1) {
            epi_hidden<true>(tacc, sb, 4, a_base, ec, (skip_here && !rh) ? hrow : nullptr, ctl);
            if (rh && l + 1 == P.skip_layer) {     // release the PE(xyz) chunks once more for the skip term
              epi_signal_chunk(ctl, 4, lane, ec.remote_a_ready);
              epi_signal_chunk(ctl, 5, lane, ec.remote_a_ready);
            }
          } else {
            // last expert layer (no activation) -> combine: y = bf16(gate * bf16(out)) -> ReLU -> A;
            // sigma head accumulated on the fly (nerf_moe.py:384-400)
            for (int c = 0; c < 4; ++c) {
              const int col0 = c * 64 + ec.cs * 16;
              uint32_t v[16];
              tmem_ld16(tacc + (uint32_t)col0, v);
              tmem_ld_wait();
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float f0 = bf16_round(__uint_as_float(v[j]) + sb[col0 + j]);
                float f1 = bf16_round(__uint_as_float(v[j + 1]) + sb[col0 + j + 1]);
                f0 = fmaxf(bf16_round(f0 * g), 0.f);
                f1 = fmaxf(bf16_round(f1 * g), 0.f);
                sig_acc = fmaf(f0, s_wsig[col0 + j], sig_acc);
                sig_acc = fmaf(f1, s_wsig[col0 + j + 1], sig_acc);
                pk[j / 2] = pack_bf16x2(f0, f1);
              }
              st_shared_v4(a_chunk_addr(a_base, row, col0 / 8), pk[0], pk[1], pk[2], pk[3]);
              st_shared_v4(a_chunk_addr(a_base, row, col0 / 8 + 1), pk[4], pk[5], pk[6], pk[7]);
              epi_signal_chunk(ctl, c, lane, ec.remote_a_ready);
            }
          }
          tl_mark(tl, 0, tn, 30 + l);
        }
      }
      sred[(0 * 4 + ec.cs) * 128 + row] = sig_acc;          // sigma partial of this column sub-slice
      // ---- layer "1" (act none): bf16 -> A[:, 0:256); then release the [dir | appearance] chunks ----
      {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.back[0].b_off, MW, sbias, buf, ec.et);
        epi_wait_acc(ctl, pp, buf);
        epi_hidden<false>(tmem_base + ec.lane_base + (uint32_t)buf * 256u, sbias + buf * 256, 4, a_base, ec, nullptr, ctl);
        const int nchunk2 = ((int)P.back[1].K16 + 63) / 64;
        for (int c = 4; c < nchunk2; ++c) epi_signal_chunk(ctl, c, lane, ec.remote_a_ready);
        tl_mark(tl, 0, tn, 50);
        ++li;
      }
      // ---- layer "2" (ReLU) + colour head partial dot products ----
      {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.back[1].b_off, H2, sbias, buf, ec.et);
        epi_wait_acc(ctl, pp, buf);
        const uint32_t tacc = tmem_base + ec.lane_base + (uint32_t)buf * 256u;
        const float* sb = sbias + buf * 256;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int c = 0; c < H2 / 64; ++c) {
          const int col0 = c * 64 + ec.cs * 16;
          uint32_t v[16];
          tmem_ld16(tacc + (uint32_t)col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = col0 + j;
            const float h2 = bf16_round(fmaxf(__uint_as_float(v[j]) + sb[k], 0.f));
            c0 = fmaf(h2, s_wcol[k], c0);
            c1 = fmaf(h2, s_wcol[H2 + k], c1);
            c2 = fmaf(h2, s_wcol[2 * H2 + k], c2);
          }
        }
        tc_fence_before();
        sred[(1 * 4 + ec.cs) * 128 + row] = c0;
        sred[(2 * 4 + ec.cs) * 128 + row] = c1;
        sred[(3 * 4 + ec.cs) * 128 + row] = c2;
        epi_bar_sync();
        if (ec.cs == 0 && valid) {
          auto rsum = [&](int v) { return sred[(v * 4 + 0) * 128 + row] + sred[(v * 4 + 1) * 128 + row] +
                                          sred[(v * 4 + 2) * 128 + row] + sred[(v * 4 + 3) * 128 + row]; };
          // sigma = softplus(bf16(W_sigma h + b) + noise - 1)
          // reference under cuda autocast: the Linear output, `sigma += noise` and `x - 1` (nerf.py:68) are bf16 tensor
          // ops (each rounds to bf16); only F.softplus runs in fp32 -- pinned by tests/golden/model_*_bf16cuda.npz
          float sr = bf16_round(rsum(0) + b_sig);
          if (noise) sr = bf16_round(sr + noise[(int64_t)sidx * io.n_stride]);
          const float tt_ = bf16_round(sr - 1.f);
          const float sigma = (tt_ > 20.f) ? tt_ : log1pf(expf(tt_));
          auto sg = [](float v) { return bf16_round(1.f / (1.f + expf(-bf16_round(v)))); };
          float4 o = make_float4(sg(rsum(1) + b_col0), sg(rsum(2) + b_col1), sg(rsum(3) + b_col2), sigma);
          if (io.ep) {      // expert-parallel: the row goes back to the rank that owns the sample (P2P store)
            const float* rec = x + (int64_t)sidx * io.x_stride;
            reinterpret_cast<float4*>(io.ret[__float_as_int(rec[10])])[__float_as_int(rec[9])] = o;
          } else {
            reinterpret_cast<float4*>(io.out)[sidx] = o;
          }
        }
        epi_bar_sync();                       // sred is reused by the next tile
        tl_mark(tl, 0, tn, 51);
        ++li;
      }
    }
    if (io.ep) __threadfence_system();      // this thread's P2P result stores are ordered before the CTA's count below
  }
  cta_teardown<CG>(tmem_base, warp);
  if (io.ep && threadIdx.x == 0) {          // last CTA to finish: raise flag B on every peer
    if (atomicAdd(io.done, 1) == (int)gridDim.x - 1) {
      *io.done = 0;
      __threadfence_system();
      for (int w = 0; w < io.world; ++w)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(io.flag_b[w]), "r"(io.epoch) : "memory");
    }
  }
}

#include "snb_tc_ts.cuh"
#include "snb_tc_wide.cuh"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
unsigned long long* g_timeline = nullptr;
int tc_timeline_read(unsigned long long* host, int n) {
  if (!g_timeline) return 0;
  if (n > 5 * TL_N) n = 5 * TL_N;
  cudaDeviceSynchronize();
  cudaMemcpy(host, g_timeline, (size_t)n * 8, cudaMemcpyDeviceToHost);
  return n;
}

size_t tc_workspace_bytes(const Model* m, int64_t S, double max_cf) {
  (void)max_cf;
  const int E = m->d.num_experts;
  if (S < 1) S = 1;
  int64_t max_rows = S + (int64_t)TILE * (E + 2);
  int64_t max_tiles = cdiv(S, TILE) + 2 * (E + 2);
  if (m->ep) {
    if (ep_max_rows(m->ep, S) > max_rows) max_rows = ep_max_rows(m->ep, S);
    if (ep_max_tiles(m->ep, S) > max_tiles) max_tiles = ep_max_tiles(m->ep, S);
  }
  size_t b = 0;
  b += align_up((size_t)S * MW * 2, 256);            // H
  b += align_up((size_t)S * E * 4, 256);             // gates
  b += 3 * align_up((size_t)S * 4, 256);             // idx, loc, gate
  b += align_up((size_t)max_rows * 4, 256);          // row2sample
  b += 3 * align_up((size_t)max_tiles * 4, 256);     // tile tables
  b += 4096;                                          // small scalars
  b += align_up((size_t)S * 4, 256);                  // packed routing words (snb_select.cuh)
  b += align_up((size_t)SEL_PM_RECORDS(S) * SEL_PM_STRIDE * 4, 256);   // partial column sums of the gates
  b += align_up((size_t)SEL_MAX_E * SEL_PM_STRIDE * 8, 256);            // per-CTA partial column sums of k_select
  b += route_workspace_bytes(S, E);
  return b + 4096;
}

// ---- one model_chunk as three separately enqueueable phases (so render_rays can software-pipeline
// them: the routing kernels of chunk c run on a side stream underneath launch #1 of chunk c+1) ----
struct TcChunk {
  TcParams Pf, Pb;
  const float* x;
  int64_t S;
  const float* noise;
  snb_route_opts o;
  float* out;
  int32_t* moe_idx;
  float* l_aux;
  float* dbg_gates;
  int32_t* dbg_loc;
  __nv_bfloat16* H;
  float* gates;
  int *idx, *loc;
  float* gate;
  TileTable tt;
  int *counts, *cap_dev;
  char* rws;
  size_t rbytes;
  uint32_t* wsel;           // packed routing words, written by k_front_ts (or k_pack_top1 from `gates`)
  float* pm;                // partial column sums of the gates
  int* hist0;               // [SEL_MAX_E][SEL_HBINS] level-0 key histogram per expert (+ the ticket of k_select)
  double* lpart;
  int npm;
  bool wide;                // wide kernels (width 512 / mip models)
  bool front_packed;        // launch #1 wrote wsel / pm itself
  bool select;              // routing = k_select (kept set only); false = full-order route_top1 (SNB_ROUTE_FULL=1)
  int64_t max_rows, max_tiles;
  PhaseEvents* pe;
  int set;                  // expert-parallel buffer set of this chunk
  int grid_cap;             // CTAs of the persistent kernels (<= SM count)
  int cg;                   // launch #1: 1 = independent CTAs, 2 = CTA pairs (tcgen05 cta_group::2)
  int cg_back;              // launch #2 (SNB_CG sets both, SNB_CG_FRONT / SNB_CG_BACK one of them)
};

static int tc_chunk_init(Model* m, TcChunk& c, const float* x, int64_t S, const float* sigma_noise,
                         const snb_route_opts* o, float* out, int32_t* moe_idx, float* l_aux, float* dbg_gates,
                         int32_t* dbg_loc, Arena& ws, cudaStream_t st, int set = 0) {
  SNB_REQUIRE(m->tc_blob, "tc_forward: weights were not packed");
  c.Pf = ((TcOwner*)m->tc_blob)->h.p;
  c.Pb = c.Pf;
  {
    // debug only: SNB_TIMELINE=1 records clock marks of CTA 0 (front: slots [0,2*TL_N), back: [2*TL_N, 4*TL_N))
    static unsigned long long* tl_buf = nullptr;
    static std::once_flag tl_once;
    std::call_once(tl_once, [] {
      if (getenv("SNB_TIMELINE")) { cudaMalloc((void**)&tl_buf, 5 * TL_N * 8); g_timeline = tl_buf; }
    });
    if (tl_buf) {
      cudaMemsetAsync(tl_buf, 0, 5 * TL_N * 8, st);
      c.Pf.tl = tl_buf;
      c.Pb.tl = tl_buf + 2 * TL_N;
    }
  }
  const int E = m->d.num_experts;
  c.x = x; c.S = S; c.noise = sigma_noise; c.o = *o; c.out = out; c.moe_idx = moe_idx; c.l_aux = l_aux;
  c.dbg_gates = dbg_gates; c.dbg_loc = dbg_loc;
  c.max_rows = S + (int64_t)TILE * (E + 2);
  c.max_tiles = cdiv(S, TILE) + 2 * (E + 2);
  c.set = 0;
  c.wide = tc_use_wide(m);
  if (c.wide) {
    SNB_REQUIRE(((TcOwner*)m->tc_blob)->wblob_w, "tc_forward: the wide weight images were not packed");
    SNB_REQUIRE(!m->ep, "expert parallelism is implemented for the width-256 NeRFMoE kernels only");
  }
  if (m->ep) {
    SNB_REQUIRE(c.Pf.recompute_h, "expert-parallel launch #2 recomputes h (SNB_GATHER_H is a single-GPU debug switch)");
    SNB_REQUIRE(!o->no_batch, "expert-parallel mode implements the capacity (batched) dispatch only");
    SNB_REQUIRE(!dbg_gates && !dbg_loc, "expert-parallel mode has no debug taps");
    if (ep_max_rows(m->ep, S) > c.max_rows) c.max_rows = ep_max_rows(m->ep, S);
    if (ep_max_tiles(m->ep, S) > c.max_tiles) c.max_tiles = ep_max_tiles(m->ep, S);
  }
  c.H = ws.take<__nv_bfloat16>((size_t)S * MW);
  c.gates = ws.take<float>((size_t)S * E);
  c.idx = ws.take<int>(S);
  c.loc = ws.take<int>(S);
  c.gate = ws.take<float>(S);
  c.tt.row2sample = ws.take<int>(c.max_rows);
  c.tt.tile_expert = ws.take<int>(c.max_tiles);
  c.tt.tile_row0 = ws.take<int>(c.max_tiles);
  c.tt.tile_rows = ws.take<int>(c.max_tiles);
  int* small = ws.take<int>(1024);
  c.wsel = ws.take<uint32_t>(S);
  c.pm = ws.take<float>((size_t)SEL_PM_RECORDS(S) * SEL_PM_STRIDE);
  // zero-between-uses scratch of the routing (level-0 histogram written by launch #1, barrier counters, level
  // histograms): owned by the model, zeroed once here and cleaned by k_select itself after every use
  if (!m->sel_zero) {
    SNB_CHECK_CUDA(cudaMalloc((void**)&m->sel_zero, (size_t)4 * SEL_ZERO_INTS * sizeof(int)));
    SNB_CHECK_CUDA(cudaMemsetAsync(m->sel_zero, 0, (size_t)4 * SEL_ZERO_INTS * sizeof(int), st));
  }
  c.hist0 = m->sel_zero + (size_t)(set & 3) * SEL_ZERO_INTS;
  c.lpart = ws.take<double>((size_t)SEL_MAX_E * SEL_PM_STRIDE);
  c.npm = 0;
  c.front_packed = false;
  c.select = !m->tune.route_full && E <= SEL_MAX_E;
  c.rbytes = route_workspace_bytes(S, E);
  c.rws = ws.take<char>(c.rbytes);
  if (!ws.ok) { set_error("tc_forward: workspace too small"); return SNB_EWORKSPACE; }
  c.counts = small; c.cap_dev = small + E;
  c.tt.seg_start = small + E + 1;
  c.tt.n_tiles = small + 2 * E + 4;
  c.tt.drop_counter = small + 2 * E + 5;
  static std::atomic<bool> attr_done_dev[64];     // function attributes are per device; setting them twice is harmless
  int dev = 0;
  cudaGetDevice(&dev);
  std::atomic<bool>& attr_done = attr_done_dev[dev & 63];
  if (!attr_done.load(std::memory_order_acquire)) {
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_front<12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_back<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_front<12, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_back<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_front_ts<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSM_TOTAL));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_back_ts<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSB_TOTAL));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_back_ts<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSB_TOTAL));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_front_wide<1, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes<1>()));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_back_wide<1, 12, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes<1>()));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_front_wide<2, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes<2>()));
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_back_wide<2, 12, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes<2>()));
    attr_done.store(true, std::memory_order_release);
  }
  c.cg = (m->tune.cta_group_front == 2 && !c.wide) ? 2 : 1;
  c.cg_back = (m->tune.cta_group_back == 2 && !c.wide) ? 2 : 1;
  c.pe = profile_next();
  c.grid_cap = m->sm_count;
  return SNB_OK;
}

static int tc_front(Model* m, TcChunk& c, cudaStream_t st) {
  const int n_front_tiles = (int)cdiv(c.S, TILE);
  int grid1 = n_front_tiles < c.grid_cap ? n_front_tiles : c.grid_cap;
  if (c.pe) cudaEventRecord(c.pe->e[0], st);
  if (c.wide) {
    const bool pack = c.select && (int64_t)cdiv(n_front_tiles, grid1) * TILE < 65536;
    float* gates_out = pack ? c.dbg_gates : c.gates;
    if (m->d.width == 512)
      k_front_wide<2, 12><<<grid1, THREADS, wide_smem_bytes<2>(), st>>>(c.Pf, c.x, c.S, gates_out, pack ? c.wsel : nullptr,
                                                                        c.hist0, pack ? c.pm : nullptr, pack ? c.moe_idx : nullptr);
    else
      k_front_wide<1, 12><<<grid1, THREADS, wide_smem_bytes<1>(), st>>>(c.Pf, c.x, c.S, gates_out, pack ? c.wsel : nullptr,
                                                                        c.hist0, pack ? c.pm : nullptr, pack ? c.moe_idx : nullptr);
    c.front_packed = pack;
    c.npm = 4 * n_front_tiles;
  } else if (c.cg == 2) {
    grid1 = (grid1 + 1) & ~1;
    if (grid1 > (c.grid_cap & ~1)) grid1 = c.grid_cap & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid1); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SM_TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SNB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_front<12, 2>, c.Pf, c.x, c.S, c.H, c.gates));
  } else {
    // SNB_TS_FRONT=0 / SNB_TS=0: shared-memory A operand (k_front)
    const bool ts_front = m->tune.ts != 0 && m->tune.ts_front != 0;
    if (ts_front && c.Pf.recompute_h && c.Pf.front[0].K16 <= TS_CAT_COLS) {
      // the [S,E] gates only leave the kernel when a caller taps them (or the full-order routing needs them)
      // the kernel keeps 16-bit row counters per CTA: a CTA must see fewer than 65536 rows (else k_pack_top1 does it)
      const bool pack = c.select && (int64_t)cdiv(n_front_tiles, grid1) * TILE < 65536;
      c.Pf.ab = m->tune.front_ab;
      float* gates_out = pack ? c.dbg_gates : c.gates;
      k_front_ts<12><<<grid1, THREADS, TSM_TOTAL, st>>>(c.Pf, c.x, c.S, gates_out, pack ? c.wsel : nullptr, c.hist0,
                                                         pack ? c.pm : nullptr, pack ? c.moe_idx : nullptr);
      c.front_packed = pack;
      c.npm = 4 * n_front_tiles;
    } else
      k_front<12, 1><<<grid1, THREADS, SM_TOTAL, st>>>(c.Pf, c.x, c.S, c.H, c.gates);
  }
  SNB_CHECK_LAUNCH("k_front");
  if (c.pe) cudaEventRecord(c.pe->e[1], st);
  return SNB_OK;
}

static int tc_route(Model* m, TcChunk& c, cudaStream_t st) {
  const int E = m->d.num_experts;
  if (c.pe) cudaEventRecord(c.pe->e[2], st);
  int rc;
  if (c.select) {
    // ONE launch: k_select (E CTAs) turns the packed words into the row table + tile plan of launch #2 (and, for
    // expert parallelism, the per-sample idx / loc / gate the record scatter reads)
    if (!c.front_packed) {
      if ((rc = route_pack_top1(c.gates, c.S, E, c.wsel, c.hist0, c.pm, &c.npm, st))) return rc;
      if (c.dbg_gates) SNB_CHECK_CUDA(cudaMemcpyAsync(c.dbg_gates, c.gates, sizeof(float) * c.S * E, cudaMemcpyDeviceToDevice, st));
    }
    SelectArgs a = {};
    a.w = c.wsel; a.zero = c.hist0; a.pm = c.pm; a.npm = c.npm; a.lpart = c.lpart; a.self_clean = 1;
    a.S = c.S; a.E = E; a.cf = c.o.capacity_factor; a.bpr = c.o.no_batch ? 0 : c.o.bpr; a.no_batch = c.o.no_batch;
    a.pair = c.cg_back == 2;
    a.counts = c.counts; a.cap_dev = c.cap_dev; a.l_aux = c.l_aux;
    a.moe_idx = c.front_packed ? nullptr : c.moe_idx;      // k_front_ts already wrote the expert ids
    a.loc = c.dbg_loc;
    if (m->ep) { a.idx = c.idx; a.loc = c.loc; a.gate = c.gate; }     // per-sample taps of the record scatter
    else a.tt = c.tt;                                                  // launch #2 decodes the gate from the packed word
    if (g_timeline) a.tl = g_timeline + 4 * TL_N;     // debug: phase marks of k_select's CTA 0
    if ((rc = route_select_launch(a, st))) return rc;
    if (m->ep) {
      rc = ep_dispatch_plan(m->ep, c.set, c.x, m->x_cols, c.gate, c.noise, c.idx, c.loc, c.counts, c.cap_dev, c.S,
                            capacity_of(c.S, E, c.o.capacity_factor), c.cg_back == 2, c.tt, st);
      if (rc) return rc;
    }
    if (c.pe) cudaEventRecord(c.pe->e[3], st);
    return SNB_OK;
  }
  rc = route_top1(c.gates, c.S, E, c.o.capacity_factor, c.o.no_batch ? 0 : c.o.bpr, c.idx, c.loc, c.gate, c.counts,
                  c.cap_dev, c.l_aux, c.rws, c.rbytes, st);
  if (rc) return rc;
  SNB_CHECK_CUDA(cudaMemsetAsync(c.tt.row2sample, 0xFF, (size_t)c.max_rows * sizeof(int), st));
  if (m->ep) {
    // experts live on other ranks: one kernel scatters this rank's records into the owners' memory (P2P stores +
    // per-peer flags), one plans the tiles over what the peers sent here
    rc = ep_dispatch_plan(m->ep, c.set, c.x, m->x_cols, c.gate, c.noise, c.idx, c.loc, c.counts, c.cap_dev, c.S,
                          capacity_of(c.S, E, c.o.capacity_factor), c.cg_back == 2, c.tt, st);
    if (rc) return rc;
  } else {
    k_tile_plan<<<1, 256, 0, st>>>(c.counts, c.cap_dev, E, c.o.no_batch, c.S, c.cg_back == 2, c.tt);
    SNB_CHECK_LAUNCH("k_tile_plan");
    k_scatter_rows<<<(unsigned)cdiv(c.S, 256), 256, 0, st>>>(c.idx, c.loc, c.cap_dev, E, c.o.no_batch, c.S, c.tt);
    SNB_CHECK_LAUNCH("k_scatter_rows");
  }
  if (c.moe_idx) SNB_CHECK_CUDA(cudaMemcpyAsync(c.moe_idx, c.idx, sizeof(int) * c.S, cudaMemcpyDeviceToDevice, st));
  if (c.dbg_gates) SNB_CHECK_CUDA(cudaMemcpyAsync(c.dbg_gates, c.gates, sizeof(float) * c.S * E, cudaMemcpyDeviceToDevice, st));
  if (c.dbg_loc) SNB_CHECK_CUDA(cudaMemcpyAsync(c.dbg_loc, c.loc, sizeof(int) * c.S, cudaMemcpyDeviceToDevice, st));
  if (c.pe) cudaEventRecord(c.pe->e[3], st);
  return SNB_OK;
}

static int tc_back(Model* m, TcChunk& c, cudaStream_t st, bool finish_inline = true) {
  int grid2 = (int)(c.max_tiles < c.grid_cap ? c.max_tiles : c.grid_cap);
  RowIO io = {};
  if (m->ep) {
    ep_row_io(m->ep, c.set, &io);
  } else {
    io.x = c.x; io.x_stride = m->x_cols;
    io.gate = c.gate; io.g_stride = 1;
    io.wsel = c.select ? c.wsel : nullptr;
    io.noise = c.noise; io.n_stride = 1;
    io.out = c.out;
  }
  if (c.pe) cudaEventRecord(c.pe->e[4], st);
  // launch #2 variants: SNB_TS=0 falls back to the shared-memory A operand (k_back); SNB_CG=2 runs CTA pairs
  const bool use_ts = m->tune.ts != 0;
  const bool ts_ok = use_ts && c.Pb.recompute_h && c.Pb.back[1].K16 > MW && c.Pb.back[1].K16 - MW <= TS_CAT_COLS &&
                     c.Pb.front[0].K16 <= TS_CAT_COLS;
  if (c.wide) {
    if (m->d.width == 512) k_back_wide<2, 12, 4><<<grid2, THREADS, wide_smem_bytes<2>(), st>>>(c.Pb, c.tt, io);
    else k_back_wide<1, 12, 4><<<grid2, THREADS, wide_smem_bytes<1>(), st>>>(c.Pb, c.tt, io);
  } else if (c.cg_back == 2) {
    grid2 &= ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid2); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = ts_ok ? TSB_TOTAL : SM_TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (ts_ok) SNB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_back_ts<4, 2>, c.Pb, c.tt, io));
    else SNB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_back<4, 2>, c.Pb, c.tt, io, (const __nv_bfloat16*)c.H));
  } else if (ts_ok) {
    k_back_ts<4, 1><<<grid2, THREADS, TSB_TOTAL, st>>>(c.Pb, c.tt, io);
  } else {
    k_back<4, 1><<<grid2, THREADS, SM_TOTAL, st>>>(c.Pb, c.tt, io, c.H);
  }
  SNB_CHECK_LAUNCH("k_back");
  if (m->ep && finish_inline) {
    int rc = ep_finish(m->ep, c.set, c.out, c.S, st);
    if (rc) return rc;
  }
  if (c.pe) cudaEventRecord(c.pe->e[5], st);
  return SNB_OK;
}

bool tc_ray_source_ok(const Model* m) {
  if (!m->tc_blob || m->ep || tc_use_wide(m)) return false;
  const TcParams& p = ((TcOwner*)m->tc_blob)->h.p;
  const snb_tuning& t = m->tune;
  const bool ts_off = !t.ts || !t.ts_front || t.cta_group_front == 2 || t.cta_group_back == 2 || t.no_ray_source;
  return !ts_off && p.recompute_h && p.front[0].K16 <= TS_CAT_COLS && p.back[1].K16 > (uint32_t)MW &&
         p.back[1].K16 - MW <= TS_CAT_COLS;
}

int tc_forward(Model* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* o, float* out,
               int32_t* moe_idx, float* l_aux, float* dbg_gates, int32_t* dbg_loc, Arena& ws, cudaStream_t st) {
  TcChunk c;
  int rc = tc_chunk_init(m, c, x, S, sigma_noise, o, out, moe_idx, l_aux, dbg_gates, dbg_loc, ws, st);
  if (rc) return rc;
  if ((rc = tc_front(m, c, st))) return rc;
  if ((rc = tc_route(m, c, st))) return rc;
  return tc_back(m, c, st);
}

// All model chunks of one render pass, software-pipelined over the caller's stream `st` and the model's
// side stream: st = front(0) front(1) back(0) front(2) back(1) ... ; side = route(0) route(1) ...
// Two workspace sets alternate between consecutive chunks.
int tc_forward_chunks(Model* m, const float* x, int64_t B, int64_t chunk, const snb_route_opts* o, float* out,
                      int32_t* moe_idx, float* l_aux, void* ws_base, size_t ws_stride, int nsets, cudaStream_t st,
                      const float* sigma_noise, const RaySource* rs) {
  SNB_REQUIRE(x || rs, "tc_forward_chunks: neither rows nor a ray source");
  SNB_REQUIRE(!rs || tc_ray_source_ok(m), "tc_forward_chunks: the kernels selected for this model read materialised rows");
  if (B <= 0) return SNB_OK;
  constexpr int MAXSETS = 4;
  static_assert(MAXSETS <= EP_SETS, "one expert-parallel buffer set per workspace set");
  if (!m->side_stream) {
    int prio_lo = 0, prio_hi = 0;
    SNB_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    SNB_CHECK_CUDA(cudaStreamCreateWithPriority(&m->side_stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < MAXSETS; ++i) {
      SNB_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_front[i], cudaEventDisableTiming));
      SNB_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_route[i], cudaEventDisableTiming));
    }
  }
  // tuning / A-B switches (read once): SNB_NO_OVERLAP routes on the caller's stream; SNB_PIPE_DEPTH = how many
  // launch #1 run ahead of launch #2 (1..3); SNB_ROUTE_SMS = SMs left free for the routing kernels
  const bool no_overlap = m->tune.no_overlap != 0;
  // defaults from the sweeps in profiles/: local experts depth 2 / 28 SMs (r1n: 12/20/28/36 SMs); expert-parallel
  // depth 3 / 36 SMs (r1q/r1r: the routing stage also scatters records to the peers and waits for theirs)
  const int depth_env = m->tune.pipe_depth, route_env = m->tune.route_sms;
  const bool back_full = !m->tune.back_partition;   // launch #2 keeps every SM (tile rounds!)
  // local experts + select routing: k_select is E CTAs of 1024 threads (one per SM); expert-parallel: the routing
  // stage also scatters records to the peers and waits for theirs (r1q/r1r)
  const bool route_full_env = m->tune.route_full != 0;
  const int sel_sms = 8;      // k_select: SEL_P = 8 CTAs of 1024 threads, one per SM
  // (r2t, N=2 expert-parallel: 8 / 16 / 36 SMs -> 729 / 726 / 682 M samples/s: the record scatter and the plan kernel are
  // short and follow k_select on the same SMs)
  const int route_sms = route_env >= 0 ? route_env : (route_full_env ? (m->ep ? 36 : 28) : sel_sms);
  int D = depth_env >= 1 ? depth_env : (m->ep ? 3 : 2);
  if (D > nsets - 1) D = nsets - 1;
  if (D > MAXSETS - 1) D = MAXSETS - 1;
  const int NS = D + 1;
  cudaStream_t sr = no_overlap ? st : m->side_stream;
  const int grid_cap = (!no_overlap && route_sms > 0 && route_sms < m->sm_count / 2) ? (m->sm_count - route_sms) & ~1 : m->sm_count;
  TcChunk cc[MAXSETS];
  int ci = 0;
  int rc;
  // expert parallelism: the wait for the peers' result rows of chunk c (flag B of every peer) and the copy of the rows
  // into `out` run on a third stream, so a slower peer never stalls this rank's launches; the main stream only joins
  // at the end of the pass (the routing stream joins before the buffer set is reused, NS chunks later)
  const bool fin_async = m->ep != nullptr;
  bool fin_pending[MAXSETS] = {false, false, false, false};
  if (fin_async && !m->fin_stream) {
    int prio_lo = 0, prio_hi = 0;
    SNB_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    SNB_CHECK_CUDA(cudaStreamCreateWithPriority(&m->fin_stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < MAXSETS; ++i) {
      SNB_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_back[i], cudaEventDisableTiming));
      SNB_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_fin[i], cudaEventDisableTiming));
    }
  }
  auto back_of = [&](int kb) -> int {
    SNB_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_route[kb], 0));
    int r = tc_back(m, cc[kb], st, !fin_async);
    if (r) return r;
    if (fin_async) {
      SNB_CHECK_CUDA(cudaEventRecord(m->ev_back[kb], st));
      SNB_CHECK_CUDA(cudaStreamWaitEvent(m->fin_stream, m->ev_back[kb], 0));
      if ((r = ep_finish(m->ep, cc[kb].set, cc[kb].out, cc[kb].S, m->fin_stream))) return r;
      SNB_CHECK_CUDA(cudaEventRecord(m->ev_fin[kb], m->fin_stream));
      fin_pending[kb] = true;
    }
    return SNB_OK;
  };
  for (int64_t i = 0; i < B; i += chunk, ++ci) {
    const int64_t rows = (B - i < chunk) ? (B - i) : chunk;
    const int k = ci % NS;
    Arena a((char*)ws_base + (size_t)k * ws_stride, ws_stride);
    if ((rc = tc_chunk_init(m, cc[k], x ? x + i * m->x_cols : nullptr, rows, sigma_noise ? sigma_noise + i : nullptr, o, out + i * 4,
                            moe_idx ? moe_idx + i : nullptr,
                            l_aux ? l_aux + ci : nullptr, nullptr, nullptr, a, st, k)))
      return rc;
    cc[k].set = k;
    cc[k].grid_cap = grid_cap;
    if (rs) {
      for (TcParams* P : {&cc[k].Pf, &cc[k].Pb}) {
        P->ray_src = rs->rays; P->ray_z = rs->z; P->ray_img = rs->img; P->ray_sn = rs->Sn; P->ray_s0 = i;
      }
      cc[k].x = nullptr;
    }
    if ((rc = tc_front(m, cc[k], st))) return rc;
    if (back_full) cc[k].grid_cap = m->sm_count;
    SNB_CHECK_CUDA(cudaEventRecord(m->ev_front[k], st));
    SNB_CHECK_CUDA(cudaStreamWaitEvent(sr, m->ev_front[k], 0));
    if (fin_pending[k]) {
      // the expert-parallel buffers of this set are about to be reused: this rank may only send the records of the
      // new chunk after every peer finished the old one (flag B seen => the peers no longer read rx[k]) and the old
      // result rows were copied out of ret[k].  Only the routing stream waits; launch #1 above does not.
      SNB_CHECK_CUDA(cudaStreamWaitEvent(sr, m->ev_fin[k], 0));
      fin_pending[k] = false;
    }
    if ((rc = tc_route(m, cc[k], sr))) return rc;
    SNB_CHECK_CUDA(cudaEventRecord(m->ev_route[k], sr));
    if (ci >= D && (rc = back_of((ci - D) % NS))) return rc;
  }
  for (int j = (ci - D > 0 ? ci - D : 0); j < ci; ++j)
    if ((rc = back_of(j % NS))) return rc;
  for (int k = 0; k < MAXSETS; ++k)
    if (fin_pending[k]) SNB_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_fin[k], 0));
  return SNB_OK;
}

}  // namespace snb

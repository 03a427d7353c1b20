// SNB_PREC_FP32 path: the NeRFMoE forward of one model_chunk in plain fp32 on the CUDA cores.
// This is the un-fused, reference-shaped pipeline (encode -> Linear ... -> dispatch buffer ->
// per-expert Linear stack -> combine -> heads).  It exists for (1) BASELINE.json configs[0]
// (fp32 parity with the reference's CPU path to <=1e-3, in practice ~1e-6) and (2) as the
// on-device cross-check of the fused tcgen05 path at full chunk sizes.  It is a CUDA path, not
// a CPU fallback: the library refuses to run without a GPU.
//
// Reference: models/nerf_moe.py:320-455, modules/tutel_moe_ext/tutel_moe_layer_nobatch.py:98-235,
// 887-924, tutel_fast_dispatch.py:15-63.
#include "snb_common.cuh"

namespace snb {

// ------------------------------------------------------------------------------------------
// k_encode: positional encodings (models/nerf.py:21-26 / 28-56) + appearance lookup.
//   pe  [S, xyz_in]           = Embedding(F)(xyz)  or MipEmbedder(F)(mean, cov)
//   cat [S, ld_cat] cols [M, M+dir_in)           = Embedding(Fd)(dir)
//                   cols [M+dir_in, M+dir_in+A)  = embedding_a[int(x[:, -1])]
// ------------------------------------------------------------------------------------------
__global__ void k_encode(const float* __restrict__ x, int64_t S, int x_cols, int mip, int F, int Fd, int A,
                         int appearance_count, const float* __restrict__ emb_a, float* __restrict__ pe,
                         int ld_pe, float* __restrict__ cat, int ld_cat, int cat_off) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (s >= S) return;
  const float* xr = x + s * x_cols;
  const int xd = mip ? 6 : 3;
  const int n_pe = 3 + 6 * F, n_dir = 3 + 6 * Fd;
  for (int c = threadIdx.x; c < n_pe + n_dir + A; c += blockDim.x) {
    if (c < n_pe) {
      float v;
      if (c < 3) {
        v = xr[c];
      } else {
        int k = (c - 3) / 6, r = (c - 3) % 6, ax = r % 3;
        float f = (float)(1 << k);
        float a = f * xr[ax];
        float t = (r < 3) ? sinf(a) : cosf(a);
        if (mip) t *= expf(-0.5f * (f * f) * xr[3 + ax]);
        v = t;
      }
      pe[s * ld_pe + c] = v;
    } else if (c < n_pe + n_dir) {
      int cc = c - n_pe;
      float v;
      if (cc < 3) {
        v = xr[xd + cc];
      } else {
        int k = (cc - 3) / 6, r = (cc - 3) % 6, ax = r % 3;
        float a = (float)(1 << k) * xr[xd + ax];
        v = (r < 3) ? sinf(a) : cosf(a);
      }
      cat[s * ld_cat + cat_off + cc] = v;
    } else {
      int cc = c - n_pe - n_dir;
      int ai = (int)xr[x_cols - 1];                 // x[:, -1].long()  (nerf_moe.py:427)
      ai = min(max(ai, 0), appearance_count - 1);
      cat[s * ld_cat + cat_off + n_dir + cc] = emb_a[(int64_t)ai * A + cc];
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_linear: C[r, n] = act( sum_k A[r,k] * W[n,k] + b[n] (+ R[r,n]) (+ noise[r]) )
// 64x64 tile, BK=16, 256 threads, 4x4 micro-tile.  grid.z = expert (batched mode):
//   A += ebase[z]*lda, rows = erows[z], W += z*N*K, b += z*N, C/R += ebase[z]*ldc.
// ------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(256) k_linear(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                const float* __restrict__ bias, const float* __restrict__ R, int ldr,
                                                float* __restrict__ C, int ldc, int64_t rows, int N, int K,
                                                const int* __restrict__ ebase, const int* __restrict__ erows,
                                                const float* __restrict__ noise) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float sA[BK][BM + 4];
  __shared__ float sW[BK][BN + 4];
  if (ebase) {
    const int z = blockIdx.z;
    const int64_t base = ebase[z];
    rows = erows[z];
    A += base * lda;
    C += base * ldc;
    if (R) R += base * ldr;
    W += (int64_t)z * N * K;
    if (bias) bias += (int64_t)z * N;
  }
  const int64_t r0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  if (r0 >= rows) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    // load A tile: 64 rows x 16 k  (256 threads x 4 elements)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = threadIdx.x + i * 256;
      int rr = e >> 4, kk = e & 15;
      int64_t r = r0 + rr;
      int k = k0 + kk;
      sA[kk][rr] = (r < rows && k < K) ? A[r * lda + k] : 0.f;
      int n = n0 + rr;
      sW[kk][rr] = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; w[i] = sW[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = r0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (R) v += R[r * ldr + n];
      if (noise) v += noise[r];
      if (ACT == ACT_RELU) v = fmaxf(v, 0.f);
      if (ACT == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      if (ACT == ACT_SOFTPLUS_SHIFT) {           // F.softplus(x - 1, beta=1, threshold=20)
        float t = v - 1.f;
        v = (t > 20.f) ? t : log1pf(expf(t));
      }
      C[r * ldc + n] = v;
    }
  }
}

static int launch_linear(int act, const float* A, int lda, const float* W, const float* b, const float* R, int ldr,
                         float* C, int ldc, int64_t rows, int N, int K, int nz, const int* ebase, const int* erows,
                         const float* noise, cudaStream_t st) {
  if (rows <= 0) return SNB_OK;
  dim3 grid((unsigned)cdiv(rows, 64), (unsigned)cdiv(N, 64), (unsigned)nz);
  switch (act) {
    case ACT_NONE: k_linear<ACT_NONE><<<grid, 256, 0, st>>>(A, lda, W, b, R, ldr, C, ldc, rows, N, K, ebase, erows, noise); break;
    case ACT_RELU: k_linear<ACT_RELU><<<grid, 256, 0, st>>>(A, lda, W, b, R, ldr, C, ldc, rows, N, K, ebase, erows, noise); break;
    case ACT_SIGMOID: k_linear<ACT_SIGMOID><<<grid, 256, 0, st>>>(A, lda, W, b, R, ldr, C, ldc, rows, N, K, ebase, erows, noise); break;
    default: k_linear<ACT_SOFTPLUS_SHIFT><<<grid, 256, 0, st>>>(A, lda, W, b, R, ldr, C, ldc, rows, N, K, ebase, erows, noise); break;
  }
  SNB_CHECK_LAUNCH("k_linear");
  return SNB_OK;
}

// the same tiles for callers in other translation units (snb_bg.cu)
int linear_launch(int act, const float* A, int lda, const float* W, const float* b, const float* R, int ldr, float* C,
                  int ldc, int64_t rows, int N, int K, const float* noise, cudaStream_t st) {
  return launch_linear(act, A, lda, W, b, R, ldr, C, ldc, rows, N, K, 1, nullptr, nullptr, noise, st);
}

// ------------------------------------------------------------------------------------------
// k_ln_gate: gate_input = LayerNorm(g) (eps 1e-5, biased var); logits = wg @ gate_input;
// gates = softmax(logits).  One warp per row.  (nerf_moe.py:372; tutel_moe_layer_nobatch.py:105-126)
// ln_w == nullptr: g already is the gate input (the standalone MoE-layer operator, snb_moe_layer_forward).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ln_gate(const float* __restrict__ g, int64_t S, int M, int E,
                                                 const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                 const float* __restrict__ wg, float* __restrict__ gates) {
  extern __shared__ float s_wg[];  // [E][M]
  for (int i = threadIdx.x; i < E * M; i += blockDim.x) s_wg[i] = wg[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
  if (s >= S) return;
  const float* row = g + s * M;
  const bool ln = (ln_w != nullptr);
  float mean = 0.f, rstd = 1.f;
  if (ln) {
    float sum = 0.f;
    for (int k = lane; k < M; k += 32) sum += row[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean = sum / (float)M;
    float var = 0.f;
    for (int k = lane; k < M; k += 32) { float d = row[k] - mean; var += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    rstd = rsqrtf(var / (float)M + 1e-5f);
  }
  float mx = -INFINITY;
  for (int e0 = 0; e0 < E; e0 += 32) {
    float keep = -INFINITY;
    for (int e = e0; e < min(E, e0 + 32); ++e) {
      float acc = 0.f;
      for (int k = lane; k < M; k += 32) {
        const float v = ln ? (row[k] - mean) * rstd * ln_w[k] + ln_b[k] : row[k];
        acc = fmaf(v, s_wg[e * M + k], acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == e - e0) keep = acc;
      mx = fmaxf(mx, acc);
    }
    // stash the logits of this group in the output row; normalised below
    if (e0 + lane < E) gates[s * E + e0 + lane] = keep;
  }
  __syncwarp();
  float den = 0.f;
  for (int e = lane; e < E; e += 32) den += expf(gates[s * E + e] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  for (int e = lane; e < E; e += 32) gates[s * E + e] = expf(gates[s * E + e] - mx) / den;
}

// ------------------------------------------------------------------------------------------
// Dispatch / combine (Tutel K4/K5 and in-tree K1/K2 addressing; see public header).
// ------------------------------------------------------------------------------------------
__global__ void k_dispatch(const float* __restrict__ x, const int* __restrict__ idx, const int* __restrict__ loc,
                           const int* __restrict__ begin, const int* __restrict__ cap_dev, int cap_host, int64_t S,
                           int H, int64_t rows_out, float* __restrict__ out) {
  const int cap = cap_dev ? *cap_dev : cap_host;
  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    const int e = idx[s], l = loc[s];
    if (e < 0) continue;
    int64_t row;
    if (begin) row = (int64_t)begin[e] + l;
    else { if (l >= cap) continue; row = (int64_t)e * cap + l; }
    if (row >= rows_out) continue;
    for (int j = threadIdx.x; j < H; j += blockDim.x) out[row * H + j] = x[s * H + j];
  }
}

template <bool RELU>
__global__ void k_combine(const float* __restrict__ buf, const int* __restrict__ idx, const int* __restrict__ loc,
                          const int* __restrict__ begin, const float* __restrict__ gate,
                          const int* __restrict__ cap_dev, int cap_host, int64_t S, int H, int64_t rows_buf,
                          float* __restrict__ y) {
  const int cap = cap_dev ? *cap_dev : cap_host;
  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    const int e = idx[s], l = loc[s];
    int64_t row = -1;
    if (e >= 0) {
      if (begin) row = (int64_t)begin[e] + l;
      else if (l < cap) row = (int64_t)e * cap + l;
      if (row >= rows_buf) row = -1;
    }
    const float g = gate ? gate[s] : 1.f;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
      float v = (row >= 0) ? g * buf[row * H + j] : 0.f;
      if (RELU) v = fmaxf(v, 0.f);
      y[s * H + j] = v;
    }
  }
}

// ebase/erows for the expert GEMMs: padded (e*cap, min(count,cap)) or contiguous (begin, count)
__global__ void k_expert_ranges(const int* __restrict__ counts, const int* __restrict__ cap_dev, int E, int no_batch,
                                int* __restrict__ ebase, int* __restrict__ erows, int* __restrict__ begin) {
  if (threadIdx.x != 0) return;
  const int cap = *cap_dev;
  int run = 0;
  for (int e = 0; e < E; ++e) {
    const int c = counts[e];
    begin[e] = run;
    if (no_batch) { ebase[e] = run; erows[e] = c; }
    else { ebase[e] = e * cap; erows[e] = min(c, cap); }
    run += c;
  }
}

// ------------------------------------------------------------------------------------------
int snb_dispatch_impl(const float* x, const int* idx, const int* loc, const int* begin, const int* cap_dev,
                      int cap_host, int64_t S, int H, int64_t rows_out, float* out, bool zero, cudaStream_t st) {
  if (zero) SNB_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)rows_out * H * sizeof(float), st));
  if (S == 0) return SNB_OK;
  int grid = (int)(S < 4096 ? S : 4096);
  k_dispatch<<<grid, 128, 0, st>>>(x, idx, loc, begin, cap_dev, cap_host, S, H, rows_out, out);
  SNB_CHECK_LAUNCH("k_dispatch");
  return SNB_OK;
}

int snb_combine_impl(const float* buf, const int* idx, const int* loc, const int* begin, const float* gate,
                     const int* cap_dev, int cap_host, int64_t S, int H, int64_t rows_buf, float* y, bool relu,
                     cudaStream_t st) {
  if (S == 0) return SNB_OK;
  int grid = (int)(S < 4096 ? S : 4096);
  if (relu) k_combine<true><<<grid, 128, 0, st>>>(buf, idx, loc, begin, gate, cap_dev, cap_host, S, H, rows_buf, y);
  else k_combine<false><<<grid, 128, 0, st>>>(buf, idx, loc, begin, gate, cap_dev, cap_host, S, H, rows_buf, y);
  SNB_CHECK_LAUNCH("k_combine");
  return SNB_OK;
}

size_t fp32_workspace_bytes(const Model* m, int64_t S, double max_cf) {
  const int M = m->d.width, E = m->d.num_experts;
  if (S < 1) S = 1;
  int64_t cap = capacity_of(S, E, max_cf > 1.0 ? max_cf : 1.0);
  int64_t rows = (int64_t)E * cap;
  if (rows < S) rows = S;
  size_t b = 0;
  auto add = [&](size_t n) { b += align_up(n * sizeof(float), 256); };
  add((size_t)S * m->xyz_in);       // pe
  add((size_t)S * M);               // h
  add((size_t)S * M);               // t0
  add((size_t)S * M);               // t1
  add((size_t)S * E);               // gates
  add((size_t)S * 3);               // idx, loc, gate
  add((size_t)rows * M * 3);        // buf0, buf1, bufx
  add((size_t)S * m->cat_in);       // cat
  add((size_t)S * m->d.hidden2);    // h2
  add((size_t)S);                   // sigma
  add((size_t)S * 3);               // rgb
  b += 4096;
  return b + route_workspace_bytes(S, E);
}

// ------------------------------------------------------------------------------------------
// Standalone MoE-layer operator (fp32): MOELayer.forward -> TopKGate.apply_on_expert_fn[_nobatch] with an external
// gate input (tutel_moe_layer_nobatch.py:733-797, 98-352): gates = softmax(gate_input @ wg^T) -> routing ->
// dispatch -> expert stack -> combine (no activation: the ReLU of nerf_moe.py:384 belongs to the caller).
// ------------------------------------------------------------------------------------------
size_t fp32_moe_layer_workspace_bytes(const Model* m, int64_t S, double max_cf) {
  const int M = m->d.width, E = m->d.num_experts;
  if (S < 1) S = 1;
  int64_t cap = capacity_of(S, E, max_cf > 1.0 ? max_cf : 1.0);
  int64_t rows = (int64_t)E * cap;
  if (rows < S) rows = S;
  size_t b = 0;
  auto add = [&](size_t n) { b += align_up(n * sizeof(float), 256); };
  add((size_t)S * E);               // gates
  add((size_t)S); add((size_t)S); add((size_t)S);   // idx, loc, gate
  add((size_t)rows * M); add((size_t)rows * M); add((size_t)rows * M);   // buf0, buf1, bufx
  b += 4096 + 4096;
  return b + route_workspace_bytes(S, E);
}

int fp32_moe_layer(Model* m, const float* input, const float* gate_input, int64_t S, const snb_route_opts* o, float* y,
                   int32_t* moe_idx, float* l_aux, Arena& ws, cudaStream_t st) {
  const int M = m->d.width, E = m->d.num_experts, L = m->d.expert_layers;
  if (S == 0) return SNB_OK;
  const double cf = o->capacity_factor;
  const int64_t cap_host = capacity_of(S, E, cf);
  int64_t rows = o->no_batch ? S : (int64_t)E * cap_host;
  if (rows < 1) rows = 1;
  float* gates = ws.take<float>((size_t)S * E);
  int* idx = ws.take<int>(S);
  int* loc = ws.take<int>(S);
  float* gate = ws.take<float>(S);
  float* buf0 = ws.take<float>((size_t)rows * M);
  float* buf1 = ws.take<float>((size_t)rows * M);
  float* bufx = ws.take<float>((size_t)rows * M);
  int* small = ws.take<int>(1024);
  const size_t rbytes = route_workspace_bytes(S, E);
  char* rws = ws.take<char>(rbytes);
  if (!ws.ok) { set_error("snb_moe_layer_forward: workspace too small"); return SNB_EWORKSPACE; }
  SNB_REQUIRE(4 * E + 8 <= 1024, "too many experts");
  int *counts = small, *cap_dev = small + E, *ebase = small + E + 1, *erows = ebase + E, *begin = erows + E;
  {
    const size_t smem = (size_t)E * M * sizeof(float);
    SNB_REQUIRE(smem <= 200 * 1024, "wg too large for shared memory");
    if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_ln_gate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ln_gate<<<(unsigned)cdiv(S, 8), 256, smem, st>>>(gate_input, S, M, E, nullptr, nullptr, m->wg, gates);
    SNB_CHECK_LAUNCH("k_ln_gate");
  }
  int rc;
  if ((rc = route_top1(gates, S, E, cf, o->no_batch ? 0 : o->bpr, idx, loc, gate, counts, cap_dev, l_aux, rws, rbytes, st))) return rc;
  k_expert_ranges<<<1, 32, 0, st>>>(counts, cap_dev, E, o->no_batch, ebase, erows, begin);
  SNB_CHECK_LAUNCH("k_expert_ranges");
  if (moe_idx) SNB_CHECK_CUDA(cudaMemcpyAsync(moe_idx, idx, sizeof(int) * S, cudaMemcpyDeviceToDevice, st));
  if ((rc = snb_dispatch_impl(input, idx, loc, o->no_batch ? begin : nullptr, cap_dev, 0, S, M, rows, bufx, false, st))) return rc;
  const float* in = bufx;
  float* outb = buf0;
  const int64_t max_rows = o->no_batch ? S : cap_host;
  for (int j = 0; j < L; ++j) {
    const bool skip = (j == m->d.skip_layer);
    if (max_rows > 0) {
      dim3 grid((unsigned)cdiv(max_rows, 64), (unsigned)cdiv(M, 64), (unsigned)E);
      if (j < L - 1)
        k_linear<ACT_RELU><<<grid, 256, 0, st>>>(in, M, m->exp_w[j], m->exp_b[j], skip ? bufx : nullptr, M, outb, M, 0, M, M, ebase, erows, nullptr);
      else
        k_linear<ACT_NONE><<<grid, 256, 0, st>>>(in, M, m->exp_w[j], m->exp_b[j], skip ? bufx : nullptr, M, outb, M, 0, M, M, ebase, erows, nullptr);
      SNB_CHECK_LAUNCH("k_linear(expert)");
    }
    in = outb;
    outb = (outb == buf0) ? buf1 : buf0;
  }
  return snb_combine_impl(in, idx, loc, o->no_batch ? begin : nullptr, gate, cap_dev, 0, S, M, rows, y, false, st);
}

__global__ void k_pack_out(const float* __restrict__ rgb, const float* __restrict__ sigma, int64_t S,
                           float* __restrict__ out) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float4 v = make_float4(rgb[s * 3 + 0], rgb[s * 3 + 1], rgb[s * 3 + 2], sigma[s]);
  reinterpret_cast<float4*>(out)[s] = v;
}

int fp32_forward(Model* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* o, float* out,
                 int32_t* moe_idx, float* l_aux, float* dbg_gates, int32_t* dbg_loc, Arena& ws, cudaStream_t st) {
  const int M = m->d.width, E = m->d.num_experts, L = m->d.expert_layers, H2 = m->d.hidden2;
  if (S == 0) return SNB_OK;
  const double cf = o->capacity_factor;
  int64_t cap_host = capacity_of(S, E, cf);
  int64_t rows = o->no_batch ? S : (int64_t)E * cap_host;
  if (rows < 1) rows = 1;
  float* pe = ws.take<float>((size_t)S * m->xyz_in);
  float* h = ws.take<float>((size_t)S * M);
  float* t0 = ws.take<float>((size_t)S * M);
  float* t1 = ws.take<float>((size_t)S * M);
  float* gates = ws.take<float>((size_t)S * E);
  int* idx = ws.take<int>(S);
  int* loc = ws.take<int>(S);
  float* gate = ws.take<float>(S);
  float* buf0 = ws.take<float>((size_t)rows * M);
  float* buf1 = ws.take<float>((size_t)rows * M);
  float* bufx = ws.take<float>((size_t)rows * M);
  float* cat = ws.take<float>((size_t)S * m->cat_in);
  float* h2 = ws.take<float>((size_t)S * H2);
  float* sigma = ws.take<float>(S);
  float* rgb = ws.take<float>((size_t)S * 3);
  int* small = ws.take<int>(1024);  // counts[E], capacity, ebase[E], erows[E], begin[E]
  size_t rbytes = route_workspace_bytes(S, E);
  char* rws = ws.take<char>(rbytes);
  if (!ws.ok) { set_error("fp32_forward: workspace too small"); return SNB_EWORKSPACE; }
  SNB_REQUIRE(4 * E + 8 <= 1024, "too many experts");
  int *counts = small, *cap_dev = small + E, *ebase = small + E + 1, *erows = ebase + E, *begin = erows + E;

  // 1. encodings  (nerf_moe.py:330, 424-429)
  {
    dim3 blk(32, 8);
    k_encode<<<(unsigned)cdiv(S, 8), blk, 0, st>>>(x, S, m->x_cols, m->d.mip, m->d.pos_xyz_freqs, m->d.pos_dir_freqs,
                                                   m->d.appearance_dim, m->d.appearance_count, m->emb_a, pe,
                                                   m->xyz_in, cat, m->cat_in, M);
    SNB_CHECK_LAUNCH("k_encode");
  }
  int rc;
  // 2. xyz linear (act none), external gate MLP  (nerf_moe.py:333, 348)
  if ((rc = launch_linear(ACT_NONE, pe, m->xyz_in, m->xyz_w, m->xyz_b, nullptr, 0, h, M, S, M, m->xyz_in, 1, nullptr, nullptr, nullptr, st))) return rc;
  const float* gin = h;
  float* gout = t0;
  for (int i = 0; i < m->d.gate_layers; ++i) {
    int act = (i < m->d.gate_layers - 1) ? ACT_RELU : ACT_NONE;
    if ((rc = launch_linear(act, gin, M, m->gate_w[i], m->gate_b[i], nullptr, 0, gout, M, S, M, M, 1, nullptr, nullptr, nullptr, st))) return rc;
    gin = gout;
    gout = (gout == t0) ? t1 : t0;
  }
  // 3. LayerNorm + fp32 gate + softmax  (nerf_moe.py:372; tutel_moe_layer_nobatch.py:105-126)
  {
    size_t smem = (size_t)E * M * sizeof(float);
    SNB_REQUIRE(smem <= 200 * 1024, "wg too large for shared memory");
    if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_ln_gate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ln_gate<<<(unsigned)cdiv(S, 8), 256, smem, st>>>(gin, S, M, E, m->ln_w, m->ln_b, m->wg, gates);
    SNB_CHECK_LAUNCH("k_ln_gate");
  }
  // 4. routing  (tutel_fast_dispatch.py:176-217)
  if ((rc = route_top1(gates, S, E, cf, o->no_batch ? 0 : o->bpr, idx, loc, gate, counts, cap_dev, l_aux, rws, rbytes, st))) return rc;
  k_expert_ranges<<<1, 32, 0, st>>>(counts, cap_dev, E, o->no_batch, ebase, erows, begin);
  SNB_CHECK_LAUNCH("k_expert_ranges");
  if (moe_idx) SNB_CHECK_CUDA(cudaMemcpyAsync(moe_idx, idx, sizeof(int) * S, cudaMemcpyDeviceToDevice, st));
  if (dbg_gates) SNB_CHECK_CUDA(cudaMemcpyAsync(dbg_gates, gates, sizeof(float) * S * E, cudaMemcpyDeviceToDevice, st));
  if (dbg_loc) SNB_CHECK_CUDA(cudaMemcpyAsync(dbg_loc, loc, sizeof(int) * S, cudaMemcpyDeviceToDevice, st));
  // 5. dispatch (K4 / K1), expert stack (ExpertMLP.forward :887-924), combine (K5 / K2) + ReLU (:384-386)
  if ((rc = snb_dispatch_impl(h, idx, loc, o->no_batch ? begin : nullptr, cap_dev, 0, S, M, rows, bufx, false, st))) return rc;
  {
    const float* in = bufx;
    float* outb = buf0;
    // rows per expert are bounded by max(cap, S) on the host side; blocks beyond erows[e] exit early
    int64_t max_rows = o->no_batch ? S : cap_host;
    for (int j = 0; j < L; ++j) {
      const bool skip = (j == m->d.skip_layer);
      const int act = (j < L - 1) ? ACT_RELU : ACT_NONE;
      if (max_rows > 0) {
        dim3 grid((unsigned)cdiv(max_rows, 64), (unsigned)cdiv(M, 64), (unsigned)E);
        if (act == ACT_RELU)
          k_linear<ACT_RELU><<<grid, 256, 0, st>>>(in, M, m->exp_w[j], m->exp_b[j], skip ? bufx : nullptr, M, outb, M, 0, M, M, ebase, erows, nullptr);
        else
          k_linear<ACT_NONE><<<grid, 256, 0, st>>>(in, M, m->exp_w[j], m->exp_b[j], skip ? bufx : nullptr, M, outb, M, 0, M, M, ebase, erows, nullptr);
        SNB_CHECK_LAUNCH("k_linear(expert)");
      }
      in = outb;
      outb = (outb == buf0) ? buf1 : buf0;
    }
    // NOTE: after the skip layer the reference sets x = h (:916); with a single skip this never matters again.
    if ((rc = snb_combine_impl(in, idx, loc, o->no_batch ? begin : nullptr, gate, cap_dev, 0, S, M, rows, t0, true, st))) return rc;
  }
  // 6. sigma head (+noise, softplus(x-1))  (nerf_moe.py:392-416)
  if ((rc = launch_linear(ACT_SOFTPLUS_SHIFT, t0, M, m->sigma_w, m->sigma_b, nullptr, 0, sigma, 1, S, 1, M, 1, nullptr, nullptr, sigma_noise, st))) return rc;
  // 7. layer "1" (act none) written straight into the concat buffer, layer "2" (ReLU), colour (sigmoid)
  if ((rc = launch_linear(ACT_NONE, t0, M, m->l1_w, m->l1_b, nullptr, 0, cat, m->cat_in, S, M, M, 1, nullptr, nullptr, nullptr, st))) return rc;
  if ((rc = launch_linear(ACT_RELU, cat, m->cat_in, m->l2_w, m->l2_b, nullptr, 0, h2, H2, S, H2, m->cat_in, 1, nullptr, nullptr, nullptr, st))) return rc;
  if ((rc = launch_linear(ACT_SIGMOID, h2, H2, m->color_w, m->color_b, nullptr, 0, rgb, 3, S, 3, H2, 1, nullptr, nullptr, nullptr, st))) return rc;
  k_pack_out<<<(unsigned)cdiv(S, 256), 256, 0, st>>>(rgb, sigma, S, out);
  SNB_CHECK_LAUNCH("k_pack_out");
  return SNB_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// forward recomputation for the backward pass (snb_backward.cu): the same kernels as fp32_forward, but every
// intermediate the backward sweep needs stays in its own buffer
// ---------------------------------------------------------------------------------------------------------------

// gate input of the reference: LayerNorm(g) (eps 1e-5, biased variance), one warp per row
__global__ void __launch_bounds__(256) k_ln_fwd(const float* __restrict__ g, int64_t S, int M, const float* __restrict__ w,
                                                const float* __restrict__ b, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= S) return;
  const float* row = g + s * M;
  float sum = 0.f;
  for (int k = lane; k < M; k += 32) sum += row[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)M;
  float var = 0.f;
  for (int k = lane; k < M; k += 32) { const float d = row[k] - mean; var += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / (float)M + 1e-5f);
  for (int k = lane; k < M; k += 32) out[s * M + k] = (row[k] - mean) * rstd * w[k] + b[k];
}
int ln_forward_launch(const float* g, int64_t S, int M, const float* w, const float* b, float* out, cudaStream_t st) {
  k_ln_fwd<<<(unsigned)cdiv(S, 8), 256, 0, st>>>(g, S, M, w, b, out);
  SNB_CHECK_LAUNCH("k_ln_fwd");
  return SNB_OK;
}

int fp32_forward_saved(Model* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* o, Fp32Saved* sv,
                       Arena& ws, cudaStream_t st) {
  const int M = m->d.width, E = m->d.num_experts, L = m->d.expert_layers, H2 = m->d.hidden2, NG = m->d.gate_layers;
  const double cf = o->capacity_factor;
  const int64_t cap_host = capacity_of(S, E, cf);
  int64_t rows = (int64_t)E * cap_host;
  if (rows < 1) rows = 1;
  sv->rows = rows; sv->cap_host = cap_host;
  sv->pe = ws.take<float>((size_t)S * m->xyz_in);
  sv->h = ws.take<float>((size_t)S * M);
  sv->ga[0] = sv->h;
  for (int i = 1; i < NG; ++i) sv->ga[i] = ws.take<float>((size_t)S * M);
  sv->g = ws.take<float>((size_t)S * M);
  sv->gates = ws.take<float>((size_t)S * E);
  sv->idx = ws.take<int>(S);
  sv->loc = ws.take<int>(S);
  sv->gate = ws.take<float>(S);
  sv->bufx = ws.take<float>((size_t)rows * M);
  for (int j = 1; j < L; ++j) sv->act[j] = ws.take<float>((size_t)rows * M);
  sv->act[0] = sv->bufx;
  sv->out_rows = ws.take<float>((size_t)rows * M);
  sv->hr = ws.take<float>((size_t)S * M);
  sv->sig_pre = ws.take<float>(S);
  sv->cat = ws.take<float>((size_t)S * m->cat_in);
  sv->h2 = ws.take<float>((size_t)S * H2);
  sv->rgb = ws.take<float>((size_t)S * 3);
  int* small = ws.take<int>(1024);
  float* l_aux = ws.take<float>(64);
  const size_t rbytes = route_workspace_bytes(S, E);
  char* rws = ws.take<char>(rbytes);
  if (!ws.ok) { set_error("snb_moe_backward: workspace too small"); return SNB_EWORKSPACE; }
  int *counts = small, *cap_dev = small + E, *ebase = small + E + 1, *erows = ebase + E, *begin = erows + E;
  sv->counts = counts; sv->cap_dev = cap_dev; sv->ebase = ebase; sv->erows = erows;
  {
    dim3 blk(32, 8);
    k_encode<<<(unsigned)cdiv(S, 8), blk, 0, st>>>(x, S, m->x_cols, m->d.mip, m->d.pos_xyz_freqs, m->d.pos_dir_freqs,
                                                   m->d.appearance_dim, m->d.appearance_count, m->emb_a, sv->pe, m->xyz_in,
                                                   sv->cat, m->cat_in, M);
    SNB_CHECK_LAUNCH("k_encode");
  }
  int rc;
  if ((rc = launch_linear(ACT_NONE, sv->pe, m->xyz_in, m->xyz_w, m->xyz_b, nullptr, 0, sv->h, M, S, M, m->xyz_in, 1, nullptr, nullptr, nullptr, st))) return rc;
  for (int i = 0; i < NG; ++i) {
    float* outp = (i < NG - 1) ? sv->ga[i + 1] : sv->g;
    if ((rc = launch_linear(i < NG - 1 ? ACT_RELU : ACT_NONE, sv->ga[i], M, m->gate_w[i], m->gate_b[i], nullptr, 0, outp, M, S, M, M, 1, nullptr, nullptr, nullptr, st))) return rc;
  }
  {
    const size_t smem = (size_t)E * M * sizeof(float);
    SNB_REQUIRE(smem <= 200 * 1024, "wg too large for shared memory");
    if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_ln_gate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ln_gate<<<(unsigned)cdiv(S, 8), 256, smem, st>>>(sv->g, S, M, E, m->ln_w, m->ln_b, m->wg, sv->gates);
    SNB_CHECK_LAUNCH("k_ln_gate");
  }
  if ((rc = route_top1(sv->gates, S, E, cf, o->bpr, sv->idx, sv->loc, sv->gate, counts, cap_dev, l_aux, rws, rbytes, st))) return rc;
  k_expert_ranges<<<1, 32, 0, st>>>(counts, cap_dev, E, 0, ebase, erows, begin);
  SNB_CHECK_LAUNCH("k_expert_ranges");
  // the padded dispatch buffer is zero where no sample landed (the skip connection reads it)
  if ((rc = snb_dispatch_impl(sv->h, sv->idx, sv->loc, nullptr, cap_dev, 0, S, M, rows, sv->bufx, true, st))) return rc;
  for (int j = 0; j < L; ++j) {
    const bool skip = (j == m->d.skip_layer);
    float* outb = (j < L - 1) ? sv->act[j + 1] : sv->out_rows;
    if (cap_host > 0) {
      dim3 grid((unsigned)cdiv(cap_host, 64), (unsigned)cdiv(M, 64), (unsigned)E);
      if (j < L - 1)
        k_linear<ACT_RELU><<<grid, 256, 0, st>>>(sv->act[j], M, m->exp_w[j], m->exp_b[j], skip ? sv->bufx : nullptr, M, outb, M, 0, M, M, ebase, erows, nullptr);
      else
        k_linear<ACT_NONE><<<grid, 256, 0, st>>>(sv->act[j], M, m->exp_w[j], m->exp_b[j], skip ? sv->bufx : nullptr, M, outb, M, 0, M, M, ebase, erows, nullptr);
      SNB_CHECK_LAUNCH("k_linear(expert)");
    }
  }
  if ((rc = snb_combine_impl(sv->out_rows, sv->idx, sv->loc, nullptr, sv->gate, cap_dev, 0, S, M, rows, sv->hr, true, st))) return rc;
  if ((rc = launch_linear(ACT_NONE, sv->hr, M, m->sigma_w, m->sigma_b, nullptr, 0, sv->sig_pre, 1, S, 1, M, 1, nullptr, nullptr, nullptr, st))) return rc;
  if ((rc = launch_linear(ACT_NONE, sv->hr, M, m->l1_w, m->l1_b, nullptr, 0, sv->cat, m->cat_in, S, M, M, 1, nullptr, nullptr, nullptr, st))) return rc;
  if ((rc = launch_linear(ACT_RELU, sv->cat, m->cat_in, m->l2_w, m->l2_b, nullptr, 0, sv->h2, H2, S, H2, m->cat_in, 1, nullptr, nullptr, nullptr, st))) return rc;
  if ((rc = launch_linear(ACT_SIGMOID, sv->h2, H2, m->color_w, m->color_b, nullptr, 0, sv->rgb, 3, S, 3, H2, 1, nullptr, nullptr, nullptr, st))) return rc;
  (void)sigma_noise;
  return SNB_OK;
}

}  // namespace snb

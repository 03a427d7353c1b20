// Top-1 routing on device: argmax / gate value / load-balance loss / location-in-expert
// (plain order or batch-prioritised order), bit-exact against extract_critical
// (reference modules/tutel_moe_ext/tutel_fast_dispatch.py:176-217, k = 1) under the
// tie-break contract of SURVEY.md Appendix A: argmax -> lowest expert id; BPR order ->
// descending max-gate, ties by ascending sample index (== a stable sort).
//
// Everything is integer / ballot work bound by HBM and launch latency, not by math:
//   k_top1        1 thread / sample, coalesced [S,E] fp32 read (4E B/sample), writes idx/gate/key
//   radix passes  stable LSD radix sort of (key, sample) pairs, 9-bit digits, only for BPR
//   k_loc         warp-ballot (match_any) rank inside a 256-sample block + scanned block offsets
// No host round trip: capacity / counts / l_aux stay in device memory.
#include "snb_common.cuh"

namespace snb {

static constexpr int RB = 256;      // samples per block in top1 / hist / loc kernels
static constexpr int ST = 2048;     // sort tile (elements per block)
static constexpr int SORT_THREADS = 256;
static constexpr int DBITS = 9;            // radix digit width
static constexpr int NBINS = 1 << DBITS;

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RB) k_top1(const float* __restrict__ gates, int64_t S, int E,
                                             int* __restrict__ idx, float* __restrict__ gate,
                                             uint32_t* __restrict__ ukey, int softmax_keys,
                                             float* __restrict__ pm, int* __restrict__ pc) {
  extern __shared__ float sm[];  // [RB/32][E] partial me + [RB/32][E] partial counts (as int)
  const int warps = RB / 32;
  float* s_me = sm;
  int* s_ce = (int*)(sm + warps * E);
  const int64_t s = (int64_t)blockIdx.x * RB + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int best = -1;
  float bv = 0.f;
  const bool valid = s < S;
  for (int e = 0; e < E; ++e) {
    float g = valid ? gates[s * E + e] : 0.f;
    if (valid && (best < 0 || g > bv)) { best = e; bv = g; }   // strict > : lowest index wins ties
    // warp-reduce column sums (fp32 tree; final cross-block sum is done in double)
    float v = g;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_me[w * E + e] = v;
  }
  if (valid) {
    idx[s] = best;
    gate[s] = bv;
    if (ukey) {
      uint32_t b = __float_as_uint(bv);
      uint32_t k;
      if (softmax_keys) {
        // gates in (0, 1]: float bits are monotone; descending order key = bits(1.0f) - bits(g)
        k = (b <= 0x3F800000u) ? (0x3F800000u - b) : 0u;
      } else {
        uint32_t asc = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
        k = ~asc;
      }
      ukey[s] = k;
    }
  }
  // per-warp expert histogram through match_any
  unsigned m = __match_any_sync(0xffffffffu, best);
  for (int e = lane; e < E; e += 32) s_ce[w * E + e] = 0;
  __syncwarp();
  if (best >= 0 && (m & ((1u << lane) - 1)) == 0) s_ce[w * E + best] = __popc(m);
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += RB) {
    float a = 0.f;
    int c = 0;
    for (int ww = 0; ww < warps; ++ww) { a += s_me[ww * E + e]; c += s_ce[ww * E + e]; }
    pm[(int64_t)blockIdx.x * E + e] = a;
    pc[(int64_t)blockIdx.x * E + e] = c;
  }
}

// per-block expert histogram over a permuted sequence (BPR): e(p) = idx[order[p]]
__global__ void __launch_bounds__(RB) k_expert_hist(const int* __restrict__ idx, const uint32_t* __restrict__ order,
                                                    int64_t S, int E, int* __restrict__ pc) {
  extern __shared__ int s_ce[];  // [RB/32][E]
  const int warps = RB / 32;
  const int64_t p = (int64_t)blockIdx.x * RB + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int e = -1;
  if (p < S) e = idx[order ? order[p] : p];
  unsigned m = __match_any_sync(0xffffffffu, e);
  for (int k = lane; k < E; k += 32) s_ce[w * E + k] = 0;
  __syncwarp();
  if (e >= 0 && (m & ((1u << lane) - 1)) == 0) s_ce[w * E + e] = __popc(m);
  __syncthreads();
  for (int k = threadIdx.x; k < E; k += RB) {
    int c = 0;
    for (int ww = 0; ww < warps; ++ww) c += s_ce[ww * E + k];
    pc[(int64_t)blockIdx.x * E + k] = c;
  }
}

// One block: exclusive scan of the per-block expert counts (-> blockoff), totals, capacity, l_aux.
// Warp w scans the blocks of expert e = w, w+nwarps, ... with shuffle prefix sums (coalesced in b).
__global__ void __launch_bounds__(1024) k_finalize(const float* __restrict__ pm, const int* __restrict__ pc, int nblk,
                                                   int E, int64_t S, double cf, int* __restrict__ counts,
                                                   int* __restrict__ capacity, float* __restrict__ l_aux,
                                                   int* __restrict__ blockoff, int write_stats) {
  __shared__ float s_prod[1024];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_prod[i] = 0.f;
  __syncthreads();
  for (int e = w; e < E; e += nw) {
    int run = 0;
    double me = 0.0;
    for (int b0 = 0; b0 < nblk; b0 += 32) {
      const int b = b0 + lane;
      int c = (b < nblk) ? pc[(int64_t)b * E + e] : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (b < nblk) blockoff[(int64_t)b * E + e] = run + inc - c;
      run += __shfl_sync(0xffffffffu, inc, 31);
      if (write_stats) {
        double m = (b < nblk) ? (double)pm[(int64_t)b * E + e] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
        me += m;
      }
    }
    if (write_stats && lane == 0) {
      counts[e] = run;
      s_prod[e] = (float)me * (float)run;   // me * ce in fp32 (tutel_fast_dispatch.py:143-145)
    }
  }
  if (!write_stats) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int k = 0; k < E; ++k) acc += s_prod[k];
    double scale = (double)E / ((double)S * (double)S);
    if (l_aux) *l_aux = (float)((double)acc * scale);
    if (capacity) *capacity = capacity_of(S, E, cf);
  }
}

// loc[s] = blockoff[block][e] + (#same-expert samples earlier in this block)
__global__ void __launch_bounds__(RB) k_loc(const int* __restrict__ idx, const uint32_t* __restrict__ order, int64_t S,
                                            int E, const int* __restrict__ blockoff, int* __restrict__ loc) {
  extern __shared__ int s_cnt[];  // [RB/32][E]
  const int warps = RB / 32;
  const int64_t p = (int64_t)blockIdx.x * RB + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int e = -1;
  int64_t s = -1;
  if (p < S) { s = order ? (int64_t)order[p] : p; e = idx[s]; }
  unsigned m = __match_any_sync(0xffffffffu, e);
  int r = __popc(m & ((1u << lane) - 1));
  for (int k = lane; k < E; k += 32) s_cnt[w * E + k] = 0;
  __syncwarp();
  if (e >= 0 && r == 0) s_cnt[w * E + e] = __popc(m);
  __syncthreads();
  if (e >= 0) {
    int pre = blockoff[(int64_t)blockIdx.x * E + e];
    for (int ww = 0; ww < w; ++ww) pre += s_cnt[ww * E + e];
    loc[s] = pre + r;
  }
}

// ------------------------------- stable LSD radix sort ---------------------------------
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint32_t* __restrict__ keys, int64_t S, int shift,
                                                            int nblk, int* __restrict__ hist /*[NBINS][nblk]*/) {
  __shared__ int h[NBINS];
  for (int i = threadIdx.x; i < NBINS; i += SORT_THREADS) h[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * ST;
  for (int j = 0; j < ST / SORT_THREADS; ++j) {
    int64_t i = base + j * SORT_THREADS + threadIdx.x;
    if (i < S) atomicAdd(&h[(keys[i] >> shift) & (NBINS - 1)], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NBINS; i += SORT_THREADS) hist[(int64_t)i * nblk + blockIdx.x] = h[i];
}

// exclusive scan of hist[NBINS*nblk] in place (single block: per-thread runs + shuffle block scan)
__global__ void __launch_bounds__(1024) k_sort_scan(int* __restrict__ hist, int n) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per = (n + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(n, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += hist[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = wsum[lane], vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= o) vi += t;
    }
    wsum[lane] = vi - v;
  }
  __syncthreads();
  int run = wsum[w] + inc - sum;
  for (int i = lo; i < hi; ++i) { int t = hist[i]; hist[i] = run; run += t; }
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint32_t* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in, int64_t S,
                                                               int shift, int nblk, const int* __restrict__ hist,
                                                               uint32_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out) {
  __shared__ int run[NBINS];                       // elements of each digit already placed by this block
  __shared__ int wcnt[SORT_THREADS / 32][NBINS];   // per-warp digit counts of the current round
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < NBINS; k += SORT_THREADS) run[k] = 0;
  for (int k = threadIdx.x; k < (SORT_THREADS / 32) * NBINS; k += SORT_THREADS) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * ST;
  for (int j = 0; j < ST / SORT_THREADS; ++j) {
    const int64_t i = base + j * SORT_THREADS + threadIdx.x;
    const bool valid = i < S;
    uint32_t key = valid ? keys_in[i] : 0u;
    int d = valid ? (int)((key >> shift) & (NBINS - 1)) : -1;
    unsigned m = __match_any_sync(0xffffffffu, d);
    int r = __popc(m & ((1u << lane) - 1));
    if (valid && r == 0) wcnt[w][d] = __popc(m);
    __syncthreads();
    if (valid) {
      int off = run[d] + r;
      for (int ww = 0; ww < w; ++ww) off += wcnt[ww][d];
      int64_t dst = (int64_t)hist[(int64_t)d * nblk + blockIdx.x] + off;
      keys_out[dst] = key;
      vals_out[dst] = vals_in ? vals_in[i] : (uint32_t)i;
    }
    __syncthreads();
    for (int d2 = threadIdx.x; d2 < NBINS; d2 += SORT_THREADS) {
      int c = 0;
#pragma unroll
      for (int ww = 0; ww < SORT_THREADS / 32; ++ww) { c += wcnt[ww][d2]; wcnt[ww][d2] = 0; }
      run[d2] += c;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
size_t route_workspace_bytes(int64_t S, int32_t E) {
  const int64_t nblk = cdiv(S > 0 ? S : 1, RB), nsort = cdiv(S > 0 ? S : 1, ST);
  size_t b = 0;
  b += align_up((size_t)nblk * E * sizeof(float), 256);   // pm
  b += align_up((size_t)nblk * E * sizeof(int), 256);     // pc
  b += align_up((size_t)nblk * E * sizeof(int), 256);     // blockoff
  b += 4 * align_up((size_t)(S > 0 ? S : 1) * sizeof(uint32_t), 256);  // keys x2, vals x2
  b += align_up((size_t)NBINS * nsort * sizeof(int), 256);  // sort hist
  return b + 1024;
}

static int route_impl(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int softmax_keys,
                      int32_t* idx, int32_t* loc, float* gate, int32_t* counts, int32_t* capacity,
                      float* l_aux, void* ws, size_t ws_bytes, cudaStream_t st) {
  SNB_REQUIRE(E >= 1 && E <= 1024, "route: E=%d out of range [1,1024]", E);
  SNB_REQUIRE(S >= 0 && S < (1ll << 31), "route: S=%lld out of range", (long long)S);
  if (S == 0) {
    if (counts) SNB_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * E, st));
    if (capacity) SNB_CHECK_CUDA(cudaMemsetAsync(capacity, 0, sizeof(int), st));
    if (l_aux) SNB_CHECK_CUDA(cudaMemsetAsync(l_aux, 0, sizeof(float), st));
    return SNB_OK;
  }
  SNB_REQUIRE(gates && idx && loc && gate && counts, "route: NULL pointer");
  Arena a(ws, ws_bytes);
  const int nblk = (int)cdiv(S, RB), nsort = (int)cdiv(S, ST);
  float* pm = a.take<float>((size_t)nblk * E);
  int* pc = a.take<int>((size_t)nblk * E);
  int* blockoff = a.take<int>((size_t)nblk * E);
  uint32_t* k0 = a.take<uint32_t>(S);
  uint32_t* k1 = a.take<uint32_t>(S);
  uint32_t* v0 = a.take<uint32_t>(S);
  uint32_t* v1 = a.take<uint32_t>(S);
  int* shist = a.take<int>((size_t)NBINS * nsort);
  if (!a.ok) { set_error("route: workspace too small (%zu bytes given)", ws_bytes); return SNB_EWORKSPACE; }

  const size_t smem = (size_t)(RB / 32) * E * (sizeof(float) + sizeof(int));
  k_top1<<<nblk, RB, smem, st>>>(gates, S, E, idx, gate, bpr ? k0 : nullptr, softmax_keys, pm, pc);
  SNB_CHECK_LAUNCH("k_top1");
  const int fin_threads = (E >= 32) ? 1024 : 32 * E;
  k_finalize<<<1, fin_threads, 0, st>>>(pm, pc, nblk, E, S, cf, counts, capacity, l_aux, blockoff, 1);
  SNB_CHECK_LAUNCH("k_finalize");
  const uint32_t* order = nullptr;
  if (bpr) {
    // number of significant key bits: softmax keys are < bits(1.0f) - bits(tiny) ; generic keys use 32
    int nbits = 32;
    if (softmax_keys) {
      // max gate >= 1/E (up to rounding of the softmax)  =>  key <= bits(1.0) - bits(0.99/E)
      float lo = 0.99f / (float)E;
      uint32_t lob;
      memcpy(&lob, &lo, 4);
      uint32_t maxk = 0x3F800000u - lob;
      nbits = 32 - __builtin_clz(maxk | 1u);
    }
    const int passes = (nbits + DBITS - 1) / DBITS;
    uint32_t *kin = k0, *kout = k1, *vin = nullptr, *vout = v0;
    for (int p = 0; p < passes; ++p) {
      const int shift = DBITS * p;
      k_sort_hist<<<nsort, SORT_THREADS, 0, st>>>(kin, S, shift, nsort, shist);
      SNB_CHECK_LAUNCH("k_sort_hist");
      k_sort_scan<<<1, 1024, 0, st>>>(shist, NBINS * nsort);
      SNB_CHECK_LAUNCH("k_sort_scan");
      k_sort_scatter<<<nsort, SORT_THREADS, 0, st>>>(kin, vin, S, shift, nsort, shist, kout, vout);
      SNB_CHECK_LAUNCH("k_sort_scatter");
      uint32_t* t = kin; kin = kout; kout = t;
      vin = vout;
      vout = (vout == v0) ? v1 : v0;
    }
    order = vin;
    k_expert_hist<<<nblk, RB, (size_t)(RB / 32) * E * sizeof(int), st>>>(idx, order, S, E, pc);
    SNB_CHECK_LAUNCH("k_expert_hist");
    k_finalize<<<1, fin_threads, 0, st>>>(pm, pc, nblk, E, S, cf, counts, capacity, l_aux, blockoff, 0);
    SNB_CHECK_LAUNCH("k_finalize2");
  }
  k_loc<<<nblk, RB, (size_t)(RB / 32) * E * sizeof(int), st>>>(idx, order, S, E, blockoff, loc);
  SNB_CHECK_LAUNCH("k_loc");
  return SNB_OK;
}

int route_top1(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int32_t* idx, int32_t* loc,
               float* gate, int32_t* counts, int32_t* capacity, float* l_aux, void* ws, size_t ws_bytes,
               cudaStream_t st) {
  // internal callers always pass softmax outputs
  return route_impl(gates, S, E, cf, bpr, 1, idx, loc, gate, counts, capacity, l_aux, ws, ws_bytes, st);
}

int route_top1_generic(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int32_t* idx,
                       int32_t* loc, float* gate, int32_t* counts, int32_t* capacity, float* l_aux, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  return route_impl(gates, S, E, cf, bpr, 0, idx, loc, gate, counts, capacity, l_aux, ws, ws_bytes, st);
}

}  // namespace snb

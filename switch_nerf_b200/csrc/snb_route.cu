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
#include <atomic>
#include "snb_common.cuh"
#include "snb_ep.cuh"
#include "snb_select.cuh"

namespace snb {

static constexpr int RB = 256;      // samples per block in top1 / hist / loc kernels
static constexpr int ST = 2048;     // sort tile (elements per block)
static constexpr int SORT_THREADS = 256;
static constexpr int DBITS = 9;            // radix digit width
static constexpr int NBINS = 1 << DBITS;

// ---------------------------------------------------------------------------------------
// k_top1: one CTA per sort tile (ST = 2048 samples = 8 rounds of 256).  Per sample: argmax, gate value,
// BPR key.  Per 256-sample sub-block: expert histogram pc (for the location scan).  Per tile: partial
// column sums pm (load-balance loss) and -- for BPR -- the radix-sort bookkeeping of ALL passes that does
// not depend on the element order: digit totals of every pass, and the per-tile histogram of pass 0.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RB) k_top1(const float* __restrict__ gates, int64_t S, int E,
                                             int* __restrict__ idx, float* __restrict__ gate,
                                             uint32_t* __restrict__ ukey, int softmax_keys, int npass,
                                             float* __restrict__ pm, int* __restrict__ pc,
                                             int* __restrict__ totals /*[npass][NBINS]*/,
                                             int* __restrict__ thist0 /*[ntile][NBINS]*/) {
  extern __shared__ float sm[];  // [8 warps][E] floats (me) + [8][E] ints (ce) + [3][NBINS] ints (digit hists)
  const int warps = RB / 32;
  float* s_me = sm;
  int* s_ce = (int*)(sm + warps * E);
  int* s_h = s_ce + warps * E;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (ukey)
    for (int i = threadIdx.x; i < npass * NBINS; i += RB) s_h[i] = 0;
  for (int e = lane; e < E; e += 32) s_me[w * E + e] = 0.f;
  __syncthreads();
  for (int j = 0; j < ST / RB; ++j) {
    const int64_t s = (int64_t)blockIdx.x * ST + (int64_t)j * RB + threadIdx.x;
    const bool valid = s < S;
    if ((int64_t)blockIdx.x * ST + (int64_t)j * RB >= S) break;   // block-uniform
    int best = -1;
    float bv = 0.f;
    for (int e = 0; e < E; ++e) {
      float g = valid ? gates[s * E + e] : 0.f;
      if (valid && (best < 0 || g > bv)) { best = e; bv = g; }   // strict > : lowest index wins ties
      float v = g;                                                // warp-reduce the column sum
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_me[w * E + e] += v;
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
        for (int p = 0; p < npass; ++p) atomicAdd(&s_h[p * NBINS + ((k >> (DBITS * p)) & (NBINS - 1))], 1);
      }
    }
    // expert histogram of this 256-sample sub-block through match_any
    unsigned m = __match_any_sync(0xffffffffu, best);
    for (int e = lane; e < E; e += 32) s_ce[w * E + e] = 0;
    __syncwarp();
    if (best >= 0 && (m & ((1u << lane) - 1)) == 0) s_ce[w * E + best] = __popc(m);
    __syncthreads();
    const int64_t sub = (int64_t)blockIdx.x * (ST / RB) + j;
    for (int e = threadIdx.x; e < E; e += RB) {
      int c = 0;
      for (int ww = 0; ww < warps; ++ww) c += s_ce[ww * E + e];
      pc[sub * E + e] = c;
    }
    __syncthreads();
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += RB) {
    float a = 0.f;
    for (int ww = 0; ww < warps; ++ww) a += s_me[ww * E + e];
    pm[(int64_t)blockIdx.x * E + e] = a;
  }
  if (ukey) {
    for (int i = threadIdx.x; i < NBINS; i += RB) thist0[(int64_t)blockIdx.x * NBINS + i] = s_h[i];
    for (int i = threadIdx.x; i < npass * NBINS; i += RB)
      if (s_h[i]) atomicAdd(&totals[i], s_h[i]);
  }
}

// One block: exclusive scan of the per-block expert counts (-> blockoff), totals, capacity, l_aux.
// Warp w scans the blocks of expert e = w, w+nwarps, ... with shuffle prefix sums (coalesced in b).
__global__ void __launch_bounds__(1024) k_finalize(const float* __restrict__ pm, int npm, const int* __restrict__ pc,
                                                   int nblk, int E, int64_t S, double cf, int* __restrict__ counts,
                                                   int* __restrict__ capacity, float* __restrict__ l_aux,
                                                   int* __restrict__ blockoff, int write_stats) {
  __shared__ float s_prod[1024];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_prod[i] = 0.f;
  __syncthreads();
  for (int e = w; e < E; e += nw) {
    int run = 0;
    double me = 0.0;
    for (int bb = 0; bb < nblk; bb += 32 * 8) {
      int cv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {           // independent loads first (L2 latency paid once per batch)
        const int b = bb + u * 32 + lane;
        cv[u] = (b < nblk) ? pc[(int64_t)b * E + e] : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = bb + u * 32 + lane;
        const int c = cv[u];
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (b < nblk) blockoff[(int64_t)b * E + e] = run + inc - c;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    if (write_stats) {
      for (int b0 = 0; b0 < npm; b0 += 32) {
        const int b = b0 + lane;
        double m = (b < npm) ? (double)pm[(int64_t)b * E + e] : 0.0;
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
// One kernel per pass.  Tile b (2048 consecutive elements of the current order) derives its own
// destination offsets: exclusive scan of the digit totals (order independent, from k_top1) plus
// the sum of the per-tile histograms of the tiles before it; ranks inside the tile come from
// warp match_any ballots (stable).  While scattering it builds the NEXT pass's per-tile histogram
// (global atomics keyed by the destination tile) or, on the last pass, the per-256-block expert
// histogram of the sorted sequence that the location scan needs.
__global__ void __launch_bounds__(SORT_THREADS) k_sort_pass(const uint32_t* __restrict__ keys_in,
                                                            const uint32_t* __restrict__ vals_in, int64_t S, int pass,
                                                            int last, const int* __restrict__ totals,
                                                            const int* __restrict__ thist, int* __restrict__ thist_next,
                                                            const int* __restrict__ idx, int E, int* __restrict__ pc2,
                                                            uint32_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ vals_out) {
  constexpr int NW = SORT_THREADS / 32, PER = ST / SORT_THREADS;   // 8 warps x 8 rounds: warp w owns elements [w*256, +256)
  __shared__ int off0[NBINS];            // global offset of the first element of each digit from this tile
  __shared__ int whist[NW][NBINS];       // per-warp digit counts -> exclusive prefix over the warps of the tile
  __shared__ int wsum[NW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int shift = DBITS * pass;
  for (int k = threadIdx.x; k < NW * NBINS; k += SORT_THREADS) (&whist[0][0])[k] = 0;
  {
    const int d0 = 2 * threadIdx.x;                // NBINS == 2 * SORT_THREADS
    const int v0 = totals[pass * NBINS + d0], v1 = totals[pass * NBINS + d0 + 1];
    int inc = v0 + v1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int wbase = 0;
    for (int ww = 0; ww < w; ++ww) wbase += wsum[ww];
    const int base0 = wbase + inc - (v0 + v1);
    int p0 = 0, p1 = 0;
    const int2* th = reinterpret_cast<const int2*>(thist);
#pragma unroll 8
    for (int b = 0; b < (int)blockIdx.x; ++b) {
      int2 t = th[(int64_t)b * (NBINS / 2) + threadIdx.x];
      p0 += t.x; p1 += t.y;
    }
    off0[d0] = base0 + p0;
    off0[d0 + 1] = base0 + v0 + p1;
  }
  __syncthreads();
  // phase 1: ranks inside the warp's own 256 elements (warp-synchronous, no block barriers)
  const int64_t wbase_i = (int64_t)blockIdx.x * ST + (int64_t)w * (PER * 32);
  uint32_t key[PER], val[PER];
  int rk[PER];                                    // rank among equal digits inside this warp, -1 = out of range
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int64_t i = wbase_i + j * 32 + lane;
    const bool valid = i < S;
    key[j] = valid ? keys_in[i] : 0u;
    val[j] = valid ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
    const int d = valid ? (int)((key[j] >> shift) & (NBINS - 1)) : -1;
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int r = __popc(m & ((1u << lane) - 1));
    int prev = 0;
    if (valid) prev = whist[w][d];
    __syncwarp();
    if (valid && r == 0) whist[w][d] = prev + __popc(m);
    __syncwarp();
    rk[j] = valid ? prev + r : -1;
  }
  __syncthreads();
  // phase 2: exclusive prefix over the warps, per digit (2 digits per thread)
  for (int d = threadIdx.x; d < NBINS; d += SORT_THREADS) {
    int run = off0[d];
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) { const int c = whist[ww][d]; whist[ww][d] = run; run += c; }
  }
  __syncthreads();
  // phase 3: scatter
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    if (rk[j] < 0) continue;
    const int d = (int)((key[j] >> shift) & (NBINS - 1));
    const int off = whist[w][d] + rk[j];
    keys_out[off] = key[j];
    vals_out[off] = val[j];
    if (!last) atomicAdd(&thist_next[(int64_t)(off / ST) * NBINS + ((key[j] >> (shift + DBITS)) & (NBINS - 1))], 1);
    else atomicAdd(&pc2[(int64_t)(off / RB) * E + idx[val[j]]], 1);
  }
}

static constexpr int MAX_PASS = 4;
static_assert(NBINS == 2 * SORT_THREADS, "k_sort_pass assumes two digits per thread");

size_t route_workspace_bytes(int64_t S, int32_t E) {
  const int64_t Sx = S > 0 ? S : 1;
  const int64_t nblk = cdiv(Sx, RB), ntile = cdiv(Sx, ST);
  size_t b = 0;
  b += align_up((size_t)ntile * E * sizeof(float), 256);  // pm
  b += align_up((size_t)nblk * E * sizeof(int), 256);     // pc
  b += align_up((size_t)nblk * E * sizeof(int), 256);     // blockoff
  b += 4 * align_up((size_t)Sx * sizeof(uint32_t), 256);  // keys x2, vals x2
  b += align_up(((size_t)MAX_PASS * NBINS + (size_t)MAX_PASS * ntile * NBINS + (size_t)nblk * E) * sizeof(int), 256);
  return b + 1024;
}

static int route_impl(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int softmax_keys,
                      int32_t* idx, int32_t* loc, float* gate, int32_t* counts, int32_t* capacity,
                      float* l_aux, void* ws, size_t ws_bytes, cudaStream_t st) {
  SNB_REQUIRE(E >= 1 && E <= 512, "route: E=%d out of range [1,512]", E);
  SNB_REQUIRE(S >= 0 && S < (1ll << 31), "route: S=%lld out of range", (long long)S);
  if (S == 0) {
    if (counts) SNB_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * E, st));
    if (capacity) SNB_CHECK_CUDA(cudaMemsetAsync(capacity, 0, sizeof(int), st));
    if (l_aux) SNB_CHECK_CUDA(cudaMemsetAsync(l_aux, 0, sizeof(float), st));
    return SNB_OK;
  }
  SNB_REQUIRE(gates && idx && loc && gate && counts, "route: NULL pointer");
  Arena a(ws, ws_bytes);
  const int nblk = (int)cdiv(S, RB), ntile = (int)cdiv(S, ST);
  float* pm = a.take<float>((size_t)ntile * E);
  int* pc = a.take<int>((size_t)nblk * E);
  int* blockoff = a.take<int>((size_t)nblk * E);
  uint32_t* k0 = a.take<uint32_t>(S);
  uint32_t* k1 = a.take<uint32_t>(S);
  uint32_t* v0 = a.take<uint32_t>(S);
  uint32_t* v1 = a.take<uint32_t>(S);
  const size_t zero_ints = (size_t)MAX_PASS * NBINS + (size_t)MAX_PASS * ntile * NBINS + (size_t)nblk * E;
  int* zeroed = a.take<int>(zero_ints);
  if (!a.ok) { set_error("route: workspace too small (%zu bytes given)", ws_bytes); return SNB_EWORKSPACE; }
  int* totals = zeroed;                                   // [MAX_PASS][NBINS]
  int* thist = zeroed + MAX_PASS * NBINS;                 // [MAX_PASS][ntile][NBINS]
  int* pc2 = thist + (size_t)MAX_PASS * ntile * NBINS;    // [nblk][E]

  // number of significant key bits: softmax keys are <= bits(1.0) - bits(~1/E); generic keys use all 32
  int npass = 0;
  if (bpr) {
    int nbits = 32;
    if (softmax_keys) {
      float lo = 0.99f / (float)E;     // max gate >= 1/E up to rounding of the softmax
      uint32_t lob;
      memcpy(&lob, &lo, 4);
      nbits = 32 - __builtin_clz((0x3F800000u - lob) | 1u);
    }
    npass = (nbits + DBITS - 1) / DBITS;
    SNB_CHECK_CUDA(cudaMemsetAsync(zeroed, 0, zero_ints * sizeof(int), st));
  }
  const size_t smem = (size_t)(RB / 32) * E * (sizeof(float) + sizeof(int)) + (size_t)MAX_PASS * NBINS * sizeof(int);
  k_top1<<<ntile, RB, smem, st>>>(gates, S, E, idx, gate, bpr ? k0 : nullptr, softmax_keys, npass, pm, pc, totals, thist);
  SNB_CHECK_LAUNCH("k_top1");
  const int fin_threads = (E >= 32) ? 1024 : 32 * E;
  k_finalize<<<1, fin_threads, 0, st>>>(pm, ntile, pc, nblk, E, S, cf, counts, capacity, l_aux, blockoff, 1);
  SNB_CHECK_LAUNCH("k_finalize");
  const uint32_t* order = nullptr;
  if (bpr) {
    uint32_t *kin = k0, *kout = k1, *vin = nullptr, *vout = v0;
    for (int p = 0; p < npass; ++p) {
      const int last = (p == npass - 1);
      k_sort_pass<<<ntile, SORT_THREADS, 0, st>>>(kin, vin, S, p, last, totals, thist + (size_t)p * ntile * NBINS,
                                                  thist + (size_t)(p + 1 < MAX_PASS ? p + 1 : p) * ntile * NBINS, idx,
                                                  E, pc2, kout, vout);
      SNB_CHECK_LAUNCH("k_sort_pass");
      uint32_t* t = kin; kin = kout; kout = t;
      vin = vout;
      vout = (vout == v0) ? v1 : v0;
    }
    order = vin;
    k_finalize<<<1, fin_threads, 0, st>>>(pm, ntile, pc2, nblk, E, S, cf, counts, capacity, l_aux, blockoff, 0);
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

// =========================================================================================
// Routing as a per-expert radix SELECT (fused bf16 path): launch #2 only needs, per expert, the SET of samples
// whose batch-prioritised rank is below the capacity -- not their order.  Input: one packed word per sample
// (expert id << 26 | key, key = bits(1.0f) - bits(max gate): ascending key == descending gate), written either by
// launch #1 itself (k_front_ts) or by k_pack_top1 below.
// The threshold of an expert is the composite value (key << 32 | sample) of its capacity-th element: ties between
// equal gates resolve to the lower sample index exactly as the stable sort of the full path does.  Kept samples get
// their row in index order, dropped samples a row of the dropped bucket: no atomics on the output, deterministic row
// order; the kernel also writes counts, capacity, l_aux and the tile table of launch #2.
// reference: tutel_fast_dispatch.py:136-139, 176-217 (the kept set == {s : locations_s < capacity}).
// =========================================================================================
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// gates [S,E] -> packed words, level-0 histogram, partial column sums (what k_front_ts does in its softmax epilogue);
// one block per 2048 samples.  pm: [gridDim.x][SEL_PM_STRIDE]
__global__ void __launch_bounds__(256) k_pack_top1(const float* __restrict__ gates, int64_t S, int E,
                                                   uint32_t* __restrict__ w, int* __restrict__ hist0,
                                                   float* __restrict__ pm) {
  __shared__ float s_me[8][SEL_MAX_E];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int e = lane; e < SEL_MAX_E; e += 32) s_me[wid][e] = 0.f;
  __syncwarp();
  for (int j = 0; j < 8; ++j) {
    const int64_t s = (int64_t)blockIdx.x * 2048 + j * 256 + threadIdx.x;
    const bool valid = s < S;
    int best = 0;
    float bv = 0.f;
    for (int e = 0; e < E; ++e) {
      const float g = valid ? gates[s * E + e] : 0.f;
      if (valid && (e == 0 || g > bv)) { best = e; bv = g; }
      float v = g;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_me[wid][e] += v;
    }
    if (valid) {
      const uint32_t key = sel_key(bv);
      w[s] = sel_pack(best, key);
      atomicAdd(&hist0[best * SEL_HBINS + (int)(key >> SEL_L1_SHIFT)], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < SEL_PM_STRIDE) {
    float a = 0.f;
    if ((int)threadIdx.x < E)
      for (int ww = 0; ww < 8; ++ww) a += s_me[ww][threadIdx.x];
    pm[(int64_t)blockIdx.x * SEL_PM_STRIDE + threadIdx.x] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k_select: SEL_P CTAs, each owning a contiguous range of the samples (every word is classified ONCE per sweep, by
// one CTA, for whatever expert it belongs to) -- the ranges' results are combined through a few words of global
// memory and grid barriers (the CTAs are co-resident: SEL_P <= 16 CTAs, launched on SMs launch #1 leaves free).
//   1. thresholds, level by level, identically in every CTA: level 0 from the histogram launch #1 accumulated; a deeper
//      level only for experts whose threshold bucket is still split: every CTA histograms the next digit of ITS
//      matching words (shared memory, warp-aggregated), adds the non-zero bins to the global level histogram, barrier,
//      every CTA picks the digit.  Digits: key bits [19,26) [9,19) [0,9), then the sample index 9 bits at a time.
//   2. per-(CTA, expert, kept|dropped) totals -> global, barrier -> this CTA's first row in every bucket.
//   3. ordered write: rank inside a 32-sample visit from __match_any_sync, visit prefixes from a shared-memory scan.
// ---------------------------------------------------------------------------------------------------------------
static constexpr int SEL_THREADS = 1024;
static constexpr int SEL_P = 8;                      // CTAs
static constexpr int SEL_CACHE = 16384;              // words of the CTA's range kept in shared memory / per group
static constexpr int SEL_VIS = SEL_CACHE / 32;       // 32-sample visits per group
static constexpr int SEL_NLEV = 7;
static_assert(SEL_LVH_INTS == (SEL_NLEV - 1) * SEL_MAX_E * 1024, "level histograms in the zeroed region");

struct SelSmem {
  uint32_t words[SEL_CACHE];
  int hist[SEL_MAX_E][1024];
  uint16_t segcnt[SEL_VIS][32];       // per visit and key (expert * 2 + dropped): count, then exclusive prefix
  int part[32][33];
  double red[32][SEL_MAX_E];
  unsigned long long pval[SEL_MAX_E], pmask[SEL_MAX_E];
  long long Ts[SEL_MAX_E];
  uint32_t Tk[SEL_MAX_E];
  int cnt[SEL_MAX_E], need[SEL_MAX_E], done[SEL_MAX_E];
  int seg0[SEL_MAX_E + 1], dropb[SEL_MAX_E];
  int tot[32], base[32], run[32];
  int flag;
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sel_grid_barrier(int* ctr, int n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1);
    while (ld_acquire_gpu(ctr) < n) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void sel_level_digit(int lv, int sbits, int& dshift, int& dbits) {
  if (lv == 0) { dshift = 32 + SEL_L1_SHIFT; dbits = SEL_KEY_BITS - SEL_L1_SHIFT; }
  else if (lv == 1) { dshift = 32 + SEL_L2_SHIFT; dbits = SEL_L1_SHIFT - SEL_L2_SHIFT; }
  else if (lv == 2) { dshift = 32; dbits = SEL_L2_SHIFT; }
  else { const int hi = sbits - 9 * (lv - 3); dshift = hi > 9 ? hi - 9 : 0; dbits = hi - dshift; }
}

__global__ void __launch_bounds__(SEL_THREADS, 1) k_select(SelectArgs a) {
  extern __shared__ __align__(16) uint8_t sel_smem_raw[];
  SelSmem& sh = *reinterpret_cast<SelSmem*>(sel_smem_raw);
  const int p = blockIdx.x, P = gridDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int E = a.E;
  const int64_t S = a.S;
  const uint32_t* __restrict__ w = a.w;
  int* bar = a.zero + SEL_Z_BAR;
  int tn = 0, nbar = 0, lv_used = 0;
  auto mark = [&](int tag) {
    if (a.tl && p == 0 && tid == 0) a.tl[tn++] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
  };
  mark(1);
  // this CTA's range of 32-sample visits
  const int64_t nvis = (S + 31) / 32;
  const int64_t v0 = nvis * p / P, v1 = nvis * (p + 1) / P;
  const bool cached = (v1 - v0) <= SEL_VIS;
  if (cached)
    for (int64_t i = tid; i < (v1 - v0) * 32; i += SEL_THREADS) {
      const int64_t s = v0 * 32 + i;
      sh.words[i] = (s < S) ? w[s] : 0xffffffffu;
    }
  auto word_at = [&](int64_t v) -> uint32_t {       // the word of (visit v, this lane); 0xffffffff past the end
    if (cached) return sh.words[(v - v0) * 32 + lane];
    const int64_t s = v * 32 + lane;
    return (s < S) ? w[s] : 0xffffffffu;
  };
  for (int i = tid; i < SEL_MAX_E * 1024; i += SEL_THREADS) (&sh.hist[0][0])[i] = 0;
  // ---- per-expert totals from the level-0 histogram ----
  for (int ee = wid; ee < SEL_MAX_E; ee += SEL_THREADS / 32) {
    int c = 0;
    if (ee < E)
      for (int i = lane; i < SEL_HBINS; i += 32) c += __ldcg(&a.zero[ee * SEL_HBINS + i]);
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) sh.cnt[ee] = c;
  }
  __syncthreads();
  const int cap = capacity_of(S, E, a.cf);
  const int keep_cap = a.no_batch ? 0x7fffffff : cap;
  const bool by_rank = !(a.bpr && !a.no_batch);     // plain order: kept iff the index-order rank is below the capacity
  if (tid < SEL_MAX_E) {
    const bool thr = !by_rank && tid < E && sh.cnt[tid] > cap;
    sh.done[tid] = thr ? 0 : 1;
    sh.need[tid] = cap;
    sh.pval[tid] = 0; sh.pmask[tid] = 0;
    sh.Tk[tid] = SEL_KEY_MASK + 1u;      // no threshold: every word of the expert counts as kept
    sh.Ts[tid] = -1;
  }
  if (tid == 0) {
    int row = 0, db = 0;
    for (int ee = 0; ee < E; ++ee) {
      const int kc = min(sh.cnt[ee], keep_cap);
      sh.seg0[ee] = row;
      sh.dropb[ee] = db;
      db += sh.cnt[ee] - kc;
      row += (kc + EP_TILE - 1) / EP_TILE * EP_TILE;
    }
    sh.seg0[E] = row;
  }
  __syncthreads();
  mark(2);
  // ---- thresholds: T[e] = composite value (key << 32 | sample) of the capacity-th element in batch-prioritised order ----
  {
    int sbits = 1;
    while (sbits < 31 && (1ll << sbits) < S) ++sbits;
    for (int lv = 0; lv < SEL_NLEV; ++lv) {
      int dshift, dbits;
      sel_level_digit(lv, sbits, dshift, dbits);
      const int nb = 1 << dbits;
      const unsigned long long dmask = (unsigned long long)nb - 1;
      bool any = false;
      for (int ee = 0; ee < E; ++ee) any |= !sh.done[ee];
      if (!any) break;
      const int* gh = (lv == 0) ? a.zero : a.zero + SEL_Z_LVH + (lv - 1) * SEL_MAX_E * 1024;
      const int gstride = (lv == 0) ? SEL_HBINS : 1024;
      if (lv > 0) {
        lv_used = lv;
        // histogram of this level's digit over MY words that still match their expert's prefix
        for (int64_t v = v0 + wid; v < v1; v += SEL_THREADS / 32) {
          const uint32_t wv = word_at(v);
          const int ee = (int)(wv >> SEL_KEY_BITS);
          bool on = false;
          uint32_t mk = 0x80000000u | (uint32_t)lane;
          if (ee < E && !sh.done[ee]) {
            const unsigned long long c = ((unsigned long long)(wv & SEL_KEY_MASK) << 32) | (unsigned long long)(v * 32 + lane);
            if ((c & sh.pmask[ee]) == sh.pval[ee]) { on = true; mk = ((uint32_t)ee << 10) | (uint32_t)((c >> dshift) & dmask); }
          }
          if (__any_sync(0xffffffffu, on)) {
            const unsigned m = __match_any_sync(0xffffffffu, mk);
            if (on && lane == __ffs(m) - 1) atomicAdd(&sh.hist[mk >> 10][mk & 1023u], __popc(m));
          }
        }
        __syncthreads();
        mark(20 + lv);
        for (int i = tid; i < E * 1024; i += SEL_THREADS) {
          const int v = (&sh.hist[0][0])[i];
          if (v) { atomicAdd(&a.zero[SEL_Z_LVH + (lv - 1) * SEL_MAX_E * 1024 + i], v); (&sh.hist[0][0])[i] = 0; }
        }
        mark(30 + lv);
        sel_grid_barrier(&bar[nbar++], P);
        mark(40 + lv);
      }
      // every CTA picks the digit of every open expert: one warp per expert, bins read 32 at a time (coalesced, all
      // loads of a warp in flight together), running inclusive scan over the rows of 32
      for (int ee = wid; ee < E; ee += SEL_THREADS / 32) {
        if (sh.done[ee]) continue;
        const int need = sh.need[ee];
        int carry = 0, d = -1, hc = 0, before = 0;
        for (int r0 = 0; r0 < nb; r0 += 32 * 8) {
          int v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = r0 + u * 32 + lane;
            v[u] = (i < nb) ? __ldcg(&gh[ee * gstride + i]) : 0;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int inc = carry + warp_incl_scan(v[u], lane);
            const bool here = d < 0 && inc >= need && inc - v[u] < need;
            const unsigned hm = __ballot_sync(0xffffffffu, here);
            if (hm) {
              const int src = __ffs(hm) - 1;
              d = r0 + u * 32 + src;
              hc = __shfl_sync(0xffffffffu, v[u], src);
              before = __shfl_sync(0xffffffffu, inc - v[u], src);
            }
            carry = __shfl_sync(0xffffffffu, inc, 31);
          }
          if (d >= 0) break;
        }
        if (lane == 0) {
          const int need2 = need - before;
          const unsigned long long pv = sh.pval[ee] | ((unsigned long long)d << dshift);
          sh.pval[ee] = pv;
          sh.pmask[ee] |= dmask << dshift;
          sh.need[ee] = need2;
          if (need2 == hc || dshift == 0) {       // the whole digit bucket is kept
            const unsigned long long T = pv | ((1ull << dshift) - 1ull);
            sh.Tk[ee] = (uint32_t)(T >> 32);
            sh.Ts[ee] = (long long)(T & 0xffffffffull);
            sh.done[ee] = 1;
          }
        }
      }
      __syncthreads();
      mark(10 + lv);
    }
  }
  mark(3);
  // key of a word: expert * 2 + (dropped by the threshold); 63 = nobody's
  auto key_of = [&](uint32_t wv, int64_t s) -> int {
    const int ee = (int)(wv >> SEL_KEY_BITS);
    if (ee >= E || s >= S) return 63;
    const uint32_t d = wv & SEL_KEY_MASK;
    const bool kp = (d < sh.Tk[ee]) || (d == sh.Tk[ee] && (long long)s <= sh.Ts[ee]);
    return 2 * ee + (kp ? 0 : 1);
  };
  // ---- totals of my range per key -> global, barrier, my first rank in every bucket ----
  if (tid < 32) { sh.tot[tid] = 0; sh.run[tid] = 0; }
  __syncthreads();
  for (int64_t v = v0 + wid; v < v1; v += SEL_THREADS / 32) {
    const int k = key_of(word_at(v), v * 32 + lane);
    const unsigned m = __match_any_sync(0xffffffffu, k);
    if (k < 32 && lane == __ffs(m) - 1) atomicAdd(&sh.tot[k], __popc(m));
  }
  __syncthreads();
  int* cnt_pe = a.zero + SEL_Z_CNT;
  if (tid < 32) cnt_pe[p * 32 + tid] = sh.tot[tid];
  sel_grid_barrier(&bar[nbar++], P);
  if (tid < 32) {
    int b = 0;
    for (int q = 0; q < p; ++q) b += __ldcg(&cnt_pe[q * 32 + tid]);
    sh.base[tid] = b;
  }
  __syncthreads();
  mark(4);
  // ---- ordered write, in groups of SEL_VIS visits ----
  const bool taps = a.loc || a.idx || a.moe_idx || a.gate;
  for (int64_t g0 = v0; g0 < v1; g0 += SEL_VIS) {
    const int nv = (int)min((int64_t)SEL_VIS, v1 - g0);
    for (int i = tid; i < nv * 16; i += SEL_THREADS) reinterpret_cast<uint32_t*>(&sh.segcnt[0][0])[i] = 0u;
    __syncthreads();
    for (int v = wid; v < nv; v += SEL_THREADS / 32) {
      const int k = key_of(word_at(g0 + v), (g0 + v) * 32 + lane);
      const unsigned m = __match_any_sync(0xffffffffu, k);
      if (k < 32 && lane == __ffs(m) - 1) sh.segcnt[v][k] = (uint16_t)__popc(m);
    }
    __syncthreads();
    {
      // exclusive prefix over the visits, per key: thread (key = tid & 31, part = tid >> 5) owns visits [part*per, +per)
      const int k = tid & 31, part = tid >> 5, per = (nv + 31) / 32;
      const int a0 = part * per, a1 = min(nv, a0 + per);
      int sum = 0;
      for (int v = a0; v < a1; ++v) sum += sh.segcnt[v][k];
      sh.part[part][k] = sum;
      __syncthreads();
      int pre = 0;
      for (int q = 0; q < part; ++q) pre += sh.part[q][k];
      for (int v = a0; v < a1; ++v) { const int c = sh.segcnt[v][k]; sh.segcnt[v][k] = (uint16_t)pre; pre += c; }
      __syncthreads();
    }
    for (int v = wid; v < nv; v += SEL_THREADS / 32) {
      const int64_t s = (g0 + v) * 32 + lane;
      const uint32_t wv = word_at(g0 + v);
      const int k = key_of(wv, s);
      const unsigned m = __match_any_sync(0xffffffffu, k);
      if (k < 32) {
        const int ee = k >> 1;
        int r = sh.base[k] + sh.run[k] + (int)sh.segcnt[v][k] + __popc(m & ((1u << lane) - 1u));
        bool kept = !(k & 1);
        if (by_rank) { kept = r < keep_cap; if (!kept) r -= keep_cap; }
        if (a.tt.row2sample) a.tt.row2sample[kept ? sh.seg0[ee] + r : sh.seg0[E] + sh.dropb[ee] + r] = (int)s;
        if (taps) {
          if (a.loc) a.loc[s] = kept ? r : cap + r;
          if (a.idx) a.idx[s] = ee;
          if (a.moe_idx) a.moe_idx[s] = ee;
          if (a.gate) a.gate[s] = sel_gate(wv);
        }
      }
    }
    __syncthreads();
    if (g0 + SEL_VIS < v1) {
      if (tid < 32) {
        int t = 0;
        for (int q = 0; q < 32; ++q) t += sh.part[q][tid];
        sh.run[tid] += t;
      }
      __syncthreads();
    }
  }
  mark(5);
  // ---- load-balance loss: l_aux = E / S^2 * sum_e me_e * ce_e (tutel_fast_dispatch.py:143-145).  Every CTA folds its
  //      share of the partial column-sum records (fixed order, double); the last CTA to arrive adds the shares in CTA
  //      order, so the result does not depend on timing or on the grid of launch #1 ----
  if (a.l_aux) {
    const int col = tid & (SEL_MAX_E - 1);
    const int r0 = (int)((int64_t)a.npm * p / P), r1 = (int)((int64_t)a.npm * (p + 1) / P);
    double acc = 0.0;
    for (int i = r0 + (tid >> 4); i < r1; i += SEL_THREADS / SEL_MAX_E) acc += (double)a.pm[(int64_t)i * SEL_PM_STRIDE + col];
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);         // lanes c and c+16 of a warp share column c
    if (lane < SEL_MAX_E) sh.red[wid][lane] = acc;
    __syncthreads();
    if (tid < SEL_MAX_E) {
      double me = 0.0;
      for (int ww = 0; ww < 32; ++ww) me += sh.red[ww][tid];
      a.lpart[p * SEL_PM_STRIDE + tid] = me;
    }
  }
  // ---- CTA 0: counts, capacity and the tile table of launch #2 ----
  if (p == 0) {
    if (tid < E && a.counts) a.counts[tid] = sh.cnt[tid];
    if (tid == 0 && a.cap_dev) *a.cap_dev = cap;
    if (a.tt.n_tiles) {
      int row = 0, nt = 0, kept_total = 0;
      for (int ee = 0; ee <= E; ++ee) {
        int kc;
        if (ee < E) { kc = min(sh.cnt[ee], keep_cap); kept_total += kc; }
        else kc = (int)S - kept_total;
        int n = (kc + EP_TILE - 1) / EP_TILE;
        if (a.pair) n = (n + 1) & ~1;
        for (int i = tid; i < n; i += SEL_THREADS) {
          a.tt.tile_expert[nt + i] = (ee < E) ? ee : -1;
          a.tt.tile_row0[nt + i] = row + i * EP_TILE;
          a.tt.tile_rows[nt + i] = max(0, min(EP_TILE, kc - i * EP_TILE));
        }
        if (tid == 0) a.tt.seg_start[ee] = row;
        row += (kc + EP_TILE - 1) / EP_TILE * EP_TILE;
        nt += n;
      }
      if (tid == 0) { *a.tt.n_tiles = nt; *a.tt.drop_counter = 0; }
    }
  }
  // ---- the last CTA to get here combines the l_aux shares and cleans the zero-between-uses region ----
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.flag = (atomicAdd(&a.zero[SEL_Z_TICKET], 1) == P - 1);
  __syncthreads();
  if (sh.flag) {
    __threadfence();
    if (a.l_aux && tid == 0) {
      float tot = 0.f;
      for (int ee = 0; ee < E; ++ee) {
        double me = 0.0;
        for (int c = 0; c < P; ++c) me += ((volatile double*)a.lpart)[c * SEL_PM_STRIDE + ee];
        tot += (float)me * (float)sh.cnt[ee];            // me * ce in fp32
      }
      *a.l_aux = (float)((double)tot * ((double)E / ((double)S * (double)S)));
    }
    if (a.self_clean) {
      for (int i = tid; i < SEL_Z_LVH + lv_used * SEL_MAX_E * 1024; i += SEL_THREADS) a.zero[i] = 0;
    }
  }
  mark(6);
}

int route_select_launch(const SelectArgs& a, cudaStream_t st) {
  SNB_REQUIRE(a.E >= 1 && a.E <= SEL_MAX_E, "route_select: E=%d out of range [1,%d]", a.E, SEL_MAX_E);
  SNB_REQUIRE(a.S >= 1 && a.S < (1ll << 31), "route_select: S=%lld out of range", (long long)a.S);
  static std::atomic<bool> attr_done_dev[64];      // per device; setting the attribute twice is harmless
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done_dev[dev & 63].load(std::memory_order_acquire)) {
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelSmem)));
    attr_done_dev[dev & 63].store(true, std::memory_order_release);
  }
  k_select<<<SEL_P, SEL_THREADS, sizeof(SelSmem), st>>>(a);
  SNB_CHECK_LAUNCH("k_select");
  return SNB_OK;
}

size_t route_select_workspace_bytes(int64_t S) {
  const int64_t Sx = S > 0 ? S : 1;
  return align_up((size_t)Sx * 4, 256) + align_up((size_t)cdiv(Sx, 2048) * SEL_PM_STRIDE * 4, 256) +
         align_up((size_t)SEL_ZERO_INTS * 4, 256) + align_up((size_t)SEL_MAX_E * SEL_PM_STRIDE * 8, 256) + 1024;
}

// hist0 must be zero on entry (the caller enqueues the memset)
int route_pack_top1(const float* gates, int64_t S, int32_t E, uint32_t* w, int* hist0, float* pm, int* npm, cudaStream_t st) {
  const int nblk = (int)cdiv(S, 2048);
  k_pack_top1<<<nblk, 256, 0, st>>>(gates, S, E, w, hist0, pm);
  SNB_CHECK_LAUNCH("k_pack_top1");
  *npm = nblk;
  return SNB_OK;
}

// stand-alone form (gates in global memory): pack + select.  Used by the parity tests that compare the kept set with
// the full-order routing (snb_route_select through the C ABI).
int route_select_from_gates(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int32_t no_batch,
                            int32_t* idx, int32_t* loc, float* gate, int32_t* counts, int32_t* capacity, float* l_aux,
                            void* ws, size_t ws_bytes, cudaStream_t st) {
  SNB_REQUIRE(E >= 1 && E <= SEL_MAX_E, "route_select: E=%d out of range [1,%d]", E, SEL_MAX_E);
  SNB_REQUIRE(S >= 1 && S < (1ll << 31), "route_select: S=%lld out of range", (long long)S);
  Arena ar(ws, ws_bytes);
  uint32_t* w = ar.take<uint32_t>(S);
  const int nblk = (int)cdiv(S, 2048);
  float* pm = ar.take<float>((size_t)nblk * SEL_PM_STRIDE);
  int* hist0 = ar.take<int>((size_t)SEL_ZERO_INTS);
  double* lpart = ar.take<double>((size_t)SEL_MAX_E * SEL_PM_STRIDE);
  if (!ar.ok) { set_error("route_select: workspace too small (%zu bytes given)", ws_bytes); return SNB_EWORKSPACE; }
  SNB_CHECK_CUDA(cudaMemsetAsync(hist0, 0, (size_t)SEL_ZERO_INTS * sizeof(int), st));
  SelectArgs a = {};
  int rc = route_pack_top1(gates, S, E, w, hist0, pm, &a.npm, st);
  if (rc) return rc;
  a.w = w; a.pm = pm; a.zero = hist0; a.lpart = lpart;
  a.S = S; a.E = E; a.cf = cf; a.bpr = bpr; a.no_batch = no_batch;
  a.idx = idx; a.loc = loc; a.gate = gate; a.counts = counts; a.cap_dev = capacity; a.l_aux = l_aux;
  return route_select_launch(a, st);
}

}  // namespace snb

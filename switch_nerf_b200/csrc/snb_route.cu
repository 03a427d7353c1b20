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
// One CTA per expert, ONE launch: a counting pass (all experts' totals + this expert's histogram of the top 9 key
// bits), radix descent to the composite threshold T over the unique 58-bit value (key << 32 | sample) -- ties
// between equal gates resolve to the lower sample index exactly as the stable sort of the full path does -- then ONE
// ordered pass that hands every kept sample its row (index order) and every dropped sample a row of the dropped
// bucket.  No atomics on the output, no memset, deterministic row order; CTA 0 also writes counts, capacity, l_aux
// and the tile table of launch #2.
// reference: tutel_fast_dispatch.py:136-139, 176-217 (the kept set == {s : locations_s < capacity}).
// =========================================================================================
static constexpr int SEL_THREADS = 1024;
static constexpr int SEL_LIST = 3072;       // candidates of the threshold digit bucket resolved in shared memory
static_assert(SEL_MAX_E <= 32, "one lane per expert in the count reduction");

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

static constexpr int SEL_SEG = 128;          // samples per warp-visit of the ordered passes (4 per lane)
static constexpr int SEL_SEG_SMEM = 2048;    // segments whose (kept, dropped) counts are scanned in shared memory
static_assert((SEL_SEG_SMEM + 32 * SEL_SEG) * 4 <= SEL_LIST * 8, "ordered-pass scratch aliases the candidate list");

struct SelShared {
  int cnt[SEL_MAX_E];
  int hist[1024];                 // digit histogram of the current level (<= 10 bits)
  int wtot[32];
  double red[32][SEL_MAX_E];
  unsigned long long list[SEL_LIST];
  uint32_t seg[SEL_SEG_SMEM];       // per 128-sample segment: kept count | dropped count << 16, then exclusive prefixes
  int n_list;
  int ch_d, ch_need, ch_hc;       // digit chosen by the current level
};

// visit the words of this chunk, 4 per 16-byte load, 4 independent loads in flight per thread.  f(word, sample, valid)
// is called the same number of times by every lane of a warp (warp-collective code is allowed inside f).
template <typename F>
__device__ __forceinline__ void sel_for_each(const uint32_t* __restrict__ w, int64_t S, int tid, F&& f) {
  const int64_t nq = S >> 2;
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
  for (int64_t q0 = 0; q0 < nq; q0 += 4 * SEL_THREADS) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t q = q0 + u * SEL_THREADS + tid;
      v[u] = (q < nq) ? w4[q] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t q = q0 + u * SEL_THREADS + tid;
      const bool ok = q < nq;
      f(v[u].x, 4 * q, ok); f(v[u].y, 4 * q + 1, ok); f(v[u].z, 4 * q + 2, ok); f(v[u].w, 4 * q + 3, ok);
    }
  }
  const int64_t s = (nq << 2) + tid;
  f(s < S ? w[s] : 0u, s, s < S);
}

// smallest digit d with (inclusive prefix count up to d) >= need; returns through shared memory (all threads sync)
__device__ __forceinline__ void sel_choose(SelShared& sh, int nbins, int need, int tid, int lane, int wid) {
  const int v = (tid < nbins) ? sh.hist[tid] : 0;
  int inc = warp_incl_scan(v, lane);
  if (lane == 31) sh.wtot[wid] = inc;
  __syncthreads();
  int base = 0;
  for (int ww = 0; ww < wid; ++ww) base += sh.wtot[ww];
  inc += base;
  if (tid < nbins && inc >= need && inc - v < need) { sh.ch_d = tid; sh.ch_need = need - (inc - v); sh.ch_hc = v; }
  __syncthreads();
}

__global__ void __launch_bounds__(SEL_THREADS) k_select(SelectArgs a) {
  __shared__ SelShared sh;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int E = a.E;
  const int64_t S = a.S;
  const uint32_t* __restrict__ w = a.w;
  int tn = 0;
  auto mark = [&](int tag) {
    if (a.tl && e == 0 && tid == 0) a.tl[tn++] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
  };
  mark(1);
  if (tid == 0) sh.n_list = 0;
  // ---- per-expert totals (segment starts need all of them) + this expert's level-0 histogram, both from the
  //      global histogram launch #1 accumulated with fire-and-forget reductions ----
  for (int ee = wid; ee < E; ee += SEL_THREADS / 32) {
    int c = 0;
    for (int i = lane; i < SEL_HBINS; i += 32) {
      const int v = a.hist0[ee * SEL_HBINS + i];
      if (ee == e) sh.hist[i] = v;
      c += v;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) sh.cnt[ee] = c;
  }
  __syncthreads();
  mark(2);
  const int cap = capacity_of(S, E, a.cf);
  const int keep_cap = a.no_batch ? 0x7fffffff : cap;
  const int cnt = sh.cnt[e];
  int seg0 = 0, drop0 = 0, drop_before = 0;        // first row of my segment / of the dropped segment; dropped rows of experts < e
  {
    int row = 0;
    for (int ee = 0; ee < E; ++ee) {
      const int kc = min(sh.cnt[ee], keep_cap);
      if (ee == e) seg0 = row;
      if (ee < e) drop_before += sh.cnt[ee] - kc;
      row += (kc + EP_TILE - 1) / EP_TILE * EP_TILE;
    }
    drop0 = row;
  }
  // ---- threshold: T = composite value (key << 32 | sample) of the keep_cap-th element in batch-prioritised order ----
  unsigned long long T = ~0ull;
  if (a.bpr && !a.no_batch && cnt > cap) {
    int sbits = 1;
    while (sbits < 32 && (1ll << sbits) < S) ++sbits;
    unsigned long long pmask = 0, pval = 0;
    int need = cap, level = 0;
    bool from_list = false, have_hist = true;
    while (true) {
      int dshift, dbits;
      if (level == 0) { dshift = 32 + SEL_L1_SHIFT; dbits = SEL_KEY_BITS - SEL_L1_SHIFT; }
      else if (level == 1) { dshift = 32 + SEL_L2_SHIFT; dbits = SEL_L1_SHIFT - SEL_L2_SHIFT; }
      else if (level == 2) { dshift = 32; dbits = SEL_L2_SHIFT; }
      else { const int hi = sbits - 9 * (level - 3); dshift = hi > 9 ? hi - 9 : 0; dbits = hi - dshift; }
      const unsigned long long dmask = (1ull << dbits) - 1;
      if (!have_hist) {
        for (int i = tid; i < 1024; i += SEL_THREADS) sh.hist[i] = 0;
        __syncthreads();
        // candidates are few here (one digit bucket of the level above): plain shared atomics
        if (from_list) {
          for (int i = tid; i < sh.n_list; i += SEL_THREADS) {
            const unsigned long long c = sh.list[i];
            if ((c & pmask) == pval) atomicAdd(&sh.hist[(int)((c >> dshift) & dmask)], 1);
          }
        } else {
          sel_for_each(w, S, tid, [&](uint32_t wv, int64_t s, bool ok) {
            if (!ok || (int)(wv >> SEL_KEY_BITS) != e) return;
            const unsigned long long c = ((unsigned long long)(wv & SEL_KEY_MASK) << 32) | (unsigned long long)s;
            if ((c & pmask) == pval) atomicAdd(&sh.hist[(int)((c >> dshift) & dmask)], 1);
          });
        }
        __syncthreads();
      }
      sel_choose(sh, 1 << dbits, need, tid, lane, wid);
      mark(10 + level);
      const int d = sh.ch_d, hc = sh.ch_hc;
      need = sh.ch_need;
      pval |= (unsigned long long)d << dshift;
      pmask |= dmask << dshift;
      if (need == hc || dshift == 0) { T = pval | ((1ull << dshift) - 1); break; }   // the whole digit bucket is kept
      if (!from_list && hc <= SEL_LIST) {
        // the undecided bucket fits in shared memory: gather it once, finish the descent there
        // (only reached from level 0: the prefix is the level-0 digit of the key)
        const uint32_t want = (((uint32_t)e << SEL_KEY_BITS) >> SEL_L1_SHIFT) | (uint32_t)(pval >> (32 + SEL_L1_SHIFT));
        const bool narrow = (pmask == (((1ull << (SEL_KEY_BITS - SEL_L1_SHIFT)) - 1) << (32 + SEL_L1_SHIFT)));
        sel_for_each(w, S, tid, [&](uint32_t wv, int64_t s, bool ok) {
          bool hit;
          if (narrow) hit = ok && (wv >> SEL_L1_SHIFT) == want;
          else {
            const unsigned long long c = ((unsigned long long)(wv & SEL_KEY_MASK) << 32) | (unsigned long long)s;
            hit = ok && (int)(wv >> SEL_KEY_BITS) == e && (c & pmask) == pval;
          }
          const unsigned long long c = ((unsigned long long)(wv & SEL_KEY_MASK) << 32) | (unsigned long long)s;
          const unsigned m = __ballot_sync(0xffffffffu, hit);       // one shared atomic per warp, not per candidate
          if (m) {
            int base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(&sh.n_list, __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (hit) sh.list[base + __popc(m & ((1u << lane) - 1u))] = c;
          }
        });
        from_list = true;
      }
      have_hist = false;
      ++level;
      __syncthreads();
    }
  }
  mark(3);
  // ---- ordered write: rows for the kept samples (index order) and for the dropped ones.  Two sweeps over the words,
  //      one warp per 128-sample segment and no block barrier inside a sweep: (1) kept / dropped counts per segment,
  //      (block-wide exclusive scan), (2) rows = segment base + ballot rank.  Chunks larger than SEL_SEG_SMEM
  //      segments are processed in groups with running bases. ----
  {
    const bool by_rank = !(a.bpr && !a.no_batch);     // plain order: kept iff the index-order rank is below the capacity
    const int64_t nseg = (S + SEL_SEG - 1) / SEL_SEG;
    int run_keep = 0, run_drop = 0;                   // block-uniform running totals (by_rank: run_keep = rank base)
    auto load_seg = [&](int64_t g, uint32_t (&wv)[4]) {
      const int64_t s0 = g * SEL_SEG + 4 * lane;
      wv[0] = wv[1] = wv[2] = wv[3] = 0xffffffffu;
      if (s0 + 3 < S) {
        const uint4 q = *reinterpret_cast<const uint4*>(w + s0);
        wv[0] = q.x; wv[1] = q.y; wv[2] = q.z; wv[3] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (s0 + j < S) wv[j] = w[s0 + j];
      }
    };
    // kept iff mine && (key, sample) <= T.  32-bit form: d = word - (e << 26) is the key when the word is mine and
    // >= 2^26 otherwise; d < Tk: kept, d > Tk: dropped, d == Tk (ties of the threshold key, rare): by sample index.
    // Padding words (0xffffffff) are nobody's: the host requires E < 63.
    const uint32_t ebase = (uint32_t)e << SEL_KEY_BITS;
    const bool keep_all = by_rank || T == ~0ull;       // no threshold: every word of mine is "kept" at this stage
    const uint32_t Tk = keep_all ? (SEL_KEY_MASK + 1u) : (uint32_t)(T >> 32);
    const long long Ts = keep_all ? -1 : (long long)(T & 0xffffffffull);
    // per lane: bit j of `kp` / `dr` = word j of this lane is mine and kept / mine and dropped
    auto classify = [&](int g, const uint32_t (&wv)[4], uint32_t& kp, uint32_t& dr) {
      const long long lim = Ts - ((long long)g * SEL_SEG + 4 * lane);        // tie at sample s0 + j kept iff j <= lim
      kp = dr = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t d = wv[j] - ebase;
        const bool k = (d < Tk) || (d == Tk && (long long)j <= lim);
        kp |= (k ? 1u : 0u) << j;
        dr |= ((d <= SEL_KEY_MASK && !k) ? 1u : 0u) << j;
      }
    };
    // packed (kept | dropped << 16) counts of the lanes below this one, and the segment totals
    auto lane_prefix = [&](uint32_t kp, uint32_t dr, uint32_t& total) {
      const uint32_t mine = (uint32_t)__popc(kp) | ((uint32_t)__popc(dr) << 16);
      uint32_t inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      total = __shfl_sync(0xffffffffu, inc, 31);
      return inc - mine;
    };
    const bool taps = a.loc || a.idx || a.moe_idx || a.gate;
    constexpr int NW = SEL_THREADS / 32, UF = 4;      // UF segments in flight per warp (independent 16-byte loads)
    for (int64_t g0 = 0; g0 < nseg; g0 += SEL_SEG_SMEM) {
      const int ng = (int)min((int64_t)SEL_SEG_SMEM, nseg - g0);
      for (int gb = wid; gb < ng; gb += NW * UF) {
        uint32_t wv[UF][4];
#pragma unroll
        for (int u = 0; u < UF; ++u) if (gb + u * NW < ng) load_seg(g0 + gb + u * NW, wv[u]);
#pragma unroll
        for (int u = 0; u < UF; ++u) {
          const int g = gb + u * NW;
          if (g >= ng) break;
          uint32_t kp, dr, total;
          classify((int)g0 + g, wv[u], kp, dr);
          lane_prefix(kp, dr, total);
          if (lane == 0) sh.seg[g] = total;
        }
      }
      __syncthreads();
      // exclusive scan of the (<= 2048) packed counts: 2 per thread
      {
        const uint32_t v0 = (2 * tid < ng) ? sh.seg[2 * tid] : 0u, v1 = (2 * tid + 1 < ng) ? sh.seg[2 * tid + 1] : 0u;
        int k = (int)((v0 & 0xffffu) + (v1 & 0xffffu)), d = (int)((v0 >> 16) + (v1 >> 16));
        int ki = warp_incl_scan(k, lane), di = warp_incl_scan(d, lane);
        if (lane == 31) { sh.wtot[wid] = ki; sh.hist[wid] = di; }
        __syncthreads();
        int kb = 0, db = 0;
        for (int ww = 0; ww < wid; ++ww) { kb += sh.wtot[ww]; db += sh.hist[ww]; }
        int ktot = 0, dtot = 0;
        for (int ww = 0; ww < 32; ++ww) { ktot += sh.wtot[ww]; dtot += sh.hist[ww]; }
        __syncthreads();
        const int ke = run_keep + kb + ki - k, de = run_drop + db + di - d;     // exclusive prefixes of segment 2*tid
        // kept prefix goes back into seg[], dropped prefix into the (now unused) candidate list
        int* dpre = reinterpret_cast<int*>(sh.list);
        if (2 * tid < ng) { sh.seg[2 * tid] = (uint32_t)ke; dpre[2 * tid] = de; }
        if (2 * tid + 1 < ng) { sh.seg[2 * tid + 1] = (uint32_t)(ke + (int)(v0 & 0xffffu)); dpre[2 * tid + 1] = de + (int)(v0 >> 16); }
        run_keep += ktot;
        run_drop += dtot;
        __syncthreads();
      }
      const int* dpre = reinterpret_cast<const int*>(sh.list);
      int* stage = reinterpret_cast<int*>(sh.list) + SEL_SEG_SMEM + wid * SEL_SEG;     // per-warp, behind dpre[]
      for (int gb = wid; gb < ng; gb += NW * UF) {
        uint32_t wv[UF][4];
#pragma unroll
        for (int u = 0; u < UF; ++u) if (gb + u * NW < ng) load_seg(g0 + gb + u * NW, wv[u]);
#pragma unroll
        for (int u = 0; u < UF; ++u) {
          const int g = gb + u * NW;
          if (g >= ng) break;
          uint32_t kp, dr, total;
          classify((int)g0 + g, wv[u], kp, dr);
          const uint32_t pre = lane_prefix(kp, dr, total);
          int kb = (int)sh.seg[g], db = dpre[g];       // first kept / dropped slot of this segment
          int nk = (int)(total & 0xffffu), nd = (int)(total >> 16);
          int kl = (int)(pre & 0xffffu), dl = (int)(pre >> 16);     // my first kept / dropped rank inside the segment
          if (by_rank) {
            // every word of mine came out as "kept" with its index-order rank kb + kl; the capacity cuts the ranks
            const int cut = max(0, min(nk, keep_cap - kb));          // ranks [0, cut) of this segment are kept
            dr = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (((kp >> j) & 1u) && kl + __popc(kp & ((1u << j) - 1u)) >= cut) { dr |= 1u << j; }
            kp &= ~dr;
            dl = max(0, kl - cut);
            kl = min(kl, cut);
            db = max(kb, keep_cap) - keep_cap;
            nd = nk - cut;
            nk = cut;
          }
          const int s_lane = ((int)g0 + g) * SEL_SEG + 4 * lane;
          // rows of one segment are consecutive per class: stage the sample ids in shared memory by rank (kept from
          // the front, dropped from the back), the warp stores them coalesced below
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if ((kp >> j) & 1u) stage[kl + __popc(kp & ((1u << j) - 1u))] = s_lane + j;
            if ((dr >> j) & 1u) stage[SEL_SEG - 1 - (dl + __popc(dr & ((1u << j) - 1u)))] = s_lane + j;
          }
          if (taps) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool k = (kp >> j) & 1u, d = (dr >> j) & 1u;
              if (k || d) {
                const int s = s_lane + j;
                const int slot = k ? kb + kl + __popc(kp & ((1u << j) - 1u)) : db + dl + __popc(dr & ((1u << j) - 1u));
                if (a.loc) a.loc[s] = k ? slot : cap + slot;
                if (a.idx) a.idx[s] = e;
                if (a.moe_idx) a.moe_idx[s] = e;
                if (a.gate) a.gate[s] = sel_gate(wv[u][j]);
              }
            }
          }
          if (a.tt.row2sample) {
            __syncwarp();
            if (lane < nk) a.tt.row2sample[seg0 + kb + lane] = stage[lane];
            for (int i = lane + 32; i < nk; i += 32) a.tt.row2sample[seg0 + kb + i] = stage[i];
            if (lane < nd) a.tt.row2sample[drop0 + drop_before + db + lane] = stage[SEL_SEG - 1 - lane];
            for (int i = lane + 32; i < nd; i += 32) a.tt.row2sample[drop0 + drop_before + db + i] = stage[SEL_SEG - 1 - i];
            __syncwarp();
          }
        }
      }
      __syncthreads();
    }
  }
  mark(4);
  // ---- load-balance loss: l_aux = E / S^2 * sum_e me_e * ce_e (tutel_fast_dispatch.py:143-145).  Every CTA folds its
  //      share of the partial column-sum records (fixed order, double); the last CTA to arrive adds the shares in CTA
  //      order, so the result does not depend on timing or on the grid of launch #1 ----
  if (a.l_aux) {
    const int col = tid & (SEL_MAX_E - 1);
    const int r0 = (int)((int64_t)a.npm * e / E), r1 = (int)((int64_t)a.npm * (e + 1) / E);
    double acc = 0.0;
    for (int i = r0 + (tid >> 4); i < r1; i += SEL_THREADS / SEL_MAX_E) acc += (double)a.pm[(int64_t)i * SEL_PM_STRIDE + col];
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);         // lanes c and c+16 of a warp share column c
    if (lane < SEL_MAX_E) sh.red[wid][lane] = acc;
    __syncthreads();
    if (tid < SEL_MAX_E) {
      double me = 0.0;
      for (int ww = 0; ww < 32; ++ww) me += sh.red[ww][tid];
      a.lpart[e * SEL_PM_STRIDE + tid] = me;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) sh.ch_d = (atomicAdd(a.ticket, 1) == E - 1);
    __syncthreads();
    if (sh.ch_d && tid == 0) {
      __threadfence();
      float tot = 0.f;
      for (int ee = 0; ee < E; ++ee) {
        double me = 0.0;
        for (int c = 0; c < E; ++c) me += ((volatile double*)a.lpart)[c * SEL_PM_STRIDE + ee];
        tot += (float)me * (float)sh.cnt[ee];            // me * ce in fp32
      }
      *a.l_aux = (float)((double)tot * ((double)E / ((double)S * (double)S)));
      *a.ticket = 0;
    }
  }
  // ---- CTA 0: counts, capacity and the tile table of launch #2 ----
  if (e == 0) {
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
  mark(5);
}

int route_select_launch(const SelectArgs& a, cudaStream_t st) {
  SNB_REQUIRE(a.E >= 1 && a.E <= SEL_MAX_E, "route_select: E=%d out of range [1,%d]", a.E, SEL_MAX_E);
  SNB_REQUIRE(a.S >= 1 && a.S < (1ll << 31), "route_select: S=%lld out of range", (long long)a.S);
  SNB_REQUIRE((reinterpret_cast<uintptr_t>(a.w) & 15) == 0, "route_select: packed words must be 16-byte aligned");
  k_select<<<a.E, SEL_THREADS, 0, st>>>(a);
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
  a.w = w; a.pm = pm; a.hist0 = hist0; a.ticket = hist0 + SEL_MAX_E * SEL_HBINS; a.lpart = lpart;
  a.S = S; a.E = E; a.cf = cf; a.bpr = bpr; a.no_batch = no_batch;
  a.idx = idx; a.loc = loc; a.gate = gate; a.counts = counts; a.cap_dev = capacity; a.l_aux = l_aux;
  return route_select_launch(a, st);
}

}  // namespace snb

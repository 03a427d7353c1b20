// Launch #2, "TS" variant: hidden activations live in TENSOR MEMORY instead of shared memory.
// (included by snb_tc.cu after k_back; shares its constants, TcParams, packed weight images and epilogue helpers)
//
// Why (profiles/r1i_*): in k_back a layer costs ~5.7K clk against 2K clk of MMA work, and the time goes to the
// epilogue, not the tensor pipe: writing the next A operand to shared memory (st.shared + fence.proxy.async per
// 64-column chunk) paces the chunks at ~950 clk each, and the MMAs read that A back from shared memory at 171 clk
// per 128x256x16 instruction (floor 128).  Here the epilogue packs its bf16 result with tcgen05.st straight into
// tensor memory (~340 clk per chunk, no proxy fence) and the MMA takes the A operand from there (138 clk).
//
// TMEM map: two 256-column buffers B0 / B1 as in k_back; layer li accumulates into B[li & 1].  Its epilogue drains
// the fp32 accumulator and writes the packed bf16 result IN PLACE: the thread that read fp32 columns
// [64c + 16t, +16) of its row writes the 16 bf16 values as 8 packed columns [64c + 16t, +8) -- exactly the 8
// columns one K=16 tcgen05.mma reads as its A operand, so the operand of instruction (chunk c, step t) of layer
// li+1 sits at column 64c + 16t of B[li & 1] while that layer accumulates into the other buffer.  A thread only
// ever overwrites columns it has itself already loaded: no cross-warp hazard, no barrier.
// Tried and rejected (profiles/): a 64x64-block schedule with N=64 instructions (a tcgen05.mma costs ~123 clk for
// any N <= 128, r1i_umma_microbench.json) and cta_group::2 pairs of this kernel (pair MMAs ran at ~380 clk, r1k).
//
// The PE(xyz) / [PE(dir) | appearance] columns stay in shared memory (2 x 24 KB, alternating per tile so the next
// tile's PE(xyz) is staged while this tile waits for the tensor pipe); the weight ring has 5 x 32 KB stages.
#pragma once

static constexpr uint32_t TS_CAT_COLS = 96;                       // PE(xyz) / [PE(dir) | appearance] block
static constexpr uint32_t TS_SBO = TS_CAT_COLS / 8 * 128;         // 1536 B between 8-row groups
static constexpr uint32_t TS_ACAT_BYTES = TILE / 8 * TS_SBO;      // 24576
static constexpr int TS_NST = 5;
#ifndef TS_BACK_NST
#define TS_BACK_NST 5
#endif
static constexpr int TS_NST_MAX = 6;

struct __align__(16) TsCtl {
  uint64_t full[2 * TS_NST_MAX];      // CTA pairs use 2*NST half-size ring slots (each CTA streams half of a slice)
  uint64_t empty[2 * TS_NST_MAX];
  uint64_t peer_ok[2 * TS_NST_MAX];   // CTA pairs, leader only: the peer's half of the weight slice landed
  uint64_t acc_full[2];
  uint64_t a_ready[4];      // epilogue -> MMA: K-chunk of the next A operand packed into TMEM
  uint64_t s_ready[2][2];   // epilogue -> MMA: shared-memory K-chunk written, [cat block][chunk]
  uint32_t tmem_base;
  uint32_t pad;
};
static constexpr size_t TSM_ACAT = 0;
static constexpr size_t TSM_RING = 2 * TS_ACAT_BYTES;      // two cat blocks: tile t+1's PE(xyz) is staged during tile t
static constexpr size_t TSM_BIAS = TSM_RING + (size_t)TS_NST * STAGE_BYTES;
static constexpr size_t TSM_VEC = TSM_BIAS + 2 * 256 * 4;
static constexpr size_t TSM_RED = TSM_VEC + SM_VEC_FLOATS * 4;
static constexpr size_t TSM_CTL = TSM_RED + SM_RED_FLOATS * 4;
static constexpr size_t TSM_TOTAL = TSM_CTL + sizeof(TsCtl) + 1024;
static_assert(TSM_TOTAL <= 227 * 1024, "shared memory budget (TS kernel)");
// launch #2 may use a deeper ring (TS_BACK_NST) -- timing diagnostic for now: with 6 slots the two cat blocks and the
// reduction scratch alias (results are garbage, the schedule is that of a 6-slot ring)
static constexpr bool TSB_DIAG = (TS_BACK_NST > 5);
static constexpr uint32_t TSB_CAT_STRIDE = TSB_DIAG ? 0u : TS_ACAT_BYTES;
static constexpr size_t TSB_RING = TSB_DIAG ? TS_ACAT_BYTES : 2 * TS_ACAT_BYTES;
static constexpr size_t TSB_BIAS = TSB_RING + (size_t)TS_BACK_NST * STAGE_BYTES;
static constexpr size_t TSB_VEC = TSB_DIAG ? TSB_BIAS : TSB_BIAS + 2 * 256 * 4;
static constexpr size_t TSB_RED = TSB_DIAG ? TSB_BIAS : TSB_VEC + SM_VEC_FLOATS * 4;
static constexpr size_t TSB_CTL = TSB_RED + SM_RED_FLOATS * 4;
static constexpr size_t TSB_TOTAL = TSB_CTL + sizeof(TsCtl) + 1024;
static_assert(TSB_TOTAL <= 227 * 1024, "shared memory budget (TS kernel, launch #2)");

struct TsPipe {
  uint32_t stage = 0, phase = 0;       // ring position of the next weight slot to produce / consume
  uint32_t a_use[4] = {0, 0, 0, 0};
  uint32_t s_use[2][2] = {{0, 0}, {0, 0}};
  uint32_t acc_use[2] = {0, 0};
};

__device__ __forceinline__ uint32_t ts_cat_addr(uint32_t base, int row, int col8) {
  return base + (uint32_t)(row >> 3) * TS_SBO + (uint32_t)col8 * 128u + (uint32_t)(row & 7) * 16u;
}
__device__ __forceinline__ void ts_cat_store_row(uint32_t base, int row, const __nv_bfloat16* vals, int n8) {
  const uint4* v = reinterpret_cast<const uint4*>(vals);
  for (int g = 0; g < n8; ++g) {
    const uint4 t = v[g];
    st_shared_v4(ts_cat_addr(base, row, g), t.x, t.y, t.z, t.w);
  }
}
// tcgen05.wait::ld that also "produces" the loaded registers, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// round two fp32 values to bf16 (RN, optional ReLU after rounding): returns the packed pair (lo = a) and the rounded
// values as fp32 -- one cvt instead of two round trips
template <bool RELU>
__device__ __forceinline__ uint32_t ts_round2(float a, float b, float& ra, float& rb) {
  const uint32_t p = pack2<RELU>(a, b);
  ra = __uint_as_float(p << 16);
  rb = __uint_as_float(p & 0xffff0000u);
  return p;
}

// ---- producer: the K-slices of one packed layer image (same images as k_back) ----
// CG = 2 (CTA pair): each CTA streams its own half of the slice (output rows [rank*N/2, +N/2) are the contiguous half)
// A ring slot holds as many consecutive 64-wide K-slices as fit in it (1 for N = 256, 2 for N = 128, 8 for the 32-wide
// gate GEMM): the cost of a bulk copy hardly depends on its size (profiles/r3a_ring_experiments.md), so narrow layers
// are fed by full-size copies too.  Segments of a layer start on a group boundary (K = 256 is 4 slices).
__device__ __forceinline__ uint32_t ts_group(uint32_t N, int CG) {
#ifdef TS_NO_GROUP
  return 1u;
#else
  const uint32_t g = STAGE_BYTES / (N * 64u * 2u);
  return (CG == 2 || g < 1u) ? 1u : g;
#endif
}
template <int CG, int NST = TS_NST, bool BACK = false>
__device__ __forceinline__ void ts_produce(const uint8_t* wsrc, uint32_t N, uint32_t K16, uint8_t* ring, TsCtl* ctl, TsPipe& pp,
                                           uint32_t rank) {
  const uint32_t G = ts_group(N, CG);
  const uint32_t nsl = (K16 + 63) / 64;
  constexpr uint32_t NS = NST * CG, SB = STAGE_BYTES / CG;
  for (uint32_t j = 0; j < nsl; j += G) {
    const uint32_t klen = min(64u * G, K16 - 64u * j);
    const uint32_t bytes = N * klen * 2 / CG;
    const uint32_t stage = pp.stage;
    mbar_wait(&ctl->empty[stage], pp.phase ^ 1);
    if (++pp.stage == NS) { pp.stage = 0; pp.phase ^= 1; }
#if defined(TS_DIAG_NOCOPY)      // timing diagnostic, launch #2 only: the weights "arrive" instantly (results are garbage)
    if (BACK) { mbar_arrive(&ctl->full[stage]); continue; }
#endif
    mbar_arrive_expect_tx(&ctl->full[stage], bytes);
    bulk_g2s(ring + (size_t)stage * SB, wsrc + (size_t)N * 64 * 2 * j + (size_t)rank * bytes, bytes, &ctl->full[stage]);
  }
}

// ---- MMA issuer: one segment of a layer = consecutive K-slices whose A operand is all in TMEM (ts) or all in the
// shared-memory cat block.  `a_tmem`: first column of the packed A operand; `cont`: accumulate onto an earlier segment.
// Executed by ALL lanes of warp 1 with warp-uniform values; only the tcgen05.mma / tcgen05.commit themselves are
// predicated on `leader` (one elected lane, the same for the whole kernel).  The issuing thread is what bounds the
// tensor pipe here: written for lane 0 alone, every instruction carried a vector->uniform register waterfall
// (ELECT + 4 x R2UR + BRA.U.ANY) and a slice cost ~830 clk of issue for 552 clk of tensor work
// (profiles/r3g_issue_path.md).  Uniform control flow lets the descriptors live in uniform registers.
// CG = 2: the leader CTA (rank 0) issues M = 256 instructions for the pair (each CTA's own TMEM / cat block supplies
// its 128 rows of A, each CTA's ring slot half of B); the peer's warp 1 only forwards "my weight half landed".
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int CG, int NST = TS_NST, uint32_t CAT_STRIDE = TS_ACAT_BYTES>
__device__ __forceinline__ void ts_mma_seg(uint32_t N, uint32_t K, bool ts, uint32_t a_tmem, uint32_t acat_base,
                                           uint32_t ring_base, uint32_t d_tmem, TsCtl* ctl, TsPipe& pp, bool cont,
                                           uint32_t rank, bool leader, unsigned long long* tl, int* tn, uint32_t cb = 0) {
  const uint32_t nsl = (K + 63) / 64;
  const uint32_t idesc = umma_idesc_bf16(TILE * CG, (int)N);
  const uint32_t G = ts_group(N, CG);
  constexpr uint32_t NS = NST * CG, SB = STAGE_BYTES / CG;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) {
    if (j >= nsl) break;
    const uint32_t klen = min(64u, K - 64u * j);
    const uint32_t stage = pp.stage, phase = pp.phase;
    const bool first = (j % G) == 0, last = ((j + 1) % G) == 0 || j + 1 == nsl;
    if (last) { if (++pp.stage == NS) { pp.stage = 0; pp.phase ^= 1; } }
    if (CG == 2 && rank != 0) {
      mbar_wait(&ctl->full[stage], phase);
      if (leader) mbar_arrive_remote(mapa_shared(smem_u32(&ctl->peer_ok[stage]), 0));
      continue;
    }
    if (ts) { mbar_wait(&ctl->a_ready[j], pp.a_use[j] & 1); ++pp.a_use[j]; }
    else if (j < 2) { mbar_wait(&ctl->s_ready[cb][j], pp.s_use[cb][j] & 1); ++pp.s_use[cb][j]; }
    if (tn) tl_mark(tl, 1, *tn, 100 + (int)j);
    if (first) {
      mbar_wait(&ctl->full[stage], phase);
      if (CG == 2) mbar_wait(&ctl->peer_ok[stage], phase);
    }
    if (tn) tl_mark(tl, 1, *tn, 110 + (int)j);
    tc_fence_after();
    // descriptors of the slice: the start-address field counts 16-byte units, one K = 16 step is two 128-byte core
    // matrices further (+16); addresses stay below 256 KB, so the 14-bit field never carries
    const uint64_t db0 = op_desc(ring_base + stage * SB + (j % G) * (N * 128u), 128u, klen * 16u);
    const uint64_t da0 = op_desc(acat_base + cb * CAT_STRIDE + (8u * j) * 128u, 128u, TS_SBO);
    const uint32_t at0 = a_tmem + j * 64u;
    const uint32_t acc0 = (cont || j) ? 1u : 0u;
    auto issue = [&](uint32_t t) {
      const uint32_t acc = t ? 1u : acc0;
      if (ts) {
        if (CG == 2) umma_bf16_ts_pair(d_tmem, at0 + t * 16u, db0 + 16u * t, idesc, acc);
        else umma_bf16_ts(d_tmem, at0 + t * 16u, db0 + 16u * t, idesc, acc);
      } else {
        if (CG == 2) umma_bf16_pair(d_tmem, da0 + 16u * t, db0 + 16u * t, idesc, acc);
        else umma_bf16(d_tmem, da0 + 16u * t, db0 + 16u * t, idesc, acc);
      }
    };
    if (leader) {
      if (klen == 64u) {
        issue(0); issue(1); issue(2); issue(3);
      } else {
        for (uint32_t t = 0; t < klen / 16; ++t) issue(t);
      }
      if (last) {
        if (CG == 2) umma_commit_pair(&ctl->empty[stage], 3);    // frees the ring slot in BOTH CTAs
        else umma_commit(&ctl->empty[stage]);
      }
    }
  }
}
template <int CG>
__device__ __forceinline__ void ts_commit_acc(TsCtl* ctl, uint32_t buf, uint32_t rank, bool leader) {
  if (!leader) return;
  if (CG == 2) {
    if (rank == 0) umma_commit_pair(&ctl->acc_full[buf], 3);
  } else {
    umma_commit(&ctl->acc_full[buf]);
  }
}

// ---- epilogue helpers ----
__device__ __forceinline__ void ts_wait_acc(TsCtl* ctl, TsPipe& pp, int buf) {
#ifdef TS_ACC_SPIN
  mbar_wait(&ctl->acc_full[buf], pp.acc_use[buf] & 1);
#else
  mbar_wait_backoff(&ctl->acc_full[buf], pp.acc_use[buf] & 1);
#endif
  ++pp.acc_use[buf];
  tc_fence_after();
}
// my tcgen05.ld of the accumulator chunk and tcgen05.st of A chunk c have completed (wait::ld / wait::st done)
// CTA pairs: the peer CTA's warps arrive on the LEADER's barriers (`remote` = cluster address of the leader's
// a_ready[0] / s_ready[0]; 0 = arrive locally)
__device__ __forceinline__ void ts_signal(TsCtl* ctl, int c, int lane, uint32_t remote) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (remote) mbar_arrive_remote(remote + (uint32_t)c * 8u);
    else mbar_arrive(&ctl->a_ready[c]);
  }
}
__device__ __forceinline__ void ts_signal_smem(TsCtl* ctl, int cb, int i, int lane, uint32_t remote) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (remote) mbar_arrive_remote(remote + (uint32_t)(cb * 2 + i) * 8u);
    else mbar_arrive(&ctl->s_ready[cb][i]);
  }
}
// hidden layer: y = act(acc + bias) -> bf16, packed in place.  Thread owns fp32 columns [64c + 16cs, +16) of every
// chunk c and writes the packed values to columns [64c + 16cs, +8).  The load of chunk c+1 is in flight while chunk c
// is converted and stored.
template <bool RELU>
__device__ __forceinline__ void ts_epi_hidden(uint32_t tbuf /* tmem_base + lane_base + buf*256 */, const float* sb,
                                              const EpiCtx& ec, TsCtl* ctl, unsigned long long* tl, int* tn) {
  uint32_t v[2][16];
  tmem_ld16(tbuf + (uint32_t)(ec.cs * 16), v[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait16(v[c & 1]);
    if (c < 3) tmem_ld16(tbuf + (uint32_t)((c + 1) * 64 + ec.cs * 16), v[(c + 1) & 1]);
    const int col0 = c * 64 + ec.cs * 16;
    const float4* b4 = reinterpret_cast<const float4*>(sb + col0);
    uint32_t pk[8];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 b = b4[j4];
      pk[2 * j4] = pack2<RELU>(__uint_as_float(v[c & 1][4 * j4 + 0]) + b.x, __uint_as_float(v[c & 1][4 * j4 + 1]) + b.y);
      pk[2 * j4 + 1] = pack2<RELU>(__uint_as_float(v[c & 1][4 * j4 + 2]) + b.z, __uint_as_float(v[c & 1][4 * j4 + 3]) + b.w);
    }
    tmem_st8(tbuf + (uint32_t)col0, pk);
    tmem_st_wait();
    ts_signal(ctl, c, ec.lane, ec.remote_a_ready);
    if (tn) tl_mark(tl, 0, *tn, 60 + c);
  }
}

template <int FD, int CG>
__global__ void __launch_bounds__(THREADS, 1) k_back_ts(TcParams P, TileTable tt, RowIO io) {
  const float* __restrict__ x = io.x;
  const float* __restrict__ gate = io.gate;
  const float* __restrict__ noise = io.noise;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int t_first0 = CG * ((int)blockIdx.x / CG), t_stride = (int)gridDim.x;   // pairs of tiles share the expert
  TsCtl* ctl = reinterpret_cast<TsCtl*>(smem + TSB_CTL);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * TS_NST_MAX; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); mbar_init(&ctl->peer_ok[i], 1); }
    mbar_init(&ctl->acc_full[0], 1);
    mbar_init(&ctl->acc_full[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&ctl->a_ready[i], EPI_WARPS * CG);
    for (int i = 0; i < 4; ++i) mbar_init(&ctl->s_ready[i >> 1][i & 1], EPI_WARPS * CG);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_pair<512>(&ctl->tmem_base);
    else tmem_alloc<512>(&ctl->tmem_base);
  }
  float* sbias = reinterpret_cast<float*>(smem + TSB_BIAS);
  float* svec = reinterpret_cast<float*>(smem + TSB_VEC);
  float* sred = reinterpret_cast<float*>(smem + TSB_RED);
  float *s_wsig = svec, *s_wcol = svec + 256;
  const int H2 = P.hidden2;
  for (int i = threadIdx.x; i < MW; i += THREADS) s_wsig[i] = P.fblob[P.o_wsig + i];
  for (int i = threadIdx.x; i < 3 * H2; i += THREADS) s_wcol[i] = P.fblob[P.o_wcol + i];
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t acat_base = smem_u32(smem + TSM_ACAT), ring_base = smem_u32(smem + TSB_RING);
  const int n_tiles = *tt.n_tiles;
  const int NE = P.n_expert;
  const uint32_t K_xyz = P.front[0].K16, K_cat = P.back[1].K16 - MW;     // shared-memory operand widths (80, 80)
  TsPipe pp;

  if (warp == 0) {
    if (lane == 0)
      for (int tb = t_first0; tb < n_tiles; tb += t_stride) {
        const int e = tt.tile_expert[tb + (int)rank];
        if (e >= 0) {
          ts_produce<CG, TS_BACK_NST, true>(P.wblob + P.front[0].w_off, MW, K_xyz, smem + TSB_RING, ctl, pp, rank);
          for (int l = 0; l < NE; ++l) {
            ts_produce<CG, TS_BACK_NST, true>(P.wblob + P.expert[l].w_off + (size_t)e * P.expert_w_stride, MW, MW, smem + TSB_RING, ctl, pp, rank);
            if (l == P.skip_layer) ts_produce<CG, TS_BACK_NST, true>(P.wblob + P.front[0].w_off, MW, K_xyz, smem + TSB_RING, ctl, pp, rank);
          }
        }
        ts_produce<CG, TS_BACK_NST, true>(P.wblob + P.back[0].w_off, P.back[0].N, P.back[0].K16, smem + TSB_RING, ctl, pp, rank);
        ts_produce<CG, TS_BACK_NST, true>(P.wblob + P.back[1].w_off, P.back[1].N, P.back[1].K16, smem + TSB_RING, ctl, pp, rank);
      }
  } else if (warp == 1) {
    // all 32 lanes run the issue loop on warp-uniform values (see ts_mma_seg); lane `leader` issues
    const bool leader = elect_one();
    unsigned long long* tlm = leader ? P.tl : nullptr;
    uint32_t li = 0;
    int tn = 0;
    // layer li accumulates into buffer li & 1; a TMEM A operand sits in columns [0,128) of the other buffer
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    auto acc_of = [&](uint32_t l) { return tmem_u + (l & 1u) * 256u; };
    auto a_of = [&](uint32_t l) { return tmem_u + ((l & 1u) ^ 1u) * 256u; };
    uint32_t cb = 0;                 // cat block of this tile (alternates)
    const int n_tiles_u = __shfl_sync(0xffffffffu, n_tiles, 0);
    for (int tb = t_first0; tb < n_tiles_u; tb += t_stride, cb ^= 1u) {
      const int e = __shfl_sync(0xffffffffu, tt.tile_expert[tb + (int)rank], 0);
      tl_mark(tlm, 1, tn, 1);
      if (e >= 0) {
        ts_mma_seg<CG, TS_BACK_NST, TSB_CAT_STRIDE>(MW, K_xyz, false, 0, acat_base, ring_base, acc_of(li), ctl, pp, false, rank, leader, tlm, &tn, cb);
        ts_commit_acc<CG>(ctl, li & 1, rank, leader);
        tl_mark(tlm, 1, tn, 120);
        ++li;
        for (int l = 0; l < NE; ++l, ++li) {
          ts_mma_seg<CG, TS_BACK_NST, TSB_CAT_STRIDE>(MW, MW, true, a_of(li), acat_base, ring_base, acc_of(li), ctl, pp, false, rank, leader, tlm, &tn);
          if (l == P.skip_layer)
            ts_mma_seg<CG, TS_BACK_NST, TSB_CAT_STRIDE>(MW, K_xyz, false, 0, acat_base, ring_base, acc_of(li), ctl, pp, true, rank, leader, tlm, &tn, cb);
          ts_commit_acc<CG>(ctl, li & 1, rank, leader);
          tl_mark(tlm, 1, tn, 120);
        }
      }
      ts_mma_seg<CG, TS_BACK_NST, TSB_CAT_STRIDE>(P.back[0].N, MW, true, a_of(li), acat_base, ring_base, acc_of(li), ctl, pp, false, rank, leader, tlm, &tn);
      ts_commit_acc<CG>(ctl, li & 1, rank, leader);
      tl_mark(tlm, 1, tn, 120);
      ++li;
      ts_mma_seg<CG, TS_BACK_NST, TSB_CAT_STRIDE>(P.back[1].N, MW, true, a_of(li), acat_base, ring_base, acc_of(li), ctl, pp, false, rank, leader, tlm, &tn);
      ts_mma_seg<CG, TS_BACK_NST, TSB_CAT_STRIDE>(P.back[1].N, K_cat, false, 0, acat_base, ring_base, acc_of(li), ctl, pp, true, rank, leader, tlm, &tn, cb);
      ts_commit_acc<CG>(ctl, li & 1, rank, leader);
      tl_mark(tlm, 1, tn, 120);
      ++li;
    }
  } else {
    EpiCtx ec;
    ec.remote_a_ready = (CG == 2 && rank != 0) ? mapa_shared(smem_u32(&ctl->a_ready[0]), 0) : 0u;
    const uint32_t remote_s = (CG == 2 && rank != 0) ? mapa_shared(smem_u32(&ctl->s_ready[0][0]), 0) : 0u;
    ec.lane = lane; ec.q = warp & 3; ec.cs = (warp - 2) >> 2; ec.row = ec.q * 32 + lane;
    ec.et = (int)threadIdx.x - 64; ec.lane_base = (uint32_t)(ec.q * 32) << 16;
    const int row = ec.row;
    const float b_sig = P.fblob[P.o_bsig];
    const float b_col0 = P.fblob[P.o_bcol], b_col1 = P.fblob[P.o_bcol + 1], b_col2 = P.fblob[P.o_bcol + 2];
    uint32_t li = 0;
    int tn = 0;
    unsigned long long* tl = (warp == 2 && lane == 0) ? P.tl : nullptr;
    auto tbuf_of = [&](uint32_t l) { return tmem_base + ec.lane_base + (l & 1u) * 256u; };
    struct RowIn { int e, sidx; float g, d0, d1, d2, x0, x1, x2; int ai; int rows, row0; };
    // A tile's rows arrive through a chain of dependent global loads (tile table -> sample index -> ray / depth / gate
    // word, ~1K clk per level).  The three levels are issued in three different idle windows of the previous tile so that
    // no epilogue warp ever waits for one of them in front of a hand-off.
    auto fetch_meta = [&](int t) {
      RowIn r;
      r.e = -1; r.sidx = -1; r.g = 0.f; r.d0 = r.d1 = r.d2 = 0.f; r.x0 = r.x1 = r.x2 = 0.f; r.ai = 0; r.rows = 0; r.row0 = 0;
      if (t < n_tiles) { r.e = tt.tile_expert[t]; r.rows = tt.tile_rows[t]; r.row0 = tt.tile_row0[t]; }
      return r;
    };
    auto fetch_idx = [&](RowIn& r) {
      if (row < r.rows) r.sidx = tt.row2sample[r.row0 + row];
    };
    auto fetch_data = [&](RowIn& r) {
      if (r.sidx >= 0) {
        const float* xr = x + (int64_t)r.sidx * io.x_stride;
        if (r.e >= 0) r.g = io.wsel ? sel_gate(io.wsel[r.sidx]) : gate[(int64_t)r.sidx * io.g_stride];
        // xyz: the cs == 0 thread of the row stages PE(xyz); direction + appearance index: all four threads of the row
        // build a quarter of the [PE(dir) | appearance] block each (write_cat)
        if (P.ray_src) {
          if (ec.cs == 0 && r.e >= 0) ray_row_xyz(P, r.sidx, r.x0, r.x1, r.x2);
          ray_row_dir(P, r.sidx, r.d0, r.d1, r.d2, r.ai);
        } else {
          if (ec.cs == 0 && r.e >= 0) { r.x0 = xr[0]; r.x1 = xr[1]; r.x2 = xr[2]; }
          r.d0 = xr[P.x_cols - 4]; r.d1 = xr[P.x_cols - 3]; r.d2 = xr[P.x_cols - 2];
          r.ai = (int)xr[P.x_cols - 1];
        }
        r.ai = min(max(r.ai, 0), P.appearance_count - 1);
      }
    };
    auto fetch_row = [&](int t) {
      RowIn r = fetch_meta(t);
      fetch_idx(r);
      fetch_data(r);
      return r;
    };
    const int n_sxyz = ((int)K_xyz + 63) / 64, n_scat = ((int)K_cat + 63) / 64;
    // PE(xyz) of a tile -> cat block `blk` (cs == 0 threads) + release of its chunks to the MMA issuer
    auto stage_pe = [&](const RowIn& r, int blk) {
      if (ec.cs == 0) {
        constexpr int NPE = 3 + 6 * 12, NPAD = (NPE + 15) / 16 * 16;
        float pxyz[3] = {r.x0, r.x1, r.x2};
        __align__(16) __nv_bfloat16 pe[NPAD];
        pe_to_bf16<12>(pxyz, pe);
#pragma unroll
        for (int i = NPE; i < NPAD; ++i) pe[i] = __float2bfloat16_rn(0.f);
        ts_cat_store_row(acat_base + (uint32_t)blk * TSB_CAT_STRIDE, row, pe, NPAD / 8);
      }
      for (int i = 0; i < n_sxyz; ++i) ts_signal_smem(ctl, blk, i, lane, remote_s);
    };
    RowIn nxt = fetch_row(t_first0 + (int)rank);
    if (nxt.e >= 0) stage_pe(nxt, 0);          // first tile of this CTA; later tiles are staged one tile ahead
    int cb = 0;
    bool xyz_bias_ready = false;
    const int cat_first = P.skip_layer > 0 ? P.skip_layer : 0;   // first layer after which the PE(xyz) block is free
    for (int tb = t_first0; tb < n_tiles; tb += t_stride, cb ^= 1) {
      const int t = tb + (int)rank;
      const RowIn cur = nxt;
      const int e = cur.e, sidx = cur.sidx;
      const bool valid = sidx >= 0;
      const float g = cur.g;
      const uint32_t acat_cur = acat_base + (uint32_t)cb * TSB_CAT_STRIDE;
      tl_mark(tl, 0, tn, 1);
      // [PE(dir) | appearance | 0-pad] -> cat block: 16-byte chunk g (columns [8g, 8g + 8)) is built by the thread
      // cs == g % 4 of the row, every index a compile-time constant (the first version filled a local array on 4 of the
      // 16 warps: ~5K clk next to the skip layer, profiles/r3g_issue_path.md)
      // In three parts (chunks g with g / 4 == part), one per layer after the skip layer: each part is one round of
      // embedding loads (~1K clk of L2 latency) and fits under a layer's MMAs
      auto write_cat = [&](int part) {
        constexpr int NDIR = 3 + 6 * FD;
        static_assert(NDIR <= 32, "PE(dir) lies in the first four chunks (part 0)");
        float pe[NDIR];
#pragma unroll
        for (int i = 0; i < NDIR; ++i) pe[i] = 0.f;
        if (part == 0) {
          float sn[3], cs_[3];
          const float dvec[3] = {cur.d0, cur.d1, cur.d2};
#pragma unroll
          for (int a = 0; a < 3; ++a) { pe[a] = dvec[a]; sincosf(dvec[a], &sn[a], &cs_[a]); }
#pragma unroll
          for (int k = 0; k < FD; ++k) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {                      // same recurrence as pe_to_bf16
              pe[3 + 6 * k + a] = sn[a];
              pe[3 + 6 * k + 3 + a] = cs_[a];
              const float s2 = 2.f * sn[a] * cs_[a];
              const float c2 = 1.f - 2.f * sn[a] * sn[a];
              sn[a] = s2;
              cs_[a] = c2;
            }
          }
        }
        // the appearance part comes from the pre-shifted bf16 table (snb_tc.cu: k_pack_emb_cat): one 16-byte load per
        // chunk.  A gather of 128 random rows is bound by the L1 at about one lane-request per clock -- fp32 rows read with
        // scalar loads cost 4K clk per tile, this layout 7 requests per row
        const uint4* er = reinterpret_cast<const uint4*>(P.emb_cat + (int64_t)cur.ai * P.cat_cols);
        const int n8 = (int)K_cat / 8;
#pragma unroll
        for (int g = 0; g < (int)TS_CAT_COLS / 8; ++g) {
          if ((g >> 2) != part || (g & 3) != ec.cs || g >= n8) continue;
          uint4 q = make_uint4(0u, 0u, 0u, 0u);
          if (8 * g + 8 > NDIR && valid && P.emb_cat) q = __ldg(er + g);
          if (8 * g < NDIR) {                                   // PE(dir) columns of this chunk (the table has zeros there)
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c0 = 8 * g + 2 * i, c1 = c0 + 1;
              const uint32_t pp2 = pack2<false>(c0 < NDIR ? pe[c0 < NDIR ? c0 : 0] : 0.f, c1 < NDIR ? pe[c1 < NDIR ? c1 : 0] : 0.f);
              w[i] = valid ? (w[i] | pp2) : 0u;
            }
            q = make_uint4(w[0], w[1], w[2], w[3]);
          }
          st_shared_v4(ts_cat_addr(acat_cur, row, g), q.x, q.y, q.z, q.w);
        }
      };
      float sig_acc = 0.f;
      nxt = fetch_meta(t + t_stride);
      if (e >= 0) {
        // PE(xyz) of this tile was staged into block cb one tile ago (operand of the xyz layer and of the skip term)
        tl_mark(tl, 0, tn, 2);
        // ---- xyz layer (act none): h -> packed A ----
        {
          const int buf = (int)(li & 1);
          // its bias was put into this slot during layer "2" of the previous tile (the tensor pipe ran this tile's xyz
          // layer behind that layer's MMAs and has been idle since)
          if (!xyz_bias_ready) epi_load_bias(P.fblob + P.front[0].b_off, MW, sbias, buf, ec.et);
          ts_wait_acc(ctl, pp, buf);
          ts_epi_hidden<false>(tbuf_of(li), sbias + buf * 256, ec, ctl, tl, &tn);
          if (P.skip_layer == 0) for (int i = 0; i < n_sxyz; ++i) ts_signal_smem(ctl, cb, i, lane, remote_s);
          ++li;
        }
        fetch_idx(nxt);              // under the first expert layer's MMAs; fetch_data follows under the second layer's
        for (int l = 0; l < NE; ++l, ++li) {
          const int buf = (int)(li & 1);
          const bool skip_here = (l == P.skip_layer);
          epi_load_bias(skip_here ? (P.fblob + P.o_b3x + (size_t)e * MW)
                                  : (P.fblob + P.expert[l].b_off + (size_t)e * P.expert_b_stride), MW, sbias, buf, ec.et);
          const float* sb = sbias + buf * 256;
          tl_mark(tl, 0, tn, 10 + l);
          ts_wait_acc(ctl, pp, buf);
          tl_mark(tl, 0, tn, 20 + l);
          const uint32_t tb = tbuf_of(li);
          if (l < NE - 1) {
            ts_epi_hidden<true>(tb, sb, ec, ctl, tl, &tn);
            if (l + 1 == P.skip_layer) for (int i = 0; i < n_sxyz; ++i) ts_signal_smem(ctl, cb, i, lane, remote_s);
            // every MMA of the skip layer has retired (acc_full above): the PE(xyz) block is free.  Filled AFTER the
            // hand-off, under the following layers' MMAs
            if (l >= cat_first && l < cat_first + 3) write_cat(l - cat_first);
            if (l == 0) fetch_data(nxt);
          } else {
            if (l == 0) fetch_data(nxt);
            for (int part = (l >= cat_first ? l - cat_first : 0); part < 3; ++part) write_cat(part);
            // last expert layer (no activation) -> combine: y = bf16(gate * bf16(out)) -> ReLU -> packed A; sigma head
            uint32_t v[2][16];
            tmem_ld16(tb + (uint32_t)(ec.cs * 16), v[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tmem_ld_wait16(v[c & 1]);
              if (c < 3) tmem_ld16(tb + (uint32_t)((c + 1) * 64 + ec.cs * 16), v[(c + 1) & 1]);
              const int col0 = c * 64 + ec.cs * 16;
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float f0, f1;
                ts_round2<false>(__uint_as_float(v[c & 1][j]) + sb[col0 + j], __uint_as_float(v[c & 1][j + 1]) + sb[col0 + j + 1], f0, f1);
                pk[j / 2] = ts_round2<true>(f0 * g, f1 * g, f0, f1);      // relu(bf16(gate * bf16(out)))
                sig_acc = fmaf(f0, s_wsig[col0 + j], sig_acc);
                sig_acc = fmaf(f1, s_wsig[col0 + j + 1], sig_acc);
              }
              tmem_st8(tb + (uint32_t)col0, pk);
              tmem_st_wait();
              ts_signal(ctl, c, lane, ec.remote_a_ready);
            }
          }
          tl_mark(tl, 0, tn, 30 + l);
        }
      } else {
        // dropped bucket: h = relu(0) = 0 -> zero A operand for layer "1" (it accumulates into B[li&1], reads B[~li&1])
        fetch_idx(nxt);
        fetch_data(nxt);
        const uint32_t ta = tbuf_of(li + 1);
        const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 4; ++c) tmem_st8(ta + (uint32_t)(c * 64 + ec.cs * 16), z);
        tmem_st_wait();
        for (int c = 0; c < 4; ++c) ts_signal(ctl, c, lane, ec.remote_a_ready);
        for (int part = 0; part < 3; ++part) write_cat(part);
      }
      sred[(0 * 4 + ec.cs) * 128 + row] = sig_acc;
      // ---- layer "1" (act none) -> packed A; then release the cat chunks ----
      {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.back[0].b_off, MW, sbias, buf, ec.et);
        tl_mark(tl, 0, tn, 70);
        // the tensor pipe is busy with layer "1": stage the next tile's PE(xyz) into the other cat block now
        // (its last reader, layer "2" of the previous tile, retired before this tile started)
        if (nxt.e >= 0) stage_pe(nxt, cb ^ 1);
        tl_mark(tl, 0, tn, 71);
        ts_wait_acc(ctl, pp, buf);
        tl_mark(tl, 0, tn, 72);
        ts_epi_hidden<false>(tbuf_of(li), sbias + buf * 256, ec, ctl, tl, &tn);
        for (int i = 0; i < n_scat; ++i) ts_signal_smem(ctl, cb, i, lane, remote_s);
        tl_mark(tl, 0, tn, 50);
        ++li;
      }
      // ---- layer "2" (ReLU) + colour head ----
      {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.back[1].b_off, H2, sbias, buf, ec.et);
        // bias of the next tile's xyz layer -> the other slot (its last readers, layer "1", are past the barrier above)
        xyz_bias_ready = nxt.e >= 0;
        if (xyz_bias_ready) epi_load_bias(P.fblob + P.front[0].b_off, MW, sbias, buf ^ 1, ec.et);
        const float* sb = sbias + buf * 256;
        tl_mark(tl, 0, tn, 73);
        ts_wait_acc(ctl, pp, buf);
        tl_mark(tl, 0, tn, 74);
        const uint32_t tb = tbuf_of(li);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int c = 0; c < H2 / 64; ++c) {
          const int col0 = c * 64 + ec.cs * 16;
          uint32_t v[16];
          tmem_ld16(tb + (uint32_t)col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const int k = col0 + j;
            float ha, hb;                               // bf16(relu(x)) == relu(bf16(x))
            ts_round2<true>(__uint_as_float(v[j]) + sb[k], __uint_as_float(v[j + 1]) + sb[k + 1], ha, hb);
            c0 = fmaf(ha, s_wcol[k], c0);
            c1 = fmaf(ha, s_wcol[H2 + k], c1);
            c2 = fmaf(ha, s_wcol[2 * H2 + k], c2);
            c0 = fmaf(hb, s_wcol[k + 1], c0);
            c1 = fmaf(hb, s_wcol[H2 + k + 1], c1);
            c2 = fmaf(hb, s_wcol[2 * H2 + k + 1], c2);
          }
        }
        tc_fence_before();
        tl_mark(tl, 0, tn, 75);
        sred[(1 * 4 + ec.cs) * 128 + row] = c0;
        sred[(2 * 4 + ec.cs) * 128 + row] = c1;
        sred[(3 * 4 + ec.cs) * 128 + row] = c2;
        epi_bar_sync();
        tl_mark(tl, 0, tn, 76);
        if (ec.cs == 0 && valid) {
          auto rsum = [&](int v) { return sred[(v * 4 + 0) * 128 + row] + sred[(v * 4 + 1) * 128 + row] +
                                          sred[(v * 4 + 2) * 128 + row] + sred[(v * 4 + 3) * 128 + row]; };
          // reference under cuda autocast: the Linear output, `sigma += noise` and `x - 1` (nerf.py:68) are bf16 tensor
          // ops (each rounds to bf16); only F.softplus runs in fp32 -- pinned by tests/golden/model_*_bf16cuda.npz
          float sr = bf16_round(rsum(0) + b_sig);
          if (noise) sr = bf16_round(sr + noise[(int64_t)sidx * io.n_stride]);
          const float tt_ = bf16_round(sr - 1.f);
          const float sigma = (tt_ > 20.f) ? tt_ : log1pf(expf(tt_));
          auto sg = [](float v) { return bf16_round(1.f / (1.f + expf(-bf16_round(v)))); };
          float4 o = make_float4(sg(rsum(1) + b_col0), sg(rsum(2) + b_col1), sg(rsum(3) + b_col2), sigma);
          if (io.ep) {
            const float* rec = x + (int64_t)sidx * io.x_stride;
            reinterpret_cast<float4*>(io.ret[__float_as_int(rec[10])])[__float_as_int(rec[9])] = o;
          } else {
            reinterpret_cast<float4*>(io.out)[sidx] = o;
          }
        }
        epi_bar_sync();
        tl_mark(tl, 0, tn, 51);
        ++li;
      }
    }
    if (io.ep) __threadfence_system();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();     // the peer may still arrive on this CTA's barriers
  else __syncthreads();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
  if (io.ep && threadIdx.x == 0) {
    if (atomicAdd(io.done, 1) == (int)gridDim.x - 1) {
      *io.done = 0;
      __threadfence_system();
      for (int w = 0; w < io.world; ++w)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(io.flag_b[w]), "r"(io.epoch) : "memory");
    }
  }
}

// ---- softmax of the gate logits + routing hand-off, shared by the front kernels -------------------------------------
// logits = rstd * (G_hi + G_lo - mean * c1) + c0 from the folded LayerNorm + gate GEMM accumulator (32 columns at
// `tacc`); LayerNorm partial sums in sred[0..7]; s_gc = [c0 | c1]; s_hist = the CTA's level-0 key histogram.
__device__ __forceinline__ void front_softmax_select(const TcParams& P, const float* sred, const float* s_gc,
                                                     uint32_t* s_hist, uint32_t tacc, const EpiCtx& ec, int row, int lane,
                                                     bool valid, int64_t s, int t, float inv_w,
                                                     float* __restrict__ gates, uint32_t* __restrict__ wsel,
                                                     float* __restrict__ pm, int32_t* __restrict__ moe_idx) {
  // every one of the four threads of a row (cs = 0..3, different warps) holds all 32 accumulator columns, so
  // each computes the whole softmax of its row itself: no exchange rounds.  Thread cs stores gates
  // [4cs, 4cs+4) (debug tap only); thread 0 owns the routing word and the column sums.
  // ... which only matters for the debug tap: without it the routing word, the histogram and the column sums all come
  // from thread 0 of the row, and the other three warps of the lane quarter have nothing to do here (they used to
  // compute 16 expf each: the softmax was issue-bound at ~2.6K clk per tile)
  if (ec.cs != 0 && gates == nullptr) return;
  const float tsum = sred[0 * 128 + row] + sred[1 * 128 + row] + sred[2 * 128 + row] + sred[3 * 128 + row];
  const float tsq = sred[4 * 128 + row] + sred[5 * 128 + row] + sred[6 * 128 + row] + sred[7 * 128 + row];
  const float mean = tsum * inv_w;
  const float var = fmaxf(tsq * inv_w - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
    uint32_t hi[16], lo[16];
  tmem_ld16(tacc, hi);
  tmem_ld16(tacc + 16u, lo);
  tmem_ld_wait();
  float lg[MAX_E];
  float mx = -INFINITY;
#pragma unroll
  for (int e = 0; e < MAX_E; ++e) {
    lg[e] = rstd * (__uint_as_float(hi[e]) + __uint_as_float(lo[e]) - mean * s_gc[MAX_E + e]) + s_gc[e];
    if (e < P.E) mx = fmaxf(mx, lg[e]);
  }
  float den = 0.f;
#pragma unroll
  for (int e = 0; e < MAX_E / 2; ++e) {
    lg[e] = (e < P.E) ? expf(lg[e] - mx) : 0.f;
    den += lg[e];
  }
  if (P.E > MAX_E / 2) {                       // uniform branch: 8 experts never evaluate the upper 8 exponentials
#pragma unroll
    for (int e = MAX_E / 2; e < MAX_E; ++e) {
      lg[e] = (e < P.E) ? expf(lg[e] - mx) : 0.f;
      den += lg[e];
    }
  } else {
#pragma unroll
    for (int e = MAX_E / 2; e < MAX_E; ++e) lg[e] = 0.f;
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int e = 0; e < MAX_E; ++e) lg[e] *= inv;
  if (valid) {
    if (gates) {
      const int e0 = 4 * ec.cs;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        if (ec.cs == q4) {            // static register indices
          if ((P.E & 3) == 0 && e0 < P.E) {
            *reinterpret_cast<float4*>(gates + s * P.E + e0) =
                make_float4(lg[4 * q4], lg[4 * q4 + 1], lg[4 * q4 + 2], lg[4 * q4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (e0 + j < P.E) gates[s * P.E + e0 + j] = lg[4 * q4 + j];
          }
        }
    }
    if (ec.cs == 0) {
      // argmax of the fp32 gates, lowest expert id on ties (torch.argmax / extract_critical)
      int best = 0;
      float bv = lg[0];
#pragma unroll
      for (int e = 1; e < MAX_E; ++e)
        if (e < P.E && lg[e] > bv) { bv = lg[e]; best = e; }
      if (wsel) {
        const uint32_t key = sel_key(bv);
        wsel[s] = sel_pack(best, key);
        const int bin = best * SEL_HBINS + (int)(key >> SEL_L1_SHIFT);
        if (!(P.ab & 1)) atomicAdd(&s_hist[bin >> 1], 1u << (16 * (bin & 1)));
        if (moe_idx) moe_idx[s] = best;
      }
    }
  }
  if (pm && ec.cs == 0 && !(P.ab & 2)) {
    // column sums of the gates over the 32 rows of this warp (load-balance loss), in a fixed order that does
    // not depend on the grid: 16 values x 32 lanes folded by a transpose-reduction (16 shuffles), after which
    // even lane l holds the sum of expert ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1).
    if (!valid) {
#pragma unroll
      for (int e = 0; e < MAX_E; ++e) lg[e] = 0.f;
    }
    float a8[8], b4[4], c2[2];
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float send = hi ? lg[i] : lg[i + 8], keep = hi ? lg[i + 8] : lg[i];
        a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {
      const bool hi = lane & 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = hi ? a8[i] : a8[i + 4], keep = hi ? a8[i + 4] : a8[i];
        b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    {
      const bool hi = lane & 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = hi ? b4[i] : b4[i + 2], keep = hi ? b4[i + 2] : b4[i];
        c2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    }
    const bool hi2 = lane & 2;
    float d1 = (hi2 ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, hi2 ? c2[0] : c2[1], 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    if ((lane & 1) == 0) {
      const int ecol = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      pm[((int64_t)t * 4 + ec.q) * SEL_PM_STRIDE + ecol] = d1;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Launch #1, TS variant: PE -> xyz layer -> external gate MLP -> folded LayerNorm + gate GEMM -> softmax, with the
// hidden activations packed in tensor memory exactly as in k_back_ts (h is not written: launch #2 recomputes it).
// ------------------------------------------------------------------------------------------------------------
// Routing hand-off (snb_select.cuh): the thread that owns a row's softmax writes the packed word
// (expert id << 26 | key of the max gate) of its sample and counts it in the per-expert histogram of the top 9 key
// bits (`hist0`, zeroed by the host before the launch), and every CTA leaves the column sums of the gates over its
// rows (`pm`, one record of SEL_PM_STRIDE floats per 32 rows = per (tile, TMEM lane quarter)) for the load-balance loss.
// `gates` ([S,E] fp32) is only written when a caller asked for it (debug tap).
template <int FX>
__global__ void __launch_bounds__(THREADS, 1) k_front_ts(TcParams P, const float* __restrict__ x, int64_t S,
                                                         float* __restrict__ gates, uint32_t* __restrict__ wsel,
                                                         int* __restrict__ hist0, float* __restrict__ pm,
                                                         int32_t* __restrict__ moe_idx) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  TsCtl* ctl = reinterpret_cast<TsCtl*>(smem + TSM_CTL);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * TS_NST_MAX; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); mbar_init(&ctl->peer_ok[i], 1); }
    mbar_init(&ctl->acc_full[0], 1);
    mbar_init(&ctl->acc_full[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&ctl->a_ready[i], EPI_WARPS);
    for (int i = 0; i < 4; ++i) mbar_init(&ctl->s_ready[i >> 1][i & 1], EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  float* sbias = reinterpret_cast<float*>(smem + TSM_BIAS);
  float* sred = reinterpret_cast<float*>(smem + TSM_RED);
  float* s_gc = reinterpret_cast<float*>(smem + TSM_VEC);     // [0,16) = c0, [16,32) = c1 of the folded gate GEMM
  if (threadIdx.x < 2 * MAX_E)
    s_gc[threadIdx.x] = P.fblob[(threadIdx.x < MAX_E ? P.o_c0 : P.o_c1 - MAX_E) + threadIdx.x];
  // level-0 key histogram of this CTA's rows: 16-bit counters (a CTA sees < 65536 rows per launch: checked on the host),
  // two per word; flushed into the global histogram with one reduction per non-zero counter at the end
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_gc + 2 * MAX_E);
  static_assert((2 * MAX_E + MAX_E * SEL_HBINS / 2) * 4 <= SM_VEC_FLOATS * 4, "histogram fits in the head-vector block");
  for (int i = threadIdx.x; i < MAX_E * SEL_HBINS / 2; i += THREADS) s_hist[i] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t acat_base = smem_u32(smem + TSM_ACAT), ring_base = smem_u32(smem + TSM_RING);
  const int n_tiles = (int)((S + TILE - 1) / TILE);
  const int NL = P.n_front;                       // xyz layer + gate MLP layers
  const uint32_t K_xyz = P.front[0].K16;
  const int n_sxyz = ((int)K_xyz + 63) / 64;
  TsPipe pp;

  if (warp == 0) {
    if (lane == 0)
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        for (int l = 0; l < NL; ++l)
          ts_produce<1>(P.wblob + P.front[l].w_off, P.front[l].N, P.front[l].K16, smem + TSM_RING, ctl, pp, 0);
        ts_produce<1>(P.wblob + P.gate.w_off, GATE_N, MW, smem + TSM_RING, ctl, pp, 0);
      }
  } else if (warp == 1) {
    const bool leader = elect_one();               // warp-uniform issue loop, see ts_mma_seg
    unsigned long long* tlm = leader ? P.tl : nullptr;
    uint32_t li = 0;
    int tn = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    auto acc_of = [&](uint32_t l) { return tmem_u + (l & 1u) * 256u; };
    auto a_of = [&](uint32_t l) { return tmem_u + ((l & 1u) ^ 1u) * 256u; };
    uint32_t cb = 0;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, cb ^= 1u) {
      tl_mark(tlm, 1, tn, 1);
      ts_mma_seg<1>(MW, K_xyz, false, 0, acat_base, ring_base, acc_of(li), ctl, pp, false, 0, leader, tlm, &tn, cb);
      ts_commit_acc<1>(ctl, li & 1, 0, leader);
      ++li;
      for (int l = 1; l < NL; ++l, ++li) {
        ts_mma_seg<1>(MW, MW, true, a_of(li), acat_base, ring_base, acc_of(li), ctl, pp, false, 0, leader, tlm, &tn);
        ts_commit_acc<1>(ctl, li & 1, 0, leader);
        tl_mark(tlm, 1, tn, 120);
      }
      ts_mma_seg<1>(GATE_N, MW, true, a_of(li), acat_base, ring_base, acc_of(li), ctl, pp, false, 0, leader, tlm, &tn);
      ts_commit_acc<1>(ctl, li & 1, 0, leader);
      tl_mark(tlm, 1, tn, 120);
      ++li;
    }
  } else {
    EpiCtx ec;
    ec.remote_a_ready = 0; ec.lane = lane; ec.q = warp & 3; ec.cs = (warp - 2) >> 2; ec.row = ec.q * 32 + lane;
    ec.et = (int)threadIdx.x - 64; ec.lane_base = (uint32_t)(ec.q * 32) << 16;
    const int row = ec.row;
    uint32_t li = 0;
    int tn = 0;
    unsigned long long* tl = (warp == 2 && lane == 0) ? P.tl : nullptr;
    auto tbuf_of = [&](uint32_t l) { return tmem_base + ec.lane_base + (l & 1u) * 256u; };
    constexpr int NPE = 3 + 6 * FX, NPAD = (NPE + 15) / 16 * 16;
    auto load_xyz = [&](int tile, float (&p)[3]) {
      p[0] = p[1] = p[2] = 0.f;
      const int64_t sr = (int64_t)tile * TILE + row;
      if (ec.cs == 0 && tile < n_tiles && sr < S) {
        if (P.ray_src) ray_row_xyz(P, sr, p[0], p[1], p[2]);
        else { p[0] = x[sr * P.x_cols]; p[1] = x[sr * P.x_cols + 1]; p[2] = x[sr * P.x_cols + 2]; }
      }
    };
    // PE(xyz) of a tile -> cat block `blk` (cs == 0 threads) + release of its chunks to the MMA issuer
    auto stage_pe = [&](const float (&p)[3], int blk) {
      if (ec.cs == 0) {
        __align__(16) __nv_bfloat16 pe[NPAD];
        pe_to_bf16<FX>(p, pe);
#pragma unroll
        for (int i = NPE; i < NPAD; ++i) pe[i] = __float2bfloat16_rn(0.f);
        ts_cat_store_row(acat_base + (uint32_t)blk * TS_ACAT_BYTES, row, pe, NPAD / 8);
      }
      for (int i = 0; i < n_sxyz; ++i) ts_signal_smem(ctl, blk, i, lane, 0);
    };
    float pn[3];                             // xyz of this thread's row in the NEXT tile (staged one tile ahead)
    load_xyz((int)blockIdx.x, pn);
    if ((int)blockIdx.x < n_tiles) stage_pe(pn, 0);
    load_xyz((int)blockIdx.x + (int)gridDim.x, pn);
    int cb = 0;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, cb ^= 1) {
      const int64_t s = (int64_t)t * TILE + row;
      const bool valid = s < S;
      tl_mark(tl, 0, tn, 1);
      tl_mark(tl, 0, tn, 2);
      float sum = 0.f, sq = 0.f;
      for (int l = 0; l < NL; ++l, ++li) {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.front[l].b_off, MW, sbias, buf, ec.et);
        const float* sb = sbias + buf * 256;
        if (l == 1 && t + (int)gridDim.x < n_tiles) {
          // the tensor pipe is busy with the first gate-MLP layer: stage the next tile's PE(xyz) into the other cat
          // block (its reader, the xyz layer of the previous tile, retired long ago) and prefetch the one after
          stage_pe(pn, cb ^ 1);
          load_xyz(t + 2 * (int)gridDim.x, pn);
        }
        tl_mark(tl, 0, tn, 10 + l);
        ts_wait_acc(ctl, pp, buf);
        tl_mark(tl, 0, tn, 20 + l);
        const uint32_t tb = tbuf_of(li);
        if (l == 0) {
          ts_epi_hidden<false>(tb, sb, ec, ctl, tl, &tn);       // h = xyz Linear (act none)
        } else if (l < NL - 1) {
          ts_epi_hidden<true>(tb, sb, ec, ctl, tl, &tn);
        } else {
          // last gate-MLP layer: g = bf16(Linear) -> packed A of the folded gate GEMM + LayerNorm statistics
          uint32_t v[2][16];
          tmem_ld16(tb + (uint32_t)(ec.cs * 16), v[0]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmem_ld_wait16(v[c & 1]);
            if (c < 3) tmem_ld16(tb + (uint32_t)((c + 1) * 64 + ec.cs * 16), v[(c + 1) & 1]);
            const int col0 = c * 64 + ec.cs * 16;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float g0, g1;
              pk[j / 2] = ts_round2<false>(__uint_as_float(v[c & 1][j]) + sb[col0 + j],
                                           __uint_as_float(v[c & 1][j + 1]) + sb[col0 + j + 1], g0, g1);
              sum += g0 + g1;
              sq = fmaf(g0, g0, fmaf(g1, g1, sq));
            }
            tmem_st8(tb + (uint32_t)col0, pk);
            tmem_st_wait();
            ts_signal(ctl, c, lane, 0);
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
        ts_wait_acc(ctl, pp, buf);
        tl_mark(tl, 0, tn, 40);
        front_softmax_select(P, sred, s_gc, s_hist, tbuf_of(li), ec, row, lane, valid, s, t, 1.f / MW, gates, wsel, pm, moe_idx);
        tc_fence_before();
        epi_bar_sync();                       // sred is reused by the next tile
        tl_mark(tl, 0, tn, 41);
        ++li;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
  if (wsel) {
    for (int i = threadIdx.x; i < MAX_E * SEL_HBINS / 2; i += THREADS) {
      const uint32_t v = s_hist[i];
      if (v & 0xffffu) atomicAdd(&hist0[2 * i], (int)(v & 0xffffu));
      if (v >> 16) atomicAdd(&hist0[2 * i + 1], (int)(v >> 16));
    }
  }
}

// Launch #2, "TS" variant: hidden activations live in TENSOR MEMORY instead of shared memory.
// (included by snb_tc.cu after k_back; shares its constants, TcParams and epilogue helpers)
//
// Why.  k_back is shared-memory-bandwidth bound: per 128x256x16 MMA (128 clk at the tensor-pipe floor) the SM
// moves A 4 KB + B 8 KB operand reads, 8 KB of weight landing and 4 KB of epilogue stores = 192 B/clk against
// 128 B/clk of shared-memory bandwidth, so the MMAs run at ~250 clk.  Here the epilogue writes the next layer's
// A operand with tcgen05.st into TMEM and the MMA takes it from there (tcgen05.mma with the A operand in tensor
// memory): shared memory only carries the weights (and the 80 PE / [dir|appearance] columns).
//
// TMEM map (512 columns): accumulator [0,256) as four 64-column quarters, A ping [256,384), A pong [384,512)
// (128 columns = 256 bf16 of K, two per 32-bit column).  There is no second accumulator buffer to hide the
// epilogue behind, so a layer is scheduled in 64x64 blocks instead: the epilogue of layer l drains quarter c
// and writes K-chunk c of the next A; the moment that lands, layer l+1 may issue every block that needs only
// quarters <= c and chunks <= c.  The weights are packed as a stream of (quarter, K-chunk) blocks in exactly
// that order, one 8 KB ring slot per block, consumed once, in order:
//     for c = 0 .. max(nq, nts)-1:
//         if c < nts: for q < min(c, nq):  (q, TMEM chunk c)          -- old quarters take the new chunk
//         if c < nq:  for s < nss:         (c, smem chunk s)           -- the new quarter: PE / cat chunks ...
//                     for k <= min(c, nts-1): (c, TMEM chunk k)        -- ... and every TMEM chunk so far
// nq = N/64 quarters, nts = K-chunks taken from TMEM (0 or 4), nss = K-chunks taken from shared memory.
#pragma once

static constexpr uint32_t TS_CAT_COLS = 96;                       // PE(xyz) / [PE(dir) | appearance] block
static constexpr uint32_t TS_SBO = TS_CAT_COLS / 8 * 128;         // 1536 B between 8-row groups
static constexpr uint32_t TS_ACAT_BYTES = TILE / 8 * TS_SBO;      // 24576
static constexpr int TS_NSLOT = 20;
static constexpr uint32_t TS_SLOT = 64 * 64 * 2;                  // one (quarter, 64-wide K chunk) weight block
static constexpr uint32_t TS_ACOL = 256;                          // first TMEM column of the A ping buffer

struct TsShape { int nq, nts, nss, kss; };     // kss = shared-memory K columns (multiple of 16)
__host__ __device__ inline int ts_ss_klen(const TsShape& s, int i) { const int r = s.kss - 64 * i; return r < 64 ? r : 64; }
__host__ __device__ inline int ts_imin(int a, int b) { return a < b ? a : b; }
// f(q, is_smem_chunk, k)
template <typename F>
__host__ __device__ inline void ts_for_each_block(const TsShape& s, F&& f) {
  const int n = s.nq > s.nts ? s.nq : s.nts;
  for (int c = 0; c < n; ++c) {
    if (c < s.nts) for (int q = 0; q < ts_imin(c, s.nq); ++q) f(q, 0, c);
    if (c < s.nq) {
      for (int i = 0; i < s.nss; ++i) f(c, 1, i);
      for (int k = 0; k <= ts_imin(c, s.nts - 1); ++k) f(c, 0, k);
    }
  }
}
inline size_t ts_stream_bytes(const TsShape& s) {
  size_t b = 0;
  ts_for_each_block(s, [&](int, int ss, int k) { b += (size_t)64 * (ss ? ts_ss_klen(s, k) : 64) * 2; });
  return b;
}

// ---- packing: fp32 [N][K] row-major sources -> bf16 block stream --------------------------------------------
struct TsPackTab {
  int n;
  struct { short q, ss, k, klen; int dst; } b[32];
};
// main: TMEM-chunk columns (k*64 ..), aux: shared-memory-chunk columns (k*64 .. of the aux matrix, zero padded)
__global__ void k_pack_ts(TsPackTab tab, const float* __restrict__ main_w, int main_ld, const float* __restrict__ aux_w,
                          int aux_ld, int aux_k, uint8_t* __restrict__ dst) {
  const int bi = blockIdx.x;
  const int q = tab.b[bi].q, ss = tab.b[bi].ss, k = tab.b[bi].k, klen = tab.b[bi].klen;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst + tab.b[bi].dst);
  for (int i = threadIdx.x; i < 64 * klen; i += blockDim.x) {
    const int n = i / klen, kk = i % klen;
    float v;
    if (ss) {
      const int col = k * 64 + kk;
      v = (col < aux_k) ? aux_w[(size_t)(q * 64 + n) * aux_ld + col] : 0.f;
    } else {
      v = main_w[(size_t)(q * 64 + n) * main_ld + k * 64 + kk];
    }
    // canonical K-major core-matrix image of a 64 x klen block
    out[((size_t)(n / 8) * (klen * 16) + (size_t)(kk / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2) / 2] = __float2bfloat16_rn(v);
  }
}
static int ts_pack_stream(const TsShape& s, const float* main_w, int main_ld, const float* aux_w, int aux_ld, int aux_k,
                          uint8_t* dst, cudaStream_t st) {
  TsPackTab tab;
  tab.n = 0;
  int off = 0;
  ts_for_each_block(s, [&](int q, int ss, int k) {
    const int klen = ss ? ts_ss_klen(s, k) : 64;
    tab.b[tab.n].q = (short)q; tab.b[tab.n].ss = (short)ss; tab.b[tab.n].k = (short)k; tab.b[tab.n].klen = (short)klen;
    tab.b[tab.n].dst = off;
    off += 64 * klen * 2;
    ++tab.n;
  });
  k_pack_ts<<<tab.n, 256, 0, st>>>(tab, main_w, main_ld, aux_w, aux_ld, aux_k, dst);
  SNB_CHECK_LAUNCH("k_pack_ts");
  return SNB_OK;
}

// ---- shared-memory carve-up ------------------------------------------------------------------------------------
struct __align__(16) TsCtl {
  uint64_t full[TS_NSLOT];
  uint64_t empty[TS_NSLOT];
  uint64_t acc_full[4];     // MMA -> epilogue: accumulator quarter complete
  uint64_t acc_free[4];     // epilogue -> MMA: accumulator quarter drained
  uint64_t a_ready[4];      // epilogue -> MMA: K-chunk of the next A operand written to TMEM
  uint64_t s_ready[2];      // epilogue -> MMA: shared-memory K-chunk (PE / cat block) written
  uint32_t tmem_base;
  uint32_t pad;
};
static constexpr size_t TSM_ACAT = 0;
static constexpr size_t TSM_RING = TS_ACAT_BYTES;
static constexpr size_t TSM_BIAS = TSM_RING + (size_t)TS_NSLOT * TS_SLOT;
static constexpr size_t TSM_VEC = TSM_BIAS + 2 * 256 * 4;
static constexpr size_t TSM_RED = TSM_VEC + SM_VEC_FLOATS * 4;
static constexpr size_t TSM_CTL = TSM_RED + SM_RED_FLOATS * 4;
static constexpr size_t TSM_TOTAL = TSM_CTL + sizeof(TsCtl) + 1024;
static_assert(TSM_TOTAL <= 227 * 1024, "shared memory budget (TS kernel)");

struct TsPipe {
  uint32_t blk = 0;                    // weight blocks produced / consumed
  uint32_t a_use[4] = {0, 0, 0, 0};    // a_ready waits (MMA) / -
  uint32_t s_use[2] = {0, 0};
  uint32_t accw[4] = {0, 0, 0, 0};     // MMA: writes started into quarter q; epilogue: acc_full waits of quarter q
  uint32_t abuf = 0;                   // A tiles produced (epilogue) / consumed (MMA): buffer = abuf & 1
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

// ---- producer: one layer's block stream through the ring ----
__device__ __forceinline__ void ts_produce(const uint8_t* src, const TsShape s, uint8_t* ring, TsCtl* ctl, TsPipe& pp) {
  ts_for_each_block(s, [&](int, int ss, int k) {
    const uint32_t bytes = 64u * (uint32_t)(ss ? ts_ss_klen(s, k) : 64) * 2u;
    const uint32_t slot = pp.blk % TS_NSLOT, phase = (pp.blk / TS_NSLOT) & 1;
    mbar_wait(&ctl->empty[slot], phase ^ 1);
    mbar_arrive_expect_tx(&ctl->full[slot], bytes);
    bulk_g2s(ring + (size_t)slot * TS_SLOT, src, bytes, &ctl->full[slot]);
    src += bytes;
    ++pp.blk;
  });
}

// ---- MMA issuer: one layer ----
__device__ __forceinline__ void ts_mma_layer(const TsShape s, uint32_t acat_base, uint32_t ring_base, uint32_t tmem_base,
                                             TsCtl* ctl, TsPipe& pp, unsigned long long* tl = nullptr, int* tn = nullptr,
                                             int lid = 0) {
  const uint32_t idesc = umma_idesc_bf16(TILE, 64);
  const uint32_t a_tmem = tmem_base + TS_ACOL + (pp.abuf & 1u) * 128u;
  auto block = [&](int q, int ss, int k) {
    const uint32_t klen = (uint32_t)(ss ? ts_ss_klen(s, k) : 64);
    const uint32_t slot = pp.blk % TS_NSLOT, phase = (pp.blk / TS_NSLOT) & 1;
    mbar_wait(&ctl->full[slot], phase);
    tc_fence_after();
    const uint32_t b_base = ring_base + slot * TS_SLOT;
    const bool first = ss ? (k == 0) : (s.nss == 0 && k == 0);    // first block of this quarter in this layer
    for (uint32_t t = 0; t < klen / 16; ++t) {
      const uint64_t db = op_desc(b_base + (2u * t) * 128u, 128u, klen * 16u);
      const uint32_t acc = (first && t == 0) ? 0u : 1u;
      if (ss) umma_bf16(tmem_base + (uint32_t)q * 64u, op_desc(acat_base + (8u * (uint32_t)k + 2u * t) * 128u, 128u, TS_SBO), db, idesc, acc);
      else umma_bf16_ts(tmem_base + (uint32_t)q * 64u, a_tmem + (uint32_t)k * 32u + t * 8u, db, idesc, acc);
    }
    umma_commit(&ctl->empty[slot]);
    ++pp.blk;
    const bool last = ss ? (s.nts == 0 && k == s.nss - 1) : (k == s.nts - 1);
    if (last) umma_commit(&ctl->acc_full[q]);
  };
  const int n = s.nq > s.nts ? s.nq : s.nts;      // <= 4
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= n) break;
    if (c < s.nts) { mbar_wait(&ctl->a_ready[c], pp.a_use[c] & 1); ++pp.a_use[c]; }
    if (c < s.nq) { mbar_wait(&ctl->acc_free[c], (pp.accw[c] & 1) ^ 1); ++pp.accw[c]; }
    if (c == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (i < s.nss) { mbar_wait(&ctl->s_ready[i], pp.s_use[i] & 1); ++pp.s_use[i]; }
    }
    tc_fence_after();
    if (tn) tl_mark(tl, 1, *tn, 100 * lid + 10 + c);
    if (c < s.nts) for (int q = 0; q < ts_imin(c, s.nq); ++q) block(q, 0, c);
    if (c < s.nq) {
      for (int i = 0; i < s.nss; ++i) block(c, 1, i);
      for (int k = 0; k <= ts_imin(c, s.nts - 1); ++k) block(c, 0, k);
    }
  }
  if (s.nts) ++pp.abuf;
  if (tn) tl_mark(tl, 1, *tn, 100 * lid + 20);
}

// ---- epilogue helpers ----
__device__ __forceinline__ void ts_wait_acc(TsCtl* ctl, TsPipe& pp, int q) {
  mbar_wait_backoff(&ctl->acc_full[q], pp.accw[q] & 1);
  ++pp.accw[q];
  tc_fence_after();
}
// all lanes have finished their tcgen05.ld of quarter c (and tcgen05.st of A chunk c when `wrote_a`)
__device__ __forceinline__ void ts_signal(TsCtl* ctl, int c, int lane, bool wrote_a, bool drained) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (wrote_a) mbar_arrive(&ctl->a_ready[c]);
    if (drained) mbar_arrive(&ctl->acc_free[c]);
  }
}
__device__ __forceinline__ void ts_signal_smem(TsCtl* ctl, int i, int lane) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(&ctl->s_ready[i]);
}
// hidden layer: y = act(acc + bias) -> bf16 -> TMEM A buffer (pp.abuf & 1); thread owns columns [64c + 16cs, +16)
template <bool RELU>
__device__ __forceinline__ void ts_epi_hidden(uint32_t tmem_base, const float* sb, const EpiCtx& ec, TsCtl* ctl, TsPipe& pp,
                                              unsigned long long* tl = nullptr, int* tn = nullptr, int lid = 0) {
  const uint32_t a_w = tmem_base + ec.lane_base + TS_ACOL + (pp.abuf & 1u) * 128u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ts_wait_acc(ctl, pp, c);
    if (tn) tl_mark(tl, 0, *tn, 100 * lid + 1 + c);
    const int col0 = c * 64 + ec.cs * 16;
    uint32_t v[16];
    tmem_ld16(tmem_base + ec.lane_base + (uint32_t)col0, v);
    tmem_ld_wait();
    const float4* b4 = reinterpret_cast<const float4*>(sb + col0);
    uint32_t pk[8];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 b = b4[j4];
      pk[2 * j4] = pack2<RELU>(__uint_as_float(v[4 * j4 + 0]) + b.x, __uint_as_float(v[4 * j4 + 1]) + b.y);
      pk[2 * j4 + 1] = pack2<RELU>(__uint_as_float(v[4 * j4 + 2]) + b.z, __uint_as_float(v[4 * j4 + 3]) + b.w);
    }
    tmem_st8(a_w + (uint32_t)(c * 32 + ec.cs * 8), pk);
    tmem_st_wait();
    ts_signal(ctl, c, ec.lane, true, true);
    if (tn) tl_mark(tl, 0, *tn, 100 * lid + 5 + c);
  }
  ++pp.abuf;
}

template <int FD>
__global__ void __launch_bounds__(THREADS, 1) k_back_ts(TcParams P, TileTable tt, RowIO io) {
  const float* __restrict__ x = io.x;
  const float* __restrict__ gate = io.gate;
  const float* __restrict__ noise = io.noise;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TsCtl* ctl = reinterpret_cast<TsCtl*>(smem + TSM_CTL);
  if (threadIdx.x == 0) {
    for (int i = 0; i < TS_NSLOT; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_free[i], EPI_WARPS);
      mbar_init(&ctl->a_ready[i], EPI_WARPS);
    }
    mbar_init(&ctl->s_ready[0], EPI_WARPS);
    mbar_init(&ctl->s_ready[1], EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  float* sbias = reinterpret_cast<float*>(smem + TSM_BIAS);
  float* svec = reinterpret_cast<float*>(smem + TSM_VEC);
  float* sred = reinterpret_cast<float*>(smem + TSM_RED);
  float *s_wsig = svec, *s_wcol = svec + 256;
  const int H2 = P.hidden2;
  for (int i = threadIdx.x; i < MW; i += THREADS) s_wsig[i] = P.fblob[P.o_wsig + i];
  for (int i = threadIdx.x; i < 3 * H2; i += THREADS) s_wcol[i] = P.fblob[P.o_wcol + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t acat_base = smem_u32(smem + TSM_ACAT), ring_base = smem_u32(smem + TSM_RING);
  const int n_tiles = *tt.n_tiles;
  const int NE = P.n_expert;
  const int kss_xyz = (int)P.front[0].K16, kss_cat = (int)P.back[1].K16 - MW;
  const TsShape SH_XYZ = {4, 0, (kss_xyz + 63) / 64, kss_xyz};
  const TsShape SH_EXP = {4, 4, 0, 0};
  const TsShape SH_SKIP = {4, 4, (kss_xyz + 63) / 64, kss_xyz};
  const TsShape SH_L1 = {4, 4, 0, 0};
  const TsShape SH_L2 = {H2 / 64, 4, (kss_cat + 63) / 64, kss_cat};
  TsPipe pp;

  if (warp == 0) {
    if (lane == 0)
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        const int e = tt.tile_expert[t];
        if (e >= 0) {
          ts_produce(P.tsblob + P.ts_xyz_off, SH_XYZ, smem + TSM_RING, ctl, pp);
          for (int l = 0; l < NE; ++l)
            ts_produce(P.tsblob + P.ts_expert_off[l] + (size_t)e * P.ts_expert_stride, l == P.skip_layer ? SH_SKIP : SH_EXP,
                       smem + TSM_RING, ctl, pp);
        }
        ts_produce(P.tsblob + P.ts_back_off[0], SH_L1, smem + TSM_RING, ctl, pp);
        ts_produce(P.tsblob + P.ts_back_off[1], SH_L2, smem + TSM_RING, ctl, pp);
      }
  } else if (warp == 1) {
    if (lane == 0) {
      int tn = 0;
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        const int e = tt.tile_expert[t];
        tl_mark(P.tl, 1, tn, 1);
        if (e >= 0) {
          ts_mma_layer(SH_XYZ, acat_base, ring_base, tmem_base, ctl, pp, P.tl, &tn, 1);
          for (int l = 0; l < NE; ++l)
            ts_mma_layer(l == P.skip_layer ? SH_SKIP : SH_EXP, acat_base, ring_base, tmem_base, ctl, pp, P.tl, &tn, 2 + l);
        }
        ts_mma_layer(SH_L1, acat_base, ring_base, tmem_base, ctl, pp, P.tl, &tn, 20);
        ts_mma_layer(SH_L2, acat_base, ring_base, tmem_base, ctl, pp, P.tl, &tn, 21);
      }
    }
  } else {
    EpiCtx ec;
    ec.remote_a_ready = 0; ec.lane = lane; ec.q = warp & 3; ec.cs = (warp - 2) >> 2; ec.row = ec.q * 32 + lane;
    ec.et = (int)threadIdx.x - 64; ec.lane_base = (uint32_t)(ec.q * 32) << 16;
    const int row = ec.row;
    const float b_sig = P.fblob[P.o_bsig];
    const float b_col0 = P.fblob[P.o_bcol], b_col1 = P.fblob[P.o_bcol + 1], b_col2 = P.fblob[P.o_bcol + 2];
    uint32_t li = 0;      // bias double buffer
    struct RowIn { int e, sidx; float g, d0, d1, d2, x0, x1, x2; int ai; };
    auto fetch_row = [&](int t) {
      RowIn r;
      r.e = -1; r.sidx = -1; r.g = 0.f; r.d0 = r.d1 = r.d2 = 0.f; r.x0 = r.x1 = r.x2 = 0.f; r.ai = 0;
      if (t < n_tiles) {
        r.e = tt.tile_expert[t];
        if (row < tt.tile_rows[t]) r.sidx = tt.row2sample[tt.tile_row0[t] + row];
        if (r.sidx >= 0) {
          const float* xr = x + (int64_t)r.sidx * io.x_stride;
          if (r.e >= 0) r.g = gate[(int64_t)r.sidx * io.g_stride];
          if (ec.cs == 0 && r.e >= 0) { r.x0 = xr[0]; r.x1 = xr[1]; r.x2 = xr[2]; }
          if (ec.cs == 1) {
            r.d0 = xr[P.x_cols - 4]; r.d1 = xr[P.x_cols - 3]; r.d2 = xr[P.x_cols - 2];
            r.ai = min(max((int)xr[P.x_cols - 1], 0), P.appearance_count - 1);
          }
        }
      }
      return r;
    };
    RowIn nxt = fetch_row((int)blockIdx.x);
    int tn = 0;
    unsigned long long* tl = (warp == 2 && lane == 0) ? P.tl : nullptr;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
      const RowIn cur = nxt;
      const int e = cur.e, sidx = cur.sidx;
      tl_mark(tl, 0, tn, 1);
      const bool valid = sidx >= 0;
      const float g = cur.g;
      // [PE(dir) | appearance | 0-pad] -> cat block (cs == 1 threads)
      auto write_cat = [&]() {
        if (ec.cs == 1) {
          constexpr int NDIR = 3 + 6 * FD;
          __align__(16) __nv_bfloat16 cat[TS_CAT_COLS];
#pragma unroll
          for (int i = 0; i < (int)TS_CAT_COLS; ++i) cat[i] = __float2bfloat16_rn(0.f);
          if (valid) {
            float dvec[3] = {cur.d0, cur.d1, cur.d2};
            pe_to_bf16<FD>(dvec, cat);
            const float4* er = reinterpret_cast<const float4*>(P.emb_a + (int64_t)cur.ai * P.appearance_dim);
            for (int i = 0; i < P.appearance_dim / 4; ++i) {
              const float4 f = er[i];
              cat[NDIR + 4 * i + 0] = __float2bfloat16_rn(f.x);
              cat[NDIR + 4 * i + 1] = __float2bfloat16_rn(f.y);
              cat[NDIR + 4 * i + 2] = __float2bfloat16_rn(f.z);
              cat[NDIR + 4 * i + 3] = __float2bfloat16_rn(f.w);
            }
          }
          ts_cat_store_row(acat_base, row, cat, kss_cat / 8);
        }
      };
      float sig_acc = 0.f;
      if (e >= 0) {
        // ---- PE(xyz) -> cat block: operand of the xyz layer now and of the skip term later ----
        if (ec.cs == 0) {
          constexpr int NPE = 3 + 6 * 12, NPAD = (NPE + 15) / 16 * 16;
          float pxyz[3] = {cur.x0, cur.x1, cur.x2};
          __align__(16) __nv_bfloat16 pe[NPAD];
          pe_to_bf16<12>(pxyz, pe);
#pragma unroll
          for (int i = NPE; i < NPAD; ++i) pe[i] = __float2bfloat16_rn(0.f);
          ts_cat_store_row(acat_base, row, pe, NPAD / 8);
        }
        for (int i = 0; i < SH_XYZ.nss; ++i) ts_signal_smem(ctl, i, lane);
        nxt = fetch_row(t + (int)gridDim.x);
        // ---- xyz layer (act none): h -> A ----
        {
          const int buf = (int)(li & 1);
          epi_load_bias(P.fblob + P.front[0].b_off, MW, sbias, buf, ec.et);
          tl_mark(tl, 0, tn, 2);
          ts_epi_hidden<false>(tmem_base, sbias + buf * 256, ec, ctl, pp, tl, &tn, 1);
          if (P.skip_layer == 0) for (int i = 0; i < SH_XYZ.nss; ++i) ts_signal_smem(ctl, i, lane);
          ++li;
        }
        for (int l = 0; l < NE; ++l, ++li) {
          const int buf = (int)(li & 1);
          const bool skip_here = (l == P.skip_layer);
          epi_load_bias(skip_here ? (P.fblob + P.o_b3x + (size_t)e * MW)
                                  : (P.fblob + P.expert[l].b_off + (size_t)e * P.expert_b_stride), MW, sbias, buf, ec.et);
          const float* sb = sbias + buf * 256;
          if (l < NE - 1) {
            ts_epi_hidden<true>(tmem_base, sb, ec, ctl, pp, tl, &tn, 2 + l);
            if (skip_here) write_cat();          // every MMA of the skip layer has retired: the PE(xyz) block is free
            if (l + 1 == P.skip_layer) for (int i = 0; i < SH_XYZ.nss; ++i) ts_signal_smem(ctl, i, lane);
          } else {
            // last expert layer (no activation) -> combine: y = bf16(gate * bf16(out)) -> ReLU -> A; sigma head on the fly
            const uint32_t a_w = tmem_base + ec.lane_base + TS_ACOL + (pp.abuf & 1u) * 128u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              ts_wait_acc(ctl, pp, c);
              const int col0 = c * 64 + ec.cs * 16;
              uint32_t v[16];
              tmem_ld16(tmem_base + ec.lane_base + (uint32_t)col0, v);
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
              tmem_st8(a_w + (uint32_t)(c * 32 + ec.cs * 8), pk);
              tmem_st_wait();
              ts_signal(ctl, c, lane, true, true);
            }
            ++pp.abuf;
            if (skip_here) write_cat();
          }
        }
      } else {
        // dropped bucket: h = relu(0) = 0 -> zero A operand for layer "1"; the cat block is still needed
        nxt = fetch_row(t + (int)gridDim.x);
        const uint32_t a_w = tmem_base + ec.lane_base + TS_ACOL + (pp.abuf & 1u) * 128u;
        const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 4; ++c) tmem_st8(a_w + (uint32_t)(c * 32 + ec.cs * 8), z);
        tmem_st_wait();
        for (int c = 0; c < 4; ++c) ts_signal(ctl, c, lane, true, false);
        ++pp.abuf;
        write_cat();
      }
      sred[(0 * 4 + ec.cs) * 128 + row] = sig_acc;
      // ---- layer "1" (act none) -> A; then release the cat chunks ----
      {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.back[0].b_off, MW, sbias, buf, ec.et);
        ts_epi_hidden<false>(tmem_base, sbias + buf * 256, ec, ctl, pp, tl, &tn, 20);
        for (int i = 0; i < SH_L2.nss; ++i) ts_signal_smem(ctl, i, lane);
        ++li;
      }
      // ---- layer "2" (ReLU) + colour head ----
      {
        const int buf = (int)(li & 1);
        epi_load_bias(P.fblob + P.back[1].b_off, H2, sbias, buf, ec.et);
        const float* sb = sbias + buf * 256;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c >= H2 / 64) break;
          ts_wait_acc(ctl, pp, c);
          const int col0 = c * 64 + ec.cs * 16;
          uint32_t v[16];
          tmem_ld16(tmem_base + ec.lane_base + (uint32_t)col0, v);
          tmem_ld_wait();
          ts_signal(ctl, c, lane, false, true);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = col0 + j;
            const float h2 = bf16_round(fmaxf(__uint_as_float(v[j]) + sb[k], 0.f));
            c0 = fmaf(h2, s_wcol[k], c0);
            c1 = fmaf(h2, s_wcol[H2 + k], c1);
            c2 = fmaf(h2, s_wcol[2 * H2 + k], c2);
          }
        }
        sred[(1 * 4 + ec.cs) * 128 + row] = c0;
        sred[(2 * 4 + ec.cs) * 128 + row] = c1;
        sred[(3 * 4 + ec.cs) * 128 + row] = c2;
        epi_bar_sync();
        if (ec.cs == 0 && valid) {
          auto rsum = [&](int v) { return sred[(v * 4 + 0) * 128 + row] + sred[(v * 4 + 1) * 128 + row] +
                                          sred[(v * 4 + 2) * 128 + row] + sred[(v * 4 + 3) * 128 + row]; };
          float sr = bf16_round(rsum(0) + b_sig);
          if (noise) sr += noise[(int64_t)sidx * io.n_stride];
          const float tt_ = sr - 1.f;
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
        tl_mark(tl, 0, tn, 2200);
        ++li;
      }
    }
    if (io.ep) __threadfence_system();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
  if (io.ep && threadIdx.x == 0) {
    if (atomicAdd(io.done, 1) == (int)gridDim.x - 1) {
      *io.done = 0;
      __threadfence_system();
      for (int w = 0; w < io.world; ++w)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(io.flag_b[w]), "r"(io.epoch) : "memory");
    }
  }
}

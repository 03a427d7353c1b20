// "Wide" fused kernels: model width W = 256 * NH (NH = 2: the Mission-Bay topology, mission_bay.yaml, width 512;
// NH = 1 exists so that the same code can be checked against the width-256 fixtures).
// (included by snb_tc.cu after snb_tc_ts.cuh; shares TcParams, the epilogue helpers and the routing hand-off)
//
// Why a different data flow than the TS kernels: a 128-row tile of a 512-wide layer needs 256 TMEM columns for its
// packed bf16 A operand AND 512 columns for its fp32 accumulators -- more than the 512 columns an SM has.  Here
//   * the accumulators of ONE layer own all of tensor memory: output columns [256h, 256h+256) of half h at TMEM
//     columns [256h, +256);
//   * the A operand lives in shared memory (128 x (W + 96) bf16, UMMA canonical K-major image, 152 KB at W = 512):
//     columns [0, W) = hidden activations, [W, W+96) = PE(xyz) / [PE(dir) | appearance];
//   * a layer is issued half by half (N = 256, K = W in 32-wide weight slices of 16 KB, 3-slot ring); the epilogue
//     drains half 0 into REGISTERS (bias / activation / bf16 pack: 32 registers per thread) while the tensor pipe is
//     still busy with half 1, and writes it into the A tile the moment the layer's last MMA has retired -- the next
//     layer's half-0 MMAs start right away while half 1 is drained, 64 columns at a time.
// Weight images: per (layer, half, K-slice of 32) one contiguous block in the canonical core-matrix layout
// (k_pack_layer_wide), streamed with 1-D bulk copies.
#pragma once

template <int NH>
struct Wide {
  static constexpr int W = 256 * NH;
  static constexpr int CAT = 96;
  static constexpr int KA = W + CAT;
  static constexpr uint32_t SBO = KA / 8 * 128;
  static constexpr uint32_t A_BYTES = TILE / 8 * SBO;
  static constexpr int NCH = W / 64;
  static constexpr int KS = 32;
  static constexpr uint32_t STAGE = 256 * KS * 2;
  static constexpr int NST = (NH == 2) ? 3 : 6;
  static constexpr size_t O_A = 0;
  static constexpr size_t O_RING = A_BYTES;
  static constexpr size_t O_BIAS = O_RING + (size_t)NST * STAGE;
  static constexpr size_t VEC_FLOATS = W + 3 * 256 + 64;
  static constexpr size_t O_VEC = O_BIAS + 2 * W * 4;
  static constexpr size_t O_RED = O_VEC + VEC_FLOATS * 4;
  static constexpr size_t O_CTL = O_RED + SM_RED_FLOATS * 4;
};

struct __align__(16) WCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t acc_full[2];
  uint64_t a_ready[8];      // epilogue -> MMA: 64-column chunk of the hidden part of the A tile written
  uint64_t s_ready[2];      // epilogue -> MMA: chunk of the cat block written ([0,64) and [64,96))
  uint32_t tmem_base;
  uint32_t pad;
};
template <int NH>
constexpr size_t wide_smem_bytes() { return Wide<NH>::O_CTL + sizeof(WCtl) + 1024; }
static_assert(wide_smem_bytes<2>() <= 227 * 1024, "shared memory budget (wide kernel)");

struct WPipe {
  uint32_t stage = 0, phase = 0;       // ring position of the next weight slot to produce / consume
  uint32_t a_use[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t s_use[2] = {0, 0};
  uint32_t acc_use[2] = {0, 0};
};

// wide image of a layer: half h = rows [256h, 256h + Nh) at byte offset h * 256 * K16 * 2; inside a half the K-slices of
// 32 follow each other (slice j at Nh * 64 * j bytes); inside a slice the canonical layout
//   (n/8) * (klen*16) + (kk/8) * 128 + (n%8) * 16 + (kk%8) * 2   bytes
__global__ void k_pack_layer_wide(const float* __restrict__ w, int N, int K, int K16, __nv_bfloat16* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K16; i += gridDim.x * blockDim.x) {
    const int n = i / K16, k = i % K16;
    const float v = (k < K) ? w[(size_t)n * K + k] : 0.f;
    const int h = n / 256, nl = n % 256, Nh = min(256, N - 256 * h);
    const int j = k / 32, kk = k % 32, klen = min(32, K16 - 32 * j);
    const size_t off = (size_t)h * 256 * K16 + (size_t)Nh * 32 * j +
                       ((size_t)(nl / 8) * (klen * 16) + (size_t)(kk / 8) * 128 + (nl % 8) * 16 + (kk % 8) * 2) / 2;
    dst[off] = __float2bfloat16_rn(v);
  }
}
// folded LayerNorm + gate operand (hi / lo rows, see k_pack_gate) in the wide image
__global__ void k_pack_gate_wide(const float* __restrict__ ln_w, const float* __restrict__ wg, int E, int K,
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
    const int j = k / 32, kk = k % 32, klen = min(32, K - 32 * j);
    const size_t off = (size_t)GATE_N * 32 * j + ((size_t)(n / 8) * (klen * 16) + (size_t)(kk / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2) / 2;
    dst[off] = __float2bfloat16_rn(v);
  }
}

// ---- producer: one half of a layer ----
template <int NH>
__device__ __forceinline__ void w_produce(const uint8_t* img, uint32_t Nh, uint32_t K16, uint8_t* ring, WCtl* ctl, WPipe& pp) {
  using C = Wide<NH>;
  const uint32_t nsl = (K16 + C::KS - 1) / C::KS;
  for (uint32_t j = 0; j < nsl; ++j) {
    const uint32_t klen = min((uint32_t)C::KS, K16 - C::KS * j);
    const uint32_t bytes = Nh * klen * 2;
    const uint32_t stage = pp.stage;
    mbar_wait(&ctl->empty[stage], pp.phase ^ 1);
    if (++pp.stage == (uint32_t)C::NST) { pp.stage = 0; pp.phase ^= 1; }
    mbar_arrive_expect_tx(&ctl->full[stage], bytes);
    bulk_g2s(ring + (size_t)stage * C::STAGE, img + (size_t)Nh * C::KS * 2 * j, bytes, &ctl->full[stage]);
  }
}
template <int NH>
__device__ __forceinline__ void w_produce_layer(const uint8_t* wblob, const TcLayer& L, size_t extra, uint8_t* ring, WCtl* ctl,
                                                WPipe& pp, const TcLayer* seg2 = nullptr) {
  for (uint32_t h = 0; h * 256 < L.N; ++h) {
    const uint32_t Nh = min(256u, L.N - 256u * h);
    w_produce<NH>(wblob + L.w_off_w + extra + (size_t)h * 256 * L.K16 * 2, Nh, L.K16, ring, ctl, pp);
    if (seg2) w_produce<NH>(wblob + seg2->w_off_w + (size_t)h * 256 * seg2->K16 * 2, Nh, seg2->K16, ring, ctl, pp);
  }
}

// ---- MMA issuer: a run of K16 input columns starting at A column a_col0 into accumulator d_tmem ----
// wait: first pass over these A columns in this layer (half 0): wait for the chunk barriers
// Executed by all lanes of warp 1 on warp-uniform values; the tcgen05.mma / tcgen05.commit are predicated on `leader`
// (one elected lane): see ts_mma_seg in snb_tc_ts.cuh and profiles/r3g_issue_path.md.
template <int NH>
__device__ __forceinline__ void w_mma_run(uint32_t Nh, uint32_t K16, uint32_t a_col0, uint32_t a_base, uint32_t ring_base,
                                          uint32_t d_tmem, bool first, bool wait, WCtl* ctl, WPipe& pp, bool leader) {
  using C = Wide<NH>;
  const uint32_t nsl = (K16 + C::KS - 1) / C::KS;
  const uint32_t idesc = umma_idesc_bf16(TILE, (int)Nh);
  for (uint32_t j = 0; j < nsl; ++j) {
    const uint32_t klen = min((uint32_t)C::KS, K16 - C::KS * j);
    const uint32_t col = a_col0 + C::KS * j;
    if (wait && (col & 63u) == 0u) {
      if (col < (uint32_t)C::W) {
        const uint32_t c = col >> 6;
        mbar_wait(&ctl->a_ready[c], pp.a_use[c] & 1);
        ++pp.a_use[c];
      } else {
        const uint32_t c = (col - C::W) >> 6;
        mbar_wait(&ctl->s_ready[c], pp.s_use[c] & 1);
        ++pp.s_use[c];
      }
    }
    const uint32_t stage = pp.stage;
    mbar_wait(&ctl->full[stage], pp.phase);
    if (++pp.stage == (uint32_t)C::NST) { pp.stage = 0; pp.phase ^= 1; }
    tc_fence_after();
    // one K = 16 step = two 128-byte core matrices further in both operands: +16 in the descriptors' 16-byte address field
    const uint64_t da0 = op_desc(a_base + (col >> 3) * 128u, 128u, C::SBO);
    const uint64_t db0 = op_desc(ring_base + stage * C::STAGE, 128u, klen * 16u);
    const uint32_t acc0 = (first && j == 0) ? 0u : 1u;
    if (leader) {
      if (klen == 32u) {
        umma_bf16(d_tmem, da0, db0, idesc, acc0);
        umma_bf16(d_tmem, da0 + 16u, db0 + 16u, idesc, 1u);
      } else {
        for (uint32_t t = 0; t < klen / 16; ++t) umma_bf16(d_tmem, da0 + 16u * t, db0 + 16u * t, idesc, t ? 1u : acc0);
      }
      umma_commit(&ctl->empty[stage]);
    }
  }
}
// one layer: every half (N = 256 or the whole narrow layer) over A columns [a_col0, a_col0 + K16) (+ a second run `seg2`
// over the cat block: the skip term), accumulators of half h at TMEM columns [256h, ...)
template <int NH>
__device__ __forceinline__ void w_mma_layer(const TcLayer& L, uint32_t a_col0, uint32_t a_base, uint32_t ring_base,
                                            uint32_t tmem_base, WCtl* ctl, WPipe& pp, bool leader, const TcLayer* seg2 = nullptr) {
  using C = Wide<NH>;
  for (uint32_t h = 0; h * 256 < L.N; ++h) {
    const uint32_t Nh = min(256u, L.N - 256u * h);
    w_mma_run<NH>(Nh, L.K16, a_col0, a_base, ring_base, tmem_base + 256u * h, true, h == 0, ctl, pp, leader);
    if (seg2) w_mma_run<NH>(Nh, seg2->K16, (uint32_t)C::W, a_base, ring_base, tmem_base + 256u * h, false, h == 0, ctl, pp, leader);
    if (leader) umma_commit(&ctl->acc_full[h]);
  }
}

// ---- epilogue helpers ----
template <int NH>
__device__ __forceinline__ uint32_t w_a_addr(uint32_t a_base, int row, int col8) {
  return a_base + (uint32_t)(row >> 3) * Wide<NH>::SBO + (uint32_t)col8 * 128u + (uint32_t)(row & 7) * 16u;
}
__device__ __forceinline__ void w_wait_acc(WCtl* ctl, WPipe& pp, int h) {
  mbar_wait_backoff(&ctl->acc_full[h], pp.acc_use[h] & 1);
  ++pp.acc_use[h];
  tc_fence_after();
}
// my st.shared of an A chunk -> visible to the tensor pipe; my tcgen05.ld of the accumulator it replaces are done
__device__ __forceinline__ void w_signal(uint64_t* bar, int lane) {
  fence_proxy_async_smem();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
template <int NH>
__device__ __forceinline__ void w_store_row(uint32_t a_base, int row, int col8, const __nv_bfloat16* vals, int n8) {
  const uint4* v = reinterpret_cast<const uint4*>(vals);
  for (int g = 0; g < n8; ++g) {
    const uint4 t = v[g];
    st_shared_v4(w_a_addr<NH>(a_base, row, col8 + g), t.x, t.y, t.z, t.w);
  }
}
template <int NH>
__device__ __forceinline__ void w_load_bias(const float* __restrict__ bias, int N, float* sbias, int slot, int et) {
  float* dst = sbias + slot * Wide<NH>::W;
  if (et < N) dst[et] = bias[et];
  epi_bar_sync();
}

// epilogue of a hidden layer.  MODE 0: y = acc + b; 1: relu(acc + b); 2: combine: relu(bf16(gate * bf16(acc + b))) and the
// sigma dot product (r0 += y . wsig); 3: y = bf16(acc + b) and the LayerNorm statistics (r0 += y, r1 += y^2).
// `after_mma()` runs once every MMA of the layer has retired (the A tile and the cat block are free).
template <int NH, int MODE, typename F>
__device__ __forceinline__ void w_epi_layer(uint32_t tlane /* tmem_base + lane_base */, const float* sb, uint32_t a_base,
                                            const EpiCtx& ec, WCtl* ctl, WPipe& pp, float g, const float* s_wsig, float& r0,
                                            float& r1, F&& after_mma) {
  auto convert = [&](const uint32_t (&v)[16], int col0, uint32_t (&pk)[8]) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float f0 = __uint_as_float(v[j]) + sb[col0 + j], f1 = __uint_as_float(v[j + 1]) + sb[col0 + j + 1];
      if (MODE == 0) pk[j / 2] = pack2<false>(f0, f1);
      else if (MODE == 1) pk[j / 2] = pack2<true>(f0, f1);
      else if (MODE == 2) {
        float a, b;
        ts_round2<false>(f0, f1, a, b);
        pk[j / 2] = ts_round2<true>(a * g, b * g, a, b);
        r0 = fmaf(a, s_wsig[col0 + j], r0);
        r0 = fmaf(b, s_wsig[col0 + j + 1], r0);
      } else {
        float a, b;
        pk[j / 2] = ts_round2<false>(f0, f1, a, b);
        r0 += a + b;
        r1 = fmaf(a, a, fmaf(b, b, r1));
      }
    }
  };
  auto store = [&](int col0, const uint32_t (&pk)[8]) {
    st_shared_v4(w_a_addr<NH>(a_base, ec.row, col0 / 8), pk[0], pk[1], pk[2], pk[3]);
    st_shared_v4(w_a_addr<NH>(a_base, ec.row, col0 / 8 + 1), pk[4], pk[5], pk[6], pk[7]);
  };
  w_wait_acc(ctl, pp, 0);
  if (NH == 1) {
    // one accumulator buffer: the next layer's first MMA overwrites all 256 columns, so every chunk is drained into
    // registers before the first chunk is released
    after_mma();
    uint32_t keep[4][8];
    uint32_t v[2][16];
    tmem_ld16(tlane + (uint32_t)(ec.cs * 16), v[0]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld_wait16(v[c & 1]);
      if (c < 3) tmem_ld16(tlane + (uint32_t)((c + 1) * 64 + ec.cs * 16), v[(c + 1) & 1]);
      convert(v[c & 1], c * 64 + ec.cs * 16, keep[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) store(c * 64 + ec.cs * 16, keep[c]);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (ec.lane == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) mbar_arrive(&ctl->a_ready[c]);
    }
  } else {
    // half 0 -> registers while the tensor pipe works on half 1
    uint32_t keep[4][8];
    {
      uint32_t v[2][16];
      tmem_ld16(tlane + (uint32_t)(ec.cs * 16), v[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait16(v[c & 1]);
        if (c < 3) tmem_ld16(tlane + (uint32_t)((c + 1) * 64 + ec.cs * 16), v[(c + 1) & 1]);
        convert(v[c & 1], c * 64 + ec.cs * 16, keep[c]);
      }
    }
    w_wait_acc(ctl, pp, 1);
    after_mma();
#pragma unroll
    for (int c = 0; c < 4; ++c) store(c * 64 + ec.cs * 16, keep[c]);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (ec.lane == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) mbar_arrive(&ctl->a_ready[c]);
    }
    uint32_t v[2][16];
    tmem_ld16(tlane + 256u + (uint32_t)(ec.cs * 16), v[0]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld_wait16(v[c & 1]);
      if (c < 3) tmem_ld16(tlane + 256u + (uint32_t)((c + 1) * 64 + ec.cs * 16), v[(c + 1) & 1]);
      uint32_t pk[8];
      convert(v[c & 1], 256 + c * 64 + ec.cs * 16, pk);
      store(256 + c * 64 + ec.cs * 16, pk);
      w_signal(&ctl->a_ready[4 + c], ec.lane);
    }
  }
}

// mip-NeRF integrated positional encoding (models/nerf.py:28-56): [x, sin(2^k x) e^(-4^k cov / 2), cos(2^k x) e^(-4^k cov / 2)]_k
template <int F>
__device__ __forceinline__ void pe_mip_to_bf16(const float (&p)[3], const float (&cv)[3], __nv_bfloat16* dst) {
  float s[3], c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    dst[a] = __float2bfloat16_rn(p[a]);
    sincosf(p[a], &s[a], &c[a]);
  }
  float fw = 1.f;
#pragma unroll
  for (int k = 0; k < F; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float dmp = expf((-0.5f * fw) * cv[a]);
      dst[3 + 6 * k + a] = __float2bfloat16_rn(s[a] * dmp);
      dst[3 + 6 * k + 3 + a] = __float2bfloat16_rn(c[a] * dmp);
      const float s2 = 2.f * s[a] * c[a];
      const float c2 = 1.f - 2.f * s[a] * s[a];
      s[a] = s2;
      c[a] = c2;
    }
    fw *= 4.f;
  }
}

template <int NH>
__device__ __forceinline__ WCtl* w_setup(uint8_t* smem, int warp) {
  WCtl* ctl = reinterpret_cast<WCtl*>(smem + Wide<NH>::O_CTL);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); mbar_init(&ctl->a_ready[i], EPI_WARPS); }
    mbar_init(&ctl->acc_full[0], 1);
    mbar_init(&ctl->acc_full[1], 1);
    mbar_init(&ctl->s_ready[0], EPI_WARPS);
    mbar_init(&ctl->s_ready[1], EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  return ctl;
}

// ------------------------------------------------------------------------------------------------------------
// launch #2, wide: gather -> PE -> xyz layer -> expert stack (skip as a second run over the PE block) -> combine ->
// sigma head -> layer "1" -> [PE(dir) | appearance] -> layer "2" -> colour head -> scatter
// ------------------------------------------------------------------------------------------------------------
template <int NH, int FX, int FD>
__global__ void __launch_bounds__(THREADS, 1) k_back_wide(TcParams P, TileTable tt, RowIO io) {
  using C = Wide<NH>;
  const float* __restrict__ x = io.x;
  const float* __restrict__ gate = io.gate;
  const float* __restrict__ noise = io.noise;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  WCtl* ctl = w_setup<NH>(smem, warp);
  float* sbias = reinterpret_cast<float*>(smem + C::O_BIAS);
  float* svec = reinterpret_cast<float*>(smem + C::O_VEC);
  float* sred = reinterpret_cast<float*>(smem + C::O_RED);
  float *s_wsig = svec, *s_wcol = svec + C::W;
  const int H2 = P.hidden2;
  for (int i = threadIdx.x; i < C::W; i += THREADS) s_wsig[i] = P.fblob[P.o_wsig + i];
  for (int i = threadIdx.x; i < 3 * H2; i += THREADS) s_wcol[i] = P.fblob[P.o_wcol + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t a_base = smem_u32(smem + C::O_A), ring_base = smem_u32(smem + C::O_RING);
  const int n_tiles = *tt.n_tiles;
  const int NE = P.n_expert;
  const uint32_t K_xyz = P.front[0].K16, K_cat = P.back[1].K16 - C::W;
  WPipe pp;

  if (warp == 0) {
    if (lane == 0)
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        const int e = tt.tile_expert[t];
        if (e >= 0) {
          w_produce_layer<NH>(P.wblob_w, P.front[0], 0, smem + C::O_RING, ctl, pp);
          for (int l = 0; l < NE; ++l)
            w_produce_layer<NH>(P.wblob_w, P.expert[l], (size_t)e * P.expert_w_stride_w, smem + C::O_RING, ctl, pp,
                                l == P.skip_layer ? &P.front[0] : nullptr);
        }
        w_produce_layer<NH>(P.wblob_w, P.back[0], 0, smem + C::O_RING, ctl, pp);
        w_produce_layer<NH>(P.wblob_w, P.back[1], 0, smem + C::O_RING, ctl, pp);
      }
  } else if (warp == 1) {
    const bool leader = elect_one();               // warp-uniform issue loop, see w_mma_run
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int n_tiles_u = __shfl_sync(0xffffffffu, n_tiles, 0);
    for (int t = (int)blockIdx.x; t < n_tiles_u; t += (int)gridDim.x) {
      const int e = __shfl_sync(0xffffffffu, tt.tile_expert[t], 0);
      if (e >= 0) {
        w_mma_layer<NH>(P.front[0], (uint32_t)C::W, a_base, ring_base, tmem_u, ctl, pp, leader);
        for (int l = 0; l < NE; ++l)
          w_mma_layer<NH>(P.expert[l], 0u, a_base, ring_base, tmem_u, ctl, pp, leader, l == P.skip_layer ? &P.front[0] : nullptr);
      }
      w_mma_layer<NH>(P.back[0], 0u, a_base, ring_base, tmem_u, ctl, pp, leader);
      w_mma_layer<NH>(P.back[1], 0u, a_base, ring_base, tmem_u, ctl, pp, leader);
    }
  } else {
    EpiCtx ec;
    ec.remote_a_ready = 0; ec.lane = lane; ec.q = warp & 3; ec.cs = (warp - 2) >> 2; ec.row = ec.q * 32 + lane;
    ec.et = (int)threadIdx.x - 64; ec.lane_base = (uint32_t)(ec.q * 32) << 16;
    const int row = ec.row;
    const uint32_t tlane = tmem_base + ec.lane_base;
    const float b_sig = P.fblob[P.o_bsig];
    const float b_col0 = P.fblob[P.o_bcol], b_col1 = P.fblob[P.o_bcol + 1], b_col2 = P.fblob[P.o_bcol + 2];
    const int xd = P.mip ? 6 : 3;
    uint32_t li = 0;
    struct RowIn { int e, sidx; float g, d0, d1, d2, x0, x1, x2, v0, v1, v2; int ai; };
    auto fetch_row = [&](int t) {
      RowIn r;
      r.e = -1; r.sidx = -1; r.g = 0.f; r.d0 = r.d1 = r.d2 = 0.f; r.x0 = r.x1 = r.x2 = 0.f; r.v0 = r.v1 = r.v2 = 0.f; r.ai = 0;
      if (t < n_tiles) {
        r.e = tt.tile_expert[t];
        if (row < tt.tile_rows[t]) r.sidx = tt.row2sample[tt.tile_row0[t] + row];
        if (r.sidx >= 0) {
          const float* xr = x + (int64_t)r.sidx * io.x_stride;
          if (r.e >= 0) r.g = io.wsel ? sel_gate(io.wsel[r.sidx]) : gate[(int64_t)r.sidx * io.g_stride];
          if (ec.cs == 0 && r.e >= 0) {
            r.x0 = xr[0]; r.x1 = xr[1]; r.x2 = xr[2];
            if (P.mip) { r.v0 = xr[3]; r.v1 = xr[4]; r.v2 = xr[5]; }
          }
          if (ec.cs == 1) {
            r.d0 = xr[xd]; r.d1 = xr[xd + 1]; r.d2 = xr[xd + 2];
            r.ai = min(max((int)xr[P.x_cols - 1], 0), P.appearance_count - 1);
          }
        }
      }
      return r;
    };
    const int n_sxyz = ((int)K_xyz + 63) / 64, n_scat = ((int)K_cat + 63) / 64;
    auto signal_cat = [&](int n) { for (int i = 0; i < n; ++i) w_signal(&ctl->s_ready[i], lane); };
    RowIn nxt = fetch_row((int)blockIdx.x);
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
      const RowIn cur = nxt;
      const int e = cur.e, sidx = cur.sidx;
      const bool valid = sidx >= 0;
      nxt = fetch_row(t + (int)gridDim.x);
      // [PE(dir) | appearance | 0-pad] -> cat block: the appearance part comes as 16-byte chunks from the pre-shifted bf16
      // table (snb_tc.cu: k_pack_emb_cat; a gather of random fp32 rows through a local array cost ~5K clk per tile,
      // profiles/r3g_issue_path.md), PE(dir) is merged into the chunks below column NDIR
      auto write_cat = [&]() {
        if (ec.cs == 1) {
          constexpr int NDIR = 3 + 6 * FD;
          float pe[NDIR];
          {
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
          const uint4* er = reinterpret_cast<const uint4*>(P.emb_cat + (int64_t)cur.ai * P.cat_cols);
          const int n8 = (int)K_cat / 8;
#pragma unroll
          for (int g = 0; g < C::CAT / 8; ++g) {
            if (g >= n8) continue;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (8 * g + 8 > NDIR && valid && P.emb_cat) q = __ldg(er + g);
            if (8 * g < NDIR) {
              uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c0 = 8 * g + 2 * i, c1 = c0 + 1;
                const uint32_t pp2 = pack2<false>(c0 < NDIR ? pe[c0 < NDIR ? c0 : 0] : 0.f, c1 < NDIR ? pe[c1 < NDIR ? c1 : 0] : 0.f);
                w[i] = valid ? (w[i] | pp2) : 0u;
              }
              q = make_uint4(w[0], w[1], w[2], w[3]);
            }
            st_shared_v4(w_a_addr<NH>(a_base, row, C::W / 8 + g), q.x, q.y, q.z, q.w);
          }
        }
      };
      float sig_acc = 0.f, dummy = 0.f;
      if (e >= 0) {
        // PE(xyz) -> cat block (the previous tile's layer "2" retired before this tile started)
        if (ec.cs == 0) {
          constexpr int NPE = 3 + 6 * FX, NPAD = (NPE + 15) / 16 * 16;
          float pxyz[3] = {cur.x0, cur.x1, cur.x2};
          __align__(16) __nv_bfloat16 pe[NPAD];
          if (P.mip) { float cv[3] = {cur.v0, cur.v1, cur.v2}; pe_mip_to_bf16<FX>(pxyz, cv, pe); }
          else pe_to_bf16<FX>(pxyz, pe);
#pragma unroll
          for (int i = NPE; i < NPAD; ++i) pe[i] = __float2bfloat16_rn(0.f);
          w_store_row<NH>(a_base, row, C::W / 8, pe, NPAD / 8);
        }
        signal_cat(n_sxyz);
        // ---- xyz layer (act none) ----
        {
          w_load_bias<NH>(P.fblob + P.front[0].b_off, C::W, sbias, (int)(li & 1), ec.et);
          w_epi_layer<NH, 0>(tlane, sbias + (li & 1) * C::W, a_base, ec, ctl, pp, 0.f, s_wsig, dummy, dummy, [] {});
          if (P.skip_layer == 0) signal_cat(n_sxyz);
          ++li;
        }
        for (int l = 0; l < NE; ++l, ++li) {
          const bool skip_here = (l == P.skip_layer);
          w_load_bias<NH>(skip_here ? (P.fblob + P.o_b3x + (size_t)e * C::W)
                                    : (P.fblob + P.expert[l].b_off + (size_t)e * P.expert_b_stride), C::W, sbias, (int)(li & 1), ec.et);
          const float* sb = sbias + (li & 1) * C::W;
          auto after = [&]() { if (skip_here) write_cat(); };      // the PE(xyz) block is free: [PE(dir) | appearance]
          if (l < NE - 1) {
            w_epi_layer<NH, 1>(tlane, sb, a_base, ec, ctl, pp, 0.f, s_wsig, dummy, dummy, after);
            if (l + 1 == P.skip_layer) signal_cat(n_sxyz);
          } else {
            w_epi_layer<NH, 2>(tlane, sb, a_base, ec, ctl, pp, cur.g, s_wsig, sig_acc, dummy, after);
          }
        }
      } else {
        // dropped bucket: h = relu(0) = 0 -> zero A operand for layer "1"
        for (int c = 0; c < C::NCH; ++c) {
          st_shared_v4(w_a_addr<NH>(a_base, row, (c * 64 + ec.cs * 16) / 8), 0u, 0u, 0u, 0u);
          st_shared_v4(w_a_addr<NH>(a_base, row, (c * 64 + ec.cs * 16) / 8 + 1), 0u, 0u, 0u, 0u);
        }
        for (int c = 0; c < C::NCH; ++c) w_signal(&ctl->a_ready[c], lane);
        write_cat();
      }
      sred[(0 * 4 + ec.cs) * 128 + row] = sig_acc;
      // ---- layer "1" (act none); then release the cat chunks ----
      {
        w_load_bias<NH>(P.fblob + P.back[0].b_off, C::W, sbias, (int)(li & 1), ec.et);
        w_epi_layer<NH, 0>(tlane, sbias + (li & 1) * C::W, a_base, ec, ctl, pp, 0.f, s_wsig, dummy, dummy, [] {});
        signal_cat(n_scat);
        ++li;
      }
      // ---- layer "2" (ReLU) + colour head ----
      {
        w_load_bias<NH>(P.fblob + P.back[1].b_off, H2, sbias, (int)(li & 1), ec.et);
        const float* sb = sbias + (li & 1) * C::W;
        w_wait_acc(ctl, pp, 0);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int c = 0; c < H2 / 64; ++c) {
          const int col0 = c * 64 + ec.cs * 16;
          uint32_t v[16];
          tmem_ld16(tlane + (uint32_t)col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const int k = col0 + j;
            float ha, hb;
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
        sred[(1 * 4 + ec.cs) * 128 + row] = c0;
        sred[(2 * 4 + ec.cs) * 128 + row] = c1;
        sred[(3 * 4 + ec.cs) * 128 + row] = c2;
        epi_bar_sync();
        if (ec.cs == 0 && valid) {
          auto rsum = [&](int v) { return sred[(v * 4 + 0) * 128 + row] + sred[(v * 4 + 1) * 128 + row] +
                                          sred[(v * 4 + 2) * 128 + row] + sred[(v * 4 + 3) * 128 + row]; };
          // rounding points of the reference under cuda autocast: see k_back_ts
          float sr = bf16_round(rsum(0) + b_sig);
          if (noise) sr = bf16_round(sr + noise[(int64_t)sidx * io.n_stride]);
          const float tt_ = bf16_round(sr - 1.f);
          const float sigma = (tt_ > 20.f) ? tt_ : log1pf(expf(tt_));
          auto sg = [](float v) { return bf16_round(1.f / (1.f + expf(-bf16_round(v)))); };
          reinterpret_cast<float4*>(io.out)[sidx] = make_float4(sg(rsum(1) + b_col0), sg(rsum(2) + b_col1), sg(rsum(3) + b_col2), sigma);
        }
        epi_bar_sync();
        ++li;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// launch #1, wide: PE -> xyz layer -> external gate MLP -> folded LayerNorm + gate GEMM -> softmax + routing hand-off
// ------------------------------------------------------------------------------------------------------------
template <int NH, int FX>
__global__ void __launch_bounds__(THREADS, 1) k_front_wide(TcParams P, const float* __restrict__ x, int64_t S,
                                                           float* __restrict__ gates, uint32_t* __restrict__ wsel,
                                                           int* __restrict__ hist0, float* __restrict__ pm,
                                                           int32_t* __restrict__ moe_idx) {
  using C = Wide<NH>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  WCtl* ctl = w_setup<NH>(smem, warp);
  float* sbias = reinterpret_cast<float*>(smem + C::O_BIAS);
  float* sred = reinterpret_cast<float*>(smem + C::O_RED);
  float* s_gc = reinterpret_cast<float*>(smem + C::O_VEC);
  if (threadIdx.x < 2 * MAX_E)
    s_gc[threadIdx.x] = P.fblob[(threadIdx.x < MAX_E ? P.o_c0 : P.o_c1 - MAX_E) + threadIdx.x];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_gc + 2 * MAX_E);
  static_assert((2 * MAX_E + MAX_E * SEL_HBINS / 2) * 4 <= C::VEC_FLOATS * 4, "histogram fits in the head-vector block");
  for (int i = threadIdx.x; i < MAX_E * SEL_HBINS / 2; i += THREADS) s_hist[i] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t a_base = smem_u32(smem + C::O_A), ring_base = smem_u32(smem + C::O_RING);
  const int n_tiles = (int)((S + TILE - 1) / TILE);
  const int NL = P.n_front;
  const uint32_t K_xyz = P.front[0].K16;
  const int n_sxyz = ((int)K_xyz + 63) / 64;
  WPipe pp;

  if (warp == 0) {
    if (lane == 0)
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        for (int l = 0; l < NL; ++l) w_produce_layer<NH>(P.wblob_w, P.front[l], 0, smem + C::O_RING, ctl, pp);
        w_produce_layer<NH>(P.wblob_w, P.gate, 0, smem + C::O_RING, ctl, pp);
      }
  } else if (warp == 1) {
    const bool leader = elect_one();               // warp-uniform issue loop, see w_mma_run
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
      w_mma_layer<NH>(P.front[0], (uint32_t)C::W, a_base, ring_base, tmem_u, ctl, pp, leader);
      for (int l = 1; l < NL; ++l) w_mma_layer<NH>(P.front[l], 0u, a_base, ring_base, tmem_u, ctl, pp, leader);
      w_mma_layer<NH>(P.gate, 0u, a_base, ring_base, tmem_u, ctl, pp, leader);
    }
  } else {
    EpiCtx ec;
    ec.remote_a_ready = 0; ec.lane = lane; ec.q = warp & 3; ec.cs = (warp - 2) >> 2; ec.row = ec.q * 32 + lane;
    ec.et = (int)threadIdx.x - 64; ec.lane_base = (uint32_t)(ec.q * 32) << 16;
    const int row = ec.row;
    const uint32_t tlane = tmem_base + ec.lane_base;
    uint32_t li = 0;
    constexpr int NPE = 3 + 6 * FX, NPAD = (NPE + 15) / 16 * 16;
    auto load_xyz = [&](int tile, float (&p)[6]) {
#pragma unroll
      for (int i = 0; i < 6; ++i) p[i] = 0.f;
      const int64_t sr = (int64_t)tile * TILE + row;
      if (ec.cs == 0 && tile < n_tiles && sr < S) {
        const float* xr = x + sr * P.x_cols;
        p[0] = xr[0]; p[1] = xr[1]; p[2] = xr[2];
        if (P.mip) { p[3] = xr[3]; p[4] = xr[4]; p[5] = xr[5]; }
      }
    };
    float pn[6];
    load_xyz((int)blockIdx.x, pn);
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
      const int64_t s = (int64_t)t * TILE + row;
      const bool valid = s < S;
      // PE(xyz) -> cat block (its last reader, the xyz layer of the previous tile, retired long ago)
      if (ec.cs == 0) {
        float pxyz[3] = {pn[0], pn[1], pn[2]};
        __align__(16) __nv_bfloat16 pe[NPAD];
        if (P.mip) { float cv[3] = {pn[3], pn[4], pn[5]}; pe_mip_to_bf16<FX>(pxyz, cv, pe); }
        else pe_to_bf16<FX>(pxyz, pe);
#pragma unroll
        for (int i = NPE; i < NPAD; ++i) pe[i] = __float2bfloat16_rn(0.f);
        w_store_row<NH>(a_base, row, C::W / 8, pe, NPAD / 8);
      }
      for (int i = 0; i < n_sxyz; ++i) w_signal(&ctl->s_ready[i], lane);
      load_xyz(t + (int)gridDim.x, pn);
      float sum = 0.f, sq = 0.f, dummy = 0.f;
      for (int l = 0; l < NL; ++l, ++li) {
        w_load_bias<NH>(P.fblob + P.front[l].b_off, C::W, sbias, (int)(li & 1), ec.et);
        const float* sb = sbias + (li & 1) * C::W;
        if (l == 0) w_epi_layer<NH, 0>(tlane, sb, a_base, ec, ctl, pp, 0.f, nullptr, dummy, dummy, [] {});
        else if (l < NL - 1) w_epi_layer<NH, 1>(tlane, sb, a_base, ec, ctl, pp, 0.f, nullptr, dummy, dummy, [] {});
        else {
          w_epi_layer<NH, 3>(tlane, sb, a_base, ec, ctl, pp, 0.f, nullptr, sum, sq, [] {});
          sred[(0 * 4 + ec.cs) * 128 + row] = sum;
          sred[(1 * 4 + ec.cs) * 128 + row] = sq;
        }
      }
      {
        epi_bar_sync();                       // LayerNorm partial sums of all 4 column sub-slices are in sred
        w_wait_acc(ctl, pp, 0);
        front_softmax_select(P, sred, s_gc, s_hist, tlane, ec, row, lane, valid, s, t, 1.f / C::W, gates, wsel, pm, moe_idx);
        tc_fence_before();
        epi_bar_sync();                       // sred is reused by the next tile
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

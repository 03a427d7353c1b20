// snb_umma_selftest: one 128 x N x K bf16 GEMM tile through exactly the machinery the fused
// kernels rely on -- canonical no-swizzle K-major core-matrix operands in shared memory, UMMA
// shared-memory + instruction descriptors, single-thread tcgen05.mma issue, tcgen05.commit ->
// mbarrier, TMEM allocation and tcgen05.ld epilogue, and (variant bit 1) a 1-D bulk async copy
// of a pre-packed B image, and (variant bit 2) the A operand taken from tensor memory (tcgen05.st by the
// row-owning threads, TS-form tcgen05.mma).  tests/test_gpu_parity.py checks it against a float reference,
// which pins the descriptor encodings on real hardware before the large kernels depend on them.
#include "snb_common.cuh"
#include "snb_umma.cuh"

namespace snb {
using namespace ptx;

// canonical offset of element (r, k) (bf16) for a tile whose 8-row groups are `sbo` bytes apart and whose
// K-adjacent core matrices are `lbo` bytes apart
__device__ __forceinline__ uint32_t canon_off(int r, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)(k >> 3) * lbo + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 7) * 2u;
}

__global__ void k_pack_canonical(const __nv_bfloat16* __restrict__ src, int R, int K, uint32_t lbo, uint32_t sbo,
                                 __nv_bfloat16* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * K; i += gridDim.x * blockDim.x) {
    int r = i / K, k = i % K;
    dst[canon_off(r, k, lbo, sbo) / 2] = src[i];
  }
}

__global__ void __launch_bounds__(128) k_umma_selftest(const __nv_bfloat16* __restrict__ A,
                                                       const __nv_bfloat16* __restrict__ B,
                                                       const __nv_bfloat16* __restrict__ B_packed, int N, int K,
                                                       float* __restrict__ D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_copy;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lbo = 128, sboA = (uint32_t)(K / 8) * 128, sboB = (uint32_t)(K / 8) * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)128 * K * 2;

  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_copy, 1);
    fence_mbar_init();
  }
  const bool a_tmem = (variant & 4) != 0;      // TS form: accumulator columns [0,256), A columns [256, 256+K/2)
  if (warp == 0) {
    if (a_tmem) tmem_alloc<512>(&tmem_base_s);
    else tmem_alloc<256>(&tmem_base_s);
  }
  // A: generic-proxy stores into the canonical layout (what the epilogue warps do in the fused kernels)
  for (int i = threadIdx.x; i < 128 * (K / 8); i += blockDim.x) {
    int r = i / (K / 8), kc = i % (K / 8);
    uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kc * 8);
    *reinterpret_cast<uint4*>(sA + canon_off(r, kc * 8, lbo, sboA)) = v;
  }
  const bool use_bulk = (variant & 2) != 0;
  if (!use_bulk) {
    for (int i = threadIdx.x; i < N * (K / 8); i += blockDim.x) {
      int r = i / (K / 8), kc = i % (K / 8);
      uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)r * K + kc * 8);
      *reinterpret_cast<uint4*>(sB + canon_off(r, kc * 8, lbo, sboB)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (a_tmem) {
    // thread (warp, lane) owns row 32*warp + lane = TMEM lane; two bf16 per 32-bit column, K ascending
    const int r = warp * 32 + lane;
    for (int k = 0; k < K / 16; ++k) {
      const uint4 lo = *reinterpret_cast<const uint4*>(A + (size_t)r * K + k * 16);
      const uint4 hi = *reinterpret_cast<const uint4*>(A + (size_t)r * K + k * 16 + 8);
      uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      tmem_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)k * 8u, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (warp == 1 && lane == 0) {
    if (use_bulk) {
      const uint32_t bytes = (uint32_t)N * K * 2;
      mbar_arrive_expect_tx(&bar_copy, bytes);
      bulk_g2s(sB, B_packed, bytes, &bar_copy);
      mbar_wait(&bar_copy, 0);
    }
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const bool swap = (variant & 1) != 0;
    for (int k = 0; k < K / 16; ++k) {
      // one MMA consumes K=16 = two core matrices along K
      uint32_t a_addr = smem_u32(sA) + (uint32_t)k * 2 * lbo;
      uint32_t b_addr = smem_u32(sB) + (uint32_t)k * 2 * lbo;
      uint64_t da = swap ? umma_smem_desc(a_addr, sboA, lbo) : umma_smem_desc(a_addr, lbo, sboA);
      uint64_t db = swap ? umma_smem_desc(b_addr, sboB, lbo) : umma_smem_desc(b_addr, lbo, sboB);
      if (a_tmem) umma_bf16_ts(tmem_base, tmem_base + 256u + (uint32_t)k * 8u, db, idesc, k > 0 ? 1u : 0u);
      else umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  // epilogue: warp w reads TMEM lanes [32w, 32w+32)
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (a_tmem) tmem_dealloc<512>(tmem_base);
    else tmem_dealloc<256>(tmem_base);
  }
}

int umma_selftest(const void* a, const void* b, int N, int K, float* d, int variant, cudaStream_t st) {
  SNB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "selftest: N=%d must be a multiple of 16 in [16,256]", N);
  SNB_REQUIRE(K >= 16 && K <= 512 && K % 16 == 0, "selftest: K=%d must be a multiple of 16 in [16,512]", K);
  size_t smem = (size_t)(128 + N) * K * 2 + 1024;
  SNB_REQUIRE(smem <= 227 * 1024, "selftest: tile does not fit shared memory");
  SNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  __nv_bfloat16* packed = nullptr;
  if (variant & 2) {
    SNB_CHECK_CUDA(cudaMallocAsync((void**)&packed, (size_t)N * K * 2, st));
    k_pack_canonical<<<64, 256, 0, st>>>((const __nv_bfloat16*)b, N, K, 128, (uint32_t)(K / 8) * 128, packed);
    SNB_CHECK_LAUNCH("k_pack_canonical");
  }
  k_umma_selftest<<<1, 128, smem, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, packed, N, K, d, variant);
  SNB_CHECK_LAUNCH("k_umma_selftest");
  if (packed) SNB_CHECK_CUDA(cudaFreeAsync(packed, st));
  return SNB_OK;
}

}  // namespace snb

// ------------------------------------------------------------------------------------------------------------
// snb_umma_microbench: issue rate of tcgen05.mma (M = 128, K = 16, bf16) for a given N with the A operand in
// shared memory (SS) or tensor memory (TS), optionally under the two kinds of traffic the fused kernels add:
// a stream of 8 KB bulk copies landing in shared memory (flags & 1) and epilogue-style tcgen05.ld readers
// (flags & 2).  Operand contents are irrelevant (uninitialised shared memory); only clocks are reported.
// ------------------------------------------------------------------------------------------------------------
namespace snb {
using namespace ptx;

__global__ void __launch_bounds__(192) k_umma_bench(int N, int ts, int flags, int reps, const uint8_t* __restrict__ src,
                                                    unsigned long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_full[4], bar_dummy, bar_set;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sA = smem;                       // 128 x 64 bf16 canonical (16 KB)
  uint8_t* sB = smem + 16384;               // 256 x 64 bf16 canonical (32 KB)
  uint8_t* sL = smem + 49152;               // landing zone 4 x (8 KB << ((flags >> 4) & 3))
  const uint32_t cbytes = 8192u << ((flags >> 4) & 3);
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_full[i], 1);
    mbar_init(&bar_dummy, 1);
    mbar_init(&bar_set, 1);
    fence_mbar_init();
    mbar_arrive(&bar_set);          // phase 0 of bar_set is complete from now on
    done = 0;
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N);
      const unsigned long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        // flags: 0x200 tcgen05.commit after every 4th instruction (the ring-slot release of the fused kernels),
        // 0x400 A operand columns at a stride of 16 (the in-place packed layout), 0x1000 a wait on an already
        // completed mbarrier before every group of 4, 0x2000 B operand cycling through four 32 KB slots (needs copy size 2),
        // 0x4000 the group's first instruction overwrites D (acc = 0)
        const uint32_t t = (uint32_t)(i & 3);
        if ((flags & 0x1000) && t == 0) mbar_wait(&bar_set, 0);
        const uint32_t bsrc = (flags & 0x2000) ? smem_u32(sL) + (uint32_t)((i >> 2) & 3) * 32768u : smem_u32(sB);
        const uint64_t db = umma_smem_desc(bsrc + 2u * t * 128u, 128u, 64u * 16u);
        const uint32_t d = tmem_base + (uint32_t)((i >> 2) % (256 / N)) * (uint32_t)N;
        const uint32_t acc = ((flags & 0x4000) && (i & 15) == 0) ? 0u : 1u;
        if (ts) umma_bf16_ts(d, tmem_base + 256u + (uint32_t)((i >> 2) & 3) * ((flags & 0x400) ? 64u : 0u) + t * ((flags & 0x400) ? 16u : 8u), db, idesc, acc);
        else umma_bf16(d, umma_smem_desc(smem_u32(sA) + 2u * t * 128u, 128u, 64u * 16u), db, idesc, acc);
        if ((flags & 0x200) && t == 3) umma_commit(&bar_dummy);
      }
      umma_commit(&bar_mma);
      mbar_wait(&bar_mma, 0);
      const unsigned long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      done = 1;
    }
  } else if (warp == 1) {
    if (lane == 0 && (flags & 1)) {
      // keep four copies in flight (one per landing slot), like the weight ring of the fused kernels
      uint32_t n = 0;
      for (uint32_t s2 = 0; s2 < 4; ++s2) {
        mbar_arrive_expect_tx(&bar_full[s2], cbytes);
        bulk_g2s(sL + s2 * cbytes, src + (size_t)s2 * cbytes, cbytes, &bar_full[s2]);
      }
      while (!done) {
        const uint32_t s2 = n & 3, ph = (n >> 2) & 1;
        mbar_wait(&bar_full[s2], ph);
        ++n;
        mbar_arrive_expect_tx(&bar_full[s2], cbytes);
        bulk_g2s(sL + s2 * cbytes, src + (size_t)((n + 3 + 7 * blockIdx.x) & 127) * cbytes, cbytes, &bar_full[s2]);
      }
      for (uint32_t k = 0; k < 4; ++k) {            // drain what is still in flight before the CTA exits
        const uint32_t s2 = (n + k) & 3, ph = ((n + k) >> 2) & 1;
        mbar_wait(&bar_full[s2], ph);
      }
      if (blockIdx.x == 0) out[1] = n;
    }
  } else if (flags & 2) {
    unsigned long long n = 0;
    uint32_t acc = 0;
    while (!done) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)((warp - 2) * 32) << 16) + (uint32_t)((n & 7) * 32), v);
      tmem_ld_wait();
      acc ^= v[0];
      ++n;
    }
    if (lane == 0 && blockIdx.x == 0) out[2 + (warp - 2)] = n + (acc == 0x12345u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// CTA-pair variant (flags & 4): a 2-CTA cluster, the leader issues M = 256 cta_group::2 instructions (each CTA supplies
// its 128 rows of A and half of B), multicast commit.  out[0] = clocks on the leader.
__global__ void __launch_bounds__(192) k_umma_bench_pair(int N, int ts, int reps, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* sA = smem;                       // 128 x 64 bf16 canonical (16 KB)
  uint8_t* sB = smem + 16384;               // this CTA's N/2 x 64 half of B (<= 16 KB)
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair<512>(&tmem_base_s);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0 && lane == 0) {
    unsigned long long t0 = 0;
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, N);
      t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t t = (uint32_t)(i & 3);
        const uint64_t db = umma_smem_desc(smem_u32(sB) + 2u * t * 128u, 128u, 64u * 16u);
        if (ts) umma_bf16_ts_pair(tmem_base, tmem_base + 256u + t * 16u, db, idesc, 1u);
        else umma_bf16_pair(tmem_base, umma_smem_desc(smem_u32(sA) + 2u * t * 128u, 128u, 64u * 16u), db, idesc, 1u);
      }
      umma_commit_pair(&bar_mma, 3);
    }
    mbar_wait(&bar_mma, 0);
    if (rank == 0) out[0] = clock64() - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem_base);
}

int umma_microbench(int N, int ts, int flags, int reps, unsigned long long* host_out6, cudaStream_t st) {
  SNB_REQUIRE(N == 64 || N == 128 || N == 256, "microbench: N must be 64, 128 or 256");
  SNB_REQUIRE(reps > 0 && reps <= (1 << 20), "microbench: bad reps");
  const size_t smem = 49152 + 4 * (8192u << ((flags >> 4) & 3)) + 1024;
  SNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  uint8_t* src = nullptr;
  unsigned long long* out = nullptr;
  SNB_CHECK_CUDA(cudaMalloc((void**)&src, (size_t)1024 * 8192));      // 8 MB: 128 slots of up to 64 KB
  SNB_CHECK_CUDA(cudaMalloc((void**)&out, 6 * sizeof(unsigned long long)));
  SNB_CHECK_CUDA(cudaMemsetAsync(out, 0, 6 * sizeof(unsigned long long), st));
  if (flags & 4) {
    SNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_bench_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SNB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_umma_bench_pair, N, ts, reps, out));
  } else {
    // flags & 0x100: one CTA on every SM (full-chip L2 load), CTA 0 reports
    int grid = 1;
    if (flags & 0x100) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&grid, cudaDevAttrMultiProcessorCount, dev); }
    k_umma_bench<<<grid, 192, smem, st>>>(N, ts, flags, reps, src, out);
  }
  SNB_CHECK_LAUNCH("k_umma_bench");
  SNB_CHECK_CUDA(cudaMemcpyAsync(host_out6, out, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  SNB_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFree(src);
  cudaFree(out);
  return SNB_OK;
}

}  // namespace snb

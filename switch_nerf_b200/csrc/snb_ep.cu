// Expert-parallel dispatch / combine over peer memory (see snb_ep.cuh for the protocol).
//
// Per chunk and per rank, in stream order:
//   k_ep_dispatch  (source)   every sample's 48-byte record -> rx[owner of its expert][this rank][e_local][loc]
//                             (P2P stores; dropped samples stay local), last block publishes the per-expert kept
//                             counts to the owners and raises flag A on every peer (st.release.sys)
//   k_ep_plan      (owner)    waits for flag A of every peer (ld.acquire.sys), builds the tile plan over the
//                             received records (one bucket per local expert, sources concatenated; + the local
//                             dropped bucket)
//   k_back         (owner)    launch #2 of snb_tc.cu reads records, stores {rgb, sigma} to ret[source rank][sample]
//                             (P2P), its last CTA raises flag B on every peer
//   k_ep_wait + D2D copy (source)  waits for flag B of every peer, copies ret -> out
// Buffer reuse is safe without further handshakes: a peer can only write set k for chunk c+NS after it has seen
// this rank's flag B of chunk c (its own stream order), i.e. after this rank finished reading set k.
#include "snb_ep.cuh"

namespace snb {

namespace {

constexpr int MAX_LOCAL_EXPERTS_PLUS = 18;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// bounded spin: a peer that never arrives (crashed rank, mismatched call sequence) traps after 20 s instead of
// hanging the GPU
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t epoch) {
  const unsigned long long t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
    __nanosleep(200);
    if (globaltimer_ns() - t0 > 20000000000ull) {
      printf("snb expert-parallel: timeout waiting for a peer flag (epoch %u, have %u)\n", epoch, ld_acquire_sys(p));
      __trap();
    }
  }
}

struct EpDev {
  int rank, world, E, E_local, capmax;
  char* peer[EP_MAX_WORLD];            // base of every rank's region of THIS set
  size_t o_rx, o_cnt, o_flag, o_local;
};

// local scratch ints of a set: [0] drop counter, [1] dispatch block counter, [2] launch #2 CTA counter
__global__ void __launch_bounds__(256) k_ep_dispatch(EpDev d, const float* __restrict__ x, int x_cols,
                                                     const float* __restrict__ gate, const float* __restrict__ noise,
                                                     const int* __restrict__ idx, const int* __restrict__ loc,
                                                     const int* __restrict__ counts, const int* __restrict__ cap_dev,
                                                     int64_t S, uint32_t epoch) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cap = *cap_dev;
  int* local = reinterpret_cast<int*>(d.peer[d.rank] + d.o_local);
  if (s < S) {
    const float* xr = x + s * x_cols;
    const int e = idx[s], l = loc[s];
    float4 r0 = make_float4(xr[0], xr[1], xr[2], xr[3]);
    float4 r1 = make_float4(xr[4], xr[5], xr[6], gate[s]);
    float4 r2 = make_float4(noise ? noise[s] : 0.f, __int_as_float((int)s), __int_as_float(d.rank), 0.f);
    char* base;
    int64_t slot;
    if (l < cap) {
      const int owner = e / d.E_local, el = e - owner * d.E_local;
      base = d.peer[owner];
      slot = ((int64_t)d.rank * d.E_local + el) * d.capmax + l;
    } else {
      base = d.peer[d.rank];
      slot = (int64_t)d.world * d.E_local * d.capmax + atomicAdd(&local[0], 1);
    }
    float4* p = reinterpret_cast<float4*>(base + d.o_rx) + slot * 3;
    p[0] = r0; p[1] = r1; p[2] = r2;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(&local[1], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last) {            // every block's records are ordered before this block's flag stores (fence + atomic chain)
    if ((int)threadIdx.x < d.E) {
      const int e = threadIdx.x, owner = e / d.E_local, el = e - owner * d.E_local;
      reinterpret_cast<int*>(d.peer[owner] + d.o_cnt)[d.rank * d.E_local + el] = min(counts[e], cap);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < d.world)
      st_release_sys(reinterpret_cast<uint32_t*>(d.peer[threadIdx.x] + d.o_flag) + d.rank, epoch);
    if (threadIdx.x == 0) local[1] = 0;
  }
}

// Several small CTAs (256 threads, 32 registers: they fit next to a resident launch-#2 CTA, so the plan never has to wait
// for a free SM): every CTA waits for the peers' flag A and derives the same segment table; the rows are filled with a
// grid-stride loop; the last CTA to finish resets the per-set scratch.  local[3] = CTA ticket.
constexpr int EP_PLAN_CTAS = 32;
__global__ void __launch_bounds__(256) k_ep_plan(EpDev d, TileTable tt, int pair, uint32_t epoch) {
  __shared__ int s_row[MAX_LOCAL_EXPERTS_PLUS], s_tile[MAX_LOCAL_EXPERTS_PLUS], s_kc[MAX_LOCAL_EXPERTS_PLUS],
      s_nt[MAX_LOCAL_EXPERTS_PLUS];
  __shared__ int s_last;
  char* mine = d.peer[d.rank];
  if ((int)threadIdx.x < d.world) wait_flag(reinterpret_cast<const uint32_t*>(mine + d.o_flag) + threadIdx.x, epoch);
  __syncthreads();
  const int* cnt = reinterpret_cast<const int*>(mine + d.o_cnt);     // [world][E_local] kept rows per source
  int* local = reinterpret_cast<int*>(mine + d.o_local);
  const int EL = d.E_local;
  if (threadIdx.x == 0) {
    int row = 0, nt = 0;
    for (int el = 0; el <= EL; ++el) {
      int kc = 0;
      if (el < EL) for (int w = 0; w < d.world; ++w) kc += cnt[w * EL + el];
      else kc = *reinterpret_cast<volatile int*>(&local[0]);         // dropped bucket (local samples only)
      int n = (kc + EP_TILE - 1) / EP_TILE;
      if (pair) n = (n + 1) & ~1;
      s_row[el] = row; s_tile[el] = nt; s_kc[el] = kc; s_nt[el] = n;
      if (blockIdx.x == 0) tt.seg_start[el] = row;
      row += (kc + EP_TILE - 1) / EP_TILE * EP_TILE;
      nt += n;
    }
    if (blockIdx.x == 0) { *tt.n_tiles = nt; *tt.drop_counter = 0; }
  }
  __syncthreads();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  for (int el = 0; el <= EL; ++el) {
    const int t0 = s_tile[el], nt = s_nt[el];
    for (int i = gtid; i < nt; i += gsz) {
      tt.tile_expert[t0 + i] = (el < EL) ? d.rank * EL + el : -1;
      tt.tile_row0[t0 + i] = s_row[el] + i * EP_TILE;
      tt.tile_rows[t0 + i] = max(0, min(EP_TILE, s_kc[el] - i * EP_TILE));
    }
    if (el < EL) {
      int off = s_row[el];
      for (int w = 0; w < d.world; ++w) {
        const int kcw = cnt[w * EL + el];
        const int slot0 = (w * EL + el) * d.capmax;
        for (int l = gtid; l < kcw; l += gsz) tt.row2sample[off + l] = slot0 + l;
        off += kcw;
      }
    } else {
      const int slot0 = d.world * EL * d.capmax;
      for (int j = gtid; j < s_kc[el]; j += gsz) tt.row2sample[s_row[el] + j] = slot0 + j;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&local[3], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) { local[0] = 0; local[3] = 0; }     // drop counter / ticket of the next chunk that uses this set
}

__global__ void k_ep_wait(EpDev d, uint32_t epoch) {
  if ((int)threadIdx.x < d.world)
    wait_flag(reinterpret_cast<const uint32_t*>(d.peer[d.rank] + d.o_flag) + EP_MAX_WORLD + threadIdx.x, epoch);
}

EpDev dev_view(const Ep* ep, int set) {
  EpDev d;
  d.rank = ep->rank; d.world = ep->world; d.E = ep->E; d.E_local = ep->E_local; d.capmax = ep->capmax;
  for (int w = 0; w < EP_MAX_WORLD; ++w)
    d.peer[w] = (w < ep->world && ep->peer_base[w]) ? ep->peer_base[w] + (size_t)set * ep->set_stride : nullptr;
  d.o_rx = ep->o_rx; d.o_cnt = ep->o_cnt; d.o_flag = ep->o_flag; d.o_local = ep->o_local;
  return d;
}

}  // namespace

int ep_create(int rank, int world, int num_experts, int64_t max_chunk_rows, double max_cf, Ep** out) {
  SNB_REQUIRE(out, "snb_a2a_init: NULL out");
  SNB_REQUIRE(world >= 1 && world <= EP_MAX_WORLD, "snb_a2a_init: world size %d not in [1, %d]", world, EP_MAX_WORLD);
  SNB_REQUIRE(rank >= 0 && rank < world, "snb_a2a_init: rank %d outside world %d", rank, world);
  SNB_REQUIRE(num_experts >= world && num_experts % world == 0 && num_experts <= 16,
              "snb_a2a_init: %d experts do not shard evenly over %d ranks", num_experts, world);
  SNB_REQUIRE(max_chunk_rows >= 1 && max_chunk_rows < (1ll << 27), "snb_a2a_init: bad max_chunk_rows");
  SNB_REQUIRE(max_cf > 0, "snb_a2a_init: capacity factor must be > 0");
  Ep* ep = new Ep();
  ep->rank = rank; ep->world = world; ep->E = num_experts; ep->E_local = num_experts / world;
  ep->smax = max_chunk_rows;
  ep->capmax = capacity_of(max_chunk_rows, num_experts, max_cf);
  if (ep->capmax < 1) ep->capmax = 1;
  cudaGetDevice(&ep->device);
  size_t o = 0;
  ep->o_rx = o;    o += align_up((size_t)ep->rx_rows() * EP_REC_FLOATS * 4, 256);
  ep->o_cnt = o;   o += align_up((size_t)world * ep->E_local * 4, 256);
  ep->o_ret = o;   o += align_up((size_t)ep->smax * 16, 256);
  ep->o_flag = o;  o += align_up(2 * EP_MAX_WORLD * 4, 256);
  ep->o_local = o; o += 256;
  ep->set_stride = align_up(o, 4096);
  ep->bytes = ep->set_stride * EP_SETS;
  cudaError_t e = cudaMalloc((void**)&ep->base, ep->bytes);
  if (e != cudaSuccess) {
    set_error("snb_a2a_init: cudaMalloc(%zu) failed: %s", ep->bytes, cudaGetErrorString(e));
    delete ep;
    return SNB_ECUDA;
  }
  e = cudaMemset(ep->base, 0, ep->bytes);      // flags, counters (synchronous: done before the handle leaves)
  if (e != cudaSuccess) {
    set_error("snb_a2a_init: cudaMemset failed: %s", cudaGetErrorString(e));
    cudaFree(ep->base);
    delete ep;
    return SNB_ECUDA;
  }
  ep->peer_base[rank] = ep->base;
  ep->connected = (world == 1);
  *out = ep;
  return SNB_OK;
}

int ep_export(Ep* ep, void* handle64) {
  SNB_REQUIRE(ep && handle64, "snb_a2a_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  SNB_CHECK_CUDA(cudaIpcGetMemHandle(&h, ep->base));
  memcpy(handle64, &h, sizeof(h));
  return SNB_OK;
}

int ep_connect_ipc(Ep* ep, const void* handles) {
  SNB_REQUIRE(ep && handles, "snb_a2a_connect: NULL argument");
  for (int w = 0; w < ep->world; ++w) {
    if (w == ep->rank || ep->peer_base[w]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)w * sizeof(h), sizeof(h));
    void* p = nullptr;
    SNB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ep->peer_base[w] = (char*)p;
    ep->ipc_opened[w] = true;
  }
  ep->connected = true;
  return SNB_OK;
}

int ep_connect_ptrs(Ep* ep, void* const* bases) {
  SNB_REQUIRE(ep && bases, "snb_a2a_connect_ptrs: NULL argument");
  for (int w = 0; w < ep->world; ++w) {
    if (w == ep->rank) continue;
    SNB_REQUIRE(bases[w], "snb_a2a_connect_ptrs: NULL base for rank %d", w);
    ep->peer_base[w] = (char*)bases[w];
  }
  ep->connected = true;
  return SNB_OK;
}

int ep_disconnect(Ep* ep) {
  if (!ep) return SNB_OK;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaSetDevice(ep->device);
  cudaDeviceSynchronize();
  for (int w = 0; w < ep->world; ++w) {
    if (ep->ipc_opened[w]) cudaIpcCloseMemHandle(ep->peer_base[w]);
    ep->ipc_opened[w] = false;
    if (w != ep->rank) ep->peer_base[w] = nullptr;
  }
  ep->connected = (ep->world == 1);
  cudaSetDevice(dev);
  return SNB_OK;
}

int ep_destroy(Ep* ep) {
  if (!ep) return SNB_OK;
  ep_disconnect(ep);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaSetDevice(ep->device);
  cudaFree(ep->base);
  cudaSetDevice(dev);
  delete ep;
  return SNB_OK;
}

int64_t ep_max_rows(const Ep* ep, int64_t S) {
  return (int64_t)ep->world * ep->E_local * ep->capmax + S + (int64_t)EP_TILE * (ep->E_local + 2);
}
int64_t ep_max_tiles(const Ep* ep, int64_t S) {
  return cdiv((int64_t)ep->world * ep->E_local * ep->capmax, EP_TILE) + cdiv(S, EP_TILE) + 2 * (ep->E_local + 2);
}

int ep_dispatch_plan(Ep* ep, int set, const float* x, int x_cols, const float* gate, const float* noise,
                     const int* idx, const int* loc, const int* counts, const int* cap_dev, int64_t S, int cap_host,
                     int pair, TileTable tt, cudaStream_t st) {
  SNB_REQUIRE(ep->connected, "expert-parallel group is not connected (snb_a2a_connect)");
  SNB_REQUIRE(set >= 0 && set < EP_SETS, "expert-parallel: bad buffer set %d", set);
  SNB_REQUIRE(S >= 1 && S <= ep->smax, "expert-parallel: chunk of %lld rows exceeds the %lld the group was sized for",
              (long long)S, (long long)ep->smax);
  SNB_REQUIRE(cap_host <= ep->capmax, "expert-parallel: capacity %d exceeds the %d the group was sized for", cap_host,
              ep->capmax);
  SNB_REQUIRE(x_cols == 7, "expert-parallel: records carry [xyz, dir, image index] rows only");
  const uint32_t epoch = ++ep->epoch[set];
  const EpDev d = dev_view(ep, set);
  k_ep_dispatch<<<(unsigned)cdiv(S, 256), 256, 0, st>>>(d, x, x_cols, gate, noise, idx, loc, counts, cap_dev, S, epoch);
  SNB_CHECK_LAUNCH("k_ep_dispatch");
  k_ep_plan<<<EP_PLAN_CTAS, 256, 0, st>>>(d, tt, pair, epoch);
  SNB_CHECK_LAUNCH("k_ep_plan");
  return SNB_OK;
}

void ep_row_io(const Ep* ep, int set, RowIO* io) {
  char* mine = ep->peer_base[ep->rank] + (size_t)set * ep->set_stride;
  const float* rec = reinterpret_cast<const float*>(mine + ep->o_rx);
  io->x = rec;        io->x_stride = EP_REC_FLOATS;
  io->gate = rec + 7; io->g_stride = EP_REC_FLOATS;
  io->noise = rec + 8; io->n_stride = EP_REC_FLOATS;
  io->out = nullptr;
  io->ep = 1; io->world = ep->world; io->rank = ep->rank;
  for (int w = 0; w < EP_MAX_WORLD; ++w) {
    char* pw = (w < ep->world) ? ep->peer_base[w] + (size_t)set * ep->set_stride : nullptr;
    io->ret[w] = pw ? reinterpret_cast<float*>(pw + ep->o_ret) : nullptr;
    io->flag_b[w] = pw ? reinterpret_cast<uint32_t*>(pw + ep->o_flag) + EP_MAX_WORLD + ep->rank : nullptr;
  }
  io->done = reinterpret_cast<int*>(mine + ep->o_local) + 2;
  io->epoch = ep->epoch[set];
}

int ep_finish(Ep* ep, int set, float* out, int64_t S, cudaStream_t st) {
  const EpDev d = dev_view(ep, set);
  k_ep_wait<<<1, 32, 0, st>>>(d, ep->epoch[set]);
  SNB_CHECK_LAUNCH("k_ep_wait");
  char* mine = ep->peer_base[ep->rank] + (size_t)set * ep->set_stride;
  SNB_CHECK_CUDA(cudaMemcpyAsync(out, mine + ep->o_ret, (size_t)S * 16, cudaMemcpyDeviceToDevice, st));
  return SNB_OK;
}

}  // namespace snb

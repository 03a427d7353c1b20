// Packed routing word shared by launch #1 (k_front_ts writes it) and k_select (which hands the gate value to launch #2):
//   w = (expert id << 26) | key,   key = bits(1.0f) - bits(max gate)  (26 bits for E <= 16: max gate >= ~1/E)
// Ascending key == descending gate, so the batch-prioritised order (tutel_fast_dispatch.py:136-139, 186-188) is the
// ascending order of (key, sample index).
#pragma once
#include "snb_common.cuh"
#include "snb_ep.cuh"

namespace snb {

constexpr int SEL_MAX_E = 16;
constexpr int SEL_KEY_BITS = 26;
constexpr uint32_t SEL_KEY_MASK = (1u << SEL_KEY_BITS) - 1u;
constexpr int SEL_L1_SHIFT = 19;                 // level-0 digit = key bits [19, 26): 128 bins per expert
constexpr int SEL_L2_SHIFT = 9;                  // level-1 digit = key bits [9, 19), level-2 digit = key bits [0, 9)
constexpr int SEL_HBINS = 1 << (SEL_KEY_BITS - SEL_L1_SHIFT);
constexpr int SEL_PM_STRIDE = 16;                // floats per partial-column-sum record
// region zeroed by the host (one memset) before launch #1 of a chunk, in ints:
constexpr int SEL_Z_TICKET = SEL_MAX_E * SEL_HBINS;          // [0, SEL_Z_TICKET) = level-0 histogram; then the l_aux ticket,
constexpr int SEL_Z_BAR = SEL_Z_TICKET + 16;                 // the grid-barrier counters of k_select,
constexpr int SEL_Z_CNT = SEL_Z_BAR + 16;                    // the per-(CTA, key) totals [16][32],
constexpr int SEL_Z_LVH = SEL_Z_CNT + 16 * 32;               // the level histograms [6][SEL_MAX_E][1024]
constexpr int SEL_LVH_INTS = 6 * SEL_MAX_E * 1024;
constexpr int SEL_ZERO_INTS = SEL_Z_LVH + SEL_LVH_INTS;
// records a chunk can produce: one per 32 rows (launch #1) or one per 2048 samples (k_pack_top1)
__host__ __device__ inline int64_t SEL_PM_RECORDS(int64_t S) { return 4 * ((S + 127) / 128) + 4; }

__host__ __device__ __forceinline__ uint32_t sel_key_bits(uint32_t gate_bits) {
  const uint32_t k = (gate_bits <= 0x3F800000u) ? (0x3F800000u - gate_bits) : 0u;
  return k > SEL_KEY_MASK ? SEL_KEY_MASK : k;
}
__device__ __forceinline__ uint32_t sel_key(float g) { return sel_key_bits(__float_as_uint(g)); }
__device__ __forceinline__ uint32_t sel_pack(int e, uint32_t key) { return ((uint32_t)e << SEL_KEY_BITS) | key; }
__device__ __forceinline__ float sel_gate(uint32_t w) { return __uint_as_float(0x3F800000u - (w & SEL_KEY_MASK)); }

struct SelectArgs {
  const uint32_t* w;        // [S] packed words (16-byte aligned)
  int* zero;                // the zeroed region (SEL_Z_*): level-0 key histogram per expert accumulated by launch #1 first
  int self_clean;           // 1: the last CTA zeroes what this chunk dirtied (the region is zero again when the kernel ends)
  const float* pm;          // [npm][SEL_PM_STRIDE] partial column sums of the gates (load-balance loss)
  int npm;
  double* lpart;            // [16][SEL_PM_STRIDE] per-CTA partial column sums (scratch)
  int64_t S;
  int E;
  double cf;
  int bpr, no_batch;
  TileTable tt;             // local mode: row2sample + tile table of launch #2 (all pointers nullable together)
  int pair;
  int* idx;                 // per-sample taps (expert-parallel dispatch, debug): nullable
  int* loc;                 //   kept: row inside the expert bucket (index order); dropped: capacity + drop slot
  float* gate;
  int* counts;              // [E]
  int* cap_dev;             // [1]
  float* l_aux;             // [1]
  int* moe_idx;             // [S] nullable
  unsigned long long* tl;   // debug (SNB_TIMELINE): (tag << 48 | clock) marks of CTA 0, nullable
};

int route_select_launch(const SelectArgs& a, cudaStream_t st);
int route_pack_top1(const float* gates, int64_t S, int32_t E, uint32_t* w, int* hist0, float* pm, int* npm, cudaStream_t st);
size_t route_select_workspace_bytes(int64_t S);
int route_select_from_gates(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int32_t no_batch,
                            int32_t* idx, int32_t* loc, float* gate, int32_t* counts, int32_t* capacity, float* l_aux,
                            void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace snb

// Expert-parallel exchange (BASELINE.json configs[2]): experts sharded E/W per GPU, one process per
// GPU, symmetric device buffers mapped into every peer (CUDA IPC or same-process peer access), data moved by
// plain st.global over NVLink with per-peer release/acquire flags -- no NCCL on the data path.
//
// What is exchanged.  The reference (tutel_moe_layer_nobatch.py:157-218, C1/C2 of SURVEY §2.3) ships the
// dispatched activations: [W, E_local, cap, 256] bf16 = 512 B per sample each way.  Every weight outside the
// experts is replicated on all ranks, and launch #2 already recomputes h = xyz-Linear(PE(xyz)) on the tensor
// core, so this implementation ships the *inputs of launch #2* instead: one 48-byte record per sample
// ({x[7], gate, sigma_noise, source sample, source rank}) to the rank that owns the sample's expert, and one
// 16-byte {rgb, sigma} row back -- 64 B instead of 1024 B per sample, results bit-identical to the
// single-GPU path (row-wise math does not depend on which tile a row sits in).
#pragma once
#include "snb_common.cuh"

namespace snb {

constexpr int EP_MAX_WORLD = 8;
constexpr int EP_REC_FLOATS = 12;     // 48-byte record = 3 x 16-byte stores
constexpr int EP_SETS = 4;            // one per workspace set of the chunk pipeline (tc_forward_chunks)
constexpr int EP_TILE = 128;          // rows per tile of launch #2 (== TILE in snb_tc.cu)

// Tile plan of launch #2 (device arrays in the chunk workspace).
struct TileTable {
  int* n_tiles;       // [1]
  int* tile_expert;   // [max_tiles]  global expert id (-1 = dropped bucket)
  int* tile_row0;     // [max_tiles]
  int* tile_rows;     // [max_tiles]  valid rows in the tile
  int* seg_start;     // [E+1] first row of each expert segment (+ dropped segment)
  int* drop_counter;  // [1]
  int* row2sample;    // [max_rows]   local mode: sample index; expert-parallel mode: record slot
};

// Where launch #2 reads the per-row inputs and writes the per-row result.
//   local mode : x[s*x_stride + c], gate[s], noise[s] (nullable), out[s*4]
//   EP mode    : the same three arrays alias the received records (stride EP_REC_FLOATS); the result goes to
//                ret[source rank] + 4*source sample (a P2P store), and the last CTA raises flag B on every peer.
struct RowIO {
  const float* x;
  int x_stride;
  const float* gate;
  int g_stride;
  const uint32_t* wsel;                // local mode with select routing: gate value decoded from the packed routing word (gate unused)
  const float* noise;
  int n_stride;
  float* out;
  int ep;                              // 0 = local
  int world, rank;
  float* ret[EP_MAX_WORLD];
  uint32_t* flag_b[EP_MAX_WORLD];      // &peer w's flag B slot of this rank
  int* done;                           // CTA completion counter (local, self-resetting)
  uint32_t epoch;
};

struct Ep {
  int rank = 0, world = 1, E = 0, E_local = 0, capmax = 0, device = 0;
  int64_t smax = 0;
  char* base = nullptr;                // this rank's region (cudaMalloc)
  size_t bytes = 0, set_stride = 0;
  size_t o_rx = 0, o_cnt = 0, o_ret = 0, o_flag = 0, o_local = 0;   // offsets inside one set
  char* peer_base[EP_MAX_WORLD] = {};
  bool ipc_opened[EP_MAX_WORLD] = {};
  bool connected = false;
  uint32_t epoch[EP_SETS] = {};
  int64_t rx_rows() const { return (int64_t)world * E_local * capmax + smax; }
};

int ep_create(int rank, int world, int num_experts, int64_t max_chunk_rows, double max_cf, Ep** out);
int ep_export(Ep* ep, void* handle64);
int ep_connect_ipc(Ep* ep, const void* handles);
int ep_connect_ptrs(Ep* ep, void* const* bases);
int ep_disconnect(Ep* ep);   // unmap the peers (every rank, before any rank frees its region)
int ep_destroy(Ep* ep);

// rows / tiles the tile plan of one chunk can need on this rank
int64_t ep_max_rows(const Ep* ep, int64_t S);
int64_t ep_max_tiles(const Ep* ep, int64_t S);

// Source side + plan: scatter the records of this rank's S samples to the owners of their experts, raise flag A on
// every peer; then wait for every peer's flag A and build the tile plan over the received records.
int ep_dispatch_plan(Ep* ep, int set, const float* x, int x_cols, const float* gate, const float* noise,
                     const int* idx, const int* loc, const int* counts, const int* cap_dev, int64_t S, int cap_host,
                     int pair, TileTable tt, cudaStream_t st);
// Fill the RowIO of launch #2 for this set (after ep_dispatch_plan of the same chunk).
void ep_row_io(const Ep* ep, int set, RowIO* io);
// Wait for flag B of every peer (all results of this rank's samples have landed) and copy them to out[S][4].
int ep_finish(Ep* ep, int set, float* out, int64_t S, cudaStream_t st);

}  // namespace snb

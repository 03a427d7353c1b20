// C-ABI entry points (include/switch_nerf_b200.h): argument checking, model object, workspace
// carving and the host-side orchestration of render_rays.  No host synchronisation anywhere.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "snb_common.cuh"
#include "snb_ep.cuh"
#include "snb_select.cuh"

namespace snb {

static thread_local char g_err[1024] = "";

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// process-wide measurement state (snb_profile_*): guarded, so that two host threads rendering with different models
// do not corrupt the event pool (the header's single-stream-per-model rule covers everything model-owned)
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<PhaseEvents> g_prof_pool;
static size_t g_prof_used = 0;
PhaseEvents* profile_next() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on) return nullptr;
  if (g_prof_pool.capacity() < 65536) g_prof_pool.reserve(65536);   // callers keep pointers: never reallocate
  if (g_prof_used >= g_prof_pool.size()) {
    if (g_prof_pool.size() >= 65536) return nullptr;
    PhaseEvents pe;
    for (int i = 0; i < 6; ++i)
      if (cudaEventCreate(&pe.e[i]) != cudaSuccess) return nullptr;
    g_prof_pool.push_back(pe);
  }
  return &g_prof_pool[g_prof_used++];
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// implemented in snb_fp32.cu / snb_render.cu / snb_route.cu
int snb_dispatch_impl(const float* x, const int* idx, const int* loc, const int* begin, const int* cap_dev,
                      int cap_host, int64_t S, int H, int64_t rows_out, float* out, bool zero, cudaStream_t st);
int snb_combine_impl(const float* buf, const int* idx, const int* loc, const int* begin, const float* gate,
                     const int* cap_dev, int cap_host, int64_t S, int H, int64_t rows_buf, float* y, bool relu,
                     cudaStream_t st);
int route_top1_generic(const float* gates, int64_t S, int32_t E, double cf, int32_t bpr, int32_t* idx,
                       int32_t* loc, float* gate, int32_t* counts, int32_t* capacity, float* l_aux, void* ws,
                       size_t ws_bytes, cudaStream_t st);
int composite_launch(const float* z, const float* raw, const float* last_delta, int64_t N, int S, int white_bkgd,
                     float* rgb, float* depth, float* var, float* lam, float* weights, cudaStream_t st);
int sample_pdf_launch(const float* bins, int ld_bins, const float* weights, int ld_w, int w_off, const float* u,
                      int64_t N, int nb, int nf, uint64_t seed, int det, float* zf, cudaStream_t st);
int coarse_z_launch(const float* rays, int64_t N, int Sc, float perturb, uint64_t seed, float* z, cudaStream_t st);
int get_rays_launch(int W, int H, float fx, float fy, float cx, float cy, int center_pixels, const float* c2w, float near,
                    float far, const float* alt, float* rays, cudaStream_t st);
int fill_x_launch(const float* rays, const int* image_indices, const float* z, int64_t N, int Sn, float* x,
                  cudaStream_t st);
int zmid_launch(const float* z, int64_t N, int S, float* mid, cudaStream_t st);
int last_delta_adj_launch(const float* last_delta, const float* z, int64_t N, int S, float* out, cudaStream_t st);
int merge_composite_launch(const float* zf, const float* zc, const float* raw_f, const float* raw_c,
                           const float* last_delta, int64_t N, int Sf, int Sc, int presorted, int white_bkgd,
                           float* rgb, float* depth, float* var, float* lam, cudaStream_t st);
int umma_selftest(const void* a, const void* b, int N, int K, float* d, int variant, cudaStream_t st);
int umma_microbench(int N, int ts, int flags, int reps, unsigned long long* host_out6, cudaStream_t st);
int mip_fill_x_launch(const float* rays, const float* radii, const int* image_indices, const float* ze, int64_t N,
                      int Se, float* x, cudaStream_t st);
int mip_composite_launch(const float* ze, const float* raw, const float* last_delta, int64_t N, int Se,
                         float rgb_padding, int white_bkgd, float* rgb, float* depth, float* var, float* weights,
                         cudaStream_t st);
int mip_resample_launch(const float* ze, const float* weights, int64_t N, int Se, int nf, float resample_padding,
                        int randomized, uint64_t seed, float* zf, cudaStream_t st);
int tc_timeline_read(unsigned long long* host, int n);

// [E][K][N] -> [E][N][K]
__global__ void k_transpose_expert(const float* __restrict__ in, float* __restrict__ out, int E, int K, int N) {
  __shared__ float t[32][33];
  const int e = blockIdx.z;
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const float* src = in + (size_t)e * K * N;
  float* dst = out + (size_t)e * K * N;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int k = k0 + i, n = n0 + threadIdx.x;
    if (k < K && n < N) t[i][threadIdx.x] = src[(size_t)k * N + n];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int n = n0 + i, k = k0 + threadIdx.x;
    if (k < K && n < N) dst[(size_t)n * K + k] = t[threadIdx.x][i];
  }
}

static int check_desc(const snb_model_desc* d) {
  SNB_REQUIRE(d, "desc is NULL");
  SNB_REQUIRE(d->num_experts >= 1 && d->num_experts <= 64, "num_experts=%d unsupported (1..64)", d->num_experts);
  SNB_REQUIRE(d->width >= 16 && d->width <= 1024 && d->width % 16 == 0, "width=%d unsupported", d->width);
  SNB_REQUIRE(d->expert_layers >= 1 && d->expert_layers <= 16, "expert_layers=%d unsupported", d->expert_layers);
  SNB_REQUIRE(d->skip_layer < d->expert_layers, "skip_layer=%d out of range", d->skip_layer);
  SNB_REQUIRE(d->gate_layers >= 1 && d->gate_layers <= 4, "gate_layers=%d unsupported", d->gate_layers);
  SNB_REQUIRE(d->pos_xyz_freqs >= 0 && d->pos_xyz_freqs <= 16, "pos_xyz_freqs=%d unsupported", d->pos_xyz_freqs);
  SNB_REQUIRE(d->pos_dir_freqs >= 0 && d->pos_dir_freqs <= 16, "pos_dir_freqs=%d unsupported", d->pos_dir_freqs);
  SNB_REQUIRE(d->appearance_dim >= 0 && d->appearance_dim <= 256, "appearance_dim=%d unsupported", d->appearance_dim);
  SNB_REQUIRE(d->appearance_count >= 1, "appearance_count must be >= 1");
  SNB_REQUIRE(d->hidden2 >= 1 && d->hidden2 <= 1024, "hidden2=%d unsupported", d->hidden2);
  return SNB_OK;
}

static int upload(Model* m, const snb_weights* w, cudaStream_t st) {
  const snb_model_desc& d = m->d;
  const int M = d.width, E = d.num_experts;
  auto cp = [&](float* dst, const float* src, size_t n) -> int {
    SNB_REQUIRE(src, "a weight pointer is NULL");
    SNB_CHECK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SNB_OK;
  };
  int rc;
  if ((rc = cp(m->xyz_w, w->xyz_w, (size_t)M * m->xyz_in))) return rc;
  if ((rc = cp(m->xyz_b, w->xyz_b, M))) return rc;
  for (int i = 0; i < d.gate_layers; ++i) {
    if ((rc = cp(m->gate_w[i], w->gate_w[i], (size_t)M * M))) return rc;
    if ((rc = cp(m->gate_b[i], w->gate_b[i], M))) return rc;
  }
  if ((rc = cp(m->ln_w, w->ln_w, M))) return rc;
  if ((rc = cp(m->ln_b, w->ln_b, M))) return rc;
  if ((rc = cp(m->wg, w->wg, (size_t)E * M))) return rc;
  for (int j = 0; j < d.expert_layers; ++j) {
    SNB_REQUIRE(w->expert_w[j] && w->expert_b[j], "expert weight pointer %d is NULL", j);
    dim3 grid((unsigned)cdiv(M, 32), (unsigned)cdiv(M, 32), (unsigned)E), blk(32, 8);
    k_transpose_expert<<<grid, blk, 0, st>>>(w->expert_w[j], m->exp_w[j], E, M, M);
    SNB_CHECK_LAUNCH("k_transpose_expert");
    if ((rc = cp(m->exp_b[j], w->expert_b[j], (size_t)E * M))) return rc;
  }
  if ((rc = cp(m->l1_w, w->l1_w, (size_t)M * M))) return rc;
  if ((rc = cp(m->l1_b, w->l1_b, M))) return rc;
  if ((rc = cp(m->l2_w, w->l2_w, (size_t)d.hidden2 * m->cat_in))) return rc;
  if ((rc = cp(m->l2_b, w->l2_b, d.hidden2))) return rc;
  if ((rc = cp(m->sigma_w, w->sigma_w, M))) return rc;
  if ((rc = cp(m->sigma_b, w->sigma_b, 1))) return rc;
  if ((rc = cp(m->color_w, w->color_w, (size_t)3 * d.hidden2))) return rc;
  if ((rc = cp(m->color_b, w->color_b, 3))) return rc;
  if (d.appearance_dim > 0)
    if ((rc = cp(m->emb_a, w->emb_a, (size_t)d.appearance_count * d.appearance_dim))) return rc;
  if (tc_supported(m)) return tc_pack_weights(m, w, st);
  return SNB_OK;
}

struct BgModel;
int bg_create(const snb_bg_desc* d, const snb_bg_weights* w, cudaStream_t st, BgModel** out);
int bg_update(BgModel* m, const snb_bg_weights* w, cudaStream_t st);
void bg_destroy(BgModel* m);
size_t bg_workspace_bytes(const BgModel* m, int64_t S);
int bg_forward(BgModel* m, const float* x, int64_t S, const float* noise, float* out, Arena& ws, cudaStream_t st);
int intersect_sphere_launch(const float* rays, int64_t N, const float* c, const float* rad, float* fg_far, int* bad, cudaStream_t st);
int depth2pts_outside_launch(const float* rays, const float* c, const float* rad, const float* z, int64_t N, int S, float* pts,
                             float* depth_real, cudaStream_t st);
}  // namespace snb

using namespace snb;

extern "C" {

const char* snb_last_error(void) { return g_err; }
int snb_version(void) { return 100; }
int64_t snb_launch_count(void) { return (int64_t)g_launches.load(); }

int snb_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  g_prof_used = 0;
  return SNB_OK;
}

int snb_debug_timeline(uint64_t* host_out, int32_t n) { return tc_timeline_read((unsigned long long*)host_out, n); }

int snb_profile_collect(double* out4) {
  SNB_REQUIRE(out4, "snb_profile_collect: NULL");
  SNB_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double f = 0, r = 0, b = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    float t;
    PhaseEvents& pe = g_prof_pool[i];
    SNB_CHECK_CUDA(cudaEventElapsedTime(&t, pe.e[0], pe.e[1])); f += t;
    SNB_CHECK_CUDA(cudaEventElapsedTime(&t, pe.e[2], pe.e[3])); r += t;
    SNB_CHECK_CUDA(cudaEventElapsedTime(&t, pe.e[4], pe.e[5])); b += t;
  }
  out4[0] = f; out4[1] = r; out4[2] = b; out4[3] = (double)g_prof_used;
  g_prof_used = 0;
  return SNB_OK;
}

int snb_model_create(const snb_model_desc* desc, const snb_weights* w, void* stream, snb_model_t** out) {
  SNB_REQUIRE(out && w, "snb_model_create: NULL argument");
  int rc = check_desc(desc);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("snb_model_create: no CUDA device -- switch_nerf_b200 has no CPU path");
    return SNB_ECUDA;
  }
  Model* m = new (std::nothrow) Model();
  if (m) tuning_from_env(&m->tune);
  SNB_REQUIRE(m, "out of host memory");
  m->d = *desc;
  m->xyz_in = 3 + 6 * desc->pos_xyz_freqs;
  m->dir_in = 3 + 6 * desc->pos_dir_freqs;
  m->cat_in = desc->width + m->dir_in + desc->appearance_dim;
  m->x_cols = (desc->mip ? 6 : 3) + 3 + 1;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev);
  const int M = desc->width, E = desc->num_experts, L = desc->expert_layers;
  size_t n = 0;
  auto cnt = [&](size_t k) { size_t o = n; n += (k + 63) / 64 * 64; return o; };
  size_t o_xyz_w = cnt((size_t)M * m->xyz_in), o_xyz_b = cnt(M);
  size_t o_gw[4], o_gb[4];
  for (int i = 0; i < desc->gate_layers; ++i) { o_gw[i] = cnt((size_t)M * M); o_gb[i] = cnt(M); }
  size_t o_lnw = cnt(M), o_lnb = cnt(M), o_wg = cnt((size_t)E * M);
  size_t o_ew[16], o_eb[16];
  for (int j = 0; j < L; ++j) { o_ew[j] = cnt((size_t)E * M * M); o_eb[j] = cnt((size_t)E * M); }
  size_t o_l1w = cnt((size_t)M * M), o_l1b = cnt(M), o_l2w = cnt((size_t)desc->hidden2 * m->cat_in), o_l2b = cnt(desc->hidden2);
  size_t o_sw = cnt(M), o_sb = cnt(1), o_cw = cnt((size_t)3 * desc->hidden2), o_cb = cnt(3);
  size_t o_emb = cnt((size_t)desc->appearance_count * (desc->appearance_dim > 0 ? desc->appearance_dim : 1));
  m->f32_bytes = n * sizeof(float);
  cudaError_t e = cudaMalloc(&m->f32_blob, m->f32_bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", m->f32_bytes, cudaGetErrorString(e)); delete m; return SNB_ECUDA; }
  float* b = m->f32_blob;
  m->xyz_w = b + o_xyz_w; m->xyz_b = b + o_xyz_b;
  for (int i = 0; i < desc->gate_layers; ++i) { m->gate_w[i] = b + o_gw[i]; m->gate_b[i] = b + o_gb[i]; }
  m->ln_w = b + o_lnw; m->ln_b = b + o_lnb; m->wg = b + o_wg;
  for (int j = 0; j < L; ++j) { m->exp_w[j] = b + o_ew[j]; m->exp_b[j] = b + o_eb[j]; }
  m->l1_w = b + o_l1w; m->l1_b = b + o_l1b; m->l2_w = b + o_l2w; m->l2_b = b + o_l2b;
  m->sigma_w = b + o_sw; m->sigma_b = b + o_sb; m->color_w = b + o_cw; m->color_b = b + o_cb; m->emb_a = b + o_emb;
  rc = upload(m, w, (cudaStream_t)stream);
  if (rc) { snb_model_destroy((snb_model_t*)m); return rc; }
  *out = (snb_model_t*)m;
  return SNB_OK;
}

int snb_model_update(snb_model_t* mm, const snb_weights* w, void* stream) {
  SNB_REQUIRE(mm && w, "snb_model_update: NULL argument");
  return upload((Model*)mm, w, (cudaStream_t)stream);
}

int snb_model_get_tuning(const snb_model_t* mm, snb_tuning* out) {
  SNB_REQUIRE(mm && out, "snb_model_get_tuning: NULL argument");
  *out = ((const Model*)mm)->tune;
  return SNB_OK;
}

int snb_model_set_tuning(snb_model_t* mm, const snb_tuning* t) {
  Model* m = (Model*)mm;
  SNB_REQUIRE(m && t, "snb_model_set_tuning: NULL argument");
  SNB_REQUIRE((t->cta_group_front == 1 || t->cta_group_front == 2) && (t->cta_group_back == 1 || t->cta_group_back == 2),
              "snb_model_set_tuning: cta_group must be 1 or 2");
  SNB_REQUIRE(t->pipe_depth <= 3 && t->route_sms < m->sm_count / 2, "snb_model_set_tuning: pipe_depth <= 3, route_sms < SMs / 2");
  const bool repack = (t->gather_h != 0) != (m->tune.gather_h != 0) || (t->wide != 0) != (m->tune.wide != 0);
  SNB_REQUIRE(!repack || !m->tc_blob, "snb_model_set_tuning: gather_h / wide change the packed weight images: set them "
                                      "through the environment (SNB_GATHER_H / SNB_WIDE) before the model is created");
  m->tune = *t;
  return SNB_OK;
}

void snb_model_destroy(snb_model_t* mm) {
  Model* m = (Model*)mm;
  if (!m) return;
  if (m->f32_blob) cudaFree(m->f32_blob);
  tc_release(m);
  if (m->sel_zero) { cudaFree(m->sel_zero); m->sel_zero = nullptr; }
  if (m->side_stream) {
    cudaStreamDestroy(m->side_stream);
    for (int i = 0; i < 4; ++i) { cudaEventDestroy(m->ev_front[i]); cudaEventDestroy(m->ev_route[i]); }
    if (m->fin_stream) cudaStreamDestroy(m->fin_stream);
    for (int i = 0; i < 4; ++i) { if (m->ev_back[i]) cudaEventDestroy(m->ev_back[i]); if (m->ev_fin[i]) cudaEventDestroy(m->ev_fin[i]); }
  }
  delete m;
}

// ---- expert-parallel group (snb_ep.cu) ----
int snb_a2a_init(int32_t rank, int32_t world, int32_t num_experts, int64_t max_chunk_rows, double max_cf,
                 snb_a2a_t** out) {
  return ep_create(rank, world, num_experts, max_chunk_rows, max_cf, (Ep**)out);
}
int32_t snb_a2a_handle_bytes(void) { return 64; }
int snb_a2a_export(snb_a2a_t* g, void* handle_out) { return ep_export((Ep*)g, handle_out); }
int snb_a2a_connect(snb_a2a_t* g, const void* handles, size_t handles_bytes) {
  SNB_REQUIRE(g && handles_bytes >= (size_t)((Ep*)g)->world * 64, "snb_a2a_connect: need world x 64 handle bytes");
  return ep_connect_ipc((Ep*)g, handles);
}
int snb_a2a_connect_ptrs(snb_a2a_t* g, void* const* bases, int32_t n) {
  SNB_REQUIRE(g && n >= ((Ep*)g)->world, "snb_a2a_connect_ptrs: need one base per rank");
  return ep_connect_ptrs((Ep*)g, bases);
}
void* snb_a2a_local_base(snb_a2a_t* g) { return g ? ((Ep*)g)->base : nullptr; }
size_t snb_a2a_region_bytes(const snb_a2a_t* g) { return g ? ((const Ep*)g)->bytes : 0; }
int snb_a2a_disconnect(snb_a2a_t* g) { return ep_disconnect((Ep*)g); }
int snb_a2a_finalize(snb_a2a_t* g) { return ep_destroy((Ep*)g); }
int snb_model_attach_a2a(snb_model_t* mm, snb_a2a_t* g) {
  Model* m = (Model*)mm;
  SNB_REQUIRE(m, "snb_model_attach_a2a: NULL model");
  Ep* ep = (Ep*)g;
  if (ep) {
    SNB_REQUIRE(ep->E == m->d.num_experts, "snb_model_attach_a2a: group was built for %d experts, model has %d", ep->E,
                m->d.num_experts);
    SNB_REQUIRE(tc_supported(m) && m->x_cols == 7, "snb_model_attach_a2a: expert-parallel mode needs the tcgen05 path "
                "(width 256, NeRFMoE rows)");
    SNB_REQUIRE(ep->connected, "snb_model_attach_a2a: group is not connected");
  }
  m->ep = ep;
  return SNB_OK;
}

size_t snb_workspace_bytes(const snb_model_t* mm, int64_t max_chunk_samples, double max_cf) {
  const Model* m = (const Model*)mm;
  if (!m) return 0;
  size_t a = fp32_workspace_bytes(m, max_chunk_samples, max_cf);      // >= fp32_moe_layer_workspace_bytes
  size_t b = tc_supported(m) ? tc_workspace_bytes(m, max_chunk_samples, max_cf) : 0;
  return (a > b ? a : b) + 4096;
}

size_t snb_route_workspace_bytes(int64_t S, int32_t E) { return route_workspace_bytes(S, E); }

int snb_route_top1(const float* gates, int64_t S, int32_t E, double capacity_factor, int32_t bpr, int32_t* idx,
                   int32_t* loc, float* gate, int32_t* counts, int32_t* capacity, float* l_aux, void* workspace,
                   size_t workspace_bytes, void* stream) {
  SNB_REQUIRE(capacity_factor > 0, "capacity_factor must be > 0 (dynamic capacity is not part of the hot path)");
  return route_top1_generic(gates, S, E, capacity_factor, bpr, idx, loc, gate, counts, capacity, l_aux, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

size_t snb_route_select_workspace_bytes(int64_t S) { return route_select_workspace_bytes(S); }

int snb_route_select(const float* gates, int64_t S, int32_t E, double capacity_factor, int32_t bpr, int32_t no_batch,
                     int32_t* idx, int32_t* loc, float* gate, int32_t* counts, int32_t* capacity, float* l_aux,
                     void* workspace, size_t workspace_bytes, void* stream) {
  SNB_REQUIRE(capacity_factor > 0, "capacity_factor must be > 0 (dynamic capacity is not part of the hot path)");
  SNB_REQUIRE(gates && workspace, "snb_route_select: NULL pointer");
  return route_select_from_gates(gates, S, E, capacity_factor, bpr, no_batch, idx, loc, gate, counts, capacity, l_aux,
                                 workspace, workspace_bytes, (cudaStream_t)stream);
}

int snb_dispatch_fwd(const float* x, const int32_t* idx, const int32_t* loc, const int32_t* begin, int64_t S,
                     int32_t H, int32_t capacity, int64_t rows_out, float* out, void* stream) {
  SNB_REQUIRE(out && (S == 0 || (x && idx && loc)), "snb_dispatch_fwd: NULL pointer");
  return snb_dispatch_impl(x, idx, loc, begin, nullptr, capacity, S, H, rows_out, out, true, (cudaStream_t)stream);
}

int snb_combine(const float* buf, const int32_t* idx, const int32_t* loc, const int32_t* begin, const float* gate,
                int64_t S, int32_t H, int32_t capacity, int64_t rows_buf, float* y, void* stream) {
  SNB_REQUIRE(S == 0 || (buf && idx && loc && y), "snb_combine: NULL pointer");
  return snb_combine_impl(buf, idx, loc, begin, gate, nullptr, capacity, S, H, rows_buf, y, false, (cudaStream_t)stream);
}

int snb_moe_forward(snb_model_t* mm, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* opts,
                    int32_t precision, float* out, int32_t* moe_idx, float* l_aux, float* dbg_gates,
                    int32_t* dbg_loc, void* workspace, size_t workspace_bytes, void* stream) {
  Model* m = (Model*)mm;
  SNB_REQUIRE(m && opts, "snb_moe_forward: NULL model/opts");
  SNB_REQUIRE(S >= 0, "snb_moe_forward: negative S");
  SNB_REQUIRE(S == 0 || (x && out), "snb_moe_forward: NULL x/out");
  SNB_REQUIRE(opts->capacity_factor > 0, "capacity_factor must be > 0");
  SNB_REQUIRE(S < (1ll << 31) / 16, "snb_moe_forward: S too large for one chunk");
  Arena ws(workspace, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 0) {
    if (l_aux) SNB_CHECK_CUDA(cudaMemsetAsync(l_aux, 0, sizeof(float), st));
    return SNB_OK;
  }
  if (m->ep && precision != SNB_PREC_BF16) {
    set_error("expert-parallel mode runs on the bf16 tcgen05 path only (precision=%d)", precision);
    return SNB_EUNSUPPORTED;
  }
  if (precision == SNB_PREC_FP32)
    return fp32_forward(m, x, S, sigma_noise, opts, out, moe_idx, l_aux, dbg_gates, dbg_loc, ws, st);
  if (precision == SNB_PREC_BF16) {
    if (!tc_supported(m)) {
      set_error("SNB_PREC_BF16 (tcgen05) supports the Building / Mission-Bay topologies only: width 256 or 512, hidden2 <= 256, "
                "pos_xyz_dim 12, pos_dir_dim 4, a skip layer for width 512 / mip (got width=%d, hidden2=%d)",
                m->d.width, m->d.hidden2);
      return SNB_EUNSUPPORTED;
    }
    return tc_forward(m, x, S, sigma_noise, opts, out, moe_idx, l_aux, dbg_gates, dbg_loc, ws, st);
  }
  set_error("unknown precision %d", precision);
  return SNB_EINVAL;
}

int snb_moe_layer_forward(snb_model_t* mm, const float* input, const float* gate_input, int64_t S,
                          const snb_route_opts* opts, float* y, int32_t* moe_idx, float* l_aux, void* workspace,
                          size_t workspace_bytes, void* stream) {
  Model* m = (Model*)mm;
  SNB_REQUIRE(m && opts, "snb_moe_layer_forward: NULL model/opts");
  SNB_REQUIRE(S >= 0 && S < (1ll << 31) / 16, "snb_moe_layer_forward: bad S");
  SNB_REQUIRE(S == 0 || (input && y), "snb_moe_layer_forward: NULL input/output");
  SNB_REQUIRE(opts->capacity_factor > 0, "capacity_factor must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 0) {
    if (l_aux) SNB_CHECK_CUDA(cudaMemsetAsync(l_aux, 0, sizeof(float), st));
    return SNB_OK;
  }
  Arena ws(workspace, workspace_bytes);
  return fp32_moe_layer(m, input, gate_input ? gate_input : input, S, opts, y, moe_idx, l_aux, ws, st);
}

int snb_composite(const float* z, const float* raw, const float* last_delta, int64_t n_rays, int32_t n_samples,
                  int32_t white_bkgd, float* rgb, float* depth, float* depth_variance, float* bg_lambda,
                  float* weights, void* stream) {
  SNB_REQUIRE(n_rays == 0 || (z && raw), "snb_composite: NULL pointer");
  return composite_launch(z, raw, last_delta, n_rays, n_samples, white_bkgd, rgb, depth, depth_variance, bg_lambda,
                          weights, (cudaStream_t)stream);
}

int snb_sample_pdf(const float* bins, const float* weights, const float* u, int64_t n_rays, int32_t n_bins_minus1,
                   int32_t n_fine, float* z_fine, void* stream) {
  SNB_REQUIRE(n_rays == 0 || (bins && weights && z_fine), "snb_sample_pdf: NULL pointer");
  return sample_pdf_launch(bins, n_bins_minus1 + 1, weights, n_bins_minus1, 0, u, n_rays, n_bins_minus1, n_fine, 0,
                           u == nullptr, z_fine, (cudaStream_t)stream);
}

static size_t render_ws_layout(const Model* m, int64_t N, const snb_render_opts* o, size_t* model_ws) {
  const int Sc = o->coarse_samples, Sf = o->fine_samples;
  const int64_t Smax = (int64_t)N * (Sc > Sf ? Sc : Sf);
  int64_t chunk = o->model_chunk_size < Smax ? o->model_chunk_size : Smax;
  if (chunk < 1) chunk = 1;
  size_t mw = align_up(snb_workspace_bytes((const snb_model_t*)m, chunk, o->route.capacity_factor), 256);
  if (model_ws) *model_ws = mw;
  size_t b = 4 * mw;            // four chunk workspaces: consecutive chunks are software-pipelined
  auto add = [&](size_t n) { b += align_up(n * sizeof(float), 256); };
  add((size_t)N * Sc);            // zc
  add((size_t)N * Sc);            // weights_c
  add((size_t)N * (Sc > 1 ? Sc - 1 : 1));  // zmid
  add((size_t)N * (Sf > 0 ? Sf : 1));      // zf
  add((size_t)Smax * m->x_cols);  // x
  add((size_t)N * Sc * 4);        // raw_c
  add((size_t)N * (Sf > 0 ? Sf : 1) * 4);  // raw_f
  add((size_t)N * (Sc > Sf ? Sc : Sf));    // moe idx scratch (int)
  return b + 4096;
}

size_t snb_moe_backward_workspace_bytes(const snb_model_t* mm, int64_t S, double capacity_factor) {
  if (!mm) return 0;
  return fp32_backward_workspace_bytes((const Model*)mm, S, capacity_factor) + 4096;
}

int snb_moe_backward(snb_model_t* mm, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* opts,
                     const float* d_out, const float* d_l_aux, const snb_grads* grads, void* workspace, size_t workspace_bytes,
                     void* stream) {
  Model* m = (Model*)mm;
  SNB_REQUIRE(m && opts && grads && workspace, "snb_moe_backward: NULL argument");
  SNB_REQUIRE(S >= 0 && (S == 0 || (x && d_out)), "snb_moe_backward: NULL input");
  SNB_REQUIRE(opts->capacity_factor > 0, "capacity_factor must be > 0");
  SNB_REQUIRE(!m->d.mip, "snb_moe_backward: MipNeRFMoE backward is not implemented");
  SNB_REQUIRE(!m->ep, "snb_moe_backward: expert-parallel backward is not implemented");
  Arena ws(workspace, workspace_bytes);
  return fp32_backward(m, x, S, sigma_noise, opts, d_out, d_l_aux, grads, ws, (cudaStream_t)stream);
}

int snb_composite_backward(const float* z, const float* raw, const float* last_delta, int64_t n_rays, int32_t n_samples,
                           const float* d_rgb, float* d_raw, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 0, "snb_composite_backward: bad sizes");
  SNB_REQUIRE(n_rays == 0 || n_samples == 0 || (z && raw && d_rgb && d_raw), "snb_composite_backward: NULL pointer");
  return composite_backward_launch(z, raw, last_delta, n_rays, n_samples, d_rgb, d_raw, (cudaStream_t)stream);
}

int snb_get_rays(int32_t W, int32_t H, float fx, float fy, float cx, float cy, int32_t center_pixels, const float* c2w,
                 float near, float far, const float* altitude_range, float* rays, void* stream) {
  SNB_REQUIRE(W >= 0 && H >= 0 && (int64_t)W * H < (1ll << 31), "snb_get_rays: bad image size %d x %d", W, H);
  SNB_REQUIRE(W == 0 || H == 0 || (c2w && rays), "snb_get_rays: NULL pointer");
  SNB_REQUIRE(fx != 0.f && fy != 0.f, "snb_get_rays: zero focal length");
  return get_rays_launch(W, H, fx, fy, cx, cy, center_pixels, c2w, near, far, altitude_range, rays, (cudaStream_t)stream);
}

size_t snb_render_workspace_bytes(const snb_model_t* mm, int64_t n_rays, const snb_render_opts* o) {
  if (!mm || !o) return 0;
  return render_ws_layout((const Model*)mm, n_rays, o, nullptr);
}

int snb_render_rays(snb_model_t* mm, const float* rays, const int32_t* image_indices, const float* last_delta,
                    int64_t N, const snb_render_opts* o, const snb_render_out* out, void* workspace,
                    size_t workspace_bytes, void* stream) {
  Model* m = (Model*)mm;
  SNB_REQUIRE(m && o && out, "snb_render_rays: NULL argument");
  SNB_REQUIRE(!m->d.mip, "snb_render_rays: this entry point renders NeRFMoE; mip models use snb_render_rays_mip");
  SNB_REQUIRE(N >= 0 && (N == 0 || rays), "snb_render_rays: bad rays");
  SNB_REQUIRE(o->coarse_samples >= 2 && o->fine_samples >= 0, "snb_render_rays: bad sample counts");
  SNB_REQUIRE(o->model_chunk_size >= 1, "snb_render_rays: bad model_chunk_size");
  SNB_REQUIRE(o->fine_samples == 0 || o->coarse_samples >= 3, "fine sampling needs >= 3 coarse samples");
  if (N == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int Sc = o->coarse_samples, Sf = o->fine_samples;
  size_t model_ws = 0;
  size_t need = render_ws_layout(m, N, o, &model_ws);
  if (workspace_bytes < need) { set_error("snb_render_rays: workspace %zu < required %zu", workspace_bytes, need); return SNB_EWORKSPACE; }
  Arena a(workspace, workspace_bytes);
  char* mws = a.take<char>(4 * model_ws);
  float* zc = a.take<float>((size_t)N * Sc);
  float* wc = a.take<float>((size_t)N * Sc);
  float* zmid = a.take<float>((size_t)N * (Sc - 1));
  float* zf = a.take<float>((size_t)N * (Sf > 0 ? Sf : 1));
  const int64_t Smax = (int64_t)N * (Sc > Sf ? Sc : Sf);
  float* x = a.take<float>((size_t)Smax * m->x_cols);
  float* raw_c = a.take<float>((size_t)N * Sc * 4);
  float* raw_f = a.take<float>((size_t)N * (Sf > 0 ? Sf : 1) * 4);
  if (!a.ok) { set_error("snb_render_rays: workspace carve failed"); return SNB_EWORKSPACE; }
  if (out->raw_coarse) raw_c = out->raw_coarse;
  if (out->raw_fine && Sf > 0) raw_f = out->raw_fine;
  if (out->z_fine && Sf > 0) zf = out->z_fine;
  if (out->z_coarse) zc = out->z_coarse;

  auto run_pass = [&](const float* z, int Sn, float* raw, int32_t* gates_out, float* loss_out, const float* noise) -> int {
    const int64_t B = N * Sn;
    int rc;
    if (o->precision == SNB_PREC_BF16 && tc_supported(m) && tc_ray_source_ok(m)) {
      // the [N*Sn, 7] point tensor of rendering.py:357-362 is never built: both fused kernels rebuild a row from its ray
      // (32 B, cached) and its depth (4 B) where they consume it
      const RaySource rs = {rays, z, image_indices, Sn};
      return tc_forward_chunks(m, nullptr, B, o->model_chunk_size, &o->route, raw, gates_out, loss_out, mws, model_ws, 4, st,
                               noise, &rs);
    }
    if ((rc = fill_x_launch(rays, image_indices, z, N, Sn, x, st))) return rc;
    if (o->precision == SNB_PREC_BF16 && tc_supported(m))
      return tc_forward_chunks(m, x, B, o->model_chunk_size, &o->route, raw, gates_out, loss_out, mws, model_ws, 4, st, noise);
    int ci = 0;
    for (int64_t i = 0; i < B; i += o->model_chunk_size, ++ci) {       // rendering.py:354
      const int64_t rows = (B - i < o->model_chunk_size) ? (B - i) : o->model_chunk_size;
      rc = snb_moe_forward(mm, x + i * m->x_cols, rows, noise ? noise + i : nullptr, &o->route, o->precision, raw + i * 4,
                           gates_out ? gates_out + i : nullptr, loss_out ? loss_out + ci : nullptr, nullptr, nullptr,
                           mws, model_ws, st);
      if (rc) return rc;
    }
    return SNB_OK;
  };

  int rc;
  if ((rc = coarse_z_launch(rays, N, Sc, o->perturb, o->seed, zc, st))) return rc;
  if ((rc = run_pass(zc, Sc, raw_c, out->moe_gates_coarse, out->gate_loss_coarse, o->sigma_noise_coarse))) return rc;
  // bg-NeRF rays: last_delta - max(z) per level (rendering.py:215-216, 250-251); zmid / wc are free at those points
  const bool adj = o->last_delta_minus_zmax && last_delta;
  const float* ld_c = last_delta;
  if (adj) {
    float* t = (Sf == 0) ? wc : zmid;
    if ((rc = last_delta_adj_launch(last_delta, zc, N, Sc, t, st))) return rc;
    ld_c = t;
  }
  if (Sf == 0) {
    return composite_launch(zc, raw_c, ld_c, N, Sc, o->white_bkgd, out->rgb, out->depth, out->depth_variance,
                            out->bg_lambda, nullptr, st);
  }
  // coarse weights -> pdf over the interior bins (rendering.py:237-241)
  if ((rc = composite_launch(zc, raw_c, ld_c, N, Sc, 0, nullptr, nullptr, nullptr, nullptr, wc, st))) return rc;
  if ((rc = zmid_launch(zc, N, Sc, zmid, st))) return rc;
  if ((rc = sample_pdf_launch(zmid, Sc - 1, wc, Sc, 1, nullptr, N, Sc - 2, Sf, o->seed, o->perturb == 0.f, zf, st))) return rc;
  if ((rc = run_pass(zf, Sf, raw_f, out->moe_gates_fine, out->gate_loss_fine, o->sigma_noise_fine))) return rc;
  const float* ld_f = last_delta;
  if (adj) {
    if ((rc = last_delta_adj_launch(last_delta, zf, N, Sf, wc, st))) return rc;   // the max is over the fine samples only
    ld_f = wc;
  }
  // perturb == 0: z_fine comes from an ascending u through a monotone cdf, z_coarse is a linspace -> both sorted
  return merge_composite_launch(zf, zc, raw_f, raw_c, ld_f, N, Sf, Sc, o->perturb == 0.f, o->white_bkgd, out->rgb, out->depth,
                                out->depth_variance, out->bg_lambda, st);
}


static size_t render_mip_ws_layout(const Model* m, int64_t N, const snb_render_opts* o, size_t* model_ws) {
  const int Sc = o->coarse_samples, Sf = o->fine_samples;
  const int Smax = (Sc > Sf ? Sc : Sf) - 1;
  const int64_t Bmax = (int64_t)N * Smax;
  int64_t chunk = o->model_chunk_size < Bmax ? o->model_chunk_size : Bmax;
  if (chunk < 1) chunk = 1;
  size_t mw = align_up(snb_workspace_bytes((const snb_model_t*)m, chunk, o->route.capacity_factor), 256);
  if (model_ws) *model_ws = mw;
  size_t b = 4 * mw;                         // workspace sets of the chunk pipeline (tc_forward_chunks)
  auto add = [&](size_t n) { b += align_up(n * sizeof(float), 256); };
  add((size_t)N * Sc);                       // coarse edges
  add((size_t)N * (Sf > 0 ? Sf : 1));        // fine edges
  add((size_t)N * (Sc - 1));                 // coarse weights
  add((size_t)Bmax * m->x_cols);             // x
  add((size_t)N * (Sc - 1) * 4);             // raw coarse
  add((size_t)N * (Sf > 1 ? Sf - 1 : 1) * 4);  // raw fine
  return b + 4096;
}

size_t snb_render_mip_workspace_bytes(const snb_model_t* mm, int64_t n_rays, const snb_render_opts* o) {
  if (!mm || !o) return 0;
  return render_mip_ws_layout((const Model*)mm, n_rays, o, nullptr);
}

int snb_render_rays_mip(snb_model_t* mm, const float* rays, const float* radii, const int32_t* image_indices,
                        const float* last_delta, int64_t N, const snb_render_opts* o, float weights_resample_padding,
                        float rgb_padding, const snb_render_out* out, void* workspace, size_t workspace_bytes,
                        void* stream) {
  Model* m = (Model*)mm;
  SNB_REQUIRE(m && o && out, "snb_render_rays_mip: NULL argument");
  SNB_REQUIRE(m->d.mip, "snb_render_rays_mip needs a MipNeRFMoE model (desc.mip = 1)");
  SNB_REQUIRE(N >= 0 && (N == 0 || (rays && radii)), "snb_render_rays_mip: bad rays/radii");
  SNB_REQUIRE(o->coarse_samples >= 3 && (o->fine_samples == 0 || o->fine_samples >= 2), "snb_render_rays_mip: bad sample counts");
  SNB_REQUIRE(o->model_chunk_size >= 1, "snb_render_rays_mip: bad model_chunk_size");
  if (N == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int Sc = o->coarse_samples, Sf = o->fine_samples;
  size_t model_ws = 0;
  size_t need = render_mip_ws_layout(m, N, o, &model_ws);
  if (workspace_bytes < need) { set_error("snb_render_rays_mip: workspace %zu < required %zu", workspace_bytes, need); return SNB_EWORKSPACE; }
  Arena a(workspace, workspace_bytes);
  char* mws = a.take<char>(4 * model_ws);
  float* zc = a.take<float>((size_t)N * Sc);
  float* zf = a.take<float>((size_t)N * (Sf > 0 ? Sf : 1));
  float* wc = a.take<float>((size_t)N * (Sc - 1));
  const int Smax = (Sc > Sf ? Sc : Sf) - 1;
  float* x = a.take<float>((size_t)N * Smax * m->x_cols);
  float* raw_c = a.take<float>((size_t)N * (Sc - 1) * 4);
  float* raw_f = a.take<float>((size_t)N * (Sf > 1 ? Sf - 1 : 1) * 4);
  if (!a.ok) { set_error("snb_render_rays_mip: workspace carve failed"); return SNB_EWORKSPACE; }
  if (out->raw_coarse) raw_c = out->raw_coarse;
  if (out->raw_fine && Sf > 0) raw_f = out->raw_fine;
  if (out->z_fine && Sf > 0) zf = out->z_fine;
  const float pad = rgb_padding < 0.f ? 0.f : rgb_padding;

  auto run_pass = [&](const float* ze, int Se, float* raw, int32_t* gates_out, float* loss_out) -> int {
    int rc = mip_fill_x_launch(rays, radii, image_indices, ze, N, Se, x, st);
    if (rc) return rc;
    const int64_t B = N * (Se - 1);
    if (o->precision == SNB_PREC_BF16 && tc_supported(m))
      return tc_forward_chunks(m, x, B, o->model_chunk_size, &o->route, raw, gates_out, loss_out, mws, model_ws, 4, st);
    int ci = 0;
    for (int64_t i = 0; i < B; i += o->model_chunk_size, ++ci) {       // rendering_mip.py:327
      const int64_t rows = (B - i < o->model_chunk_size) ? (B - i) : o->model_chunk_size;
      rc = snb_moe_forward(mm, x + i * m->x_cols, rows, nullptr, &o->route, o->precision, raw + i * 4,
                           gates_out ? gates_out + i : nullptr, loss_out ? loss_out + ci : nullptr, nullptr, nullptr,
                           mws, model_ws, st);
      if (rc) return rc;
    }
    return SNB_OK;
  };
  int rc;
  if ((rc = coarse_z_launch(rays, N, Sc, o->perturb, o->seed, zc, st))) return rc;       // rendering_mip.py:147-160
  if ((rc = run_pass(zc, Sc, raw_c, out->moe_gates_coarse, out->gate_loss_coarse))) return rc;
  const bool only_coarse = (Sf == 0);
  if ((rc = mip_composite_launch(zc, raw_c, last_delta, N, Sc, pad, o->white_bkgd,
                                 only_coarse ? (out->rgb ? out->rgb : out->rgb_coarse) : out->rgb_coarse,
                                 only_coarse ? out->depth : nullptr, only_coarse ? out->depth_variance : nullptr,
                                 only_coarse ? nullptr : wc, st)))
    return rc;
  if (only_coarse) return SNB_OK;
  if ((rc = mip_resample_launch(zc, wc, N, Sc, Sf, weights_resample_padding, o->resample_randomized, o->seed, zf, st))) return rc;
  if ((rc = run_pass(zf, Sf, raw_f, out->moe_gates_fine, out->gate_loss_fine))) return rc;
  return mip_composite_launch(zf, raw_f, last_delta, N, Sf, pad, o->white_bkgd, out->rgb, out->depth,
                              out->depth_variance, nullptr, st);
}

int snb_umma_selftest(const void* a_bf16, const void* b_bf16, int32_t N, int32_t K, float* d, int32_t variant,
                      void* stream) {
  SNB_REQUIRE(a_bf16 && b_bf16 && d, "snb_umma_selftest: NULL pointer");
  return umma_selftest(a_bf16, b_bf16, N, K, d, variant, (cudaStream_t)stream);
}

int snb_umma_microbench(int32_t N, int32_t a_in_tmem, int32_t flags, int32_t reps, uint64_t* out6, void* stream) {
  SNB_REQUIRE(out6, "snb_umma_microbench: NULL out");
  return umma_microbench(N, a_in_tmem, flags, reps, (unsigned long long*)out6, (cudaStream_t)stream);
}

// ---- f3: background NeRF + sphere parametrisation (snb_bg.cu) ----
int snb_bg_create(const snb_bg_desc* desc, const snb_bg_weights* w, void* stream, snb_bg_model_t** out) {
  return bg_create(desc, w, (cudaStream_t)stream, (BgModel**)out);
}
int snb_bg_update(snb_bg_model_t* m, const snb_bg_weights* w, void* stream) {
  SNB_REQUIRE(m && w, "snb_bg_update: NULL argument");
  return bg_update((BgModel*)m, w, (cudaStream_t)stream);
}
void snb_bg_destroy(snb_bg_model_t* m) { bg_destroy((BgModel*)m); }
size_t snb_bg_workspace_bytes(const snb_bg_model_t* m, int64_t S) { return m ? bg_workspace_bytes((const BgModel*)m, S) : 0; }
int snb_bg_forward(snb_bg_model_t* m, const float* x, int64_t S, const float* sigma_noise, float* out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  SNB_REQUIRE(m && S >= 0 && (S == 0 || (x && out)), "snb_bg_forward: bad argument");
  SNB_REQUIRE(S == 0 || workspace, "snb_bg_forward: NULL workspace");
  Arena a(workspace, workspace_bytes);
  return bg_forward((BgModel*)m, x, S, sigma_noise, out, a, (cudaStream_t)stream);
}
int snb_intersect_sphere(const float* rays, int64_t N, const float* sphere_center, const float* sphere_radius, float* fg_far,
                         int32_t* bad, void* stream) {
  SNB_REQUIRE(N >= 0 && (N == 0 || (rays && fg_far)), "snb_intersect_sphere: bad argument");
  SNB_REQUIRE((sphere_center == nullptr) == (sphere_radius == nullptr), "snb_intersect_sphere: centre and radius come together");
  return intersect_sphere_launch(rays, N, sphere_center, sphere_radius, fg_far, bad, (cudaStream_t)stream);
}
int snb_depth2pts_outside(const float* rays, const float* sphere_center, const float* sphere_radius, const float* z, int64_t N,
                          int32_t S, float* pts, float* depth_real, void* stream) {
  SNB_REQUIRE(N >= 0 && S >= 0 && (N == 0 || S == 0 || (rays && z && pts && depth_real)), "snb_depth2pts_outside: bad argument");
  SNB_REQUIRE((sphere_center == nullptr) == (sphere_radius == nullptr), "snb_depth2pts_outside: centre and radius come together");
  return depth2pts_outside_launch(rays, sphere_center, sphere_radius, z, N, S, pts, depth_real, (cudaStream_t)stream);
}

}  // extern "C"

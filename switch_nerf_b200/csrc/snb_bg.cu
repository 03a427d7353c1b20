// Background NeRF + sphere parametrisation: fp32 CUDA-core path.
//   models/nerf.py:75-191     NeRF.forward with xyz_dim = 4: Embedding(4 dims) -> `layers` ReLU layers (the encoded input
//                             concatenated in front of the hidden state at `skip_layer`) -> sigma head (+noise, activation),
//                             xyz_encoding_final -> [final | PE(dir) | appearance] -> dir_a_encoding (ReLU) -> rgb (sigmoid)
//   rendering.py:497-518      _intersect_sphere
//   rendering.py:521-570      _depth2pts_outside (include_xyz_real = False)
// The GEMMs are the 64x64x16 fp32 tiles of the fp32 model path (snb_fp32.cu: linear_launch); the concatenation at the
// skip layer is two GEMMs, the second one adding the first as a residual before the ReLU.
#include "snb_common.cuh"

namespace snb {

int linear_launch(int act, const float* A, int lda, const float* W, const float* b, const float* R, int ldr, float* C,
                  int ldc, int64_t rows, int N, int K, const float* noise, cudaStream_t st);

struct BgModel {
  snb_bg_desc d;
  int in0, dir_in, cat_in;
  float* blob = nullptr;
  size_t floats = 0;
  float *w[16], *b[16], *w_skip_pe, *w_skip_h, *final_w, *final_b, *dir_w, *dir_b, *sigma_w, *sigma_b, *rgb_w, *rgb_b, *emb_a;
};

// [x(4), sin(2^k x), cos(2^k x)]_k -> pe [S, in0]; [PE(dir) | appearance] -> cat[:, W:]
__global__ void k_bg_encode(const float* __restrict__ x, int64_t S, int F, int Fd, int A, int count,
                            const float* __restrict__ emb_a, float* __restrict__ pe, int ld_pe, float* __restrict__ cat,
                            int ld_cat, int cat_off) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (s >= S) return;
  const float* xr = x + s * 8;
  const int n_pe = 4 + 8 * F, n_dir = 3 + 6 * Fd;
  for (int c = threadIdx.x; c < n_pe + n_dir + A; c += blockDim.x) {
    if (c < n_pe) {
      float v;
      if (c < 4) v = xr[c];
      else {
        const int k = (c - 4) / 8, r = (c - 4) % 8, ax = r % 4;
        const float a = (float)(1 << k) * xr[ax];
        v = (r < 4) ? sinf(a) : cosf(a);
      }
      pe[s * ld_pe + c] = v;
    } else if (c < n_pe + n_dir) {
      const int cc = c - n_pe;
      float v;
      if (cc < 3) v = xr[4 + cc];
      else {
        const int k = (cc - 3) / 6, r = (cc - 3) % 6, ax = r % 3;
        const float a = (float)(1 << k) * xr[4 + ax];
        v = (r < 3) ? sinf(a) : cosf(a);
      }
      cat[s * ld_cat + cat_off + cc] = v;
    } else {
      const int cc = c - n_pe - n_dir;
      int ai = (int)xr[7];
      ai = min(max(ai, 0), count - 1);
      cat[s * ld_cat + cat_off + n_dir + cc] = emb_a[(int64_t)ai * A + cc];
    }
  }
}

__global__ void k_bg_pack_out(const float* __restrict__ rgb, const float* __restrict__ sigma, int64_t S, int relu_sigma,
                              float* __restrict__ out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float sg = relu_sigma ? fmaxf(sigma[s], 0.f) : sigma[s];
  reinterpret_cast<float4*>(out)[s] = make_float4(rgb[s * 3], rgb[s * 3 + 1], rgb[s * 3 + 2], sg);
}

// W [N, in0 + width] (encoded input first, models/nerf.py:154 cat([input_xyz, xyz_])) -> two row-major matrices
__global__ void k_split_cols(const float* __restrict__ w, int N, int K0, int K1, float* __restrict__ a, float* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * (K0 + K1)) return;
  const int n = i / (K0 + K1), k = i % (K0 + K1);
  if (k < K0) a[n * K0 + k] = w[i];
  else b[n * K1 + (k - K0)] = w[i];
}

static int bg_upload(BgModel* m, const snb_bg_weights* w, cudaStream_t st) {
  const snb_bg_desc& d = m->d;
  const int W = d.width;
  auto cp = [&](float* dst, const float* src, size_t n) -> int {
    SNB_REQUIRE(src, "snb_bg: a weight pointer is NULL");
    SNB_CHECK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SNB_OK;
  };
  int rc;
  for (int i = 0; i < d.layers; ++i) {
    const int in_i = (i == 0) ? m->in0 : W;
    if (i == d.skip_layer && i > 0) {
      SNB_REQUIRE(w->w[i], "snb_bg: a weight pointer is NULL");
      k_split_cols<<<(unsigned)cdiv((int64_t)W * (m->in0 + W), 256), 256, 0, st>>>(w->w[i], W, m->in0, W, m->w_skip_pe, m->w_skip_h);
      SNB_CHECK_LAUNCH("k_split_cols");
    } else if ((rc = cp(m->w[i], w->w[i], (size_t)W * in_i))) return rc;
    if ((rc = cp(m->b[i], w->b[i], W))) return rc;
  }
  if ((rc = cp(m->final_w, w->final_w, (size_t)W * W))) return rc;
  if ((rc = cp(m->final_b, w->final_b, W))) return rc;
  if ((rc = cp(m->dir_w, w->dir_w, (size_t)(W / 2) * m->cat_in))) return rc;
  if ((rc = cp(m->dir_b, w->dir_b, W / 2))) return rc;
  if ((rc = cp(m->sigma_w, w->sigma_w, W))) return rc;
  if ((rc = cp(m->sigma_b, w->sigma_b, 1))) return rc;
  if ((rc = cp(m->rgb_w, w->rgb_w, (size_t)3 * (W / 2)))) return rc;
  if ((rc = cp(m->rgb_b, w->rgb_b, 3))) return rc;
  if (d.appearance_dim > 0 && (rc = cp(m->emb_a, w->emb_a, (size_t)d.appearance_count * d.appearance_dim))) return rc;
  return SNB_OK;
}

int bg_create(const snb_bg_desc* d, const snb_bg_weights* w, cudaStream_t st, BgModel** out) {
  SNB_REQUIRE(d && w && out, "snb_bg_create: NULL argument");
  SNB_REQUIRE(d->layers >= 1 && d->layers <= 16 && d->width >= 2 && d->width % 2 == 0 && d->skip_layer < d->layers,
              "snb_bg_create: bad topology (layers %d, width %d, skip %d)", d->layers, d->width, d->skip_layer);
  SNB_REQUIRE(d->pos_xyz_freqs >= 0 && d->pos_xyz_freqs <= 16 && d->pos_dir_freqs > 0 && d->pos_dir_freqs <= 16,
              "snb_bg_create: the background model of the hot path encodes position and direction (pos_dir_dim > 0)");
  SNB_REQUIRE(d->appearance_dim >= 0 && (d->appearance_dim == 0 || d->appearance_count > 0), "snb_bg_create: bad appearance table");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("snb_bg_create: no CUDA device (there is no CPU path)"); return SNB_ECUDA; }
  BgModel* m = new BgModel();
  m->d = *d;
  const int W = d->width;
  m->in0 = 4 + 8 * d->pos_xyz_freqs;
  m->dir_in = 3 + 6 * d->pos_dir_freqs;
  m->cat_in = W + m->dir_in + d->appearance_dim;
  size_t n = 0;
  auto add = [&](size_t k) { size_t o = n; n += (k + 63) / 64 * 64; return o; };
  size_t ow[16], ob[16];
  for (int i = 0; i < d->layers; ++i) { ow[i] = add((size_t)W * ((i == 0) ? m->in0 : W)); ob[i] = add(W); }
  const size_t o_sp = add((size_t)W * m->in0), o_sh = add((size_t)W * W);
  const size_t o_fw = add((size_t)W * W), o_fb = add(W), o_dw = add((size_t)(W / 2) * m->cat_in), o_db = add(W / 2);
  const size_t o_sw = add(W), o_sb = add(1), o_rw = add((size_t)3 * (W / 2)), o_rb = add(3);
  const size_t o_e = add((size_t)d->appearance_count * d->appearance_dim + 1);
  cudaError_t e = cudaMalloc((void**)&m->blob, n * sizeof(float));
  if (e != cudaSuccess) { set_error("snb_bg_create: cudaMalloc failed: %s", cudaGetErrorString(e)); delete m; return SNB_ECUDA; }
  m->floats = n;
  for (int i = 0; i < d->layers; ++i) { m->w[i] = m->blob + ow[i]; m->b[i] = m->blob + ob[i]; }
  m->w_skip_pe = m->blob + o_sp; m->w_skip_h = m->blob + o_sh;
  m->final_w = m->blob + o_fw; m->final_b = m->blob + o_fb; m->dir_w = m->blob + o_dw; m->dir_b = m->blob + o_db;
  m->sigma_w = m->blob + o_sw; m->sigma_b = m->blob + o_sb; m->rgb_w = m->blob + o_rw; m->rgb_b = m->blob + o_rb;
  m->emb_a = m->blob + o_e;
  int rc = bg_upload(m, w, st);
  if (rc) { cudaFree(m->blob); delete m; return rc; }
  *out = m;
  return SNB_OK;
}
int bg_update(BgModel* m, const snb_bg_weights* w, cudaStream_t st) { return bg_upload(m, w, st); }
void bg_destroy(BgModel* m) {
  if (!m) return;
  if (m->blob) cudaFree(m->blob);
  delete m;
}
size_t bg_workspace_bytes(const BgModel* m, int64_t S) {
  if (S < 1) S = 1;
  const int W = m->d.width;
  size_t b = 0;
  auto add = [&](size_t k) { b += align_up(k * sizeof(float), 256); };
  add((size_t)S * m->in0); add((size_t)S * W); add((size_t)S * W); add((size_t)S * W);
  add((size_t)S * m->cat_in); add((size_t)S * (W / 2)); add(S); add((size_t)S * 3);
  return b + 4096;
}

int bg_forward(BgModel* m, const float* x, int64_t S, const float* noise, float* out, Arena& ws, cudaStream_t st) {
  const snb_bg_desc& d = m->d;
  const int W = d.width;
  if (S == 0) return SNB_OK;
  float* pe = ws.take<float>((size_t)S * m->in0);
  float* t0 = ws.take<float>((size_t)S * W);
  float* t1 = ws.take<float>((size_t)S * W);
  float* tp = ws.take<float>((size_t)S * W);
  float* cat = ws.take<float>((size_t)S * m->cat_in);
  float* hd = ws.take<float>((size_t)S * (W / 2));
  float* sigma = ws.take<float>(S);
  float* rgb = ws.take<float>((size_t)S * 3);
  if (!ws.ok) { set_error("snb_bg_forward: workspace too small"); return SNB_EWORKSPACE; }
  {
    dim3 blk(32, 8);
    k_bg_encode<<<(unsigned)cdiv(S, 8), blk, 0, st>>>(x, S, d.pos_xyz_freqs, d.pos_dir_freqs, d.appearance_dim, d.appearance_count,
                                                      m->emb_a, pe, m->in0, cat, m->cat_in, W);
    SNB_CHECK_LAUNCH("k_bg_encode");
  }
  int rc;
  const float* in = pe;
  int in_k = m->in0;
  float* o = t0;
  for (int i = 0; i < d.layers; ++i) {
    if (i == d.skip_layer && i > 0) {
      // relu(W [pe | h] + b) = relu(W_h h + b + W_pe pe)
      if ((rc = linear_launch(ACT_NONE, pe, m->in0, m->w_skip_pe, nullptr, nullptr, 0, tp, W, S, W, m->in0, nullptr, st))) return rc;
      if ((rc = linear_launch(ACT_RELU, in, in_k, m->w_skip_h, m->b[i], tp, W, o, W, S, W, W, nullptr, st))) return rc;
    } else {
      if ((rc = linear_launch(ACT_RELU, in, in_k, m->w[i], m->b[i], nullptr, 0, o, W, S, W, in_k, nullptr, st))) return rc;
    }
    in = o;
    in_k = W;
    o = (o == t0) ? t1 : t0;
  }
  // sigma head: Linear (+noise) then the activation (softplus(x - 1) fused; ReLU applied when the row is packed)
  if ((rc = linear_launch(d.shifted_softplus ? ACT_SOFTPLUS_SHIFT : ACT_NONE, in, W, m->sigma_w, m->sigma_b, nullptr, 0, sigma, 1, S, 1, W,
                          noise, st))) return rc;
  if ((rc = linear_launch(ACT_NONE, in, W, m->final_w, m->final_b, nullptr, 0, cat, m->cat_in, S, W, W, nullptr, st))) return rc;
  if ((rc = linear_launch(ACT_RELU, cat, m->cat_in, m->dir_w, m->dir_b, nullptr, 0, hd, W / 2, S, W / 2, m->cat_in, nullptr, st))) return rc;
  if ((rc = linear_launch(ACT_SIGMOID, hd, W / 2, m->rgb_w, m->rgb_b, nullptr, 0, rgb, 3, S, 3, W / 2, nullptr, st))) return rc;
  k_bg_pack_out<<<(unsigned)cdiv(S, 256), 256, 0, st>>>(rgb, sigma, S, d.shifted_softplus ? 0 : 1, out);
  SNB_CHECK_LAUNCH("k_bg_pack_out");
  return SNB_OK;
}

// ---- sphere geometry: one thread per ray / per sample, formulas and operation order of the reference ----
__device__ __forceinline__ void sphere_ray(const float* ray, const float* c, const float* rad, float (&o)[3], float (&d)[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    o[a] = ray[a];
    d[a] = ray[3 + a];
    if (rad) { o[a] = __fdiv_rn(o[a] - c[a], rad[a]); d[a] = __fdiv_rn(d[a], rad[a]); }
  }
}

__global__ void k_intersect_sphere(const float* __restrict__ rays, int64_t N, const float* __restrict__ c,
                                   const float* __restrict__ rad, float* __restrict__ fg_far, int* __restrict__ bad) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  float o[3], d[3];
  sphere_ray(rays + r * 8, c, rad, o, d);
  const float dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const float d1 = -__fdiv_rn(d[0] * o[0] + d[1] * o[1] + d[2] * o[2], dd);
  const float p0 = o[0] + d1 * d[0], p1 = o[1] + d1 * d[1], p2 = o[2] + d1 * d[2];
  const float cosv = __fdiv_rn(1.f, sqrtf(dd));
  const float pn = p0 * p0 + p1 * p1 + p2 * p2;
  if (pn >= 1.f && bad) *bad = 1;
  fg_far[r] = d1 + sqrtf(1.f - pn) * cosv;
}

__global__ void k_depth2pts_outside(const float* __restrict__ rays, const float* __restrict__ c, const float* __restrict__ rad,
                                    const float* __restrict__ z, int64_t N, int S, float* __restrict__ pts,
                                    float* __restrict__ depth_real) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * S) return;
  const int64_t r = i / S;
  float o[3], d[3];
  sphere_ray(rays + r * 8, c, rad, o, d);
  const float depth = z[i];
  const float dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const float d1 = -__fdiv_rn(d[0] * o[0] + d[1] * o[1] + d[2] * o[2], dd);
  float pm[3], ps[3], ax[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) pm[a] = o[a] + d1 * d[a];
  const float pmn = sqrtf(pm[0] * pm[0] + pm[1] * pm[1] + pm[2] * pm[2]);
  const float cosv = __fdiv_rn(1.f, sqrtf(dd));
  const float d2 = sqrtf(1.f - pmn * pmn) * cosv;
#pragma unroll
  for (int a = 0; a < 3; ++a) ps[a] = o[a] + (d1 + d2) * d[a];
  ax[0] = o[1] * ps[2] - o[2] * ps[1];
  ax[1] = o[2] * ps[0] - o[0] * ps[2];
  ax[2] = o[0] * ps[1] - o[1] * ps[0];
  const float an = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]) + 1e-8f;
#pragma unroll
  for (int a = 0; a < 3; ++a) ax[a] = __fdiv_rn(ax[a], an);
  const float phi = asinf(pmn), theta = asinf(pmn * depth);
  const float ang = phi - theta, ca = cosf(ang), sa = sinf(ang);
  const float cr[3] = {ax[1] * ps[2] - ax[2] * ps[1], ax[2] * ps[0] - ax[0] * ps[2], ax[0] * ps[1] - ax[1] * ps[0]};
  const float dot = ax[0] * ps[0] + ax[1] * ps[1] + ax[2] * ps[2];
  float q[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) q[a] = ps[a] * ca + cr[a] * sa + ax[a] * dot * (1.f - ca);
  const float qn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  float4 p = make_float4(__fdiv_rn(q[0], qn), __fdiv_rn(q[1], qn), __fdiv_rn(q[2], qn), depth);
  reinterpret_cast<float4*>(pts)[i] = p;
  depth_real[i] = __fdiv_rn(1.f, depth + 1e-8f) * cosf(theta) + d1;
}

int intersect_sphere_launch(const float* rays, int64_t N, const float* c, const float* rad, float* fg_far, int* bad, cudaStream_t st) {
  if (N == 0) return SNB_OK;
  k_intersect_sphere<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(rays, N, c, rad, fg_far, bad);
  SNB_CHECK_LAUNCH("k_intersect_sphere");
  return SNB_OK;
}
int depth2pts_outside_launch(const float* rays, const float* c, const float* rad, const float* z, int64_t N, int S, float* pts,
                             float* depth_real, cudaStream_t st) {
  if (N == 0 || S == 0) return SNB_OK;
  k_depth2pts_outside<<<(unsigned)cdiv(N * S, 256), 256, 0, st>>>(rays, c, rad, z, N, S, pts, depth_real);
  SNB_CHECK_LAUNCH("k_depth2pts_outside");
  return SNB_OK;
}

}  // namespace snb

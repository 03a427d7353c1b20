// Ray-side kernels of rendering.render_rays (reference rendering.py:15-196, 199-274, 277-494,
// 573-637) for bg_nerf=None / use_cascade=False: stratified coarse depths, point generation,
// alpha / exclusive transmittance / weights, inverse-CDF fine sampling, sorted coarse+fine merge
// and the RGB / depth / depth-variance composite.  All of it is HBM-bound elementwise / scan work
// (~20-44 B per point sample); one warp (composite, pdf) or one CTA (merge) per ray.
#include "snb_common.cuh"

namespace snb {

__device__ __forceinline__ float u01(uint64_t seed, uint64_t a, uint64_t b) {
  // splitmix64-style counter hash -> U[0,1) with 24 random bits (torch.rand_like semantics)
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (a * 0x100000001B3ull + b + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float linspace01(int i, int n) {
  // ATen linspace: step=(end-start)/(n-1); first half from the start, second half from the end
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(1.0f, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
}

// z = near*(1-t) + far*t  (+ stratified perturbation)   rendering.py:85-88, 573-584
__global__ void k_coarse_z(const float* __restrict__ rays, int64_t N, int Sc, float perturb, uint64_t seed,
                           float* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Sc) return;
  const int64_t r = i / Sc;
  const int j = (int)(i % Sc);
  const float near = rays[r * 8 + 6], far = rays[r * 8 + 7];
  auto zt = [&](int jj) {
    float t = linspace01(jj, Sc);
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
  };
  float zv = zt(j);
  if (perturb > 0.f) {
    float lo = (j == 0) ? zv : __fmul_rn(0.5f, __fadd_rn(zt(j - 1), zv));
    float hi = (j == Sc - 1) ? zv : __fmul_rn(0.5f, __fadd_rn(zv, zt(j + 1)));
    float u = perturb * u01(seed, (uint64_t)r, (uint64_t)j);
    zv = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u));
  }
  z[i] = zv;
}

// x[r*Sn + j] = [o + d*z, d, image_index]    rendering.py:90, 306-314, 357-362
__global__ void k_fill_x(const float* __restrict__ rays, const int* __restrict__ image_indices,
                         const float* __restrict__ z, int64_t N, int Sn, float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Sn) return;
  const int64_t r = i / Sn;
  const float* ray = rays + r * 8;
  const float zv = z[i];
  float* xr = x + i * 7;
  xr[0] = __fadd_rn(ray[0], __fmul_rn(ray[3], zv));
  xr[1] = __fadd_rn(ray[1], __fmul_rn(ray[4], zv));
  xr[2] = __fadd_rn(ray[2], __fmul_rn(ray[5], zv));
  xr[3] = ray[3];
  xr[4] = ray[4];
  xr[5] = ray[5];
  xr[6] = image_indices ? (float)image_indices[r] : 0.f;
}

// ------------------------------------------------------------------------------------------
// composite over already-ordered samples; one warp per ray.   rendering.py:436-494 (flip=False)
// fetch(j) -> (z, r, g, b, sigma) abstracts direct ([N,S]) vs merged (smem order) access.
// ------------------------------------------------------------------------------------------
template <typename Fetch>
__device__ __forceinline__ void warp_composite(int S, float last_delta, int white_bkgd, Fetch fetch, float* w_smem,
                                               float* z_smem, float* rgb_out, float* depth_out, float* var_out,
                                               float* lam_out, float* weights_out) {
  const int lane = threadIdx.x & 31;
  float carry = 1.f;                       // running inclusive product of (1 - alpha + 1e-8)
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_w = 0.f;
  for (int j0 = 0; j0 < S; j0 += 32) {
    const int j = j0 + lane;
    float z = 0.f, zn = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, sg = 0.f;
    if (j < S) fetch(j, z, cr, cg, cb, sg);
    zn = __shfl_down_sync(0xffffffffu, z, 1);
    if (lane == 31 && j + 1 < S) { float a, b, c, d; fetch(j + 1, zn, a, b, c, d); }
    float delta = (j == S - 1) ? last_delta : (zn - z);
    float alpha = (j < S) ? (1.f - expf(-delta * sg)) : 0.f;
    float f = (j < S) ? (1.f - alpha + 1e-8f) : 1.f;
    // inclusive product scan across the warp
    float p = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float q = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) p *= q;
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.f;
    const float T = carry * excl;
    const float w = alpha * T;
    if (j < S) {
      acc_r += w * cr; acc_g += w * cg; acc_b += w * cb; acc_d += w * z; acc_w += w;
      if (w_smem) { w_smem[j] = w; z_smem[j] = z; }
      if (weights_out) weights_out[j] = w;
    }
    carry *= __shfl_sync(0xffffffffu, p, 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_r += __shfl_xor_sync(0xffffffffu, acc_r, o);
    acc_g += __shfl_xor_sync(0xffffffffu, acc_g, o);
    acc_b += __shfl_xor_sync(0xffffffffu, acc_b, o);
    acc_d += __shfl_xor_sync(0xffffffffu, acc_d, o);
    acc_w += __shfl_xor_sync(0xffffffffu, acc_w, o);
  }
  float var = 0.f;
  if (var_out && w_smem) {
    __syncwarp();
    for (int j = lane; j < S; j += 32) { float d = z_smem[j] - acc_d; var += w_smem[j] * d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  }
  if (lane == 0) {
    if (rgb_out) {
      float bg = white_bkgd ? (1.f - acc_w) : 0.f;
      rgb_out[0] = acc_r + bg; rgb_out[1] = acc_g + bg; rgb_out[2] = acc_b + bg;
    }
    if (depth_out) *depth_out = acc_d;
    if (var_out) *var_out = var;
    if (lam_out) *lam_out = carry;          // T[..., -1]  (bg_lambda, rendering.py:456-457)
  }
}

__global__ void __launch_bounds__(128) k_composite(const float* __restrict__ z, const float* __restrict__ raw,
                                                   const float* __restrict__ last_delta, int64_t N, int S,
                                                   int white_bkgd, float* __restrict__ rgb, float* __restrict__ depth,
                                                   float* __restrict__ var, float* __restrict__ lam,
                                                   float* __restrict__ weights) {
  extern __shared__ float sm[];  // [4 warps][2][S]
  const int w = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 4 + w;
  if (r >= N) return;
  float* ws = sm + (size_t)w * 2 * S;
  const float* zr = z + r * S;
  const float4* rr = reinterpret_cast<const float4*>(raw) + r * S;
  auto fetch = [&](int j, float& zz, float& cr, float& cg, float& cb, float& sg) {
    zz = zr[j];
    float4 v = rr[j];
    cr = v.x; cg = v.y; cb = v.z; sg = v.w;
  };
  warp_composite(S, last_delta ? last_delta[r] : 1e10f, white_bkgd, fetch, ws, ws + S, rgb ? rgb + r * 3 : nullptr,
                 depth ? depth + r : nullptr, var ? var + r : nullptr, lam ? lam + r : nullptr,
                 weights ? weights + r * S : nullptr);
}

// ------------------------------------------------------------------------------------------
// inverse-CDF sampling; one warp per ray.   rendering.py:587-637
//   bins [N, nb+1] (z mid points), weights [N, nb] (already sliced [1:-1]), u [N, nf] or null (det)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sample_pdf(const float* __restrict__ bins, int ld_bins,
                                                    const float* __restrict__ weights, int ld_w, int w_off,
                                                    const float* __restrict__ u_in, int64_t N, int nb, int nf,
                                                    uint64_t seed, int det, float* __restrict__ zf) {
  extern __shared__ float sm[];  // [4 warps][(nb+1) cdf + (nb+1) bins]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + w;
  if (r >= N) return;
  float* cdf = sm + (size_t)w * 2 * (nb + 1);
  float* sb = cdf + (nb + 1);
  const float* wr = weights + r * ld_w + w_off;
  float sum = 0.f;
  for (int j = lane; j < nb; j += 32) sum += wr[j] + 1e-8f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int j = lane; j <= nb; j += 32) sb[j] = bins[r * ld_bins + j];
  for (int j = lane; j < nb; j += 32) cdf[j + 1] = __fdiv_rn(wr[j] + 1e-8f, sum);   // pdf, scanned below
  __syncwarp();
  if (lane == 0) {
    // torch CPU cumsum accumulates float in double (acc_type<float,false>) sequentially
    double run = 0.0;
    cdf[0] = 0.f;
    for (int j = 1; j <= nb; ++j) { run += (double)cdf[j]; cdf[j] = (float)run; }
  }
  __syncwarp();
  for (int j = lane; j < nf; j += 32) {
    float u;
    if (u_in) u = u_in[r * nf + j];
    else if (det) u = linspace01(j, nf);
    else u = u01(seed ^ 0x5DEECE66Dull, (uint64_t)r, (uint64_t)j);
    // searchsorted(cdf, u, right=True): number of entries <= u
    int lo = 0, hi = nb + 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    const int below = max(lo - 1, 0), above = min(lo, nb);
    float denom = cdf[above] - cdf[below];
    if (denom < 1e-8f) denom = 1.f;
    zf[r * nf + j] = __fadd_rn(sb[below], __fmul_rn(__fdiv_rn(u - cdf[below], denom), sb[above] - sb[below]));
  }
}

// z_mid[r, j] = 0.5*(z[j] + z[j+1])   rendering.py:238
__global__ void k_zmid(const float* __restrict__ z, int64_t N, int S, float* __restrict__ mid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * (S - 1)) return;
  const int64_t r = i / (S - 1);
  const int j = (int)(i % (S - 1));
  mid[i] = __fmul_rn(0.5f, __fadd_rn(z[r * S + j], z[r * S + j + 1]));
}

// ------------------------------------------------------------------------------------------
// merge coarse + fine by depth (stable: cat order [fine, coarse], rendering.py:421) and composite.
// One CTA per ray: bitonic sort of 64-bit (ordered z bits << 32 | position) keys in shared memory.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_merge_composite(const float* __restrict__ zf, const float* __restrict__ zc,
                                                         const float* __restrict__ raw_f, const float* __restrict__ raw_c,
                                                         const float* __restrict__ last_delta, int64_t N, int Sf, int Sc,
                                                         int P2, int presorted, int white_bkgd,
                                                         float* __restrict__ rgb, float* __restrict__ depth,
                                                         float* __restrict__ var, float* __restrict__ lam) {
  extern __shared__ unsigned long long keys[];  // [P2] keys, then 2*S floats
  const int64_t r = blockIdx.x;
  const int S = Sf + Sc;
  float* wsm = reinterpret_cast<float*>(keys + P2);
  if (presorted) {
    // both lists ascending (deterministic sampling): merged position by rank -- fine element i goes to
    // i + #{coarse < zf[i]}, coarse element j to j + #{fine <= zc[j]} (ties: fine first == the stable sort)
    float* szf = wsm;
    float* szc = wsm + Sf;
    for (int i = threadIdx.x; i < Sf; i += blockDim.x) szf[i] = zf[r * Sf + i];
    for (int i = threadIdx.x; i < Sc; i += blockDim.x) szc[i] = zc[r * Sc + i];
    __syncthreads();
    // The rank formulas below need monotone lists.  Rounding in the lerp / linspace can leave a 1-ulp inversion
    // between neighbours; a running maximum (warps 0 and 1) removes it (the reference's sort would have swapped
    // the two samples instead -- a <= 1 ulp difference in one delta).
    if (threadIdx.x < 64) {
      float* lst = (threadIdx.x < 32) ? szf : szc;
      const int n = (threadIdx.x < 32) ? Sf : Sc;
      const int lane = threadIdx.x & 31;
      float carry = -INFINITY;
      for (int j0 = 0; j0 < n; j0 += 32) {
        float v = (j0 + lane < n) ? lst[j0 + lane] : -INFINITY;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float t = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v = fmaxf(v, t);
        }
        v = fmaxf(v, carry);
        if (j0 + lane < n) lst[j0 + lane] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      float z;
      int pos;
      if (i < Sf) {
        z = szf[i];
        int lo = 0, hi = Sc;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (szc[mid] < z) lo = mid + 1; else hi = mid; }
        pos = i + lo;
      } else {
        const int j = i - Sf;
        z = szc[j];
        int lo = 0, hi = Sf;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (szf[mid] <= z) lo = mid + 1; else hi = mid; }
        pos = j + lo;
      }
      uint32_t b = __float_as_uint(z);
      uint32_t asc = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
      keys[pos] = ((unsigned long long)asc << 32) | (unsigned)i;
    }
    __syncthreads();
  } else
  for (int i = threadIdx.x; i < P2; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < S) {
      float z = (i < Sf) ? zf[r * Sf + i] : zc[r * Sc + (i - Sf)];
      uint32_t b = __float_as_uint(z);
      uint32_t asc = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
      k = ((unsigned long long)asc << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= (presorted ? 0 : P2); k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x >= 32) return;
  const float4* rf = reinterpret_cast<const float4*>(raw_f) + r * Sf;
  const float4* rc = reinterpret_cast<const float4*>(raw_c) + r * Sc;
  auto fetch = [&](int j, float& zz, float& cr, float& cg, float& cb, float& sg) {
    unsigned long long k = keys[j];
    uint32_t asc = (uint32_t)(k >> 32);
    uint32_t b = (asc & 0x80000000u) ? (asc ^ 0x80000000u) : ~asc;
    zz = __uint_as_float(b);
    int pos = (int)(k & 0xffffffffu);
    float4 v = (pos < Sf) ? rf[pos] : rc[pos - Sf];
    cr = v.x; cg = v.y; cb = v.z; sg = v.w;
  };
  warp_composite(S, last_delta ? last_delta[r] : 1e10f, white_bkgd, fetch, wsm, wsm + S, rgb ? rgb + r * 3 : nullptr,
                 depth ? depth + r : nullptr, var ? var + r : nullptr, lam ? lam + r : nullptr, (float*)nullptr);
}

// ------------------------------------------------------------------------------------------
int composite_launch(const float* z, const float* raw, const float* last_delta, int64_t N, int S, int white_bkgd,
                     float* rgb, float* depth, float* var, float* lam, float* weights, cudaStream_t st) {
  if (N == 0) return SNB_OK;
  SNB_REQUIRE(S >= 1 && S <= 4096, "composite: samples per ray %d out of range", S);
  size_t smem = (size_t)4 * 2 * S * sizeof(float);
  if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_composite, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_composite<<<(unsigned)cdiv(N, 4), 128, smem, st>>>(z, raw, last_delta, N, S, white_bkgd, rgb, depth, var, lam, weights);
  SNB_CHECK_LAUNCH("k_composite");
  return SNB_OK;
}

int sample_pdf_launch(const float* bins, int ld_bins, const float* weights, int ld_w, int w_off, const float* u,
                      int64_t N, int nb, int nf, uint64_t seed, int det, float* zf, cudaStream_t st) {
  if (N == 0 || nf == 0) return SNB_OK;
  SNB_REQUIRE(nb >= 1 && nb <= 4096, "sample_pdf: bins %d out of range", nb);
  size_t smem = (size_t)4 * 2 * (nb + 1) * sizeof(float);
  if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_sample_pdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_sample_pdf<<<(unsigned)cdiv(N, 4), 128, smem, st>>>(bins, ld_bins, weights, ld_w, w_off, u, N, nb, nf, seed, det, zf);
  SNB_CHECK_LAUNCH("k_sample_pdf");
  return SNB_OK;
}

// ------------------------------------------------------------------------------------------
// camera rays of one image: ray_utils.py:6-84 (get_ray_directions + get_rays + _get_rays_inner +
// _truncate_with_plane_intersection), one thread per pixel.  rays [H*W, 8] = [o, d, near, far].
// ------------------------------------------------------------------------------------------
__global__ void k_get_rays(int W, int H, float fx, float fy, float cx, float cy, int center_pixels, const float* __restrict__ c2w,
                           float near, float far, int has_alt, float alt0, float alt1, float* __restrict__ rays) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (int64_t)W * H) return;
  const int i = (int)(p % W), j = (int)(p / W);                  // meshgrid(indexing='xy'): i = column, j = row
  const float fi = (float)i + (center_pixels ? 0.5f : 0.f), fj = (float)j + (center_pixels ? 0.5f : 0.f);
  float dx = __fdiv_rn(fi - cx, fx), dy = -__fdiv_rn(fj - cy, fy), dz = -1.f;      // ray_utils.py:14-15
  float n = sqrtf(dx * dx + dy * dy + dz * dz);
  dx = __fdiv_rn(dx, n); dy = __fdiv_rn(dy, n); dz = __fdiv_rn(dz, n);            // :16
  float r[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) r[a] = dx * c2w[a * 4 + 0] + dy * c2w[a * 4 + 1] + dz * c2w[a * 4 + 2];   // directions @ c2w[:, :3].T
  n = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
#pragma unroll
  for (int a = 0; a < 3; ++a) r[a] = __fdiv_rn(r[a], n);                                              // :25
  const float o[3] = {c2w[3], c2w[7], c2w[11]};
  float nb = near, fb = far;
  if (has_alt) {
    // _truncate_with_plane_intersection (:69-90): plane x = altitude, normal (-1, 0, 0); only for rays that start
    // before the plane and head towards it
    if (o[0] < alt0 && r[0] > 0.f) {
      const float w0 = o[0] - alt0, si = __fdiv_rn(w0, -r[0]);       // si = -w.n / (d.n) with n = (-1,0,0)
      const float q0 = w0 + si * r[0], q1 = o[1] + si * r[1], q2 = o[2] + si * r[2];   // plane_intersection - plane_point (+ plane_point)
      const float e0 = o[0] - (q0 + alt0), e1 = o[1] - q1, e2 = o[2] - q2;
      nb = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    }
    nb = fmaxf(nb, near);                                                                             // :52
    if (o[0] < alt1 && r[0] > 0.f) {
      const float w0 = o[0] - alt1, si = __fdiv_rn(w0, -r[0]);
      const float q0 = w0 + si * r[0], q1 = o[1] + si * r[1], q2 = o[2] + si * r[2];
      const float e0 = o[0] - (q0 + alt1), e1 = o[1] - q1, e2 = o[2] - q2;
      fb = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    }
    fb = fminf(fb, far);                                                                              // :55
    fb = fmaxf(nb, fb);                                                                               // :56
  }
  float* out = rays + p * 8;
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
  out[3] = r[0]; out[4] = r[1]; out[5] = r[2];
  out[6] = nb; out[7] = fb;
}

int get_rays_launch(int W, int H, float fx, float fy, float cx, float cy, int center_pixels, const float* c2w, float near,
                    float far, const float* alt, float* rays, cudaStream_t st) {
  if (W <= 0 || H <= 0) return SNB_OK;
  k_get_rays<<<(unsigned)cdiv((int64_t)W * H, 256), 256, 0, st>>>(W, H, fx, fy, cx, cy, center_pixels, c2w, near, far,
                                                                   alt != nullptr, alt ? alt[0] : 0.f, alt ? alt[1] : 0.f, rays);
  SNB_CHECK_LAUNCH("k_get_rays");
  return SNB_OK;
}

int coarse_z_launch(const float* rays, int64_t N, int Sc, float perturb, uint64_t seed, float* z, cudaStream_t st) {
  if (N == 0) return SNB_OK;
  k_coarse_z<<<(unsigned)cdiv(N * Sc, 256), 256, 0, st>>>(rays, N, Sc, perturb, seed, z);
  SNB_CHECK_LAUNCH("k_coarse_z");
  return SNB_OK;
}

int fill_x_launch(const float* rays, const int* image_indices, const float* z, int64_t N, int Sn, float* x,
                  cudaStream_t st) {
  if (N == 0 || Sn == 0) return SNB_OK;
  k_fill_x<<<(unsigned)cdiv(N * Sn, 256), 256, 0, st>>>(rays, image_indices, z, N, Sn, x);
  SNB_CHECK_LAUNCH("k_fill_x");
  return SNB_OK;
}

int zmid_launch(const float* z, int64_t N, int S, float* mid, cudaStream_t st) {
  if (N == 0 || S < 2) return SNB_OK;
  k_zmid<<<(unsigned)cdiv(N * (S - 1), 256), 256, 0, st>>>(z, N, S, mid);
  SNB_CHECK_LAUNCH("k_zmid");
  return SNB_OK;
}

// rendering.py:215-216 / 250-251: rays that continue into the background model carry last_delta = fg_far; the composite
// of a level uses last_delta - max(z of that level) for them (1e10 rays are left alone).  One warp per ray.
__global__ void __launch_bounds__(128) k_last_delta_adj(const float* __restrict__ last_delta, const float* __restrict__ z,
                                                        int64_t N, int S, float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= N) return;
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int j = lane; j < S; j += 32) mx = fmaxf(mx, z[r * S + j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    const float ld = last_delta[r];
    out[r] = (ld < 1e10f) ? (ld - mx) : ld;
  }
}
int last_delta_adj_launch(const float* last_delta, const float* z, int64_t N, int S, float* out, cudaStream_t st) {
  if (N == 0) return SNB_OK;
  k_last_delta_adj<<<(unsigned)cdiv(N, 4), 128, 0, st>>>(last_delta, z, N, S, out);
  SNB_CHECK_LAUNCH("k_last_delta_adj");
  return SNB_OK;
}

int merge_composite_launch(const float* zf, const float* zc, const float* raw_f, const float* raw_c,
                           const float* last_delta, int64_t N, int Sf, int Sc, int presorted, int white_bkgd,
                           float* rgb, float* depth, float* var, float* lam, cudaStream_t st) {
  if (N == 0) return SNB_OK;
  const int S = Sf + Sc;
  int P2 = 1;
  while (P2 < S) P2 <<= 1;
  SNB_REQUIRE(P2 <= 8192, "merge: %d samples per ray is too many", S);
  size_t smem = (size_t)P2 * 8 + (size_t)2 * S * sizeof(float);
  if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_merge_composite, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_merge_composite<<<(unsigned)N, presorted ? 128 : 256, smem, st>>>(zf, zc, raw_f, raw_c, last_delta, N, Sf, Sc, P2, presorted, white_bkgd, rgb, depth, var, lam);
  SNB_CHECK_LAUNCH("k_merge_composite");
  return SNB_OK;
}

}  // namespace snb

// ==========================================================================================
// mip renderer (reference rendering_mip.py): conical-frustum samples, composite on interval
// mid points with rgb padding, blurred-weight piecewise-constant resampling.
// ==========================================================================================
namespace snb {

__device__ __forceinline__ float linspace_to(int i, int n, float end) {   // ATen linspace(0, end, n)
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(end, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(end, __fmul_rn(step, (float)(n - 1 - i)));
}

// x[r*(Se-1) + j] = [mean(3), cov_diag(3), dir(3), image_index]   rendering_mip.py:15-25, 287, 333-338
__global__ void k_mip_fill_x(const float* __restrict__ rays, const float* __restrict__ radii,
                             const int* __restrict__ image_indices, const float* __restrict__ ze, int64_t N, int Se,
                             float* __restrict__ x) {
  const int Sn = Se - 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Sn) return;
  const int64_t r = i / Sn;
  const int j = (int)(i % Sn);
  const float* ray = rays + r * 8;
  const float t0 = ze[r * Se + j], t1 = ze[r * Se + j + 1];
  const float radius = radii[r];
  // same operation order as the reference expressions (no FMA contraction)
  const float c = __fdiv_rn(__fadd_rn(t0, t1), 2.f), d = __fdiv_rn(__fsub_rn(t1, t0), 2.f);
  const float c2 = __fmul_rn(c, c), d2 = __fmul_rn(d, d), d4 = __fmul_rn(d2, d2);
  const float den = __fadd_rn(__fmul_rn(3.f, c2), d2);
  const float t_mean = __fadd_rn(c, __fdiv_rn(__fmul_rn(__fmul_rn(2.f, c), d2), den));
  const float t_var = __fsub_rn(__fdiv_rn(d2, 3.f),
                                __fmul_rn(4.f / 15.f, __fdiv_rn(__fmul_rn(d4, __fsub_rn(__fmul_rn(12.f, c2), d2)), __fmul_rn(den, den))));
  const float r_var = __fmul_rn(__fmul_rn(radius, radius),
                                __fsub_rn(__fadd_rn(__fdiv_rn(c2, 4.f), __fmul_rn(5.f / 12.f, d2)),
                                          __fdiv_rn(__fmul_rn(4.f / 15.f, d4), den)));
  const float dx = ray[3], dy = ray[4], dz = ray[5];
  const float dd[3] = {__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz)};
  const float dn = __fadd_rn(__fadd_rn(dd[0], dd[1]), dd[2]);
  float* xr = x + i * 10;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    xr[a] = __fadd_rn(ray[a], __fmul_rn(ray[3 + a], t_mean));
    const float nod = __fsub_rn(1.f, __fdiv_rn(dd[a], dn));
    xr[3 + a] = __fadd_rn(__fmul_rn(t_var, dd[a]), __fmul_rn(r_var, nod));
    xr[6 + a] = ray[3 + a];
  }
  xr[9] = image_indices ? (float)image_indices[r] : 0.f;
}

// composite on mid points of the Se edges; raw [N, Se-1, 4]; rgb' = rgb*(1+2p) - p   rendering_mip.py:382-425
__global__ void __launch_bounds__(128) k_mip_composite(const float* __restrict__ ze, const float* __restrict__ raw,
                                                       const float* __restrict__ last_delta, int64_t N, int Se,
                                                       float rgb_padding, int white_bkgd, float* __restrict__ rgb,
                                                       float* __restrict__ depth, float* __restrict__ var,
                                                       float* __restrict__ weights) {
  extern __shared__ float sm[];  // [4 warps][2][Se-1]
  const int S = Se - 1;
  const int w = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 4 + w;
  if (r >= N) return;
  float* ws = sm + (size_t)w * 2 * S;
  const float* zr = ze + r * Se;
  const float4* rr = reinterpret_cast<const float4*>(raw) + r * S;
  const float a = 1.f + 2.f * rgb_padding;
  auto fetch = [&](int j, float& zz, float& cr, float& cg, float& cb, float& sg) {
    zz = __fmul_rn(0.5f, __fadd_rn(zr[j + 1], zr[j]));
    float4 v = rr[j];
    cr = __fsub_rn(__fmul_rn(v.x, a), rgb_padding);
    cg = __fsub_rn(__fmul_rn(v.y, a), rgb_padding);
    cb = __fsub_rn(__fmul_rn(v.z, a), rgb_padding);
    sg = v.w;
  };
  warp_composite(S, last_delta ? last_delta[r] : 1e10f, white_bkgd, fetch, ws, ws + S, rgb ? rgb + r * 3 : nullptr,
                 depth ? depth + r : nullptr, var ? var + r : nullptr, (float*)nullptr,
                 weights ? weights + r * S : nullptr);
}

// weights [N, n] (n = Se-1 intervals of the Se edges) -> blurred + padded pdf -> nf sorted samples
// rendering_mip.py:217-226 + sorted_piecewise_constant_pdf1 (75-131).  randomized = 0: u = linspace(0, 1 - eps, nf);
// randomized = 1 (:97-105): stratified u_j = j/nf + rand * (1/nf - eps), capped at 1 - eps (ascending, so the samples
// come out sorted as the reference's sort leaves them).  One warp per ray.
__global__ void __launch_bounds__(128) k_mip_resample(const float* __restrict__ ze, const float* __restrict__ weights,
                                                      int64_t N, int Se, int nf, float resample_padding, int randomized,
                                                      uint64_t seed, float* __restrict__ zf) {
  extern __shared__ float sm[];  // [4 warps][(n) wprime + (n+1) cdf + (n+1) edges]
  const int n = Se - 1;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + w;
  if (r >= N) return;
  float* wp = sm + (size_t)w * (3 * n + 2);
  float* cdf = wp + n;
  float* eb = cdf + n + 1;
  const float* wr = weights + r * n;
  for (int j = lane; j <= n; j += 32) eb[j] = ze[r * Se + j];
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) {
    // weights_pad = [w0, w, w_last]; max of neighbours; blur = mean of consecutive maxima
    const float wm1 = wr[max(j - 1, 0)], w0 = wr[j], wp1 = wr[min(j + 1, n - 1)];
    const float m0 = fmaxf(wm1, w0), m1 = fmaxf(w0, wp1);
    const float v = __fadd_rn(__fmul_rn(0.5f, __fadd_rn(m0, m1)), resample_padding);
    wp[j] = v;
    sum += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float pad = fmaxf(0.f, 1e-5f - sum);
  const float wsum = sum + pad;
  __syncwarp();
  if (lane == 0) {
    double run = 0.0;                       // torch CPU cumsum: sequential, double accumulator
    cdf[0] = 0.f;
    for (int j = 0; j < n - 1; ++j) {
      const float pdf = __fdiv_rn(__fadd_rn(wp[j], __fdiv_rn(pad, (float)n)), wsum);
      run += (double)pdf;
      cdf[j + 1] = fminf(1.f, (float)run);
    }
    cdf[n] = 1.f;
  }
  __syncwarp();
  const float end = 1.f - 1.1920928955078125e-07f;   // 1 - finfo(float32).eps
  for (int j = lane; j < nf; j += 32) {
    float u;
    if (randomized) {
      const float sN = 1.f / (float)nf;
      u = fminf(__fadd_rn((float)j * sN, __fmul_rn(u01(seed ^ 0x2545F4914F6CDD1Dull, (uint64_t)r, (uint64_t)j), sN - 1.1920928955078125e-07f)), end);
    } else {
      u = linspace_to(j, nf, end);
    }
    int lo = 0, hi = n + 1;                  // first index with cdf > u
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    const int i0 = min(max(lo - 1, 0), n), i1 = min(lo, n);
    const float c0 = cdf[i0], c1 = cdf[i1];
    float t = __fdiv_rn(u - c0, c1 - c0);
    if (!(t == t) ) t = 0.f;                 // nan_to_num(nan -> 0); +-inf are clipped below
    t = fminf(fmaxf(t, 0.f), 1.f);
    zf[r * nf + j] = __fadd_rn(eb[i0], __fmul_rn(t, eb[i1] - eb[i0]));
  }
}

int mip_fill_x_launch(const float* rays, const float* radii, const int* image_indices, const float* ze, int64_t N,
                      int Se, float* x, cudaStream_t st) {
  if (N == 0 || Se < 2) return SNB_OK;
  k_mip_fill_x<<<(unsigned)cdiv(N * (Se - 1), 256), 256, 0, st>>>(rays, radii, image_indices, ze, N, Se, x);
  SNB_CHECK_LAUNCH("k_mip_fill_x");
  return SNB_OK;
}

int mip_composite_launch(const float* ze, const float* raw, const float* last_delta, int64_t N, int Se,
                         float rgb_padding, int white_bkgd, float* rgb, float* depth, float* var, float* weights,
                         cudaStream_t st) {
  if (N == 0) return SNB_OK;
  SNB_REQUIRE(Se >= 2 && Se <= 4097, "mip composite: %d edges out of range", Se);
  size_t smem = (size_t)4 * 2 * (Se - 1) * sizeof(float);
  if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_mip_composite, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mip_composite<<<(unsigned)cdiv(N, 4), 128, smem, st>>>(ze, raw, last_delta, N, Se, rgb_padding, white_bkgd, rgb, depth, var, weights);
  SNB_CHECK_LAUNCH("k_mip_composite");
  return SNB_OK;
}

int mip_resample_launch(const float* ze, const float* weights, int64_t N, int Se, int nf, float resample_padding,
                        int randomized, uint64_t seed, float* zf, cudaStream_t st) {
  if (N == 0 || nf == 0) return SNB_OK;
  size_t smem = (size_t)4 * (3 * (Se - 1) + 2) * sizeof(float);
  if (smem > 48 * 1024) SNB_CHECK_CUDA(cudaFuncSetAttribute(k_mip_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mip_resample<<<(unsigned)cdiv(N, 4), 128, smem, st>>>(ze, weights, N, Se, nf, resample_padding, randomized, seed, zf);
  SNB_CHECK_LAUNCH("k_mip_resample");
  return SNB_OK;
}

}  // namespace snb

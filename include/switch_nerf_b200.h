/*
 * switch_nerf_b200 -- C ABI of the B200-native Switch-NeRF forward/render hot path.
 *
 * Every entry point takes plain device pointers + sizes + a CUDA stream
 * (`void*` == cudaStream_t), enqueues work on that stream, never synchronises
 * the host, never allocates or frees caller memory (scratch comes from a
 * caller-provided workspace sized by snb_workspace_bytes), and returns an int
 * status: 0 = ok, otherwise an SNB_E* code whose text is available through
 * snb_last_error() (thread-local).  There are NO torch types here: the Python
 * mirror (switch_nerf_b200/*.py) binds these symbols with ctypes exactly as a
 * maintainer of the reference would.
 *
 * Each function cites the reference interface it replaces; paths are relative
 * to /root/reference/switch_nerf/.
 */
#ifndef SWITCH_NERF_B200_H_
#define SWITCH_NERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_OK 0
#define SNB_EINVAL 1      /* bad argument (shape, NULL pointer, unsupported topology) */
#define SNB_ECUDA 2       /* a CUDA runtime call or launch failed */
#define SNB_EWORKSPACE 3  /* workspace too small */
#define SNB_EUNSUPPORTED 4 /* precision / topology not supported by the selected path */

/* Compute precision of the model forward.
 *  SNB_PREC_FP32 : every GEMM in fp32 on the CUDA cores (what the reference does
 *                  without autocast; used for BASELINE.json configs[0] parity).
 *  SNB_PREC_BF16 : the autocast(bf16) map of the reference (README.md:77
 *                  --amp_use_bfloat16): bf16 operands, fp32 accumulation in TMEM
 *                  (tcgen05), fp32 LayerNorm / gate / softmax / softplus /
 *                  compositing. */
#define SNB_PREC_FP32 0
#define SNB_PREC_BF16 1

typedef struct snb_model snb_model_t; /* opaque: packed weights + tile tables (library-owned) */

/* Topology of one NeRFMoE / MipNeRFMoE network: the `model:` block of
 * configs/switch_nerf/{building,mission_bay}.yaml + opts.py defaults. */
typedef struct snb_model_desc {
  int32_t num_experts;      /* E: layers."0" local_expert_num (opts.py moe_local_expert_num)   */
  int32_t width;            /* M: layers."0".in_ch (256 Building, 512 Mission Bay)             */
  int32_t expert_layers;    /* layers."0".num (7)                                              */
  int32_t skip_layer;       /* layers."0".skips[0] (3), -1 = none                              */
  int32_t gate_layers;      /* layers.moe_external_gate.num (2)                                */
  int32_t pos_xyz_freqs;    /* hparams.pos_xyz_dim (12)                                        */
  int32_t pos_dir_freqs;    /* hparams.pos_dir_dim (4)                                         */
  int32_t appearance_dim;   /* hparams.appearance_dim (48)                                     */
  int32_t appearance_count; /* rows of embedding_a                                             */
  int32_t hidden2;          /* layers."2".out_ch (128)                                         */
  int32_t mip;              /* 0: NeRFMoE (x = [xyz3,dir3,idx1]); 1: MipNeRFMoE (x = [mean3,cov3,dir3,idx1]) */
} snb_model_desc;

/* Device pointers to fp32 weights in the reference's state_dict layout
 * (names from a live models/nerf_moe.py:103-310 instance). */
typedef struct snb_weights {
  const float* xyz_w;        /* layers.xyz.fcs.0.weight              [M, 3+6*pos_xyz_freqs]    */
  const float* xyz_b;        /* layers.xyz.fcs.0.bias                [M]                       */
  const float* gate_w[4];    /* layers.moe_external_gate.fcs.{i}.weight [M, M]                 */
  const float* gate_b[4];    /* layers.moe_external_gate.fcs.{i}.bias   [M]                    */
  const float* ln_w;         /* layers.gate_input_norm.weight        [M]                       */
  const float* ln_b;         /* layers.gate_input_norm.bias          [M]                       */
  const float* wg;           /* layers.0.gates.0.wg.weight           [E, M] (no bias)          */
  const float* expert_w[16]; /* layers.0.experts.0.weights.{j}       [E, M(in), M(out)]        */
  const float* expert_b[16]; /* layers.0.experts.0.bias.{j}          [E, 1, M]                 */
  const float* l1_w;         /* layers.1.fcs.0.weight                [M, M]                    */
  const float* l1_b;         /* layers.1.fcs.0.bias                  [M]                       */
  const float* l2_w;         /* layers.2.fcs.0.weight                [H2, M+3+6*pos_dir_freqs+appearance_dim] */
  const float* l2_b;         /* layers.2.fcs.0.bias                  [H2]                      */
  const float* sigma_w;      /* layers.sigma.fcs.0.weight            [1, M]                    */
  const float* sigma_b;      /* layers.sigma.fcs.0.bias              [1]                       */
  const float* color_w;      /* layers.color.fcs.0.weight            [3, H2]                   */
  const float* color_b;      /* layers.color.fcs.0.bias              [3]                       */
  const float* emb_a;        /* embedding_a.weight                   [appearance_count, appearance_dim] */
} snb_weights;

/* Routing options of one MoE forward: TopKGate ctor args
 * (modules/tutel_moe_ext/tutel_moe_layer_nobatch.py:33-96). */
typedef struct snb_route_opts {
  double capacity_factor; /* hparams.moe_capacity_factor (>0); double like the Python float it mirrors */
  int32_t bpr;            /* batch_prioritized_routing                                          */
  int32_t no_batch;       /* MOELayer.moe_no_batch: 1 = capacity-free eval mode (nothing dropped) */
} snb_route_opts;

/* ---- library ---------------------------------------------------------------------------- */
const char* snb_last_error(void);
int snb_version(void);
/* Number of kernels this library has launched in this process (all streams); bench.py reports the
 * delta over its timed region as `gpu_launches`. */
int64_t snb_launch_count(void);
/* Phase timing of snb_moe_forward (SNB_PREC_BF16): when enabled, CUDA events are recorded on the
 * caller's stream around launch #1 (front), the routing kernels and launch #2 (back); no host sync.
 * snb_profile_collect synchronises the device and returns the sums since the last enable/collect:
 * out[0..3] = {front_ms, route_ms, back_ms, n_chunks}.  Used by bench.py for the roofline line. */
int snb_profile_enable(int32_t on);
int snb_profile_collect(double* out4);
/* Debug only (env SNB_TIMELINE=1): clock marks of CTA 0 of the last fused forward; returns the number of
 * 64-bit words copied (0 when disabled).  Synchronises the device. */
int snb_debug_timeline(uint64_t* host_out, int32_t n);

/* ---- model object -----------------------------------------------------------------------
 * Concurrency: every entry point only enqueues on the caller's stream, but a model object owns the side streams /
 * events of its chunk pipeline and the zero-between-uses scratch of the routing kernel, so ONE caller stream may use a
 * given model at a time (serialise host threads that share a model; different models are independent).
 * The only process-global state left is debug tooling: the launch counter, the phase-event pool of snb_profile_enable and
 * the SNB_TIMELINE buffer -- not thread-safe, not meant for production. */
/* Replaces models/nerf_moe.py:1004-1041 get_nerf_moe_inner + load_state_dict: copies/packs the
 * caller's fp32 weights into kernel-native layouts (fp32 [N,K] for the CUDA-core path; bf16
 * UMMA canonical K-major core-matrix tiles for the tcgen05 path).  Call again after an
 * optimizer step (snb_model_update) -- packing is one pass over ~16 MB. */
int snb_model_create(const snb_model_desc* desc, const snb_weights* w, void* stream, snb_model_t** out);
int snb_model_update(snb_model_t* m, const snb_weights* w, void* stream);

/* Kernel-selection / pipeline knobs of a model object (A/B switches and the tuning found in profiles/).  A model starts
 * with the defaults below, overridden once at creation by the SNB_* environment variables named in the comments (kept for
 * the A/B scripts); snb_model_set_tuning changes them afterwards, and every forward call reads the model's copy -- no
 * process-global state.  Fields <= -1 mean "default".  (gather_h needs the weights re-packed: call snb_model_update.) */
typedef struct snb_tuning {
  int32_t cta_group_front;  /* SNB_CG / SNB_CG_FRONT: 1 (default) or 2 = cta_group::2 CTA pairs for launch #1            */
  int32_t cta_group_back;   /* SNB_CG / SNB_CG_BACK: the same for launch #2                                              */
  int32_t ts;               /* SNB_TS: 1 (default) = hidden activations in tensor memory, 0 = shared-memory A operand     */
  int32_t ts_front;         /* SNB_TS_FRONT: the same for launch #1 only                                                  */
  int32_t wide;             /* SNB_WIDE: 1 = force the wide kernels at width 256 (always used for width 512 / mip)        */
  int32_t route_full;       /* SNB_ROUTE_FULL: 1 = full-order routing (route_top1 + tile plan) instead of k_select        */
  int32_t no_overlap;       /* SNB_NO_OVERLAP: 1 = routing on the caller's stream (no side stream / SM partition)         */
  int32_t pipe_depth;       /* SNB_PIPE_DEPTH: launch #1 of how many chunks run ahead of launch #2 (default 2, EP 3)      */
  int32_t route_sms;        /* SNB_ROUTE_SMS: SMs launch #1 leaves free for the routing kernels (default 8)               */
  int32_t back_partition;   /* SNB_BACK_PART: 1 = launch #2 also leaves them free                                         */
  int32_t no_ray_source;    /* SNB_NO_RAY_SOURCE: 1 = the render path materialises the [N*S,7] point tensor               */
  int32_t gather_h;         /* SNB_GATHER_H: 1 = launch #2 gathers h from HBM instead of recomputing it (read when packing) */
  int32_t front_ab;         /* SNB_FRONT_AB: debug bit mask for launch #1 (1 = no key histogram, 2 = no column sums)      */
} snb_tuning;
int snb_model_get_tuning(const snb_model_t* m, snb_tuning* out);
int snb_model_set_tuning(snb_model_t* m, const snb_tuning* t);
void snb_model_destroy(snb_model_t* m);

/* Scratch bytes needed by snb_moe_forward / snb_render_rays for chunks of up to
 * `max_chunk_samples` rows and capacity factors up to `max_capacity_factor`. */
size_t snb_workspace_bytes(const snb_model_t* m, int64_t max_chunk_samples, double max_capacity_factor);

/* ---- (e) expert parallelism: experts sharded E/W per GPU, one process per GPU ------------------ */
/* Replaces the two all_to_all_single calls around the experts (tutel_moe_layer_nobatch.py:164-218 via
 * tutel_communicate_nobatch.py:15-54) and the `group`/`moe_local_expert_num` plumbing of
 * moe_layer (tutel_moe_layer_nobatch.py:443-460, runner.py:100-101).  Each rank routes its own chunk
 * (capacity from the local S and the global E, as the reference does); the exchange is done by the routing
 * stage and by launch #2 themselves with st.global into peer memory + per-peer release/acquire flags --
 * no NCCL call, no host synchronisation.  Because every non-expert weight is replicated, a sample travels
 * as its 48-byte launch-#2 input record and comes back as its 16-byte {rgb, sigma} row (the reference ships
 * two 512-byte activation rows).  Results are bit-identical to the single-GPU path.
 *
 *   snb_a2a_init        allocate this rank's symmetric region (device memory of the current device), sized
 *                       for chunks of up to max_chunk_rows rows and capacity factors up to max_capacity_factor
 *   snb_a2a_export      64-byte CUDA IPC handle of the region; exchange them with any host-side all-gather
 *                       (torch.distributed.all_gather_object, MPI, files) ...
 *   snb_a2a_connect     ... and map every peer (handles = world x 64 bytes, rank order).  A host barrier
 *                       between snb_a2a_connect on all ranks and the first forward is the caller's job.
 *   snb_a2a_connect_ptrs same-process peers / an external symmetric heap: base pointers, rank order
 *                       (peer access must already be enabled)
 *   snb_model_attach_a2a every following bf16 snb_moe_forward / snb_render_rays of the model runs
 *                       expert-parallel: experts [rank*E/W, (rank+1)*E/W) are evaluated here for all ranks
 *                       (NULL detaches).  All ranks must make the same sequence of model-chunk calls
 *                       (the reference's all_to_all has the same requirement); a peer that never arrives
 *                       traps the waiting kernel after 20 s.
 *   snb_a2a_disconnect  unmap the peers (device-synchronising).  CUDA requires every importer to unmap a
 *                       region before its owner frees it: disconnect on all ranks, host barrier, finalize.
 *   snb_a2a_finalize    disconnect if still connected, free the region.                                     */
typedef struct snb_a2a snb_a2a_t;
int snb_a2a_init(int32_t rank, int32_t world, int32_t num_experts, int64_t max_chunk_rows,
                 double max_capacity_factor, snb_a2a_t** out);
int32_t snb_a2a_handle_bytes(void);
int snb_a2a_export(snb_a2a_t* g, void* handle_out);
int snb_a2a_connect(snb_a2a_t* g, const void* handles, size_t handles_bytes);
int snb_a2a_connect_ptrs(snb_a2a_t* g, void* const* bases, int32_t n);
void* snb_a2a_local_base(snb_a2a_t* g);
size_t snb_a2a_region_bytes(const snb_a2a_t* g);
int snb_model_attach_a2a(snb_model_t* m, snb_a2a_t* g);
int snb_a2a_disconnect(snb_a2a_t* g);
int snb_a2a_finalize(snb_a2a_t* g);

/* ---- a9: routing ------------------------------------------------------------------------ */
/* Replaces extract_critical (tutel_fast_dispatch.py:176-217, k=1) incl. one_hot (131-134),
 * compute_sorted_location (136-139), load_balance (141-150) and Tutel's fast_cumsum_sub_one:
 *   idx[s]  = argmax_e gates[s,e]  (lowest e wins exact ties)
 *   gate[s] = gates[s, idx[s]]
 *   loc[s]  = #{s' : idx[s']==idx[s], s' before s}  where "before" is ascending s (bpr=0) or
 *             descending max-gate with ties by ascending s (bpr=1)
 *   counts[e] = #{s : idx[s]==e};  *capacity = int(cf * ceil(S/E));  *l_aux = E/S^2 * sum_e me_e*ce_e
 * gates fp32 [S,E] row-major; idx/loc int32 [S]; gate fp32 [S]; counts int32 [E];
 * capacity int32 [1] and l_aux fp32 [1] are DEVICE scalars (no host round trip).
 * workspace: snb_route_workspace_bytes(S, E). */
size_t snb_route_workspace_bytes(int64_t S, int32_t E);
int snb_route_top1(const float* gates, int64_t S, int32_t E, double capacity_factor, int32_t bpr,
                   int32_t* idx, int32_t* loc, float* gate, int32_t* counts, int32_t* capacity,
                   float* l_aux, void* workspace, size_t workspace_bytes, void* stream);

/* Routing as the fused path runs it (one launch, E CTAs): the same top-1 decision as snb_route_top1, but only the
 * kept SET per expert is computed -- {s : locations_s < capacity} of tutel_fast_dispatch.py:136-139, 176-217 -- not
 * the batch-prioritised order inside it.  idx, gate, counts, capacity, l_aux as snb_route_top1;
 *   loc[s] <  capacity : kept; the sample's rank among the kept samples of its expert in SAMPLE-INDEX order
 *   loc[s] >= capacity : dropped (capacity + a running number)
 * so (loc < capacity) equals the reference's kept mask bit for bit (ties between equal gates go to the lower sample
 * index, as with a stable argsort), while the value of a kept loc is a different -- equally valid -- slot numbering.
 * no_batch = 1: nothing is dropped (extract_critical_nobatch, tutel_fast_dispatch_nobatch.py).  E <= 16.
 * workspace: snb_route_select_workspace_bytes(S). */
size_t snb_route_select_workspace_bytes(int64_t S);
int snb_route_select(const float* gates, int64_t S, int32_t E, double capacity_factor, int32_t bpr,
                     int32_t no_batch, int32_t* idx, int32_t* loc, float* gate, int32_t* counts,
                     int32_t* capacity, float* l_aux, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a10/a12/a13: dispatch + combine (Tutel K4/K5, in-tree K1/K2) ----------------------- */
/* Replaces GatingEncoder.forward (tutel_fast_dispatch.py:15-28; _nobatch.py:16-37):
 *   out = zeros[rows_out, H];  out[row(s)] = x[s]  for kept samples
 *   row(s) = idx[s]*capacity + loc[s], kept iff loc[s] < capacity          (begin == NULL)
 *   row(s) = begin[idx[s]] + loc[s],  always kept                           (begin != NULL)
 * and GatingDecoder.forward (48-63): y[s] = gate[s]*buf[row(s)] if kept else 0. */
int snb_dispatch_fwd(const float* x, const int32_t* idx, const int32_t* loc, const int32_t* begin,
                     int64_t S, int32_t H, int32_t capacity, int64_t rows_out, float* out, void* stream);
int snb_combine(const float* buf, const int32_t* idx, const int32_t* loc, const int32_t* begin,
                const float* gate, int64_t S, int32_t H, int32_t capacity, int64_t rows_buf,
                float* y, void* stream);

/* ---- a4..a14: one model_chunk of NeRFMoE.forward / MipNeRFMoE.forward -------------------- */
/* Replaces `nerf(x, sigma_noise=...)` (models/nerf_moe.py:320-455 / 675-810) for the
 * Building / Mission-Bay topology:
 *   x [S, 7] (mip: [S,10]) fp32 = [xyz(3) (cov_diag(3)), dir(3), image_index(1 as float)]
 *   out [S,4] fp32 = [sigmoid(rgb)(3), softplus(sigma-1)(1)]
 *   moe_idx (nullable) int32 [S] = extras["moe_gates"][0][:,0]; l_aux fp32 device scalar =
 *   extras["moe_loss"][0].  dbg_* are optional taps for parity tests (may be NULL):
 *   dbg_gates fp32 [S,E], dbg_loc int32 [S]. */
int snb_moe_forward(snb_model_t* m, const float* x, int64_t S, const float* sigma_noise,
                    const snb_route_opts* opts, int32_t precision, float* out, int32_t* moe_idx,
                    float* l_aux, float* dbg_gates, int32_t* dbg_loc, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---- f1: backward ------------------------------------------------------------------------- */
/* Gradient buffers, fp32 device pointers in the reference's state_dict layouts (same fields as snb_weights; experts
 * [E, in, out] / [E, 1, out]); every pointer nullable (that gradient is skipped).  Gradients are ACCUMULATED. */
typedef struct snb_grads {
  float* xyz_w; float* xyz_b;
  float* gate_w[4]; float* gate_b[4];
  float* ln_w; float* ln_b;
  float* wg;
  float* exp_w[16]; float* exp_b[16];
  float* l1_w; float* l1_b; float* l2_w; float* l2_b;
  float* sigma_w; float* sigma_b; float* color_w; float* color_b;
  float* emb_a;
} snb_grads;

/* Parameter gradients of one model chunk: backward of snb_moe_forward (NeRFMoE.forward, models/nerf_moe.py:320-455, with
 * GatingEncoder/Decoder.backward of tutel_fast_dispatch.py:30-45, 65-78 and the ExpertMLP dgrad / wgrad,
 * tutel_moe_layer_nobatch.py:887-924; the caller is runner.py:677-690).  fp32, capacity (batched) dispatch; the forward
 * intermediates are recomputed inside.  d_out [S,4] = dL/d[rgb, sigma]; d_l_aux: DEVICE scalar dL/d(l_aux of this
 * chunk) or NULL.  Nothing flows into x. */
size_t snb_moe_backward_workspace_bytes(const snb_model_t* m, int64_t S, double capacity_factor);
int snb_moe_backward(snb_model_t* m, const float* x, int64_t S, const float* sigma_noise, const snb_route_opts* opts,
                     const float* d_out, const float* d_l_aux, const snb_grads* grads, void* workspace,
                     size_t workspace_bytes, void* stream);
/* Backward of the volumetric composite (rendering.py:436-494; depth / variance are computed under no_grad there):
 * z [N,S] sorted depths, raw [N,S,4] = [rgb, sigma] in the same order, last_delta [N] or NULL (1e10), d_rgb [N,3];
 * writes d_raw [N,S,4]. */
int snb_composite_backward(const float* z, const float* raw, const float* last_delta, int64_t n_rays, int32_t n_samples,
                           const float* d_rgb, float* d_raw, void* stream);

/* ---- f3: background NeRF + sphere parametrisation ------------------------------------------ */
/* The background model of the reference: models/nerf.py:75-191 `NeRF` with xyz_dim = 4 (a point on the unit sphere +
 * inverse distance), `layers` ReLU layers of `width` with the encoded input concatenated again at `skip_layer`,
 * xyz_encoding_final, dir_a_encoding (width + dir + appearance -> width/2, ReLU), sigma / rgb heads.
 * x [S, 8] = [pts(4), dir(3), image index]; out [S,4] = [sigmoid(rgb), sigma_activation(sigma (+noise))]. */
typedef struct snb_bg_desc {
  int32_t layers;           /* hparams.layers (8)                 */
  int32_t skip_layer;       /* hparams.skip_layers[0] (4), -1 none */
  int32_t width;            /* hparams.bg_layer_dim (256)          */
  int32_t pos_xyz_freqs;    /* hparams.pos_xyz_dim (12)            */
  int32_t pos_dir_freqs;    /* hparams.pos_dir_dim (4)             */
  int32_t appearance_dim;   /* hparams.appearance_dim (48)         */
  int32_t appearance_count;
  int32_t shifted_softplus; /* sigma activation: 1 = softplus(x - 1) (models/nerf.py:58-72), 0 = ReLU */
} snb_bg_desc;
typedef struct snb_bg_weights {
  const float* w[16];       /* xyz_encodings.{i}.0.weight [width, in_i]; in_0 = 4 + 8*freqs, in_skip = in_0 + width (encoded input first) */
  const float* b[16];
  const float* final_w; const float* final_b;    /* xyz_encoding_final   [width, width]            */
  const float* dir_w; const float* dir_b;        /* dir_a_encoding.0     [width/2, width + dir + A] */
  const float* sigma_w; const float* sigma_b;    /* sigma                [1, width]                */
  const float* rgb_w; const float* rgb_b;        /* rgb                  [3, width/2]              */
  const float* emb_a;                            /* embedding_a.weight   [count, A]                */
} snb_bg_weights;
typedef struct snb_bg_model snb_bg_model_t;
int snb_bg_create(const snb_bg_desc* desc, const snb_bg_weights* w, void* stream, snb_bg_model_t** out);
int snb_bg_update(snb_bg_model_t* m, const snb_bg_weights* w, void* stream);
void snb_bg_destroy(snb_bg_model_t* m);
size_t snb_bg_workspace_bytes(const snb_bg_model_t* m, int64_t S);
int snb_bg_forward(snb_bg_model_t* m, const float* x, int64_t S, const float* sigma_noise, float* out, void* workspace,
                   size_t workspace_bytes, void* stream);
/* rendering.py:497-518 `_intersect_sphere`: fg_far [N] = depth at which a ray leaves the (unit, after centre / radius)
 * sphere; *bad (DEVICE int, nullable) is set to 1 if a ray's closest point to the centre lies outside the sphere (the
 * reference raises 'Not all your cameras are bounded by the unit sphere'). */
int snb_intersect_sphere(const float* rays, int64_t N, const float* sphere_center, const float* sphere_radius,
                         float* fg_far, int32_t* bad, void* stream);
/* rendering.py:521-570 `_depth2pts_outside` (include_xyz_real = False): for ray r and inverse depth z[r, j] in [0,1] the
 * 4-D background point [p_sphere_rotated(3), z] and the conventional depth depth_real.  sphere_center / sphere_radius:
 * DEVICE [3] each (nullable together). */
int snb_depth2pts_outside(const float* rays, const float* sphere_center, const float* sphere_radius, const float* z,
                          int64_t N, int32_t S, float* pts, float* depth_real, void* stream);

/* ---- f2: ray generation ------------------------------------------------------------------- */
/* Replaces ray_utils.get_ray_directions + get_rays (ray_utils.py:6-84) for one image: rays [H*W, 8] fp32 =
 * [origin(3), unit direction(3), near, far], pixel (row j, column i) at index j*W + i.
 * c2w: DEVICE pointer to the 3x4 camera-to-world matrix (row-major); altitude_range: HOST pointer to
 * [max_altitude, min_altitude] (hparams.ray_altitude_range) or NULL. */
int snb_get_rays(int32_t W, int32_t H, float fx, float fy, float cx, float cy, int32_t center_pixels,
                 const float* c2w, float near, float far, const float* altitude_range, float* rays, void* stream);

/* ---- a1..a3: rendering.render_rays ------------------------------------------------------- */
typedef struct snb_render_opts {
  int32_t coarse_samples;   /* hparams.coarse_samples                                          */
  int32_t fine_samples;     /* hparams.fine_samples (0 = coarse only)                          */
  int64_t model_chunk_size; /* hparams.model_chunk_size: routing/capacity are per chunk (F7)   */
  float perturb;            /* hparams.perturb if nerf.training else 0                         */
  uint64_t seed;            /* Philox seed for perturb / stochastic pdf sampling               */
  int32_t white_bkgd;       /* hparams.white_bkgd                                              */
  int32_t precision;        /* SNB_PREC_*                                                      */
  snb_route_opts route;
  /* rendering.py:316-322 (`sigma_noise = randn * sigma_noise_std` per chunk when hparams.use_sigma_noise and
   * nerf.training): caller-drawn noise added to the raw sigma before the softplus, one value per point-sample
   * in ray-major order; device pointers, nullable (= no noise). */
  const float* sigma_noise_coarse; /* [N * coarse_samples] */
  const float* sigma_noise_fine;   /* [N * fine_samples]   */
  /* mip renderer only: rendering_mip.py:225 passes `randomized=hparams.perturb` to the fine resampling -- in eval too
   * (the coarse perturb above is training-only) -- so it is a separate switch: 1 = stratified random u (seeded) */
  int32_t resample_randomized;
  /* bg-NeRF rays (rendering.py:41-46, 215-216, 250-251): last_delta [N] holds the reference's raw value (fg_far for rays
   * that continue into the background, 1e10 otherwise) and the composite of each level uses last_delta - max(z of that
   * level) where last_delta < 1e10 */
  int32_t last_delta_minus_zmax;
} snb_render_opts;

/* Per-ray outputs (all nullable; device pointers). Keys of the reference `results` dict
 * (rendering.py:466-494; runner.py:1089-1121). */
typedef struct snb_render_out {
  float* rgb;            /* rgb_{fine|coarse}            [N,3] */
  float* depth;          /* depth_*                      [N]   */
  float* depth_variance; /* depth_variance_*             [N]   */
  float* bg_lambda;      /* bg_lambda_* = T[...,-1]      [N]   */
  float* gate_loss_coarse; /* [ceil(N*coarse/chunk)]           */
  float* gate_loss_fine;   /* [ceil(N*fine/chunk)]             */
  int32_t* moe_gates_coarse; /* [N, coarse] expert ids          */
  int32_t* moe_gates_fine;   /* [N, fine]                       */
  float* z_fine;         /* tap: fine z values           [N, fine]      */
  float* raw_coarse;     /* tap: per-sample [rgb,sigma]  [N, coarse, 4] */
  float* raw_fine;       /* tap: per-sample [rgb,sigma]  [N, fine, 4]   */
  float* rgb_coarse;     /* mip renderer only: rgb_coarse [N,3] (rendering_mip.py composites both levels) */
  float* z_coarse;       /* tap: coarse z values (after perturb) [N, coarse]; the backward needs them */
} snb_render_out;

size_t snb_render_workspace_bytes(const snb_model_t* m, int64_t n_rays, const snb_render_opts* o);

/* Replaces rendering.render_rays (rendering.py:15-196) with bg_nerf=None, use_cascade=False:
 * coarse z (linspace + perturb, :85-88, 573-584), points o+d*z (:90), chunked model calls
 * (:354-383), alpha/transmittance/weights (:436-461), _sample_pdf (:587-637), fine pass,
 * sorted merge of coarse+fine (:419-429) and the composite (:466-494).
 * rays [N,8] fp32 = (o3,d3,near,far); image_indices int32 [N]; last_delta nullable [N]
 * (defaults to 1e10, rendering.py:33). */
int snb_render_rays(snb_model_t* m, const float* rays, const int32_t* image_indices,
                    const float* last_delta, int64_t n_rays, const snb_render_opts* opts,
                    const snb_render_out* out, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces rendering_mip.render_rays (rendering_mip.py:133-174, 177-261, 264-425) for MipNeRFMoE models
 * (snb_model_desc.mip = 1), eval mode (perturb = 0 -> deterministic resampling):
 * coarse edges linspace(near, far, coarse_samples) -> mip_cast_rays (15-25) per interval -> model chunks ->
 * composite on interval mid points with rgb padding (382-425) -> blurred weights + padding (217-224) ->
 * sorted_piecewise_constant_pdf1 (75-131; the O(N*S^2) mask replaced by a binary search over the same cdf) ->
 * fine edges -> second model pass -> composite.  coarse_samples / fine_samples count EDGES (hparams values),
 * i.e. coarse_samples-1 and fine_samples-1 network evaluations per ray.  radii [N] (or [N,1]).
 * rgb_padding < 0 means hparams.rgb_padding = None.  out->rgb/depth/depth_variance are the fine level (coarse
 * when fine_samples == 0), out->rgb_coarse the coarse composite; moe_gates_* are [N, samples-1]. */
size_t snb_render_mip_workspace_bytes(const snb_model_t* m, int64_t n_rays, const snb_render_opts* o);
int snb_render_rays_mip(snb_model_t* m, const float* rays, const float* radii, const int32_t* image_indices,
                        const float* last_delta, int64_t n_rays, const snb_render_opts* opts,
                        float weights_resample_padding, float rgb_padding, const snb_render_out* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- a7/a8: the MoE layer as an operator ------------------------------------------------- */
/* Replaces MOELayer.forward (tutel_moe_layer_nobatch.py:733-797) -> TopKGate.apply_on_expert_fn /
 * apply_on_expert_fn_nobatch (98-352) for one layer with an external gate input: gates =
 * softmax(gate_input @ wg^T) in fp32 (105-126), routing (snb_route_top1 semantics), dispatch, the expert stack,
 * combine with the gate value (dropped rows = 0).  input / gate_input / y: fp32 [S, width]; gate_input == NULL uses
 * `input` (the reference's default).  No activation is applied to y (NeRFMoE applies its ReLU afterwards,
 * nerf_moe.py:384).  fp32 CUDA path; workspace: snb_workspace_bytes(m, S, capacity_factor).  moe_idx int32[S]
 * (nullable) = topk indices of every sample; l_aux fp32[1]. */
int snb_moe_layer_forward(snb_model_t* m, const float* input, const float* gate_input, int64_t S,
                          const snb_route_opts* opts, float* y, int32_t* moe_idx, float* l_aux, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Stand-alone composite (rendering.py:436-494, flip=False): z [N,S] ascending, raw [N,S,4]. */
int snb_composite(const float* z, const float* raw, const float* last_delta, int64_t n_rays,
                  int32_t n_samples, int32_t white_bkgd, float* rgb, float* depth,
                  float* depth_variance, float* bg_lambda, float* weights, void* stream);

/* Stand-alone inverse-CDF sampling (rendering.py:587-637); u nullable => det linspace. */
int snb_sample_pdf(const float* bins, const float* weights, const float* u, int64_t n_rays,
                   int32_t n_bins_minus1, int32_t n_fine, float* z_fine, void* stream);

/* ---- self-test of the tcgen05 building block (used by tests/ and smoke) ------------------ */
/* D[128,N] = A[128,K] * B[N,K]^T with bf16 operands / fp32 accumulate through the same UMMA
 * descriptor + TMEM + bulk-copy machinery the fused kernels use.  A,B bf16 row-major (K
 * contiguous) device pointers, D fp32 row-major. K%16==0, N%16==0, N<=256. */
int snb_umma_selftest(const void* a_bf16, const void* b_bf16, int32_t N, int32_t K, float* d,
                      int32_t variant, void* stream);

/* Measurement aid: clocks for `reps` back-to-back tcgen05.mma (M=128, K=16, bf16, given N) with the A operand in
 * shared memory (a_in_tmem=0) or tensor memory (1); flags&1 adds a stream of 8 KB bulk copies landing in shared
 * memory, flags&2 adds four warps of tcgen05.ld readers, flags&4 runs a 2-CTA cluster issuing cta_group::2 (M=256)
 * instructions instead.  out6 = {clocks, copies, ld iterations x4}.  Synchronises
 * the stream. */
int snb_umma_microbench(int32_t N, int32_t a_in_tmem, int32_t flags, int32_t reps, uint64_t* out6, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWITCH_NERF_B200_H_ */

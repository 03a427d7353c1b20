#!/usr/bin/env python
"""bench.py -- headline benchmark of the Switch-NeRF forward/render hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): point-samples/sec (rays x network evaluations per ray) per box.
Workload at every N (BASELINE.json configs[1], per GPU): Building topology, 8 experts, width 256,
8192 rays x (257 coarse + 257 fine) samples = 4 210 688 point-samples per step, model_chunk_size
131072, capacity_factor 1.0, batch-prioritised routing, bf16 tensor-core compute, synthetic rays and
random-init weights (no datasets/checkpoints offline).  One "step" = one rendering.render_rays forward
over the ray batch.  Rays are independent units: ranks shard them with no data-path collective (the
reference ships with --no_expert_parallel), so scaling is "weak" (8192 rays per GPU).

`value`  : device-timed (CUDA events, max over ranks) with the ray batch already resident in HBM.
`e2e`    : the same step through the public API (switch_nerf_b200.rendering.render_rays) with HOST
           pinned rays: the H2D copy of rays/indices and the D2H read of the per-ray result are inside
           the timed region.
`--impl reference`: the reference's own CPU implementation of the path (the oracle port, pinned
           bit-exact against the unmodified reference) on the host cores, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, COARSE, FINE, CHUNK, EXPERTS = 8192, 257, 257, 131072, 8
FLOPS_PER_SAMPLE = 1_439_232            # SURVEY.md 8d, Building, forward, useful work only
FLOPS_BACK_KEPT = 917_504 + 512 + 131_072 + 84_736 + 768      # launch #2, sample that reaches an expert
FLOPS_BACK_DROPPED = 512 + 131_072 + 84_736 + 768             # launch #2, dropped sample (no expert stack)
WORKLOAD = "building: 8192 rays x (257+257) samples, 8 experts, width 256, chunk 131072, cf=1.0, BPR"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-comparator", action="store_true")
    ap.add_argument("--no-ep", action="store_true", help="N > 1: skip the expert-parallel leg (`ep` key)")
    ap.add_argument("--workload", default="building", choices=["building", "mission_bay"],
                    help="building (default, BASELINE.json configs[1], the line the driver records) or mission_bay "
                         "(configs[3]: MipNeRFMoE width 512, 13312 rays x (256+256) intervals, chunk 212992)")
    ap.add_argument("--sweep", default=None, choices=["cf"],
                    help="cf: BASELINE.json configs[4] -- capacity factor 0.5/1/2 x batch_prioritized_routing on/off x gate "
                         "balance (raw random init ... balanced) on the Building workload, one JSON line with every point")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--parallelism", default="dp", choices=["dp", "ep"],
                    help="dp: every expert on every GPU, rays sharded (the reference's shipped mode). "
                         "ep: experts sharded E/N per GPU, P2P record exchange (BASELINE.json configs[2])")
    return ap.parse_args()


def ncu_traffic():
    """dram__bytes_read+write per k_back launch from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "k_back_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get("dram_bytes_per_launch")
    return None


def measured_peaks(clocks=None):
    """Roofline denominator: MEASURED_PEAKS.json (driver-written).  The timed region of this bench is a fraction of a
    second at (normally) the maximum SM clock, i.e. burst conditions, so the like-for-like peak is the burst figure
    `bf16_tflops`; the sustained figure is used only when the sampled clock sat visibly below the maximum."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    burst_like = True
    if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz"):
        burst_like = clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"]
    if os.path.exists(p):
        d = json.load(open(p))
        if burst_like:
            return d.get("bf16_tflops", 1682.7), d.get("hbm_gbs", 6534.5), "MEASURED_PEAKS.json bf16_tflops (burst: SM clock at max during the timed region)"
        return d.get("bf16_tflops_sustained", 1377.8), d.get("hbm_gbs", 6534.5), "MEASURED_PEAKS.json bf16_tflops_sustained (SM clock below max during the timed region)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(rank, device):
    from switch_nerf_b200 import synthetic as O
    from switch_nerf_b200.configs import make_hparams
    from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
    sd = O.benchmark_state_dict(num_experts=EXPERTS, appearance_count=2048, seed=0, n_rays=N_RAYS, coarse=COARSE)
    hp = make_hparams(num_experts=EXPERTS, capacity_factor=1.0, bpr=True, model_chunk_size=CHUNK,
                      coarse_samples=COARSE, fine_samples=FINE, amp_bf16=True, moe_return_gates=False)
    model = get_nerf_moe_inner(hp, 2048, 3)
    model.load_state_dict(sd)
    model = model.to(device).eval()
    rays, idx = O.synthetic_rays(N_RAYS, 2048, seed=100 + rank)
    return model, hp, rays, idx, sd


def pick_cpu_threads(step_small):
    """Use as many host threads as actually help: time a tiny sample at several thread counts (more threads than
    the container's CPU quota, or than the small GEMMs can use, make the torch CPU ops slower)."""
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (avail, 64, 32, 16, 8) if 1 <= c <= avail}, reverse=True)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        step_small()
        t0 = time.perf_counter()
        step_small()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    from oracle import switch_nerf_oracle as O
    from switch_nerf_b200.synthetic import benchmark_state_dict
    sd = benchmark_state_dict(num_experts=EXPERTS, appearance_count=2048, seed=0, n_rays=N_RAYS, coarse=COARSE)
    cfg = O.default_cfg(sd, 1.0, True)
    n = 1024                                 # bounded sample of the workload: 1024 of the 8192 rays per step
    rays, idx = O.synthetic_rays(N_RAYS, 2048, seed=100)
    with torch.no_grad():
        cores = pick_cpu_threads(lambda: O.render_rays(sd, cfg, rays[:32], idx[:32], coarse_samples=COARSE,
                                                       fine_samples=FINE, model_chunk_size=CHUNK))
    rays, idx = rays[:n], idx[:n]
    samples = n * (COARSE + FINE)

    def step():
        with torch.no_grad():
            O.render_rays(sd, cfg, rays, idx, coarse_samples=COARSE, fine_samples=FINE, model_chunk_size=CHUNK)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = samples * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "point-samples/sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n} of {N_RAYS} rays per step ({samples} point-samples)"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{n} rays x {COARSE + FINE} samples per step, fp32, torch CPU ops, {cores} threads",
                         "note": "oracle port of the reference's CPU path: bit-identical to the unmodified reference in fp32 and "
                                 "faster than it on the same cores (conservative baseline)"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline():
    from oracle import switch_nerf_oracle as O
    from switch_nerf_b200.synthetic import benchmark_state_dict
    sd = benchmark_state_dict(num_experts=EXPERTS, appearance_count=2048, seed=0, n_rays=N_RAYS, coarse=COARSE)
    cfg = O.default_cfg(sd, 1.0, True)
    n = 4096
    rays, idx = O.synthetic_rays(N_RAYS, 2048, seed=100)
    with torch.no_grad():
        cores = pick_cpu_threads(lambda: O.render_rays(sd, cfg, rays[:32], idx[:32], coarse_samples=COARSE,
                                                       fine_samples=FINE, model_chunk_size=CHUNK))
    rays, idx = rays[:n], idx[:n]
    with torch.no_grad():
        t0 = time.perf_counter()
        ref = O.render_rays(sd, cfg, rays, idx, coarse_samples=COARSE, fine_samples=FINE, model_chunk_size=CHUNK)
        dt = time.perf_counter() - t0
    return {"value": n * (COARSE + FINE) / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{n} of {N_RAYS} rays x {COARSE + FINE} samples, one pass, fp32 oracle port, {cores} threads",
            "note": "the port is bit-identical to the unmodified reference in fp32 (tests/test_oracle_vs_reference.py) and "
                    "faster than it on the same cores (no dispatch buffer, no python MoE plumbing): a conservative baseline"}, ref, n


def psnr_db(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * __import__("math").log10(mse)


def parity_vs_oracle(model, hp, render_rays, rays_d, idx_d, ref, n):
    """The GPU arm's per-ray outputs on the SAME n rays the cpu_baseline leg just rendered through the oracle (fp32)."""
    model.args.moe_return_gates = True
    hp2 = __import__("copy").copy(hp)
    hp2.moe_return_gates = True
    with torch.no_grad():
        res = render_rays(model, None, rays_d[:n].contiguous(), idx_d[:n].contiguous(), hp2, None, None, True, True, False)[0]
    torch.cuda.synchronize()
    model.args.moe_return_gates = False
    rgb, depth = res["rgb_fine"].cpu(), res["depth_fine"].cpu()
    out = {"vs": f"oracle fp32 restatement (bit-identical to the unmodified reference on CPU), first {n} rays of the batch",
           "max_abs": float((rgb - ref["rgb_fine"]).abs().max()), "mean_abs": float((rgb - ref["rgb_fine"]).abs().mean()),
           "psnr_db": psnr_db(rgb, ref["rgb_fine"]),
           "depth_max_rel": float(((depth - ref["depth_fine"]).abs() / ref["depth_fine"].abs().clamp_min(1e-3)).max()),
           "route_agree": float((res["moe_gates_coarse"].view(-1).cpu() == ref["moe_gates_coarse"].view(-1).long()).float().mean())}
    return out


def parity_vs_reference_cuda(res):
    """Per-ray outputs of the timed workload (rank 0's 8192 rays) vs the committed golden of the UNMODIFIED reference run
    on a B200 under torch.autocast('cuda', bf16) (tests/golden/render_bench_building_bf16cuda.npz)."""
    import numpy as np
    p = os.path.join(ROOT, "tests", "golden", "render_bench_building_bf16cuda.npz")
    if not os.path.exists(p):
        return None
    g = np.load(p)
    rgb, ref = res["rgb_fine"].cpu(), torch.from_numpy(g["rgb_fine"])
    depth, dref = res["depth_fine"].cpu(), torch.from_numpy(g["depth_fine"])
    return {"vs": "unmodified reference, cuda autocast bf16, B200 (committed golden, all 8192 rays)",
            "max_abs": float((rgb - ref).abs().max()), "mean_abs": float((rgb - ref).abs().mean()),
            "psnr_db": psnr_db(rgb, ref), "frac_le_1e-3": float(((rgb - ref).abs() <= 1e-3).float().mean()),
            "depth_max_rel": float(((depth - dref).abs() / dref.abs().clamp_min(1e-3)).max())}


def gpu_comparator(device):
    """The UNMODIFIED reference (baseline/_ref copy of switch_nerf: models.nerf_moe.NeRFMoE + rendering.render_rays) on this
    GPU under torch.autocast('cuda', bf16), Tutel's external kernels replaced by the torch restatements of
    oracle/ref_shims.py running on the GPU.  Warm-up 3, median of 10.  Comparator only: nothing of it is on the timed arm."""
    try:
        from oracle.install_ref import reference_root
        if reference_root() is None:
            return {"unavailable": "baseline/_ref not present (oracle/install_ref.py copies it where /root/reference exists)"}
        from oracle import ref_shims as R
        R.install_shims()
        from switch_nerf import rendering
        from switch_nerf_b200 import synthetic as SY
        torch.backends.cuda.matmul.allow_tf32 = False
        sd = SY.benchmark_state_dict(num_experts=EXPERTS, appearance_count=2048, seed=0, n_rays=N_RAYS, coarse=COARSE)
        hp = R.make_hparams(num_experts=EXPERTS, capacity_factor=1.0, bpr=True, model_chunk_size=CHUNK,
                            coarse_samples=COARSE, fine_samples=FINE, amp_bf16=True, moe_return_gates=False)
        m = R.build_reference_model(hp, appearance_count=2048).eval()
        m.load_state_dict(sd)
        m = m.to(device)
        rays, idx = SY.synthetic_rays(N_RAYS, 2048, seed=100)
        rays, idx = rays.to(device), idx.to(device)

        def step():
            with torch.no_grad(), R.stable_argsort(), torch.autocast("cuda", dtype=torch.bfloat16):
                return rendering.render_rays(m, None, rays, idx, hp, None, None, True, True, False)[0]

        for _ in range(3):
            step()
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = 0.5 * (ts[4] + ts[5])
        peak_mb = torch.cuda.max_memory_allocated(device) / 2 ** 20
        del m
        torch.cuda.empty_cache()
        return {"value": N_RAYS * (COARSE + FINE) / (med * 1e-3), "unit": "samples/s", "ms_per_step_median": med,
                "kind": "unmodified reference modules (baseline/_ref), torch.autocast cuda bf16, cuBLAS GEMMs, Tutel kernels "
                        "as torch ops on the GPU (oracle/ref_shims.py)", "warmup": 3, "steps": 10, "peak_mem_mib": peak_mb}
    except Exception as e:      # the comparator must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ---------------------------------------------------------------------------------------------------------------
# --workload mission_bay: BASELINE.json configs[3] (mission_bay.yaml: MipNeRFMoE, width 512; README recipe: 13312 rays per
# iteration, coarse / fine 257 edges = 256 + 256 intervals, model_chunk_size 212992).  Same JSON contract; rays shard
# over the ranks with no data-path collective.
# ---------------------------------------------------------------------------------------------------------------
MB_RAYS, MB_EDGES, MB_CHUNK = 13312, 257, 212992
MB_FLOPS_PER_SAMPLE = 5_479_936
MB_FLOPS_BACK_KEPT = 3_670_016 + 1_024 + 524_288 + 150_272 + 768
MB_FLOPS_BACK_DROPPED = 1_024 + 524_288 + 150_272 + 768
MB_WORKLOAD = "mission_bay: 13312 rays x (256+256) intervals, MipNeRFMoE width 512, 8 experts, chunk 212992, cf=1.0, BPR"


def mission_bay_inputs(rank, device):
    from switch_nerf_b200 import synthetic as SY
    from switch_nerf_b200.configs import make_hparams
    from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
    sd = SY.mission_bay_state_dict(num_experts=EXPERTS, appearance_count=2048, seed=0, n_rays=MB_RAYS, coarse=MB_EDGES)
    hp = make_hparams(num_experts=EXPERTS, capacity_factor=1.0, bpr=True, model_chunk_size=MB_CHUNK, coarse_samples=MB_EDGES,
                      fine_samples=MB_EDGES, width=512, amp_bf16=True, moe_return_gates=False, nerfmoe_class_name="MipNeRFMoE")
    hp.perturb = 0
    model = get_nerf_moe_inner(hp, 2048, 3)
    model.load_state_dict(sd)
    model = model.to(device).eval()
    rays, radii, idx = SY.mission_bay_rays(MB_RAYS, 2048, seed=100 + rank)
    return model, hp, rays, radii, idx, sd


def mission_bay_oracle(sd, rays, radii, idx, n, mode):
    from oracle import switch_nerf_oracle as O
    cfg = O.default_cfg(sd, 1.0, True, mip=True)
    with torch.no_grad():
        return O.render_rays_mip(sd, cfg, rays[:n], radii[:n], idx[:n], coarse_samples=MB_EDGES, fine_samples=MB_EDGES,
                                 model_chunk_size=MB_CHUNK, mode=mode, flavor="cuda")


def main_mission_bay(args, rank, local_rank, world):
    import ctypes as C
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from switch_nerf_b200 import _lib as L
    from switch_nerf_b200.rendering_mip import render_rays as render_rays_mip
    model, hp, rays_h, radii_h, idx_h, sd = mission_bay_inputs(rank, device)
    model.precision = args.precision
    lib = L.lib()
    rays_pin, radii_pin, idx_pin = rays_h.pin_memory(), radii_h.pin_memory(), idx_h.to(torch.int32).pin_memory()
    rays_d, radii_d, idx_d = rays_pin.to(device), radii_pin.to(device), idx_pin.to(device)
    samples_per_step = MB_RAYS * 2 * (MB_EDGES - 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    @torch.no_grad()
    def step_resident():
        return render_rays_mip(model, rays_d, radii_d, idx_d, hp, True, True)[0]

    @torch.no_grad()
    def step_e2e():
        res = render_rays_mip(model, rays_pin.to(device, non_blocking=True), radii_pin.to(device, non_blocking=True),
                              idx_pin.to(device, non_blocking=True), hp, True, True)[0]
        return res["rgb_fine"].cpu(), res["depth_fine"].cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        barrier()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    step_e2e()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.snb_profile_enable(1)
    launches0 = lib.snb_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = lib.snb_launch_count() - launches0
    prof = (C.c_double * 4)()
    L.check(lib.snb_profile_collect(prof))
    lib.snb_profile_enable(0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        hp2 = __import__("copy").copy(hp)
        hp2.moe_return_gates = True
        model.args.moe_return_gates = True
        res = render_rays_mip(model, rays_d, radii_d, idx_d, hp2, True, True)[0]
        model.args.moe_return_gates = False
        dropped, total, hist = 0, 0, torch.zeros(EXPERTS, dtype=torch.float64)
        for key in ("moe_gates_coarse", "moe_gates_fine"):
            g = res[key].view(-1).cpu()
            for i in range(0, g.numel(), MB_CHUNK):
                c = torch.bincount(g[i:i + MB_CHUNK], minlength=EXPERTS)
                capc = int(1.0 * ((min(MB_CHUNK, g.numel() - i) + EXPERTS - 1) // EXPERTS))
                dropped += int(torch.clamp(c - capc, min=0).sum())
                total += int(c.sum())
                hist += c.double()
        kept_frac = 1.0 - dropped / max(total, 1)
        shares = [round(float(v), 4) for v in (hist / hist.sum())]
        ms_per_step = ms_total / args.steps
        peak_tf, _, which = measured_peaks(clocks)
        front_ms, route_ms, back_ms, n_chunks = prof[0], prof[1], prof[2], prof[3]
        roof = None
        if n_chunks > 0 and back_ms > 0:
            per_launch = samples_per_step * args.steps / n_chunks
            flops_back = kept_frac * MB_FLOPS_BACK_KEPT + (1.0 - kept_frac) * MB_FLOPS_BACK_DROPPED
            tf = per_launch * flops_back / (back_ms / n_chunks * 1e-3) / 1e12
            roof = {"kernel": "k_back_wide<2> (recompute h + 7 expert layers of 512x512 + combine + sigma/colour heads)",
                    "bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                    "peak_source": which, "avg_launch_ms": back_ms / n_chunks, "traffic": None,
                    "algorithmic_gflop_per_launch": per_launch * flops_back / 1e9,
                    "phase_ms_per_step": {"front": front_ms / args.steps, "route": route_ms / args.steps, "back": back_ms / args.steps},
                    "step_tflops_per_gpu": samples_per_step * (MB_FLOPS_PER_SAMPLE - (1.0 - kept_frac) * 3_670_016) / (ms_per_step * 1e-3) / 1e12}
        out = {"metric": "point-samples/sec", "value": world * samples_per_step / (ms_per_step * 1e-3), "unit": "samples/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
               "config": {"workload": MB_WORKLOAD, "rays_per_gpu": MB_RAYS, "l2": "256 MiB buffer written between timed steps",
                          "weights": "seeded random init, gate LayerNorm bias balanced on the ray batch",
                          "expert_shares": shares, "kept_fraction": round(kept_frac, 4),
                          "parallelism": f"dp{world} over rays, no data-path collective"},
               "clocks": clocks,
               "e2e": {"value": world * samples_per_step * args.steps / e2e_s, "unit": "samples/s",
                       "h2d_bytes_per_step": int((rays_pin.numel() + radii_pin.numel() + idx_pin.numel()) * 4),
                       "d2h_bytes_per_step": int(MB_RAYS * 4 * 4)},
               "gpu_launches": int(launches), "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            n = 1024
            mission_bay_oracle(sd, rays_h, radii_h, idx_h, 8, "fp32")
            cores = pick_cpu_threads(lambda: mission_bay_oracle(sd, rays_h, radii_h, idx_h, 8, "fp32"))
            t0 = time.perf_counter()
            ref = mission_bay_oracle(sd, rays_h, radii_h, idx_h, n, "fp32")
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": n * 2 * (MB_EDGES - 1) / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                                   "sample": f"{n} of {MB_RAYS} rays x {2 * (MB_EDGES - 1)} intervals, one pass, fp32 oracle port"}
            with torch.no_grad():
                mine = render_rays_mip(model, rays_d[:n].contiguous(), radii_d[:n].contiguous(), idx_d[:n].contiguous(), hp, True, True)[0]
            rgb = mine["rgb_fine"].cpu()
            out["parity"] = {"vs": f"oracle fp32 restatement of rendering_mip + MipNeRFMoE, first {n} rays",
                             "max_abs": float((rgb - ref["rgb_fine"]).abs().max()), "mean_abs": float((rgb - ref["rgb_fine"]).abs().mean()),
                             "psnr_db": psnr_db(rgb, ref["rgb_fine"])}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@torch.no_grad()
def main_sweep_cf(args, local_rank):
    """BASELINE.json configs[4]: the routing-imbalance throughput curve on the current kernels (single GPU)."""
    from switch_nerf_b200 import synthetic as SY
    from switch_nerf_b200.configs import make_hparams
    from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
    from switch_nerf_b200.rendering import render_rays
    torch.cuda.set_device(local_rank)
    rays_c, idx_c = SY.synthetic_rays(N_RAYS, 2048, seed=100)
    rays, idx = rays_c.cuda(), idx_c.cuda()
    base_sd = SY.synthetic_state_dict(num_experts=EXPERTS, appearance_count=2048, seed=0, gate_scale=4.0)
    pick = torch.randperm(N_RAYS, generator=torch.Generator().manual_seed(7))[:512]
    tt = torch.linspace(0, 1, 64)
    zz = rays_c[pick, 6:7] * (1 - tt) + rays_c[pick, 7:8] * tt
    pts = (rays_c[pick, None, 0:3] + rays_c[pick, None, 3:6] * zz[..., None]).reshape(-1, 3)
    rows = []
    # gate skew: 0 balance iterations = raw random init (3 experts take ~95 %), 3 / 8 = partially, 60 = balanced
    for iters in (0, 3, 8, 60):
        sd = SY.balance_gate(base_sd, pts, iters=iters) if iters > 0 else base_sd
        for cf in (0.5, 1.0, 2.0):
            for bpr in (True, False):
                hp = make_hparams(num_experts=EXPERTS, capacity_factor=cf, bpr=bpr, model_chunk_size=CHUNK, coarse_samples=COARSE,
                                  fine_samples=FINE, amp_bf16=True, moe_return_gates=True)
                model = get_nerf_moe_inner(hp, 2048, 3)
                model.load_state_dict(sd)
                model = model.cuda().eval()
                for _ in range(3):
                    res = render_rays(model, None, rays, idx, hp, None, None, True, True, False)[0]
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(max(args.steps, 3)):
                    res = render_rays(model, None, rays, idx, hp, None, None, True, True, False)[0]
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / max(args.steps, 3)
                gates = torch.cat([res["moe_gates_coarse"].view(-1), res["moe_gates_fine"].view(-1)])
                share = torch.bincount(gates, minlength=EXPERTS).float() / gates.numel()
                dropped_n, total_n = 0, 0
                for key in ("moe_gates_coarse", "moe_gates_fine"):
                    gk = res[key].view(-1)
                    for i in range(0, gk.numel(), CHUNK):
                        c = torch.bincount(gk[i:i + CHUNK], minlength=EXPERTS)
                        capc = int(cf * ((min(CHUNK, gk.numel() - i) + EXPERTS - 1) // EXPERTS))
                        dropped_n += int(torch.clamp(c - capc, min=0).sum())
                        total_n += int(c.sum())
                rows.append({"balance_iters": iters, "capacity_factor": cf, "bpr": bpr, "ms_per_step": round(ms, 4),
                             "msamples_per_s": round(N_RAYS * (COARSE + FINE) / ms / 1e3, 2),
                             "max_expert_share": round(float(share.max()), 4), "dropped_fraction": round(dropped_n / total_n, 4)})
                model.release()
    print(json.dumps({"sweep": "cf", "metric": "point-samples/sec", "unit": "M samples/s", "n_gpus": 1, "dtype": "bf16",
                      "data": "synthetic", "config": {"workload": WORKLOAD.replace("cf=1.0, BPR", "cf x BPR x gate balance sweep")},
                      "rows": rows}))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: switch_nerf_b200 has no CPU path")
    if args.workload == "mission_bay":
        return main_mission_bay(args, rank, local_rank, world)
    if args.sweep == "cf":
        if rank == 0:
            main_sweep_cf(args, local_rank)
        return
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from switch_nerf_b200 import _lib as L
    from switch_nerf_b200.rendering import render_rays
    import ctypes as C

    model, hp, rays_h, idx_h, sd = build_inputs(rank, device)
    model.precision = args.precision
    ep_group = None
    if args.parallelism == "ep" and world > 1:
        from switch_nerf_b200.expert_parallel import ExpertParallelGroup
        ep_group = ExpertParallelGroup(EXPERTS, CHUNK, 1.0)
        ep_group.attach(model)
    lib = L.lib()
    rays_pin, idx_pin = rays_h.pin_memory(), idx_h.to(torch.int32).pin_memory()
    rays_d, idx_d = rays_pin.to(device), idx_pin.to(device)
    samples_per_step = N_RAYS * (COARSE + FINE)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > 126 MB L2

    # inference, as the reference's eval scripts run it (no autograd graph: with grad enabled render_rays also keeps the
    # per-sample taps its backward needs)
    @torch.no_grad()
    def step_resident():
        return render_rays(model, None, rays_d, idx_d, hp, None, None, True, True, False)[0]

    @torch.no_grad()
    def step_e2e():
        r = rays_pin.to(device, non_blocking=True)
        i = idx_pin.to(device, non_blocking=True)
        res = render_rays(model, None, r, i, hp, None, None, True, True, False)[0]
        return res["rgb_fine"].cpu(), res["depth_fine"].cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_events=True):
        evs = []
        barrier()
        for _ in range(steps):
            flush.zero_()                                  # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    min_warm = int(os.environ.get("SNB_BENCH_MIN_WARMUP", 3))   # profiling runs only; the contract is W >= 3
    for _ in range(max(args.warmup, min_warm)):
        step_resident()
    step_e2e()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.snb_profile_enable(1)
    launches0 = lib.snb_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = lib.snb_launch_count() - launches0
    prof = (C.c_double * 4)()
    L.check(lib.snb_profile_collect(prof))
    lib.snb_profile_enable(0)
    # e2e: host wall clock around H2D + step + D2H, max over ranks
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # expert-parallel leg (BASELINE.json configs[2]) in the same process group: experts sharded E/N per GPU, records /
    # result rows exchanged by P2P stores inside the kernels.  An untimed bit-equality check of every per-ray output
    # against the all-local run on each rank comes first (the parity test a 1-GPU test box cannot run).
    ep_report = None
    if world > 1 and ep_group is None and not args.no_ep and EXPERTS % world == 0 and args.precision == "bf16":
        from switch_nerf_b200.expert_parallel import ExpertParallelGroup
        dp_out = {k: v.clone() for k, v in step_resident().items() if torch.is_tensor(v)}
        barrier()
        grp = ExpertParallelGroup(EXPERTS, CHUNK, 1.0)
        grp.attach(model)
        for _ in range(3):
            step_resident()
        ep_out = step_resident()
        torch.cuda.synchronize()
        mism = {k: {"mismatched": int((dp_out[k] != ep_out[k]).sum()), "numel": int(dp_out[k].numel()),
                    "max_abs": float((dp_out[k].double() - ep_out[k].double()).abs().max())}
                for k in dp_out if not torch.equal(dp_out[k], ep_out[k])}
        eq = not mism
        if mism:
            print(f"[rank {rank}] expert-parallel output differs from the all-local run: {json.dumps(mism)}", file=sys.stderr, flush=True)
        flag = torch.tensor([1 if eq else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ms_ep = timed(step_resident, args.steps)
        barrier()
        grp.detach(model)
        grp.close()
        ep_report = {"value": world * samples_per_step / (ms_ep / args.steps * 1e-3), "unit": "samples/s",
                     "ms_per_step": ms_ep / args.steps, "bit_equal_to_dp": bool(flag.item()),
                     "compared": sorted(dp_out), "experts_per_gpu": EXPERTS // world,
                     "data_path": "48 B records out / 16 B rows back by P2P st.global over NVLink inside the routing and "
                                  "launch-#2 kernels, release/acquire flags, no NCCL per chunk"}
        del dp_out, ep_out

    # routing statistics of this workload (one untimed pass): expert shares and the dropped fraction, per chunk
    if ep_group is not None:
        barrier()
        ep_group.detach(model)      # rank 0 alone renders below: routing is per source rank, identical either way
    kept_frac, shares = 1.0, None
    parity_ref = None
    if rank == 0:
        model.args.moe_return_gates = True
        with torch.no_grad():
            res = render_rays(model, None, rays_d, idx_d, hp, None, None, True, True, False)[0]
        model.args.moe_return_gates = False
        parity_ref = parity_vs_reference_cuda(res) if args.precision == "bf16" else None
        cap = int(1.0 * ((CHUNK + EXPERTS - 1) // EXPERTS))
        dropped, total, hist = 0, 0, torch.zeros(EXPERTS, dtype=torch.float64)
        for key in ("moe_gates_coarse", "moe_gates_fine"):
            g = res[key].view(-1).cpu()
            for i in range(0, g.numel(), CHUNK):
                c = torch.bincount(g[i:i + CHUNK], minlength=EXPERTS)
                capc = int(1.0 * ((min(CHUNK, g.numel() - i) + EXPERTS - 1) // EXPERTS))
                dropped += int(torch.clamp(c - capc, min=0).sum())
                total += int(c.sum())
                hist += c.double()
        kept_frac = 1.0 - dropped / max(total, 1)
        shares = [round(float(v), 4) for v in (hist / hist.sum())]
        del res

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = world * samples_per_step / (ms_per_step * 1e-3)
        peak_tf, peak_hbm, which = measured_peaks(clocks)
        front_ms, route_ms, back_ms, n_chunks = prof[0], prof[1], prof[2], prof[3]
        roof = None
        if n_chunks > 0 and back_ms > 0:
            # dominant kernel = launch #2 (k_back).  Algorithmic FLOPs per launch = samples of the chunk x the
            # back-end figure (upper bound uses every sample as kept; dropped samples skip the expert stack).
            per_launch_samples = samples_per_step * args.steps / n_chunks
            avg_ms = back_ms / n_chunks
            flops_back = kept_frac * FLOPS_BACK_KEPT + (1.0 - kept_frac) * FLOPS_BACK_DROPPED   # per sample, this workload
            tf = per_launch_samples * flops_back / (avg_ms * 1e-3) / 1e12
            roof = {"kernel": "k_back (recompute h + expert stack + combine + sigma/colour heads)", "bound": "tensor", "achieved": tf,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "peak_source": which,
                    "avg_launch_ms": avg_ms, "traffic": ncu_traffic(), "traffic_unit": "bytes/launch (dram read+write, ncu)",
                    "algorithmic_gflop_per_launch": per_launch_samples * flops_back / 1e9,
                    "phase_ms_per_step": {"front": front_ms / args.steps, "route": route_ms / args.steps,
                                          "back": back_ms / args.steps},
                    "step_tflops_per_gpu": samples_per_step * (FLOPS_PER_SAMPLE - (1.0 - kept_frac) * 917_504) / (ms_per_step * 1e-3) / 1e12}
        out = {
            "metric": "point-samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, min_warm), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": N_RAYS, "l2": "256 MiB buffer written between timed steps",
                       "weights": "seeded random init, gate LayerNorm bias balanced on the ray batch (emulates the l_aux-trained gate)",
                       "expert_shares": shares, "kept_fraction": round(kept_frac, 4),
                       "parallelism": (f"ep{world}: rays sharded, experts sharded {EXPERTS // world}/GPU, 48 B records out / "
                                       "16 B rows back by P2P stores (no NCCL on the data path)") if ep_group is not None
                       else f"dp{world} over rays, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": world * samples_per_step * args.steps / e2e_s, "unit": "samples/s",
                    "h2d_bytes_per_step": int(rays_pin.numel() * 4 + idx_pin.numel() * 4),
                    "d2h_bytes_per_step": int(N_RAYS * 3 * 4 + N_RAYS * 4)},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if ep_report is not None:
            out["ep"] = ep_report
        if parity_ref is not None:
            out["parity_vs_reference_cuda"] = parity_ref
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"], ref, n = cpu_baseline()
            out["parity"] = parity_vs_oracle(model, hp, render_rays, rays_d, idx_d, ref, n)
        if world == 1 and not args.no_gpu_comparator:
            out["gpu_comparator"] = gpu_comparator(device)
        print(json.dumps(out))
    if ep_group is not None:
        ep_group.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

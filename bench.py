"""Benchmark of the RoHM denoising hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

  python bench.py --gpus 1 --steps 3 --warmup 3            # one process, cuda:0
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W            # one rank per GPU, NCCL
  python bench.py --impl reference ...                     # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): PoseNet denoiser, batch 32 x 145-frame clips (T = 144 motion frames, 145
tokens), 1000 DDPM steps, per GPU.  One "step" of this benchmark = one complete 1000-step p_sample_loop over the
batch (what `eval_losses` runs for the drivers); metric = denoised clips / second, whole job.  Multi-GPU = independent
clips sharded over ranks (weak scaling: 32 clips per rank) + one NCCL all-gather of the final outputs per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

B_PER_GPU = 32
T_FRAMES = 144          # 145 raw frames -> 144 motion-representation frames (+1 timestep token = 145 tokens)
DIFFUSION_STEPS = 1000
C_FEATS = 294
METRIC = "denoised motion clips/sec (145-frame, 1000-step PoseNet p_sample_loop)"
# algorithmic FLOPs of the tensor-core GEMMs of one PoseNet forward, per clip, S = 145 tokens:
# 8 x (QKV 226.49 + out 75.50 + FFN 301.99) MFLOP @S=144 scaled to 145 tokens + embed + head
def gemm_flops_per_forward(B, S, D=512, F=1024, C=294, Cout=272, L=8):
    per_tok = L * (2 * D * 3 * D + 2 * D * D + 2 * D * F + 2 * F * D) + 2 * C * D + 2 * D * Cout
    return float(B) * S * per_tok


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop.wait(0.5)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def build_posenet(device):
    from rohm_b200 import synthetic
    from rohm_b200.posenet import PoseNet
    ds = synthetic.make_dataset('pose')
    model = PoseNet(dataset=ds, body_feat_dim=C_FEATS, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4,
                    device=device, traj_feat_dim=22)
    sd = synthetic.synth_state_dict(model, 1)
    model.load_state_dict(sd)
    if device is not None:
        model.to(device)
    return model.eval(), sd


# -------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference algorithm on the host cores
# -------------------------------------------------------------------------------------------------------------
def cpu_port_clips_per_s(sd, n_clips, n_steps, threads):
    """Times `n_steps` consecutive ancestral steps (denoiser + posterior update + RNG) of the oracle on `n_clips`
    clips with `threads` host threads and extrapolates linearly to the 1000-step chain (step cost is homogeneous)."""
    from oracle import diffusion_oracle as do
    from oracle import posenet_oracle
    from rohm_b200 import synthetic
    torch.set_num_threads(threads)
    tables, tmap = do.create_diffusion('cosine', DIFFUSION_STEPS, '')
    cond = synthetic.posenet_batch(n_clips, T_FRAMES, 3)['cond']
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_clips, C_FEATS, 1, T_FRAMES, generator=g)
    with torch.no_grad():
        x0 = posenet_oracle.posenet_forward(sd, x, cond, torch.full((n_clips,), 999, dtype=torch.long))  # warm-up
        t0 = time.perf_counter()
        for i in range(DIFFUSION_STEPS - 1, DIFFUSION_STEPS - 1 - n_steps, -1):
            x0 = posenet_oracle.posenet_forward(sd, x, cond, torch.full((n_clips,), tmap[i], dtype=torch.long))
            x = do.p_sample_step(tables, i, x, x0, torch.randn(x.shape, generator=g))
        dt = time.perf_counter() - t0
    return n_clips / (dt / n_steps * DIFFUSION_STEPS), dt


def host_threads():
    """Threads for the CPU legs: every core up to 32 (torch's intra-op pool stops scaling, then collapses, on this
    workload's GEMM sizes beyond that -- measured 25 s/step with 128 threads on the 128-core GPU host)."""
    return max(1, min(os.cpu_count() or 1, 32))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    _, sd = build_posenet(None)
    n_clips, n_steps = B_PER_GPU, 16
    vals = []
    for _ in range(args.warmup):
        cpu_port_clips_per_s(sd, n_clips, 1, cores)
    t_all = 0.0
    for _ in range(args.steps):
        v, dt = cpu_port_clips_per_s(sd, n_clips, n_steps, cores)
        vals.append(v)
        t_all += dt
    value = float(np.mean(vals))
    sample = (f"{n_clips} clips x {n_steps} consecutive DDPM steps (PoseNet forward + posterior update + RNG) per "
              f"bench step, extrapolated x{DIFFUSION_STEPS // n_steps} to the 1000-step chain")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * B_PER_GPU / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "cpu"),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure python and cannot travel to the GPU box; this is the oracle port (pinned to the "
                "reference by tests/golden) on the host cores, extrapolated",
    }
    print(json.dumps(line))


def workload_config(n_gpus, device_kind):
    return {"workload": "BASELINE configs[1]: PoseNet denoiser, batch 32 x 145-frame clips (T=144 motion frames, 145 "
                        "tokens, 294 channels), 1000 DDPM steps, p_sample (no guidance)",
            "clips_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * n_gpus, "frames": T_FRAMES,
            "diffusion_steps": DIFFUSION_STEPS, "parallelism": f"clip-sharded x{n_gpus} (no intra-step collective)",
            "precision_mode": os.environ.get("ROHM_B200_PRECISION", "f16x2"),
            "l2": "flushed (256 MiB write) between timed iterations", "device": device_kind}


# -------------------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rohm_b200", choices=["rohm_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    from rohm_b200 import diffusion, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    distributed = world > 1
    if distributed:
        # stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    model, sd = build_posenet(dev)
    dargs = argparse.Namespace(noise_schedule='cosine', sigma_small=True)
    diff = diffusion.create_gaussian_diffusion(dargs, diffusion, diffusion.SpacedDiffusionPoseNet, DIFFUSION_STEPS, '', dev)
    B, T = B_PER_GPU, T_FRAMES
    shape = [B, C_FEATS, 1, T]
    cond_host = synthetic.posenet_batch(B, T, 100 + rank)['cond'].pin_memory()
    cond_dev = cond_host.to(dev, non_blocking=True)
    out_host = torch.empty(shape, dtype=torch.float32).pin_memory()
    gathered = torch.empty([world * B] + shape[1:], device=dev) if distributed else None
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    torch.manual_seed(1234 + rank)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step_resident():
        batch = {'cond': cond_dev}
        out = diff.p_sample_loop(model, batch, shape, clip_denoised=False, cond_fn_with_grad=False)
        if distributed:
            dist.all_gather_into_tensor(gathered, out)
        return out

    def one_step_e2e():
        batch = {'cond': cond_host.to(dev, non_blocking=True)}
        _, out = diff.eval_losses(model=model, batch=batch, shape=shape, progress=False, clip_denoised=False,
                                  cond_fn_with_grad=False, compute_loss=False)
        if distributed:
            dist.all_gather_into_tensor(gathered, out)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def timed(fn, k):
        """k iterations, each bracketed by CUDA events on the launching stream, L2 flushed in between (untimed)."""
        evs = []
        for _ in range(k):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    for _ in range(args.warmup):
        one_step_resident()
    barrier()
    with ClockSampler(local) as clocks:
        ms_total = timed(one_step_resident, args.steps)
        barrier()
    tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step / 1000.0)

    # end-to-end through the public API with host buffers
    one_step_e2e()
    barrier()
    e2e_ms = timed(one_step_e2e, args.steps)
    barrier()
    te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(te.item()) / args.steps / 1000.0)

    # per-kernel event timing of the denoiser (sampled forwards, same process, same data)
    engine = model._engine
    x = torch.randn(shape, device=dev)
    ts = torch.full((B,), 500, device=dev, dtype=torch.int64)
    cat_ms = {"gemm": 0.0, "attention": 0.0, "layernorm": 0.0, "other": 0.0}
    cat_n = dict.fromkeys(cat_ms, 0)
    reps = 10
    for _ in range(reps):
        ms, n = engine.profile(x, ts)
        for k in cat_ms:
            cat_ms[k] += ms[k] / reps
            cat_n[k] = n[k]
    launches_fwd = engine.launches_per_forward
    peaks = read_peaks()
    flops = gemm_flops_per_forward(B, T + 1)
    gemm_s = cat_ms["gemm"] / 1000.0
    achieved = flops / gemm_s / 1e12 if gemm_s > 0 else None
    prec = engine.precision  # 3 = TF32 hi/lo x 3 products, 2 = fp16 hi/lo x 3 products, 1 = single-pass TF32
    passes = 1 if prec == 1 else 3
    kernel_kind = {3: "tcgen05 kind::tf32 on TF32 hi/lo pairs, 3 products", 2: "tcgen05 kind::f16 on fp16 hi/lo pairs, 3 products",
                   1: "tcgen05 kind::tf32, single pass"}[prec]
    # tensor-pipe work per algorithmic flop: 3 products; fp16 products run at the bf16 rate, TF32 ones at half of it
    pipe_peak = peaks["bf16_tflops"] if prec == 2 else peaks["bf16_tflops"] / 2.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj["dram_bytes_per_launch"], f"profiles/r1_gemm_traffic.json ({tj['source']}, cold-cache ncu capture)"
    roofline = {
        "kernel": f"gemm_tile_kernel ({kernel_kind}), {cat_n['gemm']} launches per PoseNet forward",
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": (achieved / peaks["bf16_tflops"]) if achieved else None, "traffic": traffic,
        "traffic_source": traffic_src,
        "peak_source": peaks["source"],
        "algorithmic_flops_per_forward": flops, "avg_launch_us": 1000.0 * cat_ms["gemm"] / max(cat_n["gemm"], 1),
        "tensor_pipe_frac": (achieved * passes / pipe_peak) if achieved else None,
        "tensor_pipe_frac_note": "issued tensor work (3 products per algorithmic flop) / peak of that operand type",
        "share_of_forward": {k: cat_ms[k] / max(sum(cat_ms.values()), 1e-9) for k in cat_ms},
        "forward_ms_by_kernel_class": cat_ms,
    }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        v, dt = cpu_port_clips_per_s(sd, B, 64, cores)
        cpu_baseline = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                        "sample": f"{B} clips x 64 consecutive DDPM steps of the oracle port ({dt:.1f} s, {cores} of "
                                  f"{os.cpu_count()} host threads), extrapolated linearly to 1000 steps"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {3: "f32 (TF32 hi/lo error-compensated tensor-core GEMMs)", 2: "f32 (fp16 hi/lo error-compensated tensor-core GEMMs)",
                                                   1: "tf32"}[prec],
            "data": "synthetic", "config": workload_config(world, "B200"),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": cond_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": args.steps * DIFFUSION_STEPS * (launches_fwd + 1),
            "roofline": roofline,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

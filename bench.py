#!/usr/bin/env python
"""Benchmark of the obman_train hot path on B200 (driver contract).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

One "step" = one training step (HandNet forward + backward + Adam) on a synthetic batch of
BASELINE.json configs[1]: per-GPU batch 64, 256x256 images, ResNet-18 (shared encoder) + ManoLayer(778 v)
+ AtlasNet (1 patch, ico-3 = 642 points) + Chamfer (M = 600 GT points) + Mano/Atlas losses; weak scaling
(per-GPU batch fixed).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PRECISION_NOTE = [""]
CFG = dict(resnet_version=18, mano_root="synthetic", mano_comps=30, mano_use_shape=True, mano_use_pca=True,
           mano_neurons=[1024, 256], mano_center_idx=0, mano_lambda_verts=0.167, mano_lambda_joints3d=0.167,
           mano_lambda_shape=0.167, mano_lambda_pose_reg=0.167, atlas_lambda=0.167, atlas_final_lambda=0.167,
           atlas_mesh=True, atlas_predict_trans=True, atlas_predict_scale=True, atlas_trans_weight=0.167,
           atlas_scale_weight=0.167, atlas_separate_encoder=False, atlas_ico_divisions=3, atlas_points_nb=600)
PER_GPU_BATCH = 64
IMG = 256
N_GT = 600
WORKLOAD = "BASELINE.json configs[1]: train step fwd+bwd+Adam, ResNet-18 + ManoLayer(778v) + AtlasNet(642 pts) + " \
           "Chamfer(600 GT) + Mano/Atlas losses, 256x256"


def use_config3():
    """BASELINE.json configs[2]: batch 256, separate Atlas encoder, ico-4 sphere (2562 points / 5120 faces),
    2500 GT points, Chamfer + contact_zones loss (dist_tanh, thresh 10/20)."""
    global PER_GPU_BATCH, N_GT, WORKLOAD
    PER_GPU_BATCH, N_GT = 256, 2500
    CFG.update(atlas_separate_encoder=True, atlas_ico_divisions=4, atlas_points_nb=2500, contact_lambda=1,
               collision_lambda=1, contact_zones="zones", contact_mode="dist_tanh", collision_mode="dist_tanh",
               contact_thresh=10, collision_thresh=20)
    WORKLOAD = ("BASELINE.json configs[2]: train step fwd+bwd+Adam, 2x ResNet-18 (separate Atlas encoder) + "
                "ManoLayer + AtlasNet(ico-4, 2562 pts) + Chamfer(2500 GT) + contact_zones loss, 256x256")


def use_config4():
    """BASELINE.json configs[3]: global batch 1024 = 8 x 128 per GPU, the full loss stack: configs[2] plus the
    edge-length regulariser on the 5120-face mesh; weak scaling like the headline."""
    global PER_GPU_BATCH, WORKLOAD
    use_config3()
    PER_GPU_BATCH = 128
    CFG.update(atlas_lambda_regul_edges=0.1)
    WORKLOAD = ("BASELINE.json configs[3]: per-GPU batch 128 (1024 on 8 GPUs), full loss stack: 2x ResNet-18 + ManoLayer + "
                "AtlasNet(ico-4) + Chamfer(2500 GT) + contact_zones + edge regulariser, 256x256")


def _hand_targets(B, g):
    if CFG.get("contact_lambda"):
        # hand-shaped targets (MANO template, mm) so that the contact masks are non-trivial
        from obman_train_b200.assets import load_contacts
        verts, _ = load_contacts()
        hand = torch.tensor(verts * 1000, dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
        return hand + torch.randn(B, 778, 3, generator=g) * 5
    return torch.randn(B, 778, 3, generator=g) * 40


def synthetic_sample(B, seed):
    """images U(0,1)-0.5; verts/joints N(0,40^2); object points N(0,40^2)+30 (mm)."""
    g = torch.Generator().manual_seed(seed)
    return {
        "images": torch.rand(B, 3, IMG, IMG, generator=g) - 0.5,
        "sides": ["right" if i % 2 == 0 else "left" for i in range(B)],
        "root": "wrist",
        "joints3d": torch.randn(B, 21, 3, generator=g) * 40,
        "verts3d": _hand_targets(B, g),
        "objpoints3d": torch.randn(B, N_GT, 3, generator=g) * 40 + 30,
    }


# ---------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md): nvidia-smi polled during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port; the reference itself cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_steps(batch, steps, warmup):
    """Train step (fwd + bwd + Adam) of the reference algorithm on the host cores: oracle/nets.py, pinned
    against the reference's own files by tests/test_oracle_vs_reference.py.  Returns seconds per step."""
    from oracle import nets
    from obman_train_b200.networks.handnet import HandNet
    # all host cores up to 32 threads: beyond that torch's CPU conv/bmm kernels slow down on this workload
    # (measured on the 128-core GPU box: 128 threads gave 0.26 img/s where 8 threads give ~19 img/s)
    threads = min(os.cpu_count() or 1, int(os.environ.get("OBMAN_BENCH_CPU_THREADS", "32")))
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = HandNet(**CFG).eval()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    leaves = []
    for k, v in state.items():
        if v.is_floating_point() and "running_" not in k and "th_" not in k and ".fc." not in k:
            v.requires_grad_(True)
            leaves.append(v)
    opt = torch.optim.Adam(leaves, lr=1e-4)
    tables = {s: {k: v.detach() for k, v in getattr(model.mano_branch, "mano_layer_" + s).named_buffers()
                  if k != "th_faces"} for s in ("right", "left")}
    grid, faces = model.atlas_branch.test_verts, model.atlas_branch.test_faces
    from obman_train_b200.assets import load_contacts
    zones = load_contacts()[1]
    sample = synthetic_sample(batch, 0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        total, _, _ = nets.handnet_forward(state, CFG, sample, tables, grid, faces, zones)
        total.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = int(os.environ.get("OBMAN_BENCH_CPU_BATCH", "8"))
    sec, threads = cpu_reference_steps(batch, max(1, min(args.steps, 5)), min(1, args.warmup))
    val = batch / sec
    line = {
        "impl": "reference", "metric": "train-step images/sec", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "batch %d of the same workload, fwd+bwd+Adam, oracle/nets.py on torch CPU" % batch},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": WORKLOAD,
            "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * n_gpus,
            "parallelism": "dp%d" % n_gpus, "precision": PRECISION_NOTE[0],
            "l2_policy": "activations per step (~3 GB) exceed the 126 MB L2; no explicit flush",
            "launch": "whole step replayed as one CUDA graph (--no-graph for eager launches); weight-gradient / "
                      "BN-gradient / weight-folding kernels on a second captured stream (OBMAN_OVERLAP=0 disables)"}


# ---------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------
def chamfer_report(peaks):
    """Second half of BASELINE.json's metric ("Chamfer kernel HBM GB/s vs peak"): the fused nearest-neighbour
    kernel at the workload's own shapes and at a large sweep point, on algorithmic bytes (20 B per point: 12 read,
    8 written).  The fused kernel does 200-3300 flop per algorithmic byte, so its bound is the FP32 issue rate
    (reported as well); the full sweep is profiles/chamfer_sweep_r1.json (scripts/bench_chamfer.py)."""
    from obman_train_b200 import functional as Fb
    hbm = peaks.get("hbm_gbs", 6650.0)
    issue_peak = 148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6
    out = {"unit": "GB/s on algorithmic bytes", "hbm_peak_gbs": hbm}
    g = torch.Generator(device="cuda").manual_seed(7)
    for tag, (b, n, m, reps) in {"workload": (PER_GPU_BATCH, 642 if N_GT == 600 else 2562, N_GT, 50),
                                 "sweep_2048x10000": (2048, 10000, 10000, 2)}.items():
        x = torch.randn(b, n, 3, device="cuda", generator=g) * 60
        y = torch.randn(b, m, 3, device="cuda", generator=g) * 60
        for _ in range(3):
            Fb.nearest_neighbours(x, y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            Fb.nearest_neighbours(x, y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 20.0 * b * (n + m) / ms / 1e6
        out[tag] = {"B": b, "N": n, "M": m, "ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm,
                    "pairs_per_s": 2.0 * b * n * m / ms * 1e3,
                    "frac_of_fp32_issue_peak": 2.0 * b * n * m * 7.3 / (ms * 1e-3) / issue_peak}
        del x, y
    return out


def run_b200(args):
    import torch.distributed as dist
    from obman_train_b200 import _lib, dense
    from obman_train_b200.networks.handnet import HandNet
    from obman_train_b200.trainer import FlatAdamTrainer
    from obman_train_b200.queries import TransQueries, BaseQueries

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    if lib.obman_device_ok() != 0:
        raise RuntimeError(lib.obman_get_last_error().decode())
    dense.set_precision(args.precision, args.precision)
    PRECISION_NOTE[0] = {
        "bf16x3": "fp32 operands split into bf16 hi+lo, 3 tensor-core products per contraction (fp32-equivalent to ~2^-17), fwd+bwd",
        "tf32x3": "3xTF32 tensor-core passes (fp32-equivalent) fwd+bwd", "tf32": "single-pass TF32"}[args.precision]

    torch.manual_seed(0)  # identical replicas on every rank
    model = HandNet(**CFG).eval().cuda()
    trainer = FlatAdamTrainer(model, lr=1e-4, world_size=world)
    B = PER_GPU_BATCH
    host = synthetic_sample(B, 1000 + rank)
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}

    def to_device(src):
        return {TransQueries.images: src["images"].cuda(non_blocking=True), BaseQueries.sides: src["sides"],
                "root": src["root"], TransQueries.joints3d: src["joints3d"].cuda(non_blocking=True),
                TransQueries.verts3d: src["verts3d"].cuda(non_blocking=True),
                TransQueries.objpoints3d: src["objpoints3d"].cuda(non_blocking=True)}

    resident = to_device(pinned)
    h2d = sum(v.numel() * 4 for v in host.values() if torch.is_tensor(v))

    def timed(n_steps, fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return ms.item()

    use_graph = not args.no_graph
    if use_graph:
        trainer.capture(resident)  # whole step (fwd + bwd + all-reduce + Adam) as one CUDA graph

    def step_resident():
        if use_graph:
            trainer.replay()
        else:
            trainer.step(resident)

    loss_host = torch.zeros(1).pin_memory()
    tensor_keys = [(TransQueries.images, "images"), (TransQueries.joints3d, "joints3d"),
                   (TransQueries.verts3d, "verts3d"), (TransQueries.objpoints3d, "objpoints3d")]

    feeder = None
    if use_graph:
        from obman_train_b200.trainer import PinnedFeeder
        feeder = PinnedFeeder(trainer)
        pinned_q = {q: pinned[name] for q, name in tensor_keys}

    def step_e2e():
        # every step: this step's inputs travel pinned host -> device (graph mode: on a copy stream, overlapped with
        # the previous step, trainer.PinnedFeeder), and the step's loss is read back by the host
        if use_graph:
            loss = feeder.step(next_host_sample=pinned_q)
        else:
            loss = trainer.step(to_device(pinned))
        loss_host.copy_(loss.detach(), non_blocking=False)  # the user's read of the step result

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # sampled through warm-up and the timed region (both under load)
    for _ in range(max(3, args.warmup)):
        step_resident()
    kern0 = _lib.kernel_count
    ms = timed(args.steps, step_resident)
    kernels = _lib.kernel_count - kern0
    if use_graph:  # replays issue no Python-side calls: count the kernels of one eager step instead
        k0 = _lib.kernel_count
        trainer.step(resident)
        kernels = (_lib.kernel_count - k0) * args.steps
    clocks = sampler.stop() if rank == 0 else None
    if feeder is not None:
        feeder.prefetch(pinned_q)  # primes the pipeline: step i runs while step i+1's batch is in flight
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(args.steps, step_e2e)

    # roofline of the dominant kernel (gemm_tc_kernel): per-launch CUDA events + algorithmic FLOPs
    from obman_train_b200 import streams
    overlap = streams.set_enabled(False)  # per-launch timing needs every launch alone on the GPU: one stream
    dense.profile_begin()
    nprof = min(args.steps, 3)
    for _ in range(nprof):
        trainer.step(resident)  # eager (not the graph): per-launch events need individual launches
    dense.profile_end.steps = nprof
    prof = dense.profile_end()
    streams.set_enabled(overlap)
    if rank == 0 and args.dump_launches:
        per = len(dense.last_profile) // nprof
        with open(args.dump_launches, "w") as f:
            f.write("# tensor-core launches of one training step (CUDA events, passes=%s): call shape | GFLOP | ms | TFLOP/s\n" % args.precision)
            for tag, fl, t in dense.last_profile[-per:]:
                f.write("%-60s %9.3f %8.4f %8.1f\n" % (tag, fl / 1e9, t, fl / t / 1e9 if t > 0 else 0))

    in_sync = None
    if world > 1:
        # replicas must still hold identical parameters after the run (weights are never re-broadcast)
        chk = torch.stack([trainer.flat_p.double().sum(), trainer.flat_p.double().pow(2).sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured" if peaks else "fallback"
    # peak of the tensor-core instruction kind the contractions are issued as: kind::f16 (bf16 operands) runs at the
    # measured cuBLAS bf16 rate, kind::tf32 at half of it
    tc_peak = bf16 if args.precision == "bf16x3" else bf16 / 2.0
    passes = 1 if args.precision == "tf32" else 3
    # DRAM traffic of the dominant kernel family per launch, from the committed ncu capture of this command
    # (profiles/gemm_traffic_r1j.json <- profiles/launches_r1j_bench_step.csv); only valid for the default workload
    traffic, traffic_src = None, None
    if args.config == 2 and args.precision == "bf16x3":
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic_r1j.json")))
            traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
        except Exception:  # noqa: BLE001
            pass
    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)
    line = {
        "metric": "train-step images/sec", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": workload_config(world), "clocks": clocks,
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(kernels), "replicas_in_sync": in_sync,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (all conv/GEMM launches of a step)",
                     "achieved": prof["tflops"], "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": prof["tflops"] / tc_peak,
                     "peak_source": "%s bf16_tflops_sustained%s" % (
                         peak_src, "" if args.precision == "bf16x3" else " / 2 (TF32 rate)"),
                     "achieved_note": "algorithmic fp32 FLOPs (2*M*N*K per contraction); each is issued as %d "
                                      "tensor-core products" % passes,
                     "tensor_pipe_tflops_incl_passes": prof["tflops"] * passes,
                     "tensor_pipe_frac": prof["tflops"] * passes / tc_peak,
                     "gemm_ms_per_step": prof["ms_per_step"], "gemm_launches_per_step": prof["launches_per_step"],
                     "share_of_step": prof["ms_per_step"] / (ms / args.steps),
                     "share_note": "per-launch times are measured with every launch alone on one stream (eager); in "
                                   "the timed graph the weight-gradient launches overlap the data-gradient chain",
                     "flops_per_launch": prof["flops"] / max(1, prof["launches"]),
                     "traffic": traffic, "traffic_unit": "bytes/launch (dram read+write, ncu)",
                     "traffic_source": traffic_src},
    }
    line["chamfer"] = chamfer_report(peaks)
    if not args.no_cpu_baseline:
        sec, threads = cpu_reference_steps(8, 2, 1)
        line["cpu_baseline"] = {"value": 8 / sec, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": "batch 8 of the same workload (fwd+bwd+Adam), 2 timed steps, oracle/nets.py"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "tf32x3", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="2 = BASELINE configs[1] (default, the headline); 3 = configs[2] (B=256, contact); "
                         "4 = configs[3] (128 per GPU, full loss stack, meant for --gpus 8)")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch tensor-core profile of one step here")
    args = ap.parse_args()
    if args.config == 3:
        use_config3()
    elif args.config == 4:
        use_config4()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
        # No dist.barrier() / destroy_process_group() here: with the captured step graphs (which hold NCCL kernels)
        # still alive the teardown was observed to hang at N=2 (round 1k); the processes simply exit.


if __name__ == "__main__":
    main()

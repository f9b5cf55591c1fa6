#!/usr/bin/env python
"""Benchmark of the obman_train hot path on B200 (driver contract).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

One "step" = one training step (HandNet forward + backward + Adam) on a synthetic batch.  Workloads are
BASELINE.json's configs (0-based indices as in the file; ``--config`` takes the index + 1):

    --config 1  configs[0]  B=1 hand-only forward (image_demo.py path): latency only
    --config 2  configs[1]  B=64, shared encoder, ico-3 (642 pts), Chamfer(600 GT), Mano + Atlas losses
    --config 3  configs[2]  B=256, separate Atlas encoder, ico-4 (2562 pts), Chamfer(2500 GT) + contact_zones
    --config 4  configs[3]  128 per GPU (1024 on 8 GPUs), configs[2] + edge regulariser ("full loss stack")

Default (``--config 0``): configs[2] on one GPU - the configuration the north_star quotes its single-GPU targets
on - and configs[3] under torchrun (N > 1), weak scaling.  At N = 1 the JSON line also carries the other
single-GPU configurations as secondary objects (``secondary``), the eager-PyTorch GPU comparator
(``gpu_eager_baseline``) and the CPU port (``cpu_baseline``).  Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

IMG = 256
PRECISION_NOTES = {
    "bf16x3": "fp32 operands split into bf16 hi+lo, 3 tensor-core products per contraction (fp32-equivalent to ~2^-17), fwd+bwd",
    "tf32x3": "3xTF32 tensor-core passes (fp32-equivalent) fwd+bwd", "tf32": "single-pass TF32"}
BASE_CFG = dict(resnet_version=18, mano_root="synthetic", mano_comps=30, mano_use_shape=True, mano_use_pca=True,
                mano_neurons=[1024, 256], mano_center_idx=0, mano_lambda_verts=0.167, mano_lambda_joints3d=0.167,
                mano_lambda_shape=0.167, mano_lambda_pose_reg=0.167, atlas_lambda=0.167, atlas_final_lambda=0.167,
                atlas_mesh=True, atlas_predict_trans=True, atlas_predict_scale=True, atlas_trans_weight=0.167,
                atlas_scale_weight=0.167, atlas_separate_encoder=False, atlas_ico_divisions=3, atlas_points_nb=600)
CONTACT_CFG = dict(atlas_separate_encoder=True, atlas_ico_divisions=4, atlas_points_nb=2500, contact_lambda=1,
                   collision_lambda=1, contact_zones="zones", contact_mode="dist_tanh", collision_mode="dist_tanh",
                   contact_thresh=10, collision_thresh=20)


class Workload(object):
    """One BASELINE.json configuration: HandNet kwargs, per-GPU batch, number of GT object points, label."""

    def __init__(self, config):
        self.config = config
        self.cfg = dict(BASE_CFG)
        self.hand_only = False
        if config == 1:
            self.batch, self.n_gt = 1, 0
            self.hand_only = True
            self.name = ("BASELINE.json configs[0]: single 256x256 image, hand-only ManoBranch forward "
                         "(image_demo.py path, random-init weights), latency")
        elif config == 2:
            self.batch, self.n_gt = 64, 600
            self.name = ("BASELINE.json configs[1]: train step fwd+bwd+Adam, ResNet-18 + ManoLayer(778v) + "
                         "AtlasNet(642 pts) + Chamfer(600 GT) + Mano/Atlas losses, 256x256")
        elif config == 3:
            self.batch, self.n_gt = 256, 2500
            self.cfg.update(CONTACT_CFG)
            self.name = ("BASELINE.json configs[2]: train step fwd+bwd+Adam, 2x ResNet-18 (separate Atlas encoder) + "
                         "ManoLayer + AtlasNet(ico-4, 2562 pts) + Chamfer(2500 GT) + contact_zones loss, 256x256")
        elif config == 4:
            self.batch, self.n_gt = 128, 2500
            self.cfg.update(CONTACT_CFG)
            self.cfg.update(atlas_lambda_regul_edges=0.1)
            self.name = ("BASELINE.json configs[3]: per-GPU batch 128 (1024 on 8 GPUs), full loss stack: 2x ResNet-18 + "
                         "ManoLayer + AtlasNet(ico-4) + Chamfer(2500 GT) + contact_zones + edge regulariser, 256x256")
        else:
            raise ValueError("unknown config %r" % (config,))
        self.n_obj = 642 if self.cfg["atlas_ico_divisions"] == 3 else 2562

    def sample(self, B, seed):
        """images U(0,1)-0.5; joints N(0,40^2); object points N(0,40^2)+30 (mm); hand vertices
        N(0,40^2), or MANO-template-shaped hands when the contact loss is on (so its masks are non-trivial)."""
        g = torch.Generator().manual_seed(seed)
        out = {"images": torch.rand(B, 3, IMG, IMG, generator=g) - 0.5,
               "sides": ["right" if i % 2 == 0 else "left" for i in range(B)], "root": "wrist",
               "joints3d": torch.randn(B, 21, 3, generator=g) * 40}
        if self.hand_only:
            out["sides"] = ["left"] * B
            out["joints3d"] = torch.ones(B, 21, 3)
            return out
        if self.cfg.get("contact_lambda"):
            from obman_train_b200.assets import load_contacts
            verts, _ = load_contacts()
            hand = torch.tensor(verts * 1000, dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
            out["verts3d"] = hand + torch.randn(B, 778, 3, generator=g) * 5
        else:
            out["verts3d"] = torch.randn(B, 778, 3, generator=g) * 40
        out["objpoints3d"] = torch.randn(B, self.n_gt, 3, generator=g) * 40 + 30
        return out

    def describe(self, n_gpus, precision):
        return {"workload": self.name, "per_gpu_batch": self.batch, "global_batch": self.batch * n_gpus,
                "parallelism": "dp%d" % n_gpus, "precision": PRECISION_NOTES[precision],
                "l2_policy": "activations per step (> 3 GB) exceed the 126 MB L2; no explicit flush",
                "launch": "whole step replayed as one CUDA graph (--no-graph for eager launches); weight-gradient / "
                          "BN-gradient / weight-folding kernels on auxiliary captured streams (OBMAN_OVERLAP=0 disables)"}


def default_config(world):
    return 3 if world <= 1 else 4


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
# baselines: the reference algorithm (oracle port; the reference itself cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------------------
def oracle_train_steps(wl, batch, steps, warmup, device="cpu", threads=None):
    """Train step (fwd + bwd + torch.optim.Adam) of the reference algorithm: oracle/nets.py, pinned against the
    reference's own files by tests/test_oracle_vs_reference.py.  ``device`` "cpu" = the CPU arm, "cuda" = the
    eager-PyTorch (cuDNN / cuBLAS / ATen) comparator.  Returns seconds per step."""
    from oracle import nets
    from obman_train_b200.networks.handnet import HandNet
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = HandNet(**wl.cfg).eval()
    state = {k: v.detach().clone().to(device) for k, v in model.state_dict().items()}
    leaves = []
    for k, v in state.items():
        if v.is_floating_point() and "running_" not in k and "th_" not in k and ".fc." not in k:
            v.requires_grad_(True)
            leaves.append(v)
    opt = torch.optim.Adam(leaves, lr=1e-4)
    tables = {s: {k: v.detach().to(device) for k, v in getattr(model.mano_branch, "mano_layer_" + s).named_buffers()
                  if k != "th_faces"} for s in ("right", "left")}
    grid, faces = model.atlas_branch.test_verts.to(device), model.atlas_branch.test_faces
    from obman_train_b200.assets import load_contacts
    zones = load_contacts()[1]
    sample = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in wl.sample(batch, 0).items()}
    times = []
    for it in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad()
        total, _, _ = nets.handnet_forward(state, wl.cfg, sample, tables, grid, faces, zones)
        total.backward()
        opt.step()
        if device != "cpu":
            torch.cuda.synchronize()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def cpu_threads():
    # all host cores up to 32 threads: beyond that torch's CPU conv/bmm kernels slow down on this workload
    # (measured on the 128-core GPU box: 128 threads gave 0.26 img/s where 8 threads give ~19 img/s)
    return min(os.cpu_count() or 1, int(os.environ.get("OBMAN_BENCH_CPU_THREADS", "32")))


def run_reference(args):
    """--impl reference: the reference algorithm (kind "port": oracle/nets.py, torch CPU) on the host cores, on the
    same workload as the CUDA arm would run at this N, EXACTLY --steps timed steps after --warmup warm-up steps; each
    step is a bounded sample (2-8 images, calibrated on one untimed step, or ``OBMAN_BENCH_CPU_BATCH``) so the whole run ends
    within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = Workload(args.config or default_config(world))
    if wl.hand_only:
        raise SystemExit("--impl reference: use scripts/config1_cpu.py for configs[0]")
    threads = cpu_threads()
    if "OBMAN_BENCH_CPU_BATCH" in os.environ:
        batch = int(os.environ["OBMAN_BENCH_CPU_BATCH"])
    else:
        # calibration (not part of the reported steps): one step at batch 2, then the largest batch in [2, 8] that keeps
        # the requested --warmup + --steps within ~200 s on this host
        t2 = oracle_train_steps(wl, 2, 1, 0, "cpu", threads)
        batch = max(2, min(8, int(2 * 200.0 / (max(1, args.steps + args.warmup) * t2))))
    sec = oracle_train_steps(wl, batch, args.steps, args.warmup, "cpu", threads)
    val = batch / sec
    line = {
        "impl": "reference", "metric": "train-step images/sec", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.describe(max(1, args.gpus), args.precision),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "batch": batch,
                         "sample": "batch %d of the same workload per step, fwd+bwd+Adam, %d timed steps after %d "
                                   "warm-up, oracle/nets.py on torch CPU" % (batch, args.steps, args.warmup)},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------
def chamfer_report(peaks, wl):
    """Second half of BASELINE.json's metric ("Chamfer kernel HBM GB/s vs peak"): the fused nearest-neighbour
    kernel at the workload's own shapes and at a large sweep point, on algorithmic bytes (20 B per point: 12 read,
    8 written), plus the backward kernel (the one genuinely HBM-bound piece: 16 B read per point + 12 B written per
    predicted point).  The fused forward does 200-3300 flop per algorithmic byte, so its bound is the FP32 pipe (128 lanes
    per SM and clock; reported as well, next to the issue-slot utilisation); the full sweep is profiles/chamfer_sweep_*.json (scripts/bench_chamfer.py)."""
    from obman_train_b200 import functional as Fb
    hbm = peaks.get("hbm_gbs", 6650.0)
    issue_peak = 148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6
    out = {"unit": "GB/s on algorithmic bytes", "hbm_peak_gbs": hbm}
    g = torch.Generator(device="cuda").manual_seed(7)

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for tag, (b, n, m, reps) in {"workload": (wl.batch, wl.n_obj, wl.n_gt, 50),
                                 "sweep_2048x10000": (2048, 10000, 10000, 2)}.items():
        x = torch.randn(b, n, 3, device="cuda", generator=g) * 60
        y = torch.randn(b, m, 3, device="cuda", generator=g) * 60
        ms = timeit(lambda: Fb.nearest_neighbours(x, y), reps)
        gbs = 20.0 * b * (n + m) / ms / 1e6
        out[tag] = {"B": b, "N": n, "M": m, "ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm,
                    "pairs_per_s": 2.0 * b * n * m / ms * 1e3,
                    # per pair: 3 subtractions, 1 multiply, 2 fused multiply-adds = 6 fp32 lane operations (issued two
                    # at a time as FADD2 / FMUL2 / FFMA2 by nn_packed_kernel) and ~4.75 issue slots in all
                    "frac_of_fp32_pipe_peak": 2.0 * b * n * m * 6.0 / (ms * 1e-3) / issue_peak,
                    "frac_of_issue_peak": 2.0 * b * n * m * 4.75 / (ms * 1e-3) / issue_peak}
        # backward: gradient of mean_b(loss_1 + loss_2) w.r.t. the predicted cloud
        xr = x.clone().requires_grad_(True)
        l1, l2 = Fb.chamfer(xr, y)
        gl = torch.full_like(l1, 1.0 / b)
        ms_b = timeit(lambda: torch.autograd.grad((l1, l2), xr, (gl, gl), retain_graph=True), max(2, reps // 2))
        bytes_b = 16.0 * b * (n + m) + 12.0 * b * n
        out[tag]["bwd"] = {"ms": ms_b, "achieved_gbs": bytes_b / ms_b / 1e6, "frac_of_hbm_peak": bytes_b / ms_b / 1e6 / hbm,
                           "note": "includes the autograd dispatch of one backward call"}
        del x, y, xr, l1, l2
    return out


def traffic_record(precision):
    """DRAM bytes per tensor-core launch from the committed ncu launch list of this command - only reported while
    the kernel sources still hash to the build that was profiled (the file records the hash), else null."""
    try:
        path = os.path.join(ROOT, "profiles", "gemm_traffic_current.json")
        tr = json.load(open(path))
        if tr.get("precision", "bf16x3") != precision:
            return None, None
        if tr.get("kernel_source_sha256") != kernel_source_hash():
            return None, "stale: " + os.path.basename(path) + " was captured on a different build of csrc/"
        return tr, tr.get("source")
    except Exception:  # noqa: BLE001
        return None, None


def kernel_source_hash():
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "obman_train_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()


def measure_train(wl, args, world, rank, local, full):
    """Times ``args.steps`` training steps of workload ``wl`` on this rank's GPU.  ``full`` adds the e2e pass, the
    per-launch roofline pass and the clocks sampling (the headline workload); without it only the device-timed
    value is measured (secondary workloads)."""
    import torch.distributed as dist
    from obman_train_b200 import _lib, dense, streams
    from obman_train_b200.networks.handnet import HandNet
    from obman_train_b200.trainer import FlatAdamTrainer, PinnedFeeder
    from obman_train_b200.queries import TransQueries, BaseQueries

    torch.manual_seed(0)  # identical replicas on every rank
    model = HandNet(**wl.cfg).eval().cuda()
    trainer = FlatAdamTrainer(model, lr=1e-4, world_size=world)
    B = wl.batch
    host = wl.sample(B, 1000 + rank)
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

    warm = max(3, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0 and full:
        sampler.start()  # sampled through warm-up and the timed region (both under load)
    for _ in range(warm):
        step_resident()
    ms = timed(args.steps, step_resident)
    k0 = _lib.kernel_count
    trainer.step(resident)  # replays issue no Python-side calls: count the kernels of one eager step
    kernels = (_lib.kernel_count - k0) * args.steps
    res = {"B": B, "ms": ms, "ms_per_step": ms / args.steps, "value": B * world * args.steps / (ms / 1e3),
           "kernels": int(kernels), "warmup": warm, "h2d": h2d}
    if not full:
        trainer.release_graph()
        return res
    res["clocks"] = sampler.stop() if rank == 0 else None

    # ---- e2e: inputs travel from pinned host memory every step, the loss is read back by the host ----
    loss_host = torch.zeros(1).pin_memory()
    tensor_keys = [(TransQueries.images, "images"), (TransQueries.joints3d, "joints3d"),
                   (TransQueries.verts3d, "verts3d"), (TransQueries.objpoints3d, "objpoints3d")]
    feeder = None
    if use_graph:
        feeder = PinnedFeeder(trainer)
        pinned_q = {q: pinned[name] for q, name in tensor_keys}
        feeder.prefetch(pinned_q)  # primes the pipeline: step i runs while step i+1's batch is in flight

    def step_e2e():
        if use_graph:
            loss = feeder.step(next_host_sample=pinned_q)
        else:
            loss = trainer.step(to_device(pinned))
        loss_host.copy_(loss.detach(), non_blocking=False)  # the user's read of the step result

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(args.steps, step_e2e)
    res["e2e"] = {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps}

    # ---- roofline of the dominant kernel family: per-launch CUDA events + algorithmic FLOPs ----
    overlap = streams.set_enabled(False)  # per-launch timing needs every launch alone on the GPU: one stream
    dense.profile_begin()
    nprof = min(args.steps, 3)
    for _ in range(nprof):
        trainer.step(resident)  # eager (not the graph): per-launch events need individual launches
    dense.profile_end.steps = nprof
    res["prof"] = dense.profile_end()
    streams.set_enabled(overlap)
    if rank == 0 and args.dump_launches:
        per = len(dense.last_profile) // nprof
        with open(args.dump_launches, "w") as f:
            f.write("# tensor-core launches of one training step (CUDA events, passes=%s): call shape | GFLOP | ms | TFLOP/s\n" % args.precision)
            for tag, fl, t in dense.last_profile[-per:]:
                f.write("%-60s %9.3f %8.4f %8.1f\n" % (tag, fl / 1e9, t, fl / t / 1e9 if t > 0 else 0))
    res["in_sync"] = None
    if world > 1:
        # replicas must still hold identical parameters after the run (weights are never re-broadcast)
        chk = torch.stack([trainer.flat_p.double().sum(), trainer.flat_p.double().pow(2).sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res["in_sync"] = bool(torch.equal(lo, hi))
    if world > 1:
        # communication evidence (no nsys in this image): the exchange alone, and the same captured step with the
        # exchange switched off - the difference is the exposed (non-overlapped) communication time.  Runs last: the
        # replicas diverge once gradients are no longer summed.
        nbytes = trainer.flat_g.numel() * 4
        for _ in range(3):
            dist.all_reduce(trainer.flat_g)
        ms_ar = timed(10, lambda: dist.all_reduce(trainer.flat_g)) / 10
        trainer.release_graph()
        trainer.skip_allreduce = True
        if use_graph:
            trainer.capture(resident)
        for _ in range(3):
            step_resident()
        ms_nc = timed(args.steps, step_resident) / args.steps
        trainer.skip_allreduce = False
        exposed = max(0.0, res["ms_per_step"] - ms_nc)
        res["comm"] = {"allreduce_bytes": nbytes, "allreduce_alone_ms": ms_ar,
                       "allreduce_bus_gbs": 2.0 * (world - 1) / world * nbytes / ms_ar / 1e6,
                       "step_ms": res["ms_per_step"], "step_ms_without_exchange": ms_nc, "exposed_ms": exposed,
                       "overlap_frac": max(0.0, 1.0 - exposed / ms_ar) if ms_ar > 0 else None,
                       "how": "CUDA events, max over ranks; 'without exchange' = the same captured step with the "
                              "all-reduce launches removed (replicas no longer in sync afterwards)"}
    res["trainer"] = trainer
    return res


def measure_latency(wl, steps):
    """configs[0]: B=1 hand-only forward (no_loss=True, eval), eager launches and as a CUDA graph: ms per image."""
    from obman_train_b200.networks.handnet import HandNet
    from obman_train_b200.queries import TransQueries, BaseQueries
    torch.manual_seed(0)
    model = HandNet(**wl.cfg).eval().cuda()
    s = wl.sample(1, 0)
    sample = {TransQueries.images: s["images"].cuda(), BaseQueries.sides: s["sides"], "root": s["root"],
              TransQueries.joints3d: s["joints3d"].cuda()}

    def fwd():
        with torch.no_grad():
            return model.forward(sample, no_loss=True)[1]["verts"]

    def timeit(fn, n):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    n = max(20, steps)
    eager = timeit(fwd, n)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fwd()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        fwd()
    replay = timeit(graph.replay, n)
    del graph
    return {"workload": wl.name, "ms_eager": eager, "ms_graph": replay, "images_per_s_graph": 1e3 / replay}


def run_b200(args):
    import torch.distributed as dist
    from obman_train_b200 import _lib, dense

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
    wl = Workload(args.config or default_config(world))
    scaling = "weak"
    if args.strong:
        wl = Workload(4)
        wl.batch = 1024 // world
        wl.name += " [strong scaling: global batch 1024, %d per GPU]" % wl.batch
        scaling = "strong"
    if wl.hand_only:
        if rank == 0:
            print(json.dumps({"metric": "forward latency", "unit": "ms", "n_gpus": 1, "higher_is_better": False,
                              "data": "synthetic", **measure_latency(wl, args.steps)}))
        return None

    res = measure_train(wl, args, world, rank, local, full=True)
    trainer = res.pop("trainer")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    if rank != 0:
        return trainer
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured" if peaks else "fallback"
    # peak of the tensor-core instruction kind the contractions are issued as: kind::f16 (bf16 operands) runs at the
    # measured cuBLAS bf16 rate, kind::tf32 at half of it
    tc_peak = bf16 if args.precision == "bf16x3" else bf16 / 2.0
    passes = 1 if args.precision == "tf32" else 3
    prof = res["prof"]
    tr, traffic_src = (traffic_record(args.precision) if (args.config or default_config(world)) == 3 else (None, None))
    line = {
        "metric": "train-step images/sec", "value": res["value"], "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": res["warmup"], "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": wl.describe(world, args.precision), "clocks": res["clocks"],
        "e2e": res["e2e"], "gpu_launches": res["kernels"], "replicas_in_sync": res["in_sync"],
        "comm": res.get("comm"),
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel / wgrad_bf16_kernel (all conv/GEMM launches of a step)",
                     "achieved": prof["tflops"], "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": prof["tflops"] / tc_peak,
                     "peak_source": "%s bf16_tflops_sustained%s" % (
                         peak_src, "" if args.precision == "bf16x3" else " / 2 (TF32 rate)"),
                     "achieved_note": "algorithmic fp32 FLOPs (2*M*N*K per contraction); each is issued as %d "
                                      "tensor-core products" % passes,
                     "tensor_pipe_tflops_incl_passes": prof["tflops"] * passes,
                     "tensor_pipe_frac": prof["tflops"] * passes / tc_peak,
                     "gemm_ms_per_step": prof["ms_per_step"], "gemm_launches_per_step": prof["launches_per_step"],
                     "share_of_step": prof["ms_per_step"] / res["ms_per_step"],
                     "share_note": "per-launch times are measured with every launch alone on one stream (eager); in "
                                   "the timed graph the weight-gradient launches overlap the data-gradient chain",
                     "flops_per_launch": prof["flops"] / max(1, prof["launches"]),
                     "traffic": tr["dram_bytes_per_launch"] if tr else None,
                     "traffic_unit": "bytes/launch (dram read+write, ncu)", "traffic_source": traffic_src},
    }
    trainer.release_graph()
    del trainer
    torch.cuda.empty_cache()
    line["chamfer"] = chamfer_report(peaks, wl)
    if world == 1 and not args.no_secondary:
        # the other single-GPU configurations of BASELINE.json, device-timed the same way (graph replay, CUDA events)
        sec = {}
        sargs = argparse.Namespace(**vars(args))
        sargs.steps = min(args.steps, 10)
        for cid, key in ((2, "configs[1]"), (4, "configs[3] shard (128 per GPU on one GPU: the weak-scaling base of the N>1 runs)")):
            if cid == wl.config:
                continue
            r = measure_train(Workload(cid), sargs, 1, 0, local, full=False)
            sec[key] = {"workload": Workload(cid).name, "per_gpu_batch": r["B"], "value": r["value"],
                        "unit": "images/s", "ms_per_step": r["ms_per_step"], "steps": sargs.steps}
            torch.cuda.empty_cache()
        sec["configs[0]"] = measure_latency(Workload(1), args.steps)
        line["secondary"] = sec
    if world == 1 and not args.no_gpu_eager:
        # practical "beat this" number: the reference algorithm in eager PyTorch (cuDNN / cuBLAS /
        # ATen, TF32 convolutions = PyTorch's default) on the same GPU; bounded batch - its ray-casting temporaries
        # are (B, 778*F, 3) tensors
        torch.cuda.empty_cache()
        eb = int(os.environ.get("OBMAN_BENCH_EAGER_BATCH", "32" if wl.cfg.get("contact_lambda") else str(wl.batch)))
        try:
            s = oracle_train_steps(wl, eb, 3, 2, "cuda")
            line["gpu_eager_baseline"] = {"value": eb / s, "unit": "images/s", "ms_per_step": s * 1e3, "batch": eb,
                                          "kind": "port", "what": "oracle/nets.py on CUDA tensors: eager PyTorch "
                                          "(cuDNN TF32 convolutions, cuBLAS fp32, ATen), fwd+bwd+torch.optim.Adam, "
                                          "same workload at batch %d, 3 timed steps" % eb}
        except Exception as e:  # noqa: BLE001
            line["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        cb = int(os.environ.get("OBMAN_BENCH_CPU_BATCH", "4"))
        threads = cpu_threads()
        s = oracle_train_steps(wl, cb, 2, 1, "cpu", threads)
        line["cpu_baseline"] = {"value": cb / s, "unit": "images/s", "cores": threads, "kind": "port", "batch": cb,
                                "sample": "batch %d of the same workload (fwd+bwd+Adam), 2 timed steps after 1 warm-up, "
                                          "oracle/nets.py on torch CPU" % cb}
    print(json.dumps(line))
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "tf32x3", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations (N = 1)")
    ap.add_argument("--quick", action="store_true", help="headline numbers only: no baselines, no secondary workloads")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configs index + 1; 0 (default) = configs[2] on one GPU, configs[3] under torchrun")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch tensor-core profile of one step here")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling of configs[3]: global batch fixed at 1024, per-GPU batch = 1024 / N")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_gpu_eager = args.no_secondary = True
    if args.impl == "reference":
        run_reference(args)
        return
    keep = run_b200(args)  # noqa: F841  (ranks > 0 keep their trainer alive until every rank is done)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        teardown_distributed(keep)


def teardown_distributed(trainer):
    """Orderly exit at N > 1: every rank has finished its last collective (barrier inside the timed helper), the
    captured graphs (which hold NCCL kernels) are destroyed FIRST, then the process group.  Destroying the group with
    live graphs hung at N=2 in round 1."""
    import torch.distributed as dist
    sys.stdout.flush()
    # belt and braces: the result line is out; if the teardown stalls anyway, leave without it after 30 s
    wd = threading.Timer(30.0, lambda: os._exit(0))
    wd.daemon = True
    wd.start()
    torch.cuda.synchronize()
    if trainer is not None:
        trainer.release_graph()
    torch.cuda.synchronize()
    try:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    main()

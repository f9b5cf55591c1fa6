"""Data-parallel training step around the drop-in HandNet.

Replaces the reference's single-process ``torch.nn.DataParallel`` wrap + per-tensor Adam
(/root/reference/traineval.py:113-116,130; /root/reference/mano_train/netscripts/epochpass3d.py:80-91) with:

* one process per GPU, weights replicated once (never re-broadcast per step);
* all trainable parameters and their gradients living in two flat fp32 buffers, so the gradient exchange
  is ONE ``ncclAllReduce`` over NVLink per step and the optimiser is ONE fused Adam kernel
  (``obman_adam_step``) instead of ~130 per-tensor updates;
* no ``.item()`` syncs inside the step: the loss stays on the device until the caller reads it.
"""
import gc

import torch
import torch.distributed as dist

from ._lib import call, ptr, stream_ptr


def complement_ranges(done, total):
    """[lo, hi) ranges of [0, total) not covered by the (possibly unordered, non-overlapping) ranges in ``done``."""
    out, pos = [], 0
    for lo, hi in sorted(done):
        if lo > pos:
            out.append((pos, lo))
        pos = max(pos, hi)
    if pos < total:
        out.append((pos, total))
    return out


class _GradSink(object):
    """Hands the encoder views of the flat gradient buffer (encoder.set_grad_sink) and starts the all-reduce of a
    finished range while the rest of the backward pass is still running."""

    def __init__(self, trainer):
        self.t = trainer
        self.by_ptr = {}
        for p, off, view in zip(trainer.params, trainer.offsets, trainer.grad_views):
            self.by_ptr[p.data_ptr()] = (off, p.numel(), view)
        self.written = set()
        self.reduced = []
        self.comm_stream = None

    def begin(self):
        self.written.clear()
        self.reduced = []

    def view(self, data_ptr):
        e = self.by_ptr.get(data_ptr)
        return None if e is None else e[2]

    def wrote(self, *ptrs):
        self.written.update(ptrs)

    def stage_done(self, first_ptr, last_ptr, producer_stream):
        t = self.t
        if t.world_size <= 1 or not t.overlap_allreduce:
            return
        a, b = self.by_ptr.get(first_ptr), self.by_ptr.get(last_ptr)
        if a is None or b is None:
            return
        lo, hi = a[0], b[0] + b[1]
        hi = min(t.numel, (hi + 63) // 64 * 64)   # up to the slot boundary (padding entries are zero and stay zero)
        # every parameter slot inside the range must have been written by the encoder
        for ptr_, (off, n, _) in self.by_ptr.items():
            if lo <= off < hi and ptr_ not in self.written:
                return
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        src = producer_stream if producer_stream is not None else torch.cuda.current_stream()
        self.comm_stream.wait_stream(src)
        with torch.cuda.stream(self.comm_stream):
            if not t.skip_allreduce:
                dist.all_reduce(t.flat_g[lo:hi], op=dist.ReduceOp.SUM)
            if t.adam_per_range:
                # the optimiser step of this range follows its all-reduce on the communication stream, under the rest
                # of the backward pass (nothing later in the step reads these parameters: the forward pass folded its
                # own copies of the weights, and their gradients are complete)
                t.adam_range(lo, hi)
        self.reduced.append((lo, hi))


class FlatAdamTrainer(object):
    """``step(sample)`` = zero grads -> HandNet.forward -> backward -> [all-reduce] -> fused Adam."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, world_size=1,
                 direct_grads=True, overlap_allreduce=True):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.world_size = world_size
        # ``fc`` of the encoders is constructed and checkpointed by the reference but never runs
        # (resnet.py:123,184-186): it gets no gradient and is left out of the exchange.
        params = [(n, p) for n, p in model.named_parameters()
                  if p.requires_grad and ".fc." not in "." + n + "."]
        self.names = [n for n, _ in params]
        # every parameter starts on a 256-byte boundary (TMA operands must be 16-byte aligned)
        align = 64
        offsets, total = [], 0
        for _, p in params:
            offsets.append(total)
            total += (p.numel() + align - 1) // align * align
        dev = params[0][1].device
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad_views = []
        for (_, p), off in zip(params, offsets):
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view_as(p)
            self.grad_views.append(self.flat_g[off:off + n].view_as(p))
            p.grad = self.grad_views[-1]
        self.params = [p for _, p in params]
        self.numel = total
        self.param_numel = sum(p.numel() for _, p in params)
        # direct_grads: the encoder's weight-gradient kernels write into the flat buffer (no gather copy);
        # overlap_allreduce: the layer3 + layer4 range (94 % of an encoder) is all-reduced under the rest of backward
        self.direct_grads = direct_grads and dev.type == "cuda"
        self.overlap_allreduce = overlap_allreduce
        # Adam per all-reduced range (bucketed optimiser): the early range is updated on the communication stream as
        # soon as it is reduced, the rest after the backward pass; False = one Adam launch over the whole buffer
        self.adam_per_range = True
        self._skipped = []            # parameter indices without a gradient in the current step
        self._ever_updated = set()    # parameter indices that have received a gradient at least once
        self.skip_allreduce = False   # measurement knob (bench.py: step time without the exchange)
        self.offsets = offsets
        self._sink = _GradSink(self) if self.direct_grads else None
        self.step_count = 0
        # {step number, lr multiplier} in device memory: a captured graph reads both at replay time
        self._hyper_dev = torch.tensor([0.0, 1.0], device=dev, dtype=torch.float32)
        self._step_view, self._scale_view = self._hyper_dev[0:1], self._hyper_dev[1:2]
        self.lr_scale = 1.0
        self._graph = None
        # index of every optimised parameter in the reference's optimizer (traineval.py:105-116: Adam over
        # filter(requires_grad, model.parameters()), which still counts the never-used ``fc`` parameters)
        order = {id(p): i for i, p in enumerate(q for q in model.parameters() if q.requires_grad)}
        self.optim_index = [order[id(p)] for p in self.params]
        self.optim_len = len(order)

    def zero_grad(self):
        self.flat_g.zero_()

    def step(self, sample, return_all=False):
        """Returns the (device) total loss of this rank's shard; with ``return_all`` the model's full
        ``(total_loss, results, losses)`` triple (what epochpass3d.py:80-82 unpacks)."""
        # Gradients are produced by autograd as fresh tensors (p.grad = None, so AccumulateGrad adopts them without
        # an add kernel per parameter) and gathered into the flat buffer with one fused multi-tensor copy.
        for p in self.params:
            p.grad = None
        self._begin_step()
        if self._sink is not None:
            from . import encoder
            self._sink.begin()
            prev = encoder.set_grad_sink(self._sink)
            try:
                loss, results, losses = self.model.forward(sample)
                loss.backward()
            finally:
                encoder.set_grad_sink(prev)
        else:
            loss, results, losses = self.model.forward(sample)
            loss.backward()
        self.gather_grads()
        self.reduce_and_update()
        return (loss, results, losses) if return_all else loss

    def gather_grads(self):
        written = self._sink.written if self._sink is not None else ()
        have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
        missing = [v for v, p in zip(self.grad_views, self.params)
                   if p.grad is None and p.data_ptr() not in written]
        # torch.optim.Adam SKIPS a parameter whose .grad is None (no moment decay, no weight decay, no step): remember
        # which slots those are; reduce_and_update puts their parameter / moment values back after the fused kernel
        self._skipped = [i for i, p in enumerate(self.params) if p.grad is None and p.data_ptr() not in written]
        for i, p in enumerate(self.params):
            if p.grad is not None or p.data_ptr() in written:
                self._ever_updated.add(i)
        if missing:
            torch._foreach_zero_(missing)
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(self.grad_views, self.params):
            p.grad = v

    def _begin_step(self):
        """The step number enters the Adam kernels through device memory; with per-range updates the first of them
        runs in the middle of the backward pass, so the counter advances at the START of a step (a captured graph gets
        it from ``replay``)."""
        self._pending_ranges = None
        if self.flat_p.is_cuda and not torch.cuda.is_current_stream_capturing():
            self.step_count += 1
            self._push_hyper()
        self._stepped = True

    def all_reduce_grads(self):
        """The only data-path collective of the step: a sum all-reduce of the flat gradient buffer - the range that
        went out early on the communication stream (see _GradSink.stage_done) plus one launch per remaining
        contiguous range.  Returns the ranges reduced here."""
        done = list(self._sink.reduced) if self._sink is not None else []
        rest = complement_ranges(done, self.numel)
        if self.world_size > 1 and not self.skip_allreduce:
            for lo, hi in rest:
                dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM)
        if done:  # ranges that went out early on the communication stream (and, per range, their Adam update)
            torch.cuda.current_stream().wait_stream(self._sink.comm_stream)
            self._sink.reduced = []
        return rest, done

    def reduce_and_update(self):
        keep = None
        if self._skipped:
            sl = [slice(self.offsets[i], self.offsets[i] + self.params[i].numel()) for i in self._skipped]
            keep = [(s_, self.flat_p[s_].clone(), self.exp_avg[s_].clone(), self.exp_avg_sq[s_].clone()) for s_ in sl]
        self._reduce_and_update()
        if keep is not None:
            for s_, p_, m_, v_ in keep:
                self.flat_p[s_].copy_(p_)
                self.exp_avg[s_].copy_(m_)
                self.exp_avg_sq[s_].copy_(v_)

    def _reduce_and_update(self):
        rest, done = self.all_reduce_grads()
        if done and self.adam_per_range:
            for lo, hi in rest:
                self.adam_range(lo, hi)
        else:
            self.adam_range(0, self.numel)

    def adam_range(self, lo, hi):
        """Fused Adam on the slots [lo, hi) of the flat buffers (slot boundaries are 256-byte aligned)."""
        call("obman_adam_step", ptr(self.flat_p) + 4 * lo, ptr(self.flat_g) + 4 * lo, ptr(self.exp_avg) + 4 * lo,
             ptr(self.exp_avg_sq) + 4 * lo, hi - lo, float(self.lr), float(self.betas[0]), float(self.betas[1]),
             float(self.eps), float(self.weight_decay), ptr(self._hyper_dev), 1.0 / float(self.world_size), stream_ptr())

    def adam_update(self):
        """One fused Adam step over the whole flat buffer (advances the step counter; ``step`` / ``replay`` use the
        per-range form instead)."""
        if not self.flat_p.is_cuda:
            raise RuntimeError("FlatAdamTrainer.adam_update: the fused Adam kernel needs CUDA buffers (no CPU path)")
        if not torch.cuda.is_current_stream_capturing():
            self.step_count += 1
            self._push_hyper()
        self.adam_range(0, self.numel)

    def _push_hyper(self):
        # fill kernels carry the value as a launch argument: safe with many steps in flight (a pinned staging
        # buffer would be overwritten by the host before the queued copies ran)
        self._step_view.fill_(float(self.step_count))

    def set_lr_scale(self, scale):
        """Multiplier on ``lr`` read from device memory by the Adam kernel (works under graph replay).
        ``StepLR(step_size, gamma)`` of traineval.py:179-182 is ``set_lr_scale(gamma ** (epoch // step_size))``."""
        self.lr_scale = float(scale)
        self._scale_view.fill_(self.lr_scale)

    def step_lr(self, epoch, step_size, gamma):
        self.set_lr_scale(gamma ** (epoch // step_size))

    # ---- CUDA-graph mode: the whole step (forward, backward, all-reduce, Adam) replayed as one graph ----------
    def capture(self, sample, warmup=3):
        """Capture ``step`` on ``sample`` (whose tensors become the static input buffers).  ~1000 kernel
        launches per step otherwise cost more CPU time than the GPU needs to execute them."""
        if self.world_size > 1:
            # NCCL inside a captured graph needs the communicator warmed up outside of capture
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
        # the captured step reads OWNED device buffers: host tensors of the loader's collate (the reference's DataLoader
        # contract) are moved once here, replay() copies every later batch into the same buffers
        for k, v in list(sample.items()):
            if torch.is_tensor(v) and not v.is_cuda:
                sample[k] = v.cuda()
        sides = self._sides_of(sample)
        if sides is not None and "sides_mask" not in sample:
            # graph mode: the left/right pattern enters through a device mask instead of the launch sequence
            sample["sides_mask"] = torch.tensor([s == "right" for s in sides], dtype=torch.bool).cuda()
        # the warm-up steps (allocator / cuBLAS-free lazy initialisation outside of capture) must not train: the
        # parameters, the Adam moments and the step counter are restored afterwards
        saved = (self.flat_p.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_count)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(sample)
        torch.cuda.current_stream().wait_stream(side)
        self.flat_p.copy_(saved[0])
        self.exp_avg.copy_(saved[1])
        self.exp_avg_sq.copy_(saved[2])
        self.step_count = saved[3]
        del saved
        torch.cuda.synchronize()
        self._static_sample = sample
        self._graph = torch.cuda.CUDAGraph()
        steps_before = self.step_count
        # Destructors that run while a capture is in progress can invalidate it: cudaGraphExecDestroy of a dead trainer's
        # graph is "not permitted when stream is capturing" in the default global mode, and the cyclic garbage collector
        # may well decide to free one in the middle of the step (seen once in the test-suite, in the autograd thread).
        # So: collect now, keep the collector off until the capture ends, and restrict the unsafe-call check to the
        # capturing thread.
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                loss, results, losses = self.step(sample, return_all=True)
        finally:
            if gc_was_on:
                gc.enable()
        # Static output tensors of the captured step, rewritten by every replay.  Detached views: holding the autograd
        # graph would keep its AccumulateGrad nodes (and their stream affinity) alive across steps.
        det = lambda d: {k: (v.detach() if torch.is_tensor(v) else v) for k, v in d.items()}  # noqa: E731
        self._static_loss = loss.detach()
        self._static_all = (self._static_loss, det(results), det(losses))
        # the Adam bias corrections are baked into the captured launch; keep them exact by passing the
        # step number through device memory instead (see adam_update)
        self.step_count = steps_before
        return self._graph

    def replay(self, sample=None):
        """Run the captured step; ``sample`` tensors (if given) are copied into the static input buffers."""
        if sample is not None and sample is not self._static_sample:
            if not self.matches_captured(sample):
                raise RuntimeError("FlatAdamTrainer.replay: the batch does not have the tensor keys / shapes of the "
                                   "captured step (a partial last batch?); run it through step() instead")
            for k, v in sample.items():
                if torch.is_tensor(v):
                    self._static_sample[k].copy_(v, non_blocking=True)
            self.set_sides(self._sides_of(sample), sample.get("root"))
        self.step_count += 1
        self._push_hyper()
        self._graph.replay()
        return self._static_loss

    def matches_captured(self, sample):
        """True when ``sample`` can be copied into the captured step's static buffers (same tensor keys and shapes)."""
        if self._graph is None:
            return False
        static = self._static_sample
        keys = {k for k, v in sample.items() if torch.is_tensor(v)}
        skeys = {k for k, v in static.items() if torch.is_tensor(v) and k != "sides_mask"}
        if keys - {"sides_mask"} != skeys:
            return False
        return all(tuple(sample[k].shape) == tuple(static[k].shape) for k in skeys)

    def release_graph(self):
        """Drop the captured step (graph, static inputs / outputs).  Must run before the NCCL process group is
        destroyed: a live graph holds NCCL kernels, and tearing the communicator down underneath it hangs."""
        self._graph = None
        self._static_all = self._static_loss = None
        self._static_sample = None

    # ---- checkpoint interchange with torch.optim.Adam (modelio.load_checkpoint / save_checkpoint) ---------------
    def state_dict(self):
        """Optimizer state in ``torch.optim.Adam.state_dict()`` format, indexed like the reference's optimizer, so a
        checkpoint written here resumes under the reference's traineval.py and vice versa."""
        state = {}
        if self.step_count > 0:
            for i, (p, off, idx) in enumerate(zip(self.params, self.offsets, self.optim_index)):
                if self._ever_updated and i not in self._ever_updated:
                    continue   # never received a gradient: torch.optim.Adam holds no state for it
                n = p.numel()
                state[idx] = {"step": torch.tensor(float(self.step_count)),
                              "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                              "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        group = {"lr": self.lr * self.lr_scale, "betas": tuple(self.betas), "eps": self.eps,
                 "weight_decay": self.weight_decay, "amsgrad": False, "maximize": False, "foreach": None,
                 "capturable": False, "differentiable": False, "fused": None, "initial_lr": self.lr,
                 "params": list(range(self.optim_len))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts ``torch.optim.Adam.state_dict()`` of the reference's optimizer (or this class's own).  Raises
        ``ValueError`` on a parameter-count / shape mismatch like torch does (modelio.load_checkpoint catches it)."""
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != self.optim_len:
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of "
                             "optimizer's group")
        g = groups[0]
        self.lr = float(g.get("initial_lr", g["lr"]))
        self.lr_scale = float(g["lr"]) / self.lr if self.lr != 0 else 1.0
        self.betas, self.eps = tuple(g["betas"]), float(g["eps"])
        self.weight_decay = float(g.get("weight_decay", 0.0))
        if g.get("amsgrad", False):
            raise ValueError("FlatAdamTrainer: amsgrad checkpoints are not supported")
        ids = g["params"]  # torch re-maps by position, not by id value
        pos = {pid: i for i, pid in enumerate(ids)}
        state = {pos[k]: v for k, v in sd["state"].items()}
        steps = set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self._ever_updated = set()
        for i, (p, off, idx) in enumerate(zip(self.params, self.offsets, self.optim_index)):
            st = state.get(idx)
            if st is None:
                continue
            self._ever_updated.add(i)
            n = p.numel()
            if st["exp_avg"].numel() != n:
                raise ValueError("FlatAdamTrainer: optimizer state %d has %d elements, parameter has %d" % (
                    idx, st["exp_avg"].numel(), n))
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        # the fused kernel keeps ONE step counter; the reference's per-parameter counters are all equal except for
        # parameters that never received a gradient (no state at all)
        if len(steps) > 1:
            raise ValueError("FlatAdamTrainer: per-parameter step counts differ (%s)" % sorted(steps))
        self.step_count = steps.pop() if steps else 0
        self._scale_view.fill_(self.lr_scale)

    @staticmethod
    def _sides_of(sample):
        for k, v in sample.items():
            if getattr(k, "value", k) == "sides" and isinstance(v, (list, tuple)):
                return list(v)
        return None

    def set_sides(self, sides, root=None):
        """Left/right pattern of the next replayed batch (list of "left"/"right"); ``root`` must be the captured one."""
        if root is not None and root != self._static_sample.get("root"):
            raise RuntimeError("FlatAdamTrainer: the captured step was built for root={!r}, got {!r}".format(
                self._static_sample.get("root"), root))
        if sides is None or "sides_mask" not in self._static_sample:
            return
        mask = self._static_sample["sides_mask"]
        if len(sides) != mask.numel():
            raise RuntimeError("FlatAdamTrainer: captured batch size {}, got {} sides".format(mask.numel(), len(sides)))
        mask.copy_(torch.tensor([s == "right" for s in sides], dtype=torch.bool))

    def static_outputs(self):
        """``(total_loss, results, losses)`` of the captured step: the same device tensors after every replay."""
        if self._graph is None:
            raise RuntimeError("FlatAdamTrainer.static_outputs: call capture(sample) first")
        return self._static_all

    def grads_are_views(self):
        """Autograd must have accumulated in place into the flat buffer (sanity check for tests)."""
        base = self.flat_g.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


class PinnedFeeder(object):
    """Host -> device input pipeline for the captured step (the replacement for the reference's
    ``DataParallel.scatter`` of the collated batch, /root/reference/mano_train/netscripts/epochpass3d.py:80).

    The next batch travels from PINNED host memory into a staging set on a copy stream while the current step
    runs; at the start of a step the staging set is moved into the graph's static input buffers (a device-to-device
    copy, ~20 us for a 51 MB batch) and the staging set is handed back to the copy stream.  Every step still pays
    one full host->device copy of its own inputs; it is just not serialised with the compute.

        feeder = PinnedFeeder(trainer)          # after trainer.capture(static_sample)
        feeder.prefetch(batch0)
        for batch in batches[1:] + [None]:
            loss = feeder.step(next_host_sample=batch)
    """

    def __init__(self, trainer):
        if trainer._graph is None:
            raise RuntimeError("PinnedFeeder: call trainer.capture(sample) first")
        self.trainer = trainer
        self.static = trainer._static_sample
        self.keys = [k for k, v in self.static.items() if torch.is_tensor(v) and k != "sides_mask"]
        self.staging = {k: torch.empty_like(self.static[k]) for k in self.keys}
        self.copy_stream = torch.cuda.Stream()
        self.ready = torch.cuda.Event()
        self.free = torch.cuda.Event()
        self._staged = False
        self._consumed_once = False
        self.h2d_bytes = sum(self.static[k].numel() * self.static[k].element_size() for k in self.keys)

    def prefetch(self, host_sample):
        """Enqueue the host->device copy of ``host_sample`` (pinned tensors under the static sample's keys)."""
        if self._staged:
            raise RuntimeError("PinnedFeeder.prefetch: the staged batch has not been consumed by step() yet")
        for k in self.keys:
            v = host_sample[k]
            if v.is_cuda or not v.is_pinned():
                raise RuntimeError("PinnedFeeder: input %r must be a pinned host tensor" % (k,))
            if v.shape != self.static[k].shape or v.dtype != self.static[k].dtype:
                raise RuntimeError("PinnedFeeder: input %r has shape %s, the captured step was built for %s" % (
                    k, tuple(v.shape), tuple(self.static[k].shape)))
        if self._consumed_once:
            self.copy_stream.wait_event(self.free)
        with torch.cuda.stream(self.copy_stream):
            for k in self.keys:
                self.staging[k].copy_(host_sample[k], non_blocking=True)
            self.ready.record(self.copy_stream)
        self._staged_sides = FlatAdamTrainer._sides_of(host_sample)
        self._staged = True

    def step(self, next_host_sample=None):
        """Run one captured step on the staged batch; start copying ``next_host_sample`` behind it."""
        if not self._staged:
            raise RuntimeError("PinnedFeeder.step: nothing staged, call prefetch(host_sample) first")
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready)
        for k in self.keys:
            self.static[k].copy_(self.staging[k], non_blocking=True)
        self.free.record(cur)
        self.trainer.set_sides(self._staged_sides)
        self._staged, self._consumed_once = False, True
        if next_host_sample is not None:
            self.prefetch(next_host_sample)
        return self.trainer.replay()

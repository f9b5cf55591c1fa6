"""Data-parallel training step around the drop-in HandNet.

Replaces the reference's single-process ``torch.nn.DataParallel`` wrap + per-tensor Adam
(/root/reference/traineval.py:113-116,130; /root/reference/mano_train/netscripts/epochpass3d.py:80-91) with:

* one process per GPU, weights replicated once (never re-broadcast per step);
* all trainable parameters and their gradients living in two flat fp32 buffers, so the gradient exchange
  is ONE ``ncclAllReduce`` over NVLink per step and the optimiser is ONE fused Adam kernel
  (``obman_adam_step``) instead of ~130 per-tensor updates;
* no ``.item()`` syncs inside the step: the loss stays on the device until the caller reads it.
"""
import torch
import torch.distributed as dist

from ._lib import call, ptr, stream_ptr


class FlatAdamTrainer(object):
    """``step(sample)`` = zero grads -> HandNet.forward -> backward -> [all-reduce] -> fused Adam."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, world_size=1):
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
        self.step_count = 0
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.float32)  # step number for captured graphs
        self._graph = None

    def zero_grad(self):
        self.flat_g.zero_()

    def step(self, sample):
        """Returns the (device) total loss of this rank's shard."""
        # Gradients are produced by autograd as fresh tensors (p.grad = None, so AccumulateGrad adopts them without
        # an add kernel per parameter) and gathered into the flat buffer with one fused multi-tensor copy.
        for p in self.params:
            p.grad = None
        loss, _, _ = self.model.forward(sample)
        loss.backward()
        self.gather_grads()
        self.reduce_and_update()
        return loss

    def gather_grads(self):
        have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
        missing = [v for v, p in zip(self.grad_views, self.params) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(self.grad_views, self.params):
            p.grad = v

    def all_reduce_grads(self):
        """The only data-path collective of the step: one sum all-reduce of the flat gradient buffer."""
        if self.world_size > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)

    def reduce_and_update(self):
        self.all_reduce_grads()
        self.adam_update()

    def adam_update(self):
        if not self.flat_p.is_cuda:
            raise RuntimeError("FlatAdamTrainer.adam_update: the fused Adam kernel needs CUDA buffers (no CPU path)")
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.step_count += 1
            self._step_dev.fill_(float(self.step_count))
        call("obman_adam_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             self.numel, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
             float(self.weight_decay), ptr(self._step_dev), 1.0 / float(self.world_size), stream_ptr())

    # ---- CUDA-graph mode: the whole step (forward, backward, all-reduce, Adam) replayed as one graph ----------
    def capture(self, sample, warmup=3):
        """Capture ``step`` on ``sample`` (whose tensors become the static input buffers).  ~1000 kernel
        launches per step otherwise cost more CPU time than the GPU needs to execute them."""
        if self.world_size > 1:
            # NCCL inside a captured graph needs the communicator warmed up outside of capture
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(sample)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._static_sample = sample
        self._graph = torch.cuda.CUDAGraph()
        steps_before = self.step_count
        with torch.cuda.graph(self._graph):
            self._static_loss = self.step(sample)
        # the Adam bias corrections are baked into the captured launch; keep them exact by passing the
        # step number through device memory instead (see adam_update)
        self.step_count = steps_before
        return self._graph

    def replay(self, sample=None):
        """Run the captured step; ``sample`` tensors (if given) are copied into the static input buffers."""
        if sample is not None and sample is not self._static_sample:
            for k, v in sample.items():
                if torch.is_tensor(v):
                    self._static_sample[k].copy_(v, non_blocking=True)
        self.step_count += 1
        self._step_dev.fill_(float(self.step_count))
        self._graph.replay()
        return self._static_loss

    def grads_are_views(self):
        """Autograd must have accumulated in place into the flat buffer (sanity check for tests)."""
        base = self.flat_g.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

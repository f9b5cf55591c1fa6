"""obman_adam_step (csrc/convaux.cu::adam_kernel) against torch.optim.Adam - the optimiser of the reference's
training loop (/root/reference/traineval.py:113-116) with its StepLR schedule (:179-182): several steps, gradient
scale (1 / world size of the all-reduced sum), learning-rate multiplier read from device memory, weight decay, bias
correction from the device-side step counter, eagerly and under CUDA-graph replay."""
import pytest
import torch

from obman_train_b200._lib import call, ptr, stream_ptr

pytestmark = pytest.mark.gpu

N = 100003          # not a multiple of 4: exercises the scalar tail of the float4 kernel
LR, BETAS, EPS = 1e-3, (0.9, 0.999), 1e-8


def _reference_run(p0, grads, gscale, lr_scales, wd):
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=LR, betas=BETAS, eps=EPS, weight_decay=wd)
    out = []
    for g, s in zip(grads, lr_scales):
        for group in opt.param_groups:
            group["lr"] = LR * s
        p.grad = g * gscale
        opt.step()
        st = opt.state[p]
        out.append((p.detach().clone(), st["exp_avg"].clone(), st["exp_avg_sq"].clone()))
    return out


def _launch(p, g, m, v, hyper, gscale, wd):
    call("obman_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), N, LR, BETAS[0], BETAS[1], EPS, wd, ptr(hyper), gscale,
         stream_ptr())


def _close(a, b, rtol, what):
    # relative to the magnitude of the reference entry, with an absolute floor at the scale of one ulp of the tensor
    err = (a - b).abs()
    bound = rtol * b.abs() + rtol * b.abs().mean()
    assert bool((err <= bound).all()), "%s: max abs err %.3e (bound %.3e)" % (what, err.max().item(), bound.max().item())


@pytest.mark.parametrize("wd", [0.0, 0.01])
@pytest.mark.parametrize("use_graph", [False, True])
def test_adam_kernel_matches_torch_adam_over_six_steps(wd, use_graph):
    gen = torch.Generator(device="cuda").manual_seed(3)
    p0 = torch.randn(N, device="cuda", generator=gen)
    grads = [torch.randn(N, device="cuda", generator=gen) * (10.0 ** (i % 3 - 1)) for i in range(6)]
    gscale = 0.5                                  # 1 / world_size
    lr_scales = [1.0, 1.0, 1.0, 0.5, 0.5, 0.25]   # StepLR(gamma=0.5) stepping twice
    ref = _reference_run(p0, grads, gscale, lr_scales, wd)

    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    g = torch.empty_like(p0)
    hyper = torch.tensor([0.0, 1.0], device="cuda")
    graph = None
    if use_graph:
        g.copy_(grads[0])
        hyper.copy_(torch.tensor([1.0, 1.0]))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _launch(p.clone(), g, m.clone(), v.clone(), hyper, gscale, wd)   # warm-up on scratch copies
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            _launch(p, g, m, v, hyper, gscale, wd)
        # capture does not execute: state is still the initial one
    for i, (gi, s) in enumerate(zip(grads, lr_scales)):
        g.copy_(gi)
        hyper.copy_(torch.tensor([float(i + 1), s]))   # {step number, lr multiplier}: read at replay time
        if use_graph:
            graph.replay()
        else:
            _launch(p, g, m, v, hyper, gscale, wd)
        torch.cuda.synchronize()
        rp, rm, rv = ref[i]
        _close(m, rm, 2e-6, "exp_avg step %d" % (i + 1))
        _close(v, rv, 2e-6, "exp_avg_sq step %d" % (i + 1))
        _close(p, rp, 1e-6, "param step %d" % (i + 1))


def test_flat_trainer_adam_update_matches_torch_adam():
    """The trainer's own wiring of the kernel (flat buffers, step counter, lr scale, grad_scale = 1 / world) on a small
    parameter set, against torch.optim.Adam over the same per-parameter gradients."""
    from obman_train_b200.trainer import FlatAdamTrainer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).cuda()
    twin = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).cuda()
    twin.load_state_dict(model.state_dict())
    opt = torch.optim.Adam(twin.parameters(), lr=3e-4, weight_decay=0.02)
    tr = FlatAdamTrainer(model, lr=3e-4, weight_decay=0.02)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for step in range(5):
        scale = 0.5 ** (step // 2)
        tr.set_lr_scale(scale)
        for group in opt.param_groups:
            group["lr"] = 3e-4 * scale
        for p, q in zip(tr.params, twin.parameters()):
            gr = torch.randn(p.shape, device="cuda", generator=gen)
            p.grad.copy_(gr)
            q.grad = gr.clone()
        tr.adam_update()
        opt.step()
        for p, q in zip(tr.params, twin.parameters()):
            _close(p.data, q.data, 1e-6, "step %d" % step)
    sd = tr.state_dict()
    assert int(sd["state"][0]["step"]) == 5


def test_parameters_without_gradient_are_skipped_like_torch_adam():
    """torch.optim.Adam skips a parameter whose .grad is None (no moment decay, no weight decay, no state); the fused flat
    kernel updates every slot, so the trainer restores such slots afterwards and leaves them out of its state dict
    (with weight decay an unused parameter would otherwise drift)."""
    from obman_train_b200.trainer import FlatAdamTrainer

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.used = torch.nn.Linear(8, 8)
            self.unused = torch.nn.Linear(8, 8)

        def forward(self, sample):
            loss = self.used(sample["x"]).pow(2).sum().view(1)
            return loss, {}, {"total_loss": loss}

    torch.manual_seed(0)
    model, twin = Net().cuda(), Net().cuda()
    twin.load_state_dict(model.state_dict())
    tr = FlatAdamTrainer(model, lr=1e-2, weight_decay=0.1, direct_grads=False)
    opt = torch.optim.Adam(twin.parameters(), lr=1e-2, weight_decay=0.1)
    x = torch.randn(4, 8, device="cuda")
    unused0 = model.unused.weight.detach().clone()
    for _ in range(3):
        tr.step({"x": x})
        opt.zero_grad(set_to_none=True)
        twin({"x": x})[0].backward()
        opt.step()
    assert torch.equal(model.unused.weight.detach(), unused0)
    for p, q in zip(model.parameters(), twin.parameters()):
        assert torch.allclose(p.detach(), q.detach(), rtol=1e-5, atol=1e-7)
    sd = tr.state_dict()
    assert sorted(sd["state"].keys()) == [0, 1]          # used.weight, used.bias only
    assert sorted(opt.state_dict()["state"].keys()) == [0, 1]

"""Diagnostic: encoder gradients - ours vs fp64 oracle vs torch fp32 (CUDA, TF32 off) on the same weights."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nets  # noqa: E402
from tests.test_gpu_encoder import _randomise_bn  # noqa: E402
from obman_train_b200.networks.bases.resnet import resnet18  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def grads_ref(state, images, wts, dtype, device):
    st = {k: v.detach().to(device=device, dtype=dtype if v.is_floating_point() else v.dtype).clone() for k, v in state.items()}
    for k, v in st.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    out = nets.resnet18_features(st, images.to(device=device, dtype=dtype), "base_net", False)
    (out * wts.to(device=device, dtype=dtype)).sum().backward()
    return out.detach(), {k: v.grad.detach() for k, v in st.items() if v.is_floating_point() and v.grad is not None}


def main():
    for B, H, seed in ((2, 64, 0), (4, 128, 0), (3, 96, 0), (2, 64, 7)):
        torch.manual_seed(seed)
        model = resnet18()
        _randomise_bn(model, seed + 1)
        model.eval()
        state = {"base_net." + k: v.detach().clone() for k, v in model.state_dict().items()}
        g = torch.Generator().manual_seed(seed + 2)
        images = torch.rand(B, 3, H, H, generator=g) - 0.5
        wts = torch.randn(B, 512, generator=g)
        f64, g64 = grads_ref(state, images, wts, torch.float64, "cuda")
        f32, g32 = grads_ref(state, images, wts, torch.float32, "cuda")
        model = model.cuda()
        runs = []
        for _ in range(2):
            model.zero_grad()
            feats, _ = model(images.cuda())
            (feats * wts.cuda()).sum().backward()
            runs.append({n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
        def rel(a, b):
            return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
        rows = []
        for n in runs[0]:
            k = "base_net." + n
            rows.append((rel(runs[0][n], g64[k]), rel(g32[k], g64[k]), rel(runs[0][n], g32[k]), rel(runs[1][n], runs[0][n]), n))
        rows.sort(reverse=True)
        print("B=%d H=%d seed=%d  feat err ours %.2e torch32 %.2e" % (B, H, seed, rel(feats, f64), rel(f32, f64)))
        for r in rows[:4]:
            print("   ours-vs-64 %.2e | torch32-vs-64 %.2e | ours-vs-torch32 %.2e | run2-vs-run1 %.2e  %s" % r)


if __name__ == "__main__":
    main()

"""tcgen05 dense kernels (obman_gemm / obman_conv_nhwc / obman_wgrad_nhwc) vs fp64 torch references.
passes=3 (3xTF32, "_p3") and passes=2 (3xBF16, "_b3") must reach fp32-class accuracy, passes=1 plain TF32
accuracy."""
import pytest

pytestmark = pytest.mark.gpu


def _load_cases():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "probe_dense.py")
    spec = importlib.util.spec_from_file_location("probe_dense", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return {k: v for k, v in mod.CASES.items() if not k.startswith("wgradv_")}


CASES = _load_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_dense_case(name):
    res = CASES[name]()
    assert not res["nan"]
    tol = 5e-5 if ("_p3" in name or "_b3" in name) else 3e-3
    assert res["rel"] < tol, res


def test_linear_and_point_decoder_autograd_vs_torch():
    import torch
    import torch.nn.functional as F
    from obman_train_b200 import mlp
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 256, generator=g)
    w = torch.randn(33, 256, generator=g) / 16
    b = torch.randn(33, generator=g)
    xs, ws, bs = [t.cuda().requires_grad_(True) for t in (x, w, b)]
    y = mlp.linear(xs, ws, bs, relu=True)
    wy = torch.randn(37, 33, generator=g)
    (y * wy.cuda()).sum().backward()
    xd, wd, bd = [t.double().requires_grad_(True) for t in (x, w, b)]
    yd = F.relu(F.linear(xd, wd, bd))
    (yd * wy.double()).sum().backward()
    assert (y.detach().cpu().double() - yd.detach()).abs().max() < 1e-5 * yd.abs().max()
    for a, r in ((xs, xd), (ws, wd), (bs, bd)):
        assert (a.grad.cpu().double() - r.grad).abs().max() < 1e-4 * r.grad.abs().max()

"""Diagnostic: intermediate activation gradients of the encoder, ours vs torch fp32 autograd (CUDA)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_encoder import _randomise_bn  # noqa: E402
from obman_train_b200 import encoder  # noqa: E402
from obman_train_b200.networks.bases.resnet import resnet18  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def bn(x, st, p):
    return F.batch_norm(x, st[p + ".running_mean"], st[p + ".running_var"], st[p + ".weight"], st[p + ".bias"], False, 0.0, 1e-5)


def ref(st, images, wts):
    keep = {}
    x = F.conv2d(images, st["conv1.weight"], None, stride=2, padding=3)
    c1 = F.relu(bn(x, st, "bn1")); c1.retain_grad(); keep["c1"] = c1
    p = F.max_pool2d(c1, 3, 2, 1); p.retain_grad(); keep["p"] = p
    x = p
    bidx = 0
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in range(2):
            q = "layer%d.%d." % (li, bi)
            s = stride if bi == 0 else 1
            a = F.relu(bn(F.conv2d(x, st[q + "conv1.weight"], None, stride=s, padding=1), st, q + "bn1"))
            a.retain_grad(); keep["a%d" % bidx] = a
            o = bn(F.conv2d(a, st[q + "conv2.weight"], None, stride=1, padding=1), st, q + "bn2")
            r = bn(F.conv2d(x, st[q + "downsample.0.weight"], None, stride=s), st, q + "downsample.1") if (q + "downsample.0.weight") in st else x
            x = F.relu(o + r); x.retain_grad(); keep["out%d" % bidx] = x
            bidx += 1
    feats = x.mean(3).mean(2)
    (feats * wts).sum().backward()
    return keep


def main():
    for B, H, seed in ((2, 64, 0), (4, 128, 0)):
        torch.manual_seed(seed)
        model = resnet18()
        _randomise_bn(model, seed + 1)
        model.eval().cuda()
        st = {k: v.detach().clone() for k, v in model.state_dict().items()}
        for k, v in st.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
        g = torch.Generator().manual_seed(seed + 2)
        images = (torch.rand(B, 3, H, H, generator=g) - 0.5).cuda()
        wts = torch.randn(B, 512, generator=g).cuda()
        keep = ref(st, images, wts)
        encoder.DEBUG = {}
        feats, _ = model(images)
        (feats * wts).sum().backward()
        dbg = encoder.DEBUG

        def rel(mine_nhwc, t, mask=None):
            r = t.grad
            if mask is not None:
                r = r * (mask > 0)
            m = mine_nhwc.permute(0, 3, 1, 2)
            return ((m.double() - r.double()).norm() / (r.double().norm() + 1e-30)).item()
        print("B=%d H=%d" % (B, H))
        for b in range(7, -1, -1):
            print("  block %d: g_out %.2e   g_a %.2e" % (b, rel(dbg["g_out_%d" % b], keep["out%d" % b], keep["out%d" % b]),
                                                      rel(dbg["g_a_%d" % b], keep["a%d" % b], keep["a%d" % b])))
        print("  g_p %.2e (masked ref)  g_c1 %.2e" % (rel(dbg["g_p"], keep["p"], keep["p"]), rel(dbg["g_c1"], keep["c1"], keep["c1"])))

        def pattern(name, mine_nhwc, t):
            r = (t.grad * (t > 0)).permute(0, 2, 3, 1)
            err = (mine_nhwc - r).abs()
            bad = err > 1e-3 * r.abs().max()
            idx = bad.nonzero()
            print("  %s: %d bad of %d; ref max %.3e; max err %.3e" % (name, idx.shape[0], bad.numel(), r.abs().max().item(), err.max().item()))
            if idx.shape[0]:
                for d, nm in enumerate(("n", "h", "w")):
                    print("     %s values: %s" % (nm, torch.unique(idx[:, d]).tolist()))
                ch = idx[:, 3]
                print("     channel hist by 32: %s" % torch.bincount(ch // 32, minlength=r.shape[3] // 32).tolist())
                k = idx[:8]
                for row in k:
                    n, h, w, c = row.tolist()
                    print("       (%d,%d,%d,%d) mine %.5e ref %.5e act %.5e" % (n, h, w, c, mine_nhwc[n, h, w, c].item(), r[n, h, w, c].item(), t[n, c, h, w].item()))
        if H == 64:
            pattern("g_a_7", dbg["g_a_7"], keep["a7"])
            pattern("g_out_3", dbg["g_out_3"], keep["out3"])
        else:
            pattern("g_out_6", dbg["g_out_6"], keep["out6"])


if __name__ == "__main__":
    main()

"""Per-block parity against PURE fp32 torch modules on identical inputs (north_star gate:
per-parameter gradient cosine >= 0.999): one residual block is shallow enough that bf16 storage
noise does not get chaotically amplified, so this is the strict kernel-correctness gate."""
import pytest
import torch
import torchvision

from sota_imagenet_b200 import modules, ops

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a = a.double().flatten().cpu()
    b = b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("inplanes,planes,stride,down,hw,batch", [
    (256, 64, 1, False, 56, 4),     # identity block, layer1
    (64, 64, 1, True, 56, 4),       # layer1.0: 1x1 stride-1 projection shortcut
    (256, 128, 2, True, 56, 4),     # layer2.0: strided 3x3 + strided 1x1 shortcut
    (1024, 512, 2, True, 14, 16),   # layer4.0
    (2048, 512, 1, False, 7, 16),   # layer4 identity
])
def test_bottleneck_forward_backward(inplanes, planes, stride, down, hw, batch):
    torch.manual_seed(0)
    ds = None
    if down:
        ds = torch.nn.Sequential(torch.nn.Conv2d(inplanes, planes * 4, 1, stride, bias=False),
                                 torch.nn.BatchNorm2d(planes * 4))
    ref = torchvision.models.resnet.Bottleneck(inplanes, planes, stride, ds).cuda().train()
    for n, p in ref.named_parameters():          # non-trivial affine parameters
        if "bn" in n or "downsample.1" in n:
            p.data.uniform_(0.5, 1.5) if n.endswith("weight") else p.data.normal_(0, 0.2)
    blk = modules.Bottleneck(inplanes, planes, stride, downsample=down)
    blk.load_state_dict(ref.state_dict())
    blk = blk.cuda().train()
    # the oracle sees the same bf16-rounded input and filters the kernels see
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if p.dim() == 4:
                p.copy_(p.bfloat16().float())
    x = torch.relu(torch.randn(batch, inplanes, hw, hw, device="cuda")).bfloat16()
    xr = x.float().requires_grad_(True)
    out_ref = ref(xr)
    dy = torch.randn_like(out_ref).bfloat16()
    out_ref.backward(dy.float())
    xb = ops.to_nhwc_bf16(x).requires_grad_(True)
    out = blk(xb)
    out.backward(ops.to_nhwc_bf16(dy))
    torch.cuda.synchronize()
    assert (out.float() - out_ref).norm() / out_ref.norm() < 1e-2
    assert _cos(xb.grad, xr.grad) >= 0.999
    ref_params = dict(ref.named_parameters())
    for name, p in blk.named_parameters():
        c = _cos(p.grad, ref_params[name].grad)
        assert c >= 0.999, (name, c)
    for name, b in blk.named_buffers():
        if "running" in name:
            r = dict(ref.named_buffers())[name]
            assert (b - r).norm() / (r.norm() + 1e-12) < 1e-2, name


def test_stem_and_head_layers():
    torch.manual_seed(1)
    # stem conv 7x7/2 + BN + ReLU + maxpool vs torch
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda()
    stem = modules.StemConv(64, 7, 3)
    stem.weight.data.copy_(conv.weight.data.cpu())
    stem = stem.cuda()
    x = torch.randn(4, 3, 64, 64, device="cuda")
    with torch.no_grad():
        conv.weight.copy_(conv.weight.bfloat16().float())
    y_ref = conv(x.bfloat16().float())
    y = stem(x)
    assert (y.float() - y_ref).norm() / y_ref.norm() < 1e-2
    dy = torch.randn_like(y_ref).bfloat16()
    y_ref.backward(dy.float())
    y.backward(ops.to_nhwc_bf16(dy))
    assert _cos(stem.weight.grad, conv.weight.grad) >= 0.999
    # Linear head
    lin_ref = torch.nn.Linear(2048, 1000).cuda()
    lin = modules.Linear(2048, 1000)
    lin.load_state_dict(lin_ref.state_dict())
    lin = lin.cuda()
    with torch.no_grad():
        lin_ref.weight.copy_(lin_ref.weight.bfloat16().float())
    f = torch.randn(32, 2048, device="cuda").bfloat16()
    fr = f.float().requires_grad_(True)
    o_ref = lin_ref(fr)
    fb = f.view(32, 2048, 1, 1).clone().requires_grad_(True)
    o = lin(fb)
    assert (o.float() - o_ref).norm() / o_ref.norm() < 1e-2
    g = torch.randn_like(o_ref).bfloat16()
    o_ref.backward(g.float())
    o.backward(g)
    assert _cos(lin.weight.grad, lin_ref.weight.grad) >= 0.999
    assert _cos(lin.bias.grad, lin_ref.bias.grad) >= 0.999
    assert _cos(fb.grad, fr.grad) >= 0.999

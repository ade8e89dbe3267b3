"""Device-timed training-step rate of every BASELINE.json configuration on ONE GPU (eager launches,
CUDA events, 3 warm-up + N timed steps, batch 256): the headline config is bench.py's job; this
script gives the other configs a measured line each.

    python scripts/gpu_configs_bench.py [steps] > gpurun_out/configs.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from sota_imagenet_b200 import losses, models, optimizers, runner

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = 256


def timed(name, net, crit, opt, size, mix=None, labels="index", graph=True):
    torch.manual_seed(0)
    # the loader's native layout: bf16 channels_last [N,4,H,W], 4th channel zero (GpuAugment)
    x = torch.zeros(B, size, size, 4, device="cuda", dtype=torch.bfloat16)
    x[..., :3] = torch.randn(B, size, size, 3, device="cuda")
    x = x.permute(0, 3, 1, 2)
    y = torch.randint(0, 1000, (B,), device="cuda")
    xb = x

    def step():
        xi, yi = (x, y)
        if mix is not None:
            mix.state.input = (xb, y)
            mix.on_batch_begin()
            xi, yi = mix.state.input
        opt.zero_grad()
        loss = crit(net(xi), yi)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()        # result dropped: a live loss pins the AccumulateGrad nodes of plain nn.Parameters
                      # (the sphere-linear head) to this stream and the capture below would fail
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rec = {"config": name, "image_size": size, "batch": B}
    if graph and mix is None:            # same order as bench.py: warm-up, side-stream step, capture
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                step()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(STEPS):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / STEPS
            rec["ms_per_step_graph"] = round(ms, 3)
            rec["images_per_s_graph"] = round(B / ms * 1e3, 1)
        except Exception as e:           # capture is an optimisation; the eager number stands
            import traceback
            traceback.print_exc()
            rec["graph_error"] = repr(e)[:200]
            torch.cuda.synchronize()
    for _ in range(2):
        loss = step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(STEPS):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    rec.update({"ms_per_step_eager": round(ms, 3), "images_per_s_eager": round(B / ms * 1e3, 1),
                "loss": round(float(loss.detach()), 4),
                "mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)})
    print(json.dumps(rec), flush=True)
    torch.cuda.reset_peak_memory_stats()


def sgd(params):
    return optimizers.SGD(params, lr=0.01, momentum=0.9, weight_decay=3e-5, nesterov=True)


def main():
    only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
    want = lambda k: only is None or k in only
    ce = losses.CrossEntropyLoss(smoothing=0.1)
    if want("r50"):
        # C2 / C3: ResNet-50 at the progressive-resize sizes (reference resnet50_better.yaml:12-14)
        net = models.resnet50().cuda().train()
        opt = sgd(net.parameters())
        for size in (224, 192, 128):
            timed("resnet50 + smooth CE + SGD-Nesterov", net, ce, opt, size)
        del net, opt
    if want("novograd"):
        net = models.resnet50().cuda().train()
        timed("resnet50 + smooth CE + MyNovograd (reference optimizers.py:35-161)", net, ce,
              optimizers.MyNovograd(net.parameters(), lr=1e-3, weight_decay=1e-3), 224)
        del net
    if want("mix"):
        net = models.resnet50().cuda().train()

        class S:
            is_train, input = True, None
        np.random.seed(0)
        mix = runner.CutmixMixup(1.0, 0.2, prob=1.0)
        mix.set_state(S())
        timed("resnet50 + CutmixMixup(1.0, 0.2, prob 1) + smooth CE on soft targets", net, ce,
              sgd(net.parameters()), 224, mix=mix)
        del net
    if want("heads"):
        # C5: 512-d embedding -> SphereLinearLayer -> ArcFace / CosFace + smoothing 0.1
        for name, crit in (
            ("resnet50_embedding(512) + ArcFace(s=10,m=0.2)",
             losses.AdditiveAngularMarginLoss(final_criterion=ce, s=10.0, m=0.2)),
            ("resnet50_embedding(512) + CosFace = AdaCos(fixed_s=10, margin=0.2)",
             losses.AdaCos(final_criterion=ce, margin=0.2, fixed_s=10)),
        ):
            net = models.resnet50_embedding(512, 1000).cuda().train()
            timed(name, net, crit, sgd(net.parameters()), 224)
            del net
    if want("bresnet"):
        # C4: BResNet-50 encoder (reference configs/_old_configs/_first_attempts/BResNet50_encoder.yaml:44-60)
        net = models.resnet50(stem_type="deep", antialias=True, attn_type="eca", norm_layer="inplaceabn",
                              norm_act="leaky_relu", drop_rate=0.2, drop_connect_rate=0.2,
                              weight_standardization=True).cuda().train()
        timed("BResNet-50 encoder (deep stem, blur-pool, ECA, leaky ABN, WS, drop-connect .2)", net, ce,
              sgd(net.parameters()), 224, graph=False)


if __name__ == "__main__":
    main()

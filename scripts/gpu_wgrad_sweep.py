import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from sota_imagenet_b200 import ops
B = 256
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
shapes = [(64,64,56,1,1,0,1),(64,64,56,3,1,1,3),(64,256,56,1,1,0,4),(256,64,56,1,1,0,2),(128,128,28,3,1,1,3),(128,512,28,1,1,0,4),(512,128,28,1,1,0,3),
          (256,256,14,3,1,1,5),(256,1024,14,1,1,0,6),(1024,256,14,1,1,0,5),(512,512,7,3,1,1,2),(512,2048,7,1,1,0,3),(2048,512,7,1,1,0,2),
          (128,128,56,3,2,1,1),(256,512,56,1,2,0,1),(256,256,28,3,2,1,1),(512,1024,28,1,2,0,1),(512,512,14,3,2,1,1),(1024,2048,14,1,2,0,1),
          (256,128,56,1,1,0,1),(512,256,28,1,1,0,1),(1024,512,14,1,1,0,1)]
tot = 0
for (c,k,h,r,stride,pad,cnt) in shapes:
    x = ops.to_nhwc_bf16(torch.randn(B, c, h, h, device="cuda"))
    oh = (h + 2*pad - r)//stride + 1
    dy = ops.to_nhwc_bf16(torch.randn(B, k, oh, oh, device="cuda"))
    dw = torch.zeros(k, r, r, c, device="cuda").permute(0, 3, 1, 2)
    t = timeit(lambda: ops.conv2d_wgrad(x, dy, dw, stride=stride, pad=pad))
    tot += t * cnt
    print("%-26s %.3f ms x%d" % (str((c,k,h,r)), t, cnt))
print("weighted total %.3f ms (waves=%s)" % (tot, os.environ.get("SIB_WGRAD_WAVES", "4")))

import sys, torch
sys.path.insert(0, '.')
from videonavqa_b200 import ops
from videonavqa_b200.engine import Plan
B, C, T = 64, 512, 35
frames = [torch.randn(B, C, 10, 13, device='cuda').clamp_min_(0) for _ in range(T)]
view = torch.stack(frames).permute(1, 2, 3, 4, 0)[torch.randperm(B, device='cuda')]
del frames
plan = Plan([35] * B, B, T, torch.device('cuda'))
cont = view.contiguous()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print('K0 on contiguous (T innermost)      %.3f ms' % t(lambda: ops.frames_to_nhwc(cont, plan.img_of, plan.n_img, 1)))
print('K0 on the drivers frame-blocked view %.3f ms' % t(lambda: ops.frames_to_nhwc(view, plan.img_of, plan.n_img, 1)))
print('.contiguous() of that view (torch)   %.3f ms' % t(lambda: view.contiguous()))

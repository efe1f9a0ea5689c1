"""cuDNN conv3d fprop / dgrad / wgrad timings for the encoder's two big layers (64^3 x batch 4), fp32(TF32) vs bf16."""
import torch
torch.backends.cudnn.benchmark = True
aten = torch.ops.aten

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for name, (ci, co, d) in {"conv_0": (16, 32, 64), "conv_0_1": (32, 32, 64), "conv_1": (32, 64, 32), "conv_1_1": (64, 64, 32)}.items():
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        x = torch.randn(4, ci, d, d, d, device="cuda").to(dt).contiguous(memory_format=torch.channels_last_3d)
        w = (torch.randn(co, ci, 3, 3, 3, device="cuda") * 0.05).to(dt).contiguous(memory_format=torch.channels_last_3d)
        gy = torch.randn(4, co, d, d, d, device="cuda").to(dt).contiguous(memory_format=torch.channels_last_3d)
        args = (gy, x, w, [co], [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1)
        t_f = timeit(lambda: torch.nn.functional.conv3d(x, w, None, padding=1))
        t_d = timeit(lambda: aten.convolution_backward(*args, [True, False, False]))
        t_w = timeit(lambda: aten.convolution_backward(*args, [False, True, False]))
        print(f"{name:9s} {str(dt):15s} fprop {t_f:.3f} ms  dgrad {t_d:.3f} ms  wgrad {t_w:.3f} ms")
    # wgrad precision: bf16 vs fp32 (TF32 off) reference
    torch.backends.cudnn.allow_tf32 = False
    x = torch.randn(4, ci, d, d, d, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    gy = torch.randn(4, co, d, d, d, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    w = torch.randn(co, ci, 3, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    ref = aten.convolution_backward(gy, x, w, [co], [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1, [False, True, False])[1]
    torch.backends.cudnn.allow_tf32 = True
    tf = aten.convolution_backward(gy, x, w, [co], [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1, [False, True, False])[1]
    bf = aten.convolution_backward(gy.bfloat16(), x.bfloat16(), w.bfloat16(), [co], [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1,
                                   [False, True, False])[1].float()
    rel = lambda a: float((a - ref).norm() / ref.norm())
    print(f"{name:9s} wgrad rel-L2 error vs fp32: tf32 {rel(tf):.2e}  bf16 {rel(bf):.2e}")

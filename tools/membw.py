import torch, time
def t(f, n=20):
    f(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
x=torch.randn(544*960*192, device="cuda").to(torch.bfloat16); y=torch.empty_like(x)
ms=t(lambda: y.copy_(x)); print(f"copy 200MB: {ms*1e3:.1f} us {2*x.numel()*2/ms/1e6:.0f} GB/s")
ms=t(lambda: y.fill_(1.0)); print(f"fill 200MB: {ms*1e3:.1f} us {x.numel()*2/ms/1e6:.0f} GB/s")
ms=t(lambda: torch.nn.functional.gelu(x, ) ); print(f"gelu 200MB: {ms*1e3:.1f} us {2*x.numel()*2/ms/1e6:.0f} GB/s")
x4=x.view(1,544,960,192).permute(0,3,1,2)  # NCHW view of NHWC memory (channels_last)
w=torch.randn(192,1,3,3,device="cuda").to(torch.bfloat16)
ms=t(lambda: torch.nn.functional.conv2d(x4,w,padding=1,groups=192)); print(f"cudnn dw3x3 channels_last: {ms*1e3:.1f} us")
w1=torch.randn(192,192,1,1,device="cuda").to(torch.bfloat16)
ms=t(lambda: torch.nn.functional.conv2d(x4,w1)); print(f"cudnn pw 192->192 channels_last: {ms*1e3:.1f} us")
xm=x.view(-1,192)
ms=t(lambda: xm@w1.view(192,192).t()); print(f"cublas [522240,192]x[192,192]: {ms*1e3:.1f} us")

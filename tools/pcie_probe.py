import torch, time
a=torch.empty((64,1080,1920),dtype=torch.float32).pin_memory()
d=torch.empty_like(a,device='cuda')
for _ in range(3): d.copy_(a,non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): d.copy_(a,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('H2D GB/s', 10*a.numel()*4/dt/1e9)
h=torch.empty((64,1080,1920),dtype=torch.float32).pin_memory()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): h.copy_(d,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('D2H GB/s', 10*a.numel()*4/dt/1e9)

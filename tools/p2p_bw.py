import torch, time
a = torch.empty(256*1024*1024//8, dtype=torch.int64, device="cuda:0")
b = torch.empty_like(a, device="cuda:1")
c = torch.empty_like(a, device="cuda:1"); d = torch.empty_like(a, device="cuda:0")
for _ in range(3): b.copy_(a); d.copy_(c)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.device(0):
    e0.record()
    for _ in range(10): b.copy_(a)
    e1.record(); torch.cuda.synchronize(0)
print("one direction 256MB:", 10*a.numel()*8/ (e0.elapsed_time(e1)*1e-3)/1e9, "GB/s", e0.elapsed_time(e1)/10, "ms")
# both directions at once
s0, s1 = torch.cuda.Stream(0), torch.cuda.Stream(1)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s0): b.copy_(a, non_blocking=True)
    with torch.cuda.stream(s1): d.copy_(c, non_blocking=True)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
dt = time.perf_counter()-t0
print("both directions: per direction", 10*a.numel()*8/dt/1e9, "GB/s", dt/10*1e3, "ms per pair")

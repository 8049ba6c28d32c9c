import os, time, torch
print("ALLOC_CONF", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), os.environ.get("PYTORCH_ALLOC_CONF"))
dev = torch.device("cuda")
torch.cuda.init(); torch.zeros(1, device=dev); torch.cuda.synchronize()
def t_alloc(fn, n=8):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); x = fn(); out.append(1e6 * (time.perf_counter() - t0))
    return [round(v) for v in out]
N = 135_000_000
print("empty, drop prev        ", t_alloc(lambda: torch.empty(N, dtype=torch.uint8, device=dev)))
keep = []
def two_gen():
    keep.append(torch.empty(N, dtype=torch.uint8, device=dev))
    if len(keep) > 2: keep.pop(0)
print("empty, 2 generations    ", t_alloc(two_gen, 10))
def resize():
    t = torch.empty(0, dtype=torch.uint8, device=dev); t.resize_(N); return t
print("empty(0)+resize_        ", t_alloc(resize))
print("zeros 74MB              ", t_alloc(lambda: torch.zeros((9,1080,1920), device=dev)))
print(torch.cuda.memory_summary(abbreviated=True)[:1500])

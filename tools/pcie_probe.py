import time, torch
dev = torch.device("cuda", 0)
N = 10 * 1024**3
h = torch.empty(N, dtype=torch.uint8, pin_memory=True)
d = torch.empty(N, dtype=torch.uint8, device=dev)
for ns in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step = N // ns
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i*step:(i+1)*step].copy_(h[i*step:(i+1)*step], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("H2D streams=%d: %.1f GB/s" % (ns, N / dt / 1e9), flush=True)
# D2H concurrently with H2D
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(N // 4, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1):
    d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2):
    h2.copy_(d[:N // 4], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D 10GB + concurrent D2H 2.5GB: %.3f s (H2D alone would be %.3f)" % (dt, N / 50e9))
import subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])

#!/usr/bin/env python3
"""A/B the host pipeline chunk size of coast_run_host on ONE box (PCIe rates differ between boxes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coast_b200 as cb
rt = cb.Runtime(0)
n = 1 << 20
h_in = torch.randint(0, 255, (n * 64,), dtype=torch.uint8).pin_memory()
h_out = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
for mib in (4, 8, 16):
    os.environ["COAST_HOST_CHUNK_BYTES"] = str(mib << 20)
    for _ in range(3):
        rt.run_host(cb.K_SHA256, 3, h_in, h_out, n, unit_bytes=64, flags=3)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20):
        rt.run_host(cb.K_SHA256, 3, h_in, h_out, n, unit_bytes=64, flags=3)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
    print(f"chunk {mib:3d} MiB: {dt * 1e3:.3f} ms/step  {n * 32 / dt / 1e9:.2f} GB/s voted output")
d = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); o = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): d.copy_(h_in, non_blocking=True)
torch.cuda.synchronize(); print("bare H2D", n * 64 * 10 / (time.perf_counter() - t) / 1e9, "GB/s")
t = time.perf_counter()
for _ in range(10): h_out.copy_(o, non_blocking=True)
torch.cuda.synchronize(); print("bare D2H", n * 32 * 10 / (time.perf_counter() - t) / 1e9, "GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(o, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print(f"concurrent H2D(64MiB)+D2H(32MiB): {dt*1e3:.3f} ms per pair")

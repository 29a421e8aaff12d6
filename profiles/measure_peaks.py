"""Measured cuBLAS peaks the driver's MEASURED_PEAKS.json does not hold (BASELINE.md section 2 / SURVEY.md section 6: "the
builder must measure"): dense TF32 and FP64 GEMM throughput, plus bf16 / fp16 for reference, same method as the driver
(torch.matmul N^3, best of 10 = burst; back to back for 3 s = sustained; CUDA events).
usage: python profiles/measure_peaks.py > profiles/r02_measured_peaks_tf32_fp64.json"""
import json
import time

import torch


def gemm_tflops(dtype, n, tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = float("inf")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 3.0:
        for _ in range(10):
            a @ b
        reps += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    flop = 2.0 * n ** 3
    return {"burst_tflops": flop / (best * 1e-3) / 1e12, "sustained_tflops": flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
            "n": n}


out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
       "how": "torch.matmul n^3 (cuBLAS), best of 10 (burst) / back to back for 3 s (sustained), CUDA events",
       "bf16": gemm_tflops(torch.bfloat16, 8192), "fp16": gemm_tflops(torch.float16, 8192),
       "tf32": gemm_tflops(torch.float32, 8192, tf32=True), "fp32_simt": gemm_tflops(torch.float32, 8192, tf32=False),
       "fp64": gemm_tflops(torch.float64, 4096)}
print(json.dumps(out, indent=1))

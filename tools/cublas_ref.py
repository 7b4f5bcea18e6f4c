"""Context only: cuBLAS (torch.matmul bf16) on the ViT-B/16 GEMM shapes of a 256-image batch, to know the ceiling
this machine reaches on the same shapes (not used by the product)."""
import torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
M = 50432
for name, N, K in [("qkv", 2304, 768), ("out", 768, 768), ("fc", 3072, 768), ("proj", 768, 3072), ("square8k", 8192, 8192)]:
    m = 8192 if name == "square8k" else M
    a = torch.randn(m, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    for _ in range(3): c = a @ w.t()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c = a @ w.t()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cublas {name:8s} M={m} N={N} K={K}: {ms:.3f} ms -> {2*m*N*K/ms/1e9:.1f} TFLOP/s (plain GEMM, no bias/epilogue)")

"""Dev: attention forward throughput (algorithmic 4*Tq*Tk*d FLOP per head, halved when causal). MPL_ATTN_TC=0 selects
the mma.sync kernels.  python tools/attn_bench.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medplib_b200 import ops
dev = torch.device("cuda:0")
for (B, H, T, d, causal) in ((8, 32, 639, 128, True), (1, 32, 615, 128, True), (1, 32, 1299, 128, True), (8, 32, 2048, 128, True),
                             (4, 16, 577, 64, False), (1, 16, 577, 64, False)):
    q, k, v = (torch.randn(B, T, H, d, device=dev).to(torch.bfloat16) for _ in range(3))
    for _ in range(3):
        ops.attention(q, k, v, 1 / math.sqrt(d), causal=causal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.attention(q, k, v, 1 / math.sqrt(d), causal=causal)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * B * H * T * T * d * (0.5 if causal else 1.0)
    print(f"B={B} H={H} T={T} d={d} causal={causal}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  (MPL_ATTN_TC={os.environ.get('MPL_ATTN_TC')})")

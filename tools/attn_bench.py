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

# backward (d = 128, causal self-attention as in the train step): algorithmic 10*T*T*d FLOP per head, halved when causal
# (5 GEMMs: S, dP, dV, dK, dQ). MPL_ATTN_BWD_TC=0 selects the mma.sync FA2-style kernel.
from medplib_b200 import train_ops as Tr
for (B, H, T, d) in ((8, 32, 639, 128), (1, 32, 615, 128), (8, 32, 2048, 128)):
    qkv = torch.randn(B, T, 3, H, d, device=dev).to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    d_o = torch.randn(B, T, H, d, device=dev).to(torch.bfloat16)
    o, lse = Tr.attention_fwd_lse(q, k, v, 1 / math.sqrt(d), True, None)
    dqkv = torch.zeros_like(qkv)
    for _ in range(3):
        Tr.attention_bwd(q, k, v, o, d_o, lse, 1 / math.sqrt(d), dqkv[:, :, 1], dqkv[:, :, 2], True, None)
    torch.cuda.synchronize()
    n = 10
    e0.record()
    for _ in range(n):
        Tr.attention_bwd(q, k, v, o, d_o, lse, 1 / math.sqrt(d), dqkv[:, :, 1], dqkv[:, :, 2], True, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 10.0 * B * H * T * T * d * 0.5
    print(f"bwd B={B} H={H} T={T} d={d}: {ms * 1e3:8.1f} us (incl. dq zero-fill + delta)  {fl / ms / 1e9:7.1f} TFLOP/s  "
          f"(MPL_ATTN_BWD_TC={os.environ.get('MPL_ATTN_BWD_TC')})")

"""CPU oracle for the MedPLIB multimodal hot path — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU) restatement of the reference's algorithm for the path named in BASELINE.json:north_star.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product (medplib_b200/) never does and has no CPU fallback.

Why a restatement: the reference's arithmetic lives partly in un-vendored third-party packages that are absent
from /root/reference and from this image:
  * transformers==4.31.0  (requirements.txt:137)  LLaMA + CLIP modules, greedy_search       -> oracle/llama.py, oracle/clip.py
  * deepspeed==0.13.1     (requirements.txt:22)   deepspeed.moe.{layer,sharded_moe,experts}  -> oracle/moe.py
  * peft==0.10.0          (requirements.txt:80)   LoRA Linear                                -> (training, later round)
and partly in-tree (SAM-Med2D, projector/compressor/mask-encoder/region sampler, heads, losses), restated in
oracle/sam.py, oracle/arch.py, oracle/heads.py so the checker can travel to the GPU box, where /root/reference does
not exist.

Parity pinning: the reference ships NO golden vectors, known-answer tests or fixtures for this path,
so the oracle is pinned against outputs of the reference itself, generated in the authoring container by importing
/root/reference (tests/golden/make_golden.py, vectors committed under tests/golden/) for every in-tree component, and
against the installed transformers-5.5 eager LLaMA / CLIP modules in fp32 for the HF pieces. The DeepSpeed MoE layer
cannot be run anywhere in this environment: oracle/moe.py is "parity unpinned" against DeepSpeed itself and is
checked only through identities (eval path == gate-prob * FFN of the argmax expert when nothing is dropped).

All functions are written functionally over a state dict that uses the reference's parameter names, compute in the
dtype of the tensors they are given (bf16 reproduces the reference's eager bf16 cast points; fp32 is exact math), and
cite the reference file:line they follow.
"""

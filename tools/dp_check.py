"""Data-parallel check of the train step on N GPUs (run under torchrun, NCCL): every rank trains on its own batch;
the bucketed all-reduce must leave mean(grad_rank) in every rank's arena, and after Trainer.step() all ranks must hold
identical parameters. Each rank recomputes the other ranks' gradients locally (reducer off) as the expectation."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_train_gpu as tt  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
m, sd, ocfg = tt.build(dev, cf=1.5, aux=0.0)
tr = m.trainer(lr=1e-2, bucket_elems=1 << 16)
assert tr.reducer.on and tr.reducer.world == world and len(tr.reducer.bounds) > 4


def run(seed, reduce):
    ids, labels, am, clip_img, sam_img, gts = tt.batch(seg=True, seed=seed)
    S = ids.shape[0] * (ids.shape[1] - 1 + 16)
    g = torch.Generator().manual_seed(11)
    noise = [torch.rand(S, 2, generator=g).to(dev) for _ in range(2)]
    tr.zero_grad()
    tr.reducer.on = reduce
    out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
            labels=labels.to(dev), attention_mask=am.to(dev), offset=None, masks_list=[x.to(dev) for x in gts],
            label_list=[x.to(dev) for x in gts], resize_list=[(256, 256)] * len(gts), inference=False, seg_flag=True,
            moe_noise=noise)
    out["loss"].backward()
    scale = tr.reducer.finish() if reduce else 1.0
    torch.cuda.synchronize()
    return tr.arena.flat.clone() * scale


expect = sum(run(5 + r, False) for r in range(world)) / world
launched_early = []
orig_ready = tr.reducer.ready


def spy(upto):
    orig_ready(upto)
    launched_early.append(tr.reducer.next)


tr.reducer.ready = spy
got = run(5 + rank, True)
err = (got - expect).abs().max().item() / expect.abs().max().item()
# fp32 atomics make the per-rank gradients run-to-run non-bit-identical; the reduced mean must match to the noise of the
# wire format (bf16 by default, like the reference's bf16 engine: one rounding of every addend, 2^-9 relative)
tol = 2e-3 if tr.reducer.wire == torch.float32 else 6e-3
assert err < tol, f"rank {rank}: reduced gradient differs from the mean of per-rank gradients: {err}"
assert max(launched_early[:-1] or [0]) > 0, "no bucket was launched before the backward finished"
tr.reducer.ready = orig_ready
got0 = got.clone()
dist.broadcast(got0, 0)
assert torch.equal(got, got0), (f"rank {rank}: the all-reduced arena differs from rank 0's: max diff "
                                f"{(got - got0).abs().max().item():.3e} at {int((got - got0).abs().argmax())} of {got.numel()}")
# one optimizer step from the reduced arena: parameters must stay identical across ranks
tr.arena.flat.copy_(got / (1.0 / world))
tr.reducer.reset()
tr.reducer.on = False
tr.opt.step(grad_scale=1.0 / world)
flat = torch.cat([p.detach().float().reshape(-1) for p in tr.arena.params])
ref = flat.clone()
dist.broadcast(ref, 0)
assert torch.equal(flat, ref), f"rank {rank}: parameters diverged after the step"
if rank == 0:
    print(f"dp_check ok: world={world}, reduced-gradient rel err {err:.2e}, buckets {len(tr.reducer.bounds)}, "
          f"buckets in flight before the backward ended: {max(launched_early[:-1] or [0])}")
dist.destroy_process_group()

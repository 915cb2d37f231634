"""The part of ``deepspeed`` 0.13.1 the reference's drivers touch, routed to medplib_b200's own train step:

  deepspeed.init_distributed(dist_backend=...)                      model/eval/vqa_infer.py:185, model/serve/model_worker.py:40
  deepspeed.moe.layer.MoE                                            model/MedPLIB.py:21,253-263, medplib_moe_llama.py:604-614
  deepspeed.moe.utils.split_params_into_different_moe_groups_for_optimizer     train_ds_medplib.py:422-431
  deepspeed.initialize(model=, model_parameters=, training_data=, collate_fn=, config=)   train_ds_medplib.py:439-448
      -> engine(**batch), engine.backward(loss), engine.step(), engine.global_steps, engine.train()/eval(),
         engine.save_checkpoint(dir), engine.load_checkpoint(dir)    train_ds_medplib.py:452-470,517-521,599,624-625,693-698

What it maps to: ``engine.backward(loss)`` = ``loss.backward()`` through medplib_b200's three tape nodes (gradients land
in the fp32 arena; during the last micro-step of an accumulation window the arena is all-reduced bucket by bucket over
NCCL while the backward is still running), ``engine.step()`` = fused clip + AdamW on fp32 masters with the config's
WarmupDecayLR schedule. ZeRO-2's optimizer-state sharding is NOT reproduced (every rank keeps the 318 M-parameter
masters and moments: 3.8 GB of 180) — the exchange step of the path is the gradient all-reduce.
"""
import math
import os
import sys
import types

import torch
import torch.nn as nn

from .. import _lib
from ..model import modules as _modules

__version__ = "0.13.1+medplib_b200.compat"


# ------------------------------------------------------------------------------------------------ deepspeed.moe.*
def split_params_into_different_moe_groups_for_optimizer(param_groups, max_group_size=None):
    """deepspeed/moe/utils.py: expert parameters (``allreduce == False``) leave the dense groups and get groups of their
    own named after their expert-parallel group. With ep_size = 1 (the only value the reference uses) the expert data
    parallel group is every rank, so the split only changes bookkeeping — the groups are returned for parity of shape."""
    if isinstance(param_groups, dict):
        param_groups = [param_groups]
    out = []
    for g in param_groups:
        dense = dict(g)
        dense["params"] = [p for p in g["params"] if getattr(p, "allreduce", True)]
        out.append(dense)
        by_group = {}
        for p in g["params"]:
            if not getattr(p, "allreduce", True):
                by_group.setdefault(getattr(p, "group_name", "ep_size_1"), []).append(p)
        for name, ps in by_group.items():
            moe_g = {k: v for k, v in g.items() if k != "params"}
            moe_g.update(name=name, moe=True, params=ps)
            out.append(moe_g)
    return out


moe = types.ModuleType("deepspeed.moe")
moe.layer = types.ModuleType("deepspeed.moe.layer")
moe.layer.MoE = _modules.MoE
moe.utils = types.ModuleType("deepspeed.moe.utils")
moe.utils.split_params_into_different_moe_groups_for_optimizer = split_params_into_different_moe_groups_for_optimizer


# ------------------------------------------------------------------------------------------------ process group
def init_distributed(dist_backend="nccl", auto_mpi_discovery=False, timeout=None, init_method=None, **unused):
    """One process per GPU under torchrun / the deepspeed launcher (RANK, WORLD_SIZE, MASTER_* in the environment).
    A single process needs no group: every collective of the path is skipped when none is initialised."""
    if torch.distributed.is_available() and not torch.distributed.is_initialized() \
            and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        backend = dist_backend if (dist_backend != "nccl" or torch.cuda.is_available()) else "gloo"
        torch.distributed.init_process_group(backend, init_method=init_method)


# ------------------------------------------------------------------------------------------------ lr schedule
class WarmupDecayLR:
    """deepspeed/runtime/lr_schedules.py::WarmupDecayLR: linear (or log) warm-up from warmup_min_lr to warmup_max_lr
    over warmup_num_steps, then linear decay to 0 at total_num_steps."""

    def __init__(self, total_num_steps, warmup_min_lr=0.0, warmup_max_lr=1e-3, warmup_num_steps=1000,
                 warmup_type="log", last_batch_iteration=-1):
        self.total, self.lo, self.hi = int(total_num_steps), float(warmup_min_lr), float(warmup_max_lr)
        self.warm = max(2, int(warmup_num_steps))
        self.kind = warmup_type
        self.last_batch_iteration = last_batch_iteration
        self.inv_log = 1.0 / math.log(self.warm)

    def _gamma(self):
        it = self.last_batch_iteration
        if it < self.warm:
            return self.inv_log * math.log(it + 1) if self.kind == "log" else min(1.0, it / self.warm)
        return max(0.0, (self.total - it) / max(1.0, self.total - self.warm))

    def get_lr(self):
        if self.last_batch_iteration < 0:
            return [0.0]
        return [self.lo + (self.hi - self.lo) * self._gamma()]

    def get_last_lr(self):
        return self.get_lr()

    def step(self, last_batch_iteration=None):
        self.last_batch_iteration = self.last_batch_iteration + 1 if last_batch_iteration is None \
            else last_batch_iteration

    def state_dict(self):
        return {"last_batch_iteration": self.last_batch_iteration}

    def load_state_dict(self, sd):
        self.last_batch_iteration = sd["last_batch_iteration"]


class _ConstantLR(WarmupDecayLR):
    def __init__(self, lr):
        super().__init__(1, 0.0, lr, 2, "linear", 0)

    def get_lr(self):
        return [self.hi]


# ------------------------------------------------------------------------------------------------ engine
def _base_model(model):
    """The MedPLIBForCausalLM under a (peft / peft-stand-in) wrapper."""
    m = model
    for _ in range(4):
        if hasattr(type(m), "model_forward") and hasattr(type(m), "trainer"):  # (class attributes: wrappers forward
            return m                                                           #  instance lookups to what they wrap)
        if hasattr(type(m), "get_base_model"):
            m = m.get_base_model()
        elif isinstance(getattr(m, "base_model", None), nn.Module) and hasattr(m.base_model, "model"):
            m = m.base_model.model
        elif isinstance(getattr(m, "module", None), nn.Module):
            m = m.module
        else:
            break
    if not hasattr(type(m), "trainer"):
        raise _lib.MplError("deepspeed.initialize (medplib_b200 stand-in) needs a medplib_b200 model")
    return m


class DeepSpeedEngine(nn.Module):
    def __init__(self, model, config, training_data=None, collate_fn=None):
        super().__init__()
        self.module = model
        self._config = config or {}
        self.global_steps = 0
        self.micro_steps = 0
        self._gas = int(self._config.get("gradient_accumulation_steps", 1) or 1)
        self._micro_bs = int(self._config.get("train_micro_batch_size_per_gpu", 1) or 1)
        opt = (self._config.get("optimizer") or {})
        if opt.get("type", "AdamW").lower() not in ("adamw", "adam"):
            raise _lib.MplError(f"optimizer {opt.get('type')!r} is not built (the reference trains with AdamW)")
        op = opt.get("params") or {}
        self._opt_kw = dict(lr=float(op.get("lr", 3e-4)), betas=tuple(op.get("betas", (0.9, 0.999))),
                            eps=float(op.get("eps", 1e-8)), weight_decay=float(op.get("weight_decay", 0.0)),
                            max_grad_norm=float(self._config.get("gradient_clipping", 0.0) or 0.0))
        sch = self._config.get("scheduler") or {}
        if sch.get("type") == "WarmupDecayLR":
            self.lr_scheduler = WarmupDecayLR(**sch.get("params", {}))
            self.lr_scheduler.step()  # DeepSpeed steps the schedule once at construction (iteration 0)
        elif not sch:
            self.lr_scheduler = _ConstantLR(self._opt_kw["lr"])
        else:
            raise _lib.MplError(f"lr scheduler {sch.get('type')!r} is not built (the reference uses WarmupDecayLR)")
        self._trainer = None
        self.training_dataloader = self._loader(training_data, collate_fn) if training_data is not None else None

    # -- data
    def _loader(self, dataset, collate_fn):
        sampler = None
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            sampler = torch.utils.data.distributed.DistributedSampler(dataset, shuffle=True, drop_last=False)
        return torch.utils.data.DataLoader(dataset, batch_size=self._micro_bs, shuffle=sampler is None, sampler=sampler,
                                           collate_fn=collate_fn, num_workers=0, drop_last=False)

    # -- nn.Module surface
    def forward(self, *a, **k):
        return self.module(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)

    def train(self, mode=True):
        self.module.train(mode)
        return self

    def eval(self):
        return self.train(False)

    # -- optimisation
    @property
    def optimizer(self):
        return self.trainer().opt

    def trainer(self):
        if self._trainer is None:
            self._trainer = _base_model(self.module).trainer(**self._opt_kw)
        return self._trainer

    def gradient_accumulation_steps(self):
        return self._gas

    def is_gradient_accumulation_boundary(self):
        return (self.micro_steps + 1) % self._gas == 0

    def backward(self, loss, **unused):
        tr = self.trainer()
        if self._gas > 1:
            loss = loss / self._gas
        if self.is_gradient_accumulation_boundary():
            loss.backward()
        else:
            with tr.no_sync():
                loss.backward()
        return loss

    def step(self):
        boundary = self.is_gradient_accumulation_boundary()
        self.micro_steps += 1
        if not boundary:
            return
        self.trainer().step(lr=self.lr_scheduler.get_lr()[0])
        self.lr_scheduler.step()
        self.global_steps += 1

    def get_lr(self):
        return self.lr_scheduler.get_lr()

    # -- checkpoints (the directory layout params_bf16_to_f32.py and the merge scripts read)
    def save_checkpoint(self, save_dir, tag=None, client_state=None, save_latest=True):
        from .. import checkpoint as ck
        tag = tag or f"global_step{self.global_steps}"
        path = os.path.join(save_dir, str(tag))
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        if rank == 0:
            ck.save_deepspeed_layout(self.module, path)
            tr = self._trainer
            extra = {"global_steps": self.global_steps, "micro_steps": self.micro_steps,
                     "lr_scheduler": self.lr_scheduler.state_dict(), "client_state": client_state or {}}
            if tr is not None:
                extra["optimizer"] = {"names": list(tr.arena.names), "master": tr.opt.master.cpu(), "m": tr.opt.m.cpu(),
                                      "v": tr.opt.v.cpu(), "t": tr.opt.t}
            torch.save(extra, os.path.join(path, "medplib_b200_optim_states.pt"))
            if save_latest:
                with open(os.path.join(save_dir, "latest"), "w") as f:
                    f.write(str(tag))
        if torch.distributed.is_initialized():
            torch.distributed.barrier()
        return True

    def load_checkpoint(self, load_dir, tag=None, load_optimizer_states=True, **unused):
        from .. import checkpoint as ck
        if tag is None:
            latest = os.path.join(load_dir, "latest")
            if not os.path.exists(latest):
                return None, None
            tag = open(latest).read().strip()
        path = os.path.join(load_dir, str(tag))
        sd = ck.merge_deepspeed_states(path, dtype=None)
        ck.load_into(_base_model(self.module), sd, lora="keep" if ck.lora_keys(ck.normalize_keys(sd)) else "auto")
        extra_p = os.path.join(path, "medplib_b200_optim_states.pt")
        client = {}
        if os.path.exists(extra_p):
            extra = torch.load(extra_p, map_location="cpu", weights_only=False)
            self.global_steps, self.micro_steps = extra["global_steps"], extra["micro_steps"]
            self.lr_scheduler.load_state_dict(extra["lr_scheduler"])
            client = extra.get("client_state", {})
            if load_optimizer_states and "optimizer" in extra:
                tr = self.trainer()
                o = extra["optimizer"]
                if o["names"] == list(tr.arena.names):
                    dev = tr.opt.master.device
                    tr.opt.master.copy_(o["master"].to(dev))
                    tr.opt.m.copy_(o["m"].to(dev))
                    tr.opt.v.copy_(o["v"].to(dev))
                    tr.opt.t = o["t"]
        return path, client


def initialize(args=None, model=None, optimizer=None, model_parameters=None, training_data=None, lr_scheduler=None,
               mpu=None, dist_init_required=None, collate_fn=None, config=None, config_params=None, **unused):
    """deepspeed.initialize -> (engine, optimizer, training_dataloader, lr_scheduler)."""
    if optimizer is not None or lr_scheduler is not None:
        raise _lib.MplError("pass the optimizer / scheduler through the DeepSpeed config (as train_ds_medplib.py does)")
    init_distributed()
    engine = DeepSpeedEngine(model, config if config is not None else config_params, training_data, collate_fn)
    # the optimizer is built lazily (first backward): it needs the model in bf16 on its GPU
    return engine, None, engine.training_dataloader, engine.lr_scheduler


sys.modules.setdefault(__name__ + ".moe", moe)

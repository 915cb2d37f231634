"""Stand-ins for the two un-vendored packages the reference's drivers import next to the model classes — ``peft``
(``train_ds_medplib.py:15,294-302``) and ``deepspeed`` (``train_ds_medplib.py:9,422-448,624-625``,
``model/eval/vqa_infer.py:185``, ``model/MedPLIB.py:21``) — so that those drivers run UNCHANGED against
``medplib_b200.model`` on a box that has neither (this image: both absent).

They are not re-implementations of peft / DeepSpeed: they expose exactly the surface the reference touches and route it
to medplib_b200's own train step (``medplib_b200/train.py``: LoRA adapters, bucketed NCCL all-reduce, fused AdamW).
With the real packages installed nothing here is used — see INTEGRATION.md for what then differs (ZeRO-2 sharding is
DeepSpeed's; the model hands it ordinary ``.grad`` tensors through ``Trainer(foreign_grads=True)``).

    import medplib_b200.compat as compat
    compat.install()            # registers `peft`, `deepspeed`, `deepspeed.moe.layer`, `deepspeed.moe.utils` if missing
"""
import importlib
import importlib.util
import sys


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install(force=False):
    """Register the stand-ins under the names the reference imports. Returns the list of names installed."""
    done = []
    if force or _missing("peft"):
        from . import peft_shim
        sys.modules["peft"] = peft_shim
        done.append("peft")
    if force or _missing("deepspeed"):
        from . import deepspeed_shim as ds
        sys.modules["deepspeed"] = ds
        sys.modules["deepspeed.moe"] = ds.moe
        sys.modules["deepspeed.moe.layer"] = ds.moe.layer
        sys.modules["deepspeed.moe.utils"] = ds.moe.utils
        done.append("deepspeed")
    return done

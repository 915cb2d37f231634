"""Config classes with the reference's model_type strings and fields
(model/medplib/model/language_model/medplib_moe_llama.py:48-80, medplib_llama.py:28-30)."""
from transformers import LlamaConfig


class MedPLIBMoELlamaConfig(LlamaConfig):
    model_type = "medplib_moe_llama"

    def __init__(self, moe_enable=True, moe_mode="sparse", moe_layers_idx=None, ep_size=1, top_k_experts=2,
                 capacity_factor=1.0, eval_capacity_factor=1.0, min_capacity=4, use_residual=False,
                 router_aux_loss_coef=0.01, **kwargs):
        moe = kwargs.pop("moe", None)
        self.moe = moe if moe is not None else dict(
            moe_enable=moe_enable, moe_mode=moe_mode, moe_layers_idx=moe_layers_idx, ep_size=ep_size,
            top_k_experts=top_k_experts, capacity_factor=capacity_factor, eval_capacity_factor=eval_capacity_factor,
            min_capacity=min_capacity, use_residual=use_residual, router_aux_loss_coef=router_aux_loss_coef,
            train_modules=[])
        super().__init__(**kwargs)


class LlavaConfig(LlamaConfig):
    model_type = "medplib"


def llama_dims(config):
    """The plain-dict view of a config that medplib_b200.engine.LlamaEngine consumes."""
    return dict(hidden_size=config.hidden_size, intermediate_size=config.intermediate_size,
                num_layers=config.num_hidden_layers, num_heads=config.num_attention_heads,
                vocab_size=config.vocab_size, rms_norm_eps=config.rms_norm_eps,
                max_position_embeddings=getattr(config, "max_position_embeddings", 4096),
                rope_theta=_rope_theta(config), moe=getattr(config, "moe", None))


def _rope_theta(config):
    t = getattr(config, "rope_theta", None)
    if t is None:
        rp = getattr(config, "rope_parameters", None) or {}
        t = rp.get("rope_theta", 10000.0) if isinstance(rp, dict) else 10000.0
    return float(t)

"""LISAForCausalLM — the non-MoE twin of MedPLIBForCausalLM (reference: model/LISA.py:180-555).

Same grounding head, losses and methods; dense LlamaMLP in every layer. Differences the reference has and this class
keeps: the constructor also accepts ``vision_tower``, ``region_fea_adapter``, ``region_geo_sampler``,
``max_sample_point``, ``sampler_pooler_mode`` (LISA.py:189-208), and ``model_forward`` takes ``attention_masks``
(LISA.py:267) — both spellings are accepted because the collator emits ``attention_mask``.
"""
from .config import LlavaConfig
from .MedPLIB import MedPLIBForCausalLM


class LISAForCausalLM(MedPLIBForCausalLM):
    config_class = LlavaConfig

    def __init__(self, config, **kwargs):
        config.mm_vision_tower = kwargs.get("vision_tower", getattr(config, "mm_vision_tower", None))
        config.max_sample_point = kwargs.get("max_sample_point", getattr(config, "max_sample_point", 512))
        config.region_fea_adapter = kwargs.get("region_fea_adapter", getattr(config, "region_fea_adapter", True))
        config.region_geo_sampler = kwargs.get("region_geo_sampler", False)
        config.sampler_pooler_mode = kwargs.get("sampler_pooler_mode", getattr(config, "sampler_pooler_mode", "max"))
        kwargs = dict(kwargs)
        kwargs["test_only"] = False
        for k in ("num_experts", "top_k_experts", "capacity_factor", "use_residual", "router_aux_loss_coef",
                  "eval_capacity_factor", "moe_layers_idx", "min_capacity", "ep_size"):
            kwargs.pop(k, None)
        super().__init__(config, **kwargs)
        config.moe = None

    def model_forward(self, images, images_clip, input_ids, region_masks=None, labels=None, attention_masks=None,
                      offset=None, masks_list=None, label_list=None, resize_list=None, inference=False, **kwargs):
        am = attention_masks if attention_masks is not None else kwargs.pop("attention_mask", None)
        kwargs.pop("attention_mask", None)
        return super().model_forward(images, images_clip, input_ids, region_masks=region_masks, labels=labels,
                                     attention_mask=am, offset=offset, masks_list=masks_list, label_list=label_list,
                                     resize_list=resize_list, inference=inference, **kwargs)

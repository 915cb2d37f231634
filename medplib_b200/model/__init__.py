"""Drop-in model classes (same names as the reference's model/__init__.py:2-3)."""
from .MedPLIB import MedPLIBForCausalLM  # noqa: F401
from .LISA import LISAForCausalLM  # noqa: F401
from .config import MedPLIBMoELlamaConfig, LlavaConfig  # noqa: F401

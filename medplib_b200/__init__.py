"""medplib_b200 — B200-native (sm_100a) implementation of MedPLIB's multimodal forward hot path.

The arithmetic lives in ``libmedplib_b200.so`` (hand-written CUDA behind the C ABI in ``include/medplib_b200.h``);
this package is the Python host side that mirrors the reference's module API (``model/MedPLIB.py``,
``model/LISA.py``). There is no CPU fallback: importing :mod:`medplib_b200.ops` without the built library raises.
"""
__version__ = "0.1.0"

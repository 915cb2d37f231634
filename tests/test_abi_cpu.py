"""CPU: the C-ABI library loads and exports every symbol include/medplib_b200.h declares (no compute calls), the
ctypes struct layouts match the header, and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "medplib_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(mpl_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from medplib_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == syms, "medplib_b200/_lib.py:EXPORTS must list exactly the header's entry points"
    assert lib.mpl_version() == 1
    assert lib.mpl_launch_count() == 0


def test_struct_sizes_match_header():
    """sizeof() of every ctypes mirror equals the C struct's size as compiled by gcc from the header itself."""
    import subprocess
    import tempfile
    from medplib_b200 import _lib
    pairs = {"mpl_gemm_args": _lib.GemmArgs, "mpl_attn_args": _lib.AttnArgs, "mpl_attn_bwd_args": _lib.AttnBwdArgs, "mpl_moe_route_args": _lib.MoeRouteArgs,
             "mpl_llama_layer": _lib.LlamaLayer, "mpl_llama_model": _lib.LlamaModel, "mpl_llama_io": _lib.LlamaIO,
             "mpl_clip_layer": _lib.ClipLayer, "mpl_clip_model": _lib.ClipModel, "mpl_sam_block": _lib.SamBlock,
             "mpl_sam_encoder": _lib.SamEncoder, "mpl_sam_attn": _lib.SamAttn,
             "mpl_sam_twoway_layer": _lib.SamTwoWayLayer, "mpl_sam_mask_decoder": _lib.SamMaskDecoder,
             "mpl_preprocess_job": __import__("medplib_b200.preprocess", fromlist=["PreprocessJob"]).PreprocessJob}
    prog = '#include <stdio.h>\n#include "medplib_b200.h"\nint main(void){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in pairs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o",
                        os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert ctypes.sizeof(pairs[name]) == int(size), f"{name}: ctypes {ctypes.sizeof(pairs[name])} != C {size}"


def test_no_cpu_fallback():
    from medplib_b200 import _lib, ops
    with pytest.raises(_lib.MplError):
        ops.linear(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(_lib.MplError):
        ops.rmsnorm(torch.zeros(4, 8, dtype=torch.bfloat16), torch.ones(8, dtype=torch.bfloat16), 1e-5)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for top in ("medplib_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".cuh")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{top}/{f} imports the oracle"
    # bench.py: only inside the CPU legs (the reference arm's weights + one-image pass, the input pipeline's port), never
    # at module level
    bench = open(os.path.join(ROOT, "bench.py")).read()
    assert not re.search(r"^(from|import)\s+oracle\b", bench, re.M)
    for m in re.finditer(r"^\s+from oracle import", bench, re.M):
        fn = re.findall(r"^def (\w+)\(", bench[:m.start()], re.M)[-1]
        assert fn in ("_cpu_state", "cpu_image", "cpu_preprocess", "_pre_one"), f"bench.py:{fn} imports the oracle"

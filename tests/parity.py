"""One comparison helper for every GPU parity test: max |got - ref| against rtol * max|ref|, and — when MPL_PARITY_LOG
names a file — a JSON record of the MEASURED error per check, from which profiles/r02_parity_errors.md is produced
(tools/parity_table.py). The stated tolerance of a test is the `rtol` it passes here."""
import json
import os

BF16_ULP = 2.0 ** -8  # relative spacing of bf16 at the top of a binade (half an ulp of rounding error = 2^-9)


def close(got, ref, rtol, name=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item() if ref.numel() else 0.0
    log = os.environ.get("MPL_PARITY_LOG")
    if log:
        with open(log, "a") as f:
            f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "check": name,
                                "err": err, "scale": scale, "rel": err / scale, "rtol": rtol,
                                "n": int(ref.numel())}) + "\n")
    assert err <= rtol * scale, f"{name}: max err {err:.4e} > {rtol} * scale {scale:.4e}"
    return err / scale


def close_rows(got, ref, rtol, name="", allow_frac=0.0):
    """Row-wise variant for [.., rows, D] activations behind a hard top-1 router: a token whose two router logits are a
    near-tie goes to the other expert under ANY change of rounding (the reference's own bf16 and fp32 runs disagree on
    such tokens), and its row then differs wholesale. Rows whose max error exceeds rtol * scale are counted as flipped
    and must stay below `allow_frac` of all rows; the logged error is the max over the remaining rows."""
    got, ref = got.float().cpu(), ref.float().cpu()
    got, ref = got.reshape(-1, got.shape[-1]), ref.reshape(-1, ref.shape[-1])
    scale = max(ref.abs().max().item(), 1e-6)
    row_err = (got - ref).abs().amax(-1)
    bad = row_err > rtol * scale
    frac = bad.float().mean().item()
    err = row_err[~bad].max().item() if (~bad).any() else float("inf")
    log = os.environ.get("MPL_PARITY_LOG")
    if log:
        with open(log, "a") as f:
            f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "check": name,
                                "err": err, "scale": scale, "rel": err / scale, "rtol": rtol, "n": int(ref.numel()),
                                "rows": int(ref.shape[0]), "rows_flipped": int(bad.sum()),
                                "worst_flipped_rel": (row_err.max().item() / scale) if bad.any() else None}) + "\n")
    assert frac <= allow_frac, f"{name}: {int(bad.sum())} of {ref.shape[0]} rows exceed {rtol} * scale {scale:.4e}"
    return err / scale

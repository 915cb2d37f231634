"""Fluid model of the persistent decode kernel's weight stream (llama_decode.cu) — a planning tool, no GPU needed.

Every CTA's producer fetches its tile list in order, at most RING bytes ahead of its consumers; consumers use a tile the
moment it has arrived but cannot enter phase k+1 before every CTA has finished phase k (grid barrier, + T_BAR); the
expert tiles (P5/P6) cannot be requested before the CTA has routed (barrier of P3 + T_ROUTE).  HBM bandwidth BW is shared
equally by the CTAs that have a request outstanding, each capped at CAP.  Compares tile-to-CTA assignments:
  static   tile = blockIdx + i*G in every phase (what the kernel does today)
  rotate   the CTA's virtual index is shifted by a per-phase offset, so the CTAs that get the extra tile differ by phase
  continue one round-robin over the whole step: each phase starts at the CTA after the one that got the previous
           phase's last tile, so the CTAs with the extra tile in phase k+1 are (as far as possible) those without it in k
Usage: python tests/dev/decode_schedule_model.py [B]
"""
import sys

import numpy as np

G, D, F = 148, 4096, 11008
RING = 6 * 16384
BW, CAP = 6.5e6, 90e3          # bytes per microsecond: 6.5 TB/s total, 90 GB/s per SM
T_BAR, T_ROUTE, T_ATTN = 2.5, 3.0, 6.0   # microseconds


def phases(nact):
    t128 = 16 * D * 2
    return [("P1 qkv", 3 * (D // 16), t128, 0.0), ("P3 wo", D // 16, t128, T_ATTN),
            ("P5 gate/up", nact * ((F + 7) // 8), t128, T_ROUTE), ("P6 down", nact * (D // 16), 16 * F * 2, 0.0)]


def simulate(nact, layers, offsets):
    ph = phases(nact) * layers
    nph = len(ph)
    # per CTA: cumulative byte boundaries of its tiles per phase
    need = np.zeros((nph, G))
    for k, (_, n, tb, _) in enumerate(ph):
        if offsets == "continue":   # one round-robin over the whole step: a phase starts where the previous one stopped
            v = (np.arange(G) - sum(q[1] for q in ph[:k])) % G
        else:
            v = (np.arange(G) + offsets[k % len(offsets)] * (k + 1)) % G
        need[k] = (n // G + (v < n % G)) * tb
    cum = np.cumsum(need, axis=0)                       # bytes a CTA must have consumed by the end of phase k
    fetched = np.zeros(G)
    t, dt = 0.0, 0.05
    open_k = 0            # phase the consumers are in
    open_at = 0.0         # time it opened
    fetch_ok = np.zeros(nph, bool)   # phase tiles may be requested (router dependency)
    fetch_ok_at = np.full(nph, np.inf)
    for k in range(nph):
        if not ph[k][0].startswith("P5"):
            fetch_ok_at[k] = 0.0
    fetch_ok_at[2] = 0.0 if layers == 0 else np.inf
    ends = []
    while open_k < nph:
        # which bytes may each CTA request: up to the end of the last fetchable phase, and within the ring window
        k_lim = open_k
        while k_lim + 1 < nph and fetch_ok_at[k_lim + 1] <= t and (not ph[k_lim + 1][0].startswith("P6") or fetch_ok_at[k_lim] <= t):
            k_lim += 1
        if fetch_ok_at[open_k] > t:
            k_lim = open_k - 1
        limit = cum[k_lim] if k_lim >= 0 else np.zeros(G)
        consumed = np.minimum(fetched, cum[open_k]) if t >= open_at else (cum[open_k - 1] if open_k else np.zeros(G))
        want = np.minimum(limit, consumed + RING) - fetched
        active = want > 1e-9
        if active.any():
            rate = min(CAP, BW / active.sum())
            fetched = fetched + np.where(active, np.minimum(want, rate * dt), 0.0)
        t += dt
        if t >= open_at and (fetched >= cum[open_k] - 1e-6).all():
            ends.append(t)
            nxt = open_k + 1
            if nxt < nph:
                open_at = t + T_BAR + ph[nxt][3]
                if ph[nxt][0].startswith("P5"):
                    fetch_ok_at[nxt] = t + T_BAR + T_ROUTE
                    if nxt + 1 < nph:
                        fetch_ok_at[nxt + 1] = fetch_ok_at[nxt]
            open_k = nxt
    total = cum[-1].sum()
    return t, total


if __name__ == "__main__":
    nact = 1 if len(sys.argv) < 2 or int(sys.argv[1]) == 1 else 2
    layers = 4
    for name, offs in [("static", [0]), ("rotate +37/phase", [37]), ("continue", "continue")]:
        t, total = simulate(nact, layers, offs)
        print(f"{name:20s} {t / layers:8.1f} us/layer   {total / t / 1e6:5.2f} TB/s  ({total / t / BW:.2f} of the stream peak)")
    if len(sys.argv) > 2:
        for cap in (50e3, 60e3, 90e3, 150e3):
            CAP = cap
            print("CAP", cap / 1e3, "GB/s:", *[f"{n}={simulate(nact, layers, o)[0] / layers:.1f}" for n, o in [("static", [0]), ("rot37", [37]), ("continue", "continue")]])

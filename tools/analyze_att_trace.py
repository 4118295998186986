"""Per-step intervals (SM clocks, medians over the steady-state steps) of an attention timestamp trace (DK_ATT_TRACE):
  python tools/analyze_att_trace.py profiles/r02_att_trace_streamed_call22.txt
Columns of a trace row (one row per K/V step): 0-19 softmax group g = 2 * tile + half (5 each: S ready, max pass done,
partner max read, first part of P published, P published), 20-27 MMA issuer (4 per tile: first part of P seen, PV part 0
issued, P seen, PV part 1 + next QK^T issued)."""
import sys

import numpy as np


def summarize(path):
    a = np.loadtxt(path, dtype=np.int64)[:, :28].astype(np.float64)
    a[a <= 0] = np.nan
    J = slice(4, 30)
    S = lambda g, e: a[:, 5 * g + e]
    M = lambda w, e: a[:, 20 + 4 * w + e]
    med = lambda x: float(np.nanmedian(x[J]))
    out = {"period": med(np.diff(S(0, 0), append=np.nan)), "tile_offset": med(S(2, 0) - S(0, 0)), "tiles": []}
    for w in range(2):
        g = 2 * w
        ph = np.fmax(S(g, 3), S(g + 1, 3))
        pf = np.fmax(S(g, 4), S(g + 1, 4))
        s_next = np.roll(S(g, 0), -1)
        out["tiles"].append({
            "max_pass": med(S(g, 1) - S(g, 0)), "exchange": med(S(g, 2) - S(g, 1)), "exp_first": med(S(g, 3) - S(g, 2)),
            "exp_second": med(S(g, 4) - S(g, 3)), "softmax_total": med(pf - S(g, 0)),
            "p_half_to_issuer": med(M(w, 0) - ph), "pv0_issue": med(M(w, 1) - M(w, 0)), "p_full_to_issuer": med(M(w, 2) - pf),
            "pv1_qk_issue": med(M(w, 3) - M(w, 2)), "issue_end_to_s_ready": med(s_next - M(w, 3))})
    return out


def main(path):
    r = summarize(path)
    print("period (tile 0 S-ready to S-ready)      ", r["period"], f"-> tensor work 2048 / period = {2048 / r['period']:.2f}")
    for w, t in enumerate(r["tiles"]):
        print(f"tile {w}: max pass {t['max_pass']:6.0f} | exchange {t['exchange']:6.0f} | exp first part {t['exp_first']:6.0f} | "
              f"exp second part {t['exp_second']:6.0f} | softmax total {t['softmax_total']:6.0f}")
        print(f"        P part -> issuer sees it {t['p_half_to_issuer']:6.0f} | PV part 0 issue {t['pv0_issue']:6.0f} | P full -> "
              f"issuer sees it {t['p_full_to_issuer']:6.0f} | PV part 1 + QK issue {t['pv1_qk_issue']:6.0f} | issue end -> S ready "
              f"{t['issue_end_to_s_ready']:6.0f}")
    print("tile 1 S-ready minus tile 0 S-ready       ", r["tile_offset"])


if __name__ == "__main__":
    main(sys.argv[1])

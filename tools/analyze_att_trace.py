"""Per-step intervals (SM clocks, medians over the steady-state steps) of an attention timestamp trace (DK_ATT_TRACE)."""
import sys
import numpy as np

a = np.loadtxt(sys.argv[1], dtype=np.int64)[:, :28].astype(np.float64)
a[a <= 0] = np.nan
J = slice(4, 30)
S = lambda g, e: a[:, 5 * g + e]
M = lambda w, e: a[:, 20 + 4 * w + e]
med = lambda x: float(np.nanmedian(x[J]))
print("period (tile 0 S-ready to S-ready)      ", med(np.diff(S(0, 0), append=np.nan)))
for w in range(2):
    g = 2 * w
    ph = np.fmax(S(g, 3), S(g + 1, 3))
    pf = np.fmax(S(g, 4), S(g + 1, 4))
    s_next = np.roll(S(g, 0), -1)
    print(f"tile {w}: max pass {med(S(g,1)-S(g,0)):6.0f} | exchange {med(S(g,2)-S(g,1)):6.0f} | exp first half {med(S(g,3)-S(g,2)):6.0f} | "
          f"exp second half {med(S(g,4)-S(g,3)):6.0f} | softmax total {med(pf-S(g,0)):6.0f}")
    print(f"        P half -> issuer sees it {med(M(w,0)-ph):6.0f} | PV part 0 issue {med(M(w,1)-M(w,0)):6.0f} | P full -> issuer sees it "
          f"{med(M(w,2)-pf):6.0f} | PV part 1 + QK issue {med(M(w,3)-M(w,2)):6.0f} | issue end -> S ready {med(s_next-M(w,3)):6.0f}")
print("tile 1 S-ready minus tile 0 S-ready      ", med(S(2, 0) - S(0, 0)))

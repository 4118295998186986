"""Discrete-event model of the synchronisation protocol and arithmetic of csrc/attention_v4.cu (the experimental
one-Q-tile attention kernel), run under random interleavings.

The kernel has never executed on hardware, so this checks what can be checked without it:
  * the mbarrier protocol (phases / parities / arrival counts) neither deadlocks nor lets an agent touch a buffer in the
    wrong state: S[b] must hold the scores of step j when set j&1 reads them, P(j) when PV(j) reads it, the K / V ring
    stages must hold tile j, O must not be rescaled while a PV is in flight, m_row must be the value of step j-1;
  * the arithmetic — lazy running max shared between the two softmax sets, per-agent partial row sums tagged with the
    max they were accumulated against, O rescaling — reproduces softmax(Q K^T * scale) V.

Agents mirror the kernel's roles one to one (same barrier names, same wait parities, same order of operations):
TMA producer, MMA issuer + an in-order asynchronous tensor pipe, and 2 sets x 2 halves of softmax agents (an agent stands
for the 128 threads of one (set, half)).  `python tools/sim_attention_v4.py [seeds]`; also run by tests/test_host_cpu.py.
"""
import random
import sys

import numpy as np

KS = 2            # ring depth of the D = 128 kernel; main() also runs 4 (D = 64)
TH = 8.0          # lazy-rescale threshold (log2 units), as in the kernel


class Deadlock(Exception):
    pass


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):                 # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class NamedBar:
    def __init__(self, n):
        self.n, self.waiting, self.gen = n, 0, 0


class Sim:
    def __init__(self, S, d, seed, big_scores=False):
        self.rng = random.Random(seed)
        r = np.random.RandomState(seed)
        self.S, self.d = S, d
        self.n = (S + 127) // 128
        self.Q = r.randn(128, d).astype(np.float64)
        self.K = r.randn(S, d).astype(np.float64)
        self.V = r.randn(S, d).astype(np.float64)
        if big_scores:                        # running max keeps growing: exercises the O rescale path
            self.K *= np.linspace(0.2, 6.0, S)[:, None]
        self.sl2 = (1.0 / np.sqrt(d)) * np.log2(np.e)
        B = MBar
        self.bar = {"q_full": B(1), "pv_done": B(1), "o_full": B(1)}
        for i in range(KS):
            for nm in ("k_full", "k_empty", "v_full", "v_empty"):
                self.bar[f"{nm}{i}"] = B(1)
        for i in range(2):
            self.bar[f"s_full{i}"] = B(1)
            self.bar[f"p_full{i}"] = B(2)      # 2 agents per set (256 threads in the kernel)
            self.bar[f"m_ready{i}"] = B(2)
        self.pair = {(s): NamedBar(2) for s in range(2)}
        self.final = NamedBar(4)
        # state with tags for hazard detection
        self.kst = [None] * KS                 # tile index held by K stage
        self.vst = [None] * KS
        self.Sbuf = [None, None]               # ("S", j, array) or ("P", j, array[128,128])
        self.O = np.zeros((128, d))
        self.O_tag = -1                        # last PV accumulated
        self.m_row = [np.full(128, -np.inf), np.full(128, -np.inf)]      # one slot per set
        self.m_row_step = [-1, -1]
        self.xmax = {}
        self.lpart = {}
        self.pipe = []                         # issued, not yet executed tensor ops (in order)
        self.pv_inflight = False
        self.out = None
        self.rescales = 0

    # ------------------------------------------------------------------ agents (generators yield a wait condition)
    def wait(self, name, parity):
        b = self.bar[name]
        while not b.passed(parity):
            yield ("wait", name, parity)

    def producer(self):
        self.bar["q_full"].arrive()
        st, par = 0, 0
        for j in range(self.n):
            yield from self.wait(f"k_empty{st}", par ^ 1)
            self.kst[st] = j
            yield None
            self.bar[f"k_full{st}"].arrive()
            yield from self.wait(f"v_empty{st}", par ^ 1)
            self.vst[st] = j
            yield None
            self.bar[f"v_full{st}"].arrive()
            st += 1
            if st == KS:
                st, par = 0, par ^ 1

    def issuer(self):
        yield from self.wait("q_full", 0)
        yield from self.wait("k_full0", 0)
        self.pipe.append(("qk", 0, 0, 0))
        st, par = 0, 0
        for j in range(self.n):
            st_n = 0 if st + 1 == KS else st + 1
            par_n = par ^ 1 if st + 1 == KS else par
            if j + 1 < self.n:
                yield from self.wait(f"k_full{st_n}", par_n)
                self.pipe.append(("qk", j + 1, (j + 1) & 1, st_n))
                yield None
            yield from self.wait(f"v_full{st}", par)
            yield from self.wait(f"p_full{j & 1}", (j >> 1) & 1)
            self.pipe.append(("pv", j, j & 1, st))
            yield None
            st, par = st_n, par_n

    def tensor_pipe(self):
        """executes issued MMAs in order, asynchronously; commits fire at completion"""
        while True:
            while not self.pipe:
                yield ("idle",)
            kind, j, buf, st = self.pipe[0]
            lo, hi = j * 128, min(self.S, j * 128 + 128)
            if kind == "qk":
                assert self.kst[st] == j, f"QK({j}) read K stage {st} holding tile {self.kst[st]}"
                prev = self.Sbuf[buf]
                assert prev is None or (prev[0] == "Pdone"), f"QK({j}) overwrote S[{buf}] in state {prev and prev[:2]}"
                s = np.full((128, 128), -np.inf)
                s[:, : hi - lo] = self.Q @ self.K[lo:hi].T
                yield None                                          # takes time
                self.Sbuf[buf] = ("S", j, s)
                self.pipe.pop(0)
                self.bar[f"s_full{buf}"].arrive()
                self.bar[f"k_empty{st}"].arrive()
            else:
                assert self.vst[st] == j, f"PV({j}) read V stage {st} holding tile {self.vst[st]}"
                tag = self.Sbuf[buf]
                assert tag[0] == "P" and tag[1] == j, f"PV({j}) read S[{buf}] in state {tag[:2]}"
                assert self.O_tag == j - 1
                self.pv_inflight = True
                yield None
                v = np.zeros((128, self.d))
                v[: hi - lo] = self.V[lo:hi]
                self.O = (self.O if j > 0 else 0) + tag[2] @ v
                self.O_tag = j
                self.Sbuf[buf] = ("Pdone", j)
                self.pv_inflight = False
                self.pipe.pop(0)
                self.bar[f"v_empty{st}"].arrive()
                self.bar["pv_done"].arrive()
                if j + 1 == self.n:
                    self.bar["o_full"].arrive()

    def named(self, nb):
        gen = nb.gen
        nb.waiting += 1
        if nb.waiting == nb.n:
            nb.waiting = 0
            nb.gen += 1
        while nb.gen == gen:
            yield ("named",)

    def softmax(self, s, hh):
        m_loc = np.full(128, -np.inf)
        l_loc = np.zeros(128)
        cols = slice(hh * 64, hh * 64 + 64)
        for j in range(s, self.n, 2):
            yield from self.wait(f"s_full{s}", (j >> 1) & 1)
            tag = self.Sbuf[s]
            assert tag[0] in ("S", "Phalf") and tag[1] == j, f"set {s} step {j} read S[{s}] in state {tag[:2]}"
            sc = tag[2][:, cols].copy()
            self.xmax[(s, hh)] = sc.max(axis=1)
            yield from self.named(self.pair[s])
            m_tile = np.maximum(self.xmax[(s, hh)], self.xmax[(s, hh ^ 1)]) * self.sl2
            m_prev = np.full(128, -np.inf)
            if j > 0:
                yield from self.wait(f"m_ready{s ^ 1}", ((j - 1) >> 1) & 1)
                assert self.m_row_step[s ^ 1] == j - 1, f"step {j} read m_row of step {self.m_row_step[s ^ 1]}"
                m_prev = self.m_row[s ^ 1].copy()
            m_new = np.maximum(m_prev, m_tile)
            with np.errstate(invalid="ignore"):
                need = (m_new - m_prev) > TH
            m_use = m_prev
            if need.any():                                        # __any_sync: the whole warp takes the raise
                m_use = m_new
                if j > 0:
                    yield from self.wait("pv_done", (j - 1) & 1)
                    assert self.O_tag == j - 1 and not self.pv_inflight, "O rescaled while a PV was in flight"
                    alpha = np.exp2(m_prev - m_new)
                    oc = slice(hh * self.d // 2, (hh + 1) * self.d // 2)
                    yield None
                    self.O[:, oc] *= alpha[:, None]
                    self.rescales += 1
            if hh == 0:
                self.m_row[s] = m_use.copy()
                self.m_row_step[s] = j
            self.bar[f"m_ready{s}"].arrive()
            with np.errstate(invalid="ignore"):
                l_loc = np.where(m_loc == m_use, l_loc, l_loc * np.exp2(m_loc - m_use))
            l_loc = np.nan_to_num(l_loc, nan=0.0)
            m_loc = m_use
            yield None
            tag = self.Sbuf[s]                                     # second TMEM read of the scores
            assert tag[0] in ("S", "Phalf") and tag[1] == j
            p = np.exp2(tag[2][:, cols] * self.sl2 - m_use[:, None])
            l_loc = l_loc + p.sum(axis=1)
            cur = self.Sbuf[s]
            full = cur[2].copy()
            full[:, cols] = p                                      # P written over this half's own score columns
            self.Sbuf[s] = ("Phalf" if cur[0] == "S" else "P", j, full)
            yield None
            self.bar[f"p_full{s}"].arrive()
        yield from self.wait("o_full", 0)
        assert self.O_tag == self.n - 1
        yield from self.named(self.final)
        m_fin = self.m_row[(self.n - 1) & 1].copy()
        with np.errstate(invalid="ignore"):
            self.lpart[(s, hh)] = np.where(np.isinf(m_loc), 0.0, l_loc * np.exp2(m_loc - m_fin))
        yield from self.named(self.final)
        if (s, hh) == (0, 0):
            l = sum(self.lpart.values())
            self.out = self.O / l[:, None]

    # ------------------------------------------------------------------ scheduler
    def run(self):
        agents = {"tma": self.producer(), "mma": self.issuer(), "pipe": self.tensor_pipe()}
        for s in range(2):
            for hh in range(2):
                agents[f"sm{s}{hh}"] = self.softmax(s, hh)
        blocked = {}
        done = set()
        idle_rounds = 0
        while len(done) < len(agents) - 1:                        # the tensor pipe never finishes by itself
            name = self.rng.choice([a for a in agents if a not in done])
            try:
                res = next(agents[name])
            except StopIteration:
                done.add(name)
                continue
            progressed = res is None
            blocked[name] = res
            idle_rounds = 0 if progressed else idle_rounds + 1
            if idle_rounds > 20000:
                raise Deadlock({a: blocked.get(a) for a in agents if a not in done})
        return self.out

    def reference(self):
        s = (self.Q @ self.K.T) / np.sqrt(self.d)
        p = np.exp(s - s.max(axis=1, keepdims=True))
        return (p / p.sum(axis=1, keepdims=True)) @ self.V


def main(seeds=40):
    global KS
    worst, rescales = 0.0, 0
    for seed in range(seeds):
      for KS in (2, 4):
        for S in (1, 100, 128, 129, 256, 300, 640, 1000):
            for big in (False, True):
                sim = Sim(S, 64, seed * 131 + S + KS, big_scores=big)
                out = sim.run()
                err = float(np.abs(out - sim.reference()).max())
                worst = max(worst, err)
                rescales += sim.rescales if big else 0
                assert err < 1e-9, (seed, S, big, KS, err)
    assert rescales > 0, "the O-rescale path was never taken"
    return worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    print("attention v4 protocol model: all interleavings consistent, worst |error| =", main(n))

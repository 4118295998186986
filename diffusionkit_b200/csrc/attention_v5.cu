// K3 (v5) — attention forward softmax(scale * Q K^T) V for sm_100a.
// Replaces mx.fast.scaled_dot_product_attention (reference mlx/mmdit.py:562-563,643,687-688,736).
//
// Same tensor-side structure as v3 (attention.cu): two 128-row Q tiles per CTA, K / V tiles through TMA rings shared by
// both, S and O accumulators in TMEM, P written back over the score columns and consumed by the PV MMA straight from
// TMEM, V as an MN-major operand.  What changed is the softmax leg, which ncu showed to be the serial chain that kept
// the tensor pipe at 56 % (QK^T -> softmax -> PV per Q tile; the leg took ~2600 of a 3650-clock period):
//
//   * ONE TMEM read per step: the thread's 64 scores stay in registers from the max pass to the exponentials
//     (streamed through two 16-column register buffers, the next TMEM read in flight behind the current one);
//   * the two threads that share a row (key halves hh = 0 / 1, in different warps of the same scheduler) no longer
//     meet at a named barrier between the max pass and the exponentials: each publishes its partial row max as a
//     tagged 8-byte word in shared memory and starts the exponentials of its first 32 keys SPECULATIVELY against the
//     running max; only then does it read the partner's word.  The lazy-rescale decision (running max raised only when
//     it grows by more than 2^8) is the same function of the same two numbers in both threads; when it fires (the
//     first step and a handful of later ones) the speculative half is recomputed — the arithmetic is exactly v3's;
//   * row max through 3-input FMNMX3 (half the ALU-pipe work);
//   * one mbarrier arrival per WARP on p_full (8 instead of 256 same-address shared-memory atomics per tile and step)
//     and one polling lane per warp on s_full;
//   * persistent CTAs: grid = #SMs, each CTA walks a static list of (batch, head, Q-tile-pair) work items, heaviest
//     first, so the O write-out of one item overlaps the Q / K loads of the next and the 12th, nearly empty wave of the
//     non-persistent grid (1632 CTAs on 148 SMs) disappears into the tail of a balanced schedule.
#include "attention.cuh"

namespace dk {

// warps 0-15 softmax (tile = warp >> 3, half = (warp >> 2) & 1), 16 TMA, 17 MMA.  Registers: the SM's file is 4 x 16384,
// one quarter per scheduler, and with 18 warps two schedulers host five of them -> at most 96 registers per thread
// (a 640-thread layout with setmaxnreg, 112 for the softmax warps and 32 for the other two, was measured: correct, but
// the MMA issuer then reloads its spilled descriptors from local memory before every batch of MMAs — 794 vs 1117 TFLOP/s).
// The softmax leg is therefore written to fit 96: scores stream through two 16-column register buffers.
constexpr int ATT5_THREADS = 576;

template <int D>
struct Att5Cfg {
  static constexpr int KS = (D == 128) ? 2 : 4;      // K / V ring depth
  static constexpr int TILE_BYTES = 128 * D * 2;
  static constexpr int OFF_Q = 0;                    // Q_A, Q_B
  static constexpr int OFF_K = 2 * TILE_BYTES;
  static constexpr int OFF_V = OFF_K + KS * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_V + KS * TILE_BYTES;
  static constexpr int OFF_XCH = OFF_BAR + 256;      // [tag parity 2][tile 2][half 2][row 128] x 8 bytes (value, tag)
  static constexpr int SMEM_BYTES = OFF_XCH + 2 * 2 * 2 * 128 * 8 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int TMEM_S = 0;     // + 128 * w
  static constexpr int TMEM_O = 256;   // + 128 * w
};

__device__ __forceinline__ void xch_put(uint32_t slot, float v, uint32_t tag) {
  asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" ::"r"(slot), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// spin until the partner has published its word for `tag` (it always has, or is a few instructions away)
__device__ __forceinline__ float xch_get(uint32_t slot, uint32_t tag) {
  uint32_t v, t, spins = 0;
  for (;;) {
    asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "r"(slot) : "memory");
    if (t == tag) break;
    if (++spins == (1u << 28)) __trap();
  }
  return __uint_as_float(v);
}

// One 16-score chunk: [tail mask] -> [partial row max] -> [exponentials against mref -> packed P, partial sums].
// MASK is a template flag: only the ragged LAST K/V tile of a sequence pays for the per-element compare/select
// (ncu on v3: ISETP + FSEL of the tail mask were 26 % of all instructions issued, on every step).
template <typename H16, bool MASK, bool DO_MAX, bool DO_EXP>
__device__ __forceinline__ void att5_chunk(const uint32_t (&sc)[16], int c, int kv_valid, float sl2, float mref,
                                           float (&m2)[2], uint32_t (&pk)[2][16], float& s0, float& s1) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a0 = __uint_as_float(sc[2 * i]), a1 = __uint_as_float(sc[2 * i + 1]);
    if (MASK) {
      if (c * 16 + 2 * i >= kv_valid) a0 = -INFINITY;
      if (c * 16 + 2 * i + 1 >= kv_valid) a1 = -INFINITY;
    }
    if (DO_MAX) m2[i & 1] = fmax3(m2[i & 1], a0, a1);
    if (DO_EXP) {
      const float e0 = ex2_approx(fmaf(a0, sl2, -mref));
      const float e1 = ex2_approx(fmaf(a1, sl2, -mref));
      s0 += e0;
      s1 += e1;
      pk[c >> 1][(c & 1) * 8 + i] = H16::pack(e0, e1);
    }
  }
}

// The 64 scores of this thread stream through two 16-column register buffers (the next TMEM read is in flight while the
// current chunk is processed).
template <typename H16, bool MASK, bool DO_EXP>
__device__ __forceinline__ void att5_pass(uint32_t t_s, int kv_valid, float sl2, float mref, float (&m2)[2],
                                          uint32_t (&pk)[2][16], float& s0, float& s1) {
  uint32_t sa[16], sb[16];
  tmem_ld_32x16(t_s, sa);
  tmem_ld_wait();
  tmem_ld_32x16(t_s + 16, sb);
  att5_chunk<H16, MASK, true, DO_EXP>(sa, 0, kv_valid, sl2, mref, m2, pk, s0, s1);
  tmem_ld_wait();
  tmem_ld_32x16(t_s + 32, sa);
  att5_chunk<H16, MASK, true, DO_EXP>(sb, 1, kv_valid, sl2, mref, m2, pk, s0, s1);
  tmem_ld_wait();
  tmem_ld_32x16(t_s + 48, sb);
  att5_chunk<H16, MASK, true, DO_EXP>(sa, 2, kv_valid, sl2, mref, m2, pk, s0, s1);
  tmem_ld_wait();
  att5_chunk<H16, MASK, true, DO_EXP>(sb, 3, kv_valid, sl2, mref, m2, pk, s0, s1);
}

// One K/V step of one softmax thread: scores S (TMEM) -> probabilities P (TMEM, over the thread's own score columns),
// running max / sum updated.
//   1. pass over the scores: partial row max and — SPECULATIVELY, against the running max — the exponentials
//      (mask-free: steps that need the tail mask, and the first step of a work item, do not speculate);
//   2. publish the partial max, read the partner's (the other 64 keys of the same rows);
//   3. lazy rescale decision (running max raised only when it grows by more than 2^8): the same function of the same
//      two numbers in both threads of a row.  When it fires (rare), or when the step did not speculate, O and l are
//      rescaled and the exponentials are (re)done from the scores, which are still in TMEM because P has not been
//      written yet — this path carries the tail mask.
//   nospec: first step of a work item (running max = -inf, O not yet written) or the ragged last K/V tile.
template <typename H16, int OC>
__device__ __forceinline__ void att5_step(uint32_t t_s, uint32_t t_o, bool first, bool nospec, int kv_valid, float sl2,
                                          float& m_run, float& l_run, uint32_t my_slot, uint32_t peer_slot,
                                          uint32_t tag) {
  uint32_t pk[2][16];
  float m2[2] = {-INFINITY, -INFINITY};
  float s0 = 0.f, s1 = 0.f;
  if (!nospec)
    att5_pass<H16, false, true>(t_s, kv_valid, sl2, m_run, m2, pk, s0, s1);
  else
    att5_pass<H16, true, false>(t_s, kv_valid, sl2, m_run, m2, pk, s0, s1);
  const float mx_half = fmaxf(m2[0], m2[1]);
  xch_put(my_slot, mx_half, tag);
  const float mx = fmaxf(mx_half, xch_get(peer_slot, tag)) * sl2;
  const float m_new = fmaxf(m_run, mx);
  const bool need = (m_new - m_run) > 8.0f;
  if (__any_sync(0xffffffffu, need) || nospec) {
    const float alpha = ex2_approx(m_run - m_new);   // first step: m_run = -inf -> alpha = 0, l_run = 0
    m_run = m_new;
    l_run *= alpha;
    if (!first) {   // PV(j-1) has retired: QK(j), issued after it by the same thread, has
#pragma unroll 1
      for (int c = 0; c < OC / 16; ++c) {
        uint32_t o[16];
        tmem_ld_32x16(t_o + c * 16, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_32x16(t_o + c * 16, o);
      }
    }
    s0 = 0.f;
    s1 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t sc[16];
      tmem_ld_32x16(t_s + c * 16, sc);
      tmem_ld_wait();
      att5_chunk<H16, true, false, true>(sc, c, kv_valid, sl2, m_run, m2, pk, s0, s1);
    }
  }
  tmem_st_32x16(t_s, pk[0]);
  tmem_st_32x16(t_s + 16, pk[1]);
  l_run += s0 + s1;
}

struct Att5Work {
  int q0, head, b;
};

// Static schedule: item i of CTA c is global item c + i * gridDim.x.  Items are ordered with the Q-tile pair varying
// FASTEST inside a (batch, head): the ~148 items in flight then cover ~9 (batch, head) problems whose K / V (2.2 MB
// each at S = 4352) stay L2-resident while their 17 Q-tile pairs stream past.  (Ordering by Q-tile pair first — all 96
// (batch, head) problems in flight at once — was measured: K / V are then re-read from HBM for every item, 860 vs
// 1125 TFLOP/s.)
__device__ __forceinline__ bool att5_work(const AttParams& p, int n_qpairs, int it, Att5Work& w) {
  const int total = n_qpairs * p.heads * p.B;
  const int idx = blockIdx.x + it * gridDim.x;
  if (idx >= total) return false;
  int qp, r;
  if (p.debug & 1) {   // experiment: Q-tile-pair major
    const int bh = p.heads * p.B;
    qp = idx / bh;
    r = idx - qp * bh;
  } else {
    r = idx / n_qpairs;
    qp = idx - r * n_qpairs;
  }
  w.q0 = qp * (2 * ATT_BQ);
  w.head = r % p.heads;
  w.b = r / p.heads;
  return true;
}

template <typename T, int D>
__global__ void __launch_bounds__(ATT5_THREADS, 1)
attention_fwd_v5_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttParams p) {
  using H16 = Half16<T>;
  using Cfg = Att5Cfg<D>;
  constexpr int KS = Cfg::KS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;             // [2]  Q_w of the current work item landed
  uint64_t* q_empty = q_full + 2;          // [2]  every QK_w of the current item has retired (Q_w may be overwritten)
  uint64_t* k_full = q_empty + 2;          // [KS]
  uint64_t* k_empty = k_full + KS;         // [KS]
  uint64_t* v_full = k_empty + KS;         // [KS]
  uint64_t* v_empty = v_full + KS;         // [KS]
  uint64_t* s_full = v_empty + KS;         // [2]  QK_w(step) retired
  uint64_t* p_full = s_full + 2;           // [2]  softmax_w(step) published P_w (8 warp arrivals)
  uint64_t* o_full = p_full + 2;           // [2]  last PV_w of the item retired
  uint64_t* o_empty = o_full + 2;          // [2]  O_w of the item drained to global (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  uint32_t* xch = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_XCH);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = p.heads * D;
  const int n_tiles = (p.S + ATT_BKV - 1) / ATT_BKV;
  const int n_qpairs = (p.S + 2 * ATT_BQ - 1) / (2 * ATT_BQ);
  const int total_items = n_qpairs * p.heads * p.B;
  const int my_items = (total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], (p.debug & 4) ? 8 : 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 8);
    }
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp < 16) {   // exchange tags start at 1
    xch[threadIdx.x * 2 + 1] = 0u;
    xch[1024 + threadIdx.x * 2 + 1] = 0u;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected issue)
    int st = 0;
    uint32_t par = 0;
    Att5Work wk;
    for (int it = 0; att5_work(p, n_qpairs, it, wk); ++it) {
      const int row_base = wk.b * p.S;
      // Q tiles of this item: wait until the previous item's QK MMAs no longer read the buffers
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        mbar_wait(&q_empty[w], (it & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&q_full[w], Cfg::TILE_BYTES);
#pragma unroll
          for (int a = 0; a < D / 64; ++a)
            tma_load_2d(sQ + w * Cfg::TILE_BYTES + a * 16384, &tmQKV, &q_full[w], wk.head * D + a * 64,
                        row_base + wk.q0 + w * ATT_BQ);
        }
        __syncwarp();
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int kv_row = row_base + j * ATT_BKV;
        mbar_wait(&k_empty[st], par ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int a = 0; a < D / 64; ++a)
            tma_load_2d(sK + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &k_full[st], h + wk.head * D + a * 64, kv_row);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], par ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int a = 0; a < D / 64; ++a)
            tma_load_2d(sV + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &v_full[st], 2 * h + wk.head * D + a * 64,
                        kv_row);
        }
        __syncwarp();
        if (++st == KS) {
          st = 0;
          par ^= 1;
        }
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected issue)
    constexpr uint32_t idesc_qk = make_idesc_f16(ATT_BQ, ATT_BKV, H16::is_bf16, false, false);
    constexpr uint32_t idesc_pv = make_idesc_f16(ATT_BQ, D, H16::is_bf16, false, true);
    const uint32_t desc_hi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo0 = smem_desc_lo(smem_u32(sQ), 0);
    const uint32_t k_lo0 = smem_desc_lo(smem_u32(sK), 0);
    const uint32_t v_lo0 = smem_desc_lo(smem_u32(sV), 16384);   // MN-major: LBO = stride between 64-wide d atoms
    constexpr uint32_t TILE16 = Cfg::TILE_BYTES >> 4;
    // S_w = Q_w K^T : K = d in 16-wide slices (slice k lives in 64-column atom k>>2 at +32 B * (k&3))
    auto issue_qk = [&](int w, int st) {
      const uint32_t q_lo = q_lo0 + w * TILE16;
      const uint32_t k_lo = k_lo0 + st * TILE16;
      const uint32_t d_tmem = tmem_base + Cfg::TMEM_S + w * 128;
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
        umma_ss(d_tmem, smem_desc_join(q_lo + off, desc_hi), smem_desc_join(k_lo + off, desc_hi), idesc_qk,
                k != 0 ? 1u : 0u);
      }
      umma_commit(&s_full[w]);
    };
    // O_w += P_w V : A = P_w from TMEM (16 keys = 8 columns per slice), B = V slice of 16 key rows (2048 B apart)
    auto issue_pv = [&](int w, int st, bool first) {
      const uint32_t v_lo = v_lo0 + st * TILE16;
      const uint32_t p_tmem = tmem_base + Cfg::TMEM_S + w * 128;
      const uint32_t d_tmem = tmem_base + Cfg::TMEM_O + w * 128;
#pragma unroll
      for (int k = 0; k < ATT_BKV / 16; ++k)   // keys 0-63 -> P columns [0,32), keys 64-127 -> P columns [64,96)
        umma_ts(d_tmem, p_tmem + (k >> 2) * 64 + (k & 3) * 8, smem_desc_join(v_lo + k * (2048 >> 4), desc_hi),
                idesc_pv, (!first || k != 0) ? 1u : 0u);
    };
    int st = 0;
    uint32_t par = 0;
    uint32_t step = 0;          // global K/V step counter of this CTA: parity of s_full / p_full phases
    for (int it = 0; it < my_items; ++it) {
      // first QK of the item: needs Q, K(0); S_w is free because the previous item's last PV_w (which read P_w from
      // those columns) was issued earlier by this thread
      mbar_wait(&k_full[st], par);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        mbar_wait(&q_full[w], it & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_qk(w, st);
          if (w == 1) umma_commit(&k_empty[st]);
        }
        __syncwarp();
      }
      for (int j = 0; j < n_tiles; ++j, ++step) {
        const int st_n = (st + 1 == KS) ? 0 : st + 1;
        const uint32_t par_n = (st + 1 == KS) ? (par ^ 1) : par;
        const bool more = j + 1 < n_tiles;
        mbar_wait(&v_full[st], par);
        if (j == 0) mbar_wait(&o_empty[0], (it & 1) ^ 1);   // O_0 of the previous item drained
        mbar_wait(&p_full[0], step & 1);
        if (more) mbar_wait(&k_full[st_n], par_n);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_pv(0, st, j == 0);
          if (more) {
            issue_qk(0, st_n);
          } else {
            umma_commit(&o_full[0]);
            umma_commit(&q_empty[0]);
          }
        }
        __syncwarp();
        if (j == 0) mbar_wait(&o_empty[1], (it & 1) ^ 1);
        mbar_wait(&p_full[1], step & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_pv(1, st, j == 0);
          umma_commit(&v_empty[st]);
          if (more) {
            issue_qk(1, st_n);
            umma_commit(&k_empty[st_n]);
          } else {
            umma_commit(&o_full[1]);
            umma_commit(&q_empty[1]);
          }
        }
        __syncwarp();
        st = st_n;
        par = par_n;
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warps
    const int w = warp >> 3;          // Q tile
    const int hh = (warp >> 2) & 1;   // key half of every 128-key tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + Cfg::TMEM_S + w * 128 + hh * 64;      // this thread's 64 scores
    constexpr int OC = D / 2;                                           // O columns owned by this half
    const uint32_t t_o = t_lane + Cfg::TMEM_O + w * 128 + hh * OC;
    // Exchange words are double-buffered by tag parity: a thread may publish word t+1 (first step of the next work item,
    // which does not depend on its partner) before the partner has read word t; word t+2 needs the partner's p_full
    // arrival for t+1, which comes after its read of t.
    const uint32_t my_slot0 = smem_u32(xch + ((w * 2 + hh) * 128 + r) * 2);
    const uint32_t peer_slot0 = smem_u32(xch + ((w * 2 + (hh ^ 1)) * 128 + r) * 2);
    const float sl2 = p.scale_log2;
    const bool ragged = (p.S % ATT_BKV) != 0;   // only then does the last K/V tile need the per-element mask
    uint32_t step = 0;
    uint32_t tag = 0;
    Att5Work wk;
    for (int it = 0; att5_work(p, n_qpairs, it, wk); ++it) {
      float m_run = -INFINITY;
      float l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++step) {
        ++tag;
        mbar_wait(&s_full[w], step & 1);
        tc_fence_after();
        const int kv_valid = p.S - j * ATT_BKV - hh * 64;   // valid keys in this half (tail tile only matters)
        const uint32_t par_off = (tag & 1u) * 4096u;
        att5_step<H16, OC>(t_s, t_o, j == 0, j == 0 || (ragged && j == n_tiles - 1), kv_valid, sl2, m_run, l_run,
                           my_slot0 + par_off, peer_slot0 + par_off, tag);
        tmem_st_wait();
        tc_fence_before();
        if (p.debug & 4) {          // experiment: one arrival per warp (measured slower than 256 per-thread arrivals)
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[w]);
        } else {
          mbar_arrive(&p_full[w]);
        }
      }

      // epilogue of the item: combine the two partial row sums, O_w / l -> global (each half writes its D/2 columns)
      ++tag;
      xch_put(my_slot0 + (tag & 1u) * 4096u, l_run, tag);
      mbar_wait(&o_full[w], it & 1);
      tc_fence_after();
      const float inv_l = 1.0f / (l_run + xch_get(peer_slot0 + (tag & 1u) * 4096u, tag));
      const int s_idx = wk.q0 + w * ATT_BQ + r;
      const bool row_ok = s_idx < p.S;
      T* dst = nullptr;
      if (row_ok) {
        if (s_idx < p.split)
          dst = reinterpret_cast<T*>(p.out0) + (static_cast<long long>(wk.b) * p.split + s_idx) * p.ld0 + wk.head * D +
                hh * OC;
        else
          dst = reinterpret_cast<T*>(p.out1) +
                (static_cast<long long>(wk.b) * (p.S - p.split) + (s_idx - p.split)) * p.ld1 + wk.head * D + hh * OC;
      }
#pragma unroll
      for (int c = 0; c < OC / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32(t_o + c * 32, o);
        tmem_ld_wait();
        if (c == OC / 32 - 1) {   // O_w drained into registers: the next item's first PV_w may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_empty[w]);
        }
        if (row_ok) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            uint4 pk4;
            pk4.x = H16::pack(__uint_as_float(o[gq * 8 + 0]) * inv_l, __uint_as_float(o[gq * 8 + 1]) * inv_l);
            pk4.y = H16::pack(__uint_as_float(o[gq * 8 + 2]) * inv_l, __uint_as_float(o[gq * 8 + 3]) * inv_l);
            pk4.z = H16::pack(__uint_as_float(o[gq * 8 + 4]) * inv_l, __uint_as_float(o[gq * 8 + 5]) * inv_l);
            pk4.w = H16::pack(__uint_as_float(o[gq * 8 + 6]) * inv_l, __uint_as_float(o[gq * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c * 32 + gq * 8) = pk4;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <typename T, int D>
static int launch_attention_v5_t(dk_ctx* ctx, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  using Cfg = Att5Cfg<D>;
  auto kern = attention_fwd_v5_kernel<T, D>;
  static bool configured = false;
  if (!configured) {
    DK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const long long items = static_cast<long long>(dk_ceil_div(p.S, 2 * ATT_BQ)) * p.heads * p.B;
  const int grid = items < ctx->sm_count ? static_cast<int>(items) : ctx->sm_count;
  kern<<<grid, ATT5_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, p);
  DK_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace dk

using namespace dk;

int dk_launch_attention_v5(dk_ctx* ctx, int dtype, int d, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  if (dtype == DK_BF16) {
    if (d == 128) return launch_attention_v5_t<__nv_bfloat16, 128>(ctx, tm, p, stream);
    return launch_attention_v5_t<__nv_bfloat16, 64>(ctx, tm, p, stream);
  }
  if (d == 128) return launch_attention_v5_t<__half, 128>(ctx, tm, p, stream);
  return launch_attention_v5_t<__half, 64>(ctx, tm, p, stream);
}

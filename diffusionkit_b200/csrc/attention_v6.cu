// K3 (v6) — attention forward softmax(scale * Q K^T) V for sm_100a with DOUBLE-BUFFERED score accumulators.
// Replaces mx.fast.scaled_dot_product_attention (reference mlx/mmdit.py:562-563,643,687-688,736).
//
// Why (timestamp traces of v3): with 128-key steps S and O of the two Q tiles fill all 512 TMEM columns, P
// overwrites S, and therefore QK^T(j+1) of a tile cannot be issued before PV(j) has consumed P(j): per tile the chain
// QK^T -> softmax -> PV is serial (~3100 clocks per step against 2048 clocks of tensor work for both tiles) and the
// tensor pipe idles 35-40 % of the time.  Here a step is 64 keys: four 64-column score buffers (tile w, parity b) plus
// the two O accumulators are exactly 512 columns, QK^T(j+2) goes into the buffer PV(j) has just released, and the
// softmax of step j+1 finds its scores already computed when it finishes step j — the softmax warps never wait for the
// tensor pipe and the tensor pipe never waits for more than one softmax leg.
//
// Other differences from v3 that follow from the 64-key step:
//   * a thread owns 32 scores of a row per step (two threads per row, halves hh = 0 / 1, as in v3): they stay in the 32
//     registers of the max pass, so there is ONE TMEM read per step and no wait inside the exponential pass;
//   * the lazy O rescale at step j must wait for PV(j-1) (it was implied by the serial chain before): one commit per
//     PV on pv_done[w][b], waited for only on the rare steps that rescale;
//   * K and V travel as 64-row boxes (their own tensor map), four-stage rings.
// Barrier phases: every per-buffer barrier (s_full, p_full, pv_done) completes once per TWO steps, and neither side can
// be more than one completion ahead of the other on the same buffer, so parity (j >> 1) & 1 is unambiguous.
#include "attention.cuh"

namespace dk {

// ONE = 0: 576 threads, warps 0-15 softmax (g = warp >> 2: tile = g >> 1, half = g & 1; two threads per row), 16 TMA, 17 MMA
// ONE = 1: 320 threads, warps 0-7 softmax (tile = warp >> 2; ONE thread per row: 64 scores per step in registers, no
//          maximum exchange, half the mbarrier arrivals; up to 200 registers per thread), 8 TMA, 9 MMA
constexpr int ATT6_THREADS = 576;
constexpr int ATT6_THREADS_ONE = 320;
// ONE = 2: as 1 plus a SECOND MMA issuer warp (warp 10): one issuer per Q tile, so a tile's hand-overs no longer queue
//          behind the other tile's (ordering is only needed within a tile: PV_w(j) before QK_w(j+2), same thread)
constexpr int ATT6_THREADS_TWO = 352;
constexpr int ATT6_BKV = 64;

template <int D>
struct Att6Cfg {
  static constexpr int KS = 4;                        // K / V ring depth (64-key stages)
  static constexpr int Q_BYTES = 128 * D * 2;         // one Q tile
  static constexpr int KV_BYTES = ATT6_BKV * D * 2;   // one K or V stage
  static constexpr int ATOM_Q = 128 * 128;            // bytes between the 64-wide d atoms of a 128-row tile
  static constexpr int ATOM_KV = ATT6_BKV * 128;      // ... of a 64-row tile
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + KS * KV_BYTES;
  static constexpr int OFF_BAR = OFF_V + KS * KV_BYTES;
  static constexpr int OFF_XCH = OFF_BAR + 512;       // [step parity 2][tile 2][half 2][row 128] floats
  static constexpr int SMEM_BYTES = OFF_XCH + 2 * 2 * 2 * 128 * 4 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int TMEM_S = 0;     // + (w * 2 + b) * 64
  static constexpr int TMEM_O = 256;   // + w * 128
};

template <typename T, int D, int POLY4, int ONE>
__global__ void __launch_bounds__(ONE == 2 ? ATT6_THREADS_TWO : (ONE ? ATT6_THREADS_ONE : ATT6_THREADS), 1)
attention_fwd_v6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const AttParams p) {
  using H16 = Half16<T>;
  using Cfg = Att6Cfg<D>;
  constexpr int KS = Cfg::KS;
  constexpr int W_TMA = ONE ? 8 : 16, W_MMA = ONE ? 9 : 17;
  constexpr int N_ISS = ONE == 2 ? 2 : 1;   // MMA issuer warps: W_MMA .. W_MMA + N_ISS - 1 (issuer i serves tile i when there are two)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;             // [KS]
  uint64_t* k_empty = k_full + KS;         // [KS]
  uint64_t* v_full = k_empty + KS;         // [KS]
  uint64_t* v_empty = v_full + KS;         // [KS]
  uint64_t* s_full = v_empty + KS;         // [w][b]  QK_w(j) retired into buffer b = j & 1
  uint64_t* p_full = s_full + 4;           // [w][b]  softmax_w(j) published P_w(j) (256 arrivals)
  uint64_t* pv_done = p_full + 4;          // [w][b]  PV_w(j) retired
  uint64_t* o_full = pv_done + 4;          // [w]     last PV_w retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  float* xch = reinterpret_cast<float*>(smem + Cfg::OFF_XCH);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * ATT_BQ);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int h = p.heads * D;
  const int n_steps = (p.S + ATT6_BKV - 1) / ATT6_BKV;
  const int row_base = b * p.S;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], N_ISS);   // one commit per issuer
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], N_ISS);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], ONE ? 128 : 256);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // -------------------------------------------------------------------- TMA producer (converged warp, elected issue)
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_BYTES);
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sQ + w * Cfg::Q_BYTES + a * Cfg::ATOM_Q, &tmQ, q_full, head * D + a * 64, row_base + q0 + w * ATT_BQ);
    }
    __syncwarp();
    int st = 0;
    uint32_t par = 0;
    for (int j = 0; j < n_steps; ++j) {
      const int kv_row = row_base + j * ATT6_BKV;
      mbar_wait(&k_empty[st], par ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&k_full[st], Cfg::KV_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sK + st * Cfg::KV_BYTES + a * Cfg::ATOM_KV, &tmKV, &k_full[st], h + head * D + a * 64, kv_row);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], par ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&v_full[st], Cfg::KV_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sV + st * Cfg::KV_BYTES + a * Cfg::ATOM_KV, &tmKV, &v_full[st], 2 * h + head * D + a * 64, kv_row);
      }
      __syncwarp();
      if (++st == KS) {
        st = 0;
        par ^= 1;
      }
    }
  } else if (warp >= W_MMA && warp < W_MMA + N_ISS) {
    // -------------------------------------------------------------------- MMA issuer(s) (converged warp, elected issue)
    const int w_lo = N_ISS == 2 ? warp - W_MMA : 0;   // tiles this warp serves: [w_lo, w_hi)
    const int w_hi = N_ISS == 2 ? warp - W_MMA + 1 : 2;
    constexpr uint32_t idesc_qk = make_idesc_f16(ATT_BQ, ATT6_BKV, H16::is_bf16, false, false);
    constexpr uint32_t idesc_pv = make_idesc_f16(ATT_BQ, D, H16::is_bf16, false, true);
    const uint32_t desc_hi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo0 = smem_desc_lo(smem_u32(sQ), 0);
    const uint32_t k_lo0 = smem_desc_lo(smem_u32(sK), 0);
    const uint32_t v_lo0 = smem_desc_lo(smem_u32(sV), Cfg::ATOM_KV);   // MN-major: LBO = stride between 64-wide d atoms
    // S_w(buffer bb) = Q_w K^T over d in 16-wide slices (slice k: 64-column atom k >> 2, + 32 B * (k & 3))
    auto issue_qk = [&](int w, int st, int bb) {
      const uint32_t q_lo = q_lo0 + w * (Cfg::Q_BYTES >> 4);
      const uint32_t k_lo = k_lo0 + st * (Cfg::KV_BYTES >> 4);
      const uint32_t d_tmem = tmem_base + Cfg::TMEM_S + (w * 2 + bb) * 64;
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off_q = ((k >> 2) * Cfg::ATOM_Q + (k & 3) * 32) >> 4;
        const uint32_t off_k = ((k >> 2) * Cfg::ATOM_KV + (k & 3) * 32) >> 4;
        umma_ss(d_tmem, smem_desc_join(q_lo + off_q, desc_hi), smem_desc_join(k_lo + off_k, desc_hi), idesc_qk,
                k != 0 ? 1u : 0u);
      }
      umma_commit(&s_full[w * 2 + bb]);
    };
    // O_w += P_w V: A = P_w from TMEM (16 keys = 8 packed columns per slice: keys 0-31 in columns [0,16) of the buffer,
    // keys 32-63 in [32,48); ONE: keys 0-63 in columns [0,32)), B = V slice of 16 key rows (2048 B apart)
    auto issue_pv = [&](int w, int st, int bb, bool first) {
      const uint32_t v_lo = v_lo0 + st * (Cfg::KV_BYTES >> 4);
      const uint32_t p_tmem = tmem_base + Cfg::TMEM_S + (w * 2 + bb) * 64;
      const uint32_t d_tmem = tmem_base + Cfg::TMEM_O + w * 128;
#pragma unroll
      for (int k = 0; k < ATT6_BKV / 16; ++k)
        umma_ts(d_tmem, p_tmem + (ONE ? k * 8 : (k >> 1) * 32 + (k & 1) * 8), smem_desc_join(v_lo + k * (2048 >> 4), desc_hi), idesc_pv,
                (!first || k != 0) ? 1u : 0u);
      umma_commit(&pv_done[w * 2 + bb]);
    };
    // prologue: scores of steps 0 and 1
    mbar_wait(q_full, 0);
    for (int j = 0; j < 2 && j < n_steps; ++j) {
      mbar_wait(&k_full[j], 0);
      tc_fence_after();
      if (elect_one_sync()) {
        for (int w = w_lo; w < w_hi; ++w) issue_qk(w, j, j);
        umma_commit(&k_empty[j]);
      }
      __syncwarp();
    }
    for (int j = 0; j < n_steps; ++j) {
      const int bb = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int st_v = j % KS;
      const uint32_t par_v = (j / KS) & 1;
      const int jn = j + 2;
      const bool more = jn < n_steps;
      const int st_k = jn % KS;
      const uint32_t par_k = (jn / KS) & 1;
      mbar_wait(&v_full[st_v], par_v);
      for (int w = w_lo; w < w_hi; ++w) {
        mbar_wait(&p_full[w * 2 + bb], ph);
        if (more && w == w_lo) mbar_wait(&k_full[st_k], par_k);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_pv(w, st_v, bb, j == 0);
          if (w == w_hi - 1) umma_commit(&v_empty[st_v]);
          if (j == n_steps - 1) umma_commit(&o_full[w]);
          if (more) {
            issue_qk(w, st_k, bb);   // buffer bb: PV(j) above is ordered before it (same issuing thread)
            if (w == w_hi - 1) umma_commit(&k_empty[st_k]);
          }
        }
        __syncwarp();
      }
    }
  } else if constexpr (ONE != 0) {
    // -------------------------------------------------------------------- softmax, one thread per row
    const int w = warp >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = t_lane + Cfg::TMEM_O + w * 128;
    float m_run = -INFINITY;
    float l_run = 0.f;
    const float sl2 = p.scale_log2;
    for (int j = 0; j < n_steps; ++j) {
      const int bb = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const uint32_t t_s = t_lane + Cfg::TMEM_S + (w * 2 + bb) * 64;
      mbar_wait(&s_full[w * 2 + bb], ph);
      tc_fence_after();
      const int kv_valid = p.S - j * ATT6_BKV;
      uint32_t sa[32], sb[32];
      tmem_ld_32x32(t_s, sa);
      tmem_ld_32x32(t_s + 32, sb);
      tmem_ld_wait();
      if (kv_valid < 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) sa[i] = 0xff800000u;
          if (32 + i >= kv_valid) sb[i] = 0xff800000u;
        }
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        m4[0] = fmaxf(m4[0], __uint_as_float(sa[i]));
        m4[1] = fmaxf(m4[1], __uint_as_float(sa[i + 1]));
        m4[2] = fmaxf(m4[2], __uint_as_float(sb[i]));
        m4[3] = fmaxf(m4[3], __uint_as_float(sb[i + 1]));
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sl2;
      const float m_new = fmaxf(m_run, mx);
      const bool need = (m_new - m_run) > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = ex2_approx(m_run - m_new);
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
          mbar_wait(&pv_done[w * 2 + (bb ^ 1)], ((j - 1) >> 1) & 1);   // O is quiescent once PV(j-1) has retired
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < D / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(t_o + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(t_o + c * 16, o);
          }
        }
      }
      float ls0 = 0.f, ls1 = 0.f;
      {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(sa[2 * i]), sl2, -m_run));
          const float x1 = fmaf(__uint_as_float(sa[2 * i + 1]), sl2, -m_run);
          const float e1 = (POLY4 == 2 || (POLY4 == 1 && (i & 1))) ? ex2_poly(x1) : ex2_approx(x1);
          ls0 += e0;
          ls1 += e1;
          pk[i] = H16::pack(e0, e1);
        }
        tmem_st_32x16(t_s, pk);          // P columns [0,16): keys 0-31 (scores already in registers)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(sb[2 * i]), sl2, -m_run));
          const float x1 = fmaf(__uint_as_float(sb[2 * i + 1]), sl2, -m_run);
          const float e1 = (POLY4 == 2 || (POLY4 == 1 && (i & 1))) ? ex2_poly(x1) : ex2_approx(x1);
          ls0 += e0;
          ls1 += e1;
          pk[i] = H16::pack(e0, e1);
        }
        tmem_st_32x16(t_s + 16, pk);     // P columns [16,32): keys 32-63
      }
      l_run += ls0 + ls1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[w * 2 + bb]);
    }
    mbar_wait(&o_full[w], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int s_idx = q0 + w * ATT_BQ + r;
    const bool row_ok = s_idx < p.S;
    T* dst = nullptr;
    if (row_ok) {
      if (s_idx < p.split)
        dst = reinterpret_cast<T*>(p.out0) + (static_cast<long long>(b) * p.split + s_idx) * p.ld0 + head * D;
      else
        dst = reinterpret_cast<T*>(p.out1) +
              (static_cast<long long>(b) * (p.S - p.split) + (s_idx - p.split)) * p.ld1 + head * D;
    }
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
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
  } else {
    // -------------------------------------------------------------------- softmax warpgroups: g = 2 * w + hh
    const int g = warp >> 2;
    const int w = g >> 1;       // Q tile
    const int hh = g & 1;       // 32-key half of every 64-key step
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr int OC = D / 2;                                           // O columns owned by this half
    const uint32_t t_o = t_lane + Cfg::TMEM_O + w * 128 + hh * OC;
    // partial-maximum exchange slots, double-buffered by step parity: the write of step j+2 to a slot is separated from
    // the partner's read of step j by the named barrier of step j+1
    float* my_x0 = xch + (w * 2 + hh) * 128 + r;
    const float* peer_x0 = xch + (w * 2 + (hh ^ 1)) * 128 + r;
    float m_run = -INFINITY;
    float l_run = 0.f;
    const float sl2 = p.scale_log2;

    for (int j = 0; j < n_steps; ++j) {
      const int bb = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const uint32_t t_s = t_lane + Cfg::TMEM_S + (w * 2 + bb) * 64 + hh * 32;   // this thread's 32 scores
      mbar_wait(&s_full[w * 2 + bb], ph);
      tc_fence_after();
      const int kv_valid = p.S - j * ATT6_BKV - hh * 32;   // valid keys in this half (last step only matters)
      uint32_t sr[32];
      tmem_ld_32x32(t_s, sr);
      tmem_ld_wait();
      if (kv_valid < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= kv_valid) sr[i] = 0xff800000u;   // -inf: exp2 -> 0 below, the scores never leave the registers
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
      }
      const float mx_half = fmaxf(mx0, mx1);
      my_x0[bb * 512] = mx_half;
      named_bar_sync(1 + w, 256);   // both halves of tile w: partial maxima visible
      const float mx = fmaxf(mx_half, peer_x0[bb * 512]) * sl2;
      const float m_new = fmaxf(m_run, mx);
      const bool need = (m_new - m_run) > 8.0f;   // identical in both halves (same inputs)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = ex2_approx(m_run - m_new);
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
          // O is quiescent only once PV(j-1) has retired (with 128-key steps the serial chain implied it)
          mbar_wait(&pv_done[w * 2 + (bb ^ 1)], ((j - 1) >> 1) & 1);
          tc_fence_after();
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
      }
      float ls0 = 0.f, ls1 = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x0 = fmaf(__uint_as_float(sr[2 * i]), sl2, -m_run);
        const float x1 = fmaf(__uint_as_float(sr[2 * i + 1]), sl2, -m_run);
        const float e0 = ex2_approx(x0);
        const float e1 = (POLY4 == 2 || (POLY4 == 1 && (i & 1))) ? ex2_poly(x1) : ex2_approx(x1);
        ls0 += e0;
        ls1 += e1;
        pk[i] = H16::pack(e0, e1);
      }
      tmem_st_32x16(t_s, pk);   // P over this thread's own (already consumed) score columns
      l_run += ls0 + ls1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[w * 2 + bb]);
    }

    // epilogue: combine the two partial row sums, O_w / l -> global (each half writes its D/2 columns)
    mbar_wait(&o_full[w], 0);
    tc_fence_after();
    named_bar_sync(1 + w, 256);   // the partner has read its last maxima before the slot is reused for the sum
    my_x0[0] = l_run;
    named_bar_sync(1 + w, 256);
    const float inv_l = 1.0f / (l_run + peer_x0[0]);
    const int s_idx = q0 + w * ATT_BQ + r;
    const bool row_ok = s_idx < p.S;
    T* dst = nullptr;
    if (row_ok) {
      if (s_idx < p.split)
        dst = reinterpret_cast<T*>(p.out0) + (static_cast<long long>(b) * p.split + s_idx) * p.ld0 + head * D + hh * OC;
      else
        dst = reinterpret_cast<T*>(p.out1) +
              (static_cast<long long>(b) * (p.S - p.split) + (s_idx - p.split)) * p.ld1 + head * D + hh * OC;
    }
#pragma unroll
    for (int c = 0; c < OC / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
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

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <typename T, int D, int POLY4, int ONE>
static int launch_attention_v6p(dk_ctx* ctx, const CUtensorMap& tmQ, const CUtensorMap& tmKV, const AttParams& p,
                                cudaStream_t stream) {
  using Cfg = Att6Cfg<D>;
  auto kern = attention_fwd_v6_kernel<T, D, POLY4, ONE>;
  static bool configured = false;
  if (!configured) {
    DK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(dk_ceil_div(p.S, 2 * ATT_BQ), p.heads, p.B);
  kern<<<grid, ONE == 2 ? ATT6_THREADS_TWO : (ONE ? ATT6_THREADS_ONE : ATT6_THREADS), Cfg::SMEM_BYTES, stream>>>(tmQ, tmKV, p);
  DK_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace dk

// tmQ: the [B*S, 3*heads*d] tensor map with 128-row boxes (dk_attention_fwd builds it); tmKV: the same tensor with
// 64-row boxes.  poly: exponentials per four on the FMA pipe (0..2).  one_thread_per_row: 1 = the 320-thread form, 2 = the same with one MMA issuer warp per Q tile.
int dk_launch_attention_v6(dk_ctx* ctx, int dtype, int d, int poly, int one_thread_per_row, const CUtensorMap& tmQ,
                           const CUtensorMap& tmKV, const dk::AttParams& p, cudaStream_t stream) {
  using namespace dk;
#define DK_V6(TT, DD)                                                                          \
  do {                                                                                         \
    if (one_thread_per_row == 2) {                                                             \
      if (poly <= 0) return launch_attention_v6p<TT, DD, 0, 2>(ctx, tmQ, tmKV, p, stream);     \
      return launch_attention_v6p<TT, DD, 1, 2>(ctx, tmQ, tmKV, p, stream);                    \
    }                                                                                          \
    if (one_thread_per_row) {                                                                  \
      if (poly <= 0) return launch_attention_v6p<TT, DD, 0, 1>(ctx, tmQ, tmKV, p, stream);     \
      if (poly == 1) return launch_attention_v6p<TT, DD, 1, 1>(ctx, tmQ, tmKV, p, stream);     \
      return launch_attention_v6p<TT, DD, 2, 1>(ctx, tmQ, tmKV, p, stream);                    \
    }                                                                                          \
    if (poly <= 0) return launch_attention_v6p<TT, DD, 0, 0>(ctx, tmQ, tmKV, p, stream);       \
    if (poly == 1) return launch_attention_v6p<TT, DD, 1, 0>(ctx, tmQ, tmKV, p, stream);       \
    return launch_attention_v6p<TT, DD, 2, 0>(ctx, tmQ, tmKV, p, stream);                      \
  } while (0)
  if (dtype == DK_BF16) {
    if (d == 128) DK_V6(__nv_bfloat16, 128);
    DK_V6(__nv_bfloat16, 64);
  }
  if (d == 128) DK_V6(__half, 128);
  DK_V6(__half, 64);
#undef DK_V6
}

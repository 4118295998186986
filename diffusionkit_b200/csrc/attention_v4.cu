// EXPERIMENTAL (DK_ATTENTION_IMPL=4) — written against the analysis in DESIGN.md §8, compiled, NOT yet run on hardware.
//
// v4: ONE 128-row Q tile per CTA with a DOUBLE-BUFFERED score accumulator, two softmax warp sets alternating K/V steps.
//
// Why: in v3 each Q tile is a serial chain per K/V step, QK^T -> softmax -> PV, because P overwrites S in TMEM and two
// Q tiles (2 x S + 2 x O) fill all 512 TMEM columns; the tensor pipe idles whenever one tile's softmax leg (~2600 clk)
// does not fit into the other tile's 1024 tensor clocks (measured 56 % tensor activity).  With one Q tile per CTA the
// TMEM holds S0, S1 and O (384 columns), so QK^T of step j+1 is issued BEFORE the softmax of step j has finished, and
// step j+1's softmax set runs its TMEM read / row max / exchange while step j's set is still in its exponentials — the
// MUFU and the tensor pipe are both fed continuously.  Price: every K/V tile serves 128 query rows instead of 256
// (L2 -> SM traffic 64 B/clk/SM, the pair GEMM's level).
//
//   warps 0-7    softmax set 0 : K/V steps 0, 2, 4, ...  on S0   (quarter = warp & 3, key half hh = (warp >> 2) & 1)
//   warps 8-15   softmax set 1 : K/V steps 1, 3, 5, ...  on S1
//   warp 16      TMA producer  : Q once; K_j / V_j through KS-stage rings
//   warp 17      MMA issuer    : QK(0) | QK(1) PV(0) | QK(2) PV(1) | ...   (QK(j+1) overwrites the buffer whose P was
//                                consumed by PV(j-1): tcgen05.mma of one thread executes in issue order)
//   TMEM: S0 [0,128)  S1 [128,256)  O [256,256+D);  P(j) aliases the first half of each 64-column half of S[j&1].
//
// Running max: per row in shared memory, one slot per set (m_row[set]).  The set of step j reads the slot published by
// the set of step j-1 (mbarrier m_ready), raises it lazily (only when the tile max exceeds it by more than 2^8), publishes it
// BEFORE its exponentials so the next set is not held up, and — in the rare raise — rescales O after PV(j-1) has
// retired (mbarrier pv_done) and before it lets PV(j) start (p_full).  Row sums stay per thread, tagged with the max
// they were accumulated against, and are brought to the final max once at the end (4 partials per row).
#include "attention.cuh"

namespace dk {

constexpr int ATT4_THREADS = 576;

template <int D>
struct Att4Cfg {
  static constexpr int KS = (D == 128) ? 2 : 4;      // K / V ring depth
  static constexpr int TILE_BYTES = 128 * D * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = TILE_BYTES;
  static constexpr int OFF_V = OFF_K + KS * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_V + KS * TILE_BYTES;
  static constexpr int OFF_XCH = OFF_BAR + 512;      // float xmax[2][2][128], m_row[2][128], lpart[4][128]
  static constexpr int XCH_BYTES = (4 * 128 + 2 * 128 + 4 * 128) * 4;
  static constexpr int SMEM_BYTES = OFF_XCH + XCH_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int TMEM_S = 0;     // + 128 * (step & 1)
  static constexpr int TMEM_O = 256;
};

// 576 x 112 registers = 64512 of the SM's 65536 (__launch_bounds__(576) would make ptxas round the CTA to 640 threads
// and cap at 96 registers)
template <typename T, int D>
__global__ void __maxnreg__(112)
attention_fwd_v4_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttParams p) {
  using H16 = Half16<T>;
  using Cfg = Att4Cfg<D>;
  constexpr int KS = Cfg::KS;
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
  uint64_t* s_full = v_empty + KS;         // [2]  QK(j) retired into S[j&1]
  uint64_t* p_full = s_full + 2;           // [2]  set j&1 published P(j)                         (256 arrivals)
  uint64_t* m_ready = p_full + 2;          // [2]  set j&1 published the running max after step j (256 arrivals)
  uint64_t* pv_done = m_ready + 2;         // [1]  PV(j) retired (one phase per step; only waited on by the step-j+1 set)
  uint64_t* o_full = pv_done + 1;          // [1]  the LAST PV retired (single phase: a parity wait on pv_done would be
                                           //      ambiguous for a set that finished its steps two or more phases ago)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* xmax = reinterpret_cast<float*>(smem + Cfg::OFF_XCH);   // [set 2][half 2][row 128]
  float* m_row = xmax + 4 * 128;                                  // [set 2][row 128]: slot s is written by set s only —
                                                                  // with one slot the slower half of a set could read the
                                                                  // value its own twin has just published for this step
  float* lpart = m_row + 2 * 128;                                 // [set 2][half 2][row 128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int h = p.heads * D;
  const int n_tiles = (p.S + ATT_BKV - 1) / ATT_BKV;
  const int row_base = b * p.S;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&m_ready[i], 256);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected issue)
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
      for (int a = 0; a < D / 64; ++a)
        tma_load_2d(sQ + a * 16384, &tmQKV, q_full, head * D + a * 64, row_base + q0);
    }
    __syncwarp();
    int st = 0;
    uint32_t par = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int kv_row = row_base + j * ATT_BKV;
      mbar_wait(&k_empty[st], par ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sK + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &k_full[st], h + head * D + a * 64, kv_row);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], par ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sV + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &v_full[st], 2 * h + head * D + a * 64, kv_row);
      }
      __syncwarp();
      if (++st == KS) {
        st = 0;
        par ^= 1;
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected issue)
    constexpr uint32_t idesc_qk = make_idesc_f16(ATT_BQ, ATT_BKV, H16::is_bf16, false, false);
    constexpr uint32_t idesc_pv = make_idesc_f16(ATT_BQ, D, H16::is_bf16, false, true);
    const uint32_t desc_hi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 0);
    const uint32_t k_lo0 = smem_desc_lo(smem_u32(sK), 0);
    const uint32_t v_lo0 = smem_desc_lo(smem_u32(sV), 16384);   // MN-major: LBO = stride between 64-wide d atoms
    constexpr uint32_t TILE16 = Cfg::TILE_BYTES >> 4;
    auto issue_qk = [&](int buf, int st) {
      const uint32_t k_lo = k_lo0 + st * TILE16;
      const uint32_t d_tmem = tmem_base + Cfg::TMEM_S + buf * 128;
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
        umma_ss(d_tmem, smem_desc_join(q_lo + off, desc_hi), smem_desc_join(k_lo + off, desc_hi), idesc_qk,
                k != 0 ? 1u : 0u);
      }
      umma_commit(&s_full[buf]);
      umma_commit(&k_empty[st]);
    };
    auto issue_pv = [&](int buf, int st, bool first, bool last) {
      const uint32_t v_lo = v_lo0 + st * TILE16;
      const uint32_t p_tmem = tmem_base + Cfg::TMEM_S + buf * 128;
      const uint32_t d_tmem = tmem_base + Cfg::TMEM_O;
#pragma unroll
      for (int k = 0; k < ATT_BKV / 16; ++k)   // keys 0-63 -> P columns [0,32), keys 64-127 -> P columns [64,96)
        umma_ts(d_tmem, p_tmem + (k >> 2) * 64 + (k & 3) * 8, smem_desc_join(v_lo + k * (2048 >> 4), desc_hi),
                idesc_pv, (!first || k != 0) ? 1u : 0u);
      umma_commit(&v_empty[st]);
      umma_commit(pv_done);
      if (last) umma_commit(o_full);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one_sync()) issue_qk(0, 0);
    __syncwarp();
    int st = 0;            // ring stage of step j (K and V advance together)
    uint32_t par = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int st_n = (st + 1 == KS) ? 0 : st + 1;
      const uint32_t par_n = (st + 1 == KS) ? (par ^ 1) : par;
      if (j + 1 < n_tiles) {
        // S[(j+1)&1] last held P(j-1); PV(j-1) was issued by this thread before -> ordered
        mbar_wait(&k_full[st_n], par_n);
        tc_fence_after();
        if (elect_one_sync()) issue_qk((j + 1) & 1, st_n);
        __syncwarp();
      }
      mbar_wait(&v_full[st], par);
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) issue_pv(j & 1, st, j == 0, j + 1 == n_tiles);
      __syncwarp();
      st = st_n;
      par = par_n;
    }
  } else {
    // -------------------------------------------------------------------- softmax sets
    const int set = warp >> 3;
    const int hh = (warp >> 2) & 1;   // key half of every 128-key tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + Cfg::TMEM_S + set * 128 + hh * 64;   // this thread's 64 scores of its set's buffer
    constexpr int OC = D / 2;                                            // O columns rescaled by this half
    const uint32_t t_o = t_lane + Cfg::TMEM_O + hh * OC;
    float* my_x = xmax + (set * 2 + hh) * 128 + r;
    const float* peer_x = xmax + (set * 2 + (hh ^ 1)) * 128 + r;
    const uint32_t pair_bar = 1 + set * 4 + quarter;                     // named barriers 1..8: the two halves of a row
    float m_loc = -INFINITY;   // the max this thread's partial row sum is scaled against
    float l_loc = 0.f;
    const float sl2 = p.scale_log2;

    for (int j = set; j < n_tiles; j += 2) {
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[set], ph);
      tc_fence_after();
      const int kv_valid = p.S - j * ATT_BKV - hh * 64;   // valid keys in this half (tail tile only matters)
      // pass 1: partial row max over this half's 64 scores
      float mx_half;
      {
        uint32_t sr[2][32];
        tmem_ld_32x32(t_s, sr[0]);
        tmem_ld_32x32(t_s + 32, sr[1]);
        tmem_ld_wait();
        if (kv_valid < 64) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= kv_valid) sr[c][i] = 0xff800000u;  // -inf
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          m4[0] = fmaxf(m4[0], __uint_as_float(sr[0][i]));
          m4[1] = fmaxf(m4[1], __uint_as_float(sr[0][16 + i]));
          m4[2] = fmaxf(m4[2], __uint_as_float(sr[1][i]));
          m4[3] = fmaxf(m4[3], __uint_as_float(sr[1][16 + i]));
        }
        mx_half = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      *my_x = mx_half;
      named_bar_sync(pair_bar, 64);
      const float m_tile = fmaxf(mx_half, *peer_x) * sl2;
      // running max published by the set of step j-1
      float m_prev = -INFINITY;
      if (j > 0) {
        mbar_wait(&m_ready[set ^ 1], ((j - 1) >> 1) & 1);
        m_prev = m_row[(set ^ 1) * 128 + r];
      }
      const float m_new = fmaxf(m_prev, m_tile);
      const bool need = (m_new - m_prev) > 8.0f;   // identical in both halves of the row (same inputs); true for j == 0
      float m_use = m_prev;
      if (__any_sync(0xffffffffu, need)) {
        m_use = m_new;
        if (j > 0) {
          // O holds PV(0..j-1) scaled against m_prev: wait until PV(j-1) has retired, then rescale this half's columns.
          // PV(j) cannot start before this set arrives on p_full below.
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
          const float alpha = ex2_approx(m_prev - m_new);
#pragma unroll
          for (int c = 0; c < OC / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_o + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      // publish the running max before the exponentials: the next step's set only needs this value
      if (hh == 0) m_row[set * 128 + r] = m_use;
      mbar_arrive(&m_ready[set]);
      // this thread's partial row sum follows the max it is scaled against
      if (m_loc != m_use) {
        l_loc *= ex2_approx(m_loc - m_use);   // first own step: m_loc = -inf -> factor 0 on l_loc = 0
        m_loc = m_use;
      }
      // pass 2: P = exp2(s * sl2 - m_use) written over this half's OWN score columns
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sc[32];
        tmem_ld_32x32(t_s + c * 32, sc);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = __uint_as_float(sc[2 * i]), a1 = __uint_as_float(sc[2 * i + 1]);
          if (kv_valid < 64) {
            if (c * 32 + 2 * i >= kv_valid) a0 = -INFINITY;
            if (c * 32 + 2 * i + 1 >= kv_valid) a1 = -INFINITY;
          }
          const float e0 = ex2_approx(fmaf(a0, sl2, -m_use));
          const float e1 = ex2_approx(fmaf(a1, sl2, -m_use));
          ls0 += e0;
          ls1 += e1;
          pk[i] = H16::pack(e0, e1);
        }
        tmem_st_32x16(t_s + c * 16, pk);
      }
      l_loc += ls0 + ls1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[set]);
    }

    // epilogue: all 512 softmax threads.  Final running max, the four partial row sums, then O / l -> global
    mbar_wait(o_full, 0);
    tc_fence_after();
    named_bar_sync(9, 512);                     // every m_row write (made before the writers' last exponentials) is visible
    const float m_fin = m_row[((n_tiles - 1) & 1) * 128 + r];
    lpart[(set * 2 + hh) * 128 + r] = (m_loc == -INFINITY) ? 0.f : l_loc * ex2_approx(m_loc - m_fin);
    named_bar_sync(9, 512);
    const float inv_l = 1.0f / (lpart[r] + lpart[128 + r] + lpart[256 + r] + lpart[384 + r]);
    const int s_idx = q0 + r;
    const bool row_ok = s_idx < p.S;
    constexpr int OQ = D / 4;                   // output columns written by this thread
    const int col0 = (set * 2 + hh) * OQ;
    T* dst = nullptr;
    if (row_ok) {
      if (s_idx < p.split)
        dst = reinterpret_cast<T*>(p.out0) + (static_cast<long long>(b) * p.split + s_idx) * p.ld0 + head * D + col0;
      else
        dst = reinterpret_cast<T*>(p.out1) +
              (static_cast<long long>(b) * (p.S - p.split) + (s_idx - p.split)) * p.ld1 + head * D + col0;
    }
    const uint32_t t_out = t_lane + Cfg::TMEM_O + col0;
    if constexpr (OQ == 32) {
      uint32_t o[32];
      tmem_ld_32x32(t_out, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          uint4 pk4;
          pk4.x = H16::pack(__uint_as_float(o[gq * 8 + 0]) * inv_l, __uint_as_float(o[gq * 8 + 1]) * inv_l);
          pk4.y = H16::pack(__uint_as_float(o[gq * 8 + 2]) * inv_l, __uint_as_float(o[gq * 8 + 3]) * inv_l);
          pk4.z = H16::pack(__uint_as_float(o[gq * 8 + 4]) * inv_l, __uint_as_float(o[gq * 8 + 5]) * inv_l);
          pk4.w = H16::pack(__uint_as_float(o[gq * 8 + 6]) * inv_l, __uint_as_float(o[gq * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + gq * 8) = pk4;
        }
      }
    } else {
      uint32_t o[16];
      tmem_ld_32x16(t_out, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          uint4 pk4;
          pk4.x = H16::pack(__uint_as_float(o[gq * 8 + 0]) * inv_l, __uint_as_float(o[gq * 8 + 1]) * inv_l);
          pk4.y = H16::pack(__uint_as_float(o[gq * 8 + 2]) * inv_l, __uint_as_float(o[gq * 8 + 3]) * inv_l);
          pk4.z = H16::pack(__uint_as_float(o[gq * 8 + 4]) * inv_l, __uint_as_float(o[gq * 8 + 5]) * inv_l);
          pk4.w = H16::pack(__uint_as_float(o[gq * 8 + 6]) * inv_l, __uint_as_float(o[gq * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + gq * 8) = pk4;
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
static int launch_attention_v4_t(dk_ctx* ctx, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  using Cfg = Att4Cfg<D>;
  auto kern = attention_fwd_v4_kernel<T, D>;
  static bool configured = false;
  if (!configured) {
    DK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(dk_ceil_div(p.S, ATT_BQ), p.heads, p.B);
  kern<<<grid, ATT4_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, p);
  DK_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace dk

using namespace dk;

int dk_launch_attention_v4(dk_ctx* ctx, int dtype, int d, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  if (dtype == DK_BF16) {
    if (d == 128) return launch_attention_v4_t<__nv_bfloat16, 128>(ctx, tm, p, stream);
    return launch_attention_v4_t<__nv_bfloat16, 64>(ctx, tm, p, stream);
  }
  if (d == 128) return launch_attention_v4_t<__half, 128>(ctx, tm, p, stream);
  return launch_attention_v4_t<__half, 64>(ctx, tm, p, stream);
}

// Earlier generations of the attention forward kernel, kept selectable (DK_ATTENTION_IMPL=1 / 2 / 2a) for same-box A/B
// measurements against the current kernel in attention.cu and exercised by the GPU checks.
//
// One CTA = 128 query rows of one (batch, head).  192 threads:
//   warp 0     TMA producer : Q once, then K_j / V_j tiles (128 keys) through 2-stage rings
//   warp 1     MMA issuer   : S_j = Q K_j^T   (M128 x N128 x K=d, both operands K-major)      -> TMEM S[j&1]
//                             O  += P_j V_j   (M128 x N=d x K128; P K-major from smem, V MN-major as loaded) -> TMEM O
//                             issue order QK_0, QK_1, PV_0, QK_2, PV_1, ... so QK_{j+1} overlaps softmax_j
//   warps 2-5  softmax      : thread = query row (TMEM lane).  tcgen05.ld S row, running max / sum in the log2 domain,
//                             P written to smem in the swizzled K-major layout the MMA expects, O rescaled in TMEM
//                             only when the running max grew by more than 2^8 (lazy rescale), final O / l -> global.
#include "attention.cuh"

namespace dk {

constexpr int ATT_THREADS = 192;

template <int D>
struct AttCfg {
  static constexpr int TILE_BYTES = 128 * D * 2;     // Q / K / V tile
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = TILE_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;
  static constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_COLS = 512;  // S0 [0,128) S1 [128,256) O [256, 256+D)
  static constexpr int TMEM_O = 256;
};

template <typename T, int D>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttParams p) {
  using H16 = Half16<T>;
  using Cfg = AttCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sP = smem + Cfg::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* p_empty = bars + 15;  // [2]  (also: "PV_j retired")
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int h = p.heads * D;
  const int n_tiles = (p.S + ATT_BKV - 1) / ATT_BKV;
  const int row_base = b * p.S;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      mbar_arrive_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
      for (int a = 0; a < D / 64; ++a)
        tma_load_2d(sQ + a * 16384, &tmQKV, q_full, head * D + a * 64, row_base + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t par = static_cast<uint32_t>(j >> 1) & 1u;
        const int kv_row = row_base + j * ATT_BKV;
        mbar_wait(&k_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sK + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &k_full[st], h + head * D + a * 64, kv_row);
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sV + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &v_full[st], 2 * h + head * D + a * 64, kv_row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc_qk = make_idesc_f16(ATT_BQ, ATT_BKV, H16::is_bf16, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(ATT_BQ, D, H16::is_bf16, false, true);
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_qk = [&](int i) {
        const int st = i & 1;
        const uint32_t par = static_cast<uint32_t>(i >> 1) & 1u;
        mbar_wait(&k_full[st], par);
        mbar_wait(&s_empty[st], par ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * Cfg::TILE_BYTES);
        const uint32_t d_tmem = tmem_base + st * ATT_BKV;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(d_tmem, make_smem_desc_sw128(q_addr + off, 0, 1024), make_smem_desc_sw128(k_addr + off, 0, 1024),
                  idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };
      auto issue_pv = [&](int i) {
        const int st = i & 1;
        const uint32_t par = static_cast<uint32_t>(i >> 1) & 1u;
        mbar_wait(&v_full[st], par);
        mbar_wait(&p_full[st], par);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + st * Cfg::P_BYTES);
        const uint32_t v_addr = smem_u32(sV + st * Cfg::TILE_BYTES);
        const uint32_t d_tmem = tmem_base + Cfg::TMEM_O;
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k) {
          const uint32_t a_off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(d_tmem, make_smem_desc_sw128(p_addr + a_off, 0, 1024),
                  make_smem_desc_sw128(v_addr + k * 2048, 16384, 1024), idesc_pv, (i != 0 || k != 0) ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(&p_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        issue_pv(j);
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax / correction / epilogue
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row in the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float m_run = -INFINITY;  // running max (log2 domain) that P and l are expressed against
    float l_run = 0.f;
    const float sl2 = p.scale_log2;

    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const uint32_t par = static_cast<uint32_t>(j >> 1) & 1u;
      mbar_wait(&s_full[st], par);
      tc_fence_after();
      uint32_t sr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_lane + st * ATT_BKV + c * 32, sr[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[st]);  // S[st] may be overwritten by QK_{j+2}

      // scores in the log2 domain; keys beyond the sequence end (tail tile, or rows of the next batch / TMA zero fill)
      // are masked out
      const int kv_valid = p.S - j * ATT_BKV;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = __uint_as_float(sr[c][i]) * sl2;
          if (c * 32 + i >= kv_valid) v = -INFINITY;
          sr[c][i] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      }
      // lazy rescale: keep the stale max while the new one is within 2^8 of it (P stays <= 256, fine for fp32 sums
      // and 16-bit P); decision is warp-uniform because the TMEM ld/st below are warp-collective
      const float m_new = fmaxf(m_run, mx);
      const bool need = (m_new - m_run) > 8.0f;  // also true on the first tile (m_run = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = exp2f(m_run - m_new);  // 0 on the first tile
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
          // O must hold PV_{j-1} before it is rescaled
          mbar_wait(&p_empty[(j - 1) & 1], static_cast<uint32_t>((j - 1) >> 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_lane + Cfg::TMEM_O + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_lane + Cfg::TMEM_O + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      // P = exp2(s - m_run) -> 16-bit, into smem in the K-major 128B-swizzled layout (two 64-key atoms of 128 rows)
      mbar_wait(&p_empty[st], par ^ 1);  // PV_{j-2} has finished reading P[st]
      uint8_t* p_row = sP + st * Cfg::P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 8 keys per 16-byte chunk
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = exp2f(__uint_as_float(sr[c][g * 8 + i]) - m_run);
            lsum += e[i];
          }
          uint4 pk;
          pk.x = H16::pack(e[0], e[1]);
          pk.y = H16::pack(e[2], e[3]);
          pk.z = H16::pack(e[4], e[5]);
          pk.w = H16::pack(e[6], e[7]);
          const int key = c * 32 + g * 8;
          const int atom = key >> 6;
          const int chunk = (key & 63) >> 3;
          *reinterpret_cast<uint4*>(p_row + atom * 16384 + ((chunk ^ (r & 7)) << 4)) = pk;
        }
      }
      l_run += lsum;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      mbar_arrive(&p_full[st]);
    }

    // epilogue: O / l -> global.  PV_{n-1} retired <=> p_empty[(n-1)&1] completed its phase.
    mbar_wait(&p_empty[(n_tiles - 1) & 1], static_cast<uint32_t>((n_tiles - 1) >> 1) & 1u);
    tc_fence_after();
    const int s_idx = q0 + r;
    const bool row_ok = s_idx < p.S;
    T* dst = nullptr;
    if (row_ok) {
      if (s_idx < p.split)
        dst = reinterpret_cast<T*>(p.out0) + (static_cast<long long>(b) * p.split + s_idx) * p.ld0 + head * D;
      else
        dst = reinterpret_cast<T*>(p.out1) +
              (static_cast<long long>(b) * (p.S - p.split) + (s_idx - p.split)) * p.ld1 + head * D;
    }
    const float inv_l = 1.0f / l_run;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_lane + Cfg::TMEM_O + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = H16::pack(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          pk.y = H16::pack(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          pk.z = H16::pack(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          pk.w = H16::pack(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ================================================================================================
// v2: two 128-row Q tiles per CTA, one softmax warpgroup per tile (ping-pong), P kept in TMEM.
//
//   warp 0      TMA producer : Q_A, Q_B once; K_j / V_j through KS-stage rings (shared by both Q tiles)
//   warp 1      MMA issuer   : QK_A(0) QK_B(0) | PV_A(0) QK_A(1) | PV_B(0) QK_B(1) | PV_A(1) QK_A(2) | ...
//                              while warpgroup A runs softmax_A(j) the tensor pipe executes PV_B(j-1) + QK_B(j)
//   warps 4-7   softmax A    : thread = query row.  S row <- tcgen05.ld, exp2 in the log2 domain with lazy rescale,
//   warps 8-11  softmax B      P (16-bit, two per column) -> tcgen05.st over the first 64 columns of its own S
//                              accumulator; PV then reads A = P from TMEM (no smem round trip for P) and B = V as the
//                              MN-major tile TMA delivered.
//   TMEM: S_A [0,128) S_B [128,256) O_A [256,256+D) O_B [384,384+D);  P_w aliases S_w[0,64).
// Ordering relies on tcgen05.mma of one thread executing in issue order: QK_w(j+1) (overwrites S_w/P_w) is issued
// after PV_w(j) (reads P_w); softmax_w(j) finished reading S_w(j) before it published P_w(j).
// ================================================================================================
constexpr int ATT2_THREADS = 384;

// ---- v2a: the first v2 implementation (single-lane issue loops, one TMEM pass, all-MUFU exponentials), kept
// ---- selectable (DK_ATTENTION_IMPL=2a) for same-box A/B measurements against the current kernel
template <typename T, int D>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attention_fwd_v2a_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttParams p) {
  using H16 = Half16<T>;
  using Cfg = Att2Cfg<D>;
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
  uint64_t* s_full = v_empty + KS;         // [2]  QK_w(j) retired
  uint64_t* p_full = s_full + 2;           // [2]  softmax_w(j) published P_w(j) (128 arrivals)
  uint64_t* o_full = p_full + 2;           // [2]  PV_w(n-1) retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * ATT_BQ);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int h = p.heads * D;
  const int n_tiles = (p.S + ATT_BKV - 1) / ATT_BKV;
  const int row_base = b * p.S;

  if (warp == 0 && lane == 0) {
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
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // producer warpgroup: give its registers to the softmax warpgroups
    setmaxnreg_dec<112>();
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      mbar_arrive_expect_tx(q_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sQ + w * Cfg::TILE_BYTES + a * 16384, &tmQKV, q_full, head * D + a * 64,
                      row_base + q0 + w * ATT_BQ);
      int st = 0;
      uint32_t par = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int kv_row = row_base + j * ATT_BKV;
        mbar_wait(&k_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sK + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &k_full[st], h + head * D + a * 64, kv_row);
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D / 64; ++a)
          tma_load_2d(sV + st * Cfg::TILE_BYTES + a * 16384, &tmQKV, &v_full[st], 2 * h + head * D + a * 64, kv_row);
        if (++st == KS) {
          st = 0;
          par ^= 1;
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc_qk = make_idesc_f16(ATT_BQ, ATT_BKV, H16::is_bf16, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(ATT_BQ, D, H16::is_bf16, false, true);
      auto issue_qk = [&](int w, int st) {
        const uint32_t q_addr = smem_u32(sQ + w * Cfg::TILE_BYTES);
        const uint32_t k_addr = smem_u32(sK + st * Cfg::TILE_BYTES);
        const uint32_t d_tmem = tmem_base + Cfg::TMEM_S + w * 128;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(d_tmem, make_smem_desc_sw128(q_addr + off, 0, 1024), make_smem_desc_sw128(k_addr + off, 0, 1024),
                  idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[w]);
      };
      auto issue_pv = [&](int w, int st, bool first) {
        const uint32_t v_addr = smem_u32(sV + st * Cfg::TILE_BYTES);
        const uint32_t p_tmem = tmem_base + Cfg::TMEM_S + w * 128;   // P_w: 16-bit pairs in S_w's first 64 columns
        const uint32_t d_tmem = tmem_base + Cfg::TMEM_O + w * 128;
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k)
          umma_ts(d_tmem, p_tmem + k * 8, make_smem_desc_sw128(v_addr + k * 2048, 16384, 1024), idesc_pv,
                  (!first || k != 0) ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      umma_commit(&k_empty[0]);
      int st = 0;
      uint32_t par = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int st_n = (st + 1 == KS) ? 0 : st + 1;
        const uint32_t par_n = (st + 1 == KS) ? (par ^ 1) : par;
        const bool more = j + 1 < n_tiles;
        mbar_wait(&v_full[st], par);
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
        issue_pv(0, st, j == 0);
        if (!more) umma_commit(&o_full[0]);
        if (more) {
          mbar_wait(&k_full[st_n], par_n);
          tc_fence_after();
          issue_qk(0, st_n);
        }
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        issue_pv(1, st, j == 0);
        umma_commit(&v_empty[st]);
        if (!more) umma_commit(&o_full[1]);
        if (more) {
          issue_qk(1, st_n);
          umma_commit(&k_empty[st_n]);
        }
        st = st_n;
        par = par_n;
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warpgroups (w = 0: A, 1: B)
    setmaxnreg_inc<192>();
    const int w = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + Cfg::TMEM_S + w * 128;
    const uint32_t t_o = t_lane + Cfg::TMEM_O + w * 128;
    float m_run = -INFINITY;
    float l_run = 0.f;
    const float sl2 = p.scale_log2;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[w], j & 1);
      tc_fence_after();
      const int kv_valid = p.S - j * ATT_BKV;   // keys beyond the sequence end exist only on the tail tile

      // pass 1 over the S row: running max of the raw scores (the positive scale commutes with max).
      // TMEM reads are cheap (16 TB/s per SM), so the row is read twice instead of being held in 128 registers.
      float mx;
      {
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, sr[c]);
        tmem_ld_wait();
        if (kv_valid < ATT_BKV) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= kv_valid) sr[c][i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[2][i]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[3][i]));
        }
        mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
      }
      const float m_new = fmaxf(m_run, mx);
      const bool need = (m_new - m_run) > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = ex2_approx(m_run - m_new);
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
          // PV_w(j-1) was issued before QK_w(j), whose completion s_full signalled: O_w is quiescent
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_o + c * 32, o);
          }
        }
      }
      // pass 2: P = exp2(s * sl2 - m_run), packed two 16-bit values per 32-bit TMEM column (key 2i in the low half),
      // written over the first 64 columns of this warpgroup's own S accumulator
      float ls0 = 0.f, ls1 = 0.f;
      uint32_t pk[2][32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sa[32], sb[32];
        tmem_ld_32x32(t_s + half * 64, sa);
        tmem_ld_32x32(t_s + half * 64 + 32, sb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = __uint_as_float(sa[2 * i]), a1 = __uint_as_float(sa[2 * i + 1]);
          float b0 = __uint_as_float(sb[2 * i]), b1 = __uint_as_float(sb[2 * i + 1]);
          if (kv_valid < ATT_BKV) {
            const int k0 = half * 64 + 2 * i;
            if (k0 >= kv_valid) a0 = -INFINITY;
            if (k0 + 1 >= kv_valid) a1 = -INFINITY;
            if (k0 + 32 >= kv_valid) b0 = -INFINITY;
            if (k0 + 33 >= kv_valid) b1 = -INFINITY;
          }
          const float e0 = ex2_approx(fmaf(a0, sl2, -m_run));
          const float e1 = ex2_approx(fmaf(a1, sl2, -m_run));
          const float f0 = ex2_approx(fmaf(b0, sl2, -m_run));
          const float f1 = ex2_approx(fmaf(b1, sl2, -m_run));
          ls0 += e0 + f0;
          ls1 += e1 + f1;
          pk[half][i] = H16::pack(e0, e1);
          pk[half][16 + i] = H16::pack(f0, f1);
        }
      }
      // all reads of S_w by this thread are complete (wait::ld above) before P overwrites its first 64 columns
      tmem_st_32x32(t_s, pk[0]);
      tmem_st_32x32(t_s + 32, pk[1]);
      l_run += ls0 + ls1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[w]);
    }

    // epilogue: O_w / l -> global
    mbar_wait(&o_full[w], 0);
    tc_fence_after();
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
    const float inv_l = 1.0f / l_run;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = H16::pack(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          pk.y = H16::pack(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          pk.z = H16::pack(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          pk.w = H16::pack(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


template <typename T, int D>
static int launch_attention_v2a(dk_ctx* ctx, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  using Cfg = Att2Cfg<D>;
  auto kern = attention_fwd_v2a_kernel<T, D>;
  static bool configured = false;
  if (!configured) {
    DK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(dk_ceil_div(p.S, 2 * ATT_BQ), p.heads, p.B);
  kern<<<grid, ATT2_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, p);
  DK_LAUNCH_CHECK(ctx);
  return 0;
}

// POLY: how many of every four exponentials are evaluated on the FMA pipe (ex2_poly) instead of MUFU.EX2
template <typename T, int D, int POLY>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attention_fwd_v2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttParams p) {
  using H16 = Half16<T>;
  using Cfg = Att2Cfg<D>;
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
  uint64_t* s_full = v_empty + KS;         // [2]  QK_w(j) retired
  uint64_t* p_full = s_full + 2;           // [2]  softmax_w(j) published P_w(j) (128 arrivals)
  uint64_t* o_full = p_full + 2;           // [2]  PV_w(n-1) retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * ATT_BQ);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int h = p.heads * D;
  const int n_tiles = (p.S + ATT_BKV - 1) / ATT_BKV;
  const int row_base = b * p.S;

  if (warp == 0 && lane == 0) {
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
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // producer warpgroup: give its registers to the softmax warpgroups
    setmaxnreg_dec<112>();
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer (converged warp, elected issue)
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
          for (int a = 0; a < D / 64; ++a)
            tma_load_2d(sQ + w * Cfg::TILE_BYTES + a * 16384, &tmQKV, q_full, head * D + a * 64,
                        row_base + q0 + w * ATT_BQ);
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
    } else if (warp == 1) {
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
        for (int k = 0; k < ATT_BKV / 16; ++k)
          umma_ts(d_tmem, p_tmem + k * 8, smem_desc_join(v_lo + k * (2048 >> 4), desc_hi), idesc_pv,
                  (!first || k != 0) ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (elect_one_sync()) {
        issue_qk(0, 0);
        issue_qk(1, 0);
        umma_commit(&k_empty[0]);
      }
      __syncwarp();
      int st = 0;
      uint32_t par = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int st_n = (st + 1 == KS) ? 0 : st + 1;
        const uint32_t par_n = (st + 1 == KS) ? (par ^ 1) : par;
        const bool more = j + 1 < n_tiles;
        mbar_wait(&v_full[st], par);
        if (p.debug < 4) mbar_wait(&p_full[0], j & 1);   // debug 4: tensor-side throughput without the softmax round trip
        if (more) mbar_wait(&k_full[st_n], par_n);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_pv(0, st, j == 0);
          if (!more) umma_commit(&o_full[0]);
          if (more) issue_qk(0, st_n);
        }
        __syncwarp();
        if (p.debug < 4) mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_pv(1, st, j == 0);
          umma_commit(&v_empty[st]);
          if (!more) umma_commit(&o_full[1]);
          if (more) {
            issue_qk(1, st_n);
            umma_commit(&k_empty[st_n]);
          }
        }
        __syncwarp();
        st = st_n;
        par = par_n;
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warpgroups (w = 0: A, 1: B)
    setmaxnreg_inc<192>();
    const int w = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + Cfg::TMEM_S + w * 128;
    const uint32_t t_o = t_lane + Cfg::TMEM_O + w * 128;
    float m_run = -INFINITY;
    float l_run = 0.f;
    const float sl2 = p.scale_log2;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[w], j & 1);
      tc_fence_after();
      if (p.debug >= 3) {   // timing experiment: pure barrier round trip, no TMEM traffic
        tc_fence_before();
        mbar_arrive(&p_full[w]);
        continue;
      }
      const int kv_valid = p.S - j * ATT_BKV;   // keys beyond the sequence end exist only on the tail tile

      // pass 1 over the S row: running max of the raw scores (the positive scale commutes with max).
      // TMEM reads are cheap (16 TB/s per SM), so the row is read twice instead of being held in 128 registers.
      float mx = 0.f;
      if (p.debug < 2) {
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, sr[c]);
        tmem_ld_wait();
        if (kv_valid < ATT_BKV) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= kv_valid) sr[c][i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[2][i]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[3][i]));
        }
        mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
      }
      const float m_new = fmaxf(m_run, mx);
      const bool need = (m_new - m_run) > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = ex2_approx(m_run - m_new);
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
          // PV_w(j-1) was issued before QK_w(j), whose completion s_full signalled: O_w is quiescent
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_o + c * 32, o);
          }
        }
      }
      // pass 2: P = exp2(s * sl2 - m_run), packed two 16-bit values per 32-bit TMEM column (key 2i in the low half),
      // written over the first 64 columns of this warpgroup's own S accumulator
      float ls0 = 0.f, ls1 = 0.f;
      uint32_t pk[2][32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sa[32], sb[32];
        tmem_ld_32x32(t_s + half * 64, sa);
        tmem_ld_32x32(t_s + half * 64 + 32, sb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = __uint_as_float(sa[2 * i]), a1 = __uint_as_float(sa[2 * i + 1]);
          float b0 = __uint_as_float(sb[2 * i]), b1 = __uint_as_float(sb[2 * i + 1]);
          if (kv_valid < ATT_BKV) {
            const int k0 = half * 64 + 2 * i;
            if (k0 >= kv_valid) a0 = -INFINITY;
            if (k0 + 1 >= kv_valid) a1 = -INFINITY;
            if (k0 + 32 >= kv_valid) b0 = -INFINITY;
            if (k0 + 33 >= kv_valid) b1 = -INFINITY;
          }
          // (packed fma.rn.f32x2 / add.rn.f32x2 were measured slower here: 799 vs 947 TFLOP/s at C4)
          const float xa0 = fmaf(a0, sl2, -m_run), xa1 = fmaf(a1, sl2, -m_run);
          const float xb0 = fmaf(b0, sl2, -m_run), xb1 = fmaf(b1, sl2, -m_run);
          float e0, e1, f0, f1;
          if (p.debug >= 1) {
            e0 = xa0; e1 = xa1; f0 = xb0; f1 = xb1;
          } else {
            e0 = ex2_approx(xa0);
            e1 = (POLY >= 2) ? ex2_poly(xa1) : ex2_approx(xa1);
            f0 = ex2_approx(xb0);
            f1 = (POLY >= 1) ? ex2_poly(xb1) : ex2_approx(xb1);
          }
          ls0 += e0 + f0;
          ls1 += e1 + f1;
          pk[half][i] = H16::pack(e0, e1);
          pk[half][16 + i] = H16::pack(f0, f1);
        }
      }
      // all reads of S_w by this thread are complete (wait::ld above) before P overwrites its first 64 columns
      tmem_st_32x32(t_s, pk[0]);
      tmem_st_32x32(t_s + 32, pk[1]);
      l_run += ls0 + ls1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[w]);
    }

    // epilogue: O_w / l -> global
    mbar_wait(&o_full[w], 0);
    tc_fence_after();
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
    const float inv_l = 1.0f / l_run;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = H16::pack(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          pk.y = H16::pack(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          pk.z = H16::pack(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          pk.w = H16::pack(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <typename T, int D, int POLY>
static int launch_attention_v2p(dk_ctx* ctx, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  using Cfg = Att2Cfg<D>;
  auto kern = attention_fwd_v2_kernel<T, D, POLY>;
  static bool configured = false;
  if (!configured) {
    DK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(dk_ceil_div(p.S, 2 * ATT_BQ), p.heads, p.B);
  kern<<<grid, ATT2_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, p);
  DK_LAUNCH_CHECK(ctx);
  return 0;
}

template <typename T, int D>
static int launch_attention(dk_ctx* ctx, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  using Cfg = AttCfg<D>;
  auto kern = attention_fwd_kernel<T, D>;
  static bool configured = false;
  if (!configured) {
    DK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(dk_ceil_div(p.S, ATT_BQ), p.heads, p.B);
  kern<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, p);
  DK_LAUNCH_CHECK(ctx);
  return 0;
}

template <typename T, int D>
static int launch_attention_v2(dk_ctx* ctx, const CUtensorMap& tm, const AttParams& p, cudaStream_t stream) {
  static const int poly = [] {
    const char* e = getenv("DK_ATT_POLY");
    return e ? atoi(e) : 0;   // same-box A/B at the C4 shape: 0 -> 1099, 1 -> 842 TFLOP/s (MUFU is not the limiter)
  }();
  if (poly <= 0) return launch_attention_v2p<T, D, 0>(ctx, tm, p, stream);
  if (poly == 1) return launch_attention_v2p<T, D, 1>(ctx, tm, p, stream);
  return launch_attention_v2p<T, D, 2>(ctx, tm, p, stream);
}

}  // namespace dk

using namespace dk;

int dk_launch_attention_legacy(dk_ctx* ctx, int impl, int dtype, int d, const CUtensorMap& tm, const AttParams& p,
                               cudaStream_t stream) {
  if (impl == 3) {
    if (dtype == DK_BF16) {
      if (d == 128) return launch_attention_v2a<__nv_bfloat16, 128>(ctx, tm, p, stream);
      return launch_attention_v2a<__nv_bfloat16, 64>(ctx, tm, p, stream);
    }
    if (d == 128) return launch_attention_v2a<__half, 128>(ctx, tm, p, stream);
    return launch_attention_v2a<__half, 64>(ctx, tm, p, stream);
  }
  if (impl == 1) {
    if (dtype == DK_BF16) {
      if (d == 128) return launch_attention<__nv_bfloat16, 128>(ctx, tm, p, stream);
      return launch_attention<__nv_bfloat16, 64>(ctx, tm, p, stream);
    }
    if (d == 128) return launch_attention<__half, 128>(ctx, tm, p, stream);
    return launch_attention<__half, 64>(ctx, tm, p, stream);
  }
  if (dtype == DK_BF16) {
    if (d == 128) return launch_attention_v2<__nv_bfloat16, 128>(ctx, tm, p, stream);
    return launch_attention_v2<__nv_bfloat16, 64>(ctx, tm, p, stream);
  }
  if (d == 128) return launch_attention_v2<__half, 128>(ctx, tm, p, stream);
  return launch_attention_v2<__half, 64>(ctx, tm, p, stream);
}

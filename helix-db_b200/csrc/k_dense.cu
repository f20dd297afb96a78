// k_dense.cu — tensor-core batched distance path (configs C4/C5): exhaustive top-k of B queries against every row.
//
// The only place on this hot path where the work is a dense contraction: S[b][i] = <q_b, x_i> for B = 1024..4096
// queries against the whole shard is a (B x d) x (d x N) GEMM with arithmetic intensity ~B flop/byte, far beyond the
// HBM ridge.  It runs on the 5th-generation tensor cores:
//   * operands in bf16 (a bf16 copy of the corpus is kept when hx_index_config.storage == 1), fp32 accumulation in TMEM;
//   * CTA PAIRS (thread-block clusters of 2 = the two SMs of a TPC): one tcgen05.mma.cta_group::2 covers 256 queries x 256
//     corpus rows x 16 (K) — each CTA holds its own 128 queries (A, and the 128 TMEM lanes of their accumulators) and
//     stages only HALF of the corpus tile (B); the tensor cores of both SMs read both halves.  With cta_group::1 every SM
//     staged the whole 48 KB k-block itself: 94 B/clk of TMA writes + 94 B/clk of operand reads against 128 B/clk of shared
//     memory bandwidth — ncu: tensor pipe 65 % of active (profiles/r02_ncu_dense_v2_details.txt).  Pairing cuts both to
//     64 + 64 KB per 2 SMs;
//   * 5-stage TMA ring of (16 KB A | 16 KB half-B) per CTA (cp.async.bulk.tensor.cta_group::2, 128-byte swizzle, completion
//     counted on the LEADER CTA's mbarrier), MMAs issued by one elected thread of the leader, ring slots and accumulator
//     stages released in both CTAs by multicast tcgen05.commit; accumulators double-buffered in TMEM (2 x 256 columns);
//   * fused epilogue: 16 warps per CTA (one per TMEM lane quarter x column quarter) read the accumulator with tcgen05.ld,
//     turn it into the metric's score order with the stored row norms, and keep the HX_DENSE_T best rows per (query, run of
//     tiles, column quarter) by register insertion, emitting score_bits<<32|slot keys — the B x N score matrix is never
//     written.  The epilogue is what bounds the kernel at K = 768: a warp's column work is one dependent instruction
//     stream, and with 8 warps it took longer per tile than the tile's MMAs (1.60 ms, 0.78 of the burst peak); with 16
//     warps 1.45 ms, 0.86 (profiles/r02_dense_limiter_experiments.json).
// The approximate (bf16) keys only NOMINATE candidates: k_select keeps the k' best keys per query and the exact fp32
// scan kernel (bit-exact reference arithmetic) re-ranks them, so returned scores are exact and ordering follows the
// reference's (score, id) rule; recall vs the exhaustive exact scan is measured, not assumed.
//
// Reference behaviour this stands in for: the same answer VectorIndex::search_restricted gives with every id as a
// candidate (restricted_exact_scan, search/vector/restricted.rs:753-835), i.e. exact brute-force top-k.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hx_index.hpp"
#include "k_scan.cuh"
#include "k_util.cuh"

#define HXD_BM 128          // queries per CTA tile (UMMA M = 256 over the CTA pair)
#define HXD_BN 256          // corpus rows per tile (UMMA N); each CTA of the pair stages HXD_BN / 2 of them
#define HXD_BK 64           // bf16 elements per k-block = 128 bytes = one swizzle atom
#define HXD_STAGES 5
#define HXD_PARTS 4         // column quarters of a tile: one epilogue warp per (TMEM lane quarter, column quarter)
#ifndef HXD_T
#define HXD_T 8             // best rows kept per (query, run of n-tiles, column quarter): 32 per (query, run)
#endif
// Row placement inside a 256-row corpus tile.  A bucket of the fused top-T is (query, run of tiles, 64-COLUMN quarter) and keeps
// HXD_T rows.  With rows placed in id order a quarter is 64 CONSECUTIVE ids, so a query whose true neighbours are consecutive
// ids (near-duplicate rows inserted back to back: chunks of one document) could have more than HXD_T of them in one bucket and
// lose the rest.  HXD_INTERLEAVE = 1 stores row j of a tile at position (j % 4) * 64 + j / 4: consecutive ids land in different
// quarters (a run of 32 consecutive ids puts at most 8 into any bucket), at no cost to the kernel (same loads, same MMAs; only
// the slot a column stands for changes).
// Measured (profiles/r02_dense_row_placement_ab.json): on a fixture whose 11 nearest rows are consecutive ids the id-order
// placement returns 8 of the 10 true neighbours (recall 0.8), the interleaved one all of them; kernel time 1.489 -> 1.458 ms at
// the C4 shard shape (no cost).  Default on.
#ifndef HXD_INTERLEAVE
#define HXD_INTERLEAVE 1
#endif
// position of tile row j / tile row held at position c (inverse of each other)
__host__ __device__ __forceinline__ uint32_t hxd_pos_of_row(uint32_t j) {
#if HXD_INTERLEAVE
  return (j & 3u) * (HXD_BN / HXD_PARTS) + (j >> 2);
#else
  return j;
#endif
}
__host__ __device__ __forceinline__ uint32_t hxd_row_at_pos(uint32_t c) {
#if HXD_INTERLEAVE
  return (c & (HXD_BN / HXD_PARTS - 1u)) * HXD_PARTS + c / (HXD_BN / HXD_PARTS);
#else
  return c;
#endif
}
#define HXD_THREADS 640     // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warp 3: idle, warps 4-19: epilogue
#define HXD_EPI_THREADS 512
#define HXD_STAGE_CAP 12    // per-thread staging entries (shared memory) between the column test and the top-T insertion
#define HXD_A_BYTES (HXD_BM * HXD_BK * 2)
#define HXD_B_BYTES ((HXD_BN / 2) * HXD_BK * 2)   // this CTA's half of the corpus tile
#define HXD_STAGE_BYTES (HXD_A_BYTES + HXD_B_BYTES)

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hxd_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t hxd_mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void hxd_cluster_sync() {   // all threads of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// tile -> this CTA's shared memory; the bytes are counted on the mbarrier at cluster address `bar_cluster` (the leader's)
__device__ __forceinline__ void hxd_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          hx_smem_u32(smem_dst)),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void hxd_prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void hxd_mbar_arrive_cluster(uint32_t bar_cluster) {   // arrive on a barrier of any CTA of the cluster
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void hxd_tmem_alloc(uint32_t* smem_out, uint32_t cols) {   // the same warp id in both CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hx_smem_u32(smem_out)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void hxd_tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void hxd_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hxd_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem of both CTAs] (+)= A[smem of both CTAs] * B[smem halves of both CTAs]^T, M=256, N=256, K=16 (bf16), fp32 accumulate.
// Issued by one thread of the leader CTA; the descriptors are shared-memory offsets valid in both CTAs.
__device__ __forceinline__ void hxd_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the barrier at the same offset in BOTH CTAs when they have completed
__device__ __forceinline__ void hxd_umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   hx_smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = TMEM lane base+i)
__device__ __forceinline__ void hxd_tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: the caller overlaps it with work on the previous chunk, then hxd_tmem_wait_ld()
__device__ __forceinline__ void hxd_tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void hxd_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void hxd_tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile stored by TMA with the 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart.
// SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 64 |
// version [46,48) = 1 (Blackwell) | layout_type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t hxd_make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)64u << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;
  return d;
}

struct HxDenseArgs {
  uint32_t n_rows;            // corpus rows (unpadded)
  uint32_t n_queries;         // B (unpadded)
  uint32_t k_blocks;          // ldb / 64
  uint32_t m_pairs, n_tiles;  // m_pairs = pairs of 128-query tiles (one pair = one CTA pair)
  uint32_t n_split;           // work unit u of a CTA pair: query pair u % m_pairs, n-tiles [r*n_tiles/n_split, (r+1)*n_tiles/n_split), r = u / m_pairs
  const float* row_aux;       // cosine: 1/|x_i|   ; euclidean: |x_i|^2            (from the bf16-rounded rows)
  const float* q_aux;         // cosine: 1/|q_b|   ; euclidean: |q_b|^2
  uint64_t* keys;             // [B][n_split][HXD_PARTS][HXD_T]
  int32_t metric;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HXD_THREADS, 1)
k_dense_scores(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, HxDenseArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // HXD_STAGES x (A 16 KB | half-B 16 KB); the 128-byte swizzle needs 1024-byte aligned tiles: align inside the window
  // (the window starts at the same offset in both CTAs of the pair, so every object below has the same offset in both)
  unsigned char* smem = smem_raw + ((1024u - (hx_smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + HXD_STAGES * HXD_STAGE_BYTES);   // leader's copy is the live one
  uint64_t* empty = full + HXD_STAGES;      // per CTA: ring slot free (multicast commit of the MMAs that read it)
  uint64_t* tfull = empty + HXD_STAGES;     // [2] per CTA: accumulator stage ready for the epilogue (multicast commit)
  uint64_t* tempty = tfull + 2;             // [2] leader's copy: accumulator stage drained by the epilogues of BOTH CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* aux = reinterpret_cast<float*>(tmem_slot + 4);   // [2][HXD_BN] row aux of the tile being drained
  float* stg_t = aux + 2 * HXD_BN;                          // [HXD_STAGE_CAP][256] staged t values   (entry-major: conflict-free)
  uint32_t* stg_s = reinterpret_cast<uint32_t*>(stg_t + HXD_STAGE_CAP * HXD_EPI_THREADS);   // [HXD_STAGE_CAP][256] staged slots

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t rank = hxd_cta_rank();                     // 0 = leader (issues the MMAs of the pair)
  const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  if (threadIdx.x == 0) {
    hxd_prefetch_tmap(&map_q);
    hxd_prefetch_tmap(&map_x);
    for (int s = 0; s < HXD_STAGES; ++s) {
      hx_mbar_init(full + s, 1);                             // the leader's producer: one arrive.expect_tx for both CTAs' bytes
      hx_mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      hx_mbar_init(tfull + s, 1);
      hx_mbar_init(tempty + s, 2 * HXD_EPI_THREADS);
    }
    hx_fence_mbar_init();
  }
  if (warp == 2) hxd_tmem_alloc(tmem_slot, 512);
  hxd_fence_before();
  hxd_cluster_sync();       // barriers of both CTAs initialised before any remote arrive / complete_tx; TMEM allocated
  hxd_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t total_units = a.m_pairs * a.n_split;

  if (warp == 0) {
    // ===== TMA producer (one lane, in BOTH CTAs: own 128 queries + own half of the corpus tile) =====
    if (lane == 0) {
      uint32_t stage = 0, ph = 0;
      for (uint32_t u = pair; u < total_units; u += n_pairs) {
        // units with the same n-range and different query pairs run on neighbouring CTA pairs in step: the corpus tile is
        // read from HBM once and served to the others by L2
        const uint32_t mt = 2u * (u % a.m_pairs) + rank, r = u / a.m_pairs;
        const uint32_t nt0 = (uint32_t)((uint64_t)r * a.n_tiles / a.n_split), nt1 = (uint32_t)((uint64_t)(r + 1) * a.n_tiles / a.n_split);
        for (uint32_t nt = nt0; nt < nt1; ++nt)
          for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
            hx_mbar_wait(empty + stage, ph ^ 1u);           // own slot free (the multicast commit arrives in both CTAs)
            unsigned char* sa = tiles + (size_t)stage * HXD_STAGE_BYTES;
            const uint32_t lead_full = hxd_mapa(hx_smem_u32(full + stage), 0u);
            if (rank == 0) hx_mbar_expect_tx(full + stage, 2u * HXD_STAGE_BYTES);
            hxd_tma_load_2d(sa, &map_q, lead_full, (int)(kb * HXD_BK), (int)(mt * HXD_BM));
            hxd_tma_load_2d(sa + HXD_A_BYTES, &map_x, lead_full, (int)(kb * HXD_BK), (int)(nt * HXD_BN + rank * (HXD_BN / 2)));
            if (++stage == HXD_STAGES) { stage = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane of the LEADER CTA, for the pair) =====
    if (lane == 0 && rank == 0) {
      // InstrDescriptor: c_format F32 (1<<4) | a_format BF16 (1<<7) | b_format BF16 (1<<10) | K-major A,B | N>>3 at 17 | M>>4 at 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HXD_BN >> 3) << 17) | ((uint32_t)((2 * HXD_BM) >> 4) << 24);
      uint32_t stage = 0, ph = 0, acc = 0, acc_ph = 0;
      for (uint32_t u = pair; u < total_units; u += n_pairs) {
       const uint32_t r = u / a.m_pairs;
       const uint32_t nt0 = (uint32_t)((uint64_t)r * a.n_tiles / a.n_split), nt1 = (uint32_t)((uint64_t)(r + 1) * a.n_tiles / a.n_split);
       for (uint32_t nt = nt0; nt < nt1; ++nt) {
        hx_mbar_wait(tempty + acc, acc_ph ^ 1u);   // both epilogues have drained this accumulator stage
        hxd_fence_after();
        const uint32_t tmem_d = tmem_base + acc * HXD_BN;
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          hx_mbar_wait(full + stage, ph);          // both CTAs' tiles have landed
          hxd_fence_after();
          const uint32_t sa = hx_smem_u32(tiles + (size_t)stage * HXD_STAGE_BYTES);
          const uint64_t adesc = hxd_make_desc(sa), bdesc = hxd_make_desc(sa + HXD_A_BYTES);
#pragma unroll
          for (uint32_t k = 0; k < HXD_BK / 16; ++k)   // 16 bf16 = 32 bytes along K inside the swizzle atom: start address += 2
            hxd_umma(tmem_d, adesc + 2ull * k, bdesc + 2ull * k, idesc, (kb | k) ? 1u : 0u);
          hxd_umma_commit_pair(empty + stage);        // slot reusable in both CTAs once these MMAs have read it
          if (++stage == HXD_STAGES) { stage = 0; ph ^= 1u; }
        }
        hxd_umma_commit_pair(tfull + acc);            // accumulator complete (each CTA drains its own 128 lanes)
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
       }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps; thread = one query row x one quarter (64 columns) of the tile =====
    // A warp's column work is ONE dependent instruction stream (a lane is a query; whenever any of the 32 admits a column the
    // warp walks the staging path).  With 8 warps x 128 columns that stream took longer per tile than the tile's 12 k-blocks
    // of MMAs (profiles/r02_dense_limiter_experiments.json: the epilogue's column work cost 0.36 ms of a 1.6 ms kernel and
    // 4 extra instructions per 8 columns cost another 0.28 ms): 16 warps x 64 columns halve it.
    // ncu on the first version (profiles/): the tensor pipe sat at 10 % because 4 warps x 256 columns x ~30 instructions
    // per column could not drain an accumulator as fast as the MMAs filled it.  Now the test is done in the dot-product
    // domain (one FMUL/FFMA + one compare per column: the score is monotone in it), padded rows carry a sentinel that
    // can never pass, row terms are read as float4, and two warps share each TMEM lane quarter.
    const uint32_t ew = warp - 4;                     // 0..15
    const uint32_t q4 = ew & 3u;                      // TMEM lane quarter this warp may access (warp id % 4)
    const uint32_t half = ew >> 2;                    // column quarter 0..3
    const uint32_t row_in_tile = q4 * 32 + lane;
    const uint32_t et = threadIdx.x - 128;            // 0..511
    uint32_t acc = 0, acc_ph = 0;
    const uint32_t lead_tempty0 = hxd_mapa(hx_smem_u32(tempty), 0u);
    for (uint32_t u = pair; u < total_units; u += n_pairs) {
      const uint32_t mt = 2u * (u % a.m_pairs) + rank, r = u / a.m_pairs;
      const uint32_t nt0 = (uint32_t)((uint64_t)r * a.n_tiles / a.n_split), nt1 = (uint32_t)((uint64_t)(r + 1) * a.n_tiles / a.n_split);
      const uint32_t qrow = mt * HXD_BM + row_in_tile;
      const float qa = qrow < a.n_queries ? a.q_aux[qrow] : 0.f;
      // running best-T (largest t) of this query over the whole run of tiles; t = s*ax (cosine) or 2s - ax (euclidean)
      float bt[HXD_T];
      uint32_t bs[HXD_T];
#pragma unroll
      for (int i = 0; i < HXD_T; ++i) { bt[i] = -__int_as_float(0x7f800000); bs[i] = HX_ABSENT; }
      float thr = -__int_as_float(0x7f800000);
      uint32_t cnt = 0;
      for (uint32_t nt = nt0; nt < nt1; ++nt) {
        const uint32_t n0 = nt * HXD_BN;
        float* ax = aux + acc * HXD_BN;
        {
          const uint32_t c = et;                        // the first 256 threads stage the 256 row terms of the tile
          if (c < HXD_BN) {
            const bool live = n0 + hxd_row_at_pos(c) < a.n_rows;   // row_aux is indexed by POSITION, the bound is on the row there
            float v = live ? a.row_aux[n0 + c] : 0.f;
            if (!live) v = a.metric == HXM_COSINE ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);
            ax[c] = v;                                  // padded rows: t = NaN / -inf never beats the threshold
          }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        hx_mbar_wait(tfull + acc, acc_ph);
        hxd_fence_after();
        const uint32_t taddr = tmem_base + ((q4 * 32u) << 16) + acc * HXD_BN + half * (HXD_BN / HXD_PARTS);
        const float4* ax4 = reinterpret_cast<const float4*>(ax + half * (HXD_BN / HXD_PARTS));
        const uint32_t slot0 = n0 + half * (HXD_BN / HXD_PARTS);
        // The loop body is deliberately small and NOT unrolled: the first version inlined the insertion network at every
        // column (10k SASS instructions, instruction-cache bound: tensor pipe 9 %).  Columns that beat the running
        // threshold are only STAGED (two predicated shared-memory stores); the single insertion-network instance below
        // drains the staging area.
        // round 2: the accumulator is read 32 columns at a time and DOUBLE BUFFERED — the tcgen05.ld of chunk i+1 is in
        // flight while chunk i is tested — instead of sixteen serialised x8 loads per half tile: with two epilogue warps
        // per scheduler the TMEM round trip (not the instruction count) was what kept the tensor pipe at 54 % (ncu,
        // profiles/r02_ncu_dense_v1_raw.txt)
        auto drain = [&]() {   // the insertion network (one instance per call site: eight in all)
#pragma unroll 1
          for (uint32_t i = 0; i < cnt; ++i) {
            const float tvi = stg_t[i * HXD_EPI_THREADS + et];
            if (tvi > bt[HXD_T - 1]) {
              bt[HXD_T - 1] = tvi;
              bs[HXD_T - 1] = stg_s[i * HXD_EPI_THREADS + et];
#pragma unroll
              for (int j = HXD_T - 1; j > 0; --j)
                if (bt[j] > bt[j - 1]) {
                  const float tf = bt[j]; bt[j] = bt[j - 1]; bt[j - 1] = tf;
                  const uint32_t tsl = bs[j]; bs[j] = bs[j - 1]; bs[j - 1] = tsl;
                }
            }
          }
          cnt = 0;
          thr = bt[HXD_T - 1];
        };
        auto test8 = [&](const uint32_t* rr, uint32_t c0, bool last) {   // eight columns against the running threshold
          const float4 x0 = ax4[(c0 >> 2)], x1 = ax4[(c0 >> 2) + 1];
          float tv[8];
          if (a.metric == HXM_COSINE) {
            tv[0] = __uint_as_float(rr[0]) * x0.x; tv[1] = __uint_as_float(rr[1]) * x0.y;
            tv[2] = __uint_as_float(rr[2]) * x0.z; tv[3] = __uint_as_float(rr[3]) * x0.w;
            tv[4] = __uint_as_float(rr[4]) * x1.x; tv[5] = __uint_as_float(rr[5]) * x1.y;
            tv[6] = __uint_as_float(rr[6]) * x1.z; tv[7] = __uint_as_float(rr[7]) * x1.w;
          } else {
            tv[0] = fmaf(2.0f, __uint_as_float(rr[0]), -x0.x); tv[1] = fmaf(2.0f, __uint_as_float(rr[1]), -x0.y);
            tv[2] = fmaf(2.0f, __uint_as_float(rr[2]), -x0.z); tv[3] = fmaf(2.0f, __uint_as_float(rr[3]), -x0.w);
            tv[4] = fmaf(2.0f, __uint_as_float(rr[4]), -x1.x); tv[5] = fmaf(2.0f, __uint_as_float(rr[5]), -x1.y);
            tv[6] = fmaf(2.0f, __uint_as_float(rr[6]), -x1.z); tv[7] = fmaf(2.0f, __uint_as_float(rr[7]), -x1.w);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (tv[e] > thr) {
              stg_t[cnt * HXD_EPI_THREADS + et] = tv[e];
#if HXD_INTERLEAVE
              stg_s[cnt * HXD_EPI_THREADS + et] = n0 + (c0 + e) * HXD_PARTS + half;   // hxd_row_at_pos(half * 64 + c0 + e)
#else
              stg_s[cnt * HXD_EPI_THREADS + et] = slot0 + c0 + e;
#endif
              ++cnt;
            }
          if (cnt > HXD_STAGE_CAP - 8 || last) drain();
        };
        {   // 64 columns: two 32-column loads, the second in flight while the first is tested (16-column halves of `rb`
            // would save registers, but 640 threads leave 96 each and this fits)
          uint32_t ra[32], rb[32];
          hxd_tmem_ld32_nowait(taddr, ra);
          hxd_tmem_wait_ld();
          hxd_tmem_ld32_nowait(taddr + 32, rb);
          test8(ra, 0, false); test8(ra + 8, 8, false); test8(ra + 16, 16, false); test8(ra + 24, 24, false);
          hxd_tmem_wait_ld();
          test8(rb, 32, false); test8(rb + 8, 40, false); test8(rb + 16, 48, false); test8(rb + 24, 56, true);
        }
        hxd_fence_before();
        hxd_mbar_arrive_cluster(lead_tempty0 + acc * 8u);   // 2 x 256 arrivals (both CTAs) free the accumulator stage
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
      if (qrow < a.n_queries) {
        uint64_t* out = a.keys + (((size_t)qrow * a.n_split + r) * HXD_PARTS + half) * HXD_T;
#pragma unroll
        for (int i = 0; i < HXD_T; ++i) {
          uint64_t key = HX_KEY_MAX;
          if (bs[i] != HX_ABSENT) {
            float sc = a.metric == HXM_COSINE ? 0.5f - 0.5f * bt[i] * qa : qa - bt[i];   // (1-cos)/2 ; |q|^2 + |x|^2 - 2<q,x>
            sc = fmaxf(sc, 0.0f);
            key = hx_make_key(sc, bs[i]);
          }
          out[i] = key;
        }
      }
    }
  }
  hxd_fence_before();
  hxd_cluster_sync();       // the leader's MMAs wrote BOTH CTAs' TMEM: nobody frees before both are done
  if (warp == 2) hxd_tmem_dealloc(tmem_base, 512);
}

// f32 rows -> bf16 rows (round to nearest even) with zero padding to ldb, plus the per-row aux term computed from the
// ROUNDED values (so that approximate scores are self-consistent)
// `tiled`: the corpus copy — row w goes to its tile position (hxd_pos_of_row); queries keep their order
__global__ void k_to_bf16(const float* __restrict__ src, size_t rows, uint32_t dim, size_t ld_src, __nv_bfloat16* __restrict__ dst,
                          uint32_t ldb, float* __restrict__ aux, int metric, int tiled) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (w >= rows) return;
  const size_t pos = tiled ? (w / HXD_BN) * HXD_BN + hxd_pos_of_row((uint32_t)(w % HXD_BN)) : w;
  const float* s = src + w * ld_src;
  __nv_bfloat16* d = dst + pos * (size_t)ldb;
  float ss = 0.f;
  for (uint32_t j = lane; j < ldb; j += 32) {
    const float x = j < dim ? s[j] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(x);
    d[j] = b;
    const float xb = __bfloat162float(b);
    ss += xb * xb;
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 16); ss += __shfl_xor_sync(0xffffffffu, ss, 8);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  if (lane == 0) aux[pos] = metric == HXM_COSINE ? (ss > 0.f ? rsqrtf(ss) : 0.f) : ss;
}

// keys[q][j] hold GLOBAL slots in their low word; turn the selected ones into a per-query candidate slot list
__global__ void k_keys_to_slots(const uint64_t* __restrict__ sel_ids_as_slots, size_t total, uint32_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = (uint32_t)sel_ids_as_slots[i];
}

// ---- host ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static hx_status make_map(CUtensorMap* map, void* base, uint64_t rows, uint32_t ldb, uint32_t box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
      hx_set_error("cuTensorMapEncodeTiled is not available from the driver");
      return HX_ERR_CUDA;
    }
    fn = (PFN_encodeTiled)p;
  }
  const cuuint64_t dims[2] = {ldb, rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ldb * 2};
  const cuuint32_t box[2] = {HXD_BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    hx_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return HX_ERR_CUDA;
  }
  return HX_OK;
}

// lazily build the bf16 copy of the corpus
static hx_status ensure_bf16(hx_index* ix, uint32_t ldb) {
  if (ix->d_vec_bf16) return HX_OK;
  const size_t n_pad = (ix->n + HXD_BN - 1) / HXD_BN * HXD_BN;
  HX_CUDA(cudaMalloc(&ix->d_vec_bf16, n_pad * (size_t)ldb * 2));
  HX_CUDA(cudaMemset(ix->d_vec_bf16, 0, n_pad * (size_t)ldb * 2));
  HX_CUDA(cudaMalloc((void**)&ix->d_sqnorm, n_pad * sizeof(float)));
  HX_CUDA(cudaMemset(ix->d_sqnorm, 0, n_pad * sizeof(float)));
  k_to_bf16<<<(unsigned)((ix->n + 7) / 8), 256>>>(ix->d_vec, ix->n, ix->cfg.dimension, ix->ld, (__nv_bfloat16*)ix->d_vec_bf16, ldb,
                                                  ix->d_sqnorm, ix->cfg.metric, 1);
  HX_CUDA(cudaGetLastError());
  return HX_OK;
}

// iota of the nominee CSR offsets (b * k')
static __global__ void k_dense_offsets(uint64_t* __restrict__ offs, size_t B, uint32_t kprime) {
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= B) offs[b] = b * (uint64_t)kprime;
}

static hx_status dense_check(hx_index* ix, const hx_search_params* p) {
  if (!p || p->k == 0) {
    hx_set_error("result count must be non-zero");
    return HX_ERR_INVALID_PARAMETER;
  }
  if (p->k > 800) {
    hx_set_error("dense search result count must be at most 800");
    return HX_ERR_QUERY;
  }
  if (ix->cfg.storage != 1) {
    hx_set_error("hx_search_dense needs hx_index_config.storage == 1 (bf16 copy of the corpus)");
    return HX_ERR_INVALID_PARAMETER;
  }
  if (ix->cfg.metric == HX_METRIC_MANHATTAN) {
    hx_set_error("the L1 metric is not a contraction: use hx_search_restricted");
    return HX_ERR_UNSUPPORTED;
  }
  if (p->query_dimension != 0 && p->query_dimension != ix->cfg.dimension) {
    hx_set_error("invalid dimension: expected %u, got %u", ix->cfg.dimension, p->query_dimension);
    return HX_ERR_INVALID_DIMENSION;
  }
  return HX_OK;
}

// Everything below is enqueued on `stream`; nothing synchronises with the host (buffers, tensor maps and the nominee
// offsets are cached in the scratch set across calls).  d_queries: B x dim f32 on the device.  Per-query validation status
// is left in s->d_qstatus (a query that fails validation yields count 0); error flags accumulate in s->d_err[0].
hx_status hx_dense_device(hx_index* ix, HxScratch* scr, const float* d_q, size_t B, const hx_search_params* p,
                          uint64_t* d_out_ids, float* d_out_sc, uint32_t* d_out_cnt, cudaStream_t stream,
                          cudaEvent_t e0, cudaEvent_t e1, uint32_t* launches_out) {
  hx_status rc = dense_check(ix, p);
  if (rc) return rc;
  if (B == 0) return HX_OK;
  if (ix->n == 0 || !ix->populated) {
    HX_CUDA(cudaMemsetAsync(d_out_cnt, 0, B * sizeof(uint32_t), stream));
    return HX_OK;
  }
  const uint32_t dim = ix->cfg.dimension, k = p->k;
  const uint32_t ldb = (dim + HXD_BK - 1) / HXD_BK * HXD_BK;
  if ((rc = ensure_bf16(ix, ldb))) return rc;
  const size_t n = ix->n;
  const uint32_t m_pairs = (uint32_t)((B + 2 * HXD_BM - 1) / (2 * HXD_BM)), n_tiles = (uint32_t)((n + HXD_BN - 1) / HXD_BN);
  const size_t B_pad = (size_t)m_pairs * 2 * HXD_BM, n_pad = (size_t)n_tiles * HXD_BN;
  const size_t n_cta_pairs = (size_t)std::max(1, ix->sm_count / 2);   // one persistent CTA pair per TPC
  const uint32_t kprime = (uint32_t)std::min<size_t>(std::max<uint32_t>(4 * k, 64), std::min<size_t>(800, n));
  // runs per query pair: enough units to fill the CTA pairs, and enough that one run's best-T comfortably covers its share
  // of the k' nominees even when the true neighbours cluster in id space (8x head-room)
  uint32_t n_split = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_tiles,
      std::max<size_t>(std::max<size_t>(4, n_cta_pairs / m_pairs), (8 * (size_t)kprime + HXD_PARTS * HXD_T - 1) / (HXD_PARTS * HXD_T))));
  // the grid is one persistent CTA pair per TPC: make the number of work units a multiple of it so that the last wave is
  // full (1024 queries = 4 query pairs: 32 runs = 128 units = 1.73 waves of 74 pairs, the tensor pipe idles 13 % of the
  // launch; 37 runs = 148 units = exactly 2 waves)
  if ((size_t)m_pairs * n_split > n_cta_pairs)
    for (uint32_t cand = n_split; cand <= std::min<uint32_t>(n_tiles, 2 * n_split); ++cand)
      if (((size_t)m_pairs * cand) % n_cta_pairs == 0) { n_split = cand; break; }
  const size_t nkeys = B * (size_t)n_split * HXD_PARTS * HXD_T;
  int mi = 0;
  bool grew = false;
  auto take = [&](size_t bytes, void** out) -> hx_status {
    const unsigned char* before = scr->misc[mi].p;
    hx_status r = scr->misc[mi].reserve(bytes + 256);
    if (scr->misc[mi].p != before) grew = true;
    *out = scr->misc[mi].p;
    ++mi;
    return r;
  };
  float *d_qaux, *d_sel_sc;
  __nv_bfloat16* d_qb;
  uint64_t *d_keys, *d_sel_ids, *d_keys2, *d_coffs;
  uint32_t *d_sel_cnt, *d_cslots;
  if ((rc = take(B_pad * (size_t)ldb * 2, (void**)&d_qb)) || (rc = take(B_pad * 4, (void**)&d_qaux)) ||
      (rc = take(nkeys * 8, (void**)&d_keys)) || (rc = take(B * (size_t)kprime * 8, (void**)&d_sel_ids)) ||
      (rc = take(B * (size_t)kprime * 4, (void**)&d_sel_sc)) || (rc = take(B * 4, (void**)&d_sel_cnt)) ||
      (rc = take(B * (size_t)kprime * 4, (void**)&d_cslots)) || (rc = take((B + 1) * 8, (void**)&d_coffs)) ||
      (rc = take(B * (size_t)kprime * 8, (void**)&d_keys2)))
    return rc;
  if ((rc = scr->d_qhdr.reserve(B)) || (rc = scr->d_qstatus.reserve(B)) || (rc = scr->d_err.reserve(4))) return rc;
  float* d_qhdr = scr->d_qhdr.p;
  uint32_t* d_status = scr->d_qstatus.p;
  uint32_t* d_err = scr->d_err.p;
  HxDenseCache& dc = scr->dense;
  const bool shape_changed = grew || dc.B != B || dc.kprime != kprime || dc.ldb != ldb;
  HX_CUDA(cudaMemsetAsync(d_err, 0, 4, stream));
  HX_CUDA(cudaMemsetAsync(d_sel_ids, 0xFF, B * (size_t)kprime * 8, stream));   // unfilled nominee slots read as HX_ABSENT
  if (shape_changed) {   // pad rows / pad columns of the bf16 query tile and the nominee offsets: once per shape
    HX_CUDA(cudaMemsetAsync(d_qb, 0, B_pad * (size_t)ldb * 2, stream));
    HX_CUDA(cudaMemsetAsync(d_qaux, 0, B_pad * 4, stream));
    k_dense_offsets<<<(unsigned)((B + 256) / 256), 256, 0, stream>>>(d_coffs, B, kprime);
    if ((rc = make_map(reinterpret_cast<CUtensorMap*>(dc.map_q), d_qb, B_pad, ldb, HXD_BM))) return rc;
    dc.B = B;
    dc.kprime = kprime;
    dc.ldb = ldb;
  }
  if (dc.map_x_base != ix->d_vec_bf16 || dc.map_x_rows != n_pad) {
    if ((rc = make_map(reinterpret_cast<CUtensorMap*>(dc.map_x), ix->d_vec_bf16, n_pad, ldb, HXD_BN / 2))) return rc;   // a CTA stages half a corpus tile
    dc.map_x_base = ix->d_vec_bf16;
    dc.map_x_rows = n_pad;
  }
  uint32_t launches = 0;
  // validation + exact headers (same order of checks as every other entry point)
  float limit = 0.f;
  bool has_limit = false;
  if (ix->cfg.metric == HX_METRIC_EUCLIDEAN) {
    const double exact = std::sqrt((double)FLT_MAX / ((double)dim * 8.0));
    limit = (float)exact;
    if ((double)limit > exact) limit = std::nextafter(limit, 0.0f);
    has_limit = true;
  }
  k_validate_and_header<<<(unsigned)((B + 7) / 8), 256, 0, stream>>>(d_q, B, dim, dim, ix->cfg.metric, limit, has_limit ? 1 : 0,
                                                                    d_qhdr, d_status);
  k_to_bf16<<<(unsigned)((B + 7) / 8), 256, 0, stream>>>(d_q, B, dim, dim, d_qb, ldb, d_qaux, ix->cfg.metric, 0);
  launches += 2;
  HX_CUDA(cudaGetLastError());
  // ---- tensor-core pass ----
  HxDenseArgs a{};
  a.n_rows = (uint32_t)n;
  a.n_queries = (uint32_t)B;
  a.k_blocks = ldb / HXD_BK;
  a.m_pairs = m_pairs;
  a.n_tiles = n_tiles;
  a.n_split = n_split;
  a.row_aux = ix->d_sqnorm;
  a.q_aux = d_qaux;
  a.keys = d_keys;
  a.metric = ix->cfg.metric;
  const size_t smem = (size_t)HXD_STAGES * HXD_STAGE_BYTES + 16 * 8 + 16 + 2 * HXD_BN * 4 + (size_t)HXD_STAGE_CAP * HXD_EPI_THREADS * 8 + 1024;
  static std::atomic<int> attr_set[64];
  if (!attr_set[ix->device & 63].load()) {
    HX_CUDA(cudaFuncSetAttribute(k_dense_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[ix->device & 63].store(1);
  }
  const uint32_t grid = 2u * (uint32_t)std::min<size_t>((size_t)m_pairs * n_split, n_cta_pairs);   // clusters of 2 (__cluster_dims__)
  if (e0) HX_CUDA(cudaEventRecord(e0, stream));
  k_dense_scores<<<grid, HXD_THREADS, smem, stream>>>(*reinterpret_cast<CUtensorMap*>(dc.map_q),
                                                      *reinterpret_cast<CUtensorMap*>(dc.map_x), a);
  if (e1) HX_CUDA(cudaEventRecord(e1, stream));
  launches++;
  HX_CUDA(cudaGetLastError());
  // ---- k' nominees per query by approximate key, then the exact re-rank with the reference arithmetic ----
  HxDev dev = ix->dev();
  HxSelectArgs s1{};
  s1.keys = d_keys;
  s1.q_status = d_status;
  s1.B = (uint32_t)B;
  s1.k = kprime;
  s1.shared_set = 2;   // keys carry global slots in their low word (no candidate indirection)
  s1.n_shared = (uint64_t)n_split * HXD_PARTS * HXD_T;
  s1.out_ids = d_sel_ids;
  s1.out_scores = d_sel_sc;
  s1.out_counts = d_sel_cnt;
  k_select<<<(unsigned)std::min<size_t>(B, 65535), HX_SEL_THREADS, 0, stream>>>(dev, s1);
  k_keys_to_slots<<<(unsigned)((B * (size_t)kprime + 255) / 256), 256, 0, stream>>>(d_sel_ids, B * (size_t)kprime, d_cslots);
  launches += 2;
  HxScanArgs sa{};
  sa.queries = d_q;
  sa.q_hdr = d_qhdr;
  sa.q_status = d_status;
  sa.B = (uint32_t)B;
  sa.cand_slots = d_cslots;
  sa.cand_offsets = d_coffs;
  sa.keys = d_keys2;
  sa.shared_set = 0;
  sa.n_shared = 0;
  sa.chunk = 32 * ((kprime + 31) / 32);
  sa.err_flags = d_err;
  {
    dim3 g(1, (unsigned)std::min<size_t>(B, 65535));
    const uint32_t sm = ix->ld * 4u;
    if (ix->cfg.metric == HX_METRIC_COSINE) k_scan<HXM_COSINE><<<g, HX_SCAN_THREADS, sm, stream>>>(dev, sa);
    else k_scan<HXM_EUCLIDEAN><<<g, HX_SCAN_THREADS, sm, stream>>>(dev, sa);
  }
  HxSelectArgs s2{};
  s2.keys = d_keys2;
  s2.cand_slots = d_cslots;
  s2.cand_offsets = d_coffs;
  s2.q_status = d_status;
  s2.B = (uint32_t)B;
  s2.k = k;
  s2.shared_set = 0;
  s2.n_shared = 0;
  s2.out_ids = d_out_ids;
  s2.out_scores = d_out_sc;
  s2.out_counts = d_out_cnt;
  k_select<<<(unsigned)std::min<size_t>(B, 65535), HX_SEL_THREADS, 0, stream>>>(dev, s2);
  launches += 2;
  HX_CUDA(cudaGetLastError());
  if (launches_out) *launches_out = launches;
  return HX_OK;
}

// Host-buffer entry point: a private stream + scratch set per call, the queries cross PCIe with one async copy, results
// and per-query status come back with four, ONE stream synchronisation.
hx_status hx_dense_impl(hx_index* ix, const float* queries, size_t B, const hx_search_params* p, uint64_t* out_ids,
                        float* out_scores, uint32_t* out_counts, hx_stats* stats) {
  hx_status rc = dense_check(ix, p);
  if (rc) return rc;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (B == 0) return HX_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return HX_ERR_INVALID_PARAMETER;
  if (ix->n == 0 || !ix->populated) {
    for (size_t b = 0; b < B; ++b) out_counts[b] = 0;
    return HX_OK;
  }
  const uint32_t dim = ix->cfg.dimension, k = p->k;
  HxScratch* scr = nullptr;
  if ((rc = hx_acquire_scratch(ix, &scr))) return rc;
  struct Rel { hx_index* ix; HxScratch* s; ~Rel() { hx_release_scratch(ix, s); } } rel{ix, scr};
  if ((rc = scr->d_queries.reserve(B * (size_t)dim)) || (rc = scr->d_out_ids.reserve(B * (size_t)k)) ||
      (rc = scr->d_out_scores.reserve(B * (size_t)k)) || (rc = scr->d_out_counts.reserve(B)) ||
      (rc = scr->h_status.reserve(B)) || (rc = scr->h_err.reserve(1)))
    return rc;
  cudaStream_t st = scr->stream;
  HX_CUDA(cudaMemcpyAsync(scr->d_queries.p, queries, B * (size_t)dim * 4, cudaMemcpyHostToDevice, st));
  uint32_t launches = 0;
  if ((rc = hx_dense_device(ix, scr, scr->d_queries.p, B, p, scr->d_out_ids.p, scr->d_out_scores.p, scr->d_out_counts.p, st,
                            scr->ev0, scr->ev1, &launches)))
    return rc;
  HX_CUDA(cudaMemcpyAsync(out_ids, scr->d_out_ids.p, B * (size_t)k * 8, cudaMemcpyDeviceToHost, st));
  HX_CUDA(cudaMemcpyAsync(out_scores, scr->d_out_scores.p, B * (size_t)k * 4, cudaMemcpyDeviceToHost, st));
  HX_CUDA(cudaMemcpyAsync(out_counts, scr->d_out_counts.p, B * 4, cudaMemcpyDeviceToHost, st));
  HX_CUDA(cudaMemcpyAsync(scr->h_status.p, scr->d_qstatus.p, B * 4, cudaMemcpyDeviceToHost, st));
  HX_CUDA(cudaMemcpyAsync(scr->h_err.p, scr->d_err.p, 4, cudaMemcpyDeviceToHost, st));
  HX_CUDA(cudaStreamSynchronize(st));
  for (size_t b = 0; b < B; ++b)
    if (scr->h_status.p[b] != HX_ST_OK) {
      const uint32_t w = scr->h_status.p[b], code = w >> 24;
      hx_set_error_index(w & 0xffffffu);
      hx_set_error("query %zu failed validation (code %u)", b, code);
      return code == HX_ST_COMPONENT ? HX_ERR_INVALID_VECTOR_COMPONENT
                                     : (code == HX_ST_ZERO_NORM ? HX_ERR_ZERO_NORM_COSINE : HX_ERR_MAGNITUDE_EXCEEDED);
    }
  if (scr->h_err.p[0] & HXF_INVALID_SCORE) {
    hx_set_error("vector distance kernel emitted an invalid score");
    return HX_ERR_INVARIANT_VIOLATION;
  }
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, scr->ev0, scr->ev1) == cudaSuccess) {
    ix->last_kernel_ms = ms;
    ix->last_kernel_launches = 1;
  }
  if (stats) {
    const uint32_t ldb = (dim + HXD_BK - 1) / HXD_BK * HXD_BK;
    stats->kernel_launches = launches;
    stats->distance_computations = (uint64_t)B * ix->n;
    stats->algorithmic_bytes = (uint64_t)ix->n * ldb * 2;   // corpus streamed once; the bound is the tensor pipe: 2*B*N*d flop
    stats->reserved = (uint64_t)(2.0 * (double)B * (double)ix->n * (double)ldb);   // flop of the contraction
  }
  return HX_OK;
}

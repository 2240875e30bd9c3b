// Non-causal flash attention for sm_100a:  out = softmax(q k^T * scale) v,  head_dim = 128, bf16 in/out.
//
// One CTA owns 256 query rows of one head (two 128-row tiles, "ping-pong") and streams keys/values:
//   warp 0        TMA producer: Q once, then K_j / V_j tiles (128 keys x 128 dims, two SWIZZLE_128B boxes each)
//                 into a ring of kSlots 32-KiB shared-memory slots
//   warp 1        MMA issuer (one thread):  S_w = Q_w K_j^T  (SS, 128x128x16 x 8)  -> TMEM
//                                           O_w += P_w V_j   (TS: P read from TMEM, V MN-major from smem)
//   warp 2        TMEM allocator (512 columns: S0 | S1 | O0 | O1, P_w aliases the first 64 columns of S_w)
//   warps 4..7    softmax for tile 0, warps 8..11 softmax for tile 1: one thread per query row
//                 (tcgen05.ld 32x32b: no cross-thread reductions), online softmax in the log2 domain with
//                 lazy rescaling of O (only when the running max grows by more than 2^8), P written back to
//                 TMEM as packed bf16, final O / l epilogue straight from TMEM to global memory.
// While the softmax warps of one tile work, the tensor core runs the other tile's MMAs.
//
// Grid shaping: work items are (head, 256-row pair).  Items that fill whole waves of SMs run over the full key
// range; the items of the last, partial wave are split along the keys into `split` pieces so that the tail costs
// 1/split of a wave; pieces emit un-normalised (O, max, sum) partials that attn_combine_kernel merges.
// A pair whose second tile lies entirely past the last query row runs in single-tile mode.
//
// Replaces flash_attention()/attention() (inferix/models/attention/flash_attention.py:42-200) at the two call
// sites of the block: self-attention over cache[0:local_end] (causal_model.py:307-315) and text
// cross-attention (wan_base/model.py:94-95).  Numerics follow FlashAttention-2: fp32 scores and running
// sum (of the un-rounded probabilities), bf16 P for the PV product, one division by l at the end.
#include <algorithm>
#include <vector>

#include "ifx_internal.h"
#include "ifx_ptx.cuh"

namespace ifx {

constexpr int kQT = 128;         // query rows per tile
constexpr int kKT = 128;         // keys per tile
constexpr int kHD = 128;         // head dim
#ifndef IFX_ATTN_SLOTS
#define IFX_ATTN_SLOTS 4
#endif
constexpr int kSlots = IFX_ATTN_SLOTS;   // K/V ring slots (32 KiB each; 5 still fits 227 KiB with the two Q tiles)
constexpr int kTileBytes = 128 * 128 * 2;  // 32 KiB
constexpr int kHalfBytes = kTileBytes / 2; // one 64-wide SWIZZLE_128B box
constexpr int kAttnThreads = 384;
constexpr int kAttnSmem = 2 * kTileBytes + kSlots * kTileBytes + 1024 + 256;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int kMaxSplit = 8;
constexpr int kMaxExt = IFX_ATTN_MAX_EXTENTS;  // key-row extents (runs of physically consecutive cache pages) per launch
#ifndef IFX_ATTN_POLY_EVERY
#define IFX_ATTN_POLY_EVERY 0
#endif
#ifndef IFX_ATTN_PCHUNKS
#define IFX_ATTN_PCHUNKS 2
#endif
// P = exp(S - max) is handed to the PV MMA in kPChunks column chunks, each behind its own mbarrier: the MMA warp issues
// the chunk's tcgen05.mma (2 of the 8 k-steps per 32-key chunk) as soon as that chunk is in TMEM, while the softmax
// warps are still computing the exponentials of the next chunk.  With one chunk (the round-1 behaviour) the tensor
// pipe sits behind the whole 128 x 128 exponentials (1024 MUFU clocks) of a tile before its 512-clock PV can start.
constexpr int kPChunks = IFX_ATTN_PCHUNKS;
static_assert(kPChunks == 1 || kPChunks == 2 || kPChunks == 4, "P chunks: 1, 2 or 4");
#ifndef IFX_ATTN_LDSPLIT
#define IFX_ATTN_LDSPLIT 0
#endif
constexpr bool kLdSplit = IFX_ATTN_LDSPLIT != 0;   // softmax: overlap the TMEM load of S's second half with the first half's max
constexpr int kPolyEvery = IFX_ATTN_POLY_EVERY;  // 0: all exponentials on MUFU; n: one pair in n on the FMA pipe

struct AttnParams {
    int32_t q_rows;
    int32_t kv_rows;
    int32_t heads;
    int32_t kv_group;    // query heads per key/value head (grouped-query attention; 1 = MHA)
    int32_t num_q_pairs;
    int32_t n_whole;     // items [0, n_whole) cover all keys; the rest are split into `split` pieces
    int32_t split;
    float scale_log2;
    __nv_bfloat16* out;
    int64_t ldo;
    float* part_o;       // [pieces][256][128] un-normalised O
    float* part_ml;      // [pieces][256][2]   (max * scale_log2, sum)
    unsigned int* combine_ctr;  // non-null: the last piece of an item to finish merges the item's partials in-kernel
                                // ([split items][2] arrival counters, zero between launches) — no attn_combine_kernel
    float* lse;          // optional [heads][q_rows] natural-log softmax denominators (log-sum-exp of the scaled scores)
    // ---- partial mode (ifx_attention_partial): every CTA emits a partial; keys are a list of row extents
    int32_t partial;         // 1: blockIdx = item * piece_count + sub, slot = item * pieces_per_item + piece_first + sub
    int32_t pieces_per_item;
    int32_t piece_first;
    int32_t piece_count;
    int32_t n_ext;               // 0: one dense extent [0, kv_rows)
    int32_t ext_row0[kMaxExt];   // first key row of each extent
    int32_t ext_rows[kMaxExt];   // keys in each extent
    int32_t ext_tile0[kMaxExt + 1];  // cumulative 128-key tile counts
    // ---- ordered key tiles + in-kernel wait (sequence parallel over peer memory): tiles [0, n_old_tiles) are
    // resident when the kernel starts; tiles [n_old_tiles, total) are rows the peers are still storing.  Every CTA
    // walks its share of the old tiles first, then spins (producer thread only) until all `wait_world` epoch flags
    // reached `wait_epoch`, then walks its share of the new tiles.  wait_flags == nullptr: no wait.
    int32_t n_old_tiles;
    int32_t wait_world;
    const long long* wait_flags;
    long long wait_epoch;
    unsigned long long wait_timeout_ns;
    int32_t no_dep_wait;     // 1: do not griddepcontrol.wait (see the kernel prologue)
    // ---- fused exchange: warp 2 of CTAs [0, push.n_ctas) ships this rank's rows of the fresh pages to the peers
    PeerPushParams push;
};

// Monotone cursor over the extent list: key tile j (over the concatenated extents) -> first key row and number of
// valid keys in the tile.  Tiles are visited in increasing j by every role, so the extent index only moves forward.
struct TileCursor {
    int e = 0;
    __device__ __forceinline__ void locate(const AttnParams& p, int j, int& row0, int& valid) {
        if (p.n_ext == 0) {
            row0 = j * kKT;
            valid = p.kv_rows - row0;
            return;
        }
        while (e + 1 < p.n_ext && j >= p.ext_tile0[e + 1]) ++e;
        const int lt = j - p.ext_tile0[e];
        row0 = p.ext_row0[e] + lt * kKT;
        valid = p.ext_rows[e] - lt * kKT;
    }
};

// The key tiles one CTA owns: slice `sub` of `n` of the old tiles, then the same slice of the new tiles.
struct TileRange {
    int old_begin, n_old, new_begin, n_new;
    __device__ __forceinline__ int count() const { return n_old + n_new; }
    __device__ __forceinline__ int tile(int i) const { return i < n_old ? old_begin + i : new_begin + (i - n_old); }
};
__device__ __forceinline__ TileRange make_tile_range(int n_all, int n_old_tiles, int sub, int n) {
    const int n_new_tiles = n_all - n_old_tiles;
    TileRange r;
    r.old_begin = static_cast<int>(static_cast<int64_t>(n_old_tiles) * sub / n);
    r.n_old = static_cast<int>(static_cast<int64_t>(n_old_tiles) * (sub + 1) / n) - r.old_begin;
    r.new_begin = n_old_tiles + static_cast<int>(static_cast<int64_t>(n_new_tiles) * sub / n);
    r.n_new = n_old_tiles + static_cast<int>(static_cast<int64_t>(n_new_tiles) * (sub + 1) / n) - r.new_begin;
    return r;
}

// One warp's share of the exchange (sequence parallel over peer memory): CTA `cta` of `p.n_ctas` copies a contiguous
// slice of this rank's new K / V rows (16-byte vectors, 4 loads in flight per lane) from the local cache to the same
// rows of every other rank's cache over NVLink, then arrives; the last CTA publishes the epoch on every rank.  The
// NVLink stores need no SM resource the attention uses (the warp is otherwise idle after the TMEM allocation), and the
// whole box-wide exchange is done long before any rank reaches its first fresh-page tile.
__device__ __forceinline__ void push_slice(const PeerPushParams& p, int cta, int lane) {
    // A rank's rows of one frame are contiguous in the cache (pages are frames: page_tokens == world * chunk, checked
    // on the host), so the exchange is 2 * frames contiguous segments (K then V) of seg_vecs 16-byte vectors each, at
    // the same offsets on every rank.  This CTA owns vectors [begin, end) of the concatenation; the loops below are
    // pointer increments only (no division per element: this warp shares an issue port with two softmax warps).
    const int64_t seg_vecs = static_cast<int64_t>(p.chunk) * (p.C >> 3);
    const int nseg = 2 * p.frames;
    const int64_t total = seg_vecs * nseg;
    const int64_t begin = total * cta / p.n_ctas, end = total * (cta + 1) / p.n_ctas;
    for (int sg = static_cast<int>(begin / seg_vecs); sg < nseg && sg * seg_vecs < end; ++sg) {
        const int which = sg / p.frames, f = sg % p.frames;
        const int64_t tb = static_cast<int64_t>(f) * p.page_tokens + static_cast<int64_t>(p.rank) * p.chunk;
        const int64_t crow = static_cast<int64_t>(p.pl.pages[f]) * p.page_tokens + tb % p.page_tokens;
        const int64_t lo = begin > sg * seg_vecs ? begin - sg * seg_vecs : 0;
        const int64_t hi = end < (sg + 1) * seg_vecs ? end - sg * seg_vecs : seg_vecs;
        const int64_t base = crow * p.C;                                  // element offset of the segment in every cache
        const uint4* src = reinterpret_cast<const uint4*>((which ? p.peer_v[p.rank] : p.peer_k[p.rank]) + base);
        constexpr int kUnroll = 4;
        for (int64_t u = lo + lane; u < hi; u += 32 * kUnroll) {
            uint4 val[kUnroll];
#pragma unroll
            for (int k = 0; k < kUnroll; ++k)
                if (u + k * 32 < hi) val[k] = src[u + k * 32];
            for (int d = 1; d < p.world; ++d) {
                const int dst = (p.rank + d) % p.world;                   // start at the neighbour: spread the links
                uint4* out = reinterpret_cast<uint4*>((which ? p.peer_v[dst] : p.peer_k[dst]) + base);
#pragma unroll
                for (int k = 0; k < kUnroll; ++k)
                    if (u + k * 32 < hi) out[u + k * 32] = val[k];
            }
        }
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        const unsigned int prev = atomicAdd(p.done_counter, 1u);
        if (prev == static_cast<unsigned int>(p.n_ctas) - 1) {
            *p.done_counter = 0;                                         // next launch on this stream starts from zero
            __threadfence_system();
            for (int d = 0; d < p.world; ++d) st_relaxed_sys(p.peer_flags[d] + p.rank, p.epoch);
        }
    }
}

// Producer-side wait for the peers' K / V rows: acquire every rank's epoch flag at system scope, then order the TMA
// (async proxy) reads that follow after the acquired generic-proxy view.  Bounded: traps instead of hanging the GPU.
__device__ __forceinline__ void wait_peer_flags(const AttnParams& p) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int s = 0; s < p.wait_world; ++s) {
        while (ld_acquire_sys(p.wait_flags + s) < p.wait_epoch) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (now - t0 > p.wait_timeout_ns) {
                printf("ifx attention: rank %d did not publish epoch %lld (have %lld)\n", s, p.wait_epoch,
                       ld_acquire_sys(p.wait_flags + s));
                __trap();
            }
            __nanosleep(200);
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                    // [2][32 KiB]
    uint8_t* sKV = smem + 2 * kTileBytes;  // [kSlots][32 KiB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + kSlots) * kTileBytes);
    uint64_t* kv_full = bars;              // [kSlots]
    uint64_t* kv_empty = bars + kSlots;    // [kSlots]
    uint64_t* q_full = bars + 2 * kSlots;  // [1]
    uint64_t* s_full = q_full + 1;         // [2]
    uint64_t* p_full = s_full + 2;         // [2][kPChunks]
    uint64_t* o_full = p_full + 2 * kPChunks;  // [1]
    uint64_t* v_fixed = o_full + 1;        // [kSlots]  warp 3 -> MMA (extent mode): V tile checked, stale rows zeroed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_fixed + kSlots);
    // valid keys of tile i at [i & 7], written by the producer before it arms the tile's K slot (release through the
    // mbarrier chain kv_full -> s_full); the softmax warps read it after s_full.  The producer runs < 4 tiles ahead.
    volatile int32_t* tile_valid = reinterpret_cast<volatile int32_t*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- which (item, key tiles) does this CTA own
    const int n_kv_all = p.n_ext ? p.ext_tile0[p.n_ext] : (p.kv_rows + kKT - 1) / kKT;
    const int n_old_all = p.wait_flags ? p.n_old_tiles : n_kv_all;
    int item, piece = -1, sub = 0, nsub = 1;
    if (p.partial) {
        item = blockIdx.x / p.piece_count;
        sub = blockIdx.x % p.piece_count;
        nsub = p.piece_count;
        piece = item * p.pieces_per_item + p.piece_first + sub;
    } else if (static_cast<int>(blockIdx.x) < p.n_whole) {
        item = blockIdx.x;
    } else {
        const int idx = blockIdx.x - p.n_whole;
        item = p.n_whole + idx / p.split;
        sub = idx % p.split;
        nsub = p.split;
        piece = idx;
    }
    const TileRange tiles = make_tile_range(n_kv_all, n_old_all, sub, nsub);
    const int head = item / p.num_q_pairs;
    const int q0 = (item % p.num_q_pairs) * (2 * kQT);
    const int n_kv = tiles.count();
    const bool two = q0 + kQT < p.q_rows;  // second tile has at least one real row
    // extent mode: a tile at the end of an extent is followed in memory by rows of OTHER pages (unmapped, or being
    // written by a peer).  Their scores are masked, but 0 x NaN would still poison P V, so warp 3 zeroes those V rows
    // in shared memory before the MMA warp may consume the tile (v_fixed barrier).  The dense mode needs none of
    // this: its tensor map ends at kv_rows and TMA zero-fills.
    const bool fix_tails = p.n_ext != 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) mbar_init(&s_full[i], 1);
        for (int i = 0; i < 2 * kPChunks; ++i) mbar_init(&p_full[i], 128);
        mbar_init(o_full, 1);
        for (int i = 0; i < kSlots; ++i) mbar_init(&v_fixed[i], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // each role re-reads the TMEM base from shared memory into its own registers (a single kernel-wide value gets
    // spilled to local memory by ptxas and reloaded in front of every MMA)
    auto tmem_base_of = [tmem_slot]() { return *reinterpret_cast<volatile uint32_t*>(tmem_slot); };
    // Programmatic dependent launch.  Normal case: wait for the producer of q / K / V before the first global access.
    // Sequence-parallel overlap (p.no_dep_wait): the preceding kernel on the stream is the peer-push grid, which this
    // kernel deliberately overlaps — it released us only after the kernel that wrote q and the local rows had
    // completed (peer_push_kernel), and the peers' rows are ordered by the epoch flags.
    // The trigger for OUR dependents is issued at the very end of the kernel: this kernel runs for milliseconds and
    // may spin on the peers' flags, so dependents that became resident early would only hold SM resources.
    if (!p.no_dep_wait) griddep_wait();
    // TMEM columns
    // TMEM columns: S0 | S1 | O0 | O1 (128 each)

    // register re-balancing: the producer / MMA warpgroup needs few registers, the softmax warps hold a whole
    // 128-wide score row per thread (the CTA owns 384 x 168 = 64512 registers = 128 x 72 + 256 x 216; asking for more deadlocks the inc)
    if (warp < 4) {
      setmaxnreg_dec<72>();
      if (warp == 0) {
        if (lane == 0) {
            // Q: one or two tiles x two 64-wide halves
            const int nq = two ? 2 : 1;
            mbar_expect_tx(q_full, nq * kTileBytes);
            for (int w = 0; w < nq; ++w)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d_hint(sQ + w * kTileBytes + h * kHalfBytes, &tmQ, q_full, head * kHD + h * 64,
                                     q0 + w * kQT, kEvictFirst);
            int idx = 0;
            TileCursor cur;
            for (int i = 0; i < n_kv; ++i) {
                if (i == tiles.n_old && p.wait_flags != nullptr) wait_peer_flags(p);
                int krow0, kvalid;
                cur.locate(p, tiles.tile(i), krow0, kvalid);
                tile_valid[i & 7] = kvalid;
#pragma unroll
                for (int kv = 0; kv < 2; ++kv, ++idx) {
                    const int s = idx % kSlots;
                    const uint32_t ph = (idx / kSlots) & 1;
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    mbar_expect_tx(&kv_full[s], kTileBytes);
                    const CUtensorMap* tm = kv == 0 ? &tmK : &tmV;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d_hint(sKV + s * kTileBytes + h * kHalfBytes, tm, &kv_full[s],
                                         (head / p.kv_group) * kHD + h * 64, krow0, kEvictLast);
                }
            }
        }
      } else if (warp == 3) {
        if (fix_tails) {
            // every V tile passes through this warp on its way to the MMA warp (v_fixed instead of kv_full); tiles
            // that end an extent get their rows past the extent zeroed first.  Same tile walk as the producer.
            TileCursor cur;
            for (int i = 0; i < n_kv; ++i) {
                int krow0, kvalid;
                cur.locate(p, tiles.tile(i), krow0, kvalid);
                const int vi = 2 * i + 1;
                const int slot = vi % kSlots;
                mbar_wait(&kv_full[slot], static_cast<uint32_t>((vi / kSlots) & 1));
                if (kvalid < kKT) {
                    uint8_t* vt = sKV + slot * kTileBytes;
                    // row r of the tile = 128 bytes at r * 128 in each 64-dim half (the swizzle permutes 16-byte
                    // chunks inside the row only)
                    const int n16 = (kKT - kvalid) * 8;   // 16-byte chunks per half
                    for (int c = lane; c < 2 * n16; c += 32) {
                        const int h = c / n16, o = c % n16;
                        *reinterpret_cast<uint4*>(vt + h * kHalfBytes + kvalid * 128 + o * 16) = make_uint4(0, 0, 0, 0);
                    }
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&v_fixed[slot]);
            }
        }
      }
      if (warp == 2 && p.push.n_ctas > 0 && static_cast<int>(blockIdx.x) < p.push.n_ctas)
          push_slice(p.push, static_cast<int>(blockIdx.x), lane);
      if (warp == 1) {
        if (lane == 0) {
            const uint32_t tmem_base = tmem_base_of();
            auto tS = [tmem_base](int w) { return tmem_base + static_cast<uint32_t>(w) * 128u; };
            auto tO = [tmem_base](int w) { return tmem_base + 256u + static_cast<uint32_t>(w) * 128u; };
            // S = Q K^T : A = Q (K-major), B = K (K-major), M = N = 128
            constexpr uint32_t idesc_qk = make_idesc_bf16(kQT, kKT, 0, 0);
            // O += P V  : A = P (TMEM),   B = V (MN-major: dims contiguous), M = 128, N = head_dim
            constexpr uint32_t idesc_pv = make_idesc_bf16(kQT, kHD, 0, 1);
            auto issue_qk = [&](int w, int slot) {
                const uint32_t qa = smem_u32(sQ + w * kTileBytes);
                const uint32_t ka = smem_u32(sKV + slot * kTileBytes);
#pragma unroll
                for (int k = 0; k < kHD / 16; ++k) {
                    const uint32_t off = (k >> 2) * kHalfBytes + (k & 3) * 32;
                    umma_ss(tS(w), make_smem_desc_sw128(qa + off, 16, 1024), make_smem_desc_sw128(ka + off, 16, 1024),
                            idesc_qk, k != 0);
                }
            };
            auto issue_pv = [&](int w, int slot, bool accumulate, uint32_t parity) {
                const uint32_t va = smem_u32(sKV + slot * kTileBytes);
                constexpr int kStepsPerChunk = (kKT / 16) / kPChunks;
#pragma unroll
                for (int c = 0; c < kPChunks; ++c) {
                    mbar_wait(&p_full[w * kPChunks + c], parity);    // this chunk of P is in TMEM
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < kStepsPerChunk; ++kk) {
                        const int k = c * kStepsPerChunk + kk;
                        // 16 keys = 16 rows of 128 bytes; the two 64-dim halves are kHalfBytes apart (LBO), 8-key
                        // groups 1024 bytes apart (SBO).  P: 16 bf16 = 8 TMEM columns per step.
                        umma_ts(tO(w), tS(w) + k * 8, make_smem_desc_sw128(va + k * 16 * 128, kHalfBytes, 1024),
                                idesc_pv, accumulate || k != 0);
                    }
                }
            };
            auto slot_of = [](int idx) { return idx % kSlots; };
            auto phase_of = [](int idx) { return static_cast<uint32_t>((idx / kSlots) & 1); };

            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[slot_of(0)], phase_of(0));
            tc_fence_after();
            issue_qk(0, slot_of(0));
            umma_commit(&s_full[0]);
            if (two) {
                issue_qk(1, slot_of(0));
                umma_commit(&s_full[1]);
            }
            umma_commit(&kv_empty[slot_of(0)]);
            for (int j = 0; j < n_kv; ++j) {
                const int vi = 2 * j + 1;
                const int kn = 2 * j + 2;
                const bool more = (j + 1 < n_kv);
                mbar_wait(fix_tails ? &v_fixed[slot_of(vi)] : &kv_full[slot_of(vi)], phase_of(vi));
                tc_fence_after();
                issue_pv(0, slot_of(vi), j > 0, j & 1);
                if (more) {
                    mbar_wait(&kv_full[slot_of(kn)], phase_of(kn));
                    tc_fence_after();
                    issue_qk(0, slot_of(kn));
                    umma_commit(&s_full[0]);
                }
                if (two) issue_pv(1, slot_of(vi), j > 0, j & 1);
                umma_commit(&kv_empty[slot_of(vi)]);
                if (more) {
                    if (two) {
                        issue_qk(1, slot_of(kn));
                        umma_commit(&s_full[1]);
                    }
                    umma_commit(&kv_empty[slot_of(kn)]);
                }
            }
            umma_commit(o_full);
        }
      }
    } else {
        setmaxnreg_inc<216>();
        const int w = (warp - 4) >> 2;  // which query tile
        if (w == 0 || two) {
            const int quad = warp & 3;      // TMEM lane quadrant
            const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
            const uint32_t tmem_base = tmem_base_of();
            const uint32_t tS_row = tmem_base + static_cast<uint32_t>(w) * 128u + lane_sel;
            const uint32_t tO_row = tmem_base + 256u + static_cast<uint32_t>(w) * 128u + lane_sel;
            const int row_in_pair = w * kQT + quad * 32 + lane;
            const int row = q0 + row_in_pair;
            const float sl2 = p.scale_log2;

            const uint32_t tile_valid_addr = smem_u32(const_cast<int32_t*>(tile_valid));
            float m_used = -INFINITY;  // max (raw score units) the probabilities are currently referenced to
            float l = 0.f;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&s_full[w], j & 1);
                tc_fence_after();
                uint32_t s[4][32];
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                auto row_max = [&](int c) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(s[c][i]));
                        mx1 = fmaxf(mx1, __uint_as_float(s[c][i + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(s[c][i + 2]));
                        mx3 = fmaxf(mx3, __uint_as_float(s[c][i + 3]));
                    }
                };
                int valid;
                asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(valid) : "r"(tile_valid_addr + (j & 7) * 4));
                if (kLdSplit && valid >= kKT) {
                    // second half of the scores is still on its way from TMEM while the first half's max is taken
                    tmem_ld32(tS_row, s[0]);
                    tmem_ld32(tS_row + 32, s[1]);
                    tmem_wait_ld();
                    tmem_ld32(tS_row + 64, s[2]);
                    tmem_ld32(tS_row + 96, s[3]);
                    row_max(0);
                    row_max(1);
                    tmem_wait_ld();
                    row_max(2);
                    row_max(3);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld32(tS_row + c * 32, s[c]);
                    tmem_wait_ld();
                    if (valid < kKT) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (c * 32 + i >= valid) s[c][i] = __float_as_uint(-INFINITY);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) row_max(c);
                }
                const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_used);
                if (j == 0) {
                    m_used = m_new;
                } else {
                    const bool need = (m_new - m_used) * sl2 > kRescaleThreshold;
                    if (__any_sync(0xffffffffu, need)) {
                        // whole warp rescales (tcgen05.ld/st are warp-collective); rows that did not need it use
                        // their exact (possibly tiny) correction as well, which keeps every row consistent.
                        const float alpha = ex2_approx((m_used - m_new) * sl2);
                        m_used = m_new;
                        l *= alpha;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t o[32];
                            tmem_ld32(tO_row + c * 32, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st32(tO_row + c * 32, o);
                        }
                    }
                }
                const float ms = m_used * sl2;
                float l0 = 0.f, l1 = 0.f;
                if constexpr (kPChunks == 4) {
                    // 32 keys per chunk -> 16 packed TMEM columns.  The store of chunk c is followed by the
                    // exponentials of chunk c + 1; only then does the thread wait for the store and publish chunk c,
                    // so neither the TMEM store latency nor the barrier round trip sits on the MUFU stream.
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float x0 = fmaf(__uint_as_float(s[c][i]), sl2, -ms);
                            const float x1 = fmaf(__uint_as_float(s[c][i + 1]), sl2, -ms);
                            const bool poly = kPolyEvery > 0 && ((i >> 1) % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1;
                            const float p0 = poly ? ex2_poly(x0) : ex2_approx(x0);
                            const float p1 = poly ? ex2_poly(x1) : ex2_approx(x1);
                            l0 += p0;
                            l1 += p1;
                            pk[i >> 1] = pack_bf16x2(p0, p1);
                        }
                        if (c > 0) {
                            tmem_wait_st();
                            tc_fence_before();
                            mbar_arrive(&p_full[w * kPChunks + c - 1]);
                        }
                        tmem_st16(tS_row + c * 16, pk);
                    }
                    l += l0 + l1;
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(&p_full[w * kPChunks + 3]);
                } else {
                    uint32_t pk[2][32];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float x0 = fmaf(__uint_as_float(s[c][i]), sl2, -ms);
                            const float x1 = fmaf(__uint_as_float(s[c][i + 1]), sl2, -ms);
                            // every kPolyEvery-th pair takes the FMA-pipe polynomial instead of MUFU.EX2
                            const bool poly = kPolyEvery > 0 && ((i >> 1) % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1;
                            const float p0 = poly ? ex2_poly(x0) : ex2_approx(x0);
                            const float p1 = poly ? ex2_poly(x1) : ex2_approx(x1);
                            l0 += p0;
                            l1 += p1;
                            pk[c >> 1][(c & 1) * 16 + (i >> 1)] = pack_bf16x2(p0, p1);
                        }
                        if (kPChunks == 2 && c == 1) tmem_st32(tS_row, pk[0]);
                        if (kPChunks == 2 && c == 3) {
                            tmem_wait_st();                      // first half landed while the second was computed
                            tc_fence_before();
                            mbar_arrive(&p_full[w * kPChunks]);
                        }
                    }
                    l += l0 + l1;
                    if (kPChunks == 1) tmem_st32(tS_row, pk[0]);
                    tmem_st32(tS_row + 32, pk[1]);
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(&p_full[w * kPChunks + kPChunks - 1]);
                }
            }

            mbar_wait(o_full, 0);
            tc_fence_after();
            if (piece < 0) {
                // epilogue: O / l -> bf16 -> global
                const float inv_l = 1.0f / l;
                if (p.lse != nullptr && row < p.q_rows)
                    p.lse[static_cast<int64_t>(head) * p.q_rows + row] = (m_used * sl2 + log2f(l)) * 0.6931471805599453f;
                __nv_bfloat16* optr = p.out + static_cast<int64_t>(row) * p.ldo + head * kHD;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[32];
                    tmem_ld32(tO_row + c * 32, o);
                    tmem_wait_ld();
                    if (row < p.q_rows) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            uint4 pkt;
                            pkt.x = pack_bf16x2(__uint_as_float(o[v * 8 + 0]) * inv_l, __uint_as_float(o[v * 8 + 1]) * inv_l);
                            pkt.y = pack_bf16x2(__uint_as_float(o[v * 8 + 2]) * inv_l, __uint_as_float(o[v * 8 + 3]) * inv_l);
                            pkt.z = pack_bf16x2(__uint_as_float(o[v * 8 + 4]) * inv_l, __uint_as_float(o[v * 8 + 5]) * inv_l);
                            pkt.w = pack_bf16x2(__uint_as_float(o[v * 8 + 6]) * inv_l, __uint_as_float(o[v * 8 + 7]) * inv_l);
                            *reinterpret_cast<uint4*>(optr + c * 32 + v * 8) = pkt;
                        }
                    }
                }
            } else {
                // partial: un-normalised O, reference max (log2 units) and sum
                float* po = p.part_o + (static_cast<int64_t>(piece) * (2 * kQT) + row_in_pair) * kHD;
                float* pml = p.part_ml + (static_cast<int64_t>(piece) * (2 * kQT) + row_in_pair) * 2;
                pml[0] = m_used * sl2;
                pml[1] = l;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[32];
                    tmem_ld32(tO_row + c * 32, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int v = 0; v < 8; ++v)
                        *reinterpret_cast<uint4*>(po + c * 32 + v * 4) = make_uint4(o[v * 4], o[v * 4 + 1], o[v * 4 + 2], o[v * 4 + 3]);
                }
                if (p.combine_ctr != nullptr) {
                    // Fused combine ("last block" pattern): every thread publishes its partial row, the 128 threads of
                    // the query tile meet at a named barrier, one of them counts the tile's arrival; the piece that
                    // arrives last reads all pieces of its rows back (L2) and writes the final bf16 rows.  Same
                    // arithmetic, in the same order, as attn_combine_kernel.
                    volatile int32_t* last_flag = tile_valid + 8;
                    const int sitem = item - p.n_whole;
                    __threadfence();
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
                    if (quad == 0 && lane == 0) {
                        __threadfence();
                        const unsigned int old = atomicAdd(p.combine_ctr + sitem * 2 + w, 1u);
                        const bool last = old == static_cast<unsigned int>(p.split - 1);
                        if (last) p.combine_ctr[sitem * 2 + w] = 0;      // every piece has arrived: ready for the next launch
                        last_flag[w] = last ? 1 : 0;
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
                    if (last_flag[w] != 0 && row < p.q_rows) {
                        __threadfence();
                        const int64_t base0 = static_cast<int64_t>(sitem) * p.split * (2 * kQT) + row_in_pair;
                        float m = -INFINITY;
                        for (int sp = 0; sp < p.split; ++sp)
                            m = fmaxf(m, __ldcg(p.part_ml + (base0 + static_cast<int64_t>(sp) * (2 * kQT)) * 2));
                        float lsum = 0.f;
                        float acc[kHD];
#pragma unroll
                        for (int d = 0; d < kHD; ++d) acc[d] = 0.f;
                        for (int sp = 0; sp < p.split; ++sp) {
                            const int64_t base = base0 + static_cast<int64_t>(sp) * (2 * kQT);
                            const float a = ex2_approx(__ldcg(p.part_ml + base * 2) - m);
                            lsum += __ldcg(p.part_ml + base * 2 + 1) * a;
                            const float4* src = reinterpret_cast<const float4*>(p.part_o + base * kHD);
#pragma unroll
                            for (int d = 0; d < kHD / 4; ++d) {
                                const float4 o4 = __ldcg(src + d);
                                acc[4 * d] += o4.x * a;
                                acc[4 * d + 1] += o4.y * a;
                                acc[4 * d + 2] += o4.z * a;
                                acc[4 * d + 3] += o4.w * a;
                            }
                        }
                        const float inv = 1.0f / lsum;
                        if (p.lse != nullptr)
                            p.lse[static_cast<int64_t>(head) * p.q_rows + row] = (m + log2f(lsum)) * 0.6931471805599453f;
                        __nv_bfloat16* optr = p.out + static_cast<int64_t>(row) * p.ldo + head * kHD;
#pragma unroll
                        for (int v = 0; v < kHD / 8; ++v) {
                            uint4 pkt;
                            pkt.x = pack_bf16x2(acc[v * 8 + 0] * inv, acc[v * 8 + 1] * inv);
                            pkt.y = pack_bf16x2(acc[v * 8 + 2] * inv, acc[v * 8 + 3] * inv);
                            pkt.z = pack_bf16x2(acc[v * 8 + 4] * inv, acc[v * 8 + 5] * inv);
                            pkt.w = pack_bf16x2(acc[v * 8 + 6] * inv, acc[v * 8 + 7] * inv);
                            *reinterpret_cast<uint4*>(optr + v * 8) = pkt;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    griddep_launch();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base_of());
    }
}

// Merge the key-range pieces of the split items: one warp per query row, lane owns 4 of the 128 dims.
__global__ void __launch_bounds__(256)
attn_combine_kernel(const AttnParams p) {
    griddep_launch();
    griddep_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 8 + warp;  // (split item, row in pair)
    const int sitem = gw / (2 * kQT);
    const int r = gw % (2 * kQT);
    const int item = p.n_whole + sitem;
    const int head = item / p.num_q_pairs;
    const int row = (item % p.num_q_pairs) * (2 * kQT) + r;
    if (row >= p.q_rows) return;
    float m = -INFINITY;
    for (int s = 0; s < p.split; ++s)
        m = fmaxf(m, p.part_ml[((static_cast<int64_t>(sitem) * p.split + s) * (2 * kQT) + r) * 2]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f;
    for (int s = 0; s < p.split; ++s) {
        const int64_t base = (static_cast<int64_t>(sitem) * p.split + s) * (2 * kQT) + r;
        const float a = ex2_approx(p.part_ml[base * 2] - m);
        l += p.part_ml[base * 2 + 1] * a;
        const float4 o = *reinterpret_cast<const float4*>(p.part_o + base * kHD + lane * 4);
        acc.x += o.x * a;
        acc.y += o.y * a;
        acc.z += o.z * a;
        acc.w += o.w * a;
    }
    const float inv = 1.0f / l;
    if (p.lse != nullptr && lane == 0)
        p.lse[static_cast<int64_t>(head) * p.q_rows + row] = (m + log2f(l)) * 0.6931471805599453f;
    uint2 pkt;
    pkt.x = pack_bf16x2(acc.x * inv, acc.y * inv);
    pkt.y = pack_bf16x2(acc.z * inv, acc.w * inv);
    *reinterpret_cast<uint2*>(p.out + static_cast<int64_t>(row) * p.ldo + head * kHD + lane * 4) = pkt;
}

// Which key rows a launch attends, and (sequence parallel) which of them are still in flight from the peers.
struct KeySpec {
    int n_ext = 0;                 // 0: dense rows [0, kv_rows)
    int32_t row0[kMaxExt];
    int32_t rows[kMaxExt];
    int n_old_ext = -1;            // extents [0, n_old_ext) are resident, the rest sit behind the flag wait; -1: no wait
    const long long* flags = nullptr;
    int world = 0;
    long long epoch = 0;
    unsigned long long timeout_ns = 0;
    bool pdl = false;              // run next to the preceding kernel on the stream instead of after it (only behind a
                                   // kernel that releases this one explicitly, e.g. peer_push_kernel)
    float* lse = nullptr;                   // optional log-sum-exp output [heads][q_rows]
    const PeerPushParams* push = nullptr;   // fused exchange: see AttnParams::push
    int push_ctas = 0;                      // cap on the CTAs sharing the copy (0: default)
};

static void fill_defaults(AttnParams& p) {
    p.lse = nullptr;
    p.combine_ctr = nullptr;
    p.partial = 0;
    p.pieces_per_item = p.piece_count = 1;
    p.piece_first = 0;
    p.n_ext = 0;
    for (int i = 0; i < kMaxExt; ++i) p.ext_row0[i] = p.ext_rows[i] = 0;
    for (int i = 0; i <= kMaxExt; ++i) p.ext_tile0[i] = 0;
    p.n_old_tiles = 0;
    p.wait_world = 0;
    p.wait_flags = nullptr;
    p.wait_epoch = 0;
    p.wait_timeout_ns = 0;
    p.no_dep_wait = 0;
    p.push = PeerPushParams{};
}

// returns the total number of key tiles
static int fill_keys(AttnParams& p, const KeySpec* ks, int64_t kv_rows) {
    if (ks == nullptr || ks->n_ext == 0) return static_cast<int>((kv_rows + kKT - 1) / kKT);
    p.n_ext = ks->n_ext;
    int tiles = 0;
    for (int i = 0; i < ks->n_ext; ++i) {
        p.ext_row0[i] = ks->row0[i];
        p.ext_rows[i] = ks->rows[i];
        p.ext_tile0[i] = tiles;
        if (i == ks->n_old_ext) p.n_old_tiles = tiles;
        tiles += (ks->rows[i] + kKT - 1) / kKT;
    }
    for (int i = ks->n_ext; i <= kMaxExt; ++i) p.ext_tile0[i] = tiles;
    if (ks->n_old_ext >= ks->n_ext) p.n_old_tiles = tiles;
    if (ks->n_old_ext >= 0 && ks->flags != nullptr) {
        p.wait_flags = ks->flags;
        p.wait_world = ks->world;
        p.wait_epoch = ks->epoch;
        p.wait_timeout_ns = ks->timeout_ns;
    }
    return tiles;
}

static ifx_status launch_attn_kernel(int grid, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                     const AttnParams& p, bool overlap_prev, cudaStream_t stream) {
    // the opt-in shared-memory size is a per-device function attribute
    static uint64_t configured_devices = 0;
    int dev = 0;
    IFX_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 64 || !(configured_devices & (1ull << dev))) {
        IFX_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        if (dev < 64) configured_devices |= 1ull << dev;
    }
    // overlap_prev: start next to the preceding kernel instead of after it.  Needs the launch attribute; with
    // IFX_PDL=0 the launch is an ordinary one (stream order), which is still correct — only the overlap is lost.
    AttnParams pp = p;
    pp.no_dep_wait = overlap_prev ? 1 : 0;
    IFX_CUDA_OK(launch_kernel(attn_fwd_kernel, dim3(static_cast<unsigned>(grid)), dim3(kAttnThreads), kAttnSmem, stream,
                              true, tmQ, tmK, tmV, pp));
    return IFX_OK;
}

// The last piece of a split item merges the item's partials inside the attention kernel instead of a separate
// attn_combine_kernel launch (measured: 5.31 vs 5.33 ms sustained at the full shape, 787 vs 792 us at the 1350-row
// shard shape, bit-identical outputs).  IFX_ATTN_FUSED_COMBINE=0 restores the separate launch.
static bool fused_combine() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("IFX_ATTN_FUSED_COMBINE");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// scratch for the split partials of the tail wave, one per device (grown on demand; launches of one device are
// expected on one stream at a time — see the header: handles and scratch are not thread-safe)
struct PartScratch {
    float* ptr = nullptr;
    size_t bytes = 0;
};
static PartScratch g_part[64];

static ifx_status attention_launch(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                                   int64_t ldo, int64_t q_rows, int64_t kv_rows, int32_t heads, int32_t head_dim,
                                   float softmax_scale, cudaStream_t stream, int32_t kv_heads = 0,
                                   const KeySpec* keys = nullptr) {
    if (kv_heads <= 0) kv_heads = heads;
    IFX_CHECK_ARG(heads % kv_heads == 0, "ifx_attention: heads (%d) must be a multiple of kv_heads (%d)", heads, kv_heads);
    IFX_CHECK_ARG(q && k && v && out, "ifx_attention: null pointer");
    IFX_CHECK_ARG(head_dim == kHD, "ifx_attention: head_dim must be 128 (got %d)", head_dim);
    IFX_CHECK_ARG(q_rows > 0 && kv_rows > 0 && heads > 0, "ifx_attention: empty problem (q_rows=%lld kv_rows=%lld)",
                  (long long)q_rows, (long long)kv_rows);
    IFX_CHECK_ARG(q_rows < (1ll << 31) && kv_rows < (1ll << 31), "ifx_attention: sequence too long");
    const int64_t width = static_cast<int64_t>(heads) * head_dim;
    const int64_t kv_width = static_cast<int64_t>(kv_heads) * head_dim;
    IFX_CHECK_ARG(ldq >= width && ldkv >= kv_width && ldo >= width, "ifx_attention: stride smaller than heads*head_dim");
    IFX_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0, "ifx_attention: strides must be multiples of 8");
    IFX_CHECK_ARG(softmax_scale > 0.f, "ifx_attention: softmax_scale must be positive");

    CUtensorMap tmQ, tmK, tmV;
    ifx_status st = make_tmap_bf16_2d(&tmQ, q, (uint64_t)width, (uint64_t)q_rows, (uint64_t)ldq, 64, kQT);
    if (st != IFX_OK) return st;
    st = make_tmap_bf16_2d(&tmK, k, (uint64_t)kv_width, (uint64_t)kv_rows, (uint64_t)ldkv, 64, kKT);
    if (st != IFX_OK) return st;
    st = make_tmap_bf16_2d(&tmV, v, (uint64_t)kv_width, (uint64_t)kv_rows, (uint64_t)ldkv, 64, kKT);
    if (st != IFX_OK) return st;

    AttnParams p;
    fill_defaults(p);
    p.q_rows = static_cast<int32_t>(q_rows);
    p.kv_rows = static_cast<int32_t>(kv_rows);
    p.heads = heads;
    p.kv_group = heads / kv_heads;
    p.num_q_pairs = static_cast<int32_t>((q_rows + 2 * kQT - 1) / (2 * kQT));
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.part_o = nullptr;
    p.part_ml = nullptr;
    p.lse = keys != nullptr ? keys->lse : nullptr;
    const int n_kv = fill_keys(p, keys, kv_rows);
    int64_t key_rows = kv_rows;
    if (p.n_ext) {
        key_rows = 0;
        for (int i = 0; i < p.n_ext; ++i) key_rows += p.ext_rows[i];
    }

    // ---- grid shaping: split the items of the last partial wave along the keys
    const int items = p.num_q_pairs * heads;
    const int sms = sm_count();
    int rem = items % sms;
    int split = 1;
    if (rem != 0 && n_kv >= 2 * kMaxSplit) {
        // minimise ceil(rem * s / sms) / s  (cost of the tail in waves); ties -> smaller s
        double best = 1.0;
        for (int s = 2; s <= kMaxSplit; ++s) {
            const double cost = static_cast<double>((rem * s + sms - 1) / sms) / s;
            if (cost < best - 1e-9) {
                best = cost;
                split = s;
            }
        }
    }
    if (split == 1) rem = 0;
    p.n_whole = items - rem;
    p.split = split;
    const int pieces = rem * split;
    if (pieces > 0) {
        int dev = 0;
        IFX_CUDA_OK(cudaGetDevice(&dev));
        IFX_CHECK_ARG(dev < 64, "ifx_attention: device ordinal %d not supported", dev);
        PartScratch& sc = g_part[dev];
        const size_t need = static_cast<size_t>(pieces) * (2 * kQT) * (kHD + 2) * sizeof(float);
        if (need > sc.bytes) {
            if (sc.ptr) IFX_CUDA_OK(cudaFree(sc.ptr));
            sc.ptr = nullptr;
            sc.bytes = 0;
            IFX_CUDA_OK(cudaMalloc(&sc.ptr, need));
            sc.bytes = need;
        }
        p.part_o = sc.ptr;
        p.part_ml = sc.ptr + static_cast<size_t>(pieces) * (2 * kQT) * kHD;
        if (fused_combine()) {
            // arrival counters of the fused combine: two per split item (one per 128-row query tile), self-resetting
            static unsigned int* g_ctr[64] = {nullptr};
            constexpr int kMaxSplitItems = 4096;
            IFX_CHECK_ARG(rem <= kMaxSplitItems, "ifx_attention: %d split items", rem);
            if (!g_ctr[dev]) {
                IFX_CUDA_OK(cudaMalloc(&g_ctr[dev], 2 * kMaxSplitItems * sizeof(unsigned int)));
                IFX_CUDA_OK(cudaMemset(g_ctr[dev], 0, 2 * kMaxSplitItems * sizeof(unsigned int)));
            }
            p.combine_ctr = g_ctr[dev];
        }
    }
    const int grid = p.n_whole + pieces;
    if (keys != nullptr && keys->push != nullptr) {
        // The CTAs that carry a slice of the exchange must all be resident before any CTA can be blocked on the peers'
        // flags: keep them inside the first wave (lowest block indices are dispatched first) with some margin.
        static unsigned int* g_done[64] = {nullptr};
        int dev = 0;
        IFX_CUDA_OK(cudaGetDevice(&dev));
        IFX_CHECK_ARG(dev < 64, "ifx_attention: device ordinal %d not supported", dev);
        if (!g_done[dev]) {
            IFX_CUDA_OK(cudaMalloc(&g_done[dev], sizeof(unsigned int)));
            IFX_CUDA_OK(cudaMemset(g_done[dev], 0, sizeof(unsigned int)));
        }
        IFX_CHECK_ARG(keys->push->page_tokens == keys->push->world * keys->push->chunk &&
                          keys->push->pl.n == keys->push->frames,
                      "attention: the fused exchange needs frame-sized pages (page_tokens == world * chunk)");
        p.push = *keys->push;
        // 32 CTAs by default: measured at 8 ranks 966 ms / block with 16 or 48 CTAs vs 988 ms with 129 (the copy warp
        // shares an issue port with two softmax warps, so fewer, longer copies disturb less; the flags are still up
        // long before anyone reaches a fresh-page tile)
        int cap = keys->push_ctas > 0 ? keys->push_ctas : 32;
        if (cap > (sms * 7) / 8) cap = (sms * 7) / 8;
        p.push.n_ctas = grid < cap ? grid : cap;
        p.push.done_counter = g_done[dev];
    }
    {
        char label[96];
        snprintf(label, sizeof(label), "attn_fwd_kernel%s[Lq=%d,Lk=%lld,H=%d]", p.push.n_ctas ? "<push>" : "", p.q_rows,
                 (long long)key_rows, heads);
        ProfScope prof(label, stream);
        st = launch_attn_kernel(grid, tmQ, tmK, tmV, p, keys != nullptr && keys->pdl, stream);
        if (st != IFX_OK) return st;
        if (pieces > 0 && p.combine_ctr == nullptr)
            IFX_CUDA_OK(launch_kernel(attn_combine_kernel, dim3(rem * (2 * kQT) / 8), dim3(256), 0, stream, true, p));
    }
    IFX_LAUNCH_OK("attn_fwd_kernel");
    if (pieces > 0 && p.combine_ctr == nullptr) count_launch();
    return IFX_OK;
}

// Key extents of a paged cache: runs of physically consecutive valid pages, the pages of `fresh` (if any) last.
// Softmax attention does not depend on the key order, so the pages are visited in physical order.
static ifx_status kv_key_spec(const KvImpl* kv, const ifx_kv_plan* fresh, KeySpec& ks) {
    const int64_t pt = kv->page_tokens;
    std::vector<int32_t> old_pages, new_pages;
    if (fresh != nullptr) new_pages.assign(fresh->pages, fresh->pages + fresh->num_pages);
    const size_t n_valid = static_cast<size_t>(kv->local_end / pt);   // logical pages [0, n_valid) hold the window
    for (size_t lp = 0; lp < n_valid && lp < kv->table.size(); ++lp) {
        const int32_t pg = kv->table[lp];
        bool is_new = false;
        for (int32_t n : new_pages) is_new = is_new || (n == pg);
        if (!is_new) old_pages.push_back(pg);
    }
    std::sort(old_pages.begin(), old_pages.end());
    std::sort(new_pages.begin(), new_pages.end());
    ks.n_ext = 0;
    auto add_runs = [&](const std::vector<int32_t>& pages) -> bool {
        size_t i = 0;
        while (i < pages.size()) {
            size_t j = i + 1;
            while (j < pages.size() && pages[j] == pages[j - 1] + 1) ++j;
            if (ks.n_ext == kMaxExt) return false;
            ks.row0[ks.n_ext] = static_cast<int32_t>(pages[i] * pt);
            ks.rows[ks.n_ext] = static_cast<int32_t>(static_cast<int64_t>(j - i) * pt);
            ++ks.n_ext;
            i = j;
        }
        return true;
    };
    if (!add_runs(old_pages)) return set_error(IFX_ERR_UNSUPPORTED, "attention: more than %d page runs in the block table", kMaxExt);
    ks.n_old_ext = fresh != nullptr ? ks.n_ext : -1;
    if (!add_runs(new_pages)) return set_error(IFX_ERR_UNSUPPORTED, "attention: more than %d page runs in the block table", kMaxExt);
    // the common case — valid pages are the physical prefix, nothing to wait for — is one dense extent whose tensor
    // map ends at local_end (TMA zero-fills the last partial tile; no tail fix-up needed)
    if (fresh == nullptr && ks.n_ext == 1 && ks.row0[0] == 0) ks.n_ext = 0;
    return IFX_OK;
}

}  // namespace ifx

using namespace ifx;

extern "C" ifx_status ifx_attention_partial(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                            int64_t kv_rows_total, const int64_t* extents, int32_t n_ext,
                                            int64_t q_rows, int32_t heads, int32_t kv_heads, int32_t head_dim,
                                            float softmax_scale, void* workspace, int64_t workspace_bytes,
                                            int32_t pieces_per_item, int32_t piece_first, int32_t piece_count,
                                            void* stream) {
    IFX_CHECK_ARG(q && k && v && extents && workspace, "ifx_attention_partial: null pointer");
    IFX_CHECK_ARG(head_dim == kHD, "ifx_attention_partial: head_dim must be 128");
    IFX_CHECK_ARG(n_ext >= 1 && n_ext <= kMaxExt, "ifx_attention_partial: 1..%d key extents (got %d)", kMaxExt, n_ext);
    if (kv_heads <= 0) kv_heads = heads;
    IFX_CHECK_ARG(heads > 0 && heads % kv_heads == 0, "ifx_attention_partial: bad head counts");
    IFX_CHECK_ARG(q_rows > 0 && q_rows < (1ll << 31) && kv_rows_total > 0 && kv_rows_total < (1ll << 31),
                  "ifx_attention_partial: bad sizes");
    IFX_CHECK_ARG(pieces_per_item >= 1 && piece_count >= 1 && piece_first >= 0 &&
                      piece_first + piece_count <= pieces_per_item,
                  "ifx_attention_partial: piece window [%d, %d) outside [0, %d)", piece_first,
                  piece_first + piece_count, pieces_per_item);
    const int64_t width = static_cast<int64_t>(heads) * head_dim, kv_width = static_cast<int64_t>(kv_heads) * head_dim;
    IFX_CHECK_ARG(ldq >= width && ldkv >= kv_width && ldq % 8 == 0 && ldkv % 8 == 0, "ifx_attention_partial: strides");
    KeySpec ks;
    ks.n_ext = n_ext;
    for (int i = 0; i < n_ext; ++i) {
        const int64_t r0 = extents[2 * i], n = extents[2 * i + 1];
        IFX_CHECK_ARG(r0 >= 0 && n > 0 && r0 + n <= kv_rows_total, "ifx_attention_partial: extent %d out of range", i);
        ks.row0[i] = static_cast<int32_t>(r0);
        ks.rows[i] = static_cast<int32_t>(n);
    }
    AttnParams p;
    fill_defaults(p);
    const int tiles = fill_keys(p, &ks, kv_rows_total);
    IFX_CHECK_ARG(tiles >= piece_count, "ifx_attention_partial: more pieces (%d) than key tiles (%d)", piece_count, tiles);
    p.q_rows = static_cast<int32_t>(q_rows);
    p.kv_rows = static_cast<int32_t>(kv_rows_total);
    p.heads = heads;
    p.kv_group = heads / kv_heads;
    p.num_q_pairs = static_cast<int32_t>((q_rows + 2 * kQT - 1) / (2 * kQT));
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.out = nullptr;
    p.ldo = 0;
    p.n_whole = 0;
    p.split = 1;
    p.partial = 1;
    p.pieces_per_item = pieces_per_item;
    p.piece_first = piece_first;
    p.piece_count = piece_count;
    const int items = p.num_q_pairs * heads;
    const size_t slots = static_cast<size_t>(items) * pieces_per_item;
    const size_t need = slots * (2 * kQT) * (kHD + 2) * sizeof(float);
    IFX_CHECK_ARG(static_cast<size_t>(workspace_bytes) >= need, "ifx_attention_partial: workspace needs %zu bytes", need);
    p.part_o = static_cast<float*>(workspace);
    p.part_ml = p.part_o + slots * (2 * kQT) * kHD;

    CUtensorMap tmQ, tmK, tmV;
    ifx_status st = make_tmap_bf16_2d(&tmQ, q, (uint64_t)width, (uint64_t)q_rows, (uint64_t)ldq, 64, kQT);
    if (st != IFX_OK) return st;
    st = make_tmap_bf16_2d(&tmK, k, (uint64_t)kv_width, (uint64_t)kv_rows_total, (uint64_t)ldkv, 64, kKT);
    if (st != IFX_OK) return st;
    st = make_tmap_bf16_2d(&tmV, v, (uint64_t)kv_width, (uint64_t)kv_rows_total, (uint64_t)ldkv, 64, kKT);
    if (st != IFX_OK) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        char label[96];
        snprintf(label, sizeof(label), "attn_fwd_kernel<partial>[Lq=%d,tiles=%d,H=%d]", p.q_rows, tiles, heads);
        ProfScope prof(label, s);
        st = launch_attn_kernel(items * piece_count, tmQ, tmK, tmV, p, false, s);
        if (st != IFX_OK) return st;
    }
    IFX_LAUNCH_OK("attn_fwd_kernel<partial>");
    return IFX_OK;
}

extern "C" ifx_status ifx_attention_combine(const void* workspace, int32_t pieces_per_item, void* out, int64_t ldo,
                                            int64_t q_rows, int32_t heads, int32_t head_dim, void* stream) {
    IFX_CHECK_ARG(workspace && out, "ifx_attention_combine: null pointer");
    IFX_CHECK_ARG(head_dim == kHD && pieces_per_item >= 1 && q_rows > 0 && heads > 0, "ifx_attention_combine: bad args");
    AttnParams p;
    fill_defaults(p);
    p.q_rows = static_cast<int32_t>(q_rows);
    p.heads = heads;
    p.num_q_pairs = static_cast<int32_t>((q_rows + 2 * kQT - 1) / (2 * kQT));
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.n_whole = 0;
    p.split = pieces_per_item;
    const int items = p.num_q_pairs * heads;
    const size_t slots = static_cast<size_t>(items) * pieces_per_item;
    p.part_o = const_cast<float*>(static_cast<const float*>(workspace));
    p.part_ml = p.part_o + slots * (2 * kQT) * kHD;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("attn_combine_kernel", s);
        IFX_CUDA_OK(launch_kernel(attn_combine_kernel, dim3(items * (2 * kQT) / 8), dim3(256), 0, s, true, p));
    }
    IFX_LAUNCH_OK("attn_combine_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_attention(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                                    int64_t ldo, int64_t q_rows, int64_t kv_rows, int32_t heads, int32_t head_dim,
                                    float softmax_scale, void* stream) {
    return attention_launch(q, ldq, k, v, ldkv, out, ldo, q_rows, kv_rows, heads, head_dim, softmax_scale,
                            static_cast<cudaStream_t>(stream));
}

extern "C" ifx_status ifx_attention_gqa(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                        void* out, int64_t ldo, int64_t q_rows, int64_t kv_rows, int32_t heads,
                                        int32_t kv_heads, int32_t head_dim, float softmax_scale, void* stream) {
    return attention_launch(q, ldq, k, v, ldkv, out, ldo, q_rows, kv_rows, heads, head_dim, softmax_scale,
                            static_cast<cudaStream_t>(stream), kv_heads);
}

extern "C" ifx_status ifx_attention_lse(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                                        int64_t ldo, float* lse, int64_t q_rows, int64_t kv_rows, int32_t heads,
                                        int32_t kv_heads, int32_t head_dim, float softmax_scale, void* stream) {
    IFX_CHECK_ARG(lse != nullptr, "ifx_attention_lse: null lse");
    KeySpec ks;
    ks.lse = lse;
    return attention_launch(q, ldq, k, v, ldkv, out, ldo, q_rows, kv_rows, heads, head_dim, softmax_scale,
                            static_cast<cudaStream_t>(stream), kv_heads, &ks);
}

extern "C" ifx_status ifx_attention_extents(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                            int64_t kv_rows_total, const int64_t* extents, int32_t n_ext, void* out,
                                            int64_t ldo, int64_t q_rows, int32_t heads, int32_t kv_heads,
                                            int32_t head_dim, float softmax_scale, void* stream) {
    IFX_CHECK_ARG(extents != nullptr && n_ext >= 1 && n_ext <= kMaxExt, "ifx_attention_extents: 1..%d key extents (got %d)",
                  kMaxExt, n_ext);
    IFX_CHECK_ARG(kv_rows_total > 0 && kv_rows_total < (1ll << 31), "ifx_attention_extents: bad kv_rows_total");
    KeySpec ks;
    ks.n_ext = n_ext;
    for (int i = 0; i < n_ext; ++i) {
        const int64_t r0 = extents[2 * i], n = extents[2 * i + 1];
        IFX_CHECK_ARG(r0 >= 0 && n > 0 && r0 + n <= kv_rows_total, "ifx_attention_extents: extent %d out of range", i);
        ks.row0[i] = static_cast<int32_t>(r0);
        ks.rows[i] = static_cast<int32_t>(n);
    }
    return attention_launch(q, ldq, k, v, ldkv, out, ldo, q_rows, kv_rows_total, heads, head_dim, softmax_scale,
                            static_cast<cudaStream_t>(stream), kv_heads, &ks);
}

namespace ifx {
ifx_status attention_kv_launch(const void* q, int64_t ldq, const ifx_kv* kv_, void* out, int64_t ldo, int64_t q_rows,
                               float softmax_scale, const ifx_kv_plan* fresh, const int64_t* flags, int32_t world,
                               int64_t epoch, int32_t timeout_ms, bool pdl, cudaStream_t stream,
                               const PeerPushParams* push) {
    const KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_attention_kv: bad kv handle");
    IFX_CHECK_ARG(kv->local_end > 0, "ifx_attention_kv: cache is empty");
    IFX_CHECK_ARG(kv->local_end % kv->page_tokens == 0 &&
                      kv->local_end <= static_cast<int64_t>(kv->table.size()) * kv->page_tokens,
                  "ifx_attention_kv: the valid window (%lld tokens) must be whole mapped pages", (long long)kv->local_end);
    KeySpec ks;
    ifx_status st = kv_key_spec(kv, fresh, ks);
    if (st != IFX_OK) return st;
    if (fresh != nullptr) {
        IFX_CHECK_ARG(flags != nullptr && world >= 1 && world <= IFX_MAX_PEERS && epoch > 0 && timeout_ms > 0,
                      "ifx_attention_kv_wait: bad flags / world / epoch / timeout");
        ks.flags = reinterpret_cast<const long long*>(flags);
        ks.world = world;
        ks.epoch = epoch;
        ks.timeout_ns = static_cast<unsigned long long>(timeout_ms) * 1000000ull;
    }
    ks.pdl = pdl;
    IFX_CHECK_ARG(push == nullptr || fresh != nullptr, "attention: the fused exchange needs the plan of the fresh pages");
    ks.push = push;
    ks.push_ctas = push ? push->n_ctas : 0;
    const int64_t width = static_cast<int64_t>(kv->heads) * kv->head_dim;
    // dense: the tensor map ends at local_end; extents: it covers the whole buffer
    const int64_t map_rows = ks.n_ext == 0 ? kv->local_end : static_cast<int64_t>(kv->num_pages) * kv->page_tokens;
    return attention_launch(q, ldq, kv->k_base, kv->v_base, width, out, ldo, q_rows, map_rows, kv->heads, kv->head_dim,
                            softmax_scale, stream, 0, &ks);
}
}  // namespace ifx

extern "C" ifx_status ifx_attention_kv(const void* q, int64_t ldq, const ifx_kv* kv_, void* out, int64_t ldo,
                                       int64_t q_rows, float softmax_scale, void* stream) {
    return attention_kv_launch(q, ldq, kv_, out, ldo, q_rows, softmax_scale, nullptr, nullptr, 0, 0, 0, false,
                               static_cast<cudaStream_t>(stream));
}

extern "C" ifx_status ifx_attention_kv_wait(const void* q, int64_t ldq, const ifx_kv* kv_, const ifx_kv_plan* fresh,
                                            const int64_t* flags, int32_t world, int64_t epoch, int32_t timeout_ms,
                                            void* out, int64_t ldo, int64_t q_rows, float softmax_scale, void* stream) {
    IFX_CHECK_ARG(fresh != nullptr, "ifx_attention_kv_wait: the plan naming the in-flight pages is required");
    return attention_kv_launch(q, ldq, kv_, out, ldo, q_rows, softmax_scale, fresh, flags, world, epoch, timeout_ms, false,
                               static_cast<cudaStream_t>(stream));
}

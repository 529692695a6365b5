// fastq_fused.cu -- single-pass fused parse -> probe -> compact kernel for CANONICAL FASTQ.
//
// Replaces the whole of FastqCleaner::clean_reads (cleaner.rs:731-760) -- needletail framing,
// get_id (utils.rs:91-103), `read_ids.contains` and `record.write` -- with ONE kernel that
// reads every input byte once and writes every output byte once.
//
// Canonical input = LF line endings, bare "+" separator lines, a final newline, ASCII
// headers, well-formed records.  For such input the reference's output is a byte partition
// of the input, so the work is an order-preserving stream compaction of
// records.  Anything else (CRLF, "+id" separators, missing final newline, non-ASCII headers,
// parse errors, pathological line density) raises the device `fallback` flag and the caller
// re-runs the always-exact general path (fastq_general.cu); nothing is approximated.
//
// Structure: one persistent, warp-specialised CTA per SM over a ring of NS shared-memory stages.
//   producer warp : takes tile tickets in order (dynamic, so every tile a look-back spins on is owned
//                   by a running CTA) and streams tile t (TILE bytes + 16 B pre-halo + post-halo)
//                   into a free stage with the TMA engine (cp.async.bulk + mbarrier);
//   parse group   : (GT threads) P1 16-byte vector loads -> '\n' bit masks (SWAR) and counts;
//                   P2 newline positions compacted, line phase (line number mod 4) SPECULATED from
//                   the first "\n+\n" in the tile and published at once (a tile without one falls
//                   back to a real decoupled look-back over 2-bit phases); every newline classified
//                   by role: CR / "+\n" checks, record starts;
//                   P3 one thread per record start: '@', id token, hash, exact probe of the id set,
//                   seq/qual length check, block scan of kept bytes, aggregate published;
//   copy group    : (GT threads) decoupled look-back #2 over (kept bytes, keep-flag of the record that
//                   straddles the tile edge), then a destination-driven copy: every 16-byte aligned
//                   output chunk is owned by one thread (coalesced 16-byte stores), its source run
//                   found through a marker array + max-scan, the source re-aligned with funnel shifts.
// The parse group runs ahead of the copy group through the ring, so probe latency, look-back waits,
// TMA loads and output stores of different tiles overlap inside one SM.  The speculated phases and the
// cross-tile seq/qual length sums are verified exactly by a tiny follow-up kernel over per-tile
// metadata; any mismatch is a fallback, never a wrong answer.
// Records may straddle any number of tiles (ONT reads); only the id token must lie within the
// post-halo of the tile where the record starts.
#include <stdlib.h>

#include "fastq_records.cuh"

namespace sgpu {

constexpr int GT = 256;                    // threads per consumer group
constexpr int NPARSE = 2;                  // parse groups (alternate tiles)
constexpr int NTHREADS = 32 + (NPARSE + 1) * GT;  // producer warp + parse groups + copy group
constexpr int FC = 8;                      // 16-byte chunks per parse thread
constexpr int TILE = GT * FC * 16;         // 32 KiB
constexpr int PRE = 16;                    // pre-halo (previous 16 bytes)
constexpr int HALO = 1024;                 // post-halo
constexpr int BUF = PRE + TILE + HALO;     // bytes per stage buffer
constexpr int RMAX = 512;                  // record starts per tile
constexpr int LMAX = 4 * RMAX + 8;         // newline list capacity per tile
constexpr int NS = 4;                      // ring stages

constexpr uint64_t ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = 3ull << 62;
constexpr uint64_t D2_START = 1ull << 61, D2_FLAG = 1ull << 60;

struct FusedResult {
    unsigned long long fallback;   // != 0: input is not canonical, use the general path
    unsigned long long kept_total; // bytes written to out_w
    unsigned long long reads_in, reads_out;
    unsigned long long ticket;     // dynamic tile counter
    unsigned long long reason;     // first fallback reason (diagnostics)
};

struct FusedParams {
    const uint8_t *in;
    uint64_t n_in;
    uint64_t n_tiles;
    uint8_t *out_w, *out_o;
    int reverse;
    IdSetView set;
    unsigned long long *desc1, *desc2;  // per tile look-back descriptors (zero initialised)
    long long *sum_total, *sum_head;    // per tile signed newline-position sums (length check)
    uint32_t *nl_count;                 // per tile newline count (phase verification, reads_in)
    uint8_t *has_term, *phase_used;     // per tile: has a record end; line phase (mod 4) the tile assumed
    FusedResult *res;
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// named barrier over one consumer group (ids 1 .. NPARSE+1; id 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(GT) : "memory"); }

// The look-back descriptors carry their whole payload in one 64-bit word, so relaxed
// (non-fencing) gpu-scope accesses are sufficient: nothing else is ordered against them.
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_global_v4(void *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ------------------------------------------------------------------ group-wide decoupled look-back
// At >3 TB/s a 32 KiB tile retires every ~10 ns chip-wide, far faster than a 32-descriptor warp
// window can be walked (one L2 round trip per window), so a whole group looks back: GT descriptors
// per round, which normally covers every tile in flight (148 SMs x NS stages) in two or three rounds.
struct GroupSmem {
    uint32_t wmin[GT / 32], Sb[GT / 32], Fb[GT / 32], wmax[GT / 32];
    uint64_t red[2][GT / 32];
    uint64_t inc_val;
    uint64_t scan[80];
    uint32_t spec_rank, pad;
};

// look-back #1 (rare: tiles without a "\n+\n"): exclusive prefix of the newline counts, mod 4.
__device__ __forceinline__ uint32_t lookback_phase(unsigned long long *desc, uint64_t t, uint32_t mine, GroupSmem *L,
                                                   int gt, int bar) {
    const int lane = gt & 31, warp = gt >> 5;
    constexpr int NW = GT / 32;
    if (gt == 0) st_relaxed(desc + t, ST_AGG | (mine & 3));
    uint64_t acc = 0;
    int64_t base = (int64_t)t - 1;
    while (true) {
        const int64_t idx = base - gt;
        unsigned long long d = ST_INC;  // virtual tiles before the file: inclusive prefix 0
        if (idx >= 0) {
            while (((d = ld_relaxed(desc + idx)) & ST_MASK) == 0) __nanosleep(64);
        }
        const bool is_inc = (d & ST_MASK) == ST_INC;
        const unsigned b = __ballot_sync(0xffffffffu, is_inc);
        if (lane == 0) L->wmin[warp] = b ? (uint32_t)(warp * 32 + __ffs(b) - 1) : 0xFFFFFFFFu;
        group_sync(bar);
        uint32_t first = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < NW; w++) first = min(first, L->wmin[w]);
        const uint64_t v = ((uint32_t)gt <= first) ? (d & 3) : 0;
        const uint64_t sw = warp_sum(v);
        if (lane == 0) L->red[0][warp] = sw;
        group_sync(bar);
#pragma unroll
        for (int w = 0; w < NW; w++) acc += L->red[0][w];
        group_sync(bar);
        if (first != 0xFFFFFFFFu) break;
        base -= GT;
    }
    if (gt == 0) st_relaxed(desc + t, ST_INC | ((acc + mine) & 3));
    return (uint32_t)(acc & 3);
}

// look-back #2: kept bytes before the tile and the keep-flag of the record that straddles its edge.
// aggregate:  [61] has_start  [60] last_flag  [59:30] head_len  [29:0] rest_kept
// inclusive:  [60] carry flag after the tile  [59:0] kept bytes up to and including the tile
// A tile without a record start passes its predecessor's flag through and keeps head_len bytes iff
// that flag is set, so contributions are resolved against the nearest "provider" farther back.
// (The aggregate itself is published by the parse group as soon as the tile is parsed.)
__device__ __forceinline__ void lookback_kept(unsigned long long *desc, uint64_t t, bool has_start, bool last_flag,
                                              uint32_t head_len, uint32_t rest, GroupSmem *L, int gt, int bar,
                                              uint64_t *kept_before, bool *carry_flag) {
    const int lane = gt & 31, warp = gt >> 5;
    constexpr int NW = GT / 32;
    uint64_t acc = 0, pending = 0;
    bool known = false, my_flag = false;
    int64_t base = (int64_t)t - 1;
    while (true) {
        const int64_t idx = base - gt;
        unsigned long long d = ST_INC;  // virtual tiles before the file: nothing kept, flag 0
        if (idx >= 0) {
            while (((d = ld_relaxed(desc + idx)) & ST_MASK) == 0) __nanosleep(64);
        }
        const bool is_inc = (d & ST_MASK) == ST_INC;
        const unsigned b = __ballot_sync(0xffffffffu, is_inc);
        if (lane == 0) L->wmin[warp] = b ? (uint32_t)(warp * 32 + __ffs(b) - 1) : 0xFFFFFFFFu;
        group_sync(bar);
        uint32_t first = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < NW; w++) first = min(first, L->wmin[w]);
        const bool relevant = (uint32_t)gt <= first;
        const bool provides = relevant && (is_inc || (d & D2_START));
        const unsigned Sw = __ballot_sync(0xffffffffu, provides);
        const unsigned Fw = __ballot_sync(0xffffffffu, provides && (d & D2_FLAG));
        if (lane == 0) {
            L->Sb[warp] = Sw;
            L->Fb[warp] = Fw;
        }
        if ((uint32_t)gt == first) L->inc_val = d & 0x0FFFFFFFFFFFFFFFull;
        group_sync(bar);
        uint64_t contrib = 0, defer = 0;
        if (relevant && !is_inc) {
            const uint64_t hl = (d >> 30) & 0x3FFFFFFFull, rs = d & 0x3FFFFFFFull;
            const unsigned above = lane < 31 ? (Sw >> (lane + 1)) << (lane + 1) : 0u;  // providers farther back
            int fl = -1;
            if (above) {
                fl = (Fw >> (__ffs(above) - 1)) & 1u;
            } else {
                for (int w2 = warp + 1; w2 < NW; w2++) {
                    const unsigned s2 = L->Sb[w2];
                    if (s2) {
                        fl = (L->Fb[w2] >> (__ffs(s2) - 1)) & 1u;
                        break;
                    }
                }
            }
            contrib = rs + (fl > 0 ? hl : 0);
            if (fl < 0) defer = hl;  // its flag lies in a farther round
        }
        int f0 = -1;  // the nearest provider of this round resolves what earlier rounds deferred
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const unsigned s2 = L->Sb[w];
            if (f0 < 0 && s2) f0 = (L->Fb[w] >> (__ffs(s2) - 1)) & 1u;
        }
        if (f0 >= 0) {
            if (f0) acc += pending;
            pending = 0;
            if (!known) {
                known = true;
                my_flag = f0 != 0;
            }
        }
        const uint64_t cw = warp_sum(contrib), dw = warp_sum(defer);
        if (lane == 0) {
            L->red[0][warp] = cw;
            L->red[1][warp] = dw;
        }
        group_sync(bar);
#pragma unroll
        for (int w = 0; w < NW; w++) {
            acc += L->red[0][w];
            pending += L->red[1][w];
        }
        const uint64_t incv = L->inc_val;
        group_sync(bar);
        if (first != 0xFFFFFFFFu) {
            acc += incv;
            break;
        }
        base -= GT;
    }
    *kept_before = acc;
    *carry_flag = my_flag;
    const uint64_t incl = acc + (my_flag ? head_len : 0) + rest;
    const bool out_flag = has_start ? last_flag : my_flag;
    if (gt == 0) st_relaxed(desc + t, ST_INC | (out_flag ? D2_FLAG : 0) | incl);
}

// ------------------------------------------------------------------ shared memory
struct __align__(16) Stage {
    uint64_t tile;             // tile index (~0 = no more tiles)
    uint32_t n_starts, head_len, rest_total, tile_len;
    uint16_t runS[RMAX + 4];   // run r starts at runS[r]; run 0 = carried-in head, run j+1 = record j; sentinel = tile_len
    uint8_t runF[RMAX + 4];    // 1: the run goes to out_w
    uint32_t runK[RMAX + 4];   // kept bytes of the records before run r (head excluded)
    // nlp[] (parse group: local newline positions) and crun[] (copy group: per destination chunk,
    // 1 + index of the run it lies inside) are never live at the same time
    union {
        __align__(16) uint16_t nlp[LMAX];
        __align__(16) uint16_t crun[TILE / 16];
    };
    __align__(16) uint8_t buf[BUF];
};

struct __align__(16) FusedSmem {
    uint64_t full[NS], parsed[NS], empty[NS];
    GroupSmem g[NPARSE + 1];
    Stage st[NS];
};

__device__ __forceinline__ void set_fallback(FusedResult *res, int reason) {
    if (atomicExch(&res->fallback, 1ull) == 0ull) res->reason = (unsigned long long)reason;
}

// group-wide exclusive scan of two values per thread at once (one set of barriers)
__device__ __forceinline__ void group_scan2(uint64_t a, uint64_t b, uint64_t *pa, uint64_t *pb, uint64_t *ta,
                                            uint64_t *tb, uint64_t *sm /* >= 80 */, int gt, int bar) {
    const int lane = gt & 31, warp = gt >> 5;
    constexpr int NW = GT / 32;
    uint64_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t x = __shfl_up_sync(0xffffffffu, ia, d), y = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) {
            ia += x;
            ib += y;
        }
    }
    if (lane == 31) {
        sm[warp] = ia;
        sm[40 + warp] = ib;
    }
    group_sync(bar);
    uint64_t ba = 0, bb = 0, sa = 0, sb = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const uint64_t x = sm[w], y = sm[40 + w];
        if (w < warp) {
            ba += x;
            bb += y;
        }
        sa += x;
        sb += y;
    }
    group_sync(bar);
    *pa = ba + ia - a;
    *pb = bb + ib - b;
    *ta = sa;
    *tb = sb;
}

// 16-bit mask (bit i <-> byte i) of the bytes equal to '\n' in a 16-byte chunk.
// Per word: exact zero-byte test of (w ^ 0x0a0a0a0a), then the four flag bits (7,15,23,31) are
// gathered into a nibble with one multiply: ((t >> 7) * 0x10204080) >> 28.
__device__ __forceinline__ uint32_t nl_nibble(uint32_t w) {
    const uint32_t x = w ^ 0x0a0a0a0au;
    const uint32_t t = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
    return ((t >> 7) * 0x10204080u) >> 28;
}
__device__ __forceinline__ uint32_t nl_mask16_fast(uint4 v) {
    return nl_nibble(v.x) | (nl_nibble(v.y) << 4) | (nl_nibble(v.z) << 8) | (nl_nibble(v.w) << 12);
}

// 16 bytes from shared memory at an arbitrary byte address (two aligned 16-byte loads + funnel shifts)
__device__ __forceinline__ uint4 lds_unaligned16(const uint8_t *src) {
    const uint32_t sa = smem_u32(src);
    const uint32_t q = sa & 15, qw = q >> 2, qb = (q & 3) * 8;
    const uint8_t *sbase = src - q;
    const uint4 a = *reinterpret_cast<const uint4 *>(sbase);
    const uint4 b = *reinterpret_cast<const uint4 *>(sbase + 16);
    // select the five consecutive words starting at word qw of {a, b}
    const uint32_t w0 = qw == 0 ? a.x : qw == 1 ? a.y : qw == 2 ? a.z : a.w;
    const uint32_t w1 = qw == 0 ? a.y : qw == 1 ? a.z : qw == 2 ? a.w : b.x;
    const uint32_t w2 = qw == 0 ? a.z : qw == 1 ? a.w : qw == 2 ? b.x : b.y;
    const uint32_t w3 = qw == 0 ? a.w : qw == 1 ? b.x : qw == 2 ? b.y : b.z;
    const uint32_t w4 = qw == 0 ? b.x : qw == 1 ? b.y : qw == 2 ? b.z : b.w;
    uint4 o;
    o.x = __funnelshift_r(w0, w1, qb);
    o.y = __funnelshift_r(w1, w2, qb);
    o.z = __funnelshift_r(w2, w3, qb);
    o.w = __funnelshift_r(w3, w4, qb);
    return o;
}

// ------------------------------------------------------------------ producer warp
__device__ __forceinline__ void producer_loop(const FusedParams &P, FusedSmem *S) {
    if ((threadIdx.x & 31) != 0) return;
    for (uint32_t it = 0;; it++) {
        const uint32_t s = it % NS;
        mbar_wait(&S->empty[s], ((it / NS) & 1) ^ 1);  // passes at once on the first lap
        const unsigned long long t = atomicAdd(&P.res->ticket, 1ull);
        Stage *st = &S->st[s];
        if (t >= P.n_tiles) {
            // no more tiles: one sentinel per parse group travels through the ring behind the last tile
            for (uint32_t k = 0; k < (uint32_t)NPARSE; k++) {
                const uint32_t s2 = (it + k) % NS;
                if (k) mbar_wait(&S->empty[s2], (((it + k) / NS) & 1) ^ 1);
                S->st[s2].tile = ~0ull;
                mbar_arrive(&S->full[s2]);
            }
            return;
        }
        st->tile = t;
        // bytes [t*TILE - PRE, t*TILE + TILE + HALO) clipped to the file, rounded up to 16
        const uint64_t g0 = t * (uint64_t)TILE;
        const uint64_t src0 = t ? g0 - PRE : 0;
        uint64_t end = g0 + TILE + HALO;
        if (end > P.n_in) end = P.n_in;
        const uint32_t bytes = (uint32_t)(((end - src0) + 15) & ~15ull);
        mbar_expect_tx(&S->full[s], bytes);
        bulk_g2s(st->buf + (t ? 0 : PRE), P.in + src0, bytes, &S->full[s]);
    }
}

// ------------------------------------------------------------------ parse group (P1, P2, P3)
__device__ __forceinline__ void parse_loop(const FusedParams &P, FusedSmem *S, const int gt, const int pg) {
    const int BAR = 1 + pg;
    GroupSmem *G = &S->g[pg];
    unsigned long long my_reads_out = 0;  // thread 0 only
    for (uint32_t it = pg;; it += NPARSE) {  // parse group pg takes every NPARSE-th ring slot
        const uint32_t s = it % NS;
        Stage *st = &S->st[s];
        mbar_wait(&S->full[s], (it / NS) & 1);
        const uint64_t t = st->tile;
        if (t == ~0ull) {
            if (gt == 0) mbar_arrive(&S->parsed[s]);  // pass the sentinel on to the copy group
            break;
        }
        uint8_t *buf = st->buf;
        if (t == 0 && gt < PRE) buf[gt] = '\n';  // no predecessor: the pre-halo reads as a newline
        if (gt == 0) G->spec_rank = 0xFFFFFFFFu;
        const uint64_t g0 = t * (uint64_t)TILE;
        const uint32_t tile_len = (uint32_t)((P.n_in - g0) < (uint64_t)TILE ? (P.n_in - g0) : (uint64_t)TILE);
        const uint32_t avail =
            (uint32_t)((P.n_in - g0) < (uint64_t)(TILE + HALO) ? (P.n_in - g0) : (uint64_t)(TILE + HALO));
        const uint8_t *tile = buf + PRE;  // tile[-16 .. avail)
        uint32_t fb = 0;                  // this thread's fallback reason (0 = none)
        group_sync(BAR);                  // pre-halo fill of tile 0 visible

        // ---- P1: newline masks, counts, high-bit test
        uint32_t m[FC];
        uint32_t hi_or = 0;
        uint64_t packed[2] = {0, 0};
        if (tile_len == (uint32_t)TILE) {
#pragma unroll
            for (int k = 0; k < FC; k++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(tile + (uint32_t)(k * GT + gt) * 16);
                m[k] = nl_mask16_fast(v);
                hi_or |= (v.x | v.y | v.z | v.w);
                packed[k >> 2] |= (uint64_t)__popc(m[k]) << (16 * (k & 3));
            }
        } else {
#pragma unroll
            for (int k = 0; k < FC; k++) {  // last tile: bytes past the end of the file are stale
                const uint32_t pos = (uint32_t)(k * GT + gt) * 16;
                uint4 v = make_uint4(0, 0, 0, 0);
                uint32_t mm = 0;
                if (pos < tile_len) {
                    v = *reinterpret_cast<const uint4 *>(tile + pos);
                    mm = nl_mask16_fast(v);
                    const uint32_t valid = tile_len - pos;
                    if (valid < 16) {
                        mm &= (1u << valid) - 1u;
                        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int x = 0; x < 4; x++) {
                            const int rem = (int)valid - 4 * x;
                            if (rem <= 0) w[x] = 0;
                            else if (rem < 4) w[x] &= (1u << (8 * rem)) - 1u;
                        }
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                m[k] = mm;
                hi_or |= (v.x | v.y | v.z | v.w);
                packed[k >> 2] |= (uint64_t)__popc(mm) << (16 * (k & 3));
            }
        }
        if (hi_or & 0x80808080u) fb = 1;  // reason 1: non-ASCII byte, Unicode rules needed
        uint64_t pre[2], tot[2];
        group_scan2(packed[0], packed[1], &pre[0], &pre[1], &tot[0], &tot[1], G->scan, gt, BAR);
        uint32_t n_nl = 0;
        uint32_t row_base[FC];
#pragma unroll
        for (int k = 0; k < FC; k++) {
            row_base[k] = n_nl + (uint32_t)((pre[k >> 2] >> (16 * (k & 3))) & 0xFFFF);
            n_nl += (uint32_t)((tot[k >> 2] >> (16 * (k & 3))) & 0xFFFF);
        }
        const bool dense = n_nl > (uint32_t)LMAX;
        // ---- P2a: compact the newline positions (needs only the in-tile ranks)
        if (!dense) {
#pragma unroll
            for (int k = 0; k < FC; k++) {
                uint32_t mm = m[k];
                uint32_t r = row_base[k];
                const uint32_t pos = (uint32_t)(k * GT + gt) * 16;
                while (mm) {
                    st->nlp[r++] = (uint16_t)(pos + (uint32_t)(__ffs(mm) - 1));
                    mm &= mm - 1;
                }
            }
        }
        group_sync(BAR);
        // ---- line phase: the first newline followed by "+\n" ends a sequence line (role 1)
        if (!dense && t != 0) {
            uint32_t best = 0xFFFFFFFFu;
            for (uint32_t i = gt; i < n_nl; i += GT) {
                const uint32_t p = st->nlp[i];
                if (p + 2 < avail && tile[p + 1] == '+' && tile[p + 2] == '\n') {
                    best = i;
                    break;
                }
            }
            if (best != 0xFFFFFFFFu) atomicMin(&G->spec_rank, best);
        }
        group_sync(BAR);
        uint32_t c0;  // newlines before this tile, mod 4
        {
            const uint32_t sr = G->spec_rank;
            if (t == 0) {
                c0 = 0;
                if (gt == 0) st_relaxed(P.desc1 + t, ST_INC | (n_nl & 3));
            } else if (sr != 0xFFFFFFFFu) {
                c0 = (1u - sr) & 3u;  // role(sr) = (c0 + sr) & 3 == 1
                if (gt == 0) st_relaxed(P.desc1 + t, ST_INC | ((c0 + n_nl) & 3));
            } else {
                c0 = lookback_phase(P.desc1, t, n_nl, G, gt, BAR);
            }
        }
        // position 0 starts a record iff 4k newlines precede it and the previous byte is one
        const bool pos0_start = (c0 == 0) && tile[-1] == '\n';
        // terminating newlines (role 3) with in-tile rank < i: floor((c0 + i) / 4)

        // ---- P2b: one thread per newline: classify by role (line number mod 4)
        if (!dense) {
            for (uint32_t i = gt; i < n_nl; i += GT) {
                const uint32_t p = st->nlp[i];
                const uint32_t role = (c0 + i) & 3;
                if (tile[(int)p - 1] == '\r') fb = 3;  // CRLF: not canonical
                if (role == 1) {                       // end of the sequence line: "+\n" must follow
                    if (p + 2 >= avail) fb = 4;
                    else if (tile[p + 1] != '+' || tile[p + 2] != '\n') fb = 4;
                } else if (role == 3 && p + 1 < tile_len) {  // record j starts at p + 1 (run j + 1)
                    const uint32_t j = ((c0 + i) >> 2) + (pos0_start ? 1u : 0u);
                    if (j < (uint32_t)RMAX) st->runS[j + 1] = (uint16_t)(p + 1);
                    else fb = 5;
                }
            }
        } else {
            fb = 2;
        }
        if (gt == 0) {
            st->runS[0] = 0;
            if (pos0_start) st->runS[1] = 0;
        }
        // number of record starts inside the tile
        const uint32_t n_term = (c0 + n_nl) >> 2;
        uint32_t n_starts = n_term + (pos0_start ? 1u : 0u);
        if (n_term > 0 && !dense) {
            // the last terminating newline may sit on the tile's final byte: its record belongs to the next tile
            const uint32_t r_last = ((3u - c0) & 3u) + 4u * (n_term - 1);
            if ((uint32_t)st->nlp[r_last] + 1u >= tile_len) n_starts--;
        }
        if (dense || n_starts > (uint32_t)RMAX) n_starts = 0;  // (a fallback reason is already raised)
        group_sync(BAR);
        if (gt == 0) st->runS[n_starts + 1] = (uint16_t)tile_len;  // sentinel (TILE <= 32768 fits)

        // ---- P3: one thread per record start: '@', id token, exact probe, seq/qual length check
        uint64_t rest_total = 0;
        for (uint32_t jb = 0; jb < n_starts; jb += GT) {
            const uint32_t j = jb + gt;
            uint32_t my_len = 0;
            bool my_flag = false;
            if (j < n_starts) {
                const uint32_t sp = st->runS[j + 1];
                const uint32_t e = (j + 1 < n_starts) ? st->runS[j + 2] : tile_len;
                my_len = e - sp;
                uint32_t why = tile[sp] == '@' ? 0u : 6u;  // 6 '@', 7 id token, 8 seq/qual lengths
                // id token: skip leading blanks, run to the next blank / newline (ASCII: high bytes fell back)
                uint32_t i = sp + 1;
                while (i < avail && is_ws_ascii(tile[i]) && tile[i] != '\n') i++;
                uint32_t q = i;
                while (q < avail && !is_ws_ascii(tile[q])) q++;
                if (q >= avail || q == i) why = why ? why : 7u;  // token past the halo, or empty id (error 9)
                if (!why) {
                    const bool hit = idset_contains(P.set, tile + i, q - i);
                    my_flag = P.reverse ? hit : !hit;
                }
                // seq/qual length equality for records whose four newlines are inside the tile
                const int r0 = pos0_start ? 4 * (int)j - 1 : (int)((3u - c0) & 3u) + 4 * (int)j;
                if (r0 + 4 < (int)n_nl) {
                    const int sgn = -(int)st->nlp[r0 + 1] + (int)st->nlp[r0 + 2] + (int)st->nlp[r0 + 3] - (int)st->nlp[r0 + 4];
                    if (sgn != 0) why = why ? why : 8u;
                }
                if (why) fb = why;
                st->runF[j + 1] = my_flag ? 1 : 0;
            }
            uint64_t koff, kcnt, round_total, round_cnt;
            group_scan2(my_flag ? my_len : 0u, my_flag ? 1u : 0u, &koff, &kcnt, &round_total, &round_cnt, G->scan, gt, BAR);
            if (j < n_starts) st->runK[j + 1] = (uint32_t)(rest_total + koff);
            rest_total += round_total;
            my_reads_out += round_cnt;
        }
        if (fb) set_fallback(P.res, (int)fb);
        group_sync(BAR);  // runS / runF / runK complete (also when the loop ran zero times)
        if (gt == 0) {
            const uint32_t head_len = st->runS[1];  // == tile_len when no record starts in the tile
            const bool last_flag = n_starts ? (st->runF[n_starts] != 0) : false;
            // aggregate for look-back #2, visible to every later tile from here on
            st_relaxed(P.desc2 + t, ST_AGG | (n_starts ? D2_START : 0) | (last_flag ? D2_FLAG : 0) |
                                        ((uint64_t)head_len << 30) | (uint32_t)rest_total);
            st->n_starts = n_starts;
            st->head_len = head_len;
            st->rest_total = (uint32_t)rest_total;
            st->tile_len = tile_len;
            // signed newline-position sums: -p1 +p2 +p3 -p4 per record must vanish
            long long head = 0, total = 0;
            const int r_first = (int)((3u - c0) & 3u);
            if (!dense) {
                if (n_term == 0) {
                    for (uint32_t r = 0; r < n_nl; r++) {
                        const uint32_t role = (c0 + r) & 3;
                        const long long pp = (long long)(g0 + st->nlp[r]);
                        total += (role == 0 || role == 3) ? -pp : pp;
                    }
                } else {
                    for (int r = 0; r <= r_first; r++) {
                        const uint32_t role = (c0 + r) & 3;
                        const long long pp = (long long)(g0 + st->nlp[r]);
                        head += (role == 0 || role == 3) ? -pp : pp;
                    }
                    total = head;
                    for (uint32_t r = (uint32_t)r_first + 4u * (n_term - 1) + 1u; r < n_nl; r++) {
                        const uint32_t role = (c0 + r) & 3;
                        const long long pp = (long long)(g0 + st->nlp[r]);
                        total += (role == 0 || role == 3) ? -pp : pp;
                    }
                }
            }
            P.sum_total[t] = total;
            P.sum_head[t] = head;
            P.has_term[t] = n_term > 0 ? 1 : 0;
            P.nl_count[t] = n_nl;
            P.phase_used[t] = (uint8_t)c0;
            // end-of-file condition of canonical input (the line count is checked by the follow-up kernel)
            if (t + 1 == P.n_tiles && tile[tile_len - 1] != '\n') set_fallback(P.res, 9);
            mbar_arrive(&S->parsed[s]);  // hand the stage to the copy group
        }
    }
    if (gt == 0 && my_reads_out) atomicAdd(&P.res->reads_out, my_reads_out);
}

// P4 for one output stream.  The stream's bytes of this tile form ONE contiguous global range
// [base, base + total): run r contributes len_r bytes at offset off_r iff it belongs to the stream.
//   WRITTEN = true : stream out_w, runs with flag set,   off_r = K_r (kept prefix)
//   WRITTEN = false: stream out_o, runs with flag clear, off_r = S_r - K_r
template <bool WRITTEN>
__device__ __forceinline__ void emit_stream(Stage *st, GroupSmem *G, const uint8_t *tile, uint8_t *base, uint32_t total,
                                            uint32_t n_starts, bool carry, uint32_t head_kept, int gt, int bar) {
    const int lane = gt & 31, warp = gt >> 5;
    constexpr int NW = GT / 32;
    if (total == 0) return;  // uniform
    const uintptr_t b0 = (uintptr_t)base;
    const uintptr_t A0 = (b0 + 15) & ~(uintptr_t)15;     // first aligned chunk
    const uintptr_t A1 = (b0 + total) & ~(uintptr_t)15;  // end of the last aligned chunk
    const uint32_t n_chunks = A1 > A0 ? (uint32_t)((A1 - A0) >> 4) : 0u;
    group_sync(bar);  // the previous stream's chunk loop has finished reading crun[]
    // clear this thread's 8 markers (blocked: chunks 8*gt .. 8*gt+7)
    *reinterpret_cast<uint4 *>(&st->crun[8 * gt]) = make_uint4(0, 0, 0, 0);
    group_sync(bar);
    // ---- owners: one thread per run writes the run's edge bytes and marks its first interior chunk
    for (uint32_t r = gt; r <= n_starts; r += GT) {
        const uint32_t s = st->runS[r], e = st->runS[r + 1];
        const uint32_t len = e - s;
        const bool fl = r ? (st->runF[r] != 0) : carry;
        if (len == 0 || fl != WRITTEN) continue;
        const uint32_t K = r ? head_kept + st->runK[r] : 0u;
        const uint32_t off = WRITTEN ? K : s - K;
        const uintptr_t a = b0 + off, b = a + len;
        const uintptr_t a16 = (a + 15) & ~(uintptr_t)15, b16 = b & ~(uintptr_t)15;
        const uint8_t *src = tile + s;
        uint8_t *dst = base + off;
        if (a16 < b16) {
            st->crun[(a16 - A0) >> 4] = (uint16_t)(r + 1);
            const uint32_t hn = (uint32_t)(a16 - a), tn = (uint32_t)(b - b16);
            for (uint32_t i = 0; i < hn; i++) dst[i] = src[i];
            for (uint32_t i = len - tn; i < len; i++) dst[i] = src[i];
        } else {
            for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
        }
    }
    group_sync(bar);
    // ---- propagate the markers: crun[c] = last marker at or before c (group-wide max-scan)
    {
        uint4 mk = *reinterpret_cast<const uint4 *>(&st->crun[8 * gt]);
        uint32_t v[8] = {mk.x & 0xFFFF, mk.x >> 16, mk.y & 0xFFFF, mk.y >> 16,
                         mk.z & 0xFFFF, mk.z >> 16, mk.w & 0xFFFF, mk.w >> 16};
#pragma unroll
        for (int i = 1; i < 8; i++) v[i] = max(v[i], v[i - 1]);
        uint32_t inc = v[7];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc = max(inc, o);
        }
        if (lane == 31) G->wmax[warp] = inc;
        uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        group_sync(bar);
#pragma unroll
        for (int w = 0; w < NW; w++)
            if (w < warp) excl = max(excl, G->wmax[w]);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = max(v[i], excl);
        mk.x = v[0] | (v[1] << 16);
        mk.y = v[2] | (v[3] << 16);
        mk.z = v[4] | (v[5] << 16);
        mk.w = v[6] | (v[7] << 16);
        *reinterpret_cast<uint4 *>(&st->crun[8 * gt]) = mk;
    }
    group_sync(bar);
    // ---- interior chunks: thread-per-chunk, interleaved so that a warp stores 512 contiguous bytes
    for (uint32_t c = gt; c < n_chunks; c += GT) {
        const uint32_t r1 = st->crun[c];
        if (r1 == 0) continue;
        const uint32_t r = r1 - 1;
        const uint32_t s = st->runS[r], e = st->runS[r + 1];
        const uint32_t K = r ? head_kept + st->runK[r] : 0u;
        const uint32_t off = WRITTEN ? K : s - K;
        const uint32_t x = (uint32_t)(A0 - b0) + (c << 4);  // stream offset of this chunk
        if (x + 16 > off + (e - s)) continue;                // the chunk straddles the run's end: edge bytes
        const uint4 o = lds_unaligned16(tile + s + (x - off));
        st_global_v4(reinterpret_cast<void *>(A0 + ((uintptr_t)c << 4)), o);
    }
}

// ------------------------------------------------------------------ copy group (look-back #2, P4)
__device__ __forceinline__ void copy_loop(const FusedParams &P, FusedSmem *S, const int gt) {
    constexpr int BAR = 1 + NPARSE;
    GroupSmem *G = &S->g[NPARSE];
    for (uint32_t it = 0;; it++) {
        const uint32_t s = it % NS;
        Stage *st = &S->st[s];
        mbar_wait(&S->parsed[s], (it / NS) & 1);
        const uint64_t t = st->tile;
        if (t == ~0ull) break;
        const uint32_t n_starts = st->n_starts, head_len = st->head_len, rest_total = st->rest_total,
                       tile_len = st->tile_len;
        const uint8_t *tile = st->buf + PRE;
        const uint64_t g0 = t * (uint64_t)TILE;
        uint64_t kept_before;
        bool carry;
        const bool last_flag = n_starts ? (st->runF[n_starts] != 0) : false;
        lookback_kept(P.desc2, t, n_starts > 0, last_flag, head_len, rest_total, G, gt, BAR, &kept_before, &carry);
        const uint32_t head_kept = carry ? head_len : 0u;
        const uint32_t tile_kept = head_kept + rest_total;
        emit_stream<true>(st, G, tile, P.out_w + kept_before, tile_kept, n_starts, carry, head_kept, gt, BAR);
        if (P.out_o)
            emit_stream<false>(st, G, tile, P.out_o + (g0 - kept_before), tile_len - tile_kept, n_starts, carry,
                               head_kept, gt, BAR);
        if (t + 1 == P.n_tiles && gt == 0) P.res->kept_total = kept_before + tile_kept;
        group_sync(BAR);  // every read of the stage is done
        if (gt == 0) mbar_arrive(&S->empty[s]);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) fastq_fused_kernel(FusedParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FusedSmem *S = reinterpret_cast<FusedSmem *>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&S->full[s], 1);
            mbar_init(&S->parsed[s], 1);
            mbar_init(&S->empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) producer_loop(P, S);
    else if (tid < 32 + NPARSE * GT) parse_loop(P, S, (tid - 32) % GT, (tid - 32) / GT);
    else copy_loop(P, S, tid - 32 - NPARSE * GT);
}

// exact verification of what the tiles assumed: (1) the speculated line phase of every tile against the
// true prefix of newline counts, (2) the signed newline-position sums vanish at every record end
// (seq and qual lengths agree for records that straddle tiles), (3) the file's line count is a multiple of 4
__global__ void fused_verify_kernel(const uint64_t *sum_prefix, const long long *sum_head, const uint8_t *has_term,
                                    const uint64_t *nl_prefix, const uint32_t *nl_count, const uint8_t *phase_used,
                                    uint64_t n_tiles, FusedResult *res) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    if ((uint32_t)(nl_prefix[t] & 3) != (uint32_t)phase_used[t]) set_fallback(res, 11);
    if (has_term[t] && (long long)sum_prefix[t] + sum_head[t] != 0) set_fallback(res, 10);
    if (t + 1 == n_tiles) {
        const uint64_t lines = nl_prefix[t] + nl_count[t];
        if (lines & 3) set_fallback(res, 9);
        res->reads_in = lines >> 2;
    }
}

sgpu_status clean_fused(sgpu_ctx *c, const sgpu_idset *set, const uint8_t *d_in, size_t n_in, int reverse,
                        uint8_t *d_out_w, size_t cap_w, size_t *n_w, uint8_t *d_out_o, size_t cap_o, size_t *n_o,
                        sgpu_counts *counts, int *used) {
    *used = 0;
    // the fused kernel writes a byte partition of the input: both outputs must be able to hold it
    if (n_in == 0 || cap_w < n_in || (d_out_o && cap_o < n_in)) return SGPU_OK;
    cudaStream_t st = c->stream;
    static bool attr_done[64] = {false};
    const size_t smem = sizeof(FusedSmem);
    if (!attr_done[c->device & 63]) {
        SGPU_CUDA(cudaFuncSetAttribute(fastq_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[c->device & 63] = true;
    }
    uint64_t n_tiles = ceil_div(n_in, (size_t)TILE);
    DevBuf<unsigned long long> desc;
    DevBuf<long long> sums;
    DevBuf<uint64_t> prefix;
    DevBuf<uint32_t> nl_count;
    DevBuf<uint8_t> bytes;
    DevBuf<FusedResult> res;
    SGPU_TRY(desc.alloc(2 * n_tiles, st));
    SGPU_TRY(sums.alloc(2 * n_tiles, st));
    SGPU_TRY(prefix.alloc(2 * n_tiles, st));
    SGPU_TRY(nl_count.alloc(n_tiles, st));
    SGPU_TRY(bytes.alloc(2 * n_tiles, st));
    SGPU_TRY(res.alloc(1, st));
    SGPU_CUDA(cudaMemsetAsync(desc.p, 0, 2 * n_tiles * 8, st));
    SGPU_CUDA(cudaMemsetAsync(res.p, 0, sizeof(FusedResult), st));
    FusedParams P;
    P.in = d_in;
    P.n_in = n_in;
    P.n_tiles = n_tiles;
    P.out_w = d_out_w;
    P.out_o = d_out_o;
    P.reverse = reverse;
    P.set = view_of(set);
    P.desc1 = desc.p;
    P.desc2 = desc.p + n_tiles;
    P.sum_total = sums.p;
    P.sum_head = sums.p + n_tiles;
    P.nl_count = nl_count.p;
    P.has_term = bytes.p;
    P.phase_used = bytes.p + n_tiles;
    P.res = res.p;
    uint64_t grid = (uint64_t)c->sm_count;  // one persistent CTA per SM
    if (grid > n_tiles) grid = n_tiles;
    if (c->profiling) {
        if (c->prof_used == c->prof_events.size()) {
            cudaEvent_t a, b;
            SGPU_CUDA(cudaEventCreate(&a));
            SGPU_CUDA(cudaEventCreate(&b));
            c->prof_events.emplace_back(a, b);
        }
        SGPU_CUDA(cudaEventRecord(c->prof_events[c->prof_used].first, st));
    }
    fastq_fused_kernel<<<(unsigned)grid, NTHREADS, smem, st>>>(P);
    SGPU_LAUNCH(c);
    if (c->profiling) SGPU_CUDA(cudaEventRecord(c->prof_events[c->prof_used++].second, st));
    SGPU_TRY(exclusive_scan_u64(c, (const uint64_t *)P.sum_total, prefix.p, n_tiles, nullptr));
    SGPU_TRY(exclusive_scan_u32_to_u64(c, nl_count.p, prefix.p + n_tiles, n_tiles, nullptr));
    fused_verify_kernel<<<(unsigned)ceil_div(n_tiles, 256), 256, 0, st>>>(prefix.p, P.sum_head, P.has_term,
                                                                         prefix.p + n_tiles, nl_count.p, P.phase_used,
                                                                         n_tiles, res.p);
    SGPU_LAUNCH(c);
    SGPU_CUDA(cudaGetLastError());
    FusedResult h;
    SGPU_TRY(read_u64s(c, res.p, (uint64_t *)&h, sizeof(FusedResult) / 8));
    if (h.fallback) {  // *used stays 0: the general path decides (and reports errors)
        if (getenv("SGPU_DEBUG")) fprintf(stderr, "[sgpu] fused kernel fell back, reason %llu\n", h.reason);
        return SGPU_OK;
    }
    *used = 1;
    if (c->profiling) c->prof_alg_bytes += n_in + h.kept_total + (d_out_o ? n_in - h.kept_total : 0);
    *n_w = (size_t)h.kept_total;
    if (n_o) *n_o = d_out_o ? (size_t)(n_in - h.kept_total) : 0;
    counts->reads_in = h.reads_in;
    counts->reads_out = h.reads_out;
    counts->crlf = 0;
    counts->path = 1;
    return SGPU_OK;
}

}  // namespace sgpu

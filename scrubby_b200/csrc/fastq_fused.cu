// fastq_fused.cu -- single-pass fused parse -> probe -> compact kernel for CANONICAL FASTQ.
//
// Replaces the whole of FastqCleaner::clean_reads (cleaner.rs:731-760) -- needletail framing,
// get_id (utils.rs:91-103), `read_ids.contains` and `record.write` -- with ONE kernel that
// reads every input byte once and writes every output byte once.
//
// Canonical input = LF line endings, bare "+" separator lines, a final newline, ASCII
// bytes, well-formed records.  For such input the reference's output is a byte partition
// of the input, so the work is an order-preserving stream compaction of
// records.  Anything else (CRLF, "+id" separators, missing final newline, non-ASCII bytes,
// parse errors, pathological line density) raises the device `fallback` flag and the caller
// re-runs the always-exact general path (fastq_general.cu); nothing is approximated.
//
// Structure: persistent CTAs of NT threads, several per SM, each looping over dynamically
// ticketed TILE-byte tiles (tickets are taken in order, so every tile a look-back waits on is
// owned by a running CTA).  All threads of a CTA run every phase of a tile; latency (TMA load,
// set probes, look-back) is hidden by the other CTAs resident on the SM.
//   load : one thread issues a TMA bulk copy (cp.async.bulk + mbarrier) of the tile + 16 B
//          pre-halo + post-halo into shared memory and an L2 prefetch of the tile this CTA is
//          likely to take next;
//   P1   : 16-byte vector loads -> exact '\n' bit masks (SWAR + dp4a gather), counts, scan,
//          newline positions compacted in order;
//   phase: line number mod 4 SPECULATED from the first "\n+\n" in the tile and published at once
//          (a tile without one does a real decoupled look-back over 2-bit phases);
//   P2   : one thread per newline: CR / "+\n" checks by role, record starts;
//   P3   : one thread per record start: '@', id token -> slot image straight from a 16-byte
//          window, exact probe of the id set, seq/qual length check, block scan of kept bytes,
//          tile aggregate published;
//   scan : decoupled look-back by one warp over (kept bytes, state of the record that straddles
//          the tile edge), 256 descriptors per round, combined as a 3-state transducer;
//   P4   : destination-driven copy: every 16-byte aligned output chunk is owned by one thread
//          (coalesced 16-byte stores), its source run found through a marker array + max-scan,
//          the source re-aligned with funnel shifts.
// The speculated phases and the cross-tile seq/qual length sums are verified exactly by a tiny
// follow-up kernel over per-tile metadata; any mismatch is a fallback, never a wrong answer.
// Records may straddle any number of tiles (ONT reads); only the id token must lie within the
// post-halo of the tile where the record starts.
//
// Shards (multi-GPU): the buffer starts `lead` (< 16) bytes before the first owned
// record start and holds a halo after `own_len`; records that start after own_len are seen but
// belong to the next shard (state "none": written to neither stream).
#include <stdlib.h>

#include "fastq_records.cuh"

namespace sgpu {

#ifndef SGPU_FUSED_NT
#define SGPU_FUSED_NT 256
#endif
#ifndef SGPU_FUSED_FC
#define SGPU_FUSED_FC 4
#endif
#ifndef SGPU_FUSED_CTAS
#define SGPU_FUSED_CTAS 4
#endif
constexpr int NT = SGPU_FUSED_NT;       // worker threads per CTA
constexpr int NW = NT / 32;             // worker warps per CTA
constexpr int NTHREADS = NT + 32;       // + the scan warp
constexpr int FC = SGPU_FUSED_FC;       // 16-byte chunks per worker thread
constexpr int PW = (FC + 3) / 4;        // 64-bit words of packed per-round counts
constexpr int TILE = NT * FC * 16;      // 16 KiB
constexpr int PRE = 16;                 // pre-halo (previous 16 bytes)
constexpr int HALO = 1024;              // post-halo
constexpr int BUF = PRE + TILE + HALO;  // bytes of the tile buffer
constexpr int RMAX = 256;               // record starts per tile
constexpr int LMAX = 4 * RMAX + 8;      // newline list capacity per tile
constexpr int LBW = 16;                 // look-back descriptors per lane and round (window 32 * LBW tiles)
constexpr int CTAS_PER_SM = SGPU_FUSED_CTAS;  // resident CTAs the kernel is sized for (registers, shared memory)
static_assert(TILE <= 32768 && FC % 2 == 0 && FC <= 8, "tile offsets are 16-bit; markers are handled in pairs");

constexpr uint64_t ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = 3ull << 62;
// run / carry states
constexpr uint32_t F_OTHER = 0, F_KEPT = 1, F_NONE = 2;

struct FusedResult {
    unsigned long long fallback;    // != 0: input is not canonical, use the general path
    unsigned long long kept_total;  // bytes written to out_w
    unsigned long long reads_in, reads_out;
    unsigned long long ticket;      // dynamic tile counter
    unsigned long long reason;      // first fallback reason (diagnostics)
    unsigned long long owned_end;   // end of the last owned record (shards); ~0 when no foreign record was seen
    unsigned long long pad;
};

struct FusedParams {
    const uint8_t *in;
    uint64_t n_in;
    uint64_t n_tiles;
    uint64_t own_len;  // records that start after own_len belong to the next shard (ignored when is_last)
    uint32_t lead;     // bytes before the first record start (< 16; they belong to the previous shard)
    int is_last;       // the buffer ends at the end of the file
    uint8_t *out_w, *out_o;
    int reverse;
    IdSetView set;
    unsigned long long *desc1, *desc2;  // per tile look-back descriptors (zero initialised)
    long long *sum_total, *sum_head;    // per tile signed newline-position sums (length check)
    uint32_t *nl_count;                 // per tile newline count (phase verification)
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier over the worker warps only (id 1; id 0 is __syncthreads, which the scan warp never joins)
__device__ __forceinline__ void work_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// ------------------------------------------------------------------ byte classification
// flag byte (0x80 per matching byte) of the bytes of w equal to '\n'.  Exact for every byte value:
// low 7 bits zero after the xor  <=>  the add does not carry into bit 7; bit 7 of w itself must be clear.
__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {
    const uint32_t b = ((w ^ 0x0a0a0a0au) & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(b | w) & 0x80808080u;
}
// four words of 0x80 flags -> 16-bit mask in byte order (bit i <-> byte i): dp4a sums flag * weight
__device__ __forceinline__ uint32_t gather16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    const uint32_t lo = __dp4a(f0, 0x08040201u, __dp4a(f1, 0x80402010u, 0u));  // 128 * (bits 0..7)
    const uint32_t hi = __dp4a(f2, 0x08040201u, __dp4a(f3, 0x80402010u, 0u));  // 128 * (bits 8..15)
    return (lo >> 7) | (hi << 1);
}
__device__ __forceinline__ uint32_t nl_mask16_v2(uint4 v) {
    return gather16(nl_flags(v.x), nl_flags(v.y), nl_flags(v.z), nl_flags(v.w));
}
// 0x80 per byte <= 0x20 (ASCII input: every byte < 0x80, so the add cannot carry between bytes)
__device__ __forceinline__ uint32_t le20_flags(uint32_t w) { return ~(w + 0x5f5f5f5fu) & 0x80808080u; }

// ------------------------------------------------------------------ shared memory
// one tile buffer with everything the copy needs once the tile is parsed; a CTA owns two and parses tile
// n+1 BEFORE it looks back and copies tile n, so an aggregate is published a whole parse ahead of the
// look-backs that need it (predecessors are then normally ready and the look-back does not wait)
struct __align__(16) Stage {
    uint64_t full;         // mbarrier: tile load complete
    uint64_t agg_ready;    // mbarrier: workers -> scan warp, the tile's aggregate is published
    uint64_t scan_done;    // mbarrier: scan warp -> workers, kept_before / carry are valid
    uint64_t tile;         // tile index (>= n_tiles: no more tiles)
    uint64_t kept_before;  // look-back #2 result
    uint32_t carry;        // state of the record carried into the tile
    uint32_t n_starts, head_len, rest_total, tile_len;
    uint32_t last_flag;
    uint32_t none_pos;  // first record start of the tile that belongs to the next shard
    uint32_t none_cnt;
    uint16_t runS[RMAX + 4];  // run r starts at runS[r]; run 0 = carried-in head, run j+1 = record j; sentinel = tile_len
    uint8_t runF[RMAX + 4];   // F_OTHER / F_KEPT / F_NONE
    uint32_t runK[RMAX + 4];  // kept bytes of the records before run r (head excluded)
    // nlp[] (parse: local newline positions) and crun[] (copy: per destination chunk, 1 + index of
    // the run it lies inside) are never live at the same time
    union {
        __align__(16) uint16_t nlp[LMAX];
        __align__(16) uint16_t crun[TILE / 16];
    };
    __align__(16) uint8_t buf[BUF];
};

struct __align__(16) CtaSmem {
    uint32_t c0;           // newlines before the tile, mod 4
    uint32_t early;        // the previous tile's prefix is already there: copy it between the parse halves
    uint32_t pad[2];
    uint32_t warp_tot[NW], scan_tot[NW], wmax[NW];
    Stage st[2];
};

__device__ __forceinline__ void set_fallback(FusedResult *res, int reason) {
    if (atomicExch(&res->fallback, 1ull) == 0ull) res->reason = (unsigned long long)reason;
}

// CTA-wide exclusive scan of one u32 per thread; two barriers
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total, uint32_t *sm, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += x;
    }
    if (lane == 31) sm[warp] = inc;
    work_sync();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const uint32_t x = sm[w];
        if (w < warp) base += x;
        tot += x;
    }
    work_sync();
    *total = tot;
    return base + inc - v;
}

// ------------------------------------------------------------------ look-back #1 (rare): line phase
// exclusive prefix of the newline counts mod 4, by one warp, for tiles without a "\n+\n"
__device__ __forceinline__ uint32_t lookback_phase_warp(unsigned long long *desc, uint64_t t, uint32_t mine, int lane) {
    if (lane == 0) st_relaxed(desc + t, ST_AGG | (mine & 3));
    uint32_t acc = 0;
    int64_t base = (int64_t)t - 1;
    while (true) {
        const int64_t idx = base - lane;
        unsigned long long d = ST_INC;  // virtual tiles before the buffer: inclusive prefix 0
        if (idx >= 0) {
            while (((d = ld_relaxed(desc + idx)) & ST_MASK) == 0) __nanosleep(64);
        }
        const unsigned b = __ballot_sync(0xffffffffu, (d & ST_MASK) == ST_INC);
        const int first = b ? __ffs(b) - 1 : 32;
        acc += __reduce_add_sync(0xffffffffu, lane <= first ? (uint32_t)(d & 3) : 0u);
        if (b) break;
        base -= 32;
    }
    if (lane == 0) st_relaxed(desc + t, ST_INC | ((acc + mine) & 3));
    return acc & 3;
}

// ------------------------------------------------------------------ look-back #2: kept bytes + carried state
// aggregate:  [61] has_start  [60:59] state of the last record  [58:30] head_len  [29:0] rest_kept
// inclusive:  [60:59] state carried out of the tile              [58:0] kept bytes up to and including the tile
// A tile is a transducer on the carried state c: it keeps (c == KEPT ? head_len : 0) + rest bytes and
// carries out (has_start ? last : c).  A run of tiles composes into {P: bytes kept iff c == KEPT, K: bytes
// kept regardless, has, out}; composition is associative, so a warp reduces 256 descriptors per round.
struct Comp {
    uint32_t P, K, has, out;
};
__device__ __forceinline__ Comp comp_identity() { return Comp{0u, 0u, 0u, 0u}; }
__device__ __forceinline__ Comp compose(const Comp A /*earlier*/, const Comp B /*later*/) {
    Comp R;
    if (!A.has) {
        R.P = A.P + B.P;
        R.K = B.K;
        R.has = B.has;
        R.out = B.out;
    } else {
        R.P = A.P;
        R.K = A.K + B.K + (A.out == F_KEPT ? B.P : 0u);
        R.has = 1u;
        R.out = B.has ? B.out : A.out;
    }
    return R;
}
__device__ __forceinline__ uint64_t comp_pack(const Comp c) {
    return (uint64_t)c.P | ((uint64_t)c.K << 26) | ((uint64_t)c.has << 52) | ((uint64_t)c.out << 53);
}
__device__ __forceinline__ Comp comp_unpack(uint64_t v) {
    Comp c;
    c.P = (uint32_t)(v & 0x3FFFFFFu);
    c.K = (uint32_t)((v >> 26) & 0x3FFFFFFu);
    c.has = (uint32_t)((v >> 52) & 1u);
    c.out = (uint32_t)((v >> 53) & 3u);
    return c;
}

__device__ __forceinline__ void lookback_kept_warp(unsigned long long *desc, uint64_t t, bool has_start,
                                                   uint32_t last_flag, uint32_t head_len, uint32_t rest, int lane,
                                                   uint64_t *kept_before, uint32_t *carry) {
    Comp acc_all = comp_identity();  // composite of every tile visited so far (nearer rounds are later)
    uint64_t inc_total = 0;
    int64_t base = (int64_t)t - 1 - (int64_t)lane * LBW;  // nearest descriptor of this lane
    while (true) {
        unsigned long long d[LBW];
        int fi, zi;  // first inclusive / first not-ready descriptor of this lane
        unsigned binc;
        int L;
#pragma unroll
        for (int k = 0; k < LBW; k++) {
            const int64_t idx = base - k;
            // virtual tiles before the buffer: nothing kept, nothing carried in
            d[k] = idx >= 0 ? ld_relaxed(desc + idx) : (ST_INC | ((uint64_t)F_NONE << 59));
        }
        while (true) {
            fi = LBW;
            zi = LBW;
#pragma unroll
            for (int k = LBW - 1; k >= 0; k--) {
                const uint64_t s = d[k] & ST_MASK;
                if (s == ST_INC) fi = k;
                if (s == 0) zi = k;
            }
            binc = __ballot_sync(0xffffffffu, fi < LBW);
            L = binc ? __ffs(binc) - 1 : 32;
            const int rk = lane < L ? LBW : (lane == L ? fi : 0);  // descriptors of this lane that matter
            const bool need = zi < rk;
            if (!__any_sync(0xffffffffu, need)) break;
            if (need) {
                // only the missing descriptors are polled again (an idle scan warp must not flood L2)
                __nanosleep(200);
#pragma unroll
                for (int k = 0; k < LBW; k++)
                    if (k < rk && (d[k] & ST_MASK) == 0) d[k] = ld_relaxed(desc + (base - k));
            }
        }
        const int rk = lane < L ? LBW : (lane == L ? fi : 0);
        Comp acc = comp_identity();
        uint64_t my_total = 0;
        if (lane == L) {
            uint64_t di = d[0];
#pragma unroll
            for (int k = 1; k < LBW; k++)
                if (k == fi) di = d[k];
            acc = Comp{0u, 0u, 1u, (uint32_t)((di >> 59) & 3u)};
            my_total = di & ((1ull << 59) - 1);
        }
#pragma unroll
        for (int k = LBW - 1; k >= 0; k--) {  // file order: farthest first
            if (k < rk) {
                const uint64_t x = d[k];
                const Comp a = Comp{(uint32_t)((x >> 30) & 0x1FFFFFFFu), (uint32_t)(x & 0x3FFFFFFFu),
                                    (uint32_t)((x >> 61) & 1u), (uint32_t)((x >> 59) & 3u)};
                acc = compose(acc, a);
            }
        }
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {  // lane l+s holds earlier tiles than lane l
            const uint64_t o = __shfl_down_sync(0xffffffffu, comp_pack(acc), s);
            if (lane + s < 32) acc = compose(comp_unpack(o), acc);
        }
        const uint64_t w = __shfl_sync(0xffffffffu, comp_pack(acc), 0);
        acc_all = compose(comp_unpack(w), acc_all);
        if (L < 32) {
            inc_total = __shfl_sync(0xffffffffu, my_total, L);
            break;
        }
        base -= 32 * LBW;
    }
    // acc_all starts with the inclusive descriptor (has == 1, P == 0)
    *kept_before = inc_total + acc_all.K;
    *carry = acc_all.out;
    const uint64_t incl = *kept_before + (acc_all.out == F_KEPT ? head_len : 0u) + rest;
    const uint32_t out_flag = has_start ? last_flag : acc_all.out;
    if (lane == 0) st_relaxed(desc + t, ST_INC | ((uint64_t)out_flag << 59) | incl);
}

// ------------------------------------------------------------------ P4 for one output stream
// The stream's bytes of this tile form ONE contiguous global range [base, base + total): run r
// contributes len_r bytes at offset off_r iff it belongs to the stream.
//   WRITTEN = true : stream out_w, runs in state KEPT,  off_r = K_r (kept prefix)
//   WRITTEN = false: stream out_o, runs in state OTHER, off_r = S_r - K_r - none_prefix
template <bool WRITTEN>
__device__ __forceinline__ void emit_stream(Stage *S, uint32_t *wmax, const uint8_t *tile, uint8_t *base,
                                            uint32_t total, uint32_t n_starts, uint32_t carry, uint32_t head_kept,
                                            uint32_t none_prefix, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t WANT = WRITTEN ? F_KEPT : F_OTHER;
    if (total == 0) return;  // uniform
    const uintptr_t b0 = (uintptr_t)base;
    const uintptr_t A0 = (b0 + 15) & ~(uintptr_t)15;     // first aligned chunk
    const uintptr_t A1 = (b0 + total) & ~(uintptr_t)15;  // end of the last aligned chunk
    const uint32_t n_chunks = A1 > A0 ? (uint32_t)((A1 - A0) >> 4) : 0u;
    work_sync();  // the previous user of crun[] / nlp[] is done
    // clear this thread's FC markers (blocked: chunks FC*tid .. FC*tid+FC-1)
    uint32_t *const my_marks = reinterpret_cast<uint32_t *>(&S->crun[FC * tid]);
#pragma unroll
    for (int i = 0; i < FC / 2; i++) my_marks[i] = 0;
    work_sync();
    // ---- owners: one thread per run writes the run's edge bytes and marks its first interior chunk
    for (uint32_t r = tid; r <= n_starts; r += NT) {
        const uint32_t s = S->runS[r], e = S->runS[r + 1];
        const uint32_t len = e - s;
        const uint32_t fl = r ? (uint32_t)S->runF[r] : carry;
        if (len == 0 || fl != WANT) continue;
        const uint32_t K = r ? head_kept + S->runK[r] : 0u;
        const uint32_t off = WRITTEN ? K : s - K - none_prefix;
        const uintptr_t a = b0 + off, b = a + len;
        const uintptr_t a16 = (a + 15) & ~(uintptr_t)15, b16 = b & ~(uintptr_t)15;
        const uint8_t *src = tile + s;
        uint8_t *dst = base + off;
        if (a16 < b16) {
            S->crun[(a16 - A0) >> 4] = (uint16_t)(r + 1);
            const uint32_t hn = (uint32_t)(a16 - a), tn = (uint32_t)(b - b16);
            for (uint32_t i = 0; i < hn; i++) dst[i] = src[i];
            for (uint32_t i = len - tn; i < len; i++) dst[i] = src[i];
        } else {
            for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
        }
    }
    work_sync();
    // ---- propagate the markers: crun[c] = last marker at or before c (max-scan over the workers)
    {
        uint32_t v[FC];
#pragma unroll
        for (int i = 0; i < FC / 2; i++) {
            const uint32_t w = my_marks[i];
            v[2 * i] = w & 0xFFFF;
            v[2 * i + 1] = w >> 16;
        }
#pragma unroll
        for (int i = 1; i < FC; i++) v[i] = max(v[i], v[i - 1]);
        uint32_t inc = v[FC - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc = max(inc, o);
        }
        if (lane == 31) wmax[warp] = inc;
        uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        work_sync();
#pragma unroll
        for (int w = 0; w < NW; w++)
            if (w < warp) excl = max(excl, wmax[w]);
#pragma unroll
        for (int i = 0; i < FC; i++) v[i] = max(v[i], excl);
#pragma unroll
        for (int i = 0; i < FC / 2; i++) my_marks[i] = v[2 * i] | (v[2 * i + 1] << 16);
    }
    work_sync();
    // ---- interior chunks: thread per chunk, interleaved so that a warp stores 512 contiguous bytes
    const uint32_t tile_s = smem_u32(tile);
    for (uint32_t c = tid; c < n_chunks; c += NT) {
        const uint32_t r1 = S->crun[c];
        if (r1 == 0) continue;
        const uint32_t r = r1 - 1;
        const uint32_t s = S->runS[r], e = S->runS[r + 1];
        const uint32_t K = r ? head_kept + S->runK[r] : 0u;
        const uint32_t off = WRITTEN ? K : s - K - none_prefix;
        const uint32_t x = (uint32_t)(A0 - b0) + (c << 4);  // stream offset of this chunk
        if (x + 16 > off + (e - s)) continue;                // the chunk straddles the run's end: edge bytes
        // 16 source bytes at an arbitrary address: five aligned words + funnel shifts
        const uint32_t sa = tile_s + s + (x - off);
        const uint32_t wa = sa & ~3u, sh = (sa & 3u) * 8u;
        const uint32_t w0 = lds_u32(wa), w1 = lds_u32(wa + 4), w2 = lds_u32(wa + 8), w3 = lds_u32(wa + 12),
                       w4 = lds_u32(wa + 16);
        uint4 o;
        o.x = __funnelshift_r(w0, w1, sh);
        o.y = __funnelshift_r(w1, w2, sh);
        o.z = __funnelshift_r(w2, w3, sh);
        o.w = __funnelshift_r(w3, w4, sh);
        st_global_v4(reinterpret_cast<void *>(A0 + ((uintptr_t)c << 4)), o);
    }
}

// ------------------------------------------------------------------ tile load (one thread)
__device__ __forceinline__ void take_ticket_and_load(const FusedParams &P, Stage *st) {
    const unsigned long long t = atomicAdd(&P.res->ticket, 1ull);
    st->tile = t;
    st->none_pos = 0xFFFFFFFFu;
    st->none_cnt = 0;
    if (t >= P.n_tiles) return;
    // bytes [t*TILE - PRE, t*TILE + TILE + HALO) clipped to the buffer, rounded up to 16
    const uint64_t g0 = t * (uint64_t)TILE;
    const uint64_t src0 = t ? g0 - PRE : 0;
    uint64_t end = g0 + TILE + HALO;
    if (end > P.n_in) end = P.n_in;
    const uint32_t bytes = (uint32_t)(((end - src0) + 15) & ~15ull);
    // the buffer was read through the generic proxy; order those reads before the async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&st->full, bytes);
    bulk_g2s(st->buf + (t ? 0 : PRE), P.in + src0, bytes, &st->full);
    // the tile this CTA will most likely take next: pull it into L2 now
    const uint64_t tn = t + gridDim.x;
    if (tn < P.n_tiles) {
        const uint64_t p0 = tn * (uint64_t)TILE;
        uint64_t pe = p0 + TILE;
        if (pe > P.n_in) pe = P.n_in;
        const uint32_t pb = (uint32_t)((pe - p0) & ~15ull);
        if (pb) bulk_prefetch_l2(P.in + p0, pb);
    }
}

// ------------------------------------------------------------------ per-record helpers (P3)
// id token of the record that starts at tile[sp] ('@'): fast path = 16-byte window at the token start ->
// length and inline slot image in registers.  mode 1: (lo, hi) is the image; mode 2: take the general scan.
struct RecPrep {
    uint32_t mode;
    uint64_t lo, hi;
};
__device__ __forceinline__ RecPrep record_prepare(uint32_t tile_s, uint32_t sp, uint32_t avail) {
    RecPrep R;
    R.mode = 2;
    R.lo = R.hi = 0;
    const uint32_t i = sp + 1;
    if (i + 16 > avail) return R;
    const uint32_t sa = tile_s + i;
    const uint32_t wa = sa & ~3u, sh = (sa & 3u) * 8u;
    const uint32_t x0 = lds_u32(wa), x1 = lds_u32(wa + 4), x2 = lds_u32(wa + 8), x3 = lds_u32(wa + 12),
                   x4 = lds_u32(wa + 16);
    const uint32_t w0 = __funnelshift_r(x0, x1, sh), w1 = __funnelshift_r(x1, x2, sh),
                   w2 = __funnelshift_r(x2, x3, sh), w3 = __funnelshift_r(x3, x4, sh);
    const uint32_t stop = gather16(le20_flags(w0), le20_flags(w1), le20_flags(w2), le20_flags(w3));
    const uint32_t len = stop ? (uint32_t)(__ffs(stop) - 1) : 16u;
    // the stop byte must be real whitespace, the token non-empty and short enough to be inline
    const uint32_t cw = len < 4 ? w0 : len < 8 ? w1 : len < 12 ? w2 : w3;
    const uint32_t cb = (cw >> (8 * (len & 3))) & 0xFFu;
    if (!(len >= 1 && len <= IDSET_INLINE_MAX && (cb == 0x20u || (cb - 9u) < 5u))) return R;
    // image = (token << 8 | len) cut to len + 1 bytes
    uint32_t s0 = (w0 << 8) | len, s1 = __funnelshift_l(w0, w1, 8), s2 = __funnelshift_l(w1, w2, 8),
             s3 = __funnelshift_l(w2, w3, 8);
    const uint32_t nb = len + 1;  // 2..16
    const uint32_t full = nb >> 2, part = (nb & 3u) * 8u;
    const uint32_t pm = (1u << part) - 1u;  // part == 0 -> 0
    s0 = full > 0 ? s0 : s0 & pm;
    s1 = full > 1 ? s1 : (full == 1 ? s1 & pm : 0u);
    s2 = full > 2 ? s2 : (full == 2 ? s2 & pm : 0u);
    s3 = full > 3 ? s3 : (full == 3 ? s3 & pm : 0u);
    R.lo = (uint64_t)s0 | ((uint64_t)s1 << 32);
    R.hi = (uint64_t)s2 | ((uint64_t)s3 << 32);
    R.mode = 1;
    return R;
}
__device__ __forceinline__ uint64_t inline_home(uint64_t lo, uint64_t hi) {
    return mix64(lo ^ mix64(hi + 0x9E3779B97F4A7C15ULL));
}
// general token scan + probe: skip leading blanks, run to the next blank / newline
__device__ __forceinline__ bool record_probe_slow(const IdSetView &set, const uint8_t *tile, uint32_t sp,
                                                  uint32_t avail, uint32_t *why) {
    uint32_t a = sp + 1;
    while (a < avail && is_ws_ascii(tile[a]) && tile[a] != '\n') a++;
    uint32_t q = a;
    while (q < avail && !is_ws_ascii(tile[q])) q++;
    if (q >= avail || q == a) {  // token past the halo, or empty id (error 9)
        *why = *why ? *why : 7u;
        return false;
    }
    return idset_contains(set, tile + a, q - a);
}

// what a thread carries from the first half of the parse (probe issued) to the second (probe consumed)
struct ParseState {
    uint32_t fb, c0, n_nl, n_starts, n_term;
    bool pos0_start, dense;
    uint32_t mode;  // 0: nothing pending; 1: inline probe in flight (`first` holds the home slot)
    uint32_t why;
    uint64_t lo, hi;
    Slot first;
};

// ------------------------------------------------------------------ parse: P1, phase, P2, P3, aggregate
// first half: everything up to the ISSUE of the set probes (one per record start)
__device__ __forceinline__ ParseState parse_a(const FusedParams &P, CtaSmem *C, Stage *S, uint32_t parity, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const uint64_t t = S->tile;
    uint8_t *buf = S->buf;
    const uint8_t *tile = buf + PRE;  // tile[-16 .. avail)
    if (warp == 0) {
        while (!mbar_try_wait(&S->full, parity)) {
        }
    }
    if (t == 0 && tid < PRE) buf[tid] = '\n';  // no predecessor: the pre-halo reads as a newline
    work_sync();
    while (!mbar_try_wait(&S->full, parity)) {  // passes at once: every thread observes the completed phase
    }

    const uint64_t g0 = t * (uint64_t)TILE;
    const uint32_t tile_len = (uint32_t)((P.n_in - g0) < (uint64_t)TILE ? (P.n_in - g0) : (uint64_t)TILE);
    const uint32_t avail =
        (uint32_t)((P.n_in - g0) < (uint64_t)(TILE + HALO) ? (P.n_in - g0) : (uint64_t)(TILE + HALO));
    const uint32_t lead_t = t == 0 ? P.lead : 0u;
    uint32_t fb = 0;  // this thread's fallback reason (0 = none)

    // ---- P1: newline masks and counts.  Warp w owns chunks [w*FC*32, (w+1)*FC*32); lane l takes
    //      chunk k*32 + l of them in round k (conflict-free 16-byte shared loads)
    uint32_t m[FC];
    uint32_t hi_or = 0;
    const uint32_t cbase = (uint32_t)warp * (FC * 32) + (uint32_t)lane;
    if (tile_len == (uint32_t)TILE) {
#pragma unroll
        for (int k = 0; k < FC; k++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(tile + (cbase + k * 32) * 16);
            m[k] = nl_mask16_v2(v);
            hi_or |= (v.x | v.y | v.z | v.w);
        }
    } else {
#pragma unroll
        for (int k = 0; k < FC; k++) {  // last tile: bytes past the end of the buffer are stale
            const uint32_t pos = (cbase + k * 32) * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (pos < tile_len) {
                v = *reinterpret_cast<const uint4 *>(tile + pos);
                const uint32_t valid = tile_len - pos;
                if (valid < 16) {
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
            m[k] = nl_mask16_v2(v);  // zero bytes are never newlines
            hi_or |= (v.x | v.y | v.z | v.w);
        }
    }
    if (lead_t && tid == 0) m[0] &= ~((1u << lead_t) - 1u);  // the previous shard's bytes
    uint64_t packed[PW];
#pragma unroll
    for (int q = 0; q < PW; q++) packed[q] = 0;
#pragma unroll
    for (int k = 0; k < FC; k++) packed[k >> 2] |= (uint64_t)__popc(m[k]) << (16 * (k & 3));
    if (hi_or & 0x80808080u) fb = 1;  // reason 1: non-ASCII byte, Unicode rules needed
    // inclusive warp scan of the per-round counts (four 16-bit fields per word)
    uint64_t inc[PW];
#pragma unroll
    for (int q = 0; q < PW; q++) inc[q] = packed[q];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int q = 0; q < PW; q++) {
            const uint64_t x = __shfl_up_sync(0xffffffffu, inc[q], d);
            if (lane >= d) inc[q] += x;
        }
    }
    uint32_t row_base[FC];
    uint32_t wtot = 0;
#pragma unroll
    for (int q = 0; q < PW; q++) {
        const uint64_t rt = __shfl_sync(0xffffffffu, inc[q], 31);
        const uint64_t ex = inc[q] - packed[q];  // exclusive within the round
#pragma unroll
        for (int k = 4 * q; k < 4 * q + 4 && k < FC; k++) {
            row_base[k] = wtot + (uint32_t)((ex >> (16 * (k & 3))) & 0xFFFF);
            wtot += (uint32_t)((rt >> (16 * (k & 3))) & 0xFFFF);
        }
    }
    if (lane == 0) C->warp_tot[warp] = wtot;
    work_sync();
    uint32_t n_nl = 0, wbase = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const uint32_t x = C->warp_tot[w];
        if (w < warp) wbase += x;
        n_nl += x;
    }
    const bool dense = n_nl > (uint32_t)LMAX;
    // ---- compact the newline positions in order
    if (!dense) {
#pragma unroll
        for (int k = 0; k < FC; k++) {
            uint32_t mm = m[k];
            uint32_t r = wbase + row_base[k];
            const uint32_t pos = (cbase + k * 32) * 16;
            while (mm) {
                S->nlp[r++] = (uint16_t)(pos + (uint32_t)(__ffs(mm) - 1));
                mm &= mm - 1;
            }
        }
    }
    work_sync();
    // ---- line phase (warp 0): the first newline followed by "+\n" ends a sequence line (role 1)
    if (warp == 0) {
        uint32_t c0 = 0;
        if (t == 0) {
            if (lane == 0) st_relaxed(P.desc1 + t, ST_INC | (n_nl & 3));
        } else {
            bool hit = false;
            if (!dense && (uint32_t)lane < n_nl) {
                const uint32_t p = S->nlp[lane];
                hit = p + 2 < avail && tile[p + 1] == '+' && tile[p + 2] == '\n';
            }
            const unsigned b = __ballot_sync(0xffffffffu, hit);
            if (b) {
                c0 = (1u - (uint32_t)(__ffs(b) - 1)) & 3u;  // role(first) = (c0 + first) & 3 == 1
                if (lane == 0) st_relaxed(P.desc1 + t, ST_INC | ((c0 + n_nl) & 3));
            } else {
                c0 = lookback_phase_warp(P.desc1, t, n_nl, lane);
            }
        }
        if (lane == 0) C->c0 = c0;
    }
    work_sync();
    const uint32_t c0 = C->c0;
    // the first byte of the tile starts a record iff 4k newlines precede it and the previous byte is one;
    // tile 0 starts with a record by construction (at `lead`)
    const bool pos0_start = t == 0 || ((c0 == 0) && tile[-1] == '\n');

    // ---- P2: one thread per newline: classify by role (line number mod 4)
    if (!dense) {
        for (uint32_t i = tid; i < n_nl; i += NT) {
            const uint32_t p = S->nlp[i];
            const uint32_t role = (c0 + i) & 3;
            if (tile[(int)p - 1] == '\r') fb = 3;  // CRLF: not canonical
            if (role == 1) {                       // end of the sequence line: "+\n" must follow
                if (p + 2 >= avail) fb = 4;
                else if (tile[p + 1] != '+' || tile[p + 2] != '\n') fb = 4;
            } else if (role == 3 && p + 1 < tile_len) {  // record j starts at p + 1 (run j + 1)
                const uint32_t j = ((c0 + i) >> 2) + (pos0_start ? 1u : 0u);
                if (j < (uint32_t)RMAX) S->runS[j + 1] = (uint16_t)(p + 1);
                else fb = 5;
            }
        }
    } else {
        fb = 2;
    }
    if (tid == 0) {
        S->runS[0] = 0;
        if (pos0_start) S->runS[1] = (uint16_t)lead_t;
    }
    // number of record starts inside the tile
    const uint32_t n_term = (c0 + n_nl) >> 2;
    uint32_t n_starts = n_term + (pos0_start ? 1u : 0u);
    if (n_term > 0 && !dense) {
        // the last terminating newline may sit on the tile's final byte: its record belongs to the next tile
        const uint32_t r_last = ((3u - c0) & 3u) + 4u * (n_term - 1);
        if ((uint32_t)S->nlp[r_last] + 1u >= tile_len) n_starts--;
    }
    if (dense || n_starts > (uint32_t)RMAX) n_starts = 0;  // (a fallback reason is already raised)
    work_sync();
    if (tid == 0) S->runS[n_starts + 1] = (uint16_t)tile_len;  // sentinel (TILE <= 32768 fits)
    work_sync();

    // ---- P3a: one thread per record start (first NT records): id token -> slot image, probe issued
    ParseState Z;
    Z.fb = fb;
    Z.c0 = c0;
    Z.n_nl = n_nl;
    Z.n_starts = n_starts;
    Z.n_term = n_term;
    Z.pos0_start = pos0_start;
    Z.dense = dense;
    Z.mode = 0;
    Z.why = 0;
    Z.lo = Z.hi = 0;
    Z.first.lo = Z.first.hi = 0;
    if ((uint32_t)tid < n_starts) {
        const uint32_t sp = S->runS[tid + 1];
        if (P.is_last || (g0 + sp <= P.own_len)) {
            Z.why = tile[sp] == '@' ? 0u : 6u;  // 6 '@', 7 id token, 8 seq/qual lengths
            const RecPrep R = record_prepare(smem_u32(tile), sp, avail);
            Z.mode = R.mode;
            Z.lo = R.lo;
            Z.hi = R.hi;
            if (R.mode == 1 && !Z.why && P.set.table != nullptr)
                Z.first = load_slot(P.set.table + (inline_home(R.lo, R.hi) & P.set.mask));
        }
    }
    return Z;
}

// second half: probes consumed, seq/qual length check, block scan of kept bytes, aggregate published.
// Adds (thread 0 only) the number of owned / kept records to *reads_in / *reads_out
__device__ __forceinline__ void parse_b(const FusedParams &P, CtaSmem *C, Stage *S, const ParseState &Z, int tid,
                                        unsigned long long *reads_in, unsigned long long *reads_out) {
    const uint64_t t = S->tile;
    const uint8_t *tile = S->buf + PRE;
    const uint64_t g0 = t * (uint64_t)TILE;
    const uint32_t tile_len = (uint32_t)((P.n_in - g0) < (uint64_t)TILE ? (P.n_in - g0) : (uint64_t)TILE);
    const uint32_t avail =
        (uint32_t)((P.n_in - g0) < (uint64_t)(TILE + HALO) ? (P.n_in - g0) : (uint64_t)(TILE + HALO));
    uint32_t fb = Z.fb;
    const uint32_t c0 = Z.c0, n_nl = Z.n_nl, n_starts = Z.n_starts, n_term = Z.n_term;
    const bool pos0_start = Z.pos0_start, dense = Z.dense;
    uint32_t rest_total = 0, kept_recs = 0;
    for (uint32_t jb = 0; jb < n_starts; jb += NT) {
        const uint32_t j = jb + tid;
        uint32_t my_len = 0, my_flag = F_OTHER;
        if (j < n_starts) {
            const uint32_t sp = S->runS[j + 1];
            const uint32_t e = S->runS[j + 2];
            my_len = e - sp;
            const bool owned = P.is_last || (g0 + sp <= P.own_len);
            if (!owned) {
                my_flag = F_NONE;
                atomicMin(&S->none_pos, sp);
                atomicAdd(&S->none_cnt, 1u);
            } else {
                uint32_t why, mode;
                uint64_t lo, hi;
                Slot first;
                if (jb == 0) {  // prepared (and probed) by parse_a
                    why = Z.why;
                    mode = Z.mode;
                    lo = Z.lo;
                    hi = Z.hi;
                    first = Z.first;
                } else {
                    why = tile[sp] == '@' ? 0u : 6u;
                    const RecPrep R = record_prepare(smem_u32(tile), sp, avail);
                    mode = R.mode;
                    lo = R.lo;
                    hi = R.hi;
                    first.lo = first.hi = 0;
                    if (mode == 1 && !why && P.set.table != nullptr)
                        first = load_slot(P.set.table + (inline_home(lo, hi) & P.set.mask));
                }
                bool hit = false;
                if (!why) {
                    if (mode == 1) {
                        // linear probing from the home slot (already loaded)
                        Slot sl = first;
                        if ((sl.lo | sl.hi) != 0) {
                            if (sl.lo == lo && sl.hi == hi) {
                                hit = true;
                            } else {
                                uint64_t idx = inline_home(lo, hi) & P.set.mask;
                                while (true) {
                                    idx = (idx + 1) & P.set.mask;
                                    sl = load_slot(P.set.table + idx);
                                    if ((sl.lo | sl.hi) == 0) break;
                                    if (sl.lo == lo && sl.hi == hi) {
                                        hit = true;
                                        break;
                                    }
                                }
                            }
                        }
                    } else {
                        hit = record_probe_slow(P.set, tile, sp, avail, &why);
                    }
                }
                if (!why) my_flag = (P.reverse ? hit : !hit) ? F_KEPT : F_OTHER;
                // seq/qual length equality for records whose four newlines are inside the tile
                const int r0 = pos0_start ? 4 * (int)j - 1 : (int)((3u - c0) & 3u) + 4 * (int)j;
                if (r0 + 4 < (int)n_nl) {
                    const int sgn = -(int)S->nlp[r0 + 1] + (int)S->nlp[r0 + 2] + (int)S->nlp[r0 + 3] -
                                    (int)S->nlp[r0 + 4];
                    if (sgn != 0) why = why ? why : 8u;
                }
                if (why) fb = why;
            }
            S->runF[j + 1] = (uint8_t)my_flag;
        }
        uint32_t round_total;
        const uint32_t koff =
            block_excl_scan(my_flag == F_KEPT ? (my_len | (1u << 16)) : 0u, &round_total, C->scan_tot, tid);
        if (j < n_starts) S->runK[j + 1] = rest_total + (koff & 0xFFFFu);
        rest_total += round_total & 0xFFFFu;
        kept_recs += round_total >> 16;
    }
    if (fb) set_fallback(P.res, (int)fb);
    work_sync();  // runS / runF / runK complete (also when the loop ran zero times)

    if (tid == 0) {
        const uint32_t head_len = S->runS[1];  // == tile_len when no record starts in the tile
        const uint32_t last_flag = n_starts ? (uint32_t)S->runF[n_starts] : F_OTHER;
        // aggregate for look-back #2, visible to every later tile from here on
        st_relaxed(P.desc2 + t, ST_AGG | (n_starts ? (1ull << 61) : 0ull) | ((uint64_t)last_flag << 59) |
                                    ((uint64_t)head_len << 30) | rest_total);
        S->n_starts = n_starts;
        S->head_len = head_len;
        S->rest_total = rest_total;
        S->tile_len = tile_len;
        S->last_flag = last_flag;
        *reads_in += n_starts - S->none_cnt;
        *reads_out += kept_recs;
        if (S->none_pos != 0xFFFFFFFFu) atomicMin(&P.res->owned_end, (unsigned long long)(g0 + S->none_pos));
        mbar_arrive(&S->agg_ready);  // the scan warp may look back for this tile now
    }
    if (tid == 32) {
        // signed newline-position sums: -p1 +p2 +p3 -p4 per record must vanish
        long long head = 0, total = 0;
        const int r_first = (int)((3u - c0) & 3u);
        if (!dense) {
            if (n_term == 0) {
                for (uint32_t r = 0; r < n_nl; r++) {
                    const uint32_t role = (c0 + r) & 3;
                    const long long pp = (long long)(g0 + S->nlp[r]);
                    total += (role == 0 || role == 3) ? -pp : pp;
                }
            } else {
                for (int r = 0; r <= r_first; r++) {
                    const uint32_t role = (c0 + r) & 3;
                    const long long pp = (long long)(g0 + S->nlp[r]);
                    head += (role == 0 || role == 3) ? -pp : pp;
                }
                total = head;
                for (uint32_t r = (uint32_t)r_first + 4u * (n_term - 1) + 1u; r < n_nl; r++) {
                    const uint32_t role = (c0 + r) & 3;
                    const long long pp = (long long)(g0 + S->nlp[r]);
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
        if (P.is_last && t + 1 == P.n_tiles && tile[tile_len - 1] != '\n') set_fallback(P.res, 9);
    }
}

// ------------------------------------------------------------------ P4 of a parsed tile (workers)
__device__ __forceinline__ void copy_tile(const FusedParams &P, CtaSmem *C, Stage *S, uint32_t parity, int tid) {
    const int warp = tid >> 5;
    if (warp == 0) {
        while (!mbar_try_wait(&S->scan_done, parity)) {
        }
    }
    work_sync();  // (also: the stage's parse results written by thread 0 are visible)
    while (!mbar_try_wait(&S->scan_done, parity)) {  // passes at once: every thread observes the phase
    }
    const uint64_t t = S->tile;
    const uint64_t g0 = t * (uint64_t)TILE;
    const uint32_t n_starts = S->n_starts, head_len = S->head_len, rest_total = S->rest_total,
                   tile_len = S->tile_len, none_pos = S->none_pos;
    const uint8_t *tile = S->buf + PRE;
    const uint64_t kept_before = S->kept_before;
    const uint32_t carry = S->carry;
    const uint32_t head_kept = carry == F_KEPT ? head_len : 0u;
    const uint32_t tile_kept = head_kept + rest_total;
    const uint32_t none_prefix = carry == F_NONE ? head_len : 0u;
    emit_stream<true>(S, C->wmax, tile, P.out_w + kept_before, tile_kept, n_starts, carry, head_kept, none_prefix,
                      tid);
    if (P.out_o) {
        const uint32_t own_end = none_pos < tile_len ? none_pos : tile_len;
        const uint32_t other_total = own_end > none_prefix + tile_kept ? own_end - none_prefix - tile_kept : 0u;
        // bytes of the other stream before this tile = owned bytes before it - kept bytes before it
        const uint64_t other_before = t == 0 ? 0 : (g0 - P.lead) - kept_before;
        emit_stream<false>(S, C->wmax, tile, P.out_o + other_before, other_total, n_starts, carry, head_kept,
                           none_prefix, tid);
    }
    if (t + 1 == P.n_tiles && tid == 0) P.res->kept_total = kept_before + tile_kept;
}

// ------------------------------------------------------------------ the scan warp
// Runs the decoupled look-back of every tile this CTA parses, as soon as the workers have published the
// tile's aggregate, and hands (kept_before, carry) back through the stage.  While it waits on predecessors
// the workers are already parsing the next tile, so the inclusive prefix of a tile is normally published
// long before anybody needs it and look-backs stay short.
__device__ __forceinline__ void scan_warp_loop(const FusedParams &P, CtaSmem *C, int lane) {
    for (uint32_t it = 0;; it++) {
        Stage *S = &C->st[it & 1];
        while (!mbar_try_wait(&S->agg_ready, (it >> 1) & 1)) __nanosleep(256);
        const uint64_t t = S->tile;
        if (t >= P.n_tiles) return;
        uint64_t kept_before;
        uint32_t carry;
        lookback_kept_warp(P.desc2, t, S->n_starts > 0, S->last_flag, S->head_len, S->rest_total, lane, &kept_before,
                           &carry);
        if (lane == 0) {
            S->kept_before = kept_before;
            S->carry = carry;
            mbar_arrive(&S->scan_done);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM) fastq_fused_kernel(FusedParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    CtaSmem *C = reinterpret_cast<CtaSmem *>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&C->st[s].full, 1);
            mbar_init(&C->st[s].agg_ready, 1);
            mbar_init(&C->st[s].scan_done, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        C->st[1].tile = ~0ull;
        take_ticket_and_load(P, &C->st[0]);
    }
    __syncthreads();  // the only CTA-wide barrier: the scan warp goes its own way from here
    if (tid >= NT) {
        scan_warp_loop(P, C, tid & 31);
        return;
    }
    unsigned long long my_reads_out = 0, my_reads_in = 0;  // thread 0 only
    // software pipeline: parse tile n+1 up to the issue of its set probes, copy tile n while they fly,
    // finish tile n+1 (aggregate published, scan warp notified), then start the load of tile n+2
    for (uint32_t it = 0;; it++) {
        Stage *cur = &C->st[it & 1], *prv = &C->st[(it & 1) ^ 1];
        work_sync();  // cur->tile is visible; nobody still reads what the next writes overwrite
        const bool have_cur = cur->tile < P.n_tiles;
        ParseState Z;
        if (have_cur) Z = parse_a(P, C, cur, (it >> 1) & 1, tid);
        else if (tid == 0) mbar_arrive(&cur->agg_ready);  // releases the scan warp: it sees the end ticket
        const bool have_prv = it > 0 && prv->tile < P.n_tiles;
        const uint32_t prv_parity = ((it - 1) >> 1) & 1;
        // The aggregate of cur must never wait on anybody (later tiles look back on it), so the copy of prv
        // goes between the two halves of the parse -- where it hides the probe latency -- only when the scan
        // warp has already delivered prv's prefix; otherwise cur is finished first.
        bool early = false;
        if (have_prv && have_cur) {
            if (tid == 0) C->early = mbar_try_wait(&prv->scan_done, prv_parity) ? 1u : 0u;
            work_sync();
            early = C->early != 0;
        }
        if (have_prv && (early || !have_cur)) copy_tile(P, C, prv, prv_parity, tid);
        if (!have_cur) break;
        parse_b(P, C, cur, Z, tid, &my_reads_in, &my_reads_out);
        if (have_prv && !early) copy_tile(P, C, prv, prv_parity, tid);
        work_sync();  // every read of prv's buffer is done
        if (tid == 0) take_ticket_and_load(P, prv);
    }
    if (tid == 0) {
        if (my_reads_out) atomicAdd(&P.res->reads_out, my_reads_out);
        if (my_reads_in) atomicAdd(&P.res->reads_in, my_reads_in);
    }
}

// exact verification of what the tiles assumed: (1) the speculated line phase of every tile against the
// true prefix of newline counts, (2) the signed newline-position sums vanish at every record end
// (seq and qual lengths agree for records that straddle tiles), (3) the file's line count is a multiple of 4
__global__ void fused_verify_kernel(const uint64_t *sum_prefix, const long long *sum_head, const uint8_t *has_term,
                                    const uint64_t *nl_prefix, const uint32_t *nl_count, const uint8_t *phase_used,
                                    uint64_t n_tiles, int is_last, FusedResult *res) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    if ((uint32_t)(nl_prefix[t] & 3) != (uint32_t)phase_used[t]) set_fallback(res, 11);
    if (has_term[t] && (long long)sum_prefix[t] + sum_head[t] != 0) set_fallback(res, 10);
    if (t + 1 == n_tiles && is_last) {
        const uint64_t lines = nl_prefix[t] + nl_count[t];
        if (lines & 3) set_fallback(res, 9);
    }
}

// Runs the fused kernel over d_in[0..n_in).  The first owned record starts at `lead` (< 16); records that
// start after own_len (when !is_last) are left to the next shard.  *used = 0 when the input turned out
// not to be canonical (the caller then takes the general path).
sgpu_status clean_fused_range(sgpu_ctx *c, const sgpu_idset *set, const uint8_t *d_in, size_t n_in, uint32_t lead,
                              size_t own_len, int is_last, int reverse, uint8_t *d_out_w, size_t cap_w, size_t *n_w,
                              uint8_t *d_out_o, size_t cap_o, size_t *n_o, sgpu_counts *counts, int *used) {
    *used = 0;
    // the fused kernel writes a byte partition of the input: both outputs must be able to hold it
    if (n_in == 0 || lead >= 16 || cap_w < n_in || (d_out_o && cap_o < n_in)) return SGPU_OK;
    cudaStream_t st = c->stream;
    static bool attr_done[64] = {false};
    const size_t smem = sizeof(CtaSmem);
    if (!attr_done[c->device & 63]) {
        SGPU_CUDA(cudaFuncSetAttribute(fastq_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SGPU_CUDA(cudaFuncSetAttribute(fastq_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
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
    FusedResult init;
    memset(&init, 0, sizeof(init));
    init.owned_end = ~0ull;
    static_assert(sizeof(FusedResult) <= 64 * 8 - 32 * 8, "pinned staging too small");
    memcpy(c->h_pinned + 32, &init, sizeof(init));
    SGPU_CUDA(cudaMemcpyAsync(res.p, c->h_pinned + 32, sizeof(init), cudaMemcpyHostToDevice, st));
    FusedParams P;
    P.in = d_in;
    P.n_in = n_in;
    P.n_tiles = n_tiles;
    P.own_len = own_len;
    P.lead = lead;
    P.is_last = is_last;
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
    static int occ[64] = {0};
    if (!occ[c->device & 63]) {
        int o = 0;
        SGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fastq_fused_kernel, NTHREADS, smem));
        occ[c->device & 63] = o > 0 ? o : 1;
    }
    uint64_t grid = (uint64_t)c->sm_count * occ[c->device & 63];  // persistent: every CTA is resident
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
                                                                         n_tiles, is_last, res.p);
    SGPU_LAUNCH(c);
    SGPU_CUDA(cudaGetLastError());
    FusedResult h;
    SGPU_TRY(read_u64s(c, res.p, (uint64_t *)&h, sizeof(FusedResult) / 8));
    // a shard must have seen the start of a foreign record: only then is its last owned record complete
    if (!h.fallback && !is_last && h.owned_end == ~0ull) {
        h.fallback = 1;
        h.reason = 12;
    }
    if (h.fallback) {  // *used stays 0: the general path decides (and reports errors)
        if (getenv("SGPU_DEBUG")) fprintf(stderr, "[sgpu] fused kernel fell back, reason %llu\n", h.reason);
        return SGPU_OK;
    }
    *used = 1;
    const uint64_t owned_end = is_last ? n_in : h.owned_end;
    const uint64_t other_total = owned_end - lead - h.kept_total;
    if (c->profiling) c->prof_alg_bytes += n_in + h.kept_total + (d_out_o ? other_total : 0);
    *n_w = (size_t)h.kept_total;
    if (n_o) *n_o = d_out_o ? (size_t)other_total : 0;
    counts->reads_in = h.reads_in;
    counts->reads_out = h.reads_out;
    counts->crlf = 0;
    counts->path = 1;
    return SGPU_OK;
}

// position of the (k+1)-th '\n' of buf[0..n), or ~0: one warp, 512 bytes per step (shards: the first record
// boundary lies within the first record's length of the cut)
__global__ void kth_newline_kernel(const uint8_t *buf, uint64_t n, uint32_t k, unsigned long long *out) {
    const int lane = threadIdx.x;
    uint32_t seen = 0;
    for (uint64_t base = 0; base < n; base += 512) {
        const uint64_t pos = base + (uint64_t)lane * 16;
        uint32_t m = 0;
        if (pos < n) {
            m = nl_mask16_v2(ld_nc_u4(buf + pos));
            if (n - pos < 16) m &= (1u << (n - pos)) - 1u;
        }
        uint32_t inc = __popc(m);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += x;
        }
        const uint32_t before = seen + inc - __popc(m);
        if (before <= k && k < before + __popc(m)) {
            uint32_t mm = m;
            for (uint32_t i = before; i < k; i++) mm &= mm - 1;
            *out = pos + (uint64_t)(__ffs(mm) - 1);
        }
        seen += __shfl_sync(0xffffffffu, inc, 31);
        if (seen > k) return;
    }
}

// one shard (see sgpu_clean_fastq_shard_dev): locate the first owned record start, then run the fused kernel
// from the 16-byte aligned address below it.  *used = 0: take the general path.
sgpu_status clean_fused_shard(sgpu_ctx *c, const sgpu_idset *set, const uint8_t *d_in, size_t n_in, size_t own_len,
                              uint64_t newlines_before, int is_first, int is_last, int reverse, uint8_t *d_out_w,
                              size_t cap_w, size_t *n_w, uint8_t *d_out_o, size_t cap_o, size_t *n_o,
                              sgpu_counts *counts, int *used) {
    *used = 0;
    uint64_t s0 = 0;
    if (!is_first) {
        // the newline that ends the previous shard's last record has global index == 3 (mod 4)
        const uint32_t k = (uint32_t)((3 - (newlines_before & 3)) & 3);
        DevBuf<unsigned long long> pos;
        SGPU_TRY(pos.alloc(1, c->stream));
        SGPU_CUDA(cudaMemsetAsync(pos.p, 0xFF, 8, c->stream));
        kth_newline_kernel<<<1, 32, 0, c->stream>>>(d_in, n_in, k, pos.p);
        SGPU_LAUNCH(c);
        uint64_t p;
        SGPU_TRY(read_u64s(c, pos.p, &p, 1));
        if (p == ~0ull) return SGPU_OK;  // no record boundary in the buffer: the general path sorts it out
        s0 = p + 1;
        if (s0 >= n_in || s0 > own_len) return SGPU_OK;  // owns nothing (general path: zero records / errors)
    }
    const uint32_t lead = (uint32_t)(s0 & 15);
    const uint64_t skip = s0 - lead;
    return clean_fused_range(c, set, d_in + skip, n_in - skip, lead, own_len - skip, is_last, reverse, d_out_w, cap_w,
                             n_w, d_out_o, cap_o, n_o, counts, used);
}

sgpu_status clean_fused(sgpu_ctx *c, const sgpu_idset *set, const uint8_t *d_in, size_t n_in, int reverse,
                        uint8_t *d_out_w, size_t cap_w, size_t *n_w, uint8_t *d_out_o, size_t cap_o, size_t *n_o,
                        sgpu_counts *counts, int *used) {
    return clean_fused_range(c, set, d_in, n_in, 0, n_in, 1, reverse, d_out_w, cap_w, n_w, d_out_o, cap_o, n_o,
                             counts, used);
}

}  // namespace sgpu

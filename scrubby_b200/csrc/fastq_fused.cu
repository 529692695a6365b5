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
// Structure (persistent CTAs, dynamic tile tickets, deadlock-free under partial residency):
//   tile t (TILE bytes + a 16 B pre-halo + a post-halo) is bulk-copied into shared memory by the
//   TMA engine (cp.async.bulk + mbarrier); a ticket is only taken when the CTA can load and count
//   the tile at once (a tile parked behind another one would stall every successor's look-back),
//   so latency is hidden by several resident CTAs per SM rather than by prefetching tickets;
//   P1  16-byte vector loads from smem -> '\n' bit masks (SWAR) + per-chunk counts, one packed
//       block scan, decoupled look-back #1 over tile newline counts  => global line number;
//   P2  every newline is classified by (line number mod 4): CR / "+\n" checks, record starts;
//   P3  one thread per record start: '@' check, id token, hash, exact probe of the id set,
//       seq/qual length check; block scan of kept bytes; decoupled look-back #2 carries
//       (kept bytes so far, keep-flag of the record that straddles the tile edge);
//   P4  runs of kept / removed bytes are copied smem -> global by warps with 16-byte stores
//       re-aligned to the destination (funnel shifts), long runs by the whole CTA.
// Records may straddle any number of tiles (ONT reads); only the id token must lie within the
// post-halo of the tile where the record starts.
#include <stdlib.h>

#include "fastq_records.cuh"

namespace sgpu {

constexpr int FT = 256;                    // threads per CTA
constexpr int FC = 8;                      // 16-byte chunks per thread
constexpr int CTAS_PER_SM = 5;             // resident CTAs per SM (shared memory and registers sized for it)
constexpr int TILE = FT * FC * 16;         // 32 KiB
constexpr int PRE = 16;                    // pre-halo (previous 16 bytes)
constexpr int HALO = 1024;                 // post-halo
constexpr int BUF = PRE + TILE + HALO;     // bytes per smem stage
constexpr int RMAX = 512;                  // record starts per tile
constexpr int LMAX = 4 * RMAX + 8;         // newline list capacity per tile
constexpr int LONG_RUN = 2048;             // runs at least this long are copied by the whole CTA

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
    uint8_t *has_term;
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

__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ------------------------------------------------------------------ block-wide decoupled look-back
// At >3 TB/s a 32 KiB tile retires every ~10 ns chip-wide, far faster than a 32-descriptor warp
// window can be walked (one L2 round trip per window): the inclusive prefixes would lag by the
// whole in-flight population and every walk would take tens of microseconds.  So the WHOLE CTA
// looks back: FT descriptors per round, which covers every tile that can be in flight
// (grid <= FT), i.e. one round in steady state.
struct LookbackSmem {
    uint32_t wmin[FT / 32], Sb[FT / 32], Fb[FT / 32];
    uint64_t red[2][FT / 32];
    uint64_t inc_val;
};

// look-back #1: exclusive prefix of the tiles' newline counts.  Called by all threads.
__device__ __forceinline__ uint64_t lookback_sum(unsigned long long *desc, uint64_t t, uint64_t mine,
                                                 LookbackSmem *L) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = FT / 32;
    if (tid == 0) st_relaxed(desc + t, ST_AGG | mine);
    uint64_t acc = 0;
    int64_t base = (int64_t)t - 1;
    while (true) {
        const int64_t idx = base - tid;
        unsigned long long d = ST_INC;  // virtual tiles before the file: inclusive prefix 0
        if (idx >= 0) {
            while (((d = ld_relaxed(desc + idx)) & ST_MASK) == 0) __nanosleep(64);
        }
        const bool is_inc = (d & ST_MASK) == ST_INC;
        const unsigned b = __ballot_sync(0xffffffffu, is_inc);
        if (lane == 0) L->wmin[warp] = b ? (uint32_t)(warp * 32 + __ffs(b) - 1) : 0xFFFFFFFFu;
        __syncthreads();
        uint32_t first = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < NW; w++) first = min(first, L->wmin[w]);
        const uint64_t v = ((uint32_t)tid <= first) ? (d & ~ST_MASK) : 0;
        const uint64_t sw = warp_sum(v);
        if (lane == 0) L->red[0][warp] = sw;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; w++) acc += L->red[0][w];
        __syncthreads();
        if (first != 0xFFFFFFFFu) break;
        base -= FT;
    }
    if (tid == 0) st_relaxed(desc + t, ST_INC | (acc + mine));
    return acc;
}

// look-back #2: kept bytes before the tile and the keep-flag of the record that straddles its edge.
// aggregate:  [61] has_start  [60] last_flag  [59:30] head_len  [29:0] rest_kept
// inclusive:  [60] carry flag after the tile  [59:0] kept bytes up to and including the tile
// A tile without a record start passes its predecessor's flag through and keeps head_len bytes iff
// that flag is set, so contributions are resolved against the nearest "provider" farther back.
__device__ __forceinline__ void lookback_kept(unsigned long long *desc, uint64_t t, bool has_start, bool last_flag,
                                              uint32_t head_len, uint32_t rest, LookbackSmem *L,
                                              uint64_t *kept_before, bool *carry_flag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = FT / 32;
    if (tid == 0)
        st_relaxed(desc + t, ST_AGG | (has_start ? D2_START : 0) | (last_flag ? D2_FLAG : 0) |
                                 ((uint64_t)head_len << 30) | rest);
    uint64_t acc = 0, pending = 0;
    bool known = false, my_flag = false;
    int64_t base = (int64_t)t - 1;
    while (true) {
        const int64_t idx = base - tid;
        unsigned long long d = ST_INC;  // virtual tiles before the file: nothing kept, flag 0
        if (idx >= 0) {
            while (((d = ld_relaxed(desc + idx)) & ST_MASK) == 0) __nanosleep(64);
        }
        const bool is_inc = (d & ST_MASK) == ST_INC;
        const unsigned b = __ballot_sync(0xffffffffu, is_inc);
        if (lane == 0) L->wmin[warp] = b ? (uint32_t)(warp * 32 + __ffs(b) - 1) : 0xFFFFFFFFu;
        __syncthreads();
        uint32_t first = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < NW; w++) first = min(first, L->wmin[w]);
        const bool relevant = (uint32_t)tid <= first;
        const bool provides = relevant && (is_inc || (d & D2_START));
        const unsigned Sw = __ballot_sync(0xffffffffu, provides);
        const unsigned Fw = __ballot_sync(0xffffffffu, provides && (d & D2_FLAG));
        if (lane == 0) {
            L->Sb[warp] = Sw;
            L->Fb[warp] = Fw;
        }
        if ((uint32_t)tid == first) L->inc_val = d & 0x0FFFFFFFFFFFFFFFull;
        __syncthreads();
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
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; w++) {
            acc += L->red[0][w];
            pending += L->red[1][w];
        }
        const uint64_t incv = L->inc_val;
        __syncthreads();
        if (first != 0xFFFFFFFFu) {
            acc += incv;
            break;
        }
        base -= FT;
    }
    *kept_before = acc;
    *carry_flag = my_flag;
    const uint64_t incl = acc + (my_flag ? head_len : 0) + rest;
    const bool out_flag = has_start ? last_flag : my_flag;
    if (tid == 0) st_relaxed(desc + t, ST_INC | (out_flag ? D2_FLAG : 0) | incl);
}

// ------------------------------------------------------------------ smem -> global run copy
// One warp copies n bytes from shared `src` to global `dst` (both arbitrarily aligned):
// 16-byte stores on the destination's alignment, source re-aligned with funnel shifts.
__device__ __forceinline__ void copy_run_lanes(uint8_t *dst, const uint8_t *src, uint32_t n, int lane, int nlanes) {
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    for (uint32_t i = lane; i < head; i += nlanes) dst[i] = src[i];
    dst += head;
    src += head;
    n -= head;
    const uint32_t nchunks = n >> 4;
    const uint32_t sa = smem_u32(src);
    const uint32_t q = sa & 15, qw = q >> 2, qb = (q & 3) * 8;
    const uint8_t *sbase = src - q;  // 16-byte aligned
    for (uint32_t i = lane; i < nchunks; i += nlanes) {
        const uint4 a = *reinterpret_cast<const uint4 *>(sbase + (size_t)i * 16);
        uint4 o;
        if (q == 0) {
            o = a;
        } else {
            const uint4 b = *reinterpret_cast<const uint4 *>(sbase + (size_t)i * 16 + 16);
            uint32_t w0, w1, w2, w3, w4;
            switch (qw) {  // warp-uniform
            case 0: w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; break;
            case 1: w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; break;
            case 2: w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; break;
            default: w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; break;
            }
            o.x = __funnelshift_r(w0, w1, qb);
            o.y = __funnelshift_r(w1, w2, qb);
            o.z = __funnelshift_r(w2, w3, qb);
            o.w = __funnelshift_r(w3, w4, qb);
        }
        st_global_v4(dst + (size_t)i * 16, o);
    }
    const uint32_t done = nchunks << 4;
    for (uint32_t i = done + lane; i < n; i += nlanes) dst[i] = src[i];
}

struct __align__(16) FusedSmem {
    uint64_t bar;
    uint64_t scan[80];
    LookbackSmem lb;
    uint32_t cur_lo, cur_hi, fallback, has_long;
    uint16_t nlp[LMAX];        // local positions of the tile's newlines
    uint16_t rs[RMAX + 2];     // local positions of record starts
    uint8_t rflag[RMAX + 2];   // 1: the record is written to out_w
    uint32_t rkoff[RMAX + 2];  // exclusive kept-byte offsets of the records (relative to the tile)
    __align__(16) uint8_t buf[BUF];
};

__device__ __forceinline__ void set_fallback(FusedResult *res, int reason) {
    if (atomicExch(&res->fallback, 1ull) == 0ull) res->reason = (unsigned long long)reason;
}

// block-wide exclusive scan of two values per thread at once (one set of barriers)
__device__ __forceinline__ void block_scan2(uint64_t a, uint64_t b, uint64_t *pa, uint64_t *pb, uint64_t *ta,
                                            uint64_t *tb, uint64_t *sm /* >= 80 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = FT / 32;
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
    __syncthreads();
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
    __syncthreads();
    *pa = ba + ia - a;
    *pb = bb + ib - b;
    *ta = sa;
    *tb = sb;
}

__device__ __forceinline__ void issue_tile_load(const FusedParams &P, FusedSmem *S, uint64_t t) {
    // bytes [t*TILE - PRE, t*TILE + TILE + HALO) clipped to the file, rounded up to 16
    uint64_t g0 = t * (uint64_t)TILE;
    uint64_t src0 = t ? g0 - PRE : 0;
    uint64_t end = g0 + TILE + HALO;
    if (end > P.n_in) end = P.n_in;
    uint32_t bytes = (uint32_t)(((end - src0) + 15) & ~15ull);
    uint8_t *dst = S->buf + (t ? 0 : PRE);
    mbar_expect_tx(&S->bar, bytes);
    bulk_g2s(dst, P.in + src0, bytes, &S->bar);
}

__global__ void __launch_bounds__(FT, CTAS_PER_SM) fastq_fused_kernel(FusedParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FusedSmem *S = reinterpret_cast<FusedSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long my_reads_out = 0;
    uint32_t phase = 0;

    if (tid == 0) {
        mbar_init(&S->bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    while (true) {
        // ---- take a ticket and start streaming that tile in.  Tickets are handed out in order to
        //      RUNNING CTAs only, so every predecessor a look-back spins on is making progress.
        if (tid == 0) {
            const unsigned long long tn = atomicAdd(&P.res->ticket, 1ull);
            S->cur_lo = (uint32_t)tn;
            S->cur_hi = (uint32_t)(tn >> 32);
            S->fallback = 0;
            S->has_long = 0;
            if (tn < P.n_tiles) issue_tile_load(P, S, tn);
        }
        __syncthreads();
        const uint64_t t = ((uint64_t)S->cur_hi << 32) | S->cur_lo;
        if (t >= P.n_tiles) break;
        uint8_t *buf = S->buf;
        if (t == 0 && tid < PRE) buf[tid] = '\n';  // no predecessor: the pre-halo reads as a newline
        const uint64_t g0 = t * (uint64_t)TILE;
        const uint32_t tile_len = (uint32_t)((P.n_in - g0) < (uint64_t)TILE ? (P.n_in - g0) : (uint64_t)TILE);
        const uint32_t avail =
            (uint32_t)((P.n_in - g0) < (uint64_t)(TILE + HALO) ? (P.n_in - g0) : (uint64_t)(TILE + HALO));
        const uint8_t *tile = buf + PRE;  // tile[-16 .. avail)
        while (!mbar_try_wait(&S->bar, phase)) {
        }
        phase ^= 1;
        __syncthreads();  // pre-halo fill of tile 0 visible

        // ---- P1: newline masks, counts, high-bit test
        uint32_t m[FC];
        uint32_t hi_or = 0;
        uint64_t packed[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < FC; k++) {
            const uint32_t pos = (uint32_t)(k * FT + tid) * 16;
            uint4 v = *reinterpret_cast<const uint4 *>(tile + pos);
            uint32_t mm = nl_mask16(v);
            if (pos + 16 > tile_len) {  // last tile: bytes past the end of the file are stale
                if (pos >= tile_len) {
                    mm = 0;
                    v = make_uint4(0, 0, 0, 0);
                } else {
                    const uint32_t valid = tile_len - pos;
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
        if (hi_or & 0x80808080u) S->fallback = 1;  // reason 1: non-ASCII byte, Unicode rules needed
        uint64_t pre[2], tot[2];
        block_scan2(packed[0], packed[1], &pre[0], &pre[1], &tot[0], &tot[1], S->scan);
        uint32_t row_base[FC];
        uint32_t n_nl = 0;
#pragma unroll
        for (int k = 0; k < FC; k++) {
            row_base[k] = n_nl + (uint32_t)((pre[k >> 2] >> (16 * (k & 3))) & 0xFFFF);
            n_nl += (uint32_t)((tot[k >> 2] >> (16 * (k & 3))) & 0xFFFF);
        }
        // ---- look-back #1 (whole CTA): newlines before this tile
        const uint64_t L0 = lookback_sum(P.desc1, t, n_nl, &S->lb);
        const bool dense = n_nl > (uint32_t)LMAX;
        // position 0 starts a record iff 4k newlines precede it and the previous byte is one
        const bool pos0_start = ((L0 & 3) == 0) && tile[-1] == '\n';
        const uint32_t c0 = (uint32_t)(L0 & 3);

        // ---- P2: classify every newline by its role (line number mod 4)
        uint32_t bad = 0;  // fallback reason: 3 CRLF, 4 separator, 5 too many records
        if (!dense) {
#pragma unroll
            for (int k = 0; k < FC; k++) {
                uint32_t mm = m[k];
                uint32_t r = row_base[k];
                const uint32_t pos = (uint32_t)(k * FT + tid) * 16;
                while (mm) {
                    const uint32_t p = pos + (uint32_t)(__ffs(mm) - 1);
                    mm &= mm - 1;
                    S->nlp[r] = (uint16_t)p;
                    const uint32_t role = (c0 + r) & 3;
                    if (tile[(int)p - 1] == '\r') bad = 3;  // CRLF: not canonical
                    if (role == 1) {                        // end of the sequence line: "+\n" must follow
                        if (p + 2 >= avail) bad = 4;
                        else if (tile[p + 1] != '+' || tile[p + 2] != '\n') bad = 4;
                    } else if (role == 3 && p + 1 < tile_len) {  // a record starts at p + 1
                        const uint32_t j = (uint32_t)(((L0 + r) >> 2) - (L0 >> 2)) + (pos0_start ? 1u : 0u);
                        if (j < (uint32_t)RMAX) S->rs[j] = (uint16_t)(p + 1);
                        else bad = 5;
                    }
                    r++;
                }
            }
        }
        if (tid == 0 && pos0_start) S->rs[0] = 0;
        if (bad || dense) S->fallback = dense ? 2 : bad;
        __syncthreads();
        // number of record starts inside the tile
        const uint32_t n_term = (uint32_t)(((L0 + n_nl) >> 2) - (L0 >> 2));
        uint32_t n_starts = n_term + (pos0_start ? 1u : 0u);
        if (n_term > 0 && !dense) {
            // the last terminating newline may sit on the tile's final byte: its record belongs to the next tile
            const uint32_t r_last = ((3u - c0) & 3u) + 4u * (n_term - 1);
            if ((uint32_t)S->nlp[r_last] + 1u >= tile_len) n_starts--;
        }
        if (dense || n_starts > (uint32_t)RMAX) n_starts = 0;  // (the fallback flag is already raised)

        // ---- P3: one thread per record start: '@', id token, exact probe, seq/qual length check
        uint64_t rest_total = 0;
        for (uint32_t jb = 0; jb < n_starts; jb += FT) {
            const uint32_t j = jb + tid;
            uint32_t my_len = 0;
            bool my_flag = false;
            if (j < n_starts) {
                const uint32_t s = S->rs[j];
                const uint32_t e = (j + 1 < n_starts) ? S->rs[j + 1] : tile_len;
                my_len = e - s;
                uint32_t why = tile[s] == '@' ? 0u : 6u;  // 6 '@', 7 id token, 8 seq/qual lengths
                // id token: skip leading blanks, run to the next blank / newline (ASCII: high bytes fell back)
                uint32_t i = s + 1;
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
                    const int sgn = -(int)S->nlp[r0 + 1] + (int)S->nlp[r0 + 2] + (int)S->nlp[r0 + 3] - (int)S->nlp[r0 + 4];
                    if (sgn != 0) why = why ? why : 8u;
                }
                if (why) S->fallback = why;
                S->rflag[j] = my_flag ? 1 : 0;
                if (my_len >= (uint32_t)LONG_RUN) S->has_long = 1;
            }
            uint64_t koff, dummy, round_total, dummy2;
            block_scan2(my_flag ? my_len : 0u, 0, &koff, &dummy, &round_total, &dummy2, S->scan);
            if (j < n_starts) S->rkoff[j] = (uint32_t)(rest_total + koff);
            rest_total += round_total;
            const unsigned kept_ballot = __ballot_sync(0xffffffffu, my_flag);
            if (lane == 0) my_reads_out += __popc(kept_ballot);
        }
        __syncthreads();  // rs / rflag / rkoff complete (also when the loop ran zero times)
        const uint32_t head_len = n_starts ? (uint32_t)S->rs[0] : tile_len;

        // ---- the cross-tile length-check sums (one thread, before it joins the look-back)
        if (tid == 0) {
            if (head_len >= (uint32_t)LONG_RUN) S->has_long = 1;
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
            // end-of-file conditions of canonical input
            if (t + 1 == P.n_tiles) {
                if (((L0 + n_nl) & 3) != 0 || tile[tile_len - 1] != '\n') S->fallback = 9;
                P.res->reads_in = (L0 + n_nl) >> 2;
            }
        }
        // ---- look-back #2 (whole CTA): kept bytes before the tile + the straddling record's flag
        uint64_t kept_before;
        bool carry;
        {
            const bool last_flag = n_starts ? (S->rflag[n_starts - 1] != 0) : false;
            lookback_kept(P.desc2, t, n_starts > 0, last_flag, head_len, (uint32_t)rest_total, &S->lb, &kept_before,
                          &carry);
        }
        if (tid == 0 && S->fallback) set_fallback(P.res, (int)S->fallback);

        // ---- P4: copy runs.  run 0 = carried-in head, run j>=1 = record j-1's bytes inside the tile
        {
            const uint64_t head_kept = carry ? head_len : 0;
            const uint32_t n_runs = n_starts + 1;
            const uint64_t w_base = kept_before + head_kept;  // + rkoff[j]
            const uint64_t o_base = g0 - kept_before;          // + (src - head_kept - rkoff[j])
            const bool has_long = S->has_long != 0;
            // short runs: one warp each
            for (uint32_t r = warp; r < n_runs; r += FT / 32) {
                uint32_t s, e, ko;
                bool fl;
                if (r == 0) {
                    s = 0; e = head_len; fl = carry; ko = 0;
                } else {
                    s = S->rs[r - 1];
                    e = (r < n_starts) ? S->rs[r] : tile_len;
                    fl = S->rflag[r - 1] != 0;
                    ko = S->rkoff[r - 1];
                }
                const uint32_t len = e - s;
                if (len == 0 || len >= (uint32_t)LONG_RUN) continue;
                if (fl) {
                    const uint64_t d = (r == 0) ? kept_before : w_base + ko;
                    copy_run_lanes(P.out_w + d, tile + s, len, lane, 32);
                } else if (P.out_o) {
                    const uint64_t d = (r == 0) ? o_base : o_base + (s - head_kept - ko);
                    copy_run_lanes(P.out_o + d, tile + s, len, lane, 32);
                }
            }
            // long runs (ONT-sized records): the whole CTA, one after another
            if (has_long) {
                for (uint32_t r = 0; r < n_runs; r++) {
                    uint32_t s, e, ko;
                    bool fl;
                    if (r == 0) {
                        s = 0; e = head_len; fl = carry; ko = 0;
                    } else {
                        s = S->rs[r - 1];
                        e = (r < n_starts) ? S->rs[r] : tile_len;
                        fl = S->rflag[r - 1] != 0;
                        ko = S->rkoff[r - 1];
                    }
                    const uint32_t len = e - s;
                    if (len < (uint32_t)LONG_RUN) continue;
                    if (fl) {
                        const uint64_t d = (r == 0) ? kept_before : w_base + ko;
                        copy_run_lanes(P.out_w + d, tile + s, len, tid, FT);
                    } else if (P.out_o) {
                        const uint64_t d = (r == 0) ? o_base : o_base + (s - head_kept - ko);
                        copy_run_lanes(P.out_o + d, tile + s, len, tid, FT);
                    }
                }
            }
            if (t + 1 == P.n_tiles && tid == 0) P.res->kept_total = kept_before + head_kept + rest_total;
        }
        __syncthreads();  // all reads of the buffer are done before it is refilled
    }
    if (lane == 0 && my_reads_out) atomicAdd(&P.res->reads_out, my_reads_out);
}

// cross-tile seq/qual length check: prefix of the signed sums must vanish at every record end
__global__ void fused_sumcheck_kernel(const uint64_t *prefix, const long long *sum_head, const uint8_t *has_term,
                                      uint64_t n_tiles, FusedResult *res) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles || !has_term[t]) return;
    if ((long long)prefix[t] + sum_head[t] != 0) set_fallback(res, 10);
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
    DevBuf<uint8_t> has_term;
    DevBuf<FusedResult> res;
    SGPU_TRY(desc.alloc(2 * n_tiles, st));
    SGPU_TRY(sums.alloc(2 * n_tiles, st));
    SGPU_TRY(prefix.alloc(n_tiles, st));
    SGPU_TRY(has_term.alloc(n_tiles, st));
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
    P.has_term = has_term.p;
    P.res = res.p;
    int occ = 0;
    SGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fastq_fused_kernel, FT, smem));
    if (occ < 1) occ = 1;
    uint64_t grid = (uint64_t)c->sm_count * occ;
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
    fastq_fused_kernel<<<(unsigned)grid, FT, smem, st>>>(P);
    SGPU_LAUNCH(c);
    if (c->profiling) SGPU_CUDA(cudaEventRecord(c->prof_events[c->prof_used++].second, st));
    SGPU_TRY(exclusive_scan_u64(c, (const uint64_t *)P.sum_total, prefix.p, n_tiles, nullptr));
    fused_sumcheck_kernel<<<(unsigned)ceil_div(n_tiles, 256), 256, 0, st>>>(prefix.p, P.sum_head, has_term.p, n_tiles,
                                                                           res.p);
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

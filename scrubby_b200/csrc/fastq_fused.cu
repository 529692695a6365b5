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
//   TMA engine (cp.async.bulk + mbarrier), double buffered so the next tile streams in while
//   this one is processed;
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
constexpr int FC = 4;                      // 16-byte chunks per thread
constexpr int TILE = FT * FC * 16;         // 16 KiB
constexpr int PRE = 16;                    // pre-halo (previous 16 bytes)
constexpr int HALO = 1024;                 // post-halo
constexpr int BUF = PRE + TILE + HALO;     // bytes per smem stage
constexpr int LMAX = 4 * FT + 8;           // newline list capacity per tile
constexpr int RMAX = FT;                   // record starts per tile (one thread each)
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
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_global_v4(void *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ------------------------------------------------------------------ look-back #1: newline counts
__device__ __forceinline__ uint64_t lookback_sum(unsigned long long *desc, uint64_t t, uint64_t mine, int lane) {
    if (lane == 0) st_release(desc + t, ST_AGG | mine);
    uint64_t acc = 0;
    int64_t base = (int64_t)t - 1;
    while (true) {
        int64_t idx = base - lane;
        unsigned long long d;
        if (idx < 0) {
            d = ST_INC;  // virtual tile -1: inclusive prefix 0
        } else {
            d = ld_acquire(desc + idx);
        }
        // every lane must hold a non-empty descriptor before the window is interpreted
        while (__any_sync(0xffffffffu, (d & ST_MASK) == 0)) {
            if ((d & ST_MASK) == 0) d = ld_acquire(desc + idx);
        }
        unsigned inc = __ballot_sync(0xffffffffu, (d & ST_MASK) == ST_INC);
        int first = inc ? __ffs(inc) - 1 : 32;
        uint64_t v = (lane <= first) ? (d & ~ST_MASK) : 0;
        acc += warp_sum(v);
        if (inc) break;
        base -= 32;
    }
    if (lane == 0) st_release(desc + t, ST_INC | (acc + mine));
    return acc;
}

// ------------------------------------------------------------------ look-back #2: kept bytes + carried keep-flag
// aggregate:  [61] has_start  [60] last_flag  [59:30] head_len  [29:0] rest_kept
// inclusive:  [60] carry flag after the tile  [59:0] kept bytes up to and including the tile
__device__ __forceinline__ void lookback_kept(unsigned long long *desc, uint64_t t, bool has_start, bool last_flag,
                                              uint32_t head_len, uint32_t rest, int lane, uint64_t *kept_before,
                                              bool *carry_flag) {
    if (lane == 0)
        st_release(desc + t, ST_AGG | (has_start ? D2_START : 0) | (last_flag ? D2_FLAG : 0) |
                                 ((uint64_t)head_len << 30) | rest);
    uint64_t acc = 0, pending = 0;
    bool known = false, my_flag = false;
    int64_t base = (int64_t)t - 1;
    while (true) {
        int64_t idx = base - lane;
        unsigned long long d;
        if (idx < 0) {
            d = ST_INC;  // virtual tile -1: nothing kept, flag 0
        } else {
            d = ld_acquire(desc + idx);
        }
        while (__any_sync(0xffffffffu, (d & ST_MASK) == 0)) {
            if ((d & ST_MASK) == 0) d = ld_acquire(desc + idx);
        }
        const bool is_inc = (d & ST_MASK) == ST_INC;
        unsigned inc = __ballot_sync(0xffffffffu, is_inc);
        int first = inc ? __ffs(inc) - 1 : 32;
        const bool relevant = lane <= first;
        const bool provides = relevant && (is_inc || (d & D2_START));
        unsigned S = __ballot_sync(0xffffffffu, provides);
        unsigned F = __ballot_sync(0xffffffffu, provides && (d & D2_FLAG));
        uint64_t contrib = 0, defer = 0;
        if (relevant && !is_inc) {
            uint64_t hl = (d >> 30) & 0x3FFFFFFFull, rs = d & 0x3FFFFFFFull;
            unsigned above = lane < 31 ? (S >> (lane + 1)) << (lane + 1) : 0u;  // providers farther back
            if (above) {
                int q = __ffs(above) - 1;
                contrib = rs + (((F >> q) & 1u) ? hl : 0);
            } else {
                contrib = rs;
                defer = hl;  // its flag lies in a farther window
            }
        }
        if (S) {
            int q0 = __ffs(S) - 1;  // nearest provider of this window resolves what was pending
            bool f0 = (F >> q0) & 1u;
            if (f0) acc += pending;
            pending = 0;
            if (!known) {
                known = true;
                my_flag = f0;
            }
        }
        acc += warp_sum(contrib);
        pending += warp_sum(defer);
        if (inc) {
            uint64_t incv = __shfl_sync(0xffffffffu, (uint64_t)(d & 0x0FFFFFFFFFFFFFFFull), first);
            acc += incv;
            break;
        }
        base -= 32;
    }
    *kept_before = acc;
    *carry_flag = my_flag;
    uint64_t incl = acc + (my_flag ? head_len : 0) + rest;
    bool out_flag = has_start ? last_flag : my_flag;
    if (lane == 0) st_release(desc + t, ST_INC | (out_flag ? D2_FLAG : 0) | incl);
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
    uint64_t bar[2];
    uint64_t scan[40];
    // broadcast slots
    uint64_t L0;            // newlines before this tile
    uint64_t kept_before;   // kept bytes before this tile
    uint32_t n_nl, n_starts, rest_kept, carry_flag, next_tile_lo, next_tile_hi, fallback, pos0_start;
    uint16_t nlp[LMAX];     // local positions of the tile's newlines
    uint16_t rs[RMAX + 2];  // local positions of record starts
    uint8_t rflag[RMAX + 2];
    uint32_t rkoff[RMAX + 2];  // exclusive kept-byte offsets of the runs (relative to the tile)
    __align__(16) uint8_t buf[2][BUF];
};

__device__ __forceinline__ void set_fallback(FusedResult *res, int reason) {
    if (atomicExch(&res->fallback, 1ull) == 0ull) res->reason = (unsigned long long)reason;
}

__device__ __forceinline__ uint64_t block_scan_u64(uint64_t v, uint64_t *total, uint64_t *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < (FT / 32) ? sm[lane] : 0, winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        sm[lane] = winc - w;
        if (lane == 31) sm[32] = winc;
    }
    __syncthreads();
    uint64_t res = sm[warp] + inc - v;
    *total = sm[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ void issue_tile_load(const FusedParams &P, FusedSmem *S, int stage, uint64_t t) {
    // bytes [t*TILE - PRE, t*TILE + TILE + HALO) clipped to the file, rounded up to 16
    uint64_t g0 = t * (uint64_t)TILE;
    uint64_t src0 = t ? g0 - PRE : 0;
    uint64_t end = g0 + TILE + HALO;
    if (end > P.n_in) end = P.n_in;
    uint32_t bytes = (uint32_t)(((end - src0) + 15) & ~15ull);
    uint8_t *dst = S->buf[stage] + (t ? 0 : PRE);
    mbar_expect_tx(&S->bar[stage], bytes);
    bulk_g2s(dst, P.in + src0, bytes, &S->bar[stage]);
}

__global__ void __launch_bounds__(FT) fastq_fused_kernel(FusedParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FusedSmem *S = reinterpret_cast<FusedSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long my_reads_out = 0;

    if (tid == 0) {
        mbar_init(&S->bar[0], 1);
        mbar_init(&S->bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        unsigned long long t0 = atomicAdd(&P.res->ticket, 1ull);
        S->next_tile_lo = (uint32_t)t0;
        S->next_tile_hi = (uint32_t)(t0 >> 32);
        // tile 0 has no predecessor: its pre-halo reads as a newline (record start, no CR)
        if (t0 < P.n_tiles) issue_tile_load(P, S, 0, t0);
    }
    __syncthreads();
    uint64_t t = ((uint64_t)S->next_tile_hi << 32) | S->next_tile_lo;
    int stage = 0;
    uint32_t phase[2] = {0, 0};

    while (t < P.n_tiles) {
        // ---- take the next ticket and start streaming that tile into the other stage
        if (tid == 0) {
            unsigned long long tn = atomicAdd(&P.res->ticket, 1ull);
            S->next_tile_lo = (uint32_t)tn;
            S->next_tile_hi = (uint32_t)(tn >> 32);
            if (tn < P.n_tiles) issue_tile_load(P, S, stage ^ 1, tn);
            S->fallback = 0;
        }
        uint8_t *buf = S->buf[stage];
        if (t == 0 && tid < PRE) buf[tid] = '\n';
        const uint64_t g0 = t * (uint64_t)TILE;
        const uint32_t tile_len = (uint32_t)((P.n_in - g0) < (uint64_t)TILE ? (P.n_in - g0) : (uint64_t)TILE);
        const uint32_t avail = (uint32_t)((P.n_in - g0) < (uint64_t)(TILE + HALO) ? (P.n_in - g0) : (uint64_t)(TILE + HALO));
        const uint8_t *tile = buf + PRE;  // tile[-16 .. avail)
        while (!mbar_try_wait(&S->bar[stage], phase[stage])) {
        }
        phase[stage] ^= 1;
        __syncthreads();  // pre-halo fill of tile 0 + S->fallback reset visible

        // ---- P1: newline masks, counts, high-bit test
        uint32_t m[FC];
        uint32_t hi_or = 0;
        uint64_t packed = 0;
#pragma unroll
        for (int k = 0; k < FC; k++) {
            const uint32_t pos = (uint32_t)(k * FT + tid) * 16;
            uint4 v = *reinterpret_cast<const uint4 *>(tile + pos);
            uint32_t mm = nl_mask16(v);
            if (pos + 16 > tile_len) {
                mm = pos < tile_len ? (mm & ((1u << (tile_len - pos)) - 1u)) : 0u;
                if (pos >= tile_len) v = make_uint4(0, 0, 0, 0);
                // (a partially valid chunk may carry stale high bits past the end: mask them too)
                else {
                    uint32_t valid = tile_len - pos;
                    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        int rem = (int)valid - 4 * x;
                        if (rem <= 0) w[x] = 0;
                        else if (rem < 4) w[x] &= (1u << (8 * rem)) - 1u;
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            m[k] = mm;
            hi_or |= (v.x | v.y | v.z | v.w);
            packed |= (uint64_t)__popc(mm) << (16 * k);
        }
        if (hi_or & 0x80808080u) S->fallback = 1;  // reason 1: non-ASCII byte, Unicode rules needed
        uint64_t tot_packed;
        uint64_t pre_packed = block_scan_u64(packed, &tot_packed, S->scan);
        uint32_t row_base[FC];
        uint32_t n_nl = 0;
#pragma unroll
        for (int k = 0; k < FC; k++) {
            row_base[k] = n_nl + (uint32_t)((pre_packed >> (16 * k)) & 0xFFFF);
            n_nl += (uint32_t)((tot_packed >> (16 * k)) & 0xFFFF);
        }
        // ---- look-back #1 (warp 0)
        if (warp == 0) {
            uint64_t L0 = lookback_sum(P.desc1, t, n_nl, lane);
            if (lane == 0) {
                S->L0 = L0;
                S->n_nl = n_nl;
            }
        }
        __syncthreads();
        const uint64_t L0 = S->L0;
        const bool dense = n_nl > (uint32_t)LMAX;
        // position 0 starts a record iff 4k newlines precede it and the previous byte is one
        const bool pos0_start = ((L0 & 3) == 0) && tile[-1] == '\n';
        const uint32_t c0 = (uint32_t)(L0 & 3);
        // terminating newlines (role 3) with rank < r: floor((L0+r)/4) - floor(L0/4)

        // ---- P2: classify every newline
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
                    if (tile[(int)p - 1] == '\r') bad = 3;                          // CRLF: not canonical
                    if (role == 1) {                                                 // end of the sequence line
                        if (p + 2 >= avail) bad = 4;                                 // separator must be "+\n"
                        else if (tile[p + 1] != '+' || tile[p + 2] != '\n') bad = 4;
                    } else if (role == 3 && p + 1 < tile_len) {                      // a record starts at p + 1
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
        uint32_t n_term = (uint32_t)(((L0 + n_nl) >> 2) - (L0 >> 2));
        uint32_t n_starts = n_term + (pos0_start ? 1u : 0u);
        if (n_term > 0 && !dense) {
            // the last terminating newline may sit on the tile's final byte: its record belongs to the next tile
            const uint32_t r_last = ((3u - c0) & 3u) + 4u * (n_term - 1);
            if ((uint32_t)S->nlp[r_last] + 1u >= tile_len) n_starts--;
        }
        if (dense) n_starts = 0;
        if (n_starts > (uint32_t)RMAX) n_starts = 0;  // (bad was raised above)

        // ---- P3: one thread per record start
        uint32_t my_len = 0;
        bool my_flag = false;
        if (tid < (int)n_starts) {
            const uint32_t s = S->rs[tid];
            const uint32_t e = (tid + 1 < (int)n_starts) ? S->rs[tid + 1] : tile_len;
            my_len = e - s;
            uint32_t why = tile[s] == '@' ? 0u : 6u;  // 6 '@', 7 id token, 8 seq/qual lengths
            // id token: skip leading blanks, run to the next blank / newline (ASCII: high bytes fell back)
            uint32_t i = s + 1;
            while (i < avail && is_ws_ascii(tile[i]) && tile[i] != '\n') i++;
            uint32_t j = i;
            while (j < avail && !is_ws_ascii(tile[j])) j++;
            if (j >= avail || j == i) why = why ? why : 7u;  // token runs past the halo, or empty id (error 9)
            if (!why) {
                bool hit = idset_contains(P.set, tile + i, j - i);
                my_flag = P.reverse ? hit : !hit;
            }
            // seq/qual length equality for records whose four newlines are inside the tile
            const int r0 = pos0_start ? 4 * tid - 1 : (int)((3u - c0) & 3u) + 4 * (tid - 0);
            if (r0 + 4 < (int)n_nl) {
                int sgn = -(int)S->nlp[r0 + 1] + (int)S->nlp[r0 + 2] + (int)S->nlp[r0 + 3] - (int)S->nlp[r0 + 4];
                if (sgn != 0) why = why ? why : 8u;
            }
            if (why) S->fallback = why;
            S->rflag[tid] = my_flag ? 1 : 0;
        }
        uint64_t rest_total;
        uint64_t koff = block_scan_u64(my_flag ? my_len : 0u, &rest_total, S->scan);
        if (tid < (int)n_starts) S->rkoff[tid] = (uint32_t)koff;
        // packed count of kept records for the counters
        unsigned kept_ballot = __ballot_sync(0xffffffffu, my_flag);
        if (lane == 0) my_reads_out += __popc(kept_ballot);

        // ---- look-back #2 and the cross-tile length-check sums (warp 0)
        if (warp == 0) {
            const uint32_t head_len = n_starts ? (uint32_t)S->rs[0] : tile_len;
            const bool last_flag = n_starts ? (S->rflag[n_starts - 1] != 0) : false;
            uint64_t kb;
            bool cf;
            lookback_kept(P.desc2, t, n_starts > 0, last_flag, head_len, (uint32_t)rest_total, lane, &kb, &cf);
            if (lane == 0) {
                S->kept_before = kb;
                S->carry_flag = cf ? 1 : 0;
                S->n_starts = n_starts;
                // signed newline-position sums: -p1 +p2 +p3 -p4 per record must vanish
                long long head = 0, total = 0;
                const int r_first = (int)((3u - c0) & 3u);
                if (!dense) {
                    if (n_term == 0) {
                        for (uint32_t r = 0; r < n_nl; r++) {
                            uint32_t role = (c0 + r) & 3;
                            long long pp = (long long)(g0 + S->nlp[r]);
                            total += (role == 0 || role == 3) ? -pp : pp;
                        }
                    } else {
                        for (int r = 0; r <= r_first; r++) {
                            uint32_t role = (c0 + r) & 3;
                            long long pp = (long long)(g0 + S->nlp[r]);
                            head += (role == 0 || role == 3) ? -pp : pp;
                        }
                        total = head;
                        for (uint32_t r = (uint32_t)r_first + 4u * (n_term - 1) + 1u; r < n_nl; r++) {
                            uint32_t role = (c0 + r) & 3;
                            long long pp = (long long)(g0 + S->nlp[r]);
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
        }
        __syncthreads();
        if (tid == 0 && S->fallback) set_fallback(P.res, (int)S->fallback);

        // ---- P4: copy runs.  run 0 = carried-in head, run j>=1 = record j-1's bytes inside the tile
        {
            const uint64_t kept_before = S->kept_before;
            const bool carry = S->carry_flag != 0;
            const uint32_t head_len = n_starts ? (uint32_t)S->rs[0] : tile_len;
            const uint64_t head_kept = carry ? head_len : 0;
            const uint32_t n_runs = n_starts + 1;
            const uint64_t w_base = kept_before + head_kept;               // + rkoff[j]
            const uint64_t o_base = g0 - kept_before;                       // + (src - head_kept - rkoff[j])
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
                    uint64_t d = (r == 0) ? kept_before : w_base + ko;
                    copy_run_lanes(P.out_w + d, tile + s, len, lane, 32);
                } else if (P.out_o) {
                    uint64_t d = (r == 0) ? o_base : o_base + (s - head_kept - ko);
                    copy_run_lanes(P.out_o + d, tile + s, len, lane, 32);
                }
            }
            // long runs: the whole CTA, one after another
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
                    uint64_t d = (r == 0) ? kept_before : w_base + ko;
                    copy_run_lanes(P.out_w + d, tile + s, len, tid, FT);
                } else if (P.out_o) {
                    uint64_t d = (r == 0) ? o_base : o_base + (s - head_kept - ko);
                    copy_run_lanes(P.out_o + d, tile + s, len, tid, FT);
                }
            }
            if (t + 1 == P.n_tiles && tid == 0)
                P.res->kept_total = kept_before + head_kept + rest_total;
        }
        __syncthreads();  // all reads of this stage are done before it is refilled
        t = ((uint64_t)S->next_tile_hi << 32) | S->next_tile_lo;
        stage ^= 1;
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

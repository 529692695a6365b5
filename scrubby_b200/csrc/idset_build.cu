// idset_build.cu -- bulk build of the read-id set, staged through shared memory.
//
// Replaces the HashSet::insert loops of alignment.rs:74,106 and classifier.rs:284,322 when a whole evidence file is
// turned into a set at once.  Inserting key by key into a table far larger than L2 costs one random 128-byte line read
// AND written back per key (idset_insert_kernel: 256 B of DRAM traffic for a 16-byte slot).  Here the keys' slot images
// are first grouped by the PAGE of the table they belong to (an exact counting sort: one pass counts the keys of every
// page, a scan turns the counts into offsets, one pass scatters the 16-byte images -- the open tails of the page
// lists, one line each, stay in L2 until they are full), then every page is assembled in shared memory by one CTA and
// written out once, coalesced: the table is written exactly once, nothing of it is ever read, nothing needs zeroing.
//
//   count   : candidate spans -> slot image -> page; per-page counters, arena space for long ids   [idset_count_kernel]
//   scan    : page counts -> page offsets                                                           [scan.cu]
//   scatter : candidate spans -> slot image -> the page's list                                      [idset_scatter_kernel]
//   build   : page's images -> 32 KiB page in shared memory (per-bucket locks) -> table             [idset_page_kernel]
#include <math.h>

#include <algorithm>

#include "idset.cuh"

namespace sgpu {

struct BuildStats {
    unsigned long long inserted;     // distinct keys
    unsigned long long arena_used;   // bytes of arena handed out to long ids (duplicates included)
};

// key_image() for an inline key (1..15 bytes) in GLOBAL memory from five aligned 32-bit loads instead of up to fifteen
// byte loads: the words that hold the key are re-aligned with funnel shifts and cut to len + 1 bytes (same image as
// fastq_fused.cu::record_prepare builds from shared memory).  Only words that contain key bytes are read.
__device__ __forceinline__ void inline_image_global(const uint8_t *p, uint32_t len, uint64_t *lo, uint64_t *hi) {
    const uint32_t mis = (uint32_t)((uintptr_t)p & 3u), sh = mis * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - mis);
    const uint32_t need = mis + len;  // bytes from the aligned base
    const uint32_t x0 = __ldg(w), x1 = need > 4 ? __ldg(w + 1) : 0u, x2 = need > 8 ? __ldg(w + 2) : 0u,
                   x3 = need > 12 ? __ldg(w + 3) : 0u, x4 = need > 16 ? __ldg(w + 4) : 0u;
    const uint32_t w0 = __funnelshift_r(x0, x1, sh), w1 = __funnelshift_r(x1, x2, sh), w2 = __funnelshift_r(x2, x3, sh),
                   w3 = __funnelshift_r(x3, x4, sh);
    uint32_t s0 = (w0 << 8) | len, s1 = __funnelshift_l(w0, w1, 8), s2 = __funnelshift_l(w1, w2, 8),
             s3 = __funnelshift_l(w2, w3, 8);
    const uint32_t nb = len + 1;  // 2..16 bytes of image
    const uint32_t full = nb >> 2, part = (nb & 3u) * 8u;
    const uint32_t pm = (1u << part) - 1u;  // part == 0 -> 0
    s0 = full > 0 ? s0 : s0 & pm;
    s1 = full > 1 ? s1 : (full == 1 ? s1 & pm : 0u);
    s2 = full > 2 ? s2 : (full == 2 ? s2 & pm : 0u);
    s3 = full > 3 ? s3 : (full == 3 ? s3 & pm : 0u);
    *lo = (uint64_t)s0 | ((uint64_t)s1 << 32);
    *hi = (uint64_t)s2 | ((uint64_t)s3 << 32);
}

constexpr uint32_t NO_PAGE = 0xFFFFFFFFu;

// count: the page of every selected candidate (kept for the scatter pass: the hash is computed once), per-page counts;
// long ids reserve their arena bytes here (any order: the slot carries the offset)
__global__ void __launch_bounds__(256)
    idset_count_kernel(const uint8_t *src, const uint64_t *off, const uint32_t *len, const uint8_t *sel, size_t n,
                       uint64_t n_pages, uint64_t arena_base, uint32_t *page_of, uint64_t *arena_at, uint32_t *page_count,
                       BuildStats *st) {
    const int lane = threadIdx.x & 31;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // (whole warps stay in the loop: the arena reservation below is a warp-wide step)
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const size_t i = i0 + lane;
        uint32_t page = NO_PAGE, need = 0;
        if (i < n && (sel ? sel[i] != 0 : true)) {
            const uint32_t L = len[i];
            if (L >= 1 && L <= IDSET_MAX_KEY) {
                uint64_t lo, hi, home;
                if (L <= IDSET_INLINE_MAX) {
                    inline_image_global(src + off[i], L, &lo, &hi);
                    home = inline_hash(lo, hi);
                } else {  // fingerprint of the word-wise hash; the key's bytes get (16-byte aligned) arena space
                    const uint64_t h = hash_words(GlobalKeyWords(src + off[i], L), L);
                    lo = 0x80ull | (h & ~0xFFull);
                    home = mix64(lo);
                    need = arena_padded(L);
                }
                page = (uint32_t)home_page(home, n_pages);
                atomicAdd(&page_count[page], 1u);
            }
        }
        // arena space for the warp's long ids: one atomic per warp (a cursor every key bumps is a single hot word:
        // 5 M long ids took 1.6 ms that way)
        if (__any_sync(0xffffffffu, need != 0)) {
            uint32_t inc = need;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += x;
            }
            unsigned long long base = 0;
            if (lane == 31) base = atomicAdd(&st->arena_used, (unsigned long long)inc);
            base = __shfl_sync(0xffffffffu, base, 31);
            if (need) arena_at[i] = arena_base + base + (inc - need);
        }
        if (i < n) page_of[i] = page;
    }
}

// long ids: key bytes into the arena (16-byte aligned entries), one thread per candidate, word-wise
__global__ void idset_arena_fill_kernel(const uint8_t *src, const uint64_t *off, const uint32_t *len, const uint32_t *page_of,
                                        const uint64_t *arena_at, size_t n, uint8_t *arena) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || page_of[i] == NO_PAGE) return;
    const uint32_t L = len[i];
    if (L <= IDSET_INLINE_MAX) return;
    const GlobalKeyWords word(src + off[i], L);
    uint32_t *dst = reinterpret_cast<uint32_t *>(arena + arena_at[i]);
    const uint32_t nw = (L + 3) >> 2;
    for (uint32_t k = 0; k < nw; k++) dst[k] = word(k);
}

// scatter: every selected candidate's slot image into its page's list.  Four candidates per thread and round: the
// cursor atomics return a value (a round trip to L2 each), four of them in flight hide most of it
__global__ void __launch_bounds__(256)
    idset_scatter_kernel(const uint8_t *src, const uint64_t *off, const uint32_t *len, size_t n, const uint32_t *page_of,
                         const uint64_t *arena_at, const uint64_t *page_start, uint32_t *page_cursor, ulonglong2 *out) {
    constexpr int K = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * K) {
        uint32_t page[K], rank[K];
        uint64_t lo[K], hi[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const size_t i = i0 + k * stride;
            page[k] = i < n ? page_of[i] : NO_PAGE;
            lo[k] = hi[k] = 0;
            if (page[k] != NO_PAGE) {
                const uint32_t L = len[i];
                if (L <= IDSET_INLINE_MAX) {
                    inline_image_global(src + off[i], L, &lo[k], &hi[k]);
                } else {
                    const uint64_t h = hash_words(GlobalKeyWords(src + off[i], L), L);
                    lo[k] = 0x80ull | (h & ~0xFFull);
                    hi[k] = (arena_at[i] << 24) | L;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) rank[k] = page[k] != NO_PAGE ? atomicAdd(&page_cursor[page[k]], 1u) : 0u;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (page[k] != NO_PAGE) out[page_start[page[k]] + rank[k]] = make_ulonglong2(lo[k], hi[k]);
    }
}

// build: one CTA assembles one page in shared memory and writes it out (every page is written, empty ones too).
// In shared memory slot q of bucket b sits at position (q + b) & 7 of the bucket's 128 bytes: a bucket is exactly 32
// banks wide, so without the rotation every thread's first probe (slot 0 of SOME bucket) would hit the same four banks.
struct PageSmem {
    ulonglong2 slot[IDSET_PAGE_BUCKETS * IDSET_BUCKET];
    uint32_t lock[IDSET_PAGE_BUCKETS];
    uint32_t fresh;
};
__device__ __forceinline__ uint32_t page_slot(uint32_t b, uint32_t q) { return b * IDSET_BUCKET + ((q + b) & (IDSET_BUCKET - 1)); }
// one slot image into the shared-memory page: one bucket at a time under its lock, so that the occupied slots stay a
// prefix of every bucket and duplicates are seen exactly.  arena_shift: added to a long id's arena offset (the sharded
// build concatenates the ranks' arenas)
__device__ __forceinline__ void page_insert(PageSmem *S, ulonglong2 r, const uint8_t *arena, uint64_t arena_shift, uint32_t *fresh) {
    const bool is_long = (r.x & 0xFF) == 0x80;
    if (is_long) r.y += arena_shift << 24;
    uint32_t b = (uint32_t)(slot_home(r.x, r.y) & (IDSET_PAGE_BUCKETS - 1));
    bool done = false;
    while (!done) {
        if (atomicCAS(&S->lock[b], 0u, 1u) == 0u) {
            __threadfence_block();
            bool full = true;
            for (uint32_t q = 0; q < IDSET_BUCKET; q++) {
                volatile ulonglong2 *sp = &S->slot[page_slot(b, q)];
                const uint64_t olo = sp->x, ohi = sp->y;
                if ((olo | ohi) == 0) {
                    sp->x = r.x;
                    sp->y = r.y;
                    (*fresh)++;
                    done = true;
                    full = false;
                    break;
                }
                if (olo == r.x) {
                    bool same;
                    if (!is_long) same = ohi == r.y;
                    else same = (ohi & 0xFFFFFFull) == (r.y & 0xFFFFFFull) &&
                                arena_equal_words(arena + (ohi >> 24),
                                                  GlobalKeyWords(arena + (r.y >> 24), (uint32_t)(r.y & 0xFFFFFFull)),
                                                  (uint32_t)(r.y & 0xFFFFFFull));
                    if (same) {
                        done = true;
                        full = false;
                        break;
                    }
                }
            }
            __threadfence_block();
            atomicExch(&S->lock[b], 0u);
            if (full) b = (b + 1) & (uint32_t)(IDSET_PAGE_BUCKETS - 1);
        }
    }
}

#ifndef SGPU_PAGE_THREADS
#define SGPU_PAGE_THREADS 256
#endif
constexpr int PAGE_THREADS = SGPU_PAGE_THREADS;
__global__ void __launch_bounds__(PAGE_THREADS)
    idset_page_kernel(const ulonglong2 *in, const uint32_t *in_count, const uint64_t *in_start, uint64_t n_pages,
                      const uint8_t *arena, Slot *table, BuildStats *st) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PageSmem *S = reinterpret_cast<PageSmem *>(smem_raw);
    constexpr uint32_t SLOTS = IDSET_PAGE_BUCKETS * IDSET_BUCKET;
    static_assert(IDSET_BUCKET == 8, "the bank rotation assumes 8 slots of 16 bytes per bucket");
    uint32_t fresh = 0;
    // (the next page's list is looked up while this one is assembled: two dependent loads less per page)
    uint64_t n_next = blockIdx.x < n_pages ? in_count[blockIdx.x] : 0, start_next = blockIdx.x < n_pages ? in_start[blockIdx.x] : 0;
    for (uint64_t page = blockIdx.x; page < n_pages; page += gridDim.x) {
        const uint64_t n = n_next;
        const ulonglong2 *recs = in + start_next;
        if (page + gridDim.x < n_pages) {
            n_next = in_count[page + gridDim.x];
            start_next = in_start[page + gridDim.x];
        }
        for (uint32_t i = threadIdx.x; i < SLOTS; i += blockDim.x) S->slot[i] = make_ulonglong2(0ull, 0ull);
        for (uint32_t i = threadIdx.x; i < IDSET_PAGE_BUCKETS; i += blockDim.x) S->lock[i] = 0;
        __syncthreads();
        for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) page_insert(S, recs[i], arena, 0, &fresh);
        __syncthreads();
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(table) + page * SLOTS;
        for (uint32_t i = threadIdx.x; i < SLOTS; i += blockDim.x) dst[i] = S->slot[page_slot(i >> 3, i & 7u)];
        __syncthreads();
    }
    // distinct keys: one shared-memory reduction and one global atomic per CTA
    if (threadIdx.x == 0) S->fresh = 0;
    __syncthreads();
    for (int d = 16; d; d >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(&S->fresh, fresh);
    __syncthreads();
    if (threadIdx.x == 0 && S->fresh) atomicAdd(&st->inserted, (unsigned long long)S->fresh);
}

// The same assembly from SEVERAL page-sorted lists (the sharded build: one list per rank).  A list is sorted by VIRTUAL
// page (2^log2_v of them, page = top bits of the home hash); a table of n_pages = 2^(log2_v - shift) real pages takes the
// records of virtual pages [p << shift, (p + 1) << shift) of every list -- one contiguous segment per list.
constexpr int MAX_PARTS = 8;
struct PartLists {
    const ulonglong2 *recs[MAX_PARTS];
    const uint64_t *vstart[MAX_PARTS];  // [V] exclusive starts, [V] = records of the list
    int n;
    uint32_t shift;
};
__global__ void __launch_bounds__(PAGE_THREADS)
    idset_page_multi_kernel(PartLists L, uint64_t n_pages, Slot *table, BuildStats *st) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PageSmem *S = reinterpret_cast<PageSmem *>(smem_raw);
    constexpr uint32_t SLOTS = IDSET_PAGE_BUCKETS * IDSET_BUCKET;
    uint32_t fresh = 0;
    for (uint64_t page = blockIdx.x; page < n_pages; page += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < SLOTS; i += blockDim.x) S->slot[i] = make_ulonglong2(0ull, 0ull);
        for (uint32_t i = threadIdx.x; i < IDSET_PAGE_BUCKETS; i += blockDim.x) S->lock[i] = 0;
        __syncthreads();
        for (int q = 0; q < L.n; q++) {
            const uint64_t a = L.vstart[q][page << L.shift], b = L.vstart[q][(page + 1) << L.shift];
            const ulonglong2 *recs = L.recs[q] + a;
            for (uint64_t i = threadIdx.x; i < b - a; i += blockDim.x) page_insert(S, recs[i], nullptr, 0, &fresh);
        }
        __syncthreads();
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(table) + page * SLOTS;
        for (uint32_t i = threadIdx.x; i < SLOTS; i += blockDim.x) dst[i] = S->slot[page_slot(i >> 3, i & 7u)];
        __syncthreads();
    }
    if (threadIdx.x == 0) S->fresh = 0;
    __syncthreads();
    for (int d = 16; d; d >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(&S->fresh, fresh);
    __syncthreads();
    if (threadIdx.x == 0 && S->fresh) atomicAdd(&st->inserted, (unsigned long long)S->fresh);
}

static sgpu_status page_kernel_attr(sgpu_ctx *c) {
    static bool attr_done[64] = {false};
    if (!attr_done[c->device & 63]) {
        SGPU_CUDA(cudaFuncSetAttribute(idset_page_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PageSmem)));
        SGPU_CUDA(cudaFuncSetAttribute(idset_page_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PageSmem)));
        attr_done[c->device & 63] = true;
    }
    return SGPU_OK;
}

// Sharded build, step 1 (every rank, on the candidates of ITS evidence shard): slot images grouped by virtual page into
// caller-owned buffers (symmetric memory: the other ranks pull them).  d_vstart: V exclusive starts, [V] = records,
// [V + 1] = flags (1: ids of 16 bytes and more / too long: not shardable, 2: the empty id was offered).
sgpu_status idset_partition(sgpu_ctx *c, const uint8_t *d_src, const uint64_t *d_off, const uint32_t *d_len, size_t n,
                            uint32_t log2_v, ulonglong2 *d_recs, size_t cap_recs, uint64_t *d_vstart, uint64_t flags) {
    cudaStream_t st = c->stream;
    const uint64_t V = 1ull << log2_v;
    if (n > cap_recs) return SGPU_ERR_CAPACITY;
    DevBuf<uint32_t> page_of, counts;
    DevBuf<uint64_t> arena_at;
    DevBuf<BuildStats> stats;
    SGPU_TRY(page_of.alloc(n ? n : 1, st));
    SGPU_TRY(counts.alloc(2 * V, st));
    SGPU_TRY(arena_at.alloc(1, st));
    SGPU_TRY(stats.alloc(1, st));
    SGPU_CUDA(cudaMemsetAsync(counts.p, 0, 2 * V * 4, st));
    SGPU_CUDA(cudaMemsetAsync(stats.p, 0, sizeof(BuildStats), st));
    c->h_pinned[48] = flags;
    SGPU_CUDA(cudaMemcpyAsync(d_vstart + V + 1, c->h_pinned + 48, 8, cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(ceil_div(n, (size_t)256), (size_t)c->sm_count * 16));
    if (n && !(flags & 1)) {
        idset_count_kernel<<<grid, 256, 0, st>>>(d_src, d_off, d_len, nullptr, n, V, 0, page_of.p, arena_at.p, counts.p, stats.p);
        SGPU_LAUNCH(c);
    }
    SGPU_TRY(exclusive_scan_u32_to_u64(c, counts.p, d_vstart, V, d_vstart + V));
    if (n && !(flags & 1)) {
        idset_scatter_kernel<<<grid, 256, 0, st>>>(d_src, d_off, d_len, n, page_of.p, arena_at.p, d_vstart, counts.p + V, d_recs);
        SGPU_LAUNCH(c);
    }
    SGPU_CUDA(cudaGetLastError());
    SGPU_CUDA(cudaStreamSynchronize(st));  // the lists are final when this returns: the caller signals the other ranks
    return SGPU_OK;
}

// Sharded build, step 2 (every rank, on ALL ranks' lists -- pulled into local memory or mapped peer memory)
sgpu_status idset_assemble(sgpu_ctx *c, int n_parts, const ulonglong2 *const *recs, const uint64_t *const *vstart,
                           uint32_t log2_v, sgpu_idset *s) {
    cudaStream_t st = c->stream;
    if (n_parts < 1 || n_parts > MAX_PARTS || log2_v > 30) return SGPU_ERR_INVALID_ARG;
    const uint64_t V = 1ull << log2_v;
    for (int q = 0; q < n_parts; q++)
        SGPU_CUDA(cudaMemcpyAsync(c->h_pinned + 2 * q, vstart[q] + V, 16, cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(cudaStreamSynchronize(st));
    uint64_t total = 0, flags = 0;
    for (int q = 0; q < n_parts; q++) {
        total += c->h_pinned[2 * q];
        flags |= c->h_pinned[2 * q + 1];
    }
    if (flags & 1) return SGPU_ERR_NOT_SHARDABLE;
    if (flags & 2) s->has_empty = true;
    if (total == 0) return SGPU_OK;
    // pages: a power of two (a merge of virtual pages), load <= 0.2 -- or up to 0.5 when the keys outgrow the virtual pages
    uint64_t need = idset_buckets_for(total) / IDSET_PAGE_BUCKETS, n_pages = 1;
    while (n_pages < need && n_pages < V) n_pages <<= 1;
    if (total * 2 > n_pages * IDSET_PAGE_BUCKETS * IDSET_BUCKET) return SGPU_ERR_NOT_SHARDABLE;
    uint32_t shift = 0;
    while ((n_pages << shift) < V) shift++;
    Slot *nt = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&nt, n_pages * IDSET_PAGE_BUCKETS * IDSET_BUCKET * sizeof(Slot), st);
    if (e != cudaSuccess) {
        set_cuda_error(e, __FILE__, __LINE__);
        return SGPU_ERR_NOMEM;
    }
    if (s->d_table) SGPU_CUDA(cudaFreeAsync(s->d_table, st));
    s->d_table = nt;
    s->n_buckets = n_pages * IDSET_PAGE_BUCKETS;
    DevBuf<BuildStats> stats;
    SGPU_TRY(stats.alloc(1, st));
    SGPU_CUDA(cudaMemsetAsync(stats.p, 0, sizeof(BuildStats), st));
    SGPU_TRY(page_kernel_attr(c));
    PartLists L;
    memset(&L, 0, sizeof(L));
    for (int q = 0; q < n_parts; q++) {
        L.recs[q] = recs[q];
        L.vstart[q] = vstart[q];
    }
    L.n = n_parts;
    L.shift = shift;
    const unsigned per_sm = (unsigned)std::min<size_t>(16, (size_t)(220 * 1024) / (sizeof(PageSmem) + 1024));
    const unsigned g3 = (unsigned)std::min<uint64_t>(n_pages, (uint64_t)c->sm_count * per_sm);
    idset_page_multi_kernel<<<g3, PAGE_THREADS, sizeof(PageSmem), st>>>(L, n_pages, s->d_table, stats.p);
    SGPU_LAUNCH(c);
    SGPU_CUDA(cudaGetLastError());
    BuildStats h;
    SGPU_TRY(read_u64s(c, stats.p, (uint64_t *)&h, sizeof(BuildStats) / 8));
    s->count = h.inserted;
    return SGPU_OK;
}

// Builds the WHOLE table of an empty set from the selected candidates (the table and the arena are allocated, n_buckets
// is set, nothing needs to be zeroed).  *ok = 0: not applicable (more pages than 32-bit page ids) -- the table is then
// zeroed and the caller inserts key by key.
sgpu_status idset_build_paged(sgpu_ctx *c, sgpu_idset *s, const uint8_t *d_src, const uint64_t *d_off, const uint32_t *d_len,
                              const uint8_t *d_sel, size_t n, uint64_t n_sel, uint64_t *inserted, uint64_t *arena_bytes,
                              int *ok) {
    *ok = 0;
    cudaStream_t st = c->stream;
    const uint64_t n_pages = s->n_buckets / IDSET_PAGE_BUCKETS;
    if (n_pages >= (uint64_t)NO_PAGE) {
        SGPU_CUDA(cudaMemsetAsync(s->d_table, 0, s->slots() * sizeof(Slot), st));
        return SGPU_OK;
    }
    DevBuf<uint32_t> page_of, counts;  // counts: [n_pages] counts, [n_pages] cursors
    DevBuf<uint64_t> arena_at, starts;
    DevBuf<ulonglong2> recs;
    DevBuf<BuildStats> stats;
    SGPU_TRY(page_of.alloc(n, st));
    SGPU_TRY(counts.alloc(2 * n_pages, st));
    SGPU_TRY(starts.alloc(n_pages, st));
    SGPU_TRY(recs.alloc(n_sel, st));
    SGPU_TRY(stats.alloc(1, st));
    const bool any_long = s->arena_cap > s->arena_used;
    SGPU_TRY(arena_at.alloc(any_long ? n : 1, st));
    SGPU_CUDA(cudaMemsetAsync(counts.p, 0, 2 * n_pages * 4, st));
    SGPU_CUDA(cudaMemsetAsync(stats.p, 0, sizeof(BuildStats), st));
    SGPU_TRY(page_kernel_attr(c));
    const unsigned grid = (unsigned)std::min<size_t>(ceil_div(n, (size_t)256), (size_t)c->sm_count * 16);
    idset_count_kernel<<<grid, 256, 0, st>>>(d_src, d_off, d_len, d_sel, n, n_pages, s->arena_used, page_of.p, arena_at.p,
                                             counts.p, stats.p);
    SGPU_LAUNCH(c);
    SGPU_TRY(exclusive_scan_u32_to_u64(c, counts.p, starts.p, n_pages, nullptr));
    if (any_long) {
        idset_arena_fill_kernel<<<(unsigned)ceil_div(n, (size_t)256), 256, 0, st>>>(d_src, d_off, d_len, page_of.p, arena_at.p, n,
                                                                                    s->d_arena);
        SGPU_LAUNCH(c);
    }
    idset_scatter_kernel<<<grid, 256, 0, st>>>(d_src, d_off, d_len, n, page_of.p, arena_at.p, starts.p, counts.p + n_pages,
                                               recs.p);
    SGPU_LAUNCH(c);
    // resident CTAs per SM by shared memory (a page + its locks), at most 16
    const unsigned per_sm = (unsigned)std::min<size_t>(16, (size_t)(220 * 1024) / (sizeof(PageSmem) + 1024));
    const unsigned g3 = (unsigned)std::min<uint64_t>(n_pages, (uint64_t)c->sm_count * per_sm);
    idset_page_kernel<<<g3, PAGE_THREADS, sizeof(PageSmem), st>>>(recs.p, counts.p, starts.p, n_pages, s->d_arena, s->d_table, stats.p);
    SGPU_LAUNCH(c);
    SGPU_CUDA(cudaGetLastError());
    BuildStats h;
    SGPU_TRY(read_u64s(c, stats.p, (uint64_t *)&h, sizeof(BuildStats) / 8));
    *inserted = h.inserted;
    *arena_bytes = h.arena_used;
    *ok = 1;
    return SGPU_OK;
}

}  // namespace sgpu

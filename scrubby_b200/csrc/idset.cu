// idset.cu -- build / grow / dump of the exact read-id set (see idset.cuh for the layout).
// Replaces HashSet::insert at alignment.rs:74,106, classifier.rs:284,322, utils.rs:264,275.
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "idset.cuh"

namespace sgpu {

struct InsertStats {
    unsigned long long inserted;   // new distinct ids
    unsigned long long has_empty;  // an empty id was offered
    unsigned long long too_long;   // an id >= 16 MiB was offered
    unsigned long long n_sel;      // selected candidates
    unsigned long long long_bytes; // bytes of selected candidates longer than 15
};

__device__ __forceinline__ unsigned __int128 pack128(uint64_t lo, uint64_t hi) {
    return ((unsigned __int128)hi << 64) | lo;
}

// per-candidate arena length (0 for inline / unselected) and candidate statistics
__global__ void idset_measure_kernel(const uint32_t *len, const uint8_t *sel, size_t n, uint32_t *arena_len,
                                     InsertStats *st) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t l = 0;
    bool s = false;
    if (i < n) {
        s = sel ? sel[i] != 0 : true;
        uint32_t L = len[i];
        l = (s && L > IDSET_INLINE_MAX && L <= IDSET_MAX_KEY) ? arena_padded(L) : 0;  // (entries are 16-byte aligned)
        arena_len[i] = l;
        if (s && L > IDSET_MAX_KEY) st->too_long = 1;
        if (s && L == 0) st->has_empty = 1;
    }
    // one pair of atomics per block (all blocks add to the same two words)
    __shared__ unsigned long long s_cnt, s_lb;
    if (threadIdx.x == 0) s_cnt = s_lb = 0;
    __syncthreads();
    unsigned long long cnt = s ? 1 : 0, lb = l;
    for (int d = 16; d; d >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        lb += __shfl_xor_sync(0xffffffffu, lb, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(&s_cnt, cnt);
        if (lb) atomicAdd(&s_lb, lb);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt) atomicAdd(&st->n_sel, s_cnt);
        if (s_lb) atomicAdd(&st->long_bytes, s_lb);
    }
}

__global__ void idset_arena_copy_kernel(const uint8_t *src, const uint64_t *off, const uint32_t *len, const uint32_t *arena_len,
                                        const uint64_t *arena_off, size_t n, uint8_t *arena, uint64_t arena_base) {
    // one warp per candidate
    size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n) return;
    if (arena_len[w] == 0) return;
    warp_copy(arena + arena_base + arena_off[w], src + off[w], len[w], lane);
}

// walks the probe sequence of (lo, hi) from its home bucket and claims the first empty slot with a 128-bit CAS;
// returns true when the key was new.  Slots are read first (an L2 load brings the 128-byte bucket in from DRAM
// without occupying the L2 atomic unit for the whole miss) and the CAS is issued only on a slot seen empty.
template <typename SameKey>
__device__ __forceinline__ bool idset_claim(Slot *table, uint64_t n_pages, uint64_t home, uint64_t lo, uint64_t hi,
                                            SameKey same_key) {
    const unsigned __int128 mine = pack128(lo, hi);
    uint64_t b = home_bucket(home, n_pages);
    while (true) {
        Slot *bp = table + b * IDSET_BUCKET;
        for (int q = 0; q < (int)IDSET_BUCKET; q++) {
            const ulonglong2 cur = __ldcg(reinterpret_cast<const ulonglong2 *>(bp + q));
            uint64_t olo = cur.x, ohi = cur.y;
            if ((olo | ohi) == 0) {
                unsigned __int128 old =
                    atomicCAS(reinterpret_cast<unsigned __int128 *>(bp + q), (unsigned __int128)0, mine);
                if (old == 0) return true;
                olo = (uint64_t)old;
                ohi = (uint64_t)(old >> 64);
            }
            if (olo == lo && same_key(ohi)) return false;  // duplicate
        }
        b = next_bucket(b);
    }
}

__global__ void idset_insert_kernel(Slot *table, uint64_t n_pages, const uint8_t *arena, uint64_t arena_base,
                                    const uint8_t *src, const uint64_t *off, const uint32_t *len, const uint8_t *sel,
                                    const uint64_t *arena_off, size_t n, InsertStats *st) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool fresh = false;
    if (i < n && (sel ? sel[i] != 0 : true)) {
        uint32_t L = len[i];
        if (L >= 1 && L <= IDSET_MAX_KEY) {
            const uint8_t *key = src + off[i];
            uint64_t lo, hi, home;
            key_image(key, L, &lo, &hi, &home);
            const bool is_inline = L <= IDSET_INLINE_MAX;
            if (!is_inline) hi = ((arena_base + arena_off[i]) << 24) | L;
            fresh = idset_claim(table, n_pages, home, lo, hi, [&](uint64_t ohi) {
                if (is_inline) return ohi == hi;
                return (ohi & 0xFFFFFFull) == L && bytes_equal(arena + (ohi >> 24), key, L);  // bytes verified
            });
        }
    }
    __shared__ unsigned int s_fresh;
    if (threadIdx.x == 0) s_fresh = 0;
    __syncthreads();
    unsigned b = __ballot_sync(0xffffffffu, fresh);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_fresh, (unsigned)__popc(b));
    __syncthreads();
    if (threadIdx.x == 0 && s_fresh) atomicAdd(&st->inserted, (unsigned long long)s_fresh);
}

// every stored key is distinct: no comparison needed, only an empty slot
__global__ void idset_rehash_kernel(const Slot *old_table, uint64_t old_slots, Slot *table, uint64_t n_pages) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_slots) return;
    Slot s = old_table[i];
    if ((s.lo | s.hi) == 0) return;
    idset_claim(table, n_pages, slot_home(s.lo, s.hi), s.lo, s.hi, [](uint64_t) { return false; });
}

// dump: per-slot byte length (id + '\n'), then flat copy
__global__ void idset_dump_len_kernel(const Slot *table, uint64_t cap, uint32_t *out_len) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    Slot s = table[i];
    uint32_t l = 0;
    if ((s.lo | s.hi) != 0) l = ((s.lo & 0xFF) == 0x80 ? (uint32_t)(s.hi & 0xFFFFFF) : (uint32_t)(s.lo & 0xFF)) + 1;
    out_len[i] = l;
}
__global__ void idset_dump_copy_kernel(const Slot *table, uint64_t cap, const uint8_t *arena,
                                       const uint64_t *out_off, uint8_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    Slot s = table[i];
    if ((s.lo | s.hi) == 0) return;
    uint8_t *o = out + out_off[i];
    if ((s.lo & 0xFF) == 0x80) {
        uint32_t L = (uint32_t)(s.hi & 0xFFFFFF);
        const uint8_t *k = arena + (s.hi >> 24);
        for (uint32_t b = 0; b < L; b++) o[b] = k[b];
        o[L] = '\n';
    } else {
        uint32_t L = (uint32_t)(s.lo & 0xFF);
        for (uint32_t b = 0; b < L; b++) {
            uint32_t j = b + 1;
            o[b] = (uint8_t)(j < 8 ? (s.lo >> (8 * j)) : (s.hi >> (8 * (j - 8))));
        }
        o[L] = '\n';
    }
}

sgpu_status idset_create(sgpu_ctx *c, sgpu_idset **out) {
    sgpu_idset *s = new (std::nothrow) sgpu_idset();
    if (!s) return SGPU_ERR_NOMEM;
    s->ctx = c;
    c->refs.fetch_add(1);
    s->device = c->device;
    *out = s;
    return SGPU_OK;
}

// idset_build.cu
sgpu_status idset_build_paged(sgpu_ctx *c, sgpu_idset *s, const uint8_t *d_src, const uint64_t *d_off, const uint32_t *d_len,
                              const uint8_t *d_sel, size_t n, uint64_t n_sel, uint64_t *inserted, uint64_t *arena_bytes,
                              int *ok);

// *zeroed = false: a NEW table was allocated and left uninitialised (the bulk build writes every page of it)
static sgpu_status idset_reserve(sgpu_ctx *c, sgpu_idset *s, uint64_t n_new, uint64_t new_long_bytes, bool may_skip_zero,
                                 bool *zeroed) {
    *zeroed = true;
    cudaStream_t st = c->stream;
    // sized exactly for its keys at the target load; a set that grows again (diff accumulators, evidence in several
    // calls) at least doubles, so the rehash cost stays amortised
    uint64_t need = idset_buckets_for(s->count + n_new);
    if (need > s->n_buckets) {
        if (s->n_buckets) need = std::max(need, 2 * s->n_buckets);
        Slot *nt = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&nt, need * IDSET_BUCKET * sizeof(Slot), st);  // (256-byte aligned)
        if (e != cudaSuccess) {
            set_cuda_error(e, __FILE__, __LINE__);
            return SGPU_ERR_NOMEM;
        }
        if (may_skip_zero && s->count == 0) *zeroed = false;
        else SGPU_CUDA(cudaMemsetAsync(nt, 0, need * IDSET_BUCKET * sizeof(Slot), st));
        if (s->n_buckets && s->count) {
            idset_rehash_kernel<<<(unsigned)ceil_div(s->slots(), 256), 256, 0, st>>>(s->d_table, s->slots(), nt,
                                                                                     need / IDSET_PAGE_BUCKETS);
            SGPU_LAUNCH(c);
        }
        if (s->d_table) SGPU_CUDA(cudaFreeAsync(s->d_table, st));
        s->d_table = nt;
        s->n_buckets = need;
    }
    if (s->arena_used + new_long_bytes > s->arena_cap) {
        uint64_t ncap = std::max<uint64_t>(s->arena_cap * 2, s->arena_used + new_long_bytes);
        ncap = (ncap + 255) & ~255ull;
        uint8_t *na = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&na, ncap, st);
        if (e != cudaSuccess) {
            set_cuda_error(e, __FILE__, __LINE__);
            return SGPU_ERR_NOMEM;
        }
        if (s->arena_used)
            SGPU_CUDA(cudaMemcpyAsync(na, s->d_arena, s->arena_used, cudaMemcpyDeviceToDevice, st));
        if (s->d_arena) SGPU_CUDA(cudaFreeAsync(s->d_arena, st));
        s->d_arena = na;
        s->arena_cap = ncap;
    }
    return SGPU_OK;
}

sgpu_status idset_insert_spans(sgpu_ctx *c, sgpu_idset *s, const uint8_t *d_src, const uint64_t *d_off,
                               const uint32_t *d_len, const uint8_t *d_sel, size_t n, const SpanStats *known) {
    if (n == 0) return SGPU_OK;
    cudaStream_t st = c->stream;
    static const uint64_t bulk_min = getenv("SGPU_IDSET_BULK_MIN") ? (uint64_t)atoll(getenv("SGPU_IDSET_BULK_MIN")) : 32768;
    DevBuf<uint32_t> alen;
    DevBuf<uint64_t> aoff;
    DevBuf<InsertStats> stats;
    InsertStats h;
    memset(&h, 0, sizeof(h));
    // the measuring pass (and its per-candidate arena lengths) is only needed by the key-by-key path, or when the
    // producer of the spans has not counted them itself
    const bool skip_measure = known && s->count == 0 && known->n_sel >= bulk_min && !known->too_long;
    SGPU_TRY(stats.alloc(1, st));
    SGPU_CUDA(cudaMemsetAsync(stats.p, 0, sizeof(InsertStats), st));
    unsigned grid = (unsigned)ceil_div(n, 256);
    auto measure = [&]() -> sgpu_status {
        SGPU_TRY(alen.alloc(n, st));
        SGPU_TRY(aoff.alloc(n, st));
        idset_measure_kernel<<<grid, 256, 0, st>>>(d_len, d_sel, n, alen.p, stats.p);
        SGPU_LAUNCH(c);
        return read_u64s(c, stats.p, (uint64_t *)&h, sizeof(InsertStats) / 8);
    };
    if (skip_measure) {
        h.n_sel = known->n_sel;
        h.long_bytes = known->long_bytes;
        h.has_empty = known->has_empty;
    } else {
        SGPU_TRY(measure());
    }
    if (h.too_long) return SGPU_ERR_KEY_TOO_LONG;
    if (h.has_empty) s->has_empty = true;
    if (h.n_sel == 0) return SGPU_OK;
    // a whole evidence file into an empty set: the bulk build (partition by page, assemble every page in shared memory,
    // write the table once); everything else key by key with a 128-bit CAS
    const bool bulk = s->count == 0 && h.n_sel >= bulk_min;
    bool zeroed = true;
    SGPU_TRY(idset_reserve(c, s, h.n_sel, h.long_bytes, bulk, &zeroed));
    if (bulk) {
        uint64_t inserted = 0, arena_bytes = 0;
        int ok = 0;
        SGPU_TRY(idset_build_paged(c, s, d_src, d_off, d_len, d_sel, n, h.n_sel, &inserted, &arena_bytes, &ok));
        if (ok) {
            s->count += inserted;
            s->arena_used += arena_bytes;
            return SGPU_OK;
        }
        zeroed = true;  // (the build that did not apply left a zeroed table behind)
        if (skip_measure) {
            SGPU_CUDA(cudaMemsetAsync(stats.p, 0, sizeof(InsertStats), st));
            SGPU_TRY(measure());
        }
    }
    if (h.long_bytes) {
        SGPU_TRY(exclusive_scan_u32_to_u64(c, alen.p, aoff.p, n, nullptr));
        idset_arena_copy_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(d_src, d_off, d_len, alen.p, aoff.p, n,
                                                                                 s->d_arena, s->arena_used);
        SGPU_LAUNCH(c);
    }
    if (!zeroed) SGPU_CUDA(cudaMemsetAsync(s->d_table, 0, s->slots() * sizeof(Slot), st));
    idset_insert_kernel<<<grid, 256, 0, st>>>(s->d_table, s->n_buckets / IDSET_PAGE_BUCKETS, s->d_arena, s->arena_used, d_src, d_off,
                                              d_len, d_sel, aoff.p, n, stats.p);
    SGPU_LAUNCH(c);
    SGPU_CUDA(cudaGetLastError());
    SGPU_TRY(read_u64s(c, stats.p, (uint64_t *)&h, sizeof(InsertStats) / 8));
    s->count += h.inserted;
    s->arena_used += h.long_bytes;  // duplicates waste arena space; bounded by the candidates
    return SGPU_OK;
}

}  // namespace sgpu

using namespace sgpu;

extern "C" {

sgpu_status sgpu_idset_new(sgpu_ctx *c, sgpu_idset **out) {
    if (!c || !out) return SGPU_ERR_INVALID_ARG;
    return idset_create(c, out);
}

uint64_t sgpu_idset_len(const sgpu_idset *s) { return s ? s->count + (s->has_empty ? 1 : 0) : 0; }

void sgpu_idset_free(sgpu_idset *s) {
    if (!s) return;
    // the set holds a reference on its context, so the stream is still alive here
    sgpu_ctx *c = s->ctx;
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != s->device) cudaSetDevice(s->device);
    if (s->d_table) cudaFreeAsync(s->d_table, c->stream);
    if (s->d_arena) cudaFreeAsync(s->d_arena, c->stream);
    if (dev >= 0 && dev != s->device) cudaSetDevice(dev);
    ctx_release(c);
    delete s;
}

void sgpu_free(void *p) { free(p); }

sgpu_status sgpu_idset_from_ids(sgpu_ctx *c, const char *const *ids, const size_t *lens, size_t n,
                                sgpu_idset **out) {
    if (!c || !out || (n && (!ids || !lens))) return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    sgpu_idset *s = nullptr;
    SGPU_TRY(idset_create(c, &s));
    std::vector<uint64_t> off(n);
    std::vector<uint32_t> len(n);
    size_t total = 0;
    for (size_t i = 0; i < n; i++) {
        off[i] = total;
        if (lens[i] > 0xFFFFFFFFull) {
            sgpu_idset_free(s);
            return SGPU_ERR_KEY_TOO_LONG;
        }
        len[i] = (uint32_t)lens[i];
        total += lens[i];
    }
    std::vector<uint8_t> flat(total ? total : 1);
    for (size_t i = 0; i < n; i++)
        if (lens[i]) memcpy(flat.data() + off[i], ids[i], lens[i]);
    sgpu_status rc = SGPU_OK;
    if (n) {
        DevBuf<uint8_t> d_flat;
        DevBuf<uint64_t> d_off;
        DevBuf<uint32_t> d_len;
        cudaStream_t st = c->stream;
        rc = d_flat.alloc(flat.size(), st);
        if (rc == SGPU_OK) rc = d_off.alloc(n, st);
        if (rc == SGPU_OK) rc = d_len.alloc(n, st);
        if (rc == SGPU_OK) {
            cudaMemcpyAsync(d_flat.p, flat.data(), flat.size(), cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_off.p, off.data(), n * 8, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_len.p, len.data(), n * 4, cudaMemcpyHostToDevice, st);
            rc = idset_insert_spans(c, s, d_flat.p, d_off.p, d_len.p, nullptr, n);
            cudaStreamSynchronize(st);  // host vectors go out of scope
        }
    }
    if (rc != SGPU_OK) {
        sgpu_idset_free(s);
        return rc;
    }
    *out = s;
    return SGPU_OK;
}

__global__ void idset_contains_kernel(IdSetView v, const uint8_t *key, uint32_t len, unsigned long long *res) {
    *res = idset_contains(v, key, len) ? 1 : 0;
}

sgpu_status sgpu_idset_contains(sgpu_ctx *c, const sgpu_idset *s, const char *id, size_t len, int *found) {
    if (!c || !s || !found || (len && !id)) return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    if (len == 0) {
        *found = s->has_empty;
        return SGPU_OK;
    }
    if (len > IDSET_MAX_KEY || s->n_buckets == 0) {
        *found = 0;
        return SGPU_OK;
    }
    DevBuf<uint8_t> k;
    DevBuf<uint64_t> r;
    SGPU_TRY(k.alloc(len, c->stream));
    SGPU_TRY(r.alloc(1, c->stream));
    SGPU_CUDA(cudaMemcpyAsync(k.p, id, len, cudaMemcpyHostToDevice, c->stream));
    idset_contains_kernel<<<1, 1, 0, c->stream>>>(view_of(s), k.p, (uint32_t)len, (unsigned long long *)r.p);
    SGPU_LAUNCH(c);
    uint64_t h;
    SGPU_TRY(read_u64s(c, r.p, &h, 1));
    *found = (int)h;
    return SGPU_OK;
}

sgpu_status sgpu_idset_dump(sgpu_ctx *c, const sgpu_idset *s, uint8_t **out, size_t *n) {
    if (!c || !s || !out || !n) return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    std::vector<uint8_t> flat;
    if (s->n_buckets && s->count) {
        DevBuf<uint32_t> len;
        DevBuf<uint64_t> off, total;
        SGPU_TRY(len.alloc(s->slots(), st));
        SGPU_TRY(off.alloc(s->slots(), st));
        SGPU_TRY(total.alloc(1, st));
        unsigned grid = (unsigned)ceil_div(s->slots(), 256);
        idset_dump_len_kernel<<<grid, 256, 0, st>>>(s->d_table, s->slots(), len.p);
        SGPU_LAUNCH(c);
        SGPU_TRY(exclusive_scan_u32_to_u64(c, len.p, off.p, s->slots(), total.p));
        uint64_t bytes;
        SGPU_TRY(read_u64s(c, total.p, &bytes, 1));
        DevBuf<uint8_t> d_out;
        SGPU_TRY(d_out.alloc(bytes, st));
        idset_dump_copy_kernel<<<grid, 256, 0, st>>>(s->d_table, s->slots(), s->d_arena, off.p, d_out.p);
        SGPU_LAUNCH(c);
        flat.resize(bytes);
        SGPU_CUDA(cudaMemcpyAsync(flat.data(), d_out.p, bytes, cudaMemcpyDeviceToHost, st));
        SGPU_CUDA(cudaStreamSynchronize(st));
    }
    // host stage: order the ids (the reference's TSV order is HashSet-random; we emit sorted)
    std::vector<std::pair<const uint8_t *, size_t>> keys;
    keys.reserve(s->count + 1);
    size_t pos = 0;
    while (pos < flat.size()) {
        const uint8_t *b = flat.data() + pos;
        const uint8_t *e = (const uint8_t *)memchr(b, '\n', flat.size() - pos);
        size_t l = (size_t)(e - b);
        keys.emplace_back(b, l);
        pos += l + 1;
    }
    static const uint8_t empty = 0;
    if (s->has_empty) keys.emplace_back(&empty, 0);
    std::sort(keys.begin(), keys.end(), [](const auto &a, const auto &b) {
        size_t m = std::min(a.second, b.second);
        int r = m ? memcmp(a.first, b.first, m) : 0;
        return r != 0 ? r < 0 : a.second < b.second;
    });
    size_t total = flat.size() + (s->has_empty ? 1 : 0);
    uint8_t *o = (uint8_t *)malloc(total ? total : 1);
    if (!o) return SGPU_ERR_NOMEM;
    size_t w = 0;
    for (auto &k : keys) {
        if (k.second) memcpy(o + w, k.first, k.second);
        w += k.second;
        o[w++] = '\n';
    }
    *out = o;
    *n = w;
    return SGPU_OK;
}

// the set's keys as an unsorted one-column list ("id\n" per key) in a caller-owned DEVICE buffer: the exchange
// format of the multi-GPU diff (ReadDifference::get_difference's HashSets, utils.rs:251-283, united across ranks by
// an all-gather of these lists and sgpu_idset_from_txt_dev on the concatenation)
sgpu_status sgpu_idset_keys_dev(sgpu_ctx *c, const sgpu_idset *s, uint8_t *d_out, size_t cap, size_t *n) {
    if (!c || !s || !n) return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    uint64_t bytes = 0;
    DevBuf<uint32_t> len;
    DevBuf<uint64_t> off, total;
    const unsigned grid = (unsigned)ceil_div(s->slots(), 256);
    if (s->n_buckets && s->count) {
        SGPU_TRY(len.alloc(s->slots(), st));
        SGPU_TRY(off.alloc(s->slots(), st));
        SGPU_TRY(total.alloc(1, st));
        idset_dump_len_kernel<<<grid, 256, 0, st>>>(s->d_table, s->slots(), len.p);
        SGPU_LAUNCH(c);
        SGPU_TRY(exclusive_scan_u32_to_u64(c, len.p, off.p, s->slots(), total.p));
        SGPU_TRY(read_u64s(c, total.p, &bytes, 1));
    }
    *n = (size_t)bytes + (s->has_empty ? 1 : 0);
    if (*n == 0) return SGPU_OK;
    if (!d_out || cap < *n) return SGPU_ERR_CAPACITY;  // *n tells the size to come back with
    if (bytes) {
        idset_dump_copy_kernel<<<grid, 256, 0, st>>>(s->d_table, s->slots(), s->d_arena, off.p, d_out);
        SGPU_LAUNCH(c);
    }
    if (s->has_empty) SGPU_CUDA(cudaMemsetAsync(d_out + bytes, '\n', 1, st));  // the empty id: a blank line
    SGPU_CUDA(cudaStreamSynchronize(st));
    return SGPU_OK;
}

sgpu_status sgpu_idset_export(const sgpu_idset *s, sgpu_idset_image *img) {
    if (!s || !img) return SGPU_ERR_INVALID_ARG;
    img->d_table = s->d_table;
    img->table_bytes = s->slots() * sizeof(Slot);
    img->d_arena = s->d_arena;
    img->arena_bytes = s->arena_used;
    img->capacity = s->slots();
    img->count = s->count;
    img->has_empty = s->has_empty ? 1 : 0;
    return SGPU_OK;
}

sgpu_status sgpu_idset_import(sgpu_ctx *c, const sgpu_idset_image *img, sgpu_idset **out) {
    if (!c || !img || !out) return SGPU_ERR_INVALID_ARG;
    if (img->capacity % (IDSET_BUCKET * IDSET_PAGE_BUCKETS)) return SGPU_ERR_INVALID_ARG;
    if (img->table_bytes != img->capacity * sizeof(Slot)) return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    sgpu_idset *s = nullptr;
    SGPU_TRY(idset_create(c, &s));
    cudaStream_t st = c->stream;
    if (img->capacity) {
        if (cudaMallocAsync((void **)&s->d_table, img->table_bytes, st) != cudaSuccess) {
            sgpu_idset_free(s);
            return SGPU_ERR_NOMEM;
        }
        cudaMemcpyAsync(s->d_table, img->d_table, img->table_bytes, cudaMemcpyDeviceToDevice, st);
        s->n_buckets = img->capacity / IDSET_BUCKET;
    }
    if (img->arena_bytes) {
        uint64_t cap = (img->arena_bytes + 255) & ~255ull;
        if (cudaMallocAsync((void **)&s->d_arena, cap, st) != cudaSuccess) {
            sgpu_idset_free(s);
            return SGPU_ERR_NOMEM;
        }
        cudaMemcpyAsync(s->d_arena, img->d_arena, img->arena_bytes, cudaMemcpyDeviceToDevice, st);
        s->arena_cap = cap;
        s->arena_used = img->arena_bytes;
    }
    s->count = img->count;
    s->has_empty = img->has_empty != 0;
    SGPU_CUDA(cudaGetLastError());
    *out = s;
    return SGPU_OK;
}

}  // extern "C"

// idset.cuh -- device-side encoding and probe of the exact read-id set.
//
// Replaces std HashSet<String> (SipHash-1-3 + SwissTable) at alignment.rs:62,91,
// classifier.rs:135,274 and utils.rs:251,257, and `read_ids.contains(&id)` at
// cleaner.rs:747,751.
//
// Layout in HBM: open addressing, 16-byte slots in 64-byte BUCKETS of four; a key's probe sequence starts
// at the first slot of its home bucket and runs linearly from there (no deletions, so the occupied slots
// of a bucket are always a prefix of it).  Load factor <= 0.2: a lookup is then decided by the home
// bucket alone -- four independent 16-byte loads, one DRAM burst -- in > 99 % of the cases, which is
// what keeps a warp of 32 lookups to a single memory round trip (with slot-granular homes and load 0.5
// the longest of 32 chains is 5-6 dependent loads).
//   empty  : lo == 0 && hi == 0
//   inline : ids of 1..15 bytes live IN the slot: byte0 = len, bytes 1..15 = id (zero padded).
//            One 16-byte load and a 128-bit compare decide membership exactly.
//   long   : ids of >= 16 bytes: byte0 = 0x80, bytes 1..7 = top 56 bits of a 64-bit hash,
//            hi = (arena offset << 24) | len.  A fingerprint hit is verified against the
//            id bytes in the arena, so there are no false positives.
// A bucket is two 32-byte sectors of one 64-byte DRAM burst.
#pragma once
#include "common.cuh"

namespace sgpu {

constexpr uint32_t IDSET_INLINE_MAX = 15;
#ifndef SGPU_IDSET_BUCKET
#define SGPU_IDSET_BUCKET 4
#endif
#ifndef SGPU_IDSET_INV_LOAD
#define SGPU_IDSET_INV_LOAD 5
#endif
constexpr uint64_t IDSET_BUCKET = SGPU_IDSET_BUCKET;      // slots per bucket
constexpr uint64_t IDSET_INV_LOAD = SGPU_IDSET_INV_LOAD;  // capacity >= IDSET_INV_LOAD * keys
constexpr uint64_t IDSET_MAX_KEY = (1ull << 24) - 1;

struct IdSetView {
    const Slot *table;
    uint64_t mask;  // capacity - 1 (capacity == 0 -> table == nullptr)
    const uint8_t *arena;
};

#ifdef __CUDACC__

// first slot of the home bucket of a key whose hash is h
__device__ __forceinline__ uint64_t home_slot(uint64_t h, uint64_t mask) { return h & mask & ~(IDSET_BUCKET - 1); }

__device__ __forceinline__ uint64_t hash_bytes(const uint8_t *p, uint32_t len) {
    uint64_t h = 0x9E3779B97F4A7C15ULL ^ ((uint64_t)len * 0xD6E8FEB86659FD93ULL);
    uint32_t i = 0;
    for (; i + 8 <= len; i += 8) {
        uint64_t w = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) w |= (uint64_t)p[i + b] << (8 * b);
        h = (h ^ w) * 0xFF51AFD7ED558CCDULL;
        h ^= h >> 29;
    }
    uint64_t w = 0;
    for (uint32_t b = 0; i + b < len; b++) w |= (uint64_t)p[i + b] << (8 * b);
    h = (h ^ w) * 0xC4CEB9FE1A85EC53ULL;
    return mix64(h);
}

// builds the slot image of a key and its home index hash.  For long keys `hi` is left 0
// (the caller fills offset/len when inserting).
__device__ __forceinline__ void key_image(const uint8_t *p, uint32_t len, uint64_t *lo, uint64_t *hi,
                                          uint64_t *home) {
    if (len <= IDSET_INLINE_MAX) {
        uint64_t a = len, b = 0;
#pragma unroll
        for (uint32_t i = 0; i < 7; i++)
            if (i < len) a |= (uint64_t)p[i] << (8 * (i + 1));
#pragma unroll
        for (uint32_t i = 7; i < 15; i++)
            if (i < len) b |= (uint64_t)p[i] << (8 * (i - 7));
        *lo = a;
        *hi = b;
        *home = mix64(a ^ mix64(b + 0x9E3779B97F4A7C15ULL));
    } else {
        uint64_t h = hash_bytes(p, len);
        *lo = 0x80ull | (h & ~0xFFull);
        *hi = 0;
        *home = h >> 8;
    }
}

// home index recomputed from a stored slot (rehash without touching key bytes)
__device__ __forceinline__ uint64_t slot_home(uint64_t lo, uint64_t hi) {
    if ((lo & 0xFF) == 0x80) return lo >> 8;
    return mix64(lo ^ mix64(hi + 0x9E3779B97F4A7C15ULL));
}

__device__ __forceinline__ bool bytes_equal(const uint8_t *a, const uint8_t *b, uint32_t n) {
    for (uint32_t i = 0; i < n; i++)
        if (a[i] != b[i]) return false;
    return true;
}

__device__ __forceinline__ Slot load_slot(const Slot *p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    Slot s;
    s.lo = v.x;
    s.hi = v.y;
    return s;
}

// exact membership of key[0..len), len >= 1
__device__ __forceinline__ bool idset_contains(const IdSetView &v, const uint8_t *key, uint32_t len) {
    if (v.table == nullptr || len > IDSET_MAX_KEY) return false;
    uint64_t lo, hi, home;
    key_image(key, len, &lo, &hi, &home);
    uint64_t idx = home_slot(home, v.mask);
    const bool is_inline = len <= IDSET_INLINE_MAX;
    while (true) {
        Slot s = load_slot(v.table + idx);
        if ((s.lo | s.hi) == 0) return false;
        if (s.lo == lo) {
            if (is_inline) {
                if (s.hi == hi) return true;
            } else if ((s.hi & 0xFFFFFFull) == len && bytes_equal(v.arena + (s.hi >> 24), key, len)) {
                return true;
            }
        }
        idx = (idx + 1) & v.mask;
    }
}

// exact membership of an INLINE key (1..15 bytes) whose slot image (lo, hi) the caller has already built
// (byte0 = len, bytes 1..15 = id, zero padded) -- identical to key_image() + idset_contains()
__device__ __forceinline__ bool idset_contains_inline(const IdSetView &v, uint64_t lo, uint64_t hi) {
    if (v.table == nullptr) return false;
    uint64_t idx = home_slot(mix64(lo ^ mix64(hi + 0x9E3779B97F4A7C15ULL)), v.mask);
    while (true) {
        const Slot s = load_slot(v.table + idx);
        if ((s.lo | s.hi) == 0) return false;
        if (s.lo == lo && s.hi == hi) return true;
        idx = (idx + 1) & v.mask;
    }
}

#endif  // __CUDACC__

static inline IdSetView view_of(const sgpu_idset *s) {
    IdSetView v;
    v.table = s && s->capacity ? s->d_table : nullptr;
    v.mask = s && s->capacity ? s->capacity - 1 : 0;
    v.arena = s ? s->d_arena : nullptr;
    return v;
}

}  // namespace sgpu

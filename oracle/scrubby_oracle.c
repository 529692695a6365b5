/*
 * scrubby_oracle.c -- CPU restatement of the scrubby depletion hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see scrubby_oracle.h).  PARITY UNPINNED: no
 * reference test, golden vector or buildable reference exists for this path.
 *
 * Each function cites the reference file:line (under /root/reference/src) it
 * follows.  FASTQ framing and re-serialisation live in the un-vendored crate
 * needletail 0.5.1 (Cargo.toml:32); the rules restated here are that crate's
 * published fastq reader (`Reader::next/find/validate/check_end`,
 * `BufferPosition::{id,seq,qual}`, `trim_cr`, `find_line_ending`) and writer
 * (`write_fastq`), anchored on the reference's call sites utils.rs:377-383,
 * cleaner.rs:742-754 and utils.rs:256-283.
 *
 * Structured like the reference for the CPU-baseline timing: one sequential
 * pass per file, first-whitespace-token id, exact string set hashed with
 * SipHash-1-3 (Rust's default hasher), output appended record by record.
 */
#include "scrubby_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* exact string set: open addressing + SipHash-1-3 (std HashSet<String>) */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t hash;
    uint64_t off; /* offset into arena */
    uint32_t len;
    uint32_t used;
} orc_slot;

struct orc_set {
    orc_slot *slots;
    size_t cap; /* power of two */
    size_t count;
    uint8_t *arena;
    size_t arena_len, arena_cap;
};

#define ROTL(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND        \
    do {                \
        v0 += v1;       \
        v1 = ROTL(v1, 13); \
        v1 ^= v0;       \
        v0 = ROTL(v0, 32); \
        v2 += v3;       \
        v3 = ROTL(v3, 16); \
        v3 ^= v2;       \
        v0 += v3;       \
        v3 = ROTL(v3, 21); \
        v3 ^= v0;       \
        v2 += v1;       \
        v1 = ROTL(v1, 17); \
        v1 ^= v2;       \
        v2 = ROTL(v2, 32); \
    } while (0)

static uint64_t siphash13(const uint8_t *in, size_t inlen) {
    const uint64_t k0 = 0x0706050403020100ULL, k1 = 0x0f0e0d0c0b0a0908ULL;
    uint64_t v0 = 0x736f6d6570736575ULL ^ k0, v1 = 0x646f72616e646f6dULL ^ k1;
    uint64_t v2 = 0x6c7967656e657261ULL ^ k0, v3 = 0x7465646279746573ULL ^ k1;
    const uint8_t *end = in + (inlen & ~(size_t)7);
    uint64_t b = ((uint64_t)inlen) << 56;
    for (; in != end; in += 8) {
        uint64_t m;
        memcpy(&m, in, 8);
        v3 ^= m;
        SIPROUND;
        v0 ^= m;
    }
    switch (inlen & 7) {
    case 7: b |= ((uint64_t)in[6]) << 48; /* fallthrough */
    case 6: b |= ((uint64_t)in[5]) << 40; /* fallthrough */
    case 5: b |= ((uint64_t)in[4]) << 32; /* fallthrough */
    case 4: b |= ((uint64_t)in[3]) << 24; /* fallthrough */
    case 3: b |= ((uint64_t)in[2]) << 16; /* fallthrough */
    case 2: b |= ((uint64_t)in[1]) << 8;  /* fallthrough */
    case 1: b |= ((uint64_t)in[0]);
    default: break;
    }
    v3 ^= b;
    SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND;
    SIPROUND;
    SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

orc_set *orc_set_new(void) {
    orc_set *s = (orc_set *)calloc(1, sizeof(orc_set));
    s->cap = 1024;
    s->slots = (orc_slot *)calloc(s->cap, sizeof(orc_slot));
    s->arena_cap = 1 << 16;
    s->arena = (uint8_t *)malloc(s->arena_cap);
    return s;
}

void orc_set_free(orc_set *s) {
    if (!s) return;
    free(s->slots);
    free(s->arena);
    free(s);
}

void orc_free(void *p) { free(p); }

static void orc_set_grow(orc_set *s) {
    size_t ncap = s->cap * 2;
    orc_slot *ns = (orc_slot *)calloc(ncap, sizeof(orc_slot));
    for (size_t i = 0; i < s->cap; i++) {
        if (!s->slots[i].used) continue;
        size_t j = s->slots[i].hash & (ncap - 1);
        while (ns[j].used) j = (j + 1) & (ncap - 1);
        ns[j] = s->slots[i];
    }
    free(s->slots);
    s->slots = ns;
    s->cap = ncap;
}

int orc_set_contains(const orc_set *s, const uint8_t *key, size_t len) {
    uint64_t h = siphash13(key, len);
    size_t j = h & (s->cap - 1);
    while (s->slots[j].used) {
        const orc_slot *sl = &s->slots[j];
        if (sl->hash == h && sl->len == len && memcmp(s->arena + sl->off, key, len) == 0) return 1;
        j = (j + 1) & (s->cap - 1);
    }
    return 0;
}

void orc_set_insert(orc_set *s, const uint8_t *key, size_t len) {
    uint64_t h = siphash13(key, len);
    size_t j = h & (s->cap - 1);
    while (s->slots[j].used) {
        const orc_slot *sl = &s->slots[j];
        if (sl->hash == h && sl->len == len && memcmp(s->arena + sl->off, key, len) == 0) return;
        j = (j + 1) & (s->cap - 1);
    }
    if (s->arena_len + len > s->arena_cap) {
        while (s->arena_len + len > s->arena_cap) s->arena_cap *= 2;
        s->arena = (uint8_t *)realloc(s->arena, s->arena_cap);
    }
    memcpy(s->arena + s->arena_len, key, len);
    s->slots[j].hash = h;
    s->slots[j].off = s->arena_len;
    s->slots[j].len = (uint32_t)len;
    s->slots[j].used = 1;
    s->arena_len += len;
    s->count++;
    if (s->count * 2 > s->cap) orc_set_grow(s);
}

uint64_t orc_set_len(const orc_set *s) { return s->count; }

typedef struct {
    const uint8_t *p;
    size_t len;
} orc_span;

static int span_cmp(const void *a, const void *b) {
    const orc_span *x = (const orc_span *)a, *y = (const orc_span *)b;
    size_t m = x->len < y->len ? x->len : y->len;
    int c = m ? memcmp(x->p, y->p, m) : 0;
    if (c) return c;
    return (x->len > y->len) - (x->len < y->len);
}

int orc_set_dump_sorted(const orc_set *s, uint8_t **out, size_t *n) {
    orc_span *v = (orc_span *)malloc(sizeof(orc_span) * (s->count + 1));
    size_t k = 0, total = 0;
    for (size_t i = 0; i < s->cap; i++) {
        if (!s->slots[i].used) continue;
        v[k].p = s->arena + s->slots[i].off;
        v[k].len = s->slots[i].len;
        total += v[k].len + 1;
        k++;
    }
    qsort(v, k, sizeof(orc_span), span_cmp);
    uint8_t *o = (uint8_t *)malloc(total + 1);
    size_t w = 0;
    for (size_t i = 0; i < k; i++) {
        memcpy(o + w, v[i].p, v[i].len);
        w += v[i].len;
        o[w++] = '\n';
    }
    free(v);
    *out = o;
    *n = w;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* UTF-8 and Unicode White_Space (Rust core::str)                      */
/* ------------------------------------------------------------------ */

/* Rust std::str::from_utf8 acceptance (RFC 3629: no overlongs, no surrogates, <= U+10FFFF). */
static int utf8_valid(const uint8_t *s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) {
            i++;
        } else if (c >= 0xC2 && c <= 0xDF) {
            if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return 0;
            i += 2;
        } else if (c >= 0xE0 && c <= 0xEF) {
            if (i + 2 >= n) return 0;
            uint8_t c1 = s[i + 1], c2 = s[i + 2];
            uint8_t lo = 0x80, hi = 0xBF;
            if (c == 0xE0) lo = 0xA0;
            if (c == 0xED) hi = 0x9F;
            if (c1 < lo || c1 > hi || (c2 & 0xC0) != 0x80) return 0;
            i += 3;
        } else if (c >= 0xF0 && c <= 0xF4) {
            if (i + 3 >= n) return 0;
            uint8_t c1 = s[i + 1], c2 = s[i + 2], c3 = s[i + 3];
            uint8_t lo = 0x80, hi = 0xBF;
            if (c == 0xF0) lo = 0x90;
            if (c == 0xF4) hi = 0x8F;
            if (c1 < lo || c1 > hi || (c2 & 0xC0) != 0x80 || (c3 & 0xC0) != 0x80) return 0;
            i += 4;
        } else {
            return 0;
        }
    }
    return 1;
}

/* char::is_whitespace: the Unicode White_Space property. */
static int is_ws_cp(uint32_t cp) {
    if (cp >= 0x09 && cp <= 0x0D) return 1;
    if (cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680) return 1;
    if (cp >= 0x2000 && cp <= 0x200A) return 1;
    return cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}

/* decode one code point of valid UTF-8 at s[i]; returns its byte length */
static size_t utf8_decode(const uint8_t *s, size_t i, uint32_t *cp) {
    uint8_t c = s[i];
    if (c < 0x80) {
        *cp = c;
        return 1;
    }
    if (c < 0xE0) {
        *cp = ((uint32_t)(c & 0x1F) << 6) | (s[i + 1] & 0x3F);
        return 2;
    }
    if (c < 0xF0) {
        *cp = ((uint32_t)(c & 0x0F) << 12) | ((uint32_t)(s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F);
        return 3;
    }
    *cp = ((uint32_t)(c & 0x07) << 18) | ((uint32_t)(s[i + 1] & 0x3F) << 12) |
          ((uint32_t)(s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F);
    return 4;
}

/* str::trim on valid UTF-8: strips leading and trailing White_Space code points. */
static void utf8_trim(const uint8_t *s, size_t n, size_t *b, size_t *e) {
    size_t i = 0;
    while (i < n) {
        uint32_t cp;
        size_t l = utf8_decode(s, i, &cp);
        if (!is_ws_cp(cp)) break;
        i += l;
    }
    size_t j = n;
    while (j > i) {
        size_t k = j - 1;
        while (k > i && (s[k] & 0xC0) == 0x80) k--;
        uint32_t cp;
        utf8_decode(s, k, &cp);
        if (!is_ws_cp(cp)) break;
        j = k;
    }
    *b = i;
    *e = j;
}

/* utils.rs:91-103 get_id: from_utf8, split_whitespace, first token. */
int orc_get_id(const uint8_t *h, size_t n, size_t *off, size_t *id_len) {
    if (!utf8_valid(h, n)) return ORC_ERR_RECORD_NAME_UTF8;
    size_t i = 0;
    while (i < n) {
        uint32_t cp;
        size_t l = utf8_decode(h, i, &cp);
        if (!is_ws_cp(cp)) break;
        i += l;
    }
    if (i >= n) return ORC_ERR_FASTQ_HEADER; /* utils.rs:97-99 */
    size_t j = i;
    while (j < n) {
        uint32_t cp;
        size_t l = utf8_decode(h, j, &cp);
        if (is_ws_cp(cp)) break;
        j += l;
    }
    *off = i;
    *id_len = j - i;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* std::io::BufRead::lines over an in-memory buffer                    */
/* ------------------------------------------------------------------ */

typedef struct {
    const uint8_t *buf;
    size_t n, pos;
    uint64_t line_no;
} orc_lines;

/* 1 = line, 0 = end, -1 = invalid UTF-8 (io::ErrorKind::InvalidData) */
static int lines_next(orc_lines *it, const uint8_t **line, size_t *len) {
    if (it->pos >= it->n) return 0;
    const uint8_t *s = it->buf + it->pos;
    const uint8_t *nl = (const uint8_t *)memchr(s, '\n', it->n - it->pos);
    size_t raw = nl ? (size_t)(nl - s) + 1 : it->n - it->pos;
    it->pos += raw;
    it->line_no++;
    if (!utf8_valid(s, raw)) return -1;
    size_t l = raw;
    if (l && s[l - 1] == '\n') {
        l--;
        if (l && s[l - 1] == '\r') l--;
    }
    *line = s;
    *len = l;
    return 1;
}

/* str::split('\t') */
static size_t split_tabs(const uint8_t *s, size_t n, orc_span *f, size_t maxf) {
    size_t k = 0, st = 0;
    for (size_t i = 0; i <= n; i++) {
        if (i == n || s[i] == '\t') {
            if (k < maxf) {
                f[k].p = s + st;
                f[k].len = i - st;
            }
            k++;
            st = i + 1;
        }
    }
    return k;
}

/* core::num <uN as FromStr>::from_str: optional single '+', ASCII digits, overflow is an error */
static int parse_uint(const uint8_t *s, size_t n, uint64_t max, uint64_t *out) {
    if (n == 0) return -1;
    if (s[0] == '+' || s[0] == '-') {
        if (n == 1) return -1;
        if (s[0] == '-') return -1; /* unsigned: '-' is an invalid digit */
        s++;
        n--;
    }
    uint64_t v = 0;
    for (size_t i = 0; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return -1;
        uint64_t d = s[i] - '0';
        if (v > (UINT64_MAX - d) / 10) return -1;
        v = v * 10 + d;
        if (v > max) return -1;
    }
    *out = v;
    return 0;
}

/* ------------------------------------------------------------------ */
/* alignment.rs                                                        */
/* ------------------------------------------------------------------ */

/* alignment.rs:84-114 from_paf, :244-263 PafRecord::from_str, :265-275 */
int orc_set_from_paf(const uint8_t *buf, size_t n, uint64_t min_len, double min_cov,
                     uint8_t min_mapq, orc_set **out, uint64_t *err_line) {
    orc_set *set = orc_set_new();
    orc_lines it = {buf, n, 0, 0};
    const uint8_t *line;
    size_t len;
    int r, rc = ORC_OK;
    /* which of the 12 columns are parsed as integers, and their max */
    static const uint64_t maxv[12] = {0, UINT64_MAX, UINT64_MAX, UINT64_MAX, 0, 0,
                                      UINT64_MAX, UINT64_MAX, UINT64_MAX, UINT64_MAX, UINT64_MAX, 255};
    while ((r = lines_next(&it, &line, &len)) != 0) {
        if (r < 0) {
            rc = ORC_ERR_IO;
            break;
        }
        orc_span f[12];
        size_t nf = split_tabs(line, len, f, 12);
        uint64_t v[12] = {0};
        /* struct fields are evaluated in order: a missing column panics, a bad integer errors */
        for (size_t c = 0; c < 12 && rc == ORC_OK; c++) {
            if (c >= nf) {
                rc = ORC_ERR_WOULD_PANIC;
                break;
            }
            if (maxv[c] && parse_uint(f[c].p, f[c].len, maxv[c], &v[c]) != 0) rc = ORC_ERR_PAF_INTEGER;
        }
        if (rc != ORC_OK) break;
        uint64_t qlen = v[1], alen = v[3] - v[2]; /* usize wrap in release builds */
        double cov = qlen == 0 ? 0.0 : (double)alen / (double)qlen;
        if ((alen >= min_len || cov >= min_cov) && (uint8_t)v[11] >= min_mapq)
            orc_set_insert(set, f[0].p, f[0].len);
    }
    if (rc != ORC_OK) {
        if (err_line) *err_line = it.line_no - 1;
        orc_set_free(set);
        *out = NULL;
        return rc;
    }
    *out = set;
    return ORC_OK;
}

/* alignment.rs:117-146 from_bam + :154-211 (BamRecord::from, qalen_from_cigar, query_coverage), restated for
 * TEXT SAM.  The reference reads SAM/BAM/CRAM through rust-htslib (optional `htslib` feature, un-vendored C
 * library): the record fields below are what htslib's sam_parse1 stores for a well-formed line.  Restated
 * rules (parity unpinned):
 *   - lines split on '\n', one trailing '\r' dropped; lines starting with '@' are header lines (skipped);
 *   - a record has >= 11 tab-separated fields: QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL;
 *   - FLAG like strtol(.., 0): decimal, 0x hex or 0 octal, 0..65535; POS/PNEXT/TLEN signed decimal;
 *     MAPQ decimal 0..255; CIGAR "*" or (count op)+ with op in MIDNSHP=XB, count < 2^28;
 *   - SEQ "*" (length 0) or its byte length; with a CIGAR and a SEQ the CIGAR's query length (M I S = X) must
 *     equal it; QUAL "*" or as long as SEQ;  any violation -> ORC_ERR_SAM_RECORD at that line;
 *   - RNAME other than "*" is looked up among the SN: names of the header's @SQ lines (bam_name2id): with NO @SQ line
 *     at all that is a parse error ("no SQ lines present in the header") -> ORC_ERR_SAM_RECORD; a name that is not
 *     declared makes the record unmapped ("unrecognized reference name; treated as unmapped");
 *   - unmapped = FLAG & 4, or RNAME "*" / undeclared, or POS < 1 (htslib sets BAM_FUNMAP for all of them): skipped
 *     (alignment.rs:132-134);
 *   - qalen = sum of M and I counts (u32), qlen = SEQ length; pass iff (qalen >= min_len || cov >= min_cov)
 *     && mapq >= min_mapq with cov = qlen == 0 ? 0 : qalen / qlen (alignment.rs:136-140, :204-209);
 *   - QNAME must be valid UTF-8 (alignment.rs:187) -> ORC_ERR_RECORD_NAME_UTF8. */
static int sam_int(const uint8_t *s, size_t n, int allow_sign, int64_t *out) {
    size_t i = 0;
    int neg = 0;
    if (n && allow_sign && (s[0] == '-' || s[0] == '+')) {
        neg = s[0] == '-';
        i = 1;
    }
    if (i >= n || n - i > 18) return -1;
    int64_t v = 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return -1;
        v = v * 10 + (s[i] - '0');
    }
    *out = neg ? -v : v;
    return 0;
}
static int sam_flag(const uint8_t *s, size_t n, uint32_t *out) {
    uint32_t base = 10, v = 0;
    size_t i = 0;
    if (n >= 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) {
        base = 16;
        i = 2;
    } else if (n >= 2 && s[0] == '0') {
        base = 8;
        i = 1;
    }
    if (i >= n) return -1;
    for (; i < n; i++) {
        uint32_t d;
        if (s[i] >= '0' && s[i] <= '9') d = s[i] - '0';
        else if (s[i] >= 'a' && s[i] <= 'f') d = s[i] - 'a' + 10;
        else if (s[i] >= 'A' && s[i] <= 'F') d = s[i] - 'A' + 10;
        else return -1;
        if (d >= base) return -1;
        v = v * base + d;
        if (v > 65535) return -1;
    }
    *out = v;
    return 0;
}
/* query length (M I S = X) and aligned length (M I) of a CIGAR string; -1 on a malformed one */
static int sam_cigar(const uint8_t *s, size_t n, uint32_t *n_ops, uint32_t *qlen, uint32_t *qalen) {
    *n_ops = *qlen = *qalen = 0;
    if (n == 1 && s[0] == '*') return 0;
    size_t i = 0;
    if (n == 0) return -1;
    while (i < n) {
        uint64_t c = 0;
        size_t d0 = i;
        while (i < n && s[i] >= '0' && s[i] <= '9') {
            c = c * 10 + (s[i] - '0');
            if (c >= (1ull << 28)) return -1;
            i++;
        }
        if (i == d0 || i >= n) return -1;
        const uint8_t op = s[i++];
        if (!strchr("MIDNSHP=XB", op)) return -1;
        (*n_ops)++;
        if (op == 'M' || op == 'I') *qalen += (uint32_t)c;
        if (op == 'M' || op == 'I' || op == 'S' || op == '=' || op == 'X') *qlen += (uint32_t)c;
    }
    return 0;
}

int orc_set_from_sam(const uint8_t *buf, size_t n, uint64_t min_len, double min_cov, uint8_t min_mapq,
                     orc_set **out, uint64_t *err_line) {
    orc_set *set = orc_set_new();
    size_t pos = 0;
    uint64_t line_no = 0;
    int rc = ORC_OK;
    /* the reference names the header declares: SN: of every @SQ line */
    orc_set *refs = orc_set_new();
    uint64_t n_targets = 0;
    while (pos < n) {
        const uint8_t *s = buf + pos;
        const uint8_t *nl = (const uint8_t *)memchr(s, '\n', n - pos);
        size_t len = nl ? (size_t)(nl - s) : n - pos;
        pos += len + (nl ? 1 : 0);
        if (nl && len && s[len - 1] == '\r') len--;
        if (len >= 4 && s[0] == '@' && s[1] == 'S' && s[2] == 'Q' && s[3] == '\t') {
            size_t a = 4;
            while (a < len) {
                size_t b = a;
                while (b < len && s[b] != '\t') b++;
                if (b - a >= 3 && s[a] == 'S' && s[a + 1] == 'N' && s[a + 2] == ':') {
                    orc_set_insert(refs, s + a + 3, b - a - 3);
                    n_targets++;
                    break;
                }
                a = b + 1;
            }
        }
    }
    pos = 0;
    while (pos < n && rc == ORC_OK) {
        const uint8_t *s = buf + pos;
        const uint8_t *nl = (const uint8_t *)memchr(s, '\n', n - pos);
        size_t len = nl ? (size_t)(nl - s) : n - pos;
        pos += len + (nl ? 1 : 0);
        line_no++;
        if (nl && len && s[len - 1] == '\r') len--;
        if (len && s[0] == '@') continue;
        orc_span f[11];
        if (split_tabs(s, len, f, 11) < 11) {
            rc = ORC_ERR_SAM_RECORD;
            break;
        }
        uint32_t flag = 0, n_ops = 0, cq = 0, qalen = 0;
        int64_t p = 0, mapq = 0, t = 0;
        const int rname_star = f[2].len == 1 && f[2].p[0] == '*';
        if (f[0].len == 0 || sam_flag(f[1].p, f[1].len, &flag) || f[2].len == 0 || (!rname_star && n_targets == 0) ||
            sam_int(f[3].p, f[3].len, 1, &p) ||
            sam_int(f[4].p, f[4].len, 0, &mapq) || mapq > 255 || sam_cigar(f[5].p, f[5].len, &n_ops, &cq, &qalen) ||
            f[6].len == 0 || sam_int(f[7].p, f[7].len, 1, &t) || sam_int(f[8].p, f[8].len, 1, &t) || f[9].len == 0 ||
            f[10].len == 0) {
            rc = ORC_ERR_SAM_RECORD;
            break;
        }
        const int seq_star = f[9].len == 1 && f[9].p[0] == '*';
        const uint32_t qlen = seq_star ? 0u : (uint32_t)f[9].len;
        const int qual_star = f[10].len == 1 && f[10].p[0] == '*';
        if ((n_ops && !seq_star && cq != qlen) || (!qual_star && f[10].len != (seq_star ? 0 : f[9].len))) {
            rc = ORC_ERR_SAM_RECORD;
            break;
        }
        if (!utf8_valid(f[0].p, f[0].len)) {
            rc = ORC_ERR_RECORD_NAME_UTF8;
            break;
        }
        const int unmapped = (flag & 4) || rname_star || !orc_set_contains(refs, f[2].p, f[2].len) || p < 1;
        if (unmapped) continue;
        const double cov = qlen == 0 ? 0.0 : (double)qalen / (double)qlen;
        if (((uint64_t)qalen >= min_len || cov >= min_cov) && (uint8_t)mapq >= min_mapq)
            orc_set_insert(set, f[0].p, f[0].len);
    }
    orc_set_free(refs);
    if (rc != ORC_OK) {
        if (err_line) *err_line = line_no - 1;
        orc_set_free(set);
        *out = NULL;
        return rc;
    }
    *out = set;
    return ORC_OK;
}

/* alignment.rs:117-146 from_bam + BamRecord::from :180-197 for BINARY BAM records.  `buf` is the BGZF-DECOMPRESSED
 * stream (inflating the BGZF blocks -- concatenated gzip members -- is a host stage, like gz for FASTQ).  Restated from
 * the SAM/BAM specification (SAMv1 section 4.2) and htslib's bam_hdr_read / bam_read1 / bam_tag2cigar, which
 * rust_htslib::bam::Reader::records drives (htslib is an un-vendored dependency: parity unpinned):
 *   header : magic "BAM\1", l_text, text, n_ref, per reference l_name, name, l_ref; anything else -> ORC_ERR_BAM_RECORD
 *            at record 0 (bam::Reader::from_path fails);
 *   record : block_size (u32 LE) then block_size bytes; end of data exactly at a record boundary ends the file, a
 *            partial block_size or a block that runs past the end is an error (truncated file), block_size < 32 too;
 *            l_read_name >= 1, l_seq >= 0 and 32 + l_read_name + 4 n_cigar_op + (l_seq+1)/2 + l_seq <= block_size,
 *            else error -- all ORC_ERR_BAM_RECORD at that record's index, and the records before it have been seen
 *            (an error is returned through `result?`, alignment.rs:131: the run fails);
 *   unmapped (FLAG & 4) records are skipped BEFORE anything else is looked at (alignment.rs:132-134);
 *   qname  : the l_read_name - 1 bytes before the terminating NUL (rust_htslib Record::qname); a read_name whose last
 *            byte is not NUL is taken whole (htslib appends the missing NUL); must be UTF-8 (alignment.rs:187);
 *   CIGAR  : n_cigar_op u32 values (len << 4 | op); when the record carries the long-CIGAR placeholder (first op
 *            soft clip of l_seq, refID >= 0, pos >= 0) and a CG:B,I tag with at least n_cigar_op and fewer than 2^29
 *            entries, the tag's array IS the CIGAR (bam_tag2cigar);
 *   qalen  = sum of the lengths of M (op 0) and I (op 1) operations, u32 wrapping (release build); qlen = l_seq. */
static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t le16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

/* finds the CG:B,I/i array among the auxiliary fields [a, e): returns its element count and sets *arr, or -1 */
static int64_t bam_find_cg(const uint8_t *a, const uint8_t *e, const uint8_t **arr) {
    while (e - a >= 3) {
        const uint8_t t0 = a[0], t1 = a[1], ty = a[2];
        a += 3;
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'Z': case 'H': {
                const uint8_t *z = (const uint8_t *)memchr(a, 0, (size_t)(e - a));
                if (!z) return -1;
                sz = (size_t)(z - a) + 1;
                break;
            }
            case 'B': {
                if (e - a < 5) return -1;
                const uint8_t sub = a[0];
                const uint32_t cnt = le32(a + 1);
                size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
                if (!es) return -1;
                if ((uint64_t)cnt * es > (uint64_t)(e - a - 5)) return -1;
                if (t0 == 'C' && t1 == 'G') {
                    if (sub != 'I' && sub != 'i') return -1;
                    *arr = a + 5;
                    return (int64_t)cnt;
                }
                sz = 5 + (size_t)cnt * es;
                break;
            }
            default: return -1;
        }
        if (t0 == 'C' && t1 == 'G') return -1; /* a CG tag of another type is not a CIGAR */
        if ((size_t)(e - a) < sz) return -1;
        a += sz;
    }
    return -1;
}

int orc_set_from_bam(const uint8_t *buf, size_t n, uint64_t min_len, double min_cov, uint8_t min_mapq,
                     orc_set **out, uint64_t *err_record) {
    *out = NULL;
    if (err_record) *err_record = 0;
    /* header */
    if (n < 12 || memcmp(buf, "BAM\1", 4) != 0) return ORC_ERR_BAM_RECORD;
    size_t pos = 4;
    const uint32_t l_text = le32(buf + pos);
    pos += 4;
    if (l_text > n - pos || n - pos - l_text < 4) return ORC_ERR_BAM_RECORD;
    pos += l_text;
    const uint32_t n_ref = le32(buf + pos);
    pos += 4;
    for (uint32_t r = 0; r < n_ref; r++) {
        if (n - pos < 4) return ORC_ERR_BAM_RECORD;
        const uint32_t l_name = le32(buf + pos);
        pos += 4;
        if (l_name > n - pos || n - pos - l_name < 4) return ORC_ERR_BAM_RECORD;
        pos += (size_t)l_name + 4;
    }
    orc_set *set = orc_set_new();
    uint64_t rec = 0;
    int rc = ORC_OK;
    while (pos < n) {
        if (n - pos < 4) { rc = ORC_ERR_BAM_RECORD; break; }
        const uint32_t bs = le32(buf + pos);
        if (bs < 32 || bs > n - pos - 4) { rc = ORC_ERR_BAM_RECORD; break; }
        const uint8_t *b = buf + pos + 4;
        const int32_t ref_id = (int32_t)le32(b), rpos = (int32_t)le32(b + 4);
        const uint32_t l_name = b[8], mapq = b[9], n_cig = le16(b + 12), flag = le16(b + 14);
        const int32_t l_seq = (int32_t)le32(b + 16);
        if (l_name < 1 || l_seq < 0 ||
            32ull + l_name + 4ull * n_cig + (((uint64_t)l_seq + 1) >> 1) + (uint64_t)l_seq > bs) {
            rc = ORC_ERR_BAM_RECORD;
            break;
        }
        pos += 4 + (size_t)bs;
        if (flag & 4) { rec++; continue; }
        const uint8_t *name = b + 32;
        const uint32_t qn = name[l_name - 1] == 0 ? l_name - 1 : l_name;
        if (!utf8_valid(name, qn)) { rc = ORC_ERR_RECORD_NAME_UTF8; break; }
        const uint8_t *cig = name + l_name;
        uint64_t n_ops = n_cig;
        if (n_cig && ref_id >= 0 && rpos >= 0 && (le32(cig) & 15) == 4 && (le32(cig) >> 4) == (uint32_t)l_seq) {
            const uint8_t *aux = cig + 4ull * n_cig + (((uint64_t)l_seq + 1) >> 1) + (uint64_t)l_seq, *arr = NULL;
            const int64_t k = bam_find_cg(aux, b + bs, &arr);
            if (k >= (int64_t)n_cig && k < (1ll << 29)) { cig = arr; n_ops = (uint64_t)k; }
        }
        uint32_t qalen = 0;
        for (uint64_t i = 0; i < n_ops; i++) {
            const uint32_t v = le32(cig + 4 * i);
            if ((v & 15) <= 1) qalen += v >> 4;
        }
        const uint32_t qlen = (uint32_t)l_seq;
        const double cov = qlen == 0 ? 0.0 : (double)qalen / (double)qlen;
        if (((uint64_t)qalen >= min_len || cov >= min_cov) && mapq >= min_mapq) orc_set_insert(set, name, qn);
        rec++;
    }
    if (rc != ORC_OK) {
        if (err_record) *err_record = rec;
        orc_set_free(set);
        return rc;
    }
    *out = set;
    return ORC_OK;
}

/* alignment.rs:60-82 from_txt: every line verbatim */
int orc_set_from_txt(const uint8_t *buf, size_t n, orc_set **out, uint64_t *err_line) {
    orc_set *set = orc_set_new();
    orc_lines it = {buf, n, 0, 0};
    const uint8_t *line;
    size_t len;
    int r;
    while ((r = lines_next(&it, &line, &len)) != 0) {
        if (r < 0) {
            if (err_line) *err_line = it.line_no - 1;
            orc_set_free(set);
            *out = NULL;
            return ORC_ERR_IO;
        }
        orc_set_insert(set, line, len);
    }
    *out = set;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* classifier.rs                                                       */
/* ------------------------------------------------------------------ */

enum { /* classifier.rs:18-33, declaration order == PartialOrd order */
    LV_NONE, LV_UNCLASSIFIED, LV_NORANK, LV_ROOT, LV_DOMAIN, LV_KINGDOM, LV_PHYLUM,
    LV_CLASS, LV_ORDER, LV_FAMILY, LV_GENUS, LV_SPECIES, LV_UNSPECIFIED
};

static int starts_with(const uint8_t *s, size_t n, const char *p) {
    size_t l = strlen(p);
    return n >= l && memcmp(s, p, l) == 0;
}

/* classifier.rs:345-373 */
static int tax_level(const uint8_t *s, size_t n) {
    if (starts_with(s, n, "U")) return LV_UNCLASSIFIED;
    if (starts_with(s, n, "no rank")) return LV_NORANK;
    if (starts_with(s, n, "R")) return LV_ROOT;
    if (starts_with(s, n, "D") || starts_with(s, n, "superkingdom")) return LV_DOMAIN;
    if (starts_with(s, n, "K") || starts_with(s, n, "kingdom")) return LV_KINGDOM;
    if (starts_with(s, n, "P") || starts_with(s, n, "phylum")) return LV_PHYLUM;
    if (starts_with(s, n, "C") || starts_with(s, n, "class")) return LV_CLASS;
    if (starts_with(s, n, "O") || starts_with(s, n, "order")) return LV_ORDER;
    if (starts_with(s, n, "F") || starts_with(s, n, "family")) return LV_FAMILY;
    if (starts_with(s, n, "G") || starts_with(s, n, "genus")) return LV_GENUS;
    if (starts_with(s, n, "S") || starts_with(s, n, "species")) return LV_SPECIES;
    return LV_UNSPECIFIED;
}

static int list_contains(const orc_span *l, size_t n, const uint8_t *s, size_t len) {
    for (size_t i = 0; i < n; i++)
        if (l[i].len == len && memcmp(l[i].p, s, len) == 0) return 1;
    return 0;
}

/* classifier.rs:124-252 get_taxids_from_report, :449-466 KrakenReportRecord::from_str */
int orc_taxids_from_report(const uint8_t *buf, size_t n, const char *const *taxa, size_t n_taxa,
                           const char *const *taxa_direct, size_t n_direct, orc_set **out,
                           uint64_t *err_line) {
    orc_span *tx = (orc_span *)malloc(sizeof(orc_span) * (n_taxa + n_direct + 1));
    orc_span *td = tx + n_taxa;
    for (size_t i = 0; i < n_taxa + n_direct; i++) { /* classifier.rs:132-133 x.trim() */
        const char *s = i < n_taxa ? taxa[i] : taxa_direct[i - n_taxa];
        size_t b, e;
        utf8_trim((const uint8_t *)s, strlen(s), &b, &e);
        tx[i].p = (const uint8_t *)s + b;
        tx[i].len = e - b;
    }
    orc_set *set = orc_set_new();
    orc_lines it = {buf, n, 0, 0};
    const uint8_t *line;
    size_t len;
    int r, rc = ORC_OK;
    int extract_level = LV_NONE;
    int parent_empty = 1; /* extract_parent == "" */
    while ((r = lines_next(&it, &line, &len)) != 0) {
        if (r < 0) {
            rc = ORC_ERR_IO;
            break;
        }
        orc_span f[6];
        size_t nf = split_tabs(line, len, f, 6);
        uint64_t reads, direct;
        if (nf < 2) { rc = ORC_ERR_WOULD_PANIC; break; }
        if (parse_uint(f[1].p, f[1].len, UINT64_MAX, &reads)) { rc = ORC_ERR_KRAKEN_REPORT_READS; break; }
        if (nf < 3) { rc = ORC_ERR_WOULD_PANIC; break; }
        if (parse_uint(f[2].p, f[2].len, UINT64_MAX, &direct)) { rc = ORC_ERR_KRAKEN_REPORT_DIRECT; break; }
        if (nf < 6) { rc = ORC_ERR_WOULD_PANIC; break; }
        size_t b, e;
        utf8_trim(f[3].p, f[3].len, &b, &e);
        const uint8_t *lv = f[3].p + b; size_t lv_n = e - b;
        utf8_trim(f[4].p, f[4].len, &b, &e);
        const uint8_t *id = f[4].p + b; size_t id_n = e - b;
        utf8_trim(f[5].p, f[5].len, &b, &e);
        const uint8_t *nm = f[5].p + b; size_t nm_n = e - b;
        int level = tax_level(lv, lv_n);

        if (list_contains(td, n_direct, nm, nm_n) || list_contains(td, n_direct, id, id_n))
            orc_set_insert(set, id, id_n); /* :145-155 */
        if (level < LV_DOMAIN) continue;   /* :157-166 */
        if (list_contains(tx, n_taxa, nm, nm_n) || list_contains(tx, n_taxa, id, id_n)) {
            extract_level = level;         /* :168-187 */
            parent_empty = nm_n == 0;
            if (direct > 0) orc_set_insert(set, id, id_n);
        } else {
            if (extract_level == LV_NONE) continue; /* :189-199 */
            if (level <= extract_level && lv_n == 1) {
                extract_level = LV_NONE;   /* :200-208 */
            } else if (direct > 0) {       /* :210-223 */
                orc_set_insert(set, id, id_n);
                if (parent_empty) { rc = ORC_ERR_KRAKEN_REPORT_PARENT; break; }
            }
        }
    }
    free(tx);
    if (rc != ORC_OK) {
        if (err_line) *err_line = it.line_no - 1;
        orc_set_free(set);
        *out = NULL;
        return rc;
    }
    *out = set;
    return ORC_OK;
}

/* classifier.rs:270-290 / :308-328 and the two from_str at :401-419 / :497-517 */
int orc_set_from_reads(const uint8_t *buf, size_t n, int style, const orc_set *taxids,
                       orc_set **out, uint64_t *err_line) {
    size_t need = style == 0 ? 5 : 7;
    orc_set *set = orc_set_new();
    orc_lines it = {buf, n, 0, 0};
    const uint8_t *line;
    size_t len;
    int r, rc = ORC_OK;
    while ((r = lines_next(&it, &line, &len)) != 0) {
        if (r < 0) { rc = ORC_ERR_IO; break; }
        orc_span f[7];
        size_t nf = split_tabs(line, len, f, 7);
        if (nf < need) { rc = ORC_ERR_WOULD_PANIC; break; }
        size_t b, e, tb, te;
        utf8_trim(f[1].p, f[1].len, &b, &e);
        utf8_trim(f[2].p, f[2].len, &tb, &te);
        if (orc_set_contains(taxids, f[2].p + tb, te - tb)) orc_set_insert(set, f[1].p + b, e - b);
    }
    if (rc != ORC_OK) {
        if (err_line) *err_line = it.line_no - 1;
        orc_set_free(set);
        *out = NULL;
        return rc;
    }
    *out = set;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* needletail 0.5.1 fastq reader / writer                              */
/* ------------------------------------------------------------------ */

typedef struct {
    const uint8_t *buf;
    size_t n;
    size_t start;   /* BufferPosition.start of the next record */
    int finished;
    int have_le, crlf; /* Reader.line_ending */
    uint64_t index;    /* records returned so far */
    int fasta;         /* first byte '>': needletail's FASTA reader */
} orc_reader;

typedef struct {
    size_t id, id_n, seq, seq_n, qual, qual_n;
} orc_rec;

static size_t trim_cr_len(const uint8_t *s, size_t n) { return (n && s[n - 1] == '\r') ? n - 1 : n; }

/* Reader::validate: start byte, separator byte, equal lengths (in this order) */
static int rd_validate(const orc_reader *r, size_t start, size_t p1, size_t p2, size_t p3,
                       size_t end, orc_rec *rec) {
    const uint8_t *b = r->buf;
    if (b[start] != '@') return ORC_ERR_FASTQ_INVALID_START;
    if (b[p2 + 1] != '+') return ORC_ERR_FASTQ_INVALID_SEPARATOR;
    rec->id = start + 1;
    rec->id_n = trim_cr_len(b + start + 1, p1 - (start + 1));
    rec->seq = p1 + 1;
    rec->seq_n = trim_cr_len(b + p1 + 1, p2 - (p1 + 1));
    rec->qual = p3 + 1;
    rec->qual_n = trim_cr_len(b + p3 + 1, end - (p3 + 1));
    if (rec->seq_n != rec->qual_n) return ORC_ERR_FASTQ_UNEQUAL_LENGTHS;
    return ORC_OK;
}

/* needletail 0.5.1 fasta::Reader::next / find / _find (un-vendored; restated from the published source, parity
 * unpinned).  A record runs from its '>' to the newline in front of the next "\n>"; `seq_pos` in the reference is the
 * list of the record's newline positions, EXCEPT a newline that is the last byte of the input, which is only appended
 * at end of input when an earlier one exists.  Hence:
 *   id      = trim_cr(buf[start+1 .. first newline))
 *   raw_seq = trim_cr(buf[first newline + 1 .. last seq_pos)) when there are two or more positions, else empty --
 *             inner newlines (and inner CRs) of a multi-line sequence are kept verbatim (SequenceRecord::write passes
 *             raw_seq to write_fasta);
 *   a record without any position (a header with no newline, or whose only newline ends the input) is UnexpectedEnd;
 *   the line ending is taken from the first record whose bytes [start, last seq_pos) hold a newline
 *   (find_line_ending over BufferPosition::all); a record written before that uses LF.
 * 1 = record (seq / seq_n = raw_seq, qual_n = 0), 0 = end, < 0 = -(error code). */
static int fa_next(orc_reader *r, orc_rec *rec) {
    if (r->finished || r->start >= r->n) return 0;
    const uint8_t *b = r->buf;
    const size_t s = r->start, n = r->n;
    size_t first = (size_t)-1, last = 0, npos = 0, from = s, next_start = n;
    int complete = 0;
    while (from < n) {
        const uint8_t *q = (const uint8_t *)memchr(b + from, '\n', n - from);
        if (!q) break;
        const size_t pos = (size_t)(q - b);
        if (pos + 1 == n) { /* cannot look at the next byte: not pushed now; end of input appends it if others exist */
            if (npos) { last = pos; npos++; }
            from = n;
            complete = 2;
            break;
        }
        if (!npos) first = pos;
        last = pos;
        npos++;
        if (b[pos + 1] == '>') {
            next_start = pos + 1;
            complete = 1;
            break;
        }
        from = pos + 1;
    }
    if (complete == 0 && npos) { last = n; npos++; } /* no trailing newline: the end of the input is the last position */
    if (complete != 1) r->finished = 1;
    if (npos == 0) return -ORC_ERR_FASTQ_UNEXPECTED_END;
    rec->id = s + 1;
    rec->id_n = trim_cr_len(b + s + 1, first - (s + 1));
    rec->seq = first + 1;
    rec->seq_n = npos > 1 ? trim_cr_len(b + first + 1, last - (first + 1)) : 0;
    rec->qual = rec->qual_n = 0;
    if (!r->have_le) { /* find_line_ending(buf[start .. last)) */
        const uint8_t *q = (const uint8_t *)memchr(b + s, '\n', last - s);
        if (q) {
            r->have_le = 1;
            r->crlf = q > b + s && q[-1] == '\r';
        }
    }
    r->start = next_start;
    r->index++;
    return 1;
}

/* Reader::next + find + check_end.  1 = record, 0 = end of records, <0 = -(error code) */
static int rd_next(orc_reader *r, orc_rec *rec) {
    if (r->fasta) return fa_next(r, rec);
    if (r->finished) return 0;
    const uint8_t *b = r->buf;
    size_t s = r->start, n = r->n;
    size_t p[4];
    int found = 0;
    size_t from = s;
    while (found < 4 && from < n) {
        const uint8_t *q = (const uint8_t *)memchr(b + from, '\n', n - from);
        if (!q) break;
        p[found++] = (size_t)(q - b);
        from = p[found - 1] + 1;
    }
    size_t end;
    if (found == 4) {
        end = p[3];
    } else {
        /* check_end: at EOF with an incomplete record */
        r->finished = 1;
        if (found == 3) {
            end = n; /* SearchPosition::Quality: last record has no trailing newline */
        } else {
            /* a tail made only of blank lines (after trim_cr) is tolerated */
            size_t i = s;
            while (i < n) {
                const uint8_t *q = (const uint8_t *)memchr(b + i, '\n', n - i);
                size_t e = q ? (size_t)(q - b) : n;
                if (trim_cr_len(b + i, e - i) != 0) return -ORC_ERR_FASTQ_UNEXPECTED_END;
                i = e + 1;
            }
            return 0;
        }
    }
    int rc = rd_validate(r, s, p[0], p[1], p[2], end, rec);
    if (rc != ORC_OK) {
        r->finished = 1;
        return -rc;
    }
    if (!r->have_le) { /* find_line_ending(buf[start..end]) on the first record */
        r->have_le = 1;
        r->crlf = p[0] > s && b[p[0] - 1] == '\r';
    }
    r->start = end + 1;
    r->index++;
    return 1;
}

/* needletail write_fastq: '@' id E seq E '+' E qual E;  write_fasta: '>' id E raw_seq E */
static size_t wr_record(uint8_t *o, size_t w, const uint8_t *b, const orc_rec *rec, int crlf, int fasta) {
#define PUT_E()                 \
    do {                        \
        if (crlf) o[w++] = '\r'; \
        o[w++] = '\n';          \
    } while (0)
    if (fasta) {
        o[w++] = '>';
        memcpy(o + w, b + rec->id, rec->id_n);
        w += rec->id_n;
        PUT_E();
        memcpy(o + w, b + rec->seq, rec->seq_n);
        w += rec->seq_n;
        PUT_E();
        return w;
    }
    o[w++] = '@';
    memcpy(o + w, b + rec->id, rec->id_n);
    w += rec->id_n;
    PUT_E();
    memcpy(o + w, b + rec->seq, rec->seq_n);
    w += rec->seq_n;
    PUT_E();
    o[w++] = '+';
    PUT_E();
    memcpy(o + w, b + rec->qual, rec->qual_n);
    w += rec->qual_n;
    PUT_E();
#undef PUT_E
    return w;
}

/* utils.rs:359-383: niffler sniff needs 5 bytes (FileTooShort => "empty"); needletail
 * picks the parser from the first byte ('@' fastq, '>' fasta, else unknown format). */
static int rd_open(orc_reader *r, const uint8_t *in, size_t n, orc_counts *c) {
    memset(r, 0, sizeof(*r));
    r->buf = in;
    r->n = n;
    if (n < 5) {
        if (c) c->empty_input = 1;
        r->finished = 1;
        return ORC_OK;
    }
    if (in[0] == '>') {
        r->fasta = 1;
        return ORC_OK;
    }
    if (in[0] != '@') return ORC_ERR_FASTQ_UNKNOWN_FORMAT;
    return ORC_OK;
}

/* cleaner.rs:731-760 FastqCleaner::clean_reads (loop :742-754) */
int orc_clean_fastq(const uint8_t *in, size_t n_in, const orc_set *set, int reverse,
                    uint8_t *out_written, size_t *n_written, uint8_t *out_other,
                    size_t *n_other, orc_counts *c) {
    orc_reader r;
    orc_rec rec;
    size_t w = 0, wo = 0;
    memset(c, 0, sizeof(*c));
    int rc = rd_open(&r, in, n_in, c);
    if (rc == ORC_OK) {
        int k;
        while ((k = rd_next(&r, &rec)) != 0) {
            if (k < 0) {
                rc = -k;
                c->error_record = r.index;
                break;
            }
            size_t off, idn;
            rc = orc_get_id(in + rec.id, rec.id_n, &off, &idn);
            if (rc != ORC_OK) {
                c->error_record = r.index - 1;
                break;
            }
            c->reads_in++;
            int hit = orc_set_contains(set, in + rec.id + off, idn);
            if ((!reverse && !hit) || (reverse && hit)) {
                w = wr_record(out_written, w, in, &rec, r.crlf, r.fasta);
                c->reads_out++;
            } else if (out_other) {
                wo = wr_record(out_other, wo, in, &rec, r.crlf, r.fasta);
            }
        }
    }
    c->crlf = (uint32_t)r.crlf;
    *n_written = w;
    if (n_other) *n_other = wo;
    return rc;
}

/* utils.rs:250-285 ReadDifference::get_difference, one (input, output) file pair */
int orc_diff(const uint8_t *in, size_t n_in, const uint8_t *out, size_t n_out,
             orc_counts *c, orc_set *diff_ids) {
    orc_reader r;
    orc_rec rec;
    orc_counts tmp;
    memset(&tmp, 0, sizeof(tmp));
    orc_set *o_ids = orc_set_new();
    int rc = rd_open(&r, out, n_out, &tmp); /* :259-267 output first */
    int k;
    size_t off, idn;
    while (rc == ORC_OK && (k = rd_next(&r, &rec)) != 0) {
        if (k < 0) { rc = -k; c->error_record = r.index; break; }
        rc = orc_get_id(out + rec.id, rec.id_n, &off, &idn);
        if (rc != ORC_OK) { c->error_record = r.index - 1; break; }
        orc_set_insert(o_ids, out + rec.id + off, idn);
        c->reads_out++;
    }
    if (rc == ORC_OK) rc = rd_open(&r, in, n_in, &tmp); /* :269-283 */
    while (rc == ORC_OK && (k = rd_next(&r, &rec)) != 0) {
        if (k < 0) { rc = -k; c->error_record = r.index; break; }
        rc = orc_get_id(in + rec.id, rec.id_n, &off, &idn);
        if (rc != ORC_OK) { c->error_record = r.index - 1; break; }
        if (!orc_set_contains(o_ids, in + rec.id + off, idn)) {
            orc_set_insert(diff_ids, in + rec.id + off, idn);
            c->difference++;
        }
        c->reads_in++;
    }
    orc_set_free(o_ids);
    return rc;
}

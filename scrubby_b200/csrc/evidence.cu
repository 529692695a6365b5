// evidence.cu -- alignment / classifier evidence -> read-id set.
//
//   paf_parse_kernel + paf_segment_kernel  replace ReadAlignment::from_paf (alignment.rs:84-114),
//       PafRecord::from_str (:244-263), query_aligned_length / query_coverage (:265-275) and the
//       duplicate loop at cleaner.rs:669-678.  The per-read aggregation that is bit-exact with the
//       reference is a segmented logical OR of per-record predicates over runs of adjacent equal
//       qnames, NOT a sum of aligned lengths.
//   txt_lines_kernel                        replaces ReadAlignment::from_txt (alignment.rs:60-82).
//   reads_parse_kernel                      replaces get_taxid_reads_kraken / _metabuli
//       (classifier.rs:270-290, 308-328) with KrakenReadRecord / MetabuliReadRecord::from_str
//       (:401-419, :497-517); taxid membership is a bitmap lookup for canonical decimal taxids and
//       a byte compare against the (normally empty) list of non-canonical taxid strings.
// Lines follow std::io::BufRead::lines: split on '\n', one '\r' stripped before it, UTF-8 checked.
#include <string>
#include <vector>

#include "idset.cuh"

namespace sgpu {

struct LineParams {
    const uint8_t *in;
    uint64_t n;
    const uint64_t *nlpos;
    uint64_t n_nl;
};

// raw line i = [s, e) (without '\n'); returns false if there is no such line
__device__ __forceinline__ bool line_span(const LineParams &P, uint64_t i, uint64_t *s, uint64_t *e) {
    uint64_t st = i ? P.nlpos[i - 1] + 1 : 0;
    uint64_t en;
    if (i < P.n_nl) {
        en = P.nlpos[i];
    } else {
        if (st >= P.n) return false;  // the buffer ended with '\n' (or is empty)
        en = P.n;
    }
    *s = st;
    *e = en;
    return true;
}

// BufRead::lines: validates UTF-8 (InvalidData otherwise) and strips "\n" / "\r\n"
__device__ __forceinline__ bool line_content(const LineParams &P, uint64_t i, uint64_t s, uint64_t *e, bool *high) {
    const uint8_t *in = P.in;
    // any byte with its high bit set?  (bytes up to a 4-byte boundary, aligned words, tail bytes)
    uint32_t acc = 0;
    uint64_t k = s;
    const uint64_t end = *e;
    while (k < end && ((uintptr_t)(in + k) & 3)) acc |= in[k++];
    for (; k + 4 <= end; k += 4) acc |= *reinterpret_cast<const uint32_t *>(in + k);
    while (k < end) acc |= in[k++];
    const bool h = (acc & 0x80808080u) != 0;
    *high = h;
    if (h && !utf8_valid(in + s, *e - s)) return false;
    if (i < P.n_nl && *e > s && in[*e - 1] == '\r') (*e)--;
    return true;
}

struct PafParams {
    uint64_t min_len;
    double min_cov;
    uint32_t min_mapq;
};

__global__ void __launch_bounds__(128)
    paf_parse_kernel(LineParams P, PafParams F, uint64_t n_lines, uint64_t *key_off, uint32_t *key_len, uint8_t *pass,
                     uint8_t *head, unsigned long long *err_word) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint64_t s, e;
    uint8_t ok = 0, hd = 1;
    uint64_t koff = 0;
    uint32_t klen = 0;
    if (line_span(P, i, &s, &e)) {
        bool high;
        if (!line_content(P, i, s, &e, &high)) {
            report_error(err_word, i, SGPU_ERR_IO);
        } else {
            const uint8_t *in = P.in;
            int f = 0, code = 0;
            uint64_t fs = s, qlen = 0, qstart = 0, qend = 0, mapq = 0;
            for (uint64_t pos = s; pos <= e && f < 12; pos++) {
                if (pos == e || in[pos] == '\t') {
                    uint64_t v = 0;
                    // columns 2,3,4,7,8,9,10,11 as u64/usize, 12 as u8 (alignment.rs:248-259)
                    bool is_int = (f >= 1 && f <= 3) || (f >= 6 && f <= 11);
                    if (is_int && !parse_uint(in + fs, pos - fs, f == 11 ? 255ull : ~0ull, &v)) {
                        code = SGPU_ERR_PAF_INTEGER;
                        break;
                    }
                    if (f == 0) {
                        koff = fs;
                        klen = (uint32_t)(pos - fs);
                    } else if (f == 1) qlen = v;
                    else if (f == 2) qstart = v;
                    else if (f == 3) qend = v;
                    else if (f == 11) mapq = v;
                    f++;
                    fs = pos + 1;
                }
            }
            if (!code && f < 12) code = SGPU_ERR_WOULD_PANIC;  // fields[f] out of bounds
            if (code) {
                report_error(err_word, i, code);
            } else {
                uint64_t alen = qend - qstart;  // usize subtraction wraps in release builds
                double cov = qlen == 0 ? 0.0 : __ull2double_rn(alen) / __ull2double_rn(qlen);
                ok = ((alen >= F.min_len || cov >= F.min_cov) && mapq >= F.min_mapq) ? 1 : 0;
                // run head: qname differs from the previous line's first column
                if (i > 0) {
                    uint64_t ps = i >= 2 ? P.nlpos[i - 2] + 1 : 0;
                    uint64_t pe = P.nlpos[i - 1];
                    bool same = true;
                    uint32_t k = 0;
                    for (; k < klen && same; k++) same = (ps + k < pe) && in[ps + k] == in[koff + k];
                    if (same) same = (ps + klen == pe) || in[ps + klen] == '\t' ||
                                     (ps + klen + 1 == pe && in[ps + klen] == '\r');
                    hd = same ? 0 : 1;
                }
            }
        }
    }
    key_off[i] = koff;
    key_len[i] = klen;
    pass[i] = ok;
    head[i] = hd;
}

// segmented OR over runs of adjacent equal qnames: one candidate per run
__global__ void paf_segment_kernel(const uint8_t *pass, const uint8_t *head, uint64_t n_lines, uint8_t *sel) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint8_t any = 0;
    if (head[i]) {
        any = pass[i];
        for (uint64_t j = i + 1; j < n_lines && !head[j] && !any; j++) any |= pass[j];
    }
    sel[i] = any;
}

__global__ void __launch_bounds__(128)
    txt_lines_kernel(LineParams P, uint64_t n_lines, uint64_t *key_off, uint32_t *key_len, uint8_t *sel,
                     unsigned long long *err_word) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint64_t s, e;
    uint8_t ok = 0;
    uint64_t koff = 0;
    uint32_t klen = 0;
    if (line_span(P, i, &s, &e)) {
        bool high;
        if (!line_content(P, i, s, &e, &high)) {
            report_error(err_word, i, SGPU_ERR_IO);
        } else {
            ok = 1;
            koff = s;
            klen = (uint32_t)(e - s);
        }
    }
    key_off[i] = koff;
    key_len[i] = klen;
    sel[i] = ok;
}

struct TaxSet {
    const uint32_t *bitmap;  // canonical decimal taxids
    uint64_t bits;
    const uint8_t *exotic;   // non-canonical taxid strings, flat
    const uint32_t *exotic_off;  // n_exotic + 1 offsets
    uint32_t n_exotic;
};

__device__ __forceinline__ bool taxid_member(const TaxSet &T, const uint8_t *s, uint32_t n) {
    // canonical decimal: digits only, no sign, no leading zero (except "0") -> bitmap
    bool canon = n >= 1 && n <= 10 && !(n > 1 && s[0] == '0');
    uint64_t v = 0;
    for (uint32_t k = 0; k < n && canon; k++) {
        uint32_t d = (uint32_t)s[k] - '0';
        canon = d <= 9;
        v = v * 10 + d;
    }
    if (canon && v < T.bits) return (T.bitmap[v >> 5] >> (v & 31)) & 1u;
    for (uint32_t x = 0; x < T.n_exotic; x++) {
        uint32_t a = T.exotic_off[x], b = T.exotic_off[x + 1];
        if (b - a == n && bytes_equal(T.exotic + a, s, n)) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128)
    reads_parse_kernel(LineParams P, TaxSet T, int need_fields, uint64_t n_lines, uint64_t *key_off, uint32_t *key_len,
                       uint8_t *sel, unsigned long long *err_word) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint64_t s, e;
    uint8_t ok = 0;
    uint64_t koff = 0;
    uint32_t klen = 0;
    if (line_span(P, i, &s, &e)) {
        bool high;
        if (!line_content(P, i, s, &e, &high)) {
            report_error(err_word, i, SGPU_ERR_IO);
        } else {
            const uint8_t *in = P.in;
            int f = 0;
            uint64_t fs = s, f1s = 0, f1e = 0, f2s = 0, f2e = 0;
            for (uint64_t pos = s; pos <= e; pos++) {
                if (pos == e || in[pos] == '\t') {
                    if (f == 1) { f1s = fs; f1e = pos; }
                    if (f == 2) { f2s = fs; f2e = pos; }
                    f++;
                    fs = pos + 1;
                    if (f >= need_fields) break;
                }
            }
            if (f < need_fields) {
                report_error(err_word, i, SGPU_ERR_WOULD_PANIC);  // classifier.rs:412-415 / :508-513
            } else {
                // str::trim (Unicode White_Space) on read id and taxid
                if (high) {
                    size_t b, t;
                    utf8_trim(in + f1s, f1e - f1s, &b, &t);
                    f1e = f1s + t;
                    f1s += b;
                    utf8_trim(in + f2s, f2e - f2s, &b, &t);
                    f2e = f2s + t;
                    f2s += b;
                } else {
                    while (f1s < f1e && is_ws_ascii(in[f1s])) f1s++;
                    while (f1e > f1s && is_ws_ascii(in[f1e - 1])) f1e--;
                    while (f2s < f2e && is_ws_ascii(in[f2s])) f2s++;
                    while (f2e > f2s && is_ws_ascii(in[f2e - 1])) f2e--;
                }
                if (taxid_member(T, in + f2s, (uint32_t)(f2e - f2s))) {
                    ok = 1;
                    koff = f1s;
                    klen = (uint32_t)(f1e - f1s);
                }
            }
        }
    }
    key_off[i] = koff;
    key_len[i] = klen;
    sel[i] = ok;
}

enum EvidenceKind { EV_PAF, EV_TXT, EV_READS };

static sgpu_status evidence_to_set(sgpu_ctx *c, EvidenceKind kind, const uint8_t *d_buf, size_t n, PafParams F,
                                   const TaxSet *T, int need_fields, sgpu_idset **out, uint64_t *err_line) {
    cudaStream_t st = c->stream;
    sgpu_idset *set = nullptr;
    SGPU_TRY(idset_create(c, &set));
    if (n == 0) {  // is_file_empty (alignment.rs:64,93): nothing to insert
        *out = set;
        return SGPU_OK;
    }
    sgpu_status rc = SGPU_OK;
    do {
        DevBuf<uint64_t> nlpos, key_off, errw;
        DevBuf<uint32_t> key_len;
        DevBuf<uint8_t> sel, pass, head;
        uint64_t n_nl = 0;
        if ((rc = index_newlines(c, d_buf, n, nlpos, &n_nl)) != SGPU_OK) break;
        uint64_t n_lines = n_nl + 1;  // the last thread handles an unterminated final line (if any)
        LineParams P{d_buf, (uint64_t)n, nlpos.p, n_nl};
        if ((rc = key_off.alloc(n_lines, st)) != SGPU_OK) break;
        if ((rc = key_len.alloc(n_lines, st)) != SGPU_OK) break;
        if ((rc = sel.alloc(n_lines, st)) != SGPU_OK) break;
        if ((rc = errw.alloc(1, st)) != SGPU_OK) break;
        cudaMemsetAsync(errw.p, 0xFF, 8, st);
        unsigned grid = (unsigned)ceil_div(n_lines, 128);
        if (kind == EV_PAF) {
            if ((rc = pass.alloc(n_lines, st)) != SGPU_OK) break;
            if ((rc = head.alloc(n_lines, st)) != SGPU_OK) break;
            paf_parse_kernel<<<grid, 128, 0, st>>>(P, F, n_lines, key_off.p, key_len.p, pass.p, head.p,
                                                   (unsigned long long *)errw.p);
            SGPU_LAUNCH(c);
            paf_segment_kernel<<<(unsigned)ceil_div(n_lines, 256), 256, 0, st>>>(pass.p, head.p, n_lines, sel.p);
            SGPU_LAUNCH(c);
        } else if (kind == EV_TXT) {
            txt_lines_kernel<<<grid, 128, 0, st>>>(P, n_lines, key_off.p, key_len.p, sel.p,
                                                   (unsigned long long *)errw.p);
            SGPU_LAUNCH(c);
        } else {
            reads_parse_kernel<<<grid, 128, 0, st>>>(P, *T, need_fields, n_lines, key_off.p, key_len.p, sel.p,
                                                     (unsigned long long *)errw.p);
            SGPU_LAUNCH(c);
        }
        uint64_t ew;
        if ((rc = read_u64s(c, errw.p, &ew, 1)) != SGPU_OK) break;
        if (ew != ~0ull) {
            rc = (sgpu_status)(ew & 0xFF);
            if (err_line) *err_line = ew >> 8;
            break;
        }
        rc = idset_insert_spans(c, set, d_buf, key_off.p, key_len.p, sel.p, n_lines);
        // txt: a blank line is the empty id; the tail thread of a '\n'-terminated buffer offers
        // nothing because its sel is 0
    } while (0);
    if (rc != SGPU_OK) {
        sgpu_idset_free(set);
        return rc;
    }
    *out = set;
    return SGPU_OK;
}

// H2D staging for the host-pointer entry points
static sgpu_status stage_in(sgpu_ctx *c, const uint8_t *h, size_t n, DevBuf<uint8_t> &d) {
    SGPU_TRY(d.alloc(n + 16, c->stream));
    if (n) SGPU_CUDA(cudaMemcpyAsync(d.p, h, n, cudaMemcpyHostToDevice, c->stream));
    return SGPU_OK;
}

// host stage: HashSet<String> of taxids (classifier.rs:124-252 output) -> bitmap + exotic list
static sgpu_status build_taxset(sgpu_ctx *c, const char *const *taxids, const size_t *lens, size_t n,
                                DevBuf<uint32_t> &d_bitmap, DevBuf<uint8_t> &d_exotic, DevBuf<uint32_t> &d_exoff,
                                TaxSet *T) {
    const uint64_t BIT_LIMIT = 1ull << 31;
    uint64_t maxv = 0;
    std::vector<uint64_t> vals;
    std::vector<uint8_t> exotic;
    std::vector<uint32_t> exoff{0};
    for (size_t i = 0; i < n; i++) {
        const char *s = taxids[i];
        size_t l = lens[i];
        bool canon = l >= 1 && l <= 10 && !(l > 1 && s[0] == '0');
        uint64_t v = 0;
        for (size_t k = 0; k < l && canon; k++) {
            canon = s[k] >= '0' && s[k] <= '9';
            v = v * 10 + (uint64_t)(s[k] - '0');
        }
        if (canon && v < BIT_LIMIT) {
            vals.push_back(v);
            if (v > maxv) maxv = v;
        } else {
            exotic.insert(exotic.end(), s, s + l);
            exoff.push_back((uint32_t)exotic.size());
        }
    }
    uint64_t bits = vals.empty() ? 0 : maxv + 1;
    size_t words = (size_t)((bits + 31) / 32);
    std::vector<uint32_t> bm(words ? words : 1, 0);
    for (uint64_t v : vals) bm[v >> 5] |= 1u << (v & 31);
    cudaStream_t st = c->stream;
    SGPU_TRY(d_bitmap.alloc(bm.size(), st));
    SGPU_TRY(d_exotic.alloc(exotic.size() + 1, st));
    SGPU_TRY(d_exoff.alloc(exoff.size(), st));
    SGPU_CUDA(cudaMemcpyAsync(d_bitmap.p, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice, st));
    if (!exotic.empty())
        SGPU_CUDA(cudaMemcpyAsync(d_exotic.p, exotic.data(), exotic.size(), cudaMemcpyHostToDevice, st));
    SGPU_CUDA(cudaMemcpyAsync(d_exoff.p, exoff.data(), exoff.size() * 4, cudaMemcpyHostToDevice, st));
    SGPU_CUDA(cudaStreamSynchronize(st));  // host vectors die at scope exit
    T->bitmap = d_bitmap.p;
    T->bits = bits;
    T->exotic = d_exotic.p;
    T->exotic_off = d_exoff.p;
    T->n_exotic = (uint32_t)(exoff.size() - 1);
    return SGPU_OK;
}

}  // namespace sgpu

using namespace sgpu;

extern "C" {

sgpu_status sgpu_idset_from_paf_dev(sgpu_ctx *c, const uint8_t *d_buf, size_t n, uint64_t min_len, double min_cov,
                                    uint8_t min_mapq, sgpu_idset **out, uint64_t *err_line) {
    if (!c || !out || (n && !d_buf)) return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    PafParams F{min_len, min_cov, min_mapq};
    return evidence_to_set(c, EV_PAF, d_buf, n, F, nullptr, 0, out, err_line);
}

sgpu_status sgpu_idset_from_paf(sgpu_ctx *c, const uint8_t *buf, size_t n, uint64_t min_len, double min_cov,
                                uint8_t min_mapq, sgpu_idset **out, uint64_t *err_line) {
    if (!c || !out || (n && !buf)) return SGPU_ERR_INVALID_ARG;
    DevBuf<uint8_t> d;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        SGPU_CUDA(cudaSetDevice(c->device));
        SGPU_TRY(stage_in(c, buf, n, d));
    }
    sgpu_status rc = sgpu_idset_from_paf_dev(c, d.p, n, min_len, min_cov, min_mapq, out, err_line);
    cudaStreamSynchronize(c->stream);
    return rc;
}

sgpu_status sgpu_idset_from_txt_dev(sgpu_ctx *c, const uint8_t *d_buf, size_t n, sgpu_idset **out,
                                    uint64_t *err_line) {
    if (!c || !out || (n && !d_buf)) return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    return evidence_to_set(c, EV_TXT, d_buf, n, PafParams{0, 0.0, 0}, nullptr, 0, out, err_line);
}

sgpu_status sgpu_idset_from_txt(sgpu_ctx *c, const uint8_t *buf, size_t n, sgpu_idset **out, uint64_t *err_line) {
    if (!c || !out || (n && !buf)) return SGPU_ERR_INVALID_ARG;
    DevBuf<uint8_t> d;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        SGPU_CUDA(cudaSetDevice(c->device));
        SGPU_TRY(stage_in(c, buf, n, d));
    }
    sgpu_status rc = sgpu_idset_from_txt_dev(c, d.p, n, out, err_line);
    cudaStreamSynchronize(c->stream);
    return rc;
}

sgpu_status sgpu_idset_from_reads_dev(sgpu_ctx *c, const uint8_t *d_buf, size_t n, int style,
                                      const char *const *taxids, const size_t *taxid_lens, size_t n_taxids,
                                      sgpu_idset **out, uint64_t *err_line) {
    if (!c || !out || (n && !d_buf) || (n_taxids && (!taxids || !taxid_lens)) || (style != 0 && style != 1))
        return SGPU_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SGPU_CUDA(cudaSetDevice(c->device));
    DevBuf<uint32_t> bm, exoff;
    DevBuf<uint8_t> ex;
    TaxSet T;
    SGPU_TRY(build_taxset(c, taxids, taxid_lens, n_taxids, bm, ex, exoff, &T));
    return evidence_to_set(c, EV_READS, d_buf, n, PafParams{0, 0.0, 0}, &T, style == 0 ? 5 : 7, out, err_line);
}

sgpu_status sgpu_idset_from_reads(sgpu_ctx *c, const uint8_t *buf, size_t n, int style, const char *const *taxids,
                                  const size_t *taxid_lens, size_t n_taxids, sgpu_idset **out, uint64_t *err_line) {
    if (!c || !out || (n && !buf)) return SGPU_ERR_INVALID_ARG;
    DevBuf<uint8_t> d;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        SGPU_CUDA(cudaSetDevice(c->device));
        SGPU_TRY(stage_in(c, buf, n, d));
    }
    sgpu_status rc = sgpu_idset_from_reads_dev(c, d.p, n, style, taxids, taxid_lens, n_taxids, out, err_line);
    cudaStreamSynchronize(c->stream);
    return rc;
}

}  // extern "C"

// fastq_general.cu -- the GENERAL (always exact) FASTQ filter path.
//
// Replaces FastqCleaner::clean_reads (cleaner.rs:731-760), needletail 0.5.1's fastq
// Reader::next/validate/check_end + write_fastq, and get_id (utils.rs:91-103).
//
// Pipeline (all on the device):
//   1. '\n' index (scan.cu)                       -> nlpos[]
//   2. fastq_record_kernel: thread per record     -> validation, trimmed spans, id, probe,
//                                                    normalised output length
//   3. exclusive scans of the written / other lengths
//   4. fastq_copy_kernel: warp per record         -> '@' id E seq E '+' E qual E
// It handles CRLF, "+id" separator lines, a missing final newline, blank tails, non-ASCII
// headers, every parse error class, and shards cut at arbitrary byte offsets.  The fused
// single-pass kernel (fastq_fused.cu) takes over when the input is canonical.
#include "fastq_records.cuh"

namespace sgpu {

__global__ void __launch_bounds__(128)
    fastq_record_kernel(RecParams P, RecMeta *meta, uint32_t *len_w, uint32_t *len_o, unsigned long long *err_word) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > P.k_full) return;
    uint32_t lw = 0, lo = 0;
    RecMeta m = {0, 0, 0, 0};
    uint64_t start, p1, p3;
    size_t id_off = 0, id_len = 0;
    if (locate_record(P, k, &start, &p1, &p3, &m, &id_off, &id_len, err_word)) {
        bool hit = id_len <= IDSET_MAX_KEY && idset_contains(P.set, P.in + start + 1 + id_off, (uint32_t)id_len);
        bool written = P.reverse ? hit : !hit;
        uint32_t e = P.crlf ? 2u : 1u;
        uint32_t out_len = 2u + m.id_n + m.seq_n + m.qual_n + 4u * e;
        m.flags = 1u | (written ? 2u : 0u);
        if (written) lw = out_len; else lo = out_len;
    }
    meta[k] = m;
    len_w[k] = lw;
    len_o[k] = lo;
}

// records at or after the first error are not produced (the reference stops there)
__global__ void fastq_mask_kernel(RecMeta *meta, uint32_t *len_w, uint32_t *len_o, uint64_t from, uint64_t n) {
    uint64_t k = from + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    meta[k].flags = 0;
    len_w[k] = 0;
    len_o[k] = 0;
}

__global__ void fastq_count_kernel(const RecMeta *meta, uint64_t n, unsigned long long *counters) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t f = k < n ? meta[k].flags : 0;
    // one pair of atomics per block (per warp they queue up on two addresses: 0.2 ms per 5 M records)
    const int bi = __syncthreads_count(f & 1u), bo = __syncthreads_count(f & 2u);
    if (threadIdx.x == 0) {
        if (bi) atomicAdd(&counters[0], (unsigned long long)bi);
        if (bo) atomicAdd(&counters[1], (unsigned long long)bo);
    }
}

__global__ void __launch_bounds__(256)
    fastq_copy_kernel(RecParams P, const RecMeta *meta, const uint64_t *off_w, const uint64_t *off_o, uint8_t *out_w,
                      uint8_t *out_o, uint64_t n_rec) {
    uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (k >= n_rec) return;
    RecMeta m = meta[k];
    if (!(m.flags & 1u)) return;
    uint8_t *dst;
    if (m.flags & 2u) {
        dst = out_w + off_w[k];
    } else {
        if (!out_o) return;
        dst = out_o + off_o[k];
    }
    uint64_t b = P.b0 + 4 * k;
    uint64_t start = (b == 0) ? 0 : P.nlpos[b - 1] + 1;
    uint64_t p1 = P.nlpos[b], p3 = P.nlpos[b + 2];
    const uint8_t *in = P.in;
    const uint32_t e = P.crlf ? 2u : 1u;
    // '@' id E
    if (lane == 0) dst[0] = '@';
    warp_copy(dst + 1, in + start + 1, m.id_n, lane);
    uint8_t *q = dst + 1 + m.id_n;
    if (lane == 0) {
        if (P.crlf) q[0] = '\r';
        q[e - 1] = '\n';
    }
    q += e;
    // seq E '+' E
    warp_copy(q, in + p1 + 1, m.seq_n, lane);
    q += m.seq_n;
    if (lane == 0) {
        uint32_t w = 0;
        if (P.crlf) q[w++] = '\r';
        q[w++] = '\n';
        q[w++] = '+';
        if (P.crlf) q[w++] = '\r';
        q[w++] = '\n';
    }
    q += 2 * e + 1;
    // qual E
    warp_copy(q, in + p3 + 1, m.qual_n, lane);
    q += m.qual_n;
    if (lane == 0) {
        if (P.crlf) q[0] = '\r';
        q[e - 1] = '\n';
    }
}

// Runs the general path.  Device pointers in, device outputs filled, counts on the host.
sgpu_status clean_general(sgpu_ctx *c, const sgpu_idset *set, const uint8_t *d_in, size_t n_in, size_t own_len,
                          uint64_t newlines_before, int is_first, int is_last, int crlf_in, int reverse,
                          uint8_t *d_out_w, size_t cap_w, size_t *n_w, uint8_t *d_out_o, size_t cap_o, size_t *n_o,
                          sgpu_counts *counts) {
    cudaStream_t st = c->stream;
    counts->path = 2;
    DevBuf<uint64_t> nlpos;
    uint64_t n_nl = 0;
    SGPU_TRY(index_newlines(c, d_in, n_in, nlpos, &n_nl));

    RecParams P;
    P.in = d_in;
    P.n_in = n_in;
    P.own_len = own_len;
    P.nlpos = nlpos.p;
    P.is_last = is_last;
    P.reverse = reverse;
    P.set = view_of(set);
    bool any = setup_records(P, n_nl, newlines_before, is_first);
    int crlf = crlf_in;
    if (is_first && crlf_in < 0) {
        // find_line_ending on the first record: is its first '\n' preceded by '\r'?
        crlf = 0;
        if (n_nl >= 1) {
            uint64_t p0;
            SGPU_TRY(read_u64s(c, nlpos.p, &p0, 1));
            if (p0 > 0) {
                uint8_t ch;
                SGPU_CUDA(cudaMemcpyAsync(c->h_pinned, d_in + p0 - 1, 1, cudaMemcpyDeviceToHost, st));
                SGPU_CUDA(cudaStreamSynchronize(st));
                ch = *(uint8_t *)c->h_pinned;
                crlf = ch == '\r';
            }
        }
    }
    P.crlf = crlf;
    counts->crlf = (uint32_t)crlf;
    if (!any) {
        // a non-first shard without a single record boundary: it owns no record start
        *n_w = 0;
        if (n_o) *n_o = 0;
        return SGPU_OK;
    }
    uint64_t n_thr = P.k_full + 1;  // + the tail candidate

    DevBuf<RecMeta> meta;
    DevBuf<uint32_t> len_w, len_o;
    DevBuf<uint64_t> off_w, off_o, scratch;
    SGPU_TRY(meta.alloc(n_thr, st));
    SGPU_TRY(len_w.alloc(n_thr, st));
    SGPU_TRY(len_o.alloc(n_thr, st));
    SGPU_TRY(off_w.alloc(n_thr, st));
    SGPU_TRY(scratch.alloc(8, st));  // [0] err word, [1] total_w, [2] total_o, [3] reads_in, [4] reads_out
    uint64_t init[8] = {~0ull, 0, 0, 0, 0, 0, 0, 0};
    memcpy(c->h_pinned + 32, init, sizeof(init));
    SGPU_CUDA(cudaMemcpyAsync(scratch.p, c->h_pinned + 32, sizeof(init), cudaMemcpyHostToDevice, st));

    fastq_record_kernel<<<(unsigned)ceil_div(n_thr, 128), 128, 0, st>>>(P, meta.p, len_w.p, len_o.p,
                                                                         (unsigned long long *)scratch.p);
    SGPU_LAUNCH(c);
    uint64_t err_word;
    SGPU_TRY(read_u64s(c, scratch.p, &err_word, 1));
    sgpu_status rc = SGPU_OK;
    if (err_word != ~0ull) {
        rc = (sgpu_status)(err_word & 0xFF);
        counts->error_record = err_word >> 8;
        if (rc == SGPU_ERR_HALO) return rc;
        uint64_t from = err_word >> 8;
        fastq_mask_kernel<<<(unsigned)ceil_div(n_thr - from, 256), 256, 0, st>>>(meta.p, len_w.p, len_o.p, from,
                                                                                 n_thr);
        SGPU_LAUNCH(c);
    }
    fastq_count_kernel<<<(unsigned)ceil_div(n_thr, 256), 256, 0, st>>>(meta.p, n_thr,
                                                                        (unsigned long long *)(scratch.p + 3));
    SGPU_LAUNCH(c);
    SGPU_TRY(exclusive_scan_u32_to_u64(c, len_w.p, off_w.p, n_thr, scratch.p + 1));
    if (d_out_o) {
        SGPU_TRY(off_o.alloc(n_thr, st));
        SGPU_TRY(exclusive_scan_u32_to_u64(c, len_o.p, off_o.p, n_thr, scratch.p + 2));
    }
    uint64_t h[5];
    SGPU_TRY(read_u64s(c, scratch.p, h, 5));
    *n_w = (size_t)h[1];
    if (n_o) *n_o = d_out_o ? (size_t)h[2] : 0;
    counts->reads_in = h[3];
    counts->reads_out = h[4];
    if (h[1] > cap_w || (d_out_o && h[2] > cap_o)) return SGPU_ERR_CAPACITY;
    fastq_copy_kernel<<<(unsigned)ceil_div(n_thr * 32, 256), 256, 0, st>>>(P, meta.p, off_w.p, off_o.p, d_out_w,
                                                                           d_out_o, n_thr);
    SGPU_LAUNCH(c);
    SGPU_CUDA(cudaGetLastError());
    SGPU_CUDA(cudaStreamSynchronize(st));  // scratch buffers are released after this call returns
    return rc;
}

}  // namespace sgpu

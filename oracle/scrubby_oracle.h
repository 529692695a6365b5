/*
 * scrubby_oracle.h -- CPU oracle for the scrubby depletion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or
 * executed by the product (scrubby_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker or as the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (esteinig/scrubby 1.0.2) ships no tests, no
 * golden vectors and cannot be built here (no Rust toolchain, un-vendored
 * needletail 0.5.1 / niffler 2.5.0).  This file is a restatement of the
 * reference's algorithm from its source, plus needletail's published FASTQ
 * framing/writing rules; the vectors in tests/golden/ are hand-derived from the
 * cited lines and cross-checked against a second,
 * independently written Python restatement (oracle/pyoracle.py).
 */
#ifndef SCRUBBY_ORACLE_H
#define SCRUBBY_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: numerically identical to include/scrubby_gpu.h (checked by a test) */
enum {
    ORC_OK = 0,
    ORC_ERR_IO = 1,                    /* ScrubbyError::IoError (invalid UTF-8 from BufRead::lines) */
    ORC_ERR_FASTQ_INVALID_START = 3,   /* needletail ParseError InvalidStart */
    ORC_ERR_FASTQ_INVALID_SEPARATOR = 4,
    ORC_ERR_FASTQ_UNEQUAL_LENGTHS = 5,
    ORC_ERR_FASTQ_UNEXPECTED_END = 6,
    ORC_ERR_FASTQ_UNKNOWN_FORMAT = 7,
    ORC_ERR_RECORD_NAME_UTF8 = 8,      /* ScrubbyError::RecordNameUtf8Error, utils.rs:92 */
    ORC_ERR_FASTQ_HEADER = 9,          /* ScrubbyError::NeedletailFastqHeader, utils.rs:97-99 */
    ORC_ERR_PAF_INTEGER = 10,          /* ScrubbyError::PafRecordIntegerError */
    ORC_ERR_WOULD_PANIC = 11,          /* reference indexes fields[] out of bounds */
    ORC_ERR_KRAKEN_REPORT_READS = 12,  /* KrakenReportReadFieldConversion */
    ORC_ERR_KRAKEN_REPORT_DIRECT = 13, /* KrakenReportDirectReadFieldConversion */
    ORC_ERR_KRAKEN_REPORT_PARENT = 14, /* KrakenReportTaxonParent */
    ORC_ERR_FASTA_UNSUPPORTED = 15,
    ORC_ERR_SAM_RECORD = 22,           /* htslib sam_parse1 would reject the line (rust_htslib::errors::Error) */
    ORC_ERR_BAM_RECORD = 23            /* htslib bam_hdr_read / bam_read1 would fail (bad magic, truncated or inconsistent record) */
};

typedef struct orc_set orc_set;

typedef struct {
    uint64_t reads_in;      /* records parsed from the input                           */
    uint64_t reads_out;     /* records written (clean) / records in the output (diff)  */
    uint64_t difference;    /* diff only: input records whose id is absent from output */
    uint64_t error_record;  /* index of the record/line that raised the error          */
    uint32_t crlf;          /* 1 if the first record's first line ends with CRLF        */
    uint32_t empty_input;   /* 1 if the input is "empty" per utils.rs:359-375          */
} orc_counts;

orc_set *orc_set_new(void);
void orc_set_free(orc_set *);
void orc_set_insert(orc_set *, const uint8_t *key, size_t len);
int orc_set_contains(const orc_set *, const uint8_t *key, size_t len);
uint64_t orc_set_len(const orc_set *);
/* sorted (bytewise) keys, each followed by '\n'; caller frees with orc_free */
int orc_set_dump_sorted(const orc_set *, uint8_t **out, size_t *n);
void orc_free(void *);

int orc_get_id(const uint8_t *header, size_t len, size_t *off, size_t *id_len);

int orc_set_from_paf(const uint8_t *buf, size_t n, uint64_t min_len, double min_cov,
                     uint8_t min_mapq, orc_set **out, uint64_t *err_line);
int orc_set_from_txt(const uint8_t *buf, size_t n, orc_set **out, uint64_t *err_line);
/* alignment.rs:117-146 from_bam (+ BamRecord :154-211) restated for text SAM */
int orc_set_from_sam(const uint8_t *buf, size_t n, uint64_t min_len, double min_cov, uint8_t min_mapq,
                     orc_set **out, uint64_t *err_line);
/* alignment.rs:117-146 from_bam for binary BAM records (the BGZF-decompressed stream) */
int orc_set_from_bam(const uint8_t *buf, size_t n, uint64_t min_len, double min_cov, uint8_t min_mapq,
                     orc_set **out, uint64_t *err_record);
int orc_taxids_from_report(const uint8_t *buf, size_t n, const char *const *taxa, size_t n_taxa,
                           const char *const *taxa_direct, size_t n_direct, orc_set **out,
                           uint64_t *err_line);
/* style 0 = kraken2 (>=5 columns), 1 = metabuli (>=7 columns) */
int orc_set_from_reads(const uint8_t *buf, size_t n, int style, const orc_set *taxids,
                       orc_set **out, uint64_t *err_line);

/* out_other may be NULL.  Buffers must hold 2*n_in + 16 bytes. */
int orc_clean_fastq(const uint8_t *in, size_t n_in, const orc_set *set, int reverse,
                    uint8_t *out_written, size_t *n_written, uint8_t *out_other,
                    size_t *n_other, orc_counts *counts);

/* counts accumulate (+=) as in utils.rs:250-285; diff_ids accumulates across file pairs */
int orc_diff(const uint8_t *in, size_t n_in, const uint8_t *out, size_t n_out,
             orc_counts *counts, orc_set *diff_ids);

#ifdef __cplusplus
}
#endif
#endif

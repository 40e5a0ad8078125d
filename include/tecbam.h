/*
 * libtecbam -- BAM (BGZF) file -> structure-of-arrays read batches, on the host cores.
 *
 * the row next to the counting path.  It replaces the per-record Python work of
 * the reference's read loops -- pysam iteration plus attribute access, tag lookup and string
 * handling at te_count/te_count.py:65-98 (paired end), :190-214 (single end), :351-438 (single
 * cell) -- with one call per batch that fills exactly the arrays include/tecount.h takes
 * (tec_bulk_push / tec_sc_push: i32 start, i32 end, u16 chrom, u8 mapq, u8 flag bits, u32 cell id,
 * u64 UMI code).  No CUDA in here: BGZF blocks are inflated (zlib) and records parsed by a pool of
 * host threads; the output buffers are normally the pinned buffers of tec_host_alloc, so the
 * library writes straight into the memory the GPU copies from.
 *
 * Conventions as in tecount.h: plain C ABI, status return (0 / negative tbam_status), the caller
 * owns every buffer it passes, one reader per file, not thread-safe (the reader's own worker
 * threads are internal).  reference_end comes from the CIGAR field of the record.  Alignments of more than
 * 65535 operations, which BAM stores as the placeholder <l_seq>S<ref_len>N plus a CG:B,I tag (SAMv1 4.2.2), need no
 * expansion for that: the placeholder's N operation carries the reference length (tests/test_bam_htslib_shapes.py).
 * The host-side meaning of every field is the one of
 * te_counter_b200/reads.py (the Python packing this library replaces) and is cited per entry point.
 */
#ifndef TECBAM_H
#define TECBAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBAM_ABI_VERSION 1

typedef struct tbam_reader tbam_reader;

enum tbam_status {
    TBAM_OK = 0,
    TBAM_E_IO = -1,             /* open / mmap failed */
    TBAM_E_FORMAT = -2,         /* not a BAM stream, corrupt block or record, CRC mismatch, truncated */
    TBAM_E_NOT_BGZF = -3,       /* gzip without BGZF block sizes (or not gzip): use another reader */
    TBAM_E_ARG = -4,
    /* conditions under which the reference's read loop raises; the caller re-raises the same class */
    TBAM_E_NO_BARCODE_TAG = -10,    /* te_count.py:403-409  AssertionError('CB or CR tag not found!') */
    TBAM_E_NO_UMI_TAG = -11,        /* te_count.py:420-426  AssertionError('UB or UR tag not found!') */
    TBAM_E_UMI = -12,               /* UMI longer than 21 characters or outside A,C,G,N,T (reads.encode_umi) */
    TBAM_E_END_NONE = -13,          /* reference_end is None on a record that would be counted (te_count.py:223) */
    TBAM_E_CHROM_NAME = -14,        /* --sc record on a chromosome whose name holds ':' (reads.ChromMap.sc_id) */
    TBAM_E_REF_NONE = -15           /* --sc record that passes the filters but has no reference (te_count.py:431) */
};

/* values of the chrom column; ids below n_index are index chromosomes (reads.py) */
#define TBAM_CHROM_INVALID  0xFFFFu     /* record without a reference (bulk) / filtered record (sc) */
#define TBAM_CHROM_SC_SKIP  0xFFFEu     /* '_' or 'alt' in the chromosome key: silent skip, te_count.py:432 */
#define TBAM_CHROM_SC_BAD   0xFFFDu     /* in the sc map only: name the reference cannot key, -> TBAM_E_CHROM_NAME */
#define TBAM_CELL_INVALID   0xFFFFFFFFu

/* flag column bits (tecount.h TEC_F_*): 1 unmapped, 2 duplicate, 4 QC fail, 8 reverse,
 * 16 mate names differ (paired end, te_count.py:92) */

int tbam_abi_version(void);
const char *tbam_strerror(int status);

/* Opens `path`, reads the BAM header.  n_threads <= 0: one per online core (at most 64). */
int tbam_open(const char *path, int n_threads, tbam_reader **out);
void tbam_close(tbam_reader *r);
/* Text of the last failure on this reader (names the record index where that applies). */
const char *tbam_last_error(const tbam_reader *r);

/* Header reference sequences, in file order: pysam's AlignmentFile.references. */
int tbam_n_references(const tbam_reader *r);
const char *tbam_reference_name(const tbam_reader *r, int i);

/* Column value for each reference id: `bulk_ids[i]` / `sc_ids[i]` is what reads.ChromMap.bulk_id /
 * .sc_id give for reference i (te_count.py:96, :212, :431-432).  n must be tbam_n_references();
 * n_index = number of chromosomes of the annotation index (ids below it are "in the index"). */
int tbam_set_chrom_map(tbam_reader *r, const uint16_t *bulk_ids, const uint16_t *sc_ids, int32_t n, int32_t n_index);

/* Barcode whitelist: the n distinct barcodes, sorted (te_count.py:330-339), concatenated without
 * separators; barcode i is bytes [offsets[i], offsets[i+1]).  Its position i is the cell id. */
int tbam_set_whitelist(tbam_reader *r, const char *barcodes, const int64_t *offsets, int32_t n);

/*
 * Next batch of bulk records (reads.fill_bulk; te_count.py:76-102 paired, :203-218 single end).
 * Fills up to `capacity` records (an even number when paired: records are consumed two at a time
 * in file order and a trailing unpaired record is dropped, te_count.py:79).  *n_out = records
 * written, *more = 0 once the file is exhausted.  `qual` is only used for the two checks the
 * host makes before the GPU filter: the mate-name test of pairs that pass the filters
 * (te_count.py:92) and TBAM_E_END_NONE.
 */
int tbam_next_bulk(tbam_reader *r, int paired, int qual, int64_t capacity,
                   int32_t *start, int32_t *end, uint16_t *chrom, uint8_t *mapq, uint8_t *flag,
                   int64_t *n_out, int *more);

/*
 * Next batch of single-cell records (reads.fill_sc; te_count.py:393-438): as above plus
 * cell = whitelist id of the CB (else CR) tag or TBAM_CELL_INVALID, umi = order-preserving
 * 3-bit-per-character code of the UB (else UR) tag.  Records that fail the flag / MAPQ filter
 * keep mapq and flag and get start = end = -1, chrom = TBAM_CHROM_INVALID, cell invalid, umi 0;
 * their tags are not looked at (the reference only raises for records that pass).
 */
int tbam_next_sc(tbam_reader *r, int qual, int64_t capacity,
                 int32_t *start, int32_t *end, uint16_t *chrom, uint8_t *mapq, uint8_t *flag,
                 uint32_t *cell, uint64_t *umi, int64_t *n_out, int *more);

/* Counters since open: what = 0 records delivered, 1 compressed bytes consumed,
 * 2 uncompressed bytes produced, 3 worker threads, 4 nanoseconds spent inside tbam_next_*. */
int64_t tbam_counter(const tbam_reader *r, int what);

/* One raw-deflate stream (the payload of a BGZF block) into exactly n_out bytes.  engine 0: zlib;
 * engine 1: the decoder's own table-driven inflate only (bamdecode.cpp uses it first and falls back
 * to zlib whenever it declines).  Returns 0, or TBAM_E_FORMAT when the engine rejects the stream.
 * Exposed so that the two can be held against each other (tests/test_fast_inflate.py). */
int tbam_inflate_raw(const void *in, int64_t n_in, void *out, int64_t n_out, int engine);

#ifdef __cplusplus
}
#endif
#endif

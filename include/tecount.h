/*
 * tecount.h -- C ABI of libtecount.so, the B200 (sm_100a) counting engine behind te_count.
 *
 * The reference (oaxiom/te_counter) is pure Python and has no FFI; the seam this library sits
 * behind is the `measureTE` method set (reference te_count/te_count.py:14-754) called from
 * bin/te_count:77-120.  Each entry point below names the reference lines whose work it takes
 * over.  Python binds it with ctypes (te_counter_b200/_lib.py); INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every call returns 0 (TEC_OK) or a negative tec_status; nothing throws across the ABI;
 *     tec_last_error(ctx) gives the text of the last failure on that context.
 *   - plain pointers and sizes only.  The caller owns every host buffer.  The library owns all
 *     device memory inside the context, except where a parameter is documented as a device
 *     pointer (`*_dev` entry points), which the caller (e.g. a torch tensor) owns.
 *   - one context per GPU; a context is not thread-safe.  Work is enqueued on the context's
 *     stream; calls that return results to host memory synchronise before returning.
 *   - there is no CPU fallback: tec_create fails when no CUDA device is usable.
 *
 * Record layout (structure of arrays, one element per alignment record):
 *   int32  start   pysam reference_start (0-based)
 *   int32  end     pysam reference_end   (exclusive)
 *   uint16 chrom   index chromosome id (< n_chrom of tec_index_upload); any other value means
 *                  "not in the index"; 0xFFFE = name contains '_' or 'alt' (single-cell skip)
 *   uint8  mapq
 *   uint8  flag    bit0 unmapped(0x4) bit1 duplicate(0x400) bit2 qcfail(0x200) bit3 reverse(0x10)
 *                  bit4 paired-end query-name mismatch (te_count.py:92)
 *   uint32 cell    [single cell] id of the barcode in the sorted whitelist, 0xFFFFFFFF = not in it
 *   uint64 umi     [single cell] order-preserving code of the UMI string
 */
#ifndef TECOUNT_H
#define TECOUNT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEC_ABI_VERSION 1

typedef struct tec_ctx tec_ctx;

typedef enum tec_status {
    TEC_OK = 0,
    TEC_ERR_CUDA = -1,          /* a CUDA runtime call failed (text in tec_last_error) */
    TEC_ERR_ARG = -2,           /* bad argument */
    TEC_ERR_STATE = -3,         /* call out of order (no index, no begin, ...) */
    TEC_ERR_NOMEM = -4,         /* device or host allocation failed */
    TEC_ERR_LIMIT = -5,         /* input exceeds a documented limit (key width, record count) */
    TEC_ERR_UNIMPLEMENTED = -6,
    TEC_ERR_IO = -7,            /* tec_bam_open: the file cannot be opened or mapped */
    TEC_ERR_FORMAT = -8,        /* tec_bam_*: corrupt or truncated BAM / BGZF data */
    TEC_ERR_UNSUPPORTED = -9,   /* tec_bam_*: a file the device decoder refuses (not BGZF, or record boundaries it
                                   cannot establish block-parallel): decode it with libtecbam instead */
    /* tec_bam_count: a record on which the reference's read loop raises (same numbers as tecbam.h) */
    TEC_ERR_BAM_NO_BARCODE_TAG = -10, TEC_ERR_BAM_NO_UMI_TAG = -11, TEC_ERR_BAM_UMI = -12, TEC_ERR_BAM_END_NONE = -13,
    TEC_ERR_BAM_CHROM_NAME = -14, TEC_ERR_BAM_REF_NONE = -15
} tec_status;

/* feature type codes: which branch of te_count.py:134-146 / :662-682 a feature can trigger */
#define TEC_T_OTHER 0
#define TEC_T_GENE 1            /* protein_coding, lincRNA, lncRNA */
#define TEC_T_TE 2
#define TEC_T_SNRNA 3
#define TEC_T_ENHANCER 4
#define TEC_STRAND_MISSING 255  /* feature row has no 'strand' key */

#define TEC_F_UNMAPPED 1
#define TEC_F_DUP 2
#define TEC_F_QCFAIL 4
#define TEC_F_REVERSE 8
#define TEC_F_NAME_MISMATCH 16
#define TEC_CHROM_SC_SKIP 0xFFFE
#define TEC_CELL_INVALID 0xFFFFFFFFu

/* bulk statistics block, int64[TEC_BULK_NSTATS]  (te_count.py:158-163, :269-275) */
#define TEC_BULK_NSTATS 8
#define TEC_BS_UNITS 0          /* records (SE) or pairs (PE) consumed; total_reads = this + 1 */
#define TEC_BS_ASSIGNED 1       /* "Reads were assigned to a gene" */
#define TEC_BS_LOWQ 2           /* "Read quality is too low" */
#define TEC_BS_BADCHROM 3       /* "Reads mapped to an invalid chromosome" */
#define TEC_BS_QCFAIL 4         /* "Reads are QC fails" */
#define TEC_BS_CRASH_ENHANCER 5 /* units at which the reference raises NameError (te_count.py:147) */
#define TEC_BS_CRASH_NAME 6     /* units at which the reference dies in sys.quit (te_count.py:94) */

/* single-cell statistics block, int64[TEC_SC_NSTATS]  (te_count.py:696-705) */
#define TEC_SC_NSTATS 16
#define TEC_SS_UNITS 0          /* records consumed; total_reads = this + 1 */
#define TEC_SS_INVALID_BARCODE 1
#define TEC_SS_ALREADY_SEEN 2
#define TEC_SS_LOWQ 3
#define TEC_SS_QCFAIL 4
#define TEC_SS_VALID 5          /* lines kept by Part 2 ("total valid reads") */
#define TEC_SS_ASSIGNED 6       /* "Assigned N of total valid reads to features" */
#define TEC_SS_RAW_BARCODES 7   /* "Observed N raw barcodes" */
#define TEC_SS_BUNDLES 8        /* spill bundles the reference would have written */
#define TEC_SS_CRASH_STRAND 9   /* fragments at which the reference raises KeyError 'strand' (:661) */
#define TEC_SS_SURVIVORS 10     /* records that reached the UMI table */
#define TEC_SS_SEGMENTS 11      /* (cell, UMI, bundle) lines over all bundles */

/* ---- life cycle ------------------------------------------------------------------------- */
int tec_abi_version(void);
const char* tec_strerror(int status);
/* device: CUDA ordinal.  Replaces measureTE.__init__ state (te_count.py:15-29). */
int tec_create(int device, tec_ctx** out);
void tec_destroy(tec_ctx* ctx);
const char* tec_last_error(const tec_ctx* ctx);
int tec_sync(tec_ctx* ctx);
/* pinned host memory for the SoA batches (optional; any host memory is accepted) */
int tec_host_alloc(tec_ctx* ctx, uint64_t bytes, void** out);
int tec_host_free(tec_ctx* ctx, void* p);
/* the CUDA stream of the context as a cudaStream_t value (for event timing by the caller) */
void* tec_stream(tec_ctx* ctx);
/* elapsed GPU milliseconds of the kernels of the last bulk push / sc finalize on this context */
float tec_last_kernel_ms(tec_ctx* ctx);
/* number of kernel launches issued by this context since creation */
int64_t tec_launch_count(const tec_ctx* ctx);

/* release cached device scratch (the single-cell finalize keeps its temporaries for the next call) */
int tec_trim(tec_ctx* ctx);

/* tuning knobs: "bulk_algo" (-1 auto, 0 exact search kernel, 1 round-1 cell-table kernel -- set before
 * tec_index_upload --, 2 two-pass kernels), "stab_shift" (log2 of the cell size, 8..11, or 0 = default: 10 for
 * bulk, 11 for the single-cell pair table; used by the next tec_index_upload), "bulk_mode" (bit 0 table sectors
 * evict_last in L2, bit 1 sector prefetch, bit 2 tally through the per-warp hit queue, bit 3 deep pipeline, bit 4
 * queue filled behind a warp prefix sum, bit 5 768-thread CTAs, bit 6 queue drained once per tile; default 77),
 * "second_mode" (second bulk pass: the register set of a unit's distinct ensg, 0 stored by position, 1 shifted
 * in, 2 = 1 + the units that two sectors answer go through a straight-line kernel first), "second_parts" (warps
 * per segment of the deferred list), "ctas_per_sm",
 * "all_hot" (counters of every ensg in shared memory when they fit), "sc_algo" (-1 auto, 0 exact search in
 * Part 3, 1 cell table), "sc_pack_umi" (2-bit UMI sort keys when possible).
 * tec_get_info: "has_stab", "stab_bytes", "has_sc_stab", "sc_stab_bytes", "n_sm", "n_features",
 * "stab_primary", "stab_overflow", "stab_entries", "last_slow_units", "last_deferred_units", "last_left_units"
 * (deferred units the two-sector kernel left for the second pass), "stab_refused";
 * -1 for unknown keys. */
int tec_set_option(tec_ctx* ctx, const char* key, int64_t value);
int64_t tec_get_info(tec_ctx* ctx, const char* key);
/* why the last tec_index_upload could not build the bulk cell table ("" when it could): with such an index
 * (more than 65535 ensg, or an ensg that carries two feature types) bulk counting runs on the exact search kernel,
 * about ten times slower.  The caller should log it (te_counter_b200/te_count.py does). */
const char* tec_index_note(const tec_ctx* ctx);

/* ---- index ------------------------------------------------------------------------------
 * Replaces load_genome() + the structures of genelist._optimiseData that the read loops reach
 * into (te_count.py:31-35, :68-73, :586-590; miniglbase/genelist.py:332-396).
 * Features are grouped by chromosome id and sorted by (L, R) inside each chromosome:
 *   chrom_off[n_chrom + 1]  offsets into the feature arrays
 *   L, R                    loc['left'], loc['right']
 *   ensg_id                 index into sorted(set(ensg))  (== output row / column order)
 *   type_code               TEC_T_*;  strand_code 0 '+', 1 '-', 2..3 other strings, 255 missing
 *   bucket_size             miniglbase/config.py:36 (10000)
 */
int tec_index_upload(tec_ctx* ctx, int32_t n_chrom, const int64_t* chrom_off,
                     const int32_t* L, const int32_t* R, const int32_t* ensg_id,
                     const uint8_t* type_code, const uint8_t* strand_code,
                     int32_t n_ensg, int32_t bucket_size);

/* ---- bulk (parse_bamse te_count.py:167-277, parse_bampe te_count.py:42-165) ---------------
 * begin: zero the per-feature counters and statistics.  paired != 0 consumes records two at a
 * time (record 2i = read1, 2i+1 = read2); every push must then hold an even number of records.
 * push: filter + mate merge + point overlap + type rule + tally for one batch, host buffers.
 * push_dev: the same with DEVICE pointers, asynchronous on the context stream.
 * finish: counts[n_ensg] and stats[TEC_BULK_NSTATS] to host (either may be NULL).
 */
int tec_bulk_begin(tec_ctx* ctx, int paired, int qual);
int tec_bulk_push(tec_ctx* ctx, int64_t n_rec, const int32_t* start, const int32_t* end,
                  const uint16_t* chrom, const uint8_t* mapq, const uint8_t* flag);
int tec_bulk_push_dev(tec_ctx* ctx, int64_t n_rec, const int32_t* start, const int32_t* end,
                      const uint16_t* chrom, const uint8_t* mapq, const uint8_t* flag);
int tec_bulk_finish(tec_ctx* ctx, int64_t* counts, int64_t* stats);
/* device address of the int64 counters [n_ensg] followed by stats [TEC_BULK_NSTATS]; lets the
 * caller run the cross-GPU reduction (NCCL) on the library's buffer without a host round trip.
 * The counters are in the library's internal order (the same on every rank for the same index);
 * tec_bulk_finish returns them in ensg-id order. */
void* tec_bulk_counts_dev(tec_ctx* ctx);
/* ---- single cell (sc_parse_bamse te_count.py:298-707, sc_save_result :709-754) ------------ */
int tec_sc_begin(tec_ctx* ctx, int qual, int strand, int64_t n_whitelist);
int tec_sc_push(tec_ctx* ctx, int64_t n_rec, const int32_t* start, const int32_t* end,
                const uint16_t* chrom, const uint8_t* mapq, const uint8_t* flag,
                const uint32_t* cell, const uint64_t* umi);
int tec_sc_push_dev(tec_ctx* ctx, int64_t n_rec, const int32_t* start, const int32_t* end,
                    const uint16_t* chrom, const uint8_t* mapq, const uint8_t* flag,
                    const uint32_t* cell, const uint64_t* umi);
/* Parts 1-3: bundle_keys is the 1e7 of te_count.py:377, pad the +1000 of :502.
 * Returns the sizes of the two result lists. */
int tec_sc_finalize(tec_ctx* ctx, int64_t bundle_keys, int64_t maxcells, int64_t pad,
                    int64_t* n_triples, int64_t* n_hit_cells);
/* triples sorted by (ensg, cell): final_results[ensg][barcode] = count (te_count.py:668-682);
 * hit cells ascending by cell id: self.barcodes after Part 3 (te_count.py:653-655);
 * stats[TEC_SC_NSTATS].  Any pointer may be NULL. */
int tec_sc_fetch(tec_ctx* ctx, int32_t* ensg, uint32_t* cell, int64_t* count,
                 uint32_t* hit_cell, int64_t* hit_count, int64_t* stats);
/* ---- single cell on several GPUs (one process per GPU) -------------------------
 * The (cell, UMI) collapse is per cell, so after Part 1's filter the survivors are exchanged by cell
 * (all-to-all, done by the caller on the packed records of tec_sc_partition_dev), each carrying its
 * position in the job-wide survivor order.  tec_sc_finalize then runs the same pipeline per rank and calls the
 * collective at the few points that are global: the bundle boundaries of te_count.py:377 (a running
 * count over the whole file), the per-cell raw counts and first appearances behind the top-cell
 * choice (:502), the per-bundle cell presence behind the held-line rule (:528), the per-cell hit
 * counts (:653) and the statistics.  Triples stay on the rank that owns the cell.
 *   fn(user, dev_ptr, count, dtype, op): in-place all-reduce of a device buffer over the ranks;
 *   dtype 0 u32, 1 u64, 2 i64; op 0 sum, 1 min, 2 max; returns 0 on success. */
typedef int (*tec_allreduce_fn)(void* user, void* dev_ptr, int64_t count, int dtype, int op);
int tec_sc_set_collective(tec_ctx* ctx, tec_allreduce_fn fn, void* user, int rank, int world);
/* number of survivors held after the pushes */
int tec_sc_survivors(tec_ctx* ctx, int64_t* n);

/* the exchange: survivors packed as 32-byte records {u64 umi, u64 position,
 * u32 cell, u32 cs, i32 left, i32 right} grouped by owner rank cell % world (at most 8 ranks), file
 * order kept inside a group; counts[world] (host) are the group sizes, *records the device buffer.
 * After the all-to-all (received groups concatenated by source rank) hand the records back with
 * tec_sc_import_packed_dev. */
int tec_sc_partition_dev(tec_ctx* ctx, int world, int64_t gidx_base, int64_t* counts, void** records);
int tec_sc_import_packed_dev(tec_ctx* ctx, int64_t n, const void* records);

/* ---- collectives issued by the library (NCCL, bound at run time with dlopen; SURVEY.md 8b proposed
 * tec_reduce(ctx, ncclComm), 8e lists the exchange steps) ---------------------------------------
 * One communicator per context: rank 0 calls tec_comm_unique_id, the TEC_COMM_ID_BYTES bytes travel to
 * the other ranks by any means (te_counter_b200/dist.py: torch.distributed broadcast), every rank calls
 * tec_comm_init.  From then on the library runs its collectives itself, on its own stream and buffers:
 *   tec_bulk_allreduce        sum of the ranks' counter blocks, in place, behind the pushes; tec_bulk_finish
 *                             then returns the job's counts on every rank (te_count.py:128-149 summed)
 *   tec_sc_exchange           after the pushes: all-to-all of the survivors by owner rank cell % world
 *                             (tec_sc_partition_dev + grouped send / receive + tec_sc_import_packed_dev);
 *                             tec_sc_finalize then all-reduces through NCCL where tec_sc_set_collective
 *                             would have called back
 *   tec_sc_allgather_triples  after tec_sc_finalize: every rank holds the job's triples, ascending in
 *                             (ensg, cell); tec_sc_fetch returns them
 * Every rank must make the same sequence of these calls.  TEC_ERR_UNSUPPORTED: libnccl.so.2 not found. */
#define TEC_COMM_ID_BYTES 128
int tec_comm_unique_id(tec_ctx* ctx, void* id, int32_t capacity);
int tec_comm_init(tec_ctx* ctx, const void* id, int32_t rank, int32_t world);
int tec_comm_destroy(tec_ctx* ctx);
int tec_comm_info(tec_ctx* ctx, int32_t* rank, int32_t* world);
int tec_bulk_allreduce(tec_ctx* ctx);
int tec_sc_exchange(tec_ctx* ctx, int64_t* n_owned);
int tec_sc_allgather_triples(tec_ctx* ctx, int64_t* n_triples);

/* sc_save_result's choice of rows (te_count.py:724-733): hit cells by count descending, ties by
 * ascending id, at most maxcells.  cells_out must hold min(maxcells, n_hit_cells) entries. */
int tec_sc_select(tec_ctx* ctx, int64_t maxcells, uint32_t* cells_out, int64_t* n_out);

/*
 * Dense matrix rows of sc_save_result as text (te_count.py:744-754), built on the device: for each of
 * the n_rows cells, in the order given (normally tec_sc_select's), one line
 *     <barcode> '\t' <count of ensg 0> '\t' <count of ensg 1> ... '\n'
 * with every ensg of the index in id (= sorted name) order, zeros included, decimal integers.
 * cells[r] = whitelist id of row r (each at most once); barcodes = the rows' barcode strings
 * concatenated, row r is bytes [bc_off[r], bc_off[r+1]).  *n_bytes = length of the text, which stays
 * in device memory until the next tec_sc_matrix_text / tec_sc_begin; tec_sc_matrix_read copies the
 * byte range [offset, offset + n) of it into host memory (stream it to the file in chunks).
 * The header line (te_count.py:745) is the caller's: it is made of the feature names.
 */
int tec_sc_matrix_text(tec_ctx *ctx, int64_t n_rows, const uint32_t *cells, const char *barcodes,
                       const int64_t *bc_off, int64_t *n_bytes);
int tec_sc_matrix_read(tec_ctx *ctx, int64_t offset, int64_t n, char *out);

/* ---- BAM file -> counting path, decoded on the device (te_count.py:65-98, :190-214, :351-438)
 *
 * The same job as libtecbam (include/tecbam.h: BGZF inflate, record split, fields -> SoA columns, same
 * field semantics and the same error conditions) done by CUDA kernels with one thread per BGZF block,
 * with the columns handed to tec_bulk_push_dev / tec_sc_push_dev without leaving the device.  The host
 * reads the compressed file into pinned memory and walks the per-block results once per window (an
 * exact check of the record chain).  Usage: tec_bam_open, read the reference names, tec_bam_set_chrom_map
 * (and tec_bam_set_whitelist) exactly as for tbam_*, tec_bulk_begin / tec_sc_begin, tec_bam_count (decodes
 * the whole file into the running count), tec_bulk_finish / tec_sc_finalize, tec_bam_close.
 * mode: 0 single end, 1 paired end, 2 single cell.  On TEC_ERR_UNSUPPORTED start over with libtecbam.
 * Options (tec_set_option): "bam_window_blocks" (BGZF blocks per pass, default 65536: 4.3 GB of inflated bytes
 * in HBM), "bam_lanes" (blocks decoded per warp, default 1).  Limits shared with libtecbam: reference_end comes
 * from the CIGAR field of the record (for alignments of more than 65535 operations that field is the placeholder
 * <l_seq>S<ref_len>N of SAMv1 4.2.2, whose N operation carries the reference length: no expansion of the CG tag is
 * needed); tag values are compared as bytes. */
typedef struct tec_bam tec_bam;
int tec_bam_open(tec_ctx *ctx, const char *path, tec_bam **out);
void tec_bam_close(tec_bam *b);
int tec_bam_n_references(const tec_bam *b);
const char *tec_bam_reference_name(const tec_bam *b, int i);
int tec_bam_set_chrom_map(tec_bam *b, const uint16_t *bulk_ids, const uint16_t *sc_ids, int32_t n, int32_t n_index);
int tec_bam_set_whitelist(tec_bam *b, const char *barcodes, const int64_t *offsets, int32_t n);
int tec_bam_count(tec_bam *b, int mode, int qual, int64_t *n_records);
/* One BYTE RANGE of the file, for several ranks that split one BAM (single end and single cell; pairs are formed by
 * the global record parity and stay with one decoder).  Counts the records that START in the BGZF blocks whose first
 * byte lies in [byte_lo, byte_hi) into the running count.  out[TEC_BAM_RANGE_WORDS] = {records counted, start_block,
 * start_off, exit_block, exit_off, file size}: where this range's first record starts and where the chain stands
 * behind its last record, each as {file offset of the BGZF block, offset in its inflated data}.  start_block -1: the
 * range began at the BAM header (byte_lo == 0, exact); -2: no record starts in the range.  The first record of any other
 * range is the block-parallel guess, so the CALLER must check exit(rank) == start(next rank with records) for every
 * rank and exit(last) == {file size, 0}; if that fails, decode the file in one piece (te_count.py does both). */
#define TEC_BAM_RANGE_WORDS 6
int tec_bam_count_range(tec_bam *b, int mode, int qual, int64_t byte_lo, int64_t byte_hi, int64_t *out);
/* what = 0 blocks inflated by zlib on the host, 1..4 microseconds in load / inflate / chain / parse */
int64_t tec_bam_info(const tec_bam *b, int what);

#ifdef __cplusplus
}
#endif
#endif /* TECOUNT_H */

/* sstar_b200 -- C-ABI of the B200-native S* scoring engine.
 *
 * The reference (xin-huang/sstar, sstar2 2.0.0) is pure Python and has no FFI for this
 * path; the entry points below are what a ctypes binding behind its three Python seams
 * would bind (paths relative to the reference tree):
 *
 *   S3  Sstar.compute(**kwargs)                 sstar/sstar.py:43-106
 *   S2  mp_manager(job, data_generator, n)      sstar/mp_manager.py:98-188
 *   S1  preprocess(vcf_file, chr_name, ...)     sstar/preprocess.py:28-122
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every call returns 0 on success or a negative SSTAR_B200_E* code; the message is
 *     available from sstar_b200_last_error() (thread-local);
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t, NULL = default
 *     stream) of CUDA device `device`; the caller synchronises;
 *   - the callee never allocates or frees caller-visible memory: device scratch comes in
 *     through `d_workspace` (size it with sstar_b200_workspace_bytes);
 *   - no global mutable state except the thread-local error string, so the library may be
 *     driven concurrently from one host thread / process per GPU;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Data layout (exactly what the reference holds in numpy):
 *   ref_gt  int8  [V, R] row-major   reference genotypes (haplotype alleles or dosages)
 *   tgt_gt  int8  [V, T] row-major   target genotypes
 *   pos     int32 [V]                variant positions, STRICTLY ascending
 *   windows int32 [W] start / end    1-based CLOSED intervals (utils.py:480-488)
 *   score   int64 [W, T]             S* score; SSTAR_B200_NAN where the reference yields NaN
 *   nsnp    int32 [W, T]             Region_ind_SNP_number
 */
#ifndef SSTAR_B200_H
#define SSTAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSTAR_B200_ABI_VERSION 1
#define SSTAR_B200_NAN INT64_MIN

/* error codes */
#define SSTAR_B200_OK 0
#define SSTAR_B200_EINVAL (-1)    /* bad argument */
#define SSTAR_B200_ECUDA (-2)     /* CUDA runtime error (no device, launch failure, ...) */
#define SSTAR_B200_EWORKSPACE (-3) /* workspace too small */

/* algorithm selector (sstar_b200_options.algo) */
#define SSTAR_B200_ALGO_AUTO 0     /* prefix-max DP; windows with genotypes outside {0,1,2} go pairwise */
#define SSTAR_B200_ALGO_PAIRWISE 1 /* literal O(n^2) recurrence of sstar/sstar.py:208-212 for every window */

/* sstar_b200_options.flags */
#define SSTAR_B200_FLAG_STAGED_COPIES 1 /* host entry: always copy through device staging buffers,
                                           even when the host buffers are mapped (zero-copy capable) */
/* host entry, mapped buffers: tuning knobs of the pipelined path */
#define SSTAR_B200_FLAG_COPY_OUTPUTS 4      /* result tables always by D2H copies (default: only when >= 8 MB) */
#define SSTAR_B200_FLAG_SINGLE_CHUNK 8      /* no row chunking */
#define SSTAR_B200_FLAG_ZERO_COPY_INPUTS 32 /* the pack kernel reads the genotype matrices from mapped host memory */
#define SSTAR_B200_FLAG_COPY_INPUTS 64      /* ... never (default: inputs below 32 MB are read in place, larger ones copied in chunks) */
/* device entry points: which kernels score.  Default: small problems with a window-size hint (and
 * every replicate batch) take the one-launch kernel (bounds + staging + pack + DP in one CTA per few
 * windows), everything else pack -> score launches.  Both give identical results. */
#define SSTAR_B200_FLAG_NO_FUSED 128        /* never the one-launch kernel */
#define SSTAR_B200_FLAG_FORCE_FUSED 256     /* always (T <= 128), whatever the size and the hint */
/* A stream of independent calls (one chromosome after the other, sstar/preprocess.py:90-112): calls
 * that carry this flag promise that they touch no buffer a NEIGHBOURING flagged call on the same
 * stream reads or writes, so consecutive flagged one-launch kernels may overlap (programmatic
 * dependent launch, no wait: the next chromosome's CTAs load their tiles while this one's finish
 * their DP).  Everything else on the stream -- unflagged calls, copies, events -- keeps full stream
 * order after all flagged calls before it.  Ignored by the pack -> score launches (they share the
 * workspace).  No two flagged calls of an uninterrupted run may share a buffer one of them writes
 * (an unflagged call, a copy or a synchronisation ends the run).  On the host entry the flag also
 * pipelines the transfers of consecutive calls (see sstar_b200_score_windows_host). */
#define SSTAR_B200_FLAG_INDEPENDENT 512

typedef struct sstar_b200_options {
    int32_t algo;             /* SSTAR_B200_ALGO_* */
    int32_t max_win_variants; /* upper bound on variants inside any window (hi - lo): sizes the
                                 pairwise list arena and the window groups of the score kernel,
                                 and lets small inputs (and every replicate batch) run as ONE
                                 launch whose shared-memory tile it sizes; 0 = unknown (sized for
                                 V, always the pack -> score launches).  An understated hint is
                                 safe: it costs parallelism (windows larger than promised are
                                 streamed from global memory), never results */
    int32_t flags;            /* SSTAR_B200_FLAG_* */
    int32_t reserved[5];      /* must be zero */
} sstar_b200_options;

/* replaces: nothing in the reference (library identification) */
int sstar_b200_version(void);

/* replaces: the str(e) that mp_worker puts on the out queue, sstar/mp_manager.py:244-246 */
const char *sstar_b200_last_error(void);

/* Number of CUDA devices visible, or a negative error code. */
int sstar_b200_device_count(void);

/* Scratch bytes needed by sstar_b200_score_windows* for these shapes. */
size_t sstar_b200_workspace_bytes(int64_t V, int32_t R, int32_t T, int32_t W,
                                  int32_t max_win_variants);

/* Whole-chromosome scoring: every (window, target column) unit in one call.
 * replaces: WindowDataGenerator slicing      sstar/generators/window_data_generator.py:106-129
 *           mp_manager fan-out / gather      sstar/mp_manager.py:162-179
 *           FeatureVectorPreprocessor.run    sstar/feature_vector_preprocessor.py:71-137
 *           Sstar.compute / _calc_sstar      sstar/sstar.py:43-244
 * All pointers are DEVICE pointers. */
int sstar_b200_score_windows(int device, void *stream, const int8_t *d_ref_gt,
                             const int8_t *d_tgt_gt, const int32_t *d_pos, int64_t V, int32_t R,
                             int32_t T, const int32_t *d_win_start, const int32_t *d_win_end,
                             int32_t W, int32_t match_bonus, int32_t max_mismatch,
                             int32_t mismatch_penalty, int64_t *d_score, int32_t *d_nsnp,
                             void *d_workspace, size_t workspace_bytes);

/* Same, with options (algorithm choice, size hints). `opts` may be NULL. */
int sstar_b200_score_windows_ex(int device, void *stream, const int8_t *d_ref_gt,
                                const int8_t *d_tgt_gt, const int32_t *d_pos, int64_t V, int32_t R,
                                int32_t T, const int32_t *d_win_start, const int32_t *d_win_end,
                                int32_t W, int32_t match_bonus, int32_t max_mismatch,
                                int32_t mismatch_penalty, int64_t *d_score, int32_t *d_nsnp,
                                void *d_workspace, size_t workspace_bytes,
                                const sstar_b200_options *opts);

/* Replicate-batched scoring for the simulate/train path: Nrep independent replicates are
 * concatenated along the variant axis (replicate r owns rows d_rep_offset[r] .. d_rep_offset[r+1]),
 * positions ascend inside each replicate, and every replicate is scored on ONE window
 * [win_start, win_end] (the reference uses [1, seq_len]).  Outputs are [Nrep, T].
 * replaces: MsprimeSimulator._build_feature_input_from_ts + FeatureVectorPreprocessor.run per
 *           replicate, sstar/msprime_simulator.py:179-256, fanned out by sstar/simulate.py:147-151 */
int sstar_b200_score_replicates(int device, void *stream, const int8_t *d_ref_gt,
                                const int8_t *d_tgt_gt, const int32_t *d_pos, int64_t V, int32_t R,
                                int32_t T, const int32_t *d_rep_offset, int32_t Nrep,
                                int32_t win_start, int32_t win_end, int32_t match_bonus,
                                int32_t max_mismatch, int32_t mismatch_penalty, int64_t *d_score,
                                int32_t *d_nsnp, void *d_workspace, size_t workspace_bytes,
                                const sstar_b200_options *opts);

/* The two halves of the pipeline, exposed so the packing kernel can be timed alone.
 * pack:  genotype matrices -> bit-planes + absent mask inside the workspace   (sstar.py:87-92,163-164)
 * score: window bounds + chain DP on an already packed workspace              (sstar.py:183-217) */
int sstar_b200_pack(int device, void *stream, const int8_t *d_ref_gt, const int8_t *d_tgt_gt,
                    int64_t V, int32_t R, int32_t T, void *d_workspace, size_t workspace_bytes);

int sstar_b200_score_packed(int device, void *stream, const int8_t *d_tgt_gt, const int32_t *d_pos,
                            int64_t V, int32_t R, int32_t T, const int32_t *d_win_start,
                            const int32_t *d_win_end, int32_t W, int32_t match_bonus,
                            int32_t max_mismatch, int32_t mismatch_penalty, int64_t *d_score,
                            int32_t *d_nsnp, void *d_workspace, size_t workspace_bytes,
                            const sstar_b200_options *opts);

/* Host-buffer entry (the end-to-end call).  All work is enqueued on `stream` (plus an internal
 * copy stream that `stream` joins before the call's last kernel); the caller synchronises
 * `stream` before reading h_score / h_nsnp.
 *   - If every host buffer is page-locked AND mapped (cudaHostAlloc / cudaHostRegister / torch
 *     pin_memory) the call is pipelined.  Genotype matrices of 32 MB or more move host->device in
 *     row chunks on the copy stream while earlier chunks are packed and their windows scored;
 *     smaller ones are read in place by the pack kernel (its TMA tile loads are the transfer).
 *     Result tables below 8 MB are stored by the score kernels straight into h_score / h_nsnp,
 *     larger ones copied back per chunk.
 *   - SSTAR_B200_FLAG_INDEPENDENT (mapped buffers, inputs below 32 MB): one of a STREAM of calls, each
 *     with its own host buffers and its own `d_scratch` (the caller cycles through a few).  Inputs
 *     go through the copy engine on the internal copy stream, queued right behind the previous
 *     call's copies, so the link stays busy while `stream` scores that call; tables are stored in
 *     place.  The library orders a call's copies after the previous call that used the SAME scratch
 *     (an event recorded behind its kernels) -- the caller synchronises `stream` once, at the end.
 *   - Otherwise (pageable memory, or SSTAR_B200_FLAG_STAGED_COPIES) inputs are copied to device
 *     staging buffers, scored, and the tables copied back, all on `stream`.
 * `d_scratch` is device memory of at least sstar_b200_host_scratch_bytes(...) either way. */
size_t sstar_b200_host_scratch_bytes(int64_t V, int32_t R, int32_t T, int32_t W,
                                     int32_t max_win_variants);

int sstar_b200_score_windows_host(int device, void *stream, const int8_t *h_ref_gt,
                                  const int8_t *h_tgt_gt, const int32_t *h_pos, int64_t V, int32_t R,
                                  int32_t T, const int32_t *h_win_start, const int32_t *h_win_end,
                                  int32_t W, int32_t match_bonus, int32_t max_mismatch,
                                  int32_t mismatch_penalty, int64_t *h_score, int32_t *h_nsnp,
                                  void *d_scratch, size_t scratch_bytes,
                                  const sstar_b200_options *opts);

/* Result emitter: the feature table as TSV text (no header line), formatted on the device.
 * replaces: the row dicts of FeatureVectorPreprocessor._fmt_res (sstar/feature_vector_preprocessor.py:167-181)
 *           printed by pd.DataFrame(rows).to_csv(sep="\t", index=False, na_rep="NA")
 *           (sstar/preprocess.py:122, sstar/simulate.py:195) -- same bytes.
 * Row: chrom \t start \t end \t label \t score \t nsnp [\t replicate] \n ; score prints as "51470.0",
 * SSTAR_B200_NAN as "NA".  Rows are in (window, k) order and show column d_col_order[k] (NULL =
 * identity; the simulate path sorts rows by sample NAME).  d_labels holds the T labels
 * concatenated, label t = d_labels[d_label_off[t] .. d_label_off[t+1]).  d_replicate (NULL = no
 * such column) is per window.  All d_* pointers are device memory.
 * On return (after the stream syncs) the first 16 bytes of d_workspace hold
 *   int64 total_bytes; uint32 status  (bit 0: d_out too small, bit 1: |score| >= 1e16, where
 *   the reference switches to exponent notation -- nothing is written, format on the host). */
size_t sstar_b200_tsv_max_bytes(int32_t W, int32_t T, int32_t chrom_len, int32_t max_label_len,
                                int32_t has_replicate);
size_t sstar_b200_tsv_workspace_bytes(int32_t W, int32_t T);
int sstar_b200_format_tsv(int device, void *stream, const int64_t *d_score, const int32_t *d_nsnp,
                          const int32_t *d_win_start, const int32_t *d_win_end, int32_t W, int32_t T,
                          const char *chrom, const char *d_labels, const int32_t *d_label_off,
                          int32_t max_label_len, const int32_t *d_replicate,
                          const int32_t *d_col_order, char *d_out, size_t out_capacity,
                          void *d_workspace, size_t workspace_bytes);

/* The same emitter with options: row order and the prediction step of `sstar2 infer`.
 * replaces: sstar/quantile_regression.py:108-135 -- the Predicted_S*_score column (NA where the
 *           score is NA), rows sorted by (Sample, Chromosome, Start, End), and the BED file of the
 *           units whose observed score exceeds the predicted one (Start - 1, no header).
 * The model has ONE integer feature (Region_ind_SNP_number), so the caller evaluates it once per
 * distinct value on the host and passes the table: entry v holds the prediction for feature
 * value v (values beyond n_pred - 1 use the last entry).
 * max_tsv_bytes for a table with predictions: sstar_b200_tsv_max_bytes(...) + W*T*(1 + max_pred_len). */
typedef struct sstar_b200_table_options {
    int32_t mode;               /* 0: feature rows (+ prediction column when d_pred_text), 1: BED rows */
    int32_t sample_major;       /* 0: rows in (window, sample) order, 1: (sample, window) order */
    const double *d_pred_value; /* [n_pred] predictions (device), or NULL */
    const char *d_pred_text;    /* their decimal text, concatenated (device); mode 0 only */
    const int32_t *d_pred_off;  /* [n_pred + 1] (device) */
    int32_t n_pred;
    int32_t max_pred_len;
} sstar_b200_table_options;
int sstar_b200_format_table(int device, void *stream, const int64_t *d_score, const int32_t *d_nsnp,
                            const int32_t *d_win_start, const int32_t *d_win_end, int32_t W, int32_t T,
                            const char *chrom, const char *d_labels, const int32_t *d_label_off,
                            int32_t max_label_len, const int32_t *d_replicate,
                            const int32_t *d_col_order, const sstar_b200_table_options *topts,
                            char *d_out, size_t out_capacity, void *d_workspace, size_t workspace_bytes);

/* Native VCF scanner (host code, multi-threaded; plain text, gzip, or BGZF / bgzip -- whose members are
 * inflated in parallel and checked against their CRC32 and length; one plain gzip stream is inflated on one thread).
 * replaces: allel.read_vcf(vcf, alt_number=1, samples=..., region=chr_name) as used by
 *           read_geno_data (sstar/utils/utils.py:80-109) -- scikit-allel is an un-vendored dependency.
 * One streaming pass over the file (a block of lines at a time; lines of other chromosomes are
 * skipped before anything is stored) gives, per chromosome in order of first appearance: POS int32
 * [V], the REF and first-ALT allele strings, and GT int8 [V, N, 2] (-1 = missing allele) for the N
 * kept samples in header order.  chr_name NULL keeps every chromosome.
 * sstar_b200_vcf_scan keeps ALL samples of the header; sstar_b200_vcf_scan_ex keeps only the header
 * samples named in `samples` (n_samples names; NULL = all) -- what read_vcf(samples=...) does --
 * and sstar_b200_vcf_sample / _n_samples then describe the kept ones.  Pointers returned by the
 * accessors stay valid until sstar_b200_vcf_free.  Errors: negative code, text from
 * sstar_b200_vcf_last_error(). */
typedef struct sstar_b200_vcf sstar_b200_vcf;
int sstar_b200_vcf_scan(const char *path, const char *chr_name, int nthreads, sstar_b200_vcf **out);
int sstar_b200_vcf_scan_ex(const char *path, const char *chr_name, const char *const *samples,
                           int32_t n_samples, int nthreads, sstar_b200_vcf **out);
void sstar_b200_vcf_free(sstar_b200_vcf *h);
const char *sstar_b200_vcf_last_error(void);
int32_t sstar_b200_vcf_n_samples(const sstar_b200_vcf *h);
const char *sstar_b200_vcf_sample(const sstar_b200_vcf *h, int32_t i);
int32_t sstar_b200_vcf_n_chroms(const sstar_b200_vcf *h);
const char *sstar_b200_vcf_chrom(const sstar_b200_vcf *h, int32_t c);
int64_t sstar_b200_vcf_n_variants(const sstar_b200_vcf *h, int32_t c);
const int32_t *sstar_b200_vcf_pos(const sstar_b200_vcf *h, int32_t c);
const int8_t *sstar_b200_vcf_gt(const sstar_b200_vcf *h, int32_t c);
/* which: 0 = REF, 1 = first ALT; allele v = bytes[off[v] .. off[v+1]) */
const char *sstar_b200_vcf_alleles(const sstar_b200_vcf *h, int32_t c, int32_t which, const int32_t **off);

/* Ingest filters on the device: raw calls of one chromosome -> the matrices the tiler gets.
 * replaces: the array half of read_geno_data / read_data (sstar/utils/utils.py:105-107 all-missing
 *           rows per population, :167-222 shared positions, :282-291 sites fixed-alt in both
 *           populations, :293-306 phased reshape / unphased allele sum, :363-423 ancestral-allele
 *           polarisation with the keep/flip/drop decision per variant made on the host).
 * d_gt int8 [V, N, 2] (-1 = missing allele, all N samples in header order), d_pos int32 [V],
 * d_mode int8 [V] or NULL (0 keep, 1 flip g -> |g-1|, 2 drop), d_ref_cols / d_tgt_cols: sample
 * indices of the two populations.  Outputs (capacity V rows): d_ref_out int8 [V', R],
 * d_tgt_out int8 [V', T], d_pos_out int32 [V'] with R, T = n * 2 (phased) or n (unphased).
 * After the stream syncs the first 16 bytes of d_workspace hold int64 V' and int64 the number of
 * rows both populations shared before the fixed-site filter.
 * sstar_b200_ingest_ex additionally takes n_ref_listed / n_tgt_listed = len(ref_samples) /
 * len(tgt_samples) of the sample lists: the reference's fixed-site filter compares the hom-alt
 * counts with THOSE (utils.py:282-291), so a list naming a sample the VCF does not hold never drops
 * a fixed site; sstar_b200_ingest passes the column counts. */
size_t sstar_b200_ingest_workspace_bytes(int64_t V);
int sstar_b200_ingest(int device, void *stream, const int8_t *d_gt, const int32_t *d_pos,
                      const int8_t *d_mode, int64_t V, int32_t N, const int32_t *d_ref_cols,
                      int32_t n_ref, const int32_t *d_tgt_cols, int32_t n_tgt, int32_t is_phased,
                      int8_t *d_ref_out, int8_t *d_tgt_out, int32_t *d_pos_out, void *d_workspace,
                      size_t workspace_bytes);
int sstar_b200_ingest_ex(int device, void *stream, const int8_t *d_gt, const int32_t *d_pos,
                         const int8_t *d_mode, int64_t V, int32_t N, const int32_t *d_ref_cols,
                         int32_t n_ref, int32_t n_ref_listed, const int32_t *d_tgt_cols, int32_t n_tgt,
                         int32_t n_tgt_listed, int32_t is_phased, int8_t *d_ref_out, int8_t *d_tgt_out,
                         int32_t *d_pos_out, void *d_workspace, size_t workspace_bytes);

/* S* SNP sets: the sites on the best chain of every (window, column) unit -- an extension of the
 * scoring path.  The current reference stops at the score (sstar/sstar.py:205-217 never backtracks);
 * its legacy fixtures tests/exp_results/test.score.{unphased,phased}.exp.results.tsv carry the
 * columns S*_SNP_number / S*_SNPs, reproduced by the rule of SURVEY.md section 8a item 6: the
 * backpointer of site j is the FIRST i < j attaining M[j]; there the chain is extended when
 * M[i] >= 0 and starts afresh with the pair (i, j) otherwise; it ends at the FIRST j attaining
 * max_j M[j].  The result is ragged, so it takes two calls around one host read:
 *   1. sstar_b200_snp_set_offsets: packs, resolves the windows, runs the literal pairwise recurrence
 *      with backpointers; writes d_score / d_nsnp (as sstar_b200_score_windows) and
 *      d_set_off [W*T + 1] (int64): unit u = w*T + t owns entries [d_set_off[u], d_set_off[u+1]).
 *   2. the caller reads d_set_off[W*T] (total entries), provides d_set_pos [total] (int32) and calls
 *      sstar_b200_snp_set_fill with the SAME workspace and parameters: positions of the chain
 *      sites, ascending, unit after unit.  Units whose entries would pass `capacity` are skipped.
 * Workspace: sstar_b200_snp_set_workspace_bytes (a superset of sstar_b200_workspace_bytes). */
size_t sstar_b200_snp_set_workspace_bytes(int64_t V, int32_t R, int32_t T, int32_t W, int32_t max_win_variants);
int sstar_b200_snp_set_offsets(int device, void *stream, const int8_t *d_ref_gt, const int8_t *d_tgt_gt,
                               const int32_t *d_pos, int64_t V, int32_t R, int32_t T,
                               const int32_t *d_win_start, const int32_t *d_win_end, int32_t W,
                               int32_t match_bonus, int32_t max_mismatch, int32_t mismatch_penalty,
                               int64_t *d_score, int32_t *d_nsnp, int64_t *d_set_off, void *d_workspace,
                               size_t workspace_bytes, const sstar_b200_options *opts);
int sstar_b200_snp_set_fill(int device, void *stream, const int8_t *d_tgt_gt, const int32_t *d_pos, int64_t V,
                            int32_t R, int32_t T, int32_t W, int32_t match_bonus, int32_t max_mismatch,
                            int32_t mismatch_penalty, const int64_t *d_set_off, int32_t *d_set_pos,
                            int64_t capacity, void *d_workspace, size_t workspace_bytes,
                            const sstar_b200_options *opts);

/* Source match rates of inferred tracts: integer numerators per (tract, source individual).
 * replaces: the tract x source loops of run_match (sstar/match.py:497-545) around
 *           _dosage_for_position_index (:88-162) and calc_match_pct (:30-73).
 * d_gt int8 [V, N, 2] (-1 = missing allele) of one chromosome; tract k covers variant rows
 * [d_tract_lo[k], d_tract_hi[k]) (start < POS <= end), compares target column d_tract_tgt[k]
 * (diploid ALT dosage when d_tract_hap[k] < 0, else that haplotype's allele * ploidy) with each
 * source column's diploid dosage.  d_numer[k, s] = sum over sites called in both of
 * (ploidy - |x - y|); the rate is (numer / ploidy) / (hi - lo), averaged over the sources. */
int sstar_b200_match_numerators(int device, void *stream, const int8_t *d_gt, int64_t V, int32_t N,
                                const int32_t *d_tract_lo, const int32_t *d_tract_hi,
                                const int32_t *d_tract_tgt, const int32_t *d_tract_hap, int32_t K,
                                const int32_t *d_src_cols, int32_t S, int32_t ploidy, int64_t *d_numer);

/* Number of kernels the last sstar_b200_score_* / pack call on this thread enqueued. */
int sstar_b200_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SSTAR_B200_H */

/*
 * dcpgpu.h -- thin C ABI of the B200 (sm_100a) scan hot path of deciphon_b200.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * It replaces, for a whole batch of (window, profile) pairs at once, the calls the
 * reference's per-window loop makes one pair at a time:
 *
 *   reference (c-core)                                   this ABI
 *   -------------------------------------------------    ---------------------------------
 *   work_setup -> protein_setup_viterbi                  dcpgpu_pool_add + dcpgpu_profile_add
 *     (work.c:24-46, protein.c:353-394,
 *      viterbi_setup/viterbi_set_*  viterbi.c:336-444)
 *   batch_encode -> imm_eseq (batch.c:60, sequence.c:47) dcpgpu_reads_set (2-bit packing)
 *   work_reset -> xtrans_setup(_viterbi)                 computed inside score/trace from the
 *     (work.c:47-51, xtrans.c:21-68, thread.c:112)         window length (same C expressions)
 *   viterbi_null + viterbi_cost (viterbi.c:696-724,      dcpgpu_score_pairs / dcpgpu_score_grid
 *     thread.c:114-116)
 *   viterbi_path + trellis_unzip (viterbi.c:726-732,     dcpgpu_trace_pairs + dcpgpu_trace_fetch
 *     trellis.c:147-167, thread.c:126-128)
 *
 * All scores are COSTS (= -log-likelihood, fp32 min-plus), exactly what
 * viterbi_null()/viterbi_cost() return; lrt = -2*((-null) - (-alt)) (lrt.h:6-9).
 *
 * Every function returns 0 or a positive DCPGPU_E* code (the host library maps
 * them onto DCP_E* values appended after 80, see deciphon_b200.h).
 * There is NO CPU fallback: without a CUDA device dcpgpu_open() fails.
 * A context is single-threaded (like a dcp_scan object); use one per GPU.
 */
#ifndef DCPGPU_H
#define DCPGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCPGPU_NUM_CODES 1364 /* c-core/viterbi.c:13 TABLE_SIZE */
#define DCPGPU_TRANS_SIZE 7   /* c-core/trans.h: MM,MI,MD,IM,II,DM,DD */
#define DCPGPU_MAX_CORE_SIZE 16384 /* c-core/model.h:12 MODEL_MAX */
#define DCPGPU_MAX_WINDOW 100000   /* c-core/window.c:29 */

enum
{
  DCPGPU_OK = 0,
  DCPGPU_ENODEVICE = 1, /* no CUDA device / driver */
  DCPGPU_ECUDA = 2,     /* a CUDA call failed: see dcpgpu_last_error() */
  DCPGPU_ENOMEM = 3,    /* host or device allocation failed */
  DCPGPU_EINVAL = 4,    /* bad argument (range, NULL, ordering) */
  DCPGPU_ESTATE = 5,    /* call made in the wrong state (e.g. fetch before trace) */
  DCPGPU_EDECODE = 6,   /* a path fragment no codon can have produced (decoder.c:52-56), or no decode tables */
};

/* flags for score/trace: the two booleans of dcp_scan_setup (deciphon.h:11-13) */
#define DCPGPU_MULTI_HITS 1u
#define DCPGPU_HMMER3_COMPAT 2u
/* trace_pairs only: materialise the whole bit-packed trellis (for dcpgpu_trace_trellis) instead
 * of deciding just the words the path visits.  Paths are identical either way. */
#define DCPGPU_KEEP_TRELLIS 4u

typedef struct dcpgpu_ctx dcpgpu_ctx;

/* One unit of work: window [start, start+len) of sequence `seq` against `profile`. */
typedef struct dcpgpu_pair
{
  int32_t profile;
  int32_t seq;
  int32_t start;
  int32_t len;
} dcpgpu_pair;

/* ---- context ------------------------------------------------------------------------- */
int dcpgpu_open(dcpgpu_ctx **ctx, int device);
void dcpgpu_close(dcpgpu_ctx *ctx);
char const *dcpgpu_strerror(int code);
char const *dcpgpu_last_error(dcpgpu_ctx const *ctx);
/* Run on a caller-owned CUDA stream (a cudaStream_t passed as void*); NULL = own stream. */
int dcpgpu_set_stream(dcpgpu_ctx *ctx, void *cuda_stream);
int dcpgpu_sync(dcpgpu_ctx *ctx);
/* device facts: 0 = SM count, 1 = total bytes, 2 = free bytes, 3 = bytes held by profiles */
int64_t dcpgpu_device_info(dcpgpu_ctx const *ctx, int what);
/* CUDA devices visible to this process (0 when there is no driver or no device). */
int32_t dcpgpu_device_count(void);

/* ---- profiles ("work_setup": a profile becomes resident) -------------------------------
 * Nodes are uploaded in .dcp form (natural-log probabilities, protein.c:234-281):
 * emission[n][1364] and trans[n][7] = transitions OUT of the node.  The device negates
 * them into costs and re-indexes by destination node exactly like protein.c:353-394. */
int dcpgpu_pool_add(dcpgpu_ctx *ctx, int nnodes, float const *emission, float const *trans,
                    int64_t *first_node_id);
/* Assemble a profile of K nodes from pool nodes.  node_ids[K] (NULL = first_node_id + k).
 * BMk[K], null_emission[1364], bg_emission[1364] are log-probs.  Returns its index. */
int dcpgpu_profile_add(dcpgpu_ctx *ctx, int K, int64_t const *node_ids, int64_t first_node_id,
                       float const *BMk, float const *null_emission, float const *bg_emission,
                       int32_t *profile_index);
int dcpgpu_profile_count(dcpgpu_ctx const *ctx);
int dcpgpu_profile_core_size(dcpgpu_ctx const *ctx, int32_t profile);
/* Drop the node pool once all profiles are assembled (frees its device memory). */
int dcpgpu_pool_release(dcpgpu_ctx *ctx);

/* ---- reads ("batch_encode") -------------------------------------------------------------
 * symbols: 0..3 = A,C,G,T/U, already upper-cased and disambiguated (disambiguate.c);
 * sequence s = symbols[offsets[s] .. offsets[s+1]).  Packed 2 bits/nt on the device. */
int dcpgpu_reads_set(dcpgpu_ctx *ctx, int32_t nseq, uint8_t const *symbols,
                     int64_t const *offsets);
int dcpgpu_reads_count(dcpgpu_ctx const *ctx);

/* ---- score pass: viterbi_null + viterbi_cost for every pair ----------------------------
 * Explicit list (host memory in, host memory out; any of the outputs may be NULL). */
int dcpgpu_score_pairs(dcpgpu_ctx *ctx, int64_t npairs, dcpgpu_pair const *pairs,
                       uint32_t flags, float *null_cost, float *alt_cost);
/* First window (window.c:13-37 from its initial state: [0, min(50K, 100000, |seq|)) ) of
 * every sequence in [seq0, seq1) against every profile in [prof0, prof1).  Pairs are
 * generated on the device, profile-major; results stay on the device until fetched.
 * Asynchronous on the context's stream unless profiles of more than 256 nodes are present
 * (their speculative-strip kernels are followed by a host-side check of the redo queue). */
int dcpgpu_score_grid(dcpgpu_ctx *ctx, int32_t prof0, int32_t prof1, int32_t seq0, int32_t seq1,
                      uint32_t flags);
/* Copy the last score pass' results to the host ([npairs] each, NULL = skip); synchronises. */
int dcpgpu_scores_fetch(dcpgpu_ctx *ctx, int64_t npairs, float *null_cost, float *alt_cost);
/* Indices (into the last score pass) of pairs with finite lrt >= 0 (thread.c:119-121),
 * ascending; returns their count through *nhits, writes at most cap of them. */
int dcpgpu_hits_fetch(dcpgpu_ctx *ctx, int64_t cap, int64_t *hit_index, int64_t *nhits);
/* The last score pass' results of the given pairs only (e.g. the hit_index list of
 * dcpgpu_hits_fetch): null_cost[i], alt_cost[i] = costs of pair index[i]; synchronises. */
int dcpgpu_scores_gather(dcpgpu_ctx *ctx, int64_t n, int64_t const *index, float *null_cost, float *alt_cost);
/* DP cells (sum of len*K) of the last score pass, and device ms of its kernels. */
double dcpgpu_last_cells(dcpgpu_ctx const *ctx);
float dcpgpu_last_kernel_ms(dcpgpu_ctx *ctx);
/* Pairs of the last score pass whose speculative-strip run (profiles > 256 nodes) had to be
 * redone by the exact multi-warp kernel (results are identical either way). */
int64_t dcpgpu_last_redo(dcpgpu_ctx const *ctx);
/* Cumulative number of kernels this library has launched on the context. */
int64_t dcpgpu_launch_count(dcpgpu_ctx const *ctx);
/* Cumulative counters of the context: 0 = bytes copied host -> device, 1 = bytes copied device ->
 * host, 2 = kernels launched, 3 = DP cells of all score passes. */
double dcpgpu_counter(dcpgpu_ctx const *ctx, int what);

/* ---- trace pass: viterbi_path + trellis_unzip for the given (hit) pairs -----------------
 * Decides the reference's bit-packed trellis words (trellis.h:12-56) on the device -- only the
 * words the path visits unless DCPGPU_KEEP_TRELLIS asks for all of them --, walks back
 * T@L -> S@0 on the device and reports the number of steps of every path. */
int dcpgpu_trace_pairs(dcpgpu_ctx *ctx, int64_t npairs, dcpgpu_pair const *pairs, uint32_t flags,
                       float *alt_cost, int32_t *nsteps);
/* Steps of all traced paths, path i at [offsets[i], offsets[i] + nsteps[i]) in path order
 * (S first): state ids as in state.h:7-25 and emitted nucleotides per step. */
int dcpgpu_trace_fetch(dcpgpu_ctx *ctx, int64_t const *offsets, uint16_t *state_ids,
                       uint8_t *seqsizes);
/* Raw trellis words of traced pair i: xnodes[len+1], nodes[(len+1)*K] (tests, debugging).
 * Needs DCPGPU_KEEP_TRELLIS in the flags of the preceding dcpgpu_trace_pairs. */
int dcpgpu_trace_trellis(dcpgpu_ctx *ctx, int64_t i, uint32_t *xnodes, uint16_t *nodes);

/* Measured non-tensor FP32 issue rate of this GPU in tera lane-operations per second, a three-input
 * min and a packed f32x2 add counted as two operations each.  mode 0: FADD+FMNMX 1:1, 1: FADD,
 * 2: FMNMX, 3: FMNMX3, 4: FADD2, 5: FADD+FMNMX3 2:1, 6: FADD + three-input integer min (VIMNMX3)
 * 2:1 -- the instruction mix of the score row, the highest add/min rate measured on B200
 * (tools/alu_probe.cu, profiles/README.md).  The hard issue ceiling, one lane-operation per lane
 * and clock, is SMs x 128 x clock. */
int dcpgpu_alu_peak(dcpgpu_ctx *ctx, int mode, double *tera_ops_per_s);

/* ---- match strings: the post-processing of a traced path, on the device --------------------
 * Decode tables of a profile: node_dists[K][129], null_dist[129], bg_dist[129] (4 base log-probs
 * then 125 codon marginal log-probs [a][b][c], index 4 = any base -- the nuclt_dist of a .dcp
 * record, nuclt_dist.c:13-20) and the genetic code as 64 amino letters in TCAG x TCAG x TCAG order. */
int dcpgpu_profile_set_decoder(dcpgpu_ctx *ctx, int32_t profile, float const *node_dists,
                               float const *null_dist, float const *bg_dist, char const *gencode64);
/* For every pair of the preceding dcpgpu_trace_pairs: whether the path has a B..E segment, its
 * window-relative extent (thread.c:130-166) and the bytes of the row's match column
 * ("<fragment>,<state>,<codon>,<amino>" joined by ';': match.c:66-90, product_thread.c:112-148,
 * codons by the frame-state decoder of decoder.c:38-58).  text_off[npairs + 1] receives where
 * each pair's bytes start in the buffer dcpgpu_match_fetch fills.  Any output may be NULL; with
 * text_off == NULL only the extents are computed (no decode tables needed, nothing to fetch). */
int dcpgpu_match_build(dcpgpu_ctx *ctx, float epsilon, int is_rna, int32_t *hit, int32_t *hit_start,
                       int32_t *hit_stop, int64_t *text_off);
int dcpgpu_match_fetch(dcpgpu_ctx *ctx, char *text);

/* ---- press: frame-state emission tables --------------------------------------------------
 * For each of nstates states, the log-probability of emitting every 1..5-nucleotide fragment
 * (out[s][1364], the scan's code order) from its base log-probs [4], codon marginal log-probs
 * [5][5][5] (index 4 = any base) and the indel rate epsilon: what imm_score_table_scores writes
 * for an imm_frame_state at c-core/protein.c:102, protein_null.c:24, protein_background.c:19.
 * Host pointers in and out. */
int dcpgpu_frame_tables(dcpgpu_ctx *ctx, int32_t nstates, float const *nuclt_lprobs,
                        float const *codon_marg_lprobs, float epsilon, float *emission);

/* The 13 special-transition costs for a window of window_len nucleotides, in the order
 * RR,SN,NN,SB,NB,EB,JB,EJ,JJ,EC,CC,ET,CT (viterbi.h:4-19), as the device uses them. */
int dcpgpu_xtrans(int window_len, uint32_t flags, float out[13]);

#ifdef __cplusplus
}
#endif
#endif

/*
 * deciphon_b200.h -- the reference's public scan API, served by the B200 scan path.
 *
 * Same names, argument meaning and error behaviour as c-core/deciphon.h:9-32 (mirrored in
 * python-core/deciphon_core/interface.h:1-40), so that python-core's cffi layer can dlopen
 * libdeciphon_b200.so instead of libdeciphon.  Behind these entry points the per-window
 * loop of c-core/thread.c:49-208 runs as waves of batched GPU passes through the C ABI of
 * include/dcpgpu.h.  What differs from the reference, by design:
 *
 *   * dcp_scan_setup: the reference cuts the database into `num_threads` partitions, one per
 *     OpenMP thread (c-core/scan.c:95-152); here the partitions live on GPUs: min(num_threads,
 *     visible CUDA devices, $DCP_GPU_COUNT) devices starting at $DCP_GPU_DEVICE (default 0), each
 *     holding $DCP_SHARDS_PER_GPU (default 2; 1 for databases of fewer than 128 profiles per shard)
 *     shards of contiguous profiles balanced by core size, one host thread per shard -- one
 *     shard's kernels fill the GPU while the other's host thread decodes hits and formats rows.  `cache` is accepted and
 *     ignored (every profile is resident in HBM).  Rows come out in the reference's order
 *     whatever the number of shards.  The callback fires on shard 0 after every chunk of profiles
 *     (the reference: on partition 0 after every window); dcp_scan_interrupt is honoured between
 *     chunks ($DCP_CHUNK_CELLS DP cells each, default 4e11, about 0.8 s of GPU time: every chunk boundary drains the GPU, about 30 ms).
 *   * HMMER daemon (c-core/hmmer.c, thread.c:185-203): the third-party client libraries are
 *     not part of this build.  `port <= 0` runs WITHOUT the HMMER confirmation stage: every
 *     window with lrt >= 0 and a B..E segment yields a row, `evalue` is written as 0 and no
 *     .h3r files are produced.  `port > 0` returns DCP_EH3CDIAL.  With $DCP_WRITE_AMINOS=1 the
 *     amino-acid sequence of every row (what the reference sends to the daemon hit by hit,
 *     thread.c:168-190) is written to <product_dir>/hmmer/aminos.fa in row order: the input of
 *     ONE batched hmmscan --cut_ga run over all hits instead of a round trip per hit.
 *   * dcp_press_* (c-core/press.c): HMMER3 ASCII -> .dcp in the reference's current encoding
 *     (minifam.hmm -> 3,609,858 bytes, test_press.c:26); the frame-state emission tables are
 *     computed on the GPU.  entry_dist is occupancy-based (press.c:60).
 *   * new error codes are appended after DCP_EINVALNUMPROTEINS = 80.
 */
#ifndef DECIPHON_B200_H
#define DECIPHON_B200_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

struct dcp_scan;
struct dcp_batch;
struct dcp_press;

struct dcp_scan *dcp_scan_new(void);
void dcp_scan_del(struct dcp_scan const *);
int dcp_scan_setup(struct dcp_scan *, char const *dbfile, int port, int num_threads, bool multi_hits,
                   bool hmmer3_compat, bool cache, void (*callback)(void *), void *userdata);
int dcp_scan_run(struct dcp_scan *, struct dcp_batch *, char const *product_dir);
void dcp_scan_interrupt(struct dcp_scan *);
int dcp_scan_progress(struct dcp_scan const *);

struct dcp_press *dcp_press_new(void);
int dcp_press_setup(struct dcp_press *, int gencode_id, float epsilon);
int dcp_press_open(struct dcp_press *, char const *hmm, char const *db);
long dcp_press_nproteins(struct dcp_press const *);
int dcp_press_next(struct dcp_press *);
bool dcp_press_end(struct dcp_press const *);
int dcp_press_close(struct dcp_press *);
void dcp_press_del(struct dcp_press const *);

struct dcp_batch *dcp_batch_new(void);
void dcp_batch_del(struct dcp_batch *);
int dcp_batch_add(struct dcp_batch *, long id, char const *name, char const *data);
void dcp_batch_reset(struct dcp_batch *);

char const *dcp_error_string(int error_code);

/* Extension: parse a .dcp database (either float encoding) without touching
 * the GPU; reports the number of profiles, the total core size and epsilon. */
int dcpb200_db_info(char const *dbfile, int *num_proteins, long *total_core_size, float *epsilon);
/* Extension: GPUs a set-up scan runs on, and the profile shards (host threads) it is cut into. */
int dcpb200_scan_num_gpus(struct dcp_scan const *);
int dcpb200_scan_num_shards(struct dcp_scan const *);
/* Extension: cumulative counters summed over the shards' GPU contexts (dcpgpu_counter): 0 = bytes
 * copied host -> device, 1 = bytes copied device -> host, 2 = kernels launched; and of the scan
 * itself: 3 = DP cells of the windows it scanned, 4 = those windows, 5 = the ones with lrt >= 0
 * (thread.c:119-121), 6 = windows scored ahead of a hit that changed the chain and re-planned
 * (extra work, not counted in 3 and 4). */
double dcpb200_scan_counter(struct dcp_scan const *, int what);

/* Error codes 1..80 are the reference's (c-core/deciphon.h:34-116); only the ones this
 * library can return are named here.  81.. are new. */
enum
{
  DCP_EFDATA = 3,
  DCP_EFREAD = 5,
  DCP_EFUNCUSE = 8,
  DCP_EFWRITE = 9,
  DCP_EZEROSEQ = 11,
  DCP_EDECODON = 14,
  DCP_ELARGEMODEL = 15,
  DCP_EREADHMMER3 = 17,
  DCP_ENOMEM = 20,
  DCP_EOPENDB = 21,
  DCP_EOPENHMM = 22,
  DCP_EWRITEPROD = 39,
  DCP_ELONGACCESSION = 41,
  DCP_EMANYTHREADS = 42,
  DCP_EMKDIR = 45,
  DCP_ESETGENCODE = 49,
  DCP_EGENCODEID = 50,
  DCP_EH3CDIAL = 51,
  DCP_ESEQABC = 57,
  DCP_ELARGECORESIZE = 63,
  DCP_EENDOFFILE = 66,
  DCP_EENDOFNODES = 67,
  DCP_EDBVERSION = 68,
  DCP_ENOTDBFILE = 69,
  DCP_ENUCLTNOSUPPORT = 71,
  DCP_EDBDNASEQRNA = 72,
  DCP_EDBRNASEQDNA = 73,
  DCP_ENUCLTSEQTU = 74,
  DCP_EINVALNUMPROTEINS = 80,
  DCP_EGPUNODEVICE = 81, /* no CUDA device: this library has no CPU fallback */
  DCP_EGPUFAIL = 82,     /* a CUDA call failed */
  DCP_EGPUNOMEM = 83,    /* device memory exhausted */
  DCP_EGPUINTERNAL = 84, /* invalid argument / state inside the GPU layer */
};

#ifdef __cplusplus
}
#endif
#endif

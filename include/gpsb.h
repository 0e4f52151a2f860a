/*
 * gpsb.h - C ABI of the B200 correlator engine (libgpsb_cuda.so).
 *
 * This is the drop-in boundary for the reference's acquisition / E-P-L tracking hot path.  The
 * reference (iliasam/STM32F4_SDR_GPS, Firmware/project_main, "PM/") has no FFI: its seam is the set of
 * DSP primitives declared in PM/GPS/gps_misc.h:195-216 and called only from PM/GPS/acquisition.c and
 * PM/GPS/tracking.c.  Every entry point below names the reference call sequence it replaces.
 *
 * Conventions
 *   - plain C types only (pointers + sizes); no CUDA or torch types in any signature
 *   - every function returns 0 on success and a negative gpsb_status on failure; the reason is
 *     retrievable with gpsb_last_error() (thread-local).  Nothing aborts, nothing falls back to a CPU
 *     implementation: without a usable sm_100 device every compute entry point fails with
 *     GPSB_ERR_CUDA.
 *   - host pointers are borrowed for the duration of the call; results are written to caller memory
 *   - one context per GPU; calls on one context must be serialised by the caller
 *
 * Signal format (PM/signal_capture.c:9,169; PM/config.h:16,23-28): 1 bit per sample (sign of I),
 * 16.368 Msps, packed LSB-first, 2046 bytes per millisecond.  Inside the context each millisecond
 * occupies a 2048-byte frame (two zero pad bytes) so frames are 16-byte aligned in HBM.
 */
#ifndef GPSB_H
#define GPSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSB_CHIPS          1023u   /* PRN_LENGTH          PM/config.h:28 */
#define GPSB_MS_BYTES       2046u   /* PRN_SPI_WORDS_CNT*2 PM/config.h:27 */
#define GPSB_MS_SAMPLES     16368u  /* BITS_IN_PRN         PM/config.h:26 */
#define GPSB_FRAME_BYTES    2048u   /* device pitch of one millisecond    */
#define GPSB_OFFSETS        2046u   /* half-chip code phases, PM/GPS/acquisition.c:294 */

typedef enum gpsb_status {
    GPSB_OK = 0,
    GPSB_ERR_ARG = -1,     /* bad argument (null pointer, slot / offset out of range, ...) */
    GPSB_ERR_CUDA = -2,    /* CUDA runtime error or no usable device */
    GPSB_ERR_NOMEM = -3,   /* host or device allocation failed */
    GPSB_ERR_STATE = -4    /* call sequence error (e.g. code not set for slot) */
} gpsb_status;

typedef struct gpsb_ctx gpsb_ctx;

/* ---------------------------------------------------------------- life cycle ------------------ */
/* device: CUDA ordinal.  max_sv: number of satellite slots (code tables).  ring_ms: capacity of the
 * HBM signal ring in milliseconds (frame index = ms_index % ring_ms). */
int gpsb_create(gpsb_ctx** out, int device, uint32_t max_sv, uint32_t ring_ms);
void gpsb_destroy(gpsb_ctx* ctx);
const char* gpsb_last_error(void);
/* Library/version probe that never touches the GPU (used by the CPU-only test tier). */
uint32_t gpsb_abi_version(void);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
uint64_t gpsb_launch_count(const gpsb_ctx* ctx);
/* Capacity of the context's signal ring in milliseconds (the ring_ms given to gpsb_create). */
uint32_t gpsb_ring_ms(const gpsb_ctx* ctx);
/* Run all subsequent work of this context on an externally owned CUDA stream (cudaStream_t passed
 * as void*, e.g. torch.cuda.current_stream().cuda_stream).  NULL restores the context's own stream. */
int gpsb_set_stream(gpsb_ctx* ctx, void* cuda_stream);
int gpsb_synchronize(gpsb_ctx* ctx);

/* Device-side timers on the context's stream (CUDA events), so pure-C callers can time kernels the
 * way bench.py does.  slot 0..7. */
int gpsb_timer_start(gpsb_ctx* ctx, uint32_t slot);
int gpsb_timer_stop(gpsb_ctx* ctx, uint32_t slot);
int gpsb_timer_elapsed_ms(gpsb_ctx* ctx, uint32_t slot, float* ms); /* synchronises on the stop event */

/* ---------------------------------------------------------------- resident data --------------- */
/* Replaces gps_channell_prepare()'s product (PM/GPS/gps_misc.c:306-311): the 1023-chip C/A code of a
 * satellite, one byte (0/1) per chip, becomes a device-resident expanded replica table. */
int gpsb_set_code(gpsb_ctx* ctx, uint32_t sv_slot, const uint8_t chips[GPSB_CHIPS]);
/* Device-side generator of the same table from the PRN number (PM/GPS/gps_misc.c:317-372). */
int gpsb_set_code_prn(gpsb_ctx* ctx, uint32_t sv_slot, uint32_t prn);
/* Read back the chips of a slot (for parity tests of the generator). */
int gpsb_get_code(gpsb_ctx* ctx, uint32_t sv_slot, uint8_t chips[GPSB_CHIPS]);

/* Replaces the SPI/DMA capture buffers (PM/signal_capture.c:57-123): copy n_ms milliseconds of packed
 * samples (n_ms * 2046 bytes, contiguous) into ring frames ms0 .. ms0+n_ms-1.  Synchronous. */
int gpsb_upload_signal(gpsb_ctx* ctx, uint32_t ms0, uint32_t n_ms, const uint8_t* packed);
/* Same, but only enqueued on the context stream (packed must stay valid, ideally pinned). */
int gpsb_upload_signal_async(gpsb_ctx* ctx, uint32_t ms0, uint32_t n_ms, const uint8_t* packed);
/* Ingest adaptor for MAX2769-native 2-bit I / 2-bit Q sign-magnitude samples, one byte per sample
 * (bit0 = I sign, bit1 = I mag, bit2 = Q sign, bit3 = Q mag): keeps the I sign bit and packs it
 * LSB-first on the device (the reference front end wires only I1/sign, PM/config.h:16). */
int gpsb_upload_signal_iq2(gpsb_ctx* ctx, uint32_t ms0, uint32_t n_ms, const uint8_t* samples);
/* Copy ring frames back (2046 bytes per ms) - used by tests of the ingest path. */
int gpsb_download_signal(gpsb_ctx* ctx, uint32_t ms0, uint32_t n_ms, uint8_t* packed);

/* ---------------------------------------------------------------- level 1: fused cells -------- */
/*
 * One tracking integrate-and-dump = PM/GPS/tracking.c:115-138:
 *   gps_generate_prn_data2(bits) + gps_shift_to_zero_freq_track(acc0, step32) +
 *   3 x gps_correlation_iq(off_e | off_p | off_l).
 * The NCO words are computed by the host with the reference's own fp32 expression
 * (PM/GPS/gps_misc.c:250-253) so no float arithmetic is re-implemented on the device.
 */
typedef struct gpsb_epl_req {
    uint32_t sv_slot;
    uint32_t ms_index;
    uint32_t acc0;      /* trk->if_freq_accum on entry                */
    uint32_t step32;    /* (uint32_t)((uint64_t)acc_step * 32)         */
    uint16_t off_e;     /* byte (half-chip) offsets, 0..2045           */
    uint16_t off_p;
    uint16_t off_l;
    uint16_t off_bits;  /* sub-byte replica shift, 0..15               */
} gpsb_epl_req;

/* out[6*i .. 6*i+5] = IE,QE,IP,QP,IL,QL of request i (int16, popcount - 8184). Synchronous.
 * Batches of up to 128 cells take the closed-loop fast path (requests in the kernel parameter space,
 * results + completion flag in mapped pinned memory: one launch, no copies, no stream synchronise);
 * gpsb_set_realtime(ctx, 0) forces the staged-copy path used for larger batches. */
int gpsb_track_epl(gpsb_ctx* ctx, uint32_t n, const gpsb_epl_req* req, int16_t* out);
int gpsb_set_realtime(gpsb_ctx* ctx, int enabled);
/* Batches of at least n_cells cells (default 512) are run by k_epl_batch - one warp per cell on a persistent grid,
 * frames streamed from HBM straight into registers - instead of one CTA per cell; 0 = always, UINT32_MAX = never.
 * Both kernels return identical sums. */
int gpsb_set_epl_batch_min(gpsb_ctx* ctx, uint32_t n_cells);
/* Two interchangeable, bit-identical forms of that batch kernel:
 *   GPSB_BATCH_TMA        (default) every warp owns a ring of four 2048-byte frame buffers in shared memory, filled by the
 *                         TMA engine (one cp.async.bulk per cell, four cells ahead, each on its own mbarrier)
 *   GPSB_BATCH_REGISTERS  frames loaded straight into registers, one cell ahead (kept for comparison) */
#define GPSB_BATCH_TMA        0
#define GPSB_BATCH_REGISTERS  1
int gpsb_set_epl_batch_kernel(gpsb_ctx* ctx, int kernel);
/* The prompt arm alone: out[2*i], out[2*i+1] = I, Q of request i at byte offset off_p (off_e / off_l ignored) =
 * gps_generate_prn_data2(off_bits) + gps_shift_to_zero_freq_track(acc0, step32) + ONE gps_correlation_iq
 * (PM/GPS/gps_misc.c:128-145): the reference's single-arm correlation, for long recordings replayed open loop
 * (SURVEY.md section 8(d), config 1 batched). */
int gpsb_prompt_iq(gpsb_ctx* ctx, uint32_t n, const gpsb_epl_req* req, int16_t* out);

/* Tracking session: keeps one resident CTA per channel slot polling a command slot in mapped pinned
 * host memory, so that a gpsb_track_epl call of n <= n_slots cells costs one PCIe round trip instead
 * of a kernel launch (the 1-kHz loop of PM/main.c:134-156 is serial per channel: latency is the whole
 * budget).  The resident kernel leaves by itself after 250 ms without a command or 20 s of life and is
 * re-launched transparently, so a stalled host can never wedge the GPU.  While a session is open the
 * other entry points keep working (they use the context stream; the session has its own). */
int gpsb_session_begin(gpsb_ctx* ctx, uint32_t n_slots);
int gpsb_session_end(gpsb_ctx* ctx);
uint32_t gpsb_session_slots(const gpsb_ctx* ctx);
/* The two halves of one session exchange for a single slot.  A slot must be driven by one thread at a
 * time; different slots may be driven concurrently from different threads (every channel of the 1-kHz
 * loop is independent, PM/GPS/gps_misc.h:184: all state is per gps_ch_t).  req == NULL posts a keep-alive.
 * While a session is open gpsb_search and the staged gpsb_track_epl path remain usable from any thread
 * (they serialise on an internal lock). */
int gpsb_session_post(gpsb_ctx* ctx, uint32_t slot, const gpsb_epl_req* req, uint32_t* seq);
int gpsb_session_wait(gpsb_ctx* ctx, uint32_t slot, uint32_t seq, int16_t out6[6]);

/* ---------------------------------------------------------------- device-resident tracking loop - */
/*
 * The whole closed loop of PM/GPS/tracking.c:92-170 + nav_data.c:46-138 for n_ch channels over n_ms
 * consecutive milliseconds in ONE launch (k_track_run): per channel one CTA keeps the channel record in
 * shared memory, correlates each millisecond from the HBM signal ring and runs the reference's DLL / PLL /
 * FLL, false-lock check, NCO planning, bit synchronisation and word assembly on the device between two
 * correlations - no host round trip per millisecond.  "Every satellite every millisecond" schedule: slot
 * index = ms % 4, ms counter = frame index.
 *
 *   channels  n_ch records of gps_ch_t (include/gpsb_host.h; PM/GPS/gps_misc.h:184-193), channel_bytes each,
 *             in GPS_TRACKING_RUN (or GPS_PRE_TRACK_DONE); satellite slot == prn; updated in place
 *   aux       n_ch records of the per-channel scratch (gpsb_aux, core/gpsb_loop_core.h), aux_bytes each
 *   iq_log    NULL or int16 [n_ms][n_ch][6]  IE,QE,IP,QP,IL,QL
 *   nav_log   NULL or int8  [n_ms][n_ch]     -1, or the 20-ms data bit handed to the word assembler that ms
 *   results   n_ch records: milliseconds completed and why the channel stopped early (0 = it did not):
 *             1 the channel was not in a tracking state (nothing done), 2 the early+late power of millisecond
 *             done_ms was zero (0/0 in the DLL): its sums are in iq[], the filters were not run - the caller
 *             finishes that millisecond on the host; 3 (streaming runs only) the producer did not deliver a frame
 *             in time.  Rows of the logs past done_ms are undefined.
 * Record sizes are checked against the library's own (gpsb_track_loop_record_bytes).
 */
typedef struct gpsb_loop_result {
    uint32_t done_ms;
    int32_t  stop;
    int16_t  iq[6];
    uint32_t reserved;
} gpsb_loop_result;
int gpsb_track_loop(gpsb_ctx* ctx, uint32_t n_ch, void* channels, uint32_t channel_bytes, void* aux,
                    uint32_t aux_bytes, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log,
                    gpsb_loop_result* results);
void gpsb_track_loop_record_bytes(uint32_t* channel_bytes, uint32_t* aux_bytes);
/* Same loop on device-resident records (no copies, no synchronise): for callers that keep the channel
 * records in HBM between runs and for timing the kernel alone.  All pointers are device pointers. */
int gpsb_track_loop_dev(gpsb_ctx* ctx, uint32_t n_ch, void* d_channels, void* d_aux, uint32_t ms0, uint32_t n_ms,
                        int16_t* d_iq_log, int8_t* d_nav_log, gpsb_loop_result* d_results);
int gpsb_track_loop_dev_ex(gpsb_ctx* ctx, uint32_t n_ch, void* d_channels, void* d_aux, uint32_t ms0, uint32_t n_ms,
                           int16_t* d_iq_log, int8_t* d_nav_log, gpsb_loop_result* d_results, uint32_t flags);

/* ---------------------------------------------------------------- streaming ingest -------------- */
/*
 * Replaces the SPI/DMA double buffer of PM/signal_capture.c:57-123 (half/full-transfer interrupts flip
 * spi_curr_ready_rx_buf while the main loop consumes the other half): here the host keeps DMA-ing chunks of
 * packed samples into the HBM ring on a copy stream WHILE a k_track_run launch is consuming them.  After each
 * chunk the stream moves a watermark (first millisecond not yet uploaded) in device memory; a loop started with
 * GPSB_LOOP_STREAMING never fetches a frame at or above the watermark - it waits for the producer, at most the
 * stream time-out per frame (then the run ends with stop == 3, LC_STOP_STARVED, its done_ms complete and valid).
 *
 *   gpsb_stream_reset(ctx, ms)      frames below ms are declared present (synchronous); call before a run
 *   gpsb_stream_push(ctx, ms0, n, p) enqueue n milliseconds (n * 2046 bytes at p, which must stay valid - ideally
 *                                    pinned - until gpsb_stream_wait) and then the watermark ms0 + n; pushes must be
 *                                    issued in ascending order and must not lap the consumer: chunk [a, a+n) replaces
 *                                    frames a-ring_ms .., so push it only once gpsb_stream_progress() >= a+n-ring_ms
 *                                    (the millisecond every channel of the running loop has completed, updated every
 *                                    64 ms; a channel that leaves the loop reports the end of the run)
 *   gpsb_stream_wait(ctx)           all pushes so far have landed
 *   gpsb_track_loop_begin/_end      the two halves of gpsb_track_loop: _begin uploads the records and launches the loop
 *                                    without waiting, _end waits and copies records and logs back.  Between the two
 *                                    only gpsb_stream_* may be called on the context.
 */
#define GPSB_LOOP_STREAMING 1u
/* gpsb_track_loop_dev_ex only (records in device memory, which the library cannot look into): the caller states that no
 * channel has the slot-phase walk of include/gpsb_host.h (gpsb_rx_set_slot_walk) switched on, a slot phase other than 0
 * or an idle gap pending, and gets the build of the loop without it (about 3 % faster).  A record that does not qualify
 * is refused (stop == 1).  Entry points that take HOST records decide this themselves. */
#define GPSB_LOOP_FIXED_SLOTS 2u
int gpsb_stream_reset(gpsb_ctx* ctx, uint32_t ms_valid_upto);
int gpsb_stream_push(gpsb_ctx* ctx, uint32_t ms0, uint32_t n_ms, const uint8_t* packed);
/* Same for the MAX2769-native 2-bit I / 2-bit Q container of gpsb_upload_signal_iq2 (one byte per sample, n * 16368
 * bytes): copy and pack kernel run on the copy stream, then the watermark moves. */
int gpsb_stream_push_iq2(gpsb_ctx* ctx, uint32_t ms0, uint32_t n_ms, const uint8_t* samples);
int gpsb_stream_wait(gpsb_ctx* ctx);
uint32_t gpsb_stream_progress(const gpsb_ctx* ctx, uint32_t n_ch);
int gpsb_stream_set_timeout_ms(gpsb_ctx* ctx, uint32_t ms);
uint32_t gpsb_stream_timeout_ms(const gpsb_ctx* ctx);
/* 1 while the loop started by gpsb_track_loop_begin is still executing (a producer waiting for ring space checks
 * this so that it never waits for a consumer that has already left). */
int gpsb_stream_loop_running(gpsb_ctx* ctx);
/* The producer gives up: a streaming loop that is (or comes to be) waiting for a frame ends at once with stop == 3
 * instead of waiting out the time-out; its done_ms are complete and exact.  Cleared by gpsb_stream_reset.
 * gpsb_stream_copies_pending: 1 while pushes are still queued behind something on the copy stream. */
int gpsb_stream_abort(gpsb_ctx* ctx);
int gpsb_stream_copies_pending(gpsb_ctx* ctx);
int gpsb_track_loop_begin(gpsb_ctx* ctx, uint32_t n_ch, void* channels, uint32_t channel_bytes, void* aux,
                          uint32_t aux_bytes, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log,
                          gpsb_loop_result* results, uint32_t flags);
int gpsb_track_loop_end(gpsb_ctx* ctx);

/*
 * Device-resident code-phase rounds = PM/GPS/acquisition.c:134-275 (the CODE_PHASE_SEARCH1..3 states of
 * acquisition_process_channel: window search at the Doppler the sweep found, 32-cell histogram vote, narrowing) for a
 * whole span of snapshots in ONE launch, one CTA per channel - instead of one search launch and host round trip per
 * look-ahead window.  Channel and aux records as for gpsb_track_loop (host memory, copied in and out).
 *   mode[i] 0  channel i is left alone
 *           1  snapshots ms0, ms0+1, .. until the channel's acquisition state is no longer in busy_mask (bit s = state s),
 *              at most n_ms
 *           2  exactly n_ms snapshots (a channel that is served but does not keep a round alive)
 *   used[i]    snapshots channel i consumed.
 * The frames ms0 .. ms0 + n_ms - 1 must be in the ring.  The snapshot's millisecond is the reference's packet counter
 * (signal_capture_get_packet_cnt) for the round's time stamps.
 */
int gpsb_code_rounds(gpsb_ctx* ctx, uint32_t n_ch, void* channels, uint32_t channel_bytes, void* aux, uint32_t aux_bytes,
                     uint32_t ms0, uint32_t n_ms, uint32_t busy_mask, const uint8_t* mode, uint32_t* used);

/* Self-test support: the loop's two float discriminators evaluated on the device for ip in
 * [ip_lo, ip_lo + n_ip), every qp in [-8184, 8184]; out[(ip - ip_lo) * 16369 + qp + 8184] (host memory).
 * kind 0: Costas error in units of pi (PM/GPS/tracking.c:180-183), kind 1: FLL angle (tracking.c:232). */
int gpsb_l0_loop_math(gpsb_ctx* ctx, int kind, int32_t ip_lo, uint32_t n_ip, float* out);

/*
 * One acquisition / pre-track cell = PM/GPS/acquisition.c:282-294 (freq search),
 * :198-209 (code-phase search) and PM/GPS/tracking.c:403-426 (pre-track):
 *   gps_generate_prn_data2(bits) + gps_shift_to_zero_freq(step32, phase 0) +
 *   correlation_search(start, stop).
 */
typedef struct gpsb_search_req {
    uint32_t sv_slot;
    uint32_t ms_index;
    uint32_t acc0;      /* 0 for the reference's stateless mixer       */
    uint32_t step32;
    uint16_t off_bits;
    uint16_t start;     /* first offset, inclusive                      */
    uint16_t stop;      /* last offset, exclusive, <= 2046              */
    uint16_t flags;     /* reserved, 0                                  */
} gpsb_search_req;

typedef struct gpsb_search_res {
    uint16_t max;       /* correlation_search() return value            */
    uint16_t phase;     /* *phase: first offset holding the maximum     */
    uint16_t avg;       /* *aver_val: sum / 2046                        */
    uint16_t reserved;
} gpsb_search_res;

int gpsb_search(gpsb_ctx* ctx, uint32_t n, const gpsb_search_req* req, gpsb_search_res* res);

/* Per-offset (I,Q) of one cell: iq[2*k], iq[2*k+1] for offset start+k (gps_correlation_iq over a
 * window, PM/GPS/gps_misc.c:128-145). */
int gpsb_search_iq(gpsb_ctx* ctx, const gpsb_search_req* req, int16_t* iq);

/* Full-sky sweep: every (sv, bin, ms) cell with the full 2046-offset window, the cold-acquisition
 * workload.  Cells are ordered (sv, bin, ms); step32[b] is the NCO word of bin b.  Results land in
 * res[(sv*n_bins + b)*n_ms + m].  sv_slots[] lists the slots to search. */
int gpsb_sweep(gpsb_ctx* ctx, const uint32_t* sv_slots, uint32_t n_sv, const uint32_t* step32,
               uint32_t n_bins, uint32_t ms0, uint32_t n_ms, uint32_t off_bits,
               gpsb_search_res* res);

/* Two interchangeable, bit-identical implementations of the wide-window search (north_star: "batched
 * direct correlator or ... chosen by measurement"):
 *   GPSB_SWEEP_DIRECT  XOR + POPC per 32 samples for every offset (POPC-pipe bound)
 *   GPSB_SWEEP_DP4A    per-byte popcounts once per millisecond, then a +-1-chip x small-integer circular
 *                      correlation on the integer dot-product pipe (default; about 8x fewer issue slots); one
 *                      256-thread CTA per offset parity, two per SM, so epilogues overlap main loops
 * Applies to gpsb_sweep / gpsb_sweep_dev and to gpsb_search requests whose window is >= 384 offsets. */
#define GPSB_SWEEP_DIRECT 0
#define GPSB_SWEEP_DP4A   1
#define GPSB_SWEEP_DP4A_FULL 2     /* the dp4a search with one 512-thread CTA per cell group (kept for comparison) */
int gpsb_set_sweep_method(gpsb_ctx* ctx, int method);

/* ---- device-resident variants: request / result arrays already in device memory, enqueued on the
 *      context stream without synchronising.  Used for kernel-only timing and for results that are
 *      gathered across GPUs (NCCL) before they are read. */
int gpsb_track_epl_dev(gpsb_ctx* ctx, uint32_t n, const gpsb_epl_req* d_req, int16_t* d_out);
int gpsb_prompt_iq_dev(gpsb_ctx* ctx, uint32_t n, const gpsb_epl_req* d_req, int16_t* d_out);
int gpsb_search_dev(gpsb_ctx* ctx, uint32_t n, const gpsb_search_req* d_req, gpsb_search_res* d_res);
int gpsb_sweep_dev(gpsb_ctx* ctx, const uint32_t* d_sv_slots, uint32_t n_sv, const uint32_t* d_step32,
                   uint32_t n_bins, uint32_t ms0, uint32_t n_ms, uint32_t off_bits,
                   gpsb_search_res* d_res);

/* ---------------------------------------------------------------- multi-GPU: sharded sweep + all-gather ----
 * One process and one context per GPU.  The cells of a sweep are independent until the host vote (PM/GPS/acquisition.c:
 * 280-297 has no cross-cell data flow), so the (bin, ms) cell groups are dealt round-robin over the ranks - every rank
 * keeps whole 8-satellite tiles of the dp4a search - and the {max, phase, avg} triples are exchanged ONCE per sweep with
 * ncclAllGather over NVLink / NVSwitch, straight out of device memory on the context stream (no host bounce), so that
 * every rank holds the whole grid and can run the reference's votes.  Tracking needs no collective: channels are
 * independent (PM/GPS/gps_misc.h:184).  NCCL is bound at run time (dlopen of libnccl.so.2; a copy already loaded into
 * the process, e.g. torch's, is reused); a single-GPU user needs none.
 *
 *   gpsb_comm_unique_id   rank 0 makes the 128-byte rendezvous id; the caller hands it to every rank by any means
 *                         (torch.distributed broadcast, MPI, a file)
 *   gpsb_comm_init        ncclCommInitRank on the context's device; collective over all ranks
 *   gpsb_sweep_gather     gpsb_sweep, sharded and gathered: collective, every rank passes the same arguments and
 *                         receives the whole grid res[(sv*n_bins + b)*n_ms + m]; without a communicator (size 1)
 *                         it is gpsb_sweep
 *   gpsb_sweep_gather_dev the same enqueued on the context stream without synchronising; *d_grid = the grid in
 *                         device memory owned by the context (valid until the next call) */
typedef struct gpsb_nccl_id { char bytes[128]; } gpsb_nccl_id;
int gpsb_comm_unique_id(gpsb_nccl_id* out);
int gpsb_comm_init(gpsb_ctx* ctx, int rank, int n_ranks, const gpsb_nccl_id* id);
int gpsb_comm_destroy(gpsb_ctx* ctx);
int gpsb_comm_rank(const gpsb_ctx* ctx);
int gpsb_comm_size(const gpsb_ctx* ctx);
int gpsb_sweep_gather(gpsb_ctx* ctx, const uint32_t* sv_slots, uint32_t n_sv, const uint32_t* step32,
                      uint32_t n_bins, uint32_t ms0, uint32_t n_ms, uint32_t off_bits, gpsb_search_res* res);
int gpsb_sweep_gather_dev(gpsb_ctx* ctx, const uint32_t* d_sv_slots, uint32_t n_sv, const uint32_t* d_step32,
                          uint32_t n_bins, uint32_t ms0, uint32_t n_ms, uint32_t off_bits, gpsb_search_res** d_grid);

/* ---------------------------------------------------------------- level 0: the reference primitives
 * Same arithmetic and argument meaning as PM/GPS/gps_misc.h:198-216 on caller-owned HOST buffers
 * (each call round-trips through the device; meant for parity tests and for piecewise migration).
 * Buffers follow the reference: prn/data arrays are 1023 little-endian uint16 words (2046 bytes). */
/* gps_generate_prn_data2 (gps_misc.c:282-300); data receives 1023 words. */
int gpsb_l0_generate_prn_data2(gpsb_ctx* ctx, const uint8_t chips[GPSB_CHIPS], uint16_t* data,
                               uint16_t offset_bits);
/* gps_shift_to_zero_freq / _track (gps_misc.c:211-274) with explicit NCO words; writes 2044 bytes to
 * data_i and data_q (bytes 2044..2045 untouched, as in the reference); *acc_out = acc0 + 511*step32. */
int gpsb_l0_shift_to_zero_freq(gpsb_ctx* ctx, const uint8_t* signal_data, uint8_t* data_i,
                               uint8_t* data_q, uint32_t acc0, uint32_t step32, uint32_t* acc_out);
/* gps_correlation_iq (gps_misc.c:128-145) */
int gpsb_l0_correlation_iq(gpsb_ctx* ctx, const uint16_t* prn_p, const uint16_t* data_i,
                           const uint16_t* data_q, uint16_t offset, int16_t* res_i, int16_t* res_q);
/* gps_correlation8 (gps_misc.c:98-122) */
int gpsb_l0_correlation8(gpsb_ctx* ctx, const uint16_t* prn_p, const uint16_t* data_i,
                         const uint16_t* data_q, uint16_t offset, int16_t* res);
/* correlation_search (gps_misc.c:155-191) */
int gpsb_l0_correlation_search(gpsb_ctx* ctx, const uint16_t* prn_p, const uint16_t* data_i,
                               const uint16_t* data_q, uint16_t start_shift, uint16_t stop_shift,
                               uint16_t* aver_val, uint16_t* phase, uint16_t* max_val);

#ifdef __cplusplus
}
#endif
#endif /* GPSB_H */

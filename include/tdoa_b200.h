/*
 * tdoa_b200.h -- C ABI of libtdoa_b200.so, the B200 (sm_100a) TDOA correlation engine.
 *
 * Drop-in boundary for the processing stage of KX0U-Jim/tdoa-geolocation.  The
 * reference has no FFI seam of its own (a single `package main`), so each entry
 * point below names the reference method it replaces (file:line in the reference
 * tree).  Plain pointers and sizes only; the caller owns every host pointer and the
 * library never retains one after a call returns (cgo pointer rules).
 *
 * Error convention: every function returns 0 on success or a negative TDOA_E_* code;
 * tdoa_last_error() returns the message of the last failure on that engine
 * (NULL engine: the last tdoa_create failure of the calling thread).  There is no
 * CPU fallback: without an sm_100 device tdoa_create fails with TDOA_E_NODEVICE.
 *
 * Threading: one caller at a time per engine (the reference is single-threaded);
 * every entry point selects the engine's device itself, so it is safe to call from
 * migrating goroutines/threads without pinning.
 */
#ifndef TDOA_B200_H
#define TDOA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDOA_API __attribute__((visibility("default")))

/* error codes */
#define TDOA_OK 0
#define TDOA_E_INVALID (-1)   /* bad argument                                      */
#define TDOA_E_NODEVICE (-2)  /* no sm_100 CUDA device / CUDA runtime unusable     */
#define TDOA_E_CUDA (-3)      /* a CUDA call failed (message has the CUDA error)   */
#define TDOA_E_NOMEM (-4)     /* device or pinned-host allocation failed           */
#define TDOA_E_STATE (-5)     /* call sequence error (e.g. station not loaded)     */
#define TDOA_E_SINGULAR (-6)  /* solveTDOA: singular Jacobian (processor.go:997)   */
#define TDOA_E_IO (-7)        /* loadIQData: open / stat / read failed (:170-191)  */

/* processing modes (tdoa_config.mode) */
#define TDOA_MODE_SOURCE 0   /* processor.go as committed: complex64 box-car path,
                                1000-sample blocks, sqrt(N) gain, maxLag 20000      */
#define TDOA_MODE_BINARY 1   /* shipped `processor` ELF: FM discriminator / envelope /
                                weak 3-way preprocess, real correlator over a template
                                shortened by maxLag, 10000-sample blocks, 120-sample
                                sanity re-search                                     */
#define TDOA_MODE_EXTENDED 2 /* BINARY preprocessing + two-sided lag search with
                                parabolic sub-sample refinement (engine-defined)     */

/* signal kinds of the dual-frequency capture (processor.go:208-267) */
#define TDOA_KIND_REF 0 /* blocks 1 and 3 concatenated */
#define TDOA_KIND_TGT 1 /* block 2 */

/* tdoa_peak.flags */
#define TDOA_PEAK_RESEARCHED 0x1u /* sanity re-search replaced the first-pass peak   */
#define TDOA_PEAK_EDGE 0x2u       /* peak on the edge of the lag range (frac = 0)    */
#define TDOA_PEAK_BRUTE 0x4u      /* every lag was evaluated in the time domain      */
#define TDOA_PEAK_EMPTY 0x8u      /* empty input: (0, 0.0) as processor.go:622-625   */
#define TDOA_PEAK_NCAND(f) (((f) >> 16) & 0xffu)   /* lags re-evaluated exactly (FFT path)   */
#define TDOA_PEAK_BRANCH1(f) (((f) >> 8) & 3u)  /* preprocess branch of signal 1      */
#define TDOA_PEAK_BRANCH2(f) (((f) >> 10) & 3u) /* 0 strong/std, 1 moderate, 2 weak   */

typedef struct tdoa_engine tdoa_engine; /* opaque: device memory, streams, plans */

/* One correlation peak; 32 bytes; also the record exchanged between GPUs. */
typedef struct {
    int32_t lag;       /* integer lag of the peak, samples (delay of signal 2)       */
    uint32_t flags;    /* TDOA_PEAK_*                                                */
    double corr;       /* correlation at the peak, signed (the Go float64 return)    */
    float frac;        /* sub-sample offset in [-0.5,0.5]; 0 in SOURCE/BINARY modes  */
    float margin;      /* (|peak|-|runner-up|)/|peak| among exactly evaluated lags   */
    int32_t first_lag; /* first-pass lag before the sanity re-search                 */
    int32_t n_blocks;  /* whole blocks that contributed (processor.go:691-712)       */
} tdoa_peak;

typedef struct {
    double sample_rate;    /* 2e6 (processor.go:821)                                 */
    int32_t mode;          /* TDOA_MODE_*                                            */
    int32_t n_stations;    /* stations the engine holds (>= 2)                       */
    int32_t chunk_samples; /* "test chunk": 2000000 source (processor.go:772),
                              1000000 binary; 0 = whole signal                       */
    int32_t max_lag;       /* 20000 source (processor.go:633), 2000 binary;
                              EXTENDED: lags -max_lag..+max_lag                      */
    int32_t block_size;    /* 1000 source (processor.go:682), 10000 binary           */
    int32_t sanity_lag;    /* 120 (binary re-search window); 0 disables              */
    int32_t fast_demod;    /* 0: f64 products + f64 atan2 as the reference;
                              1: f32 discriminator (<= 2 ulp), EXTENDED only default;
                              2: as 0, every sample through the full-accuracy arctangent
                                 (no fast path; a test switch)                        */
    int32_t use_fft;       /* 1: FFT candidate search + exact re-evaluation (2 x 2 station
                              tiles; station spectra parked once per segment for >= 10
                              pairs per window; a 2^21-point transform for >= 8192 lags);
                              0: evaluate every lag in the time domain;
                              test switches: 2 one transform per pair, 3 no 2^21-point
                              transform, 4 no parked spectra, 5 plain one-add-at-a-
                              time walk for the sequential DC sum, 6 EXTENDED weak
                              branch as four kernels over planes (not the fused pass) */
    int32_t device;        /* CUDA device ordinal                                    */
    int32_t seq_dc_limit;  /* removeDCBias (processor.go:299-319): signals of up to this many
                              samples get the reference's sequential f32 accumulator,
                              bit for bit; longer ones (which the reference never
                              processes: it truncates to its chunk) an exactly rounded
                              sum.  0 = mode default (4194304 SOURCE/BINARY, never in
                              EXTENDED); -1 = never                                   */
    int32_t copy_chunk;    /* tdoa_load_u8_pinned: samples per host->device copy chunk
                              (multiple of 4096); 0 = 16777216 (32 MB of capture)    */
    int32_t guard_samples; /* samples dropped at the start of blocks 2 and 3 (the retune
                              transient, collector.go:85 / rtl_sdr.c:117-135); 0 = the
                              reference's split (processor.go:208-267)                */
    int32_t decimate;      /* EXTENDED mode only: D > 1 appends a decimating box-car
                              (mean of D consecutive samples, rtl_fm.c:302-322 style) to the
                              preprocessing chain; correlation at fs / D over max_lag / D
                              lags, records in samples of the capture (lag + frac)       */
    int32_t serial_kinds;  /* tdoa_process: 0 = the TGT pair loop is queued on a second stream beside
                              the REF pair loop (their kernels share the SMs); 1 = one stream, one
                              kernel at a time (what the per-kernel timings of tdoa_get_stats
                              need to mean anything)                                        */
    int32_t n_devices;     /* 0 or 1: one GPU (`device`).  N > 1: this ONE process drives the N GPUs
                              device .. device + N - 1: every capture is loaded to each of them and
                              tdoa_xcorr over >= 2 windows deals the windows round-robin over them
                              (processor.go:816-850 is the loop being sharded), the peak records
                              meeting in one NCCL all-gather over NVLink; every other call runs on
                              `device`.  One process per GPU instead: tdoa_comm_init            */
    int32_t reserved[1];
} tdoa_config;

/* Fill *cfg with the reference-matching defaults of `mode`. */
TDOA_API int tdoa_default_config(int32_t mode, tdoa_config *cfg);

TDOA_API int tdoa_create(tdoa_engine **out, const tdoa_config *cfg);
TDOA_API void tdoa_destroy(tdoa_engine *e);
TDOA_API const char *tdoa_last_error(const tdoa_engine *e);

/* Pinned host buffers for callers that want full PCIe rate (optional). */
TDOA_API void *tdoa_host_alloc(size_t nbytes);
TDOA_API void tdoa_host_free(void *p);

/* loadIQData + extractReferenceSignal + extractTargetSignal (processor.go:166-267):
 * hands one station's whole .dat capture (interleaved uint8 I,Q) to the engine.
 * The bytes are copied to the device before the call returns. */
TDOA_API int tdoa_load_u8(tdoa_engine *e, int32_t station, const uint8_t *iq, size_t nbytes);
/* loadIQData on a file (processor.go:166-205): streams the .dat capture to the device
 * through two pinned 32 MB staging buffers (disk read of one piece overlaps the PCIe copy
 * of the previous).  *n_samples (may be NULL) = file size / 2 (:183).  Errors carry the
 * reference's messages ("failed to open file: ...", TDOA_E_IO). */
TDOA_API int tdoa_load_file(tdoa_engine *e, int32_t station, const char *path, int64_t *n_samples);
/* Same for a capture in memory obtained from tdoa_host_alloc (C memory, outside the Go
 * heap, so the cgo pointer rules do not apply): returns at once.  The host -> device
 * copies are queued by the next call that needs the capture -- the blocks of the signal
 * kind that call asks for first, across all stations, then the rest -- and the FM
 * discriminator follows them chunk by chunk, so preprocessing overlaps the PCIe
 * transfer.  The buffer must stay unchanged until tdoa_synchronize() returns or both
 * kinds have been correlated.  Captures under 384 KiB are copied synchronously. */
TDOA_API int tdoa_load_u8_pinned(tdoa_engine *e, int32_t station, const uint8_t *pinned_iq, size_t nbytes);
/* Same, capture already resident in device memory; not copied, caller keeps it alive.
 * d_iq must be 16-byte aligned (cudaMalloc memory is; an offset into a buffer may not be):
 * the unpack / discriminator kernels fetch the capture in 16-byte words.  A misaligned
 * pointer is rejected with TDOA_E_INVALID. */
TDOA_API int tdoa_load_u8_device(tdoa_engine *e, int32_t station, const uint8_t *d_iq, size_t nbytes);

/* loadIQData parity probe (processor.go:193-201): complex64 samples
 * [first, first+count) of a station's capture, written to host out_c64[2*count]. */
TDOA_API int tdoa_unpack(tdoa_engine *e, int32_t station, int64_t first, int64_t count, float *out_c64);

/* preprocessSignal probe (processor.go:469-499 / ELF 0x49cd40): preprocesses samples
 * [start, start+len) of `kind` of `station` in the engine's mode and returns the
 * normalised complex64 signal, its initial power and the branch taken.  With
 * decimate = D only the first len / D entries of out_c64 are written. */
TDOA_API int tdoa_preprocess(tdoa_engine *e, int32_t station, int32_t kind, int64_t start, int64_t len,
                             float *out_c64, double *power, int32_t *branch);

/* ProcessTDOA pair loops (processor.go:816-850): correlates every station pair i<j
 * (lexicographic, = argv order) of `kind` over n_windows windows
 * [win_start + w*hop, +win_len).  win_len = 0 means the config's chunk (0 = whole
 * signal).  out[w * P + p], P = S(S-1)/2. */
TDOA_API int tdoa_xcorr(tdoa_engine *e, int32_t kind, int64_t win_start, int64_t win_len,
                        int32_t n_windows, int64_t hop, tdoa_peak *out);
/* Same, peaks left in device memory (e.g. an NCCL send buffer); d_out[n_windows*P]. */
TDOA_API int tdoa_xcorr_device(tdoa_engine *e, int32_t kind, int64_t win_start, int64_t win_len,
                               int32_t n_windows, int64_t hop, tdoa_peak *d_out);

/* Both pair loops of ProcessTDOA over windows in ONE call (processor.go:816-830 REF, then :836-850 TGT):
 * ref_out[n_ref_windows * P], tgt_out[n_tgt_windows * P] as tdoa_xcorr(TDOA_KIND_REF, ...) followed by
 * tdoa_xcorr(TDOA_KIND_TGT, ...) would fill them (same win_start / win_len / hop for both kinds; either
 * count may be 0).  On one GPU it is exactly those two calls.  On a multi-GPU engine (n_devices > 1 or
 * tdoa_comm_init) the windows of BOTH kinds are dealt over the ranks as one list and correlated before the
 * first gather, so the ranks meet once per call instead of once per kind -- with 66 + 33 windows on 8 GPUs
 * the busiest rank works 13 window times instead of 9 + 5. */
TDOA_API int tdoa_xcorr_windows(tdoa_engine *e, int64_t win_start, int64_t win_len, int32_t n_ref_windows,
                                int32_t n_tgt_windows, int64_t hop, tdoa_peak *ref_out, tdoa_peak *tgt_out);

/* ---- more than one GPU, one process per GPU (torchrun, MPI, ...).  Rank 0 asks for an id
 * (ncclGetUniqueId), the launcher hands its 128 bytes to every rank, every rank joins with its own
 * engine (ncclCommInitRank on the engine's device).  From then on tdoa_xcorr / tdoa_xcorr_device
 * over >= 2 windows are COLLECTIVE: every rank makes the same call; rank r correlates the windows w
 * with (cursor + w) % world == r -- all pairs of a window on one GPU, so no data-path exchange --
 * its peak kernel writes the records straight into the gather buffer, one ncclAllGather
 * completes the table on every rank, and every rank returns all n_windows * P records.  cursor
 * starts at 0 and advances by n_windows after every sharded call.  Each rank loads the captures it
 * correlates (all of them, or -- tdoa_load_u8_pinned -- only the bytes its windows need).
 * tdoa_xcorr_info afterwards describes the first window this rank processed. */
TDOA_API int tdoa_comm_unique_id(uint8_t id[128]);
TDOA_API int tdoa_comm_init(tdoa_engine *e, const uint8_t id[128], int32_t rank, int32_t world);
TDOA_API int tdoa_comm_rank(tdoa_engine *e, int32_t *rank, int32_t *world);
/* The dealing rule itself (host arithmetic, no GPU): rank's first window and how many it gets. */
TDOA_API int tdoa_shard_windows(int32_t n_windows, int32_t rank, int32_t world, int32_t cursor, int32_t *first,
                                int32_t *count);

/* ProcessTDOA from the pair loops to the fix as ONE call (processor.go:816-929 on window 0
 * with the config's chunk): the REF pair loop, the TGT pair loop, dt = delay / fs, the
 * mode's time differences (SOURCE: target only, :853; binary modes: target - reference,
 * pair by pair), range differences (:899-903) and solveTDOA -- all queued on the device
 * without an intermediate host synchronisation.  ref_out / tgt_out [P]; time_diffs,
 * range_diffs [P] (may be NULL); fix_llh[3]; *fix_status 0 or TDOA_E_SINGULAR.
 * The fix is processor.go's solveTDOA (:932-1020) on rd[0], rd[1] in every mode; the shipped
 * binary's own revision of the solver (filter, exactly-two rule) is tdoa_solve_binary on the
 * range_diffs returned here.  tdoa_xcorr_info() afterwards reports both kinds. */
TDOA_API int tdoa_process(tdoa_engine *e, const double *stations_llh, tdoa_peak *ref_out, tdoa_peak *tgt_out,
                          double *time_diffs, double *range_diffs, double *fix_llh, int32_t *fix_status,
                          int32_t *fix_iters);

/* crossCorrelate seam (processor.go:619-643): two host complex64 slices
 * (interleaved re,im), returns (delay, correlation).  Empty input -> (0, 0.0). */
TDOA_API int tdoa_cross_correlate(tdoa_engine *e, const float *sig1_c64, int64_t n1,
                                  const float *sig2_c64, int64_t n2, tdoa_peak *out);

/* latLonToECEF / calculateBaseline (processor.go:125-163): baselines[p] in metres
 * for pairs i<j of stations_llh[S][3] (lat deg, lon deg, elev m). */
TDOA_API int tdoa_baselines(tdoa_engine *e, const double *stations_llh, int32_t n_stations, double *baselines);

/* solveTDOA (processor.go:932-1020) batched: n_sets range-difference sets,
 * range_diffs[set * rd_stride + p].  out_llh[set][3]; status[set] = 0 ok,
 * TDOA_E_SINGULAR singular Jacobian; iters[set] (may be NULL). */
TDOA_API int tdoa_solve(tdoa_engine *e, const double *stations_llh, int32_t n_stations,
                        const double *range_diffs, int32_t n_sets, int32_t rd_stride,
                        double *out_llh, int32_t *status, int32_t *iters);

/* solveTDOA of the SHIPPED BINARY (ELF 0x4a0360; a later revision than processor.go:932-1020,
 * restated from its disassembly and pinned by the iteration traces it prints): the range
 * differences with |rd| <= 20400 m are kept in order; the binary solves only when exactly
 * two remain (res1 = (r2 - r1) - valid[0], res2 = (r3 - r1) - valid[1]; Newton step times 0.7,
 * steps over 1000 m scaled to 1000 m first; single-equation fall-back when |det| < 1e-12;
 * converged when both residuals < 1 m; at most 10 iterations; Z frozen).
 * *status: 0 a fix in out_llh[3]; 1 fewer than two valid ("insufficient valid measurements: only
 * %d of %d range differences are reliable"); 2 more than two valid ("no valid range difference
 * measurements remain"); 3 "singular Jacobian matrix at iteration %d (det=%.2e)".
 * n_valid, n_iter (iterations whose trace line the binary prints), converged, and
 * trace[10][5] = {det, res1, res2, step length, code: 0 plain step, 1 step limited, 2 / 3
 * single equation 1 / 2} may each be NULL. */
TDOA_API int tdoa_solve_binary(tdoa_engine *e, const double *stations_llh, int32_t n_stations,
                               const double *range_diffs, int32_t n_rd, double *out_llh, int32_t *status,
                               int32_t *n_valid, int32_t *n_iter, int32_t *converged, double *trace);

/* Dense lat-lon grid multilateration (no reference equivalent): for each set the
 * cell minimising sum_{i<j} ((r_j - r_i) - rd_ij)^2, the terms added pair by pair in
 * that order; ties go to the lowest linear index lat_index * nlon + lon_index.
 * grid_desc = {lat0, lon0, dlat, dlon, nlat, nlon, elev}.
 * The cells are RANKED by an expanded form of the cost (16 multiply-adds per cell and
 * set) whose rounding error is bounded; only cells within that bound of the minimum are
 * evaluated by the sum above, and the arg-min is taken over those values -- index and
 * cost are what evaluating every cell gives, bit for bit (use_fft = 0 does evaluate
 * every cell; a set holding a NaN, or a plateau of more than 64 ties per set, falls
 * back to that).  range_diffs is read on the host as well (one pass, n_sets * pairs). */
TDOA_API int tdoa_grid(tdoa_engine *e, const double *stations_llh, int32_t n_stations,
                       const double *grid_desc, const double *range_diffs, int32_t n_sets,
                       int32_t rd_stride, double *out_llh, double *out_cost, int64_t *out_index);

/* Least-squares fix over ALL S(S-1)/2 range differences (no reference
 * equivalent -- solveTDOA uses two of three measurements, freezes Z and stops after ten
 * half steps, processor.go:932-1020).  Levenberg-Marquardt in the local east/north/up
 * frame; dims = 2 keeps the start elevation (what 3 stations support), dims = 3 solves
 * it too (>= 4 stations).  init_llh[set][3] = start points (e.g. the tdoa_grid arg-min),
 * NULL: mean of the station coordinates.  out_rms[set] = sqrt(sum f^2 / P) in metres
 * (may be NULL); status[set] = 0 ok, 1 degenerate geometry. */
TDOA_API int tdoa_solve_ls(tdoa_engine *e, const double *stations_llh, int32_t n_stations,
                           const double *range_diffs, int32_t n_sets, int32_t rd_stride,
                           const double *init_llh, int32_t dims, double *out_llh, double *out_rms,
                           int32_t *status, int32_t *iters);

/* Per-capture signal quality (fast_analyzer.go:15-24 FastAnalysis, analyzer.go:18-42
 * SignalAnalysis): the numbers the reference's analyzers print and base their gain
 * recommendations on.  fast != 0 follows fast_analyzer.go (first 32768 samples of each
 * block, 8192-point Hanning spectrum, top 10 % vs bottom 40 %); fast == 0 follows
 * analyzer.go (whole blocks, dead-zone scan, 16384-point DC-corrected Blackman-Harris
 * spectrum, top 10 % vs lowest 50 %).  ref = blocks 1 + 3, tgt = block 2 of `station`. */
typedef struct {
    int64_t total_samples;
    double i_avg, q_avg, i_std, q_std;
    int32_t i_min, i_max, q_min, q_max;
    double snr_db, power_db;
    double dc_offset, iq_imbalance;            /* analyzer.go only */
    int32_t has_clipping, has_overload;
    int32_t has_dead_zones, has_noise;         /* analyzer.go only */
} tdoa_signal_quality;
TDOA_API int tdoa_analyze(tdoa_engine *e, int32_t station, int32_t fast, tdoa_signal_quality *ref,
                          tdoa_signal_quality *tgt);

/* What the reference prints while it works on a pair (shipped binary: "Initial signal
 * power", "Removed DC bias", "Normalized signal power", "Time domain correlation ... at
 * delay" before the sanity re-search; processor.go:474-497 prints the same quantities).
 * Values of window 0 of the last tdoa_xcorr(kind) call; signals in station order,
 * first_corr in pair order. */
typedef struct {
    double power0;        /* calculateSignalPower of the raw signal (processor.go:322)   */
    double dc_re, dc_im;  /* removeDCBias mean (processor.go:309)                         */
    double power1;        /* power entering normalizeSignal (processor.go:336)            */
    int32_t branch;       /* 0 strong FM / standard, 1 moderate (envelope), 2 weak        */
    int32_t reserved;
    int64_t n;            /* samples of the signal                                        */
} tdoa_signal_info;
TDOA_API int tdoa_xcorr_info(tdoa_engine *e, int32_t kind, tdoa_signal_info *signals, double *first_corr);

/* Introspection for the benchmark: kernels launched and device time of the last
 * tdoa_xcorr call, per stage (CUDA events on the engine's stream). */
typedef struct {
    int64_t launches_total;   /* kernels launched by this engine since creation   */
    int64_t launches_last;    /* ... by the last xcorr / cross_correlate call     */
    float ms_preprocess;      /* stats + demod + box-car of the last call         */
    float ms_fft;             /* segmented FFT cross-spectrum + inverse           */
    float ms_exact;           /* exact time-domain re-evaluation + peak select    */
    float ms_total;
    int64_t fft_launches;     /* launches of the segmented FFT kernel, last call  */
    float ms_fft_seg;         /* device time of those launches                    */
    int64_t fft_pair_samples; /* template samples they covered                    */
    int64_t brute_pairs;      /* pair-windows that fell back to every-lag search  */
    /* per-kernel device time of the last call (CUDA events around each launch) and the
     * units each kernel covered -- what the benchmark's per-kernel roofline is made of */
    float ms_demod;           /* fused unpack + power + FM discriminator launches */
    float ms_boxcar;          /* small box-car (DC subtract + low-pass + power)   */
    float ms_cand;            /* exact re-evaluation of the candidate lags        */
    float reserved0;
    int64_t demod_launches, demod_samples;
    int64_t boxcar_launches, boxcar_samples;
    int64_t cand_launches, cand_pair_samples;
} tdoa_stats;
TDOA_API int tdoa_get_stats(tdoa_engine *e, tdoa_stats *out);

/* Device self-tests of arithmetic shortcuts the kernels rely on.  which = 0: the
 * box-car's constant-divisor division equals a correctly rounded f32 divide for every
 * float input (exhaustive, ~10 ms); *mismatches = 0 means proven.  which = 1: the
 * production FM discriminator (fast arctangent + full-accuracy fall-back where the f32
 * rounding is not decided) against the reference statement of it over all 2^32
 * (previous, current) byte quads; tdoa_last_error then carries the counts and the
 * differing quads.  which = 2: *mismatches = how many of the 2^32 quads take the
 * fall-back.  which = 3: as 1 for the full-accuracy arctangent on every sample
 * (fast_demod = 2). */
TDOA_API int tdoa_selftest(tdoa_engine *e, int32_t which, int64_t *mismatches);

/* Raw CUDA stream of the engine (cudaStream_t as void*), for callers that time or
 * chain work on it. */
TDOA_API void *tdoa_stream(tdoa_engine *e);
/* Run on a caller-owned stream instead (cudaStream_t as void*; NULL: a fresh private
 * stream).  The benchmark uses this so that its CUDA events bracket the engine's work. */
TDOA_API int tdoa_set_stream(tdoa_engine *e, void *stream);
/* Block until everything queued on the engine's stream is done. */
TDOA_API int tdoa_synchronize(tdoa_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* TDOA_B200_H */

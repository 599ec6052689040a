/* talfe.h — C ABI of the B200-native log-mel acoustic front end ("talfe" = TAL front end).
 *
 * Drop-in boundary for ONE path of calclavia/tal-asrd: the waveform -> (T, n_mels) log-mel
 * feature extraction the reference performs in
 *     /root/reference/tal/asr/models.py:15-53   class LogMelSpec  (__init__ :22-33, forward :36-53)
 * called from
 *     /root/reference/tal/asr/models.py:154-162 ASRModel.extract_features
 *     /root/reference/tal/asr/models.py:430-438 SDModel.extract_features
 * The reference is pure Python and has no FFI of its own for this path; the entry points below
 * are what a ctypes binding placed inside LogMelSpec.forward would call.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is either a HOST pointer or a DEVICE pointer as
 *     stated per argument; sizes are int64_t / size_t.
 *   - the caller owns every buffer including the workspace; the library never allocates, frees
 *     or synchronises on the hot path (talfe_plan_create / _destroy are the only allocating calls).
 *   - all work is enqueued on the CUDA stream passed in (cudaStream_t as void*); asynchronous
 *     CUDA errors surface at the caller's next synchronisation.
 *   - return value 0 on success, negative talfe_status on error; no exceptions cross the ABI.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TALFE_H_
#define TALFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TALFE_VERSION 105 /* major * 100 + minor */

typedef enum talfe_status {
    TALFE_OK = 0,
    TALFE_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unknown enum) */
    TALFE_ERR_TOO_SHORT = -2,   /* n_samples <= 200: reflect padding undefined (reference raises RuntimeError) */
    TALFE_ERR_UNSUPPORTED = -3, /* configuration outside what the sm_100a kernels implement */
    TALFE_ERR_WORKSPACE = -4,   /* workspace missing or smaller than talfe_workspace_bytes() */
    TALFE_ERR_CUDA = -5,        /* a CUDA runtime call failed (talfe_last_cuda_error() has the code) */
    TALFE_ERR_NCCL = -6         /* NCCL unavailable or a collective failed */
} talfe_status;

/* Waveform element types.  F32 is what the reference feeds (models.py:36-45); F16 is what its
 * decoding paths produce with .half() (tal/asr/system.py:92,285); I16 is the on-disk PCM format
 * (tal/utils/audio.py:12-13), scaled by 1/32768 exactly as torchaudio.load normalises
 * (tal/asr/data/util.py:43). */
typedef enum talfe_dtype { TALFE_F32 = 0, TALFE_F16 = 1, TALFE_I16 = 2 } talfe_dtype;

/* Normalisation applied after log(mel + eps).
 *   BATCH_MEAN is the reference: one scalar mean over every element of the [B, T, M] result,
 *   padding frames included (models.py:52 `mel -= mel.mean()`).  The others are extensions. */
typedef enum talfe_norm {
    TALFE_NORM_NONE = 0,
    TALFE_NORM_BATCH_MEAN = 1,
    TALFE_NORM_ROW_MEAN = 2,        /* per-row scalar mean over that row's valid frames            */
    TALFE_NORM_ROW_MEL_MEAN = 3,    /* per-row, per-mel mean (CMN)                                   */
    TALFE_NORM_ROW_MEL_MEANVAR = 4  /* per-row, per-mel mean and variance (CMVN)                     */
} talfe_norm;

/* Output layout: TM = [B, T, M] (what LogMelSpec.forward returns, models.py:48);
 * MT = [B, M, T] (what encode_features permutes to next, models.py:167,443). */
typedef enum talfe_layout { TALFE_LAYOUT_TM = 0, TALFE_LAYOUT_MT = 1 } talfe_layout;

typedef struct talfe_plan talfe_plan; /* opaque: device-resident tables for one (device, n_mels) */

/* Statistics block layout (doubles): [0] = element count, [1] = sum, [2] = sum of squares,
 * then per-mel sums [3 .. 3+M) and per-mel sums of squares [3+M .. 3+2M).  One block per row for
 * the ROW_* modes, one block in total for NONE / BATCH_MEAN.
 * The per-mel entries are WRITTEN ONLY by the ROW_MEL_* modes (they cost one extra pass over the features);
 * NONE, BATCH_MEAN and ROW_MEAN fill entries [0..2] and leave the rest untouched. */
#define TALFE_STATS_DOUBLES(n_mels) (3 + 2 * (n_mels))

/* One unit of work: `batch` rows, frames [frame0, frame0 + n_frames) of each row. */
typedef struct talfe_job {
    const void* wave;        /* DEVICE: batch rows of samples, row r at wave + r * row_stride elements        */
    int32_t wave_dtype;      /* talfe_dtype                                                                  */
    int32_t norm;            /* talfe_norm                                                                   */
    int64_t batch;           /* B >= 1                                                                       */
    int64_t row_stride;      /* elements between consecutive rows (>= buf_len)                               */
    int64_t buf_len;         /* samples present per row in `wave`                                            */
    int64_t origin;          /* index, within the full row/episode, of wave[r][0] (0 unless streaming chunks)*/
    int64_t total_len;       /* full length L of every row: frame count 1 + L/160 and the reflection point;  */
                             /*   rows shorter than L are expected zero-padded by the caller (reference       */
                             /*   collater semantics, tal/asr/data/aligned.py:246-270)                        */
    const int64_t* lens;     /* DEVICE or NULL: per-row true lengths -> "each row as if alone" semantics      */
                             /*   (own frame count, reflection at its own end, frames beyond written as 0)    */
    int64_t frame0;          /* first frame to produce                                                        */
    int64_t n_frames;        /* frames to produce per row                                                     */
    float* out;              /* DEVICE: [B, n_frames, M] or [B, M, n_frames] float32                          */
    int64_t out_row_stride;  /* floats between rows of out; 0 = dense (n_frames * M)                          */
    int32_t out_layout;      /* talfe_layout                                                                  */
    int32_t accumulate_stats;/* nonzero: add this call's sums into `stats` instead of overwriting (streaming) */
    float eps;               /* additive floor inside the log (reference 1e-6, models.py:22)                  */
    int32_t defer_normalise; /* nonzero: compute `stats` but leave `out` un-normalised (apply later)          */
    double* stats;           /* DEVICE or NULL: TALFE_STATS_DOUBLES(M) doubles (x B for ROW_* modes)          */
    void* workspace;         /* DEVICE: at least talfe_workspace_bytes() bytes                                */
    size_t workspace_bytes;
    void* stream;            /* cudaStream_t                                                                  */
    /* --- extensions beyond LogMelSpec.forward (all NULL / 0 for the reference behaviour) ---                */
    const int64_t* out_offsets; /* DEVICE or NULL: packed ragged output [sum T_i, M]: row r starts at frame    */
                             /*   out_offsets[r]; requires lens and TM layout; no padding frames are written   */
    const int32_t* freq_bands;  /* DEVICE [B, n_bands, 2] (first mel, end mel): SpecAugment frequency masks, set */
                             /*   to 0 after normalisation like freq_mask(), tal/asr/models.py:531-548          */
    const int32_t* time_bands;  /* DEVICE [B, n_bands, 2] (first frame, end frame): time_mask(), models.py:550-566 */
    int32_t n_bands;         /* bands per row and axis, 0..16 (empty bands: first == end)                       */
    const double* given_stats;  /* DEVICE or NULL: ONE statistics block (TALFE_STATS_DOUBLES(M) doubles, e.g. the      */
                             /*   all-reduced sums of a corpus pass) to normalise EVERY row of this call with,       */
                             /*   according to `norm`, instead of the call's own statistics: dataset-level CMVN in   */
                             /*   the transform kernel itself, no sweep over the features (values identical to       */
                             /*   norm = NONE followed by talfe_apply_stats with that block).  Requires norm != NONE,*/
                             /*   n_bands == 0, defer_normalise == 0; `stats` is not written.                        */
    int32_t lens_are_padding_hint; /* nonzero (with `lens`): the rows keep the REFERENCE semantics (every row has the frames   */
                             /*   of total_len, reflection at the padded end, padding frames count in the mean) and lens[r]   */
                             /*   is the caller's guarantee that row r is zero from sample lens[r] on (what the collaters     */
                             /*   produce, tal/asr/data/aligned.py:246-270): frames that cannot see a sample below lens[r]    */
                             /*   are the constant log(eps) and are filled, not computed.  norm NONE / BATCH_MEAN only.       */
} talfe_job;

int talfe_version(void);
/* sizeof(talfe_job) as compiled into the library: a binding checks it against its own struct definition at load
 * time, so that a stale binary or a drifted field list fails loudly instead of passing wrong pointers. */
size_t talfe_job_size(void);
const char* talfe_strerror(int status);
int talfe_last_cuda_error(void); /* cudaError_t of the most recent TALFE_ERR_CUDA on this thread */

/* 1 + n_samples / 160, or TALFE_ERR_TOO_SHORT when n_samples <= 200
 * (torch.stft(center=True, pad_mode="reflect") inside MelSpectrogram, models.py:24-32;
 *  models.py:91 "15999 frames => 100 frames"). */
int64_t talfe_num_frames(int64_t n_samples);

/* Builds the device tables.  window_host: 400 floats or NULL (periodic Hann computed here);
 * fb_host: [201, n_mels] row-major floats or NULL (HTK triangular filters 0..8000 Hz computed here).
 * Passing the reference module's own buffers makes the tables bit-identical to the reference's.
 * n_mels in 1..80; fb rows 0 and 200 must be all zero and every filter's support contiguous,
 * otherwise TALFE_ERR_UNSUPPORTED. */
int talfe_plan_create(talfe_plan** plan, int device, int n_mels, const float* window_host, const float* fb_host);
/* The same for any frame geometry: LogMelSpec(sr) derives n_fft = win = int(0.025 sr) and hop = int(0.010 sr)
 * (tal/asr/models.py:24-32).  window_host: n_fft floats, fb_host: [n_fft / 2 + 1, n_mels] row-major, both required.
 * n_fft 400 / hop 160 is talfe_plan_create (the specialised kernels); anything else runs the generic kernel
 * (any filterbank, n_fft <= 1280, n_mels <= 80; talfe_stream_episode is not available for such plans). */
int talfe_plan_create_ex(talfe_plan** plan, int device, int n_fft, int hop, int n_mels, const float* window_host,
                         const float* fb_host);
int talfe_plan_geometry(const talfe_plan* plan, int* n_fft, int* hop);
/* 1 + (n_samples + 2 (n_fft / 2) - n_fft) / hop for the plan's geometry, or TALFE_ERR_TOO_SHORT when n_samples <= n_fft / 2. */
int64_t talfe_plan_num_frames(const talfe_plan* plan, int64_t n_samples);
void talfe_plan_destroy(talfe_plan* plan);
int talfe_plan_n_mels(const talfe_plan* plan);
/* Kernel launches one talfe_logmel_forward of this shape makes with this plan (for honest launch accounting):
 * 1 when the scalar-mean subtraction runs inside the transform kernel (small calls), 2 when a separate sweep follows. */
int talfe_launches_per_forward(const talfe_plan* plan, int64_t batch, int64_t n_samples);

/* Loader side (tal/asr/data/util.py:44-48): torchaudio.transforms.Resample(orig_freq, 16000) for files that are not 16 kHz.
 * orig / new_: the two rates divided by their gcd; kernel_dev: DEVICE filter table [new_][2 * width + orig] (built by the
 * binding with torchaudio's windowed-sinc formula); wave: DEVICE [batch, n_samples] (f32 / f16, or i16 PCM scaled by 1/32768
 * like torchaudio.load); out: DEVICE float [batch, out_len], out_len = ceil(new_ * n_samples / orig).  batch <= 65535. */
int talfe_resample(const void* wave, int wave_dtype, int64_t batch, int64_t n_samples, int64_t row_stride, int orig, int new_,
                   int width, const float* kernel_dev, float* out, int64_t out_len, int64_t out_row_stride, void* stream);

/* Bytes of workspace talfe_run needs for `batch` rows of `n_frames` frames. */
size_t talfe_workspace_bytes(const talfe_plan* plan, int64_t batch, int64_t n_frames);

/* The hot path. */
int talfe_run(const talfe_plan* plan, const talfe_job* job);

/* Convenience wrapper with the exact semantics of LogMelSpec.forward (models.py:36-53):
 * wave[B, n_samples] (row stride row_stride) -> out[B, 1 + n_samples/160, M], minus the batch scalar mean. */
int talfe_logmel_forward(const talfe_plan* plan, const void* wave, int wave_dtype, int64_t batch,
                         int64_t n_samples, int64_t row_stride, float* out, float eps, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Normalise features in place from a statistics block (e.g. after summing blocks of several chunks,
 * or after an all-reduce across ranks).  `norm` selects which entries of stats are used; for the
 * ROW_* modes stats holds one block per row.  valid_frames (DEVICE or NULL) limits each row. */
int talfe_apply_stats(const talfe_plan* plan, float* feats, int64_t batch, int64_t n_frames,
                      int64_t out_row_stride, int out_layout, int norm, const double* stats,
                      const int64_t* valid_frames, void* stream);

/* One long episode streamed from HOST memory (pinned for full copy speed) in chunks of `chunk_frames`
 * frames: chunk k+1 is copied host->device on the plan's side stream while chunk k is transformed on
 * `stream`; chunk boundaries sit on hop multiples with 200-sample halos, reflection happens only at the
 * true ends of the episode, statistics accumulate across chunks and the normalisation is applied once at
 * the end — the result equals the one-shot transform of the whole episode, which is what the reference
 * does (tal/baseline/reconcile.py:76-85 feeds a whole episode to LogMelSpec.forward).
 * out: DEVICE [T, M] with T = 1 + total_len/160.  stats: DEVICE, TALFE_STATS_DOUBLES(M) doubles.
 * staging: DEVICE, talfe_stream_staging_bytes() bytes.  workspace: talfe_workspace_bytes(plan, 1, chunk_frames).
 * defer_normalise != 0 leaves `out` un-normalised and `stats` holding the sums (dataset-level statistics).
 * One episode at a time per plan (the side stream and its events belong to the plan).
 * Buffer lifetime: when the call returns, every copy out of `wave_host` has COMPLETED (the call waits for its side
 * stream; the transforms may still be running on `stream`), so the caller may reuse or free the host buffer at once.
 * `wave_host` may also be a DEVICE pointer: the episode is then transformed where it lies, chunk by chunk, and
 * `staging` may be NULL. */
size_t talfe_stream_staging_bytes(int wave_dtype, int64_t chunk_frames);
int talfe_stream_episode(const talfe_plan* plan, const void* wave_host, int wave_dtype, int64_t total_len,
                         int64_t chunk_frames, float* out, int norm, int defer_normalise, double* stats, float eps,
                         void* staging, size_t staging_bytes, void* workspace, size_t workspace_bytes, void* stream);

/* Dataset-level statistics across ranks: in-place sum of `count` doubles over the communicator.
 * nccl_comm is an ncclComm_t created by the caller; NCCL is resolved at run time with dlopen
 * ("libnccl.so.2"), so the library itself has no link-time dependency on it. */
int talfe_allreduce_stats(double* stats_dev, int64_t count, void* nccl_comm, void* stream);

/* Diagnostic: measured FP32 FMA rate of `device` (lane-FMAs per second, packed FFMA2 chains on every SM), the
 * compute-side denominator of the roofline bench.py reports ("must be measured on the box").
 * Allocates, launches and synchronises: never call it on a hot path. */
int talfe_probe_fp32_fma_rate(int device, double* fma_per_second);

/* The collaters' zero padding (tal/asr/data/aligned.py:246-270) found on the device: lens[r] (DEVICE int64[batch]) = 1 + index
 * of row r's last non-zero sample, 0 for an all-zero row — the `lens` to pass with talfe_job::lens_are_padding_hint when the
 * caller has no lengths at hand (LogMelSpec.forward(audio) only gets the padded batch).  One backwards scan per row that stops
 * at the first non-zero sample: an unpadded batch costs one small launch. */
int talfe_detect_padding(const void* wave, int wave_dtype, int64_t batch, int64_t n_samples, int64_t row_stride, int64_t* lens,
                         void* stream);

/* Deterministic synthetic audio (bench / tests): fills wave[rows, n_samples] with episode
 * (first_episode + r), samples [start, start + n_samples); bit-identical to tal_asrd_b200/synth.py. */
int talfe_synth_fill(void* wave, int wave_dtype, int64_t rows, int64_t n_samples, int64_t row_stride,
                     uint64_t seed, int64_t first_episode, int64_t start, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TALFE_H_ */

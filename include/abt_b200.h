/* abt_b200.h -- C ABI of the B200-native Audio Barlow Twins hot path (libabt_b200.so).
 *
 * Drop-in boundary for the per-step data-parallel path of jonahanton/SSL_audio (reference tree
 * at /root/reference; citations below are relative to it).  The reference is pure Python and
 * has no FFI of its own: the interface these entry points replace is the set of Python calls
 * listed per function, and the binding a maintainer adds is the ctypes stub in INTEGRATION.md
 * (ssl_audio_b200/_lib.py is that stub).
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless marked [host];
 *   - every function returns 0 on success or a negative abt_status; abt_last_error() returns a
 *     thread-local, human-readable message for the last failure; no exception crosses the ABI;
 *   - functions taking abt_stream_t enqueue asynchronously on that CUDA stream (a cudaStream_t)
 *     and never allocate device memory: the caller provides outputs and workspaces;
 *   - there is no CPU fallback: on a non-sm_100 device the compute entry points return
 *     ABT_ERR_DEVICE.
 */
#ifndef ABT_B200_H
#define ABT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABT_VERSION 100 /* major*10000 + minor*100 + patch */

typedef void* abt_stream_t; /* cudaStream_t */

typedef enum {
    ABT_OK = 0,
    ABT_ERR_ARG = -1,    /* invalid argument (shape, alignment, null pointer) -> Python ValueError */
    ABT_ERR_CUDA = -2,   /* CUDA runtime / driver error                        -> Python RuntimeError */
    ABT_ERR_DEVICE = -3, /* current device is not sm_100                       -> Python RuntimeError */
    ABT_ERR_STATE = -4   /* object used in the wrong state                     -> Python RuntimeError */
} abt_status;

typedef enum { ABT_DTYPE_BF16 = 0, ABT_DTYPE_F16 = 1, ABT_DTYPE_F32 = 2 } abt_dtype;

int abt_version(void);
const char* abt_last_error(void);
/* 0 iff the current CUDA device has compute capability 10.x */
int abt_device_check(void);

/* ===================================================================================== *
 *  Frontend: wav -> power STFT -> mel -> log -> normalise
 *  replaces torchaudio.transforms.MelSpectrogram as constructed at datasets.py:39-48 plus
 *  `(mel + eps).log()` (datasets.py:115, old/data_manager/wav_to_lms.py:58-61) and the
 *  dataset z-score (datasets.py:118-119, :353-354).
 * ===================================================================================== */
typedef struct {
    int32_t sample_rate; /* 16000 */
    int32_t n_fft;       /* must be 1024 (the FFT is a fixed 32 x 32 decomposition; every configuration of the reference uses 1024) */
    int32_t win_length;  /* <= n_fft; centre-padded like torch.stft */
    int32_t hop_length;  /* 160 */
    int32_t n_mels;      /* 1..256 (reference default 64) */
    float f_min, f_max;  /* 60, 7800 */
    int32_t apply_norm;  /* 1: out = (log_mel - norm_mean) / norm_std */
    float norm_mean, norm_std;
} abt_mel_config;

typedef struct abt_logmel_plan abt_logmel_plan; /* device tables: window, sparse mel filterbank */

int abt_logmel_plan_create(const abt_mel_config* cfg, abt_logmel_plan** plan);
int abt_logmel_plan_destroy(abt_logmel_plan* plan);

/* Full log-mel ("mode F"): wav (n_clips, n_samples) fp32 row-major -> out (n_clips, n_mels, 1 + n_samples/hop). */
int abt_logmel_fwd(const abt_logmel_plan* plan, const float* wav, int n_clips, int n_samples, float* out, abt_stream_t stream);

/* Crop-first log-mel ("mode C"): only frames [frame_start[b], frame_start[b] + n_frames) of clip b
 * are computed (frames beyond the clip end are written as the normalised value of 0.0, mirroring
 * the right zero-pad at datasets.py:346-350 BEFORE normalisation).  Clip b is the n_samples-long
 * waveform at wav + b * wav_row_stride + wav_offset[b] (wav_offset may be NULL; it carries the
 * random unit crop of datasets.py:110-113, so reflect padding happens at the unit's edges as in
 * the reference).  Clip b is written to out_base + out_slot[b] * out_slot_stride (floats), laid
 * out (n_mels, n_frames); out_slot may be NULL (= identity).  This lets the frontend write
 * straight into the Mixup ring.  frame_start may be NULL (= 0). */
int abt_logmel_crop_fwd(const abt_logmel_plan* plan, const float* wav, int64_t wav_row_stride, const int32_t* wav_offset, int n_clips,
                        int n_samples, const int32_t* frame_start, int n_frames, float* out_base, const int32_t* out_slot,
                        int64_t out_slot_stride, abt_stream_t stream);

/* Crop-first input staging for HOST waveforms (the reference loads the whole clip in a DataLoader worker, datasets.py:98-116,
 * and ships it to the GPU at main.py:69).  Only the samples the cropped frames need -- abt_wav_span_len() per clip, e.g. 16 228
 * of 160 000 for a 96-frame crop of a 10 s clip -- are read.  With wav_on_host != 0, `wav` is a [host] pointer into MAPPED PINNED
 * memory and the kernel reads it in place over PCIe (no staging copy of the whole clip); otherwise it is a device pointer.
 * spans: (n_clips, span_len) fp32 device; span_origin: (n_clips) int32 device, first clip sample held by each span row.
 * Requires n_samples >= span_len.  abt_logmel_span_fwd is abt_logmel_crop_fwd reading those spans; it gives bit-identical
 * results (reflect padding is resolved in clip coordinates, and a span always contains the mirrored samples).
 * The gather runs as 4 blocks of 32 warps (one warp per clip): enough reads in flight for PCIe, and confined to 4 SMs it can run
 * beside the persistent tensor-core kernels when those leave them free (abt_set_reserved_sms). */
int abt_wav_span_len(const abt_logmel_plan* plan, int n_frames, int* span_len);
int abt_wav_span_gather(const abt_logmel_plan* plan, const float* wav, int wav_on_host, int64_t wav_row_stride, int n_clips, int n_samples,
                        const int32_t* frame_start, int n_frames, float* spans, int32_t* span_origin, abt_stream_t stream);
int abt_logmel_span_fwd(const abt_logmel_plan* plan, const float* spans, const int32_t* span_origin, int n_clips, int n_samples,
                        const int32_t* frame_start, int n_frames, float* out_base, const int32_t* out_slot, int64_t out_slot_stride,
                        abt_stream_t stream);

/* Crop / right-pad a precomputed log-mel and z-score it: datasets.py:342-354.
 * lms (n_clips, n_mels, t_full) -> out slots as above, (n_mels, n_frames) each. */
int abt_lms_crop_norm(const float* lms, int n_clips, int n_mels, int t_full, const int32_t* frame_start, int n_frames, int apply_norm,
                      float norm_mean, float norm_std, float* out_base, const int32_t* out_slot, int64_t out_slot_stride,
                      abt_stream_t stream);

/* ===================================================================================== *
 *  Views: Mixup -> [MixGaussianNoise] -> RandomResizeCrop -> RandomLinearFader, one launch for all clips and views
 *  replaces MixupBYOLA.forward (augmentations.py:103-117), log_mixup_exp (:81-85), MixGaussianNoise.forward (:133-140),
 *  RandomResizeCrop.forward (:40-55) and RandomLinearFader.forward (:69-74) as sequenced by
 *  AudioPairTransform.forward (utils/transforms.py:49-58).
 * ===================================================================================== */
typedef struct {
    int32_t z_kind; /* 0: no mixup partner, 1: partner is bank slot z_index, 2: partner is clip z_index of this batch */
    int32_t z_index;
    float w_x, w_z;           /* fp32 weights of exp(x), exp(z): float32(1-alpha), float32(1-(1-alpha)) */
    int32_t i, j, h, w;       /* crop box on the virtual canvas (RandomResizeCrop.get_params) */
    float head, tail;         /* fader end points */
    int32_t flags;            /* bit0: mixup on, bit1: resize-crop on, bit2: fader on, bit3: Gaussian-noise mix on */
    int32_t out_index;        /* which output tensor of `outs` this view is written to */
    float g_lambda;           /* MixGaussianNoise: lambd = ratio * np.random.rand() (augmentations.py:136), as float32 */
    float g_keep;             /* float32(1 - lambd) */
    int32_t reserved[2];
} abt_view_params;            /* 64 bytes */

typedef struct {
    int32_t n_clips, n_views; /* params has n_clips * n_views entries, clip-major */
    int32_t in_h, in_w;       /* 64, 96 */
    int32_t canvas_h, canvas_w; /* int(in_h * vcs[0]), int(in_w * vcs[1]) */
    int32_t out_h, out_w;     /* resize target (n_mels, crop_frames) or local_crops_size */
    int32_t param_stride;     /* params of (clip, view) at params[clip * param_stride + view_offset + view]; 0 = n_views */
    int32_t view_offset;
    const float* x;           /* clip b at x + x_slot[b] * x_slot_stride, (in_h, in_w) fp32; x_slot may be NULL */
    const int32_t* x_slot;
    int64_t x_slot_stride;
    const float* bank;        /* bank slot s at bank + s * bank_slot_stride */
    int64_t bank_slot_stride;
    const abt_view_params* params; /* device */
    float* outs[8];           /* output tensor of view k of this launch: (n_clips, 1, out_h, out_w) contiguous */
    const float* noise;       /* MixGaussianNoise: STANDARD normal draws, (n_clips, noise_views, in_h, in_w) fp32, view v of this launch reads
                                 plane view_offset + v; the kernel scales them by g_lambda (torch.normal(0, lambd, shape), augmentations.py:137).
                                 Required when a view has flag bit3, else may be NULL */
    int32_t noise_views;
} abt_views_args;

int abt_views_fwd(const abt_views_args* args, abt_stream_t stream);

/* copy clips into bank slots: bank[slot[b]] = x[b]  (MixupBYOLA's `memory_bank + [x]`, augmentations.py:115) */
int abt_bank_push(const float* x, int64_t x_stride, int n_clips, int clip_elems, float* bank, int64_t bank_slot_stride,
                  const int32_t* slot, abt_stream_t stream);

/* NormalizeBatch (augmentations.py:217-232; applied to every crop at main.py:62-66 under --post_norm): x (n_batch, n_channels, hw)
 * fp32 -> out = (x - mean_c) / clamp(std_c, eps), mean and UNBIASED std per channel over the batch and both spatial axes.
 * workspace: abt_normalize_batch_workspace_bytes() bytes of device memory. */
int abt_normalize_batch_workspace_bytes(int n_channels, size_t* bytes);
int abt_normalize_batch(const float* x, int n_batch, int n_channels, int hw, float* out, void* workspace, abt_stream_t stream);

/* Dataset statistics (datasets.py:362-376, `calculate_norm_stats`: lms_vectors.mean(), lms_vectors.std()): mean and UNBIASED
 * standard deviation of x (n_batch, elems) as two DEVICE doubles.  workspace: abt_normalize_batch_workspace_bytes(1). */
int abt_mean_std(const float* x, int n_batch, int elems, double* out2, void* workspace, abt_stream_t stream);

/* RunningNorm (augmentations.py:187-210; --pre_norm, main.py:272-277): x (n_batch, elems) fp32, the samples are processed IN ORDER with
 * the reference's running statistics (mu += (mean(x_b) - mu) / n with n the count before the update, likewise for mean((x_b - mu)^2));
 * updates stop after max_update samples.  state3: device, 3 doubles (count, mu, s2), zero-initialised by the caller and carried from
 * call to call.  out = (x_b - mu) / clamp(sqrt(s2), eps).  workspace: abt_running_norm_workspace_bytes(n_batch) bytes. */
int abt_running_norm_workspace_bytes(int n_batch, size_t* bytes);
int abt_running_norm(const float* x, int n_batch, int elems, long long max_update, double* state3, float* out, void* workspace,
                     abt_stream_t stream);

/* ===================================================================================== *
 *  Host planner [host]: replays the reference's RNG draw order (numpy legacy MT19937 global
 *  state + CPython `random` MT19937) for a whole batch and emits the parameter table.
 *  Bit-exact against np.random.* / random.randint as called at augmentations.py:34-37,70,105,108,136
 *  and datasets.py:89,112,344.
 * ===================================================================================== */
typedef struct abt_planner abt_planner;

typedef struct {
    int32_t mixup, rrc, rlf;      /* args.mixup / args.RRC / args.RLF */
    double mixup_ratio_d;         /* 0.2 (python float) */
    int32_t n_memory;             /* 2048: length of the virtual FIFO (MixupBYOLA.n) */
    int32_t ring_slots;           /* physical ring size; must be >= n_memory + largest batch */
    int32_t n_global;             /* global views per clip (2 for AudioPairTransform, 1 for a bare module) */
    int32_t in_h, in_w;           /* 64, 96 */
    int32_t canvas_h, canvas_w;   /* 64, 144 */
    double freq_scale[2], time_scale[2]; /* (0.6, 1.5) */
    int32_t n_local;              /* local crops per clip */
    int32_t local_h, local_w;     /* local_crops_size */
    double local_scale[2];        /* (0.05, 0.6) */
    double fader_gain;            /* 1.0 */
    int32_t gnoise;               /* args.Gnoise: MixGaussianNoise between Mixup and RandomResizeCrop (utils/transforms.py:20-21) */
    double gnoise_ratio_d;        /* 0.2 */
} abt_plan_config;

int abt_planner_create(const abt_plan_config* cfg, abt_planner** p);
int abt_planner_destroy(abt_planner* p);
/* numpy legacy state: key[624], pos in [0,624]  (np.random.get_state()[1:3]) */
int abt_planner_set_numpy_state(abt_planner* p, const uint32_t* key624, int pos);
int abt_planner_get_numpy_state(const abt_planner* p, uint32_t* key624, int* pos);
/* CPython random state: the 625-tuple of random.getstate()[1] (624 words + index) */
int abt_planner_set_pyrandom_state(abt_planner* p, const uint32_t* key624, int pos);
int abt_planner_get_pyrandom_state(const abt_planner* p, uint32_t* key624, int* pos);
/* bank bookkeeping (virtual FIFO of clip uids, mirrors MixupBYOLA.memory_bank) */
int abt_planner_bank_len(const abt_planner* p);
int abt_planner_bank_reset(abt_planner* p);
/* Plan one batch.  time_crop_range > 0 draws `np.random.randint(time_crop_range)` per clip before its
 * views (datasets.py:344); wav_crop_range > 0 draws `random.randint(0, wav_crop_range)` (datasets.py:112).
 * Outputs [host]: starts[n_clips] (-1 if not drawn), wav_starts[n_clips], params[n_clips*(2+n_local)],
 * slots[n_clips] = bank ring slot each clip must be stored in (uid % ring_slots). */
int abt_planner_plan_batch(abt_planner* p, int n_clips, int time_crop_range, int wav_crop_range, int32_t* starts,
                           int32_t* wav_starts, abt_view_params* params, int32_t* slots);

/* Same, written into ONE caller-provided [host] buffer (e.g. pinned memory, uploaded with a single copy):
 *   [abt_view_params x n_clips*(n_global+n_local)] [starts int32 x n_clips] [wav_starts int32 x n_clips] [slots int32 x n_clips]
 * each section starting on a 16-byte boundary; abt_planner_packed_bytes() gives the size and section offsets. */
int abt_planner_packed_bytes(const abt_planner* p, int n_clips, size_t* total, size_t* off_starts, size_t* off_wav_starts,
                             size_t* off_slots);
int abt_planner_plan_batch_packed(abt_planner* p, int n_clips, int time_crop_range, int wav_crop_range, void* out, size_t out_bytes);

/* Same, against the interpreter's GLOBAL generators in place (one call per batch): np_state -> numpy's legacy
 * `struct { uint32_t key[624]; int pos; }` (np.random.mtrand._rand._bit_generator.ctypes.state_address), py_index / py_key -> the
 * index and the 624 state words inside CPython's `random._inst` object.  Either may be NULL when that generator cannot be drawn from. */
int abt_planner_plan_batch_global(abt_planner* p, int n_clips, int time_crop_range, int wav_crop_range, void* np_state, int32_t* py_index,
                                  uint32_t* py_key, void* out, size_t out_bytes);

/* ===================================================================================== *
 *  Barlow Twins objective forward + backward
 *  replaces BarlowTwinsLoss.forward_loss (utils/loss.py:15-30: BatchNorm1d(affine=False) on
 *  both views, c = bn(z1).T @ bn(z2) / N, on/off-diagonal loss) AND its autograd backward,
 *  plus the BatchNorm running-stat side effect (utils/loss.py:13,17).
 * ===================================================================================== */
typedef struct {
    const void* z1;       /* (n_rows, n_dims) row-major, dtype below */
    const void* z2;
    int32_t dtype;        /* abt_dtype; the tensor cores consume bf16: f16/f32 inputs are rounded to bf16 */
    int32_t n_rows;       /* N >= 2 */
    int32_t n_dims;       /* D, multiple of 64 */
    float alpha, lambda;  /* cfg.alpha, cfg.lmbda */
    int32_t hsic;         /* cfg.HSIC */
    float eps;            /* BatchNorm eps (1e-5) */
    float momentum;       /* BatchNorm momentum (0.1) */
    float grad_scale;     /* gradients are multiplied by this (1.0 = d loss) */
    int32_t need_grad_mask; /* bit0: dz1, bit1: dz2 */
    float* loss_out;      /* device scalar */
    void* dz1;            /* (n_rows, n_dims) same dtype as z, or NULL */
    void* dz2;
    float* running_mean;  /* (n_dims) BatchNorm buffers updated z1 then z2, or NULL */
    float* running_var;
    void* workspace;      /* abt_bt_workspace_bytes() bytes, 256-byte aligned */
    size_t workspace_bytes;
} abt_bt_args;

int abt_bt_workspace_bytes(int n_rows, int n_dims, int dtype, size_t* bytes);
int abt_bt_loss_fwd_bwd(const abt_bt_args* args, abt_stream_t stream);

/* backward of the loss scalar: the gradients were produced by the forward call; autograd's `grad_output` (a device scalar,
 * e.g. the GradScaler factor) multiplies them in place.  a, b: (n_elems) of `dtype`, either may be NULL. */
int abt_scale_inplace(void* a, void* b, size_t n_elems, int dtype, const float* scale_dev, abt_stream_t stream);

/* Row-block form for the multi-GPU objective (replaces the D x D `torch.distributed.all_reduce(c)` of
 * utils/loss.py:20-21 with: all-gather of the embeddings -> this call -> all-to-all of the gradients -> 3-double
 * all-reduce).  zg1 / zg2 are the rank-ordered GLOBAL batches (n_rows = N_g); this rank owns dimensions
 * [row_begin, row_begin + row_count): it computes those rows of C and of C^T, the off-diagonal loss of its rows
 * of C, and d loss / d z for ALL n_rows samples restricted to its dimensions (batch-norm backward is per column).
 * Statistics (and the running-stat update) are those of the global batch and identical on every rank. */
typedef struct {
    const void* zg1;      /* (n_rows, n_dims) row-major gathered embeddings */
    const void* zg2;
    int32_t dtype, n_rows, n_dims;
    int32_t row_begin, row_count;   /* multiples of 8 */
    float alpha, lambda;
    int32_t hsic;
    float eps, momentum, grad_scale;
    int32_t need_grad_mask;
    double* loss_parts;   /* device, 3 doubles: sum_{i in rows, j != i} C_ij^2, sum_{i in rows, j != i} C_ij, sum_{all i} (C_ii - 1)^2 */
    void* dzr1;           /* (n_rows, row_count) compact, same dtype as the inputs, or NULL */
    void* dzr2;
    float* running_mean;  /* (n_dims) or NULL */
    float* running_var;
    void* workspace;      /* abt_bt_rows_workspace_bytes() bytes, 256-byte aligned */
    size_t workspace_bytes;
} abt_bt_rows_args;

int abt_bt_rows_workspace_bytes(int n_rows, int n_dims, int row_count, int dtype, size_t* bytes);
int abt_bt_loss_rows_fwd_bwd(const abt_bt_rows_args* args, abt_stream_t stream);

/* Multi-GPU objective as BASELINE.json:north_star words it (one process per GPU; the collectives themselves are NCCL calls made by
 * the host between these entry points, ssl_audio_b200/dist.py):
 *   1. abt_bt_dist_stats_local   local rows -> 7 numbers per column (shifted sums + shifts) at workspace + layout.pack_local
 *      [all-gather of the packs (7 D floats per rank) into workspace + layout.pack_all]
 *   2. abt_bt_dist_normalize     global statistics (combined in double), BatchNorm running-stat update, on-diagonal loss, and the
 *                                fp16 STANDARDISED local rows written into this rank's slot of the gather buffers (layout.zh1 / zh2)
 *      [in-place all-gather of the standardised embeddings: (world * n_local, n_dims) fp16 per view]
 *   3. abt_bt_dist_rows_fwd_bwd  rows [row_begin, row_begin + row_count) of C and C^T on the tensor cores from the gathered
 *                                standardised embeddings, their loss terms, and d loss / d z of ALL samples for those dimensions
 *      [all-to-all of the gradient slices, 2-double all-reduce of the loss]
 * The three calls of one step share ONE workspace (abt_bt_dist_layout_query gives its size and the offsets the host needs). */
typedef struct {
    size_t total_bytes;
    size_t zh1, zh2;          /* (world * n_local, n_dims) fp16 gather buffers; rank r owns rows [r * n_local, (r + 1) * n_local) */
    size_t pack_local;        /* pack_floats floats */
    size_t pack_all;          /* world * pack_floats floats */
    size_t pack_floats;       /* 7 * n_dims */
} abt_bt_dist_layout;

typedef struct {
    int32_t dtype;            /* dtype of dzr1 / dzr2 (the embeddings' dtype) */
    int32_t n_local, world, n_dims;
    int32_t row_begin, row_count;
    float alpha, lambda;
    int32_t hsic;
    float grad_scale;
    int32_t need_grad_mask;   /* gradients wanted from the whole step */
    int32_t phase;            /* 0: everything in one call.  1: statistics hand-over, CORR and the dz1 gradient pass; 2: only the dz2
                               * gradient pass (after a phase-1 call) -- lets the host overlap the all-to-all of dz1 with the dz2 GEMM */
    double* loss_parts;       /* device, 3 doubles as in abt_bt_rows_args (written by phases 0 and 1) */
    void* dzr1;               /* (world * n_local, row_count) compact */
    void* dzr2;
    void* workspace;
    size_t workspace_bytes;
} abt_bt_dist_args;

int abt_bt_dist_layout_query(int n_local, int world, int n_dims, int row_count, abt_bt_dist_layout* out);
int abt_bt_dist_stats_local(const void* z1, const void* z2, int dtype, int n_local, int world, int n_dims, int row_count, void* workspace,
                            abt_stream_t stream);
int abt_bt_dist_normalize(const void* z1, const void* z2, int dtype, int n_local, int world, int rank, int n_dims, int row_count, float eps,
                          float momentum, float* running_mean, float* running_var, void* workspace, abt_stream_t stream);
int abt_bt_dist_rows_fwd_bwd(const abt_bt_dist_args* args, abt_stream_t stream);

/* The same step issued natively: ONE host call per training step.  The collectives run on NCCL -- the libnccl.so.2 the process has
 * already loaded (PyTorch's), resolved with dlopen at run time -- over a private communicator and a private high-priority
 * communication stream; the all-to-all of the dz1 slices overlaps the dz2 GEMM, and `overlap_cb` (optional) is invoked right after
 * the embedding all-gather has been launched so that the caller can enqueue independent work (the next batch's frontend) that runs
 * while the embeddings cross NVLink.  Create the communicator once per process group:
 *     rank 0: abt_comm_unique_id(id) -> broadcast the 256 bytes (two NCCL ids: large gathers / small exchanges) -> every rank:
 *     abt_comm_create(world, rank, id, &comm).
 * Default schedule ("exchange"): view 2 is gathered first; CORR (whose A operand, the view-1 columns of this rank's dimensions, arrives
 * by a small all-to-all) and the dz1 GEMM run while view 1 is still in flight; the transposed block C[:, rows] arrives by an
 * all-to-all of C blocks instead of a second CORR pass, so every rank executes 6 N D^2 FLOP as on one GPU. */
typedef struct abt_comm abt_comm;
int abt_comm_unique_id(void* id256);
int abt_comm_create(int world, int rank, const void* id256, abt_comm** comm);
int abt_comm_destroy(abt_comm* comm);

typedef void (*abt_overlap_cb)(void* user);
typedef struct {
    const void* z1;           /* this rank's (n_local, n_dims) embeddings */
    const void* z2;
    int32_t dtype, n_local, n_dims;   /* n_dims divisible by the world size, n_dims / world a multiple of 8 */
    float alpha, lambda;
    int32_t hsic;
    float eps, momentum, grad_scale;
    int32_t need_grad_mask;
    float* loss_out;          /* device scalar: the GLOBAL-batch loss (identical on every rank) */
    void* dz1;                /* (n_local, n_dims), dtype of the inputs, or NULL */
    void* dz2;
    float* running_mean;      /* BatchNorm buffers (n_dims), updated with the global-batch statistics, or NULL */
    float* running_var;
    void* workspace;          /* abt_bt_dist_step_workspace_bytes() bytes, 256-byte aligned */
    size_t workspace_bytes;
    abt_overlap_cb overlap_cb;
    void* overlap_user;
    /* Optional copy-engine embedding exchange.  Instead of two NCCL all-gathers (SM kernels that compete with the tensor-core kernels
     * and the frontend), every rank PULLS the other ranks' standardised rows out of peer-mapped exchange buffers with cudaMemcpyAsync:
     * the copies run on the copy engines, NVLink is driven without a single SM.  exchange_peers: [host] array of `world` device
     * pointers, entry q = rank q's exchange buffer mapped into this process (torch symmetric memory, cudaIpc*, ...), entry `rank` = this
     * rank's own; every buffer holds abt_bt_dist_exchange_bytes() bytes and is zero-filled once before its first use (then a
     * cross-rank barrier).  exchange_epoch: 1, 2, 3, ... -- the number of this call among the calls that used THESE buffers, the same
     * on every rank (the buffers are double-buffered by its parity; one flag barrier per step keeps the ranks in step).
     * exchange_peers = NULL: NCCL all-gathers. */
    void* const* exchange_peers;
    size_t exchange_bytes;
    uint32_t exchange_epoch;
} abt_bt_dist_step_args;

int abt_bt_dist_step_workspace_bytes(int n_local, int world, int n_dims, size_t* bytes);
int abt_bt_dist_exchange_bytes(int n_local, int world, int n_dims, size_t* bytes);
int abt_bt_dist_step(const abt_bt_dist_step_args* args, abt_comm* comm, abt_stream_t stream);

/* ===================================================================================== *
 *  Step-adjacent optimiser math as multi-tensor kernels
 *  replaces LARS.step (utils/utils.py:162-189: per-parameter Python loop, two torch.norm and
 *  ~8 launches per tensor) and update_moving_average (utils/utils.py:328-331, BYOL EMA).
 *  All tensors of a step are described by ONE device table + a chunk map, so a step costs two
 *  launches (LARS) or one (EMA) whatever the number of parameters.  fp32 tensors only.
 * ===================================================================================== */
typedef struct {
    float* p;        /* LARS: parameter (updated in place).  EMA: moving-average parameter (updated in place) */
    const float* g;  /* LARS: gradient.                      EMA: current (online) parameter                */
    float* aux;      /* LARS: momentum buffer `mu` (updated in place).  EMA: unused                          */
    long long n;     /* elements */
    int flags;       /* LARS: bit 0 = apply weight decay, bit 1 = apply the LARS trust ratio (utils/utils.py:169,172) */
    int chunk0;      /* index of this tensor's first chunk in the chunk map */
} abt_opt_tensor;
/* elements per chunk; tensor t owns ceil(n / chunk) consecutive chunks starting at chunk0, chunk_tensor[c] = t */
int abt_opt_chunk_elems(void);
/* tensors_dev / chunk_tensor_dev: DEVICE copies of the table and the chunk map; partial_dev: scratch, 8 bytes per chunk + 4 bytes per tensor */
int abt_lars_step(const abt_opt_tensor* tensors_dev, const int* chunk_tensor_dev, int n_tensors, int n_chunks, float lr, float weight_decay,
                  float momentum, float eta, void* partial_dev, abt_stream_t stream);
/* ma = beta * ma + (1 - beta) * cur for every tensor of the table (utils/utils.py:320-331) */
int abt_ema_update(const abt_opt_tensor* tensors_dev, const int* chunk_tensor_dev, int n_tensors, int n_chunks, float beta, abt_stream_t stream);

/* ===================================================================================== *
 *  Projector tail fused with the objective's statistics pass
 *  Reference: the last bias-free nn.Linear of BarlowTwinsHead (model.py:22, 25-31) followed by
 *  BarlowTwinsLoss.forward_loss (utils/loss.py:15-30).
 * ===================================================================================== */
/* z1 = h1 W^T, z2 = h2 W^T (h*: (n_rows, k_dims) bf16 row-major, w: (n_dims, k_dims) bf16 row-major = nn.Linear.weight,
 * z*: (n_rows, n_dims) bf16, fp32 accumulation on the tensor cores) and, from the ROUNDED outputs, the per-column hand-over of
 * the statistics pass: pack[k * n_dims + j], k = 0..6 = sum z1, sum z1^2, sum z2, sum z2^2, sum z1 z2, 0, 0 -- the layout
 * abt_bt_dist_stats_local writes (5 sums + 2 shifts).  Put it at workspace + abt_bt_dist_layout.pack_all of a world = 1 layout
 * and continue with abt_bt_dist_normalize / abt_bt_dist_rows_fwd_bwd: the objective then never reads z for its statistics.
 * Deterministic.  k_dims and n_dims: multiples of 8, >= 64; pointers 16-byte aligned. */
int abt_proj_tail_workspace_bytes(int n_rows, int n_dims, size_t* bytes);
int abt_proj_tail_fwd(const void* h1, const void* h2, const void* w, int n_rows, int k_dims, int n_dims, void* z1, void* z2,
                      float* pack, void* workspace, size_t workspace_bytes, abt_stream_t stream);

/* ===================================================================================== *
 *  Co-scheduling
 * ===================================================================================== */
/* The tensor-core kernels of the objective are persistent and own every SM they run on (registers and shared memory), so
 * nothing else can start on those SMs until they end -- and they cannot start on an SM that holds any other block.  This call
 * makes them leave `n_sms` SMs (0..64, default 0) free for work that has to run BESIDE them: the pinned-host span gather
 * (abt_wav_span_gather is confined to 4 SMs for exactly this reason) in a prefetching input pipeline, or the caller's own
 * copy kernels.  abt_bt_dist_step applies its own reservation for NCCL's kernels.  Process-wide; returns the previous value. */
int abt_set_reserved_sms(int n_sms);

/* ===================================================================================== *
 *  Debug hooks (not part of the drop-in surface; used by tools/gpu_diag.py)
 * ===================================================================================== */
int abt_debug_set(int key, int value);
/* number of kernels the library has launched (optionally resetting the counter) */
long long abt_debug_launch_count(int reset);
/* device timing of the statistics, CORR and GRAD launches of abt_bt_loss_fwd_bwd (CUDA events on the launching stream) */
int abt_debug_timing(int enable);
int abt_debug_timing_read(float* stats_ms, float* corr_ms, float* grad_ms, int* n_calls);
int abt_debug_ws_offsets(int n_rows, int n_dims, int dtype, size_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* ABT_B200_H */

/*
 * ctc_b200.h -- C ABI of the B200-native CTC loss engine (libctc_b200.so).
 *
 * Drop-in boundary for jinserk/pytorch-asr's deepspeech_ctc training hot path:
 * the loss object constructed at asr/models/trainer.py:152-154
 * (nn.CTCLoss(blank=0, reduction='mean')), called at trainer.py:422 (and :508)
 * as loss(ys_hat[T,N,V], ys, frame_lens, label_lens) and back-propagated at
 * trainer.py:438 (:517).  The reference reaches its CTC arithmetic through
 * torch's Python binding; its only in-tree native binding is the pybind11
 * CppExtension asr/kaldi/src/latgen_lib.cc:278-281 built by
 * asr/kaldi/setup.py:48-71, which is the packaging this library's torch shim
 * (pytorch-asr_b200/csrc/ctc_binding.cc, module torch_asr._ctc_lib) mirrors.
 *
 * Conventions
 *  - Plain C types only.  Every pointer marked "device" is CUDA device memory
 *    on the current device; "host" is ordinary (ideally pinned) host memory.
 *  - The compute entry points never allocate, free or synchronise: the caller
 *    owns every buffer including the workspace, and all work is enqueued on
 *    the stream passed in (a cudaStream_t, declared void* here so that the
 *    header needs no CUDA include).  Stateless and re-entrant.
 *  - Every function returns a ctc_b200_status (0 = ok).  No exceptions cross
 *    this boundary.  The torch shim turns non-zero into RuntimeError, the way
 *    KALDI_ERR surfaces from the latgen binding (latgen_lib.cc:83,88).
 *  - Only fp32 activations: the reference casts to float before the loss
 *    (trainer.py:416-417).
 */
#ifndef CTC_B200_H_
#define CTC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ctc_b200_status {
    CTC_B200_OK = 0,
    CTC_B200_INVALID_ARGUMENT = 1,   /* null pointer, negative size, blank outside [0,V) ... */
    CTC_B200_WORKSPACE_TOO_SMALL = 2,
    CTC_B200_UNSUPPORTED = 3,        /* target longer than 4095 labels, vocabulary too large for smem */
    CTC_B200_CUDA_ERROR = 4,         /* launch / runtime failure, see ctc_b200_last_cuda_error() */
    CTC_B200_BAD_LABEL = 5,          /* device-side check: a label outside [0,V) */
    CTC_B200_BAD_LENGTH = 6,         /* device-side check: input length > T or target length > S_max */
    CTC_B200_PEER_TIMEOUT = 7        /* fused loss all-reduce: a peer's pair did not arrive */
} ctc_b200_status;

/* Memory layout of `acts` and `grad` (both always use the same one). */
typedef enum ctc_b200_layout {
    CTC_B200_LAYOUT_TNV = 0,         /* [T,N,V] time-major: what trainer.py:418 hands to the loss */
    CTC_B200_LAYOUT_NTV = 1          /* [N,T,V] batch-major: what the network emits (network.py:380,393-396);
                                        folds trainer.py:418's transpose(0,1).contiguous() and its backward */
} ctc_b200_layout;

/* Options of the extended calls; NULL means all defaults (time-major, no clamp). */
typedef struct ctc_b200_options {
    int layout;                      /* ctc_b200_layout */
    int use_clamp;                   /* != 0: fused Hardtanh(clamp_min, clamp_max) in front of the log_softmax
                                        (network.py:370 nn.Hardtanh(-50, 50)); the gradient is then with respect to
                                        the RAW (un-clamped) logits: 0 where a logit is outside (clamp_min, clamp_max) */
    float clamp_min, clamp_max;
    int persistent;                  /* != 0 and more utterances than co-resident clusters: launch only the co-resident
                                        clusters and let them pull utterances from a device-side queue (longest first).
                                        Off by default: measured on B200 the hardware CTA scheduler, which hands the
                                        next cluster of the grid to the first free slot, does exactly as well
                                        (B = 512: 0.437 ms both ways; B = 1024: 0.865 vs 0.857 ms).  Honoured for the
                                        headline shape class only (V = 48, targets up to 248 labels; the loop around
                                        the utterance is a separate instantiation there) and ignored otherwise. */
} ctc_b200_options;

typedef enum ctc_b200_reduction {
    CTC_B200_REDUCE_NONE = 0,
    CTC_B200_REDUCE_MEAN = 1,        /* mean_b( nll_b / max(S_b,1) )  -- trainer.py:153 */
    CTC_B200_REDUCE_SUM = 2
} ctc_b200_reduction;

/* Library ABI version (major*1000 + minor). */
int ctc_b200_version(void);

/* Static, never-freed description of a status code. */
const char* ctc_b200_status_string(int status);

/* cudaGetErrorString of the last CUDA failure this thread saw inside the library. */
const char* ctc_b200_last_cuda_error(void);

/*
 * Launch geometry the engine picks for a problem size.  Informational (bench,
 * DESIGN.md); also what ctc_b200_workspace_bytes is derived from.
 */
typedef struct ctc_b200_geometry {
    int kernel;                /* 2: linear-domain warp-specialised kernel (ctc_lin_kernel) with a log-domain
                                  per-utterance fallback pass, 1: log-domain warp-specialised kernel
                                  (ctc_pipe_kernel), 0: generic log-domain kernel (ctc_fused_kernel) */
    int rec_warps;             /* warps per CTA running the lattice recursion */
    int grad_warps;            /* helper warps per CTA: softmax, gradient rows (0: generic kernel); the 15-warp instantiation
                                  for launches of one CTA per SM reports 8 and runs four copy warps besides (threads = 480) */
    int pairs_per_thread;      /* lattice (blank,label) cell pairs per thread */
    int threads;               /* threads per CTA */
    int chunk;                 /* frames per softmax/gradient chunk */
    int row_stride;            /* floats per stored lattice row */
    int smem_bytes;            /* dynamic shared memory per CTA */
    size_t workspace_bytes;    /* for n_utt utterances */
    int variant;               /* which template instantiation of `kernel` is launched (ctc_b200_variant_name) */
    int fallback_kernel;       /* kernel == 2: the log-domain kernel that redoes flagged utterances (1 or 0); else -1 */
    int comb_groups;           /* kernel == 2: combine warps per recursion warp */
    int resident_clusters;     /* kernel == 2: 2-CTA clusters of this instantiation the current device holds at once
                                  (0 when there is no CUDA device to ask) */
} ctc_b200_geometry;

int ctc_b200_get_geometry(int T, int n_utt, int V, int S_max, ctc_b200_geometry* out);
/* "ctc_lin_kernel<8,1,80,128,4,FIX>" ...; static string, "?" for an unknown id. */
const char* ctc_b200_variant_name(int kernel, int variant);

/* Bytes of device workspace ctc_b200_fwd_bwd_f32 needs for (T, N, V, S_max). */
int ctc_b200_workspace_bytes(int T, int N, int V, int S_max, size_t* bytes);

/*
 * Fused log_softmax + CTC forward + gradient w.r.t. the logits.
 * Replaces F.log_softmax -> nn.CTCLoss.forward -> backward
 * (network.py:375, trainer.py:422, trainer.py:438).
 *
 *  acts         device [T,N,V] fp32 logits, contiguous (trainer.py:418)
 *  targets      device int32, concatenated labels (dataloader.py:71)
 *  tgt_offsets  device int32 [N], start of utterance b's labels in `targets`
 *  in_lens      device int32 [N]  (frame_lens, dataloader.py:72)
 *  tgt_lens     device int32 [N]  (label_lens, dataloader.py:73)
 *  S_max        >= max_b tgt_lens[b]  (sizes the CTA; host knows it, the
 *               reference keeps the lengths on the CPU)
 *  nll          device fp32 [N] out: -log p(l_b|x_b); +inf if infeasible
 *               (0 when zero_infinity)
 *  grad         device fp32 [T,N,V] out or NULL (forward only):
 *               grad_scale[b] * (softmax - occupancy), 0 for t >= in_lens[b];
 *               NaN rows for infeasible utterances unless zero_infinity (as torch)
 *  grad_scale   device fp32 [N] or NULL (=1): e.g. 1/(N*max(S_b,1)) for 'mean'
 *  workspace    device, >= ctc_b200_workspace_bytes(...), 256-byte aligned
 *  stream       cudaStream_t
 */
int ctc_b200_fwd_bwd_f32(const float* acts, const int32_t* targets, const int32_t* tgt_offsets,
                         const int32_t* in_lens, const int32_t* tgt_lens, int T, int N, int V,
                         int S_max, int blank, int zero_infinity, float* nll, float* grad,
                         const float* grad_scale, void* workspace, size_t workspace_bytes,
                         void* stream);

/*
 * Same, restricted to utterances [utt_begin, utt_begin + utt_count) of the
 * batch (all array arguments still describe the whole batch).  The workspace
 * only needs ctc_b200_workspace_bytes(T, utt_count, V, S_max).  Used to
 * pipeline host->device copies of batch slices against compute.
 */
int ctc_b200_fwd_bwd_range_f32(const float* acts, const int32_t* targets,
                               const int32_t* tgt_offsets, const int32_t* in_lens,
                               const int32_t* tgt_lens, int T, int N, int V, int S_max,
                               int blank, int zero_infinity, int utt_begin, int utt_count,
                               float* nll, float* grad, const float* grad_scale,
                               void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same with options: batch-major [N,T,V] logits / gradient (folds the
 * transpose + copy of trainer.py:418 and of its backward) and the fused Hardtanh of the FC head
 * (section 8(f)2, network.py:367-375).  opt == NULL: ctc_b200_fwd_bwd_range_f32.
 */
int ctc_b200_fwd_bwd_ex_f32(const float* acts, const int32_t* targets, const int32_t* tgt_offsets,
                            const int32_t* in_lens, const int32_t* tgt_lens, int T, int N, int V,
                            int S_max, int blank, int zero_infinity, int utt_begin, int utt_count,
                            float* nll, float* grad, const float* grad_scale, void* workspace,
                            size_t workspace_bytes, const ctc_b200_options* opt, void* stream);

/*
 * grad[t,b,:] *= scale[per_utt ? b : 0].  Applies autograd's grad_output
 * (trainer.py:429 loss.mul_(0); AMP loss scale, trainer.py:435-436).  Factors
 * equal to 1 are detected on the device and cost no memory traffic.
 */
int ctc_b200_scale_grad_f32(float* grad, const float* scale, int per_utt, int T, int N, int V,
                            void* stream);
/* the same for a gradient in `layout` (ctc_b200_layout) */
int ctc_b200_scale_grad_ex_f32(float* grad, const float* scale, int per_utt, int T, int N, int V,
                               int layout, void* stream);

/*
 * out2[0] = sum_b nll_b / max(S_b,1) (MEAN) or sum_b nll_b (SUM); out2[1] = N.
 * The pair is what a data-parallel job all-reduces; the
 * reduced loss is out2[0]/out2[1] for MEAN, out2[0] for SUM, and is also
 * written to loss[0] when `loss` is not NULL.  Single CTA, fixed summation
 * order: bit-reproducible.
 */
int ctc_b200_reduce_loss_f32(const float* nll, const int32_t* tgt_lens, int N, int reduction,
                             float* out2, float* loss, void* stream);

/*
 * Loss reduction PLUS the trainer's post-loss host checks (SURVEY.md section 8(f)3,
 * trainer.py:423-430) in one launch, so that the three device->host syncs of an iteration
 * (isnan, .item() == +-inf, frame_lens < 2 * label_lens) become ONE 16-byte read of `result`:
 *   result[0]  the reduced loss (as ctc_b200_reduce_loss_f32 writes to `loss`); 0 when
 *              zero_on_short applies
 *   result[1]  flags: the sum of the CTC_B200_FLAG_* bits that apply, as a float VALUE (0 ... 7)
 *   result[2]  the factor the reference applies to the loss before backward: 0 if zero_on_short and
 *              some T_b < 2 * S_b (trainer.py:427-429 loss.mul_(0)), else 1.  Feed it to
 *              ctc_b200_scale_grad_f32 as the scale: no host round trip.
 *   result[3]  number of utterances with T_b < 2 * S_b
 * in_lens may be NULL (then CTC_B200_FLAG_SHORT is never set).  out2 as in ctc_b200_reduce_loss_f32;
 * `loss` (may be NULL) receives a second copy of result[0].
 */
#define CTC_B200_FLAG_NAN 1        /* the reduced loss is NaN               (trainer.py:423 torch.isnan) */
#define CTC_B200_FLAG_INF 2        /* the reduced loss is +-inf             (trainer.py:423 loss.item() == inf) */
#define CTC_B200_FLAG_SHORT 4      /* some utterance has T_b < 2 * S_b      (trainer.py:427) */
int ctc_b200_reduce_loss_status_f32(const float* nll, const int32_t* in_lens, const int32_t* tgt_lens,
                                    int N, int reduction, int zero_on_short, float* out2,
                                    float* loss, float* result4, void* stream);

/*
 * The loss reduction FUSED with the data-parallel job's only collective (SURVEY.md
 * section 8e; replaces ctc_b200_reduce_loss_f32 + an NCCL all-reduce of the
 * (sum, count) pair).  ONE kernel, one CTA: it sums this rank's nll, stores the pair
 * with a sequence number straight into every peer's exchange buffer (P2P stores over
 * NVLink / NVSwitch), waits for the world_size pairs addressed to this rank and adds
 * them in rank order, so that every rank obtains bit-identical results:
 * out2 = (global sum, global utterance count), loss[0] = out2[0] / out2[1] for MEAN,
 * out2[0] for SUM.
 *   peer_bufs  host array of world_size DEVICE pointers: rank r's exchange buffer as
 *              mapped in this process (peer_bufs[rank] is the local one).  Each buffer
 *              is CTC_B200_EXCHANGE_BYTES long, symmetric-memory / IPC mapped by the
 *              caller, and zero-filled before the first call.
 *   seq        1, 2, 3, ... : the same value on every rank for the same step; steps
 *              alternate between two slot sets, so a rank may run one step ahead.
 *   status     workspace status word (the int at offset 0 of a workspace): a peer
 *              that does not arrive within the timeout sets CTC_B200_PEER_TIMEOUT there
 *              (and the result is NaN) instead of hanging the stream.  The timeout is
 *              ctc_b200_set_peer_timeout_ms (default 600 000 ms, the order of a process
 *              group's own timeout; 0 = wait for ever).
 * world_size <= CTC_B200_MAX_PEERS.
 */
#define CTC_B200_MAX_PEERS 8
#define CTC_B200_EXCHANGE_BYTES 256
int ctc_b200_reduce_loss_allreduce_f32(const float* nll, const int32_t* tgt_lens, int N,
                                       int reduction, void* const* peer_bufs, int rank,
                                       int world_size, unsigned seq, float* out2, float* loss,
                                       void* workspace, void* stream);

/*
 * Exchange-only form of the above for a pair that ctc_b200_reduce_loss_f32 already
 * produced (the torch shim's forward): out2 (sum, count) is replaced in place by the
 * global pair; `status_word` is a zero-initialised device int (or a workspace).
 */
int ctc_b200_allreduce_pair_f32(float* out2, int reduction, void* const* peer_bufs, int rank,
                                int world_size, unsigned seq, float* loss, void* status_word,
                                void* stream);
/* Process-wide timeout of the two calls above, in milliseconds (0: no timeout).  Returns the old value. */
long long ctc_b200_set_peer_timeout_ms(long long ms);

/* ------------------------------------------------------------------------
 * Greedy CTC decode + label-error count for `validate`:
 * replaces the per-utterance Python loops of trainer.py:450-463 (unit_validate:
 * onehot2int = arg-max, misc.py:44-51; remove_duplicates(blank=0), misc.py:78-84) and
 * trainer.py:336-343 (edit_distance, Levenshtein with unit costs).
 *
 *  acts       device fp32 logits or log-probs in `layout` (unit_validate uses the network's
 *             own [N,T,V] output: CTC_B200_LAYOUT_NTV); arg-max takes the LOWEST index on ties
 *  in_lens    device int32 [N]; frames t >= in_lens[b] are not read
 *  targets / tgt_offsets / tgt_lens   device int32, as in ctc_b200_fwd_bwd_f32
 *  hyp        device int32 [N,T] out: row b holds the hyp_len[b] decoded labels
 *  hyp_len    device int32 [N] out
 *  dist       device int32 [N] out: edit distance(hyp_b, target_b)
 *  totals     device int64 [2] out: {sum_b dist[b], sum_b tgt_lens[b]}; LER % = 100 * [0] / [1]
 *             (trainer.py:301-304)
 * Needs no workspace (hyp doubles as the arg-max buffer).  targets == NULL: decode only.
 * ------------------------------------------------------------------------ */
int ctc_b200_greedy_decode_ler_i32(const float* acts, int T, int N, int V, int layout,
                                   const int32_t* in_lens, const int32_t* targets,
                                   const int32_t* tgt_offsets, const int32_t* tgt_lens, int blank,
                                   int32_t* hyp, int32_t* hyp_len, int32_t* dist,
                                   long long* totals, void* stream);

/*
 * Device-side validation result of the launches that used `workspace` since it
 * was last cleared.  Synchronises `stream`.  Returns CTC_B200_OK,
 * CTC_B200_BAD_LABEL or CTC_B200_BAD_LENGTH.  ctc_b200_clear_status resets it
 * (asynchronously, on `stream`); a fresh workspace must be cleared once.
 */
int ctc_b200_check_status(const void* workspace, void* stream);
int ctc_b200_clear_status(void* workspace, void* stream);

/* ------------------------------------------------------------------------
 * Host-buffer session: the same path for a caller that holds HOST tensors
 * (the form bench.py's e2e figure and non-torch callers use).  A session owns
 * its device buffers, pinned staging and streams; run() slices the batch,
 * overlaps the host->device copy of slice k+1 with the kernel of slice k and
 * returns the reduced loss (and optionally nll / gradient) to host memory.
 * ------------------------------------------------------------------------ */
typedef struct ctc_b200_session ctc_b200_session;

int ctc_b200_session_create(int T, int N, int V, int S_max, int max_targets, int n_slices,
                            ctc_b200_session** out);
int ctc_b200_session_destroy(ctc_b200_session* s);

/*
 *  acts_host     host [T,N,V] fp32 (pinned for full speed); rows t >= max T_b of a batch slice are not read
 *  targets_host  host int32 [n_targets] concatenated; in_lens/tgt_lens host int32 [N]
 *  loss_host     host fp32 [1] out (reduced per `reduction`; for NONE: sum)
 *  nll_host      host fp32 [N] out or NULL
 *  grad_host     host fp32 [T,N,V] out or NULL (gradient stays on the device,
 *                see ctc_b200_session_grad_device)
 *  want_grad     compute the gradient (scaled for `reduction`) or forward only
 */
int ctc_b200_session_run_host_f32(ctc_b200_session* s, const float* acts_host,
                                  const int32_t* targets_host, int n_targets,
                                  const int32_t* in_lens_host, const int32_t* tgt_lens_host,
                                  int blank, int reduction, int zero_infinity, int want_grad,
                                  float* loss_host, float* nll_host, float* grad_host);
float* ctc_b200_session_grad_device(ctc_b200_session* s);
/* number of kernels the last run() launched */
int ctc_b200_session_last_launches(const ctc_b200_session* s);
/* bytes the last run() copied host -> device: the packed targets / lengths block plus, per batch slice, the logits
 * rows up to the slice's longest utterance (frames t >= T_b are never read, so they are not moved) */
long long ctc_b200_session_last_h2d_bytes(const ctc_b200_session* s);

#ifdef __cplusplus
}
#endif
#endif /* CTC_B200_H_ */

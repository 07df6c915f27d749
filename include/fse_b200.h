/* fse_b200.h — C ABI of the B200-native FluentSpeech spec_denoiser + HiFi-GAN hot path.
 *
 * The reference (Zain-Jiang/Speech-Editing-Toolkit) has no FFI: its seams for this path are Python
 * registries.  This header is the boundary a binding would target; the ctypes binding
 * that ships with this repo is speech_editing_toolkit_b200/_lib.py, and INTEGRATION.md shows the stub a
 * reference maintainer adds.  Each entry point cites the reference interface it replaces
 * (file:line relative to the reference tree).
 *
 * Conventions
 *   - return 0 on success, a negative FSE_E* code on failure; fse_last_error() returns the message of the
 *     last failure on the calling thread.
 *   - no ownership transfer: every I/O and workspace buffer is allocated by the caller (PyTorch);
 *     the handle owns only its repacked copy of the weights.
 *   - all device work is enqueued on the passed cudaStream_t (as void*); no host synchronisation inside,
 *     except in the fse_*_host convenience calls which take HOST buffers and synchronise before returning.
 *   - one handle per device; not thread-safe per handle; re-entrant across handles.
 *   - the persistent denoiser kernels (fse_denoise_step / fse_sample in the tensor-core modes) have CTAs that wait for each other
 *     through flags in global memory: their grid is sized to be fully resident on an otherwise idle device (occupancy query per
 *     handle), so do not run two of them, or one of them next to another kernel that pins SMs for long, concurrently on one device
 *     (separate streams of one process, MPS): a partially resident grid would spin until its bounded wait traps.
 *   - there is NO CPU fallback: every compute call fails with FSE_ECUDA when no sm_100 device is usable.
 */
#ifndef FSE_B200_H
#define FSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSE_OK 0
#define FSE_EINVAL (-1)   /* bad argument / shape / missing weight */
#define FSE_ECUDA (-2)    /* CUDA runtime or driver error (message has the detail) */
#define FSE_ESTATE (-3)   /* call order violated (e.g. weights not loaded) */

/* arithmetic mode of the contractions */
#define FSE_MODE_TC_BF16 0    /* tcgen05 tensor cores, bf16 operands, fp32 accumulate (default, the product) */
#define FSE_MODE_SIMT_F32 1   /* CUDA cores, fp32 operands: the reference's exact fp32 arithmetic contract */
#define FSE_MODE_SIMT_BF16 2  /* CUDA cores over the bf16 operands: on-device cross-check of the TC path */
#define FSE_MODE_TC_TF32 3    /* tcgen05 tensor cores, kind::tf32: fp32 operands in memory, tf32 multiplies, fp32 accumulate -
                                 the reference's own GPU arithmetic (cuDNN TF32 convolutions, torch default); denoiser + vocoder */

typedef struct fse_denoiser fse_denoiser;
typedef struct fse_vocoder fse_vocoder;
typedef struct fse_mel_encoder fse_mel_encoder;

/* Hyper-parameters of DiffNet (diffnet.py:84-108) as read from egs/spec_denoiser.yaml:76-81,128. */
typedef struct fse_denoiser_config {
  int32_t n_mels;                 /* audio_num_mel_bins, 80 */
  int32_t hidden;                 /* hidden_size (conditioner channels), 192 */
  int32_t channels;               /* residual_channels, 256 */
  int32_t layers;                 /* residual_layers, 20 */
  int32_t dilation_cycle_length;  /* dilation = 2**(layer % cycle), shipped configs: 1 */
  int32_t mode;                   /* FSE_MODE_* */
} fse_denoiser_config;

/* A named fp32 tensor in HOST memory (a reference state_dict entry). */
typedef struct fse_tensor {
  const char* name;    /* reference state_dict key, e.g. "residual_layers.3.dilated_conv.weight" */
  const float* data;   /* host pointer, contiguous, fp32 */
  int64_t numel;
} fse_tensor;

const char* fse_last_error(void);
int fse_version(void);

/* --- denoiser -------------------------------------------------------------------------------- */

/* replaces DIFF_DECODERS[hparams['diff_decoder_type']](hparams) -> DiffNet.__init__
 * (tasks/speech_editing/spec_denoiser.py:13-15, diffnet.py:84-108) */
int fse_denoiser_create(const fse_denoiser_config* cfg, fse_denoiser** out);
void fse_denoiser_destroy(fse_denoiser* h);

/* replaces load_ckpt(model, ..., 'model') for the `denoise_fn.*` keys (utils/commons/ckpt_utils.py:26-66):
 * takes the reference state_dict entries unchanged, repacks them on the device. */
int fse_denoiser_load_weights(fse_denoiser* h, const fse_tensor* tensors, int32_t n);

/* posterior buffers of GaussianDiffusion.__init__ (spec_denoiser.py:61-69), each of length timesteps+1:
 * posterior_mean_coef1, posterior_mean_coef2, posterior_log_variance_clipped (host pointers). */
int fse_denoiser_set_schedule(fse_denoiser* h, int32_t timesteps, const float* coef1, const float* coef2,
                              const float* logvar_clipped);

/* bytes of device workspace the calls below need for a [B, *, T] batch */
int64_t fse_denoiser_workspace_bytes(const fse_denoiser* h, int32_t B, int32_t T);

/* replaces DiffNet.forward(spec, diffusion_step, cond) (diffnet.py:110-132).
 *   x_t   [B, n_mels, T] fp32 device (the reference's spec[:,0])
 *   cond  [B, T, hidden] fp32 device — the PHYSICAL layout of the reference's `cond`
 *         (decoder_inp.transpose(1,2) is a view of a [B,T,H] tensor, spec_denoiser.py:167)
 *   t     [B] int64 device
 *   x0    [B, n_mels, T] fp32 device (out) */
int fse_denoise_step(fse_denoiser* h, const float* x_t, const float* cond, const int64_t* t, float* x0,
                     int32_t B, int32_t T, void* workspace, int64_t workspace_bytes, void* stream);

/* replaces GaussianDiffusion.q_posterior_sample (spec_denoiser.py:95-101):
 *   x_prev = coef1[t] x0 + coef2[t] x_t + [t != 0] exp(0.5 logvar[t]) noise.
 * noise may be NULL: normals then come from the in-kernel Philox stream (seed, step). */
int fse_posterior_step(fse_denoiser* h, const float* x0, const float* x_t, const int64_t* t, const float* noise,
                       uint64_t seed, uint32_t step, float* x_prev, int32_t B, int32_t T, void* stream);

/* replaces the infer branch of GaussianDiffusion.forward (spec_denoiser.py:177-185): x_S ~ N(0,I), then
 * p_sample for t = S-1..0 (spec_denoiser.py:103-108), output x[:,0].transpose(1,2).
 *   cond    [B, T, hidden] fp32 device
 *   noise   NULL (Philox, keyed by seed) or [(S+1), B, n_mels, T] fp32 device: noise[0] = x_S,
 *           noise[1+k] = the draw of iteration k (t = S-1-k), incl. the unused draw at t = 0
 *   ref_mel [B, T, n_mels] and mask [B, T] (0/1) optional (both or neither): when given the output is
 *           composited mel*mask + ref*(1-mask) (tasks/speech_editing/spec_denoiser.py:53,84)
 *   mel_out [B, T, n_mels] fp32 device (out)
 *   x_trace NULL or [S, B, n_mels, T] fp32 device: x after every iteration (tests) */
int fse_sample(fse_denoiser* h, const float* cond, const float* noise, uint64_t seed, const float* ref_mel,
               const float* mask, float* mel_out, float* x_trace, int32_t B, int32_t T, void* workspace,
               int64_t workspace_bytes, void* stream);

/* Host-buffer convenience of fse_sample (what a foreign-language binding calls): copies cond (and the
 * optional ref/mask/noise) host->device, samples, copies mel_out device->host, synchronises.
 * All pointers are HOST pointers; the library allocates and frees its own device scratch. */
int fse_sample_host(fse_denoiser* h, const float* cond, const float* noise, uint64_t seed, const float* ref_mel,
                    const float* mask, float* mel_out, int32_t B, int32_t T);

/* number of kernels the last fse_denoise_step / fse_sample call on this handle enqueued */
int64_t fse_denoiser_last_launches(const fse_denoiser* h);

/* --- HiFi-GAN generator ---------------------------------------------------------------------- */

/* The generator hyper-parameters of modules/vocoder/hifigan/hifigan.py:101-124 (config.yaml of the
 * vocoder checkpoint; not shipped with the reference — HiFi-GAN V1 is assumed by the benchmarks). */
typedef struct fse_vocoder_config {
  int32_t n_mels;                    /* 80 */
  int32_t upsample_initial_channel;  /* 512 */
  int32_t num_upsamples;             /* <= 8 */
  int32_t upsample_rates[8];
  int32_t upsample_kernel_sizes[8];
  int32_t num_kernels;               /* <= 4 */
  int32_t resblock_kernel_sizes[4];
  int32_t resblock_dilations[4][3];  /* ResBlock1: three dilations per block; ResBlock2: the first two */
  int32_t mode;                      /* FSE_MODE_* */
  int32_t resblock;                  /* config['resblock']: 1 (also when 0) = ResBlock1 (hifigan.py:27-64), 2 = ResBlock2 (:67-88) */
} fse_vocoder_config;

/* replaces HifiGanGenerator.__init__ (hifigan.py:101-124) */
int fse_vocoder_create(const fse_vocoder_config* cfg, fse_vocoder** out);
void fse_vocoder_destroy(fse_vocoder* h);
/* replaces load_ckpt(model, base_dir, 'model_gen') (tasks/tts/vocoder_infer/hifigan.py:13-21): takes the
 * reference state_dict (weight_g / weight_v pairs, or folded .weight) and folds weight-norm once. */
int fse_vocoder_load_weights(fse_vocoder* h, const fse_tensor* tensors, int32_t n);
int64_t fse_vocoder_workspace_bytes(const fse_vocoder* h, int32_t B, int32_t T);
/* replaces HifiGanGenerator.forward (hifigan.py:126-142) behind BaseTTSInfer.run_vocoder
 * (inference/tts/base_tts_infer.py:44-47): mel [B, T, n_mels] fp32 device -> wav [B, T*hop] fp32 device */
int fse_vocoder_forward(fse_vocoder* h, const float* mel, float* wav, int32_t B, int32_t T, void* workspace,
                        int64_t workspace_bytes, void* stream);
/* Host-buffer convenience (HifiGAN.spec2wav, tasks/tts/vocoder_infer/hifigan.py:23-31) */
int fse_vocoder_forward_host(fse_vocoder* h, const float* mel, float* wav, int32_t B, int32_t T);
int64_t fse_vocoder_last_launches(const fse_vocoder* h);

/* --- MelEncoder (modules/speech_editing/commons/mel_encoder.py:3-19) ----------------------------
 * The context-mel branch of the condition: three Linear layers (ReLU after the first two) over
 * x = ref_mels * (1 - time_mel_masks).  Weight names are the module's state_dict keys:
 * encoder.0.{weight,bias}, encoder.2.{weight,bias}, fc_out.{weight,bias}. */
typedef struct fse_mel_encoder_config {
  int32_t n_mels;   /* audio_num_mel_bins, 80 */
  int32_t hidden;   /* hidden_size, 192 (multiple of 32, <= 256) */
  int32_t mode;     /* FSE_MODE_* */
} fse_mel_encoder_config;
/* replaces MelEncoder.__init__ (mel_encoder.py:4-13) */
int fse_mel_encoder_create(const fse_mel_encoder_config* cfg, fse_mel_encoder** out);
void fse_mel_encoder_destroy(fse_mel_encoder* h);
/* replaces load_ckpt for the mel_encoder.* keys (utils/commons/ckpt_utils.py:26-66) */
int fse_mel_encoder_load_weights(fse_mel_encoder* h, const fse_tensor* tensors, int32_t n);
int64_t fse_mel_encoder_workspace_bytes(const fse_mel_encoder* h, int32_t B, int32_t T);
/* replaces MelEncoder.forward (mel_encoder.py:15-19): x [B, T, n_mels] fp32 device -> out [B, T, hidden] fp32 device.
 * With add / scale non-null it also performs the call site's arithmetic (spec_denoiser.py:162-164):
 *   out = add + MelEncoder(x) * scale[b, t]      (add = decoder_inp [B, T, hidden], scale = tgt_nonpadding [B, T]) */
int fse_mel_encoder_forward(fse_mel_encoder* h, const float* x, const float* add, const float* scale, float* out,
                            int32_t B, int32_t T, void* workspace, int64_t workspace_bytes, void* stream);
int64_t fse_mel_encoder_last_launches(const fse_mel_encoder* h);

/* --- condition encoder: FastSpeech.forward(skip_decoder=True) --------------------------------------
 * modules/speech_editing/spec_denoiser/fs.py:49-189 with encoder_type 'conv' (egs/spec_denoiser.yaml:105-137):
 * TextConvEncoder (modules/commons/conv.py:24-139), spk_embed_proj, dur_embed + DurationPredictor + LengthRegulator
 * (modules/commons/nar_tts_modules.py:8-72), expand_states (modules/tts/commons/align_ops.py:21-25), pitch_embed +
 * PitchPredictor (nar_tts_modules.py:75-100), f0_to_coarse / denorm_f0 (utils/audio/pitch/utils.py:17-28,71-82),
 * mel2token_to_dur (utils/audio/align.py:71-90).  Runs once per batch in front of the sampling loop.
 * The entry points are the pieces the reference itself calls separately (fs.py:83-105 in order; the inference script
 * calls encoder / forward_style_embed / forward_dur on their own, inference/tts/spec_denoiser.py:84-98).
 * Weight names are the reference's `fs.*` state_dict keys without the `fs.` prefix; the
 * unused `decoder.*` / `mel_out.*` entries are ignored.  All tensors are device pointers, integer tensors int64.
 * Every Conv1d is a tensor-core conv-GEMM launch (bf16 operands, fp32 accumulate; fp32 CUDA cores in FSE_MODE_SIMT_F32);
 * LayerNorm / GELU / embedding adds / the 192->1 and 192->2 heads / the integer ops are fp32 / int64 CUDA-core kernels. */
typedef struct fse_cond_encoder fse_cond_encoder;
typedef struct fse_cond_encoder_config {
  int32_t hidden;                  /* hidden_size, 192 (multiple of 64, <= 512) */
  int32_t vocab;                   /* rows of encoder.embed_tokens (len(phone_encoder)) */
  int32_t enc_layers;              /* len(enc_dilations), <= 8 */
  int32_t enc_dilations[8];        /* enc_dilations, shipped: 1,1,1,1 */
  int32_t enc_kernel_size;         /* 5 (odd) */
  int32_t layers_in_block;         /* 2 */
  int32_t enc_post_net_kernel;     /* 3 */
  int32_t dur_predictor_layers;    /* 3 */
  int32_t dur_predictor_kernel;    /* 5 */
  int32_t pitch_predictor_layers;  /* 5 (fs.py:76) */
  int32_t predictor_kernel;        /* 5 */
  int32_t use_pitch_embed;         /* egs/spec_denoiser.yaml: 1; egs/spec_denoiser_libritts.yaml: 0 */
  int32_t use_uv;                  /* pitch_type == 'frame' and use_uv (fs.py:156) */
  int32_t spk_embed_dim;           /* 256 with use_spk_embed, 0 without */
  int32_t mode;                    /* FSE_MODE_* */
} fse_cond_encoder_config;

/* replaces FastSpeech.__init__ (fs.py:49-82) */
int fse_cond_encoder_create(const fse_cond_encoder_config* cfg, fse_cond_encoder** out);
void fse_cond_encoder_destroy(fse_cond_encoder* h);
/* replaces load_ckpt for the `fs.*` keys (utils/commons/ckpt_utils.py:26-66) */
int fse_cond_encoder_load_weights(fse_cond_encoder* h, const fse_tensor* tensors, int32_t n);
/* workspace for a batch of B items, Tt tokens and T frames each (any of the calls below) */
int64_t fse_cond_encoder_workspace_bytes(const fse_cond_encoder* h, int32_t B, int32_t Tt, int32_t T);
int64_t fse_cond_encoder_last_launches(const fse_cond_encoder* h);

/* replaces TextConvEncoder.forward (conv.py:130-139 -> :99-116): txt [B,Tt] int64 -> encoder_out [B,Tt,hidden] fp32 */
int fse_cond_text_encoder(fse_cond_encoder* h, const int64_t* txt, float* encoder_out, int32_t B, int32_t Tt, void* workspace,
                          int64_t workspace_bytes, void* stream);
/* replaces FastSpeech.forward_style_embed (fs.py:114-121): spk_embed [B,spk_embed_dim] -> style [B,hidden] */
int fse_cond_style_embed(fse_cond_encoder* h, const float* spk_embed, float* style, int32_t B, void* stream);
/* fs.py:90 / inference/tts/spec_denoiser.py:93: dur_inp = (encoder_out + style) * (txt > 0); style may be NULL */
int fse_cond_dur_input(fse_cond_encoder* h, const float* encoder_out, const float* style, const int64_t* txt, float* dur_inp,
                       int32_t B, int32_t Tt, void* stream);
/* fs.py:136-138 (integer, bit-exact): masked_dur = mel2token_to_dur(mel2ph * (1 - mask).long(), Tt) * (txt != 0)
 * mel2ph [B,T] int64, mask [B,T] fp32 0/1 (NULL = no mask), masked_dur [B,Tt] int64 (out) */
int fse_cond_masked_dur(fse_cond_encoder* h, const int64_t* mel2ph, const float* mask, const int64_t* txt, int64_t* masked_dur,
                        int32_t B, int32_t T, int32_t Tt, void* stream);
/* replaces the rest of FastSpeech.forward_dur (fs.py:139-148): dur = DurationPredictor(dur_inp + dur_embed(masked_dur),
 * txt == 0) (nar_tts_modules.py:24-34); dur [B,Tt] fp32 (out) */
int fse_cond_duration(fse_cond_encoder* h, const float* dur_inp, const int64_t* masked_dur, const int64_t* txt, float* dur,
                      int32_t B, int32_t Tt, void* workspace, int64_t workspace_bytes, void* stream);
/* replaces LengthRegulator.forward (nar_tts_modules.py:42-72), alpha = 1, in two calls because the output length is data
 * dependent (the reference synchronises on dur.sum(-1).max() too):
 *   _cumsum: cumsum [B,Tt] int64 = cumsum(round_half_even(dur) * (txt != 0)), totals [B] int64 = frames per item
 *   _fill:   mel2ph [B,Tmax] int64: 1-based token index of every frame, 0 past an item's end (Tmax >= max(totals)) */
int fse_cond_length_cumsum(fse_cond_encoder* h, const float* dur, const int64_t* txt, int64_t* cumsum, int64_t* totals, int32_t B,
                           int32_t Tt, void* stream);
int fse_cond_length_fill(fse_cond_encoder* h, const int64_t* cumsum, int64_t* mel2ph, int32_t B, int32_t Tt, int32_t Tmax,
                         void* stream);
/* replaces fs.py:93-102: tgt_nonpadding, expand_states, forward_pitch (fs.py:153-189) and the final
 * decoder_inp = (expand(encoder_out) [+ pitch_embed] + style) * (mel2ph > 0).
 *   encoder_out [B,Tt,hidden], style [B,hidden] or NULL, mel2ph [B,T] int64, mask [B,T] fp32 0/1 (time_mel_masks),
 *   f0 / uv [B,T] fp32 (ignored without use_pitch_embed), use_pred_pitch as the reference flag
 *   out: decoder_inp [B,T,hidden]; with use_pitch_embed also pitch_pred [B,T,2], f0_denorm [B,T], f0_denorm_pred [B,T]
 *        and (optional, may be NULL) pitch [B,T] int64 = f0_to_coarse(f0_denorm), the embedded bins */
int fse_cond_frames(fse_cond_encoder* h, const float* encoder_out, const float* style, const int64_t* mel2ph, const float* mask,
                    const float* f0, const float* uv, int32_t use_pred_pitch, float* decoder_inp, float* pitch_pred,
                    float* f0_denorm, float* f0_denorm_pred, int64_t* pitch, int32_t B, int32_t Tt, int32_t T, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* --- CampNet mask-predict forward (BASELINE configs[3]) -------------------------------------------
 * modules/speech_editing/campnet/campnet.py:14-69: TransformerEncoder (3 EncSALayers) over the phonemes, MelEncoder over the
 * masked mel (mask_emb on the masked frames), TransformerDecoder (6 DecSALayers: self-attention, encoder-decoder attention,
 * causal-padded conv FFN; modules/speech_editing/commons/transformer.py:489-812) -> mel_out_coarse, then MelEncoder +
 * ConvBlocks (decoder_fine, modules/commons/conv.py:68-116) -> mel_out_fine.  Weight names are the reference CampNet's
 * state_dict keys (unused entries — encoder.pre_net.*, mel_out.*, *._float_tensor — are ignored). */
typedef struct fse_campnet fse_campnet;
typedef struct fse_campnet_config {
  int32_t hidden;       /* hidden_size, 192 (= heads * 96) */
  int32_t vocab;        /* ph_dict_size */
  int32_t n_mels;       /* audio_num_mel_bins, 80 */
  int32_t enc_layers;   /* 3 (campnet.py:17-19) */
  int32_t dec_layers;   /* 6 (campnet.py:26-28) */
  int32_t heads;        /* 2 */
  int32_t ffn_kernel;   /* dec_ffn_kernel_size, 9 */
  int32_t fine_blocks;  /* 5 (campnet.py:29-31) */
  int32_t fine_kernel;  /* 5 */
  int32_t mode;         /* FSE_MODE_* */
} fse_campnet_config;
/* replaces CampNet.__init__ (campnet.py:14-38) */
int fse_campnet_create(const fse_campnet_config* cfg, fse_campnet** out);
void fse_campnet_destroy(fse_campnet* h);
/* replaces load_ckpt(model, ..., 'model') for a CampNet checkpoint (utils/commons/ckpt_utils.py:26-66) */
int fse_campnet_load_weights(fse_campnet* h, const fse_tensor* tensors, int32_t n);
/* valid after load_weights */
int64_t fse_campnet_workspace_bytes(const fse_campnet* h, int32_t B, int32_t Tt, int32_t T);
int64_t fse_campnet_last_launches(const fse_campnet* h);
/* replaces CampNet.forward (campnet.py:40-69):
 *   txt [B,Tt] int64, mels [B,T,n_mels] fp32 (all-zero frames = padding), time_mel_masks [B,T] fp32 0/1
 *   out: mel_out_coarse, mel_out_fine [B,T,n_mels]; optional (may be NULL): attn [B,T,Tt] = head-averaged encoder-decoder
 *        attention of decoder layer 0 (the `attn` entry of the reference's dict), encoder_out [B,Tt,hidden] */
int fse_campnet_forward(fse_campnet* h, const int64_t* txt, const float* mels, const float* time_mel_masks, float* mel_out_coarse,
                        float* mel_out_fine, float* attn, float* encoder_out, int32_t B, int32_t Tt, int32_t T, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* --- region surgery of the inference script (inference/tts/spec_denoiser.py:88-131) ------------------
 * The integer / index work between the forced alignment of the original utterance and the model call: the reference does
 * it for one utterance with host-side tensor slicing; these three calls do it on the device for a padded batch (per-item
 * lengths and regions), bit-exactly.  All pointers are device pointers, integer tensors int64 unless noted; *_len may be
 * NULL (= every item uses the full padded length).  regions [B,4] = (w0, w1, c0, c1): the edited word span in the original
 * words and the span that replaces it in the edited words (1-based, inclusive: words_region[0], edited_words_region[0]).
 * Stateless (no handle); asynchronous on the passed stream. */
/* :88-97   masked_dur [B,Tpe] (durations of the phones before / after the span, for forward_dur(masked_dur=...)),
 *          masked_mel2ph [B,T] (0 inside the span), time_mel_masks_orig [B,T] fp32 (1 inside the span) */
int fse_edit_prepare(const int64_t* mel2ph, const int64_t* mel2word, const int64_t* T_len, const int64_t* ph2word, const int64_t* dur,
                     const int64_t* Tp_len, const int64_t* Tpe_len, const int64_t* regions, int64_t* masked_dur, int64_t* masked_mel2ph,
                     float* time_mel_masks_orig, int32_t B, int32_t T, int32_t Tp, int32_t Tpe, void* stream);
/* :99-110  from the predicted alignment edited_mel2ph [B,Te] of the edited text (forward_dur(..., use_pred_mel2ph=True)):
 *          plan [B,8] int64 = (Tn, head_idx, tail_idx, length_edited, n_edit, n_tail, tail_shift, has_tail) and the two
 *          order-preserving selections sel_edit [B,Te] / sel_tail [B,T] (int32) the assembly copies from.  The caller reads
 *          plan[:,0] (one host sync, as the reference's own slicing implies) to size the outputs of fse_edit_assemble. */
int fse_edit_plan(const int64_t* mel2ph, const int64_t* mel2word, const int64_t* T_len, const int64_t* edited_ph2word, const int64_t* Tpe_len,
                  const int64_t* regions, const int64_t* edited_mel2ph, const int64_t* Te_len, int32_t* sel_edit, int32_t* sel_tail, int64_t* plan,
                  int32_t B, int32_t T, int32_t Tpe, int32_t Te, void* stream);
/* :103-131 the model inputs, padded to Tn = max(plan[:,0]): mel2ph [B,Tn] (head copy, edited span, re-based tail), ref_mels
 *          [B,Tn,n_mels] / f0 / uv [B,Tn] (head + tail copies, zeros inside the span), time_mel_masks [B,Tn] fp32 */
int fse_edit_assemble(const int64_t* mel2ph, const int64_t* T_len, const int64_t* regions, const int64_t* plan, const int64_t* edited_mel2ph,
                      const int32_t* sel_edit, const int32_t* sel_tail, const float* mel, const float* f0, const float* uv, int64_t* out_mel2ph,
                      float* out_ref_mels, float* out_f0, float* out_uv, float* out_time_mel_masks, int32_t B, int32_t T, int32_t Te, int32_t Tn,
                      int32_t n_mels, void* stream);

/* --- mel front-end of the inference entry point: wav -> log10-mel -----------------------------------
 * utils/audio/__init__.py:34-81 `librosa_wav2spec` (inference/tts/spec_denoiser.py:258: fmin 55, fmax 7600, sr 22050, fft 1024,
 * hop 256, Hann, 80 mels, eps 1e-6): librosa.stft(center=True, pad_mode="constant") -> |.| -> librosa.filters.mel (Slaney) ->
 * log10(max(eps, .)).  The windowed DFT is a (fft_size / hop_size)-tap conv-GEMM over the waveform viewed as rows of hop_size
 * samples, the mel projection a second GEMM; both on fp32 CUDA cores (low-energy bins need fp32 operands). */
typedef struct fse_mel_frontend fse_mel_frontend;
typedef struct fse_mel_frontend_config {
  int32_t sample_rate;   /* audio_sample_rate, 22050 */
  int32_t fft_size;      /* 1024: an even multiple of hop_size */
  int32_t hop_size;      /* 256 */
  int32_t win_length;    /* win_size, 1024 (must equal fft_size) */
  int32_t num_mels;      /* audio_num_mel_bins, 80 */
  float fmin, fmax;      /* 55, 7600; -1 = 0 / sample_rate/2 as in the reference (:63-64) */
  float eps;             /* 1e-6 */
} fse_mel_frontend_config;
int fse_mel_frontend_create(const fse_mel_frontend_config* cfg, fse_mel_frontend** out);
void fse_mel_frontend_destroy(fse_mel_frontend* h);
/* 1 + n_samples / hop_size (librosa's frame count with center=True) */
int64_t fse_mel_frontend_frames(const fse_mel_frontend* h, int64_t n_samples);
int64_t fse_mel_frontend_workspace_bytes(const fse_mel_frontend* h, int32_t B, int64_t n_samples);
/* wav [B, n_samples] fp32 device (n_samples a multiple of hop_size: pad with zeros, as librosa's own padding is zeros) ->
 * mel [B, frames, num_mels] fp32 device = librosa_wav2spec(...)['mel'] per item */
int fse_mel_frontend_forward(fse_mel_frontend* h, const float* wav, float* mel, int32_t B, int64_t n_samples, void* workspace,
                             int64_t workspace_bytes, void* stream);
int64_t fse_mel_frontend_last_launches(const fse_mel_frontend* h);

/* --- training-mode DiffNet (BASELINE configs[4]) -----------------------------------------
 * The denoiser call of GaussianDiffusion.forward(infer=False) (spec_denoiser.py:168-176: x_0_pred = denoise_fn(x_t, t, cond)) with
 * the activations its backward needs kept in the workspace, and the activation-gradient chain of that backward; every GEMM of both
 * is a conv-GEMM launch of this library.  The weight gradients are GEMMs over tensors left in the workspace (fse_train_layout):
 * the caller takes them with fse_wgrad (below) and the losses with fse_mel_loss_*; optimizer and gradient all-reduce are torch
 * (speech_editing_toolkit_b200/train.py).  mode: FSE_MODE_TC_BF16 | FSE_MODE_TC_TF32 | FSE_MODE_SIMT_F32. */
typedef struct fse_trainer fse_trainer;
int fse_train_create(const fse_denoiser_config* cfg, fse_trainer** out);
void fse_train_destroy(fse_trainer* h);
/* the `denoise_fn.*` parameter tensors as they sit on the DEVICE (fse_tensor.data = device pointers, which must stay valid and
 * are read again by forward / backward for the biases): repacked into operand layouts by device kernels on `stream`; call after
 * every optimizer step */
int fse_train_load_weights_device(fse_trainer* h, const fse_tensor* tensors, int32_t n, void* stream);
int64_t fse_train_workspace_bytes(const fse_trainer* h, int32_t B, int32_t T);
/* byte offsets inside the workspace of the 18 buffers listed in csrc/denoiser_train.cuh (TrainWs): 0 x_rows, 1 h0, 2 h, 3 S, 4 y,
 * 5 hin[L], 6 sg[L], 7 tf[L], 8 u[L], 9 s, 10 r, 11 cond (bf16 copy), 12 dx_rows, 13 dz, 14 dS, 15 dh, 16 dres[L], 17 dy[B*T, L*2C] */
int fse_train_layout(const fse_trainer* h, int32_t B, int32_t T, int64_t* offsets, int32_t n);
/* replaces DiffNet.forward under autograd (diffnet.py:110-132): x_t [B,M,T], cond [B,T,H] (physical layout), d [L,B,C] =
 * diffusion_projection_l(mlp(SinusoidalPosEmb(t))) computed by the caller (a [B,256] problem) -> x0 [B,M,T] */
int fse_train_forward(fse_trainer* h, const float* x_t, const float* cond, const float* d, float* x0, int32_t B, int32_t T,
                      void* workspace, int64_t workspace_bytes, void* stream);
/* replaces the activation-gradient part of autograd's backward through DiffNet: dx0 [B,M,T] -> dcond [B,T,H]; leaves dz, dS, dh
 * (= gradient of h_0), dres[l], dy[l] in the workspace for the caller's weight-gradient GEMMs */
int fse_train_backward(fse_trainer* h, const float* dx0, float* dcond, int32_t B, int32_t T, void* workspace, int64_t workspace_bytes,
                       void* stream);
int64_t fse_train_last_launches(const fse_trainer* h);

/* --- weight-gradient GEMM of the training step ---------------------------------------------
 * Out[m, n, j] = sum over b, t of P[b, t, m] * Q[b, t + offs[j], n] (rows of Q outside [0, T) read as zero): the weight gradient
 * torch.autograd forms for every Conv1d / Linear of DiffNet (diffnet.py:60-132) with P = gradient of the layer's output, Q = the layer's
 * input, offs = the conv's tap offsets.  tcgen05 with MN-major operands straight from the [B, T, channels] buffers (no transposed
 * copies), frames split over the SMs, slices combined in a fixed order (bit-reproducible).  csrc/wgrad.cu
 *   mode: FSE_MODE_TC_BF16 (P, Q bf16) or FSE_MODE_TC_TF32 (P, Q fp32, kind::tf32); P: [B, T, ldp] with M columns used, Q: [B, T, ldq]
 *   with N columns used (row pitches in elements; pointers and pitches 16-byte aligned); out fp32, element (m, n, j) at
 *   out[m * ld_m + n * ld_n + j * ld_j].  workspace: fse_wgrad_workspace_bytes(...) bytes of scratch (the per-slice partial tiles);
 *   it may be shared by consecutive calls on one stream. */
int64_t fse_wgrad_workspace_bytes(int32_t mode, int32_t B, int32_t T, int32_t M, int32_t N, int32_t ntaps);
/* up to 4 such GEMMs over the same (B, T) frame grid in ONE launch (the four weight gradients of a residual layer): their output tiles
 * share the SMs and one reduction; same conventions as fse_wgrad */
typedef struct fse_wgrad_problem {
  const void* P; int64_t ldp;
  const void* Q; int64_t ldq;
  int32_t M, N;
  const int32_t* offs; int32_t ntaps;
  float* out; int64_t ld_m, ld_n, ld_j;
} fse_wgrad_problem;
int64_t fse_wgrad_group_workspace_bytes(int32_t mode, const fse_wgrad_problem* problems, int32_t n, int32_t B, int32_t T);
int fse_wgrad_group(int32_t mode, const fse_wgrad_problem* problems, int32_t n, int32_t B, int32_t T, void* workspace,
                    int64_t workspace_bytes, void* stream);
int fse_wgrad(int32_t mode, const void* P, int64_t ldp, const void* Q, int64_t ldq, int32_t B, int32_t T, int32_t M, int32_t N,
              const int32_t* offs, int32_t ntaps, float* out, int64_t ld_m, int64_t ld_n, int64_t ld_j, void* workspace,
              int64_t workspace_bytes, void* stream);

/* --- mel losses of the training step ------------------------------------------------------
 * SpeechBaseTask.add_mel_loss with `mel_losses: l1:0.5|ssim:0.5` (tasks/tts/speech_base.py:219-257; utils/metrics/ssim.py:24-44):
 * the weights_nonzero_speech-weighted L1 and 1 - SSIM (11 x 11 Gaussian window, bias 6) means over mel_out / target [B,T,n_mels] fp32,
 * and their gradient with respect to mel_out.  forward writes losses[0] = lambda_l1 * l1, losses[1] = lambda_ssim * ssim (device
 * floats) and, with want_grad, keeps three derivative fields in the workspace; backward (same workspace, same mel_out / target)
 * writes grad = dlosses[0] * d losses[0] / d mel_out + dlosses[1] * d losses[1] / d mel_out (dlosses: 2 device floats).
 * Deterministic (no floating-point atomics).  csrc/mel_loss.cu */
int64_t fse_mel_loss_workspace_bytes(int32_t B, int32_t T, int32_t n_mels);
int fse_mel_loss_forward(const float* mel_out, const float* target, float lambda_l1, float lambda_ssim, float* losses, int32_t want_grad,
                         int32_t B, int32_t T, int32_t n_mels, void* workspace, int64_t workspace_bytes, void* stream);
int fse_mel_loss_backward(const float* mel_out, const float* target, const float* dlosses, float lambda_l1, float lambda_ssim, float* grad,
                          int32_t B, int32_t T, int32_t n_mels, void* workspace, int64_t workspace_bytes, void* stream);

/* --- kernel timing (opt-in) -------------------------------------------------------------------
 * When enabled, every kernel the handle launches is bracketed by CUDA events on the launch stream;
 * *_profile_read waits for them and returns the summed device time (ms) and launch count per kind
 * since enabling (arrays of 8).  Denoiser kinds: 0 input projection, 1 gated dilated-conv GEMM,
 * 2 residual/skip GEMM, 3 skip projection, 4 output projection + posterior.  Vocoder kinds: 0 conv_pre,
 * 1 transposed conv, 2 ResBlock convs1, 3 ResBlock convs2 (+residual), 4 conv_post + tanh. */
int fse_denoiser_profile(fse_denoiser* h, int32_t enable);
int fse_denoiser_profile_read(fse_denoiser* h, double* ms_by_kind, int64_t* launches_by_kind);
int fse_vocoder_profile(fse_vocoder* h, int32_t enable);
int fse_vocoder_profile_read(fse_vocoder* h, double* ms_by_kind, int64_t* launches_by_kind);

/* --- test hook --------------------------------------------------------------------------------
 * The bare conv-as-shifted-GEMM primitive with a store-only epilogue (tests/test_gpu_conv_gemm.py):
 *   out[(b,t), n] = sum_{tap,c} A0[b, t+offs[tap], c] * W[n, tap*ceil(C0/KB)*KB + c]
 * A0 [B,T,C0] bf16 device, W [N, ntaps*ceil(C0/KB)*KB] bf16 device (zero padded), out [B*T, N] fp32.
 * mode is FSE_MODE_TC_BF16 or FSE_MODE_SIMT_BF16. */
int fse_debug_conv_gemm(int32_t mode, const void* A0, const void* W, float* out, int32_t B, int32_t T, int32_t C0,
                        int32_t ntaps, const int32_t* offs, int32_t N, int32_t BN, int32_t KB, void* stream,
                        int64_t* dbg_stamps /* device [32] or NULL: clock64 phase stamps of CTA 0 */,
                        int32_t shared_a /* 0: one activation load per tap; 1: one load per channel block, taps read
                                            row-shifted descriptors (2: probe variant that also sets the descriptor
                                            base-offset field; measured wrong on B200) */);

#ifdef __cplusplus
}
#endif
#endif /* FSE_B200_H */

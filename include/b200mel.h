/*
 * b200mel.h -- C ABI of libb200mel.so, the B200 (sm_100a) log-mel front end.
 *
 * This is the drop-in boundary for the one hot path this library replaces in
 * k0r1g/audio-transformers.  The reference has no FFI of its own: the path is reached through
 * two Python calls into third-party packages.  Each entry point below names the reference
 * interface it stands in for (REF = reference tree, HF = transformers, TA = torchaudio):
 *
 *   b200mel_whisper_logmel_f32   HF:models/whisper/feature_extraction_whisper.py:135-164
 *                                (_torch_extract_fbank_features) together with the pad/trim of
 *                                HF:feature_extraction_sequence_utils.py:263-278,327-332, as
 *                                invoked from REF:whisper_finetune/dataset.py:58-62 and
 *                                REF:whisper_finetune/inference.py:154,200.
 *   b200mel_whisper_frame_mask   HF:models/whisper/feature_extraction_whisper.py:328-337
 *                                (attention-mask rescale).
 *   b200mel_mel_f32              TA:transforms/_transforms.py:621-631 (MelSpectrogram.forward ->
 *                                Spectrogram -> MelScale) plus REF:urban_sounds/dataset.py:56
 *                                (torch.log(mel + 1e-9)), as built at REF:urban_sounds/dataset.py:19-24.
 *   b200mel_urban_prep_f32       REF:urban_sounds/dataset.py:26-52 (mono mean, T.Resample, pad/trim, peak
 *                                normalisation ahead of the mel transform).
 *   b200mel_get_table            the construction-time constants of both call sites
 *                                (HF:audio_utils.py:453-544 mel_filter_bank; TA:functional/
 *                                functional.py:518-587 melscale_fbanks; torch.hann_window).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller; 16-byte aligned.
 *   - Every call is asynchronous and stream-ordered on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).  No host synchronisation, no allocation, so a
 *     call may be captured into a CUDA graph.
 *   - Return value: 0 on success, a negative b200mel_status otherwise.  The message for the
 *     calling thread's last failure is available from b200mel_last_error().  Nothing throws or
 *     exits across this boundary.
 *   - A handle is immutable after creation (the optional b200mel_profile_begin/end pair excepted, which
 *     is not thread-safe); concurrent calls from several host threads or streams are safe as long as
 *     each call uses its own workspace.
 *   - The caller's current CUDA device must be the handle's device when a launching entry point is called.
 *   - There is no CPU path: on a machine without a compute-capability-10.x device
 *     b200mel_create fails with B200MEL_ERR_UNSUPPORTED_ARCH / B200MEL_ERR_CUDA.
 */
#ifndef B200MEL_H_
#define B200MEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MEL_VERSION 100 /* 0.1.0 */

typedef enum b200mel_status {
  B200MEL_OK = 0,
  B200MEL_ERR_BAD_ARG = -1,
  B200MEL_ERR_BAD_ALIGN = -2,
  B200MEL_ERR_CUDA = -3,
  B200MEL_ERR_UNSUPPORTED_ARCH = -4,
  B200MEL_ERR_WORKSPACE = -5
} b200mel_status;

typedef enum b200mel_preset {
  B200MEL_PRESET_WHISPER = 0, /* 16 kHz, n_fft 400, hop 160, 80 Slaney mels, 480000 -> 3000 frames */
  B200MEL_PRESET_URBAN = 1    /* 22.05 kHz, n_fft 1024, hop 512, 64 HTK mels, T -> 1 + T/512 frames */
} b200mel_preset;

typedef enum b200mel_table {
  B200MEL_TABLE_WINDOW = 0,     /* n_fft floats, periodic Hann                                  */
  B200MEL_TABLE_FILTERBANK = 1  /* dense (n_fft/2+1) x n_mels floats, row-major, bins x mels    */
} b200mel_table;

typedef struct b200mel_handle b200mel_handle;

/* Library version (B200MEL_VERSION of the build). */
int b200mel_version(void);

/* Message describing the calling thread's most recent failure ("" if none). */
const char* b200mel_last_error(void);

/* Create a handle bound to CUDA device `device` for one preset.  Verifies the device is
 * compute capability 10.x and configures the kernels (shared-memory carve-out, occupancy). */
int b200mel_create(int device, int preset, b200mel_handle** out);
int b200mel_destroy(b200mel_handle* h);

/* Bytes of device workspace a call with `batch` clips needs (per-tile {max, min} pairs the clip-floor pass reduces,
 * 94 x 8 float pairs per clip).  The workspace needs no initialisation by the caller. */
size_t b200mel_workspace_bytes(const b200mel_handle* h, int32_t batch);

/* Whisper preset.
 *   wave     [batch][stride_samples] float32; clip i holds lengths[i] valid samples (the rest is
 *            never read).  stride_samples % 4 == 0.
 *   lengths  [batch] int32 device array, or NULL meaning every clip has `stride_samples` samples.
 *            Clips longer than 480000 samples are truncated, shorter ones are treated as
 *            right-padded with zeros to 480000 (the extractor's padding="max_length").
 *   out      [batch][80][3000] float32: (log10(max(mel,1e-10)) clamped to clip max - 8, + 4) / 4.
 */
int b200mel_whisper_logmel_f32(b200mel_handle* h, const float* wave, int64_t stride_samples,
                               const int32_t* lengths, int32_t batch, float* out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Whisper preset, the encoder stem that consumes the feature map: replaces
 * HF:models/whisper/modeling_whisper.py:619-625 for whisper-tiny,
 *     hidden = gelu(conv2(gelu(conv1(features)))).permute(0, 2, 1) + embed_positions.weight
 * with conv1 = Conv1d(80, 384, 3, padding=1) (:567) and conv2 = Conv1d(384, 384, 3, stride=2, padding=1) (:568), on the
 * tensor cores: BF16 operands, FP32 accumulation, exact (erf) GELU, FP32 output.
 *   features   [batch][80][3000] float32 (the output of b200mel_whisper_logmel_f32)
 *   w1         [384][256] bfloat16: w1[co][tap * 80 + ci] = conv1.weight[co][ci][tap], columns 240..255 zero
 *   w2         [384][1152] bfloat16: w2[co][tap * 384 + ci] = conv2.weight[co][ci][tap]
 *   bias1, bias2   [384] float32;  positions  [1500][384] float32
 *   out        [batch][1500][384] float32
 *   workspace  b200mel_encoder_stem_workspace_bytes(h, batch) bytes (the BF16 im2col image of the features and the
 *              BF16 activations between the convolutions); no initialisation needed.
 * Every pointer 16-byte aligned.  Stream-ordered, no host synchronisation, CUDA-graph capturable. */
size_t b200mel_encoder_stem_workspace_bytes(const b200mel_handle* h, int32_t batch);
int b200mel_encoder_stem_bf16(b200mel_handle* h, const float* features, int32_t batch, const void* w1, const float* bias1,
                              const void* w2, const float* bias2, const float* positions, float* out,
                              void* workspace, size_t workspace_bytes, void* stream);

/* (B, 3000) int32 frame mask: mask[b][t] = (160*t < min(lengths[b], 480000)). */
int b200mel_whisper_frame_mask(b200mel_handle* h, const int32_t* lengths, int32_t batch,
                               int32_t* mask_out, void* stream);

/* Urban preset.
 *   wave      [batch][stride_samples] float32, n_samples valid samples per clip (n_samples >= 513,
 *             reflect padding needs it), stride_samples % 4 == 0.
 *   log_eps   >= 0: out = log(mel + log_eps) (natural log; the reference uses 1e-9);  < 0: out = mel.
 *   out       [batch][64][1 + n_samples/512] float32.
 * batch * (1 + n_samples/512) must stay below 2^31 - 256 (the kernel walks one flat list of frames).
 */
int b200mel_mel_f32(b200mel_handle* h, const float* wave, int64_t stride_samples, int32_t n_samples,
                    int32_t batch, float log_eps, float* out, void* stream);

/* Urban preset, the steps of REF:urban_sounds/dataset.py:26-52 (process_audio) that precede the mel transform:
 * mono mean over the channels (:31-34), torchaudio's sinc/Hann resampler orig_freq -> new_freq (:37-39;
 * TA:functional/functional.py _get_sinc_resample_kernel / _apply_sinc_resample_kernel), zero pad / trim to
 * out_samples (:42-48) and division by the clip's max |x| when that is > 0 (:51-52).
 *   audio        [batch][channels][in_stride] float32 (planar); in_lengths[batch] int32 valid samples per
 *                channel, or NULL meaning in_stride.
 *   orig_freq, new_freq   the two rates divided by their gcd; equal values skip the resampler.
 *   taps         [new_freq][2*width + orig_freq] float32 device table built by the caller with torchaudio's
 *                formula (audio_transformers_b200/urban.py: sinc_resample_kernel); NULL when the rates agree.
 *   out          [batch][out_stride] float32, out_samples written per clip.
 *   workspace    b200mel_urban_prep_workspace_bytes(h, batch) bytes, no initialisation needed.
 */
size_t b200mel_urban_prep_workspace_bytes(const b200mel_handle* h, int32_t batch);
int b200mel_urban_prep_f32(b200mel_handle* h, const float* audio, int64_t in_stride, const int32_t* in_lengths,
                           int32_t channels, int32_t batch, int32_t orig_freq, int32_t new_freq,
                           const float* taps, int32_t width, float* out, int64_t out_stride, int32_t out_samples,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Host-side staging for the drop-in call shapes (REF:whisper_finetune/dataset.py:57-62 hands the extractor float64
 * numpy arrays; HF converts them to float32 and pads on the host, feature_extraction_whisper.py:281-292).  Packs n
 * ragged HOST clips (float32, or float64 when src_is_f64 != 0) into one row-major float32 HOST buffer `dst`
 * (`dst_stride` floats per row; pinned memory makes the following H2D copy asynchronous), converting and copying with
 * up to `threads` host threads.  Only min(lengths[i], max_samples) samples of a clip are copied; out_lengths (may be
 * NULL) receives those counts.  No CUDA calls.  The threads belong to a persistent pool inside the library (created on
 * first use, re-created in a forked child); concurrent callers are serialised. */
int b200mel_host_pack(const void* const* clips, const int64_t* lengths, int32_t n, int32_t src_is_f64,
                      int64_t max_samples, float* dst, int64_t dst_stride, int32_t* out_lengths, int32_t threads);

/* The reference's call shape in one call (REF:whisper_finetune/dataset.py:57-62: the extractor is handed HOST numpy arrays,
 * float64 as `datasets` yields them; HF casts and pads on the host, feature_extraction_whisper.py:281-303).  n ragged HOST
 * clips (clips[i] is float64 when is_f64[i] != 0, else float32; only min(lengths[i], 480000) samples are used) are cast into
 * the caller's PINNED staging buffer `pinned` ([n][width] floats, width % 4 == 0, width >= every used length) by the
 * library's worker threads, each of which copies its piece to `dev_wave` ([n][width], device) with cudaMemcpyAsync on
 * `stream` as soon as the piece is converted; `pinned_lengths` (pinned, n int32) / `dev_lengths` (device) receive the used
 * lengths; then b200mel_whisper_logmel_f32 is enqueued on the same stream.  The call returns when everything is enqueued;
 * `pinned` and `pinned_lengths` must stay untouched until the copies have run (record an event on `stream`).  The handle's
 * device is made current on the worker threads; the caller's current device must be the handle's device. */
int b200mel_whisper_logmel_host(b200mel_handle* h, const void* const* clips, const int64_t* lengths,
                                const uint8_t* is_f64, int32_t n, float* pinned, int64_t width,
                                int32_t* pinned_lengths, float* dev_wave, int32_t* dev_lengths, float* out,
                                void* workspace, size_t workspace_bytes, int32_t threads, void* stream);

/* Optional per-kernel timing for benchmarks: between profile_begin and profile_end every call on
 * this handle brackets its dominant kernel (the fused log-mel kernel, not the clip-floor pass)
 * with a pair of CUDA events recorded on the call's stream, up to `max_launches` pairs.
 * profile_end waits for the recorded events, returns the summed kernel time in milliseconds and the
 * number of launches covered, and switches profiling off again.  Not CUDA-graph capturable and not
 * thread-safe; leave it off in production. */
int b200mel_profile_begin(b200mel_handle* h, int32_t max_launches);
int b200mel_profile_end(b200mel_handle* h, double* total_ms, int32_t* launches);

/* Copy one of the preset's constant tables to HOST memory `dst` (capacity in floats).
 * Returns the number of floats written, or a negative status. */
int64_t b200mel_get_table(int preset, int table, float* dst, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* B200MEL_H_ */

/*
 * l2s_vocoder.h -- C ABI of the B200-native multi_input_vocoder generator forward.
 *
 * This is the drop-in boundary for ONE path of DomhnallBoyle/lip2speech-unit:
 * MelCodeGenerator.forward (multi_input_vocoder/models_multi_input.py:60-97) ->
 * Generator.forward (speech-resynthesis/models.py:98-114), plus the unit-only
 * parent CodeGenerator.forward (speech-resynthesis/models.py:179-229).
 *
 * The reference has no native code and therefore no FFI of its own for this
 * path; the entry points below are what a ctypes binding inside the reference's
 * MelCodeGenerator would call.  Each one names the reference
 * interface it replaces.
 *
 * Conventions: plain C types only; every function returns an int status
 * (L2S_OK == 0) and never throws, exits or synchronises the device unless it
 * says so; all kernels are launched on the caller's stream (from the second forward of a shape on, the conv
 * chain is replayed there as one CUDA graph that was captured on an internal stream); inputs are borrowed
 * for the duration of the call; the output and workspace buffers are owned by
 * the caller; repacked weights are owned by the handle.
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef L2S_VOCODER_H
#define L2S_VOCODER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L2S_MAX_UPS 8
#define L2S_MAX_RK 4
#define L2S_MAX_DIL 4

enum l2s_status {
  L2S_OK = 0,
  L2S_ERR_INVALID = 1,      /* bad argument / config (ValueError / AttributeError upstream)        */
  L2S_ERR_SHAPE = 2,        /* 2*U != T etc.: what torch.cat raises at models_multi_input.py:73   */
  L2S_ERR_CUDA = 3,         /* CUDA runtime / driver failure, text in l2s_last_error               */
  L2S_ERR_STATE = 4,        /* weights missing or l2s_finalize not called                          */
  L2S_ERR_UNSUPPORTED = 5,  /* config outside the path (ResBlock2, odd k-u, text supervision ...)  */
  L2S_ERR_WORKSPACE = 6,    /* workspace too small / misaligned                                    */
  L2S_ERR_INDEX = 7         /* a unit / speaker id outside its table (IndexError upstream)         */
};

enum l2s_precision {
  L2S_PREC_FP32 = 0,  /* fp32 storage and CUDA-core FFMA: the on-device reference mode            */
  L2S_PREC_BF16 = 1,  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate + fp32 residuals  */
  L2S_PREC_TF32 = 2   /* fp32 storage everywhere, tf32 (10-bit mantissa) products on tcgen05       */
};

enum l2s_variant {
  L2S_VARIANT_MULTI_INPUT = 0, /* MelCodeGenerator.forward: units + mel + speaker embedding        */
  L2S_VARIANT_UNIT_ONLY = 1    /* CodeGenerator.forward: units + speaker id                        */
};

enum l2s_dtype { L2S_F32 = 0, L2S_F16 = 1, L2S_BF16 = 2 };

/* Flat copy of the hyper-parameters the reference constructor reads from `h`
 * (Generator.__init__ speech-resynthesis/models.py:73-96, MelCodeGenerator.__init__
 * models_multi_input.py:27-58). */
typedef struct l2s_config {
  int32_t variant;                       /* enum l2s_variant                                      */
  int32_t precision;                     /* enum l2s_precision                                    */
  int32_t n_ups;                         /* len(h.upsample_rates)                                 */
  int32_t up_rates[L2S_MAX_UPS];         /* h.upsample_rates                                      */
  int32_t up_ksizes[L2S_MAX_UPS];        /* h.upsample_kernel_sizes                               */
  int32_t up_init_ch;                    /* h.upsample_initial_channel                            */
  int32_t n_rk;                          /* len(h.resblock_kernel_sizes)                          */
  int32_t rk_sizes[L2S_MAX_RK];          /* h.resblock_kernel_sizes                               */
  int32_t n_dil;                         /* len(h.resblock_dilation_sizes[j]) (same for all j)    */
  int32_t rk_dils[L2S_MAX_RK][L2S_MAX_DIL]; /* h.resblock_dilation_sizes                          */
  int32_t num_embeddings;                /* h.num_embeddings (unit table rows)                    */
  int32_t embedding_dim;                 /* h.embedding_dim                                       */
  int32_t num_mels;                      /* mel bins concatenated in front (80); 0 for unit-only  */
  int32_t spk_dim;                       /* h.embedder_dim: >0 Linear(spk_dim, E); 0: id table    */
  int32_t num_speakers;                  /* rows of the speaker id table when spk_dim == 0        */
  int32_t multispkr;                     /* truthiness of h.multispkr                             */
  int32_t model_in_dim;                  /* h.model_in_dim: conv_pre input channels               */
} l2s_config;

typedef struct l2s_vocoder l2s_vocoder;

/* Replaces MelCodeGenerator.__init__ / CodeGenerator.__init__: validates the
 * config and builds the layer table.  No CUDA call is made (safe before fork). */
int l2s_create(const l2s_config* cfg, l2s_vocoder** out);
void l2s_destroy(l2s_vocoder* v);

/* Replaces load_state_dict + remove_weight_norm (speech-resynthesis/models.py:116-122):
 * the host passes FOLDED plain fp32 weights by reference key name
 * ("conv_pre.weight", "ups.0.bias", "resblocks.3.convs1.2.weight", "dict.weight",
 * "spkr.weight", "layer.0.weight", "fc.bias", "conv_post.weight" ...), host memory,
 * PyTorch layout.  Data is copied. */
int l2s_set_weight(l2s_vocoder* v, const char* name, const float* host_data, int64_t numel);

/* Repack (polyphase split of ConvTranspose1d, tap-major K-major conv weights,
 * bf16 rounding) and upload to `device`.  Synchronises. Must precede l2s_forward. */
int l2s_finalize(l2s_vocoder* v, int device);

/* Bytes of device scratch l2s_forward needs for a (B, T) batch; T = conditioning
 * frames (mel frames; for the unit-only variant T = U).  < 0 on error. */
int64_t l2s_workspace_bytes(l2s_vocoder* v, int32_t batch, int32_t frames);

/* Samples produced per conditioning frame (prod(upsample_rates)). */
int32_t l2s_hop(l2s_vocoder* v);

/* Replaces MelCodeGenerator.forward(**kwargs) / CodeGenerator.forward(**kwargs).
 *   code  int64 (B,U) device
 *   mel   (B,num_mels,T) device, dtype mel_dtype (L2S_F32 / L2S_F16 / L2S_BF16); NULL for unit-only
 *   spkr  float32 (B,spk_dim) device, or int64 (B) speaker ids when spk_dim == 0
 *   out   float32 (B,1,hop*T) device
 * Launches on `stream` (a cudaStream_t); does not synchronise.  Out-of-range
 * ids are clamped on the device and recorded in a sticky flag: the NEXT l2s_forward on this handle (or
 * l2s_poll_index_error) returns L2S_ERR_INDEX -- late rather than never, without a device synchronisation.
 * Thread safety: calls on one handle are serialised by the handle; different handles run concurrently. */
int l2s_forward(l2s_vocoder* v, void* stream, const int64_t* code, const void* mel, int32_t mel_dtype,
                const void* spkr, int32_t batch, int32_t units, int32_t frames, float* out,
                void* workspace, int64_t workspace_bytes);

/* Same, plus the int16 waveform the reference callers derive on the host
 * (inference.py:79-81: audio * 32768 -> astype('int16'), here with saturation):
 * out_i16 int16 (B,hop*T) device; out may be NULL when only int16 is wanted. */
int l2s_forward_i16(l2s_vocoder* v, void* stream, const int64_t* code, const void* mel, int32_t mel_dtype,
                    const void* spkr, int32_t batch, int32_t units, int32_t frames, float* out,
                    int16_t* out_i16, void* workspace, int64_t workspace_bytes);

/* Reads (and clears) the sticky out-of-range-id flag the front-end kernel sets.
 * The flag lives in host-mapped memory, so this never synchronises: after the stream has been synchronised it
 * covers every forward issued so far, otherwise those that have already run.  Returns L2S_ERR_INDEX if an id was
 * out of range since the last poll. */
int l2s_poll_index_error(l2s_vocoder* v);

/* Number of kernels l2s_forward launches for this batch shape (for bench accounting). */
int32_t l2s_launch_count(l2s_vocoder* v, int32_t batch, int32_t frames);

const char* l2s_last_error(l2s_vocoder* v);
const char* l2s_version(void);

#ifdef __cplusplus
}
#endif
#endif /* L2S_VOCODER_H */

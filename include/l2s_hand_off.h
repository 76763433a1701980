/*
 * l2s_hand_off.h -- host-side file I/O of the vocoder service's on-disk hand-off,
 * part of the same C-ABI library as l2s_vocoder.h.  No CUDA call is made by these functions; the buffers they
 * fill / drain are (pinned) host memory owned by the caller.
 *
 * What they replace in the reference (one Python call per file there, one native call per batch here):
 *   l2s_io_read_npy_f32   np.load of mel/<id>.npy and spk_emb/<id>.npy in dataset_multi_input.py __getitem__
 *                         (multi_input_vocoder/dataset_multi_input.py:174, :219-241) followed by the collate into a batch
 *   l2s_io_write_wav_i16  scipy.io.wavfile.write(path, 16000, int16 audio) of inference.py:152-165 and
 *                         inference_server.py:133-146, for every utterance of a batch
 *   l2s_io_units_to_ids   code_to_sequence (dataset_multi_input.py:128-141) for every row of a batch
 * Both run the files of one call on `threads` host threads and return when all files are done.
 */
#ifndef L2S_HAND_OFF_H
#define L2S_HAND_OFF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum l2s_io_status {
  L2S_IO_OK = 0,
  L2S_IO_ERR_ARG = 1,      /* NULL pointer, n < 0, cols <= 0                                                     */
  L2S_IO_ERR_OPEN = 2,     /* a file could not be opened / read / written completely                             */
  L2S_IO_ERR_FORMAT = 3    /* not an .npy of version 1-3, not C order, dtype other than <f4 / <f2, its row length */
                           /* differs from `cols`, or a L2S_IO_REQUIRE_* flag is not met: the caller falls back   */
                           /* to numpy (and to the reference's own error messages) for such a file                */
};

enum l2s_io_flags {
  L2S_IO_REQUIRE_1D = 1,   /* the array must be 1-D (cols,): the speaker embedding check of helpers.py:194          */
  L2S_IO_REQUIRE_2D = 2,   /* the array must be 2-D (rows, cols)                                                   */
  L2S_IO_REQUIRE_F32 = 4   /* the dtype on disk must be float32                                                    */
};

/* Reads n .npy arrays -- 2-D (rows_i, cols) or 1-D (cols,) = one row -- of little-endian float32 or float16 in C order
 * and stacks them as float32: the first min(rows_i, max_rows[i]) rows of file i land at dst + i * dst_stride (in floats;
 * dst_stride >= max_rows[i] * cols); rows_out[i] receives how many rows were copied.  The rest of slot i is left untouched.
 * `flags` is an OR of enum l2s_io_flags.  Returns L2S_IO_OK, or the status of the first failing file with its index in *bad (may be NULL). */
int l2s_io_read_npy_f32(const char* const* paths, int32_t n, float* dst, int64_t dst_stride, const int32_t* max_rows,
                        int32_t cols, int32_t flags, int32_t* rows_out, int32_t threads, int32_t* bad);

/* Writes n mono 16-bit PCM RIFF files at `rate` Hz: file i holds n_samples[i] samples starting at samples + i * stride
 * (in samples).  The bytes equal scipy.io.wavfile.write for a 1-D int16 array (44-byte header, no extra chunks).
 * Directories must exist.  Returns L2S_IO_OK, or the status of the first failing file with its index in *bad. */
int l2s_io_write_wav_i16(const char* const* paths, int32_t n, const int16_t* samples, int64_t stride, const int32_t* n_samples,
                         int32_t rate, int32_t threads, int32_t* bad);

/* Unit strings -> dictionary indices for n manifest rows: line i (the .unt line of the row, tokens separated by blanks, an
 * optional "name|" prefix already removed) is split, every token is looked up in dict_tokens (token j of dict.unt.txt maps
 * to j; tokens absent from the dictionary are DROPPED, the collapse_code = False branch of
 * multi_input_vocoder/dataset_multi_input.py:128-141), and the first min(count, max_ids[i]) indices land at
 * out + i * out_stride; n_out[i] receives the number of known tokens of the line (before the cap).  One call per batch
 * replaces n list comprehensions over a Python dict (5-7 ms of interpreter time per 128-utterance request). */
int l2s_io_units_to_ids(const char* const* lines, int32_t n, const char* const* dict_tokens, int32_t n_dict, int64_t* out,
                        int64_t out_stride, const int32_t* max_ids, int32_t* n_out, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* L2S_HAND_OFF_H */

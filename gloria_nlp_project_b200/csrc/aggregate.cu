// Word-piece aggregation of the text encoder (gloria/models/text_model.py:32-90, BertEncoder.aggregate_tokens) on the
// device: the step immediately before the loss path.
//
// The reference walks every token of every caption on the host (`word_id.item()` = one device sync per token) and
// builds word embeddings as sums of their word pieces ("##" continuations), per layer:
//   - a token that does not start with "##" closes the current word and starts a new one ([CLS] is the first word),
//   - a "##" token joins the current word,
//   - "[SEP]" closes the current word, is appended as a word of its own, and ends the caption,
//   - without a "[SEP]" (truncated caption) the word still open at the end is never emitted,
//   - the word axis is zero-padded back to the token count.
// Here: one small kernel derives, per caption, the token range of every word from the ids (a table look-up says which
// vocabulary entries start with "##"); the aggregation itself is one streaming pass -- every input element is read once
// and every output element written once (HBM bound: 2 x B x layers x T x D x sizeof(T) bytes).  The backward is the
// matching gather.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gloria {
namespace {

// word_range [B, T, 2] = (first token, one past the last token) of word w, (0, 0) for w >= n_words[b];
// token_word [B, T] = word index of token t, or -1 if the token is not part of any emitted word.
// One warp per caption (T <= 1024).
// cap_lens (optional, with is_bracket [vocab] = 1 where the entry's text, "##" stripped, starts with '['):
// gloria_model.py:107-109 on the device -- 1 + the number of emitted words whose string does not start with '['.
__global__ void word_ranges(const long long* __restrict__ ids, const unsigned char* __restrict__ is_cont, int vocab,
                            long long sep_id, int B, int T, int* __restrict__ word_range, int* __restrict__ token_word,
                            int* __restrict__ n_words, const unsigned char* __restrict__ is_bracket,
                            int* __restrict__ cap_lens) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const long long* id = ids + (size_t)b * T;
  int* wr = word_range + (size_t)b * T * 2;
  int* tw = token_word + (size_t)b * T;
  for (int t = lane; t < T; t += 32) { wr[2 * t] = 0; wr[2 * t + 1] = 0; tw[t] = -1; }
  __syncwarp();
  if (lane != 0) return;                       // the walk itself is sequential and tiny (T <= ~100 tokens)
  int w = -1, start = 0;
  bool closed = false;
  int plain = 0;                               // emitted words that do not start with '['
  auto is_plain = [&](int t0) {
    const long long v0 = id[t0];
    return !(is_bracket != nullptr && v0 >= 0 && v0 < vocab && is_bracket[v0] != 0);
  };
  for (int t = 0; t < T; ++t) {
    const long long v = id[t];
    const bool cont = v >= 0 && v < vocab && is_cont[v] != 0;
    if (v == sep_id) {
      if (w >= 0) { wr[2 * w] = start; wr[2 * w + 1] = t; plain += is_plain(start); }   // close the open word
      else ++plain;                            // [SEP] first: the reference appends an empty word string before it
      ++w;
      wr[2 * w] = t; wr[2 * w + 1] = t + 1;                           // [SEP] is a word of its own
      tw[t] = w;
      closed = true;
      break;
    }
    if (!cont || w < 0) {                      // starts a word (a leading "##" piece starts the first word)
      if (w >= 0) { wr[2 * w] = start; wr[2 * w + 1] = t; plain += is_plain(start); }
      ++w;
      start = t;
    }
    tw[t] = w;
  }
  int n = w + 1;
  if (!closed && w >= 0) {                     // no [SEP]: the word still open is never emitted (text_model.py:46-75)
    for (int t = start; t < T; ++t) tw[t] = -1;
    wr[2 * w] = 0; wr[2 * w + 1] = 0;
    n = w;
  }
  n_words[b] = n;
  if (cap_lens != nullptr) cap_lens[b] = plain + 1;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// out[b, layer, w, :] = sum of emb[b, layer, t, :] over the tokens of word w (zeros beyond the caption's words).
// grid (T, layers, B), block = 128 or 256 threads over D.
template <typename T>
__global__ void aggregate_fwd(const T* __restrict__ emb, const int* __restrict__ word_range, int layers, int Tn, int D,
                              T* __restrict__ out) {
  const int w = blockIdx.x, ly = blockIdx.y, b = blockIdx.z;
  const int t0 = word_range[((size_t)b * Tn + w) * 2], t1 = word_range[((size_t)b * Tn + w) * 2 + 1];
  const T* src = emb + ((size_t)b * layers + ly) * Tn * D;
  T* dst = out + (((size_t)b * layers + ly) * Tn + w) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int t = t0; t < t1; ++t) acc += to_f<T>(src[(size_t)t * D + d]);
    dst[d] = from_f<T>(acc);
  }
}

// d_emb[b, layer, t, :] = d_out[b, layer, token_word[b, t], :]  (0 for tokens outside every word)
template <typename T>
__global__ void aggregate_bwd(const T* __restrict__ d_out, const int* __restrict__ token_word, int layers, int Tn, int D,
                              T* __restrict__ d_emb) {
  const int t = blockIdx.x, ly = blockIdx.y, b = blockIdx.z;
  const int w = token_word[(size_t)b * Tn + t];
  const T* src = d_out + (((size_t)b * layers + ly) * Tn + (w < 0 ? 0 : w)) * D;
  T* dst = d_emb + (((size_t)b * layers + ly) * Tn + t) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = w < 0 ? from_f<T>(0.f) : src[d];
}

// 16-byte-per-thread variants (D * sizeof(T) a multiple of 16): one thread = one 16-byte chunk of one output row, all
// accesses coalesced; the flat chunk index is decoded into (row, chunk) and the row into (b, layer, w).
template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__half> { static constexpr int N = 8; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T, bool BWD>
__global__ void __launch_bounds__(256) aggregate_vec(const T* __restrict__ in, const int* __restrict__ map, int layers,
                                                     int Tn, int D, long long n_chunks, T* __restrict__ out) {
  constexpr int N = Vec16<T>::N;
  const int cpr = D / N;                                          // chunks per row
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cpr;
    const int c = (int)(i - row * cpr);
    const int w = (int)(row % Tn);
    const long long bl = row / Tn;                               // b * layers + layer
    const int b = (int)(bl / layers);
    const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)bl * Tn * D) + c;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (BWD) {
      const int ww = map[(size_t)b * Tn + w];                    // here `w` is the token, map = token_word
      if (ww >= 0) o = __ldg(src + (size_t)ww * cpr);
    } else {
      const int t0 = map[((size_t)b * Tn + w) * 2], t1 = map[((size_t)b * Tn + w) * 2 + 1];
      if (t1 - t0 == 1) {
        o = __ldg(src + (size_t)t0 * cpr);                       // single-piece word: a copy, no re-rounding
      } else if (t1 > t0) {
        float acc[N];
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] = 0.f;
        for (int t = t0; t < t1; ++t) {
          const uint4 v = __ldg(src + (size_t)t * cpr);
          const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
          for (int k = 0; k < N; ++k) acc[k] += to_f<T>(e[k]);
        }
        T* r = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int k = 0; k < N; ++k) r[k] = from_f<T>(acc[k]);
      }
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

template <typename T, bool BWD>
int launch_vec(const void* in, const int* map, int B, int layers, int Tn, int D, void* out, cudaStream_t st) {
  const long long n_chunks = (long long)B * layers * Tn * (D / Vec16<T>::N);
  const long long blocks = (n_chunks + 255) / 256;
  const int grid = (int)(blocks < 148 * 32 ? blocks : 148 * 32);
  aggregate_vec<T, BWD><<<grid, 256, 0, st>>>((const T*)in, map, layers, Tn, D, n_chunks, (T*)out);
  GLORIA_LAUNCHED(BWD ? "aggregate_vec(bwd)" : "aggregate_vec(fwd)");
  return GLORIA_OK;
}

template <typename T>
int launch_fwd(const void* emb, const int* wr, int B, int layers, int Tn, int D, void* out, cudaStream_t st) {
  if (D % Vec16<T>::N == 0 && (((uintptr_t)emb | (uintptr_t)out) & 15) == 0)
    return launch_vec<T, false>(emb, wr, B, layers, Tn, D, out, st);
  aggregate_fwd<T><<<dim3(Tn, layers, B), D >= 512 ? 256 : 128, 0, st>>>((const T*)emb, wr, layers, Tn, D, (T*)out);
  GLORIA_LAUNCHED("aggregate_fwd");
  return GLORIA_OK;
}
template <typename T>
int launch_bwd(const void* d_out, const int* tw, int B, int layers, int Tn, int D, void* d_emb, cudaStream_t st) {
  if (D % Vec16<T>::N == 0 && (((uintptr_t)d_out | (uintptr_t)d_emb) & 15) == 0)
    return launch_vec<T, true>(d_out, tw, B, layers, Tn, D, d_emb, st);
  aggregate_bwd<T><<<dim3(Tn, layers, B), D >= 512 ? 256 : 128, 0, st>>>((const T*)d_out, tw, layers, Tn, D, (T*)d_emb);
  GLORIA_LAUNCHED("aggregate_bwd");
  return GLORIA_OK;
}

}  // namespace
}  // namespace gloria

using namespace gloria;

extern "C" int gloria_b200_word_ranges(const long long* caption_ids, const unsigned char* is_continuation, int vocab,
                                       long long sep_id, int B, int T, int32_t* word_range, int32_t* token_word,
                                       int32_t* n_words, void* stream) {
  GLORIA_CHECK_ARG(caption_ids && is_continuation && word_range && token_word && n_words, "null pointer");
  GLORIA_CHECK_ARG(B > 0 && T > 0 && vocab > 0, "bad sizes B=%d T=%d vocab=%d", B, T, vocab);
  cudaStream_t st = (cudaStream_t)stream;
  word_ranges<<<(B + 3) / 4, 128, 0, st>>>(caption_ids, is_continuation, vocab, sep_id, B, T, word_range, token_word,
                                          n_words, nullptr, nullptr);
  GLORIA_LAUNCHED("word_ranges");
  return GLORIA_OK;
}

extern "C" int gloria_b200_word_ranges_cap_lens(const long long* caption_ids, const unsigned char* is_continuation,
                                                const unsigned char* is_bracket, int vocab, long long sep_id, int B,
                                                int T, int32_t* word_range, int32_t* token_word, int32_t* n_words,
                                                int32_t* cap_lens, void* stream) {
  GLORIA_CHECK_ARG(caption_ids && is_continuation && is_bracket && word_range && token_word && n_words && cap_lens,
                   "null pointer");
  GLORIA_CHECK_ARG(B > 0 && T > 0 && vocab > 0, "bad sizes B=%d T=%d vocab=%d", B, T, vocab);
  cudaStream_t st = (cudaStream_t)stream;
  word_ranges<<<(B + 3) / 4, 128, 0, st>>>(caption_ids, is_continuation, vocab, sep_id, B, T, word_range, token_word,
                                          n_words, is_bracket, cap_lens);
  GLORIA_LAUNCHED("word_ranges");
  return GLORIA_OK;
}

extern "C" int gloria_b200_aggregate_tokens_fwd(const void* embeddings, int dtype, const int32_t* word_range, int B,
                                                int layers, int T, int D, void* out, void* stream) {
  GLORIA_CHECK_ARG(embeddings && word_range && out, "null pointer");
  GLORIA_CHECK_ARG(B > 0 && layers > 0 && T > 0 && D > 0 && layers <= 65535 && B <= 65535, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case GLORIA_DTYPE_F32: return launch_fwd<float>(embeddings, word_range, B, layers, T, D, out, st);
    case GLORIA_DTYPE_F16: return launch_fwd<__half>(embeddings, word_range, B, layers, T, D, out, st);
    case GLORIA_DTYPE_BF16: return launch_fwd<__nv_bfloat16>(embeddings, word_range, B, layers, T, D, out, st);
  }
  return fail(GLORIA_ERR_UNSUPPORTED, "dtype %d", dtype);
}

extern "C" int gloria_b200_aggregate_tokens_bwd(const void* d_out, int dtype, const int32_t* token_word, int B,
                                                int layers, int T, int D, void* d_embeddings, void* stream) {
  GLORIA_CHECK_ARG(d_out && token_word && d_embeddings, "null pointer");
  GLORIA_CHECK_ARG(B > 0 && layers > 0 && T > 0 && D > 0 && layers <= 65535 && B <= 65535, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case GLORIA_DTYPE_F32: return launch_bwd<float>(d_out, token_word, B, layers, T, D, d_embeddings, st);
    case GLORIA_DTYPE_F16: return launch_bwd<__half>(d_out, token_word, B, layers, T, D, d_embeddings, st);
    case GLORIA_DTYPE_BF16: return launch_bwd<__nv_bfloat16>(d_out, token_word, B, layers, T, D, d_embeddings, st);
  }
  return fail(GLORIA_ERR_UNSUPPORTED, "dtype %d", dtype);
}

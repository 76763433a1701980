// Host-side file I/O of the on-disk hand-off (include/l2s_hand_off.h): batches of .npy reads and wav writes on native
// threads, so that the service's per-file work leaves the Python interpreter (the ctypes call releases the GIL).
// Plain C++ -- no CUDA call in this translation unit.
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/l2s_hand_off.h"

namespace {

// Runs job(i) for i in [0, n) on up to `threads` threads (the calling thread is one of them).  job returns a status;
// the first non-zero one (lowest file index wins on ties of time) is reported with its index.
template <class Job>
int run_jobs(int n, int threads, int32_t* bad, Job job) {
  std::atomic<int> next{0};
  std::atomic<long long> first_bad{-1};      // (index << 8) | status of the failing file with the lowest index
  auto worker = [&]() {
    for (;;) {
      const int i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n) return;
      const int st = job(i);
      if (st != L2S_IO_OK) {
        const long long mine = ((long long)i << 8) | st;
        long long cur = first_bad.load();
        while ((cur < 0 || mine < cur) && !first_bad.compare_exchange_weak(cur, mine)) {}
      }
    }
  };
  const int nt = threads < 1 ? 1 : (threads > n ? (n > 0 ? n : 1) : threads);
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  const long long fb = first_bad.load();
  if (fb < 0) return L2S_IO_OK;
  if (bad) *bad = (int32_t)(fb >> 8);
  return (int)(fb & 0xff);
}

bool read_full(int fd, void* buf, size_t bytes) {
  uint8_t* p = static_cast<uint8_t*>(buf);
  while (bytes > 0) {
    const ssize_t r = ::read(fd, p, bytes);
    if (r <= 0) return false;
    p += r;
    bytes -= (size_t)r;
  }
  return true;
}

inline float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {                                   // subnormal: renormalise
      int e = -1;
      do { ++e; man <<= 1; } while ((man & 0x400u) == 0);
      bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | man << 13;
  } else {
    bits = sign | (exp + 127 - 15) << 23 | man << 13;
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

// value of key 'name' in the header dict text: pointer just behind "'name':" (spaces skipped), or nullptr
const char* find_key(const char* hdr, const char* name) {
  char pat[32];
  std::snprintf(pat, sizeof pat, "'%s':", name);
  const char* p = std::strstr(hdr, pat);
  if (!p) return nullptr;
  p += std::strlen(pat);
  while (*p == ' ') ++p;
  return p;
}

int read_one_npy(const char* path, float* dst, int max_rows, int cols, int flags, int32_t* rows_out) {
  const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) return L2S_IO_ERR_OPEN;
  struct Closer { int fd; ~Closer() { ::close(fd); } } closer{fd};
  uint8_t pre[12];
  if (!read_full(fd, pre, 10)) return L2S_IO_ERR_OPEN;
  if (std::memcmp(pre, "\x93NUMPY", 6) != 0) return L2S_IO_ERR_FORMAT;
  size_t hlen;
  if (pre[6] == 1) {
    hlen = (size_t)pre[8] | (size_t)pre[9] << 8;
  } else if (pre[6] == 2 || pre[6] == 3) {
    if (!read_full(fd, pre + 10, 2)) return L2S_IO_ERR_OPEN;
    hlen = (size_t)pre[8] | (size_t)pre[9] << 8 | (size_t)pre[10] << 16 | (size_t)pre[11] << 24;
  } else {
    return L2S_IO_ERR_FORMAT;
  }
  if (hlen == 0 || hlen > (1u << 16)) return L2S_IO_ERR_FORMAT;
  std::vector<char> hdr(hlen + 1, '\0');
  if (!read_full(fd, hdr.data(), hlen)) return L2S_IO_ERR_OPEN;
  const char* descr = find_key(hdr.data(), "descr");
  const char* order = find_key(hdr.data(), "fortran_order");
  const char* shape = find_key(hdr.data(), "shape");
  if (!descr || !order || !shape || std::strncmp(order, "False", 5) != 0 || *shape != '(') return L2S_IO_ERR_FORMAT;
  int esize;
  if (std::strncmp(descr, "'<f4'", 5) == 0) esize = 4;
  else if (std::strncmp(descr, "'<f2'", 5) == 0) esize = 2;
  else return L2S_IO_ERR_FORMAT;
  long long dims[2] = {0, 0};
  int nd = 0;
  for (const char* p = shape + 1; *p && *p != ')';) {
    if (*p == ' ' || *p == ',') { ++p; continue; }
    if (*p < '0' || *p > '9' || nd == 2) return L2S_IO_ERR_FORMAT;
    char* end;
    dims[nd++] = std::strtoll(p, &end, 10);
    p = end;
  }
  long long rows, row_len;
  if (nd == 1) { rows = 1; row_len = dims[0]; }
  else if (nd == 2) { rows = dims[0]; row_len = dims[1]; }
  else return L2S_IO_ERR_FORMAT;
  if (row_len != cols) return L2S_IO_ERR_FORMAT;
  if (((flags & L2S_IO_REQUIRE_1D) && nd != 1) || ((flags & L2S_IO_REQUIRE_2D) && nd != 2) || ((flags & L2S_IO_REQUIRE_F32) && esize != 4))
    return L2S_IO_ERR_FORMAT;
  const long long take = rows < max_rows ? rows : max_rows;
  const size_t count = (size_t)(take > 0 ? take : 0) * (size_t)cols;
  if (esize == 4) {
    if (count && !read_full(fd, dst, count * 4)) return L2S_IO_ERR_OPEN;
  } else {
    // float16 on disk (create_dataset.vocoder() may store halves): read into the upper half of the slot, widen in place
    uint16_t* stage = reinterpret_cast<uint16_t*>(dst) + count;
    if (count && !read_full(fd, stage, count * 2)) return L2S_IO_ERR_OPEN;
    for (size_t j = 0; j < count; ++j) dst[j] = half_to_float(stage[j]);
  }
  *rows_out = (int32_t)(take > 0 ? take : 0);
  return L2S_IO_OK;
}

int write_one_wav(const char* path, const int16_t* samples, int n, int rate) {
  const uint32_t bytes = (uint32_t)n * 2u;
  uint8_t h[44];
  auto put32 = [&](int at, uint32_t v) { h[at] = v & 0xff; h[at + 1] = (v >> 8) & 0xff; h[at + 2] = (v >> 16) & 0xff; h[at + 3] = (v >> 24) & 0xff; };
  auto put16 = [&](int at, uint32_t v) { h[at] = v & 0xff; h[at + 1] = (v >> 8) & 0xff; };
  std::memcpy(h, "RIFF", 4);
  put32(4, 36u + bytes);
  std::memcpy(h + 8, "WAVEfmt ", 8);
  put32(16, 16);                 // fmt chunk size
  put16(20, 1);                  // PCM
  put16(22, 1);                  // mono
  put32(24, (uint32_t)rate);
  put32(28, (uint32_t)rate * 2u);
  put16(32, 2);                  // block align
  put16(34, 16);                 // bits per sample
  std::memcpy(h + 36, "data", 4);
  put32(40, bytes);
  const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
  if (fd < 0) return L2S_IO_ERR_OPEN;
  struct iovec iov[2] = {{h, sizeof h}, {const_cast<int16_t*>(samples), bytes}};
  size_t left = sizeof h + bytes;
  int first = 0;
  bool ok = true;
  while (left > 0) {
    const ssize_t w = ::writev(fd, iov + first, 2 - first);
    if (w <= 0) { ok = false; break; }
    left -= (size_t)w;
    size_t adv = (size_t)w;
    while (first < 2 && adv >= iov[first].iov_len) adv -= iov[first++].iov_len;
    if (first < 2) {
      iov[first].iov_base = static_cast<uint8_t*>(iov[first].iov_base) + adv;
      iov[first].iov_len -= adv;
    }
  }
  if (::close(fd) != 0) ok = false;
  return ok ? L2S_IO_OK : L2S_IO_ERR_OPEN;
}

}  // namespace

extern "C" int l2s_io_read_npy_f32(const char* const* paths, int32_t n, float* dst, int64_t dst_stride, const int32_t* max_rows,
                                   int32_t cols, int32_t flags, int32_t* rows_out, int32_t threads, int32_t* bad) {
  if (n < 0 || cols <= 0 || (n > 0 && (!paths || !dst || !max_rows || !rows_out))) return L2S_IO_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (!paths[i] || max_rows[i] < 0 || (int64_t)max_rows[i] * cols > dst_stride) return L2S_IO_ERR_ARG;
  return run_jobs(n, threads, bad, [&](int i) { return read_one_npy(paths[i], dst + (int64_t)i * dst_stride, max_rows[i], cols, flags, rows_out + i); });
}

extern "C" int l2s_io_write_wav_i16(const char* const* paths, int32_t n, const int16_t* samples, int64_t stride,
                                    const int32_t* n_samples, int32_t rate, int32_t threads, int32_t* bad) {
  if (n < 0 || rate <= 0 || (n > 0 && (!paths || !samples || !n_samples))) return L2S_IO_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (!paths[i] || n_samples[i] < 0 || n_samples[i] > stride) return L2S_IO_ERR_ARG;
  return run_jobs(n, threads, bad, [&](int i) { return write_one_wav(paths[i], samples + (int64_t)i * stride, n_samples[i], rate); });
}

extern "C" int l2s_io_units_to_ids(const char* const* lines, int32_t n, const char* const* dict_tokens, int32_t n_dict, int64_t* out,
                                   int64_t out_stride, const int32_t* max_ids, int32_t* n_out, int32_t threads) {
  if (n < 0 || n_dict < 0 || (n > 0 && (!lines || !out || !max_ids || !n_out)) || (n_dict > 0 && !dict_tokens)) return L2S_IO_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (!lines[i] || max_ids[i] < 0 || max_ids[i] > out_stride) return L2S_IO_ERR_ARG;
  std::unordered_map<std::string_view, int64_t> dict;
  dict.reserve((size_t)n_dict * 2 + 1);
  for (int j = 0; j < n_dict; ++j) {
    if (!dict_tokens[j]) return L2S_IO_ERR_ARG;
    dict[std::string_view(dict_tokens[j])] = j;            // a repeated token keeps its LAST index, like the dict comprehension
  }
  auto is_blank = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; };   // str.split()
  return run_jobs(n, threads, nullptr, [&](int i) {
    const char* p = lines[i];
    int64_t* dst = out + (int64_t)i * out_stride;
    int32_t count = 0;
    while (*p) {
      while (*p && is_blank(*p)) ++p;
      const char* b = p;
      while (*p && !is_blank(*p)) ++p;
      if (p == b) break;
      const auto it = dict.find(std::string_view(b, (size_t)(p - b)));
      if (it == dict.end()) continue;                       // unknown token: dropped
      if (count < max_ids[i]) dst[count] = it->second;
      ++count;
    }
    n_out[i] = count;
    return (int)L2S_IO_OK;
  });
}


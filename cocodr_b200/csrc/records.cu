// Token-record reader for the reference's EmbeddingCache files (SURVEY f-3) -- host code only (no kernels).
//
// File format (ANCE/utils/util.py:316-370; written by ANCE/data/msmarco_data.py:66-95, 268-295):
//   <base_path>      : total_number fixed-size records  [len : u32 big-endian][ids : embedding_size x int32 native]
//                      (evaluate/utils/util.py:338-369 group variant: [group : u32 BE][len : u32 BE][ids ...])
//   <base_path>_meta : {"type": "int32", "total_number": N, "embedding_size": L}   (parsed by the Python side)
// The reference does one Python seek()+read() per record (3 per training triplet); here the file is mapped once
// and a batch of records is gathered by a few threads straight into caller-owned (pinned) buffers as the padded
// int32 ids / byte mask / lengths the encoder consumes (GetProcessingFn, ANCE/data/msmarco_data.py:297-325:
// mask = [1] * len + [0] * pad).
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <thread>
#include <vector>

#include "cdr_common.cuh"

struct cdr_records {
  const uint8_t* base;
  size_t bytes;
  int64_t record_bytes, total;
  int32_t group, ids_per_record;
  int fd;
};

static inline uint32_t be32(const uint8_t* p) {
  return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) | (static_cast<uint32_t>(p[2]) << 8) |
         static_cast<uint32_t>(p[3]);
}

extern "C" {

cdr_records* cdr_records_open(const char* path, int64_t record_bytes, int64_t total, int32_t group) {
  const int64_t header = group ? 8 : 4;
  if (path == nullptr || record_bytes <= header || (record_bytes - header) % 4 != 0 || total < 0) {
    cdr::set_error("cdr_records_open: bad arguments (record_bytes=%lld total=%lld)", (long long)record_bytes, (long long)total);
    return nullptr;
  }
  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    cdr::set_error("cdr_records_open: cannot open %s", path);
    return nullptr;
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || static_cast<int64_t>(st.st_size) < record_bytes * total) {
    cdr::set_error("cdr_records_open: %s holds %lld bytes, %lld records of %lld bytes need %lld", path,
                   (long long)st.st_size, (long long)total, (long long)record_bytes, (long long)(record_bytes * total));
    close(fd);
    return nullptr;
  }
  void* m = st.st_size > 0 ? mmap(nullptr, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
  if (st.st_size > 0 && m == MAP_FAILED) {
    cdr::set_error("cdr_records_open: mmap of %s failed", path);
    close(fd);
    return nullptr;
  }
  cdr_records* r = new cdr_records;
  r->base = static_cast<const uint8_t*>(m);
  r->bytes = st.st_size;
  r->record_bytes = record_bytes;
  r->total = total;
  r->group = group;
  r->ids_per_record = static_cast<int32_t>((record_bytes - header) / 4);
  r->fd = fd;
  return r;
}

void cdr_records_close(cdr_records* r) {
  if (r == nullptr) return;
  if (r->base != nullptr) munmap(const_cast<uint8_t*>(r->base), r->bytes);
  close(r->fd);
  delete r;
}

int cdr_records_gather(const cdr_records* r, const int64_t* keys, int64_t n, int32_t max_len, int32_t* ids,
                       uint8_t* mask, int32_t* lens, int32_t* groups, int32_t n_threads) {
  CDR_REQUIRE(r != nullptr && keys != nullptr && ids != nullptr && n >= 0 && max_len > 0, "cdr_records_gather: bad arguments");
  for (int64_t i = 0; i < n; ++i)
    CDR_REQUIRE(keys[i] >= 0 && keys[i] < r->total, "cdr_records_gather: index %lld is out of bound for %lld records",
                (long long)keys[i], (long long)r->total);
  const int64_t header = r->group ? 8 : 4;
  const int32_t n_copy = r->ids_per_record < max_len ? r->ids_per_record : max_len;
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const uint8_t* rec = r->base + keys[i] * r->record_bytes;
      int32_t len = static_cast<int32_t>(be32(rec + (r->group ? 4 : 0)));
      if (len > max_len) len = max_len;
      if (groups != nullptr) groups[i] = r->group ? static_cast<int32_t>(be32(rec)) : -1;
      if (lens != nullptr) lens[i] = len;
      int32_t* dst = ids + i * max_len;
      memcpy(dst, rec + header, sizeof(int32_t) * n_copy);
      if (n_copy < max_len) memset(dst + n_copy, 0, sizeof(int32_t) * (max_len - n_copy));
      if (mask != nullptr) {
        memset(mask + i * max_len, 1, len);
        memset(mask + i * max_len + len, 0, max_len - len);
      }
    }
  };
  int t = n_threads > 0 ? n_threads : 1;
  if (t > 64) t = 64;
  if (n < 64 * t) t = 1;
  if (t == 1) {
    work(0, n);
    return CDR_OK;
  }
  std::vector<std::thread> pool;
  const int64_t per = (n + t - 1) / t;
  for (int k = 0; k < t; ++k) {
    const int64_t lo = k * per, hi = lo + per < n ? lo + per : n;
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return CDR_OK;
}

}  // extern "C"

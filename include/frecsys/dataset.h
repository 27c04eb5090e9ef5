// frecsys::Dataset — same interface and semantics as the reference (include/frecsys/dataset.h:24-99):
// a `uid,sid` CSV with a header line becomes by_user_[uid] / by_item_[sid] lists of
// (other id, tuple index) in FILE order.  In addition the tuple list is kept flat so the CUDA
// library can build its device-resident CSR/CSC from it (frx_dataset_create).
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "frecsys/logging.h"
#include "frecsys/types.h"

namespace frecsys {

// Ingest.  The reference parses with getline + substr + atoi and inserts every tuple into
// two hash maps of vectors (dataset.h:71-99): minutes at the ML-20M / MSD shapes, far longer than a GPU
// epoch.  Here the file is mmap'ed and parsed by all host threads into the flat tuple list in FILE order
// (what frx_dataset_create turns into the bit-exact device CSR/CSC); the by_user()/by_item() hash maps of
// the reference interface are built lazily, only for a caller that asks for them (EvaluateDataset's
// `eval_by_user` argument).  Line semantics are the reference's: the header line is dropped, every
// following line (also an empty one) yields one tuple, user = atoi(text before the first ','),
// item = atoi(text after it); a line without ',' gives atoi(line) for both (substr(npos + 1)).
// Binary cache (opt-in, FRECSYS_DATASET_CACHE=1: the reference never writes next to its inputs): the parsed
// tuple list is kept as `<csv>.frxbin`, keyed on the CSV's size and modification time, and read back with two
// freads on the next run (20 M tuples on 8 cores: 0.23 s instead of 1.05 s, most of the rest being the
// max / distinct scan of the summary line); a stale or damaged cache is ignored and rewritten.
class Dataset {
public:
  explicit Dataset(const std::string& filename);
  // Builds a Dataset from tuple arrays (synthetic data, tests).
  Dataset(const int* users, const int* items, int n);
  const SpMatrix& by_user() const { build_maps(); return maps_->by_user; }
  const SpMatrix& by_item() const { build_maps(); return maps_->by_item; }
  const int max_user() const { return max_user_; }
  const int max_item() const { return max_item_; }
  const int num_tuples() const { return num_tuples_; }
  // Flat tuple list in file order (users()[t], items()[t]).
  const std::vector<int>& users() const { return users_; }
  const std::vector<int>& items() const { return items_; }

private:
  struct Maps {
    std::once_flag once;
    SpMatrix by_user, by_item;
  };
  // atoi on [b, e): leading white space, optional sign, digits (no locale, no overflow handling: like atoi).
  static int parse_int(const char* b, const char* e) {
    while (b < e && (*b == ' ' || (*b >= '\t' && *b <= '\r'))) ++b;
    bool neg = false;
    if (b < e && (*b == '-' || *b == '+')) { neg = *b == '-'; ++b; }
    long v = 0;
    while (b < e && *b >= '0' && *b <= '9') { v = v * 10 + (*b - '0'); ++b; }
    return (int)(neg ? -v : v);
  }
  static void parse_range(const char* b, const char* e, std::vector<int>* us, std::vector<int>* is) {
    while (b < e) {
      const char* nl = static_cast<const char*>(memchr(b, '\n', (size_t)(e - b)));
      const char* le = nl ? nl : e;
      const char* comma = static_cast<const char*>(memchr(b, ',', (size_t)(le - b)));
      us->push_back(parse_int(b, comma ? comma : le));
      is->push_back(comma ? parse_int(comma + 1, le) : parse_int(b, le));
      b = nl ? nl + 1 : e;
    }
  }
  void parse(const char* data, size_t size);
  // binary cache of the tuple list (see the class comment)
  struct CacheHeader {
    char magic[8];
    unsigned long long csv_size;
    long long csv_mtime_ns;
    unsigned long long n;
  };
  static bool cache_enabled() {
    const char* e = std::getenv("FRECSYS_DATASET_CACHE");
    return e && *e && *e != '0';
  }
  bool load_cache(const std::string& path, const struct stat& st) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    CacheHeader h;
    bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, "FRXDS01", 8) == 0 &&
              h.csv_size == (unsigned long long)st.st_size &&
              h.csv_mtime_ns == (long long)st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec &&
              h.n <= (1ull << 31);
    if (ok) {
      users_.resize((size_t)h.n);
      items_.resize((size_t)h.n);
      ok = std::fread(users_.data(), sizeof(int), (size_t)h.n, f) == (size_t)h.n &&
           std::fread(items_.data(), sizeof(int), (size_t)h.n, f) == (size_t)h.n && std::fgetc(f) == EOF;
      if (!ok) { users_.clear(); items_.clear(); }
    }
    std::fclose(f);
    return ok;
  }
  void store_cache(const std::string& path, const struct stat& st) const {
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return;  // read-only directory: no cache
    CacheHeader h;
    std::memcpy(h.magic, "FRXDS01", 8);
    h.csv_size = (unsigned long long)st.st_size;
    h.csv_mtime_ns = (long long)st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec;
    h.n = users_.size();
    const bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 &&
                    std::fwrite(users_.data(), sizeof(int), users_.size(), f) == users_.size() &&
                    std::fwrite(items_.data(), sizeof(int), items_.size(), f) == items_.size();
    if (std::fclose(f) != 0 || !ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
  }
  void finish() {
    num_tuples_ = (int)users_.size();
    for (int t = 0; t < num_tuples_; ++t) {
      max_user_ = std::max(max_user_, users_[t]);
      max_item_ = std::max(max_item_, items_[t]);
    }
    log_summary();
  }
  void build_maps() const {  // dataset.h:86-91
    std::call_once(maps_->once, [this] {
      for (int t = 0; t < num_tuples_; ++t) {
        maps_->by_user[users_[t]].push_back({items_[t], t});
        maps_->by_item[items_[t]].push_back({users_[t], t});
      }
    });
  }
  static size_t distinct(const std::vector<int>& ids, int max_id) {
    if (max_id < 0) return 0;
    std::vector<char> seen((size_t)max_id + 1, 0);
    size_t n = 0;
    for (int v : ids)
      if (v >= 0 && !seen[v]) { seen[v] = 1; ++n; }
    return n;
  }
  void log_summary() const {  // dataset.h:94-98
    LOG(INFO) << "max_user=" << max_user() << "\tmax_item=" << max_item() << "\tdistinct user=" << distinct(users_, max_user_)
              << "\tdistinct item=" << distinct(items_, max_item_) << "\tnum_tuples=" << num_tuples();
  }
  std::shared_ptr<Maps> maps_ = std::make_shared<Maps>();
  std::vector<int> users_, items_;
  int max_user_ = -1;
  int max_item_ = -1;
  int num_tuples_ = 0;
};

inline void Dataset::parse(const char* data, size_t size) {
  // Discard the header line (the reference asserts on it, dataset.h:80).
  const char* end = data + size;
  const char* body = static_cast<const char*>(memchr(data, '\n', size));
  body = body ? body + 1 : end;
  const size_t n = (size_t)(end - body);
  unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  if (n < (1u << 20)) nt = 1;
  std::vector<const char*> cut(nt + 1, end);
  cut[0] = body;
  for (unsigned k = 1; k < nt; ++k) {  // chunk boundaries moved forward to the next line start
    const char* p = body + n * k / nt;
    if (p < cut[k - 1]) p = cut[k - 1];
    const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
    cut[k] = nl ? nl + 1 : end;
  }
  std::vector<std::vector<int>> us(nt), is(nt);
  std::vector<std::thread> th;
  for (unsigned k = 0; k < nt; ++k)
    th.emplace_back([&, k] {
      us[k].reserve((size_t)(cut[k + 1] - cut[k]) / 8 + 16);
      is[k].reserve((size_t)(cut[k + 1] - cut[k]) / 8 + 16);
      parse_range(cut[k], cut[k + 1], &us[k], &is[k]);
    });
  for (auto& t : th) t.join();
  size_t total = 0;
  for (unsigned k = 0; k < nt; ++k) total += us[k].size();
  users_.reserve(total);
  items_.reserve(total);
  for (unsigned k = 0; k < nt; ++k) {
    users_.insert(users_.end(), us[k].begin(), us[k].end());
    items_.insert(items_.end(), is[k].begin(), is[k].end());
  }
}

inline Dataset::Dataset(const std::string& filename) {
  const int fd = ::open(filename.c_str(), O_RDONLY);
  if (fd < 0) throw std::runtime_error("frecsys::Dataset: cannot read " + filename);
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); throw std::runtime_error("frecsys::Dataset: cannot stat " + filename); }
  const size_t size = (size_t)st.st_size;
  if (size == 0) { ::close(fd); throw std::runtime_error("frecsys::Dataset: cannot read " + filename); }
  const bool use_cache = cache_enabled();
  if (use_cache && load_cache(filename + ".frxbin", st)) {
    ::close(fd);
    finish();
    return;
  }
  void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (map != MAP_FAILED) {
    parse(static_cast<const char*>(map), size);
    munmap(map, size);
    ::close(fd);
  } else {  // not mappable (pipe, odd file system): read it
    ::close(fd);
    std::ifstream infile(filename, std::ios::binary);
    std::string all((std::istreambuf_iterator<char>(infile)), std::istreambuf_iterator<char>());
    parse(all.data(), all.size());
  }
  if (use_cache) store_cache(filename + ".frxbin", st);
  finish();
}

inline Dataset::Dataset(const int* users, const int* items, int n) {
  users_.assign(users, users + n);
  items_.assign(items, items + n);
  finish();
}

}  // namespace frecsys

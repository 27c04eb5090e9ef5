// dataset_dump <csv> <out.bin> [maps] — loads a `uid,sid` CSV through frecsys::Dataset (host mirror of the
// reference's dataset.h) and writes: int32 num_tuples, max_user, max_item, then users[], items[]; with `maps`
// also by_user as (row id, n, n x (item, tuple)) records sorted by row id.  Used by tests/test_ingest.py and
// to time the ingest (prints milliseconds on stderr).
#include <chrono>
#include <cstdio>
#include <map>

#include "frecsys/dataset.h"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: dataset_dump <csv> <out.bin> [maps]\n"); return 2; }
  const auto t0 = std::chrono::steady_clock::now();
  frecsys::Dataset ds(argv[1]);
  const auto t1 = std::chrono::steady_clock::now();
  fprintf(stderr, "ingest_ms=%.1f tuples=%d\n", std::chrono::duration<double, std::milli>(t1 - t0).count(), ds.num_tuples());
  FILE* f = fopen(argv[2], "wb");
  if (!f) return 1;
  const int hdr[3] = {ds.num_tuples(), ds.max_user(), ds.max_item()};
  fwrite(hdr, sizeof(int), 3, f);
  fwrite(ds.users().data(), sizeof(int), ds.users().size(), f);
  fwrite(ds.items().data(), sizeof(int), ds.items().size(), f);
  if (argc > 3) {
    std::map<int, const frecsys::SpVector*> rows;
    for (auto& kv : ds.by_user()) rows[kv.first] = &kv.second;
    for (auto& kv : rows) {
      const int rec[2] = {kv.first, (int)kv.second->size()};
      fwrite(rec, sizeof(int), 2, f);
      for (auto& pr : *kv.second) { const int e[2] = {pr.first, pr.second}; fwrite(e, sizeof(int), 2, f); }
    }
  }
  fclose(f);
  return 0;
}

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <random>
#include <vector>
#include "annb_internal.h"
static thread_local char g_err[512];
void annb_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
extern "C" const char *annb_last_error() { return g_err; }
struct Tab { const float *t; size_t ts; };
static const float *next(void *c, int64_t first, int64_t) { Tab *t = (Tab *)c; return t->t + (size_t)first * t->ts; }
int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 20000, TH = argc > 2 ? atoi(argv[2]) : 8, M = 8, Ks = 256;
  annb_index *h = new annb_index();
  h->M = M; h->Ks = Ks; h->code_bytes = 1;
  if (h->g.init(N, 16, 100, 100, M)) { puts(annb_last_error()); return 1; }
  std::mt19937 rng(1);
  std::vector<uint8_t> codes((size_t)N * M);
  for (auto &c : codes) c = rng() & 255;
  // one table per row would be N*8KB; reuse 64 random tables cyclically through a custom feed
  const int NT = 64;
  std::vector<float> tabs((size_t)NT * M * Ks);
  std::uniform_real_distribution<float> U(0.f, 4.f);
  for (auto &v : tabs) v = U(rng);
  std::vector<float> all((size_t)N * M * Ks);
  for (int i = 0; i < N; i++) memcpy(&all[(size_t)i * M * Ks], &tabs[(size_t)(i % NT) * M * Ks], sizeof(float) * M * Ks);
  std::vector<uint64_t> labels(N);
  for (int i = 0; i < N; i++) labels[i] = 1000 + i;
  Tab t{all.data(), (size_t)M * Ks};
  int rc = hnsw_insert_rows(h, codes.data(), labels.data(), N, TH, next, &t, N);
  printf("rc=%d count=%lld maxlevel=%d validate=%d\n", rc, (long long)h->g.count.load(), h->g.maxlevel, h->g.validate());
  return rc;
}

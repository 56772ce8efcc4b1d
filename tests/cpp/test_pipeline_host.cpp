// Host-only check of host/pipeline.h: the positional writer (regular file: several threads with pwrite; pipe: in order) and the
// host buffer's fallback to ordinary memory when no pinned memory can be had (no GPU).  Exit status 0 = all good.
//   usage: test_pipeline_host <scratch file>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "pipeline.h"

static std::string slurp(const char* path) {
  std::string s;
  FILE* f = fopen(path, "rb");
  if (!f) return s;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, n);
  fclose(f);
  return s;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  int bad = 0;
  // pieces of awkward sizes, some larger than the writer's 8 MB chunks, appended in several calls
  std::vector<std::string> parts;
  std::string want;
  const size_t sizes[] = {13, 0, 1, (9u << 20) + 7, 4096, (20u << 20) + 1, 5, (8u << 20)};
  unsigned x = 12345;
  for (size_t n : sizes) {
    std::string s(n, '\0');
    for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; s[i] = (char)(x >> 24); }
    parts.push_back(s);
    want += s;
  }
  {
    vc2cli::PositionalWriter w;
    if (!w.open(argv[1])) { printf("cannot open %s\n", argv[1]); return 2; }
    w.append({{parts[0].data(), parts[0].size()}, {parts[1].data(), parts[1].size()}, {parts[2].data(), parts[2].size()}});
    std::vector<vc2cli::PositionalWriter::Piece> rest;
    for (size_t i = 3; i < parts.size(); ++i) rest.push_back({parts[i].data(), parts[i].size()});
    w.append(rest, 4);
    if (!w.ok()) { printf("writer reports failure\n"); ++bad; }
    w.close();
    if (slurp(argv[1]) != want) { printf("file contents differ\n"); ++bad; }
  }
  {
    // ordinary memory when the driver cannot pin (no GPU): same interface
    vc2cli::HostBuf b(1000);
    memset(b.data(), 7, b.size());
    vc2cli::HostBuf c(std::move(b));
    if (c.size() != 1000 || c.data()[999] != 7 || b.data() != nullptr) { printf("HostBuf move\n"); ++bad; }
    c.resize(10);
    if (c.size() != 10) { printf("HostBuf resize\n"); ++bad; }
  }
  {
    vc2cli::Channel<int> q;
    q.push(3); q.push(4);
    if (q.pop() != 3 || q.pop() != 4) { printf("Channel order\n"); ++bad; }
  }
  printf("%d problem(s)\n", bad);
  return bad ? 1 : 0;
}

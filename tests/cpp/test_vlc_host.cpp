// include/vc2/VLC.h against the compiled reference (oracle/_ref/libvc2ref.so: ref_signed_vlc = SignedVLC(v).numOfBits() / code(),
// VLC.cpp:78-85), and code -> value -> code round trips through the bit buffer.  Host only.  usage: test_vlc_host libvc2ref.so
#include <dlfcn.h>
#include <cstdio>
#include <vector>
#include "vc2/VLC.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  typedef int (*vlc_f)(int, unsigned*, unsigned*);
  const vlc_f ref = (vlc_f)dlsym(h, "ref_signed_vlc");
  if (!ref) return 2;
  int bad = 0;
  std::vector<int> values;
  for (int v = -3000; v <= 3000; ++v) values.push_back(v);
  for (int v = 3001; v < 65535; v += 97) { values.push_back(v); values.push_back(-v); }
  values.push_back(65534); values.push_back(-65534);
  vc2::BitBuffer buf;
  for (int v : values) {
    unsigned n = 0, c = 0;
    ref(v, &n, &c);
    const vc2::SignedVLC s(v);
    if (s.numOfBits() != n || s.code() != c) { if (bad++ < 5) printf("SignedVLC(%d): %u/%x, reference %u/%x\n", v, s.numOfBits(), s.code(), n, c); }
    if ((int)s != v) { if (bad++ < 5) printf("SignedVLC(%d) decodes to %d\n", v, (int)s); }
    if (v >= 0) {
      const vc2::UnsignedVLC u((unsigned)v);
      if ((unsigned)u != (unsigned)v || (v > 0 && (u.numOfBits() + 1 != n || u.code() != (c >> 1)))) { if (bad++ < 5) printf("UnsignedVLC(%d)\n", v); }
      buf.put(u);
    }
    buf.put(s);
  }
  buf.align();
  vc2::BitBuffer rd(buf.bytes());
  for (int v : values) {
    if (v >= 0 && (unsigned)rd.getUnsigned() != (unsigned)v) { if (bad++ < 5) printf("read back unsigned %d\n", v); }
    if ((int)rd.getSigned() != v) { if (bad++ < 5) printf("read back signed %d\n", v); }
  }
  printf("%d problem(s) over %zu values\n", bad, values.size());
  return bad ? 1 : 0;
}

/* A/B probe for the slice coders (not part of the product, no torch, starts in a second on the GPU box):
 * loads TWO builds of the C-ABI library side by side (dlopen, RTLD_LOCAL), runs the same synthetic pictures through
 * the fused encoder of each, compares the payloads byte for byte (the first library is the one the GPU parity tests
 * last passed on) and prints the per-stage CUDA-event times of both (vc2_profile_read).
 *   build: gcc -O2 tools/pack_probe.c -o tools/_probe/pack_probe -ldl   (tools/_probe/ travels to the GPU box, gpurun_out/ does not)
 *   run:   pack_probe PICTURES baseline.so candidate.so [candidate.so ...]
 * Picture content: the generator of oracle/gen.py (SURVEY.md Appx B.3), restated in C. */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/vc2_cabi.h"

typedef struct {
  void* h;
  vc2_ctx* (*create)(int);
  void (*destroy)(vc2_ctx*);
  int (*synchronize)(vc2_ctx*);
  const char* (*last_error)(vc2_ctx*);
  int (*profile_enable)(vc2_ctx*, int);
  int (*profile_read)(vc2_ctx*, float*, int*, int);
  int (*make_geom)(int, int, int, int, int, int, int, int, int, vc2_geom*);
  vc2_codec* (*codec_create)(vc2_ctx*, const vc2_codec_params*);
  void (*codec_destroy)(vc2_codec*);
  size_t (*picture_in_bytes)(const vc2_codec*);
  size_t (*payload_capacity)(const vc2_codec*);
  int (*encode_dev)(vc2_codec*, int);
  int (*decode_dev)(vc2_codec*, int);
  int (*upload_picture)(vc2_codec*, int, const void*);
  int (*download_picture)(vc2_codec*, int, void*);
  int (*download_payload)(vc2_codec*, int, uint8_t*, size_t, size_t*, int32_t*, uint32_t*);
  int (*slot_status)(vc2_codec*, int);
} Lib;

#define SYM(l, field, name) do { *(void**)&(l)->field = dlsym((l)->h, name); if (!(l)->field) { fprintf(stderr, "missing %s\n", name); exit(2); } } while (0)
static void load(Lib* l, const char* path) {
  l->h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!l->h) { fprintf(stderr, "dlopen %s: %s\n", path, dlerror()); exit(2); }
  SYM(l, create, "vc2_create"); SYM(l, destroy, "vc2_destroy"); SYM(l, synchronize, "vc2_synchronize");
  SYM(l, last_error, "vc2_last_error"); SYM(l, profile_enable, "vc2_profile_enable"); SYM(l, profile_read, "vc2_profile_read");
  SYM(l, make_geom, "vc2_make_geom"); SYM(l, codec_create, "vc2_codec_create"); SYM(l, codec_destroy, "vc2_codec_destroy");
  SYM(l, picture_in_bytes, "vc2_codec_picture_in_bytes"); SYM(l, payload_capacity, "vc2_codec_payload_capacity");
  SYM(l, encode_dev, "vc2_codec_encode_dev"); SYM(l, decode_dev, "vc2_codec_decode_dev");
  SYM(l, upload_picture, "vc2_codec_upload_picture"); SYM(l, download_picture, "vc2_codec_download_picture");
  SYM(l, download_payload, "vc2_codec_download_payload"); SYM(l, slot_status, "vc2_codec_slot_status");
}

static uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
/* content: 0 = the bench content (gen.py), 1 = flat mid grey (every component of every slice is empty),
 * 2 = full-range noise (long codes, overflowing length bytes at small scalars) */
static void gen_plane(uint8_t* out, uint32_t seed, int c, int f, int H, int W, int depth, int content) {
  const int full = 1 << depth;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const uint32_t k = seed ^ ((uint32_t)c * 0x9E3779B1u) ^ ((uint32_t)f * 0x85EBCA77u) ^ ((uint32_t)y * 0xC2B2AE3Du) ^ ((uint32_t)x * 0x27D4EB2Fu);
      const uint32_t h = fmix32(k);
      int v;
      if (content == 1) v = full >> 1;
      else if (content == 2) v = (int)(h >> (32 - depth));
      else {
        const int noise = (int)(h >> 26) - 32;
        const int ramp = (int)((((uint32_t)x >> 1) + ((uint32_t)y >> 1) + (uint32_t)(4 * f + 37 * c)) % (uint32_t)full);
        const int chk = (int)((((uint32_t)x >> 6) ^ ((uint32_t)y >> 6)) & 1u) * (full >> 3);
        v = (ramp >> 1) + (full >> 2) + chk + (full >= 1024 ? noise * (full >> 10) : noise >> 2);
        if (v < 0) v = 0;
        if (v > full - 1) v = full - 1;
      }
      const unsigned w = (unsigned)v << (16 - depth);
      out[2 * ((size_t)y * W + x)] = (uint8_t)(w >> 8);
      out[2 * ((size_t)y * W + x) + 1] = (uint8_t)w;
    }
}

typedef struct {
  const char* name;
  int W, H, cf, bits, kernel, depth, u, a, prefix, scalar, mode, q, bytes, content, timed;
} Case;

static uint64_t fnv(const uint8_t* p, size_t n, uint64_t h) {
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001B3ull; }
  return h;
}

typedef struct { uint64_t hash, recon; size_t bytes; int status; float ms[VC2_NUM_STAGES]; int launches[VC2_NUM_STAGES]; } Result;

static Result run(Lib* l, const Case* c, int npic, uint8_t** pics, size_t pic_bytes) {
  Result r;
  memset(&r, 0, sizeof(r));
  vc2_ctx* ctx = l->create(0);
  if (!ctx) { fprintf(stderr, "vc2_create failed\n"); exit(3); }
  vc2_codec_params prm;
  memset(&prm, 0, sizeof(prm));
  if (l->make_geom(c->H, c->W, c->cf, c->kernel, c->depth, c->u, c->a, c->prefix, c->scalar, &prm.geom) != 0) { fprintf(stderr, "%s: bad geometry\n", c->name); exit(3); }
  prm.fmt.bytes_per_sample = 2; prm.fmt.luma_depth = c->bits; prm.fmt.chroma_depth = c->bits;
  prm.mode = c->mode; prm.qindex = c->q; prm.picture_bytes = c->bytes; prm.max_pictures = npic;
  vc2_codec* k = l->codec_create(ctx, &prm);
  if (!k) { fprintf(stderr, "%s: codec_create failed: %s\n", c->name, l->last_error(ctx)); exit(3); }
  if (l->picture_in_bytes(k) != pic_bytes) { fprintf(stderr, "%s: picture size mismatch\n", c->name); exit(3); }
  for (int i = 0; i < npic; ++i) l->upload_picture(k, i, pics[i]);
  l->synchronize(ctx);
  const int reps = c->timed ? 5 : 1;
  for (int i = 0; i < (c->timed ? 2 : 0); ++i) { l->encode_dev(k, npic); l->decode_dev(k, npic); }
  l->synchronize(ctx);
  l->profile_enable(ctx, 1);
  int st = 0;
  for (int i = 0; i < reps; ++i) { st = l->encode_dev(k, npic); if (st == 0) st = l->decode_dev(k, npic); }
  l->synchronize(ctx);
  l->profile_read(ctx, r.ms, r.launches, VC2_NUM_STAGES);
  for (int s = 0; s < VC2_NUM_STAGES; ++s) r.ms[s] /= (float)reps;
  l->profile_enable(ctx, 0);
  r.status = st;
  const size_t cap = l->payload_capacity(k);
  uint8_t* pay = (uint8_t*)malloc(cap);
  uint8_t* rec = (uint8_t*)malloc(pic_bytes);
  r.hash = 0xCBF29CE484222325ull; r.recon = r.hash;
  for (int i = 0; i < npic; ++i) {
    size_t len = 0;
    const int ds = l->download_payload(k, i, pay, cap, &len, NULL, NULL);
    const int ss = l->slot_status(k, i);
    r.hash = fnv((const uint8_t*)&ds, sizeof(ds), r.hash);
    r.hash = fnv((const uint8_t*)&ss, sizeof(ss), r.hash);
    if (ds == 0) {
      r.hash = fnv((const uint8_t*)&len, sizeof(len), r.hash);
      r.hash = fnv(pay, len, r.hash);
      r.bytes += len;
      if (l->download_picture(k, i, rec) == 0) { l->synchronize(ctx); r.recon = fnv(rec, pic_bytes, r.recon); }
    }
  }
  free(pay); free(rec);
  l->codec_destroy(k);
  l->destroy(ctx);
  return r;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: pack_probe pictures baseline.so candidate.so [candidate.so ...]\n"); return 2; }
  enum { MAXLIB = 6 };
  Lib L[MAXLIB];
  const int npic_timed = atoi(argv[1]);
  int nlib = 0;
  for (int i = 2; i < argc && nlib < MAXLIB; ++i) load(&L[nlib++], argv[i]);
  static const char* stage[VC2_NUM_STAGES] = {"dwt_l0", "dwt_deep", "pack", "unpack", "idwt_deep", "idwt_l0", "ld_dc", "assemble", "index", "search"};
  /* name, W, H, chroma format, bits, wavelet, depth, -u, -a, prefix, scalar, mode, q, bytes, content, timed */
  const Case cases[] = {
    {"C3 DD137 d4 q16 S4 (bench)", 3840, 2160, 1, 10, VC2_DD137, 4, 1, 2, 0, 4, VC2_HQ_VBR, 16, 0, 0, 1},
    {"C1 LeGall d3 q12", 1920, 1080, 1, 10, VC2_LEGALL, 3, 1, 2, 0, 1, VC2_HQ_VBR, 12, 0, 0, 0},
    {"C2 DD97 d3 CBR 2073600", 1920, 1080, 1, 10, VC2_DD97, 3, 1, 2, 0, 1, VC2_HQ_CBR, 0, 2073600, 0, 0},
    {"CBR prefix 3 scalar 2", 1920, 1080, 1, 10, VC2_LEGALL, 3, 1, 2, 3, 2, VC2_HQ_CBR, 0, 3000000, 0, 0},
    {"flat grey, prefix 5", 1920, 1080, 1, 10, VC2_HAAR0, 3, 1, 2, 5, 1, VC2_HQ_VBR, 4, 0, 1, 0},
    {"flat grey, prefix 2, CBR", 1920, 1080, 2, 8, VC2_HAAR1, 2, 2, 2, 2, 1, VC2_HQ_CBR, 0, 400000, 1, 0},
    {"noise q0 S8 4:4:4 12b", 1920, 1080, 0, 12, VC2_FIDELITY, 3, 1, 2, 1, 8, VC2_HQ_VBR, 0, 0, 2, 0},
    {"noise q0 S1 (scalar too small)", 1920, 1080, 1, 10, VC2_DAUB97, 3, 1, 2, 0, 1, VC2_HQ_VBR, 0, 0, 2, 0},
    {"noise q30 S1 prefix 1 4:2:0", 1920, 1080, 2, 10, VC2_DD97, 2, 2, 4, 1, 1, VC2_HQ_VBR, 30, 0, 2, 0},
    {"C5 LD LeGall d3 2073600", 1920, 1080, 1, 10, VC2_LEGALL, 3, 1, 2, 0, 1, VC2_LD, 0, 2073600, 0, 0},
    {"LD noise DD137 d2 tight", 1920, 1080, 0, 8, VC2_DD137, 2, 2, 4, 0, 1, VC2_LD, 0, 700000, 2, 0},
    {"odd slice starts: prefix 1 S1 q40", 1920, 1080, 1, 10, VC2_HAAR0, 3, 1, 2, 1, 1, VC2_HQ_VBR, 40, 0, 0, 0},
    {"big slices -u2 -a8 d2 4:4:4", 1920, 1080, 0, 10, VC2_HAAR1, 2, 2, 8, 0, 2, VC2_HQ_VBR, 8, 0, 0, 0},
    {"noise CBR tight", 1920, 1080, 1, 10, VC2_LEGALL, 3, 1, 2, 0, 1, VC2_HQ_CBR, 0, 500000, 2, 0},
  };
  int bad = 0;
  for (size_t ci = 0; ci < sizeof(cases) / sizeof(cases[0]); ++ci) {
    const Case* c = &cases[ci];
    const int npic = c->timed ? npic_timed : 2;
    const int cw = c->cf == 0 ? c->W : c->W / 2, chh = c->cf == 2 ? c->H / 2 : c->H;
    const size_t pic_bytes = 2 * ((size_t)c->W * c->H + 2 * (size_t)cw * chh);
    const int distinct = npic < 4 ? npic : 4;
    uint8_t** pics = (uint8_t**)malloc(sizeof(uint8_t*) * npic);
    for (int i = 0; i < npic; ++i) {
      if (i >= distinct) { pics[i] = pics[i % distinct]; continue; }
      pics[i] = (uint8_t*)malloc(pic_bytes);
      gen_plane(pics[i], 1234u, 0, i, c->H, c->W, c->bits, c->content);
      gen_plane(pics[i] + 2 * (size_t)c->W * c->H, 1234u, 1, i, chh, cw, c->bits, c->content);
      gen_plane(pics[i] + 2 * ((size_t)c->W * c->H + (size_t)cw * chh), 1234u, 2, i, chh, cw, c->bits, c->content);
    }
    Result r[MAXLIB];
    for (int l = 0; l < nlib; ++l) r[l] = run(&L[l], c, npic, pics, pic_bytes);
    printf("%-34s status %d  payload %zu B  hash %016llx  recon %016llx :", c->name, r[0].status, r[0].bytes,
           (unsigned long long)r[0].hash, (unsigned long long)r[0].recon);
    for (int l = 1; l < nlib; ++l) {
      const int same = r[0].hash == r[l].hash && r[0].recon == r[l].recon && r[0].bytes == r[l].bytes && r[0].status == r[l].status;
      if (!same) bad++;
      printf(" %s", same ? "SAME" : "DIFFERENT");
    }
    printf("\n");
    if (c->timed)
      for (int s = 0; s < VC2_NUM_STAGES; ++s)
        if (r[0].launches[s]) {
          printf("    %-10s %8.3f ms per %d pictures ->", stage[s], r[0].ms[s], npic);
          for (int l = 1; l < nlib; ++l) printf("  %8.3f (%+.1f %%)", r[l].ms[s], 100.0 * (r[l].ms[s] / r[0].ms[s] - 1.0));
          printf("\n");
        }
    for (int i = 0; i < distinct; ++i) free(pics[i]);
    free(pics);
  }
  printf(bad ? "%d comparison(s) DIFFER\n" : "all cases identical\n", bad);
  return bad ? 1 : 0;
}

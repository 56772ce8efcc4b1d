#!/usr/bin/env python3
"""Times the host-buffer entry points in isolation (encode only, decode only, both concurrently)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gen
import vc2_reference_b200 as vc2

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else B
w = dict(w=3840, h=2160, fmt="422", bits=10, kernel="DD137", depth=4, u=1, a=2, q=16, S=4)
ctx = vc2.Context(0); ctx2 = vc2.Context(0)
g = vc2.make_geom(w["h"], w["w"], w["fmt"], w["kernel"], w["depth"], w["u"], w["a"], 0, w["S"])
enc = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=w["q"], luma_depth=10, max_pictures=B)
dec = vc2.Codec(ctx2, g, "HQ_ConstQ", qindex=w["q"], luma_depth=10, max_pictures=B)
pin = lambda n: torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
pics = [pin(enc.picture_bytes) for _ in range(N)]
outs = [pin(enc.picture_bytes) for _ in range(N)]
pays = [pin(16 << 20) for _ in range(N)]
for i, p in enumerate(pics):
    p[:] = np.frombuffer(gen.frame_bytes(1234, i % 4, w["w"], w["h"], w["fmt"], w["bits"]), np.uint8)
lens = enc.encode_host(pics, pays)
dec.decode_host(pays, lens, outs)
def t(fn, reps=5):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
te = t(lambda: enc.encode_host(pics, pays))
td = t(lambda: dec.decode_host(pays, lens, outs))
def both():
    a = threading.Thread(target=lambda: enc.encode_host(pics, pays)); b = threading.Thread(target=lambda: dec.decode_host(pays, lens, outs))
    a.start(); b.start(); a.join(); b.join()
tb = t(both)
mb = enc.picture_bytes / 1e6
print("B=%d N=%d  encode_host %.2f ms (%.1f GB/s in)  decode_host %.2f ms (%.1f GB/s out)  both %.2f ms -> %.0f fps" % (
    B, N, te, N * mb / te, td, N * mb / td, tb, N / tb * 1e3))
# device-resident per-picture kernel time at n=1
ctx.profile_enable(True); ctx.profile_read()
for _ in range(5):
    enc.encode(1)
print({k: round(v[0] / 5, 3) for k, v in ctx.profile_read().items() if v[1]})
ctx2.profile_enable(True); ctx2.profile_read()
for i in range(B):
    dec.upload_payload(i, pays[i][:lens[i]].tobytes())
for _ in range(5):
    dec.decode(1)
print({k: round(v[0] / 5, 3) for k, v in ctx2.profile_read().items() if v[1]})

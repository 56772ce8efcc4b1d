#!/usr/bin/env python3
"""bench.py - the BASELINE.json metric on B200: HQ 2160p 4:2:2 10-bit encode + decode frames/s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of B synthetic pictures: the batch is encoded
(samples -> DWT -> quantise -> slice pack -> payload) and decoded again (payload -> parse -> dequantise
-> IDWT -> clip -> samples).  A frame counts once per round trip.  Workload = BASELINE config C3:
HQ_ConstQ 3840x2160 4:2:2 10-bit, DD 13/7 depth 4, -u 1 -a 2 -q 16 -S 4.

  value : device-resident round-trip frames/s, inputs already in HBM, CUDA-event timed, max over ranks
  e2e   : same through the host-buffer C-ABI (vc2_codec_encode_host / _decode_host): pinned host
          pictures in, host payloads out, host payloads in, host pictures out, all copies timed
  roofline : the slowest kernel of the step, its algorithmic bytes per launch / its CUDA-event time
  cpu_baseline : the unmodified reference (oracle/_ref EncodeStream + DecodeStream), one process per core,
          on a bounded sample of the same workload
Frames are sharded across ranks (intra-only codec): weak scaling, no collective on the data path.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

WORKLOAD = dict(name="C3: HQ_ConstQ 3840x2160 4:2:2 10-bit DD137 depth 4 q16 -u1 -a2 -S4", w=3840, h=2160, fmt="422", bits=10,
                kernel="DD137", depth=4, u=1, a=2, q=16, S=4, P=0, seed=1234)
METRIC = "HQ 2160p 4:2:2 10b encode+decode round-trip frames/s (C3: DD137 d4 q16)"


def ref_cmds(src, stream, dec):
    w = WORKLOAD
    enc = [os.path.join(ROOT, "oracle", "_ref", "EncodeStream"), "-m", "HQ_ConstQ", "-x", str(w["w"]), "-y", str(w["h"]), "-f", "4:2:2",
           "-l", str(w["bits"]), "-k", w["kernel"], "-d", str(w["depth"]), "-u", str(w["u"]), "-a", str(w["a"]), "-q", str(w["q"]),
           "-r", "6", "-S", str(w["S"]), src, stream]
    decc = [os.path.join(ROOT, "oracle", "_ref", "DecodeStream"), stream, dec]
    return enc, decc


def reference_round_trip(frames_per_proc, nproc, workdir):
    """One process per core, each encodes and decodes its own copy of a short clip.  Returns wall seconds."""
    import gen
    w = WORKLOAD
    src = os.path.join(workdir, "in.yuv")
    if not os.path.exists(src):
        with open(src, "wb") as f:
            for i in range(frames_per_proc):
                f.write(gen.frame_bytes(w["seed"], i, w["w"], w["h"], w["fmt"], w["bits"]))
    t0 = time.perf_counter()
    procs = []
    for p in range(nproc):
        enc, dec = ref_cmds(src, os.path.join(workdir, "s%d.vc2" % p), os.path.join(workdir, "d%d.yuv" % p))
        cmd = " ".join(enc) + " >/dev/null 2>&1 && " + " ".join(dec) + " >/dev/null 2>&1"
        if shutil.which("taskset"):
            cmd = "taskset -c %d sh -c '%s'" % (p, cmd)
        procs.append(subprocess.Popen(cmd, shell=True))
    rc = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    if any(rc):
        raise RuntimeError("reference process failed: %s" % rc)
    return dt


def have_reference():
    return all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", b)) for b in ("EncodeStream", "DecodeStream"))


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix="vc2bench_", dir=base)


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the same workload on the host cores."""
    if rank != 0:
        return
    if not have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    cores = usable_cores()
    wd = workdir()
    try:
        for _ in range(args.warmup):
            reference_round_trip(1, cores, wd)
        t = 0.0
        for _ in range(args.steps):
            t += reference_round_trip(1, cores, wd)
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    fps = cores * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "frames_per_step": cores, "note": "EncodeStream + DecodeStream, -O2, one process per core"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                         "sample": "each step: %d processes x 1 frame, encode then decode, clip in tmpfs" % cores},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        self.path = None
        if shutil.which("nvidia-smi"):
            fd, self.path = tempfile.mkstemp(prefix="vc2clk_")
            os.close(fd)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="pictures per step per GPU")
    ap.add_argument("--e2e-pictures", type=int, default=64, help="pictures per end-to-end step (host buffers) per GPU")
    ap.add_argument("--e2e-slots", type=int, default=4, help="device slots of the end-to-end codecs (pictures in flight)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import gen
    import vc2_reference_b200 as vc2

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL announces its version on standard output when the first communicator comes up: keep stdout for the one
        # JSON line by pointing file descriptor 1 at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = WORKLOAD
    B = args.batch
    ctx = vc2.Context(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    g = vc2.make_geom(w["h"], w["w"], w["fmt"], w["kernel"], w["depth"], w["u"], w["a"], w["P"], w["S"])
    codec = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=w["q"], luma_depth=w["bits"], max_pictures=B)
    S = w["w"] * w["h"] * 2                      # samples per 4:2:2 frame
    # 16 distinct synthetic frames per rank (0.7 s of host time each), dealt round robin over the B device slots: every
    # slot is its own 33 MB of HBM, so a step still streams B pictures from memory
    NDISTINCT = min(B, 16)
    frames = [np.frombuffer(gen.frame_bytes(w["seed"], rank * NDISTINCT + i, w["w"], w["h"], w["fmt"], w["bits"]), np.uint8) for i in range(NDISTINCT)]
    for i in range(B):
        codec.upload_picture(i, frames[i % NDISTINCT])
    # consecutive device-resident calls are ordered sub-batch by sub-batch (include/vc2_cabi.h): the serial slice
    # index walk that opens every HQ decode then overlaps the kernels of the other sub-batches
    codec.set_pipelined(True)

    # ---------------- device resident ----------------
    def step():
        codec.encode(B)
        codec.decode(B)

    for _ in range(max(args.warmup, 3)):
        step()
    ctx.synchronize()
    payload0, _, _ = codec.download_payload(0)
    C_bytes = sum(len(codec.download_payload(i)[0]) for i in range(NDISTINCT)) / NDISTINCT      # compressed bytes per frame
    # the round trip must reproduce the reference's decoded picture: cheap self-check = decode(encode(x)) is stable
    assert len(codec.download_picture(0)) == codec.picture_bytes

    ctx.kernel_launches(reset=True)
    clocks = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = ctx.kernel_launches(reset=True)

    # per-kernel CUDA-event times: the same steps again with the context's stage profiler on.  The profiler
    # serialises the sub-batch streams (overlapped kernels cannot be timed one by one), so the stage times
    # add up to slightly more than ms_per_step of the timed region above.
    ctx.profile_enable(True)
    ctx.profile_read()
    for _ in range(args.steps):
        step()
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # separate encode-only / decode-only timings (explain the round-trip number)
    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    enc_ms = timed(lambda: codec.encode(B), args.steps)
    dec_ms = timed(lambda: codec.decode(B), args.steps)

    # ---------------- end to end through the host-buffer C-ABI ----------------
    # One call moves E2E_N pictures through a codec that keeps E2E_SLOTS of them in flight on the device (the host
    # entry points pipeline copy-in, kernels and copy-out over the slots; a long call amortises the pipeline's fill
    # and drain).  Pictures repeat the B synthetic frames of the device-resident part.
    E2E_N, E2E_SLOTS = args.e2e_pictures, args.e2e_slots
    pin = lambda n: torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
    h_pics = [pin(codec.picture_bytes) for _ in range(E2E_N)]
    h_out = [pin(codec.picture_bytes) for _ in range(E2E_N)]
    cap = int(2.5 * C_bytes) + 4096
    h_pay = [pin(cap) for _ in range(E2E_N)]
    for i, dst in enumerate(h_pics):
        dst[:] = frames[i % NDISTINCT]

    # A two-stage host pipeline, as a transcoding application would run it: batch k+1 is encoded (thread A,
    # its own context, streams and codec) while batch k is decoded (thread B), so the H2D-heavy encode and the
    # D2H-heavy decode share the full-duplex PCIe link.  Every batch still makes the whole round trip
    # host pictures -> host payloads -> host pictures; the payload buffers are double buffered.
    import threading
    import queue
    ctx1, ctx2 = vc2.Context(local_rank), vc2.Context(local_rank)
    enc_codec = vc2.Codec(ctx1, g, "HQ_ConstQ", qindex=w["q"], luma_depth=w["bits"], max_pictures=E2E_SLOTS)
    dec_codec = vc2.Codec(ctx2, g, "HQ_ConstQ", qindex=w["q"], luma_depth=w["bits"], max_pictures=E2E_SLOTS)
    h_pay2 = [h_pay, [pin(cap) for _ in range(E2E_N)]]

    def e2e_run(nsteps):
        q = queue.Queue(maxsize=1)
        free = queue.Queue()
        free.put(0)
        free.put(1)
        out = {}

        def producer():
            for _ in range(nsteps):
                b = free.get()
                q.put((b, enc_codec.encode_host(h_pics, h_pay2[b])))
            q.put(None)

        def consumer():
            while True:
                item = q.get()
                if item is None:
                    return
                b, ln = item
                dec_codec.decode_host(h_pay2[b], ln, h_out)
                out["lens"], out["buf"] = ln, b
                free.put(b)
        ta, tb = threading.Thread(target=producer), threading.Thread(target=consumer)
        ta.start(); tb.start(); ta.join(); tb.join()
        return out["lens"], out["buf"]

    e2e_run(3)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    lens, lastbuf = e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    barrier()
    h_pay = h_pay2[lastbuf]
    assert h_out[0].tobytes() == codec.download_picture(0), "host round trip and device round trip disagree"

    # what the PCIe link gives a plain pinned copy, for scale (not part of any reported throughput)
    def copy_gbs(dst, src, nbytes):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        return 4 * nbytes / (a.elapsed_time(b) / 1000.0) / 1e9
    nb = 256 << 20
    hbuf = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    dbuf = torch.empty(nb, dtype=torch.uint8, device="cuda")
    pcie = {"h2d_gbs": copy_gbs(dbuf, hbuf, nb), "d2h_gbs": copy_gbs(hbuf, dbuf, nb)}
    # both directions at once (two streams): the ceiling of the concurrent encode + decode pipeline, which moves
    # h2d_bytes_per_step one way and d2h_bytes_per_step the other way every step
    hbuf2 = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    dbuf2 = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        with torch.cuda.stream(s1):
            dbuf.copy_(hbuf, non_blocking=True)
        with torch.cuda.stream(s2):
            hbuf2.copy_(dbuf2, non_blocking=True)
    torch.cuda.synchronize()
    pcie["bidir_total_gbs"] = 2 * 4 * nb / (time.perf_counter() - t0) / 1e9
    del hbuf, dbuf, hbuf2, dbuf2
    n_slices = g.slices_x * g.slices_y
    h2d = E2E_N * codec.picture_bytes + sum(lens)
    d2h = sum(lens) + E2E_N * codec.picture_bytes + 2 * E2E_N * 4 * n_slices + 4 * E2E_N
    # parity guard on the timed path: device-resident and host paths agree byte for byte
    assert h_pay[0][:lens[0]].tobytes() == payload0, "host path and device path disagree"

    tt = torch.tensor([ms, e2e_ms, enc_ms, dec_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_ms, enc_ms, dec_ms = [float(x) for x in tt.tolist()]

    if rank == 0:
        peak, peak_src = hbm_peak()
        fps = world * B * args.steps / (ms / 1000.0)
        e2e_fps = world * E2E_N * e2e_steps / (e2e_ms / 1000.0)
        # algorithmic bytes per frame: fused path reads 16-bit samples, keeps int32 coefficients
        alg = {
            "dwt_l0": 2 * S + 4 * S, "dwt_deep": 0.0,
            "pack": 4 * S + C_bytes, "unpack": C_bytes + 4 * S,
            "idwt_deep": 0.0, "idwt_l0": 4 * S + 2 * S, "ld_dc": 0.0, "assemble": 2 * C_bytes, "index": C_bytes,
        }
        deep = sum(8.0 * S / (4 ** l) for l in range(1, w["depth"]))
        alg["dwt_deep"] = deep
        alg["idwt_deep"] = deep
        stage = {}
        for name, (t_ms, cnt) in prof.items():
            if cnt:
                per_step = t_ms / args.steps
                stage[name] = {"ms_per_step": per_step, "launches_per_step": cnt / args.steps,
                               "gbs": alg[name] * B / (per_step / 1000.0) / 1e9 if per_step > 0 else None}
        dom = max(stage, key=lambda k: stage[k]["ms_per_step"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        ach = stage[dom]["gbs"]
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": w["name"], "frames_per_step_per_gpu": B, "distinct_frames": NDISTINCT, "parallelism": "frame-sharded x%d, no collective" % world,
                       "streams": "batch split into %s sub-batches on their own streams, pipelined across calls (vc2_codec_set_pipelined); decode rebuilds the slice index from the payload; stage times from a serialised pass" % os.environ.get("VC2_CODEC_SUBBATCH", "4"),
                       "cache": "inputs larger than L2 (%.0f MB of samples + %.0f MB of coefficients per step)" % (B * codec.picture_bytes / 1e6, B * 4 * S / 1e6),
                       "compressed_bytes_per_frame": C_bytes},
            "gpixel_per_s": fps * w["w"] * w["h"] / 1e9,
            "encode_fps": world * B / (enc_ms / 1000.0), "decode_fps": world * B / (dec_ms / 1000.0),
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "how": "vc2_codec_encode_host + vc2_codec_decode_host on pinned host buffers, %d pictures per call through %d device slots; batch k+1 encodes while batch k decodes (two host threads, two codecs)" % (E2E_N, E2E_SLOTS),
                    "pcie_pinned_copy": pcie,
                    "pictures_per_step": E2E_N, "device_slots": E2E_SLOTS,
                    "pcie_bound_fps": world * E2E_N * pcie["bidir_total_gbs"] * 1e9 / float(h2d + d2h)},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg[dom] * B /
                         max(1.0, stage[dom]["launches_per_step"])},
            "pipeline_roofline": {
                "encode_frac_int32_boundary": (4 * S + C_bytes) * B / (enc_ms / 1000.0) / 1e9 / peak,
                "decode_frac_int32_boundary": (4 * S + C_bytes) * B / (dec_ms / 1000.0) / 1e9 / peak,
                "encode_frac_packed_samples": (2 * S + C_bytes) * B / (enc_ms / 1000.0) / 1e9 / peak,
                "decode_frac_packed_samples": (2 * S + C_bytes) * B / (dec_ms / 1000.0) / 1e9 / peak,
            },
            "stages": stage,
        }
        if world == 1 and not args.no_cpu_baseline:
            if have_reference():
                cores = usable_cores()
                wd = workdir()
                try:
                    dt = reference_round_trip(1, cores, wd)
                finally:
                    shutil.rmtree(wd, ignore_errors=True)
                line["cpu_baseline"] = {"value": cores / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
                                        "sample": "%d processes x 1 frame of the same workload, EncodeStream then DecodeStream (-O2), %.1f s" % (cores, dt)}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
        print(json.dumps(line))
    dec_codec.close()
    ctx2.close()
    codec.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

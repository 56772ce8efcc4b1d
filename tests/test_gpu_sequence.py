"""GPU: BASELINE.json config 3 as a whole job - the 240-frame 2160p clip through the drop-in command lines, frame-sharded
over the GPUs of the box (-G N, ordered reassembly of the data units, DataUnit.cpp:112-123), stream and decoded md5 against
the unmodified reference's (tests/golden/sequence.json, made by tests/golden/make_sequence_golden.py).  Wall-clock frames/s
of both tools go to gpurun_out/sequence_fps.json."""
import ctypes
import hashlib
import json
import os
import subprocess
import time
from concurrent.futures import ProcessPoolExecutor

import pytest

import gen

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "vc2_reference_b200", "bin")
GOLD_PATH = os.path.join(HERE, "golden", "sequence.json")
GOLD = json.load(open(GOLD_PATH)) if os.path.exists(GOLD_PATH) else None


def _frame(f):
    c = GOLD["params"]
    return gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"], False)


def md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def device_count():
    return ctypes.CDLL(os.path.join(ROOT, "vc2_reference_b200", "libvc2b200.so")).vc2_device_count()


@pytest.fixture(scope="module")
def clip(tmp_path_factory):
    if GOLD is None:
        pytest.skip("tests/golden/sequence.json not generated")
    c = GOLD["params"]
    path = str(tmp_path_factory.mktemp("seq") / "in.yuv")
    h = hashlib.md5()
    with open(path, "wb") as out, ProcessPoolExecutor(min(32, os.cpu_count() or 4)) as ex:
        for b in ex.map(_frame, range(c["frames"]), chunksize=2):
            out.write(b)
            h.update(b)
    assert h.hexdigest() == GOLD["input_md5"]
    yield path
    os.remove(path)


def test_c3_sequence_sharded_over_the_gpus(clip, tmp_path):
    c = GOLD["params"]
    n = max(1, device_count())
    args = ["-m", c["mode"], "-x", str(c["w"]), "-y", str(c["h"]), "-f", "4:2:2", "-z", str(c["bits"]), "-k", c["kernel"], "-d", str(c["wdepth"]),
            "-u", str(c["u"]), "-a", str(c["a"]), "-r", str(c["r"]), "-q", str(c["q"]), "-S", str(c["S"]), "-P", str(c["P"])]
    report = {"frames": c["frames"], "gpus_on_box": n, "runs": []}
    # one GPU and all GPUs of the box; VC2_SEQ_GPUS="8" (a list) restricts the runs on a box that is charged per GPU
    wanted = [int(x) for x in os.environ.get("VC2_SEQ_GPUS", "").split()] or sorted({1, n})
    for g in wanted:
        stream, dec = str(tmp_path / ("s%d.vc2" % g)), str(tmp_path / ("d%d.yuv" % g))
        t = time.time()
        r = subprocess.run([os.path.join(BIN, "EncodeStream")] + args + ["-G", str(g), clip, stream], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        enc_s = time.time() - t
        assert r.returncode == 0, (r.stdout[-300:], r.stderr[-300:])
        assert os.path.getsize(stream) == GOLD["stream"]["bytes"]
        assert md5_file(stream) == GOLD["stream"]["md5"], "stream differs from the reference's (-G %d)" % g
        t = time.time()
        r = subprocess.run([os.path.join(BIN, "DecodeStream"), "-G", str(g), stream, dec], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        dec_s = time.time() - t
        assert r.returncode == 0, (r.stdout[-300:], r.stderr[-300:])
        assert md5_file(dec) == GOLD["decoded"]["md5"], "decoded pictures differ from the reference's (-G %d)" % g
        report["runs"].append({"gpus": g, "encode_wall_s": round(enc_s, 2), "encode_fps": round(c["frames"] / enc_s, 1),
                               "decode_wall_s": round(dec_s, 2), "decode_fps": round(c["frames"] / dec_s, 1)})
        os.remove(stream)
        os.remove(dec)
    report["reference"] = {"encode_s": GOLD.get("reference_encode_s"), "decode_s": GOLD.get("reference_decode_s"),
                           "where": "build container, one process (the reference is single threaded)"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "sequence_fps.json"), "w"), indent=1)
    print(json.dumps(report))

#!/usr/bin/env python3
"""DRAM bytes per launch of every stage from one `ncu --set full` capture of a whole encode + decode step
(tools/profile_step.py, VC2_CODEC_SUBBATCH=1), written as profiles/traffic_r2.json for bench.py's roofline.traffic.
usage: ncu_traffic.py report.ncu-rep CONFIG PICTURES BUILD_TAG [out.json]"""
import csv
import io
import json
import os
import subprocess
import sys

STAGE = [("dwt_tile_fwd", "dwt"), ("hq_pack_narrow", "pack"), ("hq_pack_kernel", "pack"), ("slice_scan", "assemble"), ("assemble_kernel", "assemble"),
         ("hq_index", "index"), ("slice_unpack", "unpack"), ("narrow_scale", "unpack"), ("dwt_tile_inv", "idwt"), ("ld_dc", "ld_dc")]


def main():
    rep, cfg, pictures, build = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    out_path = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic_r2.json")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 10]
    hdr, units = rows[0], rows[1]
    kn, rd, wr, tm = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    launches = []
    for r in rows[2:]:
        name = r[kn]
        stage = next((s for k, s in STAGE if k in name), None)
        if stage is None:
            continue
        b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
        launches.append([stage, b, float(r[tm]), name[:60]])
    # the finest level is the first forward and the last inverse lifting launch
    fw = [l for l in launches if l[0] == "dwt"]
    iv = [l for l in launches if l[0] == "idwt"]
    for i, l in enumerate(fw):
        l[0] = "dwt_l0" if i == 0 else "dwt_deep"
    for i, l in enumerate(iv):
        l[0] = "idwt_l0" if i == len(iv) - 1 else "idwt_deep"
    per_stage = {}
    for stage, b, t, _ in launches:
        per_stage.setdefault(stage, [0.0, 0.0])
        per_stage[stage][0] += b
        per_stage[stage][1] += t
    res = json.load(open(out_path)) if os.path.exists(out_path) else {}
    res["build"] = build
    res["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per stage and step (all launches of the stage), ncu --set full, %d pictures per launch" % pictures
    res[cfg] = {k: v[0] for k, v in per_stage.items()}
    res[cfg + "_pictures"] = pictures
    json.dump(res, open(out_path, "w"), indent=1, sort_keys=True)
    tot = sum(v[0] for v in per_stage.values())
    for k, v in per_stage.items():
        print("%-10s %8.3f GB  %8.1f MB/picture" % (k, v[0] / 1e9, v[0] / pictures / 1e6))
    enc = sum(per_stage.get(k, [0])[0] for k in ("dwt_l0", "dwt_deep", "pack", "assemble"))
    dec = sum(per_stage.get(k, [0])[0] for k in ("index", "unpack", "idwt_deep", "idwt_l0"))
    print("encode %.1f MB/picture, decode %.1f MB/picture (total %.2f GB)" % (enc / pictures / 1e6, dec / pictures / 1e6, tot / 1e9))


if __name__ == "__main__":
    main()

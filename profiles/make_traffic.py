#!/usr/bin/env python3
"""profiles/traffic.json from ncu launch lists that carry gpu__time_duration.sum + dram__bytes_{read,write}.sum
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv`) of
profiles/prof_step.py at a given size: per-launch DRAM bytes of every kernel of this library and of the WHOLE fwd+bwd
step (the last step of the run), next to the algorithmic bytes.  bench.py reads the file and reports
roofline.traffic / step_dram_bytes; it carries the hash of the kernel sources, so a stale capture is flagged.

    python profiles/make_traffic.py <launches.csv>:<B>x<T> [<launches.csv>:<B>x<T> ...]
"""
import collections
import csv
import hashlib
import io
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_hash():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "dl_speech_enhancement_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".inl")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def short_name(full):
    m = re.search(r"transform_eo_kernel<(?:\(int\))?(\d), (?:\(bool\))?(\w+), (?:\(int\))?(\d+)>", full)
    if m:
        return ("mel" if m.group(1) == "1" else "stft") + "_2048_eo" + ("" if m.group(2) in ("1", "true") else "_nograd")
    m = re.search(r"transform_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d), (?:\(bool\))?(\w+), (?:\(int\))?(\d+)(?:, (?:\(bool\))?\w+)?>", full)
    if m:
        return ("mel_" if m.group(2) == "1" else "stft_") + m.group(1) + ("" if m.group(3) in ("1", "true") else "_nograd")
    for k in ("combine_kernel", "reduce_finalize_kernel", "reduce_exchange_finalize_kernel", "reduce_kernel", "finalize_kernel"):
        if k in full:
            return k
    return None


def parse(path):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO(text)):
        d = per.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ns": 1e-9, "ms": 1e-3, "s": 1.0}.get(r["Metric Unit"].split("/")[0], 1)
    return list(per.values())


def main():
    out = {"csrc_hash": csrc_hash(), "workloads": {},
           "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none of "
                  "profiles/prof_step.py (cold caches, serialised launches); the LAST fwd+bwd step of the run"}
    for arg in sys.argv[1:]:
        path, size = arg.rsplit(":", 1)
        b, t = (int(v) for v in size.split("x"))
        launches = parse(path)
        ours = [(i, short_name(d["name"])) for i, d in enumerate(launches) if short_name(d["name"])]
        # the last step = from the last mel transform launch to the end
        starts = [i for i, n in ours if n.startswith("mel_")]
        step = launches[starts[-1]:]
        kernels, step_bytes, step_time, lib_time = {}, 0.0, 0.0, 0.0
        for d in step:
            byt = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            step_bytes += byt
            step_time += d.get("gpu__time_duration.sum", 0.0)
            n = short_name(d["name"])
            if n:
                lib_time += d.get("gpu__time_duration.sum", 0.0)
                k = kernels.setdefault(n, {"launches": 0, "dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0, "time_us": 0.0})
                k["launches"] += 1
                k["dram_bytes"] += byt
                k["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
                k["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
                k["time_us"] += d.get("gpu__time_duration.sum", 0.0) * 1e6
        for k in kernels.values():
            k["dram_bytes_per_launch"] = k["dram_bytes"] / k["launches"]
            k["time_us_per_launch"] = k["time_us"] / k["launches"]
        alias = {"mel_2048_eo": "mel_2048_hop300", "mel_2048": "mel_2048_hop300", "stft_2048": "stft_2048_hop240",
                 "stft_2048_eo": "stft_2048_hop240", "stft_1024": "stft_1024_hop120", "stft_512": "stft_512_hop50"}
        named = {alias.get(n, n): dict(v, kernel=n) for n, v in kernels.items()}
        out["workloads"][f"{b}x{t}"] = {
            "kernels": named, "step_dram_bytes": step_bytes, "step_algorithmic_bytes": 20.0 * b * t,
            "step_dram_over_algorithmic": step_bytes / (20.0 * b * t),
            "transform_algorithmic_bytes_per_launch": 8.0 * b * t,
            "step_kernel_time_us_serialised": step_time * 1e6, "library_share_of_step_time": lib_time / step_time,
            "source": os.path.relpath(path, ROOT)}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    for w, e in out["workloads"].items():
        print(w, f"step DRAM {e['step_dram_bytes'] / 1e6:.0f} MB = {e['step_dram_over_algorithmic']:.1f} x algorithmic; library kernels "
                 f"{100 * e['library_share_of_step_time']:.0f} % of the step's kernel time")
        for n, k in e["kernels"].items():
            print(f"   {n:24s} {k['launches']} x {k['time_us_per_launch']:8.1f} us  DRAM {k['dram_bytes_per_launch'] / 1e6:8.1f} MB/launch "
                  f"(R {k['dram_read'] / k['launches'] / 1e6:.1f} W {k['dram_write'] / k['launches'] / 1e6:.1f})")


if __name__ == "__main__":
    main()

"""Turns the ncu exports of a GPU session into the committed summaries under profiles/ (round 2).

    python tools/ncu_summarise.py gpurun_out/r02_ncu_full_raw.csv gpurun_out/r02_ncu_launches_bench.csv

  r02_ncu_full_raw.csv        `ncu -i <rep> --page raw --csv` of `ncu --set full --clock-control none python tools/ncu_all.py`
  r02_ncu_launches_bench.csv  `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ... python bench.py
                               --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained`
Writes profiles/r02_ncu_full_summary.txt, profiles/r02_ncu_traffic.json (read by bench.py as roofline.traffic) and
profiles/r02_ncu_launch_shares.txt, and copies the launch list.
"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("fire::", "")


def full_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    H = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        try:
            return float(r[H[name]].replace(",", "")) * SCALE.get(units[H[name]], 1.0)
        except ValueError:
            return float("nan")
    out = []
    for r in data:
        dur = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        out.append(dict(kernel=short(r[H["Kernel Name"]]), grid=r[H["launch__grid_size"]], block=r[H["launch__block_size"]],
                        cluster=r[H["launch__cluster_size"]], regs=r[H["launch__registers_per_thread"]], us=dur * 1e6, rd=rd, wr=wr,
                        gbps=(rd + wr) / dur / 1e9 if dur > 0 else 0.0,
                        dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        tensor_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        warps_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
    ours = [o for o in out if not o["kernel"].startswith(("at::", "<unnamed>"))]           # drop torch's setup kernels (randn, arange)
    # launch order of tools/ncu_all.py: K1 boxes, K1 copy, 33 forward launches (+ L2 norm), enrol, then 4 searches of 7 kernels.
    # The two pool branches are launched (on the side stream) BEFORE the convolutions they run beside.
    assert ours[0]["kernel"].startswith("preprocess_reference") and ours[1]["kernel"].startswith("preprocess_reference"), ours[0]
    labels = {0: "K1 configs[4]-sized boxes (256 boxes, 48..400 px, from 1080p frames)", 1: "K1 256 x 160x160 crops (copy case)"}
    fwd = ["Conv2d_1a (s2d 2x2, pixel pairs)", "Conv2d_2a (pixel pairs)", "Conv2d_2b", "MaxPool_3a + Conv2d_3b (one launch)", "Conv2d_4a", "Conv2d_4b", "Block35 x5 (fused chain)",
           "Mixed_6a pool", "Mixed_6a b0 3x3/2", "Mixed_6a b1 1x1", "Mixed_6a b1 3x3", "Mixed_6a b1 3x3/2", "Block17 x10 (fused chain)",
           "Mixed_7a pool", "Mixed_7a heads", "Mixed_7a b0 3x3/2", "Mixed_7a b1 3x3/2", "Mixed_7a b2 3x3", "Mixed_7a b2 3x3/2"]
    fwd += [f"Block8_{b + 1} {n}" for b in range(6) for n in ("heads", "1x3 + 3x1 + up" + (" + average pool" if b == 5 else ""))] + ["Bottleneck", "L2 norm"]
    assert ours[2 + fwd.index("Mixed_6a pool")]["kernel"].startswith("maxpool") and ours[2 + fwd.index("Block8_1 1x3 + 3x1 + up")]["kernel"].startswith("block8_fused"), \
        "launch order of the forward changed: update the label list"
    for i, name in enumerate(fwd):
        labels[2 + i] = name
    base = 2 + len(fwd)
    labels[base] = "kNN enrol 1M rows (normalise)"
    knn_names = ["normalise queries", "scan", "rerank", "refine", "exact scan", "exact merge", "overflow"]
    for qi, q in enumerate((4096, 1, 32, 256)):
        for j, n in enumerate(knn_names):
            labels[base + 1 + qi * 7 + j] = f"kNN Q={q}: {n}"
    lines = ["# ncu --set full --clock-control none  python tools/ncu_all.py   (B200, round 2; one launch of every kernel of the hot path)",
             "# Cold-cache, serialised replays: DRAM bytes are upper bounds for the warm step (ncu flushes L2 between replays), times are not step times.",
             "# dram% = gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed, tensor% = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,",
             "# warps% = sm__warps_active.avg.pct_of_peak_sustained_active.  Raw export (2387 metrics per launch): gpurun_out/r02_ncu_full_raw.csv, not committed.",
             f"# {'what':50s} {'kernel':28s} {'grid':>6s} {'blk':>4s} {'cl':>2s} {'regs':>4s} {'us':>9s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>6s} {'dram%':>6s} {'tensor%':>7s} {'warps%':>6s}"]
    for i, o in enumerate(ours):
        lines.append(f"  {labels.get(i, ''):50s} {o['kernel'][:28]:28s} {o['grid']:>6s} {o['block']:>4s} {o['cluster']:>2s} {o['regs']:>4s} {o['us']:9.1f} "
                     f"{o['rd'] / 1e6:8.1f} {o['wr'] / 1e6:8.1f} {o['gbps']:6.0f} {o['dram_pct']:6.1f} {o['tensor_pct']:7.1f} {o['warps_pct']:6.1f}")
    with open(os.path.join(ROOT, "profiles", "r02_ncu_full_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    by = {labels.get(i, str(i)): o for i, o in enumerate(ours)}
    tensor = [o for i, o in enumerate(ours) if 2 <= i < base and any(s in o["kernel"] for s in ("conv_strip", "conv_igemm", "block35", "block17", "block8", "pool_conv"))]
    tot = sum(o["rd"] + o["wr"] for o in tensor)
    traffic = {"source": "ncu --set full --clock-control none, python tools/ncu_all.py (one launch of every kernel; cold caches: ncu flushes L2 between "
                         "replays, so these are upper bounds for the warm step), profiles/r02_ncu_full_summary.txt",
               "conv_family_launches": len(tensor), "conv_family_dram_bytes_per_step": tot, "conv_family_dram_bytes_per_launch_avg": tot / len(tensor),
               "conv_family_us_serialised_cold": sum(o["us"] for o in tensor),
               "k1_copy_case_dram_bytes": by[labels[1]]["rd"] + by[labels[1]]["wr"], "k1_configs4_boxes_dram_bytes": by[labels[0]]["rd"] + by[labels[0]]["wr"],
               "maxpool_3a_conv2d_3b_dram_bytes": by["MaxPool_3a + Conv2d_3b (one launch)"]["rd"] + by["MaxPool_3a + Conv2d_3b (one launch)"]["wr"],
               "knn_scan_dram_bytes_per_launch": by["kNN Q=4096: scan"]["rd"] + by["kNN Q=4096: scan"]["wr"],
               "knn_scan_q1_dram_bytes": by["kNN Q=1: scan"]["rd"] + by["kNN Q=1: scan"]["wr"],
               "knn_scan_q32_dram_bytes": by["kNN Q=32: scan"]["rd"] + by["kNN Q=32: scan"]["wr"],
               "knn_scan_q256_dram_bytes": by["kNN Q=256: scan"]["rd"] + by["kNN Q=256: scan"]["wr"]}
    with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    return len(ours)


def launch_shares(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    H = {h: i for i, h in enumerate(rows[0])}
    data = rows[1:]
    names = [short(r[H["Kernel Name"]]) for r in data]
    vals = [float(r[H["Metric Value"]].replace(",", "")) for r in data]
    starts = [i for i, n in enumerate(names) if "preprocess_reference" in n]
    s0, s1 = starts[3], starts[5]                       # steps 4 and 5 of the command = its two TIMED steps (3 warm-ups before)
    fam = collections.OrderedDict()
    for n, v in zip(names[s0:s1], vals[s0:s1]):
        d = fam.setdefault(n, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(vals[s0:s1])
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: `python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained`",
           f"# the two TIMED steps of that command (launches {s0}..{s1 - 1} of the capture; {(s1 - s0) // 2} launches per step: K1 + 30 tensor launches + 2 max-pools + L2 norm).",
           f"# Cold-cache, serialised per-launch times: compare SHARES, not absolutes.  unit: ns, summed over the two steps ({tot / 2 / 1000:.1f} us per step under ncu).",
           "# Full list: profiles/r02_ncu_launches_bench.csv"]
    for n, (c, v) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n:40s} launches {c // 2:3d}/step  total {v:12.1f}  share {v / tot:6.3f}")
    tens = sum(v for n, (c, v) in fam.items() if any(s in n for s in ("conv_", "block")))       # pool_conv_fused_kernel included
    out.append(f"# tensor-kernel family (conv_igemm + conv_strip + block35_fused + block17_fused + block8_fused + pool_conv_fused): share {tens / tot:.3f} of the step under ncu")
    with open(os.path.join(ROOT, "profiles", "r02_ncu_launch_shares.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    shutil.copy(path, os.path.join(ROOT, "profiles", "r02_ncu_launches_bench.csv"))
    return tens / tot


if __name__ == "__main__":
    print("launches summarised:", full_summary(sys.argv[1]))
    if len(sys.argv) > 2:
        print("tensor share:", launch_shares(sys.argv[2]))

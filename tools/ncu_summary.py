"""Summarise ncu artefacts from gpurun_out/ into profiles/ (text that is committed; the .ncu-rep stay scratch).
usage: python tools/ncu_summary.py launches <launches.csv> <out.md>
       python tools/ncu_summary.py full <report.ncu-rep> <out.md>"""
import collections, csv, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"])})
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"] == "gpu__time_duration.sum":
            u = row["Metric Unit"]
            d["us"] = v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)
        else:
            d[row["Metric Name"]] = v
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in per.values():
        agg[d["name"]][0] += 1
        agg[d["name"]][1] += d["us"]
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path}): one eager batched step (8 frames), cold-cache serialised launches\n\n")
        f.write(f"total {tot:.1f} us over {len(per)} launches\n\n| kernel | launches | us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1]/tot:.3f} |\n")
    print(open(out).read()[:1500])


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {path}\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')][:110]}\n\n")
            for k in KEYS:
                hits = [i for i, h in enumerate(hdr) if h == k]
                if hits:
                    f.write(f"- {k}: {r[hits[0]]} {units[hits[0]]}\n")
            f.write("\n")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    with open(out, "a") as f:
        for b in blocks[:1]:
            h = b["rows"][0]
            col = {x: i for i, x in enumerate(h)}
            stalls = [x for x in h if x.startswith("stall_") and "Not" not in x]
            data = []
            for r in b["rows"][1:]:
                try:
                    data.append((int(r[col["# Samples"]]), r))
                except Exception:
                    pass
            tot = sum(d[0] for d in data) or 1
            f.write(f"## top stall sites (first kernel, {tot} samples)\n\n| samples | share | SASS | dominant stall |\n|---|---|---|---|\n")
            for n, r in sorted(data, key=lambda d: -d[0])[:12]:
                top = max(((int(r[col[s]] or 0), s) for s in stalls), default=(0, ""))
                f.write(f"| {n} | {n/tot:.3f} | `{r[col['Source']][:70]}` | {top[1]} |\n")
    print(open(out).read()[:2500])


def family(name):
    """kernel name -> the op family bench.py reports"""
    m = re.search(r"gemm_tc_kernel<\(?(?:int\))?\s*(\d+),\s*\(?(?:bool\))?\s*(\d)", name)
    if m:
        return "cofi_conv2d_nhwc" if m.group(2) == "1" else "cofi_gemm*"
    if "gemm_simt_kernel" in name:
        return "cofi_conv2d_nhwc" if "ConvA" in name else "cofi_gemm*"
    for key, fam in (("kpconv_aggregate", "cofi_kpconv_aggregate*"), ("attention_tc", "cofi_attention_vt"),
                     ("attention_simt", "cofi_attention"), ("maxpool_rows_f16", "cofi_maxpool_rows_f16"),
                     ("maxpool_rows", "cofi_maxpool_rows"), ("norm_", "cofi_norm_rows*"), ("sim_argmin", "cofi_sim_argmin")):
        if key in name:
            return fam
    return re.sub(r"_kernel.*", "", name.split("::")[-1])


def traffic(path, out):
    """per-family average DRAM bytes per launch -> profiles/traffic.json (read by bench.py)"""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"])})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            d["bytes"] = d.get("bytes", 0.0) + v * mult
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in per.values():
        f = family(d["name"])
        agg[f][0] += 1
        agg[f][1] += d.get("bytes", 0.0)
    res = {k: v[1] / v[0] for k, v in agg.items()}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    for k, v in sorted(res.items(), key=lambda kv: -kv[1] * agg[kv[0]][0]):
        print(f"{k:32s} launches={agg[k][0]:4d} avg_dram_MB_per_launch={v/1e6:9.2f} total_GB={v*agg[k][0]/1e9:7.2f}")


def traffic_calls(path, calls_json, out):
    """DRAM bytes per launch averaged over EXACTLY the launches of each op family bench.py reports (incl. the hbm- / tensor-
    bound split of the contraction entry points): the ncu kernel list of one eager step is cut into C-ABI calls with the
    per-call kernel counts that tools/profile_step.py --emit-calls recorded for the same step."""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"]), "bytes": 0.0, "us": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"].startswith("dram__bytes"):
            d["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif row["Metric Name"] == "gpu__time_duration.sum":
            d["us"] += v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)
    kernels = [k for k in per.values() if "at::" not in k["name"]]   # torch's own helpers (one cat) are not C-ABI calls
    calls = json.load(open(calls_json))
    assert sum(n for _, n in calls) == len(kernels), (sum(n for _, n in calls), len(kernels))
    agg = collections.OrderedDict()
    i = 0
    for name, n in calls:
        base, _, bound = name.partition("|")
        key = "cofi_gemm*" if base.startswith("cofi_gemm") else ("cofi_kpconv_aggregate*" if base.startswith("cofi_kpconv_aggregate") else base)
        if bound:
            key += " [" + bound + "-bound calls]"
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += sum(k["bytes"] for k in kernels[i:i + n])
        a[2] += sum(k["us"] for k in kernels[i:i + n])
        i += n
    res = {k: v[1] / v[0] for k, v in agg.items()}
    res["_meta"] = {"what": "ncu dram__bytes_read.sum + dram__bytes_write.sum per C-ABI call, averaged per op family of one eager "
                            "8-frame step (cold-cache serialised launches)", "launches": {k: v[0] for k, v in agg.items()},
                    "ncu_us_per_family": {k: round(v[2], 1) for k, v in agg.items()}}
    json.dump(res, open(out, "w"), indent=1)
    tot = sum(v[2] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"{k:44s} calls={v[0]:4d} avg_dram_MB={v[1]/v[0]/1e6:9.2f} total_GB={v[1]/1e9:7.2f} ncu_us={v[2]:9.1f} share={v[2]/tot:.3f}")


if __name__ == "__main__":
    if sys.argv[1] == "traffic_calls":
        traffic_calls(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])

"""Key metrics of every kernel in an `ncu --set full` report, as text for profiles/ and (with
--traffic out.json --config cfg3 --commit HASH) the DRAM bytes per launch that bench.py quotes
in `roofline.traffic`.  Usage: python tools/ncu_summary.py report.ncu-rep [--traffic profiles/ncu_traffic.json ...]"""
import argparse, collections, csv, io, json, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--traffic")
ap.add_argument("--config", default="cfg3")
ap.add_argument("--commit", default="")
ap.add_argument("--gpus", type=int, default=1)
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
traffic = collections.OrderedDict()
for r in rows[2:]:
    if len(r) < len(hdr) - 3:
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    print("###", name)
    for k in KEYS:
        if k in ix:
            print(f"    {k:75s} {r[ix[k]]} {units[ix[k]]}")
    stalls = sorted(((float(r[i] or 0), h.split("issue_stalled_")[1].split("_per_")[0]) for h, i in ix.items()
                     if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h), reverse=True)[:6]
    tot = sum(v for v, _ in stalls) or 1.0
    print("    top stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls))
    try:
        b = float(r[ix["dram__bytes_read.sum"]]) * UNIT[units[ix["dram__bytes_read.sum"]]] + \
            float(r[ix["dram__bytes_write.sum"]]) * UNIT[units[ix["dram__bytes_write.sum"]]]
        traffic.setdefault(name.replace("void ", "").split("<")[0], []).append(b)
    except Exception:
        pass
if a.traffic:
    try:
        j = json.load(open(a.traffic))
    except Exception:
        j = {"kernels": {}}
    j.update({"config": a.config, "gpus": a.gpus, "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, "
              "mean over the captured launches (ncu --set full --clock-control none)"})
    for k, v in traffic.items():
        j["kernels"][k] = sum(v) / len(v)
        j.setdefault("captured_on_commit", {})[k] = a.commit
    json.dump(j, open(a.traffic, "w"), indent=1)

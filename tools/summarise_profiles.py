"""Turn the raw artefacts a GPU run left in gpurun_out/ (tools/collect_profiles.sh <tag>) into the small, tracked summaries
under profiles/.  Runs in the build container (needs `ncu` to read .ncu-rep files, no GPU).

    python tools/summarise_profiles.py r20 r1      # gpurun_out/r20_* -> profiles/r1_*
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag, out = sys.argv[1], sys.argv[2]
G = lambda n: os.path.join(ROOT, "gpurun_out", f"{tag}_{n}")
P = lambda n: os.path.join(ROOT, "profiles", f"{out}_{n}")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)


def short(name):
    name = name.split("(")[0].replace("void ", "").replace("dcu::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


# ---- 1. launch list: per-kernel totals, share of the step, DRAM traffic -------------------------------------------------
rows = [l for l in open(G("launches.csv")) if not l.startswith("==")]
per = collections.OrderedDict()
launches = collections.defaultdict(dict)
for r in csv.DictReader(rows):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # -> us
    else:
        v = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1) * v  # -> bytes
    launches[r["ID"]][r["Metric Name"]] = v
    launches[r["ID"]]["name"] = short(r["Kernel Name"])
    launches[r["ID"]]["grid"] = r["Grid Size"]
agg = collections.defaultdict(lambda: dict(n=0, us=0.0, rd=0.0, wr=0.0))
for k, d in launches.items():
    a = agg[d["name"]]
    a["n"] += 1; a["us"] += d.get("gpu__time_duration.sum", 0); a["rd"] += d.get("dram__bytes_read.sum", 0); a["wr"] += d.get("dram__bytes_write.sum", 0)
tot = sum(a["us"] for a in agg.values())
with open(P("launch_list.txt"), "w") as f:
    f.write(f"ncu launch list of ONE step of the bench workload (batch 256 x 320x240, full pipeline), tools/profile_step.py\n"
            f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (cold-cache, serialised:\n"
            f"compare SHARES).  {len(launches)} launches, {tot/1e3:.2f} ms summed kernel time.\n\n")
    f.write(f"{'kernel':34s} {'launches':>8s} {'total ms':>9s} {'share':>7s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'avg us':>8s}\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        f.write(f"{name:34s} {a['n']:8d} {a['us']/1e3:9.3f} {100*a['us']/tot:6.1f}% {a['rd']/1e6:11.1f} {a['wr']/1e6:11.1f} {a['us']/a['n']:8.1f}\n")
    f.write("\nfirst launches in order (name, grid, us, DRAM read MB, DRAM write MB):\n")
    for k in list(launches)[:48]:
        d = launches[k]
        f.write(f"  {d['name']:30s} {d['grid']:>14s} {d.get('gpu__time_duration.sum',0):9.1f} {d.get('dram__bytes_read.sum',0)/1e6:9.1f} {d.get('dram__bytes_write.sum',0)/1e6:9.1f}\n")
conv = [a for n, a in agg.items() if "conv3x3_tc" in n or "conv_tc2" in n]      # single-CTA (1x1 heads) + CTA-pair kernels
summary = dict(step_kernel_ms=tot / 1e3, launches=len(launches),
               conv3x3_tc=dict(launches=sum(a["n"] for a in conv), ms=sum(a["us"] for a in conv) / 1e3, share=sum(a["us"] for a in conv) / tot,
                               dram_bytes_per_launch=sum(a["rd"] + a["wr"] for a in conv) / max(1, sum(a["n"] for a in conv))),
               kernels={n: dict(launches=a["n"], ms=a["us"] / 1e3, share=a["us"] / tot, dram_read_mb=a["rd"] / 1e6, dram_write_mb=a["wr"] / 1e6)
                        for n, a in agg.items()})
json.dump(summary, open(P("launch_summary.json"), "w"), indent=1)

# ---- 2. full captures: key metrics per captured launch --------------------------------------------------------------------
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"]
for rep, name in (("conv_tc.ncu-rep", "ncu_conv_tc2.txt"), ("small.ncu-rep", "ncu_small_kernels.txt"), ("conv1b.ncu-rep", "ncu_conv1b.txt"),
                  ("step_full_raw.csv", "ncu_step_full.txt")):
    if not os.path.isfile(G(rep)):
        continue
    if rep.endswith(".csv"):          # raw page exported on the GPU box (the report itself stays there)
        raw = open(G(rep)).read()
    else:
        raw = subprocess.run(["ncu", "-i", G(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units = rr[0], rr[1]
    with open(P(name), "w") as f:
        f.write(f"ncu --set full --clock-control none ({tag}); one block per captured launch.\n"
                f"Workload: ONE step of the bench workload (tools/profile_step.py --batch 256: 256 frames 320x240, engine defaults), profiled after a warm step.\n\n")
        if rep.endswith(".csv"):
            # per-instantiation table: launches, total time, tensor-pipe and DRAM utilisation (time-weighted)
            tab = collections.OrderedDict()
            for r in rr[2:]:
                d = dict(zip(hdr, r))
                def num(k):
                    try:
                        return float(d.get(k, "").replace(",", ""))
                    except ValueError:
                        return 0.0
                t = num("gpu__time_duration.sum")
                t = t / 1e3 if units[hdr.index("gpu__time_duration.sum")] == "ns" else (t * 1e3 if units[hdr.index("gpu__time_duration.sum")] == "ms" else t)
                a = tab.setdefault(short(d.get("Kernel Name", "")), dict(n=0, us=0.0, tens=0.0, dram=0.0, l1=0.0))
                a["n"] += 1; a["us"] += t
                a["tens"] += t * num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
                a["dram"] += t * num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
                a["l1"] += t * num("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
            f.write(f"{'kernel instantiation':44s} {'launches':>8s} {'total us':>10s} {'tensor pipe %':>14s} {'DRAM %':>8s} {'smem operand fetch %':>21s}\n")
            for k, a in sorted(tab.items(), key=lambda kv: -kv[1]["us"]):
                u = max(a["us"], 1e-9)
                f.write(f"{k[:44]:44s} {a['n']:8d} {a['us']:10.1f} {a['tens']/u:14.1f} {a['dram']/u:8.1f} {a['l1']/u:21.1f}\n")
            f.write("\n")
        for r in rr[2:]:
            d = dict(zip(hdr, r))
            f.write(f"== {short(d.get('Kernel Name',''))}   grid {d.get('Grid Size','')}  block {d.get('Block Size','')}\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write(f"   {k:88s} {d[k]:>16s} {units[hdr.index(k)]}\n")
            f.write("\n")
# ---- 3. copy the small text/JSON artefacts -------------------------------------------------------------------------------
for src, dst in (("bench.json", "bench.json"), ("bench_reference.json", "bench_reference.json"), ("parity.json", "parity.json"),
                 ("bench_strict.json", "bench_strict_accumulation.json"), ("parity_strict.json", "parity_strict_accumulation.json"),
                 ("parity_640.json", "parity_640x480.json"), ("layer_table.json", "layer_table.json"),
                 ("sanitize_memcheck.log", "sanitize_memcheck.txt"), ("sanitize_racecheck.log", "sanitize_racecheck.txt"),
                 ("sanitize_synccheck.log", "sanitize_synccheck.txt"), ("mma_probe.json", "mma_probe.json"), ("mma_probe.log", "mma_probe.txt"),
                 ("tcstats.log", "tc_role_cycles.txt"), ("tc_vs_ffma.log", "tc_vs_ffma.txt"), ("pytest.log", "pytest_gpu.txt"),
                 ("smoke.log", "smoke.txt")):
    if os.path.isfile(G(src)):
        shutil.copy(G(src), P(dst))
# clocks: one-line summary
if os.path.isfile(G("clocks.csv")):
    sm, pw, reasons = [], [], set()
    for i, l in enumerate(open(G("clocks.csv"))):
        f_ = [x.strip() for x in l.split(",")]
        if i == 0 or len(f_) < 9:
            continue
        try:
            sm.append(float(f_[1].split()[0])); pw.append(float(f_[3].split()[0]))
        except ValueError:
            continue
        for nme, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f_[5:9]):
            if v.startswith("Active"):
                reasons.add(nme)
    import statistics
    busy = [s for s, p in zip(sm, pw) if p > 300]
    open(P("clocks.txt"), "w").write(
        f"nvidia-smi during bench.py --steps 20 --warmup 5 ({tag}): {len(sm)} samples @200 ms; under load (>300 W) {len(busy)} samples, "
        f"SM clock median {statistics.median(busy) if busy else 'n/a'} MHz (max {max(sm) if sm else 'n/a'}), power max {max(pw) if pw else 'n/a'} W, "
        f"reasons seen: {sorted(reasons) or 'none'}\n")
print(open(P("launch_list.txt")).read()[:3000])

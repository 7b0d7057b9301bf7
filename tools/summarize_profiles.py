"""Turn ncu artefacts (gpurun_out/) into the small tracked summaries under profiles/."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def launch_list(path, out_md, title):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        try: v = float(row["Metric Value"].replace(",", ""))
        except Exception: continue
        unit = row["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else v * (1000.0 if unit in ("ms", "msecond") else 1.0)
        key = (row["Kernel Name"], row["Grid Size"], row["Block Size"])
        agg.setdefault(key, []).append(us); tot += us
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(path, ROOT)}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | grid | block | launches | mean us | share of listed time |\n|---|---|---|---|---|---|\n")
        for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:90]}` | {g} | {b} | {len(v)} | {sum(v)/len(v):.2f} | {100*sum(v)/tot:.1f}% |\n")

def full_report(rep, out_md, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(rep, ROOT)}` (`ncu --set full --clock-control none --import-source on`), one row per captured launch.\n\n")
        f.write("| " + " | ".join(f"{hdr[i]} ({units[i]})" if units[i] else hdr[i] for i in idx) + " |\n|" + "---|" * len(idx) + "\n")
        for row in r[2:]:
            f.write("| " + " | ".join(row[i][:70] for i in idx) + " |\n")

if __name__ == "__main__" and len(sys.argv) == 1:
    g = os.path.join(ROOT, "gpurun_out"); p = os.path.join(ROOT, "profiles")
    os.makedirs(p, exist_ok=True)
    for src, dst, title in [("r1/launches_multikernel.csv", "r1_launches_multikernel.md", "Round 1 - launch list, multi-kernel graph path, Nano-168M Q80 seq 512"),
                            ("launches_r1.csv", "r1_launches_first_version.md", "Round 1 - launch list of the first working version (before fusion fixes)")]:
        if os.path.exists(os.path.join(g, src)): launch_list(os.path.join(g, src), os.path.join(p, dst), title)
    for src, dst, title in [("r1/prof_multikernel.ncu-rep", "r1_ncu_full_multikernel.md", "Round 1 - ncu --set full, multi-kernel path kernels (Nano-168M Q80)"),
                            ("prof_mega_v1.ncu-rep", "r1_ncu_full_megakernel.md", "Round 1 - ncu --set full, persistent megakernel k_decode_mega (one launch = 496 tokens, under profiler)")]:
        if os.path.exists(os.path.join(g, src)): full_report(os.path.join(g, src), os.path.join(p, dst), title)
    for b in os.listdir(os.path.join(g, "r1")) if os.path.isdir(os.path.join(g, "r1")) else []:
        if b.startswith("bench_") and b.endswith(".json"):
            open(os.path.join(p, "r1_" + b), "w").write(open(os.path.join(g, "r1", b)).read())


def paths_table(out_md):
    """profiles/r1_paths.md: one row per committed bench line (profiles/r1_bench_*.json)."""
    p = os.path.join(ROOT, "profiles")
    rows = []
    for f in sorted(os.listdir(p)):
        if not (f.startswith("r1_bench_") and f.endswith(".json")): continue
        try: d = json.loads(open(os.path.join(p, f)).read().strip().splitlines()[-1])
        except Exception: continue
        if d.get("impl") == "reference": continue
        cfg = d["config"]; tr = d.get("token_roofline") or {}
        cb = d.get("cpu_baseline") or {}
        gbs = tr.get("achieved_gbs_per_gpu", tr.get("achieved_gbs"))
        rows.append((f, cfg["workload"].split(" greedy")[0], d["n_gpus"], ("tp" if cfg.get("parallelism", "").startswith("tp") else "replicas"), cfg.get("mode", ""), cfg["engine"].split(" (")[0],
                     d["value"], d["e2e"]["value"], gbs, tr.get("frac_of_peak"), cb.get("value"), d.get("gpu_launches")))
    with open(out_md, "w") as f:
        f.write("# Round 1 - every measured path (B200, bench.py JSON lines committed beside this file)\n\n")
        f.write("`value` = device-resident greedy decode tok/s (CUDA events, decode segment); `e2e` = per-token C-ABI calls with host buffers; "
                "GB/s = algorithmic bytes per token x tok/s per GPU; frac = that / MEASURED_PEAKS.json hbm_gbs; CPU = the unmodified reference on the box's host cores (best thread count).\n\n"
                "The JSON lines were collected over the round; `r1_bench_n168_q80.json`, `r1_bench_q4b_q80.json`, `r1_bench_q06_q80.json` and `r1_bench_tp2_q17_q80.json` are from the final kernels, the others from earlier "
                "states of the same paths.  `r1_evidence_run_stdout.md` has the values of ALL single-GPU configs on the (almost) final tree: N168 Q80 1498.7 (cluster) / 1460.0 (multi-kernel) / 828.0 (exact), "
                "N168 F32 1233.5, Q06 Q80 1020.0, Q06 Q4K 771.8, Q1.7B Q80 804.7, Q4B Q80 358.9 (372.2 with the final row-block rule).\n\n")
        f.write("| file | workload | GPUs | par. | mode | engine path | value tok/s | e2e tok/s | GB/s per GPU | frac of HBM peak | CPU reference tok/s | launches in timed region |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            fmt = lambda v, k=1: "" if v is None else (f"{v:.{k}f}" if isinstance(v, float) else str(v))
            f.write(f"| `{r[0]}` | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]} | {fmt(r[6])} | {fmt(r[7])} | {fmt(r[8])} | {fmt(r[9], 3)} | {fmt(r[10])} | {r[11]} |\n")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "paths":
    paths_table(os.path.join(ROOT, "profiles", "r1_paths.md"))


def round_outputs(src_dir, tag="r1"):
    """Evidence of tools/round_gpu_run.sh (gpurun_out/<dir>) -> tracked summaries under profiles/."""
    p = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(src_dir)):
        path = os.path.join(src_dir, f)
        if f.startswith("bench_") and f.endswith(".json") and os.path.getsize(path) > 10:
            open(os.path.join(p, f"{tag}_{f}"), "w").write(open(path).read())
        elif f.startswith("launches_") and f.endswith(".csv"):
            launch_list(path, os.path.join(p, f"{tag}_{f[:-4]}.md"), f"Round 1 - launch list ({f[9:-4]}): every kernel launched by the bench command under ncu")
        elif f.endswith(".ncu-rep"):
            full_report(path, os.path.join(p, f"{tag}_ncu_full_{f[5:-8]}.md"), f"Round 1 - ncu --set full ({f[5:-8]})")
        elif f in ("pytest_gpu.log", "smoke.log"):
            open(os.path.join(p, f"{tag}_{f}"), "w").write(open(path).read())
    paths_table(os.path.join(p, f"{tag}_paths.md"))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "round":
    round_outputs(sys.argv[2])


# ---------------------------------------------------------------- round 2
def full_report_wide(rep, out_md, title, note=""):
    """One row per captured launch with the metrics that explain a persistent kernel: time, DRAM bytes, issue utilisation, stall reasons."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(rep, ROOT)}` (`ncu --set full --clock-control none --import-source on`), one block per captured launch."
                f"  Numbers taken under the profiler are never bench values.\n\n{note}\n")
        for row in r[2:]:
            f.write(f"\n## `{row[hdr.index('Kernel Name')][:100]}`\n\n| metric | value |\n|---|---|\n")
            for i in idx[1:]:
                f.write(f"| {hdr[i]} ({units[i]}) | {row[i][:40]} |\n")
            st = sorted(((float(row[hdr.index(h)]), h) for h in stalls), reverse=True)
            f.write("| warp stall reasons per issued instruction (top 6) | " + ", ".join(
                f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st[:6]) + " |\n")
    return r


def round2(src="r2final"):
    g = os.path.join(ROOT, "gpurun_out", src); p = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(g)):
        if f.startswith("bench_") and f.endswith(".json") and os.path.getsize(os.path.join(g, f)) > 10:
            open(os.path.join(p, "r2_" + f), "w").write(open(os.path.join(g, f)).read())
        if f.startswith("micro_") or f in ("pytest_gpu.log", "smoke.log"):
            open(os.path.join(p, "r2_" + f), "w").write(open(os.path.join(g, f)).read())
    if os.path.exists(os.path.join(g, "launches_default_n168.csv")):
        launch_list(os.path.join(g, "launches_default_n168.csv"), os.path.join(p, "r2_launches_default.md"),
                    "Round 2 - launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra` (Nano-168M Q80 seq 512: the streaming kernel is the step)")
    traffic = {}
    rep = os.path.join(g, "prof_stream_n168.ncu-rep")
    if os.path.exists(rep):
        r = full_report_wide(rep, os.path.join(p, "r2_ncu_stream.md"), "Round 2 - ncu --set full, streaming kernel k_decode_stream (Nano-168M Q80)",
                             "The captured launch decodes 32 tokens (positions 15..46 of a seq-512 engine): divide time and bytes by 32 for per-token figures.")
        hdr = r[0]; row = r[2]
        rd = float(row[hdr.index("dram__bytes_read.sum")]); wr = float(row[hdr.index("dram__bytes_write.sum")])
        mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        rd *= mul[r[1][hdr.index("dram__bytes_read.sum")]]; wr *= mul[r[1][hdr.index("dram__bytes_write.sum")]]
        traffic["nano-168m-q80:stream"] = (rd + wr) / 32.0
    rep = os.path.join(g, "prof_multikernel_q06.ncu-rep")
    if os.path.exists(rep):
        r = full_report_wide(rep, os.path.join(p, "r2_ncu_multikernel.md"), "Round 2 - ncu --set full, multi-kernel path (Qwen3-0.6B Q80, the kernels of one layer early in the decode segment)")
        hdr = r[0]
        mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        best = 0.0
        for row in r[2:]:      # the W1|W3 kernel is the one with the most DRAM bytes among the layer's kernels
            b = float(row[hdr.index("dram__bytes_read.sum")]) * mul[r[1][hdr.index("dram__bytes_read.sum")]] + float(row[hdr.index("dram__bytes_write.sum")]) * mul[r[1][hdr.index("dram__bytes_write.sum")]]
            best = max(best, b)
        if best: traffic["qwen3-0.6b-q80:multikernel"] = best
    if traffic:
        json.dump(traffic, open(os.path.join(p, "r2_traffic.json"), "w"), indent=1)
    # table of every committed round-2 bench line
    rows = []
    for f in sorted(os.listdir(p)):
        if not (f.startswith("r2_bench_") and f.endswith(".json")): continue
        try: d = json.loads(open(os.path.join(p, f)).read().strip().splitlines()[-1])
        except Exception: continue
        run = d.get("run") or {}; tr = d.get("token_roofline") or {}; e2e = d.get("e2e") or {}
        rows.append((f, d.get("impl", "b200"), d["config"]["workload"].split(",")[0], d.get("n_gpus"), d.get("value"), e2e.get("value"), run.get("engine", "")[:28],
                     tr.get("frac_of_peak"), (d.get("cpu_baseline") or {}).get("value")))
        for w, x in (d.get("configs") or {}).items():
            rows.append((f + " (configs)", "b200", w, d.get("n_gpus"), x.get("value"), (x.get("e2e") or {}).get("value"), x.get("engine", "")[:28], None, None))
        if d.get("exact_mode"):
            rows.append((f + " (exact_mode)", "b200", d["config"]["workload"].split(",")[0], d.get("n_gpus"), d["exact_mode"].get("value"), None, d["exact_mode"].get("engine", "")[:28], None, None))
    with open(os.path.join(p, "r2_paths.md"), "w") as f:
        f.write("# Round 2 - every committed bench line (tokens/s, decode segment)\n\n| file | arm | workload | GPUs | value | e2e | engine | fraction of HBM peak (token roofline) | cpu_baseline |\n|---|---|---|---|---|---|---|---|---|\n")
        fmt = lambda v: "" if v is None else (f"{v:.1f}" if isinstance(v, float) and v > 2 else f"{v:.3f}" if isinstance(v, float) else str(v))
        for r_ in rows:
            f.write("| " + " | ".join(fmt(v) for v in r_) + " |\n")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "round2":
    round2(sys.argv[2] if len(sys.argv) > 2 else "r2final")

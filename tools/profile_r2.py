"""Round-2 profile summaries from a tools/run_final.sh output directory:
  python tools/profile_r2.py gpurun_out/final_<tag> <tag>
writes profiles/r2/launches_<tag>_summary.txt, launches_graphnet_<tag>_summary.txt (ncu launch lists aggregated by kernel),
profiles/r2/ncu_gnn_<tag>.txt (ncu --set full figures of the fused GraphNet kernels + per-line stall summary) and merges the
per-launch DRAM traffic of those kernels into profiles/ncu_traffic.json (read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import shutil
import sys


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        try:
            t = float(r[ci["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        k = r[ci["Kernel Name"]].split("(")[0][:70]
        agg[k] += t
        cnt[k] += 1
    tot = sum(agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none: every launch of the run (warm-up + timed steps +\n"
                f"# set-up), aggregated by kernel.  Cold-cache and serialised: compare SHARES, not absolutes.  total {tot / 1e3:.0f} us\n")
        f.write("#      ns   share  launches  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write(f"{v:10.0f}  {100 * v / tot:5.1f}%  {cnt[k]:6d}  {k}\n")


def full(raw_csv, lines_txt, out, traffic_json):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one launch of each fused GraphNet kernel\n"
                "# (tools/kt_graphnet_bf16.py: B=256, N=1024, k=20, configs/graph_net.yaml model, bf16 path)\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            f.write(f"===== {name[:100]}\n")
            for w in want:
                if w in idx:
                    f.write(f"{w:75s} {units[idx[w]]:12s} {r[idx[w]][:40]}\n")
            short = name.replace("void ", "").split("(")[0].split("<")[0].split("::")[-1]
            b = sum(float(r[idx[k]].replace(",", "")) * scale.get(units[idx[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic.setdefault(short, b)
        f.write("\n# ---- per source line: share of warp-stall samples / of executed warp instructions (tools/ncu_lines.py)\n")
        f.write(open(lines_txt).read())
    cur = json.load(open(traffic_json)) if os.path.exists(traffic_json) else {"kernels": {}}
    cur["kernels"].update(traffic)
    cur["graphnet_source"] = out
    json.dump(cur, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    d, tag = sys.argv[1], sys.argv[2]
    os.makedirs("profiles/r2", exist_ok=True)
    launches(f"{d}/launches.csv", f"profiles/r2/launches_{tag}_summary.txt", "python bench.py --steps 2 --warmup 1 --no-baselines (deepsets headline)")
    launches(f"{d}/launches_graphnet.csv", f"profiles/r2/launches_graphnet_{tag}_summary.txt",
             "python bench.py --config graphnet --steps 2 --warmup 1 --no-baselines")
    full(f"{d}/ncu_gnn_raw.csv", f"{d}/ncu_gnn_lines.txt", f"profiles/r2/ncu_gnn_{tag}.txt", "profiles/ncu_traffic.json")
    for name in ("bench", "bench_ref", "bench_yaml", "bench_ragged", "bench_graphnet", "bench_sweep"):
        if os.path.exists(f"{d}/{name}.json"):
            shutil.copy(f"{d}/{name}.json", f"profiles/r2/{name}_{tag}.json")
    for name in ("kt_gnn_bf16.txt", "bench_knn.txt", "pytest_gpu.txt"):
        if os.path.exists(f"{d}/{name}"):
            shutil.copy(f"{d}/{name}", f"profiles/r2/{name.replace('.txt', '')}_{tag}.txt")

"""Turn gpurun_out/launches_*.csv + a .ncu-rep into the text summaries committed under profiles/."""
import csv, subprocess, sys, collections

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
    data = rows[1:]
    names = [r[ci['Kernel Name']] for r in data]
    idxs = [i for i, n in enumerate(names) if 'phi_pool_fwd' in n]
    # one full step in the middle of the run: from the kernels that follow a wgrad_reduce up to the next one
    mid = idxs[len(idxs) // 2]
    start = mid
    while start > 0 and 'wgrad_reduce' not in names[start - 1]:
        start -= 1
    end = mid
    while end < len(names) - 1 and 'wgrad_reduce' not in names[end]:
        end += 1
    step = data[start:end + 1]
    tot = sum(float(r[ci['Metric Value']]) for r in step)
    with open(out, 'w') as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 3: launches of ONE train step\n")
        f.write(f"# (cold-cache, serialised: compare shares, not absolutes).  total {tot/1e3:.1f} us, {len(step)} launches\n")
        f.write("#   ns      share  kernel\n")
        for r in step:
            t = float(r[ci['Metric Value']])
            f.write(f"{t:9.0f}  {100*t/tot:5.1f}%  {r[ci['Kernel Name']][:110]}\n")
        agg = collections.defaultdict(float)
        for r in step:
            agg[r[ci['Kernel Name']].split('(')[0][:60]] += float(r[ci['Metric Value']])
        f.write("\n# aggregated by kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write(f"{v:9.0f}  {100*v/tot:5.1f}%  {k}\n")
    print(open(out).read())

def full(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
            'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
            'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active']
    idx = [i for i, h in enumerate(hdr) if h in want]
    traffic = {}
    ki, ri, wi = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    with open(out, 'w') as f:
        f.write("# ncu --set full --clock-control none, one launch of each fused kernel (B=256, N=1024, H=256, relu+max)\n")
        for r in rows[2:]:
            f.write("=====\n")
            for i in idx:
                f.write(f"{hdr[i]:80s} {units[i]:14s} {r[i][:70]}\n")
            name = r[ki].split('(')[0].split('<')[0].replace('void ', '').replace('pcc::', '').replace('phi_pool_fwd_pair_kernel', 'phi_pool_fwd_kernel')
            byts = float(r[ri].replace(',', '')) * scale.get(units[ri], 1.0) + float(r[wi].replace(',', '')) * scale.get(units[wi], 1.0)
            traffic.setdefault(name, byts)   # first launch of each kernel
    import json
    json.dump({"source": out, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "kernels": traffic},
              open("profiles/ncu_traffic.json", "w"), indent=1)
    print(open(out).read())

if __name__ == "__main__":
    tag = sys.argv[1]
    launches(f"gpurun_out/launches_{tag}.csv", f"profiles/launches_{tag}.txt")
    full(f"gpurun_out/prof_{tag}.ncu-rep", f"profiles/ncu_full_{tag}.txt")

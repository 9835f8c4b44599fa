"""Turn the artefacts of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.
    python tools/make_profiles.py r01"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
seq = []
for r in rows[hi + 1:]:
    try:
        seq.append((r[ki].split("(")[0].replace("void ", "").strip(), r[gi], float(r[vi].replace(",", "")) / 1e3))
    except Exception:
        pass
cmd = ("ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 320 python bench.py --steps 2 --warmup 3 "
       "--no-cpu --pool 2 --profile-run")
with open(os.path.join(P, tag + "_launches_full.csv"), "w") as f:
    f.write("# %s\n# config 2 (B=64 L=196 D=512 H=1024 V=10000 T=20); per-launch times are cold-cache and serialised\n" % cmd)
    f.write("index,kernel,grid,us\n")
    for i, (k, g, u) in enumerate(seq):
        f.write('%d,%s,"%s",%.3f\n' % (i, k, g, u))
agg = collections.OrderedDict()
for k, g, u in seq:
    agg.setdefault((k, g), []).append(u)
tot = sum(u for _, _, u in seq)
with open(os.path.join(P, tag + "_launches_summary.csv"), "w") as f:
    f.write("# %s\n# config 2; cold-cache, serialised: compare SHARES (the real loop overlaps kernels through programmatic dependent launch)\n" % cmd)
    f.write("kernel,grid,launches,mean_us,min_us,max_us,share_of_total\n")
    for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write('%s,"%s",%d,%.2f,%.2f,%.2f,%.3f\n' % (k, g, len(v), sum(v) / len(v), min(v), max(v), sum(v) / tot))

# ---- launch list of the training step (tools/gpu_train_list.sh)
tp = os.path.join(G, "train_launches.csv")
if os.path.exists(tp):
    rows = list(csv.reader(open(tp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    tseq = []
    for r in rows[hi + 1:]:
        try:
            name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()
            tseq.append((name, r[gi], float(r[vi].replace(",", "")) / 1e3))
        except Exception:
            pass
    ends = [i for i, (k, _, _) in enumerate(tseq) if k.startswith("adam_kernel")]
    one = tseq[ends[0] + 1:ends[1] + 1] if len(ends) >= 2 else tseq   # exactly one optimizer step
    agg = collections.OrderedDict()
    for k, g, u in one:
        agg.setdefault((k, g), []).append(u)
    tot = sum(u for _, _, u in one)
    with open(os.path.join(P, tag + "_train_launches_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 3000 (tools/gpu_train_list.sh 5000 3000) python bench.py --steps 1 --warmup 1 --no-cpu --workload 4\n")
        f.write("# training step, config 4 (B=64 L=196 D=512 H=1024 V=10000 T=20): the %d launches of ONE step (between two adam_kernel "
                "launches), %.1f us in total; cold-cache and serialised (the real step overlaps launch latencies)\n" % (len(one), tot))
        f.write("kernel,grid,launches,mean_us,total_us,share_of_total\n")
        for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write('%s,"%s",%d,%.2f,%.1f,%.3f\n' % (k, g, len(v), sum(v) / len(v), sum(v), sum(v) / tot))

# ---- full captures
def pick(rep, out, title):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, data = rr[0], rr[1], rr[2:]
    want = ["Kernel Name", "launch__grid_size", "launch__cluster", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct", "lts__t_bytes.sum ", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct",
            "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum ",
            "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_tensor", "smsp__average_warps_issue_stalled", "sm__icc", "sm__cycles_elapsed.max"]
    with open(os.path.join(P, out), "w") as f:
        f.write("# %s\n" % title)
        for i, name in enumerate(h):
            if any(w.strip() in name for w in want):
                f.write("%-92s %-12s %s\n" % (name[:92], units[i][:12], [r[i][:24] for r in data]))
    return h, data

h, d = pick("prof_att.ncu-rep", tag + "_att_wpc_ncu_full.txt",
            "ncu --set full --clock-control none --import-source on -k regex:att_wpc -s 30 -c 2 python bench.py --steps 1 --warmup 3 "
            "--no-cpu --pool 1 --profile-run   (attention launches inside the decode loop: 64 CTAs, config 2)")
rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
unit = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_att.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
u_rd, u_wr = unit[rr[1][rd]], unit[rr[1][wr]]
json.dump({"kernel": "att_wpc_kernel<1> (grid 64)", "dram_bytes_read": float(d[0][rd]) * u_rd, "dram_bytes_write": float(d[0][wr]) * u_wr,
           "source": "profiles/%s_att_wpc_ncu_full.txt (ncu --set full --clock-control none, config 2, one launch)" % tag},
          open(os.path.join(P, "att_traffic.json"), "w"), indent=1)
pick("prof_lin.ncu-rep", tag + "_lin_umma_ncu_full.txt",
     "ncu --set full --clock-control none --import-source on -k regex:lin_umma -s 200 -c 6 python bench.py --steps 1 --warmup 3 --no-cpu "
     "--pool 1 --profile-run   (grid 96 = decode fc_1 || attention state branch, 79 = vocabulary layer, 128 = LSTM)")

# ---- timelines and bench lines
with open(os.path.join(P, tag + "_kernel_timelines.txt"), "w") as f:
    for name, what in (("timeline.log", "tools/timeline.py --head: per launch, first CTA start / first CTA past its dependency wait / last "
                        "accumulator or stream done / last CTA end (device globaltimer, eager launches, config 2)"),
                       ("att_time.log", "tools/att_time.py: attention kernel duration, CUDA events around one launch vs device stamps"),
                       ("trace.log", "tools/trace.py: in-kernel stamps of the attention kernel alone"),
                       ("trace_loop.log", "tools/trace_loop.py: in-kernel stamps of the dense kernels inside the decode loop")):
        p = os.path.join(G, name)
        if os.path.exists(p):
            f.write("=" * 100 + "\n# " + what + "\n" + open(p).read() + "\n")
for src, dst in (("bench.log", "bench_config2"), ("bench_ref.log", "bench_config2_reference_arm"), ("bench_cfg3.log", "bench_cfg3"),
                 ("bench_train1.log", "bench_train1"), ("bench_cfg5.log", "bench_cfg5"), ("bench_multi.log", "bench_multi"),
                 ("bench_multi_ref.log", "bench_multi_reference_arm"), ("bench_train_multi.log", "bench_train_multi")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if lines:
            open(os.path.join(P, "%s_%s.json" % (tag, dst)), "w").write(lines[-1] + "\n")
print("profiles/ written")

"""Turns the ncu captures tools/r02_capture.sh left in gpurun_out/ into the small, committed summaries under profiles/.
    python tools/summarize_profiles.py r02
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def kernels_table(rep, out_path, title):
    rows = ncu_csv(rep, "raw")
    hdr = rows[0]
    want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "time_ms"),
            ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn_KB"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
            ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
            # tcgen05 work is not counted by sm__inst_executed_pipe_tensor (a handful of UTCHMMA instructions drive the pipe
            # for thousands of cycles): the cycle-based counters below are the ones that see it
            ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_mem_cycles_active_pct"),
            ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_cycles_active_pct"),
            ("TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor_hmma_cycles_active"),
            ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem_pipe_inst_pct"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
            ("sm__cycles_elapsed.avg", "sm_cycles")]
    with open(out_path, "w") as f:
        f.write("# " + title + "\n")
        f.write(",".join(n for _, n in want) + "\n")
        for r in rows[2:]:
            vals = []
            for k, n in want:
                v = r[hdr.index(k)] if k in hdr else ""
                if n == "kernel":
                    v = v.replace("void <unnamed>::", "").replace("<unnamed>::", "").replace(",", ";")[:64]
                vals.append('"%s"' % v)
            f.write(",".join(vals) + "\n")


def traffic_json(rep, out_path):
    """dram__bytes_read + dram__bytes_write per launch for the PQ / LRN kernels of ONE forward pass, in launch order."""
    import json
    rows = ncu_csv(rep, "raw")
    hdr = rows[0]
    order = ["conv1", "lrn1+pool1", "conv2", "lrn2+pool2", "conv3", "conv4", "conv5", "fc6", "fc7", "fc8"]
    out, i = {}, 0
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if "reduce" in name or "flatten" in name or "fc_prep" in name or i >= len(order):
            continue

        def mb(col):
            v = float(r[hdr.index(col)])
            unit = rows[1][hdr.index(col)].lower()
            return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        out[order[i]] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
        i += 1
    json.dump(out, open(out_path, "w"), indent=1)


def stage_split(rep, out_path, title):
    """Per kernel: share of stall samples / instructions / shared-memory wavefronts between consecutive barriers
    (== the staging / LUT / gather stages of the conv kernels)."""
    rows = ncu_csv(rep, "source")
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "ins": []}
            kern.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur and cur["hdr"] and len(r) >= 6:
            cur["ins"].append(r)
    seen = set()
    with open(out_path, "w") as f:
        f.write("# " + title + "\n")
        for k in kern:
            h = k["hdr"]
            if not h or "# Samples" not in h or "Instructions Executed" not in h:
                continue
            i_s, i_e = h.index("# Samples"), h.index("Instructions Executed")
            i_w = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
            tot = sum(int(r[i_s] or 0) for r in k["ins"])
            tote = sum(int(r[i_e] or 0) for r in k["ins"])
            if (k["name"], tot) in seen or tot == 0:
                continue
            seen.add((k["name"], tot))
            sass = " ".join(r[1] for r in k["ins"])
            f.write("\n## %s\nsamples %d, warp instructions %d; SASS has UTCHMMA: %s, LDTM: %s, LDGSTS: %s, FFMA2: %s, FADD2: %s\n" % (
                k["name"].replace("void <unnamed>::", "")[:90], tot, tote, "UTCHMMA" in sass, "LDTM" in sass,
                "LDGSTS" in sass, "FFMA2" in sass, "FADD2" in sass))
            reg, regs = 0, collections.OrderedDict()
            for r in k["ins"]:
                op = r[1].strip().split()[0] if r[1].strip() else ""
                if op.startswith("@") and len(r[1].strip().split()) > 1:
                    op = r[1].strip().split()[1]
                d = regs.setdefault(reg, {"s": 0, "e": 0, "w": 0, "ops": collections.Counter()})
                d["s"] += int(r[i_s] or 0)
                d["e"] += int(r[i_e] or 0)
                d["w"] += int(r[i_w] or 0) if i_w is not None else 0
                d["ops"][op.split(".")[0]] += int(r[i_e] or 0)
                if op.startswith("BAR"):
                    reg += 1
            for rg, d in regs.items():
                if d["s"] * 100 < tot:
                    continue
                top = ", ".join("%s %.0f%%" % (o, 100.0 * c / max(d["e"], 1)) for o, c in d["ops"].most_common(5))
                f.write("- segment %d: %.1f%% of stall samples, %.1f%% of instructions, %d smem wavefronts | %s\n" % (
                    rg, 100.0 * d["s"] / tot, 100.0 * d["e"] / max(tote, 1), d["w"], top))


def launch_list(csv_path, out_path, title):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        k = row["Kernel Name"].replace("void <unnamed>::", "").replace("<unnamed>::", "")[:60]
        agg.setdefault((k, row.get("Grid Size", ""), row.get("Block Size", "")), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out_path, "w") as f:
        f.write("# " + title + "\n# per-launch device time is cold-cache and serialised under ncu: compare SHARES\n")
        f.write("kernel,grid,block,launches,total_us,share_pct,avg_us\n")
        for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write('"%s","%s","%s",%d,%.1f,%.2f,%.1f\n' % (k.replace(",", ";"), g, b, len(v), sum(v) / 1e3, 100 * sum(v) / tot, sum(v) / len(v) / 1e3))


def step_metrics(csv_path, out_path, traffic_path):
    """ncu --metrics ... --csv log of one forward pass (long format: one row per launch and metric) -> one row per launch;
    DRAM bytes per launch of the PQ / LRN kernels -> traffic.json (read by bench.py for `roofline.traffic`)."""
    import json
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row["ID"], row["Kernel Name"])
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "").lower()
        v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)   # bytes / ns
        per.setdefault(key, collections.OrderedDict())[row["Metric Name"]] = v
    names = []
    for d in per.values():
        for k in d:
            if k not in names:
                names.append(k)
    with open(out_path, "w") as f:
        f.write("# ncu --metrics ... --clock-control none; tools/profile_step.py --batch 256: every kernel of one AlexNet PQ forward pass (bytes, ns)\n")
        f.write("kernel," + ",".join(names) + "\n")
        for (_, k), d in per.items():
            f.write('"%s",' % k.replace("void <unnamed>::", "").replace("<unnamed>::", "").replace(",", ";")[:64] +
                    ",".join("%.6g" % d.get(n, float("nan")) for n in names) + "\n")
    order = ["conv1", "lrn1+pool1", "conv2", "lrn2+pool2", "conv3", "conv4", "conv5", "fc6", "fc7", "fc8"]
    out, i = {}, 0
    for (_, k), d in per.items():
        if "reduce" in k or "fc_prep" in k or "maxpool_kernel" in k or "softmax" in k or "u8hwc" in k or i >= len(order):
            continue
        out[order[i]] = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        i += 1
    json.dump(out, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    # python tools/summarize_profiles.py <tag>  -- reads gpurun_out/<tag>_*.ncu-rep / .csv written by tools/r02_capture.sh
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    G = os.path.join(ROOT, "gpurun_out")
    P = os.path.join(ROOT, "profiles")
    os.makedirs(P, exist_ok=True)
    step = os.path.join(G, "%s_step_b256.ncu-rep" % tag)
    gemm = os.path.join(G, "%s_pq_gemm_b256.ncu-rep" % tag)
    chain = os.path.join(G, "%s_fc_chain_b1.ncu-rep" % tag)
    if os.path.exists(step):
        kernels_table(step, os.path.join(P, "%s_ncu_full_b256.csv" % tag),
                      "ncu --set full --clock-control none; tools/profile_step.py --batch 256: every kernel of one AlexNet PQ forward pass")
        traffic_json(step, os.path.join(P, "traffic.json"))
    stepm = os.path.join(G, "%s_step_b256_metrics.csv" % tag)
    if os.path.exists(stepm):
        step_metrics(stepm, os.path.join(P, "%s_step_b256_metrics.csv" % tag), os.path.join(P, "traffic.json"))
    if os.path.exists(gemm):
        kernels_table(gemm, os.path.join(P, "%s_ncu_pq_gemm_b256.csv" % tag),
                      "ncu --set full --clock-control none --import-source on -k regex:pq_gemm_tc; conv1..conv5 at batch 256")
        stage_split(gemm, os.path.join(P, "%s_stage_split_pq_gemm_b256.md" % tag),
                    "pq_gemm_tc (conv1..conv5, batch 256): stall samples / instructions / shared-memory wavefronts between barriers")
    if os.path.exists(chain):
        kernels_table(chain, os.path.join(P, "%s_ncu_fc_chain_b1.csv" % tag),
                      "ncu --set full --clock-control none --import-source on -k regex:fc_chain; fc6->fc7->fc8 as one launch, batch 1")
        stage_split(chain, os.path.join(P, "%s_stage_split_fc_chain_b1.md" % tag),
                    "fc_chain_kernel (batch 1): stall samples / instructions / shared-memory wavefronts between barriers")
    for name, title in (("launches_bench", "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off; "
                         "the timed steps of `python bench.py --steps 2 --warmup 3` (batch 256)"),
                        ("launches_b1", "same metric; one forward pass at batch 1 (tools/profile_step.py --batch 1)")):
        src = os.path.join(G, "%s_%s.csv" % (tag, name))
        if os.path.exists(src):
            launch_list(src, os.path.join(P, "%s_%s.csv" % (tag, name)), title)

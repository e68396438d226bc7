#!/usr/bin/env python3
"""Turn the files a `tools/evidence.sh` + `tools/launchlist.sh` run left in gpurun_out/ into the tracked artefacts under profiles/
(run in the build container, after the GPU call): bench lines, raw ncu metrics, per-line tables, DRAM traffic, launch-list summary."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def bench_lines():
    for f in ("default", "reference", "mug", "reach", "f64"):
        src = os.path.join(G, "r2_bench_%s.json" % f)
        if os.path.exists(src):
            d = json.loads(open(src).read().strip().splitlines()[-1])
            json.dump(d, open(os.path.join(P, "r2_bench_%s.json" % f), "w"), indent=1)
            print(f, round(d["value"]), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 4),
                  "fp32", d.get("roofline", {}).get("frac"), "hbm", d.get("roofline_hbm", {}).get("frac"), "kernel_ms", d.get("roofline", {}).get("kernel_ms"),
                  "cpu", d.get("cpu_baseline") and (round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"]))
            if d.get("workloads"):
                print("  workloads", {k: (round(v["value"]), round(v["mean_ncon"], 2), round(v["contact_rich_frac"], 3), round(v["fp32_frac"], 4)) for k, v in d["workloads"].items()})
            c = d.get("config", {})
            print("  ", {k: c.get(k) for k in ("mean_ncon", "mean_nefc", "contact_rich_frac", "episodes", "truncations", "kernel_ms_lite_tier", "kernel_ms_full_tier", "kernel_ms_side_stream")})


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


def metrics():
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
            "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
    out = {}
    for name in ("r2_prof_rollout", "r2_prof_mug"):
        rep = os.path.join(G, name + ".ncu-rep")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
        ks = []
        for r in rows[2:]:
            d = {"kernel": r[h.index("Kernel Name")]}
            for k in keys:
                if k in h:
                    d[k] = {"value": num(r[h.index(k)]), "unit": rows[1][h.index(k)]}
            ks.append(d)
        out[name] = ks
        st = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), rep], capture_output=True, text=True).stdout
        out[name + "_stall_samples_all_captured_kernels"] = st.split("\n")[:8]
        for d in ks:
            print(name, d["kernel"][24:70], d["gpu__time_duration.sum"], "Minst", round(d["smsp__inst_executed.sum"]["value"] / 1e6, 1),
                  "issue", round(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"], 1), "warps", round(d["sm__warps_active.avg.pct_of_peak_sustained_active"]["value"], 1),
                  "lanes", d["smsp__thread_inst_executed_per_inst_executed.ratio"]["value"], "regs", d["launch__registers_per_thread"]["value"],
                  "dram", d["dram__bytes_read.sum"]["value"], d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"])
        print("  stalls", out[name + "_stall_samples_all_captured_kernels"])
    out["how"] = ("tools/prof.sh <workload> <tag>: ncu --set full --clock-control none --import-source on -k regex:step_kernel, the four size-class launches of one env-step "
                  "inside the timed region of bench.py (after the same command exited 0 without ncu)")
    json.dump(out, open(os.path.join(P, "r2_metrics.json"), "w"), indent=1)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    lite = out["r2_prof_rollout"][0]
    tr = {"workload": "rollout", "envs": 65536,
          "capture": "profiles/r2_metrics.json: r2_prof_rollout (tools/prof.sh rollout: ncu --set full --clock-control none on `bench.py --workload rollout --settle 3000 --steps 4 --warmup 3`, the L2 flush between steps as in the bench)",
          "kernel": lite["kernel"], "dram_bytes_read": lite["dram__bytes_read.sum"]["value"] * mult[lite["dram__bytes_read.sum"]["unit"]],
          "dram_bytes_write": lite["dram__bytes_write.sum"]["value"] * mult[lite["dram__bytes_write.sum"]["unit"]]}
    json.dump(tr, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)


def by_line():
    cub = "/tmp/ur3e_cubins"
    subprocess.run("rm -rf %s && mkdir -p %s && cd %s && cuobjdump -xelf all %s > /dev/null 2>&1" % (cub, cub, cub, os.path.join(ROOT, "ur3e_b200", "libur3e_b200.so")), shell=True)
    cubin = os.path.join(cub, "inst_f32_main.sm_100a.cubin")
    for rep, kern, sect, dst in (("r2_prof_rollout", "(int)8, (int)44", "Li48ELi8ELi44ELi14ELb1EEELb0", "r2_by_line_rollout_lite.txt"),
                                 ("r2_prof_mug", "(int)16, (int)68", "Li48ELi16ELi68ELi14ELb1EEELb0", "r2_by_line_mug_grasp_tier.txt")):
        o = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), os.path.join(G, rep + ".ncu-rep"), cubin, kern, "45", sect], capture_output=True, text=True, cwd=ROOT)
        open(os.path.join(P, dst), "w").write(o.stdout)
        print(dst, o.stdout.split("\n")[:2], o.stderr[-200:])


def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, "r2_launches.csv"))) if len(r) > 5]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]

    def short(n):
        m = re.search(r"step_kernel<(float|double), ur3e::Dims<([^>]*)>", n)
        if m:
            d = m.group(2).replace("(int)", "").replace("(bool)", "").split(", ")
            return "ur3e::step_kernel<%s, %s contacts / %s rows%s>" % (m.group(1), d[6], d[7], ", exact-fit" if d[9] == "1" else "")
        if "env_kernel" in n: return "ur3e::env_kernel (reset / set_state)"
        if "stats_kernel" in n: return "ur3e::stats_kernel"
        if "state_io" in n: return "ur3e::state_io_kernel"
        if "FillFunctor" in n: return "torch fill (incl. the 192 MiB L2 flush)"
        return "torch: " + n.split("<")[0].replace("void ", "")[:60]
    agg = collections.OrderedDict()
    for n, v in data:
        a = agg.setdefault(short(n), [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    lines = ["# ncu launch list of `python bench.py --settle 100 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e` (tools/launchlist.sh;",
             "# `ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised: compare shares).",
             "# %d launches, %.1f ms of GPU time in total; first 400 rows of the raw list: profiles/r2_launches_head.csv" % (len(data), tot / 1e6),
             "kernel,launches,total_ms,share_of_gpu_time,max_us"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append('"%s",%d,%.3f,%.4f,%.1f' % (k, a[0], a[1] / 1e6, a[1] / tot, a[2] / 1e3))
    idx = [i for i, (n, v) in enumerate(data) if "FillFunctor<unsigned char>" in n]
    if len(idx) > 3:
        seg = data[idx[-3] + 1:idx[-2]]
        st = sum(v for n, v in seg if "step_kernel" in n); al = sum(v for n, v in seg)
        lines.append("# one timed env-step (between two L2-flush fills): %d launches, step kernels %.1f us of %.1f us = %.3f of the step; lite-tier kernel %.1f us" % (len(seg), st / 1e3, al / 1e3, st / al, max(v for n, v in seg) / 1e3))
        lines.append("# launches of that step: " + "; ".join("%s %.1f us" % (short(n), v / 1e3) for n, v in seg))
    open(os.path.join(P, "r2_launches_summary.csv"), "w").write("\n".join(lines) + "\n")
    raw = open(os.path.join(G, "r2_launches.csv")).read().split("\n")
    open(os.path.join(P, "r2_launches_head.csv"), "w").write("\n".join(raw[:404]) + "\n")
    print("\n".join(lines[-14:]))


if __name__ == "__main__":
    bench_lines(); metrics(); by_line(); launches()

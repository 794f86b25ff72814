"""Build profiles/<round>/SUMMARY.md from the artefacts a profiling run leaves in that directory:

    *_ncu_raw.csv        `ncu -i X.ncu-rep --page raw --csv` exports of `--set full` captures
    launches_gpu_time.csv  `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list
    bench_c5_n1.json     the bench line of the same build (CUDA-event shares)
    jk_bench_n120.json   scripts/bench_jk.py line

    python scripts/make_profile_summary.py profiles/r01
"""
import collections
import csv
import glob
import json
import os
import sys

KEYS = [
    ("duration", "gpu__time_duration.sum"),
    ("DMMA sub-pipe % of peak (active)", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    ("tensor pipe active % (elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active % (active cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("INT8 (IMMA) sub-pipe active %", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("shared-memory wavefronts of tensor-core operand reads % of peak", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("L2 -> SM read", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second"),
    ("L2 throughput % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("FP64 (non-tensor) pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("FP64 tensor ops % of peak (elapsed)", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("long-scoreboard stall (warps per issue)", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("registers/thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("warp instructions", "smsp__inst_executed.sum"),
]


def short(name):
    name = name.replace("void ", "").replace("qexxc::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.split("(")[0]


def raw_pages(d):
    out = []
    for path in sorted(glob.glob(os.path.join(d, "*ncu*raw*.csv")) + glob.glob(os.path.join(d, "ncu_*_raw.csv"))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        for row in rows[2:]:
            rec = dict(zip(h, row))
            u = dict(zip(h, units))
            out.append((os.path.basename(path), rec, u))
    seen, uniq = set(), []
    for p, rec, u in out:
        key = (p, rec.get("ID"), rec.get("Kernel Name"))
        if key not in seen:
            seen.add(key)
            uniq.append((p, rec, u))
    return uniq


def main(d):
    lines = [f"# Profiles in `{d}` (B200; c5 = 1000 AOs x 1e6 grid points unless noted)", "",
             "`ncu --set full --clock-control none --import-source on`, one launch per kernel; raw metric pages are the",
             "`*_raw.csv` files next to this summary (regenerate with `python scripts/make_profile_summary.py`).", ""]
    for page, rec, u in raw_pages(d):
        lines.append(f"### {short(rec.get('Kernel Name', '?'))}   ({page})")
        for label, key in KEYS:
            if key in rec and rec[key] != "":
                lines.append(f"- {label}: {rec[key]} {u.get(key, '')}".rstrip())
        lines.append("")
    for ll, bj_name, title in ((os.path.join(d, "launches_gpu_time_int8.csv"), "bench_c5_n1_int8.json", "INT8 contractions (default for nao >= 256)"),
                               (os.path.join(d, "launches_gpu_time.csv"), "bench_c5_n1.json", "FP64 DMMA contractions (QEXXC_I8=0)")):
        if not os.path.exists(ll):
            continue
        lines += [f"# {title}", ""]
        rows = [r for r in csv.reader(l for l in open(ll) if not l.startswith("=="))]
        h = rows[0]
        ki, vi = h.index("Kernel Name"), len(h) - 1
        agg = collections.OrderedDict()
        for r in rows[1:]:
            k = short(r[ki])
            if k.startswith("cutlass") or k.startswith("at::") or k.startswith("i8_peak"):
                continue  # bench.py's own cuBLAS DGEMM peak measurement and torch fills
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi])
        tot = sum(v[1] for v in agg.values())
        lines += ["## Launch list (ncu --metrics gpu__time_duration.sum, serialised / cold cache: compare SHARES)",
                  "kernel | launches | total ms | share", "---|---|---|---"]
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{k} | {n} | {t / 1e6:.3f} | {t / tot:.4f}")
        lines.append("")
        bj = os.path.join(d, bj_name)
        if not os.path.exists(bj):
            continue
        b = json.load(open(bj))
        lines += ["## bench.py c5, N=1 (CUDA events on the launching stream, in-run shares)",
                  f"- step: {b['ms_per_step']:.2f} ms -> {b['value']:.4g} {b['unit']}; e2e {b['e2e']['value']:.4g}; "
                  f"gpu_launches {b.get('gpu_launches')}; clocks {b.get('clocks')}"]
        for k, v in b["roofline"]["kernels"].items():
            lines.append(f"- {k}: " + ", ".join(f"{a}={(round(x, 4) if isinstance(x, float) else x)}" for a, x in v.items()))
        lines.append("")
    jj = os.path.join(d, "jk_bench_n120.json")
    if os.path.exists(jj):
        j = json.load(open(jj))
        lines += ["## scripts/bench_jk.py (N2: incore J/K, nao = 120, 1.66 GB tensor)",
                  f"- ms per call: {j['ms_per_call']}",
                  f"- roofline: {j['roofline']['achieved']:.0f} GB/s of {j['roofline']['peak']:.0f} = {j['roofline']['frac']:.3f}; "
                  f"J only {j['roofline']['j_only_frac']:.3f}; reverse {j['roofline']['vjp_frac']:.3f}; "
                  f"cuBLAS GEMV (J only) {j['roofline']['cublas_gemv_frac']:.3f}", ""]
    open(os.path.join(d, "SUMMARY.md"), "w").write("\n".join(lines))
    print("wrote", os.path.join(d, "SUMMARY.md"), len(lines), "lines")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r01")

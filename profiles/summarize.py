"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run in the dev container, no GPU).
    python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/launches_r01.md
    python profiles/summarize.py full gpurun_out/prof_x.ncu-rep > profiles/prof_x.md
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r["Kernel Name"].split("(")[0][-60:], []).append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean ns | share of profiled time |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %d | %.0f | %.1f%% |" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("### `%s`\n" % r[idx["Kernel Name"]][:110])
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in idx:
                print("| %s | %s | %s |" % (w, r[idx[w]], units[idx[w]]))
        print()


def traffic(path):
    """JSON {kernel short name: dram bytes read + written per launch} — what bench.py reports as roofline.traffic."""
    import json
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for r in rows[2:]:
        name = re.sub(r"^.*::", "", r[idx["Kernel Name"]].split("(")[0])
        b = sum(float(r[idx[k]].replace(",", "")) * mult[units[idx[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        res[name] = {"dram_bytes": b, "duration_us": float(r[idx["gpu__time_duration.sum"]].replace(",", "")),
                     "report": path.split("/")[-1]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])

"""Print the key metrics of an .ncu-rep (run here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        ]


def _num(v):
    return float(v.replace(",", ""))


def _bytes(v, unit):
    return _num(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main(path, traffic_json=None, algorithmic_bytes=None, commit=None, launch=None):
    """python scripts/ncu_summary.py file.ncu-rep [--traffic-json profiles/roofline_traffic.json
    --algorithmic-bytes N --commit SHA --launch "256->256 k3 @8x36x44x28"]: with --traffic-json the FIRST kernel of the
    report becomes bench.py's roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        print("----", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:86s} {units[i]:14s} {r[i]}")
        if traffic_json and n == 0:
            import json
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            t = _bytes(r[ir], units[ir]) + _bytes(r[iw], units[iw])
            it = hdr.index("gpu__time_duration.sum")
            json.dump({"kernel": r[hdr.index("Kernel Name")][:120], "launch": launch, "traffic_bytes_per_launch": t,
                       "algorithmic_bytes_per_launch": algorithmic_bytes, "duration": f"{r[it]} {units[it]}",
                       "tensor_pipe_active_pct": r[hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]
                       if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr else None,
                       "source": "ncu --set full --clock-control none, " + path.split("/")[-1], "commit": commit},
                      open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--traffic-json")
    ap.add_argument("--algorithmic-bytes", type=float)
    ap.add_argument("--commit")
    ap.add_argument("--launch")
    a = ap.parse_args()
    main(a.rep, a.traffic_json, a.algorithmic_bytes, a.commit, a.launch)

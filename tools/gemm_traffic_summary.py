"""Summarises `ncu -i gemm_full.ncu-rep --page raw --csv` of the four GEMMs of encoder layer 0 (M = 2*64*1568 rows) into
profiles/r01_gemm_ncu_full_summary.txt-style lines and the JSON bench.py reads for `roofline.traffic`.
usage: gemm_traffic_summary.py raw.csv out.txt out.json"""
import csv
import json
import sys

raw, out_txt, out_json = sys.argv[1:4]
rows = list(csv.reader(open(raw, errors="replace")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, body = rows[h], rows[h + 2:]
col = {n: i for i, n in enumerate(H)}
M = 2 * 64 * 1568
shapes = [("qkv (K=1024,N=3072, head-split epilogue)", 1024, 3072, 2, 0),
          ("proj+residual (K=1024,N=1024)", 1024, 1024, 4, 4),
          ("lin1+GELU (K=1024,N=4096)", 1024, 4096, 2, 0),
          ("lin2+residual (K=4096,N=1024)", 4096, 1024, 4, 4)]


def f(r, name):
    return float(r[col[name]].replace(",", "")) if name in col and r[col[name]] not in ("", "n/a") else float("nan")


def to_bytes(r, name):
    v = f(r, name)
    unit = rows[h + 1][col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


lines, launches = [], []
for r, (name, K, N, out_b, res_b) in zip(body, shapes):
    alg = M * K * 2 + N * K * 2 + M * N * (out_b + res_b)
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    t = f(r, "gpu__time_duration.sum")
    tunit = rows[h + 1][col["gpu__time_duration.sum"]]
    t_ms = t * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1}.get(tunit, 1)
    tp = f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    if tp != tp:
        tp = f(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")
    l2 = f(r, "lts__t_sector_hit_rate.pct")
    regs = f(r, "launch__registers_per_thread")
    kname = r[col["Kernel Name"]][:60]
    lines.append(f"{name:45s} time_ms={t_ms:6.3f} dram_read={rd / 1e9:6.3f} GB dram_write={wr / 1e9:6.3f} GB "
                 f"traffic={(rd + wr) / 1e9:6.3f} GB algorithmic={alg / 1e9:6.3f} GB ratio={(rd + wr) / alg:5.2f} "
                 f"tensor_pipe_active={tp:5.1f}% L2_hit={l2:5.1f}% regs={regs:.0f}   [{kname}]")
    launches.append({"kernel": name, "dram_bytes": rd + wr, "algorithmic_bytes": alg, "tensor_pipe_active_pct": tp})
avg_t = sum(x["dram_bytes"] for x in launches) / len(launches)
avg_a = sum(x["algorithmic_bytes"] for x in launches) / len(launches)
lines.append("")
lines.append(f"per-launch average: traffic={avg_t / 1e9:.3f} GB algorithmic={avg_a / 1e9:.3f} GB ratio={avg_t / avg_a:.2f}")
open(out_txt, "w").write("\n".join(lines) + "\n")
json.dump({"source": out_txt, "per_launch_avg_dram_bytes": avg_t, "per_launch_avg_algorithmic_bytes": avg_a,
           "launches": launches}, open(out_json, "w"), indent=1)
print("\n".join(lines))

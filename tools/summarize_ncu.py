"""Turn raw `ncu --csv` metric logs (gpurun_out/) into the compact tables committed under profiles/.
usage: summarize_ncu.py launches <ncu.csv> <out_list.csv> <out_by_kernel.csv> [skip]
       summarize_ncu.py stages   <ncu.csv> <out_stages.csv> <out_traffic.json>"""
import csv, json, re, sys
from collections import OrderedDict, defaultdict


def read(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
    hdr = rows[0]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
    out = OrderedDict()
    for r in rows[1:]:
        try:
            i = int(r[col["ID"]])
        except ValueError:
            continue
        d = out.setdefault(i, {"kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("sacb::", ""), "grid": r[col["Grid Size"]]})
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        name = r[col["Metric Name"]]
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        elif unit in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d[name] = v
    return out


def launches(src, out_list, out_by, skip):
    L = read(src)
    by = defaultdict(lambda: [0, 0.0, 0.0])
    with open(out_list, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "time_us", "dram_read_B", "dram_write_B"])
        for i, d in L.items():
            t, rd, wr = d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
            w.writerow([i, d["kernel"], d["grid"], round(t, 2), int(rd), int(wr)])
            if i >= skip:
                b = by[d["kernel"]]
                b[0] += 1; b[1] += t; b[2] += rd + wr
    tot = sum(b[1] for b in by.values())
    with open(out_by, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us", "dram_MB_total"])
        for k, b in sorted(by.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, b[0], round(b[1], 1), round(100 * b[1] / tot, 1), round(b[1] / b[0], 2), round(b[2] / 1e6, 2)])
    print(open(out_by).read())


def stages(src, out_stages, out_json):
    L = read(src)
    tot_t = tot_d = tot_l = 0.0
    with open(out_stages, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["stage", "grid", "time_us", "dram_read_B", "dram_write_B", "l2_bytes", "tensor_pipe_active_pct"])
        for s, (i, d) in enumerate(L.items()):
            t, rd, wr = d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
            l2 = d.get("lts__t_bytes.sum", 0.0)
            w.writerow([s, d["grid"], round(t, 2), int(rd), int(wr), int(l2), round(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0), 2)])
            tot_t += t; tot_d += rd + wr; tot_l += l2
    json.dump({"update_dram_bytes": tot_d, "update_l2_bytes": tot_l, "sum_time_us": tot_t}, open(out_json, "w"))
    print(open(out_json).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    else:
        stages(sys.argv[2], sys.argv[3], sys.argv[4])

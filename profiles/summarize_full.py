"""Per-kernel table from `ncu -i <rep> --page raw --csv` of an `ncu --set full` capture: launches, mean duration,
DRAM bytes per launch (read + write), achieved DRAM GB/s and % of the measured HBM peak, tensor-pipe and warp
activity, registers.   python profiles/summarize_full.py raw.csv [hbm_peak_GBs]"""
import collections
import csv
import re
import sys

COLS = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor2": "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "smem": "launch__shared_mem_per_block_dynamic",
    "grid": "launch__grid_size",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l2hit": "lts__t_sector_hit_rate.pct",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3,
        "msecond": 1e3, "second": 1e6}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path, peak=6538.3):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h, units = rows[hdr], rows[hdr + 1]
    ki = h.index("Kernel Name")
    idx = {k: (h.index(v) if v in h else None) for k, v in COLS.items()}
    agg = collections.OrderedDict()
    gi = h.index("Grid Size") if "Grid Size" in h else None
    for r in rows[hdr + 2:]:
        if len(r) <= ki:
            continue
        full = r[ki].replace("<unnamed>::", "").replace("void ", "")
        m = re.match(r"([\w:]+?_kernel)(<[^(]*>)?\(", full)
        name = (m.group(1) + (m.group(2) or "")) if m else full[:48]
        if name.startswith(("gemm_bf16_tc", "gemm_f32", "attention_ws", "attn_varlen", "fps_kernel", "layernorm")) and gi is not None:
            name += " grid " + r[gi].replace(" ", "")  # one row per launch shape
        vals = {}
        for k, i in idx.items():
            if i is None:
                continue
            v = num(r[i])
            u = units[i].strip()
            if k in ("dur", "rd", "wr", "smem"):
                v *= UNIT.get(u, 1.0)
            vals[k] = v
        a = agg.setdefault(name, collections.defaultdict(list))
        for k, v in vals.items():
            a[k].append(v)
    mean = lambda xs: sum(xs) / len(xs) if xs else float("nan")  # noqa: E731
    print(f"{'kernel (one row per template instantiation / launch shape)':58s} {'n':>4s} {'us':>8s} {'DRAM MB':>8s} {'GB/s':>7s} {'%HBM':>5s} {'tensor%':>7s} {'warps%':>6s} {'issue%':>6s} {'regs':>4s} {'smemKB':>6s} {'grid':>6s}")
    for name, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]["dur"])):
        dur = mean(a["dur"])
        traffic = mean(a["rd"]) + mean(a["wr"])
        gbs = traffic / (dur * 1e-6) / 1e9 if dur > 0 else 0.0
        tensor = max(mean(a.get("tensor", [])) if a.get("tensor") else 0.0, mean(a.get("tensor2", [])) if a.get("tensor2") else 0.0)
        print(f"{name[:58]:58s} {len(a['dur']):4d} {dur:8.1f} {traffic / 1e6:8.2f} {gbs:7.0f} {100 * gbs / peak:5.1f} {tensor:7.1f} "
              f"{mean(a['warps']):6.1f} {mean(a.get('issue', [float('nan')])):6.1f} {mean(a['regs']):4.0f} {mean(a.get('smem', [0])) / 1024:6.1f} {mean(a['grid']):6.0f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 6538.3)

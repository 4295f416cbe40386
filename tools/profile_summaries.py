"""Turns the raw ncu outputs a GPU call left in gpurun_out/ into the tracked summaries of profiles/.

  python tools/profile_summaries.py launches gpurun_out/launches_bench.csv profiles/rNN_launches_bench.md "<command>"
  python tools/profile_summaries.py full     profiles/rNN_ncu_tensor_kernels.md title=rep.ncu-rep[,note] ...
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("at::native::", "")
    return name[:64]


def launches(csv_path, out_path, command):
    rows = []
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        rows.append((short(r["Kernel Name"]), ms))
    agg = OrderedDict()
    for k, ms in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if not (k.startswith("at::") or "elementwise" in k or "cutlass" in k
                                                    or "reduce_kernel" in k or "cublas" in k or "nccl" in k
                                                    or "gemv" in k or "Memcpy" in k or "memset" in k.lower()))
    with open(out_path, "w") as o:
        o.write("# ncu launch list of the bench command (per-kernel device time, cold-cache / serialised)\n\n")
        o.write(f"Command (under gpurun): `{command}`\n")
        o.write(f"({len(rows)} launches captured; raw rows: `{csv_path.split('/')[-1]}` next to this file).  "
                "Shares, not absolutes, are comparable with the CUDA-event numbers of `bench.py`.\n\n")
        o.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if ms / total < 0.002:
                continue
            o.write(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / total:.1f} % |\n")
        o.write(f"| all | {len(rows)} | {total:.1f} | 100 % |\n\n")
        o.write(f"Hand-written kernels of this repo: {100 * ours / total:.1f} % of the device time; the rest are torch "
                "elementwise / reduction kernels on the tiny (B,C) head tensors, gradient accumulation and fills.\n")
    print("wrote", out_path)


def full(out_path, specs):
    with open(out_path, "w") as o:
        o.write("# `ncu --set full --clock-control none` captures of the tensor-core kernels\n\n")
        o.write("Commands: see `tools/run_round.sh` (`ncu --set full --clock-control none --import-source on -k regex:<kernel> "
                "-s 1 -c N python tools/bench_layers.py 32 up_tr64.ops.0`, `PCRL_PREC=fp32` for the TF32 build of the "
                "same launch).  The layer is the largest launch shape of the step (b=32, 128->64 channels, 64x64x32: "
                "1855 GFLOP algorithmic; it carries 30 % of the forward FLOPs).\n"
                "Algorithmic bytes of that launch in bf16: read x 32*64*65*32*128*2 = 1.09 GB, write y 0.55 GB, weights "
                "0.44 MB (fp32 storage: twice that).\n\n")
        for spec in specs:
            title, rest = spec.rsplit("=", 1)
            rep, _, note = rest.partition(",")
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader([l for l in txt.splitlines() if not l.startswith("==")]))
            hdr, units, data = rows[0], rows[1], rows[2:]
            o.write(f"## {title}\n\n")
            if note:
                o.write(note + "\n\n")
            o.write("| metric | unit | value (per captured launch) |\n|---|---|---|\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    vals = ", ".join(short(r[i]) if m == "Kernel Name" else r[i] for r in data)
                    o.write(f"| `{m}` | {units[i]} | {vals} |\n")
            o.write("\n")
    print("wrote", out_path)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3:])

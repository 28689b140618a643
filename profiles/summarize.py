#!/usr/bin/env python3
"""Turn an `ncu --set full --import-source on` capture of k_aggregate into the text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_agg_r1x.ncu-rep [libdnz_gpu.so] > profiles/r1x_k_aggregate.txt

Reads the report with `ncu -i ... --page raw/source --csv` (works without a GPU).  When the matching libdnz_gpu.so is
given, SASS rows are attributed to CUDA source lines through `nvdisasm -g` (the build uses -lineinfo)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_selected", "stall_not_selected", "stall_branch_resolving",
          "stall_math", "stall_mio", "stall_lg", "stall_no_inst", "stall_sleep", "stall_barrier", "stall_membar", "stall_dispatch",
          "stall_drain", "stall_tex", "stall_misc"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def line_map(so, kernel="_ZN3dnz11k_aggregateENS_9AggParamsE"):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith("dnz_kernels") and f.endswith(".cubin")]
    if not cub:
        return None
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub[0])], capture_output=True, text=True).stdout.split("\n")
    try:
        start = [i for i, l in enumerate(txt) if l.startswith(".text." + kernel + ":")][0]
    except IndexError:
        return None
    cur, seq = None, []
    for l in txt[start + 1:]:
        if l.startswith("//---------------------"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            seq.append(cur)
    return seq


def main():
    rep = sys.argv[1]
    so = sys.argv[2] if len(sys.argv) > 2 else None
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = dict(zip(hdr, zip(vals, units)))
    print(f"# {os.path.basename(rep)} -- kernel {d.get('Kernel Name', ('?',))[0]}")
    print("## launch metrics")
    for k in RAW_KEYS:
        if k in d:
            print(f"{k:78s} {d[k][0]:>18s} {d[k][1]}")
    src = ncu_csv(rep, "source")
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    data = src[2:]
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
    tot_s = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"\n## warp-state samples (all SASS of the kernel): {tot_s} samples, {tot_i} warp instructions executed")
    for k in STALLS:
        if k in ix:
            v = sum(int(r[ix[k]]) for r in data)
            if v:
                print(f"{k:28s} {v:8d} {100.0 * v / tot_s:5.1f} %")
    print("\n## global-memory instructions (warp-level executions / thread-level, predicated on)")
    for r in data:
        s = r[ix["Source"]].strip()
        if re.search(r"\b(LDG|STG|REDG|ATOMG|UBLKCP|LDL|STL)\b", s) and int(r[ix["Instructions Executed"]]) > 1000:
            print(f"{int(r[ix['Instructions Executed']]):>10d} {int(r[ix['Predicated-On Thread Instructions Executed']]):>11d}  samples {int(r[ix['# Samples']]):>6d}  {s[:100]}")
    print("\n## hottest SASS instructions by samples")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:20]:
        st = {k: int(r[ix[k]]) for k in STALLS if k in ix and int(r[ix[k]]) > 0}
        top = max(st, key=st.get) if st else "-"
        print(f"{int(r[ix['# Samples']]):>7d} {100.0 * int(r[ix['# Samples']]) / tot_s:5.1f} %  exec {int(r[ix['Instructions Executed']]):>9d}  {top:22s} {r[ix['Source']].strip()[:90]}")
    if so:
        seq = line_map(so)
        if seq and len(seq) == len(data):
            agg, smp = collections.Counter(), collections.Counter()
            for cur, r in zip(seq, data):
                agg[cur] += int(r[ix["Instructions Executed"]])
                smp[cur] += int(r[ix["# Samples"]])
            root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "denormalized_b200", "csrc")
            text = {}
            for f in ("dnz_kernels.cu", "dnz_device.cuh"):
                try:
                    text[f] = open(os.path.join(root, f)).read().split("\n")
                except OSError:
                    pass
            print("\n## CUDA source lines by samples (SASS attributed through -lineinfo)")
            for k, c in sorted(smp.items(), key=lambda kv: -kv[1])[:30]:
                t = text[k[0]][k[1] - 1].strip()[:100] if k and k[0] in text else ""
                print(f"{c:>7d} {100.0 * c / tot_s:5.1f} %  warp-instr {agg[k]:>10d}  {k[0] if k else '?'}:{k[1] if k else 0:<4d} {t}")
        else:
            print("\n(line attribution skipped: the .so does not match the capture)")


if __name__ == "__main__":
    main()

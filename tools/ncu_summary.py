"""Summarise an `ncu --page raw --csv` export (tools/gpu_ncu.sh) into a small per-kernel table for profiles/.

    python tools/ncu_summary.py gpurun_out/ncu_raw_<tag>.csv > profiles/<tag>_ncu_summary.csv
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor_hmma_inst_pct"),
    ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tensor_bf16_ops_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_B"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col["Kernel Name"]
    w = csv.writer(sys.stdout)
    used = [(c, n) for c, n in COLS if c in col]
    w.writerow(["id", "kernel"] + [f"{n}[{units[col[c]]}]" if units[col[c]] else n for c, n in used])
    for r in data:
        if len(r) <= name_i:
            continue
        k = r[name_i]
        k = k.replace("dcs::", "").split("(")[0][:60]
        w.writerow([r[0], k] + [r[col[c]] for c, _ in used])


if __name__ == "__main__":
    main(sys.argv[1])

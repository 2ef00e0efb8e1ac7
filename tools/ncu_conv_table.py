"""Markdown table of every conv_tc2 / wgrad_tc launch of an `ncu --set full` capture of tools/step_profile.py:

  ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|wgrad_tc_kernel" -s 60 -c 32 \
      -o gpurun_out/prof_conv_final -f python tools/step_profile.py
  ncu -i gpurun_out/prof_conv_final.ncu-rep --page raw --csv > /tmp/conv_raw.csv
  python tools/ncu_conv_table.py /tmp/conv_raw.csv > profiles/r02_ncu_conv_wgrad.md

The epilogue instantiation of a conv launch (template parameter EPI, conv_tc2.cu) names its configuration; the layer a
launch belongs to follows from it, the grid and the traffic (the order is the launch order of a multi-stream step)."""
import csv
import re
import sys

FLAGS = [(1, "mask"), (2, "accumulate"), (4, "x-phase pair"), (8, "WIDE (128 out)"), (16, "bias+ReLU"), (32, "s2d copy")]


def epi_of(name: str) -> str:
    if "wgrad" in name:
        return "wgrad_tc"
    m = re.search(r"conv_tc2_kernel<(?:\(bool\))?\d, (?:\(int\))?(-?\d+)>", name)
    if not m:
        return "conv_tc2"
    e = int(m.group(1))
    if e < 0:
        return "conv_tc2 generic epilogue"
    return "conv_tc2 [" + (", ".join(n for b, n in FLAGS if e & b) or "plain PAIR") + "]"


def main(path):
    r = list(csv.reader(open(path)))
    hdr, rows = r[0], r[2:]
    g = lambda row, k: float(row[hdr.index(k)])
    print("# ncu --set full of every conv_tc2 / wgrad_tc launch of one training step (final build of round 2)\n")
    print("`ncu --set full --clock-control none --import-source on -k regex:\"conv_tc2_kernel|wgrad_tc_kernel\" -s 60 -c 32 "
          "python tools/step_profile.py`, table by `tools/ncu_conv_table.py`")
    print("(B = 16, 304^2 internal grid; durations under ncu are serialised and cold-cache -- the shares agree with "
          "`r02_step_profile.txt` (CUDA events) and `r02_launches.csv`).\n")
    print("| # | kernel [epilogue instantiation] | grid | us | tensor pipe active % | DRAM read MB | DRAM write MB | L2 hit % | "
          "registers | dyn smem KB |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    tot = {}
    for i, row in enumerate(rows):
        name = epi_of(row[hdr.index("Kernel Name")])
        us = g(row, "gpu__time_duration.sum")
        tot[name.split(" [")[0].split(" generic")[0]] = tot.get(name.split(" [")[0].split(" generic")[0], 0.0) + us
        print(f"| {i} | {name} | {row[hdr.index('launch__grid_size')]} | {us:.1f} | "
              f"{g(row, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g(row, 'dram__bytes_read.sum'):.1f} | {g(row, 'dram__bytes_write.sum'):.1f} | "
              f"{g(row, 'lts__t_sector_hit_rate.pct'):.1f} | {row[hdr.index('launch__registers_per_thread')]} | "
              f"{g(row, 'launch__shared_mem_per_block_dynamic'):.0f} |")
    print("\nSum of the captured launches: " + ", ".join(f"{k} {v:.0f} us" for k, v in tot.items()) + ".")


if __name__ == "__main__":
    main(sys.argv[1])

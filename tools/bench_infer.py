"""configs[2]: high-res inference -- ConvNP.predict onto the 1400x1400 NZ target grid from ERA5 + stations (S3).
Thin wrapper over ``bench.py --workload infer`` (device-resident forward, predict end to end, per-kernel table):
  python tools/bench_infer.py [tasks]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    n = sys.argv[1] if len(sys.argv) > 1 else "32"
    sys.exit(subprocess.call([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "infer", "--steps", n,
                              "--warmup", "3"]))

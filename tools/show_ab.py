"""Prints one line per build from an ab_libs.py result file.  usage: python tools/show_ab.py gpurun_out/ab.json"""
import json
import sys

for name, v in json.load(open(sys.argv[1])).items():
    if name.startswith("_"):
        continue
    print(f"{name:28s} K1 {min(v['kernel_us']):7.2f} .. {max(v['kernel_us']):7.2f} us   forward {min(v['forward_us']):7.2f} us   "
          f"sha1 {v['sha1_none'][:8]} {v['sha1_batch_small'][:8]}   max|diff| vs first {v['max_abs_diff_vs_first']:.3g}")

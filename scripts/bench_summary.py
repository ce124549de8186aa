"""Prints the headline numbers of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.0f} Mpix/s  {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.0f} Mpix/s ({d['e2e']['ms_per_step']:.2f} ms)  "
      f"frac {d['roofline']['frac']:.3f} tile {d['roofline']['kernel_ms']:.2f} ms other {d['roofline']['other_kernels_ms']:.2f} ms  "
      f"launches {d['gpu_launches']}  checksum {d['canvas_checksum']}  parity {d.get('parity')}")
if d.get("cpu_baseline"):
    print("cpu_baseline", d["cpu_baseline"]["value"], "Mpix/s on", d["cpu_baseline"]["cores"], "threads,", d["cpu_baseline"].get("ms"), "ms")
for k, v in (d.get("configs") or {}).items():
    if "value_mpix_s" in v:
        e = v.get("e2e", {})
        print(f"  {k}: {v['value_mpix_s']:.0f} Mpix/s {v['value_fills_per_s'] / 1e6:.2f} Mfills/s {v['ms_per_step']:.2f} ms/step | e2e {e.get('mpix_s', 0):.0f} Mpix/s "
              f"{e.get('fills_per_s', 0) / 1e6:.2f} Mfills/s | parity {v.get('parity') and (v['parity'].get('pixels_differing'), v['parity'].get('max_channel_diff'))} "
              f"| cpu {v.get('reference_cpu') and round(v['reference_cpu']['ms_per_step'], 1)} ms")
    else:
        print(f"  {k}: e2e {v['e2e']} resident {v['resident']} parity {v['parity']}")
for k, v in (d.get("roofline_full_canvas") or {}).items():
    print(f"  full {k}: {v['achieved']:.0f} GB/s frac {v['frac']:.3f} ({v['kernel_ms']:.3f} ms) {v['kernel']}")
if d.get("band_sharded"):
    print("  band:", d["band_sharded"])
print("clocks", d.get("clocks"))

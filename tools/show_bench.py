"""Print the key numbers of bench JSON lines (diagnostic)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    ns = 1e9 / d["value"] * d["n_gpus"]
    sh = d["roofline"]["stage_time_share_warmup"]
    print(f"{f}: {d['value'] / 1e6:.3f} M/s  e2e {d['e2e']['value'] / 1e6:.3f}  {ns:.1f} ns/patch  dom {d['roofline']['kernel']} "
          f"frac {d['roofline']['frac']:.3f} whole {d['roofline']['whole_path']['frac']:.3f}")
    print("   per-stage ns/patch:", {k: round(v * ns, 1) for k, v in sh.items()}, d["clocks"])

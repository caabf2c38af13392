"""GPU check of the warpgroup-per-patch NAS tail kernel (csrc/nas_tail.cuh): parity against the CPU oracle and throughput
at batch 65 536 (BASELINE config 5) for several launch plans.

    python tools/nas_tail_check.py [--fast] [--timing-only]
"""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from tools import nas_resident_check as rc  # noqa: E402

rc.ENV_KEYS = rc.ENV_KEYS + ("HN_NAS_TAIL", "HN_NAS_TAIL_CUT", "HN_NAS_TAIL_WG", "HN_NAS_FRONT_DW", "HN_NAS_TAIL_MINOPS", "HN_NAS_FOLD")


def main():
    fast = "--fast" in sys.argv
    ok = True
    plans = [{}, {"HN_NAS_FOLD": "0"}, {"HN_NAS_TAIL_CUT": "0"}, {"HN_NAS_TAIL_WG": "2"}]
    if "--timing-only" not in sys.argv:
        for arch in ("wang2", "wang3", "wang4", "mixed_se"):
            for env in (plans[:2] if fast else plans):
                try:
                    r = rc.parity(arch, env)
                except Exception as exc:  # keep going: one broken plan must not hide the others
                    r = {"arch": arch, "env": env, "error": repr(exc), "ok": False}
                ok = ok and r["ok"]
                print("PARITY", json.dumps(r), flush=True)
    tplans = [{}, {"HN_NAS_TAIL_CUT": "0"}, {"HN_NAS_FOLD": "0"}, {"HN_NAS_TAIL": "0"}] if fast else \
        [{}, {"HN_NAS_TAIL_CUT": "0"}, {"HN_NAS_TAIL_WG": "3"}, {"HN_NAS_TAIL_WG": "2"}, {"HN_NAS_TAIL": "0"}]
    for arch in ("wang2", "wang3", "wang4"):
        for env in tplans:
            try:
                print("TIMING", json.dumps(rc.timing(arch, env)), flush=True)
            except Exception as exc:
                print("TIMING", json.dumps({"arch": arch, "env": env, "error": repr(exc)}), flush=True)
    print("ALL_OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

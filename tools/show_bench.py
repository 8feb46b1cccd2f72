"""Condense bench.py JSON lines (stdin or files) into a few readable lines."""
import json
import sys


def show(d):
    if "modes" not in d:
        print(json.dumps(d)[:600])
        return
    print("N=%d value=%.1f ms=%.4f mode=%s vs_baseline=%s e2e=%s clocks=%s" % (
        d["n_gpus"], d["value"], d["ms_per_step"], d["config"]["mode"], d["vs_baseline"],
        d.get("e2e", {}).get("value"), (d.get("clocks") or {}).get("sm_mhz")))
    for k, m in d["modes"].items():
        print("   %-5s %.4f [%.4f..%.4f] nvls=%s hbm=%s nvl=%s" % (k, m["ms_per_step"], m["ms_min"], m["ms_max"],
              m.get("nvls_multicast"), m.get("hbm_frac_of_measured"), m.get("nvlink_egress_frac_of_770")))
        print("         serial:", m.get("phases_serialised_ms"))
        if m.get("phases_overlapped_ms"):
            print("         overlap:", m.get("phases_overlapped_ms"))
    for k, b in d.get("baselines", {}).items():
        print("   base %-15s %.4f ms  -> x%.2f" % (k, b["ms_per_step"], b["ms_per_step"] / d["ms_per_step"]))


def main():
    srcs = [open(a) for a in sys.argv[1:]] or [sys.stdin]
    for f in srcs:
        for ln in f:
            ln = ln.rstrip()
            if ln.startswith("{"):
                try:
                    show(json.loads(ln))
                    continue
                except Exception:  # noqa: BLE001
                    pass
            print(ln[:240])


if __name__ == "__main__":
    main()

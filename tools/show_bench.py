"""One line per bench JSON: kernel-only value, end-to-end value / step time, link bytes, fetch-list figures.   python tools/show_bench.py FILE..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print("%-40s value %.3e (%.3f ms)  e2e %.3e (%.2f ms)  h2d %.0f MB copied %.0f MB rx %s GB/s  pairs %s gather %s ms x%s  frac %.3f" % (
            f.split("/")[-1], d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step", 0), e.get("h2d_bytes_per_step", 0) / 1e6,
            e.get("h2d_copied_bytes_per_step_rank0", 0) / 1e6, e.get("pcie_rx_gbs_rank0") and round(e["pcie_rx_gbs_rank0"], 1),
            e.get("fetch_pairs_per_step_rank0"), e.get("host_gather_ms_per_step_rank0"), e.get("host_gather_threads"), d["roofline"]["frac"]))
    except Exception as ex:  # an empty or broken line: show the tail of the matching .err
        print(f, "unreadable:", ex)
        try:
            print(open(f.replace(".json", ".err")).read()[-1500:])
        except OSError:
            pass

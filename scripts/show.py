"""prints one summary line per bench JSON in gpurun_out/: python scripts/show.py r02d_def r02d_s2 ..."""
import json
import sys

for name in sys.argv[1:]:
    path = "gpurun_out/%s.json" % name
    try:
        d = json.load(open(path))
        ch = d.get("checks") or {}
        e2e = d.get("e2e") or {}
        lk = d.get("lookup_pass") or {}
        print(name, "ms/step %.1f" % d["ms_per_step"], "G/s %.2f" % (d["value"] / 1e9),
              {k: round(v, 1) for k, v in d["profile_ms_per_step"].items()},
              "direct", d["stats"]["direct_inserts"], "checks", ch.get("sum_counts_equals_counted_instances"), ch.get("histogram_entries_equal_distinct"),
              "parity", ch.get("oracle_parity"), "e2e %.2f" % (e2e.get("value", 0) / 1e9), "lookup %.1f" % (lk.get("value", 0) / 1e9),
              "whole %.3f" % d["roofline"]["whole_pass"]["frac"])
    except Exception as e:
        try:
            err = open("gpurun_out/%s.err" % name).read()[-1500:]
        except Exception:
            err = ""
        print(name, "ERR", repr(e), err)
